"""cProfile of an S2 chain (GPU box).  usage: python tools/prof_chain.py [gibbs]"""
import cProfile
import io
import pstats
import sys
import time

sys.path.insert(0, '.')
from mitre_b200 import chains  # noqa: E402

gibbs = len(sys.argv) > 1 and sys.argv[1] == "gibbs"
data, model = chains.build_model("S2", seed=0)
r0 = chains.initial_rule_list(model)
chains.run_chain(model, 1, 5, r0, device_gibbs=gibbs)
t0 = time.perf_counter()
chains.run_chain(model, 3, 100, r0, device_gibbs=gibbs)
print('samples/s', 100 / (time.perf_counter() - t0))
pr = cProfile.Profile()
pr.enable()
chains.run_chain(model, 2, 100, r0, device_gibbs=gibbs)
pr.disable()
for key in ('tottime', 'cumtime'):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(28)
    print(s.getvalue()[:6500])

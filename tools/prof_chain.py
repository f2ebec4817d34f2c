import cProfile, pstats, sys, io, time
sys.path.insert(0, '/root/repo')
from mitre_b200 import chains
data, model = chains.build_model("S2", seed=0)
r0 = chains.initial_rule_list(model)
chains.run_chain(model, 1, 5, r0)
t0 = time.perf_counter(); chains.run_chain(model, 3, 60, r0); print('samples/s', 60 / (time.perf_counter() - t0))
pr = cProfile.Profile(); pr.enable()
chains.run_chain(model, 2, 60, r0)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(22); print(s.getvalue()[:5000])

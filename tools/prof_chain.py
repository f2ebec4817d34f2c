import cProfile, pstats, sys, io
sys.path.insert(0, '/root/repo')
from mitre_b200 import chains
data, model = chains.build_model("S2", seed=0)
r0 = chains.initial_rule_list(model)
chains.run_chain(model, 1, 3, r0)
pr = cProfile.Profile(); pr.enable()
chains.run_chain(model, 2, 30, r0)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:6000])

"""Independent chains (the reference's `[replicates]` mode, main.py:1913-1986) sharded over GPUs.

Chains are pure replicas: every rank builds the model once on its GPU (detector evaluation +
lexsort on the device), runs its share of the chains one after another with seeds base_seed + i
(main.py:1930-1938) and no collective while sampling; the per-chain point summaries are gathered
at the end.  One process per GPU:

    python -m torch.distributed.run --nproc-per-node 8 -m mitre_b200.chains --chains 64 --steps 300
    python -m torch.distributed.run --nproc-per-node 32 -m mitre_b200.chains --chains 64 --steps 300   # 4 processes per GPU
    python -m mitre_b200.chains --chains 2 --steps 50          # single GPU
"""
import argparse
import json
import os
import time

import numpy as np


def build_model(shape="S2", seed=0, variables=None, **model_kwargs):
    """Synthetic series of a named shape -> (Dataset, LogisticRuleModel) on the current GPU."""
    from . import logit_rules as LR, rules as R, synthetic
    d = synthetic.make_shape(shape, seed=seed, V=variables)
    V = len(d["variable_weights"])
    data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                     d["experiment_start"], d["experiment_end"])
    model = LR.LogisticRuleModel(data, d["tmin"], N_intervals=d["N_intervals"], tmax=d["tmax"],
                                 **model_kwargs)
    return data, model


def initial_rule_list(model):
    """main.py:952-955: the single primitive with the smallest Hamming distance to the outcomes."""
    from . import rules as R
    truth = np.asarray(model.rule_population.flat_truth, dtype=bool)
    dist = np.mean(truth != np.asarray(model.data.y, dtype=bool)[None, :], axis=1)
    return R.RuleList([[model.rule_population.flat_rules[int(np.argmin(dist))]]])


def run_chain(model, seed, steps, r0=None, device_draw=True, device_gibbs=False):
    """One chain: numpy's global RNG and the Polya-Gamma sampler are seeded with `seed`."""
    from . import logit_rules as LR, polyagamma
    np.random.seed(seed)
    LR.ppg = polyagamma.PyPolyaGamma(seed)
    sampler = LR.LogisticRuleSampler(model, r0 if r0 is not None else initial_rule_list(model),
                                     device_draw=device_draw, device_gibbs=device_gibbs)
    sampler.sample(steps)
    return sampler


def scalar_traces(sampler):
    """float64[steps, 6] per-step scalars of a chain: rules, detectors, log-likelihood, log-prior, phylogeny
    mean, window concentration -- the kind of columns the reference writes to `_mcmc_traces.csv`
    (posterior.py:81-137) and feeds to its R-hat tool."""
    rows = []
    for st, lik, pri in zip(sampler.states, sampler.likelihoods, sampler.priors):
        rows.append([len(st.rl), sum(len(r) for r in st.rl.rules), lik, pri, st.phylogeny_mean,
                     st.window_concentration])
    return np.asarray(rows, dtype=np.float64)


def split_r_hat(traces):
    """Gelman's potential scale reduction over independent runs, each split in halves, as the reference's
    `mitre_mcmc_diagnostics` computes it (mcmc_diagnostics.py:39-84: first half without its last sample,
    second half from the midpoint; within-run variances over n - 1, n = the shortest half minus one as in
    its `compute_R_hat(runs2, n - 1)` call).  traces: list of float[steps, k] -> float64[k]."""
    halves = []
    for t in traces:
        t = np.asarray(t, dtype=np.float64)
        mid = t.shape[0] // 2
        halves += [t[:mid - 1], t[mid:]]
    n = min(h.shape[0] for h in halves) - 1
    x = np.stack([h[:n] for h in halves])                    # [m, n, k]
    m = x.shape[0]
    means = x.mean(axis=1)
    within = ((x - means[:, None, :]) ** 2).sum(axis=1) / (n - 1)
    B = n * ((means - means.mean(axis=0)) ** 2).sum(axis=0) / (m - 1)
    W = within.mean(axis=0)
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.sqrt((W * (n - 1) / n + B / n) / W)


def summarize(sampler, burnin_fraction=0.05):
    from .posterior import PosteriorSummary
    s = PosteriorSummary(sampler, burnin_fraction=burnin_fraction)
    s.point_summarize()
    return [[list(map(str, p.as_tuple())) for p in rule] for rule in s.point.rl.rules]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=2)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--shape", default="S2")
    ap.add_argument("--variables", type=int, default=None)
    ap.add_argument("--base-seed", type=int, default=0)
    ap.add_argument("--host-gibbs", action="store_true",
                    help="omega/beta updates on the host (reference RNG stream) instead of the device kernel")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from . import parallel as PL
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_dev = torch.cuda.device_count()
    torch.cuda.set_device(local_rank % n_dev)
    if world > 1:
        # chains are replicas: the only exchange is the final gather of summaries.  With more processes
        # than GPUs (a step is host-bound at S2 sizes, so several processes per GPU raise the aggregate
        # rate) NCCL refuses two ranks on one device: use gloo for that gather.
        if world > n_dev:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    data, model = build_model(args.shape, seed=args.base_seed, variables=args.variables)
    r0 = initial_rule_list(model)
    mine = PL.chains_for_rank(args.chains, world, rank, args.base_seed)
    t0 = time.perf_counter()
    results = []
    for i, seed in mine:
        sampler = run_chain(model, seed, args.steps, r0, device_gibbs=not args.host_gibbs)
        results.append(dict(chain=i, seed=seed, point=summarize(sampler),
                            final_length=len(sampler.current_state.rl)))
    dt = time.perf_counter() - t0
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, (dt, results))
    else:
        gathered = [(dt, results)]
    if rank == 0:
        wall = max(g[0] for g in gathered)
        chains = sorted(sum([g[1] for g in gathered], []), key=lambda r: r["chain"])
        print(json.dumps(dict(n_gpus=min(world, n_dev), processes=world, chains=args.chains, steps=args.steps,
                              candidates=len(model.rule_population),
                              samples_per_sec=args.chains * args.steps / wall, wall_s=wall,
                              points=[c["point"] for c in chains])))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

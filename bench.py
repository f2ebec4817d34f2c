#!/usr/bin/env python
"""bench.py -- candidate rule-set likelihood evaluations per second (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

Workload (config.workload = "S5-primitive-sweep", BASELINE.json configs[4]): ONE problem of N = 35
subjects, V = 10,000 aggregated variables, n_intervals = 24 with unrestricted window length, mean +
slope detectors, full threshold grid -> 2.9e8 candidate primitives (2.35 GB packed, plus 2.35 GB of
log-prior and 2.35 GB of log-likelihoods per sweep: far larger than the 126 MB L2, so no flush is needed
between steps).  The table is built on the GPU by this repo's detector kernels from synthetic series of
that shape.

A step = one detector-update sweep of the sampler over ALL candidates (update_primitives,
logit_rules.py:794-881: likelihoods_of_all_options + primitive prior + proposal draw): base rule list of 3
rules + constant (p = 4), base system + subset-sum tables, log-likelihood of every candidate, (max,
sum-exp) of likelihood + log-prior, proposal draw -- the drawn index and the log-sum-exp are the step's
result.

  value     evals/s of that step with its inputs resident in HBM, device-timed (CUDA events)
  e2e       the same step through the public call (mitre_b200.parallel.ShardedSweep.step): sweep inputs
            from pinned host memory every step, drawn index + log-sum-exp + status back to pinned host
            memory; e2e_full_vector (N = 1) is the reference-signature call that returns all P
            log-likelihoods to the host (LogisticRuleModel.likelihoods_of_all_options, PCIe-bound)
  roofline  the step's candidate kernel (score_w1d_kernel, stats variant): algorithmic bytes 24
            B/candidate (8 B packed row + 8 B log-prior in, 8 B log-likelihood out) / its launch time;
            peak = MEASURED_PEAKS.json hbm_gbs
  N > 1     STRONG scaling: the same 10,000-variable problem, built by all ranks together
            (ShardedPopulation: detectors sharded by variable), its lexsorted candidates cut into N
            contiguous slices (the reference's own _likelihood_worker precedent, logit_rules.py:21-60).
            Every rank scores its slice; the 16-byte (max, sum-exp) pairs and the drawn index travel
            INSIDE the timed step (--exchange p2p: stores into peer memory over NVLink from the draw
            kernel; --exchange nccl: all-gather + all-reduce).  parity_check: every rank also builds
            the whole problem alone and verifies rows, log-likelihoods, log-sum-exp and drawn indices.
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "candidate_likelihood_evals_per_sec"
UNIT = "evals/s"
SIGMA2 = 100.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variables", type=int, default=10000, help="variables of the problem (S5: 10000)")
    ap.add_argument("--rules", type=int, default=3, help="rules in the base rule list (p = rules+1)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: how the (max, sum-exp) pairs and the drawn index travel")
    ap.add_argument("--check", default="auto", choices=["auto", "on", "off"],
                    help="sharded == single-GPU parity check on every rank (auto: on for N > 1)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# clocks sampling (pynvml) during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons (NVML) sampled WHILE the timed region runs on the GPU: the host first
    enqueues the whole region (asynchronous launches), then calls poll(done) which samples until done()
    says the end event has completed.  Sampling from a second thread during the launch loop instead was
    measured to stretch an 8-GPU, 7 ms region by a third (NVML calls of 8 processes contend with the
    launch path)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.sample(keep=False)        # first call pays the lazy initialisation
        except Exception:
            self.nv = None

    def sample(self, keep=True):
        if self.nv is None:
            return
        try:
            mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
            try:
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            if keep:
                self.samples.append(mhz)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
        except Exception:
            pass

    def poll(self, done, interval=0.002, limit=10000):
        """Sample until done() is true (at least once)."""
        for _ in range(limit):
            self.sample()
            if done():
                break
            time.sleep(interval)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
# sweep state shared by both arms
# ---------------------------------------------------------------------------------------------
def sweep_state(N, n_rules, row_sampler, seed):
    """omega ~ Gamma(2, 1/8) + 0.05 (PG(1, psi)-like), alternating outcomes, a base design matrix of
    `n_rules` rule columns taken from the candidate table + the constant column
    (rules.py:1173-1175) and the mask of one other primitive in the updated rule."""
    rng = np.random.RandomState(seed)
    omega = rng.gamma(2.0, 1.0 / 8.0, size=N) + 0.05
    y = (np.arange(N) % 2 == 0)
    cols = [row_sampler(rng).astype(np.float64) for _ in range(n_rules)]
    X = np.stack(cols + [np.ones(N)], axis=1)
    mask = row_sampler(rng)
    return dict(X=X, omega=omega, kappa=y - 0.5, y=y, mask=mask)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def fp64_roofline(D, P, kernel_ms):
    """Scoring is fp64-pipe work (north_star): algorithmic fp64 operations per candidate of the
    table kernel at p = 4 -- 12 adds (two 6-vector table sums) + 28 (|w|^2, Schur complement, table
    log, table reciprocal, assembly) = 40 -- over the measured DFMA rate of this device."""
    ops = 40.0
    peak = D.probe_fp64_tflops()
    achieved = ops * P / (kernel_ms * 1e-3) / 1e12
    return {"ops_per_candidate": ops, "achieved_tops": achieved, "peak_dfma_tflops": peak,
            "peak_instr_tops": peak / 2.0, "frac_of_instr_peak": achieved / (peak / 2.0)}


def ncu_traffic_per_launch(P, which):
    """dram bytes per launch of the step's candidate kernel: bytes per candidate from the committed ncu
    --set full capture of the same kernel at the full S5 size (profiles/score_w1d_traffic.json), times the
    candidates of this launch.  Not measured in this run (bench.py never runs under a profiler)."""
    path = os.path.join(ROOT, "profiles", "score_w1d_traffic.json")
    try:
        rec = json.load(open(path))
        return float(rec[which]["dram_bytes_per_candidate"]) * P
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU path on a bounded sample of the same workload
# ---------------------------------------------------------------------------------------------
def numpy_sample_rows(V_s, seed):
    """Lexsorted bool[P_s, 35] candidate rows of the S5 shape for V_s variables, generated on the
    CPU with numpy only (closed-form slopes; this is workload generation, not the timed path)."""
    from mitre_b200 import synthetic
    from oracle import detectors_oracle as DO
    d = synthetic.make_shape("S5", seed=seed, V=V_s)
    windows, ok_slope = DO.enumerate_windows(d["T"], d["experiment_start"], d["experiment_end"],
                                             d["N_intervals"], d["tmin"], d["tmax"])
    N = len(d["X"])
    rows = []
    for (t0, t1), ok in zip(windows, ok_slope):
        mid = 0.5 * (t0 + t1)
        means = np.empty((V_s, N))
        slopes = np.empty((V_s, N))
        for s, (x, t) in enumerate(zip(d["X"], d["T"])):
            m = (t >= t0) & (t <= t1)
            xs, ts = x[:, m], t[m] - mid
            means[:, s] = xs.mean(axis=1)
            if ok:
                dt = ts - ts.mean()
                slopes[:, s] = (xs - xs.mean(axis=1, keepdims=True)).dot(dt) / dt.dot(dt)
        for vals in ([slopes, means] if ok else [means]):
            sv = np.sort(vals, axis=1)
            thr = 0.5 * (sv[:, :-1] + sv[:, 1:])                       # [V_s, N-1]
            above = vals[:, None, :] > thr[:, :, None]                   # [V_s, N-1, N]
            keep = np.ones(thr.shape, dtype=bool)
            keep[:, 1:] = thr[:, 1:] != thr[:, :-1]
            above = above[keep]
            rows.append(above)
            rows.append(~above)
    truth = np.concatenate(rows, axis=0)
    return truth[np.lexsort(np.rot90(truth))]


_REF = {}


def _ref_worker(args):
    lo, hi = args
    ref = _REF["ref"]
    st = _REF["state"]
    opts = (_REF["truth"][lo:hi] * st["mask"]).astype(np.float64)       # logit_rules.py:214-217
    z = st["kappa"] / st["omega"]                                        # logit_rules.py:222-223
    out = ref.likelihoods_of_all_options(st["X"], st["omega"], z, opts, np.float64(SIGMA2))
    return float(out[0])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    from oracle import build_ref
    import oracle
    kind = "reference"
    try:
        ref = build_ref.load()
    except Exception:
        ref = oracle          # the C port has the same call signature
        kind = "port"
    cores = max(1, min(os.cpu_count() or 1, 64))
    V_s = max(8, min(50, 6 * cores))
    truth = numpy_sample_rows(V_s, args.seed)
    P_s, N = truth.shape
    state = sweep_state(N, args.rules, lambda rng: truth[rng.randint(P_s)], args.seed + 1)
    _REF.update(ref=ref, truth=truth, state=state)
    bounds = np.linspace(0, P_s, cores + 1).astype(int)
    slices = [(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(slices)) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_ref_worker, slices)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_ref_worker, slices)
        dt = time.perf_counter() - t0
    value = P_s * args.steps / dt
    # single-core figure for context (the reference itself is single-threaded: main.py:1553)
    t0 = time.perf_counter()
    _ref_worker((0, min(P_s, 400000)))
    one_core = min(P_s, 400000) / (time.perf_counter() - t0)
    sample = ("%d lexsorted candidates of the S5 shape (%d variables), bool->float64 marshal "
              "(logit_rules.py:214-217) + efficient_likelihoods kernel per step, candidate range split "
              "over %d forked processes like _likelihood_worker (logit_rules.py:21-60)"
              % (P_s, V_s, len(slices)))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "S5-primitive-sweep", "subjects": N, "variables_sampled": V_s,
                   "candidates_per_step": int(P_s), "base_columns_p": args.rules + 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": len(slices), "kind": kind,
                         "sample": sample, "one_core_value": one_core},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def cuda_time_ms(torch, fn, iters):
    start = torch.cuda.Event(enable_timing=True)
    end = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for _ in range(iters):
        fn()
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) / iters


def build_population(args, rank, torch):
    """Synthetic S5-shaped series -> device population (this repo's detector kernels), with the
    phases timed by CUDA events for the detector-evaluation roofline."""
    from mitre_b200 import rules as R, synthetic
    d = synthetic.make_shape("S5", seed=args.seed + rank, V=args.variables)
    V = args.variables
    data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                     d["experiment_start"], d["experiment_end"])
    data.device_series()
    torch.cuda.synchronize()
    marks = {}

    def timed(name, fn):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        marks[name] = marks.get(name, 0.0) + s.elapsed_time(e)

    pop = R.DiscreteRulePopulation.__new__(R.DiscreteRulePopulation)
    pop.model, pop.data, pop.tmin, pop.tmax = None, data, d["tmin"], np.inf if d["tmax"] is None else d["tmax"]
    pop.N_intervals, pop.max_thresholds, pop.applied_filters = d["N_intervals"], np.inf, []
    # mine_rules, phase by phase
    orig_calc, orig_upd, orig_sort = pop.calculate_values, pop.update_base_rules, pop.sort_base_rules
    pop.calculate_values = lambda: timed("values", orig_calc)
    pop.update_base_rules = lambda: timed("thresholds_rows", orig_upd)
    pop.sort_base_rules = lambda: timed("lexsort", orig_sort)
    t0 = time.perf_counter()
    pop.mine_rules()
    torch.cuda.synchronize()
    marks["wall_s"] = time.perf_counter() - t0
    marks["slope_flagged"] = pop.slope_groups_flagged
    del pop.calculate_values, pop.update_base_rules, pop.sort_base_rules
    # kernel-only time of detector evaluation: re-launch the fused kernel on the prepared inputs
    from mitre_b200 import device as D
    if pop._fused_args is not None:
        if pop._walk_args is not None:
            run, name = (lambda: D.detectors_walk(*pop._walk_args)), "values_walk_kernel + sort_emit_staged_kernel"
        else:
            run, name = (lambda: D.detectors_fused(*pop._fused_args)), "values_resident_kernel + sort_emit_staged_kernel"
        run()
        marks["detector_kernel_ms"] = min(cuda_time_ms(torch, run, 3) for _ in range(3))
        marks["detector_kernel"] = name
        torch.cuda.empty_cache()
        marks["detector_r1_kernels_ms"] = min(cuda_time_ms(torch, lambda: D.detectors_fused(*pop._fused_args), 3)
                                              for _ in range(2))
    torch.cuda.empty_cache()
    return data, pop, marks


def detector_bytes(pop, data):
    """B_det = 8 V S + 8 S + sum_g [8 N + 8 K_g + 2 K_g 8 W64]."""
    N = data.n_subjects
    S = data.device_series()["S"]
    W64 = (N + 63) // 64
    K = pop._K_host.astype(np.int64)
    return 8 * data.n_variables * S + 8 * S + int(np.sum(8 * N + 8 * K + 2 * K * 8 * W64))


def build_sharded(args, torch):
    """ONE S5 problem (args.variables variables in total) built by all ranks together: detectors by
    variable, one exchange for the global order, every rank keeps its contiguous slice of the sorted
    candidate table (mitre_b200.parallel.ShardedPopulation)."""
    from mitre_b200 import parallel as PL, rules as R, synthetic
    d = synthetic.make_shape("S5", seed=args.seed, V=args.variables)
    data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(args.variables)], d["variable_weights"],
                     d["experiment_start"], d["experiment_end"])
    t0 = time.perf_counter()
    spop = PL.ShardedPopulation(data, d["tmin"], d["N_intervals"], d["tmax"])
    torch.cuda.synchronize()
    return data, spop, time.perf_counter() - t0


def global_prior(P, seed, dev, torch, lo, hi):
    """A log-prior over all P candidates, a pure function of the global index: every rank generates the
    same vector and keeps its slice."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed + 7)
    full = torch.randn(P, dtype=torch.float64, device=dev, generator=g) - math.log(max(P, 1))
    out = full[lo:hi].clone()
    del full
    return out


def parity_check(args, torch, D, PL, sw, state, P, N, us):
    """Sharded == single table (run on EVERY rank, N > 1): this rank builds the whole problem by itself,
    scores it in one sweep and checks (a) its slice of the sharded table is that slice of the whole table,
    bit for bit, (b) so are its log-likelihoods, (c) the index every sharded step drew is the index the
    single-table draw gives for the same u, (d) the exchanged log-sum-exp is the single-table one."""
    # the step's output buffer was last written by the kernel-only timing loops (other kernel variants,
    # whose sums round differently in the last bit): run the sharded step itself once more, on every rank
    sw.stage(state["X"], state["omega"], state["kappa"], state["mask"], 0.5)
    sw.launch(args.rules + 1, SIGMA2)
    sw.result()
    data, pop, _ = build_population(args, 0, torch)
    if len(pop) != P:
        return "candidate count differs: %d vs %d" % (len(pop), P)
    a, b = sw.start, sw.start + sw.n
    if not torch.equal(pop.packed_truth[a:b], sw.rows):
        return "slice rows differ from the single-GPU table"
    dev = D.dev()
    prior = global_prior(P, args.seed, dev, torch, 0, P)
    buffers, draw = D.ScoreBuffers(P, N, args.rules + 1), D.DrawBuffers(P)
    dmask = torch.as_tensor(D.pack_mask(state["mask"]), device=dev)
    out, status, st = D.score_all_stats(pop.packed_truth, N, state["X"], state["omega"], state["kappa"], SIGMA2,
                                        mask=dmask, buffers=buffers, log_prior=prior, draw_buffers=draw)
    D.raise_on_status(status)
    if not torch.equal(out[a:b], sw.buffers.out[:sw.n]):
        return "slice log-likelihoods differ from the single-GPU sweep"
    lse = float(st[0] + torch.log(st[1]))
    for u, (idx, log_total) in us:
        want = int(D.draw_index(out, prior, u, stabilise=True, buffers=draw, reuse_tiles=True).item())
        if idx != want:
            return "draw differs at u=%r: sharded %d, single %d" % (u, idx, want)
        if abs(log_total - lse) > 1e-10 * abs(lse):
            return "log-sum-exp differs: %r vs %r" % (log_total, lse)
    return "ok"


def run_b200(args):
    import torch
    from mitre_b200 import device as D, parallel as PL
    from mitre_b200._lib import lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    p = args.rules + 1

    # ---- ONE problem; at N > 1 its candidates are sharded over the ranks (strong scaling) ----
    pop, build_marks, shard_build_s = None, None, None
    if world == 1:
        data, pop, build_marks = build_population(args, 0, torch)
        N, P = data.n_subjects, len(pop)
        local_rows, start = pop.packed_truth, 0
        shard_bounds_all = [(0, P)]
        n_windows = len(pop.windows)
    else:
        data, spop, shard_build_s = build_sharded(args, torch)
        N, P = data.n_subjects, spop.P
        local_rows, start = spop.local_rows, spop.start
        shard_bounds_all = spop.bounds
        n_windows = len(spop.windows)
        spop.local_population = None
        torch.cuda.empty_cache()
    n_local = int(local_rows.shape[0])
    W64 = D.words_for(N)
    prior = global_prior(P, args.seed, dev, torch, start, start + n_local)

    # identical sweep inputs on every rank (the broadcast of the base system, KBs)
    def row_sampler(rng):
        k = int(rng.randint(n_local))
        return np.array([(int(local_rows[k, 0]) >> (63 - i)) & 1 for i in range(N)], dtype=bool)
    state = sweep_state(N, args.rules, row_sampler, args.seed + 1) if rank == 0 else None
    if world > 1:
        box = [state]
        dist.broadcast_object_list(box, src=0)
        state = box[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sw = PL.ShardedSweep(local_rows, prior, P, N, start, exchange=args.exchange, max_p=p)
    rng_u = np.random.RandomState(args.seed + 11)      # the same uniforms on every rank

    def stage():
        return sw.stage(state["X"], state["omega"], state["kappa"], state["mask"], float(rng_u.rand()))

    # ---- value: the step with its inputs resident in HBM, device-timed ----
    # step = prepare (base system + tables) -> score the slice + (max, sum-exp) + tile pairs -> chunk sums ->
    #        exchange of the pairs, owner's draw, index back (one kernel over NVLink peer memory)
    stage()
    sw.launch(p, SIGMA2, copies=True)       # inputs into HBM once
    sw.result()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        sw.launch(p, SIGMA2, copies=False)
    barrier()
    clocks = ClockSampler(local_rank)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        sw.launch(p, SIGMA2, copies=False)
    ev1.record()
    clocks.poll(ev1.query)                 # clocks / throttle reasons while the region executes
    barrier()
    ms_step = ev0.elapsed_time(ev1) / args.steps
    if int(sw.buffers.status.item()):
        raise SystemExit("status flags %d after the timed region" % int(sw.buffers.status.item()))

    # ---- e2e: the same step through ShardedSweep.step -- host inputs in (pinned), index + log-sum-exp back ----
    drawn = []
    for _ in range(2):
        stage(); sw.launch(p, SIGMA2); sw.result()
    barrier()
    e2e_steps = max(3, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        stage()
        u = float(sw.h_small[-1])
        sw.launch(p, SIGMA2)
        drawn.append((u, sw.result()))
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps

    # ---- the other exchange mode north_star names: all-gather of the log-weights themselves ----
    dX, dom, dka, dmask = sw._views(p)
    scorer_out = sw.buffers.out[:n_local]

    def gather_step():
        D.score_all(local_rows, N, dX, dom, dka, SIGMA2, mask=dmask, buffers=sw.buffers, out=scorer_out)
        return PL.allgather_loglik(scorer_out, P, bounds=shard_bounds_all) if world > 1 else scorer_out
    gather_step()
    barrier()
    gather_ms = cuda_time_ms(torch, gather_step, max(3, min(args.steps, 10)))
    barrier()

    # ---- kernel-only times on this rank's slice (CUDA events), for the roofline ----
    tiny = local_rows[:1024]
    def stats_only():
        D.score_all_stats(local_rows, N, dX, dom, dka, SIGMA2, mask=dmask, buffers=sw.buffers, out=scorer_out,
                          log_prior=prior, draw_buffers=sw.draw)
    def plain_only():
        D.score_all(local_rows, N, dX, dom, dka, SIGMA2, mask=dmask, buffers=sw.buffers, out=scorer_out)
    def prepare_only():
        lib().mitre_set_large_table_min_rows(0)
        D.score_all(tiny, N, dX, dom, dka, SIGMA2, mask=dmask, buffers=sw.buffers, out=scorer_out[:1024])
        lib().mitre_set_large_table_min_rows(4 << 20)
    reps = max(args.steps, 5)
    stats_ms = cuda_time_ms(torch, stats_only, reps)
    plain_ms = cuda_time_ms(torch, plain_only, reps)
    prepare_ms = cuda_time_ms(torch, prepare_only, 50)
    lib().mitre_set_dedupe(0)
    plain_nodedupe_ms = cuda_time_ms(torch, plain_only, reps)
    lib().mitre_set_dedupe(1)
    dup_frac = float((local_rows[1:, 0] == local_rows[:-1, 0]).double().mean()) if n_local > 1 else 0.0
    barrier()

    # ---- parity of the sharded path against the single-GPU path, on every rank ----
    check = "skipped"
    if args.check == "on" or (args.check == "auto" and world > 1):
        check = parity_check(args, torch, D, PL, sw, state, P, N, drawn)
    # max over ranks / agreement
    t = torch.tensor([ms_step, e2e_ms, gather_ms, 0.0 if check in ("ok", "skipped") else 1.0], dtype=torch.float64,
                     device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        idx_t = torch.tensor([d[1][0] for d in drawn], dtype=torch.int64, device=dev)
        lo_t, hi_t = idx_t.clone(), idx_t.clone()
        dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
        if not torch.equal(lo_t, hi_t):
            check = "ranks disagree on the drawn index"
            t[3] = 1.0
    ms_step, e2e_ms, gather_ms = float(t[0]), float(t[1]), float(t[2])
    per_rank = torch.tensor([stats_ms, float(n_local), dup_frac], dtype=torch.float64, device=dev)
    if world > 1:
        all_ranks = torch.empty(3 * world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(all_ranks, per_rank)
        per_rank = all_ranks
    per_rank = per_rank.cpu().numpy().reshape(-1, 3)
    if float(t[3]) != 0.0 and check in ("ok", "skipped"):
        check = "failed on another rank"

    if rank == 0:
        peak, peak_src = peaks()
        # the candidate kernel's launch time INCLUDING its prepare launch (base system + tables, ~10 us): the
        # two are timed together because a back-to-back loop of the small launch alone is host-bound
        kern_stats_ms = stats_ms
        kern_plain_ms = plain_ms
        achieved = 24.0 * n_local / (kern_stats_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": P / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "S5-primitive-sweep", "subjects": N, "variables": args.variables,
                       "windows": n_windows, "candidates": int(P), "candidates_per_gpu": n_local,
                       "base_columns_p": p,
                       "step": "one detector-update sweep of ONE problem: base system + tables, log-likelihood of "
                               "every candidate of the rank's slice, (max, sum-exp) of likelihood + log-prior, "
                               "exchange, proposal draw, index on every rank",
                       "exchange": ("p2p: each rank stores its 16-byte (max, sum-exp) pair into every peer's memory and "
                                    "the owner of the marker stores the drawn index back, inside one kernel over "
                                    "NVLink (CUDA IPC peer memory, no NCCL in the step)" if args.exchange == "p2p" else
                                    "nccl: all-gather of the 16-byte (max, sum-exp) pairs + all-reduce(MAX) of the index"),
                       "l2_policy": "per-GPU inputs exceed the 126 MB L2 at N <= 8 (>= 0.88 GB per step)",
                       "duplicate_rows_fraction": dup_frac},
            "parity_check": check,
            "per_rank": {"candidate_kernel_ms": [float(x) for x in per_rank[:, 0]],
                         "candidates": [int(x) for x in per_rank[:, 1]],
                         "duplicate_rows_fraction": [float(x) for x in per_rank[:, 2]]},
            "e2e": {"value": P / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(sw.h_small.numel() * 8), "d2h_bytes_per_step": 32,
                    "what": "ShardedSweep.step: host (pinned) sweep inputs -> device, the step above, drawn index "
                            "+ log-sum-exp + status back to pinned host memory, every step"},
            "allgather_loglik": {"what": "plain scoring of the slice + NCCL all-gather of the fp64 log-weights so that "
                                         "every rank holds likelihoods[P] (the reference's data flow, "
                                         "logit_rules.py:861-863)",
                                 "ms_per_step": gather_ms, "value": P / (gather_ms * 1e-3), "unit": UNIT,
                                 "bytes_gathered_per_rank": int(8 * P) if world > 1 else 0},
            "gpu_launches": int(args.steps * 4),
            "clocks": clocks.summary(),
            "roofline": {"kernel": "score_w1d_kernel<6,stats> (the step's candidate kernel)", "bound": "hbm",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "kernel_ms": kern_stats_ms, "prepare_launch_ms_upper_bound": prepare_ms,
                         "algorithmic_bytes_per_candidate": 24,
                         "traffic": ncu_traffic_per_launch(n_local, "stats"),
                         "traffic_source": "profiles/score_w1d_traffic.json (ncu --set full capture of this kernel at "
                                           "the full S5 size, bytes per candidate) x candidates of this launch",
                         "plain_kernel": {"kernel": "score_w1d_kernel<6,plain> (mitre_score_all)", "kernel_ms": kern_plain_ms,
                                          "algorithmic_bytes_per_candidate": 16,
                                          "achieved": 16.0 * n_local / (kern_plain_ms * 1e-3) / 1e9,
                                          "frac": 16.0 * n_local / (kern_plain_ms * 1e-3) / 1e9 / peak,
                                          "traffic": ncu_traffic_per_launch(n_local, "plain"),
                                          "without_duplicate_awareness_ms": plain_nodedupe_ms},
                         "fp64": fp64_roofline(D, n_local, kern_stats_ms)},
        }
        if shard_build_s is not None:
            line["sharded_build_s"] = shard_build_s
        if world == 1:
            det_bytes = detector_bytes(pop, data)
            det_ms = build_marks.get("detector_kernel_ms", build_marks["values"] + build_marks["thresholds_rows"])
            line["detector_eval"] = {
                "groups": int(pop._G), "algorithmic_bytes": det_bytes, "ms": det_ms,
                "achieved_gbs": det_bytes / (det_ms * 1e-3) / 1e9,
                "frac_of_hbm_peak": det_bytes / (det_ms * 1e-3) / 1e9 / peak,
                "kernel": build_marks.get("detector_kernel", "phase-by-phase kernels + host"),
                "slope_groups_flagged": build_marks.get("slope_flagged"),
                "round1_kernels_ms": build_marks.get("detector_r1_kernels_ms"),
                "host_phase_ms": {"values": build_marks["values"], "thresholds_rows": build_marks["thresholds_rows"],
                                  "lexsort": build_marks["lexsort"]},
                "build_wall_s": build_marks["wall_s"]}
            line["e2e_full_vector"] = e2e_full_vector(args, torch, D, pop, state, N, P)
            line["e2e_dropin"] = e2e_dropin(args, torch, D, pop, state, N, P)
            if not args.no_cpu_baseline:
                line["cpu_baseline"] = cpu_baseline(args, pop, state, N, P, torch, D)
        print_line = line
    sw.close()
    del sw, prior, local_rows, scorer_out
    pop = None
    torch.cuda.empty_cache()
    chains_line = chains_throughput(torch, dist, world, args.seed)
    if rank == 0:
        print_line["chains"] = chains_line
        if world == 1 and not args.no_cpu_baseline:
            print_line["mcmc"] = mcmc_samples_per_sec(args)
        print(json.dumps(print_line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def e2e_dropin(args, torch, D, pop, state, N, P):
    """The ZERO-EDIT drop-in: mitre_b200.efficient_likelihoods.likelihoods_of_all_options with the reference's
    exact arguments (efficient_likelihoods.pyx:15-21) -- a float64[P', N] truth matrix in pageable host memory
    in, float64[P'] out -- on the first 8 Mi candidates of the benchmarked table (2.3 GB of float64 rows)."""
    from mitre_b200 import efficient_likelihoods as EL
    n = int(min(P, 8 << 20))
    rows = D.unpack_rows(pop.packed_truth[:n], N).cpu().numpy()
    opts = (rows * state["mask"]).astype(np.float64)                  # logit_rules.py:214-217
    del rows
    z = state["kappa"] / state["omega"]
    EL.likelihoods_of_all_options(state["X"], state["omega"], z, opts[:4096], SIGMA2)
    best, out = None, None
    for _ in range(3):
        t0 = time.perf_counter()
        out = EL.likelihoods_of_all_options(state["X"], state["omega"], z, opts, SIGMA2)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    dm = torch.as_tensor(D.pack_mask(state["mask"]), device=pop.packed_truth.device)
    want, _ = D.score_all(pop.packed_truth[:n].contiguous(), N, state["X"], state["omega"], state["kappa"], SIGMA2,
                          mask=dm)
    w = want.cpu().numpy()
    rel = float(np.max(np.abs(w - out) / np.abs(w)))      # not bit-equal: the entry receives z = kappa / omega
    return {"what": "efficient_likelihoods.likelihoods_of_all_options(X, omega, z, float64[P', N], sigma2) -> "
                    "float64[P'], host arrays in and out (rows packed to bits on the host, 8 B/candidate over PCIe)",
            "candidates": n, "value": n / best, "unit": UNIT, "ms": 1e3 * best,
            "host_bytes_read_per_candidate": 8 * N, "max_rel_diff_vs_device_path": rel}


def e2e_full_vector(args, torch, D, pop, state, N, P):
    """The reference-facing call a user of the drop-in makes, LogisticRuleModel.likelihoods_of_all_options:
    host inputs in, ALL P log-likelihoods back in host memory (2.35 GB per sweep over PCIe)."""
    from mitre_b200 import logit_rules as LR, rules as R
    model = LR.LogisticRuleModel.__new__(LR.LogisticRuleModel)
    model.data, model.rule_population = pop.data, pop
    model.prior_coefficient_variance = SIGMA2
    model._device_state = None
    # the public call takes a rule list + position; drive it with the staged X / mask of the bench state
    model._sweep_inputs = lambda rl, i, j: (state["X"], state["mask"])
    for _ in range(2):
        model.likelihoods_of_all_options(None, state["omega"], 0, 0)
    steps = 5
    t0 = time.perf_counter()
    for _ in range(steps):
        out = model.likelihoods_of_all_options(None, state["omega"], 0, 0)
    ms = 1e3 * (time.perf_counter() - t0) / steps
    return {"what": "LogisticRuleModel.likelihoods_of_all_options: pinned host inputs -> device, sweep, float64[P] "
                    "back to host memory", "value": P / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
            "d2h_bytes_per_step": int(8 * P), "finite": bool(np.isfinite(out[:1000]).all())}


def cpu_baseline(args, pop, state, N, P, torch, D):
    """The reference's compiled Cython kernel (oracle/_ref; C port if absent) on the host, one core
    (the reference is single-threaded), over lexsorted slices of the same candidate table."""
    import oracle
    from oracle import build_ref
    kind = "reference"
    try:
        ref = build_ref.load()
    except Exception:
        ref, kind = oracle, "port"
    chunk = min(P, 1500000)
    z = state["kappa"] / state["omega"]
    done, spent, marshal = 0, 0.0, 0.0
    rng = np.random.RandomState(args.seed + 2)
    check_rel = 0.0
    while spent < args.cpu_seconds and done < 40 * chunk:
        lo = int(rng.randint(0, max(P - chunk, 1)))
        rows = D.unpack_rows(pop.packed_truth[lo:lo + chunk], N).cpu().numpy().astype(bool)
        t0 = time.perf_counter()
        opts = (rows * state["mask"]).astype(np.float64)             # logit_rules.py:214-217
        t1 = time.perf_counter()
        ll = ref.likelihoods_of_all_options(state["X"], state["omega"], z, opts, np.float64(SIGMA2))
        t2 = time.perf_counter()
        marshal += t1 - t0
        spent += t2 - t0
        done += rows.shape[0]
        if check_rel == 0.0:        # parity spot check of the benchmarked table against the baseline
            dm = torch.as_tensor(D.pack_mask(state["mask"]), device=pop.packed_truth.device)
            got, _ = D.score_all(pop.packed_truth[lo:lo + chunk].contiguous(), N, state["X"], state["omega"],
                                 state["kappa"], SIGMA2, mask=dm)
            check_rel = float(np.max(np.abs(got.cpu().numpy() - ll) / np.abs(ll)))
    return {"value": done / spent, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "%d candidates in lexsorted slices of %d rows of the benchmarked table; per slice the "
                      "bool->float64 marshal (logit_rules.py:214-217, %.0f%% of the time) + the "
                      "efficient_likelihoods kernel" % (done, chunk, 100.0 * marshal / spent),
            "kernel_only_value": done / (spent - marshal), "max_rel_diff_vs_gpu": check_rel}


def chains_throughput(torch, dist, world, seed, chains_per_gpu=8, steps=100):
    """BASELINE configs[3] (64 independent chains of the Bokulich-shape problem over 8 GPUs, the reference's
    [replicates] mode, main.py:1913-1986): every rank builds the S2 model on its GPU and runs
    `chains_per_gpu` seeded replicas one after another, no collective while sampling; aggregate samples/s
    over the slowest rank."""
    from mitre_b200 import chains
    rank = int(os.environ.get("RANK", "0"))
    data, model = chains.build_model("S2", seed=seed)
    r0 = chains.initial_rule_list(model)
    chains.run_chain(model, 1, 3, r0, device_gibbs=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    lengths = []
    for c in range(chains_per_gpu):
        sampler = chains.run_chain(model, seed + rank * chains_per_gpu + c, steps, r0, device_gibbs=True)
        lengths.append(len(sampler.current_state.rl))
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return {"what": "independent S2 chains (replicas, seeds base + i), device draw + device Gibbs, %d per GPU "
                    "one after another, no collective while sampling" % chains_per_gpu,
            "chains": world * chains_per_gpu, "steps_per_chain": steps, "candidates": len(model.rule_population),
            "samples_per_sec": world * chains_per_gpu * steps / float(dt[0]), "wall_s": float(dt[0])}


def mcmc_samples_per_sec(args):
    """BASELINE.json's second metric: MCMC samples/s of the Bokulich-shaped problem (S2: N=35,
    V=200, 12 intervals, t in [1,180]) with the B200 likelihood path vs the same host sampler
    calling the reference's compiled Cython kernel on the host (same seed, same PG sampler)."""
    import torch
    from mitre_b200 import chains, logit_rules as LR
    from oracle import build_ref
    import oracle
    data, model = chains.build_model("S2", seed=args.seed)
    r0 = chains.initial_rule_list(model)
    chains.run_chain(model, 1, 3, r0)                      # warm-up
    torch.cuda.synchronize()
    steps = 300                                            # cfg2: 300 MCMC samples
    t0 = time.perf_counter()
    chains.run_chain(model, 2, steps, r0)
    gpu = steps / (time.perf_counter() - t0)
    chains.run_chain(model, 1, 3, r0, device_gibbs=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    chains.run_chain(model, 2, steps, r0, device_gibbs=True)
    gpu_gibbs = steps / (time.perf_counter() - t0)
    try:
        ref = build_ref.load()
        kind = "reference"
    except Exception:
        ref, kind = oracle, "port"
    flat_truth = model.rule_population.flat_truth                   # host bool[P, N], like the reference keeps it

    class ReferenceKernelModel(LR.LogisticRuleModel):
        """The reference's data flow for the two likelihood calls (logit_rules.py:171-229)."""
        def __init__(self, base):
            self.__dict__.update(base.__dict__)

        def likelihoods_z_given_Xs(self, omega, Xs):
            return np.array([oracle.likelihood_z_given_X(omega, X, self.data.y, self.prior_coefficient_variance)
                             for X in Xs])

        def likelihoods_of_all_options(self, rl, current_omega, i, j, N_workers=None):
            X, mask = self._sweep_inputs(rl, i, j)
            opts = (flat_truth * mask).astype(np.float64)
            z = (self.data.y - 0.5) / current_omega
            return ref.likelihoods_of_all_options(X.astype(np.float64), current_omega, z, opts,
                                                  np.float64(self.prior_coefficient_variance))

    cpu_model = ReferenceKernelModel(model)
    cpu_model.fast_prior = False          # the reference evaluates flat_prior over all P primitives
    cpu_steps = 8
    t0 = time.perf_counter()
    chains.run_chain(cpu_model, 2, cpu_steps, r0, device_draw=False)
    cpu = cpu_steps / (time.perf_counter() - t0)
    return {"workload": "S2 Bokulich-shape", "candidates": len(model.rule_population),
            "b200_samples_per_sec": gpu, "b200_steps": steps,
            "b200_device_gibbs_samples_per_sec": gpu_gibbs,
            "cpu_samples_per_sec": cpu, "cpu_steps": cpu_steps, "cpu_kernel": kind, "cpu_cores": 1,
            "note": "same host sampler, seed and Polya-Gamma sampler; the CPU arm follows the reference's data "
                    "flow (O(P) flat_prior, bool->float64 marshal, Cython kernel, numpy cumsum draw), the B200 "
                    "arm keeps the P-long vectors on the device (device_draw=True)"}


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())

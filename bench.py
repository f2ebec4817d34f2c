#!/usr/bin/env python
"""bench.py -- candidate rule-set likelihood evaluations per second (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

Workload (config.workload = "S5-primitive-sweep", BASELINE.json configs[4]): N = 35 subjects,
V = 10,000 aggregated variables, n_intervals = 24 with unrestricted window length (300 windows),
mean + slope detectors, full threshold grid  ->  ~4.08e8 candidate primitives per GPU, packed
3.3 GB -- far larger than the 126 MB L2, so no flush is needed between steps.  The table is
built on the GPU by this repo's detector kernels from synthetic series of that shape.

A step = one scoring sweep over ALL candidates (one `likelihoods_of_all_options` call of the
MCMC update_primitives move, logit_rules.py:851): base rule list of 3 rules + constant (p = 4),
fresh omega.

  value     evals/s with the sweep inputs already resident in HBM (kernel time only)
  e2e       evals/s through LogisticRuleModel.likelihoods_of_all_options-equivalent staging:
            every step copies (X, omega, kappa, mask) from pinned host memory to the device and
            the P log-likelihoods back to pinned host memory
  roofline  dominant kernel score_w1_kernel: algorithmic bytes 16 B/candidate (8 B packed row in,
            8 B log-likelihood out) / its launch time; peak = MEASURED_PEAKS.json hbm_gbs
  N > 1     weak scaling: every rank owns a full 10,000-variable shard of candidates (different
            series seed), receives the same sweep inputs, scores its shard, and all ranks
            all-gather the 16-byte (max, sum-exp) pair of their log-weights over NCCL -- the
            exchange step of the two-stage proposal draw.
"""
import argparse
import json
import math
import os
import sys
import threading
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
    ap.add_argument("--variables", type=int, default=10000, help="variables per GPU (S5: 10000)")
    ap.add_argument("--rules", type=int, default=3, help="rules in the base rule list (p = rules+1)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# clocks sampling (pynvml) during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

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


def ncu_traffic_per_launch(P):
    """dram bytes per launch of score_w1_kernel from the committed ncu capture, scaled to P."""
    path = os.path.join(ROOT, "profiles", "score_w1_traffic.json")
    try:
        rec = json.load(open(path))
        return float(rec["dram_bytes_per_candidate"]) * P
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
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
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
    del pop.calculate_values, pop.update_base_rules, pop.sort_base_rules
    # kernel-only time of detector evaluation: re-launch the fused kernel on the prepared inputs
    from mitre_b200 import device as D
    if pop._fused_args is not None:
        D.detectors_fused(*pop._fused_args)
        marks["detector_kernel_ms"] = cuda_time_ms(torch, lambda: D.detectors_fused(*pop._fused_args), 3)
        marks["detector_kernel"] = "values_resident_kernel + sort_emit_staged_kernel"
        from mitre_b200._lib import lib
        lib().mitre_set_detector_split(0)      # the single fused kernel, for comparison
        D.detectors_fused(*pop._fused_args)
        marks["detector_single_kernel_ms"] = cuda_time_ms(torch, lambda: D.detectors_fused(*pop._fused_args), 3)
        lib().mitre_set_detector_split(1)
    torch.cuda.empty_cache()
    return data, pop, marks


def detector_bytes(pop, data):
    """B_det = 8 V S + 8 S + sum_g [8 N + 8 K_g + 2 K_g 8 W64]."""
    N = data.n_subjects
    S = data.device_series()["S"]
    W64 = (N + 63) // 64
    K = pop._K_host.astype(np.int64)
    return 8 * data.n_variables * S + 8 * S + int(np.sum(8 * N + 8 * K + 2 * K * 8 * W64))


def run_b200(args):
    import torch
    from mitre_b200 import device as D
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

    data, pop, build_marks = build_population(args, rank, torch)
    N = data.n_subjects
    P = len(pop)
    W64 = D.words_for(N)
    packed = pop.packed_truth
    p = args.rules + 1

    # identical sweep inputs on every rank (the broadcast of the base system, KBs)
    def row_sampler(rng):
        return pop._truth_row(int(rng.randint(P)))
    state = sweep_state(N, args.rules, row_sampler, args.seed + 1) if rank == 0 else None
    if world > 1:
        box = [state]
        dist.broadcast_object_list(box, src=0)
        state = box[0]

    buffers = D.ScoreBuffers(P, N, p)
    n_small = N * p + 2 * N + W64
    h_small = torch.empty(n_small, dtype=torch.float64).pin_memory()
    hs = h_small.numpy()
    hs[:N * p] = state["X"].reshape(-1)
    hs[N * p:N * p + N] = state["omega"]
    hs[N * p + N:N * p + 2 * N] = state["kappa"]
    hs[N * p + 2 * N:] = D.pack_mask(state["mask"]).view(np.float64)
    d_small = torch.empty(n_small, dtype=torch.float64, device=dev)
    d_small.copy_(h_small)
    dX = d_small[:N * p].view(N, p)
    dom = d_small[N * p:N * p + N]
    dka = d_small[N * p + N:N * p + 2 * N]
    dmask = d_small[N * p + 2 * N:].view(dtype=torch.int64)
    h_out = torch.empty(P, dtype=torch.float64).pin_memory()
    draw = D.DrawBuffers(P)
    gathered = torch.empty(2 * world, dtype=torch.float64, device=dev)
    empty_rows = packed[:0]

    def sweep():
        # the timed step is the same at every N: each rank scores its own shard, nothing is exchanged
        out, _ = D.score_all(packed, N, dX, dom, dka, SIGMA2, mask=dmask, buffers=buffers)
        return out

    def sweep_with_exchange():
        # what the sampler's two-stage draw adds (parallel.py): (max, sum-exp) of the shard reduced in the
        # scoring kernel's epilogue, then a 16-byte-per-rank all-gather
        out, _, st = D.score_all_stats(packed, N, dX, dom, dka, SIGMA2, mask=dmask, buffers=buffers,
                                       draw_buffers=draw)
        if world > 1:
            dist.all_gather_into_tensor(gathered, st)
        return out

    def setup_only():
        D.score_all(empty_rows, N, dX, dom, dka, SIGMA2, mask=dmask, buffers=buffers)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM ----
    for _ in range(max(args.warmup, 3)):
        sweep()
    barrier()
    with ClockSampler(local_rank) as clocks:
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(args.steps):
            sweep()
        end.record()
        barrier()
        ms_total = start.elapsed_time(end)
    ms_step = ms_total / args.steps
    D.raise_on_status(buffers.status)

    # ---- e2e: host buffers in, host buffer out, every step ----
    def e2e_step():
        d_small.copy_(h_small, non_blocking=True)
        out = sweep()
        h_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        e2e_step()
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps

    # ---- the sampler's own end-to-end sweep (logit_rules.py:1052-1066 with device_draw=True): the P-long
    # vectors never leave the device; host buffers in, the drawn candidate index and log-total out ----
    g = torch.Generator(device=dev)
    g.manual_seed(args.seed + 7)
    d_prior = torch.randn(P, dtype=torch.float64, device=dev, generator=g) - math.log(max(P, 1))
    h_idx = torch.empty(1, dtype=torch.int64).pin_memory()
    h_stats = torch.empty(2, dtype=torch.float64).pin_memory()

    def e2e_draw_step(u=0.37):
        d_small.copy_(h_small, non_blocking=True)
        out, _, st = D.score_all_stats(packed, N, dX, dom, dka, SIGMA2, mask=dmask, buffers=buffers,
                                       log_prior=d_prior, draw_buffers=draw)
        idx = D.draw_index(out, d_prior, u, stabilise=True, buffers=draw, reuse_stats=True)
        h_idx.copy_(idx, non_blocking=True)
        h_stats.copy_(st, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        e2e_draw_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_draw_step()
    barrier()
    e2e_draw_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    del d_prior

    # the variant with the last subjects' table in shared memory, for comparison
    lib().mitre_set_split_tables(1)
    split_tables_ms = cuda_time_ms(torch, lambda: D.score_all(packed, N, dX, dom, dka, SIGMA2, mask=dmask,
                                                              buffers=buffers), max(args.steps, 5))
    lib().mitre_set_split_tables(0)

    # kernel-only time of the dominant kernel: the step minus the (tiny) setup launch
    setup_ms = cuda_time_ms(torch, setup_only, 50)
    kernel_ms = cuda_time_ms(torch, lambda: D.score_all(packed, N, dX, dom, dka, SIGMA2, mask=dmask,
                                                        buffers=buffers), max(args.steps, 5)) - setup_ms

    # the sampler's exchange step, timed apart from the headline (same CUDA-event recipe)
    barrier()
    exch_ms = cuda_time_ms(torch, sweep_with_exchange, max(args.steps, 5))
    barrier()

    # ---- max over ranks, totals ----
    t = torch.tensor([ms_step, e2e_ms, float(P), exch_ms, e2e_draw_ms], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_step, e2e_ms, P_total, exch_ms = float(tmax[0]), float(tmax[1]), float(tsum[2]), float(tmax[3])
        e2e_draw_ms = float(tmax[4])
    else:
        P_total = float(P)

    if rank == 0:
        peak, peak_src = peaks()
        achieved = 16.0 * P / (kernel_ms * 1e-3) / 1e9
        det_bytes = detector_bytes(pop, data)
        det_ms = build_marks.get("detector_kernel_ms", build_marks["values"] + build_marks["thresholds_rows"])
        line = {
            "metric": METRIC, "value": P_total / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "S5-primitive-sweep", "subjects": N, "variables_per_gpu": args.variables,
                       "windows": len(pop.windows), "candidates_per_gpu": int(P),
                       "base_columns_p": p, "l2_policy": "inputs (3.3 GB/GPU) exceed the 126 MB L2",
                       "exchange": "none: candidate shards are independent (see sampler_exchange)"},
            "sampler_exchange": {"what": "scoring with the (max, sum-exp) reduction fused into the epilogue + "
                                         "all-gather of one 16-byte pair per rank (two-stage draw)",
                                 "ms_per_step": exch_ms, "value": P_total / (exch_ms * 1e-3), "unit": UNIT,
                                 "bytes_on_wire_per_rank": 16 if world > 1 else 0},
            "e2e": {"value": P_total / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(n_small * 8), "d2h_bytes_per_step": int(P * 8),
                    "ms_per_step": e2e_ms},
            "e2e_sampler_sweep": {"what": "host inputs -> scoring + log-prior + (max, sum-exp) -> on-device draw -> "
                                          "drawn index and stats back to pinned host memory (the sampler's "
                                          "device_draw sweep; the P-long vectors stay in HBM)",
                                  "value": P_total / (e2e_draw_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_draw_ms,
                                  "h2d_bytes_per_step": int(n_small * 8), "d2h_bytes_per_step": 24},
            "gpu_launches": int(args.steps * 3),
            "clocks": clocks.summary(),
            "roofline": {"kernel": "score_w1g4_kernel", "step_ms_with_last_table_in_smem": split_tables_ms, "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                         "traffic": ncu_traffic_per_launch(P), "kernel_ms": kernel_ms,
                         "setup_ms": setup_ms, "algorithmic_bytes_per_candidate": 16,
                         "fp64": fp64_roofline(D, P, kernel_ms)},
            "detector_eval": {"groups": int(pop._G), "algorithmic_bytes": det_bytes,
                              "ms": det_ms, "achieved_gbs": det_bytes / (det_ms * 1e-3) / 1e9,
                              "frac_of_hbm_peak": det_bytes / (det_ms * 1e-3) / 1e9 / peak,
                              "kernel": build_marks.get("detector_kernel", "phase-by-phase kernels + host"),
                              "single_fused_kernel_ms": build_marks.get("detector_single_kernel_ms"),
                              "host_phase_ms": {"values": build_marks["values"],
                                                "thresholds_rows": build_marks["thresholds_rows"],
                                                "lexsort": build_marks["lexsort"]},
                              "build_wall_s": build_marks["wall_s"]},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, pop, state, N, P, torch, D)
            del pop, packed
            torch.cuda.empty_cache()
            line["mcmc"] = mcmc_samples_per_sec(args)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


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
    steps = 40
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

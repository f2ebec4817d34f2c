"""Torch-tensor wrappers over the C ABI (include/mitre_b200.h).

PyTorch is only the carrier here: it owns device memory and the current stream; every compute
step is a kernel in libmitre_b200.so.  All functions take / return CUDA tensors, are
asynchronous on torch's current stream, and raise if no CUDA device is present -- there is no
CPU path.

Packed truth rows are carried as torch.int64 tensors [P, W64] holding the uint64 bit patterns
(subject i -> word i//64, bit 63 - i%64; see the header).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, lib

F64 = torch.float64
I64 = torch.int64
I32 = torch.int32
U8 = torch.uint8


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("mitre_b200 needs a CUDA device (built for sm_100a / B200); "
                           "there is no CPU fallback")


def dev():
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def words_for(N):
    return (int(N) + 63) // 64


def _dev_f64(a):
    if isinstance(a, torch.Tensor):
        return a.to(device=dev(), dtype=F64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=dev())


def device_info():
    sm, ma, mi = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    smem = ctypes.c_size_t()
    check(lib().mitre_device_info(ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi),
                                  ctypes.byref(smem)), "mitre_device_info")
    return dict(sm_count=sm.value, cc=(ma.value, mi.value), smem_optin=smem.value)


def probe_fp64_tflops(iters=20000, reps=3):
    """Measured fp64 FMA throughput of the current device (TFLOP/s), CUDA-event timed."""
    sink = torch.zeros(1, dtype=F64, device=dev())
    flops = ctypes.c_double(0.0)
    check(lib().mitre_probe_fp64(_stream(), iters, _p(sink), ctypes.byref(flops)), "mitre_probe_fp64")
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        check(lib().mitre_probe_fp64(_stream(), iters, _p(sink), ctypes.byref(flops)), "mitre_probe_fp64")
        e.record()
        torch.cuda.synchronize()
        best = max(best, flops.value / (s.elapsed_time(e) * 1e-3) / 1e12)
    return best


# ---- packing ---------------------------------------------------------------------------------

def pack_mask(mask_bool):
    """Host helper: bool[N] -> packed int64[W64] numpy (bit layout of the header)."""
    m = np.asarray(mask_bool, dtype=bool)
    N = m.shape[0]
    out = np.zeros(words_for(N), dtype=np.uint64)
    idx = np.nonzero(m)[0]
    np.bitwise_or.at(out, idx // 64, np.uint64(1) << (np.uint64(63) - (idx % 64).astype(np.uint64)))
    return out.view(np.int64)


def pack_rows(truth):
    """bool / uint8 / float64 [P, N] (tensor or ndarray) -> packed int64[P, W64] on the device."""
    require_cuda()
    if not isinstance(truth, torch.Tensor):
        truth = np.asarray(truth)
        if truth.dtype == np.bool_:
            truth = truth.view(np.uint8)
        truth = torch.as_tensor(np.ascontiguousarray(truth), device=dev())
    if truth.dtype == torch.bool:
        truth = truth.view(U8)
    truth = truth.to(dev()).contiguous()
    P, N = truth.shape
    packed = torch.empty((P, words_for(N)), dtype=I64, device=dev())
    if truth.dtype == F64:
        check(lib().mitre_pack_rows_f64(_stream(), _p(truth), P, N, _p(packed)), "mitre_pack_rows_f64")
    elif truth.dtype == U8:
        check(lib().mitre_pack_rows_u8(_stream(), _p(truth), P, N, _p(packed)), "mitre_pack_rows_u8")
    else:
        raise TypeError("truth rows must be bool, uint8 or float64, got %s" % truth.dtype)
    return packed


def unpack_rows(packed, N):
    """packed int64[P, W64] -> uint8[P, N] (0/1) on the device."""
    P = packed.shape[0]
    out = torch.empty((P, N), dtype=U8, device=packed.device)
    check(lib().mitre_unpack_rows_u8(_stream(), _p(packed), P, int(N), _p(out)), "mitre_unpack_rows_u8")
    return out


# ---- scoring -----------------------------------------------------------------------------------

class ScoreBuffers:
    """Reusable per-model device buffers for mitre_score_all (workspace, status, output)."""

    def __init__(self, P, N, max_p=_lib.MAX_P):
        self.P, self.N = int(P), int(N)
        nbytes = lib().mitre_score_workspace_bytes(self.N, max_p)
        self.workspace = torch.empty(nbytes, dtype=U8, device=dev())
        self.status = torch.zeros(1, dtype=I32, device=dev())
        self.out = torch.empty(self.P, dtype=F64, device=dev())


def score_all(packed, N, X, omega, kappa, sigma2, mask=None, buffers=None, out=None):
    """log N(kappa/omega; 0, Omega^-1 + sigma2 [X|v_k][X|v_k]^T) for every packed row k.

    packed int64[P, W64]; X [N, p]; omega, kappa [N]; mask: packed int64[W64] tensor or None.
    Returns (out float64[P], status int32[1]) on the device, asynchronously."""
    P, W64 = packed.shape
    X = _dev_f64(X)
    omega = _dev_f64(omega)
    kappa = _dev_f64(kappa)
    p = X.shape[1]
    if buffers is None:
        buffers = ScoreBuffers(P, N, p)
    if out is None:
        out = buffers.out if buffers.out.shape[0] == P else torch.empty(P, dtype=F64, device=dev())
    ws = buffers.workspace
    check(lib().mitre_score_all(_stream(), _p(packed), P, W64, _p(mask), _p(X), _p(omega), _p(kappa),
                                int(N), int(p), float(sigma2), _p(ws), ws.numel(), _p(out),
                                _p(buffers.status)), "mitre_score_all")
    return out, buffers.status


def score_all_stats(packed, N, X, omega, kappa, sigma2, mask=None, buffers=None, out=None, log_prior=None,
                    draw_buffers=None):
    """score_all + (max, sum exp(w - max)) of w = out (+ log_prior) in the same launch.
    Returns (out, status, stats float64[2])."""
    P, W64 = packed.shape
    X = _dev_f64(X)
    omega = _dev_f64(omega)
    kappa = _dev_f64(kappa)
    p = X.shape[1]
    if buffers is None:
        buffers = ScoreBuffers(P, N, p)
    if out is None:
        out = buffers.out if buffers.out.shape[0] == P else torch.empty(P, dtype=F64, device=dev())
    if draw_buffers is None:
        draw_buffers = DrawBuffers(max(P, 1))
    ws = buffers.workspace
    check(lib().mitre_score_all_stats(_stream(), _p(packed), P, W64, _p(mask), _p(X), _p(omega), _p(kappa),
                                      int(N), int(p), float(sigma2), _p(ws), ws.numel(), _p(out),
                                      _p(buffers.status), _p(log_prior), _p(draw_buffers.stats),
                                      _p(draw_buffers.workspace), draw_buffers.workspace.numel()),
          "mitre_score_all_stats")
    return out, buffers.status, draw_buffers.stats


def score_groups(values, thresholds, K, row_offsets, P, X, omega, kappa, sigma2, mask=None, buffers=None,
                 out=None, group_prior=None):
    """Log-likelihood (+ group_prior[g], if given) of every candidate of G groups in ENUMERATION order,
    from the detector values (mitre_score_groups: suffix sums along each group's sorted order; no
    truth rows).  Returns (out float64[P], status)."""
    G, N = values.shape
    X = _dev_f64(X)
    omega = _dev_f64(omega)
    kappa = _dev_f64(kappa)
    p = X.shape[1]
    if buffers is None:
        buffers = ScoreBuffers(max(P, 1), N, p)
    if out is None:
        out = buffers.out[:P] if buffers.out.shape[0] >= P else torch.empty(P, dtype=F64, device=dev())
    ws = buffers.workspace
    check(lib().mitre_score_groups(_stream(), _p(values), _p(thresholds), _p(K), _p(row_offsets), G, _p(mask),
                                   _p(X), _p(omega), _p(kappa), int(N), int(p), float(sigma2), _p(ws), ws.numel(),
                                   _p(group_prior), _p(out), _p(buffers.status)), "mitre_score_groups")
    return out, buffers.status


def raise_on_status(status):
    """Synchronises; maps device status flags to numpy's LinAlgError like the reference
    (efficient_likelihoods.pyx:42 -> np.linalg.cholesky)."""
    s = int(status.item())
    if s:
        status.zero_()
        raise np.linalg.LinAlgError("Matrix is not positive definite (status flags %d)" % s)


def read_with_status(values, status):
    """values (any small device tensor) as a float64 numpy array, fetched together with the status
    word in ONE device->host copy; non-zero status raises LinAlgError like raise_on_status.  Every
    separate .cpu() / .item() is a stream synchronisation (~50 us): at S2 sizes they were a fifth of a
    sampler step."""
    packed = torch.cat([values.reshape(-1).to(F64), status.to(F64)]).cpu().numpy()
    if packed[-1] != 0:
        status.zero_()
        raise np.linalg.LinAlgError("Matrix is not positive definite (status flags %d)" % int(packed[-1]))
    return packed[:-1]


def likelihood_z_given_X_batch(Xs, omega, kappa, sigma2, p_b=None):
    """Xs [B, N, pmax] -> float64[B] dense log-likelihoods (logit_rules.py:171-177)."""
    Xs = _dev_f64(Xs)
    B, N, pmax = Xs.shape
    omega = _dev_f64(omega)
    kappa = _dev_f64(kappa)
    pb = None
    if p_b is not None:
        pb = torch.as_tensor(np.asarray(p_b, dtype=np.int32), device=dev())
    out = torch.empty(B, dtype=F64, device=dev())
    status = torch.zeros(1, dtype=I32, device=dev())
    check(lib().mitre_likelihood_z_given_X_batch(_stream(), _p(Xs), _p(pb), pmax, _p(omega),
                                                 _p(kappa), N, B, float(sigma2), _p(out),
                                                 _p(status)), "mitre_likelihood_z_given_X_batch")
    return out, status


def mc_logpdf(x, cov):
    x = _dev_f64(x)
    cov = _dev_f64(cov)
    N = x.shape[0]
    scratch = torch.empty((N, N), dtype=F64, device=dev())
    out = torch.empty(1, dtype=F64, device=dev())
    status = torch.zeros(1, dtype=I32, device=dev())
    check(lib().mitre_mc_logpdf(_stream(), _p(x), _p(cov), N, _p(scratch), _p(out), _p(status)),
          "mitre_mc_logpdf")
    return out, status


# ---- detector evaluation ---------------------------------------------------------------------

def window_ranges(times, t_offsets, windows):
    """times f64[S_pad], t_offsets i32[N+1], windows f64[W,2] -> (lo, cnt) int32[W, N]."""
    N = t_offsets.shape[0] - 1
    W = windows.shape[0]
    lo = torch.empty((W, N), dtype=I32, device=dev())
    cnt = torch.empty((W, N), dtype=I32, device=dev())
    check(lib().mitre_window_ranges(_stream(), _p(times), _p(t_offsets), N, _p(windows), W, _p(lo),
                                    _p(cnt)), "mitre_window_ranges")
    return lo, cnt


def detector_values(series, times, t_offsets, windows, lo, cnt, group_base, slope_ok, G, max_T):
    V, stride = series.shape
    N = t_offsets.shape[0] - 1
    W = windows.shape[0]
    values = torch.empty((G, N), dtype=F64, device=dev())
    check(lib().mitre_detector_values(_stream(), _p(series), stride, _p(times), _p(t_offsets), N, V,
                                      _p(windows), _p(lo), _p(cnt), _p(group_base), _p(slope_ok), W,
                                      int(max_T), _p(values)), "mitre_detector_values")
    return values


def detector_thresholds(values):
    G, N = values.shape
    thresholds = torch.empty((G, max(N - 1, 0)), dtype=F64, device=dev())
    K = torch.empty(G, dtype=I32, device=dev())
    check(lib().mitre_detector_thresholds(_stream(), _p(values), G, N, _p(thresholds), _p(K)),
          "mitre_detector_thresholds")
    return thresholds, K


def detector_row_offsets(K):
    G = K.shape[0]
    nbytes = lib().mitre_scan_workspace_bytes(G)
    ws = torch.empty(nbytes, dtype=U8, device=dev())
    offsets = torch.empty(G + 1, dtype=I64, device=dev())
    check(lib().mitre_detector_row_offsets(_stream(), _p(K), G, _p(offsets), _p(ws), nbytes),
          "mitre_detector_row_offsets")
    return offsets


def detector_rows(values, thresholds, K, row_offsets, P):
    G, N = values.shape
    packed = torch.empty((P, words_for(N)), dtype=I64, device=dev())
    check(lib().mitre_detector_rows(_stream(), _p(values), _p(thresholds), _p(K), _p(row_offsets), G,
                                    N, _p(packed)), "mitre_detector_rows")
    return packed


def window_lists(times, t_offsets, windows, vstride):
    """Variable-independent per-window tables for the fused detector kernel (mitre_window_lists)."""
    N = t_offsets.shape[0] - 1
    W = windows.shape[0]
    ld = times.shape[0]
    t = dict(lo=torch.empty((W, N), dtype=I32, device=dev()),
             cnt=torch.empty((W, N), dtype=I32, device=dev()),
             start=torch.empty((W, N), dtype=I32, device=dev()),
             inv_n=torch.empty((W, N), dtype=F64, device=dev()),
             inv_sxx=torch.empty((W, N), dtype=F64, device=dev()),
             idx=torch.zeros((W, ld), dtype=I32, device=dev()),
             dt=torch.zeros((W, ld), dtype=F64, device=dev()), ld=ld)
    check(lib().mitre_window_lists(_stream(), _p(times), _p(t_offsets), N, _p(windows), W, ld, int(vstride),
                                   _p(t['lo']),
                                   _p(t['cnt']), _p(t['start']), _p(t['inv_n']), _p(t['inv_sxx']),
                                   _p(t['idx']), _p(t['dt'])), "mitre_window_lists")
    return t


def detectors_fused(series_t, N, V, lists, group_base, slope_ok, task_window, task_type, G, group_type=None):
    """mitre_detectors_fused (values kernel + sort/emit kernel, no host round trip): values f64[G, N],
    thresholds f64[G, N-1], K i32[G], rows_padded i64[G, 2(N-1)].  With group_type (uint8[G], 0 = slope)
    also the slope guard band: returns additionally (group_flags uint8[G], flagged int64[1])."""
    L = 2 * (N - 1)
    values = torch.empty((G, N), dtype=F64, device=dev())
    thresholds = torch.empty((G, N - 1), dtype=F64, device=dev())
    K = torch.empty(G, dtype=I32, device=dev())
    rows = torch.empty((G, L), dtype=I64, device=dev())
    flags = torch.empty(G, dtype=U8, device=dev()) if group_type is not None else None
    flagged = torch.zeros(1, dtype=I64, device=dev()) if group_type is not None else None
    check(lib().mitre_detectors_fused(_stream(), _p(series_t), series_t.shape[1], int(N), int(V),
                                      _p(lists['lo']), _p(lists['cnt']), _p(lists['start']),
                                      _p(lists['inv_n']), _p(lists['inv_sxx']), _p(lists['idx']),
                                      _p(lists['dt']), lists['ld'], _p(group_base), _p(slope_ok),
                                      _p(task_window), _p(task_type), task_window.shape[0],
                                      lists['lo'].shape[0], int(G), _p(values),
                                      _p(thresholds), _p(K), _p(rows), _p(group_type), _p(flags), _p(flagged)),
          "mitre_detectors_fused")
    if group_type is not None:
        return values, thresholds, K, rows, flags, flagged
    return values, thresholds, K, rows


def window_tbar(times, windows, lo, cnt):
    """float64[W, N]: mean over each (window, subject)'s samples of t - t0 (mitre_window_tbar)."""
    W, N = lo.shape
    tbar = torch.empty((W, N), dtype=F64, device=dev())
    check(lib().mitre_window_tbar(_stream(), _p(times), N, _p(windows), W, _p(lo), _p(cnt), _p(tbar)),
          "mitre_window_tbar")
    return tbar


def values_walk_plan(N, S, V, sg_first, info=None):
    """How the walk-forward values kernel splits the start groups over blockIdx.y: (sg_per_block,
    max_block_windows) such that the per-block window tables fit in shared memory and the grid fills whole
    waves, or None when not even one start group fits (then: mitre_detectors_fused).
    sg_first: host int array [n_sg + 1].  info: {'smem_optin', 'sm_count'} (default: the current device)."""
    info = device_info() if info is None else info
    n_sg = len(sg_first) - 1
    vb = lib().mitre_walk_vars_per_block(int(N))
    tiles = (int(V) + vb - 1) // vb
    # A CTA costs a fixed ~2.5 start groups' worth of time (tile and table loads behind one barrier, the tail
    # of its slowest warps with nothing else resident on the SM) on top of the start groups it walks, and the
    # grid runs in waves of one CTA per SM: take the split with the smallest waves x (2.5 + start groups per
    # CTA).  Measured at S5 (24 start groups, 625 tiles): 12 per CTA 0.73 ms, 8: 0.77, 6: 0.81, 1: 1.46.
    best = None
    for chunks in range(1, n_sg + 1):
        per = (n_sg + chunks - 1) // chunks
        widest = max(int(sg_first[min(k + per, n_sg)] - sg_first[k]) for k in range(0, n_sg, per))
        smem = lib().mitre_values_walk_smem_bytes(int(N), int(S), widest)
        if not 0 < smem <= info['smem_optin']:
            continue
        blocks = tiles * ((n_sg + per - 1) // per)
        waves = -(-blocks // info['sm_count'])
        cost = waves * (2.5 + per)
        if best is None or cost < best[0]:
            best = (cost, per, widest)
    return None if best is None else (best[1], best[2])


def walk_stages(sg_first, group_base_host, G, n_stages):
    """Cut the start groups into ~n_stages runs of about equal group counts for the pipelined detector call:
    (stage_sg int32[k + 1], stage_group int64[k + 1]) host arrays."""
    n_sg = len(sg_first) - 1
    gb = [int(group_base_host[int(sg_first[k])]) for k in range(n_sg)] + [int(G)]
    sg_bounds, k = [0], 0
    for c in range(1, n_stages):
        target = G * c / float(n_stages)
        while k < n_sg and gb[k] < target:
            k += 1
        if sg_bounds[-1] < k < n_sg:
            sg_bounds.append(k)
    sg_bounds.append(n_sg)
    return (np.array(sg_bounds, dtype=np.int32), np.array([gb[k] for k in sg_bounds], dtype=np.int64))


def detectors_walk(series, times, t_offsets, S, N, V, windows, lists, tbar, group_base, slope_ok, sg_first, G,
                   group_type, plan, stages=None):
    """The default device path for N <= 63: mitre_detector_values_walk (one thread per (variable, subject),
    samples walked once per start group) + mitre_detector_sort_emit; with stages = walk_stages(...) the two
    kernels are pipelined over runs of start groups on two internal streams (mitre_detectors_walk).  Same
    returns as detectors_fused with group_type: (values, thresholds, K, rows_padded, group_flags, flagged).
    plan = values_walk_plan(...)."""
    L = 2 * (N - 1)
    values = torch.empty((G, N), dtype=F64, device=dev())
    thresholds = torch.empty((G, N - 1), dtype=F64, device=dev())
    K = torch.empty(G, dtype=I32, device=dev())
    rows = torch.empty((G, L), dtype=I64, device=dev())
    flags = torch.zeros(G, dtype=U8, device=dev())
    flagged = torch.zeros(1, dtype=I64, device=dev())
    W = windows.shape[0]
    if stages is not None and len(stages[0]) > 2:
        st_sg, st_g = stages
        check(lib().mitre_detectors_walk(_stream(), _p(series), series.shape[1], _p(times), _p(t_offsets), int(S),
                                         int(N), int(V), _p(windows), W, _p(lists['lo']), _p(lists['cnt']),
                                         _p(lists['inv_sxx']), _p(tbar), _p(group_base), _p(slope_ok), _p(sg_first),
                                         sg_first.shape[0] - 1, int(plan[0]), int(plan[1]),
                                         st_sg.ctypes.data_as(ctypes.c_void_p), st_g.ctypes.data_as(ctypes.c_void_p),
                                         len(st_sg) - 1, int(G), _p(values), _p(thresholds), _p(K), _p(rows),
                                         _p(group_type), _p(flags), _p(flagged)), "mitre_detectors_walk")
        return values, thresholds, K, rows, flags, flagged
    check(lib().mitre_detector_values_walk(_stream(), _p(series), series.shape[1], _p(times), _p(t_offsets), int(S),
                                           int(N), int(V), _p(windows), W, _p(lists['lo']), _p(lists['cnt']),
                                           _p(lists['inv_sxx']), _p(tbar), _p(group_base), _p(slope_ok),
                                           _p(sg_first), sg_first.shape[0] - 1, int(plan[0]), int(plan[1]),
                                           _p(values), _p(flags)),
          "mitre_detector_values_walk")
    check(lib().mitre_detector_sort_emit(_stream(), _p(values), G, int(N), _p(thresholds), _p(K), _p(rows),
                                         _p(group_type), _p(flags), _p(flagged)), "mitre_detector_sort_emit")
    return values, thresholds, K, rows, flags, flagged


def slope_guard(values, group_type):
    """Slope guard band over values f64[G, N] (any N): (group_flags uint8[G], flagged int64[1])."""
    G, N = values.shape
    flags = torch.zeros(G, dtype=U8, device=dev())
    flagged = torch.zeros(1, dtype=I64, device=dev())
    if G:
        check(lib().mitre_slope_guard(_stream(), _p(values), G, N, _p(group_type), _p(flags), _p(flagged)),
              "mitre_slope_guard")
    return flags, flagged


def selftest_divide(a, n):
    a = _dev_f64(a)
    n = torch.as_tensor(np.asarray(n, dtype=np.int32), device=dev())
    fast = torch.empty_like(a)
    exact = torch.empty_like(a)
    check(lib().mitre_selftest_divide(_stream(), _p(a), _p(n), a.shape[0], _p(fast), _p(exact)),
          "mitre_selftest_divide")
    return fast, exact


def densify_perm(perm_padded, N, row_offsets):
    P = perm_padded.shape[0]
    perm = torch.empty(P, dtype=I64, device=dev())
    check(lib().mitre_densify_perm(_stream(), _p(perm_padded), P, int(N), _p(row_offsets), _p(perm)),
          "mitre_densify_perm")
    return perm


def lexsort_rows(rows, N, want_sorted=True):
    """Stable lexicographic sort (np.lexsort(np.rot90(truth)) order).  Returns (perm int64[P],
    sorted_rows int64[P, W64] or None)."""
    P, W64 = rows.shape
    perm = torch.empty(P, dtype=I64, device=dev())
    sorted_rows = torch.empty_like(rows) if want_sorted else None
    nbytes = lib().mitre_lexsort_workspace_bytes(P, W64)
    ws = torch.empty(max(nbytes, 1), dtype=U8, device=dev())
    check(lib().mitre_lexsort_rows(_stream(), _p(rows), P, W64, int(N), _p(perm), _p(sorted_rows),
                                   _p(ws), nbytes), "mitre_lexsort_rows")
    if W64 == 1:
        check(lib().mitre_lexsort_check(_stream(), _p(ws), P, int(N)), "mitre_lexsort_check")
    return perm, sorted_rows


def lexsort_padded(rows_padded, N, P_valid, row_offsets):
    """sort_base_rules straight from the padded detector output (2(N-1) slots per group, unused = ~0):
    (perm int64[P_valid] of dense enumeration indices, sorted_rows int64[P_valid, 1]).  Tables that are not
    eligible for the packed one-sweep sort (mitre_lexsort_padded) take the general sort over N + 1 bits and
    mitre_densify_perm."""
    P_slots = rows_padded.numel()
    nbytes = lib().mitre_lexsort_workspace_bytes(P_slots, 1)
    ws = torch.empty(max(nbytes, 1), dtype=U8, device=dev())
    perm = torch.empty(P_valid, dtype=I64, device=dev())
    sorted_rows = torch.empty((P_valid, 1), dtype=I64, device=dev())
    rc = lib().mitre_lexsort_padded(_stream(), _p(rows_padded), P_slots, int(N), int(P_valid), _p(row_offsets),
                                    _p(perm), _p(sorted_rows), _p(ws), nbytes)
    if rc == -2:   # MITRE_ERR_UNSUPPORTED
        del perm, sorted_rows, ws
        perm_pad, sorted_pad = lexsort_rows(rows_padded.view(-1, 1), N + 1)
        return densify_perm(perm_pad[:P_valid], N, row_offsets), sorted_pad[:P_valid].contiguous()
    check(rc, "mitre_lexsort_padded")
    check(lib().mitre_lexsort_check(_stream(), _p(ws), P_slots, int(N)), "mitre_lexsort_check")
    return perm, sorted_rows


# ---- draw ------------------------------------------------------------------------------------------

class DrawBuffers:
    def __init__(self, P):
        nbytes = lib().mitre_draw_workspace_bytes(int(P))
        self.workspace = torch.empty(nbytes, dtype=U8, device=dev())
        self.stats = torch.empty(2, dtype=F64, device=dev())
        self.index = torch.empty(1, dtype=I64, device=dev())


def weight_stats(a, b=None, buffers=None):
    """(max, sum exp(w - max)) of w = a (+ b) as a float64[2] device tensor."""
    P = a.shape[0]
    buffers = buffers or DrawBuffers(P)
    check(lib().mitre_weight_stats(_stream(), _p(a), _p(b), P, _p(buffers.stats),
                                   _p(buffers.workspace), buffers.workspace.numel()),
          "mitre_weight_stats")
    return buffers.stats


def draw_index(a, b, u, stabilise=False, buffers=None, reuse_stats=False, reuse_tiles=False):
    """#{cumsum(exp(w)) < u * total} (rules.py:1261-1264) as an int64[1] device tensor.
    stabilise: subtract max(w) before exp; reuse_stats: buffers.stats already holds (max, sum-exp) of
    exactly these weights (e.g. from score_all_stats), so the extra pass over w is skipped;
    reuse_tiles: buffers.workspace also still holds the per-tile pairs score_all_stats left for exactly
    these weights (a = out, b = log_prior), so NO pass over the P weights is made at all."""
    P = a.shape[0]
    buffers = buffers or DrawBuffers(P)
    if reuse_tiles:
        check(lib().mitre_draw_from_tiles(_stream(), _p(a), _p(b), P, float(u), 1 if stabilise else 0,
                                          _p(buffers.stats), _p(buffers.index), _p(buffers.workspace),
                                          buffers.workspace.numel()), "mitre_draw_from_tiles")
        return buffers.index
    mode = (2 if reuse_stats else 1) if stabilise else 0
    check(lib().mitre_draw_index(_stream(), _p(a), _p(b), P, float(u), mode,
                                 _p(buffers.stats), _p(buffers.index), _p(buffers.workspace),
                                 buffers.workspace.numel()), "mitre_draw_index")
    return buffers.index


# ---- primitive prior ---------------------------------------------------------------------------

def row_groups(perm, row_offsets):
    """int32[P]: the group of every lexsorted row."""
    P = perm.shape[0]
    G = row_offsets.shape[0] - 1
    out = torch.empty(P, dtype=I32, device=dev())
    check(lib().mitre_row_groups(_stream(), _p(perm), P, _p(row_offsets), G, _p(out)), "mitre_row_groups")
    return out


def row_prior(row_group, group_variable, group_window, by_variable, by_window, norm, out=None):
    """out[i] = by_variable[gvar[g_i]] + by_window[gwin[g_i]] - norm (flat_prior, logit_rules.py:424-457)."""
    P = row_group.shape[0]
    if out is None:
        out = torch.empty(P, dtype=F64, device=dev())
    check(lib().mitre_row_prior(_stream(), _p(row_group), P, _p(group_variable), _p(group_window),
                                _p(by_variable), _p(by_window), float(norm), _p(out)), "mitre_row_prior")
    return out


# ---- omega / beta Gibbs loop ---------------------------------------------------------------------

def gibbs_omega_beta(X, kappa, prior_variance, beta, n_iters, seed, omega=None):
    """n_iters x (omega ~ PG(1, X beta), beta | omega) in one launch (omega=None), or n_iters beta
    draws for a given omega.  Returns (omega float64[N], beta float64[p], betas float64[n_iters, p],
    status) as CUDA tensors (gibbs_read_back brings all of them to the host with one copy)."""
    X = _dev_f64(X)
    kappa = _dev_f64(kappa)
    N, p = X.shape
    buf = torch.empty(n_iters * p + N + p, dtype=F64, device=dev())
    betas = buf[:n_iters * p].view(n_iters, p)
    om = buf[n_iters * p:n_iters * p + N]
    beta_io = buf[n_iters * p + N:]
    beta_io.copy_(_dev_f64(beta))
    mode = 0 if omega is None else 1
    if omega is not None:
        om.copy_(_dev_f64(omega))
    status = torch.zeros(1, dtype=I32, device=dev())
    check(lib().mitre_gibbs_omega_beta(_stream(), _p(X), _p(kappa), N, p, float(prior_variance), int(n_iters),
                                       mode, int(seed) & 0xFFFFFFFFFFFFFFFF, _p(beta_io), _p(om), _p(betas),
                                       _p(status)), "mitre_gibbs_omega_beta")
    return om, beta_io, betas, status


def gibbs_read_back(om, beta_io, betas, status):
    """(omega, beta, betas) of gibbs_omega_beta as numpy arrays with ONE device->host copy (plus the
    status word, raised as LinAlgError like raise_on_status)."""
    n_iters, p = betas.shape
    N = om.shape[0]
    packed = torch.cat([betas.reshape(-1), om, beta_io, status.to(F64)]).cpu().numpy()
    if packed[-1] != 0:
        status.zero_()
        raise np.linalg.LinAlgError("Matrix is not positive definite (status flags %d)" % int(packed[-1]))
    return (packed[n_iters * p:n_iters * p + N].copy(), packed[n_iters * p + N:n_iters * p + N + p].copy(),
            packed[:n_iters * p].reshape(n_iters, p).copy())


class GibbsWorkspace:
    """Persistent staging for the sampler's omega / beta sweeps: ONE pinned -> device copy in (X, kappa,
    beta, cleared status slot), one launch, ONE device -> pinned copy out (betas, omega, beta, status), one
    synchronisation.  gibbs_omega_beta + gibbs_read_back made three pageable host->device copies, a fill
    kernel and a concatenation per call -- a quarter of an S2 sampler step."""

    def __init__(self, N, p, n_iters):
        self.N, self.p, self.n = int(N), int(p), int(n_iters)
        n_in = self.N * self.p + self.N + self.p + 1           # X, kappa, beta, status slot
        n_out = self.n * self.p + self.N                         # betas, omega
        self.h = torch.zeros(n_in + n_out, dtype=F64).pin_memory()
        self.d = torch.zeros(n_in + n_out, dtype=F64, device=dev())
        self.n_in = n_in

    def run(self, X, kappa, prior_variance, beta, seed):
        N, p, n, n_in = self.N, self.p, self.n, self.n_in
        h = self.h.numpy()
        h[:N * p] = np.asarray(X, dtype=np.float64).reshape(-1)
        h[N * p:N * p + N] = kappa
        h[N * p + N:N * p + N + p] = beta
        h[n_in - 1] = 0.0
        d = self.d
        d[:n_in].copy_(self.h[:n_in], non_blocking=True)
        dX, dk = d[:N * p], d[N * p:N * p + N]
        beta_io = d[N * p + N:N * p + N + p]
        status = d[n_in - 1:n_in].view(I32)[:1]
        betas, om = d[n_in:n_in + n * p], d[n_in + n * p:]
        check(lib().mitre_gibbs_omega_beta(_stream(), _p(dX), _p(dk), N, p, float(prior_variance), n, 0,
                                           int(seed) & 0xFFFFFFFFFFFFFFFF, _p(beta_io), _p(om), _p(betas),
                                           _p(status)), "mitre_gibbs_omega_beta")
        self.h.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        flags = int(h[n_in - 1:n_in].view(np.int32)[0])
        if flags:
            raise np.linalg.LinAlgError("Matrix is not positive definite (status flags %d)" % flags)
        return (h[n_in + n * p:].copy(), h[N * p + N:N * p + N + p].copy(), h[n_in:n_in + n * p].reshape(n, p).copy())


def pg_draw(z, seed):
    z = _dev_f64(z)
    out = torch.empty_like(z)
    check(lib().mitre_pg_draw(_stream(), _p(z), z.shape[0], int(seed) & 0xFFFFFFFFFFFFFFFF, _p(out)),
          "mitre_pg_draw")
    return out

"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for the plumbing.

What shards, and what is exchanged:

  * candidate scoring -- contiguous [start, stop) slices of the lexsorted candidate index per rank,
    exactly the reference's own (unused) precedent, `_likelihood_worker` (logit_rules.py:21-60:
    `truth_table[start:stop]`, results returned as `[start, likelihoods]`).  Every rank gets the
    same O(N p) sweep inputs (KBs).  Two exchange modes:
        all-gather   every rank ends up with likelihoods[P] (what logit_rules.py:861-863 consumes)
        two-stage    ranks all-gather one (max, sum-exp) pair each (16 bytes), the rank owning the
                     marker draws locally, the index is broadcast -- no O(P) traffic at all
  * independent chains (replicates, main.py:1913-1986) -- pure replicas, no collective while
    sampling; seeds base_seed + i (main.py:1930-1938).

The functions take an optional process group and work on whatever device the tensors live on, so
the host logic is tested under `gloo` on CPU (tests/test_parallel_cpu.py); on the GPU box the
backend is NCCL over NVLink.  CUDA contexts must be created after spawn, never across fork.
"""
import math

import numpy as np
import torch
import torch.distributed as dist


def world(group=None):
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def shard_bounds(P, world_size, rank):
    """Contiguous slice [start, stop) of the candidate index for `rank`; sizes differ by <= 1."""
    base, extra = divmod(int(P), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_shard_bounds(P, world_size):
    return [shard_bounds(P, world_size, r) for r in range(world_size)]


def balanced_bounds(sorted_rows, world_size, tile=128, extra_round_cost=0.35):
    """Slice bounds [(start, stop)] * world_size that balance the WORK of the duplicate-aware scoring kernel
    instead of the row count.  The kernel scores the D distinct rows of a 128-row tile in ceil(D / 32)
    rounds, so a tile costs about 1 + extra_round_cost * (ceil(D / 32) - 1); the share of tiles needing a
    second round varies along a lexsorted table (few duplicates among the middle ranks), which left an
    equal-rows split 5 % out of balance at 8 GPUs.  extra_round_cost = 0.35 is fitted to the per-rank kernel
    times of an 8-GPU S5 run (with 0.5 the end slices, 91 % duplicates, ran 5 % longer than the middle ones).  Bounds are multiples of the tile; every rank computes
    the same bounds from the same (replicated) sorted table."""
    P = int(sorted_rows.shape[0])
    if world_size == 1 or P < tile * world_size * 4:
        return all_shard_bounds(P, world_size)
    heads = torch.ones(P, dtype=torch.bool, device=sorted_rows.device)
    heads[1:] = (sorted_rows[1:] != sorted_rows[:-1]).any(dim=1)
    ntiles = (P + tile - 1) // tile
    padded = torch.zeros(ntiles * tile, dtype=torch.int32, device=sorted_rows.device)
    padded[:P] = heads
    D_tile = padded.view(ntiles, tile).sum(dim=1)
    rounds = torch.clamp((D_tile + 31) // 32, min=1).to(torch.float64)
    cum = torch.cumsum(1.0 + extra_round_cost * (rounds - 1.0), dim=0)
    targets = cum[-1] * torch.arange(1, world_size, dtype=torch.float64, device=cum.device) / world_size
    cuts = (torch.searchsorted(cum, targets) + 1).clamp(max=ntiles).cpu().numpy().astype(np.int64) * tile
    edges = [0] + [int(min(c, P)) for c in cuts] + [P]
    for i in range(1, len(edges)):
        edges[i] = max(edges[i], edges[i - 1])
    return [(edges[r], edges[r + 1]) for r in range(world_size)]


def chains_for_rank(n_chains, world_size, rank, base_seed=0):
    """Chain indices and seeds (base_seed + i, main.py:1930-1938) this rank runs."""
    return [(i, base_seed + i) for i in range(n_chains) if i % world_size == rank]


def allgather_loglik(local, P, group=None, bounds=None):
    """local: float64[stop-start] on this rank -> float64[P] on every rank (shard order)."""
    ws, rank = world(group)
    if ws == 1:
        return local
    bounds = bounds if bounds is not None else all_shard_bounds(P, ws)
    width = max(b - a for a, b in bounds)
    padded = local.new_zeros(width)
    padded[:local.shape[0]] = local
    gathered = local.new_empty(ws * width)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    if all(b - a == width for a, b in bounds):
        return gathered
    return torch.cat([gathered[r * width:r * width + (b - a)] for r, (a, b) in enumerate(bounds)])


def gather_stats(local_stats, group=None):
    """local_stats: float64[2] = (max w, sum exp(w - max)) of this rank's log-weights.
    Returns float64[world, 2] on every rank (16 bytes per rank on the wire)."""
    ws, _ = world(group)
    if ws == 1:
        return local_stats.reshape(1, 2)
    out = local_stats.new_empty(ws * 2)
    dist.all_gather_into_tensor(out, local_stats.contiguous(), group=group)
    return out.reshape(ws, 2)


def owner_of_marker(stats, u, stabilise=True):
    """Host arithmetic of the two-stage draw.  stats: array[world, 2] of (max, sum exp(w - max));
    u: the uniform draw.  Returns (owner rank, u_local, log_total): the rank whose slice contains
    marker = u * total, the uniform to use inside that rank's slice, and log(sum exp(w)) over all
    candidates.  Mirrors choose_from_relative_probabilities (rules.py:1261-1264): the owner is the
    first rank whose cumulative mass is NOT < marker."""
    stats = np.asarray(stats, dtype=np.float64).reshape(-1, 2)
    m, s = stats[:, 0], stats[:, 1]
    finite = np.isfinite(m) & (s > 0)
    if not finite.any():
        return 0, float(u), -math.inf
    M = float(np.max(m[finite])) if stabilise else 0.0
    mass = np.where(finite, s * np.exp(np.where(finite, m, 0.0) - M), 0.0)
    cum = np.cumsum(mass)
    total = cum[-1]
    log_total = M + math.log(total) if total > 0 else -math.inf
    if not total > 0:
        return 0, float(u), log_total
    marker = u * total
    owner = int(np.sum(cum < marker))
    owner = min(owner, len(mass) - 1)
    while mass[owner] == 0 and owner + 1 < len(mass):
        owner += 1
    before = cum[owner] - mass[owner]
    u_local = (marker - before) / mass[owner] if mass[owner] > 0 else 0.0
    return owner, float(min(max(u_local, 0.0), np.nextafter(1.0, 0.0))), log_total


def two_stage_draw(local_weights_stats, local_draw, start, u, group=None, stabilise=True):
    """Draw one global candidate index from per-rank log-weights without moving them.

    local_weights_stats: float64[2] tensor (max, sum exp(w - max)) of this rank's slice
    local_draw(u_local) -> int: index within this rank's slice for the local uniform (on the GPU
        path: device.draw_index(..., stabilise=True) on the local log-weights)
    start: first global index of this rank's slice
    Returns (global_index, log_total) on every rank."""
    ws, rank = world(group)
    stats = gather_stats(local_weights_stats, group).cpu().numpy()
    owner, u_local, log_total = owner_of_marker(stats, u, stabilise)
    payload = torch.zeros(1, dtype=torch.int64, device=local_weights_stats.device)
    if rank == owner:
        payload[0] = int(start) + int(local_draw(u_local))
    if ws > 1:
        dist.broadcast(payload, src=dist.get_global_rank(group, owner) if group is not None else owner,
                       group=group)
    return int(payload.item()), log_total


class ShardedScorer:
    """Candidate-sharded `likelihoods_of_all_options` over the ranks of a process group.

    Every rank holds only its slice of the packed truth table (`local_rows`, int64[stop-start, W64]
    CUDA tensor).  `score` runs mitre_score_all on the slice; `gather` / `draw` are the two exchange
    modes."""

    def __init__(self, local_rows, P, N, group=None, bounds=None):
        from . import device as D
        self.D = D
        self.group = group
        self.P, self.N = int(P), int(N)
        self.ws, self.rank = world(group)
        self.bounds = bounds if bounds is not None else all_shard_bounds(self.P, self.ws)
        self.start, self.stop = self.bounds[self.rank]
        if local_rows.shape[0] != self.stop - self.start:
            raise ValueError("local slice has %d rows, expected %d" % (local_rows.shape[0],
                                                                        self.stop - self.start))
        self.local_rows = local_rows
        self.buffers = D.ScoreBuffers(max(self.stop - self.start, 1), self.N)
        self.draw_buffers = D.DrawBuffers(max(self.stop - self.start, 1))

    @classmethod
    def from_full_table(cls, packed_full, N, group=None):
        ws, rank = world(group)
        P = packed_full.shape[0]
        a, b = shard_bounds(P, ws, rank)
        return cls(packed_full[a:b].contiguous(), P, N, group)

    def score(self, X, omega, kappa, sigma2, mask=None):
        """Local slice log-likelihoods (device tensor) + status."""
        n = self.stop - self.start
        out = self.buffers.out[:n]
        return self.D.score_all(self.local_rows, self.N, X, omega, kappa, sigma2, mask=mask,
                                buffers=self.buffers, out=out)

    def gather(self, local_out):
        return allgather_loglik(local_out, self.P, self.group, self.bounds)

    def draw(self, local_out, local_prior, u, stabilise=True):
        """Two-stage draw of a global index from w = local_out (+ local_prior)."""
        D = self.D
        stats = D.weight_stats(local_out, local_prior, buffers=self.draw_buffers).clone()

        def local_draw(u_local):
            # draw_buffers.stats still holds this slice's (max, sum-exp) from weight_stats above
            return int(D.draw_index(local_out, local_prior, u_local, stabilise=True,
                                    buffers=self.draw_buffers, reuse_stats=True).item())
        return two_stage_draw(stats, local_draw, self.start, u, self.group, stabilise)


class ShardedSweep:
    """One problem's candidate sweep and proposal draw over the ranks of a process group: the sampler's
    detector update (logit_rules.py:851-863) with the candidate table cut into contiguous slices, one per
    GPU (_likelihood_worker's [start, stop) slices, logit_rules.py:21-60).

    Per step every rank runs, on its own stream and with no host involvement in between:
        H2D of the sweep inputs (X, omega, kappa, mask; KBs) and of (u, seq)
        mitre_score_all_stats on its slice   -> log-likelihoods, (max, sum-exp) of lik + prior, tile pairs
        mitre_sharded_chunk_sums
        the exchange + draw                   -> the global index and log-sum-exp on EVERY rank
    exchange='p2p'   mitre_sharded_draw: both exchanges inside one kernel over NVLink peer memory
                     (CUDA IPC); the whole step is captured once as a CUDA graph and replayed
    exchange='nccl'  all-gather of the pairs, mitre_sharded_select, all-reduce MAX of the index
    Nothing P-long ever leaves a GPU; 16 bytes per rank go out, 16 bytes come back.

    local_rows int64[n, 1] CUDA (this rank's slice of the lexsorted table), local_prior float64[n] or
    None, P the global candidate count, start the slice's first global index."""

    def __init__(self, local_rows, local_prior, P, N, start, group=None, exchange='p2p', max_p=None,
                 use_graph=True):
        import ctypes
        from . import device as D
        from ._lib import lib, check, MAX_P
        self.D, self.group = D, group
        self.ws, self.rank = world(group)
        self.P, self.N, self.start = int(P), int(N), int(start)
        self.rows, self.prior = local_rows, local_prior
        self.n = int(local_rows.shape[0])
        self.W64 = D.words_for(N)
        self.max_p = MAX_P if max_p is None else int(max_p)
        dev = local_rows.device
        self.buffers = D.ScoreBuffers(max(self.n, 1), self.N, self.max_p)
        self.draw = D.DrawBuffers(max(self.n, 1))
        n_small = self.N * self.max_p + 2 * self.N + self.W64
        self.h_small = torch.zeros(n_small + 1, dtype=torch.float64).pin_memory()   # X, omega, kappa, mask, u
        self.d_small = torch.zeros(n_small + 1, dtype=torch.float64, device=dev)
        self.d_u = self.d_small[n_small:]
        self.d_seq = torch.ones(1, dtype=torch.float64, device=dev)   # advanced by the draw kernel itself
        self.d_res = torch.zeros(4, dtype=torch.float64, device=dev)    # index (int64 bits), M, total, status
        self.h_res = torch.zeros(4, dtype=torch.float64).pin_memory()
        self._empty_stats = torch.tensor([-np.inf, 0.0], dtype=torch.float64, device=dev)
        self.gathered = torch.zeros(2 * self.ws, dtype=torch.float64, device=dev)
        self.exchange = exchange
        self.comm = None
        self._graphs = {}
        self.use_graph = use_graph
        if exchange == 'p2p':
            comm = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(64)
            check(lib().mitre_comm_create(self.ws, self.rank, ctypes.byref(comm), handle), "mitre_comm_create")
            handles = [None] * self.ws
            if self.ws > 1:
                dist.all_gather_object(handles, bytes(handle.raw), group=group)
            else:
                handles = [bytes(handle.raw)]
            check(lib().mitre_comm_connect(comm, ctypes.c_char_p(b"".join(handles))), "mitre_comm_connect")
            self.comm = comm
            if self.ws > 1:
                dist.barrier(group=group)
        elif exchange != 'nccl':
            raise ValueError("exchange must be 'p2p' or 'nccl'")

    def close(self):
        from ._lib import lib
        if self.comm is not None:
            torch.cuda.synchronize()
            lib().mitre_comm_destroy(self.comm)
            self.comm = None

    # -- host staging ---------------------------------------------------------------------------------
    def stage(self, X, omega, kappa, mask_bool, u):
        """Write one step's inputs into the pinned buffer (no CUDA call)."""
        N, W64, mp = self.N, self.W64, self.max_p
        X = np.asarray(X, dtype=np.float64)
        p = X.shape[1]
        if p > mp:
            raise ValueError('design matrix has %d columns; this sweep was sized for %d' % (p, mp))
        h = self.h_small.numpy()
        h[:N * p] = X.reshape(-1)
        o = N * mp
        h[o:o + N] = omega
        h[o + N:o + 2 * N] = kappa
        h[o + 2 * N:o + 2 * N + W64] = self.D.pack_mask(mask_bool).view(np.float64)
        h[o + 2 * N + W64] = u
        return p

    def _views(self, p):
        N, W64, o = self.N, self.W64, self.N * self.max_p
        d = self.d_small
        return (d[:N * p].view(N, p), d[o:o + N], d[o + N:o + 2 * N],
                d[o + 2 * N:o + 2 * N + W64].view(dtype=torch.int64))

    # -- device work of one step (stream-ordered, no synchronisation) ----------------------------------
    def _launch(self, p, sigma2, copies=True):
        from ._lib import lib, check
        D = self.D
        _p, _s = D._p, D._stream
        if copies:
            self.d_small.copy_(self.h_small, non_blocking=True)
        dX, dom, dka, dmask = self._views(p)
        out = self.buffers.out[:self.n]
        if self.n:
            D.score_all_stats(self.rows, self.N, dX, dom, dka, sigma2, mask=dmask, buffers=self.buffers,
                              out=out, log_prior=self.prior, draw_buffers=self.draw)
        else:
            self.draw.stats.copy_(self._empty_stats)
        wsb = self.draw.workspace
        check(lib().mitre_sharded_chunk_sums(_s(), self.n, _p(self.draw.stats), _p(wsb), wsb.numel()),
              "mitre_sharded_chunk_sums")
        idx_out = self.d_res[:1].view(torch.int64)
        if self.exchange == 'p2p':
            check(lib().mitre_sharded_draw(_s(), self.comm, _p(self.d_u), _p(self.d_seq), _p(self.draw.stats),
                                           _p(out), _p(self.prior), self.n, self.start, _p(wsb), wsb.numel(),
                                           _p(idx_out), _p(self.d_res[1:3]), _p(self.buffers.status)),
                  "mitre_sharded_draw")
        else:
            if self.ws > 1:
                dist.all_gather_into_tensor(self.gathered, self.draw.stats, group=self.group)
            else:
                self.gathered.copy_(self.draw.stats)
            check(lib().mitre_sharded_select(_s(), self.ws, self.rank, _p(self.d_u), _p(self.gathered), _p(out),
                                             _p(self.prior), self.n, self.start, _p(wsb), wsb.numel(), _p(idx_out),
                                             _p(self.d_res[1:3])), "mitre_sharded_select")
            if self.ws > 1:
                dist.all_reduce(idx_out, op=dist.ReduceOp.MAX, group=self.group)
        if copies:
            self.d_res[3:].copy_(self.buffers.status)      # status word rides along with the result
            self.h_res.copy_(self.d_res, non_blocking=True)
        return out

    def launch(self, p, sigma2, copies=True):
        """Enqueue one step.  copies=True: the step's inputs come from the pinned staging buffer (stage())
        and the result goes back to pinned memory; copies=False: inputs as they are in HBM, result left on
        the device.  p2p: captured once per (p, sigma2, copies) as a CUDA graph and replayed."""
        if self.exchange == 'p2p' and self.use_graph:
            key = (p, float(sigma2), bool(copies))
            g = self._graphs.get(key)
            if g is None:
                # one plain run (module loading, function attributes), then capture; every rank takes this
                # branch on the same call, so the extra step keeps the sequence numbers aligned
                self._launch(p, sigma2, copies)
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch(p, sigma2, copies)
                self._graphs[key] = g
            g.replay()
        else:
            self._launch(p, sigma2, copies)

    def result(self):
        """(global index, log-sum-exp over all candidates) of the last launched step (synchronises; one
        32-byte device->host copy carries index, pair and status)."""
        torch.cuda.current_stream().synchronize()
        r = self.h_res.numpy()
        status = int(r[3])
        if status:
            self.buffers.status.zero_()
            if status & 8:
                raise RuntimeError('sharded draw: a peer rank did not answer within the time limit')
            raise np.linalg.LinAlgError('Matrix is not positive definite (status flags %d)' % status)
        idx = int(r[:1].view(np.int64)[0])
        return idx, (float(r[1] + np.log(r[2])) if r[2] > 0 else -np.inf)

    def step(self, X, omega, kappa, mask_bool, u, sigma2):
        p = self.stage(X, omega, kappa, mask_bool, u)
        self.launch(p, sigma2)
        return self.result()


# ---- variable-sharded population build (SURVEY 8e, row "Detector evaluation") ------------------
#
# Every (window, variable, type) group depends only on its variable's series (rules.py:585-599),
# so detector evaluation shards by VARIABLE with no exchange.  What the ranks must agree on is the
# reference's global candidate order: enumeration is window-major over ALL variables
# (rules.py:465-511) and the lexsort is stable in that enumeration (rules.py:439-462).  The build
# below follows the survey's first option: ranks all-gather (packed row, global enumeration
# index) pairs -- 16 bytes per candidate, once per model build -- lay them out in enumeration
# order, run the same stable radix sort as the single-GPU path (replicated), and keep their own
# contiguous slice of the sorted order for scoring.  The host arithmetic (global group ids, global
# enumeration indices, descriptor lookup) is plain numpy and is tested under gloo on CPU.

def variable_bounds(V, world_size, rank):
    """Contiguous block [v0, v1) of variables for `rank`."""
    return shard_bounds(V, world_size, rank)


def global_group_ids(ok_slope, V_total, v0, v1):
    """Global group id of every local group of a rank holding variables [v0, v1).

    Local and global enumeration are both window-major, then variable, then type (slope before
    average, slope only where `ok_slope[w]`): local group (w, v_local, t) is global group
    GB[w] + (v0 + v_local) * nt_w + t with GB[w] = V_total * sum_{w' < w} nt_{w'}."""
    nt = 1 + np.asarray(ok_slope, dtype=np.int64).reshape(-1)
    GB = np.zeros(nt.shape[0] + 1, dtype=np.int64)
    np.cumsum(V_total * nt, out=GB[1:])
    n_local = int(v1 - v0)
    per_window = n_local * nt
    w_of = np.repeat(np.arange(nt.shape[0], dtype=np.int64), per_window)
    lb = np.zeros(nt.shape[0] + 1, dtype=np.int64)
    np.cumsum(per_window, out=lb[1:])
    within = np.arange(int(lb[-1]), dtype=np.int64) - lb[w_of]
    return GB[w_of] + v0 * nt[w_of] + within, GB


def describe_group(g_global, GB, ok_slope, V_total):
    """Inverse of global_group_ids for one group: (window, variable, type) with type 0 = slope,
    1 = average (the reference's ('slope', 'average') order, rules.py:585-599)."""
    w = int(np.searchsorted(GB, g_global, side='right') - 1)
    nt = 2 if ok_slope[w] else 1
    within = int(g_global - GB[w])
    return w, within // nt, (within % nt if nt == 2 else 1)


def global_enumeration_indices(K_local, g_global_local, row_offsets_global):
    """Global enumeration index of every local row, local rows in local enumeration order
    (group-major, then threshold, then direction): row_offsets_global[g] + slot."""
    counts = 2 * np.asarray(K_local, dtype=np.int64)
    first = np.repeat(np.asarray(row_offsets_global)[g_global_local], counts)
    local_off = np.zeros(counts.shape[0] + 1, dtype=np.int64)
    np.cumsum(counts, out=local_off[1:])
    slot = np.arange(int(local_off[-1]), dtype=np.int64) - np.repeat(local_off[:-1], counts)
    return first + slot


def _allgather_ragged(t, group=None):
    """all-gather of 1-D/2-D tensors whose leading sizes differ between ranks -> list per rank."""
    ws, _ = world(group)
    if ws == 1:
        return [t]
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = torch.empty(ws, dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(sizes, n, group=group)
    sizes = [int(x) for x in sizes.cpu()]
    width = max(max(sizes), 1)
    padded = t.new_zeros((width,) + tuple(t.shape[1:]))
    padded[:t.shape[0]] = t
    out = t.new_empty((ws * width,) + tuple(t.shape[1:]))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    return [out[r * width:r * width + sizes[r]] for r in range(ws)]


class ShardedPopulation:
    """The candidate table built by `world` ranks, each evaluating the detectors of its own
    variables; every rank ends with its contiguous slice of the reference's global sorted order.

    data: the FULL Dataset (same on every rank; only the rank's variables are staged on its GPU).
    After construction:
        P, start, stop      global candidate count and this rank's slice of the sorted index
        local_rows          int64[stop-start, W64] packed truth rows of the slice (CUDA)
        local_perm          int64[stop-start] global enumeration index of each of them
        scorer()            ShardedScorer over the slice
        describe(k)         (variable, window, type, direction, threshold) of global sorted index k
    """

    def __init__(self, data, tmin, N_intervals, tmax=None, group=None, balance=True):
        from . import device as D
        from . import rules as R
        self.group = group
        self.balance = balance
        self.ws, self.rank = world(group)
        self.N = data.n_subjects
        self.V = data.n_variables
        self.v0, self.v1 = variable_bounds(self.V, self.ws, self.rank)
        X_local = [np.ascontiguousarray(x[self.v0:self.v1]) for x in data.X]
        names = list(data.variable_names[self.v0:self.v1])
        weights = np.asarray(data.variable_weights)[self.v0:self.v1]
        local_data = R.Dataset(X_local, data.T, data.y, names, weights, data.experiment_start,
                               data.experiment_end)
        pop = R.DiscreteRulePopulation.__new__(R.DiscreteRulePopulation)
        pop.model, pop.data, pop.tmin = None, local_data, tmin
        pop.tmax = np.inf if tmax is None else tmax
        pop.N_intervals, pop.max_thresholds, pop.applied_filters = N_intervals, np.inf, []
        pop._enumerate_windows()
        pop.calculate_values()
        pop.update_base_rules()
        self.local_population = pop
        self.windows = list(pop.windows)
        self.ok_slope = np.array(pop.windows_ok_for_slope, dtype=bool)
        dev = D.dev()

        # -- global group ids and global row offsets (needs every rank's K) --
        g_local, self.GB = global_group_ids(self.ok_slope, self.V, self.v0, self.v1)
        self.G = int(self.GB[-1])
        K_local = pop._K if pop._G else torch.empty(0, dtype=torch.int32, device=dev)
        d_g_local = torch.as_tensor(g_local, device=dev)
        K_parts = _allgather_ragged(K_local, group)
        g_parts = _allgather_ragged(d_g_local, group)
        K_global = torch.zeros(self.G, dtype=torch.int32, device=dev)
        for Kp, gp in zip(K_parts, g_parts):
            K_global[gp] = Kp
        self.row_offsets = D.detector_row_offsets(K_global) if self.G else torch.zeros(1, dtype=torch.int64,
                                                                                       device=dev)
        self._row_offsets_host = self.row_offsets.cpu().numpy()
        self.P = int(self._row_offsets_host[-1])
        W64 = D.words_for(self.N)

        # -- local rows in local enumeration order + their global enumeration indices --
        if pop._P:
            if getattr(pop, '_rows_padded', None) is not None:
                L = pop._rows_padded.shape[1]
                live = torch.arange(L, device=dev)[None, :] < (2 * pop._K.to(torch.int64))[:, None]
                rows_local = pop._rows_padded[live].view(-1, 1)
            else:
                rows_local = pop._rows
            # global_enumeration_indices on the device (the numpy form took 2.0 of the 2.4 s of an S5 build):
            # row i of local group g has global index row_offsets[g_global] + (i - local_offset[g])
            counts = 2 * pop._K.to(torch.int64)
            shift = self.row_offsets[d_g_local] - (torch.cumsum(counts, 0) - counts)
            e_local = torch.repeat_interleave(shift, counts, output_size=int(pop._P))
            e_local += torch.arange(int(pop._P), dtype=torch.int64, device=dev)
            del counts, shift
        else:
            rows_local = torch.empty((0, W64), dtype=torch.int64, device=dev)
            e_local = torch.empty(0, dtype=torch.int64, device=dev)

        # -- the one exchange of the build: 8*W64 + 8 bytes per candidate --
        rows_enum = torch.empty((self.P, W64), dtype=torch.int64, device=dev)
        for rp, ep in zip(_allgather_ragged(rows_local.contiguous(), group), _allgather_ragged(e_local, group)):
            rows_enum[ep] = rp
        del rows_local, e_local
        if self.P:
            perm, sorted_rows = D.lexsort_rows(rows_enum, self.N)
        else:
            perm, sorted_rows = torch.empty(0, dtype=torch.int64, device=dev), rows_enum
        del rows_enum
        # slices of the sorted order: equal work for the scoring kernel (balanced_bounds), or equal rows
        self.bounds = (balanced_bounds(sorted_rows, self.ws) if self.balance and self.P
                       else all_shard_bounds(self.P, self.ws))
        self.start, self.stop = self.bounds[self.rank]
        self.local_rows = sorted_rows[self.start:self.stop].clone()
        self.local_perm = perm[self.start:self.stop].clone()
        del perm, sorted_rows
        pop._rows = None
        pop._rows_padded = None

    def scorer(self):
        return ShardedScorer(self.local_rows, self.P, self.N, self.group, bounds=self.bounds)

    def owner_of_index(self, k):
        for r, (a, b) in enumerate(self.bounds):
            if a <= k < b:
                return r
        raise IndexError(k)

    def owner_of_variable(self, v):
        for r in range(self.ws):
            a, b = variable_bounds(self.V, self.ws, r)
            if a <= v < b:
                return r
        raise IndexError(v)

    def _bcast(self, value, src, dtype):
        t = torch.zeros(1, dtype=dtype, device=self.local_rows.device)
        if self.rank == src:
            t[0] = value
        if self.ws > 1:
            dist.broadcast(t, src=dist.get_global_rank(self.group, src) if self.group is not None else src,
                           group=self.group)
        return t.item()

    def describe(self, k):
        """Global sorted index -> the reference's primitive tuple
        (variable, (t0, t1), type, direction, threshold) (rules.py:23-44), identical on every rank.
        Two 8-byte broadcasts: the enumeration index from the rank that owns slice position k, the
        threshold from the rank that owns the variable."""
        k = int(k)
        owner = self.owner_of_index(k)
        e = int(self._bcast(int(self.local_perm[k - self.start]) if self.rank == owner else 0, owner, torch.int64))
        g = int(np.searchsorted(self._row_offsets_host, e, side='right') - 1)
        slot = e - int(self._row_offsets_host[g])
        w, v, t = describe_group(g, self.GB, self.ok_slope, self.V)
        vowner = self.owner_of_variable(v)
        thr = 0.0
        if self.rank == vowner:
            pop = self.local_population
            g_loc = int(pop._group_base[w]) + (v - self.v0) * (2 if self.ok_slope[w] else 1) + (t if self.ok_slope[w] else 0)
            thr = float(pop._thresholds[g_loc, slot >> 1])
        thr = float(self._bcast(thr, vowner, torch.float64))
        return (v, tuple(self.windows[w]), ('slope', 'average')[t], ('above', 'below')[slot & 1], thr)

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


def chains_for_rank(n_chains, world_size, rank, base_seed=0):
    """Chain indices and seeds (base_seed + i, main.py:1930-1938) this rank runs."""
    return [(i, base_seed + i) for i in range(n_chains) if i % world_size == rank]


def allgather_loglik(local, P, group=None):
    """local: float64[stop-start] on this rank -> float64[P] on every rank (shard order)."""
    ws, rank = world(group)
    if ws == 1:
        return local
    bounds = all_shard_bounds(P, ws)
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

    def __init__(self, local_rows, P, N, group=None):
        from . import device as D
        self.D = D
        self.group = group
        self.P, self.N = int(P), int(N)
        self.ws, self.rank = world(group)
        self.start, self.stop = shard_bounds(self.P, self.ws, self.rank)
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
        return allgather_loglik(local_out, self.P, self.group)

    def draw(self, local_out, local_prior, u, stabilise=True):
        """Two-stage draw of a global index from w = local_out (+ local_prior)."""
        D = self.D
        stats = D.weight_stats(local_out, local_prior, buffers=self.draw_buffers).clone()

        def local_draw(u_local):
            # draw_buffers.stats still holds this slice's (max, sum-exp) from weight_stats above
            return int(D.draw_index(local_out, local_prior, u_local, stabilise=True,
                                    buffers=self.draw_buffers, reuse_stats=True).item())
        return two_stage_draw(stats, local_draw, self.start, u, self.group, stabilise)

"""The packed one-sweep lexsort (csrc/sort_packed.cu) against np.lexsort (what the reference runs,
rules.py:439-462) and against the general LSD path of csrc/sort.cu: same permutation -- ties in enumeration
order -- and same sorted rows, on tables with long runs of equal rows, for tile-boundary sizes, for every
digit plan (N = 1 .. 63) and straight from a padded table with sentinel slots."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def pack(truth):
    P, N = truth.shape
    out = np.zeros(P, dtype=np.uint64)
    for i in range(N):
        out |= truth[:, i].astype(np.uint64) << np.uint64(63 - i)
    return out


def random_truth(rng, P, N, distinct):
    pool = rng.random((distinct, N)) < rng.random((1, N))
    return pool[rng.integers(0, distinct, size=P)]


@pytest.mark.parametrize("N,P,distinct", [
    (1, 5000, 2), (3, 1, 8), (9, 4096, 40), (10, 4097, 500), (18, 100000, 3000), (27, 12289, 12289),
    (28, 70000, 50), (35, 300000, 20000), (35, 3000000, 100000), (36, 8191, 9), (45, 50000, 700),
    (48, 60000, 60000), (63, 2, 2)])
def test_packed_sort_matches_np_lexsort(N, P, distinct):
    import torch
    from mitre_b200 import device as D
    from mitre_b200._lib import lib
    rng = np.random.default_rng(N * 1000003 + P)
    truth = random_truth(rng, P, N, distinct)
    want = np.lexsort(np.rot90(truth))
    rows = torch.as_tensor(pack(truth).view(np.int64), device=D.dev()).view(-1, 1)
    perm, srt = D.lexsort_rows(rows, N)
    assert np.array_equal(perm.cpu().numpy(), want)
    assert np.array_equal(srt.cpu().numpy().ravel(), rows.cpu().numpy().ravel()[want])
    old = lib().mitre_set_packed_sort(0)
    try:
        perm0, srt0 = D.lexsort_rows(rows, N)
    finally:
        lib().mitre_set_packed_sort(old)
    assert torch.equal(perm, perm0) and torch.equal(srt, srt0)


@pytest.mark.parametrize("N,G,seed", [(4, 3, 0), (12, 700, 1), (35, 40000, 2), (35, 61, 3), (20, 9000, 4)])
def test_padded_sort_drops_sentinels_and_densifies(N, G, seed):
    """A padded table as mitre_detectors_fused leaves it: 2(N-1) slots per group, K_g valid (above, below)
    pairs first, ~0 in the rest.  Result: the dense table's np.lexsort."""
    import torch
    from mitre_b200 import device as D
    from mitre_b200._lib import lib
    rng = np.random.default_rng(seed)
    L = 2 * (N - 1)
    K = rng.integers(0, N, size=G)
    K[rng.random(G) < 0.1] = 0
    padded = np.full((G, L), ~np.uint64(0), dtype=np.uint64)
    dense = []
    valid_mask = (np.uint64(0xffffffffffffffff) << np.uint64(64 - N))
    for g in range(G):
        order = rng.permutation(N)
        above = np.uint64(0)
        rows = []
        for k in range(K[g]):
            above |= np.uint64(1) << np.uint64(63 - order[k % N])
            if rng.random() < 0.3 and rows:
                above = rows[-2]                      # repeated rows across thresholds
            rows += [above, ~above & valid_mask]
        padded[g, :len(rows)] = rows
        dense += rows
    dense = np.array(dense, dtype=np.uint64)
    row_offsets = np.concatenate([[0], np.cumsum(2 * K)]).astype(np.int64)
    P = int(row_offsets[-1])
    bits = ((dense[:, None] >> (np.uint64(63) - np.arange(N, dtype=np.uint64))[None, :]) & np.uint64(1)).astype(bool)
    want = np.lexsort(np.rot90(bits)) if P else np.zeros(0, dtype=np.int64)
    d_pad = torch.as_tensor(padded.view(np.int64), device=D.dev())
    d_off = torch.as_tensor(row_offsets, device=D.dev())
    perm, srt = D.lexsort_padded(d_pad, N, P, d_off)
    assert np.array_equal(perm.cpu().numpy(), want)
    assert np.array_equal(srt.cpu().numpy().ravel().view(np.uint64), dense[want])
    old = lib().mitre_set_packed_sort(0)
    try:
        perm0, srt0 = D.lexsort_padded(d_pad, N, P, d_off)
    finally:
        lib().mitre_set_packed_sort(old)
    assert torch.equal(perm, perm0) and torch.equal(srt, srt0)

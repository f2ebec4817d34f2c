"""Times the scoring / draw launch variants on the S5 bench population (GPU box only).
usage: python tools/score_variants.py [V]

Prints, for the lexsorted S5 table: the fraction of rows equal to their predecessor, the sweep time
with and without duplicate-aware scoring, with the fused (max, sum-exp) + per-tile pairs, the draw
from the pairs against the draw with its own pass, and the same on a table whose rows were made
all distinct (worst case for the duplicate-aware kernel)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from mitre_b200 import device as D  # noqa: E402
from mitre_b200._lib import lib  # noqa: E402


def t(fn, iters=10, reps=3):
    fn()
    return min(bench.cuda_time_ms(torch, fn, iters) for _ in range(reps))


def main():
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    sys.argv = sys.argv[:1]
    args = bench.parse_args()
    args.variables = V
    data, pop, marks = bench.build_population(args, 0, torch)
    N, P = data.n_subjects, len(pop)
    packed = pop.packed_truth
    dup = float((packed[1:, 0] == packed[:-1, 0]).double().mean())
    print("P = %d rows, %.1f %% equal their predecessor" % (P, 100 * dup))
    state = bench.sweep_state(N, args.rules, lambda rng: pop._truth_row(int(rng.randint(P))), 1)
    dev = D.dev()
    dX = torch.as_tensor(state["X"], device=dev)
    dom = torch.as_tensor(state["omega"], device=dev)
    dka = torch.as_tensor(state["kappa"], device=dev)
    dmask = torch.as_tensor(D.pack_mask(state["mask"]), device=dev)
    buffers = D.ScoreBuffers(P, N, args.rules + 1)
    draw = D.DrawBuffers(P)
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    prior = torch.randn(P, dtype=torch.float64, device=dev, generator=g) - np.log(P)

    def plain(rows=packed):
        return D.score_all(rows, N, dX, dom, dka, 100.0, mask=dmask, buffers=buffers)[0]

    def stats(rows=packed, pr=prior):
        return D.score_all_stats(rows, N, dX, dom, dka, 100.0, mask=dmask, buffers=buffers, log_prior=pr,
                                 draw_buffers=draw)

    for _ in range(30):
        plain()
    res = {}
    for dd in (0, 1, 2):
        lib().mitre_set_dedupe(dd)
        res[dd] = plain().clone()
        print("dedupe=%d  score_all        %.3f ms" % (dd, t(plain)))
        print("dedupe=%d  score_all_stats  %.3f ms" % (dd, t(stats)))
        print("dedupe=%d  stats, no prior  %.3f ms" % (dd, t(lambda: stats(pr=None))))
    print("max rel diff dedupe vs plain kernels: %.2e %.2e" % (float(((res[0] - res[1]) / res[0]).abs().max()),
                                                              float(((res[0] - res[2]) / res[0]).abs().max())))
    lib().mitre_set_dedupe(1)
    for bits in (11, 12, 13, 14, 16):
        lib().mitre_set_dedupe_table_bits(bits)
        print("dedupe=1 table bits %2d  score_all %.3f ms   score_all_stats %.3f ms" % (bits, t(plain), t(stats)))
    lib().mitre_set_dedupe_table_bits(11)
    out, _, st = stats()
    k_tiles = int(D.draw_index(out, prior, 0.37, stabilise=True, buffers=draw, reuse_tiles=True).item())
    k_pass = int(D.draw_index(out, prior, 0.37, stabilise=True, buffers=draw, reuse_stats=True).item())
    w = (out + prior).cpu().numpy()
    c = np.cumsum(np.exp(w - w.max()))
    print("draw: tiles %d  own pass %d  numpy cumsum %d" % (k_tiles, k_pass, int(np.sum(c < 0.37 * c[-1]))))
    print("draw from tiles   %.3f ms" % t(lambda: D.draw_index(out, prior, 0.37, stabilise=True, buffers=draw,
                                                               reuse_tiles=True)))
    print("draw own pass     %.3f ms" % t(lambda: D.draw_index(out, prior, 0.37, stabilise=True, buffers=draw,
                                                               reuse_stats=True)))
    print("sweep + draw (tiles) %.3f ms" % t(lambda: (stats(), D.draw_index(out, prior, 0.37, stabilise=True,
                                                                           buffers=draw, reuse_tiles=True))))
    # worst case: random rows (~99 % distinct), sorted as unsigned keys like the real table
    rnd = torch.randint(0, 2 ** N, (P,), dtype=torch.int64, device=dev) << (64 - N)
    top = torch.tensor(-2 ** 63, dtype=torch.int64, device=dev)
    uniq = (torch.sort(rnd ^ top).values ^ top).view(-1, 1).contiguous()
    del rnd
    print("random table: %.1f %% duplicates" % (100 * float((uniq[1:, 0] == uniq[:-1, 0]).double().mean())))
    for dd in (0, 1, 2):
        lib().mitre_set_dedupe(dd)
        print("dedupe=%d  score_all on distinct rows  %.3f ms" % (dd, t(lambda: plain(uniq))))
    lib().mitre_set_dedupe(1)


if __name__ == "__main__":
    main()

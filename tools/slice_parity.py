"""Bitwise parity of a slice sweep against the same rows of the whole-table sweep (one GPU)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench  # noqa: E402
from mitre_b200 import device as D, parallel as PL, synthetic  # noqa: E402
from mitre_b200._lib import lib  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
sys.argv = sys.argv[:1]
args = bench.parse_args()
args.variables = V
data, pop, _ = bench.build_population(args, 0, torch)
N, P = data.n_subjects, len(pop)
rows = pop.packed_truth
state = synthetic.make_sweep_state(N, args.rules, np.stack([pop._truth_row(i) for i in (5, 77, 1234, 4000, 99999 % P)]), seed=1)
X, omega, kappa, mask = state["X"], state["omega"], state["y"] - 0.5, state["mask"]
dmask = torch.as_tensor(D.pack_mask(mask), device=D.dev())
prior = torch.randn(P, dtype=torch.float64, device=D.dev())
p = X.shape[1]
buffers, draw = D.ScoreBuffers(P, N, p), D.DrawBuffers(P)
out, status, st = D.score_all_stats(rows, N, X, omega, kappa, 100.0, mask=dmask, buffers=buffers, log_prior=prior,
                                    draw_buffers=draw)
out = out.clone()
plain, status = D.score_all(rows, N, X, omega, kappa, 100.0, mask=dmask)
print("stats vs plain sweep identical:", torch.equal(out, plain))
bounds = PL.balanced_bounds(rows, 2)
print("P", P, "bounds", bounds)
for (a, b) in list(bounds) + [(0, P // 2), (128 * 1000, 128 * 1000 + 128 * 4096), (1000, 1000 + 500000), (3, P)]:
    sl = rows[a:b].clone()
    pr = prior[a:b].clone()
    bufs, dr = D.ScoreBuffers(b - a, N, p), D.DrawBuffers(b - a)
    o2, status2, st2 = D.score_all_stats(sl, N, X, omega, kappa, 100.0, mask=dmask, buffers=bufs, log_prior=pr,
                                         draw_buffers=dr)
    same = torch.equal(o2, out[a:b])
    nd = int((o2 != out[a:b]).sum())
    md = float(((o2 - out[a:b]).abs() / out[a:b].abs()).max())
    print("slice [%d, %d): identical=%s  differing=%d  max rel diff=%.3g" % (a, b, same, nd, md))
    if nd:
        i = int(torch.nonzero(o2 != out[a:b])[0])
        print("   first difference at slice position %d (global %d, %d mod 128): row %x" % (i, a + i, (a + i) % 128, int(rows[a + i, 0]) & (2**64 - 1)))
        # is the row's value consistent within each sweep?
        r = rows[a + i, 0]
        same_rows = torch.nonzero(rows[:, 0] == r).ravel()
        print("   whole-table values of that row: %d distinct; slice values: %d distinct" % (
            len(torch.unique(out[same_rows])), len(torch.unique(o2[(same_rows[(same_rows >= a) & (same_rows < b)] - a)]))))

"""examples/score_all_host.c: a plain-C program linked against libmitre_b200.so -- no Python, no torch in
the process -- calls the drop-in entry point with host buffers; its output must match the C oracle on
the same (regenerated) inputs."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Lcg:
    def __init__(self, seed):
        self.s = seed & (2 ** 64 - 1)

    def uniform(self):
        self.s = (self.s * 6364136223846793005 + 1442695040888963407) & (2 ** 64 - 1)
        return (self.s >> 11) * (1.0 / 9007199254740992.0)


def rebuild_inputs(N, p, P, seed):
    g = Lcg(seed)
    X = np.zeros((N, p))
    omega = np.zeros(N)
    response = np.zeros(N)
    for i in range(N):
        for j in range(p):
            X[i, j] = 1.0 if j == p - 1 else (1.0 if g.uniform() > 0.5 else 0.0)
        omega[i] = 0.05 + 0.5 * g.uniform()
        response[i] = ((1.0 if g.uniform() > 0.5 else 0.0) - 0.5) / omega[i]
    truth = np.zeros((P, N))
    for k in range(P):
        density = g.uniform()
        for i in range(N):
            truth[k, i] = 1.0 if g.uniform() < density else 0.0
    return X, omega, response, truth


@pytest.mark.parametrize("N,p,P,seed", [(35, 3, 500, 1), (70, 2, 200, 7)])
def test_plain_c_caller_matches_oracle(tmp_path, N, p, P, seed):
    lib_path = os.path.join(ROOT, "mitre_b200", "libmitre_b200.so")     # built in-tree, travels with the repo
    assert os.path.exists(lib_path), "run python -m mitre_b200.build first"
    exe = str(tmp_path / "score_all_host")
    libdir = os.path.dirname(lib_path)
    subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "score_all_host.c"), "-L" + libdir, "-lmitre_b200",
                           "-Wl,-rpath," + libdir, "-lm", "-o", exe])
    res = subprocess.run([exe, str(N), str(p), str(P), str(seed)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    lines = res.stdout.split()
    assert [int(lines[0]), int(lines[1]), int(lines[2])] == [N, p, P]
    got = np.array([float(x) for x in lines[4:]])
    X, omega, response, truth = rebuild_inputs(N, p, P, seed)
    want = oracle.likelihoods_of_all_options(X, omega, response, truth, float(lines[3]))
    assert got.shape == want.shape
    assert np.max(np.abs(got - want) / np.abs(want)) <= 1e-9
    assert "torch" not in res.stderr.lower()

"""Build libmitre_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m mitre_b200.build [--force] [-v]

The library is plain CUDA runtime + the C ABI in include/mitre_b200.h; it does not link torch.
Translation units compile in parallel; fused_inst.cu is compiled once per sorting-network size.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libmitre_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
# the detector kernels must not contract a*b+c (bit-exact numpy means)
NO_FMAD = ["-fmad=false"]
FUSED_NB = tuple(range(4, 65, 4))


def units():
    """(source, object, extra flags) for every translation unit."""
    out = []
    for src in sorted(f for f in os.listdir(CSRC) if f.endswith(".cu")):
        if src == "fused_inst.cu":
            for nb in FUSED_NB:
                out.append((src, "fused_inst_nb%d.o" % nb, NO_FMAD + ["-DMITRE_FUSED_NB=%d" % nb]))
        elif src in ("detectors.cu", "fused_detectors.cu"):
            out.append((src, src[:-3] + ".o", NO_FMAD))
        else:
            out.append((src, src[:-3] + ".o", []))
    return out


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    headers.append(os.path.join(HERE, "..", "include", "mitre_b200.h"))
    jobs, objs = [], []
    for src, obj, extra in units():
        s, o = os.path.join(CSRC, src), os.path.join(OBJ, obj)
        objs.append(o)
        if force or _newer(o, [s] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + ["-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return cmd, r.returncode, r.stdout

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
            for cmd, rc, out in pool.map(run, jobs):
                if verbose or rc:
                    print(" ".join(cmd))
                    print(out)
                if rc:
                    raise subprocess.CalledProcessError(rc, cmd)
    stale = set(os.listdir(OBJ)) - {os.path.basename(o) for o in objs}
    for f in stale:
        os.remove(os.path.join(OBJ, f))
    if force or jobs or _newer(LIB, objs):
        # link beside the target and rename: a snapshot of the tree never sees a half-written library
        tmp = LIB + ".tmp%d" % os.getpid()
        subprocess.check_call([nvcc] + ARCH + ["-shared", "-o", tmp] + objs + ["-lcudart"])
        os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

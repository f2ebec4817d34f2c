#!/usr/bin/env python
"""Build the UNMODIFIED reference scoring kernel into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The reference's one native component is mitre/efficient_likelihoods.pyx (setup.py:8-13 builds
it as the top-level extension module `efficient_likelihoods`).  The shipped .c is Cython-0.24
output that does not compile against CPython 3.12, so we regenerate C from the .pyx where it
lies under /root/reference (language_level=2 because the file uses `xrange`,
efficient_likelihoods.pyx:66) and compile it with gcc.  Nothing from the reference is copied
into the repository: the generated C and the .so land in oracle/_ref/, which is git-ignored
(but travels to the GPU box with the gpurun snapshot).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
load the result.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PYX = os.environ.get("MITRE_REFERENCE", "/root/reference") + "/mitre/efficient_likelihoods.pyx"
OUT = os.path.join(HERE, "_ref")


def so_path():
    return os.path.join(OUT, "efficient_likelihoods" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force=False):
    """Returns the path of the built extension, or None when the reference tree is absent
    and no prebuilt copy exists."""
    target = so_path()
    if os.path.exists(target) and not force:
        return target
    if not os.path.exists(REF_PYX):
        return None
    import numpy
    os.makedirs(OUT, exist_ok=True)
    c_file = os.path.join(OUT, "efficient_likelihoods.c")
    subprocess.check_call([sys.executable, "-m", "cython", "-2", REF_PYX, "-o", c_file])
    cmd = ["gcc", "-shared", "-fPIC", "-O2", "-fno-strict-aliasing", "-w",
           "-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(),
           c_file, "-o", target]
    subprocess.check_call(cmd)
    return target


def load():
    """Import the built reference module (name `efficient_likelihoods`, as in
    logit_rules.py:17) without putting oracle/_ref on sys.path permanently."""
    import importlib.util
    target = build()
    if target is None:
        raise ImportError("reference extension not built and /root/reference absent")
    spec = importlib.util.spec_from_file_location("efficient_likelihoods", target)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p if p else "reference tree absent; nothing built")

"""Load the reference's Python-2 `mitre/rules.py` under Python 3 WITHOUT modifying its logic
(TEST INFRASTRUCTURE ONLY; works only where /root/reference exists, i.e. in the build
container -- the GPU box uses the golden fixtures this produces, tests/golden/).

The file parses under py3; it needs only py2 builtins and two renamed imports:
  * `xrange`, `reduce` injected                        (rules.py:477, 1144)
  * `zip` must return a LIST: `subject_info = zip(...)` is iterated once per
    (window, variable, type) inside nested loops        (rules.py:584-596)
  * `.iteritems()` -> `.items()`                        (rules.py:660)
  * `from scipy.misc import logsumexp` -> scipy.special (rules.py:10)
"""
import builtins
import functools
import os
import types

REFERENCE = os.environ.get("MITRE_REFERENCE", "/root/reference")


def available():
    return os.path.exists(os.path.join(REFERENCE, "mitre", "rules.py"))


def load_rules():
    path = os.path.join(REFERENCE, "mitre", "rules.py")
    with open(path) as f:
        src = f.read()
    src = src.replace("from scipy.misc import logsumexp", "from scipy.special import logsumexp")
    src = src.replace(".iteritems()", ".items()")
    mod = types.ModuleType("mitre_reference_rules")
    mod.__file__ = path
    mod.xrange = range
    mod.reduce = functools.reduce
    mod.zip = lambda *a: list(builtins.zip(*a))
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def load_reference_mc_logpdf_np():
    """mc_logpdf_np's own source text (mvnpdf_cholesky.py:53-74), executed as-is; the rest of
    that file holds py2 print statements and cannot be imported."""
    import re
    import numpy as np
    import scipy.linalg
    path = os.path.join(REFERENCE, "mitre", "mvnpdf_cholesky.py")
    src = open(path).read()
    m = re.search(r"^def mc_logpdf_np\(.*?(?=^def )", src, re.S | re.M)
    ns = {"np": np, "scipy": scipy}
    exec(compile(m.group(0), path, "exec"), ns)
    return ns["mc_logpdf_np"]


def load_logit_rules(pg_module, efficient_likelihoods_module):
    """The reference's mitre/logit_rules.py under py3, unmodified in logic, with its two
    unavailable imports supplied by the caller:
      * `pypolyagamma` (third party, not vendored, not pinned: logit_rules.py:13, 62) ->
        `pg_module`, which must offer PyPolyaGamma(seed).pgdrawv(n, z, out);
      * `efficient_likelihoods` (logit_rules.py:17) -> the extension built by build_ref.py.
    py2 -> py3 shims: implicit relative imports (6-7, 12), `from Queue import Empty` (15),
    `np.array(counts.values())` (74), builtin reduce (204), xrange, .iteritems().
    Returns (rules_module, logit_rules_module)."""
    import queue
    import sys
    rules_mod = load_rules()
    mvn = types.ModuleType("mvnpdf_cholesky")
    mvn.mc_logpdf_np = load_reference_mc_logpdf_np()
    saved = {k: sys.modules.get(k) for k in
             ("rules", "mvnpdf_cholesky", "pypolyagamma", "Queue", "efficient_likelihoods")}
    sys.modules["rules"] = rules_mod
    sys.modules["mvnpdf_cholesky"] = mvn
    sys.modules["pypolyagamma"] = pg_module
    sys.modules["Queue"] = queue
    sys.modules["efficient_likelihoods"] = efficient_likelihoods_module
    try:
        path = os.path.join(REFERENCE, "mitre", "logit_rules.py")
        src = open(path).read()
        src = src.replace("from scipy.misc import logsumexp", "from scipy.special import logsumexp")
        src = src.replace(".iteritems()", ".items()")
        src = src.replace("np.array(counts.values())", "np.array(list(counts.values()))")
        mod = types.ModuleType("mitre_reference_logit_rules")
        mod.__file__ = path
        mod.xrange = range
        mod.reduce = functools.reduce
        exec(compile(src, path, "exec"), mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return rules_mod, mod

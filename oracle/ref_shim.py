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

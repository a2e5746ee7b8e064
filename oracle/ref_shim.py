"""Import the real (Python-2) reference in THIS container -- test infrastructure only.

The reference at /root/reference is Python-2-only (print statements, xrange, implicit relative
imports, unconditional matplotlib imports).  This module makes a *temporary*,
mechanically rewritten copy under a scratch directory OUTSIDE the repository (default
/tmp/bnmtf_ref_shim/BNMTF) and imports it, so that

  * oracle/*.py (our numpy restatement) can be validated against the reference's own code, and
  * tests/golden/*.npz can be generated from the reference itself (tests/golden/make_golden.py).

Nothing here is shipped and nothing is copied into the repository's history.  One more copy is made by
__graft_entry__.build() under baseline/_ref (git-ignored): it travels to the GPU box with the snapshot and is what
`bench.py --impl reference` and bench.py's cpu_baseline / parity legs time and compare against there
(/root/reference itself does not exist on the GPU box).  Only tests/, the golden generators, bench.py's CPU legs and
__graft_entry__.build() may import this file.
"""
import os
import re
import shutil
import sys
import types

REFERENCE_ROOT = "/root/reference"
DEFAULT_SCRATCH = "/tmp/bnmtf_ref_shim"

_REWRITES = [
    (re.compile(r"^(\s*)print (?!\()(.*)$", re.M), r"\1print(\2)"),
    (re.compile(r"^(\s*if .*?:)\s*print (?!\()(.*)$", re.M), r"\1 print(\2)"),
    (re.compile(r"\bxrange\("), "range("),
    (re.compile(r"\.iteritems\(\)"), ".items()"),
    (re.compile(r"itertools\.izip"), "zip"),
    (re.compile(r"^from distributions\.", re.M), "from BNMTF.code.models.distributions."),
    (re.compile(r"^from kmeans\.kmeans", re.M), "from BNMTF.code.models.kmeans.kmeans"),
    (re.compile(r"^import rtnorm\s*$", re.M), "from BNMTF.code.models.distributions import rtnorm"),
    (re.compile(r"^import mask\s*$", re.M), "from BNMTF.code.cross_validation import mask"),
    (re.compile(r"^from (line_search_bnmf|grid_search_bnmtf|greedy_search_bnmtf|matrix_cross_validation|"
                r"parallel_matrix_cross_validation|mask) import", re.M),
     r"from BNMTF.code.cross_validation.\1 import"),
    (re.compile(r"index / row_length"), "index // row_length"),
]


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "code", "models"))


def build(scratch=DEFAULT_SCRATCH, subdirs=("code", "data_toy", "data_drug_sensitivity", "tests")):
    """Create the rewritten scratch copy (library code + data only) and return its parent dir."""
    dst = os.path.join(scratch, "BNMTF")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    for sub in subdirs:
        src = os.path.join(REFERENCE_ROOT, sub)
        if os.path.isdir(src):
            shutil.copytree(src, os.path.join(dst, sub))
    for name in ("__init__.py",):
        if os.path.exists(os.path.join(REFERENCE_ROOT, name)):
            shutil.copy(os.path.join(REFERENCE_ROOT, name), os.path.join(dst, name))
    for root, _dirs, files in os.walk(dst):
        if "__init__.py" not in files and any(f.endswith(".py") for f in files):
            open(os.path.join(root, "__init__.py"), "w").close()
        for f in files:
            if not f.endswith(".py"):
                continue
            path = os.path.join(root, f)
            with open(path, "r", encoding="utf-8", errors="replace") as fh:
                text = fh.read()
            for pat, rep in _REWRITES:
                text = pat.sub(rep, text)
            with open(path, "w", encoding="utf-8") as fh:
                fh.write(text)
    # matplotlib is not installed: a do-nothing stub that satisfies "import matplotlib.pyplot as plt"
    mpl = os.path.join(scratch, "matplotlib")
    os.makedirs(mpl, exist_ok=True)
    with open(os.path.join(mpl, "__init__.py"), "w") as fh:
        fh.write("def __getattr__(name):\n    return lambda *a, **k: None\n")
    with open(os.path.join(mpl, "pyplot.py"), "w") as fh:
        fh.write("def __getattr__(name):\n    return lambda *a, **k: None\n")
    return scratch


def load(scratch=DEFAULT_SCRATCH, rebuild=False):
    """Return a namespace with the reference's model classes and distribution functions."""
    have = os.path.isdir(os.path.join(scratch, "BNMTF", "code"))
    if rebuild or not have:
        # a rewritten copy made earlier (e.g. baseline/_ref, which travels to the GPU box) is used as it is
        if not available():
            raise RuntimeError("reference tree not present at %s and no rewritten copy under %s" % (REFERENCE_ROOT, scratch))
        build(scratch)
    if scratch not in sys.path:
        sys.path.insert(0, scratch)
    import importlib
    ns = types.SimpleNamespace()
    m = "BNMTF.code.models."
    ns.bnmf_gibbs_optimised = importlib.import_module(m + "bnmf_gibbs_optimised").bnmf_gibbs_optimised
    ns.bnmf_vb_optimised = importlib.import_module(m + "bnmf_vb_optimised").bnmf_vb_optimised
    ns.bnmtf_gibbs_optimised = importlib.import_module(m + "bnmtf_gibbs_optimised").bnmtf_gibbs_optimised
    ns.bnmtf_vb_optimised = importlib.import_module(m + "bnmtf_vb_optimised").bnmtf_vb_optimised
    ns.nmf_icm = importlib.import_module(m + "nmf_icm").nmf_icm
    ns.NMF = importlib.import_module(m + "nmf_np").NMF
    ns.nmtf_icm = importlib.import_module(m + "nmtf_icm").nmtf_icm
    ns.NMTF = importlib.import_module(m + "nmtf_np").NMTF
    ns.tn = importlib.import_module(m + "distributions.truncated_normal")
    ns.tnv = importlib.import_module(m + "distributions.truncated_normal_vector")
    ns.gamma = importlib.import_module(m + "distributions.gamma")
    ns.exponential = importlib.import_module(m + "distributions.exponential")
    ns.rtnorm = importlib.import_module(m + "distributions.rtnorm")
    ns.kmeans = importlib.import_module(m + "kmeans.kmeans")
    ns.root = os.path.join(scratch, "BNMTF")
    return ns


if __name__ == "__main__":
    print(build())

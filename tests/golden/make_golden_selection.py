"""Generate tests/golden/model_selection.json from the REAL reference drivers (build container only).

    python tests/golden/make_golden_selection.py

Runs the reference's own code/cross_validation classes (through oracle/ref_shim.py, a temporary py3-rewritten copy
under /tmp) with the reference's own model classes on its toy data, with numpy and python `random` seeded, and stores
every result attribute the drivers expose plus the log files they write.  tests/test_model_selection_*.py replay the
same calls through bnmtf_b200.model_selection / bnmtf_b200.mask.
"""
import contextlib
import importlib
import io
import json
import os
import random
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_shim  # noqa: E402

sys.path.insert(0, HERE)
import selection_cases as cases  # noqa: E402


def main():
    ref = ref_shim.load()
    cv = "BNMTF.code.cross_validation."
    mods = {name: importlib.import_module(cv + name) for name in
            ("mask", "line_search_bnmf", "grid_search_bnmtf", "greedy_search_bnmtf", "line_search_cross_validation",
             "greedy_search_cross_validation", "matrix_cross_validation")}
    drivers = {
        "mask": mods["mask"],
        "LineSearch": mods["line_search_bnmf"].LineSearch,
        "GridSearch": mods["grid_search_bnmtf"].GridSearch,
        "GreedySearch": mods["greedy_search_bnmtf"].GreedySearch,
        "LineSearchCrossValidation": mods["line_search_cross_validation"].LineSearchCrossValidation,
        "GreedySearchCrossValidation": mods["greedy_search_cross_validation"].GreedySearchCrossValidation,
        "MatrixCrossValidation": mods["matrix_cross_validation"].MatrixCrossValidation,
    }
    models = {"bnmf_vb_optimised": ref.bnmf_vb_optimised, "nmf_icm": ref.nmf_icm, "bnmtf_vb_optimised": ref.bnmtf_vb_optimised,
              "nmtf_icm": ref.nmtf_icm, "NMF": ref.NMF}
    data = cases.load_data(ref.root)
    with contextlib.redirect_stdout(io.StringIO()):
        out = cases.run_all(drivers, models, data, tempfile.mkdtemp())
    with open(os.path.join(HERE, "model_selection.json"), "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)
    print("wrote model_selection.json:", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()

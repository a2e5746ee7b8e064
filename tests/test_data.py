"""Loaders and toy generators (bnmtf_b200/data.py) -- CPU.  In the build container they are also compared with the
reference's own functions on the reference's own files (marker `reference`)."""
import importlib
import os
import random

import numpy as np
import pytest

from bnmtf_b200 import data

REF = "/root/reference"


def write_gdsc(path, sep=","):
    rows = [["Cell Line", "Cancer Type", "Tissue", "drugA", "drugB", "drugC"],
            ["c1", "lung", "t1", "1.5", "", "-2.25"],
            ["c2", "skin", "t2", "", "0.75", "3.0"],
            ["c3", "lung", "t1", "-0.5", "2.0", ""]]
    with open(path, "w") as f:
        f.write("\r\n".join(sep.join(r) for r in rows) + "\r\n")


def test_load_gdsc_shifts_the_observed_entries(tmp_path):
    p = str(tmp_path / "gdsc.txt")
    write_gdsc(p)
    X, X_min, M, drugs, cells, cancers, tissues = data.load_gdsc(p)
    assert drugs == ["drugA", "drugB", "drugC"] and cells == ["c1", "c2", "c3"] and cancers[1] == "skin" and tissues[2] == "t1"
    assert np.array_equal(M, [[1, 0, 1], [0, 1, 1], [1, 1, 0]])
    assert np.array_equal(X, [[1.5, 0, -2.25], [0, 0.75, 3.0], [-0.5, 2.0, 0]])
    assert np.array_equal(X_min, np.where(M != 0, X + 3.25, 0.0)) and X_min[M != 0].min() == 1.0
    neg = data.negate_gdsc(X, M)
    assert np.array_equal(neg, np.where(M != 0, -X + 3.0, 0.0))
    with pytest.raises(AssertionError):
        data.load_gdsc(None)


def test_store_gdsc_round_trip(tmp_path):
    p, q = str(tmp_path / "a.txt"), str(tmp_path / "b.txt")
    write_gdsc(p)
    X, _, M, drugs, cells, cancers, tissues = data.load_gdsc(p)
    data.store_gdsc(q, X, M, drugs, cells, cancers, tissues)
    X2, _, M2, drugs2, cells2, _, _ = data.load_gdsc(q, sep="\t")
    assert np.array_equal(X, X2) and np.array_equal(M, M2) and drugs == drugs2 and cells == cells2


def test_load_ccle_and_matrix_pair(tmp_path):
    p = str(tmp_path / "ic50.txt")
    open(p, "w").write("1.0\t\t3.5\n\t2.0\t-1.0\n")
    X, M = data.load_ccle(p)
    assert np.array_equal(M, [[1, 0, 1], [0, 1, 1]]) and np.array_equal(X, [[1.0, 0, 3.5], [0, 2.0, -1.0]])
    np.savetxt(str(tmp_path / "R.txt"), X), np.savetxt(str(tmp_path / "M.txt"), M)
    R2, M2 = data.load_matrix_pair(str(tmp_path / "R.txt"), str(tmp_path / "M.txt"))
    assert np.array_equal(R2, X) and np.array_equal(M2, M)


def test_generators_shapes_and_noise():
    np.random.seed(0), random.seed(0)
    U, V, tau, true_R, R = data.generate_dataset(12, 9, 3, np.ones((12, 3)), 2 * np.ones((9, 3)), 4.0)
    assert U.shape == (12, 3) and V.shape == (9, 3) and np.allclose(true_R, U @ V.T) and R.shape == (12, 9)
    assert 0.2 < np.std(R - true_R) < 0.9                                   # sigma = 1/sqrt(4)
    assert np.array_equal(data.add_noise(true_R, np.inf), true_R)
    F, S, G, _, true3, R3 = data.generate_dataset_nmtf(7, 6, 3, 2, np.ones((7, 3)), np.ones((3, 2)), np.ones((6, 2)), np.inf)
    assert np.allclose(true3, F @ S @ G.T) and np.array_equal(R3, true3)
    M = data.try_generate_M(10, 8, 0.3, attempts=100)
    assert M.shape == (10, 8) and (M.sum(0) > 0).all() and (M.sum(1) > 0).all() and M.sum() == 80 - int(0.3 * 80)
    with pytest.raises(Exception, match="Tried to generate M 3 times"):
        data.try_generate_M(4, 4, 0.95, attempts=3)


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir(REF + "/data_drug_sensitivity"), reason="reference tree not present")
def test_loaders_and_generators_match_the_reference():
    import contextlib
    import io
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import ref_shim
    ref_shim.load()
    rg = importlib.import_module("BNMTF.data_drug_sensitivity.gdsc.load_data")
    rc = importlib.import_module("BNMTF.data_drug_sensitivity.ccle.load_data")
    gd = REF + "/data_drug_sensitivity/gdsc/ic50_excl_empty_filtered_cell_lines_drugs.txt"
    want, got = rg.load_gdsc(location=gd), data.load_gdsc(gd)
    for w, g in zip(want, got):
        assert np.array_equal(w, g) if isinstance(w, np.ndarray) else w == g
    assert got[0].shape == (622, 138) and abs(got[2].mean() - 0.8104) < 1e-3
    assert np.array_equal(rg.negate_gdsc(want[0], want[2]), data.negate_gdsc(got[0], got[2]))
    for name, ic50 in (("ic50.txt", True), ("ec50.txt", False)):
        wX, wM = rc.load_ccle(ic50=ic50)
        gX, gM = data.load_ccle(REF + "/data_drug_sensitivity/ccle/" + name)
        assert np.array_equal(wX, gX) and np.array_equal(wM, gM)
    g2 = importlib.import_module("BNMTF.data_toy.bnmf.generate_bnmf")
    g3 = importlib.import_module("BNMTF.data_toy.bnmtf.generate_bnmtf")
    lam = lambda *s: 0.5 + np.arange(np.prod(s), dtype=float).reshape(s) % 3
    for seed in (0, 5):
        np.random.seed(seed), random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            w2 = g2.generate_dataset(11, 7, 3, lam(11, 3), lam(7, 3), 2.0)
            w3 = g3.generate_dataset(9, 8, 3, 2, lam(9, 3), lam(3, 2), lam(8, 2), 0.5)
            wM = g2.try_generate_M(9, 8, 0.4, 50)
        np.random.seed(seed), random.seed(seed)
        o2 = data.generate_dataset(11, 7, 3, lam(11, 3), lam(7, 3), 2.0)
        o3 = data.generate_dataset_nmtf(9, 8, 3, 2, lam(9, 3), lam(3, 2), lam(8, 2), 0.5)
        oM = data.try_generate_M(9, 8, 0.4, 50)
        for w, o in zip(list(w2) + list(w3) + [wM], list(o2) + list(o3) + [oM]):
            assert np.array_equal(w, o)

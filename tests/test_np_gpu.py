"""GPU parity of the non-probabilistic models (NMF, NMTF multiplicative updates) against the reference goldens and
the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def close(a, b, rtol=1e-9, what=""):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = max(1.0, float(np.max(np.abs(b)))) if b.size else 1.0
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * 1e-2 * scale, err_msg=what)


def test_nmf_np_trajectory_matches_reference(golden):
    from bnmtf_b200.np_models import NMF
    g = golden("toy_nmf_np")
    m = NMF(g["R"], g["M"], int(g["K"]))
    m.initialise("ones")
    m.U, m.V = g["init_U"].copy(), g["init_V"].copy()
    m.run(int(g["its"]))
    close(m.all_performances["MSE"], g["trace_MSE"]), close(m.all_performances["Rp"], g["trace_Rp"])
    close(m.all_performances["R^2"], g["trace_R^2"])
    close(m.U, g["final_U"]), close(m.V, g["final_V"])
    close(m.compute_I_div(), g["final_Idiv"])
    close(m.predict(g["M"])["MSE"], g["trace_MSE"][-1])


def test_nmf_np_single_column_updates_match_oracle(golden):
    from bnmtf_b200.np_models import NMF
    from oracle import bnmtf_oracle as orc
    g = golden("toy_nmf_np")
    m = NMF(g["R"], g["M"], int(g["K"]))
    m.initialise("ones")
    m.U, m.V = g["init_U"].copy(), g["init_V"].copy()
    o = orc.OracleBNMF(g["R"], g["M"], int(g["K"]), mode="np")
    o.set_state(g["init_U"], g["init_V"])
    for k in (0, 3):
        m.update_U(k), o._np_column(k, "U")
        close(m.U, o.U, rtol=1e-12)
        m.update_V(k), o._np_column(k, "V")
        close(m.V, o.V, rtol=1e-12)


def test_nmf_np_run_requires_initialise():
    from bnmtf_b200.np_models import NMF
    m = NMF(np.ones((3, 2)), np.ones((3, 2)), 2)
    with pytest.raises(AssertionError) as e:
        m.run(1)
    assert str(e.value) == "U and V have not been initialised - please run NMF.initialise() first."


def test_nmtf_np_trajectory_matches_reference(golden):
    from bnmtf_b200.np_models import NMTF
    g = golden("toy_nmtf_np")
    m = NMTF(g["R"], g["M"], int(g["K"]), int(g["L"]))
    m.initialise("ones", "ones")
    m.F, m.S, m.G = g["init_F"].copy(), g["init_S"].copy(), g["init_G"].copy()
    m.run(int(g["its"]))
    close(m.all_performances["MSE"], g["trace_MSE"])
    close(m.F, g["final_F"]), close(m.S, g["final_S"]), close(m.G, g["final_G"])
    close(m.compute_I_div(), g["final_Idiv"])


def test_nmtf_np_single_updates_match_oracle(golden):
    from bnmtf_b200.np_models import NMTF
    from oracle import bnmtf_oracle as orc
    g = golden("toy_nmtf_np")
    K, L = int(g["K"]), int(g["L"])
    m = NMTF(g["R"], g["M"], K, L)
    m.initialise("ones", "ones")
    m.F, m.S, m.G = g["init_F"].copy(), g["init_S"].copy(), g["init_G"].copy()
    o = orc.OracleBNMTF(g["R"], g["M"], K, L, mode="np")
    o.set_state(g["init_F"], g["init_S"], g["init_G"])
    m.update_S(1, 2), o.np_update_S(1, 2)
    close(m.S, o.S, rtol=1e-12)
    m.update_F(3), o.np_update_F(3)
    close(m.F, o.F, rtol=1e-12)
    m.update_G(0), o.np_update_G(0)
    close(m.G, o.G, rtol=1e-12)

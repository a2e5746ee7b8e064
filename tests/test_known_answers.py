"""The reference's inline known-answer tests (tests/known_answers.py) applied to the reference's classes (build
container: proves the transcription) and to this package's GPU classes (-m gpu)."""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import known_answers as ka  # noqa: E402

CHECKS = {
    "gibbs_conditionals": lambda ns: ka.check_conditionals(ns.bnmf_gibbs_optimised),
    "icm_conditionals": lambda ns: ka.check_conditionals(ns.nmf_icm, icm=True),
    "gibbs_summaries": lambda ns: ka.check_gibbs_summaries(ns.bnmf_gibbs_optimised),
    "vb_elbo": lambda ns: ka.check_vb_elbo(ns.bnmf_vb_optimised),
    "vb_updates": lambda ns: ka.check_vb_updates(ns.bnmf_vb_optimised),
    "vb_moments": lambda ns: ka.check_vb_moments(ns.bnmf_vb_optimised),
    "constructors_two_factor": lambda ns: [ka.check_constructor_messages(c) for c in
                                           (ns.bnmf_gibbs_optimised, ns.bnmf_vb_optimised, ns.nmf_icm)],
    "constructors_three_factor": lambda ns: [ka.check_constructor_messages(c, three_factor=True) for c in
                                             (ns.bnmtf_gibbs_optimised, ns.bnmtf_vb_optimised, ns.nmtf_icm)],
}


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference/code/models"), reason="reference tree not present")
@pytest.mark.parametrize("check", sorted(CHECKS))
def test_known_answers_hold_for_the_reference_itself(check):
    import contextlib
    import io
    sys.path.insert(0, os.path.join(HERE, ".."))
    from oracle import ref_shim
    with contextlib.redirect_stdout(io.StringIO()):
        CHECKS[check](ref_shim.load())


@pytest.mark.gpu
@pytest.mark.parametrize("check", sorted(CHECKS))
def test_known_answers_hold_for_the_gpu_classes(check):
    import bnmtf_b200
    CHECKS[check](bnmtf_b200)

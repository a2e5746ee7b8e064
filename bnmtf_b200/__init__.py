"""bnmtf_b200 -- B200-native engine behind the ThomasBrouwer/BNMTF model-class API."""
from .bnmf import BNMF_Gibbs, BNMF_VB, bnmf_gibbs_optimised, bnmf_vb_optimised, nmf_icm  # noqa: F401

__all__ = ["bnmf_gibbs_optimised", "bnmf_vb_optimised", "nmf_icm", "BNMF_Gibbs", "BNMF_VB"]

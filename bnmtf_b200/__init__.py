"""bnmtf_b200 -- B200-native engine behind the ThomasBrouwer/BNMTF model-class API."""
from .bnmf import BNMF_Gibbs, BNMF_VB, bnmf_gibbs_optimised, bnmf_vb_optimised, nmf_icm  # noqa: F401
from .bnmtf import BNMTF_Gibbs, BNMTF_VB, bnmtf_gibbs_optimised, bnmtf_vb_optimised, nmtf_icm  # noqa: F401
from .np_models import NMF, NMTF  # noqa: F401

from . import data, mask, model_selection  # noqa: F401
from .model_selection import (DevicePool, GreedySearch, GreedySearchCrossValidation, GridSearch, LineSearch,  # noqa: F401
                              LineSearchCrossValidation, MatrixCrossValidation, MatrixNestedCrossValidation,
                              ParallelMatrixCrossValidation)

nmf_np, nmtf_np = NMF, NMTF   # BASELINE.json's names for the non-probabilistic models

__all__ = ["bnmf_gibbs_optimised", "bnmf_vb_optimised", "nmf_icm", "NMF", "bnmtf_gibbs_optimised", "bnmtf_vb_optimised",
           "nmtf_icm", "NMTF", "BNMF_Gibbs", "BNMF_VB", "BNMTF_Gibbs", "BNMTF_VB", "nmf_np", "nmtf_np",
           "data", "mask", "model_selection", "DevicePool", "LineSearch", "GridSearch", "GreedySearch", "LineSearchCrossValidation",
           "GreedySearchCrossValidation", "MatrixCrossValidation", "ParallelMatrixCrossValidation",
           "MatrixNestedCrossValidation"]

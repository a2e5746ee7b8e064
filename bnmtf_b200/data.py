"""Data path around the models: the reference's text loaders and toy generators, and the
upload of a host (R, M) pair into the device layout the kernels read.  Host code; the on-disk formats are the
reference's.

  load_gdsc / negate_gdsc / store_gdsc   data_drug_sensitivity/gdsc/load_data.py:15-86
  load_ccle                              data_drug_sensitivity/ccle/load_data.py:16-28
  generate_dataset / add_noise / try_generate_M    data_toy/bnmf/generate_bnmf.py:27-66 (two factors),
  generate_dataset_nmtf                            data_toy/bnmtf/generate_bnmtf.py:28-44 (three factors)
  load_matrix_pair                       the `numpy.loadtxt(R.txt), numpy.loadtxt(M.txt)` idiom of every experiment script
  upload                                 (R, M) -> engine.Dataset: R and R^T zero-padded, the mask as bit words in both
                                         orientations (the digit planes are built on first use)

The reference's data files are not shipped with this package: the loaders take the file location (the reference's
default paths inside its own checkout are what `location=None` meant there).  The generators consume numpy's global
random stream in the reference's order (U row-major, V row-major, then the noise row-major), so a seeded call
reproduces the reference's matrices.
"""
import numpy as np

from . import mask as _mask


# ---- GDSC (Sanger) drug sensitivity -------------------------------------------------------------------------
def load_gdsc(location, standardised=False, sep=","):
    """-> (X, X_min, M, drug_names, cell_lines, cancer_types, tissues).  First line: three header cells then the drug
    names; then one line per cell line: name, cancer type, tissue, values ('' = missing).  X has 0 at missing entries;
    X_min = X - (min(X) - 1) on the observed entries (so the smallest observed value becomes 1), 0 elsewhere.
    `standardised` only selected another default file in the reference; it is accepted and ignored."""
    assert location, "load_gdsc: give the location of the GDSC text file (the reference's data is not shipped)."
    with open(location, 'r') as fin:
        lines = [line.split("\n")[0].split("\r")[0].split(sep) for line in fin.readlines()]
    drug_names = lines[0][3:]
    cell_lines, cancer_types, tissues, X, M = [], [], [], [], []
    for line in lines[1:]:
        cell_lines.append(line[0])
        cancer_types.append(line[1])
        tissues.append(line[2])
        X.append([float(v) if v != '' else 0.0 for v in line[3:]])
        M.append([1.0 if v != '' else 0.0 for v in line[3:]])
    X, M = np.array(X, dtype=float), np.array(M, dtype=float)
    minimum = X.min() - 1
    X_min = np.where(M != 0, X - minimum, 0.0)
    return (X, X_min, M, drug_names, cell_lines, cancer_types, tissues)


def negate_gdsc(X, M):
    """-X shifted so that its smallest value (over ALL entries, like the reference) maps to 0; 0 at missing entries."""
    X = -np.asarray(X, dtype=float)
    minimum = X.min()
    return np.where(np.asarray(M) != 0, X - minimum, 0.0)


def store_gdsc(location, X, M, drug_names, cell_lines, cancer_types, tissues):
    """Tab-separated, same columns as load_gdsc reads (with sep='\\t'); nothing is written for a missing value."""
    with open(location, 'w') as fout:
        fout.write("Cell Line\tCancer Type\tTissue\t" + "\t".join(drug_names) + "\n")
        for i, (cell_line, cancer_type, tissue, row) in enumerate(zip(cell_lines, cancer_types, tissues, X)):
            data = [str(val) if M[i][j] else "" for (j, val) in enumerate(row)]
            fout.write(cell_line + "\t" + cancer_type + "\t" + tissue + "\t" + "\t".join(data) + "\n")


# ---- CCLE ------------------------------------------------------------------------------------------------------
def load_ccle(location, delim='\t'):
    """-> (X, M): a bare delimiter-separated matrix, empty / nan cells are missing (X = 0, M = 0 there)."""
    assert location, "load_ccle: give the location of ic50.txt / ec50.txt (the reference's data is not shipped)."
    data = np.genfromtxt(location, delimiter=delim, missing_values=[np.nan])
    data = np.atleast_2d(data)
    known = ~np.isnan(data)
    return np.where(known, data, 0.0), known.astype(float)


def load_matrix_pair(location_R, location_M):
    """The whitespace-separated R.txt / M.txt pair of data_toy/ (numpy.loadtxt on both)."""
    R, M = np.loadtxt(location_R), np.loadtxt(location_M)
    assert R.shape == M.shape, "R and M have different shapes: %s and %s." % (R.shape, M.shape)
    return R, M


# ---- toy generators -------------------------------------------------------------------------------------------
def add_noise(true_R, tau):
    """R_ij ~ N(true_R_ij, 1/tau); tau = inf: a copy."""
    true_R = np.asarray(true_R, dtype=float)
    if np.isinf(tau):
        return np.copy(true_R)
    return np.random.normal(loc=true_R, scale=1.0 / np.sqrt(tau))


def generate_dataset(I, J, K, lambdaU, lambdaV, tau):
    """-> (U, V, tau, true_R, R) with U_ik ~ Exp(lambdaU_ik), V_jk ~ Exp(lambdaV_jk), R = U V^T + noise."""
    U = np.random.exponential(scale=1.0 / np.asarray(lambdaU, dtype=float)[:I, :K])
    V = np.random.exponential(scale=1.0 / np.asarray(lambdaV, dtype=float)[:J, :K])
    true_R = np.dot(U, V.T)
    return (U, V, tau, true_R, add_noise(true_R, tau))


def generate_dataset_nmtf(I, J, K, L, lambdaF, lambdaS, lambdaG, tau):
    """-> (F, S, G, tau, true_R, R) with R = F S G^T + noise."""
    F = np.random.exponential(scale=1.0 / np.asarray(lambdaF, dtype=float)[:I, :K])
    S = np.random.exponential(scale=1.0 / np.asarray(lambdaS, dtype=float)[:K, :L])
    G = np.random.exponential(scale=1.0 / np.asarray(lambdaG, dtype=float)[:J, :L])
    true_R = np.dot(F, np.dot(S, G.T))
    return (F, S, G, tau, true_R, add_noise(true_R, tau))


def try_generate_M(I, J, fraction_unknown, attempts):
    """A random mask with no empty row or column (mask.generate_M, python `random`), or an Exception after `attempts`."""
    for attempt in range(1, attempts + 1):
        M = _mask.generate_M(I, J, fraction_unknown)
        if (M.sum(axis=0) != 0).all() and (M.sum(axis=1) != 0).all():
            return M
    raise Exception("Tried to generate M %s times, with I=%s, J=%s, fraction=%s, but failed." % (attempts, I, J, fraction_unknown))


# ---- host -> device ---------------------------------------------------------------------------------------------
def upload(R, M, device=None, distributed=False):
    """(R, M) host arrays -> engine.Dataset on `device` (default: the current CUDA device): fp64 R and R^T, padded,
    with the float mask packed into bit words (256 MiB instead of 16 GiB at 65536 x 32768).  distributed=True keeps
    only this rank's row shards of R and R^T (an initialised torch.distributed group is required).  The result can be
    passed to `Model.from_dataset(ds, K, priors)` so that several models share one resident copy of the matrix."""
    from .engine import Dataset, require_cuda
    dev = require_cuda(device)
    world, rank = 1, 0
    if distributed:
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
    return Dataset.from_host(np.array(R, dtype=float), np.array(M, dtype=float), dev, world, rank)

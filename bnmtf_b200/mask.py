"""Observation masks and cross-validation folds (the helpers of the reference's code/cross_validation/mask.py).

Masks are I x J float arrays of 0/1 as everywhere in the reference.  The random choices are made with python's
`random` module through exactly the calls the reference makes (`random.sample(range(I*J), n)`, `random.shuffle` of
the row-major list of observed positions), so a seeded run produces the same masks and folds; everything around
those calls is vectorised numpy instead of python loops (the reference builds 65536 x 32768 masks entry by entry).
Reference: code/cross_validation/mask.py:11-174.
"""
import random

import numpy as np


def _observed_positions(M):
    """Flat row-major positions of the non-zero entries: the order of the reference's nonzero_indices()."""
    return np.flatnonzero(np.asarray(M).ravel() != 0)


def _mask_from_flat(shape, flat):
    out = np.zeros(shape[0] * shape[1])
    out[np.asarray(flat, dtype=np.int64)] = 1.0
    return out.reshape(shape)


def check_empty_rows_columns(M):
    """True when every row and every column keeps at least one observed entry (mask.py:136-145)."""
    M = np.asarray(M)
    return bool((M.sum(axis=0) != 0).all() and (M.sum(axis=1) != 0).all())


def generate_M(I, J, fraction):
    """I x J mask with int(I*J*fraction) entries knocked out at random (mask.py:11-16)."""
    M = np.ones(I * J)
    M[random.sample(range(0, I * J), int(I * J * fraction))] = 0.0
    return M.reshape(I, J)


def generate_M_from_M(M, fraction):
    """Split the observed entries of M into (M_train, M_test) so that `fraction` of ALL entries is missing from
    M_train (mask.py:20-39)."""
    M = np.asarray(M)
    I, J = M.shape
    pos = list(_observed_positions(M))
    missing = I * J - len(pos)
    assert missing < I * J * fraction, \
        "Specified %s fraction missing, so %s entries missing, but there are already %s missing by default!" % \
        (fraction, I * J * fraction, missing)
    random.shuffle(pos)
    last = int(I * J * (1 - fraction))
    M_train, M_test = _mask_from_flat((I, J), pos[:last]), _mask_from_flat((I, J), pos[last:])
    assert np.array_equal(M, M_train + M_test), "Tried splitting M into M_test and M_train but something went wrong."
    return M_train, M_test


def try_generate_M_from_M(M, fraction, attempts):
    for _ in range(attempts):
        M_train, M_test = generate_M_from_M(M, fraction)
        if check_empty_rows_columns(M_train):
            return M_train, M_test
    assert False, "Failed to generate folds for training and test data, %s attempts, fraction %s." % (attempts, fraction)


def compute_folds(I, J, no_folds, M=None):
    """The observed entries of M (default: all), shuffled, cut into no_folds test masks (mask.py:51-70)."""
    M = np.ones((I, J)) if M is None else np.asarray(M)
    pos = list(_observed_positions(M))
    n = len(pos)
    random.shuffle(pos)
    cuts = [int(i * n / no_folds) for i in range(no_folds + 1)]
    return [_mask_from_flat((I, J), pos[cuts[i]:cuts[i + 1]]) for i in range(no_folds)]


def compute_folds_attempts(I, J, no_folds, attempts, M=None):
    """compute_folds() until no training mask (M minus a fold) has an empty row or column (mask.py:74-84)."""
    for _ in range(attempts):
        folds = compute_folds(I=I, J=J, no_folds=no_folds, M=M)
        base = np.ones((I, J)) if M is None else np.asarray(M)
        if all(check_empty_rows_columns(base - test) for test in folds):
            return folds
    assert False, "Failed to generate folds for training and test data, %s attempts." % attempts


def compute_crossval_folds_rows_attempts(M, no_rows, no_folds, attempts):
    """Folds over the first no_rows rows only; the other rows always train (mask.py:91-106)."""
    M = np.asarray(M)
    I, J = M.shape
    head, rest = M[:no_rows], M[no_rows:]
    out = []
    for test_head in compute_folds_attempts(no_rows, J, no_folds, attempts, head):
        out.append((np.concatenate((head - test_head, rest), axis=0),
                    np.concatenate((test_head, np.zeros((I - no_rows, J))), axis=0)))
    return out


def compute_crossval_folds_columns_attempts(M, no_columns, no_folds, attempts):
    """Folds over the first no_columns columns only (mask.py:108-123)."""
    M = np.asarray(M)
    I, J = M.shape
    head, rest = M[:, :no_columns], M[:, no_columns:]
    out = []
    for test_head in compute_folds_attempts(I, no_columns, no_folds, attempts, head):
        out.append((np.concatenate((head - test_head, rest), axis=1),
                    np.concatenate((test_head, np.zeros((I, J - no_columns))), axis=1)))
    return out


def compute_Ms(folds_M):
    """Training mask of each fold = sum of the other folds (mask.py:149-152)."""
    folds = [np.asarray(f) for f in folds_M]
    total = sum(folds)
    return [total - f for f in folds]


def calc_inverse_M(M):
    return (np.asarray(M) != 1).astype(float)


def nonzero_indices(M):
    M = np.asarray(M)
    return [(int(p // M.shape[1]), int(p % M.shape[1])) for p in _observed_positions(M)]


def nonzero_row_indices(M):
    return [list(np.flatnonzero(row)) for row in np.asarray(M)]


def nonzero_column_indices(M):
    return [list(np.flatnonzero(col)) for col in np.asarray(M).T]


def recover_predictions(M, X_true, X_pred):
    """(actual, predicted) pairs at the entries with M == 0, row-major (mask.py:167-174)."""
    M, X_true, X_pred = np.asarray(M), np.asarray(X_true), np.asarray(X_pred)
    sel = M == 0
    return list(zip(X_true[sel].tolist(), X_pred[sel].tolist()))

"""The model-selection / cross-validation scenarios, written once and run twice: by make_golden_selection.py with the
reference's drivers and models (-> model_selection.json), and by the tests with bnmtf_b200's drivers and models.
`drivers` maps names to the driver classes (and "mask" to the mask module), `models` maps names to model classes."""
import os
import random

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_data(reference_root=None):
    """Toy inputs: from the reference tree when generating, from the committed fixtures otherwise."""
    if reference_root is not None:
        ld = lambda p: np.loadtxt(os.path.join(reference_root, p))
        return {"bnmf": (ld("data_toy/bnmf/R.txt"), ld("data_toy/bnmf/M.txt")),
                "bnmtf": (ld("data_toy/bnmtf/R.txt"), ld("data_toy/bnmtf/M.txt"))}
    a = np.load(os.path.join(HERE, "toy_bnmf_vb.npz"))
    b = np.load(os.path.join(HERE, "toy_bnmtf_vb.npz"))
    return {"bnmf": (a["R"], a["M"]), "bnmtf": (b["R"], b["M"])}


def seed_all(s):
    np.random.seed(s)
    random.seed(s)


def _listify(x):
    if isinstance(x, dict):
        return {str(k): _listify(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_listify(v) for v in x]
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, (np.floating, np.integer)):
        return x.item()
    return x


PRI2 = {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1}
PRI3 = {"alpha": 1.0, "beta": 1.0, "lambdaF": 0.1, "lambdaS": 0.1, "lambdaG": 0.1}


def case_mask(drivers, models, data, tmp):
    mask = drivers["mask"]
    seed_all(3)
    M = mask.generate_M(10, 8, 0.3)
    tr, te = mask.generate_M_from_M(np.ones((7, 6)), 0.25)
    folds = mask.compute_folds(6, 5, 3, M=(np.arange(30).reshape(6, 5) % 4 != 0).astype(float))
    seed_all(4)
    folds_ok = mask.compute_folds_attempts(9, 7, 4, 100, M=M[:9, :7])
    rows = list(mask.compute_crossval_folds_rows_attempts(M, 6, 2, 100))
    cols = list(mask.compute_crossval_folds_columns_attempts(M, 5, 2, 100))
    return {"generate_M": M, "train": tr, "test": te, "folds": folds, "folds_ok": folds_ok,
            "Ms": mask.compute_Ms(folds), "rows": rows, "cols": cols,
            "inverse": mask.calc_inverse_M(M), "nonzero": [list(t) for t in mask.nonzero_indices(M)][:12],
            "row_idx": [[int(v) for v in r] for r in mask.nonzero_row_indices(M)],
            "col_idx": [[int(v) for v in r] for r in mask.nonzero_column_indices(M)],
            "recovered": [list(t) for t in mask.recover_predictions(M, np.arange(80.).reshape(10, 8), -np.arange(80.).reshape(10, 8))][:10],
            "check": [bool(mask.check_empty_rows_columns(M)), bool(mask.check_empty_rows_columns(np.eye(3) - np.eye(3)))]}


def case_line_search_vb(drivers, models, data, tmp):
    R, M = data["bnmf"]
    seed_all(0)
    ls = drivers["LineSearch"](classifier=models["bnmf_vb_optimised"], values_K=[8, 10, 12], R=R, M=M, priors=PRI2,
                               initUV="random", iterations=15, restarts=2)
    ls.search()
    return {"all": {m: ls.all_values(m) for m in ("BIC", "AIC", "loglikelihood", "MSE", "ELBO")},
            "best": {m: ls.best_value(m) for m in ("BIC", "AIC", "MSE")}}


def case_line_search_icm(drivers, models, data, tmp):
    R, M = data["bnmf"]
    seed_all(1)
    ls = drivers["LineSearch"](classifier=models["nmf_icm"], values_K=[6, 9], R=R, M=M, priors=PRI2, initUV="random",
                               iterations=12, restarts=1)
    ls.search(minimum_TN=0.1)
    return {"all": {m: ls.all_values(m) for m in ("BIC", "AIC", "loglikelihood", "MSE")}, "best": {"BIC": ls.best_value("BIC")}}


def case_grid_search_vb(drivers, models, data, tmp):
    R, M = data["bnmtf"]
    seed_all(2)
    gs = drivers["GridSearch"](classifier=models["bnmtf_vb_optimised"], values_K=[4, 5], values_L=[4, 6], R=R, M=M,
                               priors=PRI3, initS="random", initFG="kmeans", iterations=8, restarts=1)
    gs.search()
    return {"all": {m: gs.all_values(m) for m in ("BIC", "AIC", "loglikelihood", "MSE", "ELBO")},
            "best": {m: list(gs.best_value(m)) for m in ("BIC", "AIC", "MSE")}}


def case_greedy_search_vb(drivers, models, data, tmp):
    R, M = data["bnmtf"]
    seed_all(5)
    I, J = R.shape
    gs = drivers["GreedySearch"](classifier=models["bnmtf_vb_optimised"], values_K=[3, 4, 5], values_L=[3, 4, 5, 6], R=R, M=M,
                                 priors=PRI3, initS="random", initFG="kmeans", iterations=8, restarts=1)
    gs.search("AIC")
    return {"all": {m: [list(t) for t in gs.all_values(m)] for m in ("BIC", "AIC", "loglikelihood", "MSE")},
            "best": {m: list(gs.best_value(m)) for m in ("AIC", "BIC")}}


def case_line_search_cv(drivers, models, data, tmp):
    R, M = data["bnmf"]
    seed_all(6)
    path = os.path.join(tmp, "lscv.txt")
    cv = drivers["LineSearchCrossValidation"](classifier=models["bnmf_vb_optimised"], R=R, M=M, values_K=[8, 11], folds=3,
                                              priors=PRI2, init_UV="random", iterations=10, restarts=1, quality_metric="AIC",
                                              file_performance=path)
    cv.run()
    cv.fout.close()
    return {"log": open(path).read()}


def case_greedy_search_cv(drivers, models, data, tmp):
    R, M = data["bnmtf"]
    seed_all(7)
    path = os.path.join(tmp, "gscv.txt")
    cv = drivers["GreedySearchCrossValidation"](classifier=models["bnmtf_vb_optimised"], R=R, M=M, values_K=[4, 5],
                                                values_L=[4, 5], folds=2, priors=PRI3, init_S="random", init_FG="kmeans",
                                                iterations=6, restarts=1, quality_metric="AIC", file_performance=path)
    cv.run()
    cv.fout.close()
    return {"log": open(path).read()}


def case_matrix_cv_np(drivers, models, data, tmp):
    R, M = data["bnmf"]
    seed_all(8)
    path = os.path.join(tmp, "mcv.txt")
    search = [{"K": 4}, {"K": 7}]
    cv = drivers["MatrixCrossValidation"](method=models["NMF"], X=R, M=M, K=3, parameter_search=search,
                                          train_config={"iterations": 15, "init_UV": "random", "expo_prior": 0.1},
                                          file_performance=path)
    cv.run()
    best = cv.find_best_parameters("MSE", True)
    cv.fout.close()
    return {"performances": cv.performances, "average": cv.average_performances, "allp": cv.all_performances,
            "best": [best[0], best[1]], "log": open(path).read()}


CASES = {"mask": case_mask, "line_search_vb": case_line_search_vb, "line_search_icm": case_line_search_icm,
         "grid_search_vb": case_grid_search_vb, "greedy_search_vb": case_greedy_search_vb,
         "line_search_cv": case_line_search_cv, "greedy_search_cv": case_greedy_search_cv, "matrix_cv_np": case_matrix_cv_np}


def run_all(drivers, models, data, tmp, only=None):
    return {name: _listify(fn(drivers, models, data, tmp)) for name, fn in CASES.items() if only is None or name in only}

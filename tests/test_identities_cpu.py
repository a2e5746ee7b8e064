"""The algebra behind two device kernels, checked in numpy (no GPU): these are the identities csrc/nmtf.cu relies on, stated
against the reference's element-wise definitions."""
import numpy as np


def test_tri_factor_metric_sums_follow_from_the_column_statistics():
    """k_nmtf_mstat: with the column statistics w.r.t. F (c_j = sum_i m r F_i, FF_j = sum_i m F_i F_i^T, s_j = sum_i m F_i)
    and y_j = S G_j, the three masked sums behind MSE / R^2 / Rp (predict_while_running / compute_MSE / compute_R2 / compute_Rp, bnmtf_gibbs_optimised.py:234-258:
    they are taken from the prediction F S G^T there) are y.c, y^T FF y and y.s summed over the columns."""
    rng = np.random.RandomState(0)
    I, J, K, L = 37, 29, 4, 6
    F, S, G = rng.rand(I, K), rng.rand(K, L), rng.rand(J, L)
    R, M = rng.randn(I, J), (rng.rand(I, J) < 0.7).astype(float)
    P = F @ S @ G.T
    direct = np.array([(M * R * P).sum(), (M * P * P).sum(), (M * P).sum()])
    c = (M * R).T @ F                                   # J x K
    FF = np.einsum("ij,ik,il->jkl", M, F, F)            # J x K x K
    s = M.T @ F                                         # J x K
    y = G @ S.T                                         # J x K
    stats = np.array([(y * c).sum(), np.einsum("jk,jkl,jl->", y, FF, y), (y * s).sum()])
    np.testing.assert_allclose(stats, direct, rtol=1e-12)
    # and the squared error the tau update needs: sum m (r - p)^2 = sum m r^2 - 2 sum m r p + sum m p^2
    np.testing.assert_allclose((M * R * R).sum() - 2 * stats[0] + stats[1], (M * (R - P) ** 2).sum(), rtol=1e-11)


def test_s_update_terms_are_blocks_of_one_extended_product():
    """k_nmtf_sq_tiled: the VB update of S_kl (bnmtf_vb_optimised.py:256-266) needs, per (k, l),
        prec = sum_ij m (varF_ik + F_ik^2)(varG_jl + G_jl^2),
        rhs  = sum_ij m (r - p + F_ik S_kl G_jl) F_ik G_jl - cov terms,
    whose data-dependent parts are sums over the rows i of products of A_i = [vec(F_i F_i^T), varF_i] and
    B_i = [vec(GG_i), sv_i] with the row statistics GG_i = sum_j m G_j G_j^T, sv_i = sum_j m varG_j.  Checked here: the
    precision and the two covariance sums of the reference, written element-wise, against the blocks of C = A^T B."""
    rng = np.random.RandomState(1)
    I, J, K, L = 23, 31, 3, 4
    F, vF, G, vG = rng.rand(I, K), rng.rand(I, K) * 0.2, rng.rand(J, L), rng.rand(J, L) * 0.2
    S = rng.rand(K, L)
    M = (rng.rand(I, J) < 0.6).astype(float)
    GG = np.einsum("ij,jl,jn->iln", M, G, G)
    sv = M @ vG
    A = np.concatenate([(F[:, :, None] * F[:, None, :]).reshape(I, K * K), vF], axis=1)
    B = np.concatenate([GG.reshape(I, L * L), sv], axis=1)
    C = A.T @ B
    Mn, N = K * K, L * L
    for k in range(K):
        for l in range(L):
            # precision of S_kl / tau (bnmtf_vb_optimised.py:257)
            prec = (M * np.outer(vF[:, k] + F[:, k] ** 2, vG[:, l] + G[:, l] ** 2)).sum()
            got = C[k * K + k, l * L + l] + C[k * K + k, N + l] + C[Mn + k, l * L + l] + C[Mn + k, N + l]
            np.testing.assert_allclose(got, prec, rtol=1e-12)
            # covariance terms of the mean (cov_term_F, cov_term_G, bnmtf_vb_optimised.py:260-261):
            #   sum_ij m varF_ik G_jl (sum_{l' != l} S_kl' G_jl')   and   sum_ij m F_ik varG_jl (sum_{k' != k} F_ik' S_k'l)
            cov_F = sum(S[k, l2] * (M * np.outer(vF[:, k], G[:, l] * G[:, l2])).sum() for l2 in range(L) if l2 != l)
            cov_G = sum(S[k2, l] * (M * np.outer(F[:, k] * F[:, k2], vG[:, l])).sum() for k2 in range(K) if k2 != k)
            got_F = sum(S[k, l2] * C[Mn + k, l * L + l2] for l2 in range(L) if l2 != l)
            got_G = sum(S[k2, l] * C[k * K + k2, N + l] for k2 in range(K) if k2 != k)
            np.testing.assert_allclose(got_F, cov_F, rtol=1e-12)
            np.testing.assert_allclose(got_G, cov_G, rtol=1e-12)

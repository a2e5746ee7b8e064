// Layer 2 for the tri-factorisation R ~ F S G^T (bnmtf_gibbs_optimised.py, bnmtf_vb_optimised.py, nmtf_icm.py).
//
// Everything is derived from the same masked row statistics as the two-factor model (stats.cu), taken with respect
// to G for the rows of R (RG_i = sum_j m r G_j, GG_i = sum_j m G_j G_j^T, SVg_i = sum_j m varG_j) and with respect
// to F for the rows of R^T:
//   * F (resp. G) columns: k_nmtf_transform turns a row's L-dimensional statistics into the K-dimensional ones of
//     the effective factor X = G S^T (A_i = S GG_i S^T, c_i = S RG_i, VB covariance terms B_i = S diag(SVg_i) S^T)
//     in exactly the layout k_bnmf_row_solve consumes, so the row solver is shared with the two-factor model;
//   * S entries: k_nmtf_sq_partial reduces over rows the (KL x KL) normal matrix
//     H[(k,l),(k',l')] = sum_i F_ik F_ik' GG_i,ll' (+ VB covariance terms), the right-hand side sum_i F_ik RG_il and
//     the precisions; k_coord_solve then performs the K*L sequential scalar updates (reference: each one a full
//     pass over R, bnmtf_gibbs_optimised.py:201-205) on that small system;
//   * k_nmtf_extra: the variance terms of exp_square_diff (bnmtf_vb_optimised.py:239-243) per column of R.
#include "common.cuh"

namespace bnmtf {

// observed-set Gram entry (a,b) of one row from the packed upper-triangular 8x8 tiles
__device__ __forceinline__ double gram_at(const double* __restrict__ gpart, const double* __restrict__ gfull, int polarity,
                                          int nt, int a, int b) {
  if (a > b) { const int t = a; a = b; b = t; }
  const int ta = a >> 3, tb = b >> 3;
  const int p = ta * nt - ta * (ta - 1) / 2 + (tb - ta);
  const int idx = p * 64 + (a & 7) * 8 + (b & 7);
  const double v = gpart[idx];
  return polarity ? v : gfull[idx] - v;
}


__global__ void k_nmtf_transform(TransformArgs a) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int row = blockIdx.x * nw + warp;
  if (row >= a.rows) return;
  const int Ks = a.Ks, Lo = a.Lo;
  const int nto = tiles_for(Lo), nts = tiles_for(Ks);
  const int KPo = 8 * nto, KPs = 8 * nts;
  const int glo = nto * (nto + 1) / 2 * 64, gls = nts * (nts + 1) / 2 * 64;
  double* GG = sm + (size_t)warp * (Lo * Lo + Ks * Lo + Lo);
  double* T = GG + Lo * Lo;       // S * GG   (Ks x Lo)
  double* sv = T + Ks * Lo;       // observed variance sums (Lo)
  const double* gp = a.Go + (size_t)row * glo;
  const double* gf = a.Gfull_o;
  for (int i = lane; i < Lo * Lo; i += 32) GG[i] = gram_at(gp, gf, a.polarity, nto, i / Lo, i % Lo);
  for (int l = lane; l < Lo; l += 32)
    sv[l] = a.vb ? (a.polarity ? a.SVo[(size_t)row * KPo + l] : gf[glo + l] - a.SVo[(size_t)row * KPo + l]) : 0.0;
  __syncwarp();
  for (int i = lane; i < Ks * Lo; i += 32) {
    const int k = i / Lo, l = i % Lo;
    double s = 0.0;
    for (int q = 0; q < Lo; ++q) s = fma(a.Smat[k * Lo + q], GG[q * Lo + l], s);
    T[i] = s;
  }
  __syncwarp();
  // c_k = sum_l S_kl RX_l ;  VB precision extra  sv'_k = sum_l (S~_kl (GG_ll + sv_l) - S_kl^2 GG_ll)
  const double* rx = a.RXo + (size_t)row * KPo;
  for (int k = lane; k < KPs; k += 32) {
    double c = 0.0, e = 0.0;
    if (k < Ks) {
      for (int l = 0; l < Lo; ++l) {
        const double s = a.Smat[k * Lo + l];
        c = fma(s, rx[l], c);
        if (a.vb) {
          const double d = GG[l * Lo + l];
          e += (a.varS[k * Lo + l] + s * s) * (d + sv[l]) - s * s * d;
        }
      }
    }
    a.RXs[(size_t)row * KPs + k] = c;
    if (a.vb) a.SVs[(size_t)row * KPs + k] = e;
  }
  // effective Gram tiles: A_kk' = sum_l T_kl S_k'l ; off-diagonal VB covariance B_kk' = sum_l S_kl S_k'l sv_l
  double* go = a.Gs + (size_t)row * gls;
  for (int i = lane; i < gls; i += 32) {
    const int p = i >> 6, r = (i >> 3) & 7, c = i & 7;
    int ta = 0, rem = p;
    while (rem >= nts - ta) { rem -= nts - ta; ++ta; }
    const int k = 8 * ta + r, k2 = 8 * (ta + rem) + c;
    double v = 0.0;
    if (k < Ks && k2 < Ks) {
      for (int l = 0; l < Lo; ++l) {
        const double s2 = a.Smat[k2 * Lo + l];
        v = fma(T[k * Lo + l], s2, v);
        if (a.vb && k != k2) v = fma(a.Smat[k * Lo + l] * s2, sv[l], v);
      }
    }
    go[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------------
// S phase, reduction over rows.  Per CTA partial of  H (D x D), prec (D), rhs (D),  D = K*L, d = k*L + l.
// ---------------------------------------------------------------------------------------------------

// GMEM_ACC: the accumulators do not fit shared memory (K*L > ~150): each CTA accumulates straight into its own slice of
// `partial` (every entry is only ever touched by the same thread, so no atomics and no extra synchronisation).
template <bool GMEM_ACC>
__global__ void __launch_bounds__(256) k_nmtf_sq_partial(SqArgs a) {
  extern __shared__ double sm[];
  const int K = a.K, L = a.L, D = K * L;
  const int ntl = tiles_for(L), KPl = 8 * ntl, gll = ntl * (ntl + 1) / 2 * 64;
  double* H = GMEM_ACC ? a.partial + (size_t)blockIdx.x * ((size_t)D * D + 2 * D) : sm;   // D*D + 2D accumulators
  double* GG = GMEM_ACC ? sm : H + D * D + 2 * D; // L*L
  double* sv = GG + L * L;        // L
  double* rg = sv + L;            // L
  double* f = rg + L;             // K
  double* vf = f + K;             // K
  const int tot = D * D + 2 * D;
  for (int i = threadIdx.x; i < tot; i += 256) H[i] = 0.0;
  for (int row = blockIdx.x; row < a.rows; row += gridDim.x) {
    __syncthreads();
    const double* gp = a.Go + (size_t)row * gll;
    for (int i = threadIdx.x; i < L * L; i += 256) GG[i] = gram_at(gp, a.Gfull_o, a.polarity, ntl, i / L, i % L);
    for (int l = threadIdx.x; l < L; l += 256) {
      rg[l] = a.RXo[(size_t)row * KPl + l];
      sv[l] = a.vb ? (a.polarity ? a.SVo[(size_t)row * KPl + l] : a.Gfull_o[gll + l] - a.SVo[(size_t)row * KPl + l]) : 0.0;
    }
    for (int k = threadIdx.x; k < K; k += 256) {
      f[k] = a.F[(size_t)row * K + k];
      vf[k] = a.vb ? a.varF[(size_t)row * K + k] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D * D; i += 256) {
      const int d = i / D, d2 = i - d * D;
      const int k = d / L, l = d - k * L, k2 = d2 / L, l2 = d2 - k2 * L;
      double v = f[k] * f[k2] * GG[l * L + l2];
      if (a.vb) {
        if (l == l2 && k != k2) v = fma(f[k] * f[k2], sv[l], v);         // cov_term_G
        if (k == k2 && l != l2) v = fma(vf[k], GG[l * L + l2], v);        // cov_term_F
      }
      H[i] += v;
    }
    for (int d = threadIdx.x; d < D; d += 256) {
      const int k = d / L, l = d - k * L;
      const double gd = GG[l * L + l];
      H[D * D + d] += a.vb ? (vf[k] + f[k] * f[k]) * (gd + sv[l]) : f[k] * f[k] * gd;   // precision / tau
      H[D * D + D + d] += f[k] * rg[l];                                                 // right-hand side
    }
  }
  __syncthreads();
  if (!GMEM_ACC)
    for (int i = threadIdx.x; i < tot; i += 256) a.partial[(size_t)blockIdx.x * tot + i] = H[i];
}

// ---------------------------------------------------------------------------------------------------
// k_coord_solve: sequential scalar updates of x (D entries) on the small system (H, prec, rhs):
//   s = rhs_d - sum_{d' != d} H[d][d'] x_d' ; tau_d = tau * prec_d ; mu_d = 1/tau_d (-lambda_d + tau s)
// Single CTA.  Reproduces tauS/muS + TN_draw / TN moments / TN_mode for every (k,l) in `order`.
// ---------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_coord_solve(CoordArgs a) {
  extern __shared__ double xs[];   // D
  __shared__ double red[8];
  __shared__ double bcast;
  const int D = a.D;
  for (int i = threadIdx.x; i < D; i += 256) xs[i] = a.x[i];
  __syncthreads();
  const double tau = a.scalars[S_TAU];
  const unsigned long long it = a.iter ? *a.iter : 0ull;
  for (int o = 0; o < a.n_order; ++o) {
    const int d = a.order ? a.order[o] : o;
    double part = 0.0;
    for (int i = threadIdx.x; i < D; i += 256)
      if (i != d) part = fma(a.H[(size_t)d * D + i], xs[i], part);
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      double dot = 0.0;
      for (int w = 0; w < 8; ++w) dot += red[w];
      const double s = a.rhs[d] - dot;
      const double tau_d = tau * a.prec[d];
      const double mu_d = (1.0 / tau_d) * (-a.lambda[d] + tau * s);
      double val = xs[d], vv = 0.0;
      if (a.apply) {
        if (a.mode == MODE_GIBBS) {
          Philox rng(a.seed, it * 16ull + a.salt, (unsigned long long)d);
          val = tn_draw(mu_d, tau_d, rng);
        } else if (a.mode == MODE_VB) {
          tn_moments(mu_d, tau_d, val, vv);
          a.var[d] = vv;
        } else {
          val = (mu_d != mu_d) ? mu_d : fmax(mu_d, 0.0);
          val = (val != val) ? val : fmax(val, a.min_tn);
        }
        a.x[d] = val;
      }
      if (a.mu) a.mu[d] = mu_d;
      if (a.tauf) a.tauf[d] = tau_d;
      bcast = val;
    }
    __syncthreads();
    if (threadIdx.x == 0) xs[d] = bcast;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
// k_nmtf_extra: per row of R^T (column j of R), with statistics w.r.t. F and the current G_j, varG_j, S, varS:
//   t2 = sum_kl [ G~_jl S~_kl (FF_kk + SVf_k) - G_jl^2 S_kl^2 FF_kk ]
//   t3 = sum_k SVf_k [ (S G_j)_k^2 - sum_l S_kl^2 G_jl^2 ]
//   t4 = sum_l varG_jl [ (S^T FF S)_ll - sum_k S_kl^2 FF_kk ]
// (the three variance terms of exp_square_diff, bnmtf_vb_optimised.py:240-243, regrouped by column).
// ---------------------------------------------------------------------------------------------------

__global__ void k_nmtf_extra(ExtraArgs a) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int row = blockIdx.x * nw + warp;
  if (row >= a.rows) return;
  const int K = a.K, L = a.L;
  const int ntk = tiles_for(K), KPk = 8 * ntk, glk = ntk * (ntk + 1) / 2 * 64;
  double* FF = sm + (size_t)warp * (K * K + K);
  double* sv = FF + K * K;
  const double* gp = a.Go + (size_t)row * glk;
  for (int i = lane; i < K * K; i += 32) FF[i] = gram_at(gp, a.Gfull_o, a.polarity, ntk, i / K, i % K);
  for (int k = lane; k < K; k += 32)
    sv[k] = a.polarity ? a.SVo[(size_t)row * KPk + k] : a.Gfull_o[glk + k] - a.SVo[(size_t)row * KPk + k];
  __syncwarp();
  const double* g = a.G + (size_t)row * L;
  const double* vg = a.varG + (size_t)row * L;
  double acc = 0.0;
  for (int i = lane; i < K * L; i += 32) {                          // t2
    const int k = i / L, l = i - k * L;
    const double s = a.S[i], gl = g[l], d = FF[k * K + k];
    acc += (vg[l] + gl * gl) * (a.varS[i] + s * s) * (d + sv[k]) - gl * gl * s * s * d;
  }
  for (int k = lane; k < K; k += 32) {                              // t3
    double sg = 0.0, sq = 0.0;
    for (int l = 0; l < L; ++l) { const double s = a.S[k * L + l]; sg = fma(s, g[l], sg); sq = fma(s * s, g[l] * g[l], sq); }
    acc += sv[k] * (sg * sg - sq);
  }
  for (int l = lane; l < L; l += 32) {                              // t4
    double q = 0.0, sq = 0.0;
    for (int k = 0; k < K; ++k) {
      double t = 0.0;
      for (int k2 = 0; k2 < K; ++k2) t = fma(FF[k * K + k2], a.S[k2 * L + l], t);
      q = fma(a.S[k * L + l], t, q);
      sq = fma(a.S[k * L + l] * a.S[k * L + l], FF[k * K + k], sq);
    }
    acc += vg[l] * (q - sq);
  }
  acc = warp_sum(acc);
  if (lane == 0) a.extra[row] = acc;
}

// ---- launchers -------------------------------------------------------------------------------------------
static int pick_warps(size_t per_warp_bytes, int max_warps) {
  int w = (int)(96 * 1024 / (per_warp_bytes ? per_warp_bytes : 1));
  if (w > max_warps) w = max_warps;
  return w < 1 ? 1 : w;
}

int launch_nmtf_transform(const TransformArgs& a, cudaStream_t st) {
  const size_t per = ((size_t)a.Lo * a.Lo + (size_t)a.Ks * a.Lo + a.Lo) * sizeof(double);
  const int warps = pick_warps(per, 4);
  const size_t smem = per * warps;
  cudaFuncSetAttribute(k_nmtf_transform, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k_nmtf_transform<<<(a.rows + warps - 1) / warps, warps * 32, smem, st>>>(a);
  return check_launch("nmtf_transform");
}

int sq_partial_len(int K, int L) { const int D = K * L; return D * D + 2 * D; }

int launch_nmtf_sq(const SqArgs& a, int nparts, double* out, cudaStream_t st);   // defined below
__global__ void k_sum_partials2(const double* __restrict__ partial, int nparts, int len, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * len + i];
  out[i] = s;
}
int launch_nmtf_sq(const SqArgs& a, int nparts, double* out, cudaStream_t st) {
  const int D = a.K * a.L;
  const size_t small = ((size_t)a.L * a.L + 2 * a.L + 2 * a.K) * sizeof(double);
  const size_t smem = ((size_t)D * D + 2 * D) * sizeof(double) + small;
  if (smem <= 200 * 1024) {
    cudaFuncSetAttribute(k_nmtf_sq_partial<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_nmtf_sq_partial<false><<<nparts, 256, smem, st>>>(a);
  } else {
    // the reference has no size limit (its grid searches go to K, L = 20..30): accumulate in global memory instead
    k_nmtf_sq_partial<true><<<nparts, 256, small, st>>>(a);
  }
  const int len = D * D + 2 * D;
  k_sum_partials2<<<(len + 127) / 128, 128, 0, st>>>(a.partial, nparts, len, out);
  return check_launch("nmtf_sq");
}

int launch_coord_solve(const CoordArgs& a, cudaStream_t st) {
  k_coord_solve<<<1, 256, (size_t)a.D * sizeof(double), st>>>(a);
  return check_launch("coord_solve");
}

int launch_nmtf_extra(const ExtraArgs& a, cudaStream_t st) {
  const size_t per = ((size_t)a.K * a.K + a.K) * sizeof(double);
  const int warps = pick_warps(per, 4);
  cudaFuncSetAttribute(k_nmtf_extra, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k_nmtf_extra<<<(a.rows + warps - 1) / warps, warps * 32, per * warps, st>>>(a);
  return check_launch("nmtf_extra");
}

}  // namespace bnmtf

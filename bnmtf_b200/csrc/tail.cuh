// End of a sweep (tau update, metrics from the reduced sums, ELBO, trace row): shared by the multi-kernel sweep
// (solve.cu::k_bnmf_finish) and the single-kernel sweep of small problems (small.cu).
#pragma once
#include "common.cuh"

namespace bnmtf {

__device__ inline double gamma_draw_dev(double shape, double rate, Philox& rng) {
  // Marsaglia & Tsang (2000); shape < 1 boosted by U^(1/shape)
  double boost = 1.0;
  if (shape < 1.0) { double u0, u1; rng.uniform2(u0, u1); boost = pow(u0, 1.0 / shape); shape += 1.0; }
  const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (int tries = 0; tries < 1000; ++tries) {
    double u0, u1; rng.uniform2(u0, u1);
    const double x = normcdfinv(u0);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    if (log(u1) < 0.5 * x * x + d - d * v + d * log(v)) return boost * d * v / rate;
  }
  return boost * d / rate;
}

// one thread
__device__ inline void finish_sweep(const FinishArgs& a) {
  double* S = a.scalars;
  const double n = a.m8[6], se2 = a.m8[0], sp = a.m8[1], sp2 = a.m8[2], srp = a.m8[3], sr = a.m8[4], sr2 = a.m8[5];
  const double mean_r = sr / n, mean_p = sp / n;
  const double ss_tot = sr2 - sr * mean_r;
  const double cov = srp - sr * mean_p;
  const double var_p = sp2 - sp * mean_p;
  S[S_SUM_E2] = se2; S[S_SUM_R] = sr; S[S_SUM_R2] = sr2; S[S_OMEGA] = n;
  S[S_MSE] = se2 / n;
  S[S_R2] = (ss_tot != 0.0) ? 1.0 - se2 / ss_tot : INFINITY;
  S[S_RP] = cov / (sqrt(ss_tot) * sqrt(var_p));
  const unsigned long long it = *a.iter;
  const double alpha_s = a.alpha + 0.5 * n;
  S[S_ALPHA_S] = alpha_s;
  double elbo = 0.0;
  if (a.mode == MODE_VB) {
    const double esd = se2 + a.ex1[0];
    S[S_ESD] = esd;
    if (a.update_tau) {
      const double beta_s = a.beta + 0.5 * esd;
      S[S_BETA_S] = beta_s;
      S[S_TAU] = alpha_s / beta_s;
      S[S_LOGTAU] = a.digamma_alpha_s - log(beta_s);
    }
    const double beta_s = S[S_BETA_S], et = S[S_TAU], elt = S[S_LOGTAU];
    elbo = n / 2.0 * (elt - kLog2Pi) - et / 2.0 * esd + a.el8[0]
         + a.alpha * log(a.beta) - a.lgamma_alpha + (a.alpha - 1.0) * elt - a.beta * et
         - alpha_s * log(beta_s) + a.lgamma_alpha_s - (alpha_s - 1.0) * elt + beta_s * et
         + a.el8[1] + a.n_factor_elems / 2.0 * kLog2Pi;
    S[S_ELBO] = elbo;
  } else if (a.update_tau) {
    const double beta_s = a.beta + 0.5 * se2;
    S[S_BETA_S] = beta_s;
    if (a.mode == MODE_GIBBS) {
      Philox rng(a.seed, it * 16ull + 15ull, 0ull);
      S[S_TAU] = gamma_draw_dev(alpha_s, beta_s, rng);
    } else {
      S[S_TAU] = (alpha_s - 1.0) / beta_s;
    }
  }
  unsigned long long row = it, lim = (unsigned long long)a.trace_cap;
  if (a.trace_window) { lim = a.trace_window[1]; row = it - a.trace_window[0]; if (it < a.trace_window[0]) lim = 0; }
  if (a.trace && it < lim) {
    double* tr = a.trace + row * kTraceWidth;
    tr[0] = S[S_TAU]; tr[1] = S[S_MSE]; tr[2] = S[S_R2]; tr[3] = S[S_RP]; tr[4] = elbo; tr[5] = se2; tr[6] = S[S_ESD];
    tr[7] = S[S_LOGTAU];
  }
  *a.iter = it + 1;
}


}  // namespace bnmtf

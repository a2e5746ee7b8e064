// Layer 2 for the two-factor model R ~ U V^T: per-row sequential column updates from the row statistics,
// the masked prediction metrics, the noise-precision update and the small elementwise helpers.
#include <stdlib.h>
#include "common.cuh"
#include "tail.cuh"

namespace bnmtf {

// ---------------------------------------------------------------------------------------------------
// padded factor buffers
// ---------------------------------------------------------------------------------------------------
__global__ void k_pad_factor(const double* __restrict__ X, const double* __restrict__ Var, int n, int K, int KP,
                             int n_alloc, double* __restrict__ Xp, double* __restrict__ Vp) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n_alloc * KP) return;
  const int j = (int)(i / KP), k = (int)(i - (size_t)j * KP);
  double x = 0.0, v = 0.0;
  if (j < n) {
    if (k < K) { x = X[(size_t)j * K + k]; if (Var) v = Var[(size_t)j * K + k]; }
    else if (k == K) x = 1.0;
  }
  Xp[i] = x;
  if (Vp) Vp[i] = v;
}

// ---------------------------------------------------------------------------------------------------
// k_bnmf_row_solve: one warp per row.
//   G_obs = polarity ? sum_seg Gpart : Gfull - sum_seg Gpart      (K x K, symmetric, in shared memory)
//   for k in order:   s   = RX_k - sum_{k' != k} G_obs[k][k'] u_k'          (the reference's masked row sum)
//                     b   = G_obs[k][k] (+ sum_obs Var_k for VB)
//                     tau_k = tau * b ;  mu_k = 1/tau_k * (-lambda_k + tau * s)
//                     u_k <- draw | (mean, variance) | max(mu,0,min_tn)
// which is bnmf_gibbs_optimised.py:134-137,167-171 / bnmf_vb_optimised.py:132-134,189-204 / nmf_icm.py:126-130
// evaluated for one row with the other factor fixed.  The K updates of a row are sequential (Gauss-Seidel) as
// in the reference; different rows are independent.
// ---------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(128) k_bnmf_row_solve(RowSolveArgs a) {
  // observed_flag: the dynamic-range guard made the gated fp64 kernels recompute the statistics over the OBSERVED set
  const int pol = a.polarity | (a.observed_flag ? *a.observed_flag : 0);
  constexpr int KP = 8 * NT;
  constexpr int GS = KP + 1;
  constexpr int NTP = NT * (NT + 1) / 2;
  constexpr int CPL = (KP + 31) / 32;
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= a.rows) return;
  double* G = smem + (size_t)warp * (KP * GS);
  const int K = a.K;

  // Gram tiles -> full symmetric matrix
  for (int i = lane; i < NTP * 64; i += 32) {
    const int p = i >> 6, rc = i & 63, r = rc >> 3, c = rc & 7;
    int ta = 0, rem = p;
    while (rem >= NT - ta) { rem -= NT - ta; ++ta; }
    const int tb = ta + rem;
    double s = 0.0;
    for (int sgm = 0; sgm < a.nseg_g; ++sgm) s += a.Gpart[((size_t)sgm * a.rows + row) * (NTP * 64) + i];
    const double v = pol ? s : a.Gfull[i] - s;
    const int kr = 8 * ta + r, kc = 8 * tb + c;
    G[kr * GS + kc] = v;
    if (ta != tb) G[kc * GS + kr] = v;
  }
  double rx[CPL], sv[CPL], u[CPL], vr[CPL];
#pragma unroll
  for (int q = 0; q < CPL; ++q) {
    const int c = lane + 32 * q;
    rx[q] = sv[q] = u[q] = vr[q] = 0.0;
    if (c < KP) {
      for (int sgm = 0; sgm < a.nseg_rx; ++sgm) rx[q] += a.RXpart[((size_t)sgm * a.rows + row) * KP + c];
      if (a.mode == MODE_VB) {
        double s = 0.0;
        for (int sgm = 0; sgm < a.nseg_g; ++sgm) s += a.SVpart[((size_t)sgm * a.rows + row) * KP + c];
        sv[q] = pol ? s : a.Gfull[NTP * 64 + c] - s;
      }
    }
    if (c < K) {
      u[q] = a.fac[(size_t)row * K + c];
      if (a.mode == MODE_VB) vr[q] = a.var[(size_t)row * K + c];
    }
  }
  __syncwarp();
  const double tau = a.scalars[S_TAU];
  const unsigned long long it = a.iter ? *a.iter : 0ull;

  for (int o = 0; o < a.n_order; ++o) {
    const int k = a.order ? a.order[o] : o;
    double part = 0.0;
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      const int c = lane + 32 * q;
      if (c < K && c != k) part += G[k * GS + c] * u[q];
    }
    const double dot = warp_sum(part);
    double rxk = rx[0], svk = sv[0];
#pragma unroll
    for (int q = 1; q < CPL; ++q) if ((k >> 5) == q) { rxk = rx[q]; svk = sv[q]; }
    rxk = __shfl_sync(0xffffffffu, rxk, k & 31);
    svk = __shfl_sync(0xffffffffu, svk, k & 31);
    const double s = rxk - dot;
    const double gkk = G[k * GS + k];
    const double b = (a.mode == MODE_VB) ? gkk + svk : gkk;
    const double lam = a.lambda[(size_t)row * K + k];
    const double tau_k = tau * b;
    const double mu_k = (1.0 / tau_k) * (-lam + tau * s);
    double val = 0.0, vv = 0.0;
    if (a.apply) {
      if (a.mode == MODE_GIBBS) {
        Philox rng(a.seed, it * 16ull + a.salt, (unsigned long long)(a.row_offset + row) * K + k);
        val = tn_draw(mu_k, tau_k, rng);
      } else if (a.mode == MODE_VB) {
        tn_moments(mu_k, tau_k, val, vv);
      } else {
        val = (mu_k != mu_k) ? mu_k : fmax(mu_k, 0.0);   // numpy.maximum propagates NaN (nmf_icm.py:129)
        val = (val != val) ? val : fmax(val, a.min_tn);
      }
#pragma unroll
      for (int q = 0; q < CPL; ++q)
        if (lane + 32 * q == k) { u[q] = val; vr[q] = vv; }
    }
    if (lane == 0) {
      const size_t idx = (size_t)row * K + k;
      if (a.apply) { a.fac[idx] = val; if (a.mode == MODE_VB) a.var[idx] = vv; }
      if (a.mu) a.mu[idx] = mu_k;
      if (a.tauf) a.tauf[idx] = tau_k;
      if (a.sterm) a.sterm[idx] = s;
    }
  }
  if (a.mstat) {
    // masked prediction sums of this row from its statistics (p_ij = u_i . x_j over the observed j):
    //   sum r p = u . RX,   sum p^2 = u^T G_obs u,   sum p = u . (masked column sums = slot (k, K) of G_obs)
    __syncwarp();
    double* uu = G + K * GS;                             // row K of the staged matrix is free below column K
    double colsum[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      const int c = lane + 32 * q;
      colsum[q] = (c < K) ? G[c * GS + K] : 0.0;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      const int c = lane + 32 * q;
      if (c < K) uu[c] = u[q];
    }
    __syncwarp();
    double rp = 0.0, pp = 0.0, sp = 0.0;
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      const int c = lane + 32 * q;
      if (c < K) {
        double t = 0.0;
        for (int k2 = 0; k2 < K; ++k2) t = fma(G[c * GS + k2], uu[k2], t);
        pp = fma(u[q], t, pp);
        rp = fma(u[q], rx[q], rp);
        sp = fma(u[q], colsum[q], sp);
      }
    }
    rp = warp_sum(rp); pp = warp_sum(pp); sp = warp_sum(sp);
    if (lane == 0) *reinterpret_cast<double4*>(a.mstat + (size_t)row * 4) = make_double4(rp, pp, sp, 0.0);
    __syncwarp();
  }
  if (a.extra) {
    double e = 0.0;
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      const int c = lane + 32 * q;
      if (c < K) e += vr[q] * (G[c * GS + c] + sv[q]) + u[q] * u[q] * sv[q];
    }
    e = warp_sum(e);
    if (lane == 0) a.extra[row] = e;
  }
  // fused exchange: the finished row goes straight into every peer's replicated copy (one coalesced K-double store per
  // peer over NVLink) while the other warps are still in their chains; the host-side barrier follows the kernel
  if (a.peer_fac) {
    const size_t g0 = (size_t)(a.row_offset + row) * K;
    for (int r = 0; r < a.n_peers; ++r) {
      if (r == a.my_rank) continue;
      double* pf = a.peer_fac[r] + g0;
      double* pv = a.peer_var ? a.peer_var[r] + g0 : nullptr;
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        const int c = lane + 32 * q;
        if (c < K) { pf[c] = u[q]; if (pv) pv[c] = vr[q]; }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// k_bnmf_row_solve_lane: the same per-row update with ONE THREAD per row, for the natural column order.
//
// The K updates of a row form a serial chain whose links are a truncated-normal draw / moment evaluation (a few
// hundred dependent fp64 instructions).  With a warp per row all 32 lanes walk that chain redundantly and a
// 65536-row phase needs ~10 waves of resident warps, each as long as the chain (0.41-0.47 ms, fp64 pipe 35 % busy
// with redundant work).  With a thread per row the whole phase is one wave of 14 warps per SM, and the Gram row is
// read exactly once from its packed tiles (row k of the upper triangle, contiguous 64-byte pieces):
//     dot_k = acc[k] + sum_{c>k} G[k][c] u_c(old),   then   acc[c] += G[k][c] u_k(new)  for c > k,
// i.e. acc[k] = sum_{c<k} G[c][k] u_c(new) has been pushed by the earlier columns (G is symmetric).  The per-row
// outputs for the statistics-based metrics and the VB extra term accumulate along the way:
//     sum p^2 = u^T G u = sum_k u_k (G_kk u_k + 2 acc[k]),   sum r p = sum_k u_k RX_k,   sum p = sum_k u_k G[k][K].
// u and acc live in shared memory ([column][thread], conflict-free), so K is a runtime value.
// ---------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(32, 14) k_bnmf_row_solve_lane(RowSolveArgs a) {
  // observed_flag: the dynamic-range guard made the gated fp64 kernels recompute the statistics over the OBSERVED set
  const int pol = a.polarity | (a.observed_flag ? *a.observed_flag : 0);
  constexpr int KP = 8 * NT;
  constexpr int NTP = NT * (NT + 1) / 2;
  __shared__ double u_s[KP][32];
  __shared__ double acc_s[KP][32];
  const int lane = threadIdx.x;
  const int row0 = blockIdx.x * 32;
  const int row = row0 + lane;
  const int K = a.K;
  const bool vb = a.mode == MODE_VB;
  if (row < a.rows) {                                  // (threads past the end only join the cooperative exchange below)
  for (int c = 0; c < KP; ++c) {
    u_s[c][lane] = c < K ? a.fac[(size_t)row * K + c] : 0.0;
    acc_s[c][lane] = 0.0;
  }
  const double tau = a.scalars[S_TAU];
  const unsigned long long it = a.iter ? *a.iter : 0ull;
  const size_t gstride = (size_t)a.rows * (NTP * 64);
  const double* grow = a.Gpart + (size_t)row * (NTP * 64);
  double rp = 0.0, pp = 0.0, sp = 0.0, ex = 0.0;

  // the per-row vectors this thread will read one element per column (RX, SV, lambda, var): pull their lines towards L2
  // now; none of the loads below depends on the update chain, only their consumers do
  {
    const char* l0 = reinterpret_cast<const char*>(a.lambda + (size_t)row * K);
    const char* r0 = reinterpret_cast<const char*>(a.RXpart + (size_t)row * KP);
    for (int off = 0; off < K * 8; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(l0 + off));
    for (int off = 0; off < KP * 8; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(r0 + off));
  }
  for (int k = 0; k < K; ++k) {
    const int ta = k >> 3, r = k & 7;
    if (k + 1 < K) {                           // row k+1 of the Gram tiles: in L2 by the time the chain gets there
      const int tn = (k + 1) >> 3, rn = (k + 1) & 7;
      for (int sgm = 0; sgm < a.nseg_g; ++sgm)
        for (int tb = tn; tb < NT; ++tb)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(grow + sgm * gstride + tile_pair(tn, tb, NT) * 64 + rn * 8));
    }
    // row k of the upper triangle: tiles (ta, tb >= ta), 8 contiguous doubles each
    double g[KP];
#pragma unroll
    for (int tb = 0; tb < NT; ++tb) {
      if (tb >= ta) {
        const int off = tile_pair(ta, tb, NT) * 64 + r * 8;
        double4 lo = make_double4(0, 0, 0, 0), hi = make_double4(0, 0, 0, 0);
        for (int sgm = 0; sgm < a.nseg_g; ++sgm) {
          const double4 x = *reinterpret_cast<const double4*>(grow + sgm * gstride + off);
          const double4 y = *reinterpret_cast<const double4*>(grow + sgm * gstride + off + 4);
          lo.x += x.x; lo.y += x.y; lo.z += x.z; lo.w += x.w;
          hi.x += y.x; hi.y += y.y; hi.z += y.z; hi.w += y.w;
        }
        if (!pol) {
          const double4 x = *reinterpret_cast<const double4*>(a.Gfull + off);
          const double4 y = *reinterpret_cast<const double4*>(a.Gfull + off + 4);
          lo.x = x.x - lo.x; lo.y = x.y - lo.y; lo.z = x.z - lo.z; lo.w = x.w - lo.w;
          hi.x = y.x - hi.x; hi.y = y.y - hi.y; hi.z = y.z - hi.z; hi.w = y.w - hi.w;
        }
        g[8 * tb + 0] = lo.x; g[8 * tb + 1] = lo.y; g[8 * tb + 2] = lo.z; g[8 * tb + 3] = lo.w;
        g[8 * tb + 4] = hi.x; g[8 * tb + 5] = hi.y; g[8 * tb + 6] = hi.z; g[8 * tb + 7] = hi.w;
      }
    }
    double rxk = 0.0, svk = 0.0;
    for (int sgm = 0; sgm < a.nseg_rx; ++sgm) rxk += a.RXpart[((size_t)sgm * a.rows + row) * KP + k];
    if (vb) {
      double t = 0.0;
      for (int sgm = 0; sgm < a.nseg_g; ++sgm) t += a.SVpart[((size_t)sgm * a.rows + row) * KP + k];
      svk = pol ? t : a.Gfull[NTP * 64 + k] - t;
    }
    const double acck = acc_s[k][lane];
    double part = 0.0, gkk = 0.0, colsum = 0.0;
#pragma unroll
    for (int c = 0; c < KP; ++c) {
      if (c >= 8 * ta) {                       // (tiles below ta were not loaded)
        if (c == k) gkk = g[c];
        else if (c > k && c < K) part = fma(g[c], u_s[c][lane], part);
        if (c == K) colsum = g[c];
      }
    }
    const double s = rxk - (acck + part);
    const double b = vb ? gkk + svk : gkk;
    const size_t idx = (size_t)row * K + k;
    const double lam = a.lambda[idx];
    const double tau_k = tau * b;
    const double mu_k = (1.0 / tau_k) * (-lam + tau * s);
    double unew = u_s[k][lane], vv = vb ? a.var[idx] : 0.0;
    if (a.apply) {
      double val = 0.0;
      vv = 0.0;
      if (a.mode == MODE_GIBBS) {
        Philox rng(a.seed, it * 16ull + a.salt, (unsigned long long)(a.row_offset + row) * K + k);
        val = tn_draw(mu_k, tau_k, rng);
      } else if (vb) {
        tn_moments(mu_k, tau_k, val, vv);
      } else {
        val = (mu_k != mu_k) ? mu_k : fmax(mu_k, 0.0);   // numpy.maximum propagates NaN (nmf_icm.py:129)
        val = (val != val) ? val : fmax(val, a.min_tn);
      }
      unew = val;
      u_s[k][lane] = val;
      a.fac[idx] = val;
      if (vb) a.var[idx] = vv;
    }
    if (a.mu) a.mu[idx] = mu_k;
    if (a.tauf) a.tauf[idx] = tau_k;
    if (a.sterm) a.sterm[idx] = s;
#pragma unroll
    for (int c = 0; c < KP; ++c)
      if (c > k && c < K && c >= 8 * ta) acc_s[c][lane] = fma(g[c], unew, acc_s[c][lane]);
    rp = fma(unew, rxk, rp);
    sp = fma(unew, colsum, sp);
    pp = fma(unew, fma(gkk, unew, 2.0 * acck), pp);
    ex += vv * (gkk + svk) + unew * unew * svk;
  }
  if (a.mstat) *reinterpret_cast<double4*>(a.mstat + (size_t)row * 4) = make_double4(rp, pp, sp, 0.0);
  if (a.extra) a.extra[row] = ex;
  }
  // fused exchange: the warp's 32 finished rows are contiguous (32 K doubles); store them into every peer's replicated
  // copy with fully coalesced 256-byte stores over NVLink (values re-read from this rank's own copy, written above by
  // the lanes of this warp)
  if (a.peer_fac) {
    __syncwarp();
    const int nrow = min(32, a.rows - row0);
    const size_t l0 = (size_t)row0 * K, g0 = (size_t)(a.row_offset + row0) * K;
    for (int r = 0; r < a.n_peers; ++r) {
      if (r == a.my_rank) continue;
      double* pf = a.peer_fac[r] + g0;
      for (int e = lane; e < nrow * K; e += 32) pf[e] = __ldcg(a.fac + l0 + e);
      if (a.peer_var) {
        double* pv = a.peer_var[r] + g0;
        for (int e = lane; e < nrow * K; e += 32) pv[e] = __ldcg(a.var + l0 + e);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// k_bnmf_row_solve_sub<NT, W>: the same update with W lanes per row (W = 1, 2, 4, ..., 32), natural column order.
//
// The phase is bound by the serial chain of a row's K updates, so the only useful parallelism is ACROSS rows: W is
// chosen by the launcher so that the rows of this shard make about one wave of warps on the chip (65536 rows: W = 1,
// a thread per row; 8192 rows on an 8-way shard: W = 8; toy sizes: W = 32, a warp per row).  Fewer rows then mean
// more lanes per row, fewer columns per lane and a shorter chain: the solver keeps scaling when the matrix is sharded,
// which neither fixed mapping does (warp per row: a wave per ~6500 rows; thread per row: a 0.2 ms floor).
//
// Lane j of a row's group owns columns c = j, j+W, j+2W, ...: their u_c and acc_c live in registers, statically
// indexed because the k loop is fully unrolled (the thread-per-row kernel above spends most of its ~1300 instructions
// per column on predicated loops over the padded columns and on address arithmetic; here both are resolved at compile
// time).  Per column k: every lane loads its columns c >= k of row k of the packed Gram tiles, the group reduces
//     dot_k = acc[k] + sum_{c>k} G[k][c] u_c(old)        (acc[k] = sum_{c<k} G[c][k] u_c(new), pushed earlier)
// with log2 W shuffles, all W lanes evaluate the truncated-normal update redundantly (same inputs, same result), the
// owner of column k keeps u_k, and every lane pushes acc_c += G[k][c] u_k(new) for its columns c > k.  W = 1 performs
// exactly the operations of the thread-per-row kernel in the same order.
// ---------------------------------------------------------------------------------------------------
__device__ __noinline__ void row_update_value(int mode, double mu_k, double tau_k, double min_tn, unsigned long long seed,
                                              unsigned long long stream, unsigned long long index, double& val, double& vv) {
  vv = 0.0;
  if (mode == MODE_GIBBS) {
    Philox rng(seed, stream, index);
    val = tn_draw(mu_k, tau_k, rng);
  } else if (mode == MODE_VB) {
    tn_moments(mu_k, tau_k, val, vv);
  } else {
    val = (mu_k != mu_k) ? mu_k : fmax(mu_k, 0.0);   // numpy.maximum propagates NaN (nmf_icm.py:129)
    val = (val != val) ? val : fmax(val, min_tn);
  }
}

template <int NT, int W>
__global__ void __launch_bounds__(32, 14) k_bnmf_row_solve_sub(RowSolveArgs a) {
  // observed_flag: the dynamic-range guard made the gated fp64 kernels recompute the statistics over the OBSERVED set
  const int pol = a.polarity | (a.observed_flag ? *a.observed_flag : 0);
  constexpr int KP = 8 * NT;
  constexpr int NTP = NT * (NT + 1) / 2;
  constexpr int CPL = (KP + W - 1) / W;       // columns per lane
  constexpr int RPW = 32 / W;                 // rows per warp
  const int lane = threadIdx.x;
  const int j = lane % W, g = lane / W;
  const int row0 = blockIdx.x * RPW;
  const bool live = row0 + g < a.rows;
  const int row = live ? row0 + g : a.rows - 1;   // groups past the end shadow the last row and store nothing
  const int K = a.K;
  const bool vb = a.mode == MODE_VB;
  double u[CPL], acc[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = j + i * W;
    u[i] = c < K ? a.fac[(size_t)row * K + c] : 0.0;
    acc[i] = 0.0;
  }
  const double tau = a.scalars[S_TAU];
  const unsigned long long it = a.iter ? *a.iter : 0ull;
  const size_t gstride = (size_t)a.rows * (NTP * 64);
  const double* grow = a.Gpart + (size_t)row * (NTP * 64);
  const double* rxrow = a.RXpart + (size_t)row * KP;
  const size_t rxstride = (size_t)a.rows * KP;
  double rp = 0.0, pp = 0.0, sp = 0.0, ex = 0.0;
  {
    const char* l0 = reinterpret_cast<const char*>(a.lambda + (size_t)row * K);
    for (int off = j * 128; off < K * 8; off += W * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(l0 + off));
    for (int sgm = 0; sgm < a.nseg_rx; ++sgm)
      for (int off = j * 128; off < KP * 8; off += W * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(rxrow + sgm * rxstride) + off));
  }

  // column k = ik * W + jk: register slot ik (unrolled: every register index below is a compile-time constant),
  // owner lane jk (a run-time loop: it only enters comparisons with j and the shuffle source)
#pragma unroll
  for (int ik = 0; ik < CPL; ++ik) {
    for (int jk = 0; jk < W; ++jk) {
      const int k = ik * W + jk;
      if (k >= K) break;
      const int ta = k >> 3, r = k & 7;
      if (k + 1 < K) {                         // next row of the Gram tiles -> L2
        const int tn = (k + 1) >> 3, rn = (k + 1) & 7;
        for (int sgm = 0; sgm < a.nseg_g; ++sgm)
          for (int tb = tn + j; tb < NT; tb += W)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(grow + sgm * gstride + tile_pair(tn, tb, NT) * 64 + rn * 8));
      }
      // this lane's columns c >= k of row k: the diagonal, the columns still to be updated, and the column-sum slot K
      double gv[CPL];
      if constexpr (W == 1) {
        // a thread per row owns every column: whole tile rows, two 32-byte loads each (one sector per lane and load)
#pragma unroll
        for (int tb = 0; tb < NT; ++tb) {
          if (8 * tb + 7 >= ik) {                            // (ik == k here: compile-time)
            const int off = tile_pair(ta, tb, NT) * 64 + r * 8;
            double4 lo = make_double4(0, 0, 0, 0), hi = make_double4(0, 0, 0, 0);
            for (int sgm = 0; sgm < a.nseg_g; ++sgm) {
              const double4 x = *reinterpret_cast<const double4*>(grow + sgm * gstride + off);
              const double4 y = *reinterpret_cast<const double4*>(grow + sgm * gstride + off + 4);
              lo.x += x.x; lo.y += x.y; lo.z += x.z; lo.w += x.w;
              hi.x += y.x; hi.y += y.y; hi.z += y.z; hi.w += y.w;
            }
            if (!pol) {
              const double4 x = *reinterpret_cast<const double4*>(a.Gfull + off);
              const double4 y = *reinterpret_cast<const double4*>(a.Gfull + off + 4);
              lo.x = x.x - lo.x; lo.y = x.y - lo.y; lo.z = x.z - lo.z; lo.w = x.w - lo.w;
              hi.x = y.x - hi.x; hi.y = y.y - hi.y; hi.z = y.z - hi.z; hi.w = y.w - hi.w;
            }
            gv[8 * tb + 0] = lo.x; gv[8 * tb + 1] = lo.y; gv[8 * tb + 2] = lo.z; gv[8 * tb + 3] = lo.w;
            gv[8 * tb + 4] = hi.x; gv[8 * tb + 5] = hi.y; gv[8 * tb + 6] = hi.z; gv[8 * tb + 7] = hi.w;
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) gv[8 * tb + q] = 0.0;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          const int c = j + i * W;
          gv[i] = 0.0;
          if (i >= ik && c >= k && c <= K) {                 // (first test: compile-time)
            const int off = tile_pair(ta, c >> 3, NT) * 64 + r * 8 + (c & 7);
            double t = 0.0;
            for (int sgm = 0; sgm < a.nseg_g; ++sgm) t += grow[sgm * gstride + off];
            gv[i] = pol ? t : a.Gfull[off] - t;
          }
        }
      }
      double rxk = 0.0, svk = 0.0;
      for (int sgm = 0; sgm < a.nseg_rx; ++sgm) rxk += rxrow[sgm * rxstride + k];
      if (vb) {
        double t = 0.0;
        for (int sgm = 0; sgm < a.nseg_g; ++sgm) t += a.SVpart[((size_t)sgm * a.rows + row) * KP + k];
        svk = pol ? t : a.Gfull[NTP * 64 + k] - t;
      }
      const size_t idx = (size_t)row * K + k;
      const double lam = a.lambda[idx];
      double part = 0.0;
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const int c = j + i * W;
        if (i >= ik && c > k && c < K) part = fma(gv[i], u[i], part);
      }
      double gkk = gv[ik], acck = acc[ik], uold = u[ik], colsum = 0.0;
      if (W > 1) {
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o, W);
        gkk = __shfl_sync(0xffffffffu, gkk, jk, W);
        acck = __shfl_sync(0xffffffffu, acck, jk, W);
        uold = __shfl_sync(0xffffffffu, uold, jk, W);
      }
      if (a.mstat) {                           // slot (k, K): the masked column sum of the other factor's column k
        const int off = tile_pair(ta, K >> 3, NT) * 64 + r * 8 + (K & 7);
        double t = 0.0;
        for (int sgm = 0; sgm < a.nseg_g; ++sgm) t += grow[sgm * gstride + off];
        colsum = pol ? t : a.Gfull[off] - t;
      }
      const double s = rxk - (acck + part);
      const double b = vb ? gkk + svk : gkk;
      const double tau_k = tau * b;
      const double mu_k = (1.0 / tau_k) * (-lam + tau * s);
      double unew = uold, vv = 0.0;
      if (a.apply) {
        row_update_value(a.mode, mu_k, tau_k, a.min_tn, a.seed, it * 16ull + a.salt,
                         (unsigned long long)(a.row_offset + row) * K + k, unew, vv);
        if (j == jk) u[ik] = unew;
      } else if (vb) {
        vv = a.var[idx];
      }
      if (live && j == jk) {
        if (a.apply) { a.fac[idx] = unew; if (vb) a.var[idx] = vv; }
        if (a.mu) a.mu[idx] = mu_k;
        if (a.tauf) a.tauf[idx] = tau_k;
        if (a.sterm) a.sterm[idx] = s;
      }
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const int c = j + i * W;
        if (i >= ik && c > k && c < K) acc[i] = fma(gv[i], unew, acc[i]);
      }
      rp = fma(unew, rxk, rp);
      sp = fma(unew, colsum, sp);
      pp = fma(unew, fma(gkk, unew, 2.0 * acck), pp);
      ex += vv * (gkk + svk) + unew * unew * svk;
    }
  }
  if (live && j == 0) {
    if (a.mstat) *reinterpret_cast<double4*>(a.mstat + (size_t)row * 4) = make_double4(rp, pp, sp, 0.0);
    if (a.extra) a.extra[row] = ex;
  }
  // fused exchange (see k_bnmf_row_solve_lane): the warp's RPW finished rows are contiguous
  if (a.peer_fac) {
    __syncwarp();
    const int nrow = min(RPW, a.rows - row0);
    const size_t l0 = (size_t)row0 * K, g0 = (size_t)(a.row_offset + row0) * K;
    for (int r = 0; r < a.n_peers; ++r) {
      if (r == a.my_rank) continue;
      double* pf = a.peer_fac[r] + g0;
      for (int e = lane; e < nrow * K; e += 32) pf[e] = __ldcg(a.fac + l0 + e);
      if (a.peer_var) {
        double* pv = a.peer_var[r] + g0;
        for (int e = lane; e < nrow * K; e += 32) pv[e] = __ldcg(a.var + l0 + e);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// k_masked_metrics: partial sums over the set bits of `bits` of  e^2, p, p^2 (and, FULL, r*p, r, r^2, 1)  with
// p = A_i . B_j (predict / predict_while_running, bnmf_gibbs_optimised.py:199-223).  The prediction tile is a
// DMMA product of padded factor rows; a warp owns 16 rows and walks 32 columns at a time (4 column tiles
// permuted so that each lane reads 64 contiguous bytes of R).  The B chunk is staged k-major in shared memory
// (row stride = 4 mod 16 doubles -> conflict-free fragment reads); R is register double-buffered.
// For the training mask the r-only sums are static, and sum r*p = (sum r^2 + sum p^2 - sum e^2)/2.
// ---------------------------------------------------------------------------------------------------
template <int KS, bool FULL>  // KS = ceil(K/4) k-steps
__global__ void __launch_bounds__(256) k_masked_metrics(const double* __restrict__ R, const uint32_t* __restrict__ bits,
                                                       int rows, int ld, const double* __restrict__ Ap,
                                                       const double* __restrict__ Bp, int K, int KP, int seg_cols,
                                                       double* __restrict__ partials, const int* __restrict__ run_flag) {
  if (run_flag && *run_flag == 0) return;      // gated fallback of the statistics-based metrics (uniform exit)
  constexpr int CHM = KS <= 8 ? 128 : 64;
  constexpr int CS = CHM + 4;
  __shared__ double bs[4 * KS * CS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * 128 + warp * 16;
  const int c_begin = blockIdx.y * seg_cols, c_end = min(ld, c_begin + seg_cols);
  const int wpr = ld >> 5;
  double sums[7] = {0, 0, 0, 0, 0, 0, 0};
  double af[2][KS];
  const double* rp[2];
  const uint32_t* mp[2];
  bool live[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row0 + 8 * r + g;
    const int rc = min(row, rows - 1);
    live[r] = row < rows;
    rp[r] = R + (size_t)rc * ld + 8 * t;
    mp[r] = bits + (size_t)rc * wpr;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int k = 4 * ks + t;
      af[r][ks] = (k < K) ? Ap[(size_t)rc * KP + k] : 0.0;
    }
  }
  double2 nxt[2][4];
  if (c_begin < c_end) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int uu = 0; uu < 4; ++uu) nxt[r][uu] = *reinterpret_cast<const double2*>(rp[r] + c_begin + 2 * uu);
  }
  for (int c0 = c_begin; c0 < c_end; c0 += CHM) {
    const int ncol = min(CHM, c_end - c0);   // multiple of 32
    __syncthreads();
    for (int i = threadIdx.x; i < ncol * 4 * KS; i += 256) {
      const int col = i / (4 * KS), k = i - col * (4 * KS);
      const int rho = (col & ~31) + 8 * ((col >> 1) & 3) + 2 * ((col >> 3) & 3) + (col & 1);
      bs[k * CS + rho] = (k < K) ? Bp[(size_t)(c0 + col) * KP + k] : 0.0;
    }
    __syncthreads();
    for (int q = 0; q < (ncol >> 5); ++q) {
      const int cq = c0 + 32 * q;
      double2 cur[2][4];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) cur[r][uu] = nxt[r][uu];
      if (cq + 32 < c_end) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int uu = 0; uu < 4; ++uu) nxt[r][uu] = *reinterpret_cast<const double2*>(rp[r] + cq + 32 + 2 * uu);
      }
      uint32_t mw[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) mw[r] = live[r] ? mp[r][cq >> 5] : 0u;
#pragma unroll
      for (int uu = 0; uu < 4; ++uu) {
        // B fragment column g of tile uu <-> actual column cq + 8*(g>>1) + 2*uu + (g&1) = staged slot 32q + 8uu + g
        const double* br = bs + t * CS + 32 * q + 8 * uu + g;
        double p[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const double bf = br[4 * ks * CS];
          dmma884(p[0][0], p[0][1], af[0][ks], bf);
          dmma884(p[1][0], p[1][1], af[1][ks], bf);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint32_t two = (mw[r] >> (8 * t + 2 * uu)) & 3u;
          const double rr[2] = {cur[r][uu].x, cur[r][uu].y};
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const bool on = (two >> i) & 1u;
            const double pm = on ? p[r][i] : 0.0;
            const double em = on ? rr[i] - p[r][i] : 0.0;
            sums[0] = fma(em, em, sums[0]);
            sums[1] += pm;
            sums[2] = fma(pm, pm, sums[2]);
            if (FULL) {
              const double rm = on ? rr[i] : 0.0;
              sums[3] = fma(rm, pm, sums[3]);
              sums[4] += rm;
              sums[5] = fma(rm, rm, sums[5]);
              sums[6] += on ? 1.0 : 0.0;
            }
          }
        }
      }
    }
  }
  __shared__ double red[8][8];
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const double v = warp_sum(sums[i]);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double v = 0.0;
    if (threadIdx.x < 7) for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    partials[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + threadIdx.x] = v;
  }
}

// lean mode: complete the 7-vector from the static sums of the training mask {sum r, sum r^2, |Omega|}
__global__ void k_fill_static_sums(double* __restrict__ out8, const double* __restrict__ statics3,
                                   const int* __restrict__ run_flag) {
  if (run_flag && *run_flag == 0) return;
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const double e2 = out8[0], p2 = out8[2], r = statics3[0], r2 = statics3[1], n = statics3[2];
    out8[3] = 0.5 * (r2 + p2 - e2);
    out8[4] = r; out8[5] = r2; out8[6] = n;
  }
}

// ---------------------------------------------------------------------------------------------------
// training metrics without a pass over R: per-row sums from the solver (rows x 4) -> 64 block partials -> out4
// = {sum r p, sum p^2, sum p, 0} (fixed summation order).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mstat_partial(const double* __restrict__ mstat, int rows, double* __restrict__ partial) {
  double s[3] = {0, 0, 0};
  for (int i = blockIdx.x * 256 + threadIdx.x; i < rows; i += gridDim.x * 256) {
    const double4 v = *reinterpret_cast<const double4*>(mstat + (size_t)i * 4);
    s[0] += v.x; s[1] += v.y; s[2] += v.z;
  }
  __shared__ double red[8][4];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double v = warp_sum(s[i]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    if (threadIdx.x < 3) for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    partial[(size_t)blockIdx.x * 4 + threadIdx.x] = v;
  }
}
__global__ void k_mstat_final(const double* __restrict__ partial, int nparts, double* __restrict__ out4) {
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int p = 0; p < nparts; ++p) v += partial[(size_t)p * 4 + threadIdx.x];
    out4[threadIdx.x] = v;
  }
}
// sums4 = global {sum r p, sum p^2, sum p}; statics3 = global {sum r, sum r^2, |Omega|}  ->  the 7 metric sums, with
// sum e^2 = sum r^2 - 2 sum r p + sum p^2.  The subtraction loses log10(sum r^2 / sum e^2) digits, so when the fit
// is closer than `guard` (relative), *direct_flag is raised and the caller's direct pass over R replaces the result.
__global__ void k_metrics_from_sums(const double* __restrict__ sums4, const double* __restrict__ statics3, double guard,
                                    double* __restrict__ out8, int* __restrict__ direct_flag) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double rp = sums4[0], pp = sums4[1], sp = sums4[2], r = statics3[0], r2 = statics3[1], n = statics3[2];
  const double e2 = (r2 - rp) + (pp - rp);
  out8[0] = e2; out8[1] = sp; out8[2] = pp; out8[3] = rp; out8[4] = r; out8[5] = r2; out8[6] = n; out8[7] = 0.0;
  *direct_flag = !(e2 > guard * r2) ? 1 : 0;            // also raised for NaN
}
// if *flag: out8 = direct8 (global sums of the direct pass)
__global__ void k_select_metrics(const int* __restrict__ flag, const double* __restrict__ direct8, double* __restrict__ out8) {
  if (*flag && threadIdx.x < 8) out8[threadIdx.x] = direct8[threadIdx.x];
}

// deterministic sum of `n` partial vectors of width 8 -> out[8]
__global__ void k_reduce8(const double* __restrict__ partials, int n, double* __restrict__ out,
                          const int* __restrict__ run_flag = nullptr) {
  if (run_flag && *run_flag == 0) return;
  __shared__ double red[256][8];
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += 256)
#pragma unroll
    for (int c = 0; c < 8; ++c) s[c] += partials[(size_t)i * 8 + c];
#pragma unroll
  for (int c = 0; c < 8; ++c) red[threadIdx.x][c] = s[c];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
#pragma unroll
      for (int c = 0; c < 8; ++c) red[threadIdx.x][c] += red[threadIdx.x + o][c];
    __syncthreads();
  }
  if (threadIdx.x < 8) out[threadIdx.x] = red[0][threadIdx.x];
}

// generic deterministic sum of a vector
__global__ void k_reduce1(const double* __restrict__ x, long long n, double* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += 256) s += x[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

// sums over entries with M != 0 of {e^2, p, p^2, r p, r, r^2, 1} for dense host-layout matrices
// (compute_MSE / compute_R2 / compute_Rp helpers, bnmf_gibbs_optimised.py:208-223)
__global__ void __launch_bounds__(256) k_dense_metrics(const double* __restrict__ R, const double* __restrict__ P,
                                                      const double* __restrict__ M, long long n,
                                                      double* __restrict__ partials) {
  double s[7] = {0, 0, 0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const double m = M[i];
    if (m != 0.0) {
      const double r = R[i], p = P[i], e = r - p;
      s[0] += m * e * e; s[1] += m * p; s[2] += m * p * p; s[3] += m * r * p; s[4] += m * r; s[5] += m * r * r; s[6] += m;
    }
  }
  __shared__ double red[8][8];
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const double v = warp_sum(s[i]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double v = 0.0;
    if (threadIdx.x < 7) for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 8 + threadIdx.x] = v;
  }
}

// ---------------------------------------------------------------------------------------------------
// VB: factor-side ELBO terms of one factor matrix (bnmf_vb_optimised.py:166-167,172-177), as block partials:
//   [0] sum log(lambda) - lambda*exp   [1] -0.5 sum log tau + sum log(0.5 erfc(-mu sqrt(tau)/sqrt2)) + sum tau/2 (var + (exp-mu)^2)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vb_factor_terms(const double* __restrict__ ex, const double* __restrict__ var,
                                                        const double* __restrict__ mu, const double* __restrict__ tauf,
                                                        const double* __restrict__ lambda, long long n,
                                                        double* __restrict__ partials) {
  double s0 = 0.0, s1 = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const double l = lambda[i], e = ex[i], v = var[i], m = mu[i], t = tauf[i];
    s0 += log(l) - l * e;
    const double d = e - m;
    s1 += -0.5 * log(t) + log(0.5 * erfc_ref(-m * sqrt(t) / kSqrt2)) + t * 0.5 * (v + d * d);
  }
  __shared__ double red[8][2];
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = s0; red[threadIdx.x >> 5][1] = s1; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 8 + threadIdx.x] = s;
  }
  if (threadIdx.x >= 2 && threadIdx.x < 8) partials[(size_t)blockIdx.x * 8 + threadIdx.x] = 0.0;
}

// ---------------------------------------------------------------------------------------------------
// k_bnmf_finish: end of a sweep.  m8 = reduced metric sums over the training mask; ex1 = reduced extra term
// (VB), el8 = reduced factor ELBO terms (VB, U then V added).  Updates tau (Gibbs: Gamma draw,
// bnmf_gibbs_optimised.py:144,161-165; VB: :181-183,213-215; ICM: nmf_icm.py:137) and appends one trace row.
// ---------------------------------------------------------------------------------------------------
__global__ void k_bnmf_finish(FinishArgs a) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  finish_sweep(a);
}

// ---------------------------------------------------------------------------------------------------
// elementwise distribution kernels (code/models/distributions/*.py)
// ---------------------------------------------------------------------------------------------------
__global__ void k_tn_moments(const double* __restrict__ mu, const double* __restrict__ tau, long long n,
                             double* __restrict__ ex, double* __restrict__ var) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double e, v;
  tn_moments(mu[i], tau[i], e, v);
  if (ex) ex[i] = e;
  if (var) var[i] = v;
}

__global__ void k_tn_draw(const double* __restrict__ mu, const double* __restrict__ tau, long long n,
                          unsigned long long seed, unsigned long long stream, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Philox rng(seed, stream, (unsigned long long)i);
  out[i] = tn_draw(mu[i], tau[i], rng);
}

__global__ void k_gamma_draw(double shape, double rate, long long n, unsigned long long seed, unsigned long long stream,
                             double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Philox rng(seed, stream, (unsigned long long)i);
  out[i] = gamma_draw_dev(shape, rate, rng);
}

__global__ void k_exponential_draw(const double* __restrict__ lambda, long long n, unsigned long long seed,
                                   unsigned long long stream, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Philox rng(seed, stream, (unsigned long long)i);
  double u0, u1; rng.uniform2(u0, u1);
  out[i] = -log(u0) / lambda[i];
}

// ---------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------
int launch_pad_factor(const double* X, const double* Var, int n, int K, int n_alloc, double* Xp, double* Vp,
                      cudaStream_t st) {
  const int KP = 8 * tiles_for(K);
  const size_t total = (size_t)n_alloc * KP;
  k_pad_factor<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(X, Var, n, K, KP, n_alloc, Xp, Vp);
  return check_launch("pad_factor");
}

// Which solver, for the natural column order (K <= 31):
//   rows >= 24576        k_bnmf_row_solve_lane, a thread per row (one wave of warps from 65536 rows; 0.2 ms floor)
//   fewer rows           k_bnmf_row_solve_sub with W = 4, 8, 16 or 32 lanes per row, the smallest W that turns the rows
//                        of this launch into about one wave (>= 1536 warps): 16384 rows -> 4, 8192 -> 8, <= 2048 -> 32
// (W = 1, 2 of the sub-warp kernel would be the thread-per-row case again, but 24-32 fully unrolled column bodies per
// instantiation spill at 128 registers and take minutes to compile, so the older kernel keeps that range.)
// Explicit column orders, the white-box single-column calls and K > 31 use the warp-per-row kernel k_bnmf_row_solve.
// Same Philox counters everywhere, so the choice changes a Gibbs chain only by the rounding of the K-term dot
// products (tests: 1e-8 after 4 sweeps).  BNMTF_SOLVE = warp | lane | sub<W> forces a kernel (tests compare them).
template <int NT>
static int launch_row_solve_sub(const RowSolveArgs& a, int W, cudaStream_t st) {
  const int rpw = 32 / W, grid = (a.rows + rpw - 1) / rpw;
  switch (W) {
#define BNMTF_SUB(WW) case WW: k_bnmf_row_solve_sub<NT, WW><<<grid, 32, 0, st>>>(a); break;
    BNMTF_SUB(4) BNMTF_SUB(8) BNMTF_SUB(16) BNMTF_SUB(32)
#undef BNMTF_SUB
    default: set_error("row_solve: W=%d is not one of 4, 8, 16, 32", W); return -2;
  }
  return check_launch("row_solve_sub");
}

int launch_row_solve(const RowSolveArgs& a, cudaStream_t st) {
  const int nt = tiles_for(a.K);
  const char* pref = getenv("BNMTF_SOLVE");
  const bool natural = a.order == nullptr && a.n_order == a.K;
  int W = 0;                                             // 0: not the sub-warp kernel
  bool lane = natural && pref && pref[0] == 'l';
  if (natural && nt <= 4 && !(pref && (pref[0] == 'w' || pref[0] == 'l'))) {
    if (pref && pref[0] == 's' && pref[1] == 'u' && pref[2] == 'b') W = atoi(pref + 3);
    else if (a.rows >= 24576) lane = true;
    else { W = 4; while (W < 32 && (long long)a.rows * W < 1536ll * 32) W *= 2; }
  }
  if (W > 0) {
    switch (nt) {
      case 1: return launch_row_solve_sub<1>(a, W, st);
      case 2: return launch_row_solve_sub<2>(a, W, st);
      case 3: return launch_row_solve_sub<3>(a, W, st);
      default: return launch_row_solve_sub<4>(a, W, st);
    }
  }
  if (lane) {
    const int grid = (a.rows + 31) / 32;
    switch (nt) {
#define BNMTF_RL(N) case N: k_bnmf_row_solve_lane<N><<<grid, 32, 0, st>>>(a); break;
      BNMTF_RL(1) BNMTF_RL(2) BNMTF_RL(3) BNMTF_RL(4) BNMTF_RL(5) BNMTF_RL(6) BNMTF_RL(7) BNMTF_RL(8)
#undef BNMTF_RL
      default: set_error("row_solve: K=%d out of range", a.K); return -2;
    }
    return check_launch("row_solve_lane");
  }
  const int KP = 8 * nt;
  const size_t per_warp = (size_t)KP * (KP + 1) * sizeof(double);
  int warps = (int)(40 * 1024 / per_warp);
  if (warps > 4) warps = 4;
  if (warps < 1) warps = 1;
  const size_t smem = per_warp * warps;
  const int grid = (a.rows + warps - 1) / warps;
  switch (nt) {
#define BNMTF_RS(N) case N: k_bnmf_row_solve<N><<<grid, warps * 32, smem, st>>>(a); break;
    BNMTF_RS(1) BNMTF_RS(2) BNMTF_RS(3) BNMTF_RS(4) BNMTF_RS(5) BNMTF_RS(6) BNMTF_RS(7) BNMTF_RS(8)
#undef BNMTF_RS
    default: set_error("row_solve: K=%d out of range", a.K); return -2;
  }
  return check_launch("row_solve");
}

// partials must hold (ceil(rows/128) * nseg) * 8 doubles; out8 receives the reduced sums.  statics3 != NULL selects
// the lean kernel (training mask: {sum r, sum r^2, |Omega|} known).
int launch_masked_metrics(const double* R, const uint32_t* bits, int rows, int ld, const double* Ap, const double* Bp,
                          int K, int nseg, const double* statics3, double* partials, double* out8, const int* run_flag,
                          cudaStream_t st) {
  const int KP = 8 * tiles_for(K);
  const int ks = (K + 3) / 4;
  const int chm = ks <= 8 ? 128 : 64;
  const int seg_cols = round_up((ld + nseg - 1) / nseg, chm);
  dim3 grid((rows + 127) / 128, nseg);
  switch (ks) {
#define BNMTF_MM(N) case N: if (statics3) k_masked_metrics<N, false><<<grid, 256, 0, st>>>(R, bits, rows, ld, Ap, Bp, K, KP, seg_cols, partials, run_flag); \
                            else k_masked_metrics<N, true><<<grid, 256, 0, st>>>(R, bits, rows, ld, Ap, Bp, K, KP, seg_cols, partials, run_flag); break;
    BNMTF_MM(1) BNMTF_MM(2) BNMTF_MM(3) BNMTF_MM(4) BNMTF_MM(5) BNMTF_MM(6) BNMTF_MM(7) BNMTF_MM(8)
    BNMTF_MM(9) BNMTF_MM(10) BNMTF_MM(11) BNMTF_MM(12) BNMTF_MM(13) BNMTF_MM(14) BNMTF_MM(15) BNMTF_MM(16)
#undef BNMTF_MM
    default: set_error("masked_metrics: K=%d out of range", K); return -2;
  }
  k_reduce8<<<1, 256, 0, st>>>(partials, (int)(grid.x * grid.y), out8, run_flag);
  if (statics3) k_fill_static_sums<<<1, 32, 0, st>>>(out8, statics3, run_flag);
  return check_launch("masked_metrics");
}

// mstat: rows x 4 from the row solver -> out4 (this rank's sums); partial: >= 64*4 doubles
int launch_mstat_reduce(const double* mstat, int rows, double* partial, double* out4, cudaStream_t st) {
  int nparts = (rows + 255) / 256;
  if (nparts > 64) nparts = 64;
  if (nparts < 1) nparts = 1;
  k_mstat_partial<<<nparts, 256, 0, st>>>(mstat, rows, partial);
  k_mstat_final<<<1, 32, 0, st>>>(partial, nparts, out4);
  return check_launch("mstat_reduce");
}
int launch_metrics_from_sums(const double* sums4, const double* statics3, double guard, double* out8, int* direct_flag,
                             cudaStream_t st) {
  k_metrics_from_sums<<<1, 32, 0, st>>>(sums4, statics3, guard, out8, direct_flag);
  return check_launch("metrics_from_sums");
}
int launch_select_metrics(const int* flag, const double* direct8, double* out8, cudaStream_t st) {
  k_select_metrics<<<1, 32, 0, st>>>(flag, direct8, out8);
  return check_launch("select_metrics");
}

int launch_vb_factor_terms(const double* ex, const double* var, const double* mu, const double* tauf,
                           const double* lambda, long long n, double* partials, int nblocks, cudaStream_t st) {
  k_vb_factor_terms<<<nblocks, 256, 0, st>>>(ex, var, mu, tauf, lambda, n, partials);
  return check_launch("vb_factor_terms");
}

int launch_dense_metrics(const double* R, const double* P, const double* M, long long n, double* partials, int nblocks,
                         double* out8, cudaStream_t st) {
  k_dense_metrics<<<nblocks, 256, 0, st>>>(R, P, M, n, partials);
  k_reduce8<<<1, 256, 0, st>>>(partials, nblocks, out8);
  return check_launch("dense_metrics");
}

int launch_reduce8(const double* partials, int n, double* out8, cudaStream_t st) {
  k_reduce8<<<1, 256, 0, st>>>(partials, n, out8);
  return check_launch("reduce8");
}

int launch_reduce1(const double* x, long long n, double* out, cudaStream_t st) {
  k_reduce1<<<1, 256, 0, st>>>(x, n, out);
  return check_launch("reduce1");
}

int launch_finish(const FinishArgs& a, cudaStream_t st) {
  k_bnmf_finish<<<1, 32, 0, st>>>(a);
  return check_launch("bnmf_finish");
}

int launch_tn_moments(const double* mu, const double* tau, long long n, double* ex, double* var, cudaStream_t st) {
  if (n <= 0) return 0;
  k_tn_moments<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mu, tau, n, ex, var);
  return check_launch("tn_moments");
}
int launch_tn_draw(const double* mu, const double* tau, long long n, unsigned long long seed, unsigned long long stream,
                   double* out, cudaStream_t st) {
  if (n <= 0) return 0;
  k_tn_draw<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mu, tau, n, seed, stream, out);
  return check_launch("tn_draw");
}
int launch_gamma_draw(double shape, double rate, long long n, unsigned long long seed, unsigned long long stream,
                      double* out, cudaStream_t st) {
  if (n <= 0) return 0;
  k_gamma_draw<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(shape, rate, n, seed, stream, out);
  return check_launch("gamma_draw");
}
int launch_exponential_draw(const double* lambda, long long n, unsigned long long seed, unsigned long long stream,
                            double* out, cudaStream_t st) {
  if (n <= 0) return 0;
  k_exponential_draw<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(lambda, n, seed, stream, out);
  return check_launch("exponential_draw");
}

}  // namespace bnmtf

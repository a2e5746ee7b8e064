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
// k_nmtf_sq_tiled: the same reduction as a register-tiled product.  H[(k,l),(k',l')] = sum_i (F_ik F_ik') GG_i[l,l'] is
// C = A^T B over the rows of R with A_i = vec(F_i F_i^T) (K^2 entries) and B_i = vec(GG_i) (L^2 entries); for VB A_i
// gets K more entries (varF_i) and B_i L more (the observed variance sums sv_i), which yields the three covariance
// blocks and the precision terms from the same product:
//   C[(k,k'), N + l] = sum f_k f_k' sv_l   (cov_term_G)      C[M + k, (l,l')] = sum varF_k GG[l,l']   (cov_term_F)
//   C[M + k, N + l]  = sum varF_k sv_l     (precision only)
// 512 threads; `rb` rows are staged in shared memory per step (cp.async copies of the raw statistics, issued one step ahead
// so that they land during the products, then converted in shared memory), every warp owns SQ_NT blocks of 16 x 32
// entries of C in registers and forms them with fp64 DMMA (m8n8k4, the staged rows being the contraction index: one
// operand register feeds 8 FMAs, so the products run at the fp64 pipe's rate instead of the shared-memory rate).
// grid = (row partitions, passes): a CTA covers a PM x PN rectangle of the WTM x WTN warp
// blocks of C (pass = blockIdx.y; one pass up to K = L = 10), stages only that rectangle's columns of A and B of its
// share of the rows, and writes the rectangle into its row partition's slice of `partial` (Mp x Np, then the K*L
// right-hand sides); k_sq_assemble adds the slices and folds C into (H, prec, rhs).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
constexpr int SQ_NT = 2, SQ_THREADS = 512, SQ_BLOCKS = (SQ_THREADS / 32) * SQ_NT, SQ_RHS = 2, SQ_MAX_PASSES = 64, SQ_PAD = 4;

struct SqTiling { int Mext, Next, Mp, Np, WTM, WTN, PM, PN, npm, npn, npass, rb; size_t smem, plen; bool ok; };
__host__ __device__ inline SqTiling sq_tiling(int K, int L, int vb) {
  SqTiling t;
  t.Mext = K * K + (vb ? K : 0); t.Next = L * L + (vb ? L : 0);
  t.Mp = (t.Mext + 3) & ~3; t.Np = (t.Next + 3) & ~3;
  t.WTM = (t.Mp / 4 + 3) / 4; t.WTN = (t.Np / 4 + 7) / 8;     // warp blocks of 16 x 32 entries
  // a pass (one CTA) covers a PM x PN rectangle of warp blocks, PM * PN = SQ_BLOCKS: the shape with the fewest passes, then
  // the fewest staged columns (16 PM of A, 32 PN of B per row)
  t.PM = SQ_BLOCKS; t.PN = 1; t.npass = 1 << 30;
  int best_cols = 1 << 30;
  for (int pn = 1; pn <= SQ_BLOCKS; pn <<= 1) {
    const int pm = SQ_BLOCKS / pn;
    const int np = ((t.WTM + pm - 1) / pm) * ((t.WTN + pn - 1) / pn), cols = 16 * pm + 32 * pn;
    if (np < t.npass || (np == t.npass && cols < best_cols)) { t.PM = pm; t.PN = pn; t.npass = np; best_cols = cols; }
  }
  t.npm = (t.WTM + t.PM - 1) / t.PM; t.npn = (t.WTN + t.PN - 1) / t.PN;
  t.plen = ((size_t)t.Mp * t.Np + (size_t)K * L + 1) & ~(size_t)1;       // (even: 16-byte aligned slices)
  const size_t tables = (size_t)t.Next * (sizeof(double) + sizeof(int)) + (size_t)t.Mext * sizeof(int) + 16;
  t.rb = 16; t.smem = 0;
  for (; t.rb >= 4; t.rb >>= 1) {                               // (a multiple of the DMMA depth 4 that divides 16 warps)
    // products side: As | Bs (padded rows) | Fs | RGs;  landing area of the asynchronous copies: Braw | Fraw | VFraw | RGraw
    t.smem = (size_t)t.rb * (16 * t.PM + 2 * 32 * t.PN + 2 * SQ_PAD + 3 * K + 2 * L) * sizeof(double) + tables;
    if (t.smem <= 200 * 1024) break;
  }
  t.ok = t.rb >= 4 && t.npass <= SQ_MAX_PASSES && K * L <= SQ_THREADS * SQ_RHS;
  return t;
}

template <bool VB>
__global__ void __launch_bounds__(SQ_THREADS, 1) k_nmtf_sq_tiled(SqArgs a) {
  extern __shared__ double sm[];
  const int K = a.K, L = a.L, D = K * L, M = K * K, N = L * L;
  const SqTiling t = sq_tiling(K, L, VB ? 1 : 0);
  const int Mp = t.Mp, Np = t.Np, Mext = t.Mext, Next = t.Next, RB = t.rb;
  const int ntl = tiles_for(L), KPl = 8 * ntl, gll = ntl * (ntl + 1) / 2 * 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, pass = blockIdx.y;
  // this pass's rectangle of C: columns [acol0, acol0 + AW) of A and [bcol0, bcol0 + BW) of B, of which an / bn exist
  const int pi = pass / t.npn, pj = pass - pi * t.npn;
  const int AW = 16 * t.PM, BW = 32 * t.PN, acol0 = pi * AW, bcol0 = pj * BW;
  const int an = (acol0 + AW < Mext ? acol0 + AW : Mext) - acol0, bn = (bcol0 + BW < Next ? bcol0 + BW : Next) - bcol0;
  // row strides of the operand arrays: + 4 doubles, so that the four rows a DMMA fragment load touches (4 x 32 bytes per
  // half warp) fall into disjoint banks
  const int AWp = AW + SQ_PAD, BWp = BW + SQ_PAD;
  // shared memory: what the products read (As | Bs | Fs | RGs), the landing area of the asynchronous copies of the next
  // rows (Braw | Fraw | VFraw | RGraw), then the tables
  double* As = sm;
  double* Bs = As + RB * AWp;
  double* Fs = Bs + RB * BWp;
  double* RGs = Fs + RB * K;
  double* Braw = RGs + RB * L;
  double* Fraw = Braw + RB * BW;
  double* VFraw = Fraw + RB * K;
  double* RGraw = VFraw + RB * K;
  double* gfB = RGraw + RB * L;                                 // full-set value of entry n (polarity 0), else 0
  int* idxB = reinterpret_cast<int*>(gfB + Next);               // offset of entry n in a row's Gram tiles / variance sums
  int* kkA = idxB + Next;                                       // k | k' << 16 of entry m
  for (int n = tid; n < Next; n += SQ_THREADS) {
    if (n < N) {
      int la = n / L, lb = n - la * L;
      if (la > lb) { const int x = la; la = lb; lb = x; }
      const int ta = la >> 3, tb = lb >> 3;
      const int idx = (ta * ntl - ta * (ta - 1) / 2 + (tb - ta)) * 64 + (la & 7) * 8 + (lb & 7);
      idxB[n] = idx;
      gfB[n] = a.polarity ? 0.0 : a.Gfull_o[idx];
    } else {
      idxB[n] = n - N;
      gfB[n] = a.polarity ? 0.0 : a.Gfull_o[gll + n - N];
    }
  }
  for (int m = tid; m < Mext; m += SQ_THREADS) kkA[m] = m < M ? ((m / K) | ((m % K) << 16)) : (m - M);
  for (int i = tid; i < RB * (AWp + BWp); i += SQ_THREADS) As[i] = 0.0;        // (padding entries stay zero)
  // staging work is split by rows: 16 / RB warps share a staged row and stride over its columns
  const int wpr = (SQ_THREADS / 32) / RB, srow = warp % RB, scol = (warp / RB) * 32 + lane, sstep = 32 * wpr;
  // products: every warp owns SQ_NT blocks of 16 x 32 entries of the rectangle = 2 x 4 DMMA tiles (8 x 8) each
  const int g = lane >> 2, t4 = lane & 3;
  int aoff[SQ_NT], boff[SQ_NT];
  bool valid[SQ_NT];
#pragma unroll
  for (int q = 0; q < SQ_NT; ++q) {
    const int b = q * (SQ_THREADS / 32) + warp;                // block of the rectangle; offsets are relative to it
    const int pm = b / t.PN, pn = b - pm * t.PN;
    aoff[q] = 16 * pm;
    boff[q] = 32 * pn;
    valid[q] = acol0 + aoff[q] < Mp && bcol0 + boff[q] < Np;   // (warp-uniform)
  }
  double acc[SQ_NT][2][4][2];
#pragma unroll
  for (int q = 0; q < SQ_NT; ++q)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[q][i][j][0] = acc[q][i][j][1] = 0.0;
  // pass 0 also accumulates the right-hand sides: thread -> entries d = tid, tid + 512:  sum_i F_ik RG_il
  const bool want_rhs = pass == 0;
  double rhs[SQ_RHS];
  int dk[SQ_RHS], dl[SQ_RHS];
#pragma unroll
  for (int c = 0; c < SQ_RHS; ++c) {
    const int d = tid + c * SQ_THREADS;
    rhs[c] = 0.0;
    dk[c] = (want_rhs && d < D) ? d / L : -1;
    dl[c] = dk[c] >= 0 ? d - dk[c] * L : 0;
  }
  // asynchronous copies (cp.async, 8 bytes each) of the raw statistics of the rows [row0, row0 + RB) into the landing area
  auto fetch = [&](int row0) {
    const int row = row0 + srow;
    if (row < a.rows) {
      const double* gp = a.Go + (size_t)row * gll;
      const double* sp = a.SVo + (size_t)row * KPl;
      for (int c = scol; c < bn; c += sstep) {
        const int n = bcol0 + c;
        cp_async8(Braw + srow * BW + c, n < N ? gp + idxB[n] : sp + idxB[n]);
      }
      for (int k = scol; k < K; k += sstep) {
        cp_async8(Fraw + srow * K + k, a.F + (size_t)row * K + k);
        if (VB) cp_async8(VFraw + srow * K + k, a.varF + (size_t)row * K + k);
      }
      if (want_rhs)
        for (int l = scol; l < L; l += sstep) cp_async8(RGraw + srow * L + l, a.RXo + (size_t)row * KPl + l);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int nbatch = (a.rows + RB - 1) / RB;
  __syncthreads();                                              // tables ready
  if ((int)blockIdx.x < nbatch) fetch(blockIdx.x * RB);
  for (int bt = blockIdx.x; bt < nbatch; bt += gridDim.x) {
    const int row0 = bt * RB;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                            // copies landed; the previous step's products are done
    // raw -> what the products read: observed-set values of B, the products F_k F_k' (and varF) of A, F and RG for the rhs
    {
      const bool in = row0 + srow < a.rows;
      for (int c = scol; c < bn; c += sstep) {
        const double raw = Braw[srow * BW + c];
        Bs[srow * BWp + c] = in ? (a.polarity ? raw : gfB[bcol0 + c] - raw) : 0.0;
      }
      const double* fr = Fraw + srow * K;
      for (int ca = scol; ca < an; ca += sstep) {
        const int m = acol0 + ca, c = kkA[m];
        const double v = m < M ? fr[c & 0xffff] * fr[c >> 16] : VFraw[srow * K + c];
        As[srow * AWp + ca] = in ? v : 0.0;
      }
      if (want_rhs) {
        for (int k = scol; k < K; k += sstep) Fs[srow * K + k] = in ? fr[k] : 0.0;
        for (int l = scol; l < L; l += sstep) RGs[srow * L + l] = in ? RGraw[srow * L + l] : 0.0;
      }
    }
    __syncthreads();                                            // the landing area is free again
    if (bt + (int)gridDim.x < nbatch) fetch((bt + gridDim.x) * RB);           // in flight during the products
    for (int kk = 0; kk < RB; kk += 4) {
      const double* ar = As + (kk + t4) * AWp + g;              // A fragment: lane holds A^T[m = g][r = t4]
      const double* br = Bs + (kk + t4) * BWp + g;              // B fragment: lane holds B[r = t4][n = g]
#pragma unroll
      for (int q = 0; q < SQ_NT; ++q) {
        if (!valid[q]) continue;
        const double a0 = ar[aoff[q]], a1 = ar[aoff[q] + 8];
        double bf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = br[boff[q] + 8 * j];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dmma884(acc[q][0][j][0], acc[q][0][j][1], a0, bf[j]);
          dmma884(acc[q][1][j][0], acc[q][1][j][1], a1, bf[j]);
        }
      }
    }
    if (want_rhs)
      for (int r = 0; r < RB; ++r)
#pragma unroll
        for (int c = 0; c < SQ_RHS; ++c)
          if (dk[c] >= 0) rhs[c] = fma(Fs[r * K + dk[c]], RGs[r * L + dl[c]], rhs[c]);
  }
  double* C = a.partial + (size_t)blockIdx.x * t.plen;          // this row partition's Mp x Np (+ D) slice
#pragma unroll
  for (int q = 0; q < SQ_NT; ++q)
    if (valid[q])
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int m = acol0 + aoff[q] + 8 * i + g, n = bcol0 + boff[q] + 8 * j + 2 * t4;   // lane holds C[g][2 t4 + {0,1}]
          if (m < Mp && n < Np)
            *reinterpret_cast<double2*>(C + (size_t)m * Np + n) = make_double2(acc[q][i][j][0], acc[q][i][j][1]);
        }
#pragma unroll
  for (int c = 0; c < SQ_RHS; ++c)
    if (dk[c] >= 0) C[(size_t)Mp * Np + tid + c * SQ_THREADS] = rhs[c];
}

// adds the row partitions' slices of C and folds them into (H, prec, rhs): 256 threads = 32 entries x 8 groups of slices
template <bool VB>
__global__ void __launch_bounds__(256) k_sq_assemble(const double* __restrict__ partial, int nparts, int K, int L,
                                                      double* __restrict__ out) {
  __shared__ double red[8][32];
  const SqTiling t = sq_tiling(K, L, VB ? 1 : 0);
  const int D = K * L, M = K * K, N = L * L, Np = t.Np, len = D * D + 2 * D;
  const int e = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + e;
  size_t idx[4];
  int cnt = 0;
  if (i < D * D) {
    const int d = i / D, d2 = i - d * D;
    const int k = d / L, l = d - k * L, k2 = d2 / L, l2 = d2 - k2 * L;
    idx[cnt++] = (size_t)(k * K + k2) * Np + l * L + l2;
    if (VB && l == l2 && k != k2) idx[cnt++] = (size_t)(k * K + k2) * Np + N + l;          // cov_term_G
    if (VB && k == k2 && l != l2) idx[cnt++] = (size_t)(M + k) * Np + l * L + l2;          // cov_term_F
  } else if (i < D * D + D) {
    const int d = i - D * D, k = d / L, l = d - k * L;
    idx[cnt++] = (size_t)(k * K + k) * Np + l * L + l;
    if (VB) {
      idx[cnt++] = (size_t)(k * K + k) * Np + N + l;
      idx[cnt++] = (size_t)(M + k) * Np + l * L + l;
      idx[cnt++] = (size_t)(M + k) * Np + N + l;
    }
  } else if (i < len) {
    idx[cnt++] = (size_t)t.Mp * Np + (i - D * D - D);
  }
  double s = 0.0;
  for (int p = g; p < nparts; p += 8) {
    const double* c = partial + (size_t)p * t.plen;
    for (int q = 0; q < cnt; ++q) s += c[idx[q]];
  }
  red[g][e] = s;
  __syncthreads();
  if (g == 0 && i < len) {
    double v = red[0][e];
#pragma unroll
    for (int w = 1; w < 8; ++w) v += red[w][e];
    out[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------------
// k_coord_solve: sequential scalar updates of x (D entries) on the small system (H, prec, rhs):
//   s = rhs_d - sum_{d' != d} H[d][d'] x_d' ; tau_d = tau * prec_d ; mu_d = 1/tau_d (-lambda_d + tau s)
// Single CTA.  Reproduces tauS/muS + TN_draw / TN moments / TN_mode for every (k,l) in `order`.
// ---------------------------------------------------------------------------------------------------

template <bool STAGED>
__global__ void __launch_bounds__(256) k_coord_solve(CoordArgs a) {
  // the D updates are one dependent chain: the whole CTA stages H (when D*D doubles fit shared memory), then warp 0 runs
  // the chain alone -- warp-synchronous, every lane holds the full dot product and evaluates the same update
  extern __shared__ double xs[];   // x, rhs, prec, lambda: D each  (+ D*D: H)
  const int D = a.D;
  double *rhs_s = xs + D, *prec_s = rhs_s + D, *lam_s = prec_s + D, *Hs = lam_s + D;
  for (int i = threadIdx.x; i < D; i += 256) {
    xs[i] = a.x[i]; rhs_s[i] = a.rhs[i]; prec_s[i] = a.prec[i]; lam_s[i] = a.lambda[i];
  }
  if (STAGED)
    for (int i = threadIdx.x; i < D * D; i += 256) Hs[i] = a.H[i];
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  const double tau = a.scalars[S_TAU];
  const unsigned long long it = a.iter ? *a.iter : 0ull;
  for (int o = 0; o < a.n_order; ++o) {
    const int d = a.order ? a.order[o] : o;
    const double* Hrow = STAGED ? Hs + (size_t)d * D : a.H + (size_t)d * D;
    double part = 0.0;
    for (int i = lane; i < D; i += 32)
      if (i != d) part = fma(Hrow[i], xs[i], part);
    const double dot = warp_sum(part);
    const double s = rhs_s[d] - dot;
    const double tau_d = tau * prec_s[d];
    const double mu_d = (1.0 / tau_d) * (-lam_s[d] + tau * s);
    double val = xs[d], vv = 0.0;
    if (a.apply) {
      if (a.mode == MODE_GIBBS) {
        Philox rng(a.seed, it * 16ull + a.salt, (unsigned long long)d);
        val = tn_draw(mu_d, tau_d, rng);
      } else if (a.mode == MODE_VB) {
        tn_moments(mu_d, tau_d, val, vv);
      } else {
        val = (mu_d != mu_d) ? mu_d : fmax(mu_d, 0.0);
        val = (val != val) ? val : fmax(val, a.min_tn);
      }
    }
    __syncwarp();
    if (lane == 0) {
      if (a.apply) {
        if (a.mode == MODE_VB) a.var[d] = vv;
        a.x[d] = val;
      }
      if (a.mu) a.mu[d] = mu_d;
      if (a.tauf) a.tauf[d] = tau_d;
      xs[d] = val;
    }
    __syncwarp();
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


// ---------------------------------------------------------------------------------------------------
// k_nmtf_mstat: the three masked sums the training metrics need, per column j of R, from the column statistics
// w.r.t. F (c_j = sum_i m r F_i, FF_j = sum_i m F_i F_i^T, s_j = sum_i m F_i: slot (k, K) of the Gram tiles) and the
// current S and G_j.  With y = S G_j:   sum_i m r p = y.c_j,   sum_i m p^2 = y^T FF_j y,   sum_i m p = y.s_j
// (p = F S G^T; predict_while_running / compute_MSE / compute_R2 / compute_Rp of the reference, bnmtf_gibbs_optimised.py:234-258, takes them from a pass over R).
// mstat: rows x 4 (rp, pp, sp, 0) -- the layout k_mstat_partial of solve.cu reduces.
// ---------------------------------------------------------------------------------------------------
__global__ void k_nmtf_mstat(int rows, int K, int L, int polarity, const double* __restrict__ RXo,
                             const double* __restrict__ Go, const double* __restrict__ Gfull_o,
                             const double* __restrict__ G, const double* __restrict__ S, double* __restrict__ mstat) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int row = blockIdx.x * nw + warp;
  if (row >= rows) return;
  const int ntk = tiles_for(K), KPk = 8 * ntk, glk = ntk * (ntk + 1) / 2 * 64;
  double* y = sm + (size_t)warp * K;
  const double* g = G + (size_t)row * L;
  for (int k = lane; k < K; k += 32) {
    double t = 0.0;
    for (int l = 0; l < L; ++l) t = fma(S[k * L + l], g[l], t);
    y[k] = t;
  }
  __syncwarp();
  const double* gp = Go + (size_t)row * glk;
  double rp = 0.0, pp = 0.0, sp = 0.0;
  for (int i = lane; i < K * K; i += 32) {
    const int k = i / K, k2 = i - k * K;
    pp = fma(y[k] * y[k2], gram_at(gp, Gfull_o, polarity, ntk, k, k2), pp);
  }
  for (int k = lane; k < K; k += 32) {
    rp = fma(y[k], RXo[(size_t)row * KPk + k], rp);
    sp = fma(y[k], gram_at(gp, Gfull_o, polarity, ntk, k, K), sp);
  }
  rp = warp_sum(rp); pp = warp_sum(pp); sp = warp_sum(sp);
  if (lane == 0) *reinterpret_cast<double4*>(mstat + (size_t)row * 4) = make_double4(rp, pp, sp, 0.0);
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
// doubles of scratch per row partition of bnmtf_nmtf_sq_f64, and the number of row partitions that fills the device
long long sq_scratch_len(int K, int L, int vb) {
  const SqTiling t = sq_tiling(K, L, vb);
  const long long old_len = (long long)K * L * K * L + 2LL * K * L;
  return (t.ok && (long long)t.plen > old_len) ? (long long)t.plen : old_len;
}
int sq_parts(long long rows, int K, int L, int vb) {
  const SqTiling t = sq_tiling(K, L, vb);
  const int rb = t.ok ? t.rb : 16, passes = t.ok ? t.npass : 1;
  long long p = (rows + rb - 1) / rb;
  const long long cap = 148 / passes > 0 ? 148 / passes : 1;
  if (p > cap) p = cap;
  return p < 1 ? 1 : (int)p;
}

int launch_nmtf_sq(const SqArgs& a, int nparts, double* out, cudaStream_t st);   // defined below
// 256 threads = 32 entries x 8 groups of partials: thread (g, e) adds the partials g, g + 8, ... of its entry, then the eight
// group sums are added in a fixed order
__global__ void __launch_bounds__(256) k_sum_partials2(const double* __restrict__ partial, int nparts, int len,
                                                        double* __restrict__ out) {
  __shared__ double red[8][32];
  const int e = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + e;
  double s = 0.0;
  if (i < len)
#pragma unroll 4
    for (int p = g; p < nparts; p += 8) s += partial[(size_t)p * len + i];
  red[g][e] = s;
  __syncthreads();
  if (g == 0 && i < len) {
    double t = red[0][e];
#pragma unroll
    for (int w = 1; w < 8; ++w) t += red[w][e];
    out[i] = t;
  }
}
int launch_nmtf_sq(const SqArgs& a, int nparts, double* out, cudaStream_t st) {
  const int D = a.K * a.L;
  const SqTiling t = sq_tiling(a.K, a.L, a.vb);
  if (t.ok) {
    const int len = D * D + 2 * D;
    if (a.vb) {
      cudaFuncSetAttribute(k_nmtf_sq_tiled<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      k_nmtf_sq_tiled<true><<<dim3(nparts, t.npass), SQ_THREADS, t.smem, st>>>(a);
      k_sq_assemble<true><<<(len + 31) / 32, 256, 0, st>>>(a.partial, nparts, a.K, a.L, out);
    } else {
      cudaFuncSetAttribute(k_nmtf_sq_tiled<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      k_nmtf_sq_tiled<false><<<dim3(nparts, t.npass), SQ_THREADS, t.smem, st>>>(a);
      k_sq_assemble<false><<<(len + 31) / 32, 256, 0, st>>>(a.partial, nparts, a.K, a.L, out);
    }
    return check_launch("nmtf_sq");
  }
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
  k_sum_partials2<<<(len + 31) / 32, 256, 0, st>>>(a.partial, nparts, len, out);
  return check_launch("nmtf_sq");
}

int launch_coord_solve(const CoordArgs& a, cudaStream_t st) {
  const size_t staged = ((size_t)a.D * a.D + 4 * a.D) * sizeof(double);
  if (staged <= 200 * 1024) {
    cudaFuncSetAttribute(k_coord_solve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_coord_solve<true><<<1, 256, staged, st>>>(a);
  } else {
    cudaFuncSetAttribute(k_coord_solve<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_coord_solve<false><<<1, 256, (size_t)4 * a.D * sizeof(double), st>>>(a);
  }
  return check_launch("coord_solve");
}

int launch_nmtf_extra(const ExtraArgs& a, cudaStream_t st) {
  const size_t per = ((size_t)a.K * a.K + a.K) * sizeof(double);
  const int warps = pick_warps(per, 4);
  cudaFuncSetAttribute(k_nmtf_extra, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k_nmtf_extra<<<(a.rows + warps - 1) / warps, warps * 32, per * warps, st>>>(a);
  return check_launch("nmtf_extra");
}

int launch_nmtf_mstat(int rows, int K, int L, int polarity, const double* RXo, const double* Go, const double* Gfull_o,
                      const double* G, const double* S, double* mstat, cudaStream_t st) {
  const int warps = 8;
  k_nmtf_mstat<<<(rows + warps - 1) / warps, warps * 32, (size_t)warps * K * sizeof(double), st>>>(
      rows, K, L, polarity, RXo, Go, Gfull_o, G, S, mstat);
  return check_launch("nmtf_mstat");
}

}  // namespace bnmtf

// Single-kernel sweeps for SMALL two-factor problems (the reference's own toy 100 x 80 and GDSC 622 x 138 matrices,
// BASELINE.json configs 1 and 3): the multi-kernel sweep of the large-matrix path is ~45 launches whose own latencies
// (200-400 us per sweep even as a CUDA graph) dwarf the arithmetic of such a matrix.  Here ONE thread-block cluster
// runs `sweeps` whole iterations of bnmf_gibbs_optimised.run / bnmf_vb_optimised.run / nmf_icm.run
// (bnmf_gibbs_optimised.py:121-157, bnmf_vb_optimised.py:121-153, nmf_icm.py:114-149) without returning to the host:
//
//   * every CTA keeps its rows of R and its rows of R^T (masked, fp64) and their mask bits in shared memory for the
//     whole run;
//   * a phase = stage the other factor (L2 -> shared memory), a warp per row for the row statistics
//     c_i = sum_j m r x_j, G_i = sum_j m x_j x_j^T, SV_i = sum_j m var_j  (plain fp64 FMAs: at these sizes the tensor
//     cores have nothing to win), then a THREAD per row for the K sequential updates -- the same formulas, the same
//     Philox streams and the same truncated-normal code as the large-matrix solver (solve.cu), so both paths draw the
//     same chains;
//   * the updated rows go to global memory; a cluster barrier (release / acquire) and L2 loads (ld.global.cg) make
//     them visible to the other CTAs -- no kernel boundary between the phases;
//   * metrics over the training mask, the VB extra term and the factor-side ELBO terms are per-CTA partial sums added
//     in rank order by one thread, which also runs the common end-of-sweep code (tail.cuh: tau, trace row).
#include <cooperative_groups.h>
#include "common.cuh"
#include "tail.cuh"

namespace cg = cooperative_groups;

namespace bnmtf {

constexpr int SM_KMAX = 16;
constexpr int SM_THREADS = 512;
constexpr int SM_WARPS = SM_THREADS / 32;
constexpr int SM_PARTIAL = 16;       // doubles of scratch per CTA

struct SmallArgs {
  int mode, I, J, K, ldJ, ldI, nrow[2];      // nrow[s]: rows of R (s = 0) / of R^T (s = 1) per CTA
  const double* R; const uint32_t* bits; const double* RT; const uint32_t* bitsT;
  double* fac[2]; double* var[2]; double* mu[2]; double* tauf[2]; const double* lam[2];     // [0] = U (I x K), [1] = V (J x K)
  double* scalars; double* trace; unsigned long long* iter; int trace_cap;
  double alpha, beta, digamma_alpha_s, lgamma_alpha, lgamma_alpha_s, min_tn;
  unsigned long long seed;
  int sweeps;
  double* all_U; double* all_V;              // Gibbs: the draws of every sweep, [sweep][n][K] (or NULL)
  double* sum_U; double* sum_V; int burn_in, thinning;     // Gibbs: running sums over sweeps burn_in, burn_in + thinning, ... (or NULL)
  double* partial;                           // cluster size x SM_PARTIAL doubles
  unsigned long long* times;                 // %globaltimer at the start and after every sweep (or NULL)
};

__device__ __forceinline__ int tri_index(int a, int b, int K) { return a * K - a * (a - 1) / 2 + (b - a); }   // a <= b

__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < SM_WARPS; ++w) s += red[w];
  return s;
}

// shared-memory carve-up (doubles unless noted); sizes computed identically on the host (small_smem_bytes)
struct SmallSmem {
  double* Rs; double* RTs; uint32_t* Ms; uint32_t* MTs; double* buf; double* stats; double* red; double* ubuf;
};

__host__ __device__ inline size_t small_layout(int I, int J, int K, int vb, int nr0, int nr1, size_t off[8]) {
  const int wJ = (J + 31) / 32, wI = (I + 31) / 32, nmax = nr0 > nr1 ? nr0 : nr1, cmax = I > J ? I : J;
  const int ns = K * (K + 1) / 2 + 2 * K;
  size_t o = 0;
  off[0] = o; o += (size_t)nr0 * J * 8;
  off[1] = o; o += (size_t)nr1 * I * 8;
  off[2] = o; o += ((size_t)nr0 * wJ * 4 + 7) / 8 * 8;
  off[3] = o; o += ((size_t)nr1 * wI * 4 + 7) / 8 * 8;
  off[4] = o; o += (size_t)cmax * K * (vb ? 2 : 1) * 8;
  off[5] = o; o += (size_t)nmax * (ns > K ? ns : K) * 8;
  off[6] = o; o += 64 * 8;
  off[7] = o; o += (size_t)K * nmax * 8;                 // the update chains' current row values, [k][row]
  return o;
}

// row statistics of one phase: a warp per row, lane = (column group, k).  K is a compile-time constant here: with a
// run-time K every "if (k2 < K)" of the unrolled inner loop became a basic block of its own and the ten shared-memory
// loads of an iteration were issued one after the other (measured: 490 clk per iteration instead of ~60).
template <int K, bool VB>
__device__ void small_stats(const SmallSmem& sm, const double* __restrict__ slice, const uint32_t* __restrict__ mb, int cols, int nr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpr = (cols + 31) / 32;
  constexpr int NT = K * (K + 1) / 2, NS = NT + 2 * K, NG = 32 / K;
  const int grp = lane / K, kk = lane - grp * K;
  const bool active = lane < NG * K;
  const double* X = sm.buf;
  const double* XV = sm.buf + (size_t)cols * K;
  for (int r = warp; r < nr; r += SM_WARPS) {
    double g[K];
#pragma unroll
    for (int k2 = 0; k2 < K; ++k2) g[k2] = 0.0;
    double c = 0.0, sv = 0.0;
    if (active) {
      const double* rrow = slice + (size_t)r * cols;
      const uint32_t* mrow = mb + (size_t)r * wpr;
#pragma unroll 2
      for (int j = grp; j < cols; j += NG) {
        const bool m = (mrow[j >> 5] >> (j & 31)) & 1u;
        const double* xj = X + (size_t)j * K;
        const double x = m ? xj[kk] : 0.0;
        c = fma(rrow[j], x, c);
        if (VB) sv += m ? XV[(size_t)j * K + kk] : 0.0;
#pragma unroll
        for (int k2 = 0; k2 < K; ++k2) g[k2] = fma(x, xj[k2], g[k2]);
      }
    }
    // add the column groups (fixed order) into the lanes of group 0; the shuffled values are the groups' own partial sums
    // (they are never modified here: every lane takes part in every shuffle)
    double ct = c, svt = sv, gt[K];
#pragma unroll
    for (int k2 = 0; k2 < K; ++k2) gt[k2] = g[k2];
#pragma unroll
    for (int gg = 1; gg < NG; ++gg) {
      const int src = (kk + gg * K) & 31;
      ct += __shfl_sync(0xffffffffu, c, src);
      if (VB) svt += __shfl_sync(0xffffffffu, sv, src);
#pragma unroll
      for (int k2 = 0; k2 < K; ++k2) gt[k2] += __shfl_sync(0xffffffffu, g[k2], src);
    }
    if (lane < K) {
      double* st = sm.stats + (size_t)r * NS;
#pragma unroll
      for (int k2 = 0; k2 < K; ++k2)
        if (k2 >= kk) st[tri_index(kk, k2, K)] = gt[k2];
      st[NT + kk] = ct;
      st[NT + K + kk] = svt;
    }
  }
}

template <bool VB>
__device__ void small_stats_k(int K, const SmallSmem& sm, const double* slice, const uint32_t* mb, int cols, int nr) {
  switch (K) {
#define BNMTF_SMALL_K(KV) case KV: small_stats<KV, VB>(sm, slice, mb, cols, nr); break;
    BNMTF_SMALL_K(1) BNMTF_SMALL_K(2) BNMTF_SMALL_K(3) BNMTF_SMALL_K(4) BNMTF_SMALL_K(5) BNMTF_SMALL_K(6) BNMTF_SMALL_K(7) BNMTF_SMALL_K(8)
    BNMTF_SMALL_K(9) BNMTF_SMALL_K(10) BNMTF_SMALL_K(11) BNMTF_SMALL_K(12) BNMTF_SMALL_K(13) BNMTF_SMALL_K(14) BNMTF_SMALL_K(15)
    BNMTF_SMALL_K(16)
#undef BNMTF_SMALL_K
  }
}

// one phase: rows [r0, r0 + nr) of the side-s data (s = 0: rows of R, factor U; s = 1: rows of R^T, factor V)
template <int MODE>
__device__ void small_phase(const SmallArgs& a, const SmallSmem& sm, int s, int r0, int nr, unsigned long long it, int sweep,
                            double& ex_out, double* dbg) {
  const int K = a.K, o = 1 - s;
  const int cols = s == 0 ? a.J : a.I;
  const double* slice = s == 0 ? sm.Rs : sm.RTs;
  const uint32_t* mb = s == 0 ? sm.Ms : sm.MTs;
  constexpr bool VB = MODE == MODE_VB;
  const int tid = threadIdx.x;
  const int NT = K * (K + 1) / 2, NS = NT + 2 * K;
  // 1. the other factor, from L2 (written by the other CTAs before the last cluster barrier)
  double* X = sm.buf;
  double* XV = sm.buf + (size_t)cols * K;
  for (int i = tid; i < cols * K; i += SM_THREADS) {
    X[i] = __ldcg(a.fac[o] + i);
    if (VB) XV[i] = __ldcg(a.var[o] + i);
  }
  __syncthreads();
  if (dbg && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[0] = (double)(t & 0xffffffffffffull); }
  // 2. row statistics
  small_stats_k<VB>(K, sm, slice, mb, cols, nr);
  __syncthreads();
  if (dbg && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[1] = (double)(t & 0xffffffffffffull); }
  // 3. the K sequential updates: a thread per row (solve.cu::k_bnmf_row_solve_lane is the large-matrix form of this loop).
  // Run-time loops with the row's current values in shared memory ([k][thread]): the truncated-normal code exists once.
  double ex = 0.0;
  if (tid < nr) {
    const int row = r0 + tid;
    const double* st = sm.stats + (size_t)tid * NS;
    double* u = sm.ubuf + tid;                          // u[k * nr_stride]
    const int us = a.nrow[0] > a.nrow[1] ? a.nrow[0] : a.nrow[1];
    const double tau = __ldcg(a.scalars + S_TAU);
    for (int k = 0; k < K; ++k) u[k * us] = __ldcg(a.fac[s] + (size_t)row * K + k);
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
      double acck = 0.0, part = 0.0;
      for (int c2 = 0; c2 < k; ++c2) acck = fma(st[tri_index(c2, k, K)], u[c2 * us], acck);
      for (int c2 = k + 1; c2 < K; ++c2) part = fma(st[tri_index(k, c2, K)], u[c2 * us], part);
      const double gkk = st[tri_index(k, k, K)], rxk = st[NT + k], svk = VB ? st[NT + K + k] : 0.0;
      const double sres = rxk - (acck + part);
      const double b = VB ? gkk + svk : gkk;
      const size_t idx = (size_t)row * K + k;
      const double lam = a.lam[s][idx];
      const double tau_k = tau * b;
      const double mu_k = (1.0 / tau_k) * (-lam + tau * sres);
      double val = 0.0, vv = 0.0;
      if (MODE == MODE_GIBBS) {
        Philox rng(a.seed, it * 16ull + (unsigned long long)s, (unsigned long long)row * K + k);
        val = tn_draw(mu_k, tau_k, rng);
      } else if (VB) {
        tn_moments(mu_k, tau_k, val, vv);
      } else {
        val = (mu_k != mu_k) ? mu_k : fmax(mu_k, 0.0);   // numpy.maximum propagates NaN (nmf_icm.py:129)
        val = (val != val) ? val : fmax(val, a.min_tn);
      }
      u[k * us] = val;
      a.fac[s][idx] = val;
      if (VB) a.var[s][idx] = vv;
      a.mu[s][idx] = mu_k;
      a.tauf[s][idx] = tau_k;
      ex += vv * (gkk + svk) + val * val * svk;
      if (MODE == MODE_GIBBS) {
        double* all = s == 0 ? a.all_U : a.all_V;
        if (all) all[((size_t)sweep * (s == 0 ? a.I : a.J) + row) * K + k] = val;
        double* sums = s == 0 ? a.sum_U : a.sum_V;
        if (sums && sweep >= a.burn_in && (sweep - a.burn_in) % a.thinning == 0) sums[idx] += val;
      }
    }
  }
  ex_out = ex;
}

template <int MODE>
__global__ void __launch_bounds__(SM_THREADS, 1) k_small_sweeps(SmallArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, K = a.K;
  constexpr bool VB = MODE == MODE_VB;
  size_t off[8];
  small_layout(a.I, a.J, K, VB, a.nrow[0], a.nrow[1], off);
  SmallSmem sm;
  sm.Rs = reinterpret_cast<double*>(smem_raw + off[0]);
  sm.RTs = reinterpret_cast<double*>(smem_raw + off[1]);
  sm.Ms = reinterpret_cast<uint32_t*>(smem_raw + off[2]);
  sm.MTs = reinterpret_cast<uint32_t*>(smem_raw + off[3]);
  sm.buf = reinterpret_cast<double*>(smem_raw + off[4]);
  sm.stats = reinterpret_cast<double*>(smem_raw + off[5]);
  sm.red = reinterpret_cast<double*>(smem_raw + off[6]);
  sm.ubuf = reinterpret_cast<double*>(smem_raw + off[7]);
  // this CTA's rows of both orientations
  int r0[2], nr[2];
  for (int s = 0; s < 2; ++s) {
    const int n = s == 0 ? a.I : a.J;
    r0[s] = min(n, rank * a.nrow[s]);
    nr[s] = min(n, r0[s] + a.nrow[s]) - r0[s];
  }
  // resident data: masked entries and mask bits (global layout: rows x ld doubles, rows x ld/32 words)
  for (int s = 0; s < 2; ++s) {
    const int cols = s == 0 ? a.J : a.I, ld = s == 0 ? a.ldJ : a.ldI, wpr = (cols + 31) / 32;
    const double* src = s == 0 ? a.R : a.RT;
    const uint32_t* bsrc = s == 0 ? a.bits : a.bitsT;
    double* dst = s == 0 ? sm.Rs : sm.RTs;
    uint32_t* mdst = s == 0 ? sm.Ms : sm.MTs;
    for (int i = tid; i < nr[s] * wpr; i += SM_THREADS) {
      const int r = i / wpr, w = i - r * wpr;
      uint32_t v = bsrc[(size_t)(r0[s] + r) * (ld / 32) + w];
      if (w * 32 + 32 > cols) v &= (1u << (cols - w * 32)) - 1u;
      mdst[i] = v;
    }
    __syncthreads();
    for (int i = tid; i < nr[s] * cols; i += SM_THREADS) {
      const int r = i / cols, j = i - r * cols;
      const bool m = (mdst[r * wpr + (j >> 5)] >> (j & 31)) & 1u;
      dst[i] = m ? src[(size_t)(r0[s] + r) * ld + j] : 0.0;
    }
  }
  __syncthreads();
  if (rank == 0 && tid == 0 && a.times) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.times[0] = t;
  }
  const int NT = K * (K + 1) / 2;
  (void)NT;
  auto stamp = [&](int sweep, int q) {                 // stage boundaries of the last sweep (rank 0), for tools/small_sweep_times.py
    if (rank == 0 && tid == 0 && sweep == a.sweeps - 1) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.partial[16 * SM_PARTIAL + q] = (double)(t & 0xffffffffffffull);
      if (q == 0 || q == 8) a.partial[16 * SM_PARTIAL + 9 + (q >> 3)] = (double)(clock64() & 0xffffffffffffll);     // -> SM clock
    }
  };
  for (int sweep = 0; sweep < a.sweeps; ++sweep) {
    const unsigned long long it = __ldcg(a.iter);
    double ex0, ex1;
    stamp(sweep, 0);
    small_phase<MODE>(a, sm, 0, r0[0], nr[0], it, sweep, ex0, (rank == 0 && sweep == a.sweeps - 1) ? a.partial + 16 * SM_PARTIAL + 11 : nullptr);
    stamp(sweep, 1);
    __threadfence();
    cluster.sync();
    stamp(sweep, 2);
    small_phase<MODE>(a, sm, 1, r0[1], nr[1], it, sweep, ex1, nullptr);
    stamp(sweep, 3);
    __threadfence();
    cluster.sync();
    stamp(sweep, 4);
    // ---- metrics over this CTA's rows of R with the new factors; VB: extra term and factor-side ELBO terms ----
    double* own = sm.stats;                     // the row statistics are no longer needed
    for (int i = tid; i < nr[0] * K; i += SM_THREADS) own[i] = __ldcg(a.fac[0] + (size_t)r0[0] * K + i);
    for (int i = tid; i < a.J * K; i += SM_THREADS) sm.buf[i] = __ldcg(a.fac[1] + i);
    __syncthreads();
    double se2 = 0.0, sp = 0.0, sp2 = 0.0, srp = 0.0, sr = 0.0, sr2 = 0.0, cnt = 0.0;
    const int wJ = (a.J + 31) / 32;
    for (int i = tid; i < nr[0] * a.J; i += SM_THREADS) {
      const int r = i / a.J, j = i - r * a.J;
      if ((sm.Ms[r * wJ + (j >> 5)] >> (j & 31)) & 1u) {
        double p = 0.0;
        for (int k = 0; k < K; ++k) p = fma(own[r * K + k], sm.buf[j * K + k], p);
        const double rv = sm.Rs[i], e = rv - p;
        se2 = fma(e, e, se2); sp += p; sp2 = fma(p, p, sp2); srp = fma(rv, p, srp); sr += rv; sr2 = fma(rv, rv, sr2); cnt += 1.0;
      }
    }
    double el0 = 0.0, el1 = 0.0;
    if (VB) {
      for (int s = 0; s < 2; ++s) {
        const long long n = (long long)(s == 0 ? a.I : a.J) * K;
        for (long long i = (long long)rank * SM_THREADS + tid; i < n; i += (long long)C * SM_THREADS) {
          const double l = a.lam[s][i], e = __ldcg(a.fac[s] + i), v = __ldcg(a.var[s] + i), m = __ldcg(a.mu[s] + i),
                       t = __ldcg(a.tauf[s] + i);
          el0 += log(l) - l * e;
          const double d = e - m;
          el1 += -0.5 * log(t) + log(0.5 * erfc_ref(-m * sqrt(t) / kSqrt2)) + t * 0.5 * (v + d * d);
        }
      }
    }
    double vals[10] = {se2, sp, sp2, srp, sr, sr2, cnt, ex1, el0, el1};
    for (int q = 0; q < 10; ++q) {
      const double t = block_sum(vals[q], sm.red);
      if (tid == 0) a.partial[(size_t)rank * SM_PARTIAL + q] = t;
    }
    stamp(sweep, 5);
    __threadfence();
    cluster.sync();
    stamp(sweep, 6);
    // ---- one thread: add the partial sums in rank order, end of sweep (tau, trace, sweep counter) ----
    if (rank == 0 && tid == 0) {
      double m8[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ex1s = 0.0, el8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int c = 0; c < C; ++c) {
        const double* p = a.partial + (size_t)c * SM_PARTIAL;
        for (int q = 0; q < 7; ++q) m8[q] += __ldcg(p + q);
        ex1s += __ldcg(p + 7);
        el8[0] += __ldcg(p + 8);
        el8[1] += __ldcg(p + 9);
      }
      FinishArgs f;
      f.mode = MODE; f.alpha = a.alpha; f.beta = a.beta; f.digamma_alpha_s = a.digamma_alpha_s;
      f.lgamma_alpha = a.lgamma_alpha; f.lgamma_alpha_s = a.lgamma_alpha_s; f.n_factor_elems = (a.I + a.J) * K;
      f.m8 = m8; f.ex1 = &ex1s; f.el8 = el8; f.scalars = a.scalars; f.trace = a.trace; f.iter = a.iter;
      f.trace_cap = a.trace_cap; f.seed = a.seed; f.update_tau = 1; f.trace_window = nullptr;
      finish_sweep(f);
      if (a.times) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.times[sweep + 1] = t;
      }
      __threadfence();
    }
    stamp(sweep, 7);
    cluster.sync();
    stamp(sweep, 8);
    (void)ex0;
  }
}

static size_t small_smem_bytes(int I, int J, int K, int vb, int C) {
  size_t off[8];
  return small_layout(I, J, K, vb, (I + C - 1) / C, (J + C - 1) / C, off);
}

// cluster size for (I, J, K): 0 when the problem does not qualify for the single-kernel sweep
int small_cluster_size(int I, int J, int K, int vb) {
  if (K < 1 || K > SM_KMAX || I < 1 || J < 1) return 0;
  for (int C = 1; C <= 16; C *= 2) {
    const int nr0 = (I + C - 1) / C, nr1 = (J + C - 1) / C;
    if (nr0 > SM_THREADS || nr1 > SM_THREADS) continue;
    if (C < 16 && (nr0 > 48 || nr1 > 48) ) continue;                 // prefer short per-CTA row lists (the chains are a thread per row)
    if (small_smem_bytes(I, J, K, vb, C) <= 226 * 1024) return C;
  }
  return 0;
}

template <int MODE>
static int launch_small_mode(const SmallArgs& a, int C, size_t smem, cudaStream_t st) {
  auto kern = k_small_sweeps<MODE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (C > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C); cfg.blockDim = dim3(SM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, a);
  if (le != cudaSuccess) { set_error("small_sweeps: launch failed: %s", cudaGetErrorString(le)); cudaGetLastError(); return -1; }
  return check_launch("small_sweeps");
}

int launch_small_sweeps(SmallArgs a, cudaStream_t st) {
  const int vb = a.mode == MODE_VB;
  const int C = small_cluster_size(a.I, a.J, a.K, vb);
  if (!C) { set_error("small_sweeps: %d x %d, K=%d does not fit the single-kernel sweep", a.I, a.J, a.K); return -4; }
  a.nrow[0] = (a.I + C - 1) / C; a.nrow[1] = (a.J + C - 1) / C;
  const size_t smem = small_smem_bytes(a.I, a.J, a.K, vb, C);
  switch (a.mode) {
    case MODE_GIBBS: return launch_small_mode<MODE_GIBBS>(a, C, smem, st);
    case MODE_VB: return launch_small_mode<MODE_VB>(a, C, smem, st);
    case MODE_ICM: return launch_small_mode<MODE_ICM>(a, C, smem, st);
  }
  set_error("small_sweeps: bad mode %d", a.mode);
  return -2;
}

}  // namespace bnmtf

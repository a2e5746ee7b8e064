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


// =====================================================================================================
// The same for the tri-factorisation R ~ F S G^T (bnmtf_gibbs_optimised.py:138-180, bnmtf_vb_optimised.py:160-204,
// nmtf_icm.py:132-173): row statistics w.r.t. G -> F columns (through the S-transform of nmtf.cu::k_nmtf_transform) and the
// K*L scalar S updates on the (KL x KL) normal equations reduced over the cluster; column statistics w.r.t. F -> G columns.
// Order F, S, G (Gibbs, ICM) or S, F, G in the host's shuffled orders (VB).  Same formulas, streams and end-of-sweep code
// as the multi-kernel engine (bnmtf.py::BNMTFEngine, nmtf.cu).
// =====================================================================================================
struct TriArgs {
  int mode, I, J, K, L, ldJ, ldI, nrow[2];
  const double* R; const uint32_t* bits; const double* RT; const uint32_t* bitsT;
  double* fac[3]; double* var[3]; double* mu[3]; double* tauf[3]; const double* lam[3];     // [0] = F (I x K), [1] = G (J x L), [2] = S (K x L)
  double* scalars; double* trace; unsigned long long* iter; int trace_cap;
  double alpha, beta, digamma_alpha_s, lgamma_alpha, lgamma_alpha_s, min_tn;
  unsigned long long seed;
  int sweeps;
  const int* orders;                         // VB: [sweep][K*L + K + L] (S, F, G orders), or NULL: natural order
  double* all_F; double* all_S; double* all_G;   // Gibbs draws per sweep, or NULL
  double* partial;                           // >= 288 doubles
  double* Hpart; double* Hsum;               // C x (D^2 + 2D), D^2 + 2D
  unsigned long long* times;
};

struct TriSmem {
  double* Rs; double* RTs; uint32_t* Ms; uint32_t* MTs; double* buf; double* stats; double* eff; double* red; double* ubuf;
  double* Sm; double* Sv; double* xs; double* Hs;
};

__host__ __device__ inline size_t tri_layout(int I, int J, int K, int L, int vb, int nr0, int nr1, size_t off[13]) {
  const int wJ = (J + 31) / 32, wI = (I + 31) / 32, nmax = nr0 > nr1 ? nr0 : nr1, cmax = I > J ? I : J;
  const int dm = K > L ? K : L, ns = dm * (dm + 1) / 2 + 2 * dm, D = K * L;
  size_t o = 0;
  off[0] = o; o += (size_t)nr0 * J * 8;
  off[1] = o; o += (size_t)nr1 * I * 8;
  off[2] = o; o += ((size_t)nr0 * wJ * 4 + 7) / 8 * 8;
  off[3] = o; o += ((size_t)nr1 * wI * 4 + 7) / 8 * 8;
  off[4] = o; o += (size_t)cmax * dm * (vb ? 2 : 1) * 8;
  off[5] = o; o += (size_t)nmax * ns * 8;                // raw statistics (other factor's dimension)
  off[6] = o; o += (size_t)nmax * ns * 8;                // effective statistics (own dimension)
  off[7] = o; o += 64 * 8;
  off[8] = o; o += (size_t)dm * nmax * 8;
  off[9] = o; o += (size_t)D * 8;
  off[10] = o; o += (size_t)D * 8;
  off[11] = o; o += (size_t)D * 8;
  off[12] = o; o += ((size_t)D * D + 2 * D) * 8;          // the S-phase normal equations (used by rank 0)
  return o;
}

// raw row statistics (dimension Lo, triangle + rx + sv) -> effective statistics of dimension Ks under Smat (Ks x Lo):
// nmtf.cu::k_nmtf_transform for one row, by one thread
template <bool VB>
__device__ void tri_transform_row(const double* __restrict__ st, int Lo, int Ks, const double* __restrict__ Sm,
                                  const double* __restrict__ Sv, bool transposed, int ldS, double* __restrict__ out) {
  const int NTo = Lo * (Lo + 1) / 2, NTs = Ks * (Ks + 1) / 2;
  auto GG = [&](int a, int b) { return a <= b ? st[tri_index(a, b, Lo)] : st[tri_index(b, a, Lo)]; };
  auto S = [&](int k, int l) { return transposed ? Sm[l * ldS + k] : Sm[k * ldS + l]; };
  auto VS = [&](int k, int l) { return transposed ? Sv[l * ldS + k] : Sv[k * ldS + l]; };
  const double* rx = st + NTo;
  const double* sv = st + NTo + Lo;
  for (int k = 0; k < Ks; ++k) {
    double c = 0.0, e = 0.0;
    for (int l = 0; l < Lo; ++l) {
      const double s = S(k, l);
      c = fma(s, rx[l], c);
      if (VB) {
        const double d = GG(l, l);
        e += (VS(k, l) + s * s) * (d + sv[l]) - s * s * d;
      }
    }
    out[NTs + k] = c;
    out[NTs + Ks + k] = e;
  }
  for (int k = 0; k < Ks; ++k) {
    double T[SM_KMAX];                                    // T[l] = sum_q S[k][q] GG[q][l], once per k
    for (int l = 0; l < Lo; ++l) {
      double t = 0.0;
      for (int q = 0; q < Lo; ++q) t = fma(S(k, q), GG(q, l), t);
      T[l] = t;
    }
    for (int k2 = k; k2 < Ks; ++k2) {
      double v = 0.0;
      for (int l = 0; l < Lo; ++l) {
        const double s2 = S(k2, l);
        v = fma(T[l], s2, v);
        if (VB && k != k2) v = fma(S(k, l) * s2, sv[l], v);
      }
      out[tri_index(k, k2, Ks)] = v;
    }
  }
}

// the sequential updates of one row on its effective statistics (dimension Ks), in `order` (or natural order)
template <int MODE>
__device__ double tri_chain(const TriArgs& a, int f, int row, int Ks, const double* __restrict__ st, double* u, int us,
                            const int* order, unsigned long long it, int salt, int sweep, double* all, int nrows_total) {
  constexpr bool VB = MODE == MODE_VB;
  const int NT = Ks * (Ks + 1) / 2;
  const double tau = __ldcg(a.scalars + S_TAU);
  for (int k = 0; k < Ks; ++k) u[k * us] = __ldcg(a.fac[f] + (size_t)row * Ks + k);
#pragma unroll 1
  for (int o = 0; o < Ks; ++o) {
    const int k = order ? order[o] : o;
    double acck = 0.0, part = 0.0;
    for (int c2 = 0; c2 < k; ++c2) acck = fma(st[tri_index(c2, k, Ks)], u[c2 * us], acck);
    for (int c2 = k + 1; c2 < Ks; ++c2) part = fma(st[tri_index(k, c2, Ks)], u[c2 * us], part);
    const double gkk = st[tri_index(k, k, Ks)], rxk = st[NT + k], svk = VB ? st[NT + Ks + k] : 0.0;
    const double sres = rxk - (acck + part);
    const double b = VB ? gkk + svk : gkk;
    const size_t idx = (size_t)row * Ks + k;
    const double lam = a.lam[f][idx];
    const double tau_k = tau * b;
    const double mu_k = (1.0 / tau_k) * (-lam + tau * sres);
    double val = 0.0, vv = 0.0;
    if (MODE == MODE_GIBBS) {
      Philox rng(a.seed, it * 16ull + (unsigned long long)salt, (unsigned long long)row * Ks + k);
      val = tn_draw(mu_k, tau_k, rng);
    } else if (VB) {
      tn_moments(mu_k, tau_k, val, vv);
    } else {
      val = (mu_k != mu_k) ? mu_k : fmax(mu_k, 0.0);
      val = (val != val) ? val : fmax(val, a.min_tn);
    }
    u[k * us] = val;
    a.fac[f][idx] = val;
    if (VB) a.var[f][idx] = vv;
    a.mu[f][idx] = mu_k;
    a.tauf[f][idx] = tau_k;
    if (MODE == MODE_GIBBS && all) all[((size_t)sweep * nrows_total + row) * Ks + k] = val;
  }
  return 0.0;
}

template <int MODE>
__global__ void __launch_bounds__(SM_THREADS, 1) k_small_tri(TriArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, K = a.K, L = a.L, D = K * L;
  constexpr bool VB = MODE == MODE_VB;
  size_t off[13];
  tri_layout(a.I, a.J, K, L, VB, a.nrow[0], a.nrow[1], off);
  TriSmem sm;
  sm.Rs = reinterpret_cast<double*>(smem_raw + off[0]);
  sm.RTs = reinterpret_cast<double*>(smem_raw + off[1]);
  sm.Ms = reinterpret_cast<uint32_t*>(smem_raw + off[2]);
  sm.MTs = reinterpret_cast<uint32_t*>(smem_raw + off[3]);
  sm.buf = reinterpret_cast<double*>(smem_raw + off[4]);
  sm.stats = reinterpret_cast<double*>(smem_raw + off[5]);
  sm.eff = reinterpret_cast<double*>(smem_raw + off[6]);
  sm.red = reinterpret_cast<double*>(smem_raw + off[7]);
  sm.ubuf = reinterpret_cast<double*>(smem_raw + off[8]);
  sm.Sm = reinterpret_cast<double*>(smem_raw + off[9]);
  sm.Sv = reinterpret_cast<double*>(smem_raw + off[10]);
  sm.xs = reinterpret_cast<double*>(smem_raw + off[11]);
  sm.Hs = reinterpret_cast<double*>(smem_raw + off[12]);
  SmallSmem ss;                                        // view for small_stats
  ss.Rs = sm.Rs; ss.RTs = sm.RTs; ss.Ms = sm.Ms; ss.MTs = sm.MTs; ss.buf = sm.buf; ss.stats = sm.stats; ss.red = sm.red; ss.ubuf = sm.ubuf;
  int r0[2], nr[2];
  for (int s = 0; s < 2; ++s) {
    const int n = s == 0 ? a.I : a.J;
    r0[s] = min(n, rank * a.nrow[s]);
    nr[s] = min(n, r0[s] + a.nrow[s]) - r0[s];
  }
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
  const int us = a.nrow[0] > a.nrow[1] ? a.nrow[0] : a.nrow[1];
  const int NTl = L * (L + 1) / 2, NTk = K * (K + 1) / 2;
  const int NSl = NTl + 2 * L, NSk = NTk + 2 * K;      // doubles per row record of dimension L / K (triangle, rx, sv)
  const int HL = D * D + 2 * D;

  auto load_S = [&]() {
    for (int i = tid; i < D; i += SM_THREADS) { sm.Sm[i] = __ldcg(a.fac[2] + i); sm.Sv[i] = VB ? __ldcg(a.var[2] + i) : 0.0; }
    __syncthreads();
  };
  auto stage = [&](int f, int n, int dims) {              // factor f (n x dims) into buf: exp, then var (VB)
    for (int i = tid; i < n * dims; i += SM_THREADS) {
      sm.buf[i] = __ldcg(a.fac[f] + i);
      if (VB) sm.buf[(size_t)n * dims + i] = __ldcg(a.var[f] + i);
    }
    __syncthreads();
  };

  for (int sweep = 0; sweep < a.sweeps; ++sweep) {
    const unsigned long long it = __ldcg(a.iter);
    const int* ord = a.orders ? a.orders + (size_t)sweep * (D + K + L) : nullptr;
    // ---- row statistics w.r.t. G ----
    stage(1, a.J, L);
    small_stats_k<VB>(L, ss, sm.Rs, sm.Ms, a.J, nr[0]);
    __syncthreads();

    auto phase_F = [&]() {
      load_S();
      if (tid < nr[0]) {
        tri_transform_row<VB>(sm.stats + (size_t)tid * NSl, L, K, sm.Sm, sm.Sv, false, L, sm.eff + (size_t)tid * NSk);
        tri_chain<MODE>(a, 0, r0[0] + tid, K, sm.eff + (size_t)tid * NSk, sm.ubuf + tid, us, ord ? ord + D : nullptr, it, 0, sweep,
                        a.all_F, a.I);
      }
      __threadfence();
      cluster.sync();
    };
    auto phase_S = [&]() {
      // partial normal equations over this CTA's rows (nmtf.cu::k_nmtf_sq_partial), then the K*L scalar updates by rank 0
      double* own = sm.eff;                              // this CTA's rows of F (exp, var)
      for (int i = tid; i < nr[0] * K; i += SM_THREADS) {
        own[i] = __ldcg(a.fac[0] + (size_t)r0[0] * K + i);
        own[(size_t)us * K + i] = VB ? __ldcg(a.var[0] + (size_t)r0[0] * K + i) : 0.0;
      }
      __syncthreads();
      double* hp = a.Hpart + (size_t)rank * HL;
      for (int e = tid; e < HL; e += SM_THREADS) {
        double acc = 0.0;
        if (e < D * D) {
          const int d = e / D, d2 = e - d * D;
          const int k = d / L, l = d - k * L, k2 = d2 / L, l2 = d2 - k2 * L;
          const int gi = l <= l2 ? tri_index(l, l2, L) : tri_index(l2, l, L);
          for (int r = 0; r < nr[0]; ++r) {
            const double* st = sm.stats + (size_t)r * NSl;
            const double fk = own[r * K + k], fk2 = own[r * K + k2], gg = st[gi];
            double v = fk * fk2 * gg;
            if (VB) {
              if (l == l2 && k != k2) v = fma(fk * fk2, st[NTl + L + l], v);
              if (k == k2 && l != l2) v = fma(own[(size_t)us * K + r * K + k], gg, v);
            }
            acc += v;
          }
        } else if (e < D * D + D) {
          const int d = e - D * D, k = d / L, l = d - k * L;
          for (int r = 0; r < nr[0]; ++r) {
            const double* st = sm.stats + (size_t)r * NSl;
            const double fk = own[r * K + k], gd = st[tri_index(l, l, L)];
            acc += VB ? (own[(size_t)us * K + r * K + k] + fk * fk) * (gd + st[NTl + L + l]) : fk * fk * gd;
          }
        } else {
          const int d = e - D * D - D, k = d / L, l = d - k * L;
          for (int r = 0; r < nr[0]; ++r) acc += own[r * K + k] * sm.stats[(size_t)r * NSl + NTl + l];
        }
        hp[e] = acc;
      }
      __threadfence();
      cluster.sync();
      for (int e = rank * SM_THREADS + tid; e < HL; e += C * SM_THREADS) {       // every CTA adds a slice, in rank order
        double t = 0.0;
        for (int c = 0; c < C; ++c) t += __ldcg(a.Hpart + (size_t)c * HL + e);
        a.Hsum[e] = t;
      }
      __threadfence();
      cluster.sync();
      if (rank == 0) {
        // nmtf.cu::k_coord_solve: the K*L scalar updates are one serial chain -- a single warp on a shared-memory copy of
        // the normal equations (a lane per entry of the row, shuffles for the dot product)
        double* Hs = sm.Hs;
        for (int i = tid; i < HL; i += SM_THREADS) Hs[i] = __ldcg(a.Hsum + i);
        for (int i = tid; i < D; i += SM_THREADS) sm.xs[i] = __ldcg(a.fac[2] + i);
        __syncthreads();
        const double tau = __ldcg(a.scalars + S_TAU);
        if (tid < 32)
        for (int o = 0; o < D; ++o) {
          const int d = ord ? ord[o] : o;
          double part = 0.0;
          for (int i = tid; i < D; i += 32)
            if (i != d) part = fma(Hs[(size_t)d * D + i], sm.xs[i], part);
          const double dot = warp_sum(part);
          if (tid == 0) {
            const double sres = Hs[D * D + D + d] - dot;
            const double tau_d = tau * Hs[D * D + d];
            const double mu_d = (1.0 / tau_d) * (-a.lam[2][d] + tau * sres);
            double val = 0.0, vv = 0.0;
            if (MODE == MODE_GIBBS) {
              Philox rng(a.seed, it * 16ull + 2ull, (unsigned long long)d);
              val = tn_draw(mu_d, tau_d, rng);
            } else if (VB) {
              tn_moments(mu_d, tau_d, val, vv);
              a.var[2][d] = vv;
            } else {
              val = (mu_d != mu_d) ? mu_d : fmax(mu_d, 0.0);
              val = (val != val) ? val : fmax(val, a.min_tn);
            }
            a.fac[2][d] = val;
            a.mu[2][d] = mu_d;
            a.tauf[2][d] = tau_d;
            if (MODE == MODE_GIBBS && a.all_S) a.all_S[(size_t)sweep * D + d] = val;
            sm.xs[d] = val;
          }
          __syncwarp();
        }
        __syncthreads();
        __threadfence();
      }
      cluster.sync();
    };
    if (VB) { phase_S(); phase_F(); } else { phase_F(); phase_S(); }

    // ---- column statistics w.r.t. F, G columns, VB extra term ----
    stage(0, a.I, K);
    small_stats_k<VB>(K, ss, sm.RTs, sm.MTs, a.I, nr[1]);
    __syncthreads();
    load_S();
    double ex = 0.0;
    if (tid < nr[1]) {
      const int row = r0[1] + tid;
      const double* st = sm.stats + (size_t)tid * NSk;
      tri_transform_row<VB>(st, K, L, sm.Sm, sm.Sv, true, L, sm.eff + (size_t)tid * NSl);
      tri_chain<MODE>(a, 1, row, L, sm.eff + (size_t)tid * NSl, sm.ubuf + tid, us, ord ? ord + D + K : nullptr, it, 1, sweep, a.all_G, a.J);
      if (VB) {
        // nmtf.cu::k_nmtf_extra for this column, with the new G_j
        auto FF = [&](int x, int y) { return x <= y ? st[tri_index(x, y, K)] : st[tri_index(y, x, K)]; };
        const double* sv = st + NTk + K;
        const double* g = a.fac[1] + (size_t)row * L;
        const double* vg = a.var[1] + (size_t)row * L;
        for (int i = 0; i < D; ++i) {                                  // t2
          const int k = i / L, l = i - k * L;
          const double s = sm.Sm[i], gl = g[l], d = FF(k, k);
          ex += (vg[l] + gl * gl) * (sm.Sv[i] + s * s) * (d + sv[k]) - gl * gl * s * s * d;
        }
        for (int k = 0; k < K; ++k) {                                  // t3
          double sg = 0.0, sq = 0.0;
          for (int l = 0; l < L; ++l) { const double s = sm.Sm[k * L + l]; sg = fma(s, g[l], sg); sq = fma(s * s, g[l] * g[l], sq); }
          ex += sv[k] * (sg * sg - sq);
        }
        for (int l = 0; l < L; ++l) {                                  // t4
          double q = 0.0, sq = 0.0;
          for (int k = 0; k < K; ++k) {
            double t = 0.0;
            for (int k2 = 0; k2 < K; ++k2) t = fma(FF(k, k2), sm.Sm[k2 * L + l], t);
            q = fma(sm.Sm[k * L + l], t, q);
            sq = fma(sm.Sm[k * L + l] * sm.Sm[k * L + l], FF(k, k), sq);
          }
          ex += vg[l] * (q - sq);
        }
      }
    }
    __threadfence();
    cluster.sync();

    // ---- metrics over this CTA's rows of R with the new factors; VB: factor-side ELBO terms ----
    double* fs = sm.eff;                                 // (F S) rows of this CTA: nr0 x L
    for (int i = tid; i < nr[0] * L; i += SM_THREADS) {
      const int r = i / L, l = i - r * L;
      double t = 0.0;
      for (int k = 0; k < K; ++k) t = fma(__ldcg(a.fac[0] + (size_t)(r0[0] + r) * K + k), sm.Sm[k * L + l], t);
      fs[i] = t;
    }
    for (int i = tid; i < a.J * L; i += SM_THREADS) sm.buf[i] = __ldcg(a.fac[1] + i);
    __syncthreads();
    double se2 = 0.0, sp = 0.0, sp2 = 0.0, srp = 0.0, sr = 0.0, sr2 = 0.0, cnt = 0.0;
    const int wJ = (a.J + 31) / 32;
    for (int i = tid; i < nr[0] * a.J; i += SM_THREADS) {
      const int r = i / a.J, j = i - r * a.J;
      if ((sm.Ms[r * wJ + (j >> 5)] >> (j & 31)) & 1u) {
        double p = 0.0;
        for (int l = 0; l < L; ++l) p = fma(fs[r * L + l], sm.buf[j * L + l], p);
        const double rv = sm.Rs[i], e = rv - p;
        se2 = fma(e, e, se2); sp += p; sp2 = fma(p, p, sp2); srp = fma(rv, p, srp); sr += rv; sr2 = fma(rv, rv, sr2); cnt += 1.0;
      }
    }
    double el0 = 0.0, el1 = 0.0;
    if (VB) {
      for (int f = 0; f < 3; ++f) {
        const long long n = f == 0 ? (long long)a.I * K : (f == 1 ? (long long)a.J * L : (long long)D);
        for (long long i = (long long)rank * SM_THREADS + tid; i < n; i += (long long)C * SM_THREADS) {
          const double l = a.lam[f][i], e = __ldcg(a.fac[f] + i), v = __ldcg(a.var[f] + i), m = __ldcg(a.mu[f] + i),
                       t = __ldcg(a.tauf[f] + i);
          el0 += log(l) - l * e;
          const double d = e - m;
          el1 += -0.5 * log(t) + log(0.5 * erfc_ref(-m * sqrt(t) / kSqrt2)) + t * 0.5 * (v + d * d);
        }
      }
    }
    double vals[10] = {se2, sp, sp2, srp, sr, sr2, cnt, ex, el0, el1};
    for (int q = 0; q < 10; ++q) {
      const double t = block_sum(vals[q], sm.red);
      if (tid == 0) a.partial[(size_t)rank * SM_PARTIAL + q] = t;
    }
    __threadfence();
    cluster.sync();
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
      f.lgamma_alpha = a.lgamma_alpha; f.lgamma_alpha_s = a.lgamma_alpha_s; f.n_factor_elems = a.I * K + D + a.J * L;
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
    cluster.sync();
  }
}

static size_t tri_smem_bytes(int I, int J, int K, int L, int vb, int C) {
  size_t off[13];
  return tri_layout(I, J, K, L, vb, (I + C - 1) / C, (J + C - 1) / C, off);
}

int small_tri_cluster_size(int I, int J, int K, int L, int vb) {
  // K*L <= 64: the K*L scalar S updates are a serial chain and every row's S-transform is one thread's work -- measured on the
  // GDSC matrix (us per Gibbs / VB / ICM sweep, single kernel vs per-phase kernels): K = L = 5: 114 / 137 / 80 vs 182 / 248 /
  // 181; K = L = 7: 203 / 241 / 146 vs 240 / 318 / 192; K = L = 10: 425 / - / 689 vs 426 / 526 / 601
  if (K < 1 || L < 1 || K > SM_KMAX || L > SM_KMAX || K * L > 64 || I < 1 || J < 1) return 0;
  for (int C = 1; C <= 16; C *= 2) {
    const int nr0 = (I + C - 1) / C, nr1 = (J + C - 1) / C;
    if (nr0 > SM_THREADS || nr1 > SM_THREADS) continue;
    if (C < 16 && (nr0 > 48 || nr1 > 48)) continue;
    if (tri_smem_bytes(I, J, K, L, vb, C) <= 226 * 1024) return C;
  }
  return 0;
}

template <int MODE>
static int launch_tri_mode(const TriArgs& a, int C, size_t smem, cudaStream_t st) {
  auto kern = k_small_tri<MODE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (C > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C); cfg.blockDim = dim3(SM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, a);
  if (le != cudaSuccess) { set_error("small_tri_sweeps: launch failed: %s", cudaGetErrorString(le)); cudaGetLastError(); return -1; }
  return check_launch("small_tri_sweeps");
}

int launch_small_tri(TriArgs a, cudaStream_t st) {
  const int vb = a.mode == MODE_VB;
  const int C = small_tri_cluster_size(a.I, a.J, a.K, a.L, vb);
  if (!C) { set_error("small_tri_sweeps: %d x %d, K=%d, L=%d does not fit the single-kernel sweep", a.I, a.J, a.K, a.L); return -4; }
  a.nrow[0] = (a.I + C - 1) / C; a.nrow[1] = (a.J + C - 1) / C;
  const size_t smem = tri_smem_bytes(a.I, a.J, a.K, a.L, vb, C);
  switch (a.mode) {
    case MODE_GIBBS: return launch_tri_mode<MODE_GIBBS>(a, C, smem, st);
    case MODE_VB: return launch_tri_mode<MODE_VB>(a, C, smem, st);
    case MODE_ICM: return launch_tri_mode<MODE_ICM>(a, C, smem, st);
  }
  set_error("small_tri_sweeps: bad mode %d", a.mode);
  return -2;
}

}  // namespace bnmtf

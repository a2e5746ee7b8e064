// K-means with missing values (init_FG = 'kmeans'; code/models/kmeans/kmeans.py:105-133): the assignment step's
// distances  d(i, c) = sum_j m_ij mc_cj (x_ij - cen_cj)^2 / sum_j m_ij mc_cj   for every point i and centroid c.
//
// Near-ties between two centroids decide cluster membership and hence the whole clustering, so the device reproduces the
// BITS of the reference's numpy evaluation (bnmtf_b200/kmeans.py::_all_distances is the host statement of it): every
// term with separately rounded subtraction, square and mask product (no FMA contraction), and the sum in numpy's
// pairwise order -- blocks of <= 128 terms on eight interleaved accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))
// plus a sequential tail, blocks joined by recursive halving at multiples of eight (numpy/core/src/umath/loops_utils.h.src,
// DOUBLE_pairwise_sum; checked on the host against np.sum for lengths 1 ... 32768).
//
// A warp per point, a lane per centroid: the point's coordinates are broadcast loads, the centroid rows (K x d, small)
// stay in L1/L2.  A one-off initialisation: no attempt at the roofline.
#include "common.cuh"

namespace bnmtf {

struct KmTerm {
  const double* x; const double* m; const double* c; const double* mc;
  __device__ __forceinline__ double operator()(int j) const {
    const double both = __dmul_rn(m[j], mc[j]);
    const double diff = __dsub_rn(x[j], c[j]);
    return __dmul_rn(both, __dmul_rn(diff, diff));
  }
};

__device__ double km_block_sum(const KmTerm& t, int lo, int n) {      // n <= 128
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = __dadd_rn(r, t(lo + i));
    return r;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = t(lo + j);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], t(lo + i + j));
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, t(lo + i));
  return res;
}

// the recursion  sum(lo, n) = n <= 128 ? block : sum(lo, n2) + sum(lo + n2, n - n2),  n2 = n/2 rounded down to a multiple
// of eight, walked with an explicit stack (depth <= log2(d / 64))
__device__ double km_pairwise(const KmTerm& t, int lo0, int n0) {
  int lo[32], n[32], state[32];
  double left[32];
  int sp = 0;
  lo[0] = lo0; n[0] = n0; state[0] = 0;
  double ret = 0.0;
  while (sp >= 0) {
    if (state[sp] == 0) {
      if (n[sp] <= 128) { ret = km_block_sum(t, lo[sp], n[sp]); --sp; continue; }
      int n2 = n[sp] / 2;
      n2 -= n2 % 8;
      state[sp] = 1;
      lo[sp + 1] = lo[sp]; n[sp + 1] = n2; state[sp + 1] = 0;
      ++sp;
    } else if (state[sp] == 1) {
      left[sp] = ret;
      int n2 = n[sp] / 2;
      n2 -= n2 % 8;
      state[sp] = 2;
      lo[sp + 1] = lo[sp] + n2; n[sp + 1] = n[sp] - n2; state[sp + 1] = 0;
      ++sp;
    } else {
      ret = __dadd_rn(left[sp], ret);
      --sp;
    }
  }
  return ret;
}

__global__ void __launch_bounds__(128) k_kmeans_dist(const double* __restrict__ X, const double* __restrict__ M, int n, int d,
                                                    const double* __restrict__ C, const double* __restrict__ MC, int K,
                                                    double* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + warp;
  if (i >= n) return;
  for (int c = lane; c < K; c += 32) {
    KmTerm t{X + (size_t)i * d, M + (size_t)i * d, C + (size_t)c * d, MC + (size_t)c * d};
    double overlap = 0.0;                                  // a count: exact in any order
    for (int j = 0; j < d; ++j) overlap += t.m[j] * t.mc[j];
    const double sq = km_pairwise(t, 0, d);
    out[(size_t)i * K + c] = overlap > 0.0 ? __ddiv_rn(sq, overlap) : __longlong_as_double(0x7ff0000000000000ll);
  }
}

int launch_kmeans_dist(const double* X, const double* M, int n, int d, const double* C, const double* MC, int K, double* out,
                       cudaStream_t st) {
  if (n <= 0 || d <= 0 || K <= 0) { set_error("kmeans_distances: bad shape n=%d d=%d K=%d", n, d, K); return -2; }
  cudaFuncSetAttribute(k_kmeans_dist, cudaFuncAttributeMaxDynamicSharedMemorySize, 0);
  k_kmeans_dist<<<(n + 3) / 4, 128, 0, st>>>(X, M, n, d, C, MC, K, out);
  return check_launch("kmeans_distances");
}

}  // namespace bnmtf

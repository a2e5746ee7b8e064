// Non-probabilistic multiplicative updates (Lee & Seung I-divergence NMF, Yoo & Choi NMTF) with a mask:
// reference code/models/nmf_np.py:114-118 and code/models/nmtf_np.py:155-174.
//
// These updates divide R by the current prediction element by element, so they cannot be folded into the row
// statistics of stats.cu; the prediction P is kept in a device scratch matrix (rows x ld) and maintained
// incrementally.  Within a phase the rows are independent: one CTA owns a row and performs the K sequential
// column updates on it (two passes over the row per column, both from L1/L2).
#include "common.cuh"

namespace bnmtf {

// P[i][j] = sum_k A[i][k] * B[j][k]   (A: rows x K, B: cols x K, plain row-major), padding columns get 1
__global__ void __launch_bounds__(256) k_np_build_pred(const double* __restrict__ A, const double* __restrict__ B, int rows,
                                                      int cols, int ld, int K, double* __restrict__ P) {
  // rows on grid.x (2^31 - 1 blocks), column blocks on grid.y: gridDim.y stops at 65535, and row counts do not
  const int i = blockIdx.x;
  const double* a = A + (size_t)i * K;
  for (int j = blockIdx.y * 256 + threadIdx.x; j < ld; j += gridDim.y * 256) {
    double s = 1.0;
    if (j < cols) {
      s = 0.0;
      const double* b = B + (size_t)j * K;
      for (int k = 0; k < K; ++k) s = fma(a[k], b[k], s);
    }
    P[(size_t)i * ld + j] = s;
  }
}

__device__ __forceinline__ double block_sum_128(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

// One CTA (128 threads) per row: for k in 0..K-1:  a_k <- a_k * sum_j m r/p b_jk / sum_j m b_jk ;  p += delta b_jk
__global__ void __launch_bounds__(128) k_np_row_update(const double* __restrict__ R, const uint32_t* __restrict__ bits,
                                                      double* __restrict__ P, int rows, int cols, int ld,
                                                      double* __restrict__ A, const double* __restrict__ B, int K) {
  __shared__ double red[4];
  const int i = blockIdx.x;
  const double* r = R + (size_t)i * ld;
  double* p = P + (size_t)i * ld;
  const uint32_t* m = bits + (size_t)i * (ld >> 5);
  for (int k = 0; k < K; ++k) {
    double num = 0.0, den = 0.0;
    for (int j = threadIdx.x; j < cols; j += 128) {
      if ((m[j >> 5] >> (j & 31)) & 1u) {
        const double b = B[(size_t)j * K + k];
        num += b * (r[j] / p[j]);
        den += b;
      }
    }
    num = block_sum_128(num, red);
    den = block_sum_128(den, red);
    const double old = A[(size_t)i * K + k];
    const double nw = old * num / den;
    const double delta = nw - old;
    for (int j = threadIdx.x; j < cols; j += 128) p[j] = fma(delta, B[(size_t)j * K + k], p[j]);
    __syncthreads();
    if (threadIdx.x == 0) A[(size_t)i * K + k] = nw;
  }
}

// S update of the tri-factorisation, element (k,l): partial sums over a block of rows of
//   num = sum_ij m r F_ik G_jl / p_ij ,  den = sum_ij m F_ik G_jl
__global__ void __launch_bounds__(256) k_np_s_partial(const double* __restrict__ R, const uint32_t* __restrict__ bits,
                                                     const double* __restrict__ P, int rows, int cols, int ld,
                                                     const double* __restrict__ F, int K, int k,
                                                     const double* __restrict__ G, int L, int l,
                                                     double* __restrict__ partials) {
  double num = 0.0, den = 0.0;
  for (int i = blockIdx.x; i < rows; i += gridDim.x) {
    const double f = F[(size_t)i * K + k];
    const double* r = R + (size_t)i * ld;
    const double* p = P + (size_t)i * ld;
    const uint32_t* m = bits + (size_t)i * (ld >> 5);
    for (int j = threadIdx.x; j < cols; j += 256) {
      if ((m[j >> 5] >> (j & 31)) & 1u) {
        const double fg = f * G[(size_t)j * L + l];
        num += r[j] * fg / p[j];
        den += fg;
      }
    }
  }
  __shared__ double red[8][2];
  num = warp_sum(num); den = warp_sum(den);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = num; red[threadIdx.x >> 5][1] = den; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 2 + threadIdx.x] = s;
  }
}

// finalise S_kl from the partials (every CTA recomputes the same sum in the same order), then p += delta F_ik G_jl
__global__ void __launch_bounds__(256) k_np_s_apply(const double* __restrict__ partials, int nparts, double* __restrict__ S,
                                                   int K, int L, int k, int l, double* __restrict__ P, int rows, int cols,
                                                   int ld, const double* __restrict__ F, const double* __restrict__ G,
                                                   double* __restrict__ s_new_out) {
  double num = 0.0, den = 0.0;
  for (int q = 0; q < nparts; ++q) { num += partials[2 * q]; den += partials[2 * q + 1]; }
  const double old = S[k * L + l];
  const double nw = old * num / den;
  const double delta = nw - old;
  for (int i = blockIdx.x; i < rows; i += gridDim.x) {
    const double f = delta * F[(size_t)i * K + k];
    double* p = P + (size_t)i * ld;
    for (int j = threadIdx.x; j < cols; j += 256) p[j] = fma(f, G[(size_t)j * L + l], p[j]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) s_new_out[k * L + l] = nw;   // written to a shadow; copied after the kernel
}

// sums over observed entries of {(r-p)^2, p, p^2, r p, r, r^2, 1, r log(r/p) - r + p} with P explicit
__global__ void __launch_bounds__(256) k_np_metrics(const double* __restrict__ R, const uint32_t* __restrict__ bits,
                                                   const double* __restrict__ P, int rows, int cols, int ld,
                                                   double* __restrict__ partials) {
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x; i < rows; i += gridDim.x) {
    const double* r = R + (size_t)i * ld;
    const double* p = P + (size_t)i * ld;
    const uint32_t* m = bits + (size_t)i * (ld >> 5);
    for (int j = threadIdx.x; j < cols; j += 256) {
      if ((m[j >> 5] >> (j & 31)) & 1u) {
        const double rv = r[j], pv = p[j], e = rv - pv;
        s[0] += e * e; s[1] += pv; s[2] += pv * pv; s[3] += rv * pv; s[4] += rv; s[5] += rv * rv; s[6] += 1.0;
        s[7] += rv * log(rv / pv) - rv + pv;
      }
    }
  }
  __shared__ double red[8][8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const double v = warp_sum(s[c]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][c] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 8 + threadIdx.x] = v;
  }
}

// C (n x q) = A (n x p) * B (p x q)  or  A * B^T when transB (B is q x p); tiny operands (factor-sized)
__global__ void k_small_matmul(const double* __restrict__ A, const double* __restrict__ B, int n, int p, int q, int transB,
                               double* __restrict__ C) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * q) return;
  const int i = (int)(idx / q), c = (int)(idx - (long long)i * q);
  double s = 0.0;
  for (int k = 0; k < p; ++k) s = fma(A[(size_t)i * p + k], transB ? B[(size_t)c * p + k] : B[(size_t)k * q + c], s);
  C[idx] = s;
}

// ---- launchers -------------------------------------------------------------------------------------------
int launch_np_build_pred(const double* A, const double* B, int rows, int cols, int ld, int K, double* P, cudaStream_t st) {
  dim3 grid(rows, (ld + 255) / 256);
  if (grid.y > 64) grid.y = 64;
  k_np_build_pred<<<grid, 256, 0, st>>>(A, B, rows, cols, ld, K, P);
  return check_launch("np_build_pred");
}
int launch_np_row_update(const double* R, const uint32_t* bits, double* P, int rows, int cols, int ld, double* A,
                         const double* B, int K, cudaStream_t st) {
  k_np_row_update<<<rows, 128, 0, st>>>(R, bits, P, rows, cols, ld, A, B, K);
  return check_launch("np_row_update");
}
int launch_np_s_update(const double* R, const uint32_t* bits, double* P, int rows, int cols, int ld, const double* F, int K,
                       int k, const double* G, int L, int l, double* S, double* partials, int nparts, cudaStream_t st) {
  k_np_s_partial<<<nparts, 256, 0, st>>>(R, bits, P, rows, cols, ld, F, K, k, G, L, l, partials);
  // S itself is read by every CTA of the apply kernel, so the new value goes to a shadow slot first
  double* shadow = partials + 2 * (size_t)nparts;
  k_np_s_apply<<<nparts, 256, 0, st>>>(partials, nparts, S, K, L, k, l, P, rows, cols, ld, F, G, shadow - (k * L + l));
  cudaMemcpyAsync(S + k * L + l, shadow, sizeof(double), cudaMemcpyDeviceToDevice, st);
  return check_launch("np_s_update");
}
int launch_np_metrics(const double* R, const uint32_t* bits, const double* P, int rows, int cols, int ld, double* partials,
                      int nparts, cudaStream_t st) {
  k_np_metrics<<<nparts, 256, 0, st>>>(R, bits, P, rows, cols, ld, partials);
  return check_launch("np_metrics");
}
int launch_small_matmul(const double* A, const double* B, int n, int p, int q, int transB, double* C, cudaStream_t st) {
  const long long tot = (long long)n * q;
  if (tot <= 0) return 0;
  k_small_matmul<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(A, B, n, p, q, transB, C);
  return check_launch("small_matmul");
}

}  // namespace bnmtf

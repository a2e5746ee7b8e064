// Microbenchmark: fp64 pipe rates on B200 (sm_100a) that decide the sweep-kernel design.
//   1. DFMA peak (independent chains)            2. DFMA dependent-issue latency
//   3. DMMA (mma.sync m8n8k4 f64) peak            4. DFMA + DMMA concurrently (same SM, different warps / same warp)
//   5. streaming HBM read bandwidth               6. L2-resident read bandwidth
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template<int MODE>  // 0 dfma, 1 dmma, 2 split by warp parity, 3 interleaved in every warp
__global__ void __launch_bounds__(512) pipes(double *out, int iters, double a, double b) {
  double x[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3 + i; c[i] = i * 0.5; }
  int warp = threadIdx.x >> 5;
  bool do_fma = (MODE == 0) || (MODE == 3) || (MODE == 2 && (warp & 1) == 0);
  bool do_mma = (MODE == 1) || (MODE == 3) || (MODE == 2 && (warp & 1) == 1);
  for (int it = 0; it < iters; ++it) {
    if (do_fma) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
    if (do_mma) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) dmma(c[i], c[i + 1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + c[i];
  if (s == 123.456) out[0] = s;
}

__global__ void dfma_latency(double *out, long long *cyc, int iters, double a, double b) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x = fma(x, a, b);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; out[1] = x; }
}

__global__ void __launch_bounds__(256) stream_read(const double2 *__restrict__ p, size_t n, double *out) {
  double s = 0;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride * 4) {
    double2 v0 = p[i];
    double2 v1 = (i + stride < n) ? p[i + stride] : make_double2(0, 0);
    double2 v2 = (i + 2 * stride < n) ? p[i + 2 * stride] : make_double2(0, 0);
    double2 v3 = (i + 3 * stride < n) ? p[i + 3 * stride] : make_double2(0, 0);
    s += v0.x + v0.y + v1.x + v1.y + v2.x + v2.y + v3.x + v3.y;
  }
  if (s == 123.456) out[0] = s;
}

template<int MODE> float run_pipes(double *out, int blocks, int threads, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  pipes<MODE><<<blocks, threads>>>(out, 10, 1.0000001, 1e-9);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  pipes<MODE><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, max clock %d MHz\n", p.name, p.multiProcessorCount, clk_khz / 1000);
  double *out; CK(cudaMalloc(&out, 64)); long long *cyc; CK(cudaMalloc(&cyc, 64));
  int sms = p.multiProcessorCount;
  const int iters = 20000;
  for (int threads : {128, 256, 512}) {
    for (int bps : {1, 2, 4}) {
      if (threads * bps > 2048) continue;
      int blocks = sms * bps;
      double lanes = (double)blocks * threads;
      float t0 = run_pipes<0>(out, blocks, threads, iters);
      float t1 = run_pipes<1>(out, blocks, threads, iters);
      float t2 = run_pipes<2>(out, blocks, threads, iters);
      float t3 = run_pipes<3>(out, blocks, threads, iters);
      double f_fma = lanes * iters * 8.0 * 2.0;                 // flops of the DFMA part (all warps)
      double f_mma = (lanes / 32.0) * iters * 4.0 * 256.0 * 2;  // 4 DMMA per warp-iter, 8x8x4 FMAs each
      printf("threads %4d x %d/SM: DFMA %.2f TF/s | DMMA %.2f TF/s | split-warps %.2f TF/s (%.3f ms vs %.3f/%.3f alone-half) | interleaved %.2f TF/s\n",
             threads, bps, f_fma / t0 * 1e-9, f_mma / t1 * 1e-9,
             (f_fma / 2 + f_mma / 2) / t2 * 1e-9, t2, t0 / 2, t1 / 2, (f_fma + f_mma) / t3 * 1e-9);
    }
  }
  dfma_latency<<<1, 32>>>(out, cyc, 1000, 1.0000001, 1e-9);
  long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("DFMA dependent latency: %.2f cycles\n", (double)h / 16000.0);

  // HBM / L2 streaming read
  for (size_t mb : {(size_t)64, (size_t)8192}) {
    size_t bytes = mb << 20; double2 *buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
    size_t n = bytes / 16;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int bps : {4, 8}) {
      stream_read<<<sms * bps, 256>>>(buf, n, out); CK(cudaDeviceSynchronize());
      int reps = mb < 1024 ? 200 : 5;
      cudaEventRecord(e0);
      for (int r = 0; r < reps; ++r) stream_read<<<sms * bps, 256>>>(buf, n, out);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("stream read %zu MiB, %d CTAs/SM: %.1f GB/s\n", mb, bps, (double)bytes * reps / ms * 1e-6);
    }
    cudaFree(buf);
  }
  return 0;
}

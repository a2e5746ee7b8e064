// Cross-GPU synchronisation and the tiny exchanges of a row-sharded sweep, as plain kernels over peer-mapped memory
// (NVLink P2P loads / stores; the mapping itself -- symmetric memory or CUDA IPC -- is the caller's).
//
// SURVEY.md section 8e lists three exchanges per sweep: an all-gather of each updated factor and an all-reduce of a
// few scalars.  The factor rows are stored into the peers' copies by the solver kernels themselves (solve.cu); what
// is left is (1) one barrier per phase, so that nobody reads a factor before every rank's rows have landed, (2) the
// all-reduce of <= 32 doubles (metric sums, ELBO terms), (3) the replication of freshly uploaded factor rows.  Doing
// these with our own kernels instead of NCCL / torch calls keeps the whole sharded sweep a fixed sequence of kernel
// launches with fixed arguments -- it can be captured in a CUDA graph like the single-GPU sweep.
//
// Every rank owns one SYNC BLOCK in peer-mapped memory (bnmtf_peer_sync_bytes() bytes, zeroed once):
//     flags[channel][rank]   u64   epoch last signalled by `rank` on `channel`
//     red[parity][rank][32]  f64   all-reduce contributions, double-buffered by epoch parity
// and a private device array epoch[channel] (u64, zeroed once).  A call on a channel bumps that channel's epoch,
// publishes (payload, then flag with release semantics at system scope) into every rank's block, and waits until
// every rank's flag in the LOCAL block has reached the epoch (acquire at system scope).  Because a rank can only
// pass epoch e after every rank has published e, and publishes e+1 only after it has consumed e, two payload
// buffers suffice.  The sums are taken in rank order on every rank: bitwise identical results everywhere.
#include "common.cuh"

namespace bnmtf {

constexpr int PEER_CHANNELS = 8;
constexpr int PEER_MAX_RANKS = 64;
constexpr int PEER_RED = 32;
constexpr size_t PEER_FLAGS_U64 = (size_t)PEER_CHANNELS * PEER_MAX_RANKS;
constexpr size_t PEER_BLOCK_U64 = PEER_FLAGS_U64 + 2ull * PEER_MAX_RANKS * PEER_RED;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_f64(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

struct PeerSyncArgs {
  unsigned long long* const* blocks;   // device array [world]: base of every rank's sync block (peer-mapped)
  unsigned long long* epoch;           // private [PEER_CHANNELS]
  double* data;                        // in/out payload (n doubles) or nullptr: barrier only
  int world, rank, channel, n;
};

__global__ void __launch_bounds__(128) k_peer_sync(PeerSyncArgs a) {
  __shared__ unsigned long long e_sh;
  const int tid = threadIdx.x;
  if (tid == 0) e_sh = ++a.epoch[a.channel];
  __syncthreads();
  const unsigned long long e = e_sh;
  const size_t par = (size_t)(e & 1ull);
  if (a.data) {
    for (int idx = tid; idx < a.world * a.n; idx += blockDim.x) {
      const int p = idx / a.n, i = idx - p * a.n;
      double* dst = reinterpret_cast<double*>(a.blocks[p] + PEER_FLAGS_U64) + (par * PEER_MAX_RANKS + a.rank) * PEER_RED + i;
      st_relaxed_sys_f64(dst, a.data[i]);
    }
  }
  __syncthreads();
  if (tid < a.world) {
    // everything this GPU has written so far -- the payload above and the factor rows the preceding solver kernel
    // stored into the peers' copies -- becomes visible before the flag does
    __threadfence_system();
    st_release_sys(a.blocks[tid] + (size_t)a.channel * PEER_MAX_RANKS + a.rank, e);
    const unsigned long long* mine = a.blocks[a.rank] + (size_t)a.channel * PEER_MAX_RANKS + tid;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) < e) {
      if (clock64() - t0 > 120000000000ll) __trap();     // ~60 s: a rank that never arrives must not hang the GPU
    }
  }
  __syncthreads();
  if (a.data && tid < a.n) {
    const double* src = reinterpret_cast<const double*>(a.blocks[a.rank] + PEER_FLAGS_U64) + par * PEER_MAX_RANKS * PEER_RED + tid;
    double s = 0.0;
    for (int p = 0; p < a.world; ++p) s += ld_relaxed_sys_f64(src + (size_t)p * PEER_RED);
    a.data[tid] = s;
  }
}

// rows [0, elems) of `local` (this rank's slice of a replicated array) -> the same slice of every other rank's copy
__global__ void __launch_bounds__(256) k_peer_put(const double* __restrict__ local, double* const* __restrict__ peers,
                                                 long long offset, long long elems, int world, int rank) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < elems; i += stride) {
    const double v = local[i];
    for (int p = 0; p < world; ++p)
      if (p != rank) peers[p][offset + i] = v;
  }
}

long long peer_sync_bytes() { return (long long)(PEER_BLOCK_U64 * 8); }

int launch_peer_sync(const unsigned long long* blocks, unsigned long long* epoch, int world, int rank, int channel,
                     double* data, int n, cudaStream_t st) {
  if (world < 1 || world > PEER_MAX_RANKS || rank < 0 || rank >= world) { set_error("peer_sync: bad world/rank %d/%d", world, rank); return -2; }
  if (channel < 0 || channel >= PEER_CHANNELS) { set_error("peer_sync: channel %d out of range", channel); return -2; }
  if (data && (n < 1 || n > PEER_RED)) { set_error("peer_sync: payload of %d doubles (max %d)", n, PEER_RED); return -2; }
  PeerSyncArgs a;
  a.blocks = reinterpret_cast<unsigned long long* const*>(blocks);
  a.epoch = epoch; a.data = data; a.world = world; a.rank = rank; a.channel = channel; a.n = data ? n : 0;
  k_peer_sync<<<1, 128, 0, st>>>(a);
  return check_launch("peer_sync");
}

int launch_peer_put(const double* local, const unsigned long long* peers, long long offset, long long elems, int world,
                    int rank, cudaStream_t st) {
  if (elems <= 0) return 0;
  long long nb = (elems + 255) / 256;
  if (nb > 1184) nb = 1184;
  k_peer_put<<<(unsigned)nb, 256, 0, st>>>(local, reinterpret_cast<double* const*>(peers), offset, elems, world, rank);
  return check_launch("peer_put");
}

// dst[i] += src[i]: the running sums of the device-side posterior summaries (bnmf_gibbs_optimised.py:182-187)
__global__ void __launch_bounds__(256) k_accumulate(double* __restrict__ dst, const double* __restrict__ src, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] += src[i];
}

// out[e] = mean over it in range(burn_in, n_iter, thinning) of samples[it][e], summed in iteration order (as numpy's
// sum(axis=0) over the stacked samples does)
__global__ void __launch_bounds__(256) k_sample_mean(const double* __restrict__ samples, long long elems, int n_iter, int burn_in,
                                                    int thinning, double* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  int cnt = 0;
  for (int it = burn_in; it < n_iter; it += thinning) ++cnt;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < elems; i += stride) {
    double s = 0.0;
    for (int it = burn_in; it < n_iter; it += thinning) s += samples[(size_t)it * elems + i];
    out[i] = s / (double)cnt;
  }
}

int launch_accumulate(double* dst, const double* src, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  long long nb = (n + 255) / 256;
  if (nb > 2368) nb = 2368;
  k_accumulate<<<(unsigned)nb, 256, 0, st>>>(dst, src, n);
  return check_launch("accumulate");
}

int launch_sample_mean(const double* samples, long long elems, int n_iter, int burn_in, int thinning, double* out,
                       cudaStream_t st) {
  if (elems <= 0) return 0;
  if (thinning < 1 || burn_in < 0 || burn_in >= n_iter) { set_error("sample_mean: empty window (burn_in=%d, thinning=%d, iterations=%d)", burn_in, thinning, n_iter); return -2; }
  long long nb = (elems + 255) / 256;
  if (nb > 2368) nb = 2368;
  k_sample_mean<<<(unsigned)nb, 256, 0, st>>>(samples, elems, n_iter, burn_in, thinning, out);
  return check_launch("sample_mean");
}

}  // namespace bnmtf

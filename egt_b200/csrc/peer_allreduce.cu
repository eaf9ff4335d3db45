// peer_allreduce.cu -- one-shot all-reduce (sum) of the flat weight-gradient buffer over NVLink peer memory.
//
//   reference: the gradient all-reduce tf.distribute.MirroredStrategy performs once per step
//   (lib/training/training_base.py:230-238).  The payload is tiny (17 K floats per block at the headline
//   widths), so the collective is pure latency; a ring / tree through NCCL costs tens of microseconds at
//   8 ranks.  Here every rank publishes its gradient in a symmetric (peer-mapped) buffer, signals each peer
//   with one flag per CTA, and then sums all peers' buffers directly over NVLink / NVSwitch loads.
//
// One launch per step, capturable in a CUDA graph: everything that changes from call to call (the epoch
// and the half of the double buffer in use) lives in device memory, not in kernel arguments.
//   buffers   [world] device pointers to each rank's symmetric buffer of 2 * n floats (double buffered: a
//             rank overwrites half k%2 only after it has passed the flag exchange of call k+1, which every
//             peer enters after finishing its reads of call k)
//   pads      [world] device pointers to each rank's signal pad (uint32): slot [cta][peer] flags, then
//             [PAR_EPOCH + cta] this CTA's call counter.  Pads start zeroed.
#include <stdlib.h>
#include "common.cuh"

namespace egt {

constexpr int PAR_MAXW = 8, PAR_MAXCTA = 32, PAR_EPOCH = PAR_MAXCTA * PAR_MAXW;

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4 *p) {   // peer data: never from a stale L1 line
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) peer_allreduce_kernel(const uint64_t *buffers, const uint64_t *pads, float *grad,
                                                             int n4, int rank, int world, long long timeout_cycles) {
  __shared__ uint32_t s_epoch;
  const int tid = threadIdx.x, cta = blockIdx.x;
  pdl_trigger();
  pdl_wait();
  uint32_t *my_pad = (uint32_t *)pads[rank];
  if (tid == 0) s_epoch = my_pad[PAR_EPOCH + cta];
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const size_t half = (size_t)(epoch & 1u) * (size_t)n4;            // in float4 units
  const int per = (n4 + gridDim.x - 1) / gridDim.x;
  const int i0 = cta * per, i1 = min(n4, i0 + per);
  // 1. publish this rank's slice
  float4 *mine = (float4 *)buffers[rank] + half;
  const float4 *g4 = (const float4 *)grad;
  for (int i = i0 + tid; i < i1; i += blockDim.x) mine[i] = g4[i];
  __threadfence_system();
  __syncthreads();
  // 2. flag every peer, wait for every peer's flag (monotonic epochs: >= tolerates a peer that is already a call ahead)
  if (tid < world) {
    st_release_sys((uint32_t *)pads[tid] + cta * PAR_MAXW + rank, epoch + 1u);
    const uint32_t *slot = my_pad + cta * PAR_MAXW + tid;
    // A peer that is merely late (evaluation or a checkpoint on one rank, a data-loader stall) must be waited
    // for; only a peer that never shows up may end the launch.  timeout_cycles <= 0 waits for ever
    // (EGT_PEER_TIMEOUT_S, default 600 s -- far beyond any step, short of a box-level hang).
    long long t0 = clock64();
    while (ld_acquire_sys(slot) < epoch + 1u) {
      __nanosleep(64);
      if (timeout_cycles > 0 && clock64() - t0 > timeout_cycles) __trap();
    }
  }
  __syncthreads();
  // 3. sum the peers' slices in rank order (bitwise identical result on every rank)
  for (int i = i0 + tid; i < i1; i += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {
      const float4 v = ld_volatile_f4((const float4 *)buffers[r] + half + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    ((float4 *)grad)[i] = acc;
  }
  if (tid == 0) my_pad[PAR_EPOCH + cta] = epoch + 1u;
}

// ---- push protocol ("LL": every 8-byte word carries 4 bytes of data and a 4-byte flag) ----------------------------
// The pull kernel above needs a system-scope fence, a flag round trip and then remote LOADS: 37 us per call on 8 GPUs
// (tools/peer_allreduce_check.py), more than NCCL.  Here every rank STORES its gradient straight into a receive slot
// of every peer, each 16-byte store = {x0, flag, x1, flag}; the receiver polls its own memory until both flags of a
// word show this call's number and adds the slots in rank order (bitwise identical result on every rank).  No fence,
// no separate signal, no remote load: one NVLink store latency per call.
//   buffers[r]   rank r's symmetric buffer: [2 halves][world sources][n2 words of 16 B]  (n2 = n / 2)
//   A half is reused two calls later; a peer can only be two calls ahead after it has received this rank's data of
//   the call in between, which this rank sends after it finished reading the current one.
__device__ __forceinline__ void st_volatile_u4(uint4 *p, uint4 v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_u4(const uint4 *p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) peer_allreduce_push_kernel(const uint64_t *buffers, const uint64_t *pads, float *grad,
                                                                  int n2, int rank, int world, long long timeout_cycles) {
  __shared__ uint32_t s_epoch;
  const int tid = threadIdx.x, cta = blockIdx.x;
  pdl_trigger();
  pdl_wait();
  uint32_t *my_pad = (uint32_t *)pads[rank];
  if (tid == 0) s_epoch = my_pad[PAR_EPOCH + cta];
  __syncthreads();
  const uint32_t epoch = s_epoch, flag = epoch + 1u;                 // never 0: the buffers start zeroed
  const size_t half = (size_t)(epoch & 1u) * (size_t)world * (size_t)n2;
  const int per = (n2 + gridDim.x - 1) / gridDim.x;
  const int i0 = cta * per, i1 = min(n2, i0 + per);
  const float2 *g2 = (const float2 *)grad;
  // 1. push this rank's slice into slot [rank] of every peer (its own copy included)
  for (int i = i0 + tid; i < i1; i += blockDim.x) {
    const float2 v = g2[i];
    const uint4 w = make_uint4(__float_as_uint(v.x), flag, __float_as_uint(v.y), flag);
    for (int r = 0; r < world; ++r) {
      const int dst = (rank + r) % world;                            // spread the targets over the links
      st_volatile_u4((uint4 *)buffers[dst] + half + (size_t)rank * n2 + i, w);
    }
  }
  // 2. collect: poll the local slots, add in rank order
  const uint4 *mine = (const uint4 *)buffers[rank] + half;
  const long long t0 = clock64();
  for (int i = i0 + tid; i < i1; i += blockDim.x) {
    float2 acc = make_float2(0.f, 0.f);
    for (int r = 0; r < world; ++r) {
      const uint4 *slot = mine + (size_t)r * n2 + i;
      uint4 w = ld_volatile_u4(slot);
      while (w.y != flag || w.w != flag) {
        // a peer that is merely late must be waited for; only one that never shows up may end the launch
        if (timeout_cycles > 0 && clock64() - t0 > timeout_cycles) __trap();
        w = ld_volatile_u4(slot);
      }
      acc.x += __uint_as_float(w.x);
      acc.y += __uint_as_float(w.z);
    }
    ((float2 *)grad)[i] = acc;
  }
  if (tid == 0) my_pad[PAR_EPOCH + cta] = epoch + 1u;
}

}  // namespace egt

extern "C" int egt_peer_allreduce(const uint64_t *buffer_ptrs_dev, const uint64_t *signal_pad_ptrs_dev, float *grad,
                                  int64_t n, int rank, int world, void *stream) {
  using namespace egt;
  EGT_REQUIRE(buffer_ptrs_dev && signal_pad_ptrs_dev && grad, EGT_E_ARG, "peer_allreduce: NULL pointer");
  EGT_REQUIRE(world >= 1 && world <= PAR_MAXW && rank >= 0 && rank < world, EGT_E_ARG, "peer_allreduce: world/rank out of range");
  EGT_REQUIRE(n > 0 && n % 4 == 0 && ((uintptr_t)grad & 15) == 0, EGT_E_ALIGN, "peer_allreduce: n must be a multiple of 4 and grad 16-byte aligned");
  const int n4 = (int)(n / 4);
  int ctas = (n4 + 511) / 512;
  if (ctas > PAR_MAXCTA) ctas = PAR_MAXCTA;
  if (ctas < 1) ctas = 1;
  // seconds -> SM clock cycles at the nominal 2 GHz (an upper bound of the real clock: the wait is at least this long)
  static const long long timeout_cycles = [] {
    const char *e = getenv("EGT_PEER_TIMEOUT_S");
    const double s = e ? atof(e) : 600.0;
    return s > 0 ? (long long)(s * 2.0e9) : 0ll;
  }();
  LaunchScope _ls("peer_allreduce_kernel", (cudaStream_t)stream);
  EGT_CHECK_CUDA(launch_pdl(peer_allreduce_kernel, dim3(ctas), dim3(256), 0, (cudaStream_t)stream, buffer_ptrs_dev, signal_pad_ptrs_dev, grad, n4, rank, world, timeout_cycles));
  return EGT_OK;
}

// Push / LL form of the same collective: `buffer_floats` (the size of every rank's symmetric buffer) must be at least
// egt_peer_allreduce_push_floats(n, world).  n % 2 == 0.
extern "C" int64_t egt_peer_allreduce_push_floats(int64_t n, int world) { return 4 * n * (int64_t)world; }

extern "C" int egt_peer_allreduce_push(const uint64_t *buffer_ptrs_dev, const uint64_t *signal_pad_ptrs_dev, float *grad,
                                       int64_t n, int64_t buffer_floats, int rank, int world, void *stream) {
  using namespace egt;
  EGT_REQUIRE(buffer_ptrs_dev && signal_pad_ptrs_dev && grad, EGT_E_ARG, "peer_allreduce_push: NULL pointer");
  EGT_REQUIRE(world >= 1 && world <= PAR_MAXW && rank >= 0 && rank < world, EGT_E_ARG, "peer_allreduce_push: world/rank out of range");
  EGT_REQUIRE(n > 0 && n % 2 == 0 && ((uintptr_t)grad & 7) == 0, EGT_E_ALIGN, "peer_allreduce_push: n must be even and grad 8-byte aligned");
  EGT_REQUIRE(buffer_floats >= egt_peer_allreduce_push_floats(n, world), EGT_E_ARG,
              "peer_allreduce_push: symmetric buffer of %lld floats is smaller than %lld", (long long)buffer_floats,
              (long long)egt_peer_allreduce_push_floats(n, world));
  const int n2 = (int)(n / 2);
  int ctas = (n2 + 255) / 256;                  // one 16-byte word per thread: every store and every poll in flight at once
  if (ctas > PAR_MAXCTA) ctas = PAR_MAXCTA;
  if (ctas < 1) ctas = 1;
  static const long long timeout_cycles = [] {
    const char *e = getenv("EGT_PEER_TIMEOUT_S");
    const double s = e ? atof(e) : 600.0;
    return s > 0 ? (long long)(s * 2.0e9) : 0ll;
  }();
  LaunchScope _ls("peer_allreduce_push_kernel", (cudaStream_t)stream);
  EGT_CHECK_CUDA(launch_pdl(peer_allreduce_push_kernel, dim3(ctas), dim3(256), 0, (cudaStream_t)stream, buffer_ptrs_dev, signal_pad_ptrs_dev, grad, n2, rank, world, timeout_cycles));
  return EGT_OK;
}


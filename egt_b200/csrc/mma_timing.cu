// mma_timing.cu -- micro-benchmark behind DESIGN.md's cost model of small tcgen05.mma instructions.
//
// egt_debug_mma_timing() issues `chains` k-chains of `ksteps` tcgen05.mma (M = 128, K = 16 per instruction, N as
// given) from one warp, commits, waits, and returns the SM clock cycles from the first issue to the completion.
// a_mode: 0 = A from shared memory (K-major, 128B swizzle), 1 = A from tensor memory, 2 = A from shared memory
// MN-major (the transposed products of the backward).  ndst = number of distinct accumulators the chains rotate over
// (1: every chain accumulates into the same columns, i.e. one long dependent chain).
#include "common.cuh"
#include "umma.cuh"

namespace egt {
using namespace umma;

__global__ void __launch_bounds__(128) mma_timing_kernel(int a_mode, int N, int ksteps, int chains, int ndst, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 98304 / 16; i += 128) ((uint4 *)smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t sbase = smem_u32(smem);
  {   // zero the A operand columns in tensor memory
    const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int c = 0; c < 64; c += 8) tmem_st8(((uint32_t)(warp * 32) << 16) + 448 + c, z);
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    tc_fence_after();
    constexpr uint32_t HI_SW = desc_hi(1024, LAYOUT_SW128);
    const uint32_t idesc = idesc_bf16(128, N, a_mode == 2, 0);
    const uint32_t loA = desc_lo(sbase, a_mode == 2 ? 16384u : 16u), loB = desc_lo(sbase + 65536, 16);
    if (ksteps == 8 && ndst == 1) {   // best case: loop-invariant operands, one asm statement per 8-chain
      t0 = clock64();
      for (int c = 0; c < chains; ++c) {
        if (a_mode == 1) MmaChain<8>::ts(0, 448, loB, HI_SW, idesc, 0, 0, 0);
        else if (a_mode == 2) MmaChain<8>::ss(0, loA, HI_SW, loB, HI_SW, idesc, 0, 128, 0);
        else MmaChain<8>::ss(0, loA, HI_SW, loB, HI_SW, idesc, 0, 2, 0);
      }
      const long long ti = clock64();
      mma_commit_w(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
      t1 = clock64();
      if (tid == 0) { out[0] = t1 - t0; out[1] = ti - t0; }
    } else {
    t0 = clock64();
    for (int c = 0; c < chains; ++c) {
      const uint32_t d = (uint32_t)((c % ndst) * N);
      for (int s = 0; s < ksteps; ++s) {
        if (a_mode == 1) MmaChain<1>::ts(d, 448 + 8 * (s & 7), loB + 2 * (s & 3), HI_SW, idesc, s > 0, 0, 0);
        else if (a_mode == 2) MmaChain<1>::ss(d, loA + 128 * (s & 7), HI_SW, loB + 2 * (s & 3), HI_SW, idesc, s > 0, 0, 0);
        else MmaChain<1>::ss(d, loA + 2 * (s & 3) + 1024 * ((s >> 2) & 1), HI_SW, loB + 2 * (s & 3), HI_SW, idesc, s > 0, 0, 0);
      }
    }
    const long long ti = clock64();
    mma_commit_w(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
    if (tid == 0) { out[0] = t1 - t0; out[1] = ti - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(0, 512);
}

}  // namespace egt

extern "C" int egt_debug_mma_timing(int a_mode, int N, int ksteps, int chains, int ndst, long long *cycles_host, void *stream) {
  using namespace egt;
  EGT_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && ksteps >= 1 && chains >= 1 && ndst >= 1 && ndst * N <= 448, EGT_E_ARG,
              "mma_timing: bad arguments");
  long long *dev = nullptr;
  EGT_CHECK_CUDA(cudaMalloc(&dev, 16));
  const int smem = 98304 + 1024;
  EGT_CHECK_CUDA(cudaFuncSetAttribute(mma_timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mma_timing_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a_mode, N, ksteps, chains, ndst, dev);
  EGT_CHECK_CUDA(cudaGetLastError());
  EGT_CHECK_CUDA(cudaMemcpy(cycles_host, dev, 16, cudaMemcpyDeviceToHost));
  EGT_CHECK_CUDA(cudaFree(dev));
  return EGT_OK;
}

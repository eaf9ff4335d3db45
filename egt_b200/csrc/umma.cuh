// umma.cuh -- sm_100a primitives used by the fused kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (TMEM alloc / mma / ld / st / commit / fences) and the shared-memory / instruction
// descriptors of tcgen05.mma.  Bit layouts follow the PTX ISA "tcgen05 matrix / instruction
// descriptor" tables.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace egt {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes (or the hint
// expires) instead of returning after a few dozen cycles.  Without the hint a waiting warp re-polls ~70 times per
// handshake (ncu: 42 % of all issued instructions of the wide kernels were this loop) and, being always eligible,
// takes issue slots from the warps of the same SM sub-partition that do the arithmetic.
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(20000u)
      : "memory");
  return ok;
}
// Bounded wait: a wrong phase must never hang the GPU box.  After ~2 s without completion the kernel
// traps (the launch then reports cudaErrorLaunchFailed instead of hanging).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  int polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(32);
    if ((++polls & 63) == 0 && clock64() - t0 > 4000000000ll) __trap();
  }
}

// ---- proxies / fences -------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMA -----------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(m), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---- register rebalancing between warpgroups (all 4 warps of a warpgroup execute the same one) ----
template <int R>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

// ---- TMEM ------------------------------------------------------------------------------------
// One full warp calls alloc / dealloc.  ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// tcgen05.mma shared-memory matrix descriptor.
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version (1 on sm_100)
//   bits [61,64) layout: 0 none (interleaved 8x16B core matrices), 2 = 128B, 4 = 64B, 6 = 32B swizzle
enum : uint32_t { LAYOUT_NONE = 0, LAYOUT_SW128 = 2, LAYOUT_SW64 = 4, LAYOUT_SW32 = 6 };
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint32_t lo = ((addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
  uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
  return ((uint64_t)hi << 32) | lo;
}
// The same descriptor split into its two words: the high word depends only on (SBO, layout) and is a
// compile-time constant at every call site; the low word is (address | LBO) and advances by bytes >> 4.
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo_bytes) {
  return ((addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint64_t mkdesc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// kind::f16 instruction descriptor: bf16 x bf16 -> f32.
//   [4,6) D format (1 = f32)  [7,10) A format (1 = bf16)  [10,13) B format  [15] A major (1 = MN)
//   [16] B major  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]     (issued by ONE thread)
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-collective forms: EVERY lane of a converged warp calls them with the same (warp-uniform) operands and one
// elected lane issues.  Issuing from inside `if (lane == 0)` makes ptxas treat the descriptors as divergent values:
// each tcgen05.mma then costs a vector -> uniform register round trip inside an election loop (~14 instructions and
// >100 cycles per instruction measured); called from converged code the operands stay in uniform registers.
__device__ __forceinline__ void mma_ss_w(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_w(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}

// ---- chains of tcgen05.mma over consecutive k-steps, ONE inline-asm statement each ----------------------------------
// Every lane of a converged warp calls them (warp-uniform operands); one elected lane issues KS instructions whose
// descriptor low words advance by (a_step, b_step) per k-step (16-byte units; TMEM columns for A in tensor memory).
// One election and no C++-level register shuffling between the instructions: the issue cost per tcgen05.mma drops
// from ~18 SASS instructions (one statement per instruction) to the two adds and the uniform-register moves.
template <int KS> struct MmaChain;

template <> struct MmaChain<1> {
  static __device__ __forceinline__ void ss(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                             uint32_t acc, uint32_t astep, uint32_t bstep) {
    asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b32 al, bl;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %6, 0;\n\tsetp.eq.b32 pt, %6, %6;\n\t"
      "mov.b32 al, %1;\n\tmov.b32 bl, %3;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pa;\n\t"
      "}"
      ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc), "r"(astep), "r"(bstep) : "memory");
  }
  static __device__ __forceinline__ void ts(uint32_t d, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc,
                                             uint32_t astep, uint32_t bstep) {
    asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b32 al, bl;\n\t.reg .b64 db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %5, 0;\n\tsetp.eq.b32 pt, %5, %5;\n\t"
      "mov.b32 al, %1;\n\tmov.b32 bl, %2;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pa;\n\t"
      "}"
      ::"r"(d), "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc), "r"(astep), "r"(bstep) : "memory");
  }
};
template <> struct MmaChain<2> {
  static __device__ __forceinline__ void ss(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                             uint32_t acc, uint32_t astep, uint32_t bstep) {
    asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b32 al, bl;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %6, 0;\n\tsetp.eq.b32 pt, %6, %6;\n\t"
      "mov.b32 al, %1;\n\tmov.b32 bl, %3;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pa;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "}"
      ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc), "r"(astep), "r"(bstep) : "memory");
  }
  static __device__ __forceinline__ void ts(uint32_t d, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc,
                                             uint32_t astep, uint32_t bstep) {
    asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b32 al, bl;\n\t.reg .b64 db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %5, 0;\n\tsetp.eq.b32 pt, %5, %5;\n\t"
      "mov.b32 al, %1;\n\tmov.b32 bl, %2;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pa;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "}"
      ::"r"(d), "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc), "r"(astep), "r"(bstep) : "memory");
  }
};
template <> struct MmaChain<3> {
  static __device__ __forceinline__ void ss(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                             uint32_t acc, uint32_t astep, uint32_t bstep) {
    asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b32 al, bl;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %6, 0;\n\tsetp.eq.b32 pt, %6, %6;\n\t"
      "mov.b32 al, %1;\n\tmov.b32 bl, %3;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pa;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "}"
      ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc), "r"(astep), "r"(bstep) : "memory");
  }
  static __device__ __forceinline__ void ts(uint32_t d, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc,
                                             uint32_t astep, uint32_t bstep) {
    asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b32 al, bl;\n\t.reg .b64 db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %5, 0;\n\tsetp.eq.b32 pt, %5, %5;\n\t"
      "mov.b32 al, %1;\n\tmov.b32 bl, %2;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pa;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "}"
      ::"r"(d), "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc), "r"(astep), "r"(bstep) : "memory");
  }
};
template <> struct MmaChain<4> {
  static __device__ __forceinline__ void ss(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                             uint32_t acc, uint32_t astep, uint32_t bstep) {
    asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b32 al, bl;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %6, 0;\n\tsetp.eq.b32 pt, %6, %6;\n\t"
      "mov.b32 al, %1;\n\tmov.b32 bl, %3;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pa;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "}"
      ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc), "r"(astep), "r"(bstep) : "memory");
  }
  static __device__ __forceinline__ void ts(uint32_t d, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc,
                                             uint32_t astep, uint32_t bstep) {
    asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b32 al, bl;\n\t.reg .b64 db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %5, 0;\n\tsetp.eq.b32 pt, %5, %5;\n\t"
      "mov.b32 al, %1;\n\tmov.b32 bl, %2;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pa;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "}"
      ::"r"(d), "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc), "r"(astep), "r"(bstep) : "memory");
  }
};
template <> struct MmaChain<6> {
  static __device__ __forceinline__ void ss(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                             uint32_t acc, uint32_t astep, uint32_t bstep) {
    asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b32 al, bl;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %6, 0;\n\tsetp.eq.b32 pt, %6, %6;\n\t"
      "mov.b32 al, %1;\n\tmov.b32 bl, %3;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pa;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "}"
      ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc), "r"(astep), "r"(bstep) : "memory");
  }
  static __device__ __forceinline__ void ts(uint32_t d, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc,
                                             uint32_t astep, uint32_t bstep) {
    asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b32 al, bl;\n\t.reg .b64 db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %5, 0;\n\tsetp.eq.b32 pt, %5, %5;\n\t"
      "mov.b32 al, %1;\n\tmov.b32 bl, %2;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pa;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "}"
      ::"r"(d), "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc), "r"(astep), "r"(bstep) : "memory");
  }
};
template <> struct MmaChain<8> {
  static __device__ __forceinline__ void ss(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                             uint32_t acc, uint32_t astep, uint32_t bstep) {
    asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b32 al, bl;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %6, 0;\n\tsetp.eq.b32 pt, %6, %6;\n\t"
      "mov.b32 al, %1;\n\tmov.b32 bl, %3;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pa;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "}"
      ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc), "r"(astep), "r"(bstep) : "memory");
  }
  static __device__ __forceinline__ void ts(uint32_t d, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc,
                                             uint32_t astep, uint32_t bstep) {
    asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b32 al, bl;\n\t.reg .b64 db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %5, 0;\n\tsetp.eq.b32 pt, %5, %5;\n\t"
      "mov.b32 al, %1;\n\tmov.b32 bl, %2;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pa;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "add.u32 al, al, %6;\n\tadd.u32 bl, bl, %7;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, pt;\n\t"
      "}"
      ::"r"(d), "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc), "r"(astep), "r"(bstep) : "memory");
  }
};


// 8-step chain whose k-steps >= nsteps are predicated off (contraction over the query rows of a partly filled tile)
__device__ __forceinline__ void mma_chain8_n(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                             uint32_t acc, uint32_t astep, uint32_t bstep, uint32_t nsteps) {
  asm volatile(
      "{\n\t.reg .pred pe, pa, pt, pk;\n\t.reg .b32 al, bl;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %6, 0;\n\tsetp.eq.b32 pt, %6, %6;\n\t"
      "mov.b32 al, %1;\n\tmov.b32 bl, %3;\n\t"
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pa;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "setp.gt.u32 pk, %9, 1;\n\tand.pred pk, pk, pe;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "setp.gt.u32 pk, %9, 2;\n\tand.pred pk, pk, pe;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "setp.gt.u32 pk, %9, 3;\n\tand.pred pk, pk, pe;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "setp.gt.u32 pk, %9, 4;\n\tand.pred pk, pk, pe;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "setp.gt.u32 pk, %9, 5;\n\tand.pred pk, pk, pe;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "setp.gt.u32 pk, %9, 6;\n\tand.pred pk, pk, pe;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "add.u32 al, al, %7;\n\tadd.u32 bl, bl, %8;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "setp.gt.u32 pk, %9, 7;\n\tand.pred pk, pk, pe;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pt;\n\t"
      "}"
      ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc), "r"(astep), "r"(bstep), "r"(nsteps) : "memory");
}

// All tcgen05.mma issued so far by this thread arrive (once) on `bar` when they complete.
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM <-> registers, shape 32x32b: thread t of warp w touches lane 32*(w%4)+t, N consecutive columns.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t *r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t *r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t *r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, const uint32_t *r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(r[0]), "r"(r[1]) : "memory");
}

// ---- fast math (MUFU) and packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2) -----------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid(x) = 0.5 tanh(0.5 x) + 0.5: one MUFU op instead of ex2 + rcp
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ffma2(float2 &acc, const float2 a, const float2 b) {   // acc += a * b (lane-wise)
  unsigned long long &c = reinterpret_cast<unsigned long long &>(acc);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(c)
      : "l"(reinterpret_cast<const unsigned long long &>(a)), "l"(reinterpret_cast<const unsigned long long &>(b)));
}
__device__ __forceinline__ void fadd2(float2 &acc, const float2 a) {
  unsigned long long &c = reinterpret_cast<unsigned long long &>(acc);
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(c) : "l"(reinterpret_cast<const unsigned long long &>(a)));
}

// ---- small helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {   // element 2j in the low half
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// byte offset of element (row, col) in a K-major / MN-major 128B-swizzled tile whose rows are 128 bytes
// (64 bf16): 16-byte chunk index XOR (row & 7).  The tile base must be 1024-byte aligned.
__device__ __forceinline__ uint32_t sw128_off(uint32_t row, uint32_t col_elem) {
  return row * 128u + ((((col_elem >> 3) ^ row) & 7u) << 4) + ((col_elem & 7u) << 1);
}

}  // namespace umma
}  // namespace egt

// sm_100a building blocks written as inline PTX: mbarrier, TMA (tiled + gather4 loads),
// tcgen05 (TMEM allocation, UMMA issue/commit, TMEM loads) and the shared-memory /
// instruction descriptors they take.  Bit layouts follow the PTX ISA "tcgen05" chapter
// (shared-memory descriptor, instruction descriptor for .kind::f16).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define TC_D __device__ __forceinline__

TC_D uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------
TC_D void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
TC_D void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
TC_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes)
               : "memory");
}
TC_D void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: a warp whose phase is not complete is PARKED by the hardware (it wakes when
// the phase completes, or after the hint) instead of re-issuing the poll.  Without the hint the poll returns
// after ~70 cycles; eight waiting epilogue warps then take two thirds of the issue slots of the scheduler
// they share with the single UMMA-issuing thread, which paces the whole CTA (measured in k_topk_tc: 190 cycles
// per UMMA against 64 on the tensor pipe).
TC_D bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(tc_smem_u32(bar)), "r"(parity), "r"(2000u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol error must not hang the GPU box - after ~2 s of polling the
// kernel traps (the host sees a launch failure instead of a dead GPU).
TC_D void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
#pragma unroll 1
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// generic-proxy writes to shared memory -> visible to the async proxy (TMA / UMMA reads)
TC_D void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMA -----------------------------------------------------------------------------
TC_D void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tile: box {c0 .. , c1 ..} of the tensor map -> smem, completes `bytes` on the mbarrier
TC_D void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(tc_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(tc_smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// gather4: four rows r0..r3 of a 2-D tensor, columns c0 .. c0+box0, -> 4 consecutive smem rows
TC_D void tma_gather4(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t r0, int32_t r1,
                      int32_t r2, int32_t r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(tc_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(tc_smem_u32(bar)), "r"(c0), "r"(r0),
      "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

// ---- tcgen05: tensor memory ------------------------------------------------------------
// whole-warp, .sync.aligned: allocate `ncols` (power of two >= 32) columns; address -> *smem_slot
TC_D void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
TC_D void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
TC_D void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
TC_D void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], issued by ONE thread; accumulate=0 overwrites D
TC_D void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same with the descriptors handed over as (low, high) 32-bit words: an issue loop that advances the start
// address with 32-bit adds keeps the per-UMMA instruction count of its single thread small
TC_D void umma_f16_lohi(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-convergent issue: EVERY lane of the issuing warp executes the call with the same operands and the
// instruction itself is predicated on elect.sync.  Inside an `if (lane == 0)` region the compiler cannot prove
// the operands warp-uniform and wraps each UMMA into an ELECT / BRA.U.ANY emulation loop (~25 dependent
// instructions, ~140 cycles per UMMA for the lone thread: twice what the tensor pipe needs for M = N = 128);
// in convergent code the descriptors sit in uniform registers and a UMMA costs a handful of instructions.
TC_D void umma_f16_lohi_elect(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
      : "memory");
}
TC_D void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(tc_smem_u32(bar))
      : "memory");
}
// arrive on an mbarrier once all previously issued UMMAs of this thread have completed
TC_D void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive columns: thread t of the warp receives row (lane base + t)
TC_D void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
TC_D void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors -----------------------------------------------------------------------
enum { TC_SWZ_NONE = 0, TC_SWZ_128B = 2, TC_SWZ_64B = 4, TC_SWZ_32B = 6 };

// Shared-memory matrix descriptor (64 bit): start address, leading / stride byte offsets
// (all >> 4), descriptor version 1 (Blackwell), swizzle mode in bits [61,64).
TC_D uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t swizzle) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(swizzle & 7u) << 61;
  return d;
}
// Instruction descriptor of .kind::f16 with BF16 inputs and FP32 accumulation.
// a_mn / b_mn: 0 = K-major operand, 1 = MN-major operand.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

TC_D bool tc_elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- host: tensor-map encoding through the driver entry point (no link-time libcuda) ----
typedef CUresult (*tc_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                       CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                       CUtensorMapFloatOOBfill);
tc_encode_tiled_fn tc_get_encode_tiled();
// 2-D bf16 row-major tensor [rows, cols] (cols contiguous, row pitch `pitch_bytes`),
// box = {box_cols, box_rows}; swizzle: 0 none, 1 32B, 2 64B, 3 128B.  Returns 0 on success.
int tc_make_tensor_map_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                            uint32_t box_cols, uint32_t box_rows, int swizzle);

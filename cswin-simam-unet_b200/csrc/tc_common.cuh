// Raw-PTX building blocks for the tcgen05 / TMEM / TMA kernels (sm_100a).
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix / instruction descriptor" tables (the
// same fields CUTLASS names UMMA::SmemDescriptor and UMMA::InstrDescriptor).
#pragma once

#include <cuda.h>  // CUtensorMap (types only; the driver entry point is resolved at run time)

#include "common.cuh"

namespace csb200 {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// suspendTimeHint of try_wait: how long the thread may stay suspended inside ONE try_wait before it returns false
// and the loop retries (it wakes as soon as the phase completes).  A warp that retries at the default, short limit
// takes issue slots from the math warps of its scheduler (three of the 16 warps of the attention kernels are
// waiting on barriers almost all the time).
#ifndef CSB_MBAR_HINT_NS
#define CSB_MBAR_HINT_NS 2000u
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(CSB_MBAR_HINT_NS)
      : "memory");
}

// ---- TMA ----------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 4-D tiled load global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (TMA / tensor core reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMEM ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(slot_in_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp gets lane (warp%4)*32 + t
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ---- UMMA ---------------------------------------------------------------------------------------
// Shared-memory matrix descriptor for 64-byte-swizzled tiles whose rows are 64 B (32 bf16) wide:
// 8-row groups are 512 B apart (SBO).  Valid both as a K-major operand (rows = M/N index, the 64 B
// row = 32 K elements; advance 32 B per K=16 step) and as an MN-major operand (rows = K index, the
// 64 B row = 32 MN elements; advance 1024 B per K=16 step).
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);  // start address, bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                    // LBO (unused for swizzled layouts)
  d |= static_cast<uint64_t>(512 >> 4) << 32;             // SBO = 512 B, bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                    // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(4) << 61;                    // layout type SWIZZLE_64B
  return d;
}
// Shared-memory matrix descriptor for a 128-byte-swizzled MN-major tile: rows are the K index and
// hold 64 MN elements (128 B); 8-row groups are 1024 B apart (SBO); successive 64-element MN blocks
// are `lbo_bytes` apart.  Advance 2048 B per K=16 step.
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;  // layout type SWIZZLE_128B
  return d;
}
// Instruction descriptor, kind::f16, bf16 x bf16 -> fp32, dense, M = 128.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                                 // D format F32
         | (1u << 7) | (1u << 10)                  // A, B format BF16
         | (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16)
         | (static_cast<uint32_t>(n >> 3) << 17)   // N / 8
         | (static_cast<uint32_t>(128 >> 4) << 24);  // M / 16
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- cheap issue path --------------------------------------------------------------------------
// One MMA-issuing thread executes every descriptor instruction serially, so the issue stream itself
// bounds the pipeline when tiles are small (measured: ~490 SASS instructions per backward iteration,
// ~8 cycles each, r1_attn_bwd_tc_s3).  Two things keep it short: (1) the whole warp runs the loop and
// one lane is chosen with elect.sync, which ptxas recognises as single-thread code — a plain
// `if (lane == 0)` makes it wrap every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop;
// (2) descriptors are (lo, hi) register pairs whose hi word is a constant and whose lo word advances
// by an immediate (shared memory is < 256 KB, so address >> 4 never leaves its 14-bit field).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  __syncwarp();  // lanes leave the mbarrier spin loops at different times
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "@p mov.s32 %0, 1;\n\t}"
      : "+r"(pred));
  return pred != 0;
}
constexpr uint32_t DESC_HI_SW64 = (512u >> 4) | (1u << 14) | (4u << 29);      // SBO 512, version, SWIZZLE_64B
constexpr uint32_t DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);    // SBO 1024, version, SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo_sw64(uint32_t smem_addr) { return (smem_addr >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_lo_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (smem_addr >> 4) | ((lbo_bytes >> 4) << 16);
}
// D[tmem] (+)= A[smem] * B[smem], descriptors as (lo, hi) words
__device__ __forceinline__ void umma_ss2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                         uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when they complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for a PAIR of non-positive arguments on the FMA pipes (no MUFU): the softmax sweeps issue one ex2 per 128
// tensor-core FLOPs at head dimension 32, and MUFU runs 16 lanes per clock per SM — a quarter of the tensor peak —
// so part of the exponentials go through here instead.  Cody-Waite: x = n + f with n = round(x) (magic-number add),
// 2^f on [-0.5, 0.5] by a cubic (max relative error 7.5e-5, 25x below the rounding of the bf16 probabilities),
// 2^n by adding n to the exponent field.  ~10 issue slots per pair on packed fp32 (fma.rn.f32x2).
__device__ __forceinline__ void ex2_poly_pair(f2_t x, float& p0, float& p1) {
  float x0, x1;
  f2_split(x, x0, x1);
  x = f2_make(fmaxf(x0, -125.f), fmaxf(x1, -125.f));  // 2^-125 ~ 0; keeps the exponent arithmetic in range
  const f2_t t = f2_add(x, f2_splat(12582912.f));     // 1.5 * 2^23: the low mantissa bits of t are round(x)
  const f2_t f = f2_add(x, f2_fma(t, f2_splat(-1.f), f2_splat(12582912.f)));  // x - round(x)
  f2_t p = f2_fma(f, f2_splat(0.05517164617776871f), f2_splat(0.2426111251115799f));
  p = f2_fma(p, f, f2_splat(0.6932609677314758f));
  p = f2_fma(p, f, f2_splat(0.9999280571937561f));
  float t0, t1, q0, q1;
  f2_split(t, t0, t1);
  f2_split(p, q0, q1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

}  // namespace tc

// host: cuTensorMapEncodeTiled resolved through the runtime (libcsb200.so does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();
void ensure_context();  // binds the primary context to the calling host thread (once per thread)
struct StripeGeom;
// (channel, x, y, image) map over one token-major operand of a branch; box = (32, bx, by, 1),
// 64-byte swizzle.  Defined in stripe_attn_tc.cu.
int tc_make_map(CUtensorMap* m, const void* base, const StripeGeom& g, int64_t sb, int64_t sl, int bx,
                int by);

}  // namespace csb200

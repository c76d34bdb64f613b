// Shared device/host helpers for the csb200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/csb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "csb200 kernels are written for sm_100a (B200) only"
#endif

namespace csb200 {

// ---- host side: error text + launch accounting ---------------------------------------------------
extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CSB200_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return CSB200_OK;
}

#define CSB200_CUDA(call)                                                                   \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return ::csb200::fail(CSB200_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__));     \
  } while (0)

// ---- host side: per-device launch-setup caches (defined in stripe_attn.cu) ----------------------------
// All keyed by the CURRENT device and mutex-guarded: cudaFuncSetAttribute applies to the current device
// only, a process may drive several GPUs, and the main thread and autograd's backward thread both launch.
int device_sm_count();                                  // SM count of the current device (<= 0 on error)
cudaError_t opt_in_smem(const void* func, int bytes);   // MaxDynamicSharedMemorySize >= bytes, once per (device, func)
bool memo_get(const void* key, int tag, int* val);      // small integer memo, e.g. occupancy query results
void memo_put(const void* key, int tag, int val);

// ---- device side ------------------------------------------------------------------------------
template <typename T>
struct Vec16;  // 16-byte vector of T
template <>
struct Vec16<float> {
  static constexpr int N = 4;
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
};

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// streaming 128-bit accesses: read-once / write-once data should not pollute L1
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// unpack a 16-byte vector into fp32 lanes and back
template <typename T>
__device__ __forceinline__ void unpack(const uint4& u, float (&f)[Vec16<T>::N]);
template <>
__device__ __forceinline__ void unpack<float>(const uint4& u, float (&f)[4]) {
  f[0] = __uint_as_float(u.x);
  f[1] = __uint_as_float(u.y);
  f[2] = __uint_as_float(u.z);
  f[3] = __uint_as_float(u.w);
}
template <>
__device__ __forceinline__ void unpack<__nv_bfloat16>(const uint4& u, float (&f)[8]) {
  // bf16 -> fp32 is a 16-bit shift
  f[0] = __uint_as_float(u.x << 16);
  f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16);
  f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16);
  f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16);
  f[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <typename T>
__device__ __forceinline__ uint4 pack(const float (&f)[Vec16<T>::N]);
template <>
__device__ __forceinline__ uint4 pack<float>(const float (&f)[4]) {
  return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                    __float_as_uint(f[3]));
}
template <>
__device__ __forceinline__ uint4 pack<__nv_bfloat16>(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}

// ---- packed fp32x2 arithmetic (FFMA2: two fp32 lanes per issue slot, sm_100) ----------------------
// For kernels bound by instruction issue rather than by HBM: a 32-bit word holding two bf16 values IS
// a pair of fp32 values after (w << 16, w & 0xffff0000), so whole per-element chains run on pairs.
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_make(float lo, float hi) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f2_t f2_splat(float v) { return f2_make(v, v); }
__device__ __forceinline__ void f2_split(f2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2_t f2_from_bf16x2(uint32_t w) {
  return f2_make(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) {
  f2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f2_t f2_mul(f2_t a, f2_t b) {
  f2_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) {
  f2_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum of p[b * stride] over b = lane, lane + 32, ... < n followed by a warp reduction: the last stage of the
// deterministic two-stage reductions (per-CTA partials -> one warp per output).  Four independent
// accumulators so the L2 round trips overlap; fixed order -> deterministic.  (Measured: these last-stage
// kernels stay at 4-5 us each either way — they are launch-latency bound, ~210 of them per train step;
// folding them into their producers is listed in DESIGN.md section 9.)
__device__ __forceinline__ float strided_partial_sum(const float* __restrict__ p, int n, int64_t stride, int lane) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int b = lane;
  for (; b + 96 < n; b += 128) {
    a0 += p[(int64_t)b * stride];
    a1 += p[(int64_t)(b + 32) * stride];
    a2 += p[(int64_t)(b + 64) * stride];
    a3 += p[(int64_t)(b + 96) * stride];
  }
  for (; b < n; b += 32) a0 += p[(int64_t)b * stride];
  return warp_sum((a0 + a1) + (a2 + a3));
}

// sigmoid with the precise expf (fp32 parity target is 1e-5 relative)
__device__ __forceinline__ float sigmoidf_precise(float y) { return 1.f / (1.f + expf(-y)); }

// cluster helpers (raw PTX; cooperative_groups would do the same)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::
                   : "memory");
}
// read a float from the shared memory of CTA `rank` of this cluster (DSMEM)
__device__ __forceinline__ float dsmem_ld_f32(const float* local_smem_ptr, uint32_t rank) {
  uint32_t local = static_cast<uint32_t>(__cvta_generic_to_shared(local_smem_ptr));
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote));
  return v;
}

}  // namespace csb200

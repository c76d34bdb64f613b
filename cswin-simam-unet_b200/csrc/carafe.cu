// Fused CARAFE reassembly (forward / backward) — the content-aware upsampling of the decoder,
// reference CARAFE.forward C:406-431 (and CARAFE4, C:455-480):
//
//   kernel = softmax_over_9_taps(pixel_shuffle(encoder_out, up))          (C:408-411)
//   out[b, c, h*up+dy, w*up+dx] = sum_tap kernel[b, tap, h*up+dy, w*up+dx] * x[b, c, h+ky-1, w+kx-1]
//                                                                          (C:413-431, zero padding)
//
// The reference materialises the pixel-shuffled logits, their softmax (fp32, 9 x the OUTPUT
// resolution: 302 MB at 512^2 / batch 32), an unfold of x (9 x the activation) and runs a batched
// 9 x up^2 matmul plus two pixel shuffles; autograd replays all of it (col2im, im2col, bmm: ~12 ms of
// a 42 ms train step, profiles/r1_step_profile.md).  Here everything is index arithmetic on
// channels-last tensors: encoder logit (tap, dy, dx) of source pixel (h, w) is channel
// tap*up^2 + dy*up + dx of the encoder output; the 9 logits of an output pixel are softmaxed in
// registers and applied to the 3x3 neighbourhood of the low-resolution features.  HBM traffic is the
// inputs once + the output once (+ the 9 weights per output pixel saved for backward).
//
// Layouts (all channels-last == token-major): low [B][H][W][Co], enc [B][H][W][9*up*up],
// out / grad_out [B][H*up][W*up][Co], wt [B][H*up][W*up][9] (type T).  Co == 1 or Co % 8 == 0.

#include "common.cuh"

namespace csb200 {
namespace {

template <typename T>
struct CExp {
  static __device__ __forceinline__ float f(float x) { return expf(x); }
};
template <>
struct CExp<__nv_bfloat16> {
  static __device__ __forceinline__ float f(float x) { return __expf(x); }
};

struct CGeom {
  int B, H, W, Co, up;
};

// load CG consecutive channels as fp32 (CG = 1 or 8)
template <typename T, int CG>
__device__ __forceinline__ void ldc(const T* p, float (&f)[CG]) {
  if constexpr (CG == 1) {
    f[0] = to_f32(__ldg(p));
  } else if constexpr (sizeof(T) == 2) {
    float t[8];
    unpack<__nv_bfloat16>(__ldg(reinterpret_cast<const uint4*>(p)), t);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = t[i];
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
}
template <typename T, int CG>
__device__ __forceinline__ void stc(T* p, const float (&f)[CG]) {
  if constexpr (CG == 1) {
    p[0] = from_f32<T>(f[0]);
  } else if constexpr (sizeof(T) == 2) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                              pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  } else {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
}

// forward: one thread per (output pixel, group of CG channels); channel groups fastest
template <typename T, int CG>
__global__ void __launch_bounds__(256)
    carafe_fwd_kernel(CGeom g, const T* __restrict__ low, const T* __restrict__ enc,
                      T* __restrict__ out, T* __restrict__ wt) {
  const int ngrp = g.Co / CG, OH = g.H * g.up, OW = g.W * g.up, u2 = g.up * g.up;
  const int64_t total = (int64_t)g.B * OH * OW * ngrp;
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int cgp = (int)(idx % ngrp);
  int64_t pix = idx / ngrp;
  const int ox = (int)(pix % OW);
  pix /= OW;
  const int oy = (int)(pix % OH), b = (int)(pix / OH);
  const int h = oy / g.up, w = ox / g.up, sub = (oy % g.up) * g.up + ox % g.up;
  const T* e = enc + (((int64_t)b * g.H + h) * g.W + w) * (9 * u2) + sub;
  float k[9], mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    k[t] = to_f32(__ldg(e + t * u2));
    mx = fmaxf(mx, k[t]);
  }
  float z = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    k[t] = CExp<T>::f(k[t] - mx);
    z += k[t];
  }
  const float iz = 1.f / z;
  float acc[CG];
#pragma unroll
  for (int c = 0; c < CG; ++c) acc[c] = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    k[t] *= iz;
    const int hy = h + t / 3 - 1, wx = w + t % 3 - 1;
    if (hy < 0 || hy >= g.H || wx < 0 || wx >= g.W) continue;  // zero padding, C:415-417
    float f[CG];
    ldc<T, CG>(low + (((int64_t)b * g.H + hy) * g.W + wx) * g.Co + cgp * CG, f);
#pragma unroll
    for (int c = 0; c < CG; ++c) acc[c] = fmaf(k[t], f[c], acc[c]);
  }
  const int64_t opix = ((int64_t)b * OH + oy) * OW + ox;
  stc<T, CG>(out + opix * g.Co + cgp * CG, acc);
  if (cgp == 0 && wt != nullptr) {
#pragma unroll
    for (int t = 0; t < 9; ++t) wt[opix * 9 + t] = from_f32<T>(k[t]);
  }
}

// backward, part 1: gradient of the encoder logits.  LPP lanes share an output pixel and split its
// channels; d_wt[tap] = sum_c grad_out[c] * x[neighbour_tap][c], then the softmax Jacobian.
template <typename T, int CG, int LPP>
__global__ void __launch_bounds__(256)
    carafe_bwd_enc_kernel(CGeom g, const T* __restrict__ low, const T* __restrict__ wt,
                          const T* __restrict__ gout, T* __restrict__ d_enc) {
  const int OH = g.H * g.up, OW = g.W * g.up, u2 = g.up * g.up;
  const int64_t npix = (int64_t)g.B * OH * OW;
  const int64_t gid = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int sl = (int)(gid % LPP);
  int64_t pix = gid / LPP;
  const bool valid = pix < npix;
  if (!valid) pix = npix - 1;  // keep the lane alive for the shuffles
  const int64_t opix = pix;
  const int ox = (int)(pix % OW);
  pix /= OW;
  const int oy = (int)(pix % OH), b = (int)(pix / OH);
  const int h = oy / g.up, w = ox / g.up, sub = (oy % g.up) * g.up + ox % g.up;
  float dw[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) dw[t] = 0.f;
  for (int c0 = sl * CG; c0 < g.Co; c0 += LPP * CG) {
    float go[CG];
    ldc<T, CG>(gout + opix * g.Co + c0, go);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int hy = h + t / 3 - 1, wx = w + t % 3 - 1;
      if (hy < 0 || hy >= g.H || wx < 0 || wx >= g.W) continue;
      float f[CG];
      ldc<T, CG>(low + (((int64_t)b * g.H + hy) * g.W + wx) * g.Co + c0, f);
#pragma unroll
      for (int c = 0; c < CG; ++c) dw[t] = fmaf(go[c], f[c], dw[t]);
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) dw[t] += __shfl_xor_sync(0xffffffffu, dw[t], o);
  if (sl != 0 || !valid) return;
  float k[9], dot = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    k[t] = to_f32(__ldg(wt + opix * 9 + t));
    dot = fmaf(k[t], dw[t], dot);
  }
  T* de = d_enc + (((int64_t)b * g.H + h) * g.W + w) * (9 * u2) + sub;
#pragma unroll
  for (int t = 0; t < 9; ++t) de[t * u2] = from_f32<T>(k[t] * (dw[t] - dot));
}

// backward, part 2: gradient of the low-resolution features (gather form, no atomics):
// x[h', w'] was tap (ky, kx) of source pixel (h'-ky+1, w'-kx+1), for each of its up^2 outputs.
template <typename T, int CG>
__global__ void __launch_bounds__(256)
    carafe_bwd_low_kernel(CGeom g, const T* __restrict__ wt, const T* __restrict__ gout,
                          T* __restrict__ d_low) {
  const int ngrp = g.Co / CG, OH = g.H * g.up, OW = g.W * g.up;
  const int64_t total = (int64_t)g.B * g.H * g.W * ngrp;
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int cgp = (int)(idx % ngrp);
  int64_t pix = idx / ngrp;
  const int wq = (int)(pix % g.W);
  pix /= g.W;
  const int hq = (int)(pix % g.H), b = (int)(pix / g.H);
  float acc[CG];
#pragma unroll
  for (int c = 0; c < CG; ++c) acc[c] = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int h = hq - t / 3 + 1, w = wq - t % 3 + 1;  // the source pixel that saw us as tap t
    if (h < 0 || h >= g.H || w < 0 || w >= g.W) continue;
    for (int dy = 0; dy < g.up; ++dy)
      for (int dx = 0; dx < g.up; ++dx) {
        const int64_t opix = ((int64_t)b * OH + h * g.up + dy) * OW + w * g.up + dx;
        const float k = to_f32(__ldg(wt + opix * 9 + t));
        float go[CG];
        ldc<T, CG>(gout + opix * g.Co + cgp * CG, go);
#pragma unroll
        for (int c = 0; c < CG; ++c) acc[c] = fmaf(k, go[c], acc[c]);
      }
  }
  stc<T, CG>(d_low + (((int64_t)b * g.H + hq) * g.W + wq) * g.Co + cgp * CG, acc);
}

// ------------------------------------------------------------------------------------------------
// Tiled kernels (up = 2 or 4, W a multiple of 16): a CTA owns TW source pixels of one image row.  The
// per-pixel thread mapping above touches the encoder logits, the saved weights and the logit gradients
// as 2-byte elements at a stride of up^2 (a 32-byte sector for every 2..8 useful bytes) and ran at
// 16-18 % of the HBM roofline; here every global access of those tensors is a contiguous 16-byte
// vector and the (tap, sub-position) <-> (sub-position, tap) transposes happen in shared memory.
//   forward : enc rows -> smem (fp32) -> softmax over the 9 taps in place -> weights out (coalesced)
//             -> 3x3 reassembly from smem weights, low-resolution features through L1
//   backward (logits): per output pixel d_wt and the softmax Jacobian as before, results staged in smem
//             in the encoder layout and written as whole rows.
// ------------------------------------------------------------------------------------------------
template <typename T, int CG, int UP>
__global__ void __launch_bounds__(256)
    carafe_fwd_tiled(CGeom g, int TW, const T* __restrict__ low, const T* __restrict__ enc,
                     T* __restrict__ out, T* __restrict__ wt) {
  constexpr int U2 = UP * UP, NL = 9 * U2, PS = NL + 1, VE = Vec16<T>::N;
  extern __shared__ __align__(16) uint8_t s_raw[];
  T* s_w = reinterpret_cast<T*>(s_raw);                                    // [UP][TW*UP*9]: weights, wt layout
  float* s_e = reinterpret_cast<float*>(s_raw + (size_t)TW * NL * sizeof(T));  // [TW][PS]: [pixel][tap][sub]
  const int tiles = g.W / TW;
  const int w0 = (blockIdx.x % tiles) * TW;
  const int h = (blockIdx.x / tiles) % g.H, b = blockIdx.x / (tiles * g.H);
  const int OH = g.H * UP, OW = g.W * UP;
  // ---- encoder logits of the tile: TW * NL contiguous elements
  {
    const uint4* src = reinterpret_cast<const uint4*>(enc + (((int64_t)b * g.H + h) * g.W + w0) * NL);
    for (int v = threadIdx.x; v < TW * NL / VE; v += 256) {
      float f[VE];
      unpack<T>(ld_stream(src + v), f);
      int pl = (v * VE) / NL, l = (v * VE) % NL;
#pragma unroll
      for (int e = 0; e < VE; ++e) {
        s_e[pl * PS + l] = f[e];
        if (++l == NL) {
          l = 0;
          ++pl;
        }
      }
    }
  }
  __syncthreads();
  // ---- softmax over the 9 taps of every (pixel, sub-position): fp32 in place for the reassembly, and
  //      rounded to T in the layout of the saved-weights tensor ([dy][pixel][dx][tap]) for the copy-out
  for (int i = threadIdx.x; i < TW * U2; i += 256) {
    const int pl = i / U2, sub = i % U2;
    float* e = s_e + pl * PS + sub;
    float k[9], mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      k[t] = e[t * U2];
      mx = fmaxf(mx, k[t]);
    }
    float z = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      k[t] = CExp<T>::f(k[t] - mx);
      z += k[t];
    }
    const float iz = 1.f / z;
    T* wrow = s_w + (sub / UP) * (TW * UP * 9) + (pl * UP + sub % UP) * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float kv = k[t] * iz;
      e[t * U2] = kv;
      wrow[t] = from_f32<T>(kv);
    }
  }
  __syncthreads();
  // ---- weights for backward: [B][OH][OW][9], one contiguous run of TW*UP*9 elements per output row
  if (wt != nullptr) {
    const int runv = TW * UP * 9 / VE;  // 16-byte vectors per run
    const uint4* srcv = reinterpret_cast<const uint4*>(s_w);
#pragma unroll
    for (int dy = 0; dy < UP; ++dy) {
      uint4* dst = reinterpret_cast<uint4*>(wt + (((int64_t)b * OH + h * UP + dy) * OW + (int64_t)w0 * UP) * 9);
      for (int v = threadIdx.x; v < runv; v += 256) dst[v] = srcv[dy * runv + v];
    }
  }
  // ---- reassembly
  if constexpr (CG == 1) {
    // Co == 1: one thread per (pixel, output row of the pixel), UP outputs each
    for (int i = threadIdx.x; i < TW * UP; i += 256) {
      const int pl = i / UP, dy = i % UP, w = w0 + pl;
      float lo[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int hy = h + t / 3 - 1, wx = w + t % 3 - 1;
        lo[t] = (hy < 0 || hy >= g.H || wx < 0 || wx >= g.W)
                    ? 0.f
                    : to_f32(__ldg(low + ((int64_t)b * g.H + hy) * g.W + wx));
      }
      float o[UP];
#pragma unroll
      for (int dx = 0; dx < UP; ++dx) {
        const float* e = s_e + pl * PS + dy * UP + dx;
        float a = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) a = fmaf(e[t * U2], lo[t], a);
        o[dx] = a;
      }
      T* dst = out + ((int64_t)b * OH + h * UP + dy) * OW + (int64_t)w * UP;  // UP elements, UP-aligned
      if constexpr (sizeof(T) == 2) {
        if constexpr (UP == 4) *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
        else *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(o[0], o[1]);
      } else {
        if constexpr (UP == 4) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        else *reinterpret_cast<float2*>(dst) = make_float2(o[0], o[1]);
      }
    }
  } else {
    // one thread per (pixel, group of 8 channels, output row of the pixel): UP accumulators of 8 channels,
    // the neighbour vectors are consumed tap by tap (few registers -> enough CTAs per SM to hide the loads)
    const int ngrp = g.Co / CG;
    for (int i = threadIdx.x; i < TW * ngrp * UP; i += 256) {
      const int cgp = i % ngrp, r = i / ngrp, dy = r % UP, pl = r / UP, w = w0 + pl;
      float acc[UP][CG];
#pragma unroll
      for (int dx = 0; dx < UP; ++dx)
#pragma unroll
        for (int c = 0; c < CG; ++c) acc[dx][c] = 0.f;
      const float* e = s_e + pl * PS + dy * UP;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int hy = h + t / 3 - 1, wx = w + t % 3 - 1;
        if (hy < 0 || hy >= g.H || wx < 0 || wx >= g.W) continue;
        float lo[CG];
        ldc<T, CG>(low + (((int64_t)b * g.H + hy) * g.W + wx) * g.Co + cgp * CG, lo);
#pragma unroll
        for (int dx = 0; dx < UP; ++dx) {
          const float k = e[t * U2 + dx];
#pragma unroll
          for (int c = 0; c < CG; ++c) acc[dx][c] = fmaf(k, lo[c], acc[dx][c]);
        }
      }
      const int64_t opix = ((int64_t)b * OH + h * UP + dy) * OW + (int64_t)w * UP;
#pragma unroll
      for (int dx = 0; dx < UP; ++dx) stc<T, CG>(out + (opix + dx) * g.Co + cgp * CG, acc[dx]);
    }
  }
}

template <typename T, int CG, int LPP, int UP>
__global__ void __launch_bounds__(256)
    carafe_bwd_enc_tiled(CGeom g, int TW, const T* __restrict__ low, const T* __restrict__ wt,
                         const T* __restrict__ gout, T* __restrict__ d_enc) {
  constexpr int U2 = UP * UP, NL = 9 * U2, VE = Vec16<T>::N, IPB = 256 / LPP;  // items per CTA step
  extern __shared__ __align__(16) uint8_t s_raw[];
  T* s_d = reinterpret_cast<T*>(s_raw);  // [TW][NL]: the logit gradients of the tile, encoder layout
  const int tiles = g.W / TW;
  const int w0 = (blockIdx.x % tiles) * TW;
  const int h = (blockIdx.x / tiles) % g.H, b = blockIdx.x / (tiles * g.H);
  const int OH = g.H * UP, OW = g.W * UP;
  const int sl = threadIdx.x % LPP, slot = threadIdx.x / LPP;
  // items in output-row order (dy, pixel, dx): consecutive items are consecutive output pixels.
  // TW is a power of two: row = item >> log2(TW * UP)
  const int row_items = TW * UP, row_shift = 31 - __clz(row_items);
#pragma unroll 2
  for (int i0 = 0; i0 < TW * U2; i0 += IPB) {
    const bool valid = i0 + slot < TW * U2;  // invalid lanes stay alive for the shuffles
    const int i = valid ? i0 + slot : 0;
    const int dy = i >> row_shift, r = i & (row_items - 1), pl = r / UP, dx = r % UP, w = w0 + pl;
    const int64_t opix = ((int64_t)b * OH + h * UP + dy) * OW + (int64_t)w0 * UP + r;
    // the saved weights do not depend on the reduction below: fetch them first (one chain, not two)
    float k[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) k[t] = to_f32(__ldg(wt + opix * 9 + t));
    float dw[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) dw[t] = 0.f;
    for (int c0 = sl * CG; c0 < g.Co; c0 += LPP * CG) {
      float go[CG];
      ldc<T, CG>(gout + opix * g.Co + c0, go);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int hy = h + t / 3 - 1, wx = w + t % 3 - 1;
        if (hy < 0 || hy >= g.H || wx < 0 || wx >= g.W) continue;
        float f[CG];
        ldc<T, CG>(low + (((int64_t)b * g.H + hy) * g.W + wx) * g.Co + c0, f);
#pragma unroll
        for (int c = 0; c < CG; ++c) dw[t] = fmaf(go[c], f[c], dw[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) dw[t] += __shfl_xor_sync(0xffffffffu, dw[t], o);
    if (sl == 0 && valid) {
      float dot = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) dot = fmaf(k[t], dw[t], dot);
      T* de = s_d + pl * NL + dy * UP + dx;
#pragma unroll
      for (int t = 0; t < 9; ++t) de[t * U2] = from_f32<T>(k[t] * (dw[t] - dot));
    }
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(d_enc + (((int64_t)b * g.H + h) * g.W + w0) * NL);
  const uint4* srcv = reinterpret_cast<const uint4*>(s_d);
  for (int v = threadIdx.x; v < TW * NL / VE; v += 256) dst[v] = srcv[v];
}

// ------------------------------------------------------------------------------------------------
// Co == 1, up == 4: the last decoder step (up_x4 collapsed to one channel) — 151 MB of
// logits and 151 MB of saved weights at 512^2 / batch 32.  One thread per (source pixel, output row dy of
// the pixel), dy fastest: the four lanes of a pixel read the four 8-byte quarters of every 32-byte tap
// sector of its logits row, so the loads are sector-exact WITHOUT a shared-memory stage; the 4 x 9 saved
// weights of a thread are one contiguous 72-byte run of the weights tensor (nine 8-byte stores, the
// (tap, dx) -> (dx, tap) transpose is register renaming).  The logit-gradient kernel mirrors it.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct Quad;  // four consecutive elements of T
template <>
struct Quad<__nv_bfloat16> {
  using V = uint2;
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&f)[4]) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    f[0] = __uint_as_float(u.x << 16);
    f[1] = __uint_as_float(u.x & 0xffff0000u);
    f[2] = __uint_as_float(u.y << 16);
    f[3] = __uint_as_float(u.y & 0xffff0000u);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float a, float b, float c, float d) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
  }
};
template <>
struct Quad<float> {
  static __device__ __forceinline__ void ld(const float* p, float (&f)[4]) {
    const float4 u = __ldg(reinterpret_cast<const float4*>(p));
    f[0] = u.x; f[1] = u.y; f[2] = u.z; f[3] = u.w;
  }
  static __device__ __forceinline__ void st(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
  }
};

// value of element m (0..35) of a thread's weights run: m = dx * 9 + tap
#define CSB_RUN(arr, m) arr[(m) % 9][(m) / 9]

template <typename T>
__global__ void __launch_bounds__(256)
    carafe_fwd_c1u4(CGeom g, const T* __restrict__ low, const T* __restrict__ enc, T* __restrict__ out,
                    T* __restrict__ wt) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (int64_t)g.B * g.H * g.W * 4) return;
  const int dy = (int)(idx & 3);
  const int64_t pix = idx >> 2;  // (b, h, w) flattened
  const int w = (int)(pix % g.W), h = (int)((pix / g.W) % g.H);
  const int64_t bh0 = pix - w - (int64_t)h * g.W;  // b * H * W
  float k[9][4], lo[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    Quad<T>::ld(enc + pix * 144 + t * 16 + dy * 4, k[t]);
    const int hy = h + t / 3 - 1, wx = w + t % 3 - 1;
    lo[t] = (hy < 0 || hy >= g.H || wx < 0 || wx >= g.W) ? 0.f : to_f32(__ldg(low + bh0 + (int64_t)hy * g.W + wx));
  }
  float o[4];
#pragma unroll
  for (int dx = 0; dx < 4; ++dx) {
    float mx = k[0][dx];
#pragma unroll
    for (int t = 1; t < 9; ++t) mx = fmaxf(mx, k[t][dx]);
    float z = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      k[t][dx] = CExp<T>::f(k[t][dx] - mx);
      z += k[t][dx];
    }
    const float iz = 1.f / z;
    float a = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      k[t][dx] *= iz;
      a = fmaf(k[t][dx], lo[t], a);
    }
    o[dx] = a;
  }
  const int OW = g.W * 4;
  const int64_t opix = ((bh0 / g.W) * 4 + (int64_t)h * 4 + dy) * OW + (int64_t)w * 4;  // (b*OH + oy)*OW + ox
  Quad<T>::st(out + opix, o[0], o[1], o[2], o[3]);
  if (wt != nullptr) {
    T* run = wt + opix * 9;
#pragma unroll
    for (int j = 0; j < 9; ++j)
      Quad<T>::st(run + 4 * j, CSB_RUN(k, 4 * j), CSB_RUN(k, 4 * j + 1), CSB_RUN(k, 4 * j + 2), CSB_RUN(k, 4 * j + 3));
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    carafe_bwd_enc_c1u4(CGeom g, const T* __restrict__ low, const T* __restrict__ wt,
                        const T* __restrict__ gout, T* __restrict__ d_enc) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (int64_t)g.B * g.H * g.W * 4) return;
  const int dy = (int)(idx & 3);
  const int64_t pix = idx >> 2;
  const int w = (int)(pix % g.W), h = (int)((pix / g.W) % g.H);
  const int64_t bh0 = pix - w - (int64_t)h * g.W;
  const int OW = g.W * 4;
  const int64_t opix = ((bh0 / g.W) * 4 + (int64_t)h * 4 + dy) * OW + (int64_t)w * 4;
  float run[9][4];  // run[j][e] = element 4 j + e of the 36-element weights run, element m = dx * 9 + tap
#pragma unroll
  for (int j = 0; j < 9; ++j) Quad<T>::ld(wt + opix * 9 + 4 * j, run[j]);
  float go[4], lo[9];
  Quad<T>::ld(gout + opix, go);
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int hy = h + t / 3 - 1, wx = w + t % 3 - 1;
    lo[t] = (hy < 0 || hy >= g.H || wx < 0 || wx >= g.W) ? 0.f : to_f32(__ldg(low + bh0 + (int64_t)hy * g.W + wx));
  }
  // d_logit[t] = k[t] (d_wt[t] - sum_t' k[t'] d_wt[t']) with d_wt[t] = grad_out * low[t]
  float de[9][4];
#pragma unroll
  for (int dx = 0; dx < 4; ++dx) {
    float dot = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) dot = fmaf(run[(dx * 9 + t) / 4][(dx * 9 + t) % 4], lo[t], dot);
#pragma unroll
    for (int t = 0; t < 9; ++t) de[t][dx] = run[(dx * 9 + t) / 4][(dx * 9 + t) % 4] * go[dx] * (lo[t] - dot);
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
    Quad<T>::st(d_enc + pix * 144 + t * 16 + dy * 4, de[t][0], de[t][1], de[t][2], de[t][3]);
}

// tile width for the tiled kernels: 0 = shape not covered (the per-pixel kernels take it)
inline int carafe_tile(const CGeom& g) {
  if (g.up != 2 && g.up != 4) return 0;
  for (int tw = g.up == 2 ? 64 : 32; tw >= 16; tw >>= 1)  // <= 37 KB of shared memory in every configuration
    if (g.W % tw == 0) return tw;
  return 0;
}

int blocks_for(int64_t threads) { return (int)((threads + 255) / 256); }

template <typename T>
int carafe_fwd_t(const CGeom& g, const void* low, const void* enc, void* out, void* wt, cudaStream_t st) {
  const int64_t opix = (int64_t)g.B * g.H * g.up * g.W * g.up;
  if (g.Co == 1 && g.up == 4 &&
      ((reinterpret_cast<uintptr_t>(enc) | reinterpret_cast<uintptr_t>(wt) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    carafe_fwd_c1u4<T><<<blocks_for((int64_t)g.B * g.H * g.W * 4), 256, 0, st>>>(g, (const T*)low, (const T*)enc,
                                                                              (T*)out, (T*)wt);
    return check_launch("carafe_fwd_c1u4");
  }
  const int tw = carafe_tile(g);
  const bool al = ((reinterpret_cast<uintptr_t>(enc) | reinterpret_cast<uintptr_t>(wt) |
                    reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (tw != 0 && al && (int64_t)g.B * g.H * (g.W / tw) <= 0x7fffffff) {
    const int grid = g.B * g.H * (g.W / tw);
    const size_t smem = (size_t)tw * (9 * g.up * g.up) * sizeof(T) + (size_t)tw * (9 * g.up * g.up + 1) * sizeof(float);
#define CSB_CF(CGV, UPV)                                                                                  \
  carafe_fwd_tiled<T, CGV, UPV><<<grid, 256, smem, st>>>(g, tw, (const T*)low, (const T*)enc, (T*)out, (T*)wt)
    if (g.Co == 1) { if (g.up == 2) CSB_CF(1, 2); else CSB_CF(1, 4); }
    else { if (g.up == 2) CSB_CF(8, 2); else CSB_CF(8, 4); }
#undef CSB_CF
    return check_launch("carafe_fwd_tiled");
  }
  if (g.Co == 1)
    carafe_fwd_kernel<T, 1><<<blocks_for(opix), 256, 0, st>>>(g, (const T*)low, (const T*)enc, (T*)out, (T*)wt);
  else
    carafe_fwd_kernel<T, 8><<<blocks_for(opix * (g.Co / 8)), 256, 0, st>>>(g, (const T*)low, (const T*)enc,
                                                                          (T*)out, (T*)wt);
  return check_launch("carafe_fwd_kernel");
}

template <typename T>
int carafe_bwd_t(const CGeom& g, const void* low, const void* wt, const void* gout, void* d_low,
                 void* d_enc, cudaStream_t st) {
  const int64_t opix = (int64_t)g.B * g.H * g.up * g.W * g.up, lpix = (int64_t)g.B * g.H * g.W;
  const T *l = (const T*)low, *k = (const T*)wt, *go = (const T*)gout;
  const int tw = carafe_tile(g);
  if (g.Co == 1 && g.up == 4 &&
      ((reinterpret_cast<uintptr_t>(d_enc) | reinterpret_cast<uintptr_t>(wt) | reinterpret_cast<uintptr_t>(gout)) & 15) == 0) {
    carafe_bwd_enc_c1u4<T><<<blocks_for((int64_t)g.B * g.H * g.W * 4), 256, 0, st>>>(g, l, k, go, (T*)d_enc);
  } else if (tw != 0 && (reinterpret_cast<uintptr_t>(d_enc) & 15) == 0 && (int64_t)g.B * g.H * (g.W / tw) <= 0x7fffffff) {
    const int grid = g.B * g.H * (g.W / tw);
    const size_t smem = (size_t)tw * 9 * g.up * g.up * sizeof(T);
    const int ngrp = g.Co / 8;
#define CSB_CB(CGV, LPPV)                                                                                    \
  do {                                                                                                       \
    if (g.up == 2) carafe_bwd_enc_tiled<T, CGV, LPPV, 2><<<grid, 256, smem, st>>>(g, tw, l, k, go, (T*)d_enc); \
    else carafe_bwd_enc_tiled<T, CGV, LPPV, 4><<<grid, 256, smem, st>>>(g, tw, l, k, go, (T*)d_enc);         \
  } while (0)
    if (g.Co == 1) CSB_CB(1, 1);
    else if (ngrp >= 32) CSB_CB(8, 32);
    else if (ngrp >= 16) CSB_CB(8, 16);
    else if (ngrp >= 8) CSB_CB(8, 8);
    else CSB_CB(8, 1);
#undef CSB_CB
  } else if (g.Co == 1) {
    carafe_bwd_enc_kernel<T, 1, 1><<<blocks_for(opix), 256, 0, st>>>(g, l, k, go, (T*)d_enc);
  } else {
    const int ngrp = g.Co / 8;
    if (ngrp >= 32) carafe_bwd_enc_kernel<T, 8, 32><<<blocks_for(opix * 32), 256, 0, st>>>(g, l, k, go, (T*)d_enc);
    else if (ngrp >= 16) carafe_bwd_enc_kernel<T, 8, 16><<<blocks_for(opix * 16), 256, 0, st>>>(g, l, k, go, (T*)d_enc);
    else if (ngrp >= 8) carafe_bwd_enc_kernel<T, 8, 8><<<blocks_for(opix * 8), 256, 0, st>>>(g, l, k, go, (T*)d_enc);
    else carafe_bwd_enc_kernel<T, 8, 1><<<blocks_for(opix), 256, 0, st>>>(g, l, k, go, (T*)d_enc);
  }
  int rc = check_launch("carafe_bwd_enc_kernel");
  if (rc != CSB200_OK) return rc;
  if (g.Co == 1) carafe_bwd_low_kernel<T, 1><<<blocks_for(lpix), 256, 0, st>>>(g, k, go, (T*)d_low);
  else carafe_bwd_low_kernel<T, 8><<<blocks_for(lpix * (g.Co / 8)), 256, 0, st>>>(g, k, go, (T*)d_low);
  return check_launch("carafe_bwd_low_kernel");
}

int carafe_check(int64_t B, int64_t H, int64_t W, int64_t Co, int up, int dtype, CGeom* g) {
  if (dtype != CSB200_F32 && dtype != CSB200_BF16) return fail(CSB200_ERR_INVALID, "carafe: unknown dtype %d", dtype);
  if (B < 0 || H <= 0 || W <= 0 || Co <= 0 || up <= 0) return fail(CSB200_ERR_INVALID, "carafe: bad size");
  if (Co != 1 && Co % 8 != 0) return fail(CSB200_ERR_UNSUPPORTED, "carafe: channels must be 1 or a multiple of 8");
  if (B * H * up * W * up > (int64_t)1 << 33) return fail(CSB200_ERR_INVALID, "carafe: tensor too large");
  g->B = (int)B; g->H = (int)H; g->W = (int)W; g->Co = (int)Co; g->up = up;
  return CSB200_OK;
}

}  // namespace
}  // namespace csb200

using namespace csb200;

extern "C" int csb200_carafe_supported(int64_t channels) { return channels == 1 || channels % 8 == 0; }

extern "C" int csb200_carafe_fwd(const void* low, const void* enc, void* out, void* weights,
                                 int64_t batch, int64_t height, int64_t width, int64_t channels,
                                 int up, int dtype, void* stream) {
  CGeom g;
  int rc = carafe_check(batch, height, width, channels, up, dtype, &g);
  if (rc != CSB200_OK) return rc;
  if (batch == 0) return CSB200_OK;
  if (!low || !enc || !out) return fail(CSB200_ERR_INVALID, "carafe_fwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return dtype == CSB200_F32 ? carafe_fwd_t<float>(g, low, enc, out, weights, st)
                             : carafe_fwd_t<__nv_bfloat16>(g, low, enc, out, weights, st);
}

extern "C" int csb200_carafe_bwd(const void* low, const void* weights, const void* grad_out,
                                 void* grad_low, void* grad_enc, int64_t batch, int64_t height,
                                 int64_t width, int64_t channels, int up, int dtype, void* stream) {
  CGeom g;
  int rc = carafe_check(batch, height, width, channels, up, dtype, &g);
  if (rc != CSB200_OK) return rc;
  if (batch == 0) return CSB200_OK;
  if (!low || !weights || !grad_out || !grad_low || !grad_enc)
    return fail(CSB200_ERR_INVALID, "carafe_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return dtype == CSB200_F32 ? carafe_bwd_t<float>(g, low, weights, grad_out, grad_low, grad_enc, st)
                             : carafe_bwd_t<__nv_bfloat16>(g, low, weights, grad_out, grad_low, grad_enc, st);
}

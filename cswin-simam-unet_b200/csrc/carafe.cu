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

int blocks_for(int64_t threads) { return (int)((threads + 255) / 256); }

template <typename T>
int carafe_fwd_t(const CGeom& g, const void* low, const void* enc, void* out, void* wt, cudaStream_t st) {
  const int64_t opix = (int64_t)g.B * g.H * g.up * g.W * g.up;
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
  if (g.Co == 1) {
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

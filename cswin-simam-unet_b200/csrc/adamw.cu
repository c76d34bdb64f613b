// The optimizer step of the reference train loop (C:786 `optimizer.step()` with AdamW lr 1e-4 wd 1e-4,
// C:937-941; U:486-490 Adam with L2 decay) for ALL parameter tensors of the model in ONE launch, and the
// bf16 shadows that the Linear / conv layers read under autocast are written in the same pass.
//
// The CSWin-UNet has 463 parameter tensors, most of them a few hundred elements: ATen's fused AdamW
// walks them through 13 multi_tensor_apply launches (kernel-argument-sized tables) at ~1/5 of the HBM
// roofline (0.49 ms for 23.6 M parameters = 660 MB of traffic), followed by one more multi-tensor pass to
// refresh the shadows.  Here the table of tensors lives in device memory, a second table maps every
// CTA to (tensor, chunk of 8192 elements), and one pass reads p, g, m, v and writes p, m, v (+ shadow):
// 28 (+2) bytes per parameter, HBM-bound.
//
// Arithmetic (torch.optim.AdamW / Adam, amsgrad = false, maximize = false), t = step count incl. this step:
//   decoupled: p <- p (1 - lr wd)          |  L2: g <- g + wd p
//   m <- m + (1 - b1)(g - m);  v <- b2 v + (1 - b2) g^2
//   p <- p - (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// Hyper-parameters and the step count are read from device memory, so a captured CUDA graph follows
// a learning-rate schedule (C:943-949 ReduceLROnPlateau) without being re-captured.

#include "common.cuh"

namespace csb200 {
namespace {

constexpr int AD_CHUNK = 8192, AD_THREADS = 256;

struct AdamConst {
  float lr, b1, b2, eps, wd, bc1, rsqrt_bc2;
  int decoupled;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamConst& c) {
  if (c.wd != 0.f) {
    if (c.decoupled) p -= c.lr * c.wd * p;
    else g = fmaf(c.wd, p, g);
  }
  m = fmaf(1.f - c.b1, g - m, m);
  v = fmaf(c.b2, v, (1.f - c.b2) * g * g);
  const float denom = sqrtf(v) * c.rsqrt_bc2 + c.eps;
  p -= (c.lr / c.bc1) * (m / denom);
}

__global__ void __launch_bounds__(AD_THREADS)
    adam_multi_kernel(const csb200_adam_tensor* __restrict__ tensors, const int2* __restrict__ chunks,
                      const float* __restrict__ hyper /* [groups][8] */, const float* __restrict__ step) {
  const int2 ck = chunks[blockIdx.x];
  const csb200_adam_tensor t = tensors[ck.x];
  const float* h = hyper + 8 * t.group;
  const float steps = __ldg(step);
  AdamConst c;
  c.lr = h[0]; c.b1 = h[1]; c.b2 = h[2]; c.eps = h[3]; c.wd = h[4];
  c.decoupled = h[5] != 0.f;
  c.bc1 = 1.f - powf(c.b1, steps);
  c.rsqrt_bc2 = rsqrtf(1.f - powf(c.b2, steps));
  float* p = static_cast<float*>(t.param);
  const float* g = static_cast<const float*>(t.grad);
  float* m = static_cast<float*>(t.exp_avg);
  float* v = static_cast<float*>(t.exp_avg_sq);
  __nv_bfloat16* sh = static_cast<__nv_bfloat16*>(t.shadow);
  const int64_t lo = (int64_t)ck.y * AD_CHUNK;
  const int64_t hi = lo + AD_CHUNK < t.numel ? lo + AD_CHUNK : t.numel;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
                   (sh == nullptr || (reinterpret_cast<uintptr_t>(sh) & 7) == 0);
  int64_t i = lo;
  if (vec) {
    const int64_t nv = (hi - lo) / 4;
    for (int64_t q = threadIdx.x; q < nv; q += AD_THREADS) {
      const int64_t e = lo + 4 * q;
      float4 pp = *reinterpret_cast<const float4*>(p + e);
      const float4 gg = *reinterpret_cast<const float4*>(g + e);
      float4 mm = *reinterpret_cast<const float4*>(m + e);
      float4 vv = *reinterpret_cast<const float4*>(v + e);
      adam_elem(pp.x, gg.x, mm.x, vv.x, c);
      adam_elem(pp.y, gg.y, mm.y, vv.y, c);
      adam_elem(pp.z, gg.z, mm.z, vv.z, c);
      adam_elem(pp.w, gg.w, mm.w, vv.w, c);
      *reinterpret_cast<float4*>(p + e) = pp;
      *reinterpret_cast<float4*>(m + e) = mm;
      *reinterpret_cast<float4*>(v + e) = vv;
      if (sh != nullptr)
        *reinterpret_cast<uint2*>(sh + e) = make_uint2(pack_bf16x2(pp.x, pp.y), pack_bf16x2(pp.z, pp.w));
    }
    i = lo + 4 * nv;
  }
  for (int64_t e = i + threadIdx.x; e < hi; e += AD_THREADS) {
    float pp = p[e], mm = m[e], vv = v[e];
    adam_elem(pp, g[e], mm, vv, c);
    p[e] = pp;
    m[e] = mm;
    v[e] = vv;
    if (sh != nullptr) sh[e] = __float2bfloat16_rn(pp);
  }
}

}  // namespace
}  // namespace csb200

using namespace csb200;

extern "C" int64_t csb200_adam_chunk_elems(void) { return AD_CHUNK; }

extern "C" int csb200_adam_step(const csb200_adam_tensor* tensors_dev, const int32_t* chunks_dev,
                                int64_t n_chunks, const float* hyper_dev, const float* step_dev, void* stream) {
  if (n_chunks < 0 || n_chunks > 0x7fffffff) return fail(CSB200_ERR_INVALID, "adam_step: bad chunk count");
  if (n_chunks == 0) return CSB200_OK;
  if (!tensors_dev || !chunks_dev || !hyper_dev || !step_dev) return fail(CSB200_ERR_INVALID, "adam_step: null pointer");
  adam_multi_kernel<<<(int)n_chunks, AD_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      tensors_dev, reinterpret_cast<const int2*>(chunks_dev), hyper_dev, step_dev);
  return check_launch("adam_multi_kernel");
}

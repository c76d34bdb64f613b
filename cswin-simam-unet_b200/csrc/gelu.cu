// GELU of the Mlp hidden layer (C:188-196: fc1 -> act_layer() = nn.GELU, exact erf form) as two
// HBM-bound flat passes:
//   forward   a = h * 0.5 * (1 + erf(h / sqrt 2))                         (1 read + 1 write)
//   backward  dh = g * (cdf(h) + h * pdf(h))  AND  grad_bias(fc1) = column sums of dh, in the SAME
//             pass (2 reads + 1 write; the separate colsum pass over the 4C-wide dh disappears).
// Every thread keeps one fixed 16-byte column vector (block size = a multiple of cols / vector
// width, as in colsum.cu), so loads and stores stay perfectly contiguous while the 4 / 8 bias
// partials live in registers; per-CTA partials are added in a fixed order (deterministic).
// fp32 uses erff / expf (parity target 1e-5); bf16 shares ONE exponential between the Gaussian term
// and an Abramowitz-Stegun 7.1.26 erf (|error| < 1.5e-7, far below bf16 rounding) so the kernel stays
// under the instruction budget of the HBM roofline (~33 issue slots per element).

#include "common.cuh"
#include "gelu_math.cuh"

namespace csb200 {
namespace {

constexpr int GELU_MAX_GRID = 148 * 4;

template <typename T>
struct Gelu;
template <>
struct Gelu<float> {
  static __device__ __forceinline__ float fwd(float x) { return 0.5f * x * (1.f + erff(x * kSqrtHalf)); }
  static __device__ __forceinline__ float bwd(float g, float x) {
    const float cdf = 0.5f * (1.f + erff(x * kSqrtHalf));
    const float pdf = expf(-0.5f * x * x) * kInvSqrt2Pi;
    return g * fmaf(x, pdf, cdf);
  }
};
// element-wise forward / backward of one 16-byte vector
template <typename T>
__device__ __forceinline__ uint4 gelu_fwd_vec(const uint4& u) {
  if constexpr (sizeof(T) == 2) {
    return make_uint4(gelu_fwd2(u.x), gelu_fwd2(u.y), gelu_fwd2(u.z), gelu_fwd2(u.w));
  } else {
    float t[Vec16<T>::N];
    unpack<T>(u, t);
#pragma unroll
    for (int e = 0; e < Vec16<T>::N; ++e) t[e] = Gelu<T>::fwd(t[e]);
    return pack<T>(t);
  }
}
template <typename T>
__device__ __forceinline__ uint4 gelu_bwd_vec(const uint4& ug, const uint4& uh) {
  if constexpr (sizeof(T) == 2) {
    return make_uint4(gelu_bwd2(ug.x, uh.x), gelu_bwd2(ug.y, uh.y), gelu_bwd2(ug.z, uh.z), gelu_bwd2(ug.w, uh.w));
  } else {
    float tg[Vec16<T>::N], th[Vec16<T>::N];
    unpack<T>(ug, tg);
    unpack<T>(uh, th);
#pragma unroll
    for (int e = 0; e < Vec16<T>::N; ++e) tg[e] = Gelu<T>::bwd(tg[e], th[e]);
    return pack<T>(tg);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const T* __restrict__ h, T* __restrict__ out, int64_t nvec) {
  const uint4* hv = reinterpret_cast<const uint4*>(h);
  uint4* ov = reinterpret_cast<uint4*>(out);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; f + 3 * stride < nvec; f += 4 * stride) {
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = ld_stream(hv + f + i * stride);
#pragma unroll
    for (int i = 0; i < 4; ++i) ov[f + i * stride] = gelu_fwd_vec<T>(u[i]);  // re-read by the fc2 GEMM: leave it in L2
  }
  for (; f < nvec; f += stride) ov[f] = gelu_fwd_vec<T>(ld_stream(hv + f));
}

template <typename T>
__global__ void __launch_bounds__(256)
    gelu_bwd_colsum_kernel(const T* __restrict__ g, const T* __restrict__ h, T* __restrict__ dh,
                           float* __restrict__ partial, int64_t nvec, int cvn) {
  constexpr int VE = Vec16<T>::N;
  __shared__ float s_acc[256 * VE];
  const int tpb = blockDim.x;  // multiple of cvn: a thread always sees the same column vector
  float acc[VE];
#pragma unroll
  for (int e = 0; e < VE; ++e) acc[e] = 0.f;
  const uint4* gv = reinterpret_cast<const uint4*>(g);
  const uint4* hv = reinterpret_cast<const uint4*>(h);
  uint4* dv = reinterpret_cast<uint4*>(dh);
  const int64_t stride = (int64_t)gridDim.x * tpb;
  int64_t f = (int64_t)blockIdx.x * tpb + threadIdx.x;
  for (; f + stride < nvec; f += 2 * stride) {
    uint4 ug[2], uh[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      ug[i] = ld_stream(gv + f + i * stride);
      uh[i] = ld_stream(hv + f + i * stride);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint4 o = gelu_bwd_vec<T>(ug[i], uh[i]);
      dv[f + i * stride] = o;
      // the bias gradient sums what the GEMMs will read: the ROUNDED dh (as ATen's sum over it would)
      float tg[VE];
      unpack<T>(o, tg);
#pragma unroll
      for (int e = 0; e < VE; ++e) acc[e] += tg[e];
    }
  }
  for (; f < nvec; f += stride) {
    const uint4 o = gelu_bwd_vec<T>(ld_stream(gv + f), ld_stream(hv + f));
    dv[f] = o;
    float tg[VE];
    unpack<T>(o, tg);
#pragma unroll
    for (int e = 0; e < VE; ++e) acc[e] += tg[e];
  }
#pragma unroll
  for (int e = 0; e < VE; ++e) s_acc[threadIdx.x * VE + e] = acc[e];
  __syncthreads();
  for (int i = threadIdx.x; i < cvn * VE; i += tpb) {
    const int cv = i / VE, e = i % VE;
    float a = 0.f;
    for (int t = cv; t < tpb; t += cvn) a += s_acc[t * VE + e];
    partial[(int64_t)blockIdx.x * cvn * VE + i] = a;
  }
}

__global__ void __launch_bounds__(256)
    gelu_colsum_final(const float* __restrict__ partial, int blocks, int cols, float* __restrict__ out) {
  const int i = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= cols) return;
  const float a = strided_partial_sum(partial + i, blocks, cols, lane);
  if (lane == 0) out[i] = a;
}

bool vec_ok(int64_t cols, int dtype) {
  const int ve = dtype == CSB200_F32 ? 4 : (dtype == CSB200_BF16 ? 8 : 0);
  return ve != 0 && cols > 0 && cols % ve == 0 && cols / ve <= 256;
}

template <typename T>
int fwd_t(const void* h, void* out, int64_t numel, cudaStream_t st) {
  const int64_t nvec = numel / Vec16<T>::N;
  int64_t grid = (nvec + 256 * 8 - 1) / (256 * 8);
  grid = grid < 1 ? 1 : (grid > GELU_MAX_GRID * 4 ? GELU_MAX_GRID * 4 : grid);
  gelu_fwd_kernel<T><<<(int)grid, 256, 0, st>>>(static_cast<const T*>(h), static_cast<T*>(out), nvec);
  return check_launch("gelu_fwd");
}

template <typename T>
int bwd_t(const void* g, const void* h, void* dh, float* gb, float* partial, int64_t rows, int64_t cols,
          cudaStream_t st) {
  constexpr int VE = Vec16<T>::N;
  const int cvn = (int)(cols / VE);
  const int tpb = 256 / cvn * cvn;
  const int64_t nvec = rows * cvn;
  int64_t grid = (nvec + (int64_t)tpb * 8 - 1) / ((int64_t)tpb * 8);
  grid = grid < 1 ? 1 : (grid > GELU_MAX_GRID ? GELU_MAX_GRID : grid);
  gelu_bwd_colsum_kernel<T><<<(int)grid, tpb, 0, st>>>(static_cast<const T*>(g), static_cast<const T*>(h),
                                                       static_cast<T*>(dh), partial, nvec, cvn);
  int rc = check_launch("gelu_bwd_colsum");
  if (rc != CSB200_OK) return rc;
  gelu_colsum_final<<<(int)((cols * 32 + 255) / 256), 256, 0, st>>>(partial, (int)grid, (int)cols, gb);
  return check_launch("gelu_colsum_final");
}

}  // namespace
}  // namespace csb200

using namespace csb200;

extern "C" int csb200_gelu_supported(int64_t cols, int dtype) { return vec_ok(cols, dtype) ? 1 : 0; }

extern "C" size_t csb200_gelu_bwd_workspace_bytes(int64_t cols) {
  return (size_t)GELU_MAX_GRID * (size_t)cols * sizeof(float) + 256;
}

extern "C" int csb200_gelu_fwd(const void* h, void* out, int64_t rows, int64_t cols, int dtype, void* stream) {
  if (rows < 0 || !vec_ok(cols, dtype))
    return fail(CSB200_ERR_UNSUPPORTED, "gelu: cols=%lld dtype=%d is not tiled", (long long)cols, dtype);
  if (rows == 0) return CSB200_OK;
  if (!h || !out) return fail(CSB200_ERR_INVALID, "gelu_fwd: null pointer");
  if (((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(out)) & 15) != 0)
    return fail(CSB200_ERR_INVALID, "gelu_fwd: tensors must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return dtype == CSB200_F32 ? fwd_t<float>(h, out, rows * cols, st)
                             : fwd_t<__nv_bfloat16>(h, out, rows * cols, st);
}

extern "C" int csb200_gelu_bwd(const void* grad_out, const void* h, void* grad_h, float* grad_bias,
                               void* workspace, size_t workspace_bytes, int64_t rows, int64_t cols, int dtype,
                               void* stream) {
  if (rows < 0 || !vec_ok(cols, dtype))
    return fail(CSB200_ERR_UNSUPPORTED, "gelu: cols=%lld dtype=%d is not tiled", (long long)cols, dtype);
  if (!grad_bias || !workspace) return fail(CSB200_ERR_INVALID, "gelu_bwd: null pointer");
  if (workspace_bytes < csb200_gelu_bwd_workspace_bytes(cols))
    return fail(CSB200_ERR_WORKSPACE, "gelu_bwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rows == 0) {
    CSB200_CUDA(cudaMemsetAsync(grad_bias, 0, cols * sizeof(float), st));
    return CSB200_OK;
  }
  if (!grad_out || !h || !grad_h) return fail(CSB200_ERR_INVALID, "gelu_bwd: null pointer");
  if (((reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(h) |
        reinterpret_cast<uintptr_t>(grad_h)) & 15) != 0)
    return fail(CSB200_ERR_INVALID, "gelu_bwd: tensors must be 16-byte aligned");
  float* partial = static_cast<float*>(workspace);
  return dtype == CSB200_F32
             ? bwd_t<float>(grad_out, h, grad_h, grad_bias, partial, rows, cols, st)
             : bwd_t<__nv_bfloat16>(grad_out, h, grad_h, grad_bias, partial, rows, cols, st);
}

// Column sums of a row-major [rows][cols] matrix — the bias gradient of every Linear on the token
// path (qkv C:358, proj C:366, Mlp C:188-196, concat_linear C:658): grad_b = sum over tokens of
// grad_out.  ATen's generic reduce_kernel ran these 118 reductions per train step at ~1/9 of the HBM
// roofline (13.7 % of the step, profiles/r1_step_profile.md).  Here the matrix is read once as a FLAT
// stream of 16-byte vectors: the block size is the largest multiple of (cols / vector width) <= 256,
// so every thread keeps one fixed column vector while all loads stay perfectly contiguous.
// Two stages (per-CTA partials, warp-per-column final sum) keep the result deterministic.

#include "common.cuh"

namespace csb200 {
namespace {

constexpr int CS_MAX_GRID = 148 * 4;

template <typename T>
__global__ void __launch_bounds__(256)
    colsum_partial(const T* __restrict__ x, float* __restrict__ partial, int64_t nvec_total, int cvn) {
  constexpr int VE = Vec16<T>::N;
  __shared__ float s_acc[256 * VE];
  const int tpb = blockDim.x;  // multiple of cvn
  float acc[VE];
#pragma unroll
  for (int e = 0; e < VE; ++e) acc[e] = 0.f;
  const uint4* xv = reinterpret_cast<const uint4*>(x);
  const int64_t stride = (int64_t)gridDim.x * tpb;
  int64_t f = (int64_t)blockIdx.x * tpb + threadIdx.x;
  // 4 independent loads in flight per thread
  for (; f + 3 * stride < nvec_total; f += 4 * stride) {
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = ld_stream(xv + f + i * stride);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float t[VE];
      unpack<T>(u[i], t);
#pragma unroll
      for (int e = 0; e < VE; ++e) acc[e] += t[e];
    }
  }
  for (; f < nvec_total; f += stride) {
    float t[VE];
    unpack<T>(ld_stream(xv + f), t);
#pragma unroll
    for (int e = 0; e < VE; ++e) acc[e] += t[e];
  }
#pragma unroll
  for (int e = 0; e < VE; ++e) s_acc[threadIdx.x * VE + e] = acc[e];
  __syncthreads();
  // threads t, t + cvn, t + 2 cvn ... hold the same column vector
  for (int i = threadIdx.x; i < cvn * VE; i += tpb) {
    const int cv = i / VE, e = i % VE;
    float a = 0.f;
    for (int t = cv; t < tpb; t += cvn) a += s_acc[t * VE + e];
    partial[(int64_t)blockIdx.x * cvn * VE + i] = a;
  }
}

__global__ void __launch_bounds__(256)
    colsum_final(const float* __restrict__ partial, int blocks, int cols, float* __restrict__ out) {
  const int i = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= cols) return;
  const float a = strided_partial_sum(partial + i, blocks, cols, lane);
  if (lane == 0) out[i] = a;
}

// y[r][c] = x[r][c] + bias[c] over a flat stream of 16-byte vectors: the same "one fixed column vector
// per thread" layout as colsum_partial, so the bias vector is loaded once and stays in registers.
// Replaces the broadcasting add_ that follows every cuDNN convolution on a channels-last tensor
// (ATen runs it through the non-vectorised elementwise_kernel: 0.17 ms for the 144-channel CARAFE
// encoder output at 512^2, against 0.05 ms of HBM time).
template <typename T>
__global__ void __launch_bounds__(256)
    add_row_bias_kernel(const T* __restrict__ x, const float* __restrict__ bias, T* __restrict__ y,
                        int64_t nvec_total, int cvn) {
  constexpr int VE = Vec16<T>::N;
  const int tpb = blockDim.x;  // multiple of cvn
  float b[VE];
  const int cv = threadIdx.x % cvn;
#pragma unroll
  for (int e = 0; e < VE; ++e) b[e] = __ldg(bias + cv * VE + e);
  const uint4* xv = reinterpret_cast<const uint4*>(x);
  uint4* yv = reinterpret_cast<uint4*>(y);
  const int64_t stride = (int64_t)gridDim.x * tpb;
  int64_t f = (int64_t)blockIdx.x * tpb + threadIdx.x;
  for (; f + 3 * stride < nvec_total; f += 4 * stride) {
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = ld_stream(xv + f + i * stride);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float t[VE];
      unpack<T>(u[i], t);
#pragma unroll
      for (int e = 0; e < VE; ++e) t[e] += b[e];
      yv[f + i * stride] = pack<T>(t);
    }
  }
  for (; f < nvec_total; f += stride) {
    float t[VE];
    unpack<T>(ld_stream(xv + f), t);
#pragma unroll
    for (int e = 0; e < VE; ++e) t[e] += b[e];
    yv[f] = pack<T>(t);
  }
}

template <typename T>
int add_row_bias_t(const void* x, const float* bias, void* y, int64_t rows, int64_t cols, cudaStream_t st) {
  constexpr int VE = Vec16<T>::N;
  const int cvn = (int)(cols / VE);
  const int tpb = 256 / cvn * cvn;
  const int64_t nvec = rows * cvn;
  int64_t grid = (nvec + (int64_t)tpb * 8 - 1) / ((int64_t)tpb * 8);
  grid = grid < 1 ? 1 : (grid > CS_MAX_GRID * 4 ? CS_MAX_GRID * 4 : grid);
  add_row_bias_kernel<T><<<(int)grid, tpb, 0, st>>>(static_cast<const T*>(x), bias, static_cast<T*>(y), nvec, cvn);
  return check_launch("add_row_bias");
}

template <typename T>
int colsum_t(const void* x, float* out, float* partial, int64_t rows, int64_t cols, cudaStream_t st,
             int32_t* partial_rows = nullptr) {
  constexpr int VE = Vec16<T>::N;
  const int cvn = (int)(cols / VE);
  const int tpb = 256 / cvn * cvn;
  const int64_t nvec = rows * cvn;
  int64_t grid = (nvec + (int64_t)tpb * 8 - 1) / ((int64_t)tpb * 8);  // >= 8 vectors per thread
  grid = grid < 1 ? 1 : (grid > CS_MAX_GRID ? CS_MAX_GRID : grid);
  colsum_partial<T><<<(int)grid, tpb, 0, st>>>(static_cast<const T*>(x), partial, nvec, cvn);
  int rc = check_launch("colsum_partial");
  if (rc != CSB200_OK) return rc;
  if (partial_rows != nullptr) {  // deferred final sum (csb200_sum_rows_deferred / _flush, sum_rows.cu)
    *partial_rows = (int32_t)grid;
    return CSB200_OK;
  }
  colsum_final<<<(int)((cols * 32 + 255) / 256), 256, 0, st>>>(partial, (int)grid, (int)cols, out);
  return check_launch("colsum_final");
}

}  // namespace
}  // namespace csb200

using namespace csb200;

extern "C" int csb200_colsum_supported(int64_t cols, int dtype) {
  const int ve = dtype == CSB200_F32 ? 4 : (dtype == CSB200_BF16 ? 8 : 0);
  return ve != 0 && cols > 0 && cols % ve == 0 && cols / ve <= 256;
}

extern "C" size_t csb200_colsum_workspace_bytes(int64_t cols) {
  return (size_t)CS_MAX_GRID * (size_t)cols * sizeof(float) + 256;
}

extern "C" int csb200_colsum(const void* x, float* out, void* workspace, size_t workspace_bytes,
                             int64_t rows, int64_t cols, int dtype, void* stream) {
  if (rows < 0 || !csb200_colsum_supported(cols, dtype))
    return fail(CSB200_ERR_UNSUPPORTED, "colsum: cols=%lld dtype=%d is not tiled", (long long)cols, dtype);
  if (!x || !out || !workspace) return fail(CSB200_ERR_INVALID, "colsum: null pointer");
  if (workspace_bytes < csb200_colsum_workspace_bytes(cols))
    return fail(CSB200_ERR_WORKSPACE, "colsum: workspace too small");
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return fail(CSB200_ERR_INVALID, "colsum: x must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rows == 0) {
    CSB200_CUDA(cudaMemsetAsync(out, 0, cols * sizeof(float), st));
    return CSB200_OK;
  }
  float* partial = static_cast<float*>(workspace);
  return dtype == CSB200_F32 ? colsum_t<float>(x, out, partial, rows, cols, st)
                             : colsum_t<__nv_bfloat16>(x, out, partial, rows, cols, st);
}

// csb200_colsum without its last launch: float[*partial_rows][cols] stays in the workspace.
extern "C" int csb200_colsum_partials(const void* x, void* workspace, size_t workspace_bytes, int64_t rows,
                                      int64_t cols, int dtype, const float** partials, int32_t* partial_rows,
                                      void* stream) {
  if (rows < 1 || !csb200_colsum_supported(cols, dtype))
    return fail(CSB200_ERR_UNSUPPORTED, "colsum_partials: rows=%lld cols=%lld dtype=%d", (long long)rows,
                (long long)cols, dtype);
  if (!x || !workspace || !partials || !partial_rows) return fail(CSB200_ERR_INVALID, "colsum_partials: null pointer");
  if (workspace_bytes < csb200_colsum_workspace_bytes(cols))
    return fail(CSB200_ERR_WORKSPACE, "colsum_partials: workspace too small");
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0)
    return fail(CSB200_ERR_INVALID, "colsum_partials: x must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(workspace);
  *partials = partial;
  return dtype == CSB200_F32 ? colsum_t<float>(x, nullptr, partial, rows, cols, st, partial_rows)
                             : colsum_t<__nv_bfloat16>(x, nullptr, partial, rows, cols, st, partial_rows);
}

extern "C" int csb200_add_row_bias(const void* x, const float* bias, void* y, int64_t rows, int64_t cols,
                                   int dtype, void* stream) {
  if (rows < 0 || !csb200_colsum_supported(cols, dtype))
    return fail(CSB200_ERR_UNSUPPORTED, "add_row_bias: cols=%lld dtype=%d is not tiled", (long long)cols, dtype);
  if (rows == 0) return CSB200_OK;
  if (!x || !bias || !y) return fail(CSB200_ERR_INVALID, "add_row_bias: null pointer");
  if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) != 0)
    return fail(CSB200_ERR_INVALID, "add_row_bias: tensors must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return dtype == CSB200_F32 ? add_row_bias_t<float>(x, bias, y, rows, cols, st)
                             : add_row_bias_t<__nv_bfloat16>(x, bias, y, rows, cols, st);
}

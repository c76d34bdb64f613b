// Token-major LayerNorm forward / backward for sm_100a — the pre-norms of CSWinBlock (C:357, C:368),
// Merge_Block (C:386), the stem (C:507) and the decoder norms (C:648, C:671).
//
// Why it is here: with C = 64..512 channels and up to 524 288 token rows per call (512^2, batch 32)
// ATen's LayerNorm backward (GammaBetaBackward + grad_input) was the largest single item of the train
// step (28 % of kernel time, profiles/r1_step_launches.md).  Both passes are HBM-bound streams:
//   forward : read x, write y (+ 8 B/row of statistics)           bytes = rows*C*(sizeof x + sizeof y)
//   backward: read x and grad_y, write grad_x                      bytes = rows*C*(sx + sgy + sgx)
// Design: a persistent grid; each warp walks row groups with 16-byte loads — LPR = min(32, C*sizeof/16)
// lanes per row, 32/LPR rows per warp step — statistics by xor-shuffles inside the row's lanes, exact
// two-pass variance in registers.  The output type is independent of the input type, so under autocast
// the fp32 residual stream is normalised straight into the bf16 operand of the next GEMM (no cast pass).
// Backward keeps gamma/beta gradient accumulators in registers over all rows of the warp and reduces
// them once (shuffles -> shared memory -> per-CTA partials -> fixed-order final sum: deterministic).

#include "common.cuh"

namespace csb200 {
namespace {

constexpr int LN_THREADS = 256;
constexpr int LN_WARPS = LN_THREADS / 32;

// convert NE fp32 values to TOut and store them contiguously (NE = 4 or 8)
template <typename TOut, int NE>
__device__ __forceinline__ void store_n(TOut* p, const float (&f)[NE]) {
  if constexpr (sizeof(TOut) == 4) {
#pragma unroll
    for (int i = 0; i < NE; i += 4)
      st_stream(p + i, make_uint4(__float_as_uint(f[i]), __float_as_uint(f[i + 1]),
                                  __float_as_uint(f[i + 2]), __float_as_uint(f[i + 3])));
  } else {
    if constexpr (NE == 8) {
      st_stream(p, make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                              pack_bf16x2(f[6], f[7])));
    } else {
      *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]));
    }
  }
}
template <typename TIn, int NE>
__device__ __forceinline__ void load_n(const TIn* p, float (&f)[NE]) {
  if constexpr (sizeof(TIn) == 4) {
#pragma unroll
    for (int i = 0; i < NE; i += 4) {
      const uint4 u = ld_stream(p + i);
      f[i] = __uint_as_float(u.x);
      f[i + 1] = __uint_as_float(u.y);
      f[i + 2] = __uint_as_float(u.z);
      f[i + 3] = __uint_as_float(u.w);
    }
  } else {
    if constexpr (NE == 8) {
      float t[8];
      unpack<__nv_bfloat16>(ld_stream(p), t);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = t[i];
    } else {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
      f[0] = __uint_as_float(u.x << 16);
      f[1] = __uint_as_float(u.x & 0xffff0000u);
      f[2] = __uint_as_float(u.y << 16);
      f[3] = __uint_as_float(u.y & 0xffff0000u);
    }
  }
}

// NE elements of T kept PACKED (16 or 8 bytes): what the software pipeline holds for the next row group —
// a quarter of the registers of the unpacked fp32 form
template <typename T, int NE>
struct RawN {
  static constexpr int WORDS = NE * (int)sizeof(T) / 4;
  uint32_t w[WORDS];
  __device__ __forceinline__ void load(const T* p) {
    if constexpr (WORDS == 4) {
      const uint4 u = ld_stream(p);
      w[0] = u.x; w[1] = u.y; w[2] = u.z; w[3] = u.w;
    } else {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
      w[0] = u.x; w[1] = u.y;
    }
  }
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < WORDS; ++i) w[i] = 0u;
  }
  __device__ __forceinline__ void unpack_to(float (&f)[NE]) const {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
      for (int i = 0; i < NE; ++i) f[i] = __uint_as_float(w[i]);
    } else {
#pragma unroll
      for (int i = 0; i < NE / 2; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
      }
    }
  }
};

template <int LPR>
__device__ __forceinline__ float row_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// A lane holds CPL chunks of NE elements; chunk j of
// a lane covers channels (j * LPR + lane_in_row) * NE .. + NE  (coalesced across the row's lanes).
template <typename TIn, typename TOut, int LPR, int CPL, int NE>
__global__ void __launch_bounds__(LN_THREADS)
    layernorm_fwd_kernel(const TIn* __restrict__ x, const TIn* __restrict__ res, TIn* __restrict__ sum_out,
                         const float* __restrict__ gamma, const float* __restrict__ beta,
                         TOut* __restrict__ y, float* __restrict__ stats, int64_t rows, float eps) {
  constexpr int C = LPR * CPL * NE, RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, lr = lane % LPR, sub = lane / LPR;
  float g[CPL][NE], b[CPL][NE];
#pragma unroll
  for (int j = 0; j < CPL; ++j)
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      g[j][e] = __ldg(gamma + (j * LPR + lr) * NE + e);
      b[j][e] = __ldg(beta + (j * LPR + lr) * NE + e);
    }
  const int64_t warp_global = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int64_t warp_count = (int64_t)gridDim.x * LN_WARPS;
  // Software pipeline: the loads of the NEXT row group are issued before this one is reduced, so every
  // warp keeps two row groups in flight (one was ~70 % of the bytes in flight the HBM latency asks for).
  RawN<TIn, NE> nx[CPL], nr[CPL];
  auto fetch = [&](int64_t rr) {
    const bool okn = rr < rows;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      nx[j].zero();
      nr[j].zero();
      if (okn) {
        nx[j].load(x + rr * C + (j * LPR + lr) * NE);
        if (res != nullptr) nr[j].load(res + rr * C + (j * LPR + lr) * NE);
      }
    }
  };
  fetch(warp_global * RPW + sub);
  for (int64_t r0 = warp_global * RPW; r0 < rows; r0 += warp_count * RPW) {
    const int64_t r = r0 + sub;
    const bool ok = r < rows;
    float v[CPL][NE], rv_[CPL][NE];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      nx[j].unpack_to(v[j]);
      nr[j].unpack_to(rv_[j]);
    }
    fetch(r + warp_count * RPW);
    if (res != nullptr && ok) {
      // fused residual add (C:367 / C:369 feeding the next pre-norm): s = x + res is written once, in
      // the stream's dtype, and the statistics are taken from the ROUNDED s — exactly what a separate
      // add kernel followed by this LayerNorm would produce
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          v[j][e] += rv_[j][e];
          if constexpr (sizeof(TIn) == 2) v[j][e] = __bfloat162float(__float2bfloat16_rn(v[j][e]));
        }
        store_n<TIn, NE>(sum_out + r * C + (j * LPR + lr) * NE, v[j]);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j)
#pragma unroll
      for (int e = 0; e < NE; ++e) s += v[j][e];
    const float mean = row_sum<LPR>(s) * (1.f / C);
    float m2 = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j)
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const float t = v[j][e] - mean;
        m2 = fmaf(t, t, m2);
      }
    const float rstd = rsqrtf(row_sum<LPR>(m2) * (1.f / C) + eps);
    if (ok) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        float o[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) o[e] = fmaf((v[j][e] - mean) * rstd, g[j][e], b[j][e]);
        store_n<TOut, NE>(y + r * C + (j * LPR + lr) * NE, o);
      }
      if (lr == 0) {
        stats[2 * r] = mean;
        stats[2 * r + 1] = rstd;
      }
    }
  }
}

// RB: also accumulate the column sums of the gradient written to grad_x — the bias gradient of the
// Linear whose output was the `residual` operand of the fused add (proj C:366, Mlp.fc2 C:195): that
// Linear then needs no column-sum pass of its own.  The sums are taken over the ROUNDED values, i.e.
// exactly what a separate pass over the stored tensor would see.
template <typename TIn, typename TGy, typename TGx, int LPR, int CPL, int NE, bool RB>
__global__ void __launch_bounds__(LN_THREADS, CPL == 1 ? 3 : 1)
    layernorm_bwd_kernel(const TIn* __restrict__ x, const TGy* __restrict__ gy,
                         const float* __restrict__ gamma, const float* __restrict__ stats,
                         const TGx* __restrict__ gres, TGx* __restrict__ gx, float* __restrict__ partial,
                         int64_t rows) {
  constexpr int C = LPR * CPL * NE, RPW = 32 / LPR, K = RB ? 3 : 2;
  __shared__ float s_part[LN_WARPS][2 * C];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lr = lane % LPR, sub = lane / LPR;
  float g[CPL][NE], dg[CPL][NE], db[CPL][NE], dr[RB ? CPL : 1][NE];
#pragma unroll
  for (int j = 0; j < CPL; ++j)
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      g[j][e] = __ldg(gamma + (j * LPR + lr) * NE + e);
      dg[j][e] = db[j][e] = 0.f;
      if constexpr (RB) dr[j][e] = 0.f;
    }
  const int64_t warp_global = (int64_t)blockIdx.x * LN_WARPS + warp;
  const int64_t warp_count = (int64_t)gridDim.x * LN_WARPS;
  RawN<TIn, NE> nxv[CPL];
  RawN<TGy, NE> ngv[CPL];
  // the gradient arriving on the residual stream is prefetched with the other two operands (it used to be
  // loaded where it is added: an un-hidden L2 / DRAM round trip per row group, 52 of the 58 calls per step)
  constexpr int NEG = 16 / (int)sizeof(TGx) < NE ? 16 / (int)sizeof(TGx) : NE;  // elements per 16-byte load of gres
  RawN<TGx, NEG> nrv[CPL][NE / NEG];
  float nmean = 0.f, nrstd = 0.f;
  auto fetch = [&](int64_t rr) {
    const bool okn = rr < rows;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      nxv[j].zero();
      ngv[j].zero();
      if (okn) {
        nxv[j].load(x + rr * C + (j * LPR + lr) * NE);
        ngv[j].load(gy + rr * C + (j * LPR + lr) * NE);
      }
      if (gres != nullptr) {
#pragma unroll
        for (int h = 0; h < NE / NEG; ++h) {
          nrv[j][h].zero();
          if (okn) nrv[j][h].load(gres + rr * C + (j * LPR + lr) * NE + h * NEG);
        }
      }
    }
    nmean = okn ? __ldg(stats + 2 * rr) : 0.f;
    nrstd = okn ? __ldg(stats + 2 * rr + 1) : 0.f;
  };
  fetch(warp_global * RPW + sub);
  for (int64_t r0 = warp_global * RPW; r0 < rows; r0 += warp_count * RPW) {
    const int64_t r = r0 + sub;
    const bool ok = r < rows;
    float xv[CPL][NE], gv[CPL][NE], rsv[CPL][NE];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      nxv[j].unpack_to(xv[j]);
      ngv[j].unpack_to(gv[j]);
      if (gres != nullptr) {
#pragma unroll
        for (int h = 0; h < NE / NEG; ++h) {
          float t[NEG];
          nrv[j][h].unpack_to(t);
#pragma unroll
          for (int e = 0; e < NEG; ++e) rsv[j][h * NEG + e] = t[e];
        }
      }
    }
    const float mean = nmean, rstd = nrstd;
    fetch(r + warp_count * RPW);  // next row group in flight while this one is reduced
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j)
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const float xh = (xv[j][e] - mean) * rstd;
        const float gg = gv[j][e] * g[j][e];
        dg[j][e] = fmaf(gv[j][e], xh, dg[j][e]);  // zero rows contribute nothing
        db[j][e] += gv[j][e];
        s1 += gg;
        s2 = fmaf(gg, xh, s2);
        xv[j][e] = xh;
        gv[j][e] = gg;
      }
    const float c1 = row_sum<LPR>(s1) * (1.f / C), c2 = row_sum<LPR>(s2) * (1.f / C);
    if (ok) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        float o[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) o[e] = rstd * (gv[j][e] - c1 - xv[j][e] * c2);
        if (gres != nullptr) {  // gradient arriving on the residual stream itself: summed here, not by autograd
#pragma unroll
          for (int e = 0; e < NE; ++e) o[e] += rsv[j][e];
        }
        store_n<TGx, NE>(gx + r * C + (j * LPR + lr) * NE, o);
        if constexpr (RB) {
#pragma unroll
          for (int e = 0; e < NE; ++e)
            dr[j][e] += sizeof(TGx) == 2 ? __bfloat162float(__float2bfloat16_rn(o[e])) : o[e];
        }
      }
    }
  }
  // gamma / beta gradients: over the RPW row slots of the warp, then over the warps of the CTA
#pragma unroll
  for (int j = 0; j < CPL; ++j)
#pragma unroll
    for (int e = 0; e < NE; ++e) {
#pragma unroll
      for (int o = 16; o >= LPR; o >>= 1) {
        dg[j][e] += __shfl_xor_sync(0xffffffffu, dg[j][e], o);
        db[j][e] += __shfl_xor_sync(0xffffffffu, db[j][e], o);
      }
    }
  if (sub == 0) {
#pragma unroll
    for (int j = 0; j < CPL; ++j)
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        s_part[warp][(j * LPR + lr) * NE + e] = dg[j][e];
        s_part[warp][C + (j * LPR + lr) * NE + e] = db[j][e];
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += LN_THREADS) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) a += s_part[w][i];
    partial[(int64_t)blockIdx.x * K * C + i] = a;
  }
  if constexpr (RB) {
    __syncthreads();  // s_part is reused for the third vector
#pragma unroll
    for (int j = 0; j < CPL; ++j)
#pragma unroll
      for (int e = 0; e < NE; ++e) {
#pragma unroll
        for (int o = 16; o >= LPR; o >>= 1) dr[j][e] += __shfl_xor_sync(0xffffffffu, dr[j][e], o);
      }
    if (sub == 0) {
#pragma unroll
      for (int j = 0; j < CPL; ++j)
#pragma unroll
        for (int e = 0; e < NE; ++e) s_part[warp][(j * LPR + lr) * NE + e] = dr[j][e];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += LN_THREADS) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < LN_WARPS; ++w) a += s_part[w][i];
      partial[(int64_t)blockIdx.x * K * C + 2 * C + i] = a;
    }
  }
}

// one warp per output: lanes stride over the per-CTA partials (coalescing is across neighbouring
// warps), fixed summation order -> deterministic
__global__ void __launch_bounds__(256)
    layernorm_param_grad_final(const float* __restrict__ partial, int blocks, int C, int K,
                               float* __restrict__ ggamma, float* __restrict__ gbeta,
                               float* __restrict__ grbias) {
  const int i = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= K * C) return;
  const float a = strided_partial_sum(partial + i, blocks, (int64_t)K * C, lane);
  if (lane == 0) {
    if (i < C) ggamma[i] = a;
    else if (i < 2 * C) gbeta[i - C] = a;
    else grbias[i - 2 * C] = a;
  }
}

constexpr int LN_MAX_GRID = 148 * 4;  // persistent: 4 CTAs of 256 threads per B200 SM (forward)
// the backward kernels hold the gamma / beta (/ residual-bias) accumulators and the prefetched next row
// group: 80 registers, 3 CTAs per SM (__launch_bounds__) — a grid of 4 per SM would run as 1.33 waves
constexpr int LN_MAX_GRID_BWD = 148 * 3;
int ln_grid(int64_t rows, int rpw, int max_grid = LN_MAX_GRID) {
  const int64_t need = (rows + (int64_t)rpw * LN_WARPS - 1) / ((int64_t)rpw * LN_WARPS);
  return (int)(need < max_grid ? (need > 0 ? need : 1) : max_grid);
}

// vectors of 16 B of the INPUT per row decide the tiling: LPR lanes per row, CPL chunks per lane
template <typename TIn>
bool ln_shape(int64_t C, int* lpr, int* cpl) {
  constexpr int NE = 16 / sizeof(TIn);
  if (C % NE != 0) return false;
  const int64_t nv = C / NE;
  if (nv == 8 || nv == 16 || nv == 32) { *lpr = (int)nv; *cpl = 1; return true; }
  if (nv == 64) { *lpr = 32; *cpl = 2; return true; }
  if (nv == 128 && NE == 4) { *lpr = 32; *cpl = 4; return true; }  // fp32 C = 512
  return false;
}

#define LN_DISPATCH_SHAPE(CALL)                                  \
  if (lpr == 8 && cpl == 1) { CALL(8, 1); }                      \
  else if (lpr == 16 && cpl == 1) { CALL(16, 1); }               \
  else if (lpr == 32 && cpl == 1) { CALL(32, 1); }               \
  else if (lpr == 32 && cpl == 2) { CALL(32, 2); }               \
  else { if constexpr (NE == 4) { CALL(32, 4); } }

template <typename TIn, typename TOut>
int ln_fwd_t(const void* x, const void* res, void* sum_out, const float* gamma, const float* beta, void* y,
             float* stats, int64_t rows, int64_t C, float eps, cudaStream_t st) {
  constexpr int NE = 16 / sizeof(TIn);
  int lpr, cpl;
  if (!ln_shape<TIn>(C, &lpr, &cpl))
    return fail(CSB200_ERR_UNSUPPORTED, "layernorm: C=%lld is not tiled for this dtype", (long long)C);
#define CALL(L, P)                                                                              \
  layernorm_fwd_kernel<TIn, TOut, L, P, NE><<<ln_grid(rows, 32 / L), LN_THREADS, 0, st>>>(      \
      static_cast<const TIn*>(x), static_cast<const TIn*>(res), static_cast<TIn*>(sum_out), gamma, beta, \
      static_cast<TOut*>(y), stats, rows, eps)
  LN_DISPATCH_SHAPE(CALL)
#undef CALL
  return check_launch("layernorm_fwd_kernel");
}

template <typename TIn, typename TGy, typename TGx>
int ln_bwd_t(const void* x, const void* gy, const void* gres, const float* gamma, const float* stats,
             void* gx, float* ggamma, float* gbeta, float* grbias, float* partial, int64_t rows, int64_t C,
             cudaStream_t st, bool rb, int32_t* partial_rows) {
  constexpr int NE = 16 / sizeof(TIn);
  int lpr, cpl;
  if (!ln_shape<TIn>(C, &lpr, &cpl))
    return fail(CSB200_ERR_UNSUPPORTED, "layernorm: C=%lld is not tiled for this dtype", (long long)C);
  const int grid = ln_grid(rows, 32 / lpr, cpl == 1 ? LN_MAX_GRID_BWD : 148);  // wide rows: 150+ registers, 1 CTA / SM
#define CALL(L, P)                                                                       \
  do {                                                                                   \
    if (rb)                                                                              \
      layernorm_bwd_kernel<TIn, TGy, TGx, L, P, NE, true><<<grid, LN_THREADS, 0, st>>>(  \
          static_cast<const TIn*>(x), static_cast<const TGy*>(gy), gamma, stats,         \
          static_cast<const TGx*>(gres), static_cast<TGx*>(gx), partial, rows);          \
    else                                                                                 \
      layernorm_bwd_kernel<TIn, TGy, TGx, L, P, NE, false><<<grid, LN_THREADS, 0, st>>>( \
          static_cast<const TIn*>(x), static_cast<const TGy*>(gy), gamma, stats,         \
          static_cast<const TGx*>(gres), static_cast<TGx*>(gx), partial, rows);          \
  } while (0)
  LN_DISPATCH_SHAPE(CALL)
#undef CALL
  int rc = check_launch("layernorm_bwd_kernel");
  if (rc != CSB200_OK) return rc;
  if (partial_rows != nullptr) {  // deferred: the caller records the final sums (csb200_sum_rows_deferred)
    *partial_rows = grid;
    return CSB200_OK;
  }
  const int K = rb ? 3 : 2;
  layernorm_param_grad_final<<<(int)((K * C * 32 + 255) / 256), 256, 0, st>>>(partial, grid, (int)C, K,
                                                                         ggamma, gbeta, grbias);
  return check_launch("layernorm_param_grad_final");
}

bool ok_dtype(int d) { return d == CSB200_F32 || d == CSB200_BF16; }

}  // namespace
}  // namespace csb200

using namespace csb200;
using bf16 = __nv_bfloat16;

extern "C" int csb200_layernorm_supported(int64_t channels, int x_dtype) {
  int lpr, cpl;
  if (x_dtype == CSB200_F32) return ln_shape<float>(channels, &lpr, &cpl) ? 1 : 0;
  if (x_dtype == CSB200_BF16) return ln_shape<bf16>(channels, &lpr, &cpl) ? 1 : 0;
  return 0;
}

static int ln_fwd_dispatch(const void* x, const void* res, void* sum_out, const float* gamma,
                           const float* beta, void* y, float* stats, int64_t rows, int64_t channels,
                           int x_dtype, int y_dtype, float eps, void* stream, const char* who) {
  if (rows < 0 || channels <= 0 || !ok_dtype(x_dtype) || !ok_dtype(y_dtype))
    return fail(CSB200_ERR_INVALID, "%s: bad size or dtype", who);
  if (rows == 0) return CSB200_OK;
  if (!x || !gamma || !beta || !y || !stats || ((res == nullptr) != (sum_out == nullptr)))
    return fail(CSB200_ERR_INVALID, "%s: null pointer", who);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (x_dtype == CSB200_F32)
    return y_dtype == CSB200_F32
               ? ln_fwd_t<float, float>(x, res, sum_out, gamma, beta, y, stats, rows, channels, eps, st)
               : ln_fwd_t<float, bf16>(x, res, sum_out, gamma, beta, y, stats, rows, channels, eps, st);
  return y_dtype == CSB200_F32
             ? ln_fwd_t<bf16, float>(x, res, sum_out, gamma, beta, y, stats, rows, channels, eps, st)
             : ln_fwd_t<bf16, bf16>(x, res, sum_out, gamma, beta, y, stats, rows, channels, eps, st);
}

extern "C" int csb200_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y,
                                    float* stats, int64_t rows, int64_t channels, int x_dtype,
                                    int y_dtype, float eps, void* stream) {
  return ln_fwd_dispatch(x, nullptr, nullptr, gamma, beta, y, stats, rows, channels, x_dtype, y_dtype, eps,
                         stream, "layernorm_fwd");
}

extern "C" int csb200_add_layernorm_fwd(const void* x, const void* residual, void* sum_out,
                                        const float* gamma, const float* beta, void* y, float* stats,
                                        int64_t rows, int64_t channels, int x_dtype, int y_dtype, float eps,
                                        void* stream) {
  if (rows > 0 && (!residual || !sum_out)) return fail(CSB200_ERR_INVALID, "add_layernorm_fwd: null pointer");
  return ln_fwd_dispatch(x, residual, sum_out, gamma, beta, y, stats, rows, channels, x_dtype, y_dtype, eps,
                         stream, "add_layernorm_fwd");
}

extern "C" size_t csb200_layernorm_bwd_workspace_bytes(int64_t rows, int64_t channels) {
  (void)rows;
  return (size_t)LN_MAX_GRID * 3 * (size_t)channels * sizeof(float) + 256;  // per-CTA partials
}

static int ln_bwd_dispatch(const void* x, const void* grad_y, const void* grad_res, const float* gamma,
                           const float* stats, void* grad_x, float* grad_gamma, float* grad_beta,
                           float* grad_res_bias, void* workspace, size_t workspace_bytes, int64_t rows,
                           int64_t channels, int x_dtype, int gy_dtype, void* stream, bool rb = false,
                           int32_t* partial_rows = nullptr) {
  if (rows < 0 || channels <= 0 || !ok_dtype(x_dtype) || !ok_dtype(gy_dtype))
    return fail(CSB200_ERR_INVALID, "layernorm_bwd: bad size or dtype");
  const bool deferred = partial_rows != nullptr;
  if (!deferred) rb = grad_res_bias != nullptr;
  if (!x || !grad_y || !gamma || !stats || !grad_x || (!deferred && (!grad_gamma || !grad_beta)) || !workspace)
    return fail(CSB200_ERR_INVALID, "layernorm_bwd: null pointer");
  if (deferred && rows == 0) return fail(CSB200_ERR_INVALID, "layernorm_bwd_partials: rows == 0 has no partial rows");
  if (workspace_bytes < csb200_layernorm_bwd_workspace_bytes(rows, channels))
    return fail(CSB200_ERR_WORKSPACE, "layernorm_bwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(workspace);
  if (rows == 0) {
    CSB200_CUDA(cudaMemsetAsync(grad_gamma, 0, channels * sizeof(float), st));
    CSB200_CUDA(cudaMemsetAsync(grad_beta, 0, channels * sizeof(float), st));
    if (grad_res_bias) CSB200_CUDA(cudaMemsetAsync(grad_res_bias, 0, channels * sizeof(float), st));
    return CSB200_OK;
  }
  // grad_x (and grad_res) have the type of x
  if (x_dtype == CSB200_F32)
    return gy_dtype == CSB200_F32
               ? ln_bwd_t<float, float, float>(x, grad_y, grad_res, gamma, stats, grad_x, grad_gamma,
                                               grad_beta, grad_res_bias, partial, rows, channels, st, rb, partial_rows)
               : ln_bwd_t<float, bf16, float>(x, grad_y, grad_res, gamma, stats, grad_x, grad_gamma,
                                              grad_beta, grad_res_bias, partial, rows, channels, st, rb, partial_rows);
  return gy_dtype == CSB200_F32
             ? ln_bwd_t<bf16, float, bf16>(x, grad_y, grad_res, gamma, stats, grad_x, grad_gamma, grad_beta,
                                           grad_res_bias, partial, rows, channels, st, rb, partial_rows)
             : ln_bwd_t<bf16, bf16, bf16>(x, grad_y, grad_res, gamma, stats, grad_x, grad_gamma, grad_beta,
                                          grad_res_bias, partial, rows, channels, st, rb, partial_rows);
}

extern "C" int csb200_layernorm_bwd(const void* x, const void* grad_y, const float* gamma,
                                    const float* stats, void* grad_x, float* grad_gamma,
                                    float* grad_beta, void* workspace, size_t workspace_bytes,
                                    int64_t rows, int64_t channels, int x_dtype, int gy_dtype,
                                    void* stream) {
  return ln_bwd_dispatch(x, grad_y, nullptr, gamma, stats, grad_x, grad_gamma, grad_beta, nullptr, workspace,
                         workspace_bytes, rows, channels, x_dtype, gy_dtype, stream);
}

extern "C" int csb200_add_layernorm_bwd(const void* sum, const void* grad_y, const void* grad_sum,
                                        const float* gamma, const float* stats, void* grad_x,
                                        float* grad_gamma, float* grad_beta, void* workspace,
                                        size_t workspace_bytes, int64_t rows, int64_t channels, int x_dtype,
                                        int gy_dtype, void* stream) {
  return ln_bwd_dispatch(sum, grad_y, grad_sum, gamma, stats, grad_x, grad_gamma, grad_beta, nullptr, workspace,
                         workspace_bytes, rows, channels, x_dtype, gy_dtype, stream);
}

extern "C" int csb200_add_layernorm_bwd_rb(const void* sum, const void* grad_y, const void* grad_sum,
                                           const float* gamma, const float* stats, void* grad_x,
                                           float* grad_gamma, float* grad_beta, float* grad_res_bias,
                                           void* workspace, size_t workspace_bytes, int64_t rows,
                                           int64_t channels, int x_dtype, int gy_dtype, void* stream) {
  if (!grad_res_bias) return fail(CSB200_ERR_INVALID, "add_layernorm_bwd_rb: null pointer");
  return ln_bwd_dispatch(sum, grad_y, grad_sum, gamma, stats, grad_x, grad_gamma, grad_beta, grad_res_bias,
                         workspace, workspace_bytes, rows, channels, x_dtype, gy_dtype, stream);
}

// The same pass without its last launch: the per-CTA partial sums stay in the workspace as
// float[*partial_rows][(2 + with_res_bias) * channels] for csb200_sum_rows_deferred / _flush (sum_rows.cu).
extern "C" int csb200_layernorm_bwd_partials(const void* sum_or_x, const void* grad_y, const void* grad_sum,
                                             const float* gamma, const float* stats, void* grad_x,
                                             int with_res_bias, void* workspace, size_t workspace_bytes,
                                             int64_t rows, int64_t channels, int x_dtype, int gy_dtype,
                                             const float** partials, int32_t* partial_rows, void* stream) {
  if (!partials || !partial_rows) return fail(CSB200_ERR_INVALID, "layernorm_bwd_partials: null pointer");
  *partials = static_cast<const float*>(workspace);
  *partial_rows = 0;
  return ln_bwd_dispatch(sum_or_x, grad_y, grad_sum, gamma, stats, grad_x, nullptr, nullptr, nullptr, workspace,
                         workspace_bytes, rows, channels, x_dtype, gy_dtype, stream, with_res_bias != 0, partial_rows);
}

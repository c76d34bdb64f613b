// SimAM forward / backward for sm_100a.
//
// SimAM is not in the reference checkout; the arithmetic follows the public module
// (Yang et al., ICML 2021) restated in oracle/ops.py.  Both passes are HBM-bound: the design
// goal is algorithmic traffic only — forward 1 read + 1 write, backward 2 reads + 1 write — which
// means a plane must stay on chip between the statistics and the rescale.
//
// Fast paths ("resident"): every thread keeps its share of the plane in REGISTERS as 16-byte
// vectors, all loads are issued before the first use (deep memory-level parallelism), the mean and
// the centred second moment are reduced by warp shuffles -> shared memory -> (for planes larger than
// one CTA can hold) DSMEM across a thread-block cluster of 2/4/8 CTAs, and the result is written
// straight from the same registers.  The variance is the exact two-pass form sum((x-mean)^2) — the
// second pass costs nothing because the data is already in registers — so large-mean inputs keep
// the 1e-5 fp32 parity target.
//
//   NCHW (UNet, U:177-250): a plane is H*W contiguous elements.
//   NLC  (CSWin tokens, C:349): a plane is a column of a (L, C) matrix; a CTA cluster owns a slab of
//        32/64/128-byte-wide row segments and reduces per channel.
//
// Generic paths: any shape / alignment, three sweeps (the 2nd and 3rd normally hit L2).

#include "common.cuh"

namespace csb200 {
namespace {

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ------------------------------------------------------------------------------------------------
// shared arithmetic
// ------------------------------------------------------------------------------------------------
template <typename T>
struct Sig {
  // fp32: ex2.approx + rcp.approx based; ~3e-7 relative error, well inside the 1e-5 target
  static __device__ __forceinline__ float f(float y) { return __fdividef(1.f, 1.f + __expf(-y)); }
};
template <>
struct Sig<__nv_bfloat16> {
  // bf16 output has 8 mantissa bits; a single MUFU.TANH (2^-11 abs error) is 4x below its rounding
  static __device__ __forceinline__ float f(float y) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * y));
    return fmaf(0.5f, t, 0.5f);
  }
};

// The mean is carried as (pivot, dmean) with pivot = the plane's first element and
// dmean = mean - pivot: x - pivot is exact in fp32 for nearby values (Sterbenz), so the centred value
// t = (x - pivot) - dmean keeps full relative accuracy even when |mean| >> std.
struct FwdCoef {
  float pivot, dmean, inv4v;
};
__device__ __forceinline__ float centred(float x, float pivot, float dmean) { return (x - pivot) - dmean; }
template <typename T>
__device__ __forceinline__ float simam_fwd_elem(float x, const FwdCoef& c) {
  float t = centred(x, c.pivot, c.dmean);
  return x * Sig<T>::f(fmaf(t * t, c.inv4v, 0.5f));
}
// bf16 fast form: sigmoid(e) = 0.5 tanh(e/2) + 0.5 with the halves folded into the coefficients,
//   y = hx * tanh(t^2 * inv8v + 0.25) + hx,  hx = x / 2,  t = x - mean  (7 issue slots + 1 MUFU)
__device__ __forceinline__ float simam_fwd_fast(float x, float mean, float inv8v) {
  const float t = x - mean;
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(fmaf(t * t, inv8v, 0.25f)));
  const float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}

// The bf16 backward is bound by instruction issue, not by HBM (~36 slots per element over two sweeps):
// the whole per-element chain runs on packed fp32 pairs (common.cuh) and only the two tanh stay scalar.
__device__ __forceinline__ f2_t f2_tanh(f2_t a) {
  float lo, hi;
  f2_split(a, lo, hi);
  asm("tanh.approx.f32 %0, %0;" : "+f"(lo));
  asm("tanh.approx.f32 %0, %0;" : "+f"(hi));
  return f2_make(lo, hi);
}

// ------------------------------------------------------------------------------------------------
// block / cluster reduction of NV scalars held by every thread (NCHW: one plane per CTA/cluster)
// ------------------------------------------------------------------------------------------------
template <int THREADS, int CLUSTER, int NV>
__device__ __forceinline__ void plane_reduce(float (&v)[NV], float* s_warp /*[NV][32]*/,
                                             float* s_cta /*[NV]*/) {
  constexpr int WARPS = THREADS / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) s_warp[i * 32 + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) a += s_warp[i * 32 + w];  // broadcast reads, fixed order
    v[i] = a;
  }
  if constexpr (CLUSTER > 1) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i) s_cta[i] = v[i];
    }
    cluster_sync_all();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float a = 0.f;
#pragma unroll
      for (int r = 0; r < CLUSTER; ++r) a += dsmem_ld_f32(&s_cta[i], r);  // same order everywhere
      v[i] = a;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// NCHW resident forward.  WARP_PLANE: one warp per plane (small planes), else one CTA / cluster.
// ------------------------------------------------------------------------------------------------
template <typename T, int VPT, int THREADS, int CLUSTER, bool WARP_PLANE>
__global__ void __launch_bounds__(THREADS, (THREADS <= 256 ? 4 : 2))
    simam_nchw_fwd_resident(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ stats,
                            int64_t planes, int nvec, float S, float e_lambda) {
  constexpr int VE = Vec16<T>::N;
  __shared__ float s_warp[2 * 32];
  __shared__ float s_cta[2];

  int64_t plane;
  int v0, vstride;
  if constexpr (WARP_PLANE) {
    plane = (int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
    v0 = threadIdx.x & 31;
    vstride = 32;
    if (plane >= planes) return;  // whole warp leaves; no block-level sync on this path
  } else {
    plane = blockIdx.x / CLUSTER;
    const int rank = CLUSTER > 1 ? (int)cluster_ctarank() : 0;
    v0 = rank * (THREADS * VPT) + threadIdx.x;
    vstride = THREADS;
  }
  const uint4* xp = reinterpret_cast<const uint4*>(x) + plane * nvec;
  uint4* yp = reinterpret_cast<uint4*>(y) + plane * nvec;

  uint4 d[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    int vi = v0 + i * vstride;
    d[i] = (vi < nvec) ? ld_stream(xp + vi) : make_uint4(0, 0, 0, 0);
  }
  const float pivot = to_f32(__ldg(x + plane * nvec * VE));  // first element of the plane
  constexpr bool ONEPASS = sizeof(T) == 2;
  float dmean, v;
  FwdCoef c;
  if constexpr (ONEPASS) {
    // bf16 inputs carry 8 significant bits: shifted single-pass moments in fp32 are exact enough
    // (sum d, sum d^2 with d = x - pivot) and need ONE reduction round instead of two
    float red[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      if (v0 + i * vstride < nvec) {
        float f[VE];
        unpack<T>(d[i], f);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          const float t = f[e] - pivot;
          red[0] += t;
          red[1] = fmaf(t, t, red[1]);
        }
      }
    }
    if constexpr (WARP_PLANE) {
      red[0] = warp_sum(red[0]);
      red[1] = warp_sum(red[1]);
    } else {
      plane_reduce<THREADS, CLUSTER, 2>(red, s_warp, s_cta);
    }
    dmean = red[0] / S;
    v = fmaxf(red[1] - red[0] * dmean, 0.f) / (S - 1.f) + e_lambda;
    c = FwdCoef{0.f, pivot + dmean, 1.f / (8.f * v)};  // (unused, mean, 1/(8v)) for simam_fwd_fast
  } else {
    // pass A: sum of (x - pivot)
    float red[1] = {0.f};
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      if (v0 + i * vstride < nvec) {
        float f[VE];
        unpack<T>(d[i], f);
#pragma unroll
        for (int e = 0; e < VE; ++e) red[0] += f[e] - pivot;
      }
    }
    if constexpr (WARP_PLANE) red[0] = warp_sum(red[0]);
    else plane_reduce<THREADS, CLUSTER, 1>(red, s_warp, s_cta);
    dmean = red[0] / S;
    // pass B: centred second moment
    red[0] = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      if (v0 + i * vstride < nvec) {
        float f[VE];
        unpack<T>(d[i], f);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          float t = centred(f[e], pivot, dmean);
          red[0] = fmaf(t, t, red[0]);
        }
      }
    }
    if constexpr (WARP_PLANE) red[0] = warp_sum(red[0]);
    else plane_reduce<THREADS, CLUSTER, 1>(red, s_warp + 32, s_cta + 1);
    v = red[0] / (S - 1.f) + e_lambda;
    c = FwdCoef{pivot, dmean, 1.f / (4.f * v)};
  }

#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    int vi = v0 + i * vstride;
    if (vi < nvec) {
      float f[VE];
      unpack<T>(d[i], f);
#pragma unroll
      for (int e = 0; e < VE; ++e) {
        if constexpr (ONEPASS) f[e] = simam_fwd_fast(f[e], c.dmean, c.inv4v);
        else f[e] = simam_fwd_elem<T>(f[e], c);
      }
      st_stream(yp + vi, pack<T>(f));
    }
  }
  if (stats != nullptr && v0 == 0) {
    stats[2 * plane] = dmean;
    stats[2 * plane + 1] = v;
  }
  if constexpr (CLUSTER > 1) cluster_sync_all();  // keep s_cta alive until every peer has read it
}

// NCHW resident backward: holds x and grad_y; one reduction of (R1, R2').
template <typename T, int VPT, int THREADS, int CLUSTER, bool WARP_PLANE>
__global__ void __launch_bounds__(THREADS, (THREADS <= 256 ? (sizeof(T) == 2 && VPT <= 4 ? 3 : 2) : 1))
    simam_nchw_bwd_resident(const T* __restrict__ x, const T* __restrict__ gy,
                            const float* __restrict__ stats, T* __restrict__ gx, int64_t planes,
                            int nvec, float S) {
  constexpr int VE = Vec16<T>::N;
  __shared__ float s_warp[2 * 32];
  __shared__ float s_cta[2];

  int64_t plane;
  int v0, vstride;
  if constexpr (WARP_PLANE) {
    plane = (int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
    v0 = threadIdx.x & 31;
    vstride = 32;
    if (plane >= planes) return;
  } else {
    plane = blockIdx.x / CLUSTER;
    const int rank = CLUSTER > 1 ? (int)cluster_ctarank() : 0;
    v0 = rank * (THREADS * VPT) + threadIdx.x;
    vstride = THREADS;
  }
  const uint4* xp = reinterpret_cast<const uint4*>(x) + plane * nvec;
  const uint4* gp = reinterpret_cast<const uint4*>(gy) + plane * nvec;
  uint4* op = reinterpret_cast<uint4*>(gx) + plane * nvec;

  uint4 dx[VPT], dg[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    int vi = v0 + i * vstride;
    bool ok = vi < nvec;
    dx[i] = ok ? ld_stream(xp + vi) : make_uint4(0, 0, 0, 0);
    dg[i] = ok ? ld_stream(gp + vi) : make_uint4(0, 0, 0, 0);
  }
  const float pivot = to_f32(__ldg(x + plane * nvec * VE));
  const float dmean = __ldg(stats + 2 * plane), v = __ldg(stats + 2 * plane + 1);
  const float inv4v = 1.f / (4.f * v);

  float red[2] = {0.f, 0.f};  // R1 = sum a*d, R2' = sum a*t  (grad of zero-padded slots is 0 -> a = 0)
  if constexpr (sizeof(T) == 2) {
    // bf16: the pass-1 products g*s and a = g x s (1-s) are kept, packed to bf16, in the registers
    // that held grad_y (+ VPT more), so pass 2 is 8 issue slots per element instead of a recompute.
    const float mean = pivot + dmean, inv8v = 0.5f * inv4v;
    uint4 da[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      float fx[VE], fg[VE], fa[VE];
      unpack<T>(dx[i], fx);
      unpack<T>(dg[i], fg);
#pragma unroll
      for (int e = 0; e < VE; ++e) {
        const float t = fx[e] - mean, dd = t * t;
        float th;
        asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(fmaf(dd, inv8v, 0.25f)));
        const float s = fmaf(0.5f, th, 0.5f);
        const float gs = fg[e] * s;
        const float a = gs * fx[e] * (1.f - s);
        red[0] = fmaf(a, dd, red[0]);
        red[1] = fmaf(a, t, red[1]);
        fg[e] = gs;
        fa[e] = a;
      }
      dg[i] = pack<T>(fg);
      da[i] = pack<T>(fa);
    }
    if constexpr (WARP_PLANE) {
      red[0] = warp_sum(red[0]);
      red[1] = warp_sum(red[1]);
    } else {
      plane_reduce<THREADS, CLUSTER, 2>(red, s_warp, s_cta);
    }
    const float c1 = red[0] * inv4v / (v * (S - 1.f));  // R1 / (4 v^2 n)
    const float c2 = 2.f / S * red[1] * inv4v;          // (2/HW) R2
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int vi = v0 + i * vstride;
      if (vi < nvec) {
        float fx[VE], fg[VE], fa[VE];
        unpack<T>(dx[i], fx);
        unpack<T>(dg[i], fg);
        unpack<T>(da[i], fa);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          const float t2 = 2.f * (fx[e] - mean);
          fx[e] = fmaf(t2, fmaf(fa[e], inv4v, -c1), fg[e]) - c2;
        }
        st_stream(op + vi, pack<T>(fx));
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      float fx[VE], fg[VE];
      unpack<T>(dx[i], fx);
      unpack<T>(dg[i], fg);
#pragma unroll
      for (int e = 0; e < VE; ++e) {
        float t = centred(fx[e], pivot, dmean), dd = t * t;
        float s = Sig<T>::f(fmaf(dd, inv4v, 0.5f));
        float a = fg[e] * fx[e] * s * (1.f - s);
        red[0] = fmaf(a, dd, red[0]);
        red[1] = fmaf(a, t, red[1]);
      }
    }
    if constexpr (WARP_PLANE) {
      red[0] = warp_sum(red[0]);
      red[1] = warp_sum(red[1]);
    } else {
      plane_reduce<THREADS, CLUSTER, 2>(red, s_warp, s_cta);
    }
    const float c1 = red[0] * inv4v / (v * (S - 1.f));  // R1 / (4 v^2 n)
    const float c2 = 2.f / S * red[1] * inv4v;          // (2/HW) R2
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      int vi = v0 + i * vstride;
      if (vi < nvec) {
        float fx[VE], fg[VE];
        unpack<T>(dx[i], fx);
        unpack<T>(dg[i], fg);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          float t = centred(fx[e], pivot, dmean), dd = t * t;
          float s = Sig<T>::f(fmaf(dd, inv4v, 0.5f));
          float a = fg[e] * fx[e] * s * (1.f - s);
          fx[e] = fmaf(fg[e], s, 2.f * t * fmaf(a, inv4v, -c1)) - c2;
        }
        st_stream(op + vi, pack<T>(fx));
      }
    }
  }
  if constexpr (CLUSTER > 1) cluster_sync_all();
}

// ------------------------------------------------------------------------------------------------
// NLC resident: a cluster owns (image b, slab of CWV 16-byte vectors per row); per-channel reduce.
// Thread t holds column vector t % CWV of rows  (rank*VPT + i) * (THREADS/CWV) + t / CWV.
// ------------------------------------------------------------------------------------------------
template <int VE, int CWV, int THREADS, int CLUSTER, int NQ>
__device__ __forceinline__ void slab_reduce(float (&acc)[NQ][VE], float* s_red /*[WARPS][NQ*CW]*/,
                                            float* s_part /*[NQ*CW]*/, float* s_fin /*[NQ*CW]*/) {
  constexpr int CW = CWV * VE;  // channels in the slab
  constexpr int WARPS = THREADS / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int e = 0; e < VE; ++e)
#pragma unroll
      for (int o = 16; o >= CWV; o >>= 1) acc[q][e] += __shfl_xor_sync(0xffffffffu, acc[q][e], o);
  if (lane < CWV) {
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int e = 0; e < VE; ++e) s_red[warp * (NQ * CW) + q * CW + lane * VE + e] = acc[q][e];
  }
  __syncthreads();
  if (threadIdx.x < NQ * CW) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) a += s_red[w * (NQ * CW) + threadIdx.x];
    if constexpr (CLUSTER > 1) s_part[threadIdx.x] = a;
    else s_fin[threadIdx.x] = a;
  }
  if constexpr (CLUSTER > 1) {
    cluster_sync_all();
    if (threadIdx.x < NQ * CW) {
      float a = 0.f;
#pragma unroll
      for (int r = 0; r < CLUSTER; ++r) a += dsmem_ld_f32(&s_part[threadIdx.x], r);
      s_fin[threadIdx.x] = a;
    }
  }
  __syncthreads();
  const int cv = threadIdx.x % CWV;
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int e = 0; e < VE; ++e) acc[q][e] = s_fin[q * CW + cv * VE + e];
}

template <typename T, int VPT, int CWV, int THREADS, int CLUSTER>
__global__ void __launch_bounds__(THREADS, 2)
    simam_nlc_fwd_resident(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ stats,
                           int L, int C, int slabs, float e_lambda) {
  constexpr int VE = Vec16<T>::N;
  constexpr int CW = CWV * VE;
  constexpr int RPI = THREADS / CWV;  // rows per iteration
  constexpr int WARPS = THREADS / 32;
  __shared__ float s_red[WARPS * CW];
  __shared__ float s_part[2][CW];
  __shared__ float s_fin[CW];

  const int group = blockIdx.x / CLUSTER;  // (b, slab)
  const int rank = CLUSTER > 1 ? (int)cluster_ctarank() : 0;
  const int b = group / slabs, slab = group % slabs;
  const int cv = threadIdx.x % CWV;
  const int r0 = rank * VPT * RPI + threadIdx.x / CWV;
  const int cvec = C / VE;  // vectors per row
  const int64_t base = (int64_t)b * L * cvec + slab * CWV + cv;
  const uint4* xp = reinterpret_cast<const uint4*>(x) + base;
  uint4* yp = reinterpret_cast<uint4*>(y) + base;

  uint4 d[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    int r = r0 + i * RPI;
    d[i] = (r < L) ? ld_stream(xp + (int64_t)r * cvec) : make_uint4(0, 0, 0, 0);
  }
  float pivot[VE];  // row 0 of this thread's channel vector (same address for the whole column)
  unpack<T>(__ldg(xp), pivot);
  float acc[1][VE];
#pragma unroll
  for (int e = 0; e < VE; ++e) acc[0][e] = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    if (r0 + i * RPI < L) {
      float f[VE];
      unpack<T>(d[i], f);
#pragma unroll
      for (int e = 0; e < VE; ++e) acc[0][e] += f[e] - pivot[e];
    }
  }
  slab_reduce<VE, CWV, THREADS, CLUSTER, 1>(acc, s_red, s_part[0], s_fin);
  float mean[VE];  // mean - pivot
#pragma unroll
  for (int e = 0; e < VE; ++e) {
    mean[e] = acc[0][e] / (float)L;
    acc[0][e] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    if (r0 + i * RPI < L) {
      float f[VE];
      unpack<T>(d[i], f);
#pragma unroll
      for (int e = 0; e < VE; ++e) {
        float t = centred(f[e], pivot[e], mean[e]);
        acc[0][e] = fmaf(t, t, acc[0][e]);
      }
    }
  }
  __syncthreads();  // s_fin is re-used by the second reduction
  slab_reduce<VE, CWV, THREADS, CLUSTER, 1>(acc, s_red, s_part[1], s_fin);
  FwdCoef c[VE];
#pragma unroll
  for (int e = 0; e < VE; ++e) {
    float v = acc[0][e] / ((float)L - 1.f) + e_lambda;
    c[e].pivot = pivot[e];
    c[e].dmean = mean[e];
    c[e].inv4v = 1.f / (4.f * v);
    acc[0][e] = v;
  }
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    int r = r0 + i * RPI;
    if (r < L) {
      float f[VE];
      unpack<T>(d[i], f);
#pragma unroll
      for (int e = 0; e < VE; ++e) f[e] = simam_fwd_elem<T>(f[e], c[e]);
      st_stream(yp + (int64_t)r * cvec, pack<T>(f));
    }
  }
  if (stats != nullptr && rank == 0 && threadIdx.x < CWV) {
    const int64_t p0 = (int64_t)b * C + slab * CW + cv * VE;
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      stats[2 * (p0 + e)] = mean[e];
      stats[2 * (p0 + e) + 1] = acc[0][e];
    }
  }
  if constexpr (CLUSTER > 1) cluster_sync_all();
}

template <typename T, int VPT, int CWV, int THREADS, int CLUSTER>
__global__ void __launch_bounds__(THREADS)
    simam_nlc_bwd_resident(const T* __restrict__ x, const T* __restrict__ gy,
                           const float* __restrict__ stats, T* __restrict__ gx, int L, int C,
                           int slabs) {
  constexpr int VE = Vec16<T>::N;
  constexpr int CW = CWV * VE;
  constexpr int RPI = THREADS / CWV;
  constexpr int WARPS = THREADS / 32;
  __shared__ float s_red[WARPS * 2 * CW];
  __shared__ float s_part[2 * CW];
  __shared__ float s_fin[2 * CW];

  const int group = blockIdx.x / CLUSTER;
  const int rank = CLUSTER > 1 ? (int)cluster_ctarank() : 0;
  const int b = group / slabs, slab = group % slabs;
  const int cv = threadIdx.x % CWV;
  const int r0 = rank * VPT * RPI + threadIdx.x / CWV;
  const int cvec = C / VE;
  const int64_t base = (int64_t)b * L * cvec + slab * CWV + cv;
  const uint4* xp = reinterpret_cast<const uint4*>(x) + base;
  const uint4* gp = reinterpret_cast<const uint4*>(gy) + base;
  uint4* op = reinterpret_cast<uint4*>(gx) + base;

  uint4 dx[VPT], dg[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    int r = r0 + i * RPI;
    bool ok = r < L;
    dx[i] = ok ? ld_stream(xp + (int64_t)r * cvec) : make_uint4(0, 0, 0, 0);
    dg[i] = ok ? ld_stream(gp + (int64_t)r * cvec) : make_uint4(0, 0, 0, 0);
  }
  float pivot[VE];
  unpack<T>(__ldg(xp), pivot);
  float mean[VE], v[VE], inv4v[VE];  // mean[] holds mean - pivot
  {
    const int64_t p0 = (int64_t)b * C + slab * CW + cv * VE;
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      mean[e] = __ldg(stats + 2 * (p0 + e));
      v[e] = __ldg(stats + 2 * (p0 + e) + 1);
      inv4v[e] = 1.f / (4.f * v[e]);
    }
  }
  float acc[2][VE];
#pragma unroll
  for (int e = 0; e < VE; ++e) acc[0][e] = acc[1][e] = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    float fx[VE], fg[VE];
    unpack<T>(dx[i], fx);
    unpack<T>(dg[i], fg);
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      float t = centred(fx[e], pivot[e], mean[e]), dd = t * t;
      float s = Sig<T>::f(fmaf(dd, inv4v[e], 0.5f));
      float a = fg[e] * fx[e] * s * (1.f - s);
      acc[0][e] = fmaf(a, dd, acc[0][e]);
      acc[1][e] = fmaf(a, t, acc[1][e]);
    }
  }
  slab_reduce<VE, CWV, THREADS, CLUSTER, 2>(acc, s_red, s_part, s_fin);
  float c1[VE], c2[VE];
#pragma unroll
  for (int e = 0; e < VE; ++e) {
    c1[e] = acc[0][e] * inv4v[e] / (v[e] * ((float)L - 1.f));
    c2[e] = 2.f / (float)L * acc[1][e] * inv4v[e];
  }
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    int r = r0 + i * RPI;
    if (r < L) {
      float fx[VE], fg[VE];
      unpack<T>(dx[i], fx);
      unpack<T>(dg[i], fg);
#pragma unroll
      for (int e = 0; e < VE; ++e) {
        float t = centred(fx[e], pivot[e], mean[e]), dd = t * t;
        float s = Sig<T>::f(fmaf(dd, inv4v[e], 0.5f));
        float a = fg[e] * fx[e] * s * (1.f - s);
        fx[e] = fmaf(fg[e], s, 2.f * t * fmaf(a, inv4v[e], -c1[e])) - c2[e];
      }
      st_stream(op + (int64_t)r * cvec, pack<T>(fx));
    }
  }
  if constexpr (CLUSTER > 1) cluster_sync_all();
}

// ------------------------------------------------------------------------------------------------
// NLC two-sweep kernels: a cluster owns (image b, slab of CWV 16-byte vectors per row) and splits the
// L rows among its CTAs; sweep 1 accumulates the per-channel moments (forward) or R1 / R2'
// (backward) in registers, one slab_reduce (shuffles -> smem -> DSMEM) combines them, sweep 2
// re-reads the rows — served by L2: an image is 2-4 MB and was just read — and writes the result.
// Nothing is held between the sweeps, so any L fits, registers stay low (3 CTAs / SM) and DRAM
// traffic remains 1 read + 1 write.
// ------------------------------------------------------------------------------------------------
template <typename T, int CWV, int CLUSTER>
__global__ void __launch_bounds__(256, 2)
    simam_nlc_fwd_2pass(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ stats, int L,
                        int C, int slabs, int ngroups, float e_lambda) {
  constexpr int VE = Vec16<T>::N, CW = CWV * VE, THREADS = 256, RPI = THREADS / CWV, WARPS = THREADS / 32;
  // 16-byte loads in flight per thread: the sweeps are latency-bound (a CTA streams only ~256 KB),
  // 4 in flight left the kernel at 42 % of HBM bandwidth
  constexpr int UF = 4;  // x 2: double-buffered
  __shared__ float s_red[WARPS * 2 * CW];
  __shared__ float s_part[2 * CW];
  __shared__ float s_fin[2 * CW];
  const int rank = CLUSTER > 1 ? (int)cluster_ctarank() : 0;
  // persistent clusters: the host sizes the grid so that the groups in flight fit in L2
  for (int group = blockIdx.x / CLUSTER; group < ngroups; group += gridDim.x / CLUSTER) {
  const int b = group / slabs, slab = group % slabs, cv = threadIdx.x % CWV;
  const int cvec = C / VE;
  const int rpc = (L + CLUSTER - 1) / CLUSTER;
  const int r_begin = rank * rpc + threadIdx.x / CWV, r_end = min(L, (rank + 1) * rpc);
  const int64_t base = (int64_t)b * L * cvec + slab * CWV + cv;
  const uint4* xp = reinterpret_cast<const uint4*>(x) + base;
  uint4* yp = reinterpret_cast<uint4*>(y) + base;
  float pivot[VE];
  unpack<T>(__ldg(xp), pivot);  // row 0 of the image, this thread's channels
  float acc[2][VE];
#pragma unroll
  for (int e = 0; e < VE; ++e) acc[0][e] = acc[1][e] = 0.f;
  // software-pipelined: the loads of step i+1 are in flight while step i is reduced (all CTAs start
  // together, so without this every warp waits for DRAM and then computes, in lock-step)
  auto load_rows = [&](int r, uint4 (&u)[UF]) {
#pragma unroll
    for (int i = 0; i < UF; ++i)
      u[i] = (r + i * RPI < r_end) ? ld_stream(xp + (int64_t)(r + i * RPI) * cvec) : make_uint4(0, 0, 0, 0);
  };
  uint4 u[UF], un[UF];
  load_rows(r_begin, u);
  for (int r = r_begin; r < r_end; r += UF * RPI) {
    load_rows(r + UF * RPI, un);
#pragma unroll
    for (int i = 0; i < UF; ++i) {
      if (r + i * RPI < r_end) {
        float f[VE];
        unpack<T>(u[i], f);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          const float d = f[e] - pivot[e];
          acc[0][e] += d;
          acc[1][e] = fmaf(d, d, acc[1][e]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < UF; ++i) u[i] = un[i];
  }
  slab_reduce<VE, CWV, THREADS, CLUSTER, 2>(acc, s_red, s_part, s_fin);
  float mean[VE], k[VE];  // full mean; 1/(8v) (bf16) or 1/(4v) (fp32)
  FwdCoef cf[VE];
#pragma unroll
  for (int e = 0; e < VE; ++e) {
    const float dmean = acc[0][e] / (float)L;
    const float var = fmaxf(acc[1][e] - acc[0][e] * dmean, 0.f) / ((float)L - 1.f) + e_lambda;
    mean[e] = pivot[e] + dmean;
    k[e] = 1.f / (8.f * var);
    cf[e] = FwdCoef{pivot[e], dmean, 2.f * k[e]};
    acc[0][e] = dmean;
    acc[1][e] = var;
  }
  if (stats != nullptr && rank == 0 && threadIdx.x < CWV) {
    const int64_t p0 = (int64_t)b * C + slab * CW + cv * VE;
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      stats[2 * (p0 + e)] = acc[0][e];
      stats[2 * (p0 + e) + 1] = acc[1][e];
    }
  }
  load_rows(r_begin, u);
  for (int r = r_begin; r < r_end; r += UF * RPI) {
    load_rows(r + UF * RPI, un);
#pragma unroll
    for (int i = 0; i < UF; ++i) {
      if (r + i * RPI < r_end) {
        float f[VE];
        unpack<T>(u[i], f);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          if constexpr (sizeof(T) == 2) f[e] = simam_fwd_fast(f[e], mean[e], k[e]);
          else f[e] = simam_fwd_elem<T>(f[e], cf[e]);
        }
        st_stream(yp + (int64_t)(r + i * RPI) * cvec, pack<T>(f));
      }
    }
#pragma unroll
    for (int i = 0; i < UF; ++i) u[i] = un[i];
  }
  if constexpr (CLUSTER > 1) cluster_sync_all();  // peers have finished reading s_part
  }
}

template <typename T, int CWV, int CLUSTER>
__global__ void __launch_bounds__(256, 2)
    simam_nlc_bwd_2pass(const T* __restrict__ x, const T* __restrict__ gy,
                        const float* __restrict__ stats, T* __restrict__ gx, int L, int C, int slabs,
                        int ngroups) {
  constexpr int VE = Vec16<T>::N, CW = CWV * VE, THREADS = 256, RPI = THREADS / CWV, WARPS = THREADS / 32;
  constexpr int UB = 2;  // rows per thread per step; double-buffered: 4 x UB 16-byte loads in flight
  __shared__ float s_red[WARPS * 2 * CW];
  __shared__ float s_part[2 * CW];
  __shared__ float s_fin[2 * CW];
  const int rank = CLUSTER > 1 ? (int)cluster_ctarank() : 0;
  for (int group = blockIdx.x / CLUSTER; group < ngroups; group += gridDim.x / CLUSTER) {
  const int b = group / slabs, slab = group % slabs, cv = threadIdx.x % CWV;
  const int cvec = C / VE;
  const int rpc = (L + CLUSTER - 1) / CLUSTER;
  const int r_begin = rank * rpc + threadIdx.x / CWV, r_end = min(L, (rank + 1) * rpc);
  const int64_t base = (int64_t)b * L * cvec + slab * CWV + cv;
  const uint4* xp = reinterpret_cast<const uint4*>(x) + base;
  const uint4* gp = reinterpret_cast<const uint4*>(gy) + base;
  uint4* op = reinterpret_cast<uint4*>(gx) + base;
  float pivot[VE], mean[VE], inv4v[VE], vv[VE];  // mean[]: full mean (bf16) or mean - pivot (fp32)
  unpack<T>(__ldg(xp), pivot);
  {
    const int64_t p0 = (int64_t)b * C + slab * CW + cv * VE;
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      const float dm = __ldg(stats + 2 * (p0 + e));
      mean[e] = sizeof(T) == 2 ? pivot[e] + dm : dm;
      vv[e] = __ldg(stats + 2 * (p0 + e) + 1);
      inv4v[e] = 1.f / (4.f * vv[e]);
    }
  }
  // a4 = 4 a = g x (1 - tanh^2) (bf16) or 4 g x s (1 - s) (fp32); t is formed against (pivot, dmean)
  auto elem = [&](float xe, float ge, int e, float& t, float& dd, float& a4, float& gs) {
    if constexpr (sizeof(T) == 2) {
      t = xe - mean[e];
      dd = t * t;
      float th;
      asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(fmaf(dd, 0.5f * inv4v[e], 0.25f)));
      a4 = ge * xe * fmaf(-th, th, 1.f);
      const float hg = 0.5f * ge;
      gs = fmaf(hg, th, hg);
    } else {
      t = (xe - pivot[e]) - mean[e];
      dd = t * t;
      const float sg = Sig<T>::f(fmaf(dd, inv4v[e], 0.5f));
      a4 = 4.f * ge * xe * sg * (1.f - sg);
      gs = ge * sg;
    }
  };
  float acc[2][VE];
#pragma unroll
  for (int e = 0; e < VE; ++e) acc[0][e] = acc[1][e] = 0.f;
  auto load_rows = [&](int r, uint4 (&ux)[UB], uint4 (&ug)[UB]) {
#pragma unroll
    for (int i = 0; i < UB; ++i) {
      const bool ok = r + i * RPI < r_end;
      ux[i] = ok ? ld_stream(xp + (int64_t)(r + i * RPI) * cvec) : make_uint4(0, 0, 0, 0);
      ug[i] = ok ? ld_stream(gp + (int64_t)(r + i * RPI) * cvec) : make_uint4(0, 0, 0, 0);
    }
  };
  uint4 ux[UB], ug[UB], nx[UB], ng[UB];
  load_rows(r_begin, ux, ug);
  for (int r = r_begin; r < r_end; r += UB * RPI) {
    load_rows(r + UB * RPI, nx, ng);  // next step's rows in flight while this one is reduced
#pragma unroll
    for (int i = 0; i < UB; ++i) {
      if (r + i * RPI < r_end) {
        float fx[VE], fg[VE];
        unpack<T>(ux[i], fx);
        unpack<T>(ug[i], fg);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          float t, dd, a4, gs;
          elem(fx[e], fg[e], e, t, dd, a4, gs);
          acc[0][e] = fmaf(a4, dd, acc[0][e]);
          acc[1][e] = fmaf(a4, t, acc[1][e]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < UB; ++i) {
      ux[i] = nx[i];
      ug[i] = ng[i];
    }
  }
  slab_reduce<VE, CWV, THREADS, CLUSTER, 2>(acc, s_red, s_part, s_fin);
  float k1[VE], k2[VE], c2[VE];
#pragma unroll
  for (int e = 0; e < VE; ++e) {
    const float r1 = 0.25f * acc[0][e], r2 = 0.25f * acc[1][e];
    const float c1 = r1 * inv4v[e] / (vv[e] * ((float)L - 1.f));
    c2[e] = 2.f / (float)L * r2 * inv4v[e];
    k1[e] = 0.5f * inv4v[e];  // 2 (a inv4v - c1) = a4 k1 - k2
    k2[e] = 2.f * c1;
  }
  load_rows(r_begin, ux, ug);
  for (int r = r_begin; r < r_end; r += UB * RPI) {
    load_rows(r + UB * RPI, nx, ng);
#pragma unroll
    for (int i = 0; i < UB; ++i) {
      if (r + i * RPI < r_end) {
        float fx[VE], fg[VE];
        unpack<T>(ux[i], fx);
        unpack<T>(ug[i], fg);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          float t, dd, a4, gs;
          elem(fx[e], fg[e], e, t, dd, a4, gs);
          fx[e] = fmaf(t, fmaf(a4, k1[e], -k2[e]), gs) - c2[e];
        }
        st_stream(op + (int64_t)(r + i * RPI) * cvec, pack<T>(fx));
      }
    }
#pragma unroll
    for (int i = 0; i < UB; ++i) {
      ux[i] = nx[i];
      ug[i] = ng[i];
    }
  }
  if constexpr (CLUSTER > 1) cluster_sync_all();
  }
}

// ------------------------------------------------------------------------------------------------
// generic (any shape, any alignment): element (plane p, position i) at  base(p) + i * istride.
//   NCHW: base = p*S, istride = 1, 256 threads per plane, one plane per CTA.
//   NLC : blockDim = (32 channels, 8 row groups); a CTA covers 32 adjacent channels of one image.
// ------------------------------------------------------------------------------------------------
template <typename T, bool BWD>
__global__ void __launch_bounds__(256)
    simam_generic(const T* __restrict__ x, const T* __restrict__ gy, float* __restrict__ stats_out,
                  const float* __restrict__ stats_in, T* __restrict__ out, int64_t S, int64_t C,
                  int layout, float e_lambda) {
  __shared__ float s_red[2][8][33];
  // lane = which plane inside the CTA (NLC) / always 0 (NCHW); grp = which row group
  int lane, grp, ngrp;
  int64_t plane, base, istride;
  bool active = true;
  if (layout == CSB200_NCHW) {
    lane = 0;
    grp = threadIdx.x;
    ngrp = 256;
    plane = blockIdx.x;
    base = plane * S;
    istride = 1;
  } else {
    lane = threadIdx.x & 31;
    grp = threadIdx.x >> 5;
    ngrp = 8;
    const int64_t cblocks = (C + 31) / 32;
    const int64_t b = blockIdx.x / cblocks, c = (blockIdx.x % cblocks) * 32 + lane;
    active = c < C;
    plane = b * C + (active ? c : 0);
    base = b * S * C + (active ? c : 0);
    istride = C;
  }
  auto reduce2 = [&](float& a, float& b2) {
    if (layout == CSB200_NCHW) {
      a = warp_sum(a);
      b2 = warp_sum(b2);
      if ((threadIdx.x & 31) == 0) {
        s_red[0][threadIdx.x >> 5][0] = a;
        s_red[1][threadIdx.x >> 5][0] = b2;
      }
      __syncthreads();
      a = b2 = 0.f;
      for (int w = 0; w < 8; ++w) {
        a += s_red[0][w][0];
        b2 += s_red[1][w][0];
      }
    } else {
      s_red[0][grp][lane] = a;
      s_red[1][grp][lane] = b2;
      __syncthreads();
      a = b2 = 0.f;
      for (int w = 0; w < 8; ++w) {
        a += s_red[0][w][lane];
        b2 += s_red[1][w][lane];
      }
    }
    __syncthreads();
  };
  const float Sf = (float)S;
  const float pivot = to_f32(x[base]);
  float mean, v;  // mean holds mean - pivot
  if constexpr (!BWD) {
    float sum = 0.f, dummy = 0.f;
    if (active)
      for (int64_t i = grp; i < S; i += ngrp) sum += to_f32(x[base + i * istride]) - pivot;
    reduce2(sum, dummy);
    mean = sum / Sf;
    float m2 = 0.f;
    if (active)
      for (int64_t i = grp; i < S; i += ngrp) {
        float t = centred(to_f32(x[base + i * istride]), pivot, mean);
        m2 = fmaf(t, t, m2);
      }
    reduce2(m2, dummy);
    v = m2 / (Sf - 1.f) + e_lambda;
    FwdCoef c{pivot, mean, 1.f / (4.f * v)};
    if (active) {
      for (int64_t i = grp; i < S; i += ngrp)
        out[base + i * istride] = from_f32<T>(simam_fwd_elem<T>(to_f32(x[base + i * istride]), c));
      if (stats_out != nullptr && grp == 0) {
        stats_out[2 * plane] = mean;
        stats_out[2 * plane + 1] = v;
      }
    }
  } else {
    mean = stats_in[2 * plane];
    v = stats_in[2 * plane + 1];
    const float inv4v = 1.f / (4.f * v);
    float r1 = 0.f, r2 = 0.f;
    if (active)
      for (int64_t i = grp; i < S; i += ngrp) {
        float xv = to_f32(x[base + i * istride]), g = to_f32(gy[base + i * istride]);
        float t = centred(xv, pivot, mean), dd = t * t;
        float s = Sig<T>::f(fmaf(dd, inv4v, 0.5f));
        float a = g * xv * s * (1.f - s);
        r1 = fmaf(a, dd, r1);
        r2 = fmaf(a, t, r2);
      }
    reduce2(r1, r2);
    const float c1 = r1 * inv4v / (v * (Sf - 1.f)), c2 = 2.f / Sf * r2 * inv4v;
    if (active)
      for (int64_t i = grp; i < S; i += ngrp) {
        float xv = to_f32(x[base + i * istride]), g = to_f32(gy[base + i * istride]);
        float t = centred(xv, pivot, mean), dd = t * t;
        float s = Sig<T>::f(fmaf(dd, inv4v, 0.5f));
        float a = g * xv * s * (1.f - s);
        out[base + i * istride] = from_f32<T>(fmaf(g, s, 2.f * t * fmaf(a, inv4v, -c1)) - c2);
      }
  }
}

// ------------------------------------------------------------------------------------------------
// NCHW streaming forward for planes that are whole multiples of 16 KB (64 KB .. any size): one
// persistent CTA per SM, a producer warp feeds a 12-stage shared-memory ring with 16-KB bulk copies
// (cp.async.bulk + mbarrier: loads never wait for arithmetic), 8 consumer warps sweep each plane
// twice — moments, then rescale + store.  The second sweep re-reads the plane through the ring; with
// <= 148 planes of <= 256 KB in flight it is served by the 126-MB L2, so DRAM traffic stays
// 1 read + 1 write.  No clusters, no per-plane launch phases, registers hold only the current vector.
// ------------------------------------------------------------------------------------------------
constexpr int ST_CHUNK = 16384, ST_STAGES = 12, ST_CONSUMERS = 512;  // 16 consumer warps + 1 producer

__device__ __forceinline__ uint32_t st_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void st_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tSW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra SD_%=;\n\tbra SW_%=;\n\tSD_%=:\n\t}" ::"r"(st_smem_u32(bar)),
      "r"(parity)
      : "memory");
}

template <typename T>
__global__ void __launch_bounds__(ST_CONSUMERS + 32, 1)
    simam_nchw_fwd_stream(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ stats,
                          int planes, int chunks, float S, float e_lambda) {
  constexpr int VE = Vec16<T>::N;
  extern __shared__ __align__(128) uint8_t st_smem[];
  uint8_t* ring = st_smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(st_smem + ST_STAGES * ST_CHUNK);
  uint64_t* empty = full + ST_STAGES;
  float* s_red = reinterpret_cast<float*>(empty + ST_STAGES);  // [warps][2]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ST_STAGES; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(&full[i])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(&empty[i])), "r"(ST_CONSUMERS / 32));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int my_planes = (planes - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int64_t plane_bytes = (int64_t)chunks * ST_CHUNK;

  if (warp == ST_CONSUMERS / 32) {
    // ------------------------------- producer -------------------------------
    if (lane == 0) {
      int it = 0;
      for (int pi = 0; pi < my_planes; ++pi) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(x) +
                             ((int64_t)blockIdx.x + (int64_t)pi * gridDim.x) * plane_bytes;
        for (int pass = 0; pass < 2; ++pass)
          for (int c = 0; c < chunks; ++c, ++it) {
            const int s = it % ST_STAGES;
            st_mbar_wait(&empty[s], ((it / ST_STAGES) & 1) ^ 1);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                             st_smem_u32(&full[s])),
                         "r"(ST_CHUNK)
                         : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                    "r"(st_smem_u32(ring + s * ST_CHUNK)),
                "l"(src + (int64_t)c * ST_CHUNK), "r"(ST_CHUNK), "r"(st_smem_u32(&full[s]))
                : "memory");
          }
      }
    }
    return;
  }
  // --------------------------------- consumers ---------------------------------
  int it = 0;
  for (int pi = 0; pi < my_planes; ++pi) {
    const int64_t plane = (int64_t)blockIdx.x + (int64_t)pi * gridDim.x;
    float pivot = 0.f, sum = 0.f, sq = 0.f;
    for (int c = 0; c < chunks; ++c, ++it) {
      const int s = it % ST_STAGES;
      st_mbar_wait(&full[s], (it / ST_STAGES) & 1);
      const uint4* v = reinterpret_cast<const uint4*>(ring + s * ST_CHUNK);
      if (c == 0) pivot = to_f32(*reinterpret_cast<const T*>(v));  // plane's first element (broadcast)
#pragma unroll
      for (int i = 0; i < ST_CHUNK / 16 / ST_CONSUMERS; ++i) {
        float f[VE];
        unpack<T>(v[threadIdx.x + i * ST_CONSUMERS], f);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          const float d = f[e] - pivot;
          sum += d;
          sq = fmaf(d, d, sq);
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(st_smem_u32(&empty[s])) : "memory");
    }
    sum = warp_sum(sum);
    sq = warp_sum(sq);
    if (lane == 0) {
      s_red[warp * 2] = sum;
      s_red[warp * 2 + 1] = sq;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(ST_CONSUMERS) : "memory");
    sum = sq = 0.f;
#pragma unroll
    for (int w = 0; w < ST_CONSUMERS / 32; ++w) {
      sum += s_red[w * 2];
      sq += s_red[w * 2 + 1];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(ST_CONSUMERS) : "memory");  // s_red is reused by the next plane
    const float dmean = sum / S;
    const float var = fmaxf(sq - sum * dmean, 0.f) / (S - 1.f) + e_lambda;
    if (threadIdx.x == 0 && stats != nullptr) {
      stats[2 * plane] = dmean;
      stats[2 * plane + 1] = var;
    }
    const float mean = pivot + dmean;
    FwdCoef cf{pivot, dmean, 1.f / (4.f * var)};
    const float inv8v = 1.f / (8.f * var);
    uint4* dst = reinterpret_cast<uint4*>(y) + plane * (plane_bytes / 16);
    for (int c = 0; c < chunks; ++c, ++it) {
      const int s = it % ST_STAGES;
      st_mbar_wait(&full[s], (it / ST_STAGES) & 1);
      const uint4* v = reinterpret_cast<const uint4*>(ring + s * ST_CHUNK);
#pragma unroll
      for (int i = 0; i < ST_CHUNK / 16 / ST_CONSUMERS; ++i) {
        float f[VE];
        unpack<T>(v[threadIdx.x + i * ST_CONSUMERS], f);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          if constexpr (sizeof(T) == 2) f[e] = simam_fwd_fast(f[e], mean, inv8v);
          else f[e] = simam_fwd_elem<T>(f[e], cf);
        }
        st_stream(dst + (int64_t)c * (ST_CHUNK / 16) + threadIdx.x + i * ST_CONSUMERS, pack<T>(f));
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(st_smem_u32(&empty[s])) : "memory");
    }
  }
}

// Streaming backward: a ring stage holds an 8-KB chunk of x and the matching 8-KB chunk of grad_y.
// Sweep 1 accumulates R1 = sum a d and R2' = sum a t, sweep 2 (served by L2) writes grad_x.
// bf16 uses sigmoid = 0.5 tanh + 0.5, hence s (1 - s) = (1 - tanh^2) / 4, and folds the constants.
template <typename T>
__global__ void __launch_bounds__(ST_CONSUMERS + 32, 1)
    simam_nchw_bwd_stream(const T* __restrict__ x, const T* __restrict__ gy,
                          const float* __restrict__ stats, T* __restrict__ gx, int planes, int chunks,
                          float S) {
  constexpr int VE = Vec16<T>::N, HALF = ST_CHUNK / 2;
  extern __shared__ __align__(128) uint8_t st_smem[];
  uint8_t* ring = st_smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(st_smem + ST_STAGES * ST_CHUNK);
  uint64_t* empty = full + ST_STAGES;
  float* s_red = reinterpret_cast<float*>(empty + ST_STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ST_STAGES; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(&full[i])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(&empty[i])), "r"(ST_CONSUMERS / 32));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int my_planes = (planes - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int64_t plane_bytes = (int64_t)chunks * HALF;

  if (warp == ST_CONSUMERS / 32) {
    if (lane == 0) {
      int it = 0;
      for (int pi = 0; pi < my_planes; ++pi) {
        const int64_t off = ((int64_t)blockIdx.x + (int64_t)pi * gridDim.x) * plane_bytes;
        const uint8_t* sx = reinterpret_cast<const uint8_t*>(x) + off;
        const uint8_t* sg = reinterpret_cast<const uint8_t*>(gy) + off;
        for (int pass = 0; pass < 2; ++pass)
          for (int c = 0; c < chunks; ++c, ++it) {
            const int s = it % ST_STAGES;
            st_mbar_wait(&empty[s], ((it / ST_STAGES) & 1) ^ 1);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                             st_smem_u32(&full[s])),
                         "r"(ST_CHUNK)
                         : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                    "r"(st_smem_u32(ring + s * ST_CHUNK)),
                "l"(sx + (int64_t)c * HALF), "r"(HALF), "r"(st_smem_u32(&full[s]))
                : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                    "r"(st_smem_u32(ring + s * ST_CHUNK + HALF)),
                "l"(sg + (int64_t)c * HALF), "r"(HALF), "r"(st_smem_u32(&full[s]))
                : "memory");
          }
      }
    }
    return;
  }
  int it = 0;
  for (int pi = 0; pi < my_planes; ++pi) {
    const int64_t plane = (int64_t)blockIdx.x + (int64_t)pi * gridDim.x;
    const float dmean = __ldg(stats + 2 * plane), v = __ldg(stats + 2 * plane + 1);
    const float inv4v = 1.f / (4.f * v), inv8v = 0.5f * inv4v;
    float pivot = 0.f, mean = 0.f, r1 = 0.f, r2 = 0.f;
    f2_t r1p = f2_splat(0.f), r2p = f2_splat(0.f), nmean2 = f2_splat(0.f);
    const f2_t inv8v2 = f2_splat(inv8v), quarter2 = f2_splat(0.25f), mone2 = f2_splat(-1.f);
    for (int c = 0; c < chunks; ++c, ++it) {
      const int s = it % ST_STAGES;
      st_mbar_wait(&full[s], (it / ST_STAGES) & 1);
      const uint4* vx = reinterpret_cast<const uint4*>(ring + s * ST_CHUNK);
      const uint4* vg = reinterpret_cast<const uint4*>(ring + s * ST_CHUNK + HALF);
      if (c == 0) {
        pivot = to_f32(*reinterpret_cast<const T*>(vx));
        mean = pivot + dmean;
        nmean2 = f2_splat(-mean);
      }
#pragma unroll
      for (int i = 0; i < HALF / 16 / ST_CONSUMERS; ++i) {
        if constexpr (sizeof(T) == 2) {
          // pairs: -4a = g x (tanh^2 - 1); the sign is undone when the sums are finalised
          const uint4 ux = vx[threadIdx.x + i * ST_CONSUMERS], ug = vg[threadIdx.x + i * ST_CONSUMERS];
          const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w}, wg[4] = {ug.x, ug.y, ug.z, ug.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const f2_t x2 = f2_from_bf16x2(wx[q]), g2 = f2_from_bf16x2(wg[q]);
            const f2_t t = f2_add(x2, nmean2), dd = f2_mul(t, t);
            const f2_t th = f2_tanh(f2_fma(dd, inv8v2, quarter2));
            const f2_t na4 = f2_mul(f2_mul(g2, x2), f2_fma(th, th, mone2));
            r1p = f2_fma(na4, dd, r1p);
            r2p = f2_fma(na4, t, r2p);
          }
          continue;
        }
        float fx[VE], fg[VE];
        unpack<T>(vx[threadIdx.x + i * ST_CONSUMERS], fx);
        unpack<T>(vg[threadIdx.x + i * ST_CONSUMERS], fg);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          if constexpr (sizeof(T) == 2) {
            const float t = fx[e] - mean, dd = t * t;
            float th;
            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(fmaf(dd, inv8v, 0.25f)));
            const float a4 = fg[e] * fx[e] * fmaf(-th, th, 1.f);  // 4 a
            r1 = fmaf(a4, dd, r1);
            r2 = fmaf(a4, t, r2);
          } else {
            const float t = centred(fx[e], pivot, dmean), dd = t * t;
            const float sg_ = Sig<T>::f(fmaf(dd, inv4v, 0.5f));
            const float a4 = 4.f * fg[e] * fx[e] * sg_ * (1.f - sg_);
            r1 = fmaf(a4, dd, r1);
            r2 = fmaf(a4, t, r2);
          }
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(st_smem_u32(&empty[s])) : "memory");
    }
    if constexpr (sizeof(T) == 2) {
      float a, b;
      f2_split(r1p, a, b);
      r1 = -(a + b);  // the pairs accumulate -4a
      f2_split(r2p, a, b);
      r2 = -(a + b);
    }
    r1 = warp_sum(r1);
    r2 = warp_sum(r2);
    if (lane == 0) {
      s_red[warp * 2] = r1;
      s_red[warp * 2 + 1] = r2;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(ST_CONSUMERS) : "memory");
    r1 = r2 = 0.f;
#pragma unroll
    for (int w = 0; w < ST_CONSUMERS / 32; ++w) {
      r1 += s_red[w * 2];
      r2 += s_red[w * 2 + 1];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(ST_CONSUMERS) : "memory");
    r1 *= 0.25f;  // the sweeps accumulate 4 a
    r2 *= 0.25f;
    const float c1 = r1 * inv4v / (v * (S - 1.f));  // R1 / (4 v^2 n)
    const float c2 = 2.f / S * r2 * inv4v;          // (2/HW) R2
    const float k1 = 0.5f * inv4v, k2 = 2.f * c1;   // 2 (a inv4v - c1) = a4 k1 - k2
    const f2_t half2 = f2_splat(0.5f), nc2_2 = f2_splat(-c2), nk1_2 = f2_splat(-k1), nk2_2 = f2_splat(-k2);
    uint4* dst = reinterpret_cast<uint4*>(gx) + plane * (plane_bytes / 16);
    for (int c = 0; c < chunks; ++c, ++it) {
      const int s = it % ST_STAGES;
      st_mbar_wait(&full[s], (it / ST_STAGES) & 1);
      const uint4* vx = reinterpret_cast<const uint4*>(ring + s * ST_CHUNK);
      const uint4* vg = reinterpret_cast<const uint4*>(ring + s * ST_CHUNK + HALF);
#pragma unroll
      for (int i = 0; i < HALF / 16 / ST_CONSUMERS; ++i) {
        if constexpr (sizeof(T) == 2) {
          // grad_x = 0.5 g (1 + tanh) + t (4a k1 - k2) - c2, on pairs; -4a as in sweep 1, so k1 enters negated
          const uint4 ux = vx[threadIdx.x + i * ST_CONSUMERS], ug = vg[threadIdx.x + i * ST_CONSUMERS];
          const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w}, wg[4] = {ug.x, ug.y, ug.z, ug.w};
          uint32_t o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const f2_t x2 = f2_from_bf16x2(wx[q]), g2 = f2_from_bf16x2(wg[q]);
            const f2_t t = f2_add(x2, nmean2), dd = f2_mul(t, t);
            const f2_t th = f2_tanh(f2_fma(dd, inv8v2, quarter2));
            const f2_t na4 = f2_mul(f2_mul(g2, x2), f2_fma(th, th, mone2));
            const f2_t hg = f2_mul(g2, half2);
            const f2_t gs = f2_fma(hg, th, f2_add(hg, nc2_2));              // 0.5 g (1 + tanh) - c2
            const f2_t r = f2_fma(t, f2_fma(na4, nk1_2, nk2_2), gs);
            float lo, hi;
            f2_split(r, lo, hi);
            o[q] = pack_bf16x2(lo, hi);
          }
          st_stream(dst + (int64_t)c * (HALF / 16) + threadIdx.x + i * ST_CONSUMERS, make_uint4(o[0], o[1], o[2], o[3]));
          continue;
        }
        float fx[VE], fg[VE];
        unpack<T>(vx[threadIdx.x + i * ST_CONSUMERS], fx);
        unpack<T>(vg[threadIdx.x + i * ST_CONSUMERS], fg);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          if constexpr (sizeof(T) == 2) {
            const float t = fx[e] - mean, dd = t * t;
            float th;
            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(fmaf(dd, inv8v, 0.25f)));
            const float a4 = fg[e] * fx[e] * fmaf(-th, th, 1.f);
            const float hg = 0.5f * fg[e];
            fx[e] = fmaf(t, fmaf(a4, k1, -k2), fmaf(hg, th, hg)) - c2;
          } else {
            const float t = centred(fx[e], pivot, dmean), dd = t * t;
            const float sg_ = Sig<T>::f(fmaf(dd, inv4v, 0.5f));
            const float a = fg[e] * fx[e] * sg_ * (1.f - sg_);
            fx[e] = fmaf(fg[e], sg_, 2.f * t * fmaf(a, inv4v, -c1)) - c2;
          }
        }
        st_stream(dst + (int64_t)c * (HALF / 16) + threadIdx.x + i * ST_CONSUMERS, pack<T>(fx));
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(st_smem_u32(&empty[s])) : "memory");
    }
  }
}

template <typename T>
int nchw_bwd_stream(const T* x, const T* gy, const float* stats, T* gx, int64_t planes, int64_t S,
                    cudaStream_t st) {
  const int64_t plane_bytes = S * (int64_t)sizeof(T);
  if (plane_bytes % (ST_CHUNK / 2) != 0 || plane_bytes < 4 * (ST_CHUNK / 2) || planes > 0x7fffffff) return -1;
  if (!aligned16(x) || !aligned16(gy) || !aligned16(gx)) return -1;
  const int sms = device_sm_count();
  if (sms <= 0) return -1;
  const int smem = ST_STAGES * ST_CHUNK + 2 * ST_STAGES * 8 + 2 * (ST_CONSUMERS / 32) * 4 + 64;
  if (opt_in_smem(reinterpret_cast<const void*>(&simam_nchw_bwd_stream<T>), smem) != cudaSuccess) return -1;
  const int grid = planes < sms ? (int)planes : sms;
  simam_nchw_bwd_stream<T><<<grid, ST_CONSUMERS + 32, smem, st>>>(
      x, gy, stats, gx, (int)planes, (int)(plane_bytes / (ST_CHUNK / 2)), (float)S);
  return check_launch("simam_nchw_bwd_stream");
}

template <typename T>
int nchw_fwd_stream(const T* x, T* y, float* stats, int64_t planes, int64_t S, float e_lambda,
                    cudaStream_t st) {
  const int64_t plane_bytes = S * (int64_t)sizeof(T);
  if (plane_bytes % ST_CHUNK != 0 || plane_bytes < 2 * ST_CHUNK || planes > 0x7fffffff) return -1;
  if (!aligned16(x) || !aligned16(y)) return -1;
  const int sms = device_sm_count();
  if (sms <= 0) return -1;
  const int smem = ST_STAGES * ST_CHUNK + 2 * ST_STAGES * 8 + 2 * (ST_CONSUMERS / 32) * 4 + 64;
  if (opt_in_smem(reinterpret_cast<const void*>(&simam_nchw_fwd_stream<T>), smem) != cudaSuccess) return -1;
  const int grid = planes < sms ? (int)planes : sms;
  simam_nchw_fwd_stream<T><<<grid, ST_CONSUMERS + 32, smem, st>>>(x, y, stats, (int)planes,
                                                                   (int)(plane_bytes / ST_CHUNK), (float)S, e_lambda);
  return check_launch("simam_nchw_fwd_stream");
}

// ------------------------------------------------------------------------------------------------
// NLC streaming kernels (token layout, the CSWin skips of config 3): a cluster of 1/2/4/8 CTAs owns
// one image — L*C contiguous elements — and each CTA streams its contiguous share of the rows
// through the same bulk-copy ring as the NCHW kernels (whole 128..1024-byte rows, 16 KB per copy,
// instead of 32..128-byte row segments fetched by the threads themselves).  512 % (16-byte vectors
// per row) == 0, so a consumer thread meets the same 8 (bf16) / 4 (fp32) channels in every vector it
// reads and keeps their moments in registers.  The per-channel partials of the CTAs meet once per
// image: every CTA PUSHES its partials into the shared memory of all peers (st.shared::cluster) and
// then arrives on their mbarriers, so a CTA only ever reads its own shared memory and needs no
// hand-shake before it exits.  Only consumer threads take part; the producer thread keeps
// prefetching.  A cluster that owns several images software-pipelines them: the second sweep of
// image i (reads served by L2, writes to DRAM) is interleaved chunk by chunk with the first sweep of
// image i+1 (reads from DRAM).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t nl_mapa(const void* local_smem_ptr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(st_smem_u32(local_smem_ptr)), "r"(rank));
  return remote;
}
__device__ __forceinline__ void nl_remote_arrive(uint64_t* bar, uint32_t rank) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(nl_mapa(bar, rank)) : "memory");
}
__device__ __forceinline__ void nl_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tXW_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra XD_%=;\n\tbra XW_%=;\n\tXD_%=:\n\t}" ::"r"(st_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void nl_consumer_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(ST_CONSUMERS) : "memory");
}
// streaming store whose line is the first candidate for eviction from L2 (outputs are never re-read)
__device__ __forceinline__ void st_stream_first(void* p, const uint4& v, uint64_t policy) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy)
               : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

template <int VE, int CVEC>
struct NlSmem {
  static constexpr int GW = CVEC < 32 ? CVEC : 32;        // column vectors a warp covers
  static constexpr int NCLS = CVEC < 32 ? 1 : CVEC / 32;  // warp w covers vectors (w % NCLS) * 32 + lane
  static constexpr int CW = CVEC * VE;                    // channels
  static constexpr int WARPS = ST_CONSUMERS / 32;
  static constexpr int RED_BYTES = WARPS * 2 * VE * GW * 4;  // [WARPS][2][VE][GW] floats
  static constexpr int PART_BYTES = 2 * 8 * 2 * CW * 4;      // [image parity][source rank][2 * CW] floats
  static constexpr int FIN_BYTES = 2 * CW * 4;
  static constexpr int BAR_BYTES = 256;                      // full[STAGES], empty[STAGES], xbar
  static constexpr int MAX_SMEM = 227 * 1024;
  static constexpr int FIT = (MAX_SMEM - BAR_BYTES - RED_BYTES - PART_BYTES - FIN_BYTES) / ST_CHUNK;
  static constexpr int STAGES = FIT < 10 ? FIT : 10;
  static constexpr int BARS = STAGES * ST_CHUNK;
  static constexpr int RED = BARS + BAR_BYTES;
  static constexpr int PART = RED + RED_BYTES;
  static constexpr int FIN = PART + PART_BYTES;
  static constexpr int BYTES = FIN + FIN_BYTES;
  static_assert(STAGES >= 4 && 2 * STAGES + 1 <= BAR_BYTES / 8, "ring too shallow");
};

// Sum acc[q][e] over all consumer threads of the cluster that hold the same channel; fixed order,
// identical in every CTA of the cluster.
template <int VE, int CVEC>
__device__ __forceinline__ void nl_reduce(float (&acc)[2][VE], uint8_t* smem, int image_parity, int CL, int rank) {
  using S = NlSmem<VE, CVEC>;
  float* s_red = reinterpret_cast<float*>(smem + S::RED);
  float* s_part = reinterpret_cast<float*>(smem + S::PART) + image_parity * (8 * 2 * S::CW);
  float* s_fin = reinterpret_cast<float*>(smem + S::FIN);
  uint64_t* xbar = reinterpret_cast<uint64_t*>(smem + S::BARS) + 2 * S::STAGES;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int e = 0; e < VE; ++e) {
#pragma unroll
      for (int o = 16; o >= CVEC; o >>= 1) acc[q][e] += __shfl_xor_sync(0xffffffffu, acc[q][e], o);
    }
  if (lane < S::GW) {
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int e = 0; e < VE; ++e) s_red[((warp * 2 + q) * VE + e) * S::GW + lane] = acc[q][e];
  }
  nl_consumer_sync();
  for (int j = threadIdx.x; j < 2 * S::CW; j += ST_CONSUMERS) {
    const int q = j / S::CW, ch = j % S::CW, g = ch / VE, e = ch % VE;
    const int cls = g / S::GW, gl = g % S::GW;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < S::WARPS / S::NCLS; ++w)
      a += s_red[(((cls + w * S::NCLS) * 2 + q) * VE + e) * S::GW + gl];
    if (CL > 1) {
      for (int r = 0; r < CL; ++r)
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(nl_mapa(&s_part[rank * 2 * S::CW + j], (uint32_t)r)), "f"(a)
                     : "memory");
    } else {
      s_fin[j] = a;
    }
  }
  if (CL > 1) {
    asm volatile("fence.acq_rel.cluster;" ::: "memory");
    nl_consumer_sync();  // every partial of this CTA has been pushed
    if (threadIdx.x == 0)
      for (int r = 0; r < CL; ++r) nl_remote_arrive(xbar, (uint32_t)r);
    nl_wait_cluster(xbar, (uint32_t)image_parity);  // ... and every peer's has landed here
    for (int j = threadIdx.x; j < 2 * S::CW; j += ST_CONSUMERS) {
      float a = 0.f;
      for (int r = 0; r < CL; ++r) a += s_part[r * 2 * S::CW + j];  // same order in every CTA
      s_fin[j] = a;
    }
  }
  nl_consumer_sync();
  const int cv = threadIdx.x % CVEC;
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int e = 0; e < VE; ++e) acc[q][e] = s_fin[q * S::CW + cv * VE + e];
}

// barriers + the cluster-wide start line; returns after every CTA of the cluster has initialised
template <int VE, int CVEC>
__device__ __forceinline__ void nl_init(uint8_t* smem, int CL) {
  using S = NlSmem<VE, CVEC>;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BARS);
  uint64_t* empty = full + S::STAGES;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S::STAGES; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(&full[i])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(&empty[i])), "r"(ST_CONSUMERS / 32));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(&empty[S::STAGES])), "r"(CL));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // every thread is still here: nobody touches a peer that has not initialised
}

// Producer: one thread streams this CTA's share of the cluster's images.  Phase p interleaves, chunk by
// chunk, the second sweep of image p-1 (an L2 hit, marked evict-first: it is dead afterwards) with
// the first sweep of image p; the consumers walk the ring in the same order.  The second sweep walks
// its chunks BACKWARDS: when the batch does not fit in L2 (backward at 512^2: x and grad_y are
// 128 MB) a forward re-walk of an LRU cache misses every time, the reverse walk hits whatever the
// cache still holds.
template <int STAGES>
__device__ __forceinline__ void nl_produce(uint8_t* smem, int bars_off, const uint8_t* x, const uint8_t* g,
                                           int my_images, int first_image, int image_step, int64_t image_bytes,
                                           int64_t cta_off, int chunks, int part_bytes) {
  uint8_t* ring = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + bars_off);
  uint64_t* empty = full + STAGES;
  const uint64_t pol_first = l2_evict_first_policy();
  int ring_s = 0;
  uint32_t ring_par = 0;
  auto issue = [&](int64_t off, bool again) {
    const int s = ring_s;
    st_mbar_wait(&empty[s], ring_par ^ 1u);
    if (++ring_s == STAGES) {
      ring_s = 0;
      ring_par ^= 1u;
    }
    const uint32_t bar = st_smem_u32(&full[s]), dst = st_smem_u32(ring + s * ST_CHUNK);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(ST_CHUNK) : "memory");
    if (again) {
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
              "r"(dst), "l"(x + off), "r"(part_bytes), "r"(bar), "l"(pol_first) : "memory");
      if (g != nullptr)
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
                "r"(dst + part_bytes), "l"(g + off), "r"(part_bytes), "r"(bar), "l"(pol_first) : "memory");
    } else {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "l"(x + off), "r"(part_bytes), "r"(bar)
                   : "memory");
      if (g != nullptr)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         dst + part_bytes),
                     "l"(g + off), "r"(part_bytes), "r"(bar)
                     : "memory");
    }
  };
  for (int ph = 0; ph <= my_images; ++ph) {
    const int64_t off2 = ((int64_t)first_image + (int64_t)(ph - 1) * image_step) * image_bytes + cta_off;
    const int64_t off1 = off2 + (int64_t)image_step * image_bytes;
    for (int c = 0; c < chunks; ++c) {
      if (ph > 0) issue(off2 + (int64_t)(chunks - 1 - c) * part_bytes, true);
      if (ph < my_images) issue(off1 + (int64_t)c * part_bytes, false);
    }
  }
}

template <typename T, int CVEC, bool PIPE>
__global__ void __launch_bounds__(ST_CONSUMERS + 32, 1)
    simam_nlc_fwd_stream(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ stats, int B, int L,
                         int chunks /* 16-KB chunks per CTA per image */, float e_lambda) {
  constexpr int VE = Vec16<T>::N, NP = VE / 2, VPC = ST_CHUNK / 16 / ST_CONSUMERS;
  constexpr bool BF = sizeof(T) == 2;
  using S = NlSmem<VE, CVEC>;
  constexpr int STAGES = S::STAGES;
  extern __shared__ __align__(128) uint8_t st_smem[];
  uint8_t* ring = st_smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(st_smem + S::BARS);
  uint64_t* empty = full + STAGES;
  const int CL = (int)cluster_nctarank(), rank = (int)cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  nl_init<VE, CVEC>(st_smem, CL);
  const int cid = (int)blockIdx.x / CL, ncl = (int)gridDim.x / CL;
  const int my_images = (B - cid + ncl - 1) / ncl;
  const int64_t cta_bytes = (int64_t)chunks * ST_CHUNK, image_bytes = cta_bytes * CL;

  if (warp == ST_CONSUMERS / 32) {
    if (lane == 0)
      nl_produce<STAGES>(st_smem, S::BARS, reinterpret_cast<const uint8_t*>(x), nullptr, my_images, cid, ncl,
                         image_bytes, rank * cta_bytes, chunks, ST_CHUNK);
    return;
  }
  const int cv = threadIdx.x % CVEC;
  const uint64_t pol_first = l2_evict_first_policy();
  const f2_t quarter2 = f2_splat(0.25f), half2 = f2_splat(0.5f);
  // first-sweep state (image ph): pivot and running moments; second-sweep state (image ph-1): coefficients
  float pivot[VE], acc[2][VE];
  f2_t npivot2[NP], sum2[NP], sq2[NP];          // bf16: packed pairs
  f2_t nmean2[NP], inv8v2[NP];                  // bf16 second sweep
  float piv2[VE], dmean2[VE], inv4v2[VE];       // fp32 second sweep
  int ring_s = 0;
  uint32_t ring_par = 0;
  auto release = [&](int s) {
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(st_smem_u32(&empty[s])) : "memory");
  };
  int64_t b1 = 0;
  uint4* dst = nullptr;
  auto start_first = [&](int ph) {
    b1 = (int64_t)cid + (int64_t)ph * ncl;
    const uint4* ximg = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(x) + b1 * image_bytes);
    unpack<T>(__ldg(ximg + cv), pivot);  // row 0 of the image, this thread's channels
#pragma unroll
    for (int e = 0; e < VE; ++e) acc[0][e] = acc[1][e] = 0.f;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      npivot2[q] = f2_make(-pivot[2 * q], -pivot[2 * q + 1]);
      sum2[q] = sq2[q] = f2_splat(0.f);
    }
  };
  auto start_second = [&](int ph) {
    const int64_t b2 = (int64_t)cid + (int64_t)(ph - 1) * ncl;
    dst = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(y) + b2 * image_bytes + rank * cta_bytes);
  };
  auto second_chunk = [&](int c) {
    const int s = ring_s;
    st_mbar_wait(&full[s], ring_par);
    if (++ring_s == STAGES) {
      ring_s = 0;
      ring_par ^= 1u;
    }
    const uint4* v = reinterpret_cast<const uint4*>(ring + s * ST_CHUNK);
#pragma unroll
    for (int i = 0; i < VPC; ++i) {
      const uint4 u = v[threadIdx.x + i * ST_CONSUMERS];
      uint4 r;
      if constexpr (BF) {
        // y = hx tanh(t^2 / (8v) + 1/4) + hx, hx = x / 2, on packed pairs
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const f2_t x2 = f2_from_bf16x2(w[q]);
          const f2_t t = f2_add(x2, nmean2[q]);
          const f2_t th = f2_tanh(f2_fma(f2_mul(t, t), inv8v2[q], quarter2));
          const f2_t hx = f2_mul(x2, half2);
          float lo, hi;
          f2_split(f2_fma(hx, th, hx), lo, hi);
          o[q] = pack_bf16x2(lo, hi);
        }
        r = make_uint4(o[0], o[1], o[2], o[3]);
      } else {
        float f[VE];
        unpack<T>(u, f);
#pragma unroll
        for (int e = 0; e < VE; ++e) f[e] = simam_fwd_elem<T>(f[e], FwdCoef{piv2[e], dmean2[e], inv4v2[e]});
        r = pack<T>(f);
      }
      st_stream_first(dst + (int64_t)(chunks - 1 - c) * (ST_CHUNK / 16) + threadIdx.x + i * ST_CONSUMERS, r, pol_first);
    }
    release(s);
  };
  auto first_chunk = [&](int c) {
    const int s = ring_s;
    st_mbar_wait(&full[s], ring_par);
    if (++ring_s == STAGES) {
      ring_s = 0;
      ring_par ^= 1u;
    }
    const uint4* v = reinterpret_cast<const uint4*>(ring + s * ST_CHUNK);
#pragma unroll
    for (int i = 0; i < VPC; ++i) {
      const uint4 u = v[threadIdx.x + i * ST_CONSUMERS];
      if constexpr (BF) {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const f2_t d = f2_add(f2_from_bf16x2(w[q]), npivot2[q]);
          sum2[q] = f2_add(sum2[q], d);
          sq2[q] = f2_fma(d, d, sq2[q]);
        }
      } else {
        float f[VE];
        unpack<T>(u, f);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          const float d = f[e] - pivot[e];
          acc[0][e] += d;
          acc[1][e] = fmaf(d, d, acc[1][e]);
        }
      }
    }
    release(s);
  };
  auto finish_first = [&](int ph) {
    if constexpr (BF) {
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        f2_split(sum2[q], acc[0][2 * q], acc[0][2 * q + 1]);
        f2_split(sq2[q], acc[1][2 * q], acc[1][2 * q + 1]);
      }
    }
    nl_reduce<VE, CVEC>(acc, st_smem, ph & 1, CL, rank);
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      const float dmean = acc[0][e] / (float)L;
      const float var = fmaxf(acc[1][e] - acc[0][e] * dmean, 0.f) / ((float)L - 1.f) + e_lambda;
      acc[0][e] = dmean;
      acc[1][e] = var;
      piv2[e] = pivot[e];
      dmean2[e] = dmean;
      inv4v2[e] = 1.f / (4.f * var);
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      nmean2[q] = f2_make(-(pivot[2 * q] + acc[0][2 * q]), -(pivot[2 * q + 1] + acc[0][2 * q + 1]));
      inv8v2[q] = f2_make(1.f / (8.f * acc[1][2 * q]), 1.f / (8.f * acc[1][2 * q + 1]));
    }
    if (stats != nullptr && rank == 0 && threadIdx.x < CVEC) {
      const int64_t p0 = b1 * S::CW + cv * VE;
#pragma unroll
      for (int e = 0; e < VE; ++e) {
        stats[2 * (p0 + e)] = acc[0][e];
        stats[2 * (p0 + e) + 1] = acc[1][e];
      }
    }
  };
  if constexpr (!PIPE) {
    // one image per cluster: the two sweeps never overlap and their state shares registers
    start_first(0);
    for (int c = 0; c < chunks; ++c) first_chunk(c);
    finish_first(0);
    start_second(1);
    for (int c = 0; c < chunks; ++c) second_chunk(c);
  } else {
    for (int ph = 0; ph <= my_images; ++ph) {
      const bool first = ph < my_images, second = ph > 0;
      if (first) start_first(ph);
      if (second) start_second(ph);
      for (int c = 0; c < chunks; ++c) {
        if (second) second_chunk(c);
        if (first) first_chunk(c);
      }
      if (first) finish_first(ph);
    }
  }
}

template <typename T, int CVEC, bool PIPE>
__global__ void __launch_bounds__(ST_CONSUMERS + 32, 1)
    simam_nlc_bwd_stream(const T* __restrict__ x, const T* __restrict__ gy, const float* __restrict__ stats,
                         T* __restrict__ gx, int B, int L, int chunks /* 8-KB chunks per CTA per image */) {
  constexpr int VE = Vec16<T>::N, NP = VE / 2, HALF = ST_CHUNK / 2, VPC = HALF / 16 / ST_CONSUMERS;
  constexpr bool BF = sizeof(T) == 2;
  using S = NlSmem<VE, CVEC>;
  constexpr int STAGES = S::STAGES;
  extern __shared__ __align__(128) uint8_t st_smem[];
  uint8_t* ring = st_smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(st_smem + S::BARS);
  uint64_t* empty = full + STAGES;
  const int CL = (int)cluster_nctarank(), rank = (int)cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  nl_init<VE, CVEC>(st_smem, CL);
  const int cid = (int)blockIdx.x / CL, ncl = (int)gridDim.x / CL;
  const int my_images = (B - cid + ncl - 1) / ncl;
  const int64_t cta_bytes = (int64_t)chunks * HALF, image_bytes = cta_bytes * CL;

  if (warp == ST_CONSUMERS / 32) {
    if (lane == 0)
      nl_produce<STAGES>(st_smem, S::BARS, reinterpret_cast<const uint8_t*>(x), reinterpret_cast<const uint8_t*>(gy),
                         my_images, cid, ncl, image_bytes, rank * cta_bytes, chunks, HALF);
    return;
  }
  const int cv = threadIdx.x % CVEC;
  const float Lf = (float)L;
  const uint64_t pol_first = l2_evict_first_policy();
  const f2_t quarter2 = f2_splat(0.25f), mone2 = f2_splat(-1.f), half2 = f2_splat(0.5f);
  // first-sweep state (image ph)
  float pivot[VE], dmean[VE], inv4v[VE], vv[VE], acc[2][VE];
  f2_t r1p[NP], r2p[NP], nmean2[NP], inv8v2[NP];
  // second-sweep state (image ph-1)
  float piv_b[VE], dmean_b[VE], inv4v_b[VE], c1_b[VE], c2_b[VE];
  f2_t nmean2_b[NP], inv8v2_b[NP], nk2_b[NP], nc2_b[NP];
  int ring_s = 0;
  uint32_t ring_par = 0;
  auto release = [&](int s) {
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(st_smem_u32(&empty[s])) : "memory");
  };
  int64_t b1 = 0;
  uint4* dst = nullptr;
  auto start_first = [&](int ph) {
    b1 = (int64_t)cid + (int64_t)ph * ncl;
    const uint4* ximg = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(x) + b1 * image_bytes);
    unpack<T>(__ldg(ximg + cv), pivot);
    const int64_t p0 = b1 * S::CW + cv * VE;
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      dmean[e] = __ldg(stats + 2 * (p0 + e));
      vv[e] = __ldg(stats + 2 * (p0 + e) + 1);
      inv4v[e] = 1.f / (4.f * vv[e]);
      acc[0][e] = acc[1][e] = 0.f;
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      r1p[q] = r2p[q] = f2_splat(0.f);
      nmean2[q] = f2_make(-(pivot[2 * q] + dmean[2 * q]), -(pivot[2 * q + 1] + dmean[2 * q + 1]));
      inv8v2[q] = f2_make(0.5f * inv4v[2 * q], 0.5f * inv4v[2 * q + 1]);
    }
  };
  auto start_second = [&](int ph) {
    const int64_t b2 = (int64_t)cid + (int64_t)(ph - 1) * ncl;
    dst = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(gx) + b2 * image_bytes + rank * cta_bytes);
  };
  auto second_chunk = [&](int c) {
    const int s = ring_s;
    st_mbar_wait(&full[s], ring_par);
    if (++ring_s == STAGES) {
      ring_s = 0;
      ring_par ^= 1u;
    }
    const uint4* vx = reinterpret_cast<const uint4*>(ring + s * ST_CHUNK);
    const uint4* vg = reinterpret_cast<const uint4*>(ring + s * ST_CHUNK + HALF);
#pragma unroll
    for (int i = 0; i < VPC; ++i) {
      const uint4 ux = vx[threadIdx.x + i * ST_CONSUMERS], ug = vg[threadIdx.x + i * ST_CONSUMERS];
      uint4 res;
      if constexpr (BF) {
        // grad_x = 0.5 g (1 + tanh) + t (4a k1 - k2) - c2 with k1 = 1/(8v); the pairs hold -4a
        const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w}, wg[4] = {ug.x, ug.y, ug.z, ug.w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const f2_t x2 = f2_from_bf16x2(wx[q]), g2 = f2_from_bf16x2(wg[q]);
          const f2_t t = f2_add(x2, nmean2_b[q]), dd = f2_mul(t, t);
          const f2_t th = f2_tanh(f2_fma(dd, inv8v2_b[q], quarter2));
          const f2_t na4 = f2_mul(f2_mul(g2, x2), f2_fma(th, th, mone2));
          const f2_t hg = f2_mul(g2, half2);
          const f2_t gs = f2_fma(hg, th, f2_add(hg, nc2_b[q]));
          // t (-4a)(-k1) = t (na4 * inv8v) with the sign carried by nk2: t (na4 (-k1) - k2)
          const f2_t r = f2_fma(t, f2_fma(f2_mul(na4, mone2), inv8v2_b[q], nk2_b[q]), gs);
          float lo, hi;
          f2_split(r, lo, hi);
          o[q] = pack_bf16x2(lo, hi);
        }
        res = make_uint4(o[0], o[1], o[2], o[3]);
      } else {
        float fx[VE], fg[VE];
        unpack<T>(ux, fx);
        unpack<T>(ug, fg);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          const float t = centred(fx[e], piv_b[e], dmean_b[e]), dd = t * t;
          const float sg_ = Sig<T>::f(fmaf(dd, inv4v_b[e], 0.5f));
          const float a = fg[e] * fx[e] * sg_ * (1.f - sg_);
          fx[e] = fmaf(fg[e], sg_, 2.f * t * fmaf(a, inv4v_b[e], -c1_b[e])) - c2_b[e];
        }
        res = pack<T>(fx);
      }
      st_stream_first(dst + (int64_t)(chunks - 1 - c) * (HALF / 16) + threadIdx.x + i * ST_CONSUMERS, res, pol_first);
    }
    release(s);
  };
  auto first_chunk = [&](int c) {
    const int s = ring_s;
    st_mbar_wait(&full[s], ring_par);
    if (++ring_s == STAGES) {
      ring_s = 0;
      ring_par ^= 1u;
    }
    const uint4* vx = reinterpret_cast<const uint4*>(ring + s * ST_CHUNK);
    const uint4* vg = reinterpret_cast<const uint4*>(ring + s * ST_CHUNK + HALF);
#pragma unroll
    for (int i = 0; i < VPC; ++i) {
      const uint4 ux = vx[threadIdx.x + i * ST_CONSUMERS], ug = vg[threadIdx.x + i * ST_CONSUMERS];
      if constexpr (BF) {
        const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w}, wg[4] = {ug.x, ug.y, ug.z, ug.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const f2_t x2 = f2_from_bf16x2(wx[q]), g2 = f2_from_bf16x2(wg[q]);
          const f2_t t = f2_add(x2, nmean2[q]), dd = f2_mul(t, t);
          const f2_t th = f2_tanh(f2_fma(dd, inv8v2[q], quarter2));
          const f2_t na4 = f2_mul(f2_mul(g2, x2), f2_fma(th, th, mone2));  // -4a = g x (tanh^2 - 1)
          r1p[q] = f2_fma(na4, dd, r1p[q]);
          r2p[q] = f2_fma(na4, t, r2p[q]);
        }
      } else {
        float fx[VE], fg[VE];
        unpack<T>(ux, fx);
        unpack<T>(ug, fg);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          const float t = centred(fx[e], pivot[e], dmean[e]), dd = t * t;
          const float sg_ = Sig<T>::f(fmaf(dd, inv4v[e], 0.5f));
          const float a4 = 4.f * fg[e] * fx[e] * sg_ * (1.f - sg_);
          acc[0][e] = fmaf(a4, dd, acc[0][e]);
          acc[1][e] = fmaf(a4, t, acc[1][e]);
        }
      }
    }
    release(s);
  };
  auto finish_first = [&](int ph) {
    if constexpr (BF) {
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        float lo, hi;
        f2_split(r1p[q], lo, hi);
        acc[0][2 * q] = -lo;
        acc[0][2 * q + 1] = -hi;
        f2_split(r2p[q], lo, hi);
        acc[1][2 * q] = -lo;
        acc[1][2 * q + 1] = -hi;
      }
    }
    nl_reduce<VE, CVEC>(acc, st_smem, ph & 1, CL, rank);
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      const float r1 = 0.25f * acc[0][e], r2 = 0.25f * acc[1][e];  // the sweeps accumulate 4 a
      c1_b[e] = r1 * inv4v[e] / (vv[e] * (Lf - 1.f));              // R1 / (4 v^2 n)
      c2_b[e] = 2.f / Lf * r2 * inv4v[e];                          // (2/L) R2
      piv_b[e] = pivot[e];
      dmean_b[e] = dmean[e];
      inv4v_b[e] = inv4v[e];
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      nmean2_b[q] = nmean2[q];
      inv8v2_b[q] = inv8v2[q];
      nk2_b[q] = f2_make(-2.f * c1_b[2 * q], -2.f * c1_b[2 * q + 1]);
      nc2_b[q] = f2_make(-c2_b[2 * q], -c2_b[2 * q + 1]);
    }
  };
  if constexpr (!PIPE) {
    // one image per cluster: the two sweeps never overlap and their state shares registers
    start_first(0);
    for (int c = 0; c < chunks; ++c) first_chunk(c);
    finish_first(0);
    start_second(1);
    for (int c = 0; c < chunks; ++c) second_chunk(c);
  } else {
    for (int ph = 0; ph <= my_images; ++ph) {
      const bool first = ph < my_images, second = ph > 0;
      if (first) start_first(ph);
      if (second) start_second(ph);
      for (int c = 0; c < chunks; ++c) {
        if (second) second_chunk(c);
        if (first) first_chunk(c);
      }
      if (first) finish_first(ph);
    }
  }
}

// Launch: the largest cluster (<= 8 CTAs) that still gives every image of the batch its own cluster in
// ONE wave and divides the image into whole chunks; persistent, pipelined clusters when the batch is larger.
template <typename T, bool BWD, int CVEC, bool PIPE>
int nlc_stream_launch2(const cudaLaunchConfig_t& cfg, const T* x, const T* gy, float* stats_out, const float* stats_in,
                       T* out, int B, int L, int chunks, float e_lambda) {
  constexpr int SMEM = NlSmem<Vec16<T>::N, CVEC>::BYTES;
  cudaError_t e;
  if constexpr (BWD) e = opt_in_smem(reinterpret_cast<const void*>(&simam_nlc_bwd_stream<T, CVEC, PIPE>), SMEM);
  else e = opt_in_smem(reinterpret_cast<const void*>(&simam_nlc_fwd_stream<T, CVEC, PIPE>), SMEM);
  if (e != cudaSuccess) return fail(CSB200_ERR_CUDA, "simam_nlc_stream: %s", cudaGetErrorString(e));
  if constexpr (BWD)
    e = cudaLaunchKernelEx(&cfg, simam_nlc_bwd_stream<T, CVEC, PIPE>, x, gy, stats_in, out, B, L, chunks);
  else
    e = cudaLaunchKernelEx(&cfg, simam_nlc_fwd_stream<T, CVEC, PIPE>, x, out, stats_out, B, L, chunks, e_lambda);
  if (e != cudaSuccess) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return fail(CSB200_ERR_CUDA, "simam_nlc_stream: %s", cudaGetErrorString(e));
  }
  return check_launch(BWD ? "simam_nlc_bwd_stream" : "simam_nlc_fwd_stream");
}

// Launch: the largest cluster (<= 8 CTAs) that still gives every image of the batch its own cluster in
// ONE wave and divides the image into whole chunks; persistent, pipelined clusters when the batch is larger.
template <typename T, bool BWD, int CVEC>
int nlc_stream_launch(const T* x, const T* gy, float* stats_out, const float* stats_in, T* out, int64_t B,
                      int64_t L, int64_t image_bytes, float e_lambda, cudaStream_t st) {
  constexpr int VE = Vec16<T>::N;
  constexpr int PART = BWD ? ST_CHUNK / 2 : ST_CHUNK;
  constexpr int SMEM = NlSmem<VE, CVEC>::BYTES;
  const int sms = device_sm_count();
  if (sms <= 0) return -1;
  const void* kfun = BWD ? reinterpret_cast<const void*>(&simam_nlc_bwd_stream<T, CVEC, true>)
                         : reinterpret_cast<const void*>(&simam_nlc_fwd_stream<T, CVEC, true>);
  if (opt_in_smem(kfun, SMEM) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  int cl = 8, lg = 3;
  while (cl > 1 && (B * cl > sms || image_bytes % ((int64_t)cl * PART) != 0 ||
                    image_bytes / ((int64_t)cl * PART) < 2)) {
    cl >>= 1;
    --lg;
  }
  if (image_bytes % ((int64_t)cl * PART) != 0 || image_bytes / ((int64_t)cl * PART) > 0x7fffffff) return -1;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(ST_CONSUMERS + 32);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int resident = 0;  // clusters of this size that are co-resident on the current device (memo per device)
  if (!memo_get(kfun, 16 + lg, &resident)) {
    int n = 0;
    cfg.gridDim = dim3((unsigned)(sms / cl * cl));
    cudaError_t e;
    if constexpr (BWD) e = cudaOccupancyMaxActiveClusters(&n, simam_nlc_bwd_stream<T, CVEC, true>, &cfg);
    else e = cudaOccupancyMaxActiveClusters(&n, simam_nlc_fwd_stream<T, CVEC, true>, &cfg);
    if (e != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = sms / cl > 1 ? sms / cl / 2 : 1;  // conservative: correctness never depends on co-residency of clusters
    }
    resident = n;
    memo_put(kfun, 16 + lg, n);
  }
  // even rounds: with r = ceil(B / resident) rounds, ceil(B / r) clusters finish together
  const int64_t rounds = (B + resident - 1) / resident;
  const int64_t nclusters = (B + rounds - 1) / rounds;
  cfg.gridDim = dim3((unsigned)(nclusters * cl));
  const int chunks = (int)(image_bytes / ((int64_t)cl * PART));
  if (rounds > 1)
    return nlc_stream_launch2<T, BWD, CVEC, true>(cfg, x, gy, stats_out, stats_in, out, (int)B, (int)L, chunks, e_lambda);
  return nlc_stream_launch2<T, BWD, CVEC, false>(cfg, x, gy, stats_out, stats_in, out, (int)B, (int)L, chunks, e_lambda);
}

// Rows of 8..64 16-byte vectors (bf16: 64..512 channels, fp32: 32..256), images of >= 128 KB.
template <typename T, bool BWD>
int nlc_stream(const T* x, const T* gy, float* stats_out, const float* stats_in, T* out, int64_t B, int64_t C,
               int64_t L, float e_lambda, cudaStream_t st) {
  constexpr int VE = Vec16<T>::N;
  if (C % VE != 0 || !aligned16(x) || !aligned16(out) || (BWD && !aligned16(gy))) return -1;
  if (B > 0x7fffffff || L > 0x7fffffff || L < 2) return -1;
  const int64_t cvec = C / VE, image_bytes = L * C * (int64_t)sizeof(T);
  if (image_bytes < (128 << 10)) return -1;
#define CSB_NLS(CV) \
  if (cvec == CV) return nlc_stream_launch<T, BWD, CV>(x, gy, stats_out, stats_in, out, B, L, image_bytes, e_lambda, st);
  CSB_NLS(8)
  CSB_NLS(16)
  CSB_NLS(32)
  CSB_NLS(64)
#undef CSB_NLS
  return -1;
}

#include "simam_grid.cuh"

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
template <typename K, typename... Args>
int launch(K kernel, int64_t grid, int threads, int cluster, cudaStream_t st, const char* name,
           Args... args) {
  if (grid <= 0) return CSB200_OK;
  if (grid > 0x7fffffffLL) return fail(CSB200_ERR_INVALID, "%s: grid too large", name);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = cluster > 1 ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
  if (e != cudaSuccess) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return fail(CSB200_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e));
  }
  return check_launch(name);
}


// NCHW resident dispatch.  Returns -1 if no resident configuration fits.
template <typename T, bool BWD>
int nchw_resident(const T* x, const T* gy, float* stats_out, const float* stats_in, T* out,
                  int64_t planes, int64_t S, float e_lambda, cudaStream_t st) {
  constexpr int VE = Vec16<T>::N;
  if (S % VE != 0 || !aligned16(x) || !aligned16(out) || (BWD && !aligned16(gy))) return -1;
  const int64_t nvec64 = S / VE;
  if (nvec64 > 32768) return -1;
  const int nvec = (int)nvec64;
  const float Sf = (float)S;
#define CSB_NCHW(VPT, THREADS, CLUSTER, WARP)                                                     \
  do {                                                                                            \
    const int64_t grid = (WARP) ? (planes + (THREADS) / 32 - 1) / ((THREADS) / 32)                \
                                : planes * (CLUSTER);                                             \
    if constexpr (BWD)                                                                            \
      return launch(simam_nchw_bwd_resident<T, VPT, THREADS, CLUSTER, WARP>, grid, THREADS,       \
                    CLUSTER, st, "simam_nchw_bwd_resident", x, gy, stats_in, out, planes, nvec,   \
                    Sf);                                                                          \
    else                                                                                          \
      return launch(simam_nchw_fwd_resident<T, VPT, THREADS, CLUSTER, WARP>, grid, THREADS,       \
                    CLUSTER, st, "simam_nchw_fwd_resident", x, out, stats_out, planes, nvec, Sf,  \
                    e_lambda);                                                                    \
  } while (0)
  // CTAs of 256 threads keep 4 (forward) / 3 (backward) of them resident per SM, so the load, reduce
  // and store phases of different planes overlap; a plane larger than one CTA spans a cluster.
  if constexpr (!BWD) {
    if (nvec <= 32) CSB_NCHW(1, 256, 1, true);
    if (nvec <= 64) CSB_NCHW(2, 256, 1, true);
    if (nvec <= 128) CSB_NCHW(4, 256, 1, true);
    if (nvec <= 256) CSB_NCHW(8, 256, 1, true);
    if (nvec <= 512) CSB_NCHW(4, 128, 1, false);
    if (nvec <= 1024) CSB_NCHW(4, 256, 1, false);
    if (nvec <= 2048) CSB_NCHW(8, 256, 1, false);
    if (nvec <= 4096) CSB_NCHW(8, 256, 2, false);
    if (nvec <= 8192) CSB_NCHW(8, 256, 4, false);
    if (nvec <= 16384) CSB_NCHW(8, 256, 8, false);
    CSB_NCHW(8, 512, 8, false);
  } else {
    if (nvec <= 32) CSB_NCHW(1, 256, 1, true);
    if (nvec <= 64) CSB_NCHW(2, 256, 1, true);
    if (nvec <= 128) CSB_NCHW(4, 256, 1, true);
    if (nvec <= 512) CSB_NCHW(4, 128, 1, false);
    if (nvec <= 1024) CSB_NCHW(4, 256, 1, false);
    if (nvec <= 2048) CSB_NCHW(4, 256, 2, false);
    if (nvec <= 4096) CSB_NCHW(4, 256, 4, false);
    if (nvec <= 8192) CSB_NCHW(8, 256, 4, false);
    if (nvec <= 16384) CSB_NCHW(8, 256, 8, false);
    CSB_NCHW(8, 512, 8, false);
  }
#undef CSB_NCHW
}

// NLC two-sweep dispatch: widest slab (128/64/32 B per row) that divides the row, and the smallest
// cluster that puts >= 2 CTAs on every SM (or 8).
template <typename T, bool BWD>
int nlc_2pass(const T* x, const T* gy, float* stats_out, const float* stats_in, T* out, int64_t B,
              int64_t C, int64_t L, float e_lambda, cudaStream_t st) {
  constexpr int VE = Vec16<T>::N;
  if (C % VE != 0 || !aligned16(x) || !aligned16(out) || (BWD && !aligned16(gy))) return -1;
  if (L > 0x7fffffff / 64 || B * C > 0x7fffffff || L < 2) return -1;
  const int cvec = (int)(C / VE);
  const int cwv = cvec % 8 == 0 ? 8 : (cvec % 4 == 0 ? 4 : (cvec % 2 == 0 ? 2 : 0));
  if (cwv == 0) return -1;
  const int slabs = cvec / cwv;
  const int64_t groups = B * slabs;
  int cluster = 1;
  while (cluster < 8 && groups * cluster < 296 && L / (cluster * 2) >= 256 / cwv) cluster *= 2;
  // One cluster per group, all resident at once: measured faster than throttling the groups in
  // flight to an L2 budget (more bytes in flight beats a higher second-sweep hit rate).
  const int64_t resident = groups;
  const int ngroups = (int)groups;
#define CSB_2P(CWV, CL)                                                                           \
  do {                                                                                            \
    if constexpr (BWD)                                                                            \
      return launch(simam_nlc_bwd_2pass<T, CWV, CL>, resident * (CL), 256, CL, st,                \
                    "simam_nlc_bwd_2pass", x, gy, stats_in, out, (int)L, (int)C, slabs, ngroups); \
    else                                                                                          \
      return launch(simam_nlc_fwd_2pass<T, CWV, CL>, resident * (CL), 256, CL, st,                \
                    "simam_nlc_fwd_2pass", x, out, stats_out, (int)L, (int)C, slabs, ngroups,     \
                    e_lambda);                                                                    \
  } while (0)
#define CSB_2P_CL(CWV)                     \
  if (cluster == 1) CSB_2P(CWV, 1);        \
  if (cluster == 2) CSB_2P(CWV, 2);        \
  if (cluster == 4) CSB_2P(CWV, 4);        \
  CSB_2P(CWV, 8);
  if (cwv == 8) { CSB_2P_CL(8) }
  if (cwv == 4) { CSB_2P_CL(4) }
  CSB_2P_CL(2)
#undef CSB_2P_CL
#undef CSB_2P
}

// NLC resident dispatch: widest slab (128/64/32 B per row) whose L rows fit a cluster of <= 8 CTAs.
template <typename T, bool BWD>
int nlc_resident(const T* x, const T* gy, float* stats_out, const float* stats_in, T* out,
                 int64_t B, int64_t C, int64_t L, float e_lambda, cudaStream_t st) {
  constexpr int VE = Vec16<T>::N;
  constexpr int THREADS = 512;
  if (C % VE != 0 || !aligned16(x) || !aligned16(out) || (BWD && !aligned16(gy))) return -1;
  if (L > 0x7fffffff / 64 || B * C > 0x7fffffff) return -1;
  const int cvec = (int)(C / VE);
#define CSB_NLC(VPT, CWV, CLUSTER)                                                                    \
  do {                                                                                            \
    const int slabs = cvec / (CWV);                                                               \
    const int64_t grid = B * slabs * (CLUSTER);                                                   \
    if constexpr (BWD)                                                                            \
      return launch(simam_nlc_bwd_resident<T, VPT, CWV, THREADS, CLUSTER>, grid, THREADS,         \
                    CLUSTER, st, "simam_nlc_bwd_resident", x, gy, stats_in, out, (int)L, (int)C,  \
                    slabs);                                                                       \
    else                                                                                          \
      return launch(simam_nlc_fwd_resident<T, VPT, CWV, THREADS, CLUSTER>, grid, THREADS,         \
                    CLUSTER, st, "simam_nlc_fwd_resident", x, out, stats_out, (int)L, (int)C,     \
                    slabs, e_lambda);                                                             \
  } while (0)
#define CSB_NLC_TRY(VPT, CWV)                                                                     \
  if (cvec % (CWV) == 0) {                                                                        \
    const int64_t rows1 = (int64_t)(VPT) * (THREADS / (CWV)); /* rows one CTA can hold */         \
    if (L <= rows1) CSB_NLC(VPT, CWV, 1);                                                         \
    if (L <= 2 * rows1) CSB_NLC(VPT, CWV, 2);                                                     \
    if (L <= 4 * rows1) CSB_NLC(VPT, CWV, 4);                                                     \
    if (L <= 8 * rows1) CSB_NLC(VPT, CWV, 8);                                                     \
  }
  if constexpr (BWD) {
    CSB_NLC_TRY(4, 8)
    CSB_NLC_TRY(4, 4)
    CSB_NLC_TRY(4, 2)
    CSB_NLC_TRY(8, 2)
  } else {
    CSB_NLC_TRY(8, 8)
    CSB_NLC_TRY(8, 4)
    CSB_NLC_TRY(8, 2)
  }
#undef CSB_NLC_TRY
#undef CSB_NLC
  return -1;
}

template <typename T, bool BWD>
int simam_dispatch(const void* x_, const void* gy_, float* stats_out, const float* stats_in,
                   void* out_, int64_t B, int64_t C, int64_t S, int layout, float e_lambda,
                   cudaStream_t st, void* ws = nullptr, size_t ws_bytes = 0) {
  const T* x = static_cast<const T*>(x_);
  const T* gy = static_cast<const T*>(gy_);
  T* out = static_cast<T*>(out_);
  if (layout == CSB200_NLC && ws != nullptr) {
    const int rs = nlc_grid<T, BWD>(x, gy, stats_out, stats_in, out, B, C, S, e_lambda, ws, ws_bytes, st);
    if (rs >= 0) return rs;
  }
  if (layout == CSB200_NCHW) {
    int rs;
    if constexpr (BWD) rs = nchw_bwd_stream<T>(x, gy, stats_in, out, B * C, S, st);
    else rs = nchw_fwd_stream<T>(x, out, stats_out, B * C, S, e_lambda, st);
    if (rs >= 0) return rs;
  }
  if (layout == CSB200_NLC) {
    const int rs = nlc_stream<T, BWD>(x, gy, stats_out, stats_in, out, B, C, S, e_lambda, st);
    if (rs >= 0) return rs;
  }
  int rc = (layout == CSB200_NCHW)
               ? nchw_resident<T, BWD>(x, gy, stats_out, stats_in, out, B * C, S, e_lambda, st)
               : nlc_2pass<T, BWD>(x, gy, stats_out, stats_in, out, B, C, S, e_lambda, st);
  if (rc < 0 && layout == CSB200_NLC)
    rc = nlc_resident<T, BWD>(x, gy, stats_out, stats_in, out, B, C, S, e_lambda, st);
  if (rc >= 0) return rc;
  const int64_t grid = (layout == CSB200_NCHW) ? B * C : B * ((C + 31) / 32);
  return launch(simam_generic<T, BWD>, grid, 256, 1, st, "simam_generic", x, gy, stats_out,
                stats_in, out, S, C, layout, e_lambda);
}

int simam_check(const void* x, const void* out, int64_t B, int64_t C, int64_t S, int layout,
                int dtype) {
  if (B < 0 || C < 0 || S < 0) return fail(CSB200_ERR_INVALID, "simam: negative size");
  if (layout != CSB200_NCHW && layout != CSB200_NLC)
    return fail(CSB200_ERR_INVALID, "simam: unknown layout %d", layout);
  if (dtype != CSB200_F32 && dtype != CSB200_BF16)
    return fail(CSB200_ERR_INVALID, "simam: unknown dtype %d", dtype);
  if (B * C * S > 0 && (x == nullptr || out == nullptr))
    return fail(CSB200_ERR_INVALID, "simam: null pointer");
  return CSB200_OK;
}

}  // namespace
}  // namespace csb200

using namespace csb200;

#ifdef CSB_PROF
extern "C" __attribute__((visibility("default"))) int csb200_debug_prof_simam(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_simam_prof, sizeof(g_simam_prof));
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_simam_prof, z, sizeof(z));
  }
  return 0;
}
#endif

extern "C" size_t csb200_simam_workspace_bytes(int64_t batch, int64_t channels, int64_t spatial, int layout,
                                               int dtype) {
  if (layout != CSB200_NLC || batch <= 0 || channels <= 0 || spatial <= 0) return 0;
  const int sms = device_sm_count();
  if (sms <= 0) return 0;
  GridPlan plan;
  size_t fwd = 0, bwd = 0;
  if (dtype == CSB200_F32) {
    if (!grid_plan<float, false>(batch, channels, spatial, sms, &plan, &fwd)) fwd = 0;
    if (!grid_plan<float, true>(batch, channels, spatial, sms, &plan, &bwd)) bwd = 0;
  } else if (dtype == CSB200_BF16) {
    if (!grid_plan<__nv_bfloat16, false>(batch, channels, spatial, sms, &plan, &fwd)) fwd = 0;
    if (!grid_plan<__nv_bfloat16, true>(batch, channels, spatial, sms, &plan, &bwd)) bwd = 0;
  }
  return fwd > bwd ? fwd : bwd;
}

extern "C" int csb200_simam_fwd_ws(const void* x, void* y, float* stats, int64_t batch, int64_t channels,
                                   int64_t spatial, int layout, int dtype, float e_lambda, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  int rc = simam_check(x, y, batch, channels, spatial, layout, dtype);
  if (rc != CSB200_OK) return rc;
  if (batch * channels * spatial == 0) return CSB200_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == CSB200_F32)
    return simam_dispatch<float, false>(x, nullptr, stats, nullptr, y, batch, channels, spatial,
                                        layout, e_lambda, st, workspace, workspace_bytes);
  return simam_dispatch<__nv_bfloat16, false>(x, nullptr, stats, nullptr, y, batch, channels,
                                              spatial, layout, e_lambda, st, workspace, workspace_bytes);
}

extern "C" int csb200_simam_fwd(const void* x, void* y, float* stats, int64_t batch,
                                int64_t channels, int64_t spatial, int layout, int dtype,
                                float e_lambda, void* stream) {
  return csb200_simam_fwd_ws(x, y, stats, batch, channels, spatial, layout, dtype, e_lambda, nullptr, 0, stream);
}

extern "C" int csb200_simam_bwd_ws(const void* x, const void* grad_y, const float* stats, void* grad_x,
                                   int64_t batch, int64_t channels, int64_t spatial, int layout, int dtype,
                                   float e_lambda, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = simam_check(x, grad_x, batch, channels, spatial, layout, dtype);
  if (rc != CSB200_OK) return rc;
  if (batch * channels * spatial == 0) return CSB200_OK;
  if (grad_y == nullptr || stats == nullptr)
    return fail(CSB200_ERR_INVALID, "simam_bwd: grad_y and stats are required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == CSB200_F32)
    return simam_dispatch<float, true>(x, grad_y, nullptr, stats, grad_x, batch, channels, spatial,
                                       layout, e_lambda, st, workspace, workspace_bytes);
  return simam_dispatch<__nv_bfloat16, true>(x, grad_y, nullptr, stats, grad_x, batch, channels,
                                             spatial, layout, e_lambda, st, workspace, workspace_bytes);
}

extern "C" int csb200_simam_bwd(const void* x, const void* grad_y, const float* stats,
                                void* grad_x, int64_t batch, int64_t channels, int64_t spatial,
                                int layout, int dtype, float e_lambda, void* stream) {
  return csb200_simam_bwd_ws(x, grad_y, stats, grad_x, batch, channels, spatial, layout, dtype, e_lambda, nullptr, 0,
                             stream);
}

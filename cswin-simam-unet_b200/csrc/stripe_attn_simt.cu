// Cross-shaped stripe attention with LePE — CUDA-core engine (fp32 accumulate, any stripe shape).
//
// Replaces the body of LePEAttention.forward (reference C:271-298).  The stripe partition
// (img2windows / windows2img, C:199-217; im2cswin, C:248-254; get_lepe, C:256-269) is pure index
// arithmetic here: q/k/v are read in place from the token-major qkv buffer and the result is written
// in place at the branch's channel offset, so none of the reference's 12 full-tensor copies exist.
//
//   token l = y*W + x  ->  stripe (y / h_sp, x / w_sp), in-stripe index n = (y % h_sp)*w_sp + x % w_sp
//
// This engine is the fp32 parity anchor (<= 1e-5 relative vs the reference) and the fallback for
// stripe lengths the tcgen05 engine does not tile (N = 49, 56, 98 at 224^2).  One thread owns one
// query (or key) row of 32 channels in registers; the other operand streams through shared memory
// in 64-row chunks read as warp-wide broadcasts; softmax is the online (running max / sum) form.
// Backward recomputes the probabilities from the saved log-sum-exp (no N x N tensor ever exists).

#include <cstring>
#include <type_traits>

#include "philox.cuh"
#include "stripe_attn.cuh"

namespace csb200 {

template <typename T>
int lepe_bwd_prep_t(const StripeGeom& g, const T* v, const float* lepe_w, const float* lepe_b,
                    const T* out, const T* gout, float* delta, float* partial, float* gw, float* gb,
                    cudaStream_t st);

namespace {

constexpr int HD = 32;      // head_dim (dim // num_heads is 32 in every stage)
constexpr int ROWS = 128;   // query (key) rows per CTA == threads per CTA
constexpr int CHUNK = 64;   // rows of the streamed operand per shared-memory chunk

template <typename T>
struct Exp {
  static __device__ __forceinline__ float f(float x) { return expf(x); }  // fp32: full precision
};
template <>
struct Exp<__nv_bfloat16> {
  static __device__ __forceinline__ float f(float x) { return __expf(x); }
};

// 4 consecutive channels of one token row -> fp32
template <typename T>
__device__ __forceinline__ float4 ld4(const T* p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                     __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}
template <typename T>
__device__ __forceinline__ void st4(T* p, float4 v);
template <>
__device__ __forceinline__ void st4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <>
__device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

template <typename T>
__device__ __forceinline__ void load_row(const T* p, float (&r)[HD]) {
#pragma unroll
  for (int i = 0; i < HD / 4; ++i) {
    float4 f = ld4<T>(p + 4 * i);
    r[4 * i] = f.x;
    r[4 * i + 1] = f.y;
    r[4 * i + 2] = f.z;
    r[4 * i + 3] = f.w;
  }
}
template <typename T>
__device__ __forceinline__ void store_row(T* p, const float (&r)[HD]) {
#pragma unroll
  for (int i = 0; i < HD / 4; ++i)
    st4<T>(p + 4 * i, make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]));
}

// Which (batch, stripe, head, row tile) a CTA works on, and the token index of in-stripe row n.
struct Work {
  int b, head, wy, wx, tile;
};
__device__ __forceinline__ Work decode_work(const StripeGeom& g, int tiles) {
  int id = blockIdx.x;
  Work w;
  w.tile = id % tiles;
  id /= tiles;
  w.head = id % g.heads;
  id /= g.heads;
  w.wx = id % g.nwx;
  id /= g.nwx;
  w.wy = id % g.nwy;
  w.b = id / g.nwy;
  return w;
}
__device__ __forceinline__ int token_of(const StripeGeom& g, const Work& w, int n) {
  return (w.wy * g.hs + n / g.ws) * g.W + w.wx * g.ws + n % g.ws;
}

// Cooperative load of `cnt` stripe rows [n0, n0+cnt) of one operand into smem (fp32, [CHUNK][HD]).
template <typename T>
__device__ __forceinline__ void load_chunk(float* dst, const T* base, int64_t sl,
                                           const StripeGeom& g, const Work& w, int n0, int cnt,
                                           float mul) {
  for (int idx = threadIdx.x; idx < cnt * (HD / 4); idx += ROWS) {
    int r = idx / (HD / 4), c4 = idx % (HD / 4);
    float4 f = ld4<T>(base + (int64_t)token_of(g, w, n0 + r) * sl + 4 * c4);
    f.x *= mul; f.y *= mul; f.z *= mul; f.w *= mul;
    *reinterpret_cast<float4*>(dst + r * HD + 4 * c4) = f;
  }
}

__device__ __forceinline__ float dot_smem(const float (&a)[HD], const float* row) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < HD / 4; ++i) {
    float4 k = *reinterpret_cast<const float4*>(row + 4 * i);  // warp-wide broadcast
    s = fmaf(a[4 * i], k.x, s);
    s = fmaf(a[4 * i + 1], k.y, s);
    s = fmaf(a[4 * i + 2], k.z, s);
    s = fmaf(a[4 * i + 3], k.w, s);
  }
  return s;
}
__device__ __forceinline__ void axpy_smem(float (&acc)[HD], float a, const float* row) {
#pragma unroll
  for (int i = 0; i < HD / 4; ++i) {
    float4 v = *reinterpret_cast<const float4*>(row + 4 * i);
    acc[4 * i] = fmaf(a, v.x, acc[4 * i]);
    acc[4 * i + 1] = fmaf(a, v.y, acc[4 * i + 1]);
    acc[4 * i + 2] = fmaf(a, v.z, acc[4 * i + 2]);
    acc[4 * i + 3] = fmaf(a, v.w, acc[4 * i + 3]);
  }
}

// LePE taps for the 32 channels of one head: s_w[tap][c] (tap = ky*3+kx), s_b[c]
__device__ __forceinline__ void load_lepe_weights(float* s_w, float* s_b, const float* lepe_w,
                                                  const float* lepe_b, int head) {
  for (int i = threadIdx.x; i < 9 * HD; i += ROWS) {
    int tap = i / HD, c = i % HD;
    s_w[i] = __ldg(lepe_w + (head * HD + c) * 9 + tap);
  }
  if (threadIdx.x < HD && s_b != nullptr)
    s_b[threadIdx.x] = lepe_b ? __ldg(lepe_b + head * HD + threadIdx.x) : 0.f;
}

// acc[c] += sum over the 3x3 neighbourhood (zero padding at the STRIPE border, C:244,263-265) of
// w[c][tap] * src[neighbour][c].  FLIP selects the transposed stencil used for grad_v.
template <typename T, bool FLIP>
__device__ __forceinline__ void lepe_stencil(float (&acc)[HD], const T* src_bh, int64_t sl,
                                             const StripeGeom& g, const Work& w, int n,
                                             const float* s_w) {
  const int yy = n / g.ws, xx = n % g.ws;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int ny = FLIP ? yy - ky + 1 : yy + ky - 1;
    if (ny < 0 || ny >= g.hs) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int nx = FLIP ? xx - kx + 1 : xx + kx - 1;
      if (nx < 0 || nx >= g.ws) continue;
      const T* p = src_bh + (int64_t)token_of(g, w, ny * g.ws + nx) * sl;
      const float* wt = s_w + (ky * 3 + kx) * HD;
#pragma unroll
      for (int i = 0; i < HD / 4; ++i) {
        float4 f = ld4<T>(p + 4 * i);
        acc[4 * i] = fmaf(wt[4 * i], f.x, acc[4 * i]);
        acc[4 * i + 1] = fmaf(wt[4 * i + 1], f.y, acc[4 * i + 1]);
        acc[4 * i + 2] = fmaf(wt[4 * i + 2], f.z, acc[4 * i + 2]);
        acc[4 * i + 3] = fmaf(wt[4 * i + 3], f.w, acc[4 * i + 3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward: out = softmax(scale q k^T) v + lepe(v),  lse = log sum exp (scaled scores)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(ROWS)
    stripe_fwd_simt(StripeGeom g, const T* __restrict__ q, const T* __restrict__ k,
                    const T* __restrict__ v, const float* __restrict__ lepe_w,
                    const float* __restrict__ lepe_b, T* __restrict__ out,
                    float* __restrict__ lse, int tiles) {
  __shared__ __align__(16) float s_k[CHUNK * HD];
  __shared__ __align__(16) float s_v[CHUNK * HD];
  __shared__ __align__(16) float s_w[9 * HD];
  __shared__ float s_b[HD];
  const Work w = decode_work(g, tiles);
  const int n = w.tile * ROWS + threadIdx.x;
  const bool valid = n < g.N;
  const int co = w.head * HD;
  const T* qb = q + (int64_t)w.b * g.q_sb + co;
  const T* kb = k + (int64_t)w.b * g.k_sb + co;
  const T* vb = v + (int64_t)w.b * g.v_sb + co;

  load_lepe_weights(s_w, s_b, lepe_w, lepe_b, w.head);
  const int tok = valid ? token_of(g, w, n) : 0;
  float qs[HD], acc[HD];
  if (valid) {
    load_row<T>(qb + (int64_t)tok * g.q_sl, qs);
#pragma unroll
    for (int c = 0; c < HD; ++c) qs[c] *= g.scale;  // q = q * scale, C:287
  } else {
#pragma unroll
    for (int c = 0; c < HD; ++c) qs[c] = 0.f;
  }
#pragma unroll
  for (int c = 0; c < HD; ++c) acc[c] = 0.f;
  float m = -INFINITY, l = 0.f;
  // attention dropout (C:290): keep decisions per (query, key) from Philox, 16 keys per call; the warp's 32
  // decisions for a key are one ballot word of the TRANSPOSED mask the backward kernels read back
  const bool drop = g.drop_thr != 0;
  DropRng rng;
  uint32_t unit = 0;
  uint4 rb = make_uint4(0u, 0u, 0u, 0u);
  if (drop) {
    rng = drop_rng_load(g.rng);
    unit = drop_unit(g, w.b, w.wy, w.wx, w.head);
  }
  const int iw = (w.tile * ROWS + (int)threadIdx.x) >> 5;  // mask word of this warp's 32 query rows
  uint32_t* mrow = drop ? g.drop_mask + ((int64_t)w.b * g.heads + w.head) * g.L * g.mask_words + iw : nullptr;

  for (int j0 = 0; j0 < g.N; j0 += CHUNK) {
    const int cnt = min(CHUNK, g.N - j0);
    __syncthreads();
    load_chunk<T>(s_k, kb, g.k_sl, g, w, j0, cnt, 1.f);
    load_chunk<T>(s_v, vb, g.v_sl, g, w, j0, cnt, 1.f);
    __syncthreads();
    for (int j = 0; j < cnt; j += 8) {
      if (drop && ((j & 15) == 0)) rb = drop_bytes(rng, unit, (uint32_t)n, (uint32_t)((j0 + j) >> 4));
      float s[8];
      float mx = m;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        s[u] = (j + u < cnt) ? dot_smem(qs, s_k + (j + u) * HD) : -INFINITY;
        mx = fmaxf(mx, s[u]);
      }
      const float corr = Exp<T>::f(m - mx);  // exp(-inf) = 0 on the first group
      l *= corr;
#pragma unroll
      for (int c = 0; c < HD; ++c) acc[c] *= corr;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (j + u < cnt) {
          const float p = Exp<T>::f(s[u] - mx);
          l += p;  // the softmax normalisation is over ALL probabilities; dropout acts on the normalised ones
          bool keep = true;
          if (drop) {
            keep = drop_byte(rb, j + u) >= g.drop_thr;
            const uint32_t bits = __ballot_sync(0xffffffffu, keep);
            if ((threadIdx.x & 31) == 0 && iw < g.mask_words)
              mrow[(int64_t)token_of(g, w, j0 + j + u) * g.mask_words] = bits;
          }
          if (keep) axpy_smem(acc, p, s_v + (j + u) * HD);
        }
      }
      m = mx;
    }
  }
  if (!valid) return;
  const float inv_l = g.keep_scale / l;  // survivors are scaled by 1 / (1 - p)
  float o[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) o[c] = s_b[c];
  lepe_stencil<T, false>(o, vb, g.v_sl, g, w, n, s_w);
#pragma unroll
  for (int c = 0; c < HD; ++c) o[c] = fmaf(acc[c], inv_l, o[c]);  // attn @ v + lepe, C:292
  store_row<T>(out + (int64_t)w.b * g.o_sb + (int64_t)tok * g.o_sl + co, o);
  lse[((int64_t)w.b * g.heads + w.head) * g.L + tok] = m + logf(l);
}

// ------------------------------------------------------------------------------------------------
// backward, step 2: grad_q.  Thread = query row; K and V stream through smem.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(ROWS)
    stripe_bwd_dq_simt(StripeGeom g, const T* __restrict__ q, const T* __restrict__ k,
                       const T* __restrict__ v, const T* __restrict__ gout,
                       const float* __restrict__ lse, const float* __restrict__ delta,
                       T* __restrict__ dq, int tiles) {
  __shared__ __align__(16) float s_k[CHUNK * HD];
  __shared__ __align__(16) float s_v[CHUNK * HD];
  const Work w = decode_work(g, tiles);
  const int n = w.tile * ROWS + threadIdx.x;
  const bool valid = n < g.N;
  const int co = w.head * HD;
  const int tok = valid ? token_of(g, w, n) : 0;
  float qs[HD], go[HD], acc[HD];
  float lse_i = 0.f, delta_i = 0.f;
  if (valid) {
    load_row<T>(q + (int64_t)w.b * g.q_sb + (int64_t)tok * g.q_sl + co, qs);
    load_row<T>(gout + (int64_t)w.b * g.o_sb + (int64_t)tok * g.o_sl + co, go);
    const int64_t si = ((int64_t)w.b * g.heads + w.head) * g.L + tok;
    lse_i = lse[si];
    delta_i = delta[si];
  } else {
#pragma unroll
    for (int c = 0; c < HD; ++c) qs[c] = go[c] = 0.f;
  }
#pragma unroll
  for (int c = 0; c < HD; ++c) {
    qs[c] *= g.scale;
    acc[c] = 0.f;
  }
  const T* kb = k + (int64_t)w.b * g.k_sb + co;
  const T* vb = v + (int64_t)w.b * g.v_sb + co;
  // dropout: bit (query % 32) of the forward pass's mask word [key token][query / 32]
  const bool drop = g.drop_thr != 0;
  const int iw = (w.tile * ROWS + (int)threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const uint32_t* mrow = drop && iw < g.mask_words
                             ? g.drop_mask + ((int64_t)w.b * g.heads + w.head) * g.L * g.mask_words + iw : nullptr;
  for (int j0 = 0; j0 < g.N; j0 += CHUNK) {
    const int cnt = min(CHUNK, g.N - j0);
    __syncthreads();
    load_chunk<T>(s_k, kb, g.k_sl, g, w, j0, cnt, 1.f);
    load_chunk<T>(s_v, vb, g.v_sl, g, w, j0, cnt, 1.f);
    __syncthreads();
    for (int j = 0; j < cnt; ++j) {
      const float s = dot_smem(qs, s_k + j * HD);
      const float p = Exp<T>::f(s - lse_i);
      float dp = dot_smem(go, s_v + j * HD);
      if (mrow != nullptr) {  // d(dropped P) / dP = keep / (1 - p)
        const uint32_t bits = __ldg(mrow + (int64_t)token_of(g, w, j0 + j) * g.mask_words);
        dp = ((bits >> lane) & 1u) ? dp * g.keep_scale : 0.f;
      }
      axpy_smem(acc, p * (dp - delta_i), s_k + j * HD);
    }
  }
  if (!valid) return;
#pragma unroll
  for (int c = 0; c < HD; ++c) acc[c] *= g.scale;
  store_row<T>(dq + (int64_t)w.b * g.dq_sb + (int64_t)tok * g.dq_sl + co, acc);
}

// ------------------------------------------------------------------------------------------------
// backward, step 3: grad_k and grad_v (attention part + transposed LePE stencil on grad_out).
// Thread = key row; Q (pre-scaled) and grad_out stream through smem with their lse / delta.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(ROWS)
    stripe_bwd_dkv_simt(StripeGeom g, const T* __restrict__ q, const T* __restrict__ k,
                        const T* __restrict__ v, const float* __restrict__ lepe_w,
                        const T* __restrict__ gout, const float* __restrict__ lse,
                        const float* __restrict__ delta, T* __restrict__ dk, T* __restrict__ dv,
                        int tiles) {
  __shared__ __align__(16) float s_q[CHUNK * HD];
  __shared__ __align__(16) float s_g[CHUNK * HD];
  __shared__ __align__(16) float s_w[9 * HD];
  __shared__ float s_lse[CHUNK], s_delta[CHUNK];
  const Work w = decode_work(g, tiles);
  const int n = w.tile * ROWS + threadIdx.x;
  const bool valid = n < g.N;
  const int co = w.head * HD;
  const int tok = valid ? token_of(g, w, n) : 0;
  load_lepe_weights(s_w, nullptr, lepe_w, nullptr, w.head);
  float kr[HD], vr[HD], ak[HD], av[HD];
  if (valid) {
    load_row<T>(k + (int64_t)w.b * g.k_sb + (int64_t)tok * g.k_sl + co, kr);
    load_row<T>(v + (int64_t)w.b * g.v_sb + (int64_t)tok * g.v_sl + co, vr);
  } else {
#pragma unroll
    for (int c = 0; c < HD; ++c) kr[c] = vr[c] = 0.f;
  }
#pragma unroll
  for (int c = 0; c < HD; ++c) ak[c] = av[c] = 0.f;
  const T* qb = q + (int64_t)w.b * g.q_sb + co;
  const T* gb = gout + (int64_t)w.b * g.o_sb + co;
  const int64_t sbase = ((int64_t)w.b * g.heads + w.head) * g.L;
  for (int i0 = 0; i0 < g.N; i0 += CHUNK) {
    const int cnt = min(CHUNK, g.N - i0);
    __syncthreads();
    load_chunk<T>(s_q, qb, g.q_sl, g, w, i0, cnt, g.scale);
    load_chunk<T>(s_g, gb, g.o_sl, g, w, i0, cnt, 1.f);
    if (threadIdx.x < cnt) {
      const int t = token_of(g, w, i0 + threadIdx.x);
      s_lse[threadIdx.x] = lse[sbase + t];
      s_delta[threadIdx.x] = delta[sbase + t];
    }
    __syncthreads();
    // dropout: this key's own mask row holds the keep bits of all queries (CHUNK = 64 queries = 2 words)
    uint32_t m0 = 0xffffffffu, m1 = 0xffffffffu;
    if (g.drop_thr != 0 && valid) {
      const uint32_t* mrow = g.drop_mask + (sbase + tok) * g.mask_words + (i0 >> 5);
      m0 = __ldg(mrow);
      if ((i0 >> 5) + 1 < g.mask_words) m1 = __ldg(mrow + 1);
    }
    const float ks = g.keep_scale;
    for (int i = 0; i < cnt; ++i) {
      const float s = dot_smem(kr, s_q + i * HD);
      const float p = Exp<T>::f(s - s_lse[i]);
      const bool keep = (((i < 32 ? m0 : m1) >> (i & 31)) & 1u) != 0;
      float dp = dot_smem(vr, s_g + i * HD);
      if (keep) axpy_smem(av, p * ks, s_g + i * HD);  // dV = (dropped P)^T dO
      dp = keep ? dp * ks : 0.f;
      axpy_smem(ak, p * (dp - s_delta[i]), s_q + i * HD);
    }
  }
  if (!valid) return;
  lepe_stencil<T, true>(av, gb, g.o_sl, g, w, n, s_w);
  store_row<T>(dk + (int64_t)w.b * g.dk_sb + (int64_t)tok * g.dk_sl + co, ak);
  store_row<T>(dv + (int64_t)w.b * g.dv_sb + (int64_t)tok * g.dv_sl + co, av);
}

// ------------------------------------------------------------------------------------------------
// backward, step 1 (both engines): ONE pass over grad_out, out and the 3x3 neighbourhood of v gives
//   delta[b,h,l] = sum_c grad_out * (out - lepe)      (the row term of the softmax gradient), and
//   per-CTA partials of the depthwise weight / bias gradients of get_v (C:244):
//   gw[c][tap] = sum_tokens grad_out[l][c] * v[neighbour_tap(l)][c],  gb[c] = sum grad_out[l][c].
// A CTA owns one head and a contiguous token range; 4 adjacent lanes cover the 32 channels of a
// token (8 each, one 16/32-byte load), so a warp reads 8 whole token rows per step.  The 80
// accumulators stay in registers across the whole range and are reduced once at the end
// (shuffles -> shared memory -> partial[block][C'][10]); lepe_wgrad_final sums the partials in a
// fixed order, so the result is deterministic.
// ------------------------------------------------------------------------------------------------
constexpr int PREP_THREADS = 256;
constexpr int PREP_TOK_PER_ITER = PREP_THREADS / 8;  // 8 lanes x 4 channels per token

// arguments of one branch; a launch covers up to two (blockIdx.y runs over the heads of both)
template <typename T>
struct PrepBranch {
  StripeGeom g;
  const T *v, *out, *gout;
  const float *lepe_w, *lepe_b;
  float *delta, *partial;
};
template <typename T>
struct PrepArgs {
  PrepBranch<T> br[2];
  int heads0;  // heads of branch 0
};

template <typename T>
__global__ void __launch_bounds__(PREP_THREADS, 3)
    lepe_bwd_prep(const __grid_constant__ PrepArgs<T> args, int tok_per_cta) {
  __shared__ __align__(16) float s_w[10 * HD];       // [tap][c], bias last
  __shared__ float s_part[PREP_THREADS / 32][10 * HD];
  const int which = (int)blockIdx.y >= args.heads0 ? 1 : 0;
  const PrepBranch<T>& A = args.br[which];
  const StripeGeom& g = A.g;
  const T* __restrict__ v = A.v;
  const T* __restrict__ out = A.out;
  const T* __restrict__ gout = A.gout;
  const float* __restrict__ lepe_w = A.lepe_w;
  const float* __restrict__ lepe_b = A.lepe_b;
  float* __restrict__ delta = A.delta;
  float* __restrict__ partial = A.partial;
  const int head = (int)blockIdx.y - (which ? args.heads0 : 0), cp = g.heads * HD;
  for (int i = threadIdx.x; i < 10 * HD; i += PREP_THREADS) {
    const int tap = i / HD, c = i % HD;
    s_w[i] = tap < 9 ? __ldg(lepe_w + (head * HD + c) * 9 + tap) : __ldg(lepe_b + head * HD + c);
  }
  __syncthreads();
  const int cg = threadIdx.x & 7;              // which 4 of the head's 32 channels
  const int co = head * HD + cg * 4;
  const int total = g.B * g.L;                       // < 2^31 (checked on the host)
  const int t_begin = blockIdx.x * tok_per_cta;
  float4 acc[10];
#pragma unroll
  for (int t = 0; t < 10; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int it = threadIdx.x >> 3; it < tok_per_cta; it += PREP_TOK_PER_ITER) {
    const int gt = t_begin + it;
    const bool valid = gt < total;  // no early exit: every lane takes part in the shuffles below
    const int b = valid ? gt / g.L : 0, l = valid ? gt - b * g.L : 0;
    const int y = l / g.W, x = l - y * g.W, yy = y % g.hs, xx = x % g.ws;
    float4 go = make_float4(0.f, 0.f, 0.f, 0.f), o = go;
    if (valid) {
      go = ld4<T>(gout + (int64_t)b * g.o_sb + (int64_t)l * g.o_sl + co);
      o = ld4<T>(out + (int64_t)b * g.o_sb + (int64_t)l * g.o_sl + co);
    }
    float4 lp = *reinterpret_cast<const float4*>(s_w + 9 * HD + cg * 4);
    acc[9].x += go.x; acc[9].y += go.y; acc[9].z += go.z; acc[9].w += go.w;
    const T* vb = v + (int64_t)b * g.v_sb + co;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      if (!valid || yy + ky - 1 < 0 || yy + ky - 1 >= g.hs) continue;  // zero padding at the STRIPE border
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        if (xx + kx - 1 < 0 || xx + kx - 1 >= g.ws) continue;
        const float4 vn = ld4<T>(vb + (int64_t)((y + ky - 1) * g.W + (x + kx - 1)) * g.v_sl);
        const float4 wt = *reinterpret_cast<const float4*>(s_w + (ky * 3 + kx) * HD + cg * 4);
        lp.x = fmaf(wt.x, vn.x, lp.x); lp.y = fmaf(wt.y, vn.y, lp.y);
        lp.z = fmaf(wt.z, vn.z, lp.z); lp.w = fmaf(wt.w, vn.w, lp.w);
        float4& a = acc[ky * 3 + kx];
        a.x = fmaf(go.x, vn.x, a.x); a.y = fmaf(go.y, vn.y, a.y);
        a.z = fmaf(go.z, vn.z, a.z); a.w = fmaf(go.w, vn.w, a.w);
      }
    }
    float d = go.x * (o.x - lp.x) + go.y * (o.y - lp.y) + go.z * (o.z - lp.z) + go.w * (o.w - lp.w);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    if (cg == 0 && valid) delta[((int64_t)b * g.heads + head) * g.L + l] = d;
  }
  // reduce the 40 accumulators over the 4 token lanes of the warp, then over the 8 warps
#pragma unroll
  for (int t = 0; t < 10; ++t) {
    float* f = reinterpret_cast<float*>(&acc[t]);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float a = f[e];
      a += __shfl_xor_sync(0xffffffffu, a, 8);
      a += __shfl_xor_sync(0xffffffffu, a, 16);
      f[e] = a;
    }
  }
  if ((threadIdx.x & 31) < 8) {
#pragma unroll
    for (int t = 0; t < 10; ++t)
      *reinterpret_cast<float4*>(&s_part[threadIdx.x >> 5][t * HD + cg * 4]) = acc[t];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 10 * HD; i += PREP_THREADS) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < PREP_THREADS / 32; ++w) a += s_part[w][i];
    const int tap = i / HD, c = i % HD;
    partial[((int64_t)blockIdx.x * cp + head * HD + c) * 10 + tap] = a;
  }
}

// one warp per (channel, tap): lanes stride over the per-CTA partials, fixed order -> deterministic
struct WgradFinal {
  const float* partial;
  int cp;
  float *gw, *gb;
};
__global__ void __launch_bounds__(256)
    lepe_wgrad_final(WgradFinal f0, WgradFinal f1, int blocks) {
  const WgradFinal f = blockIdx.y ? f1 : f0;
  const int i = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;  // (c, tap)
  if (i >= f.cp * 10) return;
  const float a = strided_partial_sum(f.partial + i, blocks, (int64_t)f.cp * 10, lane);
  if (lane != 0) return;
  const int c = i / 10, tap = i % 10;
  if (tap == 9) f.gb[c] = a;
  else f.gw[c * 9 + tap] = a;
}

template <typename T>
int fwd_t(const StripeGeom& g, const void* q, const void* k, const void* v, const float* lepe_w,
          const float* lepe_b, void* out, float* lse, cudaStream_t st) {
  const int tiles = (g.N + ROWS - 1) / ROWS;
  const int64_t grid = (int64_t)g.B * g.nwy * g.nwx * g.heads * tiles;
  if (grid > 0x7fffffffLL) return fail(CSB200_ERR_INVALID, "stripe_attn: grid too large");
  stripe_fwd_simt<T><<<(unsigned)grid, ROWS, 0, st>>>(
      g, static_cast<const T*>(q), static_cast<const T*>(k), static_cast<const T*>(v), lepe_w,
      lepe_b, static_cast<T*>(out), lse, tiles);
  return check_launch("stripe_fwd_simt");
}

template <typename T>
int bwd_t(const StripeGeom& g, const void* q, const void* k, const void* v, const float* lepe_w,
          const float* lepe_b, const void* out, const void* gout, const float* lse, void* dq,
          void* dk, void* dv, float* gw, float* gb, float* delta, float* partial,
          cudaStream_t st) {
  const int tiles = (g.N + ROWS - 1) / ROWS;
  const int64_t grid = (int64_t)g.B * g.nwy * g.nwx * g.heads * tiles;
  if (grid > 0x7fffffffLL) return fail(CSB200_ERR_INVALID, "stripe_attn: grid too large");
  const T* qt = static_cast<const T*>(q);
  const T* kt = static_cast<const T*>(k);
  const T* vt = static_cast<const T*>(v);
  const T* ot = static_cast<const T*>(out);
  const T* gt = static_cast<const T*>(gout);
  int rc;
  if ((rc = lepe_bwd_prep_t<T>(g, vt, lepe_w, lepe_b, ot, gt, delta, partial, gw, gb, st)) != CSB200_OK)
    return rc;
  stripe_bwd_dq_simt<T><<<(unsigned)grid, ROWS, 0, st>>>(g, qt, kt, vt, gt, lse, delta,
                                                         static_cast<T*>(dq), tiles);
  if ((rc = check_launch("stripe_bwd_dq_simt")) != CSB200_OK) return rc;
  stripe_bwd_dkv_simt<T><<<(unsigned)grid, ROWS, 0, st>>>(
      g, qt, kt, vt, lepe_w, gt, lse, delta, static_cast<T*>(dk), static_cast<T*>(dv), tiles);
  return check_launch("stripe_bwd_dkv_simt");
}

}  // namespace

// tokens per CTA of lepe_bwd_prep: ONE wave of CTAs (3 x 148, split over the heads), each walking a
// contiguous token range — the 40 accumulators are reduced once per CTA, and the final sum reads
// at most ~444 partials per output.
static int prep_tok_per_cta(const StripeGeom& g, int total_heads) {
  const int64_t total = (int64_t)g.B * g.L;
  const int ctas_per_head = (444 + total_heads - 1) / total_heads;
  int64_t tpc = (total + ctas_per_head - 1) / ctas_per_head;
  tpc = (tpc + PREP_TOK_PER_ITER - 1) / PREP_TOK_PER_ITER * PREP_TOK_PER_ITER;
  return (int)(tpc < PREP_TOK_PER_ITER ? PREP_TOK_PER_ITER : tpc);
}
// upper bound on the CTAs per head (sizes the partial buffer): a single-branch launch is the worst case
int wgrad_blocks(const StripeGeom& g) {
  const int tpc = prep_tok_per_cta(g, g.heads);
  const int generic = (int)(((int64_t)g.B * g.L + tpc - 1) / tpc);
  return generic > lepe_prep_tma_max_blocks() ? generic : lepe_prep_tma_max_blocks();
}

template <typename T>
int lepe_bwd_prep_multi_t(int nbr, const StripeGeom* g, const PrepIO* io, cudaStream_t st, int* final_blocks) {
  PrepArgs<T> a;
  memset(&a, 0, sizeof(a));
  int heads = 0, cp_max = 0;
  for (int i = 0; i < nbr; ++i) {
    a.br[i].g = g[i];
    a.br[i].v = static_cast<const T*>(io[i].v);
    a.br[i].out = static_cast<const T*>(io[i].out);
    a.br[i].gout = static_cast<const T*>(io[i].gout);
    a.br[i].lepe_w = io[i].lepe_w;
    a.br[i].lepe_b = io[i].lepe_b;
    a.br[i].delta = io[i].delta;
    a.br[i].partial = io[i].partial;
    heads += g[i].heads;
    cp_max = g[i].heads * HD > cp_max ? g[i].heads * HD : cp_max;
  }
  a.heads0 = g[0].heads;
  // TMA-streamed kernel (lepe_prep.cu) when the shape and alignment allow, else the generic one
  int blocks = 0;
  int rc = lepe_prep_tma_launch(nbr, g, std::is_same<T, float>::value ? CSB200_F32 : CSB200_BF16, io,
                                &blocks, st);
  if (rc == CSB200_ERR_UNSUPPORTED) {
    const int tpc = prep_tok_per_cta(g[0], heads);
    blocks = (int)(((int64_t)g[0].B * g[0].L + tpc - 1) / tpc);
    lepe_bwd_prep<T><<<dim3(blocks, heads), PREP_THREADS, 0, st>>>(a, tpc);
    rc = check_launch("lepe_bwd_prep");
  }
  if (rc != CSB200_OK) return rc;
  if (final_blocks != nullptr) {  // the caller's next kernel sums the partials in its prologue
    *final_blocks = blocks;
    return CSB200_OK;
  }
  WgradFinal f[2];
  for (int i = 0; i < 2; ++i) {
    const int j = i < nbr ? i : 0;
    f[i] = WgradFinal{io[j].partial, g[j].heads * HD, io[j].gw, io[j].gb};
  }
  lepe_wgrad_final<<<dim3((cp_max * 10 * 32 + 255) / 256, nbr), 256, 0, st>>>(f[0], f[1], blocks);
  return check_launch("lepe_wgrad_final");
}

int lepe_bwd_prep_multi(int nbr, const StripeGeom* g, int dtype, const PrepIO* io, cudaStream_t st,
                        int* final_blocks) {
  return dtype == CSB200_F32 ? lepe_bwd_prep_multi_t<float>(nbr, g, io, st, final_blocks)
                             : lepe_bwd_prep_multi_t<__nv_bfloat16>(nbr, g, io, st, final_blocks);
}

template <typename T>
int lepe_bwd_prep_t(const StripeGeom& g, const T* v, const float* lepe_w, const float* lepe_b,
                    const T* out, const T* gout, float* delta, float* partial, float* gw, float* gb,
                    cudaStream_t st) {
  const PrepIO io{v, out, gout, lepe_w, lepe_b, delta, partial, gw, gb};
  return lepe_bwd_prep_multi_t<T>(1, &g, &io, st, nullptr);
}

int lepe_bwd_prep(const StripeGeom& g, int dtype, const void* v, const float* lepe_w,
                  const float* lepe_b, const void* out, const void* gout, float* delta,
                  float* partial, float* gw, float* gb, cudaStream_t st) {
  const PrepIO io{v, out, gout, lepe_w, lepe_b, delta, partial, gw, gb};
  return lepe_bwd_prep_multi(1, &g, dtype, &io, st);
}

int simt_fwd(const StripeGeom& g, int dtype, const void* q, const void* k, const void* v,
             const float* lepe_w, const float* lepe_b, void* out, float* lse, cudaStream_t st) {
  return dtype == CSB200_F32 ? fwd_t<float>(g, q, k, v, lepe_w, lepe_b, out, lse, st)
                             : fwd_t<__nv_bfloat16>(g, q, k, v, lepe_w, lepe_b, out, lse, st);
}

int simt_bwd(const StripeGeom& g, int dtype, const void* q, const void* k, const void* v,
             const float* lepe_w, const float* lepe_b, const void* out, const void* gout,
             const float* lse, void* dq, void* dk, void* dv, float* gw, float* gb, float* delta,
             float* partial, cudaStream_t st) {
  return dtype == CSB200_F32
             ? bwd_t<float>(g, q, k, v, lepe_w, lepe_b, out, gout, lse, dq, dk, dv, gw, gb, delta,
                            partial, st)
             : bwd_t<__nv_bfloat16>(g, q, k, v, lepe_w, lepe_b, out, gout, lse, dq, dk, dv, gw,
                                    gb, delta, partial, st);
}

}  // namespace csb200

// Backward step 1 of stripe attention, TMA-streamed (both engines, any stripe shape, bf16 / fp32).
//
// ONE pass over grad_out, out and the 3x3 neighbourhood of v (get_v, C:244,256-269) yields
//   delta[b,h,l]  = sum_c grad_out * (out - lepe)          (row term of the softmax gradient)
//   gw[c][tap]    = sum_l grad_out[l][c] * v[nbr_tap(l)][c]   (depthwise weight gradient)
//   gb[c]         = sum_l grad_out[l][c]
// It is HBM-bound: 3 reads of (tokens x C') + 4 B per (token, head).  A persistent CTA owns one
// (branch, channel block) and walks (image, row block) tiles.  A producer warp fetches the three
// tiles of an item with TMA — v with a one-row halo above and below, out-of-image rows zero-filled
// by the copy engine — into a 2..3-stage shared-memory ring; 12 consumer warps read them back
// 4 channels per thread.  A thread keeps its channels for the whole kernel, so the 36 taps + bias
// and the 40 gradient accumulators live in registers and are reduced once at the end into
// partial[cta][C'][10]; lepe_wgrad_final adds the per-CTA partials in a fixed order (deterministic).
// The zero padding at the STRIPE border (get_lepe convolves each window separately, C:263-265) is a
// per-row / per-column bit mask built once per CTA.

#include <cstring>

#include "stripe_attn.cuh"
#include "tc_common.cuh"

namespace csb200 {
namespace {
using namespace tc;

constexpr int HD = 32;
constexpr int CONSUMERS = 384;  // 12 warps = 3 warpgroups
// + one warpgroup for the producer warp (its other three warps only hand their registers over): 512 threads start
// with 128 registers each, the producer warpgroup drops to 40 and the consumers grow to 152 (setmaxnreg) — with
// 128 the sliding-window walk (40 weights + 40 accumulators + a 36-register window per thread) spilled
constexpr int THREADS = CONSUMERS + 128;
constexpr int MAX_STAGES = 3;
constexpr int STAGE_BUDGET = 88 * 1024;  // 3 rows of 8 KB per tile: tokens divisible by the 3 * 2^k walkers
constexpr int SMEM_BUDGET = 200 * 1024;
constexpr int PART_BYTES = (CONSUMERS / 32) * 32 * 40 * (int)sizeof(float);  // final reduction scratch

struct PrepTBranch {
  int hs, ws, heads;
  int cb, ncb;              // channels per CTA, channel blocks (cb * ncb == C')
  int R, nrb;               // rows per tile, row blocks per image
  int v_bytes, t_bytes;     // v tile (R + 2 rows) / grad_out, out tile (R rows), 128-byte multiples
  int stages;
  const float *lepe_w, *lepe_b;
  float *delta, *partial;
};
struct PrepTParams {
  int B, H, W, L;
  int ncb0;                 // channel blocks of branch 0 (blockIdx.y below this -> branch 0)
  PrepTBranch br[2];
};
struct PrepTMaps {
  CUtensorMap v[2], g[2], o[2];
};

// 4 consecutive channels of one token from shared memory as two packed fp32 PAIRS (explicit 32-bit shared
// address: a generic pointer into the dynamic ring makes the compiler emit generic loads with 64-bit address
// math).  The whole per-token chain below runs on pairs (fma.rn.f32x2): the kernel is bound by instruction
// issue at the stage-3 / stage-4 shapes (ncu: issue-active 50 %, no DRAM pressure), and 72 scalar FMAs per
// (token, 4 channels) become 36.
struct P4 {
  f2_t a, b;  // channels (0, 1) and (2, 3)
};
template <typename T>
__device__ __forceinline__ P4 lds4(uint32_t addr);
template <>
__device__ __forceinline__ P4 lds4<float>(uint32_t addr) {
  P4 v;
  asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(v.a), "=l"(v.b) : "r"(addr));
  return v;
}
template <>
__device__ __forceinline__ P4 lds4<__nv_bfloat16>(uint32_t addr) {
  uint2 u;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(u.x), "=r"(u.y) : "r"(addr));
  return P4{f2_from_bf16x2(u.x), f2_from_bf16x2(u.y)};
}

struct Walk {        // per-thread constants of the token walk inside a tile
  int t0, dt;        // first token of this thread, tokens per step (== walkers)
  int x0, ry0, dx, dry;
  uint32_t tok_bytes, row_bytes;
};

// One token: lepe recomputation, weight-gradient accumulation, and the thread's share of delta.
// HAS_X / HAS_Y: the stripe is wider / taller than one token (else those taps never exist).
template <typename T, bool HAS_X, bool HAS_Y>
__device__ __forceinline__ float prep_token(uint32_t va, uint32_t ga, uint32_t oa, const Walk& wk, int mx,
                                            int my, const P4 (&w)[10], P4 (&acc)[10]) {
  const P4 go = lds4<T>(ga);
  const P4 o = lds4<T>(oa);
  P4 lp = w[9];
  acc[9].a = f2_add(acc[9].a, go.a);
  acc[9].b = f2_add(acc[9].b, go.b);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    if (!HAS_Y && ky != 1) continue;
    if (HAS_Y && ((ky == 0 && !(my & 1)) || (ky == 2 && !(my & 2)))) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      if (!HAS_X && kx != 1) continue;
      if (HAS_X && ((kx == 0 && !(mx & 1)) || (kx == 2 && !(mx & 2)))) continue;
      const P4 vn = lds4<T>(va + ky * wk.row_bytes + kx * wk.tok_bytes);  // va: tap (0,0)
      const P4 wt = w[ky * 3 + kx];
      lp.a = f2_fma(wt.a, vn.a, lp.a);
      lp.b = f2_fma(wt.b, vn.b, lp.b);
      P4& a = acc[ky * 3 + kx];
      a.a = f2_fma(go.a, vn.a, a.a);
      a.b = f2_fma(go.b, vn.b, a.b);
    }
  }
  // go . (o - lp)
  const f2_t m1 = f2_splat(-1.f);
  const f2_t d = f2_fma(go.b, f2_fma(lp.b, m1, o.b), f2_mul(go.a, f2_fma(lp.a, m1, o.a)));
  float d0, d1;
  f2_split(d, d0, d1);
  return d0 + d1;
}

__device__ __forceinline__ float head_sum(float d) {  // 8 adjacent lanes = the 32 channels of a head
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  d += __shfl_xor_sync(0xffffffffu, d, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 4);
  return d;
}

// All items of this CTA.  Token t of a tile sits at byte t * tok_bytes of the grad_out / out tiles and
// one row lower in the v tile (whose row 0 is the halo row y0 - 1); delta[y0 * W + t] is its output.
template <typename T, bool HAS_X, bool HAS_Y>
__device__ __forceinline__ void prep_items(const PrepTParams& p, const PrepTBranch& bg, uint32_t ring,
                                           uint64_t* full, uint64_t* empty, const uint8_t* s_my,
                                           const uint8_t* s_mx, const Walk& wk, int my_items, int head,
                                           bool writer, const P4 (&w)[10], P4 (&acc)[10]) {
  const int stage_bytes = bg.v_bytes + 2 * bg.t_bytes;
  const int tokens = bg.R * p.W;
  const int lane = threadIdx.x & 31;
  for (int i = 0; i < my_items; ++i) {
    const int item = (int)blockIdx.x + i * (int)gridDim.x;
    const int b = item / bg.nrb, y0 = (item - b * bg.nrb) * bg.R;
    const int s = i % bg.stages;
    const int rows = p.H - y0 < bg.R ? p.H - y0 : bg.R;
    const int tmax = rows * p.W;  // tokens of this tile that exist
    mbar_wait(&full[s], (i / bg.stages) & 1);
    // v address of tap (ky = 0, kx = 0) of token 0: tile row 0 is y0 - 1, one token to the left
    const uint32_t vbase = ring + (uint32_t)(s * stage_bytes) - wk.tok_bytes;
    const uint32_t gbase = ring + (uint32_t)(s * stage_bytes + bg.v_bytes);
    const uint32_t obase = gbase + (uint32_t)bg.t_bytes;
    float* drow = bg.delta + ((int64_t)b * bg.heads + head) * p.L + (int64_t)y0 * p.W;
    int x = wk.x0, ry = wk.ry0;
    const int iters = (tokens + 2 * wk.dt - 1) / (2 * wk.dt);  // same trip count for every lane
    for (int k = 0, t = wk.t0; k < iters; ++k, t += 2 * wk.dt) {
      // two tokens per iteration: two independent shuffle / FMA chains in flight
      const int ta = t, tb = t + wk.dt;
      int mxa = 0, mya = 0, mxb = 0, myb = 0;
      if (HAS_X) mxa = s_mx[x];
      if (HAS_Y) mya = s_my[min(y0 + ry, 511)];
      x += wk.dx; ry += wk.dry;
      if (x >= p.W) { x -= p.W; ++ry; }
      const bool vb_ = tb < tmax;
      if (HAS_X) mxb = s_mx[x];
      if (HAS_Y) myb = s_my[min(y0 + ry, 511)];
      x += wk.dx; ry += wk.dry;
      if (x >= p.W) { x -= p.W; ++ry; }
      float da = 0.f, db = 0.f;
      if (ta < tmax) {
        const uint32_t off = (uint32_t)ta * wk.tok_bytes;
        da = prep_token<T, HAS_X, HAS_Y>(vbase + off, gbase + off, obase + off, wk, mxa, mya, w, acc);
      }
      if (vb_) {
        const uint32_t off = (uint32_t)tb * wk.tok_bytes;
        db = prep_token<T, HAS_X, HAS_Y>(vbase + off, gbase + off, obase + off, wk, mxb, myb, w, acc);
      }
      da = head_sum(da);
      db = head_sum(db);
      if (writer && ta < tmax) drow[ta] = da;
      if (writer && vb_) drow[tb] = db;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
}

// Sliding-window walk (the config-3 tile shapes: R * W * (cb / 4) == 8 * CONSUMERS, stripe width 1 or a multiple
// of 8): a thread owns 4 channels of a RUN of 8 consecutive tokens of one tile row and keeps the three v columns
// (x - 1, x, x + 1) x three rows of the 3x3 window in registers, so a token costs ONE new column (3 shared-memory
// loads) instead of 9, and there is no per-tap branching: a neighbour outside the stripe (or the image, or the rows
// a partial tile lacks) is read from a block of zeros — zero padding, C:244,263-265 — by selecting the ADDRESS.
// The generic walk above executed ~300 instructions per (token, 4 channels) at the stage-3 shape (ncu: issue-bound,
// 35 us for 50 MB); this one ~85.
constexpr int SEG = 8;
template <typename T, bool HAS_X, bool HAS_Y>
__device__ __forceinline__ void prep_items_slide(const PrepTParams& p, const PrepTBranch& bg, uint32_t ring,
                                                 uint32_t zaddr, uint64_t* full, uint64_t* empty, int cg, int walker,
                                                 int my_items, int head, bool writer, const P4 (&w)[10],
                                                 P4 (&acc)[10]) {
  constexpr int NX = HAS_X ? 3 : 1, NY = HAS_Y ? 3 : 1;
  const int stage_bytes = bg.v_bytes + 2 * bg.t_bytes;
  const int lane = threadIdx.x & 31;
  const int segs = p.W / SEG;
  const int r = walker / segs, xs = (walker - r * segs) * SEG;  // tile row and first token of this thread's run
  const uint32_t tok_bytes = (uint32_t)(bg.cb * (int)sizeof(T)), row_bytes = tok_bytes * (uint32_t)p.W;
  const bool left_ok = HAS_X && (xs % bg.ws) != 0, right_ok = HAS_X && ((xs + SEG) % bg.ws) != 0;
  const f2_t m1 = f2_splat(-1.f);
  for (int i = 0; i < my_items; ++i) {
    const int item = (int)blockIdx.x + i * (int)gridDim.x;
    const int b = item / bg.nrb, y0 = (item - b * bg.nrb) * bg.R;
    const int s = i % bg.stages;
    const int rows = p.H - y0 < bg.R ? p.H - y0 : bg.R;
    const bool row_in = r < rows;
    const int yy = (y0 + r) % bg.hs;
    // which of the three window rows exist for this thread's tile row (ky = 1 is the row itself)
    const bool rok[3] = {HAS_Y && row_in && yy != 0, row_in, HAS_Y && row_in && yy != bg.hs - 1};
    mbar_wait(&full[s], (i / bg.stages) & 1);
    // v tile row 0 is image row y0 - 1: window row ky of tile row r is v tile row r + ky
    const uint32_t vrow = ring + (uint32_t)(s * stage_bytes) + (uint32_t)r * row_bytes + (uint32_t)xs * tok_bytes;
    const uint32_t grow = ring + (uint32_t)(s * stage_bytes + bg.v_bytes) + (uint32_t)r * row_bytes + (uint32_t)xs * tok_bytes;
    const uint32_t orow = grow + (uint32_t)bg.t_bytes;
    float* drow = bg.delta + ((int64_t)b * bg.heads + head) * p.L + (int64_t)(y0 + r) * p.W + xs;
    P4 col[NY][3];  // [window row][slot]; slot (j + kx) % 3 holds column xs + j + kx - 1 while token j is computed
    auto load_col = [&](int slot, int dxs, bool ok) {  // column xs + dxs of the window rows
#pragma unroll
      for (int ky = 0; ky < NY; ++ky) {
        const int kr = HAS_Y ? ky : 1;
        const bool use = ok && rok[kr];
        col[ky][slot] = lds4<T>(use ? vrow + (uint32_t)kr * row_bytes + (uint32_t)(dxs * (int)tok_bytes) : zaddr);
      }
    };
    if (HAS_X) {
      load_col(0, -1, left_ok);
      load_col(1, 0, true);
    }
#pragma unroll
    for (int j = 0; j < SEG; ++j) {
      if (HAS_X) load_col((j + 2) % 3, j + 1, j + 1 < SEG ? true : right_ok);
      else load_col(0, j, true);
      const P4 go = lds4<T>(row_in ? grow + (uint32_t)j * tok_bytes : zaddr);
      const P4 o = lds4<T>(row_in ? orow + (uint32_t)j * tok_bytes : zaddr);
      P4 lp = w[9];
      acc[9].a = f2_add(acc[9].a, go.a);
      acc[9].b = f2_add(acc[9].b, go.b);
#pragma unroll
      for (int ky = 0; ky < NY; ++ky) {
#pragma unroll
        for (int kx = 0; kx < NX; ++kx) {
          const P4 vn = col[ky][HAS_X ? (j + kx) % 3 : 0];
          const int tap = (HAS_Y ? ky : 1) * 3 + (HAS_X ? kx : 1);
          const P4 wt = w[tap];
          lp.a = f2_fma(wt.a, vn.a, lp.a);
          lp.b = f2_fma(wt.b, vn.b, lp.b);
          P4& a = acc[tap];
          a.a = f2_fma(go.a, vn.a, a.a);
          a.b = f2_fma(go.b, vn.b, a.b);
        }
      }
      // go . (o - lp), summed over the 32 channels of the head (8 adjacent lanes)
      const f2_t d = f2_fma(go.b, f2_fma(lp.b, m1, o.b), f2_mul(go.a, f2_fma(lp.a, m1, o.a)));
      float d0, d1;
      f2_split(d, d0, d1);
      const float dsum = head_sum(d0 + d1);
      if (writer && row_in) drow[j] = dsum;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
}

template <typename T, bool SLIDE>
__global__ void __launch_bounds__(THREADS, 1)
    lepe_prep_tma(const __grid_constant__ PrepTMaps maps, const __grid_constant__ PrepTParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  __shared__ uint64_t full[MAX_STAGES], empty[MAX_STAGES];
  __shared__ uint8_t s_my[256 + 256], s_mx[256];  // s_my: slack for the rows a partial tile lacks
  __shared__ __align__(16) uint8_t s_zero[16];    // what a neighbour outside the stripe reads as (sliding-window walk)

  const int which = (int)blockIdx.y >= p.ncb0 ? 1 : 0;
  const PrepTBranch& bg = p.br[which];
  const int c0 = ((int)blockIdx.y - (which ? p.ncb0 : 0)) * bg.cb;  // first channel of this CTA
  const int items = p.B * bg.nrb;
  const int my_items = (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int stage_bytes = bg.v_bytes + 2 * bg.t_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // bit 0: the neighbour above / to the left is inside the stripe; bit 1: below / to the right
  for (int i = threadIdx.x; i < 512; i += THREADS) {
    const int yy = i % bg.hs;
    s_my[i] = i < p.H ? (uint8_t)((yy > 0 ? 1 : 0) | (yy < bg.hs - 1 ? 2 : 0)) : (uint8_t)0;
  }
  for (int i = threadIdx.x; i < p.W; i += THREADS) {
    const int xx = i % bg.ws;
    s_mx[i] = (uint8_t)((xx > 0 ? 1 : 0) | (xx < bg.ws - 1 ? 2 : 0));
  }
  if (threadIdx.x < 16) s_zero[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], CONSUMERS / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp >= CONSUMERS / 32) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    // ===================================== TMA producer =====================================
    if (warp == CONSUMERS / 32 && lane == 0) {
      prefetch_tensormap(&maps.v[which]);
      prefetch_tensormap(&maps.g[which]);
      prefetch_tensormap(&maps.o[which]);
      for (int i = 0; i < my_items; ++i) {
        const int item = (int)blockIdx.x + i * (int)gridDim.x;
        const int b = item / bg.nrb, y0 = (item - b * bg.nrb) * bg.R;
        const int s = i % bg.stages;
        mbar_wait(&empty[s], ((i / bg.stages) & 1) ^ 1);
        uint8_t* st = ring + (size_t)s * stage_bytes;
        mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
        tma_load_4d(st, &maps.v[which], &full[s], c0, 0, y0 - 1, b);
        tma_load_4d(st + bg.v_bytes, &maps.g[which], &full[s], c0, 0, y0, b);
        tma_load_4d(st + bg.v_bytes + bg.t_bytes, &maps.o[which], &full[s], c0, 0, y0, b);
      }
    }
    return;
  }

  // ======================================= consumers ========================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
  constexpr int ES = (int)sizeof(T);
  const int cgn = bg.cb >> 2;                    // channel groups (of 4) per token: 8, 16, 32 or 64
  const int cg = (int)threadIdx.x % cgn;
  const int walker = (int)threadIdx.x / cgn, walkers = CONSUMERS / cgn;
  const int ch = c0 + cg * 4, head = ch / HD;    // channel inside the branch

  P4 w[10];  // [tap] for this thread's 4 channels, bias last
  {
    float t[4][10];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
#pragma unroll
      for (int k = 0; k < 9; ++k) t[e][k] = __ldg(bg.lepe_w + (ch + e) * 9 + k);
      t[e][9] = __ldg(bg.lepe_b + ch + e);
    }
#pragma unroll
    for (int k = 0; k < 10; ++k) w[k] = P4{f2_make(t[0][k], t[1][k]), f2_make(t[2][k], t[3][k])};
  }
  P4 accp[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) accp[k] = P4{f2_splat(0.f), f2_splat(0.f)};

  const bool writer = (cg & 7) == 0;
  const bool has_x = bg.ws > 1, has_y = bg.hs > 1;
  if constexpr (SLIDE) {  // the host has checked slide_ok() for every branch of the launch
    const uint32_t ring_c = smem_u32(ring) + (uint32_t)(cg * 4 * ES), zaddr = smem_u32(s_zero);
    if (has_x && has_y)
      prep_items_slide<T, true, true>(p, bg, ring_c, zaddr, full, empty, cg, walker, my_items, head, writer, w, accp);
    else if (has_y)
      prep_items_slide<T, false, true>(p, bg, ring_c, zaddr, full, empty, cg, walker, my_items, head, writer, w, accp);
    else if (has_x)
      prep_items_slide<T, true, false>(p, bg, ring_c, zaddr, full, empty, cg, walker, my_items, head, writer, w, accp);
    else
      prep_items_slide<T, false, false>(p, bg, ring_c, zaddr, full, empty, cg, walker, my_items, head, writer, w, accp);
  } else {
    Walk wk;
    wk.t0 = walker; wk.dt = walkers;
    wk.ry0 = walker / p.W; wk.x0 = walker - wk.ry0 * p.W;
    wk.dry = walkers / p.W; wk.dx = walkers - wk.dry * p.W;
    wk.tok_bytes = (uint32_t)(bg.cb * ES);
    wk.row_bytes = wk.tok_bytes * (uint32_t)p.W;
    const uint32_t ring_a = smem_u32(ring) + (uint32_t)(cg * 4 * ES);
    if (has_x && has_y)
      prep_items<T, true, true>(p, bg, ring_a, full, empty, s_my, s_mx, wk, my_items, head, writer, w, accp);
    else if (has_y)
      prep_items<T, false, true>(p, bg, ring_a, full, empty, s_my, s_mx, wk, my_items, head, writer, w, accp);
    else if (has_x)
      prep_items<T, true, false>(p, bg, ring_a, full, empty, s_my, s_mx, wk, my_items, head, writer, w, accp);
    else
      prep_items<T, false, false>(p, bg, ring_a, full, empty, s_my, s_mx, wk, my_items, head, writer, w, accp);
  }

  float4 acc[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    f2_split(accp[k].a, acc[k].x, acc[k].y);
    f2_split(accp[k].b, acc[k].z, acc[k].w);
  }
  // ---- reduce the 40 accumulators over the walkers of this CTA (fixed order) ----
  for (int off = cgn; off < 32; off <<= 1) {
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      acc[k].x += __shfl_xor_sync(0xffffffffu, acc[k].x, off);
      acc[k].y += __shfl_xor_sync(0xffffffffu, acc[k].y, off);
      acc[k].z += __shfl_xor_sync(0xffffffffu, acc[k].z, off);
      acc[k].w += __shfl_xor_sync(0xffffffffu, acc[k].w, off);
    }
  }
  // every stage has been consumed (all TMA writes landed): reuse the ring as scratch
  asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");
  float* part = reinterpret_cast<float*>(ring);  // [warp][lane][10][4]
  const int held = cgn < 32 ? cgn : 32;          // channel groups a warp holds (lanes 0..held-1)
  if (lane < held) {
#pragma unroll
    for (int k = 0; k < 10; ++k)
      *reinterpret_cast<float4*>(part + ((warp * 32 + lane) * 10 + k) * 4) = acc[k];
  }
  asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");
  const int cp = bg.heads * HD;
  const int wstep = cgn <= 32 ? 1 : cgn / 32;    // warps wstep apart hold the same channel groups
  for (int idx = threadIdx.x; idx < cgn * 40; idx += CONSUMERS) {
    const int cgi = idx / 40, r = idx - cgi * 40, k = r >> 2, e = r & 3;
    const int w0 = cgn <= 32 ? 0 : cgi / 32, l = cgn <= 32 ? cgi : cgi & 31;
    float a = 0.f;
    for (int wv = w0; wv < CONSUMERS / 32; wv += wstep) a += part[((wv * 32 + l) * 10 + k) * 4 + e];
    bg.partial[((int64_t)blockIdx.x * cp + c0 + cgi * 4 + e) * 10 + k] = a;
  }
}

int make_tok_map(CUtensorMap* m, const void* base, int dtype, int cp, int W, int H, int B, int64_t sb,
                 int64_t sl, int box_c, int box_y) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return fail(CSB200_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
  ensure_context();
  const cuuint64_t es = dtype == CSB200_F32 ? 4 : 2;
  const cuuint64_t dims[4] = {(cuuint64_t)cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)sl * es, (cuuint64_t)sl * es * W, (cuuint64_t)sb * es};
  const cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)W, (cuuint32_t)box_y, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, dtype == CSB200_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                   4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CSB200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return CSB200_OK;
}

// tile plan of one branch; false if the shape does not fit the TMA path
bool plan_branch(const StripeGeom& g, int dtype, const PrepIO& io, PrepTBranch* b) {
  const int es = dtype == CSB200_F32 ? 4 : 2;
  const int cp = g.heads * HD;
  if (g.W > 256 || g.H > 256 || g.B < 1) return false;
  const uintptr_t al = reinterpret_cast<uintptr_t>(io.v) | reinterpret_cast<uintptr_t>(io.out) |
                       reinterpret_cast<uintptr_t>(io.gout);
  if (al & 15) return false;
  const int64_t strides[4] = {g.v_sb, g.v_sl, g.o_sb, g.o_sl};
  for (int64_t s : strides)
    if (s <= 0 || (s * es) % 16 != 0) return false;
  // the token grid of a map is (W, H) with row stride sl * W: fine for any sb
  int cb = 256;
  while (cb > 32 && (cp % cb != 0 || 5LL * g.W * cb * es > SMEM_BUDGET / 2)) cb >>= 1;
  const int64_t row = (int64_t)g.W * cb * es;
  if (cp % cb != 0 || 5 * row > SMEM_BUDGET / 2) return false;
  int R = (int)((STAGE_BUDGET / row - 2) / 3);
  if (R < 1) R = 1;
  if (R > g.H) R = g.H;
  if (R > 254) R = 254;
  b->hs = g.hs; b->ws = g.ws; b->heads = g.heads;
  b->cb = cb; b->ncb = cp / cb;
  b->R = R; b->nrb = (g.H + R - 1) / R;
  b->v_bytes = (int)(((R + 2) * row + 127) / 128 * 128);
  b->t_bytes = (int)((R * row + 127) / 128 * 128);
  const int stage = b->v_bytes + 2 * b->t_bytes;
  b->stages = SMEM_BUDGET / stage < MAX_STAGES ? SMEM_BUDGET / stage : MAX_STAGES;
  if (b->stages < 2) return false;
  b->lepe_w = io.lepe_w; b->lepe_b = io.lepe_b;
  b->delta = io.delta; b->partial = io.partial;
  return true;
}

}  // namespace

// Upper bound on blockIdx.x of the TMA path (sizes the partial buffer together with the generic one).
int lepe_prep_tma_max_blocks() { return 160; }

// Returns CSB200_OK and sets *blocks (CTAs per channel block == partials per output) when the TMA
// path ran; CSB200_ERR_UNSUPPORTED (without touching the error text) when the caller should use the
// generic kernel.
int lepe_prep_tma_launch(int nbr, const StripeGeom* g, int dtype, const PrepIO* io, int* blocks,
                         cudaStream_t st) {
  PrepTMaps maps;
  PrepTParams p;
  memset(&maps, 0, sizeof(maps));
  memset(&p, 0, sizeof(p));
  p.B = g[0].B; p.H = g[0].H; p.W = g[0].W; p.L = g[0].L;
  int ncb = 0, smem = PART_BYTES, items = 0;
  for (int i = 0; i < nbr; ++i) {
    if (g[i].B != p.B || g[i].H != p.H || g[i].W != p.W) return CSB200_ERR_UNSUPPORTED;
    if (!plan_branch(g[i], dtype, io[i], &p.br[i])) return CSB200_ERR_UNSUPPORTED;
    const PrepTBranch& b = p.br[i];
    const int need = b.stages * (b.v_bytes + 2 * b.t_bytes);
    smem = need > smem ? need : smem;
    items = i == 0 || p.B * b.nrb < items ? p.B * b.nrb : items;
    ncb += b.ncb;
  }
  p.ncb0 = p.br[0].ncb;
  const int sm_count = device_sm_count();
  if (sm_count <= 0) return fail(CSB200_ERR_CUDA, "lepe_prep_tma: cannot query the SM count");
  int gx = sm_count / ncb;
  if (gx < 1) gx = 1;
  if (gx > items) gx = items;
  if (gx > lepe_prep_tma_max_blocks()) gx = lepe_prep_tma_max_blocks();
  for (int i = 0; i < nbr; ++i) {
    const int cp = g[i].heads * HD;
    const PrepTBranch& b = p.br[i];
    int rc;
    if ((rc = make_tok_map(&maps.v[i], io[i].v, dtype, cp, p.W, p.H, p.B, g[i].v_sb, g[i].v_sl, b.cb, b.R + 2)) != CSB200_OK) return rc;
    if ((rc = make_tok_map(&maps.g[i], io[i].gout, dtype, cp, p.W, p.H, p.B, g[i].o_sb, g[i].o_sl, b.cb, b.R)) != CSB200_OK) return rc;
    if ((rc = make_tok_map(&maps.o[i], io[i].out, dtype, cp, p.W, p.H, p.B, g[i].o_sb, g[i].o_sl, b.cb, b.R)) != CSB200_OK) return rc;
  }
  smem += 128;
  // sliding-window walk: every branch's tile is 8 tokens per consumer thread and stripe widths are 1 or 8 k
  bool slide = true;
  for (int i = 0; i < nbr; ++i) {
    const PrepTBranch& b = p.br[i];
    slide = slide && b.R * p.W * (b.cb / 4) == CONSUMERS * SEG && p.W % SEG == 0 && (b.ws == 1 || b.ws % SEG == 0);
  }
  const dim3 grid(gx, ncb);
#define CSB_PREP_LAUNCH(TT, SL)                                                                              \
  do {                                                                                                       \
    CSB200_CUDA(opt_in_smem(reinterpret_cast<const void*>(&lepe_prep_tma<TT, SL>), SMEM_BUDGET + 128));      \
    lepe_prep_tma<TT, SL><<<grid, THREADS, smem, st>>>(maps, p);                                             \
  } while (0)
  if (dtype == CSB200_F32) {
    if (slide) CSB_PREP_LAUNCH(float, true);
    else CSB_PREP_LAUNCH(float, false);
  } else {
    if (slide) CSB_PREP_LAUNCH(__nv_bfloat16, true);
    else CSB_PREP_LAUNCH(__nv_bfloat16, false);
  }
#undef CSB_PREP_LAUNCH
  *blocks = gx;
  return check_launch("lepe_prep_tma");
}

}  // namespace csb200

// Stripe attention + LePE, tcgen05 / TMEM / TMA engine — forward for LONG stripes (N = 128 T, T >= 3): the
// high-resolution shapes of BASELINE config 5 (1024^2 inference: N = 512, 1024, 2048 at stripe width 8, the
// 32 x 32 full window of its last stage) that the single-pass kernels of stripe_attn_tc.cu (one S row of N
// columns in TMEM) cannot hold.
//
// A work item is one 128-row QUERY TILE of a (image, stripe, head) group; its T key/value blocks of 128 tokens
// stream through shared memory and the softmax is the online form (running row max m, running sum l, the
// 32-column fp32 output accumulator in TMEM rescaled by exp(m_old - m_new) before each P V):
//
//   warp 0      TMA producer: Q once per item, K and V once per (item, block), in the issue order below (even entries)
//   warp 1      tcgen05.mma issuer   S = Q K_j^T   (M128 x N128 x K32)
//   warp 2      TMEM allocator, then the second TMA producer (odd entries)
//   warp 3      tcgen05.mma issuer   O (+)= P_j V_j (M128 x N32 x K128, P read from TMEM)
//   warps 4..15 three softmax warpgroups; warpgroup w owns TMEM columns [160 w, 160 w + 160): S / P in
//               [0, 128), O in [128, 160), and takes every third item; the two issuing warps serve the
//               warpgroups round-robin, one block each per round, so three items are always in flight
//
// Epilogue: O / l plus the LePE depthwise 3x3 of V (zero padding at the stripe border, C:244,263-265), read from
// global memory here (the V blocks of a long stripe are not resident; 9 x 64 B per row out of L2), and
// lse = m scale + log l for a backward pass (which, for these lengths, runs on the CUDA-core engine).
// The partition copies of the reference (C:199-217, C:248-254) are the TMA address generation, as in the
// single-pass kernels.

#include <cstring>

#include "stripe_attn.cuh"
#include "tc_common.cuh"

namespace csb200 {
namespace {
using namespace tc;

constexpr int HD = 32;
#ifndef CSB_POLY_MOD
#define CSB_POLY_MOD 2
#endif
constexpr int POLY_MOD = CSB_POLY_MOD;        // every POLY_MOD-th pair of exponentials on the FMA pipes (0: none)
constexpr int TILE = 128;
constexpr int ROW_BYTES = HD * 2;
constexpr int TILE_BYTES = TILE * ROW_BYTES;  // 8 KB
constexpr int NWG = 3;
constexpr int QS = 4, KS = 8, VS = 8;         // ring depths (tiles of 8 KB)
constexpr uint32_t BUF_COLS = 160, O_COL = 128;
constexpr int THREADS = 128 + 128 * NWG;

struct KvParams {
  int B, W, L;
  int hs, ws, nwy, nwx, heads;
  int by;                   // ws <= 128: stripe rows per tile (floor(128 / ws)); the TMA box is (32, ws, by)
  int tpr;                  // ws > 128: tiles per stripe row (ceil(ws / 128)); the TMA box is (32, 128, 1)
  int box_bytes;            // bytes one box delivers (tile rows the box does not cover are never written)
  int T;                    // query tiles == key/value blocks per group
  int items;                // groups * T
  float scale, scale_log2;
  const float* lepe_w;      // [C'][9]
  const float* lepe_b;      // [C']
  const __nv_bfloat16* v;   // for the LePE stencil
  int64_t v_sb, v_sl;
  __nv_bfloat16* out;
  int64_t o_sb, o_sl;
  float* lse;
};
struct KvMaps {
  CUtensorMap q, k, v;
};
struct KvSmem {
  alignas(1024) uint8_t q[QS][TILE_BYTES];
  alignas(1024) uint8_t k[KS][TILE_BYTES];
  alignas(1024) uint8_t v[VS][TILE_BYTES];
  alignas(8) uint64_t q_full[QS], q_empty[QS], k_full[KS], k_empty[KS], v_full[VS], v_empty[VS];
  uint64_t s_full[NWG], p_full[NWG], o_full[NWG], buf_empty[NWG];
  uint32_t tmem_base;
};

struct Item {
  int b, wy, wx, head, qt;
};
__device__ __forceinline__ Item decode_item(const KvParams& p, int it) {
  Item c;
  c.qt = it % p.T;
  int g = it / p.T;
  c.head = g % p.heads;
  g /= p.heads;
  c.wx = g % p.nwx;
  g /= p.nwx;
  c.wy = g % p.nwy;
  c.b = g / p.nwy;
  return c;
}
// Tile t of a stripe (h_sp x w_sp tokens, row-major): w_sp <= 128 -> `by` whole stripe rows (by * w_sp <= 128 token
// rows of the tile are used), w_sp > 128 -> 128 consecutive tokens of one stripe row.  Tokens of the box that lie
// outside the stripe (below it, or to its right) are loaded like any others — or zero-filled outside the image —
// and MASKED: tile rows [vcount, 128) never take part (keys: S columns set to -inf; queries: not stored).  Stripe
// shapes that are not powers of two (w_sp = 7: BASELINE config 5 at 896^2, config 1) therefore tile like the rest.
__device__ __forceinline__ void tile_xy(const KvParams& p, const Item& c, int t, int& x, int& y) {
  x = c.wx * p.ws + (p.ws > TILE ? (t % p.tpr) * TILE : 0);
  y = c.wy * p.hs + (p.ws > TILE ? t / p.tpr : t * p.by);
}
__device__ __forceinline__ int tile_vcount(const KvParams& p, int t) {  // valid token rows of tile t (a prefix)
  if (p.ws > TILE) {
    const int left = p.ws - (t % p.tpr) * TILE;
    return left < TILE ? left : TILE;
  }
  const int rows = p.hs - t * p.by;
  return (rows < p.by ? rows : p.by) * p.ws;
}
// in-stripe (row, column) of token row r of tile t
__device__ __forceinline__ void tile_pos(const KvParams& p, int t, int r, int& yy, int& xx) {
  if (p.ws > TILE) {
    yy = t / p.tpr;
    xx = (t % p.tpr) * TILE + r;
  } else {
    const int q = r / p.ws;
    yy = t * p.by + q;
    xx = r - q * p.ws;
  }
}
// keys [32 ch, 32 ch + 32) of a block with `vk` valid keys: the others become -inf
__device__ __forceinline__ void mask_keys(uint32_t (&r)[32], int ch, int vk) {
  const int left = vk - 32 * ch;
  if (left >= 32) return;
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = i < left ? r[i] : 0xff800000u;
}

__global__ void __launch_bounds__(THREADS, 1)
    stripe_fwd_tc_kv(const __grid_constant__ KvMaps maps, const __grid_constant__ KvParams p) {
  extern __shared__ uint8_t smem_raw[];
  KvSmem& sm = *reinterpret_cast<KvSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_items = (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int T = p.T;
  // Issue order shared by the producer and the two issuing warps: rounds r = 0, 1, ...; in round r warpgroup w
  // works on block (r % T) of its item number (r / T), i.e. the CTA's local item w + 3 (r / T).
  const int rounds = ((my_items + NWG - 1) / NWG) * T;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&maps.q);
    prefetch_tensormap(&maps.k);
    prefetch_tensormap(&maps.v);
    for (int i = 0; i < QS; ++i) {
      mbar_init(&sm.q_full[i], 1);
      mbar_init(&sm.q_empty[i], 1);
    }
    for (int i = 0; i < KS; ++i) {
      mbar_init(&sm.k_full[i], 1);
      mbar_init(&sm.k_empty[i], 1);
    }
    for (int i = 0; i < VS; ++i) {
      mbar_init(&sm.v_full[i], 1);
      mbar_init(&sm.v_empty[i], 1);
    }
    for (int i = 0; i < NWG; ++i) {
      mbar_init(&sm.s_full[i], 1);
      mbar_init(&sm.p_full[i], 128);
      mbar_init(&sm.o_full[i], 1);
      mbar_init(&sm.buf_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&sm.tmem_base, 512);
  if (p.box_bytes < TILE_BYTES) {
    // tile rows the boxes never write must read as ZERO in V (their probabilities are zero, but 0 x NaN is NaN)
    uint4* vz = reinterpret_cast<uint4*>(&sm.v[0][0]);
    for (int i = threadIdx.x; i < VS * TILE_BYTES / 16; i += THREADS) vz[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0 || warp == 2) {
    // ================================ TMA producers (two warps) ================================
    // warp 0 takes the even (item, block) entries, warp 2 the odd ones: one thread moves an 8-KB box of 64-byte
    // rows every ~800 cycles whatever the ring depth (benchmarks/debug/tma_rows.cu); the K and V boxes of an entry
    // stay together on one warp (k | v of a token share a DRAM page)
    if (lane == 0) {
      const int mine = warp >> 1;
      int e = 0, qn = 0;  // entries (K / V blocks) and Q tiles so far (both warps count all of them)
      for (int r = 0; r < rounds; ++r) {
        const int j = r % T;
        for (int w = 0; w < NWG; ++w) {
          const int li = w + NWG * (r / T);
          if (li >= my_items) continue;
          if ((e & 1) == mine) {
            const Item c = decode_item(p, (int)blockIdx.x + li * (int)gridDim.x);
            int x, y;
            if (j == 0) {
              const int qs = qn % QS;
              mbar_wait(&sm.q_empty[qs], ((qn / QS) & 1) ^ 1);
              mbar_expect_tx(&sm.q_full[qs], p.box_bytes);
              tile_xy(p, c, c.qt, x, y);
              tma_load_4d(sm.q[qs], &maps.q, &sm.q_full[qs], c.head * HD, x, y, c.b);
            }
            const int ks = e % KS, vs = e % VS;
            tile_xy(p, c, j, x, y);
            mbar_wait(&sm.k_empty[ks], ((e / KS) & 1) ^ 1);
            mbar_expect_tx(&sm.k_full[ks], p.box_bytes);
            tma_load_4d(sm.k[ks], &maps.k, &sm.k_full[ks], c.head * HD, x, y, c.b);
            mbar_wait(&sm.v_empty[vs], ((e / VS) & 1) ^ 1);
            mbar_expect_tx(&sm.v_full[vs], p.box_bytes);
            tma_load_4d(sm.v[vs], &maps.v, &sm.v_full[vs], c.head * HD, x, y, c.b);
          }
          if (j == 0) ++qn;
          ++e;
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer 1: S = Q K_j^T ================================
    constexpr uint32_t idesc_s = umma_idesc_bf16(TILE, false, false);
    const uint32_t q_lo0 = desc_lo_sw64(smem_u32(sm.q[0])), k_lo0 = desc_lo_sw64(smem_u32(sm.k[0]));
    int e = 0, qn = 0;
    int q_of_wg[NWG] = {0, 0, 0};   // Q stage of each warpgroup's current item
    int blk[NWG] = {0, 0, 0};       // blocks issued per warpgroup (all items)
    for (int r = 0; r < rounds; ++r) {
      const int j = r % T;
#pragma unroll
      for (int w = 0; w < NWG; ++w) {
        const int li = w + NWG * (r / T);
        if (li >= my_items) continue;
        if (j == 0) {
          q_of_wg[w] = qn % QS;
          mbar_wait(&sm.q_full[q_of_wg[w]], (qn / QS) & 1);
          ++qn;
        }
        const int ks = e % KS;
        mbar_wait(&sm.k_full[ks], (e / KS) & 1);
        // the S / P columns of this warpgroup are free again once its previous P V has completed (it read P) —
        // and, for the first block of an item, once the epilogue has drained O
        if (blk[w] > 0) {
          if (j == 0) mbar_wait(&sm.buf_empty[w], ((blk[w] / T - 1) & 1));
          else mbar_wait(&sm.o_full[w], (blk[w] - 1) & 1);
        }
        fence_after_sync();
        if (elect_one_sync()) {
          const uint32_t q_lo = q_lo0 + q_of_wg[w] * (TILE_BYTES >> 4), k_lo = k_lo0 + ks * (TILE_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_ss2(tmem + w * BUF_COLS, q_lo + k * (32 >> 4), DESC_HI_SW64, k_lo + k * (32 >> 4), DESC_HI_SW64,
                     idesc_s, k > 0);
          umma_commit(&sm.s_full[w]);
          umma_commit(&sm.k_empty[ks]);
          if (j == T - 1) umma_commit(&sm.q_empty[q_of_wg[w]]);
        }
        __syncwarp();
        ++blk[w];
        ++e;
      }
    }
  } else if (warp == 3) {
    // ================================ MMA issuer 2: O (+)= P_j V_j ==============================
    constexpr uint32_t idesc_pv = umma_idesc_bf16(HD, false, true);
    const uint32_t v_lo0 = desc_lo_sw64(smem_u32(sm.v[0]));
    int e = 0;
    int blk[NWG] = {0, 0, 0};
    for (int r = 0; r < rounds; ++r) {
      const int j = r % T;
#pragma unroll
      for (int w = 0; w < NWG; ++w) {
        const int li = w + NWG * (r / T);
        if (li >= my_items) continue;
        const int vs = e % VS;
        mbar_wait(&sm.v_full[vs], (e / VS) & 1);
        mbar_wait(&sm.p_full[w], blk[w] & 1);
        fence_after_sync();
        if (elect_one_sync()) {
          const uint32_t d = tmem + w * BUF_COLS + O_COL, a = tmem + w * BUF_COLS;
          const uint32_t v_lo = v_lo0 + vs * (TILE_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < TILE / 16; ++k)  // 16 keys per step: 8 TMEM columns of P, 1024 B of V
            umma_ts2(d, a + 8 * k, v_lo + k * (1024 >> 4), DESC_HI_SW64, idesc_pv, (j > 0 || k > 0));
          umma_commit(&sm.o_full[w]);
          umma_commit(&sm.v_empty[vs]);
        }
        __syncwarp();
        ++blk[w];
        ++e;
      }
    }
  } else if (warp >= 4) {
    // ============================ softmax + epilogue warpgroups ============================
    const int wg = (warp - 4) >> 2;
    const int row = ((warp & 3) << 5) | lane;
    const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) << 5) << 16) + wg * BUF_COLS;
    int nblk = 0;  // blocks of this warpgroup so far (all items): parity of s_full / o_full
    for (int li = wg; li < my_items; li += NWG) {
      const Item c = decode_item(p, (int)blockIdx.x + li * (int)gridDim.x);
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < T; ++j, ++nblk) {
        mbar_wait(&sm.s_full[wg], nblk & 1);
        fence_after_sync();
        uint32_t ra[32], rb[32];
        float mb[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // independent chains
        const int vk = tile_vcount(p, j);  // valid keys of this block (128 unless the stripe shape leaves a ragged tile)
        tmem_ld32(lane_base, ra);
        tmem_wait_ld();
#pragma unroll
        for (int ch = 0; ch < 4; ch += 2) {
          tmem_ld32(lane_base + (ch + 1) * 32, rb);
          if (vk < TILE) mask_keys(ra, ch, vk);
#pragma unroll
          for (int i = 0; i < 32; ++i) mb[i & 3] = fmaxf(mb[i & 3], __uint_as_float(ra[i]));
          tmem_wait_ld();
          if (ch + 2 < 4) tmem_ld32(lane_base + (ch + 2) * 32, ra);
          if (vk < TILE) mask_keys(rb, ch + 1, vk);
#pragma unroll
          for (int i = 0; i < 32; ++i) mb[i & 3] = fmaxf(mb[i & 3], __uint_as_float(rb[i]));
          tmem_wait_ld();
        }
        const float m_blk = fmaxf(fmaxf(mb[0], mb[1]), fmaxf(mb[2], mb[3]));
        const float m_new = fmaxf(m_run, m_blk);
        if (j > 0) {
          // online softmax: everything accumulated so far is relative to m_run
          const float alpha = ex2((m_run - m_new) * p.scale_log2);
          l_run *= alpha;
          mbar_wait(&sm.o_full[wg], (nblk - 1) & 1);  // P V of the previous block has completed
          fence_after_sync();
          uint32_t ro[32];
          tmem_ld32(lane_base + O_COL, ro);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) ro[i] = __float_as_uint(__uint_as_float(ro[i]) * alpha);
          uint32_t lo16[16], hi16[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            lo16[i] = ro[i];
            hi16[i] = ro[16 + i];
          }
          tmem_st16(lane_base + O_COL, lo16);
          tmem_st16(lane_base + O_COL + 16, hi16);
        }
        m_run = m_new;
        const float neg_m = -m_new * p.scale_log2;
        const f2_t scale2 = f2_splat(p.scale_log2), negm2 = f2_splat(neg_m);
        float l0 = 0.f, l1 = 0.f;
        auto exp_chunk = [&](uint32_t (&r)[32], int ch) {
          uint32_t pk[16];
          if (vk < TILE) mask_keys(r, ch, vk);  // 2^(-inf) = 0: masked keys get no probability
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const f2_t x2 = f2_fma(f2_make(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), scale2, negm2);
            float p0, p1;
            if (POLY_MOD > 0 && i % (POLY_MOD > 0 ? POLY_MOD : 1) == POLY_MOD - 1) {  // FMA-pipe 2^x (tc_common.cuh)
              ex2_poly_pair(x2, p0, p1);
            } else {
              float x0, x1;
              f2_split(x2, x0, x1);
              p0 = ex2(x0);
              p1 = ex2(x1);
            }
            l0 += p0;
            l1 += p1;
            pk[i] = pack_bf16x2(p0, p1);
          }
          tmem_st16(lane_base + ch * 16, pk);  // P over S columns that were already consumed
        };
        tmem_ld32(lane_base, ra);
        tmem_wait_ld();
#pragma unroll
        for (int ch = 0; ch < 4; ch += 2) {
          tmem_ld32(lane_base + (ch + 1) * 32, rb);
          exp_chunk(ra, ch);
          tmem_wait_ld();
          if (ch + 2 < 4) tmem_ld32(lane_base + (ch + 2) * 32, ra);
          exp_chunk(rb, ch + 1);
          tmem_wait_ld();
        }
        l_run += l0 + l1;
        tmem_wait_st();
        fence_before_sync();
        mbar_arrive(&sm.p_full[wg]);
      }
      // ---- epilogue: O / l + LePE -> out, lse ----
      mbar_wait(&sm.o_full[wg], (nblk - 1) & 1);
      fence_after_sync();
      uint32_t r[32];
      tmem_ld32(lane_base + O_COL, r);
      tmem_wait_ld();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.buf_empty[wg]);  // the next item's first S may overwrite this buffer
      const float inv_l = 1.f / l_run;
      if (row >= tile_vcount(p, c.qt)) continue;  // token rows of the tile that are not queries of this stripe
      int yy, xx;
      tile_pos(p, c.qt, row, yy, xx);
      const int y0 = c.wy * p.hs, x0 = c.wx * p.ws;
      float o[HD];
#pragma unroll
      for (int cc = 0; cc < HD; ++cc)
        o[cc] = fmaf(__uint_as_float(r[cc]), inv_l, __ldg(p.lepe_b + c.head * HD + cc));
      const __nv_bfloat16* vb = p.v + (int64_t)c.b * p.v_sb + c.head * HD;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int ny = yy + ky - 1;
        if (ny < 0 || ny >= p.hs) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int nx = xx + kx - 1;
          if (nx < 0 || nx >= p.ws) continue;
          const uint4* vp = reinterpret_cast<const uint4*>(vb + (int64_t)((y0 + ny) * p.W + x0 + nx) * p.v_sl);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            float f[8];
            unpack<__nv_bfloat16>(__ldg(vp + q4), f);
#pragma unroll
            for (int e2 = 0; e2 < 8; ++e2)
              o[q4 * 8 + e2] = fmaf(__ldg(p.lepe_w + (c.head * HD + q4 * 8 + e2) * 9 + ky * 3 + kx), f[e2], o[q4 * 8 + e2]);
          }
        }
      }
      const int tok = (y0 + yy) * p.W + x0 + xx;
      uint4* dst = reinterpret_cast<uint4*>(p.out + (int64_t)c.b * p.o_sb + (int64_t)tok * p.o_sl + c.head * HD);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        float f[8];
#pragma unroll
        for (int e2 = 0; e2 < 8; ++e2) f[e2] = o[q4 * 8 + e2];
        dst[q4] = pack<__nv_bfloat16>(f);
      }
      p.lse[((int64_t)c.b * p.heads + c.head) * p.L + tok] = m_run * p.scale + __logf(l_run);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

}  // namespace

// Stripes of more than 128 tokens that the single-pass kernels do not take: bf16, stripe width <= 256, at most 64
// tiles per stripe, no attention dropout.  ANY stripe shape — tiles that a stripe does not fill are masked.  Forward
// only (inference; a backward pass for these shapes uses the CUDA-core engine).
static void kv_tiling(const StripeGeom& g, int* bx, int* by, int* tpr, int* T) {
  if (g.ws > TILE) {
    *bx = TILE; *by = 1;
    *tpr = (g.ws + TILE - 1) / TILE;
    *T = g.hs * *tpr;
  } else {
    *bx = g.ws; *by = TILE / g.ws;
    if (*by > g.hs) *by = g.hs;
    *tpr = 1;
    *T = (g.hs + *by - 1) / *by;
  }
}
bool tc_fwd_kv_supported(const StripeGeom& g, int dtype) {
  if (dtype != CSB200_BF16 || g.drop_thr != 0) return false;
  if (g.N <= TILE || g.ws > 256 || g.hs > 256) return false;
  int bx, by, tpr, T;
  kv_tiling(g, &bx, &by, &tpr, &T);
  return T >= 2 && T <= 64;
}

int tc_fwd_kv(const StripeGeom& g, const void* q, const void* k, const void* v, const float* lepe_w,
              const float* lepe_b, void* out, float* lse, cudaStream_t st) {
  KvMaps maps;
  KvParams p;
  memset(&maps, 0, sizeof(maps));
  memset(&p, 0, sizeof(p));
  int bx, by, tpr, T;
  kv_tiling(g, &bx, &by, &tpr, &T);
  int rc;
  if ((rc = tc_make_map(&maps.q, q, g, g.q_sb, g.q_sl, bx, by)) != CSB200_OK) return rc;
  if ((rc = tc_make_map(&maps.k, k, g, g.k_sb, g.k_sl, bx, by)) != CSB200_OK) return rc;
  if ((rc = tc_make_map(&maps.v, v, g, g.v_sb, g.v_sl, bx, by)) != CSB200_OK) return rc;
  p.B = g.B; p.W = g.W; p.L = g.L;
  p.hs = g.hs; p.ws = g.ws; p.nwy = g.nwy; p.nwx = g.nwx; p.heads = g.heads;
  p.by = by; p.tpr = tpr; p.T = T;
  p.box_bytes = bx * by * ROW_BYTES;
  const int64_t items = (int64_t)g.B * g.nwy * g.nwx * g.heads * p.T;
  if (items > 0x7fffffff) return fail(CSB200_ERR_INVALID, "stripe_fwd_tc_kv: too many work items");
  p.items = (int)items;
  p.scale = g.scale;
  p.scale_log2 = g.scale * 1.4426950408889634f;
  p.lepe_w = lepe_w; p.lepe_b = lepe_b;
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.v_sb = g.v_sb; p.v_sl = g.v_sl;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.o_sb = g.o_sb; p.o_sl = g.o_sl;
  p.lse = lse;
  const int sm_count = device_sm_count();
  if (sm_count <= 0) return fail(CSB200_ERR_CUDA, "stripe_fwd_tc_kv: cannot query the SM count");
  // > half of the 227 KB so that exactly one CTA (which owns all 512 TMEM columns) fits per SM
  const int smem = (int)sizeof(KvSmem) + 1024 > 120 * 1024 ? (int)sizeof(KvSmem) + 1024 : 120 * 1024;
  CSB200_CUDA(opt_in_smem(reinterpret_cast<const void*>(&stripe_fwd_tc_kv), smem));
  const int grid = p.items < sm_count ? p.items : sm_count;
  stripe_fwd_tc_kv<<<grid, THREADS, smem, st>>>(maps, p);
  return check_launch("stripe_fwd_tc_kv");
}

}  // namespace csb200

// Stripe attention + LePE, tcgen05 / TMEM / TMA engine — backward (bf16 in, fp32 accumulate).
//
// Replaces what autograd records for LePEAttention.forward (C:271-298): the N x N probabilities
// are recomputed from the saved log-sum-exp, never stored.  Per (image, stripe, head) group the
// stripe's Q, K, V and grad_out tiles (N x 32 bf16) arrive by TMA, 64-byte swizzled, and serve
// both as K-major operands (contraction over the 32 channels) and as MN-major operands
// (contraction over tokens), so five GEMMs run on the tensor core with no transposes:
//
//   per key tile kt (128 keys = TMEM lanes), per 64-query half-block qh:
//     S^T  = K_kt Q_qh^T          M128 x N64 x K32   (A, B K-major smem)      -> TMEM
//     dP^T = V_kt dO_qh^T         M128 x N64 x K32                            -> TMEM
//     convert warps, one thread per key row:
//       P^T = exp2(S^T scale log2e - lse log2e);  dS^T = scale P^T (dP^T - delta)
//       P^T, dS^T -> TMEM as bf16 (A operands);  dS^T -> smem, 128B-swizzled MN-major (A of dQ)
//     dV_kt += P^T  dO_qh         M128 x N32 x K64   (A TMEM, B MN-major smem)
//     dK_kt += dS^T Q_qh          M128 x N32 x K64
//     every second qh:  dQ_qb += dS K_kt   M128 x N32 x K128 (A MN-major smem, B MN-major smem)
//
//   delta = rowsum(grad_out * (out - lepe)) comes from stripe_bwd_delta (CUDA cores, HBM-bound);
//   the LePE transposed stencil on grad_out is added to dV in the epilogue from the grad_out tile
//   in shared memory; the depthwise weight / bias gradients use the shared reduction kernels.
//
//   warp 0 TMA producer | warp 1 MMA issuer (S^T, dP^T) | warp 2 TMEM allocator | warp 3 MMA issuer
//   (dV, dK, dQ) | warps 4-7, 8-11 convert
//   warpgroups (even / odd half-blocks, TMEM stage 0 / 1) | warps 12-15 epilogue warpgroup.
//
// TMEM (512 columns): stage s at 128 s: S^T [0,64) -> P^T bf16 [0,32); dP^T [64,128) -> dS^T bf16
// [64,96).  Accumulators: (dV, dK) sets at 256 + 64 a (a = key-tile parity); dQ sets at 384 + 64 g
// (g = group parity), 32 columns per 128-query block.

#include <cstring>

#include "stripe_attn.cuh"
#include "tc_common.cuh"

namespace csb200 {
#ifdef CSB_PROF
__device__ unsigned long long g_prof_bwd[32];
#define PROF_T(v) const long long v = clock64()
#define PROF_ADDT(tid, i, a, b) if (blockIdx.x == 0 && threadIdx.x == (tid)) atomicAdd(&g_prof_bwd[i], (unsigned long long)((b) - (a)))
#else
#define PROF_T(v)
#define PROF_ADDT(tid, i, a, b)
#endif
namespace {
using namespace tc;

constexpr int HD = 32;
constexpr int TILE = 128;
constexpr int ROW_BYTES = HD * 2;
constexpr int TILE_BYTES = TILE * ROW_BYTES;  // 8 KB
constexpr int HALF = 64;                      // queries per convert iteration
constexpr int DS_BLOCK_BYTES = TILE * 128;    // 128 key rows x 64 queries bf16 = 16 KB
constexpr int THREADS = 512;

// one branch (stripe orientation); a launch covers up to two, interleaved image by image
struct BwdBranch {
  int hs, ws, ws_log2, nwy, nwx, heads, by;
  const float* lepe_w;  // [C'][9]
  const float* lse;     // [B][heads][L]
  const float* delta;   // [B][heads][L]
  __nv_bfloat16 *dq, *dk, *dv;
  int64_t dq_sb, dq_sl, dk_sb, dk_sl, dv_sb, dv_sl;
  const uint32_t* drop_mask;  // attention dropout: the forward pass's transposed keep bits
  const float* wg_partial;    // [wg_blocks][C'][10] partial LePE gradients of lepe_prep (nullptr: already summed)
  float *gw, *gb;             // [C'][9], [C']
};
struct BwdParams {
  int B, W, L;
  int g0, gpi, groups;  // groups per image of branch 0 / of both; B * gpi
  float scale, scale_log2;
  float keep_scale;     // 1 / (1 - p) of the attention dropout (1 in the <.., false> instantiation)
  int wg_blocks, nbr;
  BwdBranch br[2];
};
struct BwdMaps {
  CUtensorMap q[2], k[2], v[2], go[2];
};

template <int NK>
struct BCfg;
template <>
struct BCfg<128> {
  static constexpr int GS = 4;  // group stages in shared memory
};
template <>
struct BCfg<256> {
  static constexpr int GS = 2;
};

template <int NK>
struct BSmem {
  static constexpr int GS = BCfg<NK>::GS;
  static constexpr int OP_BYTES = NK * ROW_BYTES;
  alignas(1024) uint8_t q[GS][OP_BYTES];
  alignas(1024) uint8_t k[GS][OP_BYTES];
  alignas(1024) uint8_t v[GS][OP_BYTES];
  alignas(1024) uint8_t go[GS][OP_BYTES];
  alignas(1024) uint8_t ds[2][2 * DS_BLOCK_BYTES];  // dS^T of one 128-query block, two 64-blocks
  alignas(16) float lse2[GS][NK];                   // -lse * log2(e)
  alignas(16) float delta[GS][NK];                  // -scale * delta
  alignas(16) float lepe[GS][9 * HD];               // [tap][c]
  alignas(16) int4 coord[GS];                       // image, first token of the stripe, head
  alignas(8) uint64_t grp_full[GS], grp_empty[GS];
  uint64_t sdp_full[2], conv_done[2], stage_free[2], ds_free[2];
  uint64_t dvdk_full[2], dvdk_empty[2], dq_full[2], dq_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ const uint4* sw64_chunk(const uint8_t* tile, int n, int chunk) {
  return reinterpret_cast<const uint4*>(tile + n * ROW_BYTES + ((chunk ^ ((n >> 1) & 3)) << 4));
}

template <int NK, bool DROP>
__global__ void __launch_bounds__(THREADS, 1)
    stripe_bwd_tc(const __grid_constant__ BwdMaps maps, const __grid_constant__ BwdParams p) {
  constexpr int T = NK / TILE;      // key tiles (and 128-query blocks) per group
  constexpr int NH = NK / HALF;     // 64-query half-blocks per key tile
  constexpr int NIT = T * NH;       // convert iterations per group (2 or 8: always even)
  constexpr int GS = BCfg<NK>::GS;
  constexpr uint32_t ACC_DVDK = 256, ACC_DQ = 384;
  extern __shared__ uint8_t smem_raw[];
  // align inside the shared window: pointer + integer offset keeps the shared address space (an
  // integer -> pointer cast makes every access a generic LD/ST with 64-bit address math)
  BSmem<NK>& sm = *reinterpret_cast<BSmem<NK>*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_groups = (p.groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_it = my_groups * NIT;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < (p.gpi > p.g0 ? 2 : 1); ++i) {
      prefetch_tensormap(&maps.q[i]);
      prefetch_tensormap(&maps.k[i]);
      prefetch_tensormap(&maps.v[i]);
      prefetch_tensormap(&maps.go[i]);
    }
    for (int i = 0; i < GS; ++i) {
      mbar_init(&sm.grp_full[i], 2);  // expect_tx arrival + one after the lse / delta / tap stores
      mbar_init(&sm.grp_empty[i], 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.sdp_full[i], 1);
      mbar_init(&sm.conv_done[i], 128);
      mbar_init(&sm.stage_free[i], 1);
      mbar_init(&sm.ds_free[i], 1);
      mbar_init(&sm.dvdk_full[i], 1);
      mbar_init(&sm.dvdk_empty[i], 4);
      mbar_init(&sm.dq_full[i], 1);
      mbar_init(&sm.dq_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&sm.tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  PROF_T(kbeg);

  if (warp == 0 || warp == 2) {
    // ============================== producers (two warps) ===================================
    // warp 0 takes the even groups of this CTA, warp 2 the odd ones.  One thread gets an 8-KB box of 64-byte rows
    // through every ~800 cycles whatever the ring depth, two warps together reach what the access pattern
    // sustains (benchmarks/debug/tma_rows.cu: 2.6 -> 4.3 TB/s); with all boxes on one warp the MMA warps waited
    // ~3 300 cycles per group for data at N = 128.  The four boxes of a group stay on one warp, back to back
    // (q | k | v of a token share a DRAM page).
    // The lse / delta / LePE-tap gathers are fetched ONE GROUP AHEAD into registers (an L2 / DRAM round trip that
    // would otherwise sit between the TMA issues of consecutive groups).
    const int first = warp >> 1;
    constexpr int PER_LANE = NK / 32;
    float r_lse[PER_LANE], r_delta[PER_LANE], r_tap[9];
    struct GC {
      int b, br, head, wx, wy, tok0;
    };
    auto decode = [&](int gi) {
      GC c;
      const int g = (int)blockIdx.x + gi * (int)gridDim.x;
      c.b = g / p.gpi;
      int r = g - c.b * p.gpi;
      c.br = r >= p.g0 ? 1 : 0;
      r -= c.br ? p.g0 : 0;
      const BwdBranch& bg = p.br[c.br];
      c.head = r % bg.heads;
      r /= bg.heads;
      c.wx = r % bg.nwx;
      c.wy = r / bg.nwx;
      c.tok0 = (c.wy * bg.hs) * p.W + c.wx * bg.ws;
      return c;
    };
    auto gather = [&](const GC& c) {
      const BwdBranch& bg = p.br[c.br];
      const float* lse = bg.lse + ((int64_t)c.b * bg.heads + c.head) * p.L;
      const float* dl = bg.delta + ((int64_t)c.b * bg.heads + c.head) * p.L;
#pragma unroll
      for (int j = 0; j < PER_LANE; ++j) {
        const int i = lane + 32 * j;
        const int tok = c.tok0 + (i >> bg.ws_log2) * p.W + (i & (bg.ws - 1));
        r_lse[j] = __ldg(lse + tok);
        r_delta[j] = __ldg(dl + tok);
      }
#pragma unroll
      for (int j = 0; j < 9; ++j) r_tap[j] = __ldg(bg.lepe_w + (c.head * HD + lane) * 9 + j);  // [tap j][c = lane]
    };
    if (first < my_groups) gather(decode(first));
    for (int gi = first; gi < my_groups; gi += 2) {
      const GC c = decode(gi);
      const BwdBranch& bg = p.br[c.br];
      const int gs = gi % GS;
      mbar_wait(&sm.grp_empty[gs], ((gi / GS) & 1) ^ 1);
      if (lane == 0) {
        mbar_expect_tx(&sm.grp_full[gs], 4 * BSmem<NK>::OP_BYTES);
        const int x0 = c.wx * bg.ws, y0 = c.wy * bg.hs;
#pragma unroll
        for (int bxi = 0; bxi < T; ++bxi) {
          const int dx = (bg.ws > TILE) ? (bxi * TILE) % bg.ws : 0;
          const int dy = (bg.ws > TILE) ? (bxi * TILE) / bg.ws : bxi * bg.by;
          const int off = bxi * TILE_BYTES;
          tma_load_4d(sm.k[gs] + off, &maps.k[c.br], &sm.grp_full[gs], c.head * HD, x0 + dx, y0 + dy, c.b);
          tma_load_4d(sm.q[gs] + off, &maps.q[c.br], &sm.grp_full[gs], c.head * HD, x0 + dx, y0 + dy, c.b);
          tma_load_4d(sm.v[gs] + off, &maps.v[c.br], &sm.grp_full[gs], c.head * HD, x0 + dx, y0 + dy, c.b);
          tma_load_4d(sm.go[gs] + off, &maps.go[c.br], &sm.grp_full[gs], c.head * HD, x0 + dx, y0 + dy, c.b);
        }
      }
#pragma unroll
      for (int j = 0; j < PER_LANE; ++j) {
        // stored NEGATED and pre-scaled, so that the convert warps form  x = s * scale * log2(e) - lse * log2(e)  and
        // scale * (dP - delta)  as one packed FMA each
        sm.lse2[gs][lane + 32 * j] = r_lse[j] * -1.4426950408889634f;
        sm.delta[gs][lane + 32 * j] = r_delta[j] * -p.scale;
      }
#pragma unroll
      for (int j = 0; j < 9; ++j) sm.lepe[gs][lane + 32 * j] = r_tap[j];
      if (lane == 0) sm.coord[gs] = make_int4(c.b, c.tok0, c.head, c.br);
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.grp_full[gs]);  // second arrival: the plain stores above are done
      if (gi + 2 < my_groups) gather(decode(gi + 2));
    }
  } else if (warp == 1) {
    // ============================ MMA issuer 1: S^T and dP^T ================================
    // Two issuing warps: measured, one warp needs ~400 cycles to issue the S^T / dP^T MMAs of an
    // iteration and ~600 for the dependent ones, and — being in order — it also held S^T(x+1) back
    // until convert(x-1) had finished.  Ordering between the two warps' MMAs is carried by the same
    // mbarriers as before (stage_free / conv_done), so nothing else changes.
    // The whole warp walks the loop; one elected lane issues (tc_common.cuh, "cheap issue path").
    constexpr uint32_t idesc_sdp = umma_idesc_bf16(HALF, false, false);
    constexpr uint32_t OP16 = BSmem<NK>::OP_BYTES >> 4, HALF16 = (HALF * ROW_BYTES) >> 4;
    const uint32_t q_lo0 = desc_lo_sw64(smem_u32(sm.q[0])), k_lo0 = desc_lo_sw64(smem_u32(sm.k[0]));
    const uint32_t v_lo0 = desc_lo_sw64(smem_u32(sm.v[0])), go_lo0 = desc_lo_sw64(smem_u32(sm.go[0]));
    for (int x = 0; x < total_it; ++x) {
      const int st = x & 1, gi = x / NIT, r = x % NIT, kt = r / NH, qh = r % NH, gs = gi % GS;
      PROF_T(e0);
      if (r == 0) mbar_wait(&sm.grp_full[gs], (gi / GS) & 1);
      PROF_T(e1);
      mbar_wait(&sm.stage_free[st], ((x >> 1) & 1) ^ 1);
      fence_after_sync();
      PROF_T(e2);
      PROF_ADDT(32, 10, e0, e1); PROF_ADDT(32, 11, e1, e2);
      if (elect_one_sync()) {
        const uint32_t k_lo = k_lo0 + gs * OP16 + kt * (TILE_BYTES >> 4);
        const uint32_t v_lo = v_lo0 + gs * OP16 + kt * (TILE_BYTES >> 4);
        const uint32_t q_lo = q_lo0 + gs * OP16 + qh * HALF16;
        const uint32_t go_lo = go_lo0 + gs * OP16 + qh * HALF16;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_ss2(tmem + st * 128, k_lo + k * (32 >> 4), DESC_HI_SW64, q_lo + k * (32 >> 4),
                   DESC_HI_SW64, idesc_sdp, k > 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_ss2(tmem + st * 128 + 64, v_lo + k * (32 >> 4), DESC_HI_SW64, go_lo + k * (32 >> 4),
                   DESC_HI_SW64, idesc_sdp, k > 0);
        umma_commit(&sm.sdp_full[st]);
      }
      __syncwarp();
      PROF_T(e3);
      PROF_ADDT(32, 12, e2, e3);
    }
  } else if (warp == 3) {
    // ==================== MMA issuer 2: dV, dK, dQ (consume the converted tiles) ==================
    constexpr uint32_t idesc_acc = umma_idesc_bf16(HD, false, true);   // A from TMEM, B MN-major
    constexpr uint32_t idesc_dq = umma_idesc_bf16(HD, true, true);     // A MN-major smem
    constexpr uint32_t OP16 = BSmem<NK>::OP_BYTES >> 4, HALF16 = (HALF * ROW_BYTES) >> 4;
    const uint32_t q_lo0 = desc_lo_sw64(smem_u32(sm.q[0])), k_lo0 = desc_lo_sw64(smem_u32(sm.k[0]));
    const uint32_t go_lo0 = desc_lo_sw64(smem_u32(sm.go[0]));
    const uint32_t ds_lo0 = desc_lo_sw128_mn(smem_u32(sm.ds[0]), DS_BLOCK_BYTES);
    for (int y = 0; y < total_it; ++y) {
      const int st = y & 1, gi = y / NIT, r = y % NIT, kt = r / NH, qh = r % NH, gs = gi % GS;
      const int ktc = gi * T + kt, aset = ktc & 1;
      const int qb = qh >> 1, qbc = ktc * T + qb, dsb = qbc & 1;
      PROF_T(d0);
      if (r == 0) mbar_wait(&sm.grp_full[gs], (gi / GS) & 1);  // this thread's own view of the TMA data
      mbar_wait(&sm.conv_done[st], (y >> 1) & 1);
      PROF_T(d1);
      if (qh == 0) mbar_wait(&sm.dvdk_empty[aset], ((ktc >> 1) & 1) ^ 1);
      if ((qh & 1) && kt == 0 && qb == 0) mbar_wait(&sm.dq_empty[gi & 1], ((gi >> 1) & 1) ^ 1);
      fence_after_sync();
      PROF_T(d2);
      PROF_ADDT(96, 8, d0, d1); PROF_ADDT(96, 9, d1, d2);
      if (elect_one_sync()) {
        const uint32_t sbase = tmem + st * 128;
        const uint32_t go_lo = go_lo0 + gs * OP16 + qh * HALF16;
        const uint32_t q_lo = q_lo0 + gs * OP16 + qh * HALF16;
#pragma unroll
        for (int k = 0; k < HALF / 16; ++k)  // dV += P^T dO : 16 queries per step
          umma_ts2(tmem + ACC_DVDK + aset * 64, sbase + 8 * k, go_lo + k * (1024 >> 4), DESC_HI_SW64,
                   idesc_acc, (qh > 0 || k > 0));
#pragma unroll
        for (int k = 0; k < HALF / 16; ++k)  // dK += dS^T Q
          umma_ts2(tmem + ACC_DVDK + aset * 64 + 32, sbase + 64 + 8 * k, q_lo + k * (1024 >> 4),
                   DESC_HI_SW64, idesc_acc, (qh > 0 || k > 0));
        umma_commit(&sm.stage_free[st]);
        if (qh & 1) {  // a 128-query block of dS^T is complete in shared memory
          const uint32_t ds_lo = ds_lo0 + dsb * ((2 * DS_BLOCK_BYTES) >> 4);
          const uint32_t k_lo = k_lo0 + gs * OP16 + kt * (TILE_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < TILE / 16; ++k)  // dQ += dS K : 16 keys per step
            umma_ss2(tmem + ACC_DQ + (gi & 1) * 64 + qb * 32, ds_lo + k * (2048 >> 4), DESC_HI_SW128,
                     k_lo + k * (1024 >> 4), DESC_HI_SW64, idesc_dq, (kt > 0 || k > 0));
          umma_commit(&sm.ds_free[dsb]);
        }
        if (qh == NH - 1) {
          umma_commit(&sm.dvdk_full[aset]);
          if (kt == T - 1) umma_commit(&sm.dq_full[gi & 1]);
        }
      }
      __syncwarp();
      PROF_T(d3);
      PROF_ADDT(96, 13, d2, d3);
    }
  } else if (warp >= 4 && warp < 12) {
    // ================================ convert warpgroups ====================================
    const int wg = (warp - 4) >> 2;             // == TMEM stage == parity of the half-block
    const int j = ((warp & 3) << 5) | lane;     // key row inside the tile == TMEM lane
    const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) << 5) << 16) + wg * 128;
    for (int x = wg; x < total_it; x += 2) {
      const int gi = x / NIT, r = x % NIT, kt = r / NH, qh = r % NH, gs = gi % GS;
      const int qbc = (gi * T + kt) * T + (qh >> 1), dsb = qbc & 1;
      PROF_T(c0);
      mbar_wait(&sm.grp_full[gs], (gi / GS) & 1);   // lse / delta visibility (already complete)
      // attention dropout: this key's mask row holds the keep bits of the 64 queries of the half-block
      uint2 mw = make_uint2(0xffffffffu, 0xffffffffu);
      if constexpr (DROP) {
        const int4 gc = sm.coord[gs];
        const BwdBranch& bg = p.br[gc.w];
        const int nj = kt * TILE + j;
        const int tokj = gc.y + (nj >> bg.ws_log2) * p.W + (nj & (bg.ws - 1));
        mw = __ldg(reinterpret_cast<const uint2*>(
            bg.drop_mask + (((int64_t)gc.x * bg.heads + gc.z) * p.L + tokj) * (NK / 32) + qh * 2));
      }
      mbar_wait(&sm.ds_free[dsb], ((qbc >> 1) & 1) ^ 1);
      PROF_T(c1);
      mbar_wait(&sm.sdp_full[wg], (x >> 1) & 1);
      fence_after_sync();
      PROF_T(c2);
      uint8_t* ds_row = sm.ds[dsb] + (qh & 1) * DS_BLOCK_BYTES + j * 128;
      const float* lse2 = sm.lse2[gs] + qh * HALF;
      const float* dlt = sm.delta[gs] + qh * HALF;
      const f2_t sl2 = f2_splat(p.scale_log2), sc2 = f2_splat(p.scale);
#pragma unroll
      for (int c = 0; c < 2; ++c) {  // 32 query columns at a time
        uint32_t rs[32], rd[32];
        tmem_ld32(lane_base + 32 * c, rs);
        tmem_ld32(lane_base + 64 + 32 * c, rd);
        tmem_wait_ld();
        uint32_t pp[16], pd[16];
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          // four queries at a time, as two packed fp32 pairs (the warps are bound by instruction issue):
          // P = 2^(s scale log2e - lse log2e),  dS = P * (scale dP - scale delta)
          const ulonglong2 l4 = *reinterpret_cast<const ulonglong2*>(lse2 + 32 * c + 4 * e4);  // broadcast: -lse log2e
          const ulonglong2 d4 = *reinterpret_cast<const ulonglong2*>(dlt + 32 * c + 4 * e4);   // -scale delta
          const f2_t xa = f2_fma(f2_make(__uint_as_float(rs[4 * e4 + 0]), __uint_as_float(rs[4 * e4 + 1])), sl2, l4.x);
          const f2_t xb = f2_fma(f2_make(__uint_as_float(rs[4 * e4 + 2]), __uint_as_float(rs[4 * e4 + 3])), sl2, l4.y);
          float x0, x1, x2, x3;
          f2_split(xa, x0, x1);
          f2_split(xb, x2, x3);
          const float p0 = ex2(x0), p1 = ex2(x1), p2 = ex2(x2), p3 = ex2(x3);
          if constexpr (DROP) {
            // dropped P (the A operand of dV) = keep / (1 - p) * P, and dP reaches P through the same factor
            const uint32_t kb = (c == 0 ? mw.x : mw.y) >> (4 * e4);
            const float k0 = kb & 1u ? p.keep_scale : 0.f, k1 = kb & 2u ? p.keep_scale : 0.f;
            const float k2 = kb & 4u ? p.keep_scale : 0.f, k3 = kb & 8u ? p.keep_scale : 0.f;
            const f2_t ka = f2_make(k0, k1), kb2 = f2_make(k2, k3);
            const f2_t da = f2_mul(f2_make(__uint_as_float(rd[4 * e4 + 0]), __uint_as_float(rd[4 * e4 + 1])), ka);
            const f2_t db = f2_mul(f2_make(__uint_as_float(rd[4 * e4 + 2]), __uint_as_float(rd[4 * e4 + 3])), kb2);
            float s0, s1, s2, s3;
            f2_split(f2_mul(f2_make(p0, p1), f2_fma(da, sc2, d4.x)), s0, s1);
            f2_split(f2_mul(f2_make(p2, p3), f2_fma(db, sc2, d4.y)), s2, s3);
            pp[2 * e4] = pack_bf16x2(p0 * k0, p1 * k1);
            pp[2 * e4 + 1] = pack_bf16x2(p2 * k2, p3 * k3);
            pd[2 * e4] = pack_bf16x2(s0, s1);
            pd[2 * e4 + 1] = pack_bf16x2(s2, s3);
            continue;
          }
          float s0, s1, s2, s3;
          f2_split(f2_mul(f2_make(p0, p1),
                          f2_fma(f2_make(__uint_as_float(rd[4 * e4 + 0]), __uint_as_float(rd[4 * e4 + 1])), sc2, d4.x)),
                   s0, s1);
          f2_split(f2_mul(f2_make(p2, p3),
                          f2_fma(f2_make(__uint_as_float(rd[4 * e4 + 2]), __uint_as_float(rd[4 * e4 + 3])), sc2, d4.y)),
                   s2, s3);
          pp[2 * e4] = pack_bf16x2(p0, p1);
          pp[2 * e4 + 1] = pack_bf16x2(p2, p3);
          pd[2 * e4] = pack_bf16x2(s0, s1);
          pd[2 * e4 + 1] = pack_bf16x2(s2, s3);
        }
        tmem_st16(lane_base + 16 * c, pp);       // P^T over S^T columns already consumed
        tmem_st16(lane_base + 64 + 16 * c, pd);  // dS^T over dP^T columns already consumed
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {         // dS^T row -> smem, 128B swizzle: chunk ^ (row & 7)
          const int chunk = 4 * c + q4;
          *reinterpret_cast<uint4*>(ds_row + ((chunk ^ (j & 7)) << 4)) =
              make_uint4(pd[4 * q4], pd[4 * q4 + 1], pd[4 * q4 + 2], pd[4 * q4 + 3]);
        }
      }
      tmem_wait_st();
      fence_before_sync();
      fence_proxy_async_smem();
      mbar_arrive(&sm.conv_done[wg]);
      PROF_T(c3);
      PROF_ADDT(128, 0, c0, c1); PROF_ADDT(128, 1, c1, c2); PROF_ADDT(128, 2, c2, c3); PROF_ADDT(128, 3, c0, c0 + 1);
    }
  } else if (warp >= 12) {
    // ================================== epilogue warpgroup ==================================
    const int row = ((warp & 3) << 5) | lane;
    const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) << 5) << 16);
    // Prologue: the depthwise weight / bias gradients of get_v (C:244) — fixed-order sums of the per-CTA partials
    // that lepe_prep left behind, one warp per output, the outputs dealt round-robin to the CTAs.  Runs while
    // the first group's tiles are still in flight (this warpgroup has nothing to drain yet).
    if (p.wg_blocks > 0) {
      for (int br = 0; br < p.nbr; ++br) {
        const BwdBranch& bw = p.br[br];
        if (bw.wg_partial == nullptr) continue;
        const int outs = bw.heads * HD * 10;
        for (int o = (int)blockIdx.x * 4 + (warp & 3); o < outs; o += (int)gridDim.x * 4) {
          const float a = strided_partial_sum(bw.wg_partial + o, p.wg_blocks, outs, lane);
          if (lane == 0) {
            const int c = o / 10, tap = o % 10;
            if (tap == 9) bw.gb[c] = a;
            else bw.gw[c * 9 + tap] = a;
          }
        }
      }
    }
    for (int gi = 0; gi < my_groups; ++gi) {
      const int gs = gi % GS;
      mbar_wait(&sm.grp_full[gs], (gi / GS) & 1);
      const int4 gc = sm.coord[gs];
      const BwdBranch& bg = p.br[gc.w];
      const float* lw = sm.lepe[gs];
      const uint8_t* got = sm.go[gs];
#pragma unroll 1
      for (int kt = 0; kt < T; ++kt) {
        const int ktc = gi * T + kt, aset = ktc & 1;
        PROF_T(f0);
        mbar_wait(&sm.dvdk_full[aset], (ktc >> 1) & 1);
        fence_after_sync();
        PROF_T(f1);
        PROF_ADDT(384, 16, f0, f1);
        uint32_t rv[32], rk[32];
        tmem_ld32(lane_base + ACC_DVDK + aset * 64, rv);
        tmem_ld32(lane_base + ACC_DVDK + aset * 64 + 32, rk);
        tmem_wait_ld();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.dvdk_empty[aset]);
        const int n = kt * TILE + row;
        const int yy = n >> bg.ws_log2, xx = n & (bg.ws - 1);
        const int tok = gc.y + yy * p.W + xx;
        float dv[HD];
#pragma unroll
        for (int cc = 0; cc < HD; ++cc) dv[cc] = __uint_as_float(rv[cc]);
        // transposed LePE stencil: v_n receives w[tap] * grad_out[n - offset(tap)] (stripe-local)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int ny = yy - ky + 1;
          if (ny < 0 || ny >= bg.hs) continue;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int nx = xx - kx + 1;
            if (nx < 0 || nx >= bg.ws) continue;
            const int nn = (ny << bg.ws_log2) + nx;
            const float* wt = lw + (ky * 3 + kx) * HD;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              float f[8];
              unpack<__nv_bfloat16>(*sw64_chunk(got, nn, q4), f);
              const float4 w0 = *reinterpret_cast<const float4*>(wt + q4 * 8);
              const float4 w1 = *reinterpret_cast<const float4*>(wt + q4 * 8 + 4);
              dv[q4 * 8 + 0] = fmaf(w0.x, f[0], dv[q4 * 8 + 0]);
              dv[q4 * 8 + 1] = fmaf(w0.y, f[1], dv[q4 * 8 + 1]);
              dv[q4 * 8 + 2] = fmaf(w0.z, f[2], dv[q4 * 8 + 2]);
              dv[q4 * 8 + 3] = fmaf(w0.w, f[3], dv[q4 * 8 + 3]);
              dv[q4 * 8 + 4] = fmaf(w1.x, f[4], dv[q4 * 8 + 4]);
              dv[q4 * 8 + 5] = fmaf(w1.y, f[5], dv[q4 * 8 + 5]);
              dv[q4 * 8 + 6] = fmaf(w1.z, f[6], dv[q4 * 8 + 6]);
              dv[q4 * 8 + 7] = fmaf(w1.w, f[7], dv[q4 * 8 + 7]);
            }
          }
        }
        uint4* dvp = reinterpret_cast<uint4*>(bg.dv + (int64_t)gc.x * bg.dv_sb + (int64_t)tok * bg.dv_sl +
                                              gc.z * HD);
        uint4* dkp = reinterpret_cast<uint4*>(bg.dk + (int64_t)gc.x * bg.dk_sb + (int64_t)tok * bg.dk_sl +
                                              gc.z * HD);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          float f[8], h[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            f[e] = dv[q4 * 8 + e];
            h[e] = __uint_as_float(rk[q4 * 8 + e]);  // scale already folded into dS^T
          }
          dvp[q4] = pack<__nv_bfloat16>(f);
          dkp[q4] = pack<__nv_bfloat16>(h);
        }
        PROF_T(f2);
        PROF_ADDT(384, 17, f1, f2);
      }
      PROF_T(h0);
      // dQ of the whole group
      mbar_wait(&sm.dq_full[gi & 1], (gi >> 1) & 1);
      fence_after_sync();
      PROF_T(h1);
      PROF_ADDT(384, 18, h0, h1);
#pragma unroll 1
      for (int qb = 0; qb < T; ++qb) {
        uint32_t rq[32];
        tmem_ld32(lane_base + ACC_DQ + (gi & 1) * 64 + qb * 32, rq);
        tmem_wait_ld();
        const int n = qb * TILE + row;
        const int tok = gc.y + (n >> bg.ws_log2) * p.W + (n & (bg.ws - 1));
        uint4* dqp = reinterpret_cast<uint4*>(bg.dq + (int64_t)gc.x * bg.dq_sb + (int64_t)tok * bg.dq_sl +
                                              gc.z * HD);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(rq[q4 * 8 + e]);
          dqp[q4] = pack<__nv_bfloat16>(f);
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&sm.dq_empty[gi & 1]);
        mbar_arrive(&sm.grp_empty[gs]);  // every MMA of the group has completed (dq_full)
      }
      PROF_T(h2);
      PROF_ADDT(384, 19, h1, h2); PROF_ADDT(384, 20, h0, h0 + 1);
    }
  }
  fence_before_sync();
  __syncthreads();
  PROF_T(kend);
  PROF_ADDT(32, 24, kbeg, kend);
  if (warp == 2) tmem_dealloc(tmem, 512);
}

template <int NK>
int launch_bwd(int nbr, const StripeGeom* g, const TcBwdIO* io, cudaStream_t st, int wg_blocks) {
  BwdMaps maps;
  BwdParams p;
  memset(&maps, 0, sizeof(maps));
  memset(&p, 0, sizeof(p));
  p.B = g[0].B; p.W = g[0].W; p.L = g[0].L;
  p.scale = g[0].scale;
  p.scale_log2 = g[0].scale * 1.4426950408889634f;
  int gpi = 0;
  for (int i = 0; i < nbr; ++i) {
    const int bx = g[i].ws < TILE ? g[i].ws : TILE, by = TILE / bx;
    int rc;
    if ((rc = tc_make_map(&maps.q[i], io[i].q, g[i], g[i].q_sb, g[i].q_sl, bx, by)) != CSB200_OK) return rc;
    if ((rc = tc_make_map(&maps.k[i], io[i].k, g[i], g[i].k_sb, g[i].k_sl, bx, by)) != CSB200_OK) return rc;
    if ((rc = tc_make_map(&maps.v[i], io[i].v, g[i], g[i].v_sb, g[i].v_sl, bx, by)) != CSB200_OK) return rc;
    if ((rc = tc_make_map(&maps.go[i], io[i].gout, g[i], g[i].o_sb, g[i].o_sl, bx, by)) != CSB200_OK) return rc;
    BwdBranch& b = p.br[i];
    b.hs = g[i].hs; b.ws = g[i].ws; b.nwy = g[i].nwy; b.nwx = g[i].nwx; b.heads = g[i].heads; b.by = by;
    b.ws_log2 = 0;
    while ((1 << b.ws_log2) < g[i].ws) ++b.ws_log2;
    b.lepe_w = io[i].lepe_w; b.lse = io[i].lse; b.delta = io[i].delta;
    b.dq = static_cast<__nv_bfloat16*>(io[i].dq);
    b.dk = static_cast<__nv_bfloat16*>(io[i].dk);
    b.dv = static_cast<__nv_bfloat16*>(io[i].dv);
    b.dq_sb = g[i].dq_sb; b.dq_sl = g[i].dq_sl; b.dk_sb = g[i].dk_sb; b.dk_sl = g[i].dk_sl;
    b.dv_sb = g[i].dv_sb; b.dv_sl = g[i].dv_sl;
    b.drop_mask = g[i].drop_mask;
    b.wg_partial = wg_blocks > 0 ? io[i].wg_partial : nullptr;
    b.gw = io[i].gw; b.gb = io[i].gb;
    if (i == 0) p.g0 = g[i].nwy * g[i].nwx * g[i].heads;
    gpi += g[i].nwy * g[i].nwx * g[i].heads;
  }
  p.gpi = gpi;
  p.groups = p.B * gpi;
  p.keep_scale = g[0].keep_scale;
  p.wg_blocks = wg_blocks;
  p.nbr = nbr;
  const bool drop = g[0].drop_thr != 0;
  const int smem = (int)sizeof(BSmem<NK>) + 1024;  // > 113 KB: one CTA (all 512 TMEM columns) per SM
  const int sm_count = device_sm_count();
  if (sm_count <= 0) return fail(CSB200_ERR_CUDA, "stripe_bwd_tc: cannot query the SM count");
  const int grid = p.groups < sm_count ? p.groups : sm_count;
  if (drop) {
    CSB200_CUDA(opt_in_smem(reinterpret_cast<const void*>(&stripe_bwd_tc<NK, true>), smem));
    stripe_bwd_tc<NK, true><<<grid, THREADS, smem, st>>>(maps, p);
  } else {
    CSB200_CUDA(opt_in_smem(reinterpret_cast<const void*>(&stripe_bwd_tc<NK, false>), smem));
    stripe_bwd_tc<NK, false><<<grid, THREADS, smem, st>>>(maps, p);
  }
  return check_launch("stripe_bwd_tc");
}

}  // namespace

#ifdef CSB_PROF
extern "C" __attribute__((visibility("default"))) int csb200_debug_prof_bwd(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_prof_bwd, sizeof(g_prof_bwd));
  if (reset) {
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(g_prof_bwd, z, sizeof(z));
  }
  return 0;
}
#endif

int tc_bwd_multi(int nbr, const StripeGeom* g, const TcBwdIO* io, cudaStream_t st, int wg_blocks) {
  static_assert(sizeof(BSmem<128>) + 1024 > 114 * 1024 && sizeof(BSmem<256>) + 1024 > 114 * 1024,
                "two CTAs must not fit one SM");
  static_assert(sizeof(BSmem<128>) + 1024 <= 227 * 1024 && sizeof(BSmem<256>) + 1024 <= 227 * 1024,
                "shared memory budget");
  return g[0].N == 128 ? launch_bwd<128>(nbr, g, io, st, wg_blocks) : launch_bwd<256>(nbr, g, io, st, wg_blocks);
}

int tc_bwd_core(const StripeGeom& g, const void* q, const void* k, const void* v, const void* gout,
                const float* lepe_w, const float* lse, const float* delta, void* dq, void* dk,
                void* dv, cudaStream_t st) {
  const TcBwdIO io{q, k, v, gout, lepe_w, lse, delta, dq, dk, dv, nullptr, nullptr, nullptr};
  return tc_bwd_multi(1, &g, &io, st, 0);
}

}  // namespace csb200

// Stripe attention + LePE, tcgen05 / TMEM / TMA engine (bf16 in, fp32 accumulate) — forward.
//
// One persistent CTA per SM walks the (image, stripe, head) groups.  Per group the stripe's K and V
// (N x 32 bf16 each) and, per 128-row query tile, Q are fetched by TMA straight out of the packed
// token-major qkv buffer: the tensor map is (channel, x, y, image) and the box is
// (32, min(w_sp,128), 128/min(w_sp,128), 1), so the reference's img2windows / im2cswin copies
// (C:199-206, C:248-254) ARE the TMA address generation.  Tiles land 64-byte-swizzled, which is at
// once the K-major layout of Q and K for S = Q K^T and the MN-major layout of V for O = P V.
//
//   warp 0      TMA producer (K, V, LePE taps per group; Q per tile); K and V in separate rings: K is
//               released as soon as S = Q K^T has completed, V after the epilogue
//   warp 1      tcgen05.mma issuer:  S = Q K^T (M128 x N{128,256} x K32) into TMEM
//   warp 2      TMEM allocator
//   warp 3      tcgen05.mma issuer:  O = P V (M128 x N32 x K{128,256}), P read from TMEM as the A operand
//   warps 4..   NWG softmax warpgroups (3 for N=128, 2 for N=256), each on its own TMEM buffer and
//               taking every NWG-th tile, so loads, MMAs, exponentials and stores of different tiles
//               overlap (the chain TMA -> S -> softmax -> PV -> epilogue is ~4 us long): one thread per query
//               row reads its S row with tcgen05.ld, max / exp2 / sum in registers, writes bf16 P back
//               over S with tcgen05.st; later reads O, scales by 1/sum, adds the LePE depthwise 3x3
//               evaluated from the V tile that is already in shared memory (zero padding at the
//               stripe border, C:244,263-265), and stores the row to out[b, token, head*32 ..] —
//               windows2img (C:209-217) and the branch concat (C:363) are this store's address.
//
// TMEM map per buffer (N columns): S fp32 [0,N) -> P bf16 [0,N/2) -> O fp32 [N/2, N/2+32).

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "philox.cuh"
#include "stripe_attn.cuh"
#include "tc_common.cuh"

namespace csb200 {

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// cuTensorMapEncodeTiled needs a current context; a host thread that has issued no runtime work yet
// (autograd's backward thread on its first call) has none bound.
void ensure_context() {
  static thread_local bool bound = false;
  if (!bound) {
    cudaFree(nullptr);
    bound = true;
  }
}

constexpr int HD_ = 32;
static CUtensorMapL2promotion l2_promotion(int heads) {
  static const int forced = [] {
    const char* e = getenv("CSB_L2PROMO");  // experiment knob: 0 none, 1 64 B, 2 128 B, 3 256 B
    return e ? atoi(e) : -1;
  }();
  if (forced >= 0) return static_cast<CUtensorMapL2promotion>(forced);
  return heads >= 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
}
// A token row of one head is 64 B.  With >= 2 heads the other half of the 128-B line is the
// neighbouring head, which the neighbouring CTA wants at the same moment: promote to 128 B.  With a
// single head it belongs to the other branch / operand: promoting would double the DRAM traffic.
int tc_make_map(CUtensorMap* m, const void* base, const StripeGeom& g, int64_t sb, int64_t sl, int bx,
             int by) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return fail(CSB200_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
  ensure_context();
  const cuuint64_t dims[4] = {(cuuint64_t)g.heads * HD_, (cuuint64_t)g.W, (cuuint64_t)g.H,
                              (cuuint64_t)g.B};
  const cuuint64_t strides[3] = {(cuuint64_t)sl * 2, (cuuint64_t)sl * 2 * g.W, (cuuint64_t)sb * 2};
  const cuuint32_t box[4] = {HD_, (cuuint32_t)bx, (cuuint32_t)by, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                   l2_promotion(g.heads), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CSB200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return CSB200_OK;
}


#ifdef CSB_PROF
__device__ unsigned long long g_prof[32];
#define PROF_T(v) const long long v = clock64()
#define PROF_ADD(i, a, b) if (blockIdx.x == 0 && (threadIdx.x & 127) == 0) atomicAdd(&g_prof[i], (unsigned long long)((b) - (a)))
#define PROF_ADD1(i, a, b) if (blockIdx.x == 0 && threadIdx.x == 32) atomicAdd(&g_prof[i], (unsigned long long)((b) - (a)))
#else
#define PROF_T(v)
#define PROF_ADD(i, a, b)
#define PROF_ADD1(i, a, b)
#endif

namespace {
using namespace tc;

constexpr int HD = 32;
// Every POLY_MOD-th pair of exponentials can run on the FMA pipes instead of MUFU (tc_common.cuh: ex2_poly_pair).
// Measured here: no gain at N = 128 / 256 (45.5 -> 45.1 us at the stage-3 shape) — the sweep is bound by the ISSUE
// slots of the two softmax warps per scheduler (~4 instructions per element with MUFU, ~8 with the polynomial), not
// by the MUFU pipe — so it is off; the key/value-tiled kernel (stripe_attn_tc_kv.cu, three softmax warpgroups,
// compute-bound) keeps it (+4 %).
#ifndef CSB_POLY_MOD
#define CSB_POLY_MOD 0
#endif
constexpr int POLY_MOD = CSB_POLY_MOD;
constexpr int TILE = 128;                   // query rows per tile == TMEM lanes
constexpr int ROW_BYTES = HD * 2;           // 64 B per token row of one head
constexpr int TILE_BYTES = TILE * ROW_BYTES;  // 8 KB
constexpr int LEPE_FLOATS = 10 * HD;        // 9 taps + bias for the 32 channels of a head

// Geometry and outputs of one branch (orientation).  A launch covers up to two branches — the
// horizontal and the vertical stripes of one CSWinBlock (C:360-363) — whose groups are interleaved
// image by image, so the two halves of every 128-byte line of the packed qkv buffer are consumed
// close together in time and one launch fills the machine where two half-size ones left a tail.
struct FwdBranch {
  int hs, ws, ws_log2, nwy, nwx, heads, by;  // by: TMA box extent in y (bx * by == 128)
  const float* lepe_w;  // [C'][9]
  const float* lepe_b;  // [C']
  __nv_bfloat16* out;
  int64_t o_sb, o_sl;
  float* lse;
  uint32_t* drop_mask;  // attention dropout: [B][heads][L][N / 32] transposed keep bits (written here)
  uint32_t drop_salt;
};
struct FwdParams {
  int B, W, L;
  int g0, gpi;         // groups per image of branch 0 / of both branches
  int groups;          // B * gpi (pair mode: the number of TILES, two groups each)
  int real_groups;     // B * gpi
  float scale_log2;    // scale * log2(e)
  float scale;
  uint32_t drop_thr;   // attention dropout (stripe_attn.cuh); 0 in the <.., false> instantiation
  float keep_scale;
  const unsigned long long* rng;
  FwdBranch br[2];
};
struct FwdMaps {
  CUtensorMap q[2], k[2], v[2];
};

// Pipeline depths per stripe length: softmax warpgroups (== TMEM buffers), K/V ring, Q ring.
template <int NK>
struct Cfg;
template <>
struct Cfg<128> {
#ifndef CSB_KS128
#define CSB_KS128 5
#define CSB_VS128 10
#define CSB_QS128 5
#endif
  static constexpr int NWG = 3, KS = CSB_KS128, VS = CSB_VS128, QS = CSB_QS128;
  static constexpr int LWG = 0, LB = 1;   // N = 128 is bound by the operand loads, not by the softmax warps
};
template <>
struct Cfg<256> {
  // LWG: a fifth warpgroup evaluates the LePE stencil (3 100 of the 3 900 epilogue cycles of a softmax
  // warpgroup at N = 256) into shared memory while the softmax warpgroups run their exponentials; LB buffers
  static constexpr int NWG = 2, KS = 3, VS = 4, QS = 6;
  static constexpr int LWG = 1, LB = 2;
};

template <int NK>
struct Smem {
  static constexpr int KV_BYTES = NK * ROW_BYTES;
  static constexpr int NWG = Cfg<NK>::NWG, KS = Cfg<NK>::KS, VS = Cfg<NK>::VS, QS = Cfg<NK>::QS;
  alignas(1024) uint8_t q[QS][TILE_BYTES];
  alignas(1024) uint8_t k[KS][KV_BYTES];
  alignas(1024) uint8_t v[VS][KV_BYTES];
  // [half][tap][c] then bias[c]; half 1 only in the pair mode (two 64-token stripes per tile)
  alignas(16) float lepe[VS][2][LEPE_FLOATS];
  alignas(16) int4 coord[VS][2];             // (image, first token of the stripe, head, branch) per V stage (and half)
  alignas(8) uint64_t q_full[QS], q_empty[QS];
  uint64_t k_full[KS], k_empty[KS], v_full[VS], v_empty[VS];
  uint64_t s_full[NWG], p_full[NWG], o_full[NWG], buf_empty[NWG];
  uint64_t l_full[Cfg<NK>::LB], l_empty[Cfg<NK>::LB];
  uint32_t tmem_base;
  // LePE warpgroup -> epilogue: bias + depthwise 3x3 of V for the 128 rows of a tile, fp32, row-major with the
  // 16-byte chunks of row r rotated by r (a thread writes / reads its own row: conflict-free), and the tile's
  // group coordinates (the V stage may be recycled before the epilogue runs)
  alignas(16) float lepe_out[Cfg<NK>::LWG ? Cfg<NK>::LB : 1][Cfg<NK>::LWG ? TILE * HD : 4];
  alignas(16) int4 lcoord[Cfg<NK>::LB];
};

// address of 16-byte chunk `chunk` (0..3) of row n inside a 64B-swizzled tile
__device__ __forceinline__ const uint4* sw64_chunk(const uint8_t* tile, int n, int chunk) {
  return reinterpret_cast<const uint4*>(tile + n * ROW_BYTES + ((chunk ^ ((n >> 1) & 3)) << 4));
}

struct GroupCoord {
  int b, wy, wx, head, br;
};
__device__ __forceinline__ GroupCoord decode_group(const FwdParams& p, int g) {
  GroupCoord c;
  c.b = g / p.gpi;
  int r = g - c.b * p.gpi;
  c.br = r >= p.g0 ? 1 : 0;
  r -= c.br ? p.g0 : 0;
  const FwdBranch& bg = p.br[c.br];
  c.head = r % bg.heads;
  r /= bg.heads;
  c.wx = r % bg.nwx;
  c.wy = r / bg.nwx;
  return c;
}

template <int A, int B>
__device__ __forceinline__ void setmaxnreg() {  // A = 1: grow to B registers per thread, A = 0: shrink
  if constexpr (A) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(B));
  else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(B));
}

// PAIR (NK = 128 only): stripes of 64 tokens, TWO (stripe, head) groups per 128-row tile — rows / keys 0..63 are
// group 2P, 64..127 group 2P + 1 (any two groups: coordinates, taps and tensor maps are per half).  S = Q K^T is
// computed for the whole tile and only its two diagonal 64 x 64 blocks are used: a row's softmax runs over the
// two 32-column chunks of its own half, the other half of its P row is written as zeros, so O = P V needs no
// change.  (BASELINE config 5, 1024^2 at stripe width 1: stage 3 has N = 64.)
template <int NK, bool DROP, bool PAIR = false>
__global__ void __launch_bounds__(128 + 128 * (Cfg<NK>::NWG + Cfg<NK>::LWG), 1)
    stripe_fwd_tc(const __grid_constant__ FwdMaps maps, const __grid_constant__ FwdParams p) {
  static_assert(!PAIR || (NK == TILE && !DROP), "pair mode: one 128-row tile, no dropout");
  constexpr int T = NK / TILE;          // query tiles per group
  constexpr int NBOX = NK / TILE;       // TMA boxes per K (or V) load
  constexpr int NWG = Cfg<NK>::NWG, KS = Cfg<NK>::KS, VS = Cfg<NK>::VS, QS = Cfg<NK>::QS;
  constexpr bool LWG = Cfg<NK>::LWG != 0;
  constexpr int LB = Cfg<NK>::LB;
  constexpr uint32_t P_COL = 0, O_COL = NK / 2, BUF_COLS = NK;
  extern __shared__ uint8_t smem_raw[];
  // align inside the shared window: pointer + integer offset keeps the shared address space (an
  // integer -> pointer cast makes every access a generic LD/ST with 64-bit address math)
  Smem<NK>& sm = *reinterpret_cast<Smem<NK>*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_groups = (p.groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int my_tiles = my_groups * T;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < (p.gpi > p.g0 ? 2 : 1); ++i) {
      prefetch_tensormap(&maps.q[i]);
      prefetch_tensormap(&maps.k[i]);
      prefetch_tensormap(&maps.v[i]);
    }
    for (int i = 0; i < QS; ++i) {
      mbar_init(&sm.q_full[i], 1);
      mbar_init(&sm.q_empty[i], 1);
    }
    for (int i = 0; i < KS; ++i) {
      mbar_init(&sm.k_full[i], 1);
      mbar_init(&sm.k_empty[i], 1);       // tcgen05.commit after the last S = Q K^T of the group
    }
    for (int i = 0; i < VS; ++i) {
      mbar_init(&sm.v_full[i], 2);        // expect_tx arrival (before the TMA) + one after the LePE taps
      // readers of a V stage: the warps that evaluate the stencil (one arrival per warp per tile of the group)
      // and, with a LePE warpgroup, the P V MMAs themselves (a tcgen05.commit after the last tile's)
      mbar_init(&sm.v_empty[i], 4 * T + (LWG ? 1 : 0));
    }
    for (int i = 0; i < LB; ++i) {
      mbar_init(&sm.l_full[i], 4);
      mbar_init(&sm.l_empty[i], 4);
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
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  PROF_T(k0);
  // With a LePE warpgroup the CTA has 512 threads, i.e. 128 registers each at launch: the copy / issue warps and
  // the stencil warpgroup hand registers to the softmax warpgroups (two 32-word TMEM chunks in flight plus the
  // output row need ~170).  Each role's code is DOMINATED by its setmaxnreg (ptxas budgets a region by the
  // setmaxnreg that dominates it) and each warpgroup executes one common instruction.
  if (warp < 4) {
    if constexpr (LWG) setmaxnreg<0, 40>();
  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // ONE producer warp: unlike the backward kernel (four boxes per group: two producer warps help there), the
    // forward pass measured slower with a second producer, whether the warps split the operands (97 us at the
    // stage-1 shape) or alternate groups (90 us) — 85 us with one.
    // K lives only until its S = Q K^T has been issued and completed; V (with the LePE taps) until the epilogue:
    // separate rings, so K and Q run far ahead of the tiles still in the softmax warpgroups.
    // LePE taps of a head ([tap][c], bias last): 10 values per lane, fetched ONE GROUP AHEAD — the loads are an
    // L2 round trip that used to sit, un-overlapped, between the TMA issues of consecutive groups (101 -> 85 us).
    constexpr int TAPS_PER_LANE = LEPE_FLOATS / 32;
    float taps[TAPS_PER_LANE];
    auto fetch_taps = [&](const GroupCoord& gc) {
      const FwdBranch& b2 = p.br[gc.br];
#pragma unroll
      for (int j = 0; j < TAPS_PER_LANE; ++j)  // element lane + 32 j: tap j (HD == 32), channel = lane
        taps[j] = j < 9 ? __ldg(b2.lepe_w + (gc.head * HD + lane) * 9 + j) : __ldg(b2.lepe_b + gc.head * HD + lane);
    };
    // pair mode: tile index P holds groups 2P and 2P + 1 (the last tile of an odd count holds its group twice:
    // both halves then compute and store the same values)
    auto half_group = [&](int tile, int h) {
      const int g = 2 * tile + h;
      return decode_group(p, g < p.real_groups ? g : p.real_groups - 1);
    };
    float taps1[PAIR ? TAPS_PER_LANE : 1];
    auto fetch_taps1 = [&](const GroupCoord& gc) {
      const FwdBranch& b2 = p.br[gc.br];
#pragma unroll
      for (int j = 0; j < (PAIR ? TAPS_PER_LANE : 1); ++j)
        taps1[j] = j < 9 ? __ldg(b2.lepe_w + (gc.head * HD + lane) * 9 + j) : __ldg(b2.lepe_b + gc.head * HD + lane);
    };
    if (my_groups > 0) {
      if constexpr (PAIR) {
        fetch_taps(half_group((int)blockIdx.x, 0));
        fetch_taps1(half_group((int)blockIdx.x, 1));
      } else {
        fetch_taps(decode_group(p, (int)blockIdx.x));
      }
    }
    int it = 0;
    for (int gi = 0; gi < my_groups; ++gi) {
      const int tile = (int)blockIdx.x + gi * (int)gridDim.x;
      const GroupCoord c = PAIR ? half_group(tile, 0) : decode_group(p, tile);
      const GroupCoord c1 = PAIR ? half_group(tile, 1) : c;
      const FwdBranch& bg = p.br[c.br];
      const int ks = gi % KS, vs = gi % VS;
      mbar_wait(&sm.k_empty[ks], ((gi / KS) & 1) ^ 1);
      mbar_wait(&sm.v_empty[vs], ((gi / VS) & 1) ^ 1);
      if (lane == 0) {
        mbar_expect_tx(&sm.k_full[ks], Smem<NK>::KV_BYTES);
        mbar_expect_tx(&sm.v_full[vs], Smem<NK>::KV_BYTES);
        if constexpr (PAIR) {
          const int qs = it % QS;
          mbar_wait(&sm.q_empty[qs], ((it / QS) & 1) ^ 1);
          mbar_expect_tx(&sm.q_full[qs], TILE_BYTES);
#pragma unroll
          for (int h = 0; h < 2; ++h) {  // 64-row boxes: the whole stripe of each half
            const GroupCoord& ch = h ? c1 : c;
            const FwdBranch& bh = p.br[ch.br];
            const int x0 = ch.wx * bh.ws, y0 = ch.wy * bh.hs, off = h * (TILE_BYTES / 2);
            tma_load_4d(sm.k[ks] + off, &maps.k[ch.br], &sm.k_full[ks], ch.head * HD, x0, y0, ch.b);
            tma_load_4d(sm.v[vs] + off, &maps.v[ch.br], &sm.v_full[vs], ch.head * HD, x0, y0, ch.b);
            tma_load_4d(sm.q[qs] + off, &maps.q[ch.br], &sm.q_full[qs], ch.head * HD, x0, y0, ch.b);
          }
          ++it;
        } else {
        const int x0 = c.wx * bg.ws, y0 = c.wy * bg.hs;
#pragma unroll
        for (int bx = 0; bx < NBOX; ++bx) {
          // box `bx` covers in-stripe rows [128 bx, 128 bx + 128)
          const int dx = (bg.ws > TILE) ? (bx * TILE) % bg.ws : 0;
          const int dy = (bg.ws > TILE) ? (bx * TILE) / bg.ws : bx * bg.by;
          tma_load_4d(sm.k[ks] + bx * TILE_BYTES, &maps.k[c.br], &sm.k_full[ks], c.head * HD,
                      x0 + dx, y0 + dy, c.b);
          tma_load_4d(sm.v[vs] + bx * TILE_BYTES, &maps.v[c.br], &sm.v_full[vs], c.head * HD,
                      x0 + dx, y0 + dy, c.b);
        }
        for (int t = 0; t < T; ++t, ++it) {
          const int qs = it % QS;
          mbar_wait(&sm.q_empty[qs], ((it / QS) & 1) ^ 1);
          mbar_expect_tx(&sm.q_full[qs], TILE_BYTES);
          const int dx = (bg.ws > TILE) ? (t * TILE) % bg.ws : 0;
          const int dy = (bg.ws > TILE) ? (t * TILE) / bg.ws : t * bg.by;
          tma_load_4d(sm.q[qs], &maps.q[c.br], &sm.q_full[qs], c.head * HD, x0 + dx, y0 + dy, c.b);
        }
        }
      }
      // taps of this head -> smem; plain stores, released by the second arrival on v_full
#pragma unroll
      for (int j = 0; j < TAPS_PER_LANE; ++j) sm.lepe[vs][0][lane + 32 * j] = taps[j];
      if constexpr (PAIR) {
#pragma unroll
        for (int j = 0; j < TAPS_PER_LANE; ++j) sm.lepe[vs][1][lane + 32 * j] = taps1[j];
      }
      if (lane == 0) {
        sm.coord[vs][0] = make_int4(c.b, (c.wy * bg.hs) * p.W + c.wx * bg.ws, c.head, c.br);
        if constexpr (PAIR) {
          const FwdBranch& b1 = p.br[c1.br];
          sm.coord[vs][1] = make_int4(c1.b, (c1.wy * b1.hs) * p.W + c1.wx * b1.ws, c1.head, c1.br);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.v_full[vs]);
      if (gi + 1 < my_groups) {
        const int nt = (int)blockIdx.x + (gi + 1) * (int)gridDim.x;
        if constexpr (PAIR) {
          fetch_taps(half_group(nt, 0));
          fetch_taps1(half_group(nt, 1));
        } else {
          fetch_taps(decode_group(p, nt));
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer 1: S = Q K^T ================================
    // Two issuing warps (S here, PV in warp 3): a single in-order issuer held PV(it-2) back while it
    // waited for the Q/K/V tiles of tile `it` (measured: the softmax warps then waited ~2400 cycles
    // for O).  The whole warp walks the loop; one elected lane issues (tc_common.cuh).
    constexpr uint32_t idesc_s = umma_idesc_bf16(NK, false, false);
    const uint32_t q_lo0 = desc_lo_sw64(smem_u32(sm.q[0])), k_lo0 = desc_lo_sw64(smem_u32(sm.k[0]));
    for (int it = 0; it < my_tiles; ++it) {
      const int buf = it % NWG, qs = it % QS, gi = it / T, ks = gi % KS;
      PROF_T(m0);
      mbar_wait(&sm.q_full[qs], (it / QS) & 1);
      if (it % T == 0) mbar_wait(&sm.k_full[ks], (gi / KS) & 1);
      PROF_T(m1);
      mbar_wait(&sm.buf_empty[buf], ((it / NWG) & 1) ^ 1);
      fence_after_sync();
      PROF_T(m2);
      PROF_ADD1(9, m0, m1); PROF_ADD1(10, m1, m2);
      if (elect_one_sync()) {
        const uint32_t q_lo = q_lo0 + qs * (TILE_BYTES >> 4), k_lo = k_lo0 + ks * (Smem<NK>::KV_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)  // 16 channels per step: 32 B inside the swizzled row
          umma_ss2(tmem + buf * BUF_COLS, q_lo + k * (32 >> 4), DESC_HI_SW64, k_lo + k * (32 >> 4),
                   DESC_HI_SW64, idesc_s, k > 0);
        umma_commit(&sm.s_full[buf]);
        umma_commit(&sm.q_empty[qs]);
        if (it % T == T - 1) umma_commit(&sm.k_empty[ks]);  // every S of the group has read K
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    // ================================ MMA issuer 2: O = P V =================================
    constexpr uint32_t idesc_pv = umma_idesc_bf16(HD, false, true);
    const uint32_t v_lo0 = desc_lo_sw64(smem_u32(sm.v[0]));
    for (int it = 0; it < my_tiles; ++it) {
      const int buf = it % NWG, gi = it / T, vs = gi % VS;
      if (it % T == 0) mbar_wait(&sm.v_full[vs], (gi / VS) & 1);
      mbar_wait(&sm.p_full[buf], (it / NWG) & 1);
      fence_after_sync();
      if (elect_one_sync()) {
        const uint32_t d = tmem + buf * BUF_COLS + O_COL, a = tmem + buf * BUF_COLS + P_COL;
        const uint32_t v_lo = v_lo0 + vs * (Smem<NK>::KV_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < NK / 16; ++k)  // 16 keys per step: 8 TMEM columns of P, 1024 B of V
          umma_ts2(d, a + 8 * k, v_lo + k * (1024 >> 4), DESC_HI_SW64, idesc_pv, k > 0);
        umma_commit(&sm.o_full[buf]);
        if (LWG && it % T == T - 1) umma_commit(&sm.v_empty[vs]);  // every P V of the group has read V
      }
      __syncwarp();
    }
  }
  } else if (LWG && warp >= 4 + 4 * NWG) {
    if constexpr (LWG) setmaxnreg<0, 96>();
    // ================================== LePE warpgroup ==================================
    // bias + depthwise 3x3 of V (zero padding at the stripe border, C:244,263-265) for the 128 rows of every
    // tile, from the V tile in shared memory into lepe_out[it % LB]
    const int row = ((warp & 3) << 5) | lane;
    for (int it = 0; it < my_tiles; ++it) {
      const int gi = it / T, t = it % T, vs = gi % VS, lb = it % LB;
      mbar_wait(&sm.v_full[vs], (gi / VS) & 1);
      mbar_wait(&sm.l_empty[lb], ((it / LB) & 1) ^ 1);
      const int4 gc = sm.coord[vs][0];
      const FwdBranch& bg = p.br[gc.w];
      const int n = t * TILE + row;
      const int yy = n >> bg.ws_log2, xx = n & (bg.ws - 1);
      float o[HD];
      const float* lw = sm.lepe[vs][0];
#pragma unroll
      for (int cc = 0; cc < HD; ++cc) o[cc] = lw[9 * HD + cc];
      const uint8_t* vt = sm.v[vs];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int ny = yy + ky - 1;
        if (ny < 0 || ny >= bg.hs) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int nx = xx + kx - 1;
          if (nx < 0 || nx >= bg.ws) continue;
          const int nn = (ny << bg.ws_log2) + nx;
          const float* wt = lw + (ky * 3 + kx) * HD;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            float f[8];
            unpack<__nv_bfloat16>(*sw64_chunk(vt, nn, q4), f);
            const float4 w0 = *reinterpret_cast<const float4*>(wt + q4 * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(wt + q4 * 8 + 4);
            o[q4 * 8 + 0] = fmaf(w0.x, f[0], o[q4 * 8 + 0]);
            o[q4 * 8 + 1] = fmaf(w0.y, f[1], o[q4 * 8 + 1]);
            o[q4 * 8 + 2] = fmaf(w0.z, f[2], o[q4 * 8 + 2]);
            o[q4 * 8 + 3] = fmaf(w0.w, f[3], o[q4 * 8 + 3]);
            o[q4 * 8 + 4] = fmaf(w1.x, f[4], o[q4 * 8 + 4]);
            o[q4 * 8 + 5] = fmaf(w1.y, f[5], o[q4 * 8 + 5]);
            o[q4 * 8 + 6] = fmaf(w1.z, f[6], o[q4 * 8 + 6]);
            o[q4 * 8 + 7] = fmaf(w1.w, f[7], o[q4 * 8 + 7]);
          }
        }
      }
      float4* dst = reinterpret_cast<float4*>(sm.lepe_out[lb] + row * HD);
#pragma unroll
      for (int c = 0; c < 8; ++c) dst[(c + row) & 7] = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
      if (row == 0) sm.lcoord[lb] = gc;
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&sm.l_full[lb]);
        mbar_arrive(&sm.v_empty[vs]);  // this warp is done with V / the taps of the group's tile
      }
    }
  } else {
    if constexpr (LWG) setmaxnreg<1, 184>();
    // ============================ softmax + epilogue warpgroups ============================
    const int wg = (warp - 4) >> 2;                  // warpgroup == TMEM buffer
    const int row = ((warp & 3) << 5) | lane;        // query row inside the tile == TMEM lane
    const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) << 5) << 16) + wg * BUF_COLS;
    constexpr int NCH = NK / 32;
    for (int it = wg; it < my_tiles; it += NWG) {
      const int gi = it / T, t = it % T, vs = gi % VS;
      const uint32_t use = (it / NWG) & 1;
      PROF_T(t0);
      mbar_wait(&sm.s_full[wg], use);
      fence_after_sync();
      PROF_T(t1);
      // Both sweeps double-buffer the TMEM reads: chunk ch+1 is in flight while chunk ch is consumed.
      uint32_t ra[32], rb[32];
      float mm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent chains (one was 16 dependent FMNMX3 per chunk)
      // pair mode: this row belongs to half `hh` of the tile and attends to the keys of that half only — the
      // two 32-column chunks 2 hh, 2 hh + 1 (warp-uniform: a warp's 32 rows lie in one half)
      const int hh = PAIR ? (row >> 6) : 0;
      tmem_ld32(lane_base, ra);
      tmem_wait_ld();
#pragma unroll
      for (int ch = 0; ch < NCH; ch += 2) {
        tmem_ld32(lane_base + (ch + 1) * 32, rb);
        if (!PAIR || (ch >> 1) == hh) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mm[i & 3] = fmaxf(mm[i & 3], __uint_as_float(ra[i]));
        }
        tmem_wait_ld();
        if (ch + 2 < NCH) tmem_ld32(lane_base + (ch + 2) * 32, ra);
        if (!PAIR || (ch >> 1) == hh) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mm[i & 3] = fmaxf(mm[i & 3], __uint_as_float(rb[i]));
        }
        tmem_wait_ld();
      }
      const float m = fmaxf(fmaxf(mm[0], mm[1]), fmaxf(mm[2], mm[3]));
      const float neg_m = -m * p.scale_log2;
      const f2_t scale2 = f2_splat(p.scale_log2), negm2 = f2_splat(neg_m);
      PROF_T(t2);
      float l0 = 0.f, l1 = 0.f;
      // attention dropout (C:290): the stripe coordinates of the group are needed before the epilogue
      DropRng rng;
      uint32_t unit = 0;
      uint32_t* mask_row = nullptr;  // this group's mask rows, offset to the word of this warp's 32 queries
      int d_tok0 = 0, d_wsl = 0, d_ws1 = 0, d_nw = 0;
      if constexpr (DROP) {
        mbar_wait(&sm.v_full[vs], (gi / VS) & 1);  // sm.coord rides on the V barrier
        const int4 gc = sm.coord[vs][0];
        const FwdBranch& bg = p.br[gc.w];
        rng = drop_rng_load(p.rng);
        const int y0 = gc.y / p.W, x0 = gc.y - y0 * p.W;
        unit = ((uint32_t)(((gc.x * bg.nwy + y0 / bg.hs) * bg.nwx + x0 / bg.ws) * bg.heads + gc.z) << 1) |
               (bg.drop_salt & 1u);
        d_nw = NK / 32;
        mask_row = bg.drop_mask + ((int64_t)gc.x * bg.heads + gc.z) * p.L * d_nw + (t * TILE + ((warp & 3) << 5)) / 32;
        d_tok0 = gc.y; d_wsl = bg.ws_log2; d_ws1 = bg.ws - 1;
      }
      auto exp_chunk = [&](const uint32_t (&r)[32], int ch) {
        uint32_t pk[16];
        if (PAIR && (ch >> 1) != hh) {  // keys of the other stripe: P = 0
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0u;
          tmem_st16(lane_base + P_COL + ch * 16, pk);
          return;
        }
        uint32_t kw = 0xffffffffu;  // keep bits of keys 32 ch .. 32 ch + 31 for this query row
        if constexpr (DROP) {
          const uint32_t thr4 = p.drop_thr * 0x01010101u;
          kw = 0u;
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            const uint4 rb = drop_bytes(rng, unit, (uint32_t)(t * TILE + row), (uint32_t)(2 * ch + hb));
            const uint32_t wds[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)  // byte >= threshold -> one keep bit per byte, gathered into a nibble
              kw |= ((((__vcmpgeu4(wds[q], thr4) & 0x01010101u) * 0x01020408u) >> 24) & 0xfu) << (16 * hb + 4 * q);
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          // x = s * scale * log2(e) - max, as a packed pair; every POLY_MOD-th pair takes the FMA-pipe 2^x
          const f2_t x2 = f2_fma(f2_make(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), scale2, negm2);
          float p0, p1;
          if (POLY_MOD > 0 && i % (POLY_MOD > 0 ? POLY_MOD : 1) == POLY_MOD - 1) {
            ex2_poly_pair(x2, p0, p1);
          } else {
            float x0, x1;
            f2_split(x2, x0, x1);
            p0 = ex2(x0);
            p1 = ex2(x1);
          }
          l0 += p0;  // the normalisation runs over ALL probabilities
          l1 += p1;
          if constexpr (DROP) {
            p0 = (kw >> (2 * i)) & 1u ? p0 : 0.f;
            p1 = (kw >> (2 * i + 1)) & 1u ? p1 : 0.f;
          }
          pk[i] = pack_bf16x2(p0, p1);
        }
        tmem_st16(lane_base + P_COL + ch * 16, pk);  // P over S columns that were already consumed
        if constexpr (DROP) {
          // transposed mask for the backward pass: the warp's 32 decisions for key j are one ballot word
          uint32_t mine = 0u;
#pragma unroll
          for (int jl = 0; jl < 32; ++jl) {
            const uint32_t bal = __ballot_sync(0xffffffffu, (kw >> jl) & 1u);
            mine = lane == jl ? bal : mine;
          }
          const int j = ch * 32 + lane;
          const int tokj = d_tok0 + (j >> d_wsl) * p.W + (j & d_ws1);
          mask_row[(int64_t)tokj * d_nw] = mine;
        }
      };
      tmem_ld32(lane_base, ra);
      tmem_wait_ld();
#pragma unroll
      for (int ch = 0; ch < NCH; ch += 2) {
        tmem_ld32(lane_base + (ch + 1) * 32, rb);
        exp_chunk(ra, ch);
        tmem_wait_ld();
        if (ch + 2 < NCH) tmem_ld32(lane_base + (ch + 2) * 32, ra);
        exp_chunk(rb, ch + 1);
        tmem_wait_ld();
      }
      const float l = l0 + l1;
      uint32_t (&r)[32] = ra;
      tmem_wait_st();
      fence_before_sync();
      mbar_arrive(&sm.p_full[wg]);
      PROF_T(t3);

      // ---- epilogue: O / l + LePE -> out, lse ----
      mbar_wait(&sm.o_full[wg], use);
      fence_after_sync();
      PROF_T(t4);
      tmem_ld32(lane_base + O_COL, r);
      tmem_wait_ld();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.buf_empty[wg]);  // S(it+2) may now overwrite this buffer

      float o[HD];
      int4 gc;
      const float inv_l = (DROP ? p.keep_scale : 1.f) / l;  // dropout: survivors are scaled by 1 / (1 - p)
      PROF_T(t4a);
      const int n = PAIR ? (row & 63) : t * TILE + row;  // in-stripe index
      if constexpr (LWG) {
        // the LePE term of this tile comes from the LePE warpgroup (chunks of row r rotated by r)
        const int lb = it % LB;
        mbar_wait(&sm.l_full[lb], (it / LB) & 1);
        gc = sm.lcoord[lb];
        const float4* src = reinterpret_cast<const float4*>(sm.lepe_out[lb] + row * HD);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 lp = src[(c + row) & 7];
          o[4 * c + 0] = fmaf(__uint_as_float(r[4 * c + 0]), inv_l, lp.x);
          o[4 * c + 1] = fmaf(__uint_as_float(r[4 * c + 1]), inv_l, lp.y);
          o[4 * c + 2] = fmaf(__uint_as_float(r[4 * c + 2]), inv_l, lp.z);
          o[4 * c + 3] = fmaf(__uint_as_float(r[4 * c + 3]), inv_l, lp.w);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.l_empty[lb]);
      } else {
      // V and the LePE taps were written by TMA / the producer warp: acquire them through the same
      // barrier the PV issuer used (already complete; cannot advance before this warp's v_empty)
      mbar_wait(&sm.v_full[vs], (gi / VS) & 1);
      gc = sm.coord[vs][hh];  // image, first token of the stripe, head, branch
      const FwdBranch& bgl = p.br[gc.w];
      const int yy = n >> bgl.ws_log2, xx = n & (bgl.ws - 1);
      const float* lw = sm.lepe[vs][hh];
#pragma unroll
      for (int cc = 0; cc < HD; ++cc) o[cc] = fmaf(__uint_as_float(r[cc]), inv_l, lw[9 * HD + cc]);
      const uint8_t* vt = sm.v[vs];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int ny = yy + ky - 1;
        if (ny < 0 || ny >= bgl.hs) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int nx = xx + kx - 1;
          if (nx < 0 || nx >= bgl.ws) continue;
          const int nn = (ny << bgl.ws_log2) + nx + (PAIR ? 64 * hh : 0);  // row of the V tile
          const float* wt = lw + (ky * 3 + kx) * HD;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            float f[8];
            unpack<__nv_bfloat16>(*sw64_chunk(vt, nn, q4), f);
            const float4 w0 = *reinterpret_cast<const float4*>(wt + q4 * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(wt + q4 * 8 + 4);
            o[q4 * 8 + 0] = fmaf(w0.x, f[0], o[q4 * 8 + 0]);
            o[q4 * 8 + 1] = fmaf(w0.y, f[1], o[q4 * 8 + 1]);
            o[q4 * 8 + 2] = fmaf(w0.z, f[2], o[q4 * 8 + 2]);
            o[q4 * 8 + 3] = fmaf(w0.w, f[3], o[q4 * 8 + 3]);
            o[q4 * 8 + 4] = fmaf(w1.x, f[4], o[q4 * 8 + 4]);
            o[q4 * 8 + 5] = fmaf(w1.y, f[5], o[q4 * 8 + 5]);
            o[q4 * 8 + 6] = fmaf(w1.z, f[6], o[q4 * 8 + 6]);
            o[q4 * 8 + 7] = fmaf(w1.w, f[7], o[q4 * 8 + 7]);
          }
        }
      }
      }
      const FwdBranch& bg = p.br[gc.w];
      const int yy = n >> bg.ws_log2, xx = n & (bg.ws - 1);
      PROF_T(t4b);
      const int tok = gc.y + yy * p.W + xx;
      uint4* dst = reinterpret_cast<uint4*>(bg.out + (int64_t)gc.x * bg.o_sb + (int64_t)tok * bg.o_sl +
                                            gc.z * HD);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = o[q4 * 8 + e];
        dst[q4] = pack<__nv_bfloat16>(f);
      }
      bg.lse[((int64_t)gc.x * bg.heads + gc.z) * p.L + tok] = m * p.scale + __logf(l);
      __syncwarp();
      if (!LWG && lane == 0) mbar_arrive(&sm.v_empty[vs]);  // this warp is done with V / LePE of the group
      PROF_T(t5);
      PROF_ADD(0, t0, t1); PROF_ADD(1, t1, t2); PROF_ADD(2, t2, t3); PROF_ADD(3, t3, t4); PROF_ADD(4, t4, t5);
      PROF_ADD(5, t0, t0 + 1);
      PROF_ADD(12, t4, t4a); PROF_ADD(13, t4a, t4b); PROF_ADD(14, t4b, t5);
    }
  }
  // teardown
  fence_before_sync();
  __syncthreads();
  PROF_T(k1);
  PROF_ADD1(11, k0, k1);
  if (warp == 2) tmem_dealloc(tmem, 512);
}

template <int NK, bool PAIR = false>
int launch_fwd(int nbr, const StripeGeom* g, const TcFwdIO* io, cudaStream_t st) {
  constexpr int ROWS = PAIR ? TILE / 2 : TILE;  // token rows of one TMA box
  FwdMaps maps;
  FwdParams p;
  memset(&maps, 0, sizeof(maps));
  memset(&p, 0, sizeof(p));
  p.B = g[0].B; p.W = g[0].W; p.L = g[0].L;
  p.scale = g[0].scale;
  p.scale_log2 = g[0].scale * 1.4426950408889634f;
  int gpi = 0;
  for (int i = 0; i < nbr; ++i) {
    const int bx = g[i].ws < ROWS ? g[i].ws : ROWS, by = ROWS / bx;
    int rc;
    if ((rc = tc_make_map(&maps.q[i], io[i].q, g[i], g[i].q_sb, g[i].q_sl, bx, by)) != CSB200_OK) return rc;
    if ((rc = tc_make_map(&maps.k[i], io[i].k, g[i], g[i].k_sb, g[i].k_sl, bx, by)) != CSB200_OK) return rc;
    if ((rc = tc_make_map(&maps.v[i], io[i].v, g[i], g[i].v_sb, g[i].v_sl, bx, by)) != CSB200_OK) return rc;
    FwdBranch& b = p.br[i];
    b.hs = g[i].hs; b.ws = g[i].ws; b.nwy = g[i].nwy; b.nwx = g[i].nwx; b.heads = g[i].heads; b.by = by;
    b.ws_log2 = 0;
    while ((1 << b.ws_log2) < g[i].ws) ++b.ws_log2;
    b.lepe_w = io[i].lepe_w; b.lepe_b = io[i].lepe_b;
    b.out = static_cast<__nv_bfloat16*>(io[i].out);
    b.o_sb = g[i].o_sb; b.o_sl = g[i].o_sl;
    b.lse = io[i].lse;
    b.drop_mask = g[i].drop_mask;
    b.drop_salt = g[i].drop_salt;
    if (i == 0) p.g0 = g[i].nwy * g[i].nwx * g[i].heads;
    gpi += g[i].nwy * g[i].nwx * g[i].heads;
  }
  p.gpi = gpi;
  p.real_groups = p.B * gpi;
  p.groups = PAIR ? (p.real_groups + 1) / 2 : p.real_groups;
  p.drop_thr = g[0].drop_thr;
  p.keep_scale = g[0].keep_scale;
  p.rng = g[0].rng;
  const bool drop = p.drop_thr != 0;

  // > half of the 227 KB so that exactly one CTA (which owns all 512 TMEM columns) fits per SM
  const int smem = (int)sizeof(Smem<NK>) + 1024 > 120 * 1024 ? (int)sizeof(Smem<NK>) + 1024 : 120 * 1024;
  const int sm_count = device_sm_count();
  if (sm_count <= 0) return fail(CSB200_ERR_CUDA, "stripe_fwd_tc: cannot query the SM count");
  const int grid = p.groups < sm_count ? p.groups : sm_count;
  constexpr int THREADS_FWD = 128 + 128 * (Cfg<NK>::NWG + Cfg<NK>::LWG);
  if constexpr (PAIR) {
    if (drop) return fail(CSB200_ERR_UNSUPPORTED, "stripe_fwd_tc: no attention dropout in the pair mode");
    CSB200_CUDA(opt_in_smem(reinterpret_cast<const void*>(&stripe_fwd_tc<NK, false, true>), smem));
    stripe_fwd_tc<NK, false, true><<<grid, THREADS_FWD, smem, st>>>(maps, p);
  } else if (drop) {
    CSB200_CUDA(opt_in_smem(reinterpret_cast<const void*>(&stripe_fwd_tc<NK, true>), smem));
    stripe_fwd_tc<NK, true><<<grid, THREADS_FWD, smem, st>>>(maps, p);
  } else {
    CSB200_CUDA(opt_in_smem(reinterpret_cast<const void*>(&stripe_fwd_tc<NK, false>), smem));
    stripe_fwd_tc<NK, false><<<grid, THREADS_FWD, smem, st>>>(maps, p);
  }
  return check_launch("stripe_fwd_tc");
}

}  // namespace

// Shapes the tcgen05 engine tiles: bf16, stripes of exactly 128 or 256 tokens whose width divides
// (or is a multiple of) 128, so that a 128-row tile is a rectangular TMA box.
bool tc_single_pass_supported(const StripeGeom& g, int dtype) {
  if (dtype != CSB200_BF16) return false;
  if (g.N != 128 && g.N != 256) return false;
  if (!((g.ws <= TILE && TILE % g.ws == 0) || (g.ws % TILE == 0))) return false;
  if (g.ws > 256 || g.hs > 256) return false;
  return true;
}
// stripes of 64 tokens (two per tile, forward only, no dropout): the stripe is one 64-row TMA box
bool tc_fwd_pair_supported(const StripeGeom& g, int dtype) {
  if (dtype != CSB200_BF16 || g.N != 64 || g.drop_thr != 0) return false;
  return g.ws <= 64 && 64 % g.ws == 0 && g.hs <= 256;
}
// forward: the single-pass kernels here (incl. the pair mode), or the key/value-tiled kernel for long stripes
// (stripe_attn_tc_kv.cu)
bool tc_fwd_supported(const StripeGeom& g, int dtype) {
  return tc_single_pass_supported(g, dtype) || tc_fwd_pair_supported(g, dtype) || tc_fwd_kv_supported(g, dtype);
}
bool tc_bwd_supported(const StripeGeom& g, int dtype) { return tc_single_pass_supported(g, dtype); }

#ifdef CSB_PROF
extern "C" __attribute__((visibility("default"))) int csb200_debug_prof_fwd(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_prof, sizeof(g_prof));
  if (reset) {
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(g_prof, z, sizeof(z));
  }
  return 0;
}
#endif

int tc_fwd_multi(int nbr, const StripeGeom* g, const TcFwdIO* io, cudaStream_t st) {
  return g[0].N == 128 ? launch_fwd<128>(nbr, g, io, st) : launch_fwd<256>(nbr, g, io, st);
}

int tc_fwd(const StripeGeom& g, const void* q, const void* k, const void* v, const float* lepe_w,
           const float* lepe_b, void* out, float* lse, cudaStream_t st) {
  const TcFwdIO io{q, k, v, lepe_w, lepe_b, out, lse};
  if (tc_single_pass_supported(g, CSB200_BF16)) return tc_fwd_multi(1, &g, &io, st);
  if (g.N == 64) return launch_fwd<128, true>(1, &g, &io, st);
  return tc_fwd_kv(g, q, k, v, lepe_w, lepe_b, out, lse, st);
}

}  // namespace csb200

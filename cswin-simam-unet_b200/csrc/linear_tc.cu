// Token-path Linear on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), bf16 in, fp32 accumulate:
//
//     Y = act(X W^T + b)        X (M, K) tokens, W (N, K) = nn.Linear.weight, b (N) fp32, Y (M, N)
//
// for the GEMMs of the CSWinBlock whose contraction is the block width (K = C in {64, 128, 256}):
// `qkv` (C:357-358, N = 3C), `proj` (C:366, N = C) and `Mlp.fc1` + `act` (C:188-196, N = 4C, exact-erf GELU
// in the epilogue, optionally also storing the pre-activation h that GELU' needs in backward).  M is the
// token count of the batch (524 288 ... 32 768 at 512^2, batch 32), so these GEMMs are HBM- / epilogue-bound,
// not tensor-bound, and the design follows from that:
//
//   * one persistent CTA per SM owns ONE n-tile of the weight (BN <= 256 output columns x K) and keeps it
//     RESIDENT in shared memory for its whole life (<= 128 KB, 128-byte-swizzled K-major), so the only
//     operand that streams is X: 16-KB TMA boxes (128 tokens x 64 channels) through a ring;
//   * warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (M128 x N{BN} x K16, smem x smem, accumulators in
//     TMEM, two buffers of 256 columns so the MMAs of tile i+1 run under the epilogue of tile i), warp 2 =
//     TMEM allocator, then FOUR epilogue warpgroups: warpgroup j drains column half (j >> 1) of TMEM buffer
//     (j & 1) — 16 warps keep tcgen05.ld / MUFU latency covered; one thread owns one token row, adds the
//     bias, applies GELU on packed fp32 pairs (gelu_math.cuh, the same arithmetic as the flat GELU pass);
//     a warp stages its 32 rows x 32 columns in shared memory and one lane hands the box to the TMA
//     store engine (thread-per-row 16-byte global stores — half a sector each, 32 lines per instruction —
//     ran the bias-only GEMM at 0.5x cuBLAS: profiles/r2_linear_bench_v1_direct_stores.jsonl).
//
// The separate GELU pass (1 read + 1 write of the 4C-wide hidden tensor) disappears: fc1 + GELU moves
// M (K + N [+ N]) elements instead of M (K + 3N).

#include <cstring>

#include "gelu_math.cuh"
#include "tc_common.cuh"

namespace csb200 {
namespace {
using namespace tc;

constexpr int BM = 128;                       // token rows per tile == TMEM lanes
constexpr int BK = 64;                        // channels per k-chunk: one 128-byte swizzled row
constexpr int A_STAGE_BYTES = BM * BK * 2;    // 16 KB
constexpr int MAX_STAGES = 8;
constexpr int NUM_EPI_WG = 4;
constexpr int THREADS = 128 + 128 * NUM_EPI_WG;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int STG_BYTES = 32 * 64;            // one staged store box: 32 rows x 32 bf16 columns

struct LinParams {
  int M, N, K;
  int BN, n_tiles, m_tiles, kchunks, stages;
  int m_stride;            // CTAs that share an n-tile (stride of the m-tile walk)
  int stage_bufs;          // staging buffers per epilogue warp (1 or 2), 2 KB each
  int w_mn;                // weight operand is MN-major: memory [K][N] (the input-gradient form y = x W)
  float* partial;          // EPI_DGELU: per-CTA column sums of the output, [m_stride][N]
  uint32_t idesc;
  uint32_t w_bytes;        // resident weight tile
  const float* bias;       // [N] or nullptr
  __nv_bfloat16* y;        // [M][N]
  __nv_bfloat16* h;        // [M][N] pre-activation (EPI_GELU_SAVE) or nullptr
};
struct LinMaps {
  CUtensorMap a, w;
  CUtensorMap y, h;        // stores: box 32 columns x 32 rows, 64-byte swizzle (rows past M are clipped)
};

// EPI_GELU_SAVE_D / EPI_DMUL: the pair that stores GELU'(h) (bf16) in forward and multiplies by it in backward
enum { EPI_BIAS = 0, EPI_GELU = 1, EPI_GELU_SAVE = 2, EPI_DGELU = 3, EPI_GELU_SAVE_D = 4, EPI_DMUL = 5 };

// shared -> global tile store (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::
                   "l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {  // <= N groups of this thread still reading shared memory
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// shared-memory carve-up (dynamic; offsets from a 1024-byte aligned base):
//   [0, w_bytes)                       weight tile: kchunks x (BN rows x 128 B), 128B-swizzled K-major
//   [w_bytes, + stages * 16 KB)        X ring
//   then 16 warps x stage_bufs x 2 KB  store staging (32 rows x 64 B, 64-byte swizzled: what the output
//                                      tensor maps read), bias (BN floats) and the mbarriers
struct Bars {
  uint64_t w_full;
  uint64_t a_full[MAX_STAGES], a_empty[MAX_STAGES];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t h_full[4 * NUM_EPI_WG];   // EPI_DGELU: one per epilogue warp (its staged pre-activation box)
  uint32_t tmem_base;
};

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1)
    linear_tc_kernel(const __grid_constant__ LinMaps maps, const __grid_constant__ LinParams p) {
  constexpr bool SAVE = EPI == EPI_GELU_SAVE || EPI == EPI_GELU_SAVE_D;  // forward: a second output tensor
  constexpr bool DG = EPI == EPI_DGELU || EPI == EPI_DMUL;               // input-gradient form of the Mlp
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw) + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t w_sm = base, a_sm = base + p.w_bytes;
  const uint32_t stg_off = p.w_bytes + p.stages * A_STAGE_BYTES;
  // 16 epilogue warps x stage_bufs output boxes, then (EPI_DGELU) 16 pre-activation boxes
  const uint32_t hstg_off = stg_off + (uint32_t)(4 * NUM_EPI_WG * p.stage_bufs * STG_BYTES);
  const uint32_t misc_off = hstg_off + (DG ? 4 * NUM_EPI_WG * STG_BYTES : 0);
  float* bias_sm = reinterpret_cast<float*>(base_ptr + misc_off);
  Bars& bar = *reinterpret_cast<Bars*>(base_ptr + misc_off + 256 * sizeof(float));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = (int)blockIdx.x % p.n_tiles, m0 = (int)blockIdx.x / p.n_tiles;
  const int my_tiles = m0 < p.m_tiles ? (p.m_tiles - m0 + p.m_stride - 1) / p.m_stride : 0;
  const int n0 = n_tile * p.BN;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&maps.a);
    prefetch_tensormap(&maps.w);
    prefetch_tensormap(&maps.y);
    if (SAVE || DG) prefetch_tensormap(&maps.h);
    mbar_init(&bar.w_full, 1);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&bar.a_full[i], 1);
      mbar_init(&bar.a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar.acc_full[i], 1);
      mbar_init(&bar.acc_empty[i], 8);  // one arrival per warp of the two warpgroups that drain the buffer
    }
    for (int i = 0; i < 4 * NUM_EPI_WG; ++i) mbar_init(&bar.h_full[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bar.tmem_base, 512);
  if (warp == 3)
    for (int i = lane; i < p.BN; i += 32) bias_sm[i] = p.bias != nullptr ? __ldg(p.bias + n0 + i) : 0.f;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = bar.tmem_base;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      mbar_expect_tx(&bar.w_full, p.w_bytes);
      if (!p.w_mn) {
        for (int kc = 0; kc < p.kchunks; ++kc)  // K-major: per k-chunk BN rows of 64 channels
          tma_load_2d(w_sm + kc * (p.BN * 128), &maps.w, &bar.w_full, kc * BK, n0);
      } else {
        for (int j = 0; j < p.BN / 64; ++j)     // MN-major: per 64 output columns K rows of 128 B
          tma_load_2d(w_sm + j * (p.K * 128), &maps.w, &bar.w_full, n0 + j * 64, 0);
      }
      int it = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int row0 = (m0 + i * p.m_stride) * BM;
        for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
          const int s = it % p.stages;
          mbar_wait(&bar.a_empty[s], ((it / p.stages) & 1) ^ 1);
          mbar_expect_tx(&bar.a_full[s], A_STAGE_BYTES);
          tma_load_2d(a_sm + s * A_STAGE_BYTES, &maps.a, &bar.a_full[s], kc * BK, row0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ========================================
    mbar_wait(&bar.w_full, 0);
    // K-major operands: (addr >> 4) | LBO = 1 (unused); an MN-major weight tile: LBO = bytes between its
    // 64-column blocks, and a K = 16 step is 16 rows of 128 B
    const uint32_t a_lo0 = desc_lo_sw64(a_sm);
    const uint32_t w_lo0 = p.w_mn ? desc_lo_sw128_mn(w_sm, (uint32_t)p.K * 128u) : desc_lo_sw64(w_sm);
    const uint32_t w_kc = p.w_mn ? (uint32_t)(BK * 128) >> 4 : (uint32_t)(p.BN * 128) >> 4;  // per k-chunk
    const uint32_t w_k = p.w_mn ? (16u * 128u) >> 4 : 2u;                                     // per K = 16 step
    int it = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int buf = i & 1;
      mbar_wait(&bar.acc_empty[buf], ((i >> 1) & 1) ^ 1);
      for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
        const int s = it % p.stages;
        mbar_wait(&bar.a_full[s], (it / p.stages) & 1);
        fence_after_sync();
        if (elect_one_sync()) {
          const uint32_t a_lo = a_lo0 + s * (A_STAGE_BYTES >> 4), w_lo = w_lo0 + kc * w_kc;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)  // 16 channels per MMA: 32 B inside the 128-byte swizzled row
            umma_ss2(tmem + buf * 256, a_lo + k * 2, DESC_HI_SW128, w_lo + k * w_k, DESC_HI_SW128, p.idesc,
                     (kc | k) != 0);
          umma_commit(&bar.a_empty[s]);
          if (kc == p.kchunks - 1) umma_commit(&bar.acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ================================= epilogue warpgroups ====================================
    const int wg = (warp - 4) >> 2;           // 0..3
    const int buf = wg & 1, half = wg >> 1;   // TMEM buffer, column half of the n-tile
    const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) << 5) << 16) + buf * 256;
    const int chunks = p.BN / 32;             // 32-column chunks of the n-tile
    const int c_lo = half == 0 ? 0 : (chunks + 1) / 2, c_hi = half == 0 ? (chunks + 1) / 2 : chunks;
    // this warp's staging buffers; a row of a staged box is 64 B, chunk q of row `lane` sits at the 64-byte
    // swizzled position the store tensor map expects
    const uint32_t stg0 = base + stg_off + (uint32_t)(warp - 4) * p.stage_bufs * STG_BYTES;
    const uint32_t my_row = stg0 + lane * 64;
    const int sw = (lane >> 1) & 3;
    int sbuf = 0;                             // staging buffer of the next store
    float colacc[4] = {0.f, 0.f, 0.f, 0.f};   // EPI_DGELU: column (32 cc + lane) of this warp's rows, all tiles
    // EPI_DGELU: the warp's 32 x 32 box of pre-activations arrives by TMA (64-byte swizzled, rows past M
    // zero-filled) one chunk AHEAD of its use: thread-per-row global loads touch 32 lines per instruction
    // and sat un-prefetched in front of every chunk (0.58x of the two-pass path)
    const uint32_t hstg = base + hstg_off + (uint32_t)(warp - 4) * STG_BYTES;
    uint64_t* hbar = &bar.h_full[warp - 4];
    uint32_t hphase = 0;
    if (DG && lane == 0 && buf < my_tiles && c_lo < c_hi) {
      mbar_expect_tx(hbar, STG_BYTES);
      tma_load_2d(hstg, &maps.h, hbar, n0 + c_lo * 32, (m0 + buf * p.m_stride) * BM + ((warp & 3) << 5));
    }
    for (int i = buf; i < my_tiles; i += 2) {
      const int row0 = (m0 + i * p.m_stride) * BM + ((warp & 3) << 5);  // first token row of this warp
      mbar_wait(&bar.acc_full[buf], (i >> 1) & 1);
      fence_after_sync();
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {  // <= 4 chunks per column half (BN <= 256)
        const int c = c_lo + cc;
        if (c >= c_hi) break;
        uint4 hv[4];
        uint32_t r[32];
        tmem_ld32(lane_base + c * 32, r);
        if (DG) {  // this row's 32 pre-activations out of the staged box, then prefetch the next box
          mbar_wait(hbar, hphase);
          hphase ^= 1;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(hv[q].x), "=r"(hv[q].y), "=r"(hv[q].z), "=r"(hv[q].w)
                         : "r"(hstg + lane * 64 + ((q ^ sw) << 4)));
          __syncwarp();
          if (lane == 0) {
            int nc = c + 1, ni = i;
            if (nc >= c_hi) { nc = c_lo; ni = i + 2; }
            if (ni < my_tiles) {
              mbar_expect_tx(hbar, STG_BYTES);
              tma_load_2d(hstg, &maps.h, hbar, n0 + nc * 32, (m0 + ni * p.m_stride) * BM + ((warp & 3) << 5));
            }
          }
        }
        tmem_wait_ld();
        uint32_t outw[16], hw[16];
        if (DG) {
          const uint32_t* hwv = reinterpret_cast<const uint32_t*>(hv);
          float d[32];
#pragma unroll
          for (int j2 = 0; j2 < 16; ++j2) {
            const f2_t g2 = f2_make(__uint_as_float(r[2 * j2]), __uint_as_float(r[2 * j2 + 1]));
            // EPI_DMUL: the staged box holds GELU'(h) itself (forward saved it): one multiplication
            const f2_t v = EPI == EPI_DMUL ? f2_mul(g2, f2_from_bf16x2(hwv[j2])) : gelu_bwd2_f32(g2, hwv[j2]);
            f2_split(v, d[2 * j2], d[2 * j2 + 1]);
            outw[j2] = pack_bf16x2(d[2 * j2], d[2 * j2 + 1]);
          }
          // column sums over the 32 rows of this warp, from the fp32 values BEFORE the bf16 rounding (one
          // rounding error less per term than a sum over the stored tensor): transpose-reduce, 31 shuffles;
          // lane l ends up with column l of the chunk
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int e = 0; e < off; ++e) {
              const float keep = upper ? d[e + off] : d[e], send = upper ? d[e] : d[e + off];
              d[e] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          colacc[cc] += d[0];
        } else {
          const float4* b4 = reinterpret_cast<const float4*>(bias_sm + c * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bb = b4[q];  // broadcast
            const float v0 = __uint_as_float(r[4 * q + 0]) + bb.x, v1 = __uint_as_float(r[4 * q + 1]) + bb.y;
            const float v2 = __uint_as_float(r[4 * q + 2]) + bb.z, v3 = __uint_as_float(r[4 * q + 3]) + bb.w;
            const uint32_t w0 = pack_bf16x2(v0, v1), w1 = pack_bf16x2(v2, v3);
            if (EPI == EPI_BIAS) {
              outw[2 * q] = w0;
              outw[2 * q + 1] = w1;
            } else {
              // GELU of the bf16-ROUNDED pre-activation: what nn.GELU sees after a bf16 Linear under autocast,
              // and exactly what the flat csb200_gelu_fwd pass computes from the stored h
              if (EPI == EPI_GELU_SAVE_D) {  // second output: GELU'(h) instead of h
                outw[2 * q] = gelu_fwd_deriv2(w0, hw[2 * q]);
                outw[2 * q + 1] = gelu_fwd_deriv2(w1, hw[2 * q + 1]);
              } else {
                hw[2 * q] = w0;
                hw[2 * q + 1] = w1;
                outw[2 * q] = gelu_fwd2(w0);
                outw[2 * q + 1] = gelu_fwd2(w1);
              }
            }
          }
        }
        // registers -> staging (conflict-free: 8 rows cover the 32 banks) -> one TMA store per 32 x 32 box
#pragma unroll
        for (int v = 0; v < (SAVE ? 2 : 1); ++v) {
          const uint32_t (&src)[16] = (v == 0 || !SAVE) ? outw : hw;
          if (lane == 0) {
            if (p.stage_bufs == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
          }
          __syncwarp();
          const uint32_t dst = my_row + sbuf * STG_BYTES;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            sts128(dst + ((q ^ sw) << 4), src[4 * q], src[4 * q + 1], src[4 * q + 2], src[4 * q + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(v == 0 ? &maps.y : &maps.h, stg0 + sbuf * STG_BYTES, n0 + c * 32, row0);
            bulk_commit();
          }
          sbuf = p.stage_bufs == 2 ? sbuf ^ 1 : 0;
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar.acc_empty[buf]);
    }
    if (lane == 0) bulk_wait_all();  // the staged boxes have been read AND written before the CTA retires
    if (DG) {
      // per-CTA column sums: 16 warps x 4 chunks x 32 lanes through the (now idle) staging area, added in a
      // fixed order; the per-CTA rows are summed by linear_colsum_final (deterministic)
      __syncwarp();
      asm volatile("bar.sync 1, %0;" ::"n"(128 * NUM_EPI_WG) : "memory");
      float* cs = reinterpret_cast<float*>(base_ptr + stg_off);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) cs[((warp - 4) * 4 + cc) * 32 + lane] = colacc[cc];
      asm volatile("bar.sync 1, %0;" ::"n"(128 * NUM_EPI_WG) : "memory");
      for (int col = (int)threadIdx.x - 128; col < p.BN; col += 128 * NUM_EPI_WG) {
        const int c = col >> 5, hf = c >= (chunks + 1) / 2 ? 1 : 0, cc = c - (hf ? (chunks + 1) / 2 : 0);
        float a = 0.f;
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int rw = 0; rw < 4; ++rw) a += cs[(((b + 2 * hf) * 4 + rw) * 4 + cc) * 32 + (col & 31)];
        p.partial[(int64_t)m0 * p.N + n0 + col] = a;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// out[c] = sum over the per-CTA rows of partial[blocks][cols] (one warp per column, fixed order)
__global__ void __launch_bounds__(256)
    linear_colsum_final(const float* __restrict__ partial, int blocks, int cols, float* __restrict__ out) {
  const int i = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= cols) return;
  const float a = strided_partial_sum(partial + i, blocks, cols, lane);
  if (lane == 0) out[i] = a;
}

int make_map_2d(CUtensorMap* m, const void* base, int64_t inner, int64_t rows, int64_t ld, int box_rows,
                int box_inner = BK, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return fail(CSB200_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
  ensure_context();
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CSB200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return CSB200_OK;
}

// n-tile width: the largest multiple of `step` that divides N, is <= 256 and whose resident weight tile
// leaves room for at least 3 ring stages and the store staging
int pick_bn(int64_t N, int64_t K, int step = 32) {
  // step 64 == the input-gradient form (EPI_DGELU): it also stages the pre-activation boxes
  const int staging = (step == 64 ? 2 : 1) * 4 * NUM_EPI_WG * STG_BYTES;
  for (int bn = 256; bn >= step; bn -= step) {
    if (N % bn != 0) continue;
    const int64_t w_bytes = (int64_t)bn * K * 2;
    if (w_bytes + 4 * A_STAGE_BYTES + staging + 4096 <= SMEM_LIMIT) return bn;
  }
  return 0;
}

bool shape_ok(int64_t M, int64_t N, int64_t K, int dtype, int step) {
  if (dtype != CSB200_BF16) return false;
  if (M < 1 || M > 0x7fffffff / 2 || N < step || N > 65536) return false;
  if (K != 64 && K != 128 && K != 256) return false;
  return pick_bn(N, K, step) != 0;
}

// x [M][K] (row stride ldx) times the weight (w_mn ? [K][N] : [N][K]) -> y [M][N], epilogue EPI
int launch_linear(int epilogue, bool w_mn, const void* x, const void* weight, const float* bias, void* y,
                  void* pre_act, float* partial, int* partial_rows, int64_t M, int64_t N, int64_t K, int64_t ldx,
                  cudaStream_t st) {
  LinMaps maps;
  LinParams p;
  memset(&maps, 0, sizeof(maps));
  memset(&p, 0, sizeof(p));
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.BN = pick_bn(N, K, w_mn ? 64 : 32);
  p.n_tiles = (int)(N / p.BN);
  p.m_tiles = (int)((M + BM - 1) / BM);
  p.kchunks = (int)(K / BK);
  p.w_bytes = (uint32_t)(p.BN * K * 2);
  p.w_mn = w_mn ? 1 : 0;
  // two staging buffers per epilogue warp when that still leaves >= 4 ring stages
  const int misc = 256 * (int)sizeof(float) + (int)sizeof(Bars) + 1024 +
                   (epilogue == EPI_DGELU || epilogue == EPI_DMUL ? 4 * NUM_EPI_WG * STG_BYTES : 0);
  p.stage_bufs = (SMEM_LIMIT - (int)p.w_bytes - misc - 2 * 4 * NUM_EPI_WG * STG_BYTES) / A_STAGE_BYTES >= 5 ? 2 : 1;
  const int fixed = (int)p.w_bytes + misc + p.stage_bufs * 4 * NUM_EPI_WG * STG_BYTES;
  p.stages = (SMEM_LIMIT - fixed) / A_STAGE_BYTES;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  p.idesc = umma_idesc_bf16(p.BN, false, w_mn);
  p.bias = bias;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.h = static_cast<__nv_bfloat16*>(pre_act);
  p.partial = partial;
  int rc;
  if ((rc = make_map_2d(&maps.a, x, K, M, ldx, BM)) != CSB200_OK) return rc;
  if (!w_mn) rc = make_map_2d(&maps.w, weight, K, N, K, p.BN);
  else rc = make_map_2d(&maps.w, weight, N, K, N, (int)K);  // box: 64 output columns x all K rows
  if (rc != CSB200_OK) return rc;
  if ((rc = make_map_2d(&maps.y, y, N, M, N, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) != CSB200_OK) return rc;
  if ((epilogue == EPI_GELU_SAVE || epilogue == EPI_DGELU || epilogue == EPI_GELU_SAVE_D || epilogue == EPI_DMUL) &&
      (rc = make_map_2d(&maps.h, pre_act, N, M, N, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) != CSB200_OK)
    return rc;
  const int sms = device_sm_count();
  if (sms <= 0) return fail(CSB200_ERR_CUDA, "csb200_linear: cannot query the SM count");
  int per_n = sms / p.n_tiles;              // CTAs per n-tile
  if (per_n < 1) per_n = 1;
  if (per_n > p.m_tiles) per_n = p.m_tiles;
  p.m_stride = per_n;
  if (partial_rows != nullptr) *partial_rows = per_n;
  const int grid = per_n * p.n_tiles;
  const int smem = fixed + p.stages * A_STAGE_BYTES;
#define CSB_LIN_LAUNCH(E)                                                                        \
  case E:                                                                                         \
    CSB200_CUDA(opt_in_smem(reinterpret_cast<const void*>(&linear_tc_kernel<E>), SMEM_LIMIT));    \
    linear_tc_kernel<E><<<grid, THREADS, smem, st>>>(maps, p);                                    \
    break;
  switch (epilogue) {
    CSB_LIN_LAUNCH(EPI_BIAS)
    CSB_LIN_LAUNCH(EPI_GELU)
    CSB_LIN_LAUNCH(EPI_GELU_SAVE)
    CSB_LIN_LAUNCH(EPI_DGELU)
    CSB_LIN_LAUNCH(EPI_GELU_SAVE_D)
    CSB_LIN_LAUNCH(EPI_DMUL)
    default:
      return fail(CSB200_ERR_INVALID, "csb200_linear: unknown epilogue %d", epilogue);
  }
#undef CSB_LIN_LAUNCH
  return check_launch("linear_tc_kernel");
}

constexpr int MAX_PARTIAL_ROWS = 160;

}  // namespace
}  // namespace csb200

using namespace csb200;

extern "C" {

CSB200_API int csb200_linear_supported(int64_t M, int64_t N, int64_t K, int dtype) {
  return shape_ok(M, N, K, dtype, 32) ? 1 : 0;
}

CSB200_API int csb200_linear_fwd(const void* x, const void* weight, const float* bias, void* y, void* pre_act, int64_t M,
                      int64_t N, int64_t K, int64_t ldx, int dtype, int epilogue, void* stream) {
  if (M == 0) return CSB200_OK;
  if (x == nullptr || weight == nullptr || y == nullptr)
    return fail(CSB200_ERR_INVALID, "csb200_linear_fwd: null pointer");
  // public codes: 0 bias, 1 GELU, 2 GELU + pre-activation h, 3 GELU + GELU'(h) (CSB200_EPI_GELU_SAVE_DERIV)
  if (epilogue < 0 || epilogue > 3 || (epilogue >= 2 && pre_act == nullptr))
    return fail(CSB200_ERR_INVALID, "csb200_linear_fwd: bad epilogue %d", epilogue);
  if (epilogue == 3) epilogue = EPI_GELU_SAVE_D;
  if (!shape_ok(M, N, K, dtype, 32))
    return fail(CSB200_ERR_UNSUPPORTED, "csb200_linear_fwd: bf16 with K in {64,128,256} and N a multiple of 32 only "
                "(M %lld, N %lld, K %lld)", (long long)M, (long long)N, (long long)K);
  if (ldx < K || (ldx * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(weight) & 15) ||
      (reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(pre_act) & 15))
    return fail(CSB200_ERR_INVALID, "csb200_linear_fwd: operands must be 16-byte aligned (ldx %lld)", (long long)ldx);
  return launch_linear(epilogue, false, x, weight, bias, y, pre_act, nullptr, nullptr, M, N, K, ldx,
                       static_cast<cudaStream_t>(stream));
}

CSB200_API int csb200_linear_dgelu_supported(int64_t M, int64_t N, int64_t K, int dtype) {
  return shape_ok(M, N, K, dtype, 64) ? 1 : 0;
}

CSB200_API size_t csb200_linear_dgelu_workspace_bytes(int64_t N) {
  return (size_t)MAX_PARTIAL_ROWS * (size_t)N * sizeof(float) + 256;
}

static int linear_dact_impl(int epi, const void* grad_y, const void* weight, const void* pre_act, void* grad_h,
                            float* grad_bias, void* workspace, size_t workspace_bytes, int64_t M, int64_t N,
                            int64_t K, int64_t ldg, int dtype, void* stream, const float** partials = nullptr,
                            int32_t* partial_rows = nullptr) {
  const bool deferred = partial_rows != nullptr;
  if (M == 0) return deferred ? fail(CSB200_ERR_INVALID, "csb200_linear_dact_bwd_partials: M == 0") : CSB200_OK;
  if (grad_y == nullptr || weight == nullptr || pre_act == nullptr || grad_h == nullptr ||
      (!deferred && grad_bias == nullptr) || workspace == nullptr)
    return fail(CSB200_ERR_INVALID, "csb200_linear_dgelu_bwd: null pointer");
  if (!shape_ok(M, N, K, dtype, 64))
    return fail(CSB200_ERR_UNSUPPORTED, "csb200_linear_dgelu_bwd: bf16 with K in {64,128,256} and N a multiple of 64 "
                "only (M %lld, N %lld, K %lld)", (long long)M, (long long)N, (long long)K);
  if (workspace_bytes < csb200_linear_dgelu_workspace_bytes(N))
    return fail(CSB200_ERR_WORKSPACE, "csb200_linear_dgelu_bwd: workspace too small");
  if (ldg < K || (ldg * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(grad_y) & 15) ||
      (reinterpret_cast<uintptr_t>(weight) & 15) || (reinterpret_cast<uintptr_t>(pre_act) & 15) ||
      (reinterpret_cast<uintptr_t>(grad_h) & 15))
    return fail(CSB200_ERR_INVALID, "csb200_linear_dgelu_bwd: operands must be 16-byte aligned (ldg %lld)", (long long)ldg);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  int rows = 0;
  int rc = launch_linear(epi, true, grad_y, weight, nullptr, grad_h, const_cast<void*>(pre_act), partial, &rows,
                         M, N, K, ldg, st);
  if (rc != CSB200_OK) return rc;
  if (rows > MAX_PARTIAL_ROWS) return fail(CSB200_ERR_WORKSPACE, "csb200_linear_dgelu_bwd: %d partial rows", rows);
  if (deferred) {  // the caller records the final sum (csb200_sum_rows_deferred / _flush, sum_rows.cu)
    *partials = partial;
    *partial_rows = rows;
    return CSB200_OK;
  }
  linear_colsum_final<<<(int)((N * 32 + 255) / 256), 256, 0, st>>>(partial, rows, (int)N, grad_bias);
  return check_launch("linear_colsum_final");
}

CSB200_API int csb200_linear_dgelu_bwd(const void* grad_y, const void* weight, const void* pre_act, void* grad_h,
                                       float* grad_bias, void* workspace, size_t workspace_bytes, int64_t M,
                                       int64_t N, int64_t K, int64_t ldg, int dtype, void* stream) {
  return linear_dact_impl(EPI_DGELU, grad_y, weight, pre_act, grad_h, grad_bias, workspace, workspace_bytes, M, N, K,
                          ldg, dtype, stream);
}

CSB200_API int csb200_linear_dact_bwd(const void* grad_y, const void* weight, const void* act_deriv, void* grad_h,
                                      float* grad_bias, void* workspace, size_t workspace_bytes, int64_t M,
                                      int64_t N, int64_t K, int64_t ldg, int dtype, void* stream) {
  return linear_dact_impl(EPI_DMUL, grad_y, weight, act_deriv, grad_h, grad_bias, workspace, workspace_bytes, M, N, K,
                          ldg, dtype, stream);
}

// Both input-gradient GEMMs without their last launch: the per-CTA column sums of grad_h stay in the workspace
// as float[*partial_rows][N] at *partials.
CSB200_API int csb200_linear_dact_bwd_partials(const void* grad_y, const void* weight, const void* act, void* grad_h,
                                               void* workspace, size_t workspace_bytes, int64_t M, int64_t N,
                                               int64_t K, int64_t ldg, int dtype, int use_saved_derivative,
                                               const float** partials, int32_t* partial_rows, void* stream) {
  if (partials == nullptr || partial_rows == nullptr)
    return fail(CSB200_ERR_INVALID, "csb200_linear_dact_bwd_partials: null pointer");
  return linear_dact_impl(use_saved_derivative ? EPI_DMUL : EPI_DGELU, grad_y, weight, act, grad_h, nullptr, workspace,
                          workspace_bytes, M, N, K, ldg, dtype, stream, partials, partial_rows);
}

}  // extern "C"

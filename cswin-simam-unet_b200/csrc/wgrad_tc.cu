// Weight (and bias) gradient of a token-path nn.Linear on the tcgen05 tensor cores:
//
//     grad_W[N][K] = grad_y[M][N]^T  x[M][K]          grad_b[N] = sum_m grad_y[m][n]
//
// (the backward of `qkv` C:357-358, `proj` C:366, `Mlp.fc1` / `fc2` C:188-196 with respect to their
// parameters).  M is the token count of the batch (32 768 ... 524 288), the output is tiny: the op streams
// both operands once and is HBM-bound.  cuBLAS runs it as a split-K GEMM plus a reduce kernel at 40-80 % of
// the HBM roofline and leaves the bias gradient to a separate column-sum pass.  Here:
//
//   * the contraction runs over TOKENS, so both operands are MN-major for the tensor core: 128-token x
//     64-column TMA boxes (128-byte swizzle) of grad_y and x are consumed in place — no transposes;
//   * a CTA owns one 128 x (<= 256) tile of grad_W and one contiguous range of tokens (the grid is
//     tiles x splits ~ one CTA per SM); warps 0, 2, 3 = TMA producers (the 3-6 boxes of a pipeline step are dealt
//     round-robin to them), warp 1 = tcgen05.mma issuer (M128 x N{<=256} x K16, accumulators in TMEM), warp 2
//     also allocates TMEM, warp 3 also writes the ones block, warps 4-7 = epilogue.  Measured with 1 / 2 / 3
//     producers (profiles/r2i_wgrad_producers.txt): no change at the stage-1..3 shapes, 3-6 % at stage 4 — the
//     kernel is NOT bound by the box-issue rate of one thread; at the short token counts of stages 3-4 the
//     time is launch + ramp, ~13 us of streaming and then ~9 us in which every CTA dumps its 128 x 256 fp32
//     partial tile through red.global.add at once (18-24 splits per output element, 19 MB of reductions);
//   * the bias gradient is one more MMA per step against a constant block of ones (N = 16): grad_y^T 1;
//   * partial tiles are added into the zeroed fp32 gradient with 16-byte vector reductions
//     (red.global.add.v4.f32): no partial buffers, no reduce kernel.  The order of those additions is not
//     fixed, so the last bits of grad_W can differ from run to run (as with any atomic reduction).

#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace csb200 {
namespace {
using namespace tc;

constexpr int TOK = 128;                      // tokens per pipeline step (the MMA contraction chunk)
constexpr int BOX_BYTES = TOK * 128;          // one 64-column x 128-token box, 16 KB
constexpr int BN = 128;                       // rows of grad_W per tile == TMEM lanes
constexpr int MAX_STAGES = 6;
constexpr int THREADS = 256;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr uint32_t BIAS_COL = 256;            // TMEM column of the bias accumulator

struct WgParams {
  int M, N, K;
  int n_tiles, k_tiles, splits, chunks, per_split;
  int bk;                  // columns of this launch's tiles (multiple of 64, <= 256)
  int stages, stage_bytes;
  int nprod;               // TMA-issuing warps (1..3)
  uint32_t idesc, idesc_bias;
  float* gw;               // [N][K]
  float* gb;               // [N] or nullptr
};
struct WgMaps {
  CUtensorMap g, x;
};
struct WgBars {
  uint64_t full[MAX_STAGES], empty[MAX_STAGES];
  uint64_t acc_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// shared memory: [ones block 16 KB][stages x (2 grad_y boxes + bk/64 x boxes)][barriers]
__global__ void __launch_bounds__(THREADS, 1)
    wgrad_tc_kernel(const __grid_constant__ WgMaps maps, const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw) + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t ones_sm = base, ring = base + BOX_BYTES;
  WgBars& bar = *reinterpret_cast<WgBars*>(base_ptr + BOX_BYTES + p.stages * p.stage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = p.n_tiles * p.k_tiles;
  const int tile = (int)blockIdx.x % tiles, split = (int)blockIdx.x / tiles;
  const int nt = tile % p.n_tiles, kt = tile / p.n_tiles;
  const int n0 = nt * BN, k0 = kt * p.bk;
  const int c_begin = split * p.per_split;
  const int c_end = c_begin + p.per_split < p.chunks ? c_begin + p.per_split : p.chunks;
  const int my_chunks = c_end > c_begin ? c_end - c_begin : 0;
  const bool with_bias = p.gb != nullptr && kt == 0;
  const int xboxes = p.bk / 64;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&maps.g);
    prefetch_tensormap(&maps.x);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&bar.full[i], (uint32_t)p.nprod);  // every producer arrives with the bytes of its own boxes
      mbar_init(&bar.empty[i], 1);
    }
    mbar_init(&bar.acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bar.tmem_base, 512);
  if (warp == 3) {  // the constant B operand of the bias MMA: 128 token rows x 64 columns of bf16 1.0
    uint4* o = reinterpret_cast<uint4*>(base_ptr);
    for (int i = lane; i < BOX_BYTES / 16; i += 32) o[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    fence_proxy_async_smem();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = bar.tmem_base;

  const int prod = warp == 0 ? 0 : warp == 2 ? 1 : warp == 3 ? 2 : -1;
  if (prod >= 0) {
    // ===================================== TMA producers ====================================
    // box b of a step: b < 2 -> grad_y columns n0 + 64 b (columns past N: zero-filled), else x columns
    // k0 + 64 (b - 2); producer j issues the boxes b == j (mod nprod)
    const int nboxes = 2 + xboxes;  // 3..6 >= nprod
    if (lane == 0 && prod < p.nprod) {
      const uint32_t my_bytes = (uint32_t)((nboxes - prod + p.nprod - 1) / p.nprod) * BOX_BYTES;
      for (int i = 0; i < my_chunks; ++i) {
        const int s = i % p.stages, tok0 = (c_begin + i) * TOK;
        mbar_wait(&bar.empty[s], ((i / p.stages) & 1) ^ 1);
        mbar_expect_tx(&bar.full[s], my_bytes);
        const uint32_t st = ring + s * p.stage_bytes;
        for (int b = prod; b < nboxes; b += p.nprod) {
          if (b < 2)
            tma_load_2d(st + b * BOX_BYTES, &maps.g, &bar.full[s], n0 + 64 * b, tok0);
          else
            tma_load_2d(st + b * BOX_BYTES, &maps.x, &bar.full[s], k0 + (b - 2) * 64, tok0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ========================================
    const uint32_t ones_lo = desc_lo_sw128_mn(ones_sm, BOX_BYTES);
    for (int i = 0; i < my_chunks; ++i) {
      const int s = i % p.stages;
      mbar_wait(&bar.full[s], (i / p.stages) & 1);
      fence_after_sync();
      if (elect_one_sync()) {
        const uint32_t st = ring + s * p.stage_bytes;
        const uint32_t g_lo = desc_lo_sw128_mn(st, BOX_BYTES), x_lo = desc_lo_sw128_mn(st + 2 * BOX_BYTES, BOX_BYTES);
#pragma unroll
        for (int k = 0; k < TOK / 16; ++k) {  // 16 tokens per MMA: 16 rows of 128 B in every box
          umma_ss2(tmem, g_lo + k * (2048 >> 4), DESC_HI_SW128, x_lo + k * (2048 >> 4), DESC_HI_SW128, p.idesc,
                   (i | k) != 0);
          if (with_bias)
            umma_ss2(tmem + BIAS_COL, g_lo + k * (2048 >> 4), DESC_HI_SW128, ones_lo, DESC_HI_SW128, p.idesc_bias,
                     (i | k) != 0);
        }
        umma_commit(&bar.empty[s]);
        if (i == my_chunks - 1) umma_commit(&bar.acc_full);
      }
      __syncwarp();
    }
  } else if (warp >= 4 && my_chunks > 0) {
    // ===================================== epilogue ===========================================
    const int row = ((warp & 3) << 5) | lane;  // row of the tile == TMEM lane
    const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) << 5) << 16);
    mbar_wait(&bar.acc_full, 0);
    fence_after_sync();
    const bool live = n0 + row < p.N;
    float* dst = p.gw + (int64_t)(n0 + row) * p.K + k0;
    for (int c = 0; c < p.bk / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(lane_base + c * 32, r);
      tmem_wait_ld();
      if (live) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          red_add_v4(dst + c * 32 + q * 4, __uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                     __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
      }
    }
    if (with_bias) {
      uint32_t r[32];
      tmem_ld32(lane_base + BIAS_COL, r);  // 16 identical columns (+ 16 unused)
      tmem_wait_ld();
      if (live) atomicAdd(p.gb + n0 + row, __uint_as_float(r[0]));
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

int make_map(CUtensorMap* m, const void* base, int64_t inner, int64_t rows, int64_t ld) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return fail(CSB200_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
  ensure_context();
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64, TOK};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CSB200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return CSB200_OK;
}

// CSB200_WGRAD_PRODUCERS=1|2|3 (experiments; default 3)
int wgrad_producers() {
  static const int n = [] {
    const char* e = getenv("CSB200_WGRAD_PRODUCERS");
    const int v = e != nullptr ? atoi(e) : 3;
    return v < 1 ? 1 : v > 3 ? 3 : v;
  }();
  return n;
}

bool shape_ok(int64_t M, int64_t N, int64_t K, int dtype) {
  if (dtype != CSB200_BF16) return false;
  if (M < 1 || M > 0x7fffffff / 2 || N < 8 || N > 65536 || K < 64 || K > 65536) return false;
  return N % 8 == 0 && K % 64 == 0;
}

}  // namespace
}  // namespace csb200

using namespace csb200;

extern "C" {

CSB200_API int csb200_linear_wgrad_supported(int64_t M, int64_t N, int64_t K, int dtype) {
  return shape_ok(M, N, K, dtype) ? 1 : 0;
}

static int wgrad_impl(const void* grad_y, const void* x, float* grad_w, float* grad_bias, int64_t M, int64_t N,
                      int64_t K, int64_t ldg, int64_t ldx, int dtype, void* stream, bool zero_outputs) {
  if (grad_y == nullptr || x == nullptr || grad_w == nullptr)
    return fail(CSB200_ERR_INVALID, "csb200_linear_wgrad: null pointer");
  if (!shape_ok(M, N, K, dtype))
    return fail(CSB200_ERR_UNSUPPORTED, "csb200_linear_wgrad: bf16 with N a multiple of 8 and K a multiple of 64 only "
                "(M %lld, N %lld, K %lld)", (long long)M, (long long)N, (long long)K);
  if (ldg < N || ldx < K || (ldg * 2) % 16 != 0 || (ldx * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(grad_y) & 15) ||
      (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(grad_w) & 15))
    return fail(CSB200_ERR_INVALID, "csb200_linear_wgrad: operands must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (zero_outputs) {
    CSB200_CUDA(cudaMemsetAsync(grad_w, 0, (size_t)N * K * sizeof(float), st));
    if (grad_bias != nullptr) CSB200_CUDA(cudaMemsetAsync(grad_bias, 0, (size_t)N * sizeof(float), st));
  }
  WgMaps maps;
  WgParams p;
  memset(&maps, 0, sizeof(maps));
  memset(&p, 0, sizeof(p));
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.bk = 64;
  for (int bk : {256, 192, 128})  // the widest tile (multiple of 64, <= 256 TMEM columns) that divides K
    if (K % bk == 0) {
      p.bk = bk;
      break;
    }
  p.n_tiles = (int)((N + BN - 1) / BN);
  p.k_tiles = (int)(K / p.bk);
  p.chunks = (int)((M + TOK - 1) / TOK);
  const int sms = device_sm_count();
  if (sms <= 0) return fail(CSB200_ERR_CUDA, "csb200_linear_wgrad: cannot query the SM count");
  const int tiles = p.n_tiles * p.k_tiles;
  p.splits = sms / tiles < 1 ? 1 : sms / tiles;
  if (p.splits > p.chunks) p.splits = p.chunks;
  p.per_split = (p.chunks + p.splits - 1) / p.splits;
  p.splits = (p.chunks + p.per_split - 1) / p.per_split;  // no empty splits
  p.stage_bytes = (2 + p.bk / 64) * BOX_BYTES;
  const int fixed = BOX_BYTES + (int)sizeof(WgBars) + 1024;
  p.stages = (SMEM_LIMIT - fixed) / p.stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  p.nprod = wgrad_producers();
  p.idesc = umma_idesc_bf16(p.bk, true, true);
  p.idesc_bias = umma_idesc_bf16(16, true, true);
  p.gw = grad_w;
  p.gb = grad_bias;
  int rc;
  if ((rc = make_map(&maps.g, grad_y, N, M, ldg)) != CSB200_OK) return rc;
  if ((rc = make_map(&maps.x, x, K, M, ldx)) != CSB200_OK) return rc;
  // more than half of the shared memory: one CTA (which owns all 512 TMEM columns) per SM
  int smem = fixed + p.stages * p.stage_bytes;
  if (smem < 120 * 1024) smem = 120 * 1024;
  CSB200_CUDA(opt_in_smem(reinterpret_cast<const void*>(&wgrad_tc_kernel), SMEM_LIMIT));
  wgrad_tc_kernel<<<tiles * p.splits, THREADS, smem, st>>>(maps, p);
  return check_launch("wgrad_tc_kernel");
}

CSB200_API int csb200_linear_wgrad(const void* grad_y, const void* x, float* grad_w, float* grad_bias, int64_t M,
                                   int64_t N, int64_t K, int64_t ldg, int64_t ldx, int dtype, void* stream) {
  return wgrad_impl(grad_y, x, grad_w, grad_bias, M, N, K, ldg, ldx, dtype, stream, true);
}

// The same pass ADDING into grad_w / grad_bias (no memsets): the caller hands over zero-filled outputs — one
// arena zeroed by ONE memset per backward pass instead of two memset nodes in front of every call — or partial
// gradients to accumulate into.
CSB200_API int csb200_linear_wgrad_acc(const void* grad_y, const void* x, float* grad_w, float* grad_bias, int64_t M,
                                       int64_t N, int64_t K, int64_t ldg, int64_t ldx, int dtype, void* stream) {
  return wgrad_impl(grad_y, x, grad_w, grad_bias, M, N, K, ldg, ldx, dtype, stream, false);
}

}  // extern "C"

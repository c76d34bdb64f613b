// Microbenchmark: TMA load throughput as a function of the box's inner extent (bytes per row) on a token-major
// buffer of 384-byte rows (the packed qkv buffer of stage 1), one persistent CTA per SM, nothing but the copies.
// Decides whether a layout that puts a head's q | k (or q | k | v) side by side is worth it (the
// 64-byte rows of one head cap stripe_fwd_tc<128> near 3 TB/s).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../cswin-simam-unet_b200/csrc \
//        -I../../include -o tma_rows tma_rows.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace csb200::tc;

constexpr int MAX_STAGES = 32;
constexpr int BUF_BYTES = 200 * 1024;

struct Sm {
  alignas(1024) uint8_t buf[BUF_BYTES];
  uint64_t full[MAX_STAGES];
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// boxes: each CTA loads `per_cta` boxes; box b of the grid is (chan0 = (b % nsplit) * inner_elems, row0 = (b / nsplit) * rows)
__global__ void __launch_bounds__(256, 1)
    k(const __grid_constant__ CUtensorMap map, int boxes, int nsplit, int inner_elems, int rows, int box_bytes, int col_mode,
      int W, int STAGES, int NPROD) {
  extern __shared__ uint8_t raw[];
  Sm& sm = *reinterpret_cast<Sm*>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_STAGES; ++i) mbar_init(&sm.full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) < NPROD) {
    const int w = threadIdx.x >> 5;
    STAGES /= NPROD;  // each producer warp owns its own slice of the ring
    uint8_t* mybuf = sm.buf + w * STAGES * box_bytes;
    uint64_t* full = sm.full + w * STAGES;
    int n = 0;
    for (int b = blockIdx.x + w * gridDim.x; b < boxes; b += gridDim.x * NPROD, ++n) {
      const int s = n % STAGES;
      if (n >= STAGES) mbar_wait(&full[s], ((n / STAGES) - 1) & 1);  // the previous load into this stage has landed
      mbar_expect_tx(&full[s], box_bytes);
      const int piece = b % nsplit, blk = b / nsplit;
      if (col_mode) {
        // column stripes: 128 rows at a stride of W tokens: tensor dims (chan, x, y)
        const int x = blk % W, y0 = (blk / W) * rows;
        tma_load_3d(mybuf + s * box_bytes, &map, &full[s], piece * inner_elems, x, y0);
      } else {
        int x0, y;
        if (rows <= W) {
          const int per_row = W / rows;
          x0 = (blk % per_row) * rows;
          y = blk / per_row;
        } else {
          x0 = 0;
          y = blk * (rows / W);
        }
        tma_load_3d(mybuf + s * box_bytes, &map, &full[s], piece * inner_elems, x0, y);
      }
    }
    // wait for everything outstanding
    for (int j = (n > STAGES ? n - STAGES : 0); j < n; ++j) mbar_wait(&full[j % STAGES], (j / STAGES) & 1);
  }
}

int main() {
  const int W = 128, H = 128, B = 32, ROW = 192;  // 192 bf16 = 384 B per token
  const size_t tokens = (size_t)B * H * W;
  void* d;
  cudaMalloc(&d, tokens * ROW * 2);
  cudaMemset(d, 1, tokens * ROW * 2);
  void* big;
  cudaMalloc(&big, 512u << 20);
  struct Case { const char* name; int inner; int rows; CUtensorMapSwizzle sw; int col; int stages; int nprod; };
  const Case cases[] = {
      {"row  64B x128 sw64  st16 p1", 32, 128, CU_TENSOR_MAP_SWIZZLE_64B, 0, 16, 1},
      {"row  64B x128 sw64  st16 p2", 32, 128, CU_TENSOR_MAP_SWIZZLE_64B, 0, 16, 2},
      {"row  64B x128 sw64  st16 p4", 32, 128, CU_TENSOR_MAP_SWIZZLE_64B, 0, 16, 4},
      {"row  64B x128 sw64  st24 p8", 32, 128, CU_TENSOR_MAP_SWIZZLE_64B, 0, 24, 8},
      {"col  64B x128 sw64  st16 p1", 32, 128, CU_TENSOR_MAP_SWIZZLE_64B, 1, 16, 1},
      {"col  64B x128 sw64  st16 p4", 32, 128, CU_TENSOR_MAP_SWIZZLE_64B, 1, 16, 4},
      {"row 128B x128 sw128 st12 p1", 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 0, 12, 1},
      {"row 128B x128 sw128 st12 p2", 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 0, 12, 2},
      {"row 128B x128 sw128 st12 p4", 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 0, 12, 4},
      {"row 128B x 64 sw128 st24 p4", 64, 64, CU_TENSOR_MAP_SWIZZLE_128B, 0, 24, 4},
      {"row  64B x 64 sw64  st24 p4", 32, 64, CU_TENSOR_MAP_SWIZZLE_64B, 0, 24, 4},
      {"row  64B x256 sw64  st12 p1", 32, 256, CU_TENSOR_MAP_SWIZZLE_64B, 0, 12, 1},
      {"row  64B x256 sw64  st12 p4", 32, 256, CU_TENSOR_MAP_SWIZZLE_64B, 0, 12, 4},
      {"row 192B x128 none  st8  p1", 96, 128, CU_TENSOR_MAP_SWIZZLE_NONE, 0, 8, 1},
      {"row 192B x128 none  st8  p4", 96, 128, CU_TENSOR_MAP_SWIZZLE_NONE, 0, 8, 4},
  };
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sm) + 1024);
  for (const Case& c : cases) {
    CUtensorMap m;
    const cuuint64_t dims[3] = {(cuuint64_t)ROW, (cuuint64_t)W, (cuuint64_t)H * B};
    const cuuint64_t strides[2] = {(cuuint64_t)ROW * 2, (cuuint64_t)ROW * 2 * W};
    const cuuint32_t box[3] = {(cuuint32_t)c.inner, c.col ? 1u : (cuuint32_t)(c.rows < W ? c.rows : W), c.col ? (cuuint32_t)c.rows : (cuuint32_t)(c.rows <= W ? 1 : c.rows / W)};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("%s encode failed %d\n", c.name, (int)r);
      continue;
    }
    const int nsplit = ROW / c.inner;             // pieces per token row: all of the buffer is read once
    const int boxes = (int)(tokens / c.rows) * nsplit;
    const int box_bytes = c.inner * 2 * c.rows;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaMemsetAsync(big, rep, 512u << 20);  // flush L2
      cudaEventRecord(e0);
      k<<<sms, 256, sizeof(Sm) + 1024>>>(m, boxes, nsplit, c.inner, c.rows, box_bytes, c.col, W, c.stages, c.nprod);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    const double bytes = (double)tokens * ROW * 2;
    printf("%s  %8.1f us  %7.1f GB/s  (%s)\n", c.name, best * 1e3, bytes / best / 1e6, cudaGetErrorString(err));
  }
  return 0;
}

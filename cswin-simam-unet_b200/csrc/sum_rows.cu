// Deferred final sums (include/csb200.h, "Deferred final sums"): the last stage of every "per-CTA partial rows ->
// one vector" reduction of the backward pass — LayerNorm gamma / beta / residual-bias gradients (C:357, C:368),
// bias gradients by column sums (C:358, C:366) and the fc1 bias gradient of the fused Mlp backward (C:188-196) —
// recorded on the host and performed by ONE launch per 120 records instead of one 4-5 us launch each.
//
// The records travel as a KERNEL PARAMETER (3.8 KB, __grid_constant__): no device table, no host -> device copy,
// and a CUDA-graph capture keeps them by value, so nothing on the host has to outlive the capture.
// One warp per output element, lanes stride over the partial rows: exactly strided_partial_sum of the immediate
// kernels (layernorm_param_grad_final, colsum_final, linear_colsum_final) -> bit-identical sums.

#include <mutex>
#include <vector>

#include "common.cuh"

namespace csb200 {
namespace {

struct SumJob {
  const float* partial;
  float* out;
  int32_t rows, cols, stride, first_col;
};
constexpr int JOBS_PER_LAUNCH = 120;
struct SumTable {
  int32_t n, total_cols;
  SumJob jobs[JOBS_PER_LAUNCH];
};
static_assert(sizeof(SumJob) == 32 && sizeof(SumTable) <= 4000, "the table must fit the 4-KB kernel parameter space");

__global__ void __launch_bounds__(256) sum_rows_jobs_kernel(const __grid_constant__ SumTable t) {
  const int g = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g >= t.total_cols) return;
  int lo = 0, hi = t.n - 1;  // the last job whose first_col <= g (warp-uniform search)
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (t.jobs[mid].first_col <= g) lo = mid;
    else hi = mid - 1;
  }
  const SumJob& j = t.jobs[lo];
  const int i = g - j.first_col;
  const float a = strided_partial_sum(j.partial + i, j.rows, (int64_t)j.stride, lane);
  if (lane == 0) j.out[i] = a;
}

std::mutex g_mu;
std::vector<SumJob> g_jobs;

}  // namespace
}  // namespace csb200

using namespace csb200;

extern "C" {

CSB200_API int csb200_sum_rows_deferred(const float* partial, int64_t rows, int64_t cols, int64_t row_stride,
                                        float* out) {
  if (partial == nullptr || out == nullptr) return fail(CSB200_ERR_INVALID, "csb200_sum_rows_deferred: null pointer");
  if (rows < 1 || rows > 0x7fffffff || cols < 1 || cols > (1 << 24) || row_stride < cols || row_stride > 0x7fffffff)
    return fail(CSB200_ERR_INVALID, "csb200_sum_rows_deferred: rows %lld, cols %lld, row stride %lld", (long long)rows,
                (long long)cols, (long long)row_stride);
  std::lock_guard<std::mutex> lk(g_mu);
  g_jobs.push_back(SumJob{partial, out, (int32_t)rows, (int32_t)cols, (int32_t)row_stride, 0});
  return CSB200_OK;
}

CSB200_API int64_t csb200_sum_rows_pending(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  return (int64_t)g_jobs.size();
}

CSB200_API int csb200_sum_rows_discard(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_jobs.clear();
  return CSB200_OK;
}

CSB200_API int csb200_sum_rows_flush(void* stream) {
  std::vector<SumJob> jobs;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    jobs.swap(g_jobs);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (size_t j0 = 0; j0 < jobs.size(); j0 += JOBS_PER_LAUNCH) {
    SumTable t;
    t.n = (int32_t)(jobs.size() - j0 < (size_t)JOBS_PER_LAUNCH ? jobs.size() - j0 : (size_t)JOBS_PER_LAUNCH);
    int64_t total = 0;
    for (int i = 0; i < t.n; ++i) {
      t.jobs[i] = jobs[j0 + i];
      t.jobs[i].first_col = (int32_t)total;
      total += t.jobs[i].cols;
    }
    for (int i = t.n; i < JOBS_PER_LAUNCH; ++i) t.jobs[i] = SumJob{nullptr, nullptr, 0, 0, 0, 0x7fffffff};
    if (total > (int64_t)0x7fffffff / 32) return fail(CSB200_ERR_INVALID, "csb200_sum_rows_flush: too many columns");
    t.total_cols = (int32_t)total;
    sum_rows_jobs_kernel<<<(int)((total * 32 + 255) / 256), 256, 0, st>>>(t);
    const int rc = check_launch("sum_rows_jobs_kernel");
    if (rc != CSB200_OK) return rc;
  }
  return CSB200_OK;
}

}  // extern "C"

// tcgen05 / TMEM / TMA engine of the stripe attention (bf16, fp32 accumulate) — placeholder that
// reports "unsupported" until the kernels land; AUTO then resolves to the CUDA-core engine.
#include "stripe_attn.cuh"

namespace csb200 {
bool tc_fwd_supported(const StripeGeom&, int) { return false; }
bool tc_bwd_supported(const StripeGeom&, int) { return false; }
int tc_fwd(const StripeGeom&, const void*, const void*, const void*, const float*, const float*,
           void*, float*, cudaStream_t) {
  return fail(CSB200_ERR_UNSUPPORTED, "tcgen05 engine not built");
}
}  // namespace csb200

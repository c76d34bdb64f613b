// Internal interface between the C-ABI entry points (stripe_attn.cu) and the attention engines.
#pragma once
#include "common.cuh"

namespace csb200 {

// Validated, kernel-friendly copy of csb200_stripe_desc (passed by value to the kernels).
struct StripeGeom {
  int B, H, W, L;      // batch, token grid, L = H*W
  int hs, ws, N;       // stripe extent, N = hs*ws tokens per stripe
  int nwy, nwx;        // stripes per image along y / x
  int heads;           // heads of this branch (head_dim is 32)
  float scale;
  int64_t q_sb, q_sl, k_sb, k_sl, v_sb, v_sl, o_sb, o_sl;
  int64_t dq_sb, dq_sl, dk_sb, dk_sl, dv_sb, dv_sl;
  // attention dropout (csb200_stripe_desc.drop_p): a probability is DROPPED when its random byte < drop_thr
  uint32_t drop_thr;        // 0 = off; p = drop_thr / 256
  float keep_scale;         // 1 / (1 - p)
  int mask_words;           // ceil(N / 32)
  uint32_t drop_salt;
  const unsigned long long* rng;  // device [2]: seed, call counter (forward)
  uint32_t* drop_mask;      // [B][heads][L][mask_words] transposed keep bits
};

// ---- CUDA-core engine (stripe_attn_simt.cu) ---------------------------------------------------
int simt_fwd(const StripeGeom& g, int dtype, const void* q, const void* k, const void* v,
             const float* lepe_w, const float* lepe_b, void* out, float* lse, cudaStream_t st);
int simt_bwd(const StripeGeom& g, int dtype, const void* q, const void* k, const void* v,
             const float* lepe_w, const float* lepe_b, const void* out, const void* gout,
             const float* lse, void* dq, void* dk, void* dv, float* gw, float* gb, float* delta,
             float* partial, cudaStream_t st);
// One pass (both engines): delta[b,h,l] = sum_c grad_out * (out - lepe) — the row term of the
// softmax gradient — plus the depthwise-3x3 weight / bias gradients (gw [C'][9], gb [C']) through
// `partial` ([wgrad_blocks][C'][10] floats of scratch).
int wgrad_blocks(const StripeGeom& g);
struct PrepIO {
  const void *v, *out, *gout;
  const float *lepe_w, *lepe_b;
  float *delta, *partial, *gw, *gb;
};
// TMA-streamed variant (lepe_prep.cu): CSB200_ERR_UNSUPPORTED means "use the generic kernel"
int lepe_prep_tma_max_blocks();
int lepe_prep_tma_launch(int nbr, const StripeGeom* g, int dtype, const PrepIO* io, int* blocks,
                         cudaStream_t st);
// up to two branches of equal (B, L) in one launch.  final_blocks == nullptr: gw / gb are complete on return
// (lepe_wgrad_final launched).  Otherwise the per-CTA partials are left in io[i].partial, *final_blocks is
// their count per output, and the CALLER sums them (the tcgen05 backward kernel does, in its prologue).
int lepe_bwd_prep_multi(int nbr, const StripeGeom* g, int dtype, const PrepIO* io, cudaStream_t st,
                        int* final_blocks = nullptr);
int lepe_bwd_prep(const StripeGeom& g, int dtype, const void* v, const float* lepe_w,
                  const float* lepe_b, const void* out, const void* gout, float* delta,
                  float* partial, float* gw, float* gb, cudaStream_t st);

// ---- tcgen05 engine (stripe_attn_tc.cu) -------------------------------------------------------
bool tc_fwd_supported(const StripeGeom& g, int dtype);
// long stripes (N = 128 T, T >= 3): key/value-tiled online-softmax forward kernel (stripe_attn_tc_kv.cu)
bool tc_fwd_kv_supported(const StripeGeom& g, int dtype);
int tc_fwd_kv(const StripeGeom& g, const void* q, const void* k, const void* v, const float* lepe_w,
              const float* lepe_b, void* out, float* lse, cudaStream_t st);
bool tc_bwd_supported(const StripeGeom& g, int dtype);
int tc_fwd(const StripeGeom& g, const void* q, const void* k, const void* v, const float* lepe_w,
           const float* lepe_b, void* out, float* lse, cudaStream_t st);
// up to two branches (the two stripe orientations of one CSWinBlock) in ONE launch; same N, batch
struct TcFwdIO {
  const void *q, *k, *v;
  const float *lepe_w, *lepe_b;
  void* out;
  float* lse;
};
int tc_fwd_multi(int nbr, const StripeGeom* g, const TcFwdIO* io, cudaStream_t st);
struct TcBwdIO {
  const void *q, *k, *v, *gout;
  const float *lepe_w, *lse, *delta;
  void *dq, *dk, *dv;
  // optional: per-CTA partials of the LePE weight / bias gradients ([blocks][C'][10], lepe_bwd_prep_multi) to be
  // summed into gw [C'][9] / gb [C'] by the kernel's prologue (nullptr: nothing to do)
  const float* wg_partial;
  float *gw, *gb;
};
int tc_bwd_multi(int nbr, const StripeGeom* g, const TcBwdIO* io, cudaStream_t st, int wg_blocks = 0);
// dq, dk, dv from q, k, v, grad_out, lse and delta (stripe_attn_tc_bwd.cu)
int tc_bwd_core(const StripeGeom& g, const void* q, const void* k, const void* v, const void* gout,
                const float* lepe_w, const float* lse, const float* delta, void* dq, void* dk,
                void* dv, cudaStream_t st);

}  // namespace csb200

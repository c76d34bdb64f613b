/*
 * csb200.h — C ABI of the B200 (sm_100a) kernels behind the CSWin-SimAM-UNet training hot path.
 *
 * The reference (TrungMasterChef/CSWin-SimAM-UNet) has no FFI of its own: its only seam is the
 * nn.Module surface in train_cswinunet_segmentation.py ("C:") and train_unet_segmentation.py ("U:").
 * Each entry point below names the reference code it replaces.  The Python host side
 * (cswin-simam-unet_b200/capi.py) binds these with ctypes; INTEGRATION.md shows the stub a reference
 * maintainer would add.
 *
 * Conventions
 *   - every function returns an int status (CSB200_OK == 0); nothing throws, nothing calls exit()
 *     (the reference's exit(0) on a bad idx, C:239-240, is deliberately not reproduced);
 *   - all data pointers are DEVICE pointers owned by the caller; no allocation happens inside;
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued and the call returns at once;
 *   - strides are in ELEMENTS; the channel stride of every token-major operand is 1;
 *   - thread safe: no mutable global state except atomically-updated counters and an error string
 *     that is thread-local.
 */
#ifndef CSB200_H_
#define CSB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSB200_ABI_VERSION 2

#if defined(__GNUC__)
#define CSB200_API __attribute__((visibility("default")))
#else
#define CSB200_API
#endif

/* status codes */
#define CSB200_OK 0
#define CSB200_ERR_INVALID 1     /* bad shape / divisibility / null pointer — mirrors the RuntimeError
                                    the reference raises from view() in img2windows, C:204 */
#define CSB200_ERR_UNSUPPORTED 2 /* valid but not built (e.g. head_dim != 32) */
#define CSB200_ERR_CUDA 3        /* a CUDA runtime / driver call failed */
#define CSB200_ERR_WORKSPACE 4   /* workspace too small */

/* element types of activations */
#define CSB200_F32 0
#define CSB200_BF16 1

/* SimAM memory layouts */
#define CSB200_NCHW 0 /* (B, C, H, W): a plane is H*W contiguous elements — UNet, U:177-250 */
#define CSB200_NLC 1  /* (B, L, C) token-major: a plane is a column of stride C — CSWin, C:349 */

/* attention engines (see csb200_stripe_attn_engine) */
#define CSB200_ENGINE_AUTO 0
#define CSB200_ENGINE_SIMT 1    /* fp32-accumulate CUDA-core kernels, any shape */
#define CSB200_ENGINE_TCGEN05 2 /* bf16 tcgen05/TMEM/TMA kernels, shapes listed in DESIGN.md */

CSB200_API int csb200_abi_version(void);
/* Thread-local text of the last error raised on this thread ("" if none). */
CSB200_API const char* csb200_last_error_string(void);
/* Number of kernels this library has launched since load (all threads). bench.py's gpu_launches. */
CSB200_API uint64_t csb200_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * SimAM — NOT present in the reference checkout; restates the public SimAM module
 * (Yang et al., ICML 2021):  n = H*W - 1;  d = (x - mean)^2;  v = sum(d)/n + e_lambda;
 *                            y = x * sigmoid(d / (4 v) + 0.5)
 * stats (optional in fwd, required in bwd): float[2 * B * C], index b*C + c — state saved for the
 * backward pass: {mean - x_first, v}, x_first being the plane's first spatial element (the mean is
 * kept relative to that pivot so that large-mean planes lose no precision).
 * `spatial` = H*W (NCHW) or L (NLC).
 * ---------------------------------------------------------------------------------------------- */
CSB200_API int csb200_simam_fwd(const void* x, void* y, float* stats, int64_t batch, int64_t channels,
                     int64_t spatial, int layout, int dtype, float e_lambda, void* stream);
CSB200_API int csb200_simam_bwd(const void* x, const void* grad_y, const float* stats, void* grad_x,
                     int64_t batch, int64_t channels, int64_t spatial, int layout, int dtype,
                     float e_lambda, void* stream);
/* The same two passes with a caller-owned workspace, which lets large token-layout (NLC) tensors take the
 * grid-resident kernels (csrc/simam_grid.cuh: all SMs work on a few images at a time, every byte crosses HBM once,
 * reads of the next images overlap the writes of the current ones).  csb200_simam_workspace_bytes returns the
 * size to allocate for a shape (0: these kernels do not apply to it; the _ws calls then behave like the plain ones,
 * as they do with workspace == NULL or a workspace that is too small).
 * CONTRACT: the workspace must be zero-filled ONCE, when it is allocated; every call leaves it ready for the
 * next call of any shape.  Calls that share a workspace must be ordered on one stream. */
CSB200_API size_t csb200_simam_workspace_bytes(int64_t batch, int64_t channels, int64_t spatial, int layout,
                     int dtype);
CSB200_API int csb200_simam_fwd_ws(const void* x, void* y, float* stats, int64_t batch, int64_t channels,
                     int64_t spatial, int layout, int dtype, float e_lambda, void* workspace,
                     size_t workspace_bytes, void* stream);
CSB200_API int csb200_simam_bwd_ws(const void* x, const void* grad_y, const float* stats, void* grad_x,
                     int64_t batch, int64_t channels, int64_t spatial, int layout, int dtype,
                     float e_lambda, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Token-major LayerNorm over the last dimension — the pre-norms at the CSWinBlock call site (C:357,
 * C:368) and C:386, C:507, C:648, C:671: y = (x - mean) * rstd * gamma + beta, biased variance.
 * x: [rows][channels] contiguous, type x_dtype; y has its own type (fp32 residual stream -> bf16 GEMM
 * operand in one pass); gamma / beta / their gradients fp32; stats: float[2*rows] = {mean, rstd}.
 * grad_x has the type of x.  csb200_layernorm_supported tells whether `channels` is tiled (16-byte
 * vectors per row in {8,16,32,64,128}); otherwise the functions return CSB200_ERR_UNSUPPORTED.
 * ---------------------------------------------------------------------------------------------- */
CSB200_API int csb200_layernorm_supported(int64_t channels, int x_dtype);
CSB200_API int csb200_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y,
                                    float* stats, int64_t rows, int64_t channels, int x_dtype,
                                    int y_dtype, float eps, void* stream);
CSB200_API size_t csb200_layernorm_bwd_workspace_bytes(int64_t rows, int64_t channels);
CSB200_API int csb200_layernorm_bwd(const void* x, const void* grad_y, const float* gamma,
                                    const float* stats, void* grad_x, float* grad_gamma,
                                    float* grad_beta, void* workspace, size_t workspace_bytes,
                                    int64_t rows, int64_t channels, int x_dtype, int gy_dtype,
                                    void* stream);
/* Residual add fused into the pre-norm that follows it (C:367 -> C:368, C:369 -> the next block's
 * C:357):  sum_out = x + residual (type of x, written once),  y = LayerNorm(sum_out).
 * Backward takes the gradient arriving on `sum` from later layers (grad_sum, type of x) and returns
 *   grad_x = LayerNorm'(sum; grad_y) + grad_sum,  which is the gradient of BOTH x and residual. */
CSB200_API int csb200_add_layernorm_fwd(const void* x, const void* residual, void* sum_out,
                                        const float* gamma, const float* beta, void* y, float* stats,
                                        int64_t rows, int64_t channels, int x_dtype, int y_dtype,
                                        float eps, void* stream);
CSB200_API int csb200_add_layernorm_bwd(const void* sum, const void* grad_y, const void* grad_sum,
                                        const float* gamma, const float* stats, void* grad_x,
                                        float* grad_gamma, float* grad_beta, void* workspace,
                                        size_t workspace_bytes, int64_t rows, int64_t channels,
                                        int x_dtype, int gy_dtype, void* stream);
/* Same pass, which also returns grad_res_bias[c] = sum over rows of grad_x[r][c] (fp32): when `residual`
 * was the output of a Linear (proj C:366 -> C:367, Mlp.fc2 C:195 -> C:369) that is the bias gradient of
 * that Linear, which then needs no column-sum pass over the same tensor. */
CSB200_API int csb200_add_layernorm_bwd_rb(const void* sum, const void* grad_y, const void* grad_sum,
                                           const float* gamma, const float* stats, void* grad_x,
                                           float* grad_gamma, float* grad_beta, float* grad_res_bias,
                                           void* workspace, size_t workspace_bytes, int64_t rows,
                                           int64_t channels, int x_dtype, int gy_dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Column sums of a row-major [rows][cols] matrix (fp32 out) — the bias gradient of the Linear layers
 * on the token path (C:358, C:366, C:188-196, C:658): grad_bias = sum over tokens of grad_out.
 * Supported when cols is a multiple of the 16-byte vector width and cols / width <= 256.
 * ---------------------------------------------------------------------------------------------- */
CSB200_API int csb200_colsum_supported(int64_t cols, int dtype);
CSB200_API size_t csb200_colsum_workspace_bytes(int64_t cols);
CSB200_API int csb200_colsum(const void* x, float* out, void* workspace, size_t workspace_bytes,
                             int64_t rows, int64_t cols, int dtype, void* stream);
/* y[r][c] = x[r][c] + bias[c] (y may alias x): the bias add of a convolution on a channels-last tensor
 * (rows = B*H*W, cols = channels), which cuDNN leaves to a broadcasting add_.  Same tiling rule. */
CSB200_API int csb200_add_row_bias(const void* x, const float* bias, void* y, int64_t rows, int64_t cols,
                                   int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GELU of the Mlp hidden layer (Mlp.forward C:188-196, act_layer() = nn.GELU in its exact erf form)
 * on a row-major [rows][cols] matrix: forward in one pass; backward writes grad_h = grad_out * gelu'(h)
 * AND the column sums of grad_h (the bias gradient of fc1) in the same pass.  Same tiling rule as
 * csb200_colsum (cols a multiple of the 16-byte vector width, cols / width <= 256).
 * ---------------------------------------------------------------------------------------------- */
CSB200_API int csb200_gelu_supported(int64_t cols, int dtype);
CSB200_API int csb200_gelu_fwd(const void* h, void* out, int64_t rows, int64_t cols, int dtype,
                               void* stream);
CSB200_API size_t csb200_gelu_bwd_workspace_bytes(int64_t cols);
CSB200_API int csb200_gelu_bwd(const void* grad_out, const void* h, void* grad_h, float* grad_bias,
                               void* workspace, size_t workspace_bytes, int64_t rows, int64_t cols,
                               int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Token-path Linear on the tcgen05 tensor cores: y = epilogue(x W^T + b) for the nn.Linear layers of the
 * CSWinBlock whose contraction is the block width — `qkv` (C:357-358), `proj` (C:366), `Mlp.fc1` + `act`
 * (C:188-196) — SURVEY.md section 8(f)-2.  x: [M][K] tokens (row stride ldx elements), weight: [N][K]
 * (nn.Linear.weight), bias: fp32 [N] or NULL, y: [M][N].  bf16 activations / weight, fp32 accumulation.
 * epilogue: CSB200_EPI_BIAS       y = x W^T + b
 *           CSB200_EPI_GELU       y = GELU(bf16(x W^T + b))   exact erf form, nn.GELU()
 *           CSB200_EPI_GELU_SAVE  as above and pre_act = bf16(x W^T + b)  (what GELU' needs in backward)
 *           CSB200_EPI_GELU_SAVE_DERIV  as CSB200_EPI_GELU and pre_act = bf16(GELU'(bf16(x W^T + b))): the
 *                                 derivative itself, for csb200_linear_dact_bwd
 * Supported (csb200_linear_supported): bf16, K in {64, 128, 256}, N a multiple of 32.
 * ---------------------------------------------------------------------------------------------- */
#define CSB200_EPI_BIAS 0
#define CSB200_EPI_GELU 1
#define CSB200_EPI_GELU_SAVE 2
#define CSB200_EPI_GELU_SAVE_DERIV 3
CSB200_API int csb200_linear_supported(int64_t M, int64_t N, int64_t K, int dtype);
CSB200_API int csb200_linear_fwd(const void* x, const void* weight, const float* bias, void* y, void* pre_act,
                                 int64_t M, int64_t N, int64_t K, int64_t ldx, int dtype, int epilogue,
                                 void* stream);
/* Backward of `Mlp.fc2(act(h))` with respect to h (C:190-195) as ONE tcgen05 GEMM with the GELU' factor and
 * the fc1 bias gradient in its epilogue:
 *     grad_h = (grad_y W2) * GELU'(pre_act),     grad_bias[n] = sum_m grad_h[m][n]   (of the ROUNDED grad_h)
 * grad_y: [M][K] (row stride ldg), weight = fc2.weight: [K][N] (nn.Linear(N, K).weight, read in place as the
 * MN-major operand: no transposed copy), pre_act / grad_h: [M][N], grad_bias: fp32 [N].  Replaces the cuBLAS
 * input-gradient GEMM + the flat csb200_gelu_bwd pass (1 write + 1 read of the 4C-wide tensor less).
 * Supported: bf16, K in {64, 128, 256}, N a multiple of 64.  workspace: csb200_linear_dgelu_workspace_bytes(N). */
CSB200_API int csb200_linear_dgelu_supported(int64_t M, int64_t N, int64_t K, int dtype);
CSB200_API size_t csb200_linear_dgelu_workspace_bytes(int64_t N);
CSB200_API int csb200_linear_dgelu_bwd(const void* grad_y, const void* weight, const void* pre_act, void* grad_h,
                                       float* grad_bias, void* workspace, size_t workspace_bytes, int64_t M,
                                       int64_t N, int64_t K, int64_t ldg, int dtype, void* stream);
/* The same input-gradient GEMM for a forward pass that saved the activation's DERIVATIVE instead of the
 * pre-activation (csb200_linear_fwd with CSB200_EPI_GELU_SAVE_DERIV writes GELU'(h), bf16, into `pre_act`):
 *     grad_h = (grad_y W2) * act_deriv,   grad_bias[n] = sum_m grad_h[m][n]
 * The epilogue is one multiplication per element instead of a second erf + exponential (C:190 backward), which
 * is what bounds csb200_linear_dgelu_bwd.  Same shapes, workspace and alignment rules. */
CSB200_API int csb200_linear_dact_bwd(const void* grad_y, const void* weight, const void* act_deriv, void* grad_h,
                                      float* grad_bias, void* workspace, size_t workspace_bytes, int64_t M,
                                      int64_t N, int64_t K, int64_t ldg, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Deferred final sums.  csb200_layernorm_bwd / _add_layernorm_bwd(_rb), csb200_colsum and
 * csb200_linear_dgelu_bwd / _dact_bwd all end with the same tiny launch: out[c] = sum over the per-CTA partial
 * rows of partial[r][c] (fixed order).  One backward pass of the 512^2 model issues 106 of them, ~4-5 us each
 * on the critical path of the stream, for results nobody reads before the optimizer step.  The `_partials`
 * twins below stop after the main kernel and hand back where the partial rows are; the caller records the sums
 * it wants with csb200_sum_rows_deferred and ONE launch per 120 records performs them all at
 * csb200_sum_rows_flush (same summation order as the immediate kernels: bit-identical results).
 * CONTRACT: the workspace holding the partial rows and the `out` vectors must stay allocated and untouched
 * until the flush; the record list is process-wide (one deferring client at a time), guarded by a mutex;
 * records and flush must be issued in stream order on the same stream (or a stream ordered after it).
 *   *partials      : float[*partial_rows][row_stride], inside `workspace`
 *   layernorm      : row_stride = (2 + with_res_bias) * channels: grad_gamma | grad_beta | grad_res_bias
 *   colsum         : row_stride = cols
 *   linear dact    : row_stride = N (fc1 bias gradient); use_saved_derivative selects _dact_ (1) or _dgelu_ (0)
 * ---------------------------------------------------------------------------------------------- */
CSB200_API int csb200_layernorm_bwd_partials(const void* sum_or_x, const void* grad_y, const void* grad_sum,
                                             const float* gamma, const float* stats, void* grad_x,
                                             int with_res_bias, void* workspace, size_t workspace_bytes,
                                             int64_t rows, int64_t channels, int x_dtype, int gy_dtype,
                                             const float** partials, int32_t* partial_rows, void* stream);
CSB200_API int csb200_colsum_partials(const void* x, void* workspace, size_t workspace_bytes, int64_t rows,
                                      int64_t cols, int dtype, const float** partials, int32_t* partial_rows,
                                      void* stream);
CSB200_API int csb200_linear_dact_bwd_partials(const void* grad_y, const void* weight, const void* act, void* grad_h,
                                               void* workspace, size_t workspace_bytes, int64_t M, int64_t N,
                                               int64_t K, int64_t ldg, int dtype, int use_saved_derivative,
                                               const float** partials, int32_t* partial_rows, void* stream);
/* record: out[c] = sum_{r < rows} partial[r * row_stride + c] for c < cols (performed at the next flush) */
CSB200_API int csb200_sum_rows_deferred(const float* partial, int64_t rows, int64_t cols, int64_t row_stride,
                                        float* out);
CSB200_API int64_t csb200_sum_rows_pending(void);
CSB200_API int csb200_sum_rows_flush(void* stream);  /* performs and clears the records */
CSB200_API int csb200_sum_rows_discard(void);        /* clears them without performing (error paths) */

/* Parameter gradients of a token-path nn.Linear (backward of C:357-358, C:366, C:188-196 w.r.t. weight / bias):
 *     grad_w[N][K] = grad_y[M][N]^T x[M][K],      grad_bias[n] = sum_m grad_y[m][n]   (nullable)
 * bf16 operands (row strides ldg / ldx elements), fp32 outputs, OVERWRITTEN (zeroed inside, then accumulated by
 * vector reductions: the summation order over the token splits is not fixed).  One streaming tcgen05 pass
 * over both operands, the bias gradient included; replaces the split-K GEMM + reduce kernel + column-sum pass.
 * Supported: bf16, N a multiple of 8, K a multiple of 64. */
CSB200_API int csb200_linear_wgrad_supported(int64_t M, int64_t N, int64_t K, int dtype);
CSB200_API int csb200_linear_wgrad(const void* grad_y, const void* x, float* grad_w, float* grad_bias, int64_t M,
                                   int64_t N, int64_t K, int64_t ldg, int64_t ldx, int dtype, void* stream);
/* The same pass ADDING into grad_w / grad_bias instead of overwriting them (no memsets inside): for outputs carved
 * from an arena the caller zeroed once for the whole backward pass, or for gradient accumulation. */
CSB200_API int csb200_linear_wgrad_acc(const void* grad_y, const void* x, float* grad_w, float* grad_bias, int64_t M,
                                   int64_t N, int64_t K, int64_t ldg, int64_t ldx, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer step for every parameter tensor of the model in one launch — `optimizer.step()` of the
 * reference train loop (C:786) with torch.optim.AdamW (C:937-941, decoupled decay) or torch.optim.Adam
 * (U:486-490, L2 decay); fp32 parameters, gradients and moments.  `shadow` (nullable) receives the bf16
 * rounding of the updated parameter in the same pass.
 *   tensors_dev : device array of csb200_adam_tensor
 *   chunks_dev  : device array of n_chunks (tensor index, chunk index) int32 pairs; a chunk is
 *                 csb200_adam_chunk_elems() consecutive elements of one tensor
 *   hyper_dev   : device array [groups][8] = {lr, beta1, beta2, eps, weight_decay, decoupled (0/1), -, -}
 *   step_dev    : device scalar, the step count INCLUDING this step (1 on the first call)
 * ---------------------------------------------------------------------------------------------- */
typedef struct csb200_adam_tensor {
  void* param;
  const void* grad;
  void* exp_avg;
  void* exp_avg_sq;
  void* shadow;
  int64_t numel;
  int32_t group;
  int32_t reserved;
} csb200_adam_tensor;
CSB200_API int64_t csb200_adam_chunk_elems(void);
CSB200_API int csb200_adam_step(const csb200_adam_tensor* tensors_dev, const int32_t* chunks_dev,
                                int64_t n_chunks, const float* hyper_dev, const float* step_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused CARAFE reassembly — CARAFE.forward C:406-431 / CARAFE4 C:455-480 without the pixel-shuffled
 * logits, the fp32 softmax tensor, the 9x unfold and the batched 9 x up^2 matmul:
 *   out[b, h*up+dy, w*up+dx, c] = sum_tap softmax_tap(enc[b, h, w, tap*up^2 + dy*up + dx])
 *                                         * low[b, h+ky-1, w+kx-1, c]        (zero padding)
 * All tensors channels-last (token-major), one type T: low [B][H][W][C], enc [B][H][W][9 up^2]
 * (the raw `encoder` conv output, C:407), out / grad_out [B][H up][W up][C], weights
 * [B][H up][W up][9] (the softmax result, written by fwd when non-NULL, required by bwd).
 * channels == 1 or a multiple of 8.
 * ---------------------------------------------------------------------------------------------- */
CSB200_API int csb200_carafe_supported(int64_t channels);
CSB200_API int csb200_carafe_fwd(const void* low, const void* enc, void* out, void* weights,
                                 int64_t batch, int64_t height, int64_t width, int64_t channels,
                                 int up, int dtype, void* stream);
CSB200_API int csb200_carafe_bwd(const void* low, const void* weights, const void* grad_out,
                                 void* grad_low, void* grad_enc, int64_t batch, int64_t height,
                                 int64_t width, int64_t channels, int up, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Cross-shaped stripe attention with LePE — replaces the body of LePEAttention.forward (C:271-298)
 * including im2cswin (C:248-254), get_lepe (C:256-269), img2windows / windows2img (C:199-217) and, on
 * the caller's side, the torch.cat of the two branches (C:363): q/k/v are read in place from the
 * (B, L, 3C) qkv buffer through strides and `out` is written at the branch's channel offset.
 *
 *   token l = y*width + x belongs to stripe (y / h_sp, x / w_sp), position (y % h_sp)*w_sp + x % w_sp
 *   channel c of the branch belongs to head c / head_dim
 *   out[b,l,:] = softmax(scale * q k^T) v  +  depthwise3x3(v; lepe_w, lepe_b) with ZERO padding at
 *   the STRIPE border (C:244,263-265)
 *
 * lepe_w: float[C'][3][3] (the get_v.weight parameter, (C',1,3,3)), lepe_b: float[C'] — always fp32.
 * lse:    float[B][heads][L], log-sum-exp of the scaled scores per query row (saved for backward).
 * ---------------------------------------------------------------------------------------------- */
typedef struct csb200_stripe_desc {
  int32_t dtype;          /* CSB200_F32 / CSB200_BF16: type of q, k, v, out and their gradients */
  int32_t batch;          /* B */
  int32_t height, width;  /* token grid; L = height * width (C:276-281) */
  int32_t h_sp, w_sp;     /* stripe extent (C:232-242); height % h_sp == 0 and width % w_sp == 0 */
  int32_t heads;          /* heads of THIS branch */
  int32_t head_dim;       /* channels per head; C' = heads * head_dim */
  float scale;            /* qk scale, head_dim^-0.5 unless qk_scale was given (C:231) */
  int32_t engine;         /* CSB200_ENGINE_* request; AUTO picks tcgen05 when the shape allows */
  int64_t q_sb, q_sl;     /* batch / token strides of q (elements) */
  int64_t k_sb, k_sl;
  int64_t v_sb, v_sl;
  int64_t o_sb, o_sl;     /* out and grad_out */
  int64_t dq_sb, dq_sl;   /* gradients (backward only) */
  int64_t dk_sb, dk_sl;
  int64_t dv_sb, dv_sl;
  /* Attention dropout (`attn = self.attn_drop(attn)`, C:290; the reference trains with attn_drop_rate 0.3,
   * C:930-932).  drop_p == 0: off, the three fields below are ignored.  Otherwise every softmax probability is
   * zeroed with probability p = round(256 drop_p) / 256 and the survivors are scaled by 1 / (1 - p); the keep
   * decisions come from Philox4x32-10 keyed by rng_state[0] (seed) at counter (key block, query, stripe-and-head
   * unit, rng_state[1] = call counter) and are written by the FORWARD call to drop_mask as transposed bit rows —
   * word [((b * heads + head) * L + key token) * ceil(N / 32) + query / 32], bit query % 32 — which the
   * BACKWARD call reads back (same mask by construction, and a test can read it too). */
  float drop_p;
  int32_t drop_salt;            /* distinguishes the branches of one block that share a call counter */
  const uint64_t* rng_state;    /* device, [2]: seed, call counter (forward only) */
  uint32_t* drop_mask;          /* device, uint32[B][heads][L][ceil(N / 32)] */
} csb200_stripe_desc;

/* Which engine a descriptor resolves to (CSB200_ENGINE_SIMT / _TCGEN05), or a negative status. */
CSB200_API int csb200_stripe_attn_engine(const csb200_stripe_desc* d, int backward);

CSB200_API int csb200_stripe_attn_fwd(const csb200_stripe_desc* d, const void* q, const void* k, const void* v,
                           const float* lepe_w, const float* lepe_b, void* out, float* lse,
                           void* stream);

/* Bytes of scratch csb200_stripe_attn_bwd needs for this descriptor. */
CSB200_API size_t csb200_stripe_attn_bwd_workspace_bytes(const csb200_stripe_desc* d);

/* grad_lepe_w: float[C'][3][3], grad_lepe_b: float[C'] — overwritten (not accumulated).
 * dq/dk/dv are overwritten. `out` is the forward result (needed for the softmax-gradient row term). */
CSB200_API int csb200_stripe_attn_bwd(const csb200_stripe_desc* d, const void* q, const void* k, const void* v,
                           const float* lepe_w, const float* lepe_b, const void* out,
                           const void* grad_out, const float* lse, void* dq, void* dk, void* dv,
                           float* grad_lepe_w, float* grad_lepe_b, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * All branches of one CSWinBlock (C:360-363: the two stripe orientations on the two channel halves,
 * or the single full-window branch of the last stage) in one call.  Two tcgen05-eligible branches of
 * equal stripe length run as ONE launch whose work items are interleaved image by image (both halves
 * of every 128-byte line of the packed qkv buffer are consumed together; no half-empty tail wave);
 * anything else runs branch by branch exactly like the single-branch entry points.
 * io[i] carries the pointers of branch i; fields not used by a direction may be NULL.
 * ---------------------------------------------------------------------------------------------- */
typedef struct csb200_branch_io {
  const void *q, *k, *v;            /* already offset to the branch's first channel */
  const float *lepe_w, *lepe_b;
  void* out;                        /* forward: written; backward: read */
  float* lse;                       /* forward: written; backward: read */
  const void* grad_out;             /* backward only from here on */
  void *dq, *dk, *dv;
  float *grad_lepe_w, *grad_lepe_b;
  void* workspace;
  size_t workspace_bytes;           /* >= csb200_stripe_attn_bwd_workspace_bytes(&descs[i]) */
} csb200_branch_io;

CSB200_API int csb200_cross_stripe_attn_fwd(int n_branches, const csb200_stripe_desc* descs,
                                            const csb200_branch_io* io, void* stream);
CSB200_API int csb200_cross_stripe_attn_bwd(int n_branches, const csb200_stripe_desc* descs,
                                            const csb200_branch_io* io, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CSB200_H_ */

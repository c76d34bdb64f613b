// C-ABI entry points of the stripe attention (include/csb200.h): validation, engine choice,
// workspace carving.  The arithmetic lives in stripe_attn_simt.cu / stripe_attn_tc.cu.

#include <mutex>
#include <unordered_map>

#include "stripe_attn.cuh"

namespace csb200 {

thread_local char g_err[512] = "";

namespace {
std::mutex g_memo_mu;
std::unordered_map<uint64_t, int> g_memo;
inline uint64_t memo_key(int dev, const void* key, int tag) {
  uint64_t h = reinterpret_cast<uint64_t>(key) * 0x9E3779B97F4A7C15ull;
  h ^= (static_cast<uint64_t>(static_cast<uint32_t>(dev)) << 48) ^ (static_cast<uint64_t>(static_cast<uint32_t>(tag)) << 32);
  return h ^ (h >> 29);
}
const char kSmCountKey = 0, kSmemKey = 0;
}  // namespace

bool memo_get(const void* key, int tag, int* val) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  std::lock_guard<std::mutex> lock(g_memo_mu);
  auto it = g_memo.find(memo_key(dev, key, tag));
  if (it == g_memo.end()) return false;
  *val = it->second;
  return true;
}
void memo_put(const void* key, int tag, int val) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  std::lock_guard<std::mutex> lock(g_memo_mu);
  g_memo[memo_key(dev, key, tag)] = val;
}
int device_sm_count() {
  int n = 0;
  if (memo_get(&kSmCountKey, 0, &n)) return n;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return -1;
  memo_put(&kSmCountKey, 0, n);
  return n;
}
cudaError_t opt_in_smem(const void* func, int bytes) {
  int have = 0;
  if (memo_get(func, 1, &have) && have >= bytes) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) memo_put(func, 1, bytes);
  (void)kSmemKey;
  return e;
}
std::atomic<uint64_t> g_launches{0};

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Mirrors the failures of the reference: a resolution that the stripe does not divide raises
// RuntimeError from view() in img2windows (C:204); here it is CSB200_ERR_INVALID + message.
int make_geom(const csb200_stripe_desc* d, bool backward, StripeGeom* g) {
  if (d == nullptr) return fail(CSB200_ERR_INVALID, "stripe_attn: null descriptor");
  if (d->dtype != CSB200_F32 && d->dtype != CSB200_BF16)
    return fail(CSB200_ERR_INVALID, "stripe_attn: unknown dtype %d", d->dtype);
  if (d->batch < 0 || d->height <= 0 || d->width <= 0 || d->h_sp <= 0 || d->w_sp <= 0 ||
      d->heads <= 0)
    return fail(CSB200_ERR_INVALID, "stripe_attn: non-positive size");
  if (d->height % d->h_sp != 0 || d->width % d->w_sp != 0)
    return fail(CSB200_ERR_INVALID,
                "stripe_attn: token grid %dx%d is not divisible by the stripe %dx%d "
                "(the reference raises from view() in img2windows for the same input)",
                d->height, d->width, d->h_sp, d->w_sp);
  if (d->head_dim != 32)
    return fail(CSB200_ERR_UNSUPPORTED, "stripe_attn: head_dim %d (only 32 is built)",
                d->head_dim);
  if ((int64_t)d->height * d->width > 0x7fffffff / 4 ||
      (int64_t)d->batch * d->height * d->width > 0x7fffffff)
    return fail(CSB200_ERR_INVALID, "stripe_attn: token grid too large");
  const int64_t strides[] = {d->q_sb, d->q_sl, d->k_sb, d->k_sl, d->v_sb, d->v_sl, d->o_sb, d->o_sl};
  for (int64_t s : strides)
    if (s % 8 != 0 || s < 0)
      return fail(CSB200_ERR_INVALID, "stripe_attn: strides must be non-negative multiples of 8");
  if (backward) {
    const int64_t gs[] = {d->dq_sb, d->dq_sl, d->dk_sb, d->dk_sl, d->dv_sb, d->dv_sl};
    for (int64_t s : gs)
      if (s % 8 != 0 || s < 0)
        return fail(CSB200_ERR_INVALID, "stripe_attn: gradient strides must be multiples of 8");
  }
  g->B = d->batch;
  g->H = d->height;
  g->W = d->width;
  g->L = d->height * d->width;
  g->hs = d->h_sp;
  g->ws = d->w_sp;
  g->N = d->h_sp * d->w_sp;
  g->nwy = d->height / d->h_sp;
  g->nwx = d->width / d->w_sp;
  g->heads = d->heads;
  g->scale = d->scale;
  g->q_sb = d->q_sb; g->q_sl = d->q_sl;
  g->k_sb = d->k_sb; g->k_sl = d->k_sl;
  g->v_sb = d->v_sb; g->v_sl = d->v_sl;
  g->o_sb = d->o_sb; g->o_sl = d->o_sl;
  g->dq_sb = d->dq_sb; g->dq_sl = d->dq_sl;
  g->dk_sb = d->dk_sb; g->dk_sl = d->dk_sl;
  g->dv_sb = d->dv_sb; g->dv_sl = d->dv_sl;
  g->drop_thr = 0;
  g->keep_scale = 1.f;
  g->mask_words = (g->N + 31) / 32;
  g->drop_salt = (uint32_t)d->drop_salt;
  g->rng = nullptr;
  g->drop_mask = nullptr;
  if (d->drop_p != 0.f) {
    if (!(d->drop_p > 0.f) || !(d->drop_p < 1.f))
      return fail(CSB200_ERR_INVALID, "stripe_attn: drop_p must be in [0, 1), got %g", (double)d->drop_p);
    const int thr = (int)(d->drop_p * 256.f + 0.5f);
    if (thr > 0) {
      if (d->drop_mask == nullptr || (!backward && d->rng_state == nullptr))
        return fail(CSB200_ERR_INVALID, "stripe_attn: drop_p > 0 needs drop_mask (and rng_state in forward)");
      g->drop_thr = (uint32_t)(thr > 255 ? 255 : thr);
      g->keep_scale = 256.f / (256.f - (float)g->drop_thr);
      g->rng = reinterpret_cast<const unsigned long long*>(d->rng_state);
      g->drop_mask = d->drop_mask;
    }
  }
  return CSB200_OK;
}

bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

int pick_engine(const csb200_stripe_desc* d, const StripeGeom& g, bool backward) {
  const bool tc_ok = backward ? tc_bwd_supported(g, d->dtype) : tc_fwd_supported(g, d->dtype);
  switch (d->engine) {
    case CSB200_ENGINE_AUTO:
      return tc_ok ? CSB200_ENGINE_TCGEN05 : CSB200_ENGINE_SIMT;
    case CSB200_ENGINE_SIMT:
      return CSB200_ENGINE_SIMT;
    case CSB200_ENGINE_TCGEN05:
      if (!tc_ok)
        return -fail(CSB200_ERR_UNSUPPORTED,
                     "stripe_attn: the tcgen05 engine does not tile this shape (N=%d, dtype=%d)",
                     g.N, d->dtype);
      return CSB200_ENGINE_TCGEN05;
    default:
      return -fail(CSB200_ERR_INVALID, "stripe_attn: unknown engine %d", d->engine);
  }
}

}  // namespace
}  // namespace csb200

using namespace csb200;

extern "C" int csb200_abi_version(void) { return CSB200_ABI_VERSION; }
extern "C" const char* csb200_last_error_string(void) { return g_err; }
extern "C" uint64_t csb200_launch_count(void) { return g_launches.load(); }

extern "C" int csb200_stripe_attn_engine(const csb200_stripe_desc* d, int backward) {
  StripeGeom g;
  int rc = make_geom(d, backward != 0, &g);
  if (rc != CSB200_OK) return -rc;
  return pick_engine(d, g, backward != 0);
}

extern "C" int csb200_stripe_attn_fwd(const csb200_stripe_desc* d, const void* q, const void* k,
                                      const void* v, const float* lepe_w, const float* lepe_b,
                                      void* out, float* lse, void* stream) {
  StripeGeom g;
  int rc = make_geom(d, false, &g);
  if (rc != CSB200_OK) return rc;
  if (g.B == 0) return CSB200_OK;
  if (!q || !k || !v || !lepe_w || !lepe_b || !out || !lse)
    return fail(CSB200_ERR_INVALID, "stripe_attn_fwd: null pointer");
  if (!aligned(q, 16) || !aligned(k, 16) || !aligned(v, 16) || !aligned(out, 16))
    return fail(CSB200_ERR_INVALID, "stripe_attn_fwd: q/k/v/out must be 16-byte aligned");
  const int engine = pick_engine(d, g, false);
  if (engine < 0) return -engine;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (engine == CSB200_ENGINE_TCGEN05) return tc_fwd(g, q, k, v, lepe_w, lepe_b, out, lse, st);
  return simt_fwd(g, d->dtype, q, k, v, lepe_w, lepe_b, out, lse, st);
}

extern "C" size_t csb200_stripe_attn_bwd_workspace_bytes(const csb200_stripe_desc* d) {
  StripeGeom g;
  if (make_geom(d, true, &g) != CSB200_OK) return 0;
  const size_t delta = align_up((size_t)g.B * g.heads * g.L * sizeof(float), 256);
  const size_t partial = align_up((size_t)wgrad_blocks(g) * g.heads * 32 * 10 * sizeof(float), 256);
  return delta + partial;
}

extern "C" int csb200_stripe_attn_bwd(const csb200_stripe_desc* d, const void* q, const void* k,
                                      const void* v, const float* lepe_w, const float* lepe_b,
                                      const void* out, const void* grad_out, const float* lse,
                                      void* dq, void* dk, void* dv, float* grad_lepe_w,
                                      float* grad_lepe_b, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  StripeGeom g;
  int rc = make_geom(d, true, &g);
  if (rc != CSB200_OK) return rc;
  if (g.B == 0) return CSB200_OK;
  if (!q || !k || !v || !lepe_w || !lepe_b || !out || !grad_out || !lse || !dq || !dk || !dv ||
      !grad_lepe_w || !grad_lepe_b || !workspace)
    return fail(CSB200_ERR_INVALID, "stripe_attn_bwd: null pointer");
  const void* ptrs[] = {q, k, v, out, grad_out, dq, dk, dv, workspace};
  for (const void* p : ptrs)
    if (!aligned(p, 16))
      return fail(CSB200_ERR_INVALID, "stripe_attn_bwd: tensors must be 16-byte aligned");
  const size_t need = csb200_stripe_attn_bwd_workspace_bytes(d);
  if (workspace_bytes < need)
    return fail(CSB200_ERR_WORKSPACE, "stripe_attn_bwd: workspace %zu < %zu bytes",
                workspace_bytes, need);
  const int engine = pick_engine(d, g, true);
  if (engine < 0) return -engine;
  float* delta = static_cast<float*>(workspace);
  float* partial = reinterpret_cast<float*>(
      static_cast<char*>(workspace) + align_up((size_t)g.B * g.heads * g.L * sizeof(float), 256));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (engine == CSB200_ENGINE_TCGEN05) {
    if ((rc = lepe_bwd_prep(g, d->dtype, v, lepe_w, lepe_b, out, grad_out, delta, partial,
                            grad_lepe_w, grad_lepe_b, st)) != CSB200_OK)
      return rc;
    return tc_bwd_core(g, q, k, v, grad_out, lepe_w, lse, delta, dq, dk, dv, st);
  }
  return simt_bwd(g, d->dtype, q, k, v, lepe_w, lepe_b, out, grad_out, lse, dq, dk, dv,
                  grad_lepe_w, grad_lepe_b, delta, partial, st);
}

// ---- all branches of a block -------------------------------------------------------------------
namespace {
bool mergeable(const csb200_stripe_desc* d, const StripeGeom* g, int n, bool backward) {
  if (n != 2) return false;
  for (int i = 0; i < 2; ++i)
    if (pick_engine(&d[i], g[i], backward) != CSB200_ENGINE_TCGEN05) return false;
  if (g[0].N != 128 && g[0].N != 256) return false;  // long stripes: one launch per branch (key/value-tiled kernel)
  return g[0].N == g[1].N && g[0].B == g[1].B && g[0].H == g[1].H && g[0].W == g[1].W &&
         g[0].scale == g[1].scale && d[0].dtype == d[1].dtype && g[0].drop_thr == g[1].drop_thr;
}
}  // namespace

extern "C" int csb200_cross_stripe_attn_fwd(int n, const csb200_stripe_desc* d,
                                            const csb200_branch_io* io, void* stream) {
  if (n < 1 || n > 2 || !d || !io) return fail(CSB200_ERR_INVALID, "cross_stripe_attn: 1 or 2 branches");
  StripeGeom g[2];
  for (int i = 0; i < n; ++i) {
    int rc = make_geom(&d[i], false, &g[i]);
    if (rc != CSB200_OK) return rc;
  }
  if (g[0].B > 0 && mergeable(d, g, n, false)) {
    TcFwdIO t[2];
    for (int i = 0; i < 2; ++i) {
      if (!io[i].q || !io[i].k || !io[i].v || !io[i].lepe_w || !io[i].lepe_b || !io[i].out || !io[i].lse)
        return fail(CSB200_ERR_INVALID, "cross_stripe_attn_fwd: null pointer");
      if (!aligned(io[i].q, 16) || !aligned(io[i].k, 16) || !aligned(io[i].v, 16) || !aligned(io[i].out, 16))
        return fail(CSB200_ERR_INVALID, "cross_stripe_attn_fwd: q/k/v/out must be 16-byte aligned");
      t[i] = TcFwdIO{io[i].q, io[i].k, io[i].v, io[i].lepe_w, io[i].lepe_b, io[i].out, io[i].lse};
    }
    return tc_fwd_multi(2, g, t, static_cast<cudaStream_t>(stream));
  }
  for (int i = 0; i < n; ++i) {
    int rc = csb200_stripe_attn_fwd(&d[i], io[i].q, io[i].k, io[i].v, io[i].lepe_w, io[i].lepe_b,
                                    io[i].out, io[i].lse, stream);
    if (rc != CSB200_OK) return rc;
  }
  return CSB200_OK;
}

extern "C" int csb200_cross_stripe_attn_bwd(int n, const csb200_stripe_desc* d,
                                            const csb200_branch_io* io, void* stream) {
  if (n < 1 || n > 2 || !d || !io) return fail(CSB200_ERR_INVALID, "cross_stripe_attn: 1 or 2 branches");
  StripeGeom g[2];
  for (int i = 0; i < n; ++i) {
    int rc = make_geom(&d[i], true, &g[i]);
    if (rc != CSB200_OK) return rc;
  }
  if (g[0].B > 0 && mergeable(d, g, n, true)) {
    PrepIO pio[2];
    TcBwdIO tio[2];
    for (int i = 0; i < 2; ++i) {
      const csb200_branch_io& b = io[i];
      if (!b.q || !b.k || !b.v || !b.lepe_w || !b.lepe_b || !b.out || !b.grad_out || !b.lse || !b.dq ||
          !b.dk || !b.dv || !b.grad_lepe_w || !b.grad_lepe_b || !b.workspace)
        return fail(CSB200_ERR_INVALID, "cross_stripe_attn_bwd: null pointer");
      const void* ptrs[] = {b.q, b.k, b.v, b.out, b.grad_out, b.dq, b.dk, b.dv, b.workspace};
      for (const void* ptr : ptrs)
        if (!aligned(ptr, 16))
          return fail(CSB200_ERR_INVALID, "cross_stripe_attn_bwd: tensors must be 16-byte aligned");
      if (b.workspace_bytes < csb200_stripe_attn_bwd_workspace_bytes(&d[i]))
        return fail(CSB200_ERR_WORKSPACE, "cross_stripe_attn_bwd: workspace too small");
      float* delta = static_cast<float*>(b.workspace);
      float* partial = reinterpret_cast<float*>(
          static_cast<char*>(b.workspace) +
          align_up((size_t)g[i].B * g[i].heads * g[i].L * sizeof(float), 256));
      pio[i] = PrepIO{b.v, b.out, b.grad_out, b.lepe_w, b.lepe_b, delta, partial, b.grad_lepe_w,
                      b.grad_lepe_b};
      tio[i] = TcBwdIO{b.q, b.k, b.v, b.grad_out, b.lepe_w, b.lse, delta, b.dq, b.dk, b.dv, partial,
                       b.grad_lepe_w, b.grad_lepe_b};
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the per-CTA partials of the depthwise gradients are summed by the prologue of the backward kernel itself
    // (one launch less per call: the separate final-sum kernel was 6 us of launch latency for 10 KB of output)
    int wg_blocks = 0;
    int rc = lepe_bwd_prep_multi(2, g, d[0].dtype, pio, st, &wg_blocks);
    if (rc != CSB200_OK) return rc;
    return tc_bwd_multi(2, g, tio, st, wg_blocks);
  }
  for (int i = 0; i < n; ++i) {
    int rc = csb200_stripe_attn_bwd(&d[i], io[i].q, io[i].k, io[i].v, io[i].lepe_w, io[i].lepe_b,
                                    io[i].out, io[i].grad_out, io[i].lse, io[i].dq, io[i].dk, io[i].dv,
                                    io[i].grad_lepe_w, io[i].grad_lepe_b, io[i].workspace,
                                    io[i].workspace_bytes, stream);
    if (rc != CSB200_OK) return rc;
  }
  return CSB200_OK;
}

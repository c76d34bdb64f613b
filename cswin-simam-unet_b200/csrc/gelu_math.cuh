// Exact-erf GELU (nn.GELU(), C:190) on packed fp32 pairs — shared by the flat passes (gelu.cu) and the
// tcgen05 Linear epilogues (linear_tc.cu), so both paths produce the same values for the same input.
#pragma once

#include "common.cuh"

namespace csb200 {

constexpr float kSqrtHalf = 0.70710678118654752440f;
constexpr float kInvSqrt2Pi = 0.39894228040143267794f;

// bf16: erf(z) = sign(z) (1 - poly(t) e^{-z^2}), t = 1 / (1 + p |z|), with e^{-z^2} = e^{-x^2 / 2}, evaluated
// on packed fp32 pairs below.
// ---- bf16 on packed fp32 pairs (FFMA2): both kernels are bound by instruction issue + MUFU --------
struct GeluPair {
  f2_t erf, gauss;  // erf(x / sqrt 2), e^{-x^2 / 2}
};
__device__ __forceinline__ GeluPair erf_and_gauss2(f2_t x) {
  const f2_t ax = x & 0x7fffffff7fffffffull;                       // |x| on both halves
  const f2_t z = f2_mul(ax, f2_splat(kSqrtHalf));
  float e0, e1, d0, d1;
  f2_split(f2_mul(f2_mul(x, x), f2_splat(-0.72134752044448170368f)), e0, e1);
  asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(e0));
  asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(e1));
  f2_split(f2_fma(z, f2_splat(0.3275911f), f2_splat(1.f)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(d0));
  asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(d1));
  const f2_t t = f2_make(d0, d1), e = f2_make(e0, e1);
  // -poly(t): the negated Abramowitz-Stegun coefficients, so that erf|x| = 1 + (-poly) e
  f2_t p = f2_fma(f2_splat(-1.061405429f), t, f2_splat(1.453152027f));
  p = f2_fma(p, t, f2_splat(-1.421413741f));
  p = f2_fma(p, t, f2_splat(0.284496736f));
  p = f2_fma(p, t, f2_splat(-0.254829592f));
  p = f2_mul(p, t);
  const f2_t erf_abs = f2_fma(p, e, f2_splat(1.f));                 // in [0, 1]: sign bit clear
  return GeluPair{erf_abs | (x & 0x8000000080000000ull), e};       // copysign on both halves
}
__device__ __forceinline__ uint32_t gelu_fwd2(uint32_t w) {
  const f2_t x = f2_from_bf16x2(w);
  const GeluPair g = erf_and_gauss2(x);
  const f2_t hx = f2_mul(x, f2_splat(0.5f));
  float lo, hi;
  f2_split(f2_fma(hx, g.erf, hx), lo, hi);
  return pack_bf16x2(lo, hi);
}
// GELU(x) and GELU'(x) = Phi(x) + x phi(x) of a packed bf16 pair in one go: the derivative costs three more packed
// operations once erf and the Gaussian are there.  The fused Mlp (linear_tc.cu) stores GELU'(h) instead of h in
// forward, so that its backward epilogue is one multiplication per element instead of a second erf + exponential.
__device__ __forceinline__ uint32_t gelu_fwd_deriv2(uint32_t w, uint32_t& deriv) {
  const f2_t x = f2_from_bf16x2(w);
  const GeluPair g = erf_and_gauss2(x);
  const f2_t hx = f2_mul(x, f2_splat(0.5f));
  const f2_t cdf = f2_fma(g.erf, f2_splat(0.5f), f2_splat(0.5f));
  float lo, hi;
  f2_split(f2_fma(f2_mul(x, g.gauss), f2_splat(kInvSqrt2Pi), cdf), lo, hi);
  deriv = pack_bf16x2(lo, hi);
  f2_split(f2_fma(hx, g.erf, hx), lo, hi);
  return pack_bf16x2(lo, hi);
}
__device__ __forceinline__ uint32_t gelu_bwd2(uint32_t wg, uint32_t wh) {
  const f2_t g = f2_from_bf16x2(wg), x = f2_from_bf16x2(wh);
  const GeluPair r = erf_and_gauss2(x);
  const f2_t cdf = f2_fma(r.erf, f2_splat(0.5f), f2_splat(0.5f));
  const f2_t xpdf = f2_mul(x, f2_mul(r.gauss, f2_splat(kInvSqrt2Pi)));
  float lo, hi;
  f2_split(f2_mul(g, f2_add(cdf, xpdf)), lo, hi);
  return pack_bf16x2(lo, hi);
}
// grad_h = g * GELU'(h) for an fp32 pair g (e.g. tensor-core accumulators) and a packed bf16 pair h; fp32 pair
__device__ __forceinline__ f2_t gelu_bwd2_f32(f2_t g, uint32_t wh) {
  const f2_t x = f2_from_bf16x2(wh);
  const GeluPair r = erf_and_gauss2(x);
  const f2_t cdf = f2_fma(r.erf, f2_splat(0.5f), f2_splat(0.5f));
  const f2_t xpdf = f2_mul(x, f2_mul(r.gauss, f2_splat(kInvSqrt2Pi)));
  return f2_mul(g, f2_add(cdf, xpdf));
}
// (Measured, round 2: trading the reciprocal for a degree-9 erfcx polynomial — one MUFU op per element instead
// of two, three more packed FMAs — made the GEMM epilogues 5-15 % SLOWER: they are bound by instruction issue,
// not by the MUFU pipe.  profiles/r2_linear_bench_erfcx_experiment.jsonl)

}  // namespace csb200

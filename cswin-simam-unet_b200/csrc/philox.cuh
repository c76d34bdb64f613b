// Philox4x32-10 (Salmon et al., SC'11) — the counter-based generator behind the attention-dropout mask.
#pragma once
#include "stripe_attn.cuh"

namespace csb200 {

__device__ __forceinline__ uint4 philox4x32_10(uint2 key, uint4 c) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ key.x, lo1, hi0 ^ c.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return c;
}

// Per-call state of the dropout generator, read once per thread from the device-resident rng_state.
struct DropRng {
  uint2 key;       // seed
  uint32_t call;   // call counter (low word; the high word is folded into the key)
};
__device__ __forceinline__ DropRng drop_rng_load(const unsigned long long* rng) {
  const unsigned long long seed = rng[0], call = rng[1];
  DropRng r;
  r.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(call >> 32));
  r.call = (uint32_t)call;
  return r;
}
// stripe-and-head unit of a (image, stripe row, stripe column, head) group of a branch
__device__ __forceinline__ uint32_t drop_unit(const StripeGeom& g, int b, int wy, int wx, int head) {
  return ((uint32_t)(((b * g.nwy + wy) * g.nwx + wx) * g.heads + head) << 1) | (g.drop_salt & 1u);
}
// 16 random bytes: the keep decisions of query `i` against keys 16 jb .. 16 jb + 15 of its stripe
__device__ __forceinline__ uint4 drop_bytes(const DropRng& r, uint32_t unit, uint32_t i, uint32_t jb) {
  return philox4x32_10(r.key, make_uint4(jb, i, unit, r.call));
}
__device__ __forceinline__ uint32_t drop_byte(const uint4& rb, int j) {  // byte (j & 15)
  const uint32_t w = (j & 8) ? ((j & 4) ? rb.w : rb.z) : ((j & 4) ? rb.y : rb.x);
  return (w >> ((j & 3) * 8)) & 0xffu;
}

}  // namespace csb200

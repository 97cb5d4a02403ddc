// Counter-based noise for the DDPM sampler: Philox4x32-10 (Salmon et al., SC'11) -> Box-Muller normals.
// Oracle: oracle/philox.py (same counter layout; the Random123 known-answer vectors pin both).
//   ctr = (column / 4, latent index, t, kPhiloxTag)    key = (seed lo, seed hi)
// One call yields the normals of 4 consecutive columns of one latent at one step; x_T uses t = steps.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sdfb {

constexpr uint32_t kPhiloxTag = 0x53444642u;

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// 24-bit uniforms: u1 in (0, 1], u2 in [0, 1); every intermediate but log / sqrt / sincos is exact in fp32
__device__ __forceinline__ void box_muller(uint32_t ra, uint32_t rb, float& za, float& zb) {
  const float u1 = (static_cast<float>(ra >> 8) + 1.0f) * 5.9604644775390625e-8f;
  const float u2 = static_cast<float>(rb >> 8) * 5.9604644775390625e-8f;
  const float rad = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  za = rad * c;
  zb = rad * s;
}

// normals of columns [4 g, 4 g + 4) of latent `row` at step `t`
__device__ __forceinline__ float4 philox_normal4(uint32_t g, uint32_t row, uint32_t t, uint2 key) {
  const uint4 r = philox4x32_10(make_uint4(g, row, t, kPhiloxTag), key);
  float4 z;
  box_muller(r.x, r.y, z.x, z.y);
  box_muller(r.z, r.w, z.z, z.w);
  return z;
}

}  // namespace sdfb

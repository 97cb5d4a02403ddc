// Small tensor-core support kernels shared by the decoder path:
//  * umma_selftest_kernel - one 128x256x64 tcgen05 product through the same shared-memory
//    descriptors, 128B swizzle, commit and TMEM-load conventions the fused decoder uses
//    (the unit test of that plumbing);
//  * fold_latent_kernel   - the per-latent constants bias0' and bias4'.
// No upstream source exists (/root/reference/README.md:1); see SURVEY.md section 8a.
#include "kernels.h"
#include "ptx.cuh"

namespace sdfb {

namespace {

// ---------------------------------------------------------------------------
// UMMA self-test: one 128x256x64 product through exactly the descriptors, swizzle, TMEM
// load and commit paths the fused kernel uses.  A and B arrive row-major; the kernel
// swizzles them into shared memory itself.
template <bool FP16>
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const uint16_t* __restrict__ a,
                                                               const uint16_t* __restrict__ b,
                                                               float* __restrict__ d, unsigned int* status) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (smem0 - smem_u32(smem_raw));
  const uint32_t sa = smem0, sb = smem0 + kAChunkBytes, bar = smem0 + kAChunkBytes + kBlockBytes;
  volatile uint32_t* misc = reinterpret_cast<volatile uint32_t*>(gen + kAChunkBytes + kBlockBytes + 8);
  const int warp = threadIdx.x >> 5;
  // stage operands: 16-byte units, swizzled
  for (int i = threadIdx.x; i < 128 * 8; i += 128) {
    const int r = i >> 3, u = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(a + r * 64 + u * 8);
    *reinterpret_cast<uint4*>(gen + r * 128 + ((u ^ (r & 7)) << 4)) = v;
  }
  for (int i = threadIdx.x; i < 256 * 8; i += 128) {
    const int r = i >> 3, u = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(b + r * 64 + u * 8);
    *reinterpret_cast<uint4*>(gen + kAChunkBytes + r * 128 + ((u ^ (r & 7)) << 4)) = v;
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    misc[1] = 0;
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc<1>(smem_u32(const_cast<uint32_t*>(misc)), 256);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = misc[0];
  Watchdog wd{misc + 1, status, 200000000ull, nullptr, nullptr};
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc(128, 256, FP16 ? 0 : 1);
    const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sb);
#pragma unroll
    for (int j = 0; j < 4; ++j) umma_ss<1>(tmem_base, adesc + 2 * j, bdesc + 2 * j, idesc, j != 0 ? 1u : 0u);
    umma_commit<1>(bar);
  }
  if (mbar_wait(bar, 0, wd, 0x70)) {
    tc_fence_after();
    const int row = threadIdx.x;
    for (int g = 0; g < 8; ++g) {
      uint32_t v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + g * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) d[row * 256 + g * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 256);
  }
}

// bias0' and bias4': the latent's contribution to layers 0 and 4 (one warp per output feature;
// lanes stride k, shuffle tree) - fp32, order fixed, independent of everything else.
__global__ void fold_latent_kernel(const float* __restrict__ W0, const float* __restrict__ b0,
                                   const float* __restrict__ W4, const float* __restrict__ b4,
                                   const float* __restrict__ z, DecConsts* __restrict__ consts) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // 0..1023
  const int lane = threadIdx.x & 31;
  if (j >= 2 * kHid) return;
  const bool l4 = j >= kHid;
  const int n = l4 ? j - kHid : j;
  const float* w = l4 ? W4 + static_cast<long long>(n) * kHid + kSkipOut
                      : W0 + static_cast<long long>(n) * kDecIn;
  float s = 0.f;
  for (int k = lane; k < kLatent; k += 32) s = fmaf(w[k], z[k], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    if (l4) consts->bias[3][n] = b4[n] + s;
    else consts->l0[n].w = b0[n] + s;
  }
}

}  // namespace

cudaError_t tc_common_init() {
  constexpr int st_bytes = kAChunkBytes + kBlockBytes + 64 + 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, st_bytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(umma_selftest_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, st_bytes);
}

cudaError_t launch_umma_selftest(const uint16_t* a, const uint16_t* b, float* d, unsigned int* status,
                                 bool fp16, cudaStream_t stream) {
  constexpr int st_bytes = kAChunkBytes + kBlockBytes + 64 + 1024;
  if (fp16)
    umma_selftest_kernel<true><<<1, 128, st_bytes, stream>>>(a, b, d, status);
  else
    umma_selftest_kernel<false><<<1, 128, st_bytes, stream>>>(a, b, d, status);
  return cudaGetLastError();
}

cudaError_t launch_fold_latent(const float* W0, const float* b0, const float* W4, const float* b4,
                               const float* z, DecConsts* consts, cudaStream_t stream) {
  fold_latent_kernel<<<(2 * kHid * 32 + 255) / 256, 256, 0, stream>>>(W0, b0, W4, b4, z, consts);
  return cudaGetLastError();
}

}  // namespace sdfb

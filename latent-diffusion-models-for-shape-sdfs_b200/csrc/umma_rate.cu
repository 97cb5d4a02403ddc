// Diagnostic micro-benchmark: how many cycles does one tcgen05.mma (kind::f16, N=256, K=16)
// take when issued back to back from shared-memory operands?  cta_group::1 (M=128) and
// cta_group::2 (M=256 per pair).  Operand contents are irrelevant (zero-filled shared memory).
#include "kernels.h"
#include "ptx.cuh"

namespace sdfb {
namespace {

template <int CG>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int iters, int k_per_commit, int n_acc, int flags, long long* out) {
  // flags bit0: no intermediate waits (commits go to a second, never-waited barrier); bit1: never
  // restart the accumulation; bit2: two commits per commit point;
  // bit3: TS form - the A operand comes from TMEM (columns 256.., packed 16-bit), N = 128 accumulators (what an
  //       activations-in-TMEM version of the fused decoder would issue);
  // bit4 / bit5: no MMAs at all - the four warps run an epilogue-shaped loop `iters` times per chunk of 32 columns
  //       (TMEM load, bias add, ReLU, round, hand the operand over) with the hand-over through tcgen05.st + wait::st (bit4)
  //       or through st.shared + fence.proxy.async (bit5); out = cycles per chunk seen by warp 0
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (smem0 - smem_u32(smem_raw));
  // A: 128 rows x 64 k (16 KiB) x 4 chunks ; B: (256/CG) rows x 64 k x 2 stages
  const uint32_t sa = smem0, sb = smem0 + 4 * kAChunkBytes, bar = sb + 2 * kBlockBytes;
  volatile uint32_t* misc = reinterpret_cast<volatile uint32_t*>(gen + 4 * kAChunkBytes + 2 * kBlockBytes + 32);
  for (int i = threadIdx.x; i < (4 * kAChunkBytes + 2 * kBlockBytes) / 16; i += 128)
    reinterpret_cast<uint4*>(gen)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5;
  uint32_t rank = 0;
  if constexpr (CG == 2) rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    misc[1] = 0;
    mbar_init(bar, 1);
    mbar_init(bar + 8, 1);
    mbar_init(bar + 16, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc<CG>(smem_u32(const_cast<uint32_t*>(misc)), 512);
    tmem_relinquish<CG>();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = misc[0];
  if (flags & 8) {
    // the TMEM operand region is written once (zeros) so the MMAs never read tensor memory nobody has stored to
    uint32_t z[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) z[j] = 0u;
    for (int c = 0; c < 32; ++c) tmem_st16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c * 16, z);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();
    tc_fence_after();
  }
  if (flags & 48) {
    // ---- epilogue-shaped loop (see the flags above) ----
    const int lane = threadIdx.x & 31;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t srow = sa + (warp * 32 + lane) * 128;
    const uint32_t row7 = (warp * 32 + lane) & 7u;
    float bias = 0.25f * lane;
    __syncthreads();
    const long long t0 = clock64();
    uint32_t v[32];
    tmem_ld32(trow, v);
    for (int it = 0; it < iters; ++it) {
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_relu<false>(__uint_as_float(v[2 * j]) + bias, __uint_as_float(v[2 * j + 1]) + bias);
      tmem_ld32(trow + ((it + 1) & 7) * 32, v);            // next chunk's accumulator in flight
      if (flags & 16) {
        tmem_st16(trow + 256 + (it & 7) * 16, pk);
        tmem_st_wait();
        tc_fence_before();
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          st_shared_v4(srow + ((it & 3) * kAChunkBytes) + ((static_cast<uint32_t>(u) ^ row7) << 4), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar + 16);
    }
    tmem_ld_wait();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x / CG] = t1 - t0;    // the host divides by iters * k_per_commit * 4
  } else if (threadIdx.x == 0 && rank == 0) {
    // bit6: the TS form with N = 256 (one accumulator); bit7: operand in columns [0, 256), accumulators above it
    const bool ts = (flags & 8) != 0, ts_wide = (flags & 64) != 0;
    const bool ss_narrow = (flags & 256) != 0;            // bit8: the SS form with N = 128
    const uint32_t idesc = ((ts && !ts_wide) || ss_narrow) ? umma_idesc(128 * CG, 128, 1) : umma_idesc(128 * CG, 256, 1);
    const uint32_t a_t = tmem_base + ((flags & 128) ? 0u : 256u), d_t = tmem_base + ((flags & 128) ? 256u : 0u);
    uint32_t parity = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < k_per_commit; ++k) {
        const uint64_t adesc = umma_desc_sw128(sa + (k & 3) * kAChunkBytes);
        const uint64_t bdesc = umma_desc_sw128(sb + (k & 1) * kBlockBytes);
        const uint32_t d = ts ? d_t + (ts_wide ? 0 : (it % n_acc) * 128) : tmem_base + (it % n_acc) * 256;
        if (ts) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            umma_ts<CG>(d, a_t + ((k & 7) * 32) + 8 * j, bdesc + 2 * j, idesc, ((k | j) != 0 || ((flags & 2) && it > 0)) ? 1u : 0u);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) umma_ss<CG>(d, adesc + 2 * j, bdesc + 2 * j, idesc, ((k | j) != 0 || ((flags & 2) && it > 0)) ? 1u : 0u);
        }
      }
      if (flags & 1) {
        umma_commit<CG>(bar + 8);                       // nobody waits on this one
        if (flags & 4) umma_commit<CG>(bar + 16);
        if (it == iters - 1) umma_commit<CG>(bar);
      } else {
        umma_commit<CG>(bar);
        if (flags & 4) umma_commit<CG>(bar + 16);
        // keep at most one commit group outstanding behind the one being issued
        if (it > 0) {
          while (!mbar_try_wait(bar, parity)) {}
          parity ^= 1u;
        }
      }
    }
    while (!mbar_try_wait(bar, parity)) {}
    const long long t1 = clock64();
    out[blockIdx.x / CG] = t1 - t0;
  } else if (threadIdx.x == 0) {
    // peer CTA: consume the multicast commits so its barrier phases stay in step
    uint32_t parity = 0;
    const int n = (flags & 1) ? 1 : iters;
    for (int it = 0; it < n; ++it) {
      while (!mbar_try_wait(bar, parity)) {}
      parity ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, 512);
  }
}

}  // namespace

// cycles for `iters` commit groups of 4*k_per_commit MMAs each, per CTA (pair); returns via out_dev[grid/cg]
cudaError_t launch_umma_rate(int cg, int grid, int iters, int k_per_commit, int n_acc, int flags, long long* out_dev,
                             cudaStream_t stream) {
  constexpr int bytes = 4 * kAChunkBytes + 2 * kBlockBytes + 64 + 1024;
  cudaError_t e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cg;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cg == 1) {
    e = cudaFuncSetAttribute(umma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    return cudaLaunchKernelEx(&cfg, umma_rate_kernel<1>, iters, k_per_commit, n_acc, flags, out_dev);
  }
  e = cudaFuncSetAttribute(umma_rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  return cudaLaunchKernelEx(&cfg, umma_rate_kernel<2>, iters, k_per_commit, n_acc, flags, out_dev);
}

}  // namespace sdfb

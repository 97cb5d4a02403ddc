// Diagnostic micro-benchmark: how many bytes per SM cycle can a CTA pull out of L2 through the TMA unit into shared memory?
// This is the operand-delivery ceiling the sampler kernel (ddpm_step.cu) and the decoder's weight ring run against.
//   mode 0: every CTA loads its own sequence of 16 KiB boxes (128 rows x 128 bytes, 128B swizzle) from an L2-resident tensor
//   mode 1: all CTAs of a cluster load the SAME sequence (what the four pairs of a latent group do with the A operand)
//   mode 2: the same sequence, but every box is loaded once and multicast to all CTAs of the cluster (the CTAs take turns)
//   mode 3: own 16 KiB pieces as ONE 1-D bulk copy each (cp.async.bulk, contiguous source); mode 4: as four 4 KiB bulk copies;
//   mode 5: CTA pairs (cluster 2) the way the tcgen05 kernels load: both CTAs load their own boxes into their own shared
//           memory with the .cta_group::2 form, completion counted on the LEADER CTA's barrier (expects 2 x 16 KiB per stage)
//   mode 6: CTA pairs again, but every CTA counts its own boxes on its OWN barrier (plain loads) and the peer forwards each
//           completed stage to the leader with one remote arrive; the leader releases a stage in both CTAs as in mode 5
// `issuers` warps (1, 2 or 4) take the boxes in turn.  `uniform` = 1: the whole issuing warp runs the loop and one elected lane
// executes the TMA instructions (what the compiler needs to keep UTMALDG out of an ELECT / BRA.U.ANY loop); 0: `if (lane == 0)`
// around the loop, the way a producer warp is usually written.
// Nothing reads the data: one thread arms and issues, one thread waits and releases.
#include <cuda.h>

#include "kernels.h"
#include "ptx.cuh"

namespace sdfb {
namespace {

constexpr int kIngestStages = 8;
constexpr uint32_t kBox = 16384;

__device__ __forceinline__ void tma_box(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_box_pair(uint32_t dst, const void* tmap, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void tma_box_multicast(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar), "h"(mask)
      : "memory");
}

__global__ void __launch_bounds__(192, 1)
tma_ingest_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* base, int iters, int col_blocks, int row_blocks, int mode, int csize, int issuers,
                  int uniform, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem0 + kIngestStages * kBox;       // full[8] | consumed[8] | empty[8]
  const uint32_t crank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kIngestStages; ++s) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 8 * (kIngestStages + s), 1);
      mbar_init(bars + 8 * (2 * kIngestStages + s), mode == 6 ? 1u : static_cast<uint32_t>(csize));
    }
    fence_mbar_init();
  }
  __syncthreads();
  cluster_sync_all();
  const int seq = (mode == 1 || mode == 2) ? static_cast<int>(blockIdx.x) / csize : static_cast<int>(blockIdx.x);
  const int nboxes = col_blocks * row_blocks;
  long long t0 = 0, t1 = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= 2 && warp < 2 + issuers && (uniform || lane == 0)) {
    // (no divisions in the loop: a lone thread runs dependent integer code at ~5 cycles per instruction)
    const int step = issuers, first = warp - 2;
    int box = static_cast<int>((static_cast<long long>(seq) * iters + first) % nboxes);
    int cb = box % col_blocks, rb = box / col_blocks;
    for (int i = first; i < iters; i += step) {
      const int s = i & (kIngestStages - 1), use = i / kIngestStages;
      if (use > 0) while (!mbar_try_wait(bars + 8 * (kIngestStages + s), (use - 1) & 1)) {}     // this CTA is done with the slot
      const bool mine = mode != 2 || static_cast<uint32_t>(i & (csize - 1)) == crank;
      if (mode == 2 && mine && use > 0) while (!mbar_try_wait_cluster(bars + 8 * (2 * kIngestStages + s), (use - 1) & 1)) {}   // ... and so is every peer
      const int c0 = cb * 64, c1 = rb * 128;
      if (!uniform || elect_one()) {
        if (mode == 5) {
          if (crank == 0) mbar_arrive_expect_tx(bars + 8 * s, 2 * kBox);
          tma_box_pair(smem0 + s * kBox, &tm, c0, c1, map_to_cta(bars + 8 * s, 0));
        } else {
        mbar_arrive_expect_tx(bars + 8 * s, kBox);
        if (mode == 3) {
          bulk_g2s(smem0 + s * kBox, base + static_cast<size_t>(box) * kBox, kBox, bars + 8 * s);
        } else if (mode == 4) {
#pragma unroll
          for (int q = 0; q < 4; ++q) bulk_g2s(smem0 + s * kBox + q * 4096, base + static_cast<size_t>(box) * kBox + q * 4096, 4096, bars + 8 * s);
        } else if (mode != 2) {
          tma_box(smem0 + s * kBox, &tm, c0, c1, bars + 8 * s);
        } else if (mine) {
          tma_box_multicast(smem0 + s * kBox, &tm, c0, c1, bars + 8 * s, static_cast<uint16_t>((1u << csize) - 1u));
        }
        }
      }
      if (uniform) __syncwarp();
      for (int q = 0; q < step; ++q) {
        if (++box == nboxes) { box = 0; cb = 0; rb = 0; }
        else if (++cb == col_blocks) { cb = 0; ++rb; }
      }
    }
  } else if (threadIdx.x == 32) {
    t0 = clock64();
    for (int i = 0; i < iters && !(mode == 5 && crank != 0); ++i) {
      const int s = i & (kIngestStages - 1), use = i / kIngestStages;
      while (!mbar_try_wait(bars + 8 * s, use & 1)) {}
      if (mode == 6) {
        if (crank != 0) { mbar_arrive_cluster(map_to_cta(bars + 8 * (2 * kIngestStages + s), 0)); continue; }
        while (!mbar_try_wait_cluster(bars + 8 * (2 * kIngestStages + s), use & 1)) {}
        mbar_arrive_cluster(map_to_cta(bars + 8 * (kIngestStages + s), 1));
      }
      mbar_arrive(bars + 8 * (kIngestStages + s));
      if (mode == 5) mbar_arrive_remote_relaxed(bars + 8 * (kIngestStages + s), 1, 1);      // (relaxed: nothing is read, and a release fence per box would pace the loop)
      if (mode == 2) mbar_arrive_cluster(map_to_cta(bars + 8 * (2 * kIngestStages + s), static_cast<uint32_t>(i & (csize - 1))));
    }
    t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  cluster_sync_all();
}

}  // namespace

// out_dev[grid] = cycles each CTA needed for `iters` boxes of 16 KiB; the tensor is [rows][cols] 16-bit, L2-resident if it fits
cudaError_t launch_tma_ingest(const void* tensor, int cols, int rows, int grid, int csize, int mode, int issuers, int uniform, int iters,
                              long long* out_dev, cudaStream_t stream) {
  alignas(64) CUtensorMap tm;
  const unsigned long long dims[2] = {static_cast<unsigned long long>(cols), static_cast<unsigned long long>(rows)};
  const unsigned long long strides[1] = {static_cast<unsigned long long>(cols) * 2ull};
  const unsigned box[2] = {64u, 128u};
  cudaError_t e = make_tensor_map(&tm, tensor, 2, 2, dims, strides, box, true);
  if (e != cudaSuccess) return e;
  constexpr int bytes = kIngestStages * kBox + 3 * kIngestStages * 8 + 1024;
  e = cudaFuncSetAttribute(tma_ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  if (csize > 8) {
    e = cudaFuncSetAttribute(tma_ingest_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, tma_ingest_kernel, tm, static_cast<const uint8_t*>(tensor), iters, cols / 64, rows / 128, mode, csize, issuers, uniform, out_dev);
}

}  // namespace sdfb

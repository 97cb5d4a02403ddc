// General 16-bit tensor-core product for the TRAINING steps (decoder weight gradients, DDPM
// training step; oracle: oracle/train.py; no upstream source exists, /root/reference/README.md:1).  The inference kernels
// (fused_decoder.cu, ddpm_step.cu) keep their activations on chip; training has to keep every layer's activations and
// deltas for the weight gradients, so it runs layer by layer through this kernel:
//
//   NT:  C[M][N] = A[M][K] . B[N][K]^T          forward (A = activations, B = W) and backward-data (A = delta, B = W^T copy)
//   TN:  C[M][N] = At[K][M]^T . Bt[K][N]         weight gradient: C = dW[out][in], At = delta[rows][out], Bt = h[rows][in];
//                                                the contraction runs over the ROWS (latents / queries) of two row-major
//                                                arrays, i.e. both operands are MN-major for the tensor core
// 16-bit operands (bf16 / fp16) from global memory through TMA boxes with 128-byte swizzle, fp32 accumulation in TMEM.
// One CTA computes 128 x BN output tiles (cta_group::1, M = 128, N = BN <= 256), K in steps of 64 through a 4-stage ring;
// two accumulators ping-pong so the epilogue of a tile overlaps the products of the next.  Persistent over
// (tile, k-split) work items.
//
//   warps 0-7  epilogue (TMEM lane quadrant = warp & 3, column half = warp >> 2): TMEM -> registers -> global, per 32-column group
//   warp 8     TMA producer 0 (arms the stage, A)      warp 9   MMA issuer + TMEM allocation
//   warps 10, 11  TMA producers 1 and 2 (a half of B each)
//
// Shared-memory operand layouts (what the TMA boxes produce and the descriptors describe):
//   K-major  (NT): rows = M or N index, 128 B per row = 64 k, 16-byte units XOR (row & 7)            (SBO = 1024 B)
//   MN-major (TN): panels of [64 k-rows][64 m or n values = 128 B], units XOR (k-row & 7); panels of one operand are
//                  8 KiB apart (LBO), 8-row groups inside a panel 1 KiB apart (SBO); one MMA (K = 16) reads two groups.
#include <cuda.h>

#include "kernels.h"
#include "ptx.cuh"

namespace sdfb {

namespace {

constexpr int kGStages = 4;
constexpr int kGBM = 128, kGBK = 64;
constexpr int kGEpiWarps = 8;
constexpr int kGThreads = (kGEpiWarps + 4) * 32;    // + producer 0, MMA issuer, producers 1 and 2
constexpr uint32_t kGAStage = kGBM * kGBK * 2;                  // 16 KiB
constexpr uint32_t kGBStageMax = 256 * kGBK * 2;                // 32 KiB
constexpr uint32_t kGStage = kGAStage + kGBStageMax;            // 48 KiB
constexpr int kGBarFull = 0, kGBarEmpty = kGStages, kGBarAccFull = 2 * kGStages, kGBarAccEmpty = 2 * kGStages + 2;
constexpr int kGNumBars = 2 * kGStages + 4;
constexpr uint32_t kGSmem = kGStages * kGStage + kGNumBars * 8 + 16 + 1024;

enum : uint32_t { kGErrFull = 0x210, kGErrEmpty = 0x220, kGErrAccFull = 0x230, kGErrAccEmpty = 0x240 };

// shared-memory descriptor, MN-major operand, 128-byte swizzle (see the header comment)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;   // next 64-wide panel along M / N
  d |= static_cast<uint64_t>(1024 >> 4) << 32;        // next group of 8 k-rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void tma_load_box(uint32_t dst_smem, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

template <bool FP16>
__device__ __forceinline__ float lowp_bits_to_float(uint32_t bits16) {
  if constexpr (FP16) {
    float f;
    asm("{.reg .b16 h; cvt.u16.u32 h, %1; cvt.f32.f16 %0, h;}" : "=f"(f) : "r"(bits16 & 0xFFFFu));
    return f;
  } else {
    return __uint_as_float(bits16 << 16);
  }
}

template <bool FP16, bool TN>
__global__ void __launch_bounds__(kGThreads, 1)
gemm_tc_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (smem0 - smem_u32(smem_raw));
  const uint32_t bars = smem0 + kGStages * kGStage;
  volatile uint32_t* misc = reinterpret_cast<volatile uint32_t*>(gen + kGStages * kGStage + kGNumBars * 8);   // [0] tmem base, [1] abort
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.bn;
  const int tiles_m = (p.M + kGBM - 1) / kGBM, tiles_n = p.N / BN;
  const int ksteps_total = (p.K + kGBK - 1) / kGBK;
  const int ksteps_per = (ksteps_total + p.ksplit - 1) / p.ksplit;
  const long long items = static_cast<long long>(tiles_m) * tiles_n * p.ksplit;

  if (threadIdx.x == 0) {
    misc[1] = 0;
    for (int s = 0; s < kGStages; ++s) {
      mbar_init(bars + 8 * (kGBarFull + s), 1);
      mbar_init(bars + 8 * (kGBarEmpty + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bars + 8 * (kGBarAccFull + b), 1);
      mbar_init(bars + 8 * (kGBarAccEmpty + b), kGEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == kGEpiWarps + 1) {
    tmem_alloc<1>(smem_u32(const_cast<uint32_t*>(misc)), 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = misc[0];
  Watchdog wd{misc + 1, p.status, p.timeout_ns, nullptr, p.status_host};

  if (warp == kGEpiWarps || warp >= kGEpiWarps + 2) {
    // ===================== producers =====================
    // THREE issuing warps.  The TMA unit works through one thread's boxes one after the other at ~32 bytes per cycle
    // (tools/tma_ingest.py, profiles/r2_tma_ingest.txt: 32 / 55 / 98 B/cycle/SM with 1 / 2 / 4 issuing warps), and a k-step of
    // this kernel needs 16 KiB of A + up to 32 KiB of B per 512 tensor cycles = 96 B/cycle: with one producer the product ran
    // at a third of the tensor rate.  Producer 0 arms the stage and brings A, producers 1 and 2 a half of B each; all three
    // visit every stage in order, so each waits on the stage's release itself.
    const int pidx = warp == kGEpiWarps ? 0 : warp - kGEpiWarps - 1;
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t tx = kGAStage + static_cast<uint32_t>(BN) * kGBK * 2;
      // B in units of 64 rows (NT: a box of up to 128 rows = 2 units; TN: one panel = 1 unit); units [u0, u1) are this warp's
      const int units = BN / 64, half = (units + 1) / 2;
      const int u0 = pidx == 1 ? 0 : half, u1 = pidx == 1 ? half : units;
      for (long long w = blockIdx.x; w < items; w += gridDim.x) {
        const int split = static_cast<int>(w % p.ksplit);
        const long long t = w / p.ksplit;
        const int tn = static_cast<int>(t % tiles_n), tm = static_cast<int>(t / tiles_n);
        const int k0 = split * ksteps_per, k1 = min(k0 + ksteps_per, ksteps_total);
        for (int ks = k0; ks < k1; ++ks) {
          if (!mbar_wait(bars + 8 * (kGBarEmpty + stage), phase ^ 1u, wd, kGErrEmpty, stage)) goto done;
          const uint32_t full = bars + 8 * (kGBarFull + stage);
          const uint32_t sa = smem0 + stage * kGStage, sb = sa + kGAStage;
          if (pidx == 0) {
            mbar_arrive_expect_tx(full, tx);
            if constexpr (!TN) {
              tma_load_box(sa, &tm_a, ks * kGBK, tm * kGBM, full);                     // [128 rows][64 k]
            } else {
              for (int pnl = 0; pnl < kGBM / 64; ++pnl)                                // panels of [64 k-rows][64 m]
                tma_load_box(sa + pnl * 8192, &tm_a, tm * kGBM + pnl * 64, ks * kGBK, full);
            }
          } else if constexpr (!TN) {
            // boxes of 128 rows (one of 64 when the tile is 64 wide): producer 1 takes the first half of them, producer 2 the rest
            const int nboxes = BN >= 128 ? BN / 128 : 1, bh = (nboxes + 1) / 2;
            for (int bx = (pidx == 1 ? 0 : bh); bx < (pidx == 1 ? bh : nboxes); ++bx)
              tma_load_box(sb + bx * 128 * 128, &tm_b, ks * kGBK, tn * BN + bx * 128, full);
          } else {
            for (int u = u0; u < u1; ++u) tma_load_box(sb + u * 8192, &tm_b, tn * BN + u * 64, ks * kGBK, full);
          }
          if (++stage == kGStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == kGEpiWarps + 1) {
    // ===================== MMA issuer (whole warp runs the loop, tcgen05 under elect) =====================
    const uint32_t idesc = umma_idesc(kGBM, BN, FP16 ? 0 : 1) | (TN ? ((1u << 15) | (1u << 16)) : 0u);
    uint32_t stage = 0, phase = 0, ephase = 0, item = 0;
    for (long long w = blockIdx.x; w < items; w += gridDim.x, ++item) {
      const int split = static_cast<int>(w % p.ksplit);
      const int k0 = split * ksteps_per, k1 = min(k0 + ksteps_per, ksteps_total);
      const uint32_t b = item & 1u;
      const uint32_t d_tmem = tmem_base + b * 256;
      if (!mbar_wait(bars + 8 * (kGBarAccEmpty + b), ((ephase >> b) & 1u) ^ 1u, wd, kGErrAccEmpty, b)) goto done;
      ephase ^= 1u << b;
      for (int ks = k0; ks < k1; ++ks) {
        if (!mbar_wait(bars + 8 * (kGBarFull + stage), phase, wd, kGErrFull, stage)) goto done;
        tc_fence_after();
        const uint32_t sa = smem0 + stage * kGStage, sb = sa + kGAStage;
        if (elect_one()) {
          if constexpr (!TN) {
            const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sb);
#pragma unroll
            for (int j = 0; j < 4; ++j) umma_ss<1>(d_tmem, adesc + 2 * j, bdesc + 2 * j, idesc, (ks > k0 || j > 0) ? 1u : 0u);
          } else {
            const uint64_t adesc = umma_desc_mn_sw128(sa, 8192), bdesc = umma_desc_mn_sw128(sb, 8192);
#pragma unroll
            for (int j = 0; j < 4; ++j) umma_ss<1>(d_tmem, adesc + 128 * j, bdesc + 128 * j, idesc, (ks > k0 || j > 0) ? 1u : 0u);
          }
          umma_commit<1>(bars + 8 * (kGBarEmpty + stage));
          if (ks == k1 - 1) umma_commit<1>(bars + 8 * (kGBarAccFull + b));
        }
        __syncwarp();
        if (++stage == kGStages) { stage = 0; phase ^= 1u; }
      }
      if (k1 <= k0 && elect_one()) umma_commit<1>(bars + 8 * (kGBarAccFull + b));      // empty k-range (cannot happen for ksplit <= ksteps)
    }
  } else {
    // ===================== epilogue warps =====================
    uint32_t acc_phase = 0, item = 0;
    const int quad = warp & 3, chalf = warp >> 2;                  // TMEM lane quadrant; which half of the tile's columns
    // 256-bit global accesses need 32-byte alignment: bases and row pitches of everything this epilogue touches
    const bool wide = (p.out_lowp == nullptr || ((reinterpret_cast<uintptr_t>(p.out_lowp) & 31u) == 0 && (p.ldo_lowp & 15) == 0)) &&
                      (p.out_f32 == nullptr || ((reinterpret_cast<uintptr_t>(p.out_f32) & 31u) == 0 && (p.ldo_f32 & 7) == 0 && (p.split_stride & 7) == 0)) &&
                      (p.mask_h == nullptr || ((reinterpret_cast<uintptr_t>(p.mask_h) & 31u) == 0 && (p.ldh & 15) == 0));
    const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    for (long long w = blockIdx.x; w < items; w += gridDim.x, ++item) {
      const int split = static_cast<int>(w % p.ksplit);
      const long long t = w / p.ksplit;
      const int tn = static_cast<int>(t % tiles_n), tm = static_cast<int>(t / tiles_n);
      const uint32_t b = item & 1u;
      const long long row = static_cast<long long>(tm) * kGBM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const int cbeg = BN >= 64 ? chalf * (BN / 2) : (chalf == 0 ? 0 : BN), cend = BN >= 64 ? cbeg + BN / 2 : BN;
      // ReLU mask of this thread's row (backward-data): the forward activations exist long before the accumulator is full,
      // so all their loads go out here and their latency hides behind the wait; one bit per column is kept
      uint32_t mbits[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
      if (p.epi == kGemmEpiMaskLowp && row_ok) {
        uint32_t hw[4][16];
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = cbeg + 32 * ci;
          if (c < cend) {
            const uint16_t* hp = p.mask_h + row * p.ldh + tn * BN + c;
            if (wide) {
              ld_global_nc_v8(hp, hw[ci]);
              ld_global_nc_v8(hp + 16, hw[ci] + 8);
            } else {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(hp) + u);
                hw[ci][4 * u] = q.x; hw[ci][4 * u + 1] = q.y; hw[ci][4 * u + 2] = q.z; hw[ci][4 * u + 3] = q.w;
              }
            }
          }
        }
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          if (cbeg + 32 * ci < cend) {
            uint32_t mb = 0;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              if (lowp_bits_to_float<FP16>(hw[ci][e] & 0xFFFFu) > 0.f) mb |= 1u << (2 * e);
              if (lowp_bits_to_float<FP16>(hw[ci][e] >> 16) > 0.f) mb |= 1u << (2 * e + 1);
            }
            mbits[ci] = mb;
          }
        }
      }
      if (!mbar_wait(bars + 8 * (kGBarAccFull + b), (acc_phase >> b) & 1u, wd, kGErrAccFull, b)) goto done;
      acc_phase ^= 1u << b;
      __syncwarp();
      tc_fence_after();
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int c = cbeg + 32 * ci;
        if (c >= cend) break;
        uint32_t v[32];
        tmem_ld32(tmem_row + b * 256 + c, v);
        tmem_ld_wait();
        const int col = tn * BN + c;
        float f[32];
        const bool empty_range = split * ksteps_per >= ksteps_total;      // a k-range past the end contributes zero (not stale TMEM)
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = empty_range ? 0.f : __uint_as_float(v[i]) * p.alpha;
        if (p.bias != nullptr) {     // scalar loads: a bias inside a parameter blob need not be 16-byte aligned
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] += __ldg(p.bias + col + i);
        }
        if (p.epi == kGemmEpiMaskLowp) {     // delta_in = (delta_out W) where the forward activation was positive
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (!((mbits[ci] >> i) & 1u)) f[i] = 0.f;
        }
        if (row_ok) {
          if (p.out_f32 != nullptr) {
            float* dst = p.out_f32 + static_cast<long long>(split) * p.split_stride + row * p.ldo_f32 + col;
            const bool post = p.epi == kGemmEpiBiasReluLowp && p.f32_post_relu;
            if (wide) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint32_t o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(post ? fmaxf(f[i + e], 0.f) : f[i + e]);
                st_global_v8(dst + i, o);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                float4 o = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
                if (post) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                *reinterpret_cast<float4*>(dst + i) = o;
              }
            }
          }
          if (p.out_lowp != nullptr) {
            uint32_t pk[16];
            if (p.epi == kGemmEpiBiasReluLowp) {
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = pack_relu<FP16>(f[2 * i], f[2 * i + 1]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = pack_plain<FP16>(f[2 * i], f[2 * i + 1]);
            }
            uint16_t* dst = p.out_lowp + row * p.ldo_lowp + col;
            if (wide) {          // 2 x 32 bytes: whole sectors (four 16-byte stores per row left every sector half written)
              st_global_v8(dst, pk);
              st_global_v8(dst + 16, pk + 8);
            } else {
#pragma unroll
              for (int u = 0; u < 4; ++u) reinterpret_cast<uint4*>(dst)[u] = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * (kGBarAccEmpty + b));
    }
  }
done:
  tc_fence_before();
  __syncthreads();
  if (warp == kGEpiWarps + 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

}  // namespace

cudaError_t gemm_tc_init() {
  const void* fns[4] = {reinterpret_cast<const void*>(gemm_tc_kernel<false, false>), reinterpret_cast<const void*>(gemm_tc_kernel<false, true>),
                        reinterpret_cast<const void*>(gemm_tc_kernel<true, false>), reinterpret_cast<const void*>(gemm_tc_kernel<true, true>)};
  for (const void* f : fns) {
    cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGSmem));
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// A and B as row-major 16-bit matrices: NT: a [M][lda >= K], b [N][ldb >= K]; TN: a [K][lda >= M], b [K][ldb >= N].
// Leading dimensions in elements, multiples of 8; N a multiple of p.bn (64, 128 or 256); for TN also M % 64 == 0.
cudaError_t launch_gemm_tc(const GemmParams& p, const uint16_t* a, int lda, const uint16_t* b, int ldb, bool tn, bool fp16,
                           int num_sms, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return cudaSuccess;
  if ((p.bn != 64 && p.bn != 128 && p.bn != 256) || p.N % p.bn != 0 || (lda & 7) || (ldb & 7) || p.ksplit < 1) return cudaErrorInvalidValue;
  const int ksteps = (p.K + kGBK - 1) / kGBK;
  if (p.ksplit > ksteps) return cudaErrorInvalidValue;
  alignas(64) unsigned char tma[128], tmb[128];
  cudaError_t e;
  if (!tn) {
    const unsigned long long ad[2] = {static_cast<unsigned long long>(p.K), static_cast<unsigned long long>(p.M)};
    const unsigned long long as[1] = {static_cast<unsigned long long>(lda) * 2};
    const unsigned abox[2] = {kGBK, kGBM};
    e = make_tensor_map(tma, a, 2, 2, ad, as, abox, true);
    if (e != cudaSuccess) return e;
    const unsigned long long bd[2] = {static_cast<unsigned long long>(p.K), static_cast<unsigned long long>(p.N)};
    const unsigned long long bs[1] = {static_cast<unsigned long long>(ldb) * 2};
    const unsigned bbox[2] = {kGBK, static_cast<unsigned>(p.bn < 128 ? p.bn : 128)};
    e = make_tensor_map(tmb, b, 2, 2, bd, bs, bbox, true);
    if (e != cudaSuccess) return e;
  } else {
    if (p.M % 64 != 0) return cudaErrorInvalidValue;
    const unsigned long long ad[2] = {static_cast<unsigned long long>(p.M), static_cast<unsigned long long>(p.K)};
    const unsigned long long as[1] = {static_cast<unsigned long long>(lda) * 2};
    const unsigned box[2] = {64, kGBK};
    e = make_tensor_map(tma, a, 2, 2, ad, as, box, true);
    if (e != cudaSuccess) return e;
    const unsigned long long bd[2] = {static_cast<unsigned long long>(p.N), static_cast<unsigned long long>(p.K)};
    const unsigned long long bs[1] = {static_cast<unsigned long long>(ldb) * 2};
    e = make_tensor_map(tmb, b, 2, 2, bd, bs, box, true);
    if (e != cudaSuccess) return e;
  }
  const long long items = static_cast<long long>((p.M + kGBM - 1) / kGBM) * (p.N / p.bn) * p.ksplit;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(items < num_sms ? items : num_sms));
  cfg.blockDim = dim3(kGThreads);
  cfg.dynamicSmemBytes = kGSmem;
  cfg.stream = stream;
  const CUtensorMap* ta = reinterpret_cast<const CUtensorMap*>(tma);
  const CUtensorMap* tb = reinterpret_cast<const CUtensorMap*>(tmb);
  if (fp16) return tn ? cudaLaunchKernelEx(&cfg, gemm_tc_kernel<true, true>, p, *ta, *tb) : cudaLaunchKernelEx(&cfg, gemm_tc_kernel<true, false>, p, *ta, *tb);
  return tn ? cudaLaunchKernelEx(&cfg, gemm_tc_kernel<false, true>, p, *ta, *tb) : cudaLaunchKernelEx(&cfg, gemm_tc_kernel<false, false>, p, *ta, *tb);
}

}  // namespace sdfb

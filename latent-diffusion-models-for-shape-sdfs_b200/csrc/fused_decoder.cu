// K1: fused persistent SDF decoder for sm_100a (tcgen05 / TMEM / bulk-async copies).
//
// What it computes (oracle: oracle/decoder.py
// decoder_forward_lowp - there is no upstream source, /root/reference/README.md:1):
//   h0 = relu(xyz W0x^T + bias0')                     fp32 FFMA in the epilogue warps
//   h1..h7 chained 512-wide layers                    tcgen05.mma, 16-bit operands, fp32 in TMEM
//   (L3 emits 253 features, L4 consumes [h3 | xyz] with the latent folded into bias4')
//   sdf = tanh(h7 . w8 + b8)                          fp32, h7 never rounded
//
// Structure: one CTA per SM, persistent over 128-query tiles (static round robin).
//   warps 0-3  epilogue: TMEM -> registers -> +bias, ReLU, pack -> shared memory (the next
//              layer's A operand, written IN PLACE over the current one), L0, head, store
//   warp 4     producer: streams the 96 weight blocks of a tile through a 3-stage ring
//   warp 5     MMA issuer: one lane issues every tcgen05.mma and the commits
//
// The activations of a tile live in eight 16 KiB shared-memory chunks (64 features each,
// K-major, 128B swizzle).  A layer's output is produced in two N=256 passes that ping-pong
// between the two 256-column halves of TMEM; the epilogue of one pass runs while the tensor
// core works on the next.  Chunk c of the next layer's input overwrites chunk c of the
// current one as soon as the last MMA reading it has committed (a_free[c]); the MMA issuer
// starts a layer as soon as the chunks it needs have been published (a_ready[c]).  The first
// layer of the NEXT tile is computed by the epilogue warps while the last layer of the
// current tile is still on the tensor core, so the pipe never drains between tiles.
#include "kernels.h"
#include "ptx.cuh"

namespace sdfb {

namespace {

constexpr int kStages = 3;
constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;

constexpr uint32_t kSmemA = 0;
constexpr uint32_t kSmemW = kAChunks * kAChunkBytes;                    // 131072
constexpr uint32_t kSmemBar = kSmemW + kStages * kBlockBytes;           // 229376
// barrier slots (8 bytes each)
constexpr int kBarWFull = 0;                   // [kStages]
constexpr int kBarWEmpty = kBarWFull + kStages;
constexpr int kBarAccFull = kBarWEmpty + kStages;   // [2]
constexpr int kBarAccEmpty = kBarAccFull + 2;       // [2]
constexpr int kBarAReady = kBarAccEmpty + 2;        // [8]
constexpr int kBarAFree = kBarAReady + kAChunks;    // [8]
constexpr int kNumBars = kBarAFree + kAChunks;
constexpr uint32_t kSmemMisc = kSmemBar + kNumBars * 8;                 // tmem ptr, abort flag
constexpr uint32_t kSmemBytes = kSmemMisc + 16;
constexpr uint32_t kSmemAlloc = kSmemBytes + 1024;                      // slack for 1024B alignment
static_assert(kSmemAlloc <= 232448, "exceeds the 227 KiB opt-in shared memory of sm_100");

// watchdog site codes
enum : uint32_t {
  kErrWFull = 0x10, kErrWEmpty = 0x20, kErrAccFull = 0x30, kErrAccEmpty = 0x40,
  kErrAReady = 0x50, kErrAFree = 0x60,
};

// static description of the 13 tensor-core passes of a tile
__device__ __forceinline__ int pass_chunks(int p) { return (p == 5 || p == 6) ? 4 : 8; }
// does pass p start a layer (must wait for published input chunks)?
__device__ __forceinline__ bool pass_first(int p) { return (0x0AB5u >> p) & 1u; }  // {0,2,4,5,7,9,11}
// is pass p the last reader of its layer's input chunks?
__device__ __forceinline__ bool pass_last(int p) { return (0x155Au >> p) & 1u; }   // {1,3,4,6,8,10,12}

template <bool FP16>
__device__ __forceinline__ uint32_t pack_relu(float lo, float hi) {
  uint32_t d;
  if constexpr (FP16)
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

struct EpiState {
  uint32_t wphase;      // bit c: parity of the next a_free[c] wait
  uint32_t acc_phase;   // bit b: parity of the next acc_full[b] wait
};

// 64 features of h0 = relu(xyz W0x^T + bias0') for this thread's query -> chunk c (in place).
template <bool FP16>
__device__ __forceinline__ bool epi_layer0_chunk(const DecConsts* __restrict__ cs, int c, float x, float y,
                                                 float z, uint32_t a_row_addr, uint32_t row7,
                                                 uint32_t bars, EpiState& st, const Watchdog& wd) {
  uint32_t packed[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int n = c * 64 + 2 * j;
    const float4 w0 = __ldg(&cs->l0[n]);
    const float4 w1 = __ldg(&cs->l0[n + 1]);
    const float f0 = fmaf(z, w0.z, fmaf(y, w0.y, fmaf(x, w0.x, w0.w)));
    const float f1 = fmaf(z, w1.z, fmaf(y, w1.y, fmaf(x, w1.x, w1.w)));
    packed[j] = pack_relu<FP16>(f0, f1);
  }
  if (!mbar_wait(bars + 8 * (kBarAFree + c), ((st.wphase >> c) & 1u) ^ 1u, wd, kErrAFree, c)) return false;
  st.wphase ^= 1u << c;
  const uint32_t base = a_row_addr + c * kAChunkBytes;
#pragma unroll
  for (int u = 0; u < 8; ++u)
    st_shared_v4(base + ((u ^ row7) << 4), packed[4 * u], packed[4 * u + 1], packed[4 * u + 2],
                 packed[4 * u + 3]);
  fence_proxy_async_smem();
  mbar_arrive(bars + 8 * (kBarAReady + c));
  return true;
}

// One hidden pass: accumulator half `b` (256 columns) -> +bias (+ xyz term for L4) -> ReLU ->
// 16-bit -> chunks [c0, c0+4) of the activation buffer.
template <bool FP16, bool XYZ>
__device__ __forceinline__ bool epi_hidden_pass(const float* __restrict__ bias, const float4* __restrict__ l4x,
                                                float x, float y, float z, int c0, uint32_t tmem_row,
                                                int b, uint32_t a_row_addr, uint32_t row7, uint32_t bars,
                                                EpiState& st, const Watchdog& wd, float* dump_row) {
  if (!mbar_wait(bars + 8 * (kBarAccFull + b), (st.acc_phase >> b) & 1u, wd, kErrAccFull, b)) return false;
  st.acc_phase ^= 1u << b;
  tc_fence_after();
#pragma unroll 1
  for (int cc = 0; cc < 4; ++cc) {
    uint32_t packed[32];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      uint32_t v[32];
      const int col = cc * 64 + g * 32;
      tmem_ld32(tmem_row + b * 256 + col, v);
      float bb[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = ldg4(bias + col + 4 * j);
        bb[4 * j] = t.x; bb[4 * j + 1] = t.y; bb[4 * j + 2] = t.z; bb[4 * j + 3] = t.w;
      }
      if constexpr (XYZ) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float4 w = __ldg(&l4x[col + j]);
          bb[j] = fmaf(z, w.z, fmaf(y, w.y, fmaf(x, w.x, bb[j])));
        }
      }
      tmem_ld_wait();
      if (dump_row != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) dump_row[col + j] = __uint_as_float(v[j]) + bb[j];
      }
#pragma unroll
      for (int j = 0; j < 16; ++j)
        packed[g * 16 + j] = pack_relu<FP16>(__uint_as_float(v[2 * j]) + bb[2 * j],
                                             __uint_as_float(v[2 * j + 1]) + bb[2 * j + 1]);
    }
    if (cc == 3) {  // every column of this half has been read: hand the accumulator back
      tc_fence_before();
      mbar_arrive(bars + 8 * (kBarAccEmpty + b));
    }
    const int c = c0 + cc;
    if (!mbar_wait(bars + 8 * (kBarAFree + c), ((st.wphase >> c) & 1u) ^ 1u, wd, kErrAFree, c)) return false;
    st.wphase ^= 1u << c;
    const uint32_t base = a_row_addr + c * kAChunkBytes;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      st_shared_v4(base + ((u ^ row7) << 4), packed[4 * u], packed[4 * u + 1], packed[4 * u + 2],
                   packed[4 * u + 3]);
    fence_proxy_async_smem();
    mbar_arrive(bars + 8 * (kBarAReady + c));
  }
  return true;
}

// One half of the last hidden layer: h7 = relu(acc + b7) stays fp32 and is reduced against w8.
__device__ __forceinline__ bool epi_head_pass(const float* __restrict__ bias, const float* __restrict__ head,
                                              uint32_t tmem_row, int b, uint32_t bars, EpiState& st,
                                              const Watchdog& wd, float& dot, float* dump_row) {
  if (!mbar_wait(bars + 8 * (kBarAccFull + b), (st.acc_phase >> b) & 1u, wd, kErrAccFull, b)) return false;
  st.acc_phase ^= 1u << b;
  tc_fence_after();
#pragma unroll 1
  for (int g = 0; g < 8; ++g) {
    uint32_t v[32];
    const int col = g * 32;
    tmem_ld32(tmem_row + b * 256 + col, v);
    float bb[32], hw[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = ldg4(bias + col + 4 * j);
      bb[4 * j] = t.x; bb[4 * j + 1] = t.y; bb[4 * j + 2] = t.z; bb[4 * j + 3] = t.w;
      const float4 h = ldg4(head + col + 4 * j);
      hw[4 * j] = h.x; hw[4 * j + 1] = h.y; hw[4 * j + 2] = h.z; hw[4 * j + 3] = h.w;
    }
    tmem_ld_wait();
    if (dump_row != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; ++j) dump_row[col + j] = __uint_as_float(v[j]) + bb[j];
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) dot = fmaf(fmaxf(__uint_as_float(v[j]) + bb[j], 0.f), hw[j], dot);
  }
  tc_fence_before();
  mbar_arrive(bars + 8 * (kBarAccEmpty + b));
  return true;
}

struct Query { float x, y, z; };

__device__ __forceinline__ Query load_query(const DecodeParams& p, long long tile, int row) {
  Query q{0.f, 0.f, 0.f};
  const long long m = tile * kTileM + row;
  if (m >= p.M) return q;
  if (p.xyz != nullptr) {
    q.x = __ldg(p.xyz + 3 * m); q.y = __ldg(p.xyz + 3 * m + 1); q.z = __ldg(p.xyz + 3 * m + 2);
  } else {
    const long long g = p.q0 + m;
    const long long t = g / p.res;
    const int ix = static_cast<int>(g - t * p.res);
    const int iz = static_cast<int>(t / p.res);
    const int iy = static_cast<int>(t - static_cast<long long>(iz) * p.res);
    const float den = static_cast<float>(p.res - 1);
    q.x = __fdiv_rn(axis_coord_num(ix, p.res), den);
    q.y = __fdiv_rn(axis_coord_num(iy, p.res), den);
    q.z = __fdiv_rn(axis_coord_num(iz, p.res), den);
  }
  return q;
}

template <bool FP16>
__global__ void __launch_bounds__(kThreads, 1) fused_decoder_kernel(const DecodeParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
  const uint32_t bars = smem0 + kSmemBar;
  volatile uint32_t* misc = reinterpret_cast<volatile uint32_t*>(smem_gen + kSmemMisc);   // [0] tmem base, [1] abort
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const long long num_tiles = (p.M + kTileM - 1) / kTileM;
  const long long first_tile = blockIdx.x;
  const long long my_tiles = first_tile < num_tiles ? (num_tiles - first_tile + gridDim.x - 1) / gridDim.x : 0;

  if (threadIdx.x == 0) {
    misc[1] = 0;
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bars + 8 * (kBarWFull + s), 1);
      mbar_init(bars + 8 * (kBarWEmpty + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bars + 8 * (kBarAccFull + b), 1);
      mbar_init(bars + 8 * (kBarAccEmpty + b), kEpiThreads);
    }
    for (int c = 0; c < kAChunks; ++c) {
      mbar_init(bars + 8 * (kBarAReady + c), kEpiThreads);
      mbar_init(bars + 8 * (kBarAFree + c), 1);
    }
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc<1>(smem0 + kSmemMisc, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = misc[0];
  long long waited[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  Watchdog wd{misc + 1, p.status, p.timeout_ns, p.prof != nullptr ? waited : nullptr};
  const long long t_start = clock64();

  if (warp == 4) {
    // ===================== producer: weight stream -> 3-stage ring =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (long long it = 0; it < my_tiles; ++it) {
        const uint8_t* src = p.wstream;
#pragma unroll 1
        for (int blk = 0; blk < kBlocksPerTile; ++blk, src += kBlockBytes) {
          if (!mbar_wait(bars + 8 * (kBarWEmpty + stage), phase ^ 1u, wd, kErrWEmpty, stage)) goto done;
          const uint32_t full = bars + 8 * (kBarWFull + stage);
          if (p.debug_flags & 1u) {
            mbar_arrive(full);
          } else {
            mbar_arrive_expect_tx(full, kBlockBytes);
            bulk_g2s(smem0 + kSmemW + stage * kBlockBytes, src, kBlockBytes, full);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(128, 256, FP16 ? 0 : 1);
      uint32_t stage = 0, phase = 0;
      uint32_t rphase = 0;      // bit c: parity of the next a_ready[c] wait
      uint32_t ephase = 0;      // bit b: parity state of acc_empty[b]
      uint32_t gpass = 0;       // global pass counter -> accumulator half
      for (long long it = 0; it < my_tiles; ++it) {
#pragma unroll 1
        for (int ps = 0; ps < kPasses; ++ps, ++gpass) {
          const int nk = pass_chunks(ps);
          const bool first = pass_first(ps), last = pass_last(ps);
          const uint32_t b = gpass & 1u;
          const uint32_t d_tmem = tmem_base + b * 256;
          if (!mbar_wait(bars + 8 * (kBarAccEmpty + b), ((ephase >> b) & 1u) ^ 1u, wd, kErrAccEmpty, b)) goto done;
          ephase ^= 1u << b;
#pragma unroll 1
          for (int k = 0; k < nk; ++k) {
            if (first) {
              if (!mbar_wait(bars + 8 * (kBarAReady + k), (rphase >> k) & 1u, wd, kErrAReady, k)) goto done;
              rphase ^= 1u << k;
            }
            if (!mbar_wait(bars + 8 * (kBarWFull + stage), phase, wd, kErrWFull, stage)) goto done;
            tc_fence_after();
            const uint64_t adesc = umma_desc_sw128(smem0 + kSmemA + k * kAChunkBytes);
            const uint64_t bdesc = umma_desc_sw128(smem0 + kSmemW + stage * kBlockBytes);
#pragma unroll
            for (int j = 0; j < 4; ++j)   // 4 x K=16 inside the 128-byte swizzle atom: +32 B each
              umma_ss<1>(d_tmem, adesc + 2 * j, bdesc + 2 * j, idesc, (k | j) != 0 ? 1u : 0u);
            umma_commit<1>(bars + 8 * (kBarWEmpty + stage));
            if (last) umma_commit<1>(bars + 8 * (kBarAFree + k));
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          umma_commit<1>(bars + 8 * (kBarAccFull + b));
        }
      }
      // L4 reads only chunks 0..3, so chunks 4..7 see one reader fewer per tile than chunks
      // 0..3; nothing to fix up: both sides count events per chunk.
    }
  } else {
    // ===================== epilogue warps (thread <-> query row) =====================
    const int row = threadIdx.x;                    // 0..127 == TMEM lane
    const uint32_t row7 = row & 7u;
    const uint32_t a_row_addr = smem0 + kSmemA + row * 128;
    const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const DecConsts* __restrict__ cs = p.consts;
    EpiState st{0u, 0u};
    uint32_t gpass = 0;
    if (my_tiles > 0) {
      Query q = load_query(p, first_tile, row);
#pragma unroll 1
      for (int c = 0; c < kAChunks; ++c)
        if (!epi_layer0_chunk<FP16>(cs, c, q.x, q.y, q.z, a_row_addr, row7, bars, st, wd)) goto done;
      for (long long it = 0; it < my_tiles; ++it) {
        const long long tile = first_tile + it * gridDim.x;
        float* dump_row = nullptr;
#pragma unroll 1
        for (int ps = 0; ps < 11; ++ps, ++gpass) {
          // pass -> (bias row, column offset, destination chunks)
          int layer, half;
          if (ps < 4) { layer = 1 + (ps >> 1); half = ps & 1; }
          else if (ps == 4) { layer = 3; half = 0; }
          else { layer = 4 + ((ps - 5) >> 1); half = (ps - 5) & 1; }
          const float* bias = cs->bias[layer - 1] + half * 256;
          dump_row = (p.dump != nullptr && tile == 0 && ps == p.dump_pass) ? p.dump + row * 256 : nullptr;
          bool ok;
          if (layer == 4)
            ok = epi_hidden_pass<FP16, true>(bias, cs->l4x + half * 256, q.x, q.y, q.z, half * 4, tmem_row,
                                             gpass & 1u, a_row_addr, row7, bars, st, wd, dump_row);
          else
            ok = epi_hidden_pass<FP16, false>(bias, nullptr, 0.f, 0.f, 0.f, half * 4, tmem_row, gpass & 1u,
                                              a_row_addr, row7, bars, st, wd, dump_row);
          if (!ok) goto done;
        }
        float dot = 0.f;
        dump_row = (p.dump != nullptr && tile == 0 && p.dump_pass == 11) ? p.dump + row * 256 : nullptr;
        if (!epi_head_pass(cs->bias[6], cs->head, tmem_row, gpass & 1u, bars, st, wd, dot, dump_row)) goto done;
        ++gpass;
        // first layer of the next tile, written behind the last readers of this tile's h6
        Query qn{0.f, 0.f, 0.f};
        if (it + 1 < my_tiles) {
          qn = load_query(p, tile + gridDim.x, row);
#pragma unroll 1
          for (int c = 0; c < kAChunks; ++c)
            if (!epi_layer0_chunk<FP16>(cs, c, qn.x, qn.y, qn.z, a_row_addr, row7, bars, st, wd)) goto done;
        }
        dump_row = (p.dump != nullptr && tile == 0 && p.dump_pass == 12) ? p.dump + row * 256 : nullptr;
        if (!epi_head_pass(cs->bias[6] + 256, cs->head + 256, tmem_row, gpass & 1u, bars, st, wd, dot, dump_row))
          goto done;
        ++gpass;
        const long long m = tile * kTileM + row;
        if (m < p.M) p.out[m] = tanhf(dot + __ldg(&cs->head_b[0]));
        q = qn;
      }
    }
  }
done:
  if (p.prof != nullptr && (lane == 0) && (warp == 0 || warp >= 4)) {
    // blocked cycles per wait class (index = site >> 4) and this role's total, per CTA and role
    const int role = warp == 0 ? 0 : warp - 3;            // 0 epilogue, 1 producer, 2 MMA issuer
    long long* dst = p.prof + (static_cast<long long>(blockIdx.x) * 3 + role) * 8;
#pragma unroll
    for (int i = 1; i < 7; ++i) dst[i] = waited[i];
    dst[0] = clock64() - t_start;
    dst[7] = my_tiles;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

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
  Watchdog wd{misc + 1, status, 200000000ull, nullptr};
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

cudaError_t fused_decoder_init() {
  cudaError_t e = cudaFuncSetAttribute(fused_decoder_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kSmemAlloc));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(fused_decoder_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(kSmemAlloc));
  if (e != cudaSuccess) return e;
  constexpr int st_bytes = kAChunkBytes + kBlockBytes + 64 + 1024;
  e = cudaFuncSetAttribute(umma_selftest_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, st_bytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(umma_selftest_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, st_bytes);
}

cudaError_t launch_fused_decoder(const DecodeParams& p, bool fp16, int num_sms, cudaStream_t stream) {
  if (p.M <= 0) return cudaSuccess;
  const long long tiles = (p.M + kTileM - 1) / kTileM;
  const unsigned grid = static_cast<unsigned>(tiles < num_sms ? tiles : num_sms);
  if (fp16)
    fused_decoder_kernel<true><<<grid, kThreads, kSmemAlloc, stream>>>(p);
  else
    fused_decoder_kernel<false><<<grid, kThreads, kSmemAlloc, stream>>>(p);
  return cudaGetLastError();
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

// K1: fused persistent SDF decoder for sm_100a (tcgen05 / TMEM / TMA, CTA pairs).
//
// What it computes (oracle: oracle/decoder.py
// decoder_forward_lowp; no upstream source exists, /root/reference/README.md:1):
//   h0 = relu(xyz W0x^T + bias0')              fp32 FFMA in the epilogue warps (latent folded)
//   h1..h7 chained 512-wide layers             tcgen05.mma, 16-bit operands, fp32 accum in TMEM
//   (L3 emits 253 features; its 3 padding columns carry xyz into L4, latent folded in bias4')
//   sdf = tanh(h7 . w8 + b8)                   fp32, h7 never rounded
// One CTA pair carries a tile of 256 queries through all layers; activations never leave the
// SMs.  Per tile the tensor core runs 13 passes (kernels.h); a layer's output is produced in
// two N=256 passes that ping-pong between the two halves of TMEM, and the epilogue writes the
// next layer's A operand IN PLACE over the current one, chunk by chunk, as soon as the last
// MMA reading a chunk has committed.  The first layer of the NEXT tile is computed while the
// last layer of the current one is still on the tensor core.
//
// Why pairs (cta_group::2): each SM supplies its own 128 rows of A and only HALF of every
// weight block, which halves the shared-memory operand traffic per MMA and the L2->SM weight
// stream (a single-CTA version of this kernel measured 54.5 ms for 256^3).
//
//   cluster = 2 CTAs = 256 queries;  CTA rank r owns rows [128r, 128r+128) of the pair tile
//   warps 0-7  epilogue  (lane quadrant = warp & 3; the two warp sets split each 64-wide chunk)
//   warp 8     producer: this CTA's half (16 KiB) of every weight block, tensor-map TMA,
//              completion counted on the LEADER's barrier (.cta_group::2)
//   warp 9     MMA issuer (leader CTA only) + TMEM allocation (both CTAs)
//
// Shared memory per CTA: activations 128 KiB (in place, 8 chunks), 5-slot weight ring 80 KiB,
// biases and head weights 16 KiB.
//
// Commit cadence: a tcgen05.commit after every 4 MMAs costs ~245 cycles per commit point
// (tools/umma_rate.py: 189 cycles/MMA instead of 128), after every 8 MMAs nothing.  So the
// issuer commits once per PAIR of weight blocks and releases both ring slots, both activation
// chunks and (at the end of a pass) the accumulator at that one point.
//
// BWD = true (oracle: oracle/decoder.py decoder_vjp_latent_lowp): the same
// tile then runs BACKWARDS through the same machinery - 13 more passes against the transposed weight
// blocks (kernels.h) - to produce what the latent's gradient needs:
//   delta7' = w8 where h7 > 0 (epilogue, like the first layer; written half by half behind the last forward
//   pass);  g = dLdy (1 - sdf^2) is applied per row to the accumulator of the first backward layer;
//   delta_{l-1} = mask_{l-1} * (delta_l W_l): the accumulator is masked instead of biased + rectified and
//   written back in place as the next A operand.  The ReLU masks of layers 1-6 are one 32-bit word per
//   (thread, 32 columns) parked in an L2-resident scratch (64 KiB per CTA, written and read by the same
//   thread); layer 7's stay in registers; layer 0's are warp ballots of the first-layer epilogue, two tiles deep.
//   The latent enters through the bias of layers 0 and 4 only, so all it needs are the COLUMN SUMS of
//   delta0 and delta4 over the queries: a 31-shuffle transpose-reduce per 32 x 32 block leaves lane l
//   with column l's sum, accumulated in registers over all tiles and written once per warp at the end
//   (vjp_finish_kernel contracts them with the fp32 latent columns of W0 and W4).
//   Loss mode (DecodeParams::target): dLdy is formed here from the value just decoded (clamped-L1 fitting loss).
#include <cuda.h>

#include "kernels.h"
#include "ptx.cuh"

namespace sdfb {

namespace {

constexpr int kStages = 5;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = kEpiThreads + 64;
constexpr uint32_t kHalfBlockBytes = kBlockBytes / 2;                  // 16 KiB: 128 weight rows x 64 k

constexpr uint32_t oA = 0;
constexpr uint32_t oW = kAChunks * kAChunkBytes;                       // 131072
constexpr uint32_t oBias = oW + kStages * kHalfBlockBytes;             // 7 x 512 floats
constexpr uint32_t oHead = oBias + 7 * kHid * 4;                       // 512 floats
constexpr uint32_t oXyz = oHead + kHid * 4;                            // 128 x float4: the tile's coordinates
constexpr uint32_t oDot = oXyz + kTileM * 16;                          // 128 floats: head partial sums
constexpr uint32_t oBar = oDot + kTileM * 4;
constexpr int kBarWFull = 0;
constexpr int kBarWEmpty = kBarWFull + kStages;
constexpr int kBarAccFull = kBarWEmpty + kStages;
constexpr int kBarAccEmpty = kBarAccFull + 2;
constexpr int kBarAReady = kBarAccEmpty + 2;
constexpr int kBarAFree = kBarAReady + kAChunks;
constexpr int kNumBars = kBarAFree + kAChunks;
constexpr uint32_t oMisc = oBar + kNumBars * 8;
constexpr uint32_t kSmemBytes = oMisc + 16;
constexpr uint32_t kSmemAlloc = kSmemBytes;     // no slack: the dynamic window is declared 1024-byte aligned
static_assert(kSmemAlloc <= 232448, "exceeds the 227 KiB opt-in shared memory of sm_100");
static_assert(oBar % 8 == 0, "barriers must be 8-byte aligned");

enum : uint32_t {
  kErrWFull = 0x10, kErrWEmpty = 0x20, kErrAccFull = 0x30, kErrAccEmpty = 0x40,
  kErrAReady = 0x50, kErrAFree = 0x60,
};

// Pass tables (kernels.h).  `first`: the pass is the first reader of a freshly written A operand (waits
// for its chunks); `last`: the last reader (releases them).  Backward passes 13-25: delta6, delta5, delta4
// in halves, 19 = delta3 (N = 256), 20/21 = delta2 (K = 256), delta1 and delta0 in halves.
constexpr uint32_t kFirstFwd = 0x0AB5u;                                           // {0,2,4,5,7,9,11}
constexpr uint32_t kLastFwd = 0x155Au;                                            // {1,3,4,6,8,10,12}
constexpr uint32_t kFirstBwd = kFirstFwd | (1u << 13) | (1u << 15) | (1u << 17) | (1u << 19) | (1u << 20) | (1u << 22) | (1u << 24);
constexpr uint32_t kLastBwd = kLastFwd | (1u << 14) | (1u << 16) | (1u << 18) | (1u << 19) | (1u << 21) | (1u << 23) | (1u << 25);
template <bool BWD>
__device__ __forceinline__ int pass_chunks(int p) { return (p == 5 || p == 6 || (BWD && (p == 20 || p == 21))) ? 4 : 8; }
template <bool BWD>
__device__ __forceinline__ bool pass_first(int p) { return ((BWD ? kFirstBwd : kFirstFwd) >> p) & 1u; }
template <bool BWD>
__device__ __forceinline__ bool pass_last(int p) { return ((BWD ? kLastBwd : kLastFwd) >> p) & 1u; }

// Diagnostics (SDFB_PROF): CTA 0 stamps clock64() at per-pass events of its tile number 5 into the tail of the profile
// buffer, [prof_tail + 32 * kind + pass]: kind 0 = issuer about to issue the pass' first MMA, 1 = issuer has issued its last
// MMA, 2 = epilogue warp 0 sees the accumulator full, 3 = epilogue warp 0 has written its last chunk of the pass.
#define SDFB_K1_TRACE(kind, pass)                                                                        \
  do {                                                                                                   \
    if (p.prof != nullptr && blockIdx.x == 0 && it == 5 && (threadIdx.x & 31) == 0)                      \
      p.prof[static_cast<size_t>(gridDim.x) * 24 + 32 * (kind) + (pass)] = clock64();                    \
  } while (0)

struct Query { float x, y, z; };

__device__ __forceinline__ Query load_query(const DecodeParams& p, long long m) {
  Query q{0.f, 0.f, 0.f};
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

// A compile-time bool usable where a runtime bool is accepted (see epi_hidden_pass).
template <bool V>
struct Flag {
  __device__ __forceinline__ constexpr operator bool() const { return V; }
};

// ReLU masks (BWD kernels) are built most-significant-bit first, two instructions per value: column i of a
// 32-column group ends up in bit 31 - i.  0 - f is negative exactly when f > 0 (+0 and -0 both give +0).
__device__ __forceinline__ uint32_t push_positive(uint32_t bits, float f) {
  return __funnelshift_l(__float_as_uint(__fsub_rn(0.f, f)), bits, 1);
}

struct Epi {
  uint32_t bars;          // shared address of the barrier array
  uint32_t tmem_row;      // TMEM address of this warp's lane quadrant, column 0
  uint32_t a_row_addr;    // shared address of this thread's row in chunk 0
  uint32_t row7;
  uint32_t wphase;        // bit c: parity of the next a_free[c] wait (tracked for all chunks)
  uint32_t acc_phase;     // bit b: parity of the next acc_full[b] wait
  int set;                // 0/1: which of the two warp sets (splits chunks / head columns)
  int lane;
  long long* trace;       // diagnostics: where to stamp "accumulator full seen" of the next pass (nullptr: off)
  uint32_t dbg;           // DecodeParams::debug_flags (timing experiments under SDFB_K1_EXPERIMENTS: bit2 / bit3, see kernels.h)
};

// Hidden pass.  The two warp sets work on the SAME chunk at the same time (set s converts
// columns [32s, 32s+32) of it), chunk after chunk, so chunks become available in the order the
// next layer's MMAs consume them.  The TMEM load of the next chunk is in flight while the
// current one is converted.  `L3`: this is layer 3, whose last three (padding) columns carry
// the query coordinates into layer 4.
// `mrow` (BWD kernels only): where this thread parks the ReLU mask of its 32 columns of chunk cc: word
// (2 cc + set) * 128 of the (layer, half) block, already offset by the row.
// `l3` / `wait_free` are Flag<true> / Flag<false> in the forward kernel (folded at compile time, three
// instances) and plain bools in the BWD kernel (ONE instance: that kernel's code must stay close to the
// instruction cache's size - the first version, 240 KB of SASS, spent 17 % of its issue slots on no_inst stalls).
template <bool FP16, bool MASK, typename L3T, typename WFT>
__device__ __forceinline__ bool epi_hidden_pass(Epi& e, const float* __restrict__ sbias, Query q, int c0, int b,
                                                const Watchdog& wd, float* dump_row, L3T l3, WFT wait_free,
                                                uint32_t* mrow = nullptr) {
  if (!mbar_wait(e.bars + 8 * (kBarAccFull + b), (e.acc_phase >> b) & 1u, wd, kErrAccFull, b)) return false;
  e.acc_phase ^= 1u << b;
  if (e.trace != nullptr && e.lane == 0) *e.trace = clock64();
  __syncwarp();
  tc_fence_after();
  const uint32_t tbase = e.tmem_row + b * 256 + e.set * 32;
  uint32_t v[2][32];
  tmem_ld32(tbase, v[0]);
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    tmem_ld_wait();
    if (cc < 3) {
      tmem_ld32(tbase + (cc + 1) * 64, v[(cc + 1) & 1]);
    } else {      // every column this warp owns has been read: hand the accumulator back
      tc_fence_before();
      __syncwarp();
      if (e.lane == 0) arrive_on_leader(e.bars + 8 * (kBarAccEmpty + b), 1);
    }
    const uint32_t(&vc)[32] = v[cc & 1];
    const int col = cc * 64 + e.set * 32;
    uint32_t packed[16];
    uint32_t bits = 0;
#ifdef SDFB_K1_EXPERIMENTS
    if (e.dbg & 4u) {      // timing experiment: no conversion, no stores - only the barrier protocol (results are garbage)
      if (wait_free) {
        if (!mbar_wait(e.bars + 8 * (kBarAFree + c0 + cc), ((e.wphase >> (c0 + cc)) & 1u) ^ 1u, wd, kErrAFree, c0 + cc)) return false;
      }
      __syncwarp();
      if (e.lane == 0) arrive_on_leader(e.bars + 8 * (kBarAReady + c0 + cc), 1);
      continue;
    }
#endif
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = *reinterpret_cast<const float4*>(sbias + col + 4 * j);
      const float f0 = __uint_as_float(vc[4 * j]) + t.x, f1 = __uint_as_float(vc[4 * j + 1]) + t.y;
      const float f2 = __uint_as_float(vc[4 * j + 2]) + t.z, f3 = __uint_as_float(vc[4 * j + 3]) + t.w;
      if (dump_row != nullptr) {
        dump_row[col + 4 * j] = f0; dump_row[col + 4 * j + 1] = f1;
        dump_row[col + 4 * j + 2] = f2; dump_row[col + 4 * j + 3] = f3;
      }
      packed[2 * j] = pack_relu<FP16>(f0, f1);
      packed[2 * j + 1] = pack_relu<FP16>(f2, f3);
      if constexpr (MASK) bits = push_positive(push_positive(push_positive(push_positive(bits, f0), f1), f2), f3);
      if (l3) {
        if (j == 7 && cc == 3 && e.set == 1) {   // features 252 | x, y | z  (x, y, z unrectified)
          packed[14] = pack_plain<FP16>(fmaxf(f0, 0.f), q.x);
          packed[15] = pack_plain<FP16>(q.y, q.z);
          if constexpr (MASK) bits &= ~7u;   // columns 253-255 (the last three pushed) are not features: nothing flows back
        }
      }
    }
    if constexpr (MASK) mrow[(2 * cc + e.set) * kTileM] = bits;
    const int c = c0 + cc;
    // Chunks written by a layer's LAST pass were read for the last time by that very pass, whose
    // completion acc_full already reported: only a first-half pass must wait for the other half's
    // MMAs to release the chunk.  (The phase bits are advanced by schedule, not by waiting.)
    if (wait_free) {
      if (!mbar_wait(e.bars + 8 * (kBarAFree + c), ((e.wphase >> c) & 1u) ^ 1u, wd, kErrAFree, c)) return false;
    }
    const uint32_t base = e.a_row_addr + c * kAChunkBytes;
#ifdef SDFB_K1_EXPERIMENTS
    if (e.dbg & 8u) {      // (bit3, timing experiment: conversion but no stores / proxy fence)
      asm volatile("" ::"r"(packed[0] ^ packed[5] ^ packed[10] ^ packed[15]));
    } else
#endif
    {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        st_shared_v4(base + (((4 * e.set + u) ^ e.row7) << 4), packed[4 * u], packed[4 * u + 1], packed[4 * u + 2],
                     packed[4 * u + 3]);
      fence_proxy_async_smem();
    }
    __syncwarp();
    if (e.lane == 0) arrive_on_leader(e.bars + 8 * (kBarAReady + c), 1);
  }
  e.wphase ^= 0xFu << c0;
  return true;
}

// Head pass: this warp reduces columns [128*set, 128*set+128) of half `b` of h7 against w8.
// `hm` (BWD kernels): receives the ReLU mask of the four 32-column groups this thread reduces.
template <bool MASK = false>
__device__ __forceinline__ bool epi_head_pass(Epi& e, const float* __restrict__ sbias, const float* __restrict__ shead,
                                              int b, const Watchdog& wd, float& dot, float* dump_row,
                                              uint32_t* hm = nullptr) {
  if (!mbar_wait(e.bars + 8 * (kBarAccFull + b), (e.acc_phase >> b) & 1u, wd, kErrAccFull, b)) return false;
  e.acc_phase ^= 1u << b;
  if (e.trace != nullptr && e.lane == 0) *e.trace = clock64();
  __syncwarp();
  tc_fence_after();
  const uint32_t tbase = e.tmem_row + b * 256 + e.set * 128;
  uint32_t v[2][32];
  tmem_ld32(tbase, v[0]);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    tmem_ld_wait();
    if (g < 3) {
      tmem_ld32(tbase + (g + 1) * 32, v[(g + 1) & 1]);
    } else {
      tc_fence_before();
      __syncwarp();
      if (e.lane == 0) arrive_on_leader(e.bars + 8 * (kBarAccEmpty + b), 1);
    }
    const uint32_t(&vc)[32] = v[g & 1];
    const int col = e.set * 128 + g * 32;
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = *reinterpret_cast<const float4*>(sbias + col + 4 * j);
      const float4 h = *reinterpret_cast<const float4*>(shead + col + 4 * j);
      const float f0 = __uint_as_float(vc[4 * j]) + t.x, f1 = __uint_as_float(vc[4 * j + 1]) + t.y;
      const float f2 = __uint_as_float(vc[4 * j + 2]) + t.z, f3 = __uint_as_float(vc[4 * j + 3]) + t.w;
      if (dump_row != nullptr) {
        dump_row[col + 4 * j] = f0; dump_row[col + 4 * j + 1] = f1;
        dump_row[col + 4 * j + 2] = f2; dump_row[col + 4 * j + 3] = f3;
      }
      dot = fmaf(fmaxf(f0, 0.f), h.x, dot);
      dot = fmaf(fmaxf(f1, 0.f), h.y, dot);
      dot = fmaf(fmaxf(f2, 0.f), h.z, dot);
      dot = fmaf(fmaxf(f3, 0.f), h.w, dot);
      if constexpr (MASK) bits = push_positive(push_positive(push_positive(push_positive(bits, f0), f1), f2), f3);
    }
    if constexpr (MASK) hm[g] = bits;
  }
  return true;
}

// ---- backward epilogues (BWD kernels) ------------------------------------------------------------
// Transpose-reduce of a 32 (rows = lanes) x 32 (columns = v[]) block in 31 shuffles: on return lane l
// holds the sum over the warp's 32 rows of column l.  v is destroyed.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float recv = __shfl_xor_sync(0xffffffffu, send, off);
      v[i] = (up ? v[i + off] : v[i]) + recv;
    }
  }
  return v[0];
}

// The operand of the first backward layer, half b: w8 where h7's pre-activation was positive.  The per-query factor
// g = dLdy (1 - sdf^2) commutes with the product and is applied to the accumulator of passes 13 / 14 instead, so
// this half can be written as soon as its head pass is done - while the other head pass is still on the tensor
// core - by the thread that holds the masks: set s writes chunks 4b + 2s, 4b + 2s + 1 of its row.  Like the
// first layer, every chunk must have been released by the last forward pass first.  (wphase: flipped by the caller
// once both halves are written.)
template <bool FP16>
__device__ __forceinline__ bool epi_delta7_half(Epi& e, const float* __restrict__ shead, const uint32_t (&hm)[4], int b,
                                                const Watchdog& wd) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int c = 4 * b + 2 * e.set + (g >> 1);
    if ((g & 1) == 0) {
      if (!mbar_wait(e.bars + 8 * (kBarAFree + c), ((e.wphase >> c) & 1u) ^ 1u, wd, kErrAFree, c)) return false;
    }
    const int col = 256 * b + 128 * e.set + 32 * g;
    const uint32_t m = hm[g];
    uint32_t packed[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 h = *reinterpret_cast<const float4*>(shead + col + 4 * j);
      const float v0 = (m >> (31 - 4 * j)) & 1u ? h.x : 0.f, v1 = (m >> (30 - 4 * j)) & 1u ? h.y : 0.f;
      const float v2 = (m >> (29 - 4 * j)) & 1u ? h.z : 0.f, v3 = (m >> (28 - 4 * j)) & 1u ? h.w : 0.f;
      packed[2 * j] = pack_plain<FP16>(v0, v1);
      packed[2 * j + 1] = pack_plain<FP16>(v2, v3);
    }
    const uint32_t base = e.a_row_addr + c * kAChunkBytes;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      st_shared_v4(base + (((4 * (g & 1) + u) ^ e.row7) << 4), packed[4 * u], packed[4 * u + 1], packed[4 * u + 2],
                   packed[4 * u + 3]);
    if (g & 1) {
      fence_proxy_async_smem();
      __syncwarp();
      if (e.lane == 0) arrive_on_leader(e.bars + 8 * (kBarAReady + c), 2);   // 4 warps x 2 = the 8 expected per CTA
    }
  }
  return true;
}

// Backward pass epilogue (ONE instance, runtime flags): delta = mask * accumulator [* row_scale], then
//   write_back: rounded and written in place as the next A operand (chunks c0 .. c0 + 3; `wait_free` as in the
//               forward pass); same thread <-> column mapping as the forward pass that built the mask words;
//   cs != nullptr: the column sums of the unrounded values are added to cs[0..3] (one column per lane and chunk).
// `mw`: the mask words of this thread's four 32-column groups (column i in bit 31 - i).
template <bool FP16>
__device__ __forceinline__ bool epi_bwd_pass(Epi& e, const uint32_t (&mw)[4], int c0, int b, const Watchdog& wd, float* cs,
                                             const float* row_scale, bool wait_free, bool write_back) {
  if (!mbar_wait(e.bars + 8 * (kBarAccFull + b), (e.acc_phase >> b) & 1u, wd, kErrAccFull, b)) return false;
  e.acc_phase ^= 1u << b;
  __syncwarp();
  tc_fence_after();
  const uint32_t tbase = e.tmem_row + b * 256 + e.set * 32;
  uint32_t v[2][32];
  tmem_ld32(tbase, v[0]);
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    tmem_ld_wait();
    if (cc < 3) {
      tmem_ld32(tbase + (cc + 1) * 64, v[(cc + 1) & 1]);
    } else {
      tc_fence_before();
      __syncwarp();
      if (e.lane == 0) arrive_on_leader(e.bars + 8 * (kBarAccEmpty + b), 1);
    }
    const uint32_t(&vc)[32] = v[cc & 1];
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = (mw[cc] >> (31 - i)) & 1u ? __uint_as_float(vc[i]) : 0.f;
    if (row_scale != nullptr) {
      const float g = *row_scale;
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] *= g;
    }
    if (write_back) {
      uint32_t packed[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) packed[k] = pack_plain<FP16>(f[2 * k], f[2 * k + 1]);
      const int c = c0 + cc;
      if (wait_free) {
        if (!mbar_wait(e.bars + 8 * (kBarAFree + c), ((e.wphase >> c) & 1u) ^ 1u, wd, kErrAFree, c)) return false;
      }
      const uint32_t base = e.a_row_addr + c * kAChunkBytes;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        st_shared_v4(base + (((4 * e.set + u) ^ e.row7) << 4), packed[4 * u], packed[4 * u + 1], packed[4 * u + 2],
                     packed[4 * u + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (e.lane == 0) arrive_on_leader(e.bars + 8 * (kBarAReady + c), 1);
    }
    if (cs != nullptr) cs[cc] += warp_colsum32(f, e.lane);
  }
  if (write_back) e.wphase ^= 0xFu << c0;
  return true;
}

// h0's masks arrive as ballot words (epi_layer0): even features in E, odd ones in O, this thread's 32 columns in bits
// 16 set .. 16 set + 15 of each.  Interleave them into the MSB-first word the backward pass expects.
__device__ __forceinline__ uint32_t spread16(uint32_t x) {
  x &= 0xFFFFu;
  x = (x | (x << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;
  x = (x | (x << 2)) & 0x33333333u;
  return (x | (x << 1)) & 0x55555555u;
}
__device__ __forceinline__ uint32_t l0_mask_word(uint32_t even, uint32_t odd, int set) {
  return __brev(spread16(even >> (16 * set)) | (spread16(odd >> (16 * set)) << 1));
}

// First layer of a tile, column-mapped.  Chunks of the activation buffer are released in pairs
// by the last layer of the previous tile, so all 8 warps work on the pair that was released
// most recently: in step i warp w takes chunk 2i + (w & 1), rows [32 (w >> 1), +32); lane l owns
// features 64c + 2l, +1 of that chunk (weights in registers: `wl[i]`), the coordinates of the
// rows sit in shared memory.
struct L0Weights { float4 a, b; };

// Packed fp32 FMA (FFMA2): two independent IEEE fused multiply-adds per instruction, each lane bit-identical to fmaf.
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long dup2(float v) {
  const unsigned long long u = __float_as_uint(v);
  return u | (u << 32);
}

// `m0` (BWD kernels): where the ReLU mask of h0 goes - per chunk c two ballot words per row, m0[(2 c + par) * 128 + row],
// bit l of word `par` = (feature 64 c + 2 l + par is positive); lane r - r0 keeps row r's words and stores them.
// The coordinates sit in shared memory as three arrays sx / sy / sz [128] (so that one 64-bit load brings a coordinate of
// two consecutive rows): every packed FMA works on rows (r, r + 1) of one feature.  Same operations in the same order as
// the scalar chain fmaf(z, wz, fmaf(y, wy, fmaf(x, wx, b))) -> identical bits.  (The first layer sits on the critical path at
// every tile boundary - the next tile cannot start before its chunks exist, and they can only be written once the last
// layer of the current tile has released them - so its instruction count matters: profiles/r2_k1_pass_trace_experiments.txt.)
template <bool FP16, bool MASK = false>
__device__ __forceinline__ bool epi_layer0(Epi& e, int warp, const L0Weights (&wl)[4],
                                           const float* __restrict__ sxyz, uint32_t smem_a, const Watchdog& wd,
                                           uint32_t* m0 = nullptr) {
  const uint32_t unit = e.lane >> 2;
  const int r0 = (warp >> 1) * 32;
  const float* sx = sxyz;
  const float* sy = sxyz + kTileM;
  const float* sz = sxyz + 2 * kTileM;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = 2 * i + (warp & 1);
    if (!mbar_wait(e.bars + 8 * (kBarAFree + c), ((e.wphase >> c) & 1u) ^ 1u, wd, kErrAFree, c)) return false;
    const uint32_t chunk = smem_a + c * kAChunkBytes + ((e.lane & 3) << 2);
    const float4 wa = wl[i].a, wb = wl[i].b;
    const unsigned long long wax = dup2(wa.x), way = dup2(wa.y), waz = dup2(wa.z), waw = dup2(wa.w);
    const unsigned long long wbx = dup2(wb.x), wby = dup2(wb.y), wbz = dup2(wb.z), wbw = dup2(wb.w);
    uint32_t keep_e = 0, keep_o = 0;
#pragma unroll 4
    for (int r = r0; r < r0 + 32; r += 2) {
      const unsigned long long qx = *reinterpret_cast<const unsigned long long*>(sx + r);
      const unsigned long long qy = *reinterpret_cast<const unsigned long long*>(sy + r);
      const unsigned long long qz = *reinterpret_cast<const unsigned long long*>(sz + r);
      const unsigned long long ta = ffma2(qz, waz, ffma2(qy, way, ffma2(qx, wax, waw)));     // feature 2 l    of rows r, r + 1
      const unsigned long long tb = ffma2(qz, wbz, ffma2(qy, wby, ffma2(qx, wbx, wbw)));     // feature 2 l + 1
      const float f0a = __uint_as_float(static_cast<uint32_t>(ta)), f0b = __uint_as_float(static_cast<uint32_t>(ta >> 32));
      const float f1a = __uint_as_float(static_cast<uint32_t>(tb)), f1b = __uint_as_float(static_cast<uint32_t>(tb >> 32));
      const uint32_t va = pack_relu<FP16>(f0a, f1a), vb = pack_relu<FP16>(f0b, f1b);
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(chunk + r * 128 + ((unit ^ (r & 7)) << 4)), "r"(va) : "memory");
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(chunk + (r + 1) * 128 + ((unit ^ ((r + 1) & 7)) << 4)), "r"(vb) : "memory");
      if constexpr (MASK) {
        const uint32_t bea = __ballot_sync(0xffffffffu, f0a > 0.f), boa = __ballot_sync(0xffffffffu, f1a > 0.f);
        const uint32_t beb = __ballot_sync(0xffffffffu, f0b > 0.f), bob = __ballot_sync(0xffffffffu, f1b > 0.f);
        if (e.lane == r - r0) { keep_e = bea; keep_o = boa; }
        if (e.lane == r + 1 - r0) { keep_e = beb; keep_o = bob; }
      }
    }
    if constexpr (MASK) {
      m0[(2 * c) * kTileM + r0 + e.lane] = keep_e;
      m0[(2 * c + 1) * kTileM + r0 + e.lane] = keep_o;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (e.lane == 0) arrive_on_leader(e.bars + 8 * (kBarAReady + c), 2);   // 4 warps x 2 = the 8 expected per CTA
  }
  e.wphase ^= 0xFFu;
  return true;
}

template <bool FP16, bool BWD>
__global__ void __launch_bounds__(kThreads, 1)
fused_decoder_kernel(const DecodeParams p, const __grid_constant__ CUtensorMap tmap) {
  constexpr int kNumPasses = BWD ? kPassesBwd : kPasses;
  constexpr int kNumBlocks = BWD ? kBlocksPerTileBwd : kBlocksPerTile;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem0 = smem_u32(smem_raw);
  uint8_t* gen = smem_raw;
  if ((smem0 & 1023u) != 0) {   // the swizzled operand layout needs it; never observed, but fail loudly
    if (threadIdx.x == 0) atomicCAS(p.status, 0u, 0xA11u);
    return;
  }
  const uint32_t bars = smem0 + oBar;
  volatile uint32_t* misc = reinterpret_cast<volatile uint32_t*>(gen + oMisc);   // [0] tmem base, [1] abort
  float* sbias = reinterpret_cast<float*>(gen + oBias);
  float* shead = reinterpret_cast<float*>(gen + oHead);
  float* sxyz = reinterpret_cast<float*>(gen + oXyz);        // sx [128] | sy [128] | sz [128]
  float* sdot = reinterpret_cast<float*>(gen + oDot);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  const long long num_tiles = (p.M + 2 * kTileM - 1) / (2 * kTileM);      // pair tiles of 256 queries
  const long long npairs = gridDim.x >> 1, pidx = blockIdx.x >> 1;
  const long long my_tiles = pidx < num_tiles ? (num_tiles - pidx + npairs - 1) / npairs : 0;
  const DecConsts* __restrict__ cs = p.consts;

  if (threadIdx.x == 0) {
    misc[1] = 0;
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bars + 8 * (kBarWFull + s), 1);
      mbar_init(bars + 8 * (kBarWEmpty + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bars + 8 * (kBarAccFull + b), 1);
      mbar_init(bars + 8 * (kBarAccEmpty + b), 2 * kEpiWarps);      // every epilogue warp of both CTAs
    }
    for (int c = 0; c < kAChunks; ++c) {
      mbar_init(bars + 8 * (kBarAReady + c), 2 * kEpiWarps);        // every epilogue warp of both CTAs
      mbar_init(bars + 8 * (kBarAFree + c), 1);
    }
    fence_mbar_init();
  }
  // epilogue constants -> shared memory (the fold kernel ran earlier on this stream)
  {
    const float* gb = &cs->bias[0][0];
    for (int i = threadIdx.x; i < 7 * kHid; i += kThreads) sbias[i] = gb[i];
  }
  for (int i = threadIdx.x; i < kHid; i += kThreads) shead[i] = cs->head[i];
  if (warp == 9) {
    tmem_alloc<2>(smem0 + oMisc, 512);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = misc[0];
  long long waited[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  Watchdog wd{misc + 1, p.status, p.timeout_ns, p.prof != nullptr ? waited : nullptr, p.status_host};
  const long long t_start = clock64();

  if (warp == 8) {
    // ===================== producer: this CTA's half of every weight block =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      if (p.debug_flags & 2u) goto done;   // test hook: no weights ever arrive, so the consumers' waits must time out
      for (long long it = 0; it < my_tiles; ++it) {
#pragma unroll 1
        for (int blk = 0; blk < kNumBlocks; ++blk) {
          if (!mbar_wait(bars + 8 * (kBarWEmpty + stage), phase ^ 1u, wd, kErrWEmpty, stage)) goto done;
          const uint32_t full = bars + 8 * (kBarWFull + stage);
          if (leader) mbar_arrive_expect_tx(full, kBlockBytes);          // both halves land on this barrier
          if (!(p.debug_flags & 1u))
            tma_load_half_block(smem0 + oW + stage * kHalfBlockBytes, &tmap, blk * kBlockRows + rank * 128,
                                map_to_cta(full, 0));
          else if (leader)
            asm volatile("mbarrier.complete_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(full), "r"(kBlockBytes) : "memory");
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer (leader CTA) =====================
    // The WHOLE warp runs the loop so that every operand is warp-uniform and only the
    // tcgen05 instructions sit under elect.sync: issued from a divergent `lane == 0` branch,
    // ptxas wraps each one in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop and the issuing
    // thread itself (~440 cycles of instructions per 512-cycle block) becomes the bottleneck.
    if (leader) {
      constexpr uint32_t idesc = umma_idesc(256, 256, FP16 ? 0 : 1);
      const uint64_t desc_hi = umma_desc_sw128(0) & 0xFFFFFFFF00000000ull;
      const uint32_t a_lo0 = ((smem0 + oA) & 0x3FFFFu) >> 4 | (1u << 16);
      const uint32_t w_lo0 = ((smem0 + oW) & 0x3FFFFu) >> 4 | (1u << 16);
      uint32_t stage = 0, phase = 0, rphase = 0, ephase = 0, gpass = 0, prev_stage = 0;
      for (long long it = 0; it < my_tiles; ++it) {
#pragma unroll 1
        for (int ps = 0; ps < kNumPasses; ++ps, ++gpass) {
          const int nk = pass_chunks<BWD>(ps);
          const bool first = pass_first<BWD>(ps), last = pass_last<BWD>(ps);
          const uint32_t b = gpass & 1u;
          const uint32_t d_tmem = tmem_base + b * 256;
          if (!mbar_wait(bars + 8 * (kBarAccEmpty + b), ((ephase >> b) & 1u) ^ 1u, wd, kErrAccEmpty, b)) goto done;
          ephase ^= 1u << b;
          SDFB_K1_TRACE(0, ps);
#pragma unroll 1
          for (int k = 0; k < nk; ++k) {
            if (first) {
              if (!mbar_wait(bars + 8 * (kBarAReady + k), (rphase >> k) & 1u, wd, kErrAReady, k)) goto done;
              rphase ^= 1u << k;
            }
            if (!mbar_wait(bars + 8 * (kBarWFull + stage), phase, wd, kErrWFull, stage)) goto done;
            tc_fence_after();
            const uint64_t adesc = desc_hi | (a_lo0 + k * (kAChunkBytes >> 4));
            const uint64_t bdesc = desc_hi | (w_lo0 + stage * (kHalfBlockBytes >> 4));
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_ss<2>(d_tmem, adesc + 2 * j, bdesc + 2 * j, idesc, (k | j) != 0 ? 1u : 0u);
              if (k & 1) {   // one commit point per pair of blocks (see the header comment)
                umma_commit<2>(bars + 8 * (kBarWEmpty + prev_stage));
                umma_commit<2>(bars + 8 * (kBarWEmpty + stage));
                if (last) {
                  umma_commit<2>(bars + 8 * (kBarAFree + k - 1));
                  umma_commit<2>(bars + 8 * (kBarAFree + k));
                }
                if (k == nk - 1) umma_commit<2>(bars + 8 * (kBarAccFull + b));
              }
            }
            __syncwarp();
            prev_stage = stage;
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          SDFB_K1_TRACE(1, ps);
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    Epi e;
    e.bars = bars;
    e.lane = lane;
    e.set = warp >> 2;
    const int row = (warp & 3) * 32 + lane;               // row of this CTA's 128 == TMEM lane
    e.tmem_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    e.a_row_addr = smem0 + oA + row * 128;
    e.row7 = row & 7u;
    e.wphase = 0;
    e.acc_phase = 0;
    e.trace = nullptr;
    e.dbg = p.debug_flags;
    L0Weights wl[4];                                      // layer-0 features of this lane, per step
    auto load_l0_weights = [&](L0Weights (&w)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n0 = (2 * i + (warp & 1)) * 64 + 2 * lane;
        w[i].a = cs->l0[n0];
        w[i].b = cs->l0[n0 + 1];
      }
    };
    if constexpr (!BWD) load_l0_weights(wl);              // BWD: reloaded per tile (registers are scarcer there)
    // BWD: h0's ReLU masks, double-buffered by tile parity (the next tile's first layer runs before this tile's last pass)
    uint32_t* const m0_base = BWD ? p.mask_scratch + static_cast<size_t>(blockIdx.x) * (8 * 16 * kTileM) + 6 * 16 * kTileM : nullptr;
    const float head_b = cs->head_b[0];
    const long long tile_stride = npairs * 2 * kTileM;
    long long row_base = pidx * 2 * kTileM + rank * kTileM;   // first query of this CTA's half tile
    uint32_t gpass = 0;
    float cs0[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};   // BWD: column sums of delta0 / delta4 (lane = column)
    float cs4[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float loss_acc_total = 0.f;
    if (my_tiles > 0) {
      float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
      Query q{0.f, 0.f, 0.f};
      if constexpr (!BWD) {
      if (e.set == 0) {
        const Query q0 = load_query(p, row_base + row);
        sxyz[row] = q0.x; sxyz[kTileM + row] = q0.y; sxyz[2 * kTileM + row] = q0.z;
      }
      named_bar_sync(1, kEpiThreads);
      if (!epi_layer0<FP16>(e, warp, wl, sxyz, smem0 + oA, wd)) goto done;
      qv = make_float4(sxyz[row], sxyz[kTileM + row], sxyz[2 * kTileM + row], 0.f);
      q = Query{qv.x, qv.y, qv.z};
      for (long long it = 0; it < my_tiles; ++it, row_base += tile_stride) {
        const bool dump_tile = p.dump != nullptr && row_base == 0;
#pragma unroll 1
        for (int ps = 0; ps < 11; ++ps, ++gpass) {
          int layer, half;
          if (ps < 4) { layer = 1 + (ps >> 1); half = ps & 1; }
          else if (ps == 4) { layer = 3; half = 0; }
          else { layer = 4 + ((ps - 5) >> 1); half = (ps - 5) & 1; }
          const float* bias = sbias + (layer - 1) * kHid + half * 256;
          float* dump_row = (dump_tile && ps == p.dump_pass) ? p.dump + row * 256 : nullptr;
          e.trace = (p.prof != nullptr && blockIdx.x == 0 && it == 5 && warp == 0) ? p.prof + static_cast<size_t>(gridDim.x) * 24 + 64 + ps : nullptr;
          bool ok;
          if (layer == 3)
            ok = epi_hidden_pass<FP16, false>(e, bias, q, 0, gpass & 1u, wd, dump_row, Flag<true>{}, Flag<false>{});
          else if (half == 0)
            ok = epi_hidden_pass<FP16, false>(e, bias, q, 0, gpass & 1u, wd, dump_row, Flag<false>{}, Flag<true>{});
          else
            ok = epi_hidden_pass<FP16, false>(e, bias, q, 4, gpass & 1u, wd, dump_row, Flag<false>{}, Flag<false>{});
          if (!ok) goto done;
          if (warp == 0) SDFB_K1_TRACE(3, ps);
        }
        float dot = 0.f;
        float* dump_row = (dump_tile && p.dump_pass == 11) ? p.dump + row * 256 : nullptr;
        if (e.trace != nullptr) e.trace = p.prof + static_cast<size_t>(gridDim.x) * 24 + 64 + 11;
        if (!epi_head_pass(e, sbias + 6 * kHid, shead, gpass & 1u, wd, dot, dump_row)) goto done;
        ++gpass;
        // first layer of the next tile, written behind the last readers of this tile's h6
        if (it + 1 < my_tiles) {
          if (e.set == 0) {
            const Query qn = load_query(p, row_base + tile_stride + row);
            sxyz[row] = qn.x; sxyz[kTileM + row] = qn.y; sxyz[2 * kTileM + row] = qn.z;
          }
          named_bar_sync(1, kEpiThreads);
          if (!epi_layer0<FP16>(e, warp, wl, sxyz, smem0 + oA, wd)) goto done;
          qv = make_float4(sxyz[row], sxyz[kTileM + row], sxyz[2 * kTileM + row], 0.f);
        }
        dump_row = (dump_tile && p.dump_pass == 12) ? p.dump + row * 256 : nullptr;
        if (e.trace != nullptr) e.trace = p.prof + static_cast<size_t>(gridDim.x) * 24 + 64 + 12;
        if (!epi_head_pass(e, sbias + 6 * kHid + 256, shead + 256, gpass & 1u, wd, dot, dump_row)) goto done;
        e.trace = nullptr;
        ++gpass;
        if (e.set == 1) sdot[row] = dot;
        named_bar_sync(2, kEpiThreads);
        if (e.set == 0) {
          const long long m = row_base + row;
          const float v = tanhf((dot + sdot[row]) + head_b);
          if (m < p.M) p.out[m] = v;
          if (p.signs != nullptr) {   // inside(v) := v < 0 of the value just stored, 32 queries per word (bit = query & 31)
            const unsigned int bits = __ballot_sync(0xffffffffu, m < p.M && v < 0.f);
            if (lane == 0 && m < p.M) p.signs[m >> 5] = bits;
          }
        }
        q = Query{qv.x, qv.y, qv.z};
      }
      } else {
        // ============ forward + backward (latent gradient) ============
        // One call site per epilogue function (code size, see epi_hidden_pass).  Iteration -1 only computes the
        // first layer of the first tile; iteration `it` ends with the first layer of tile it + 1, placed before
        // its own last pass exactly like the forward kernel's.
        uint32_t* mbase = p.mask_scratch + static_cast<size_t>(blockIdx.x) * (8 * 16 * kTileM) + row;
        const float up_scale = p.target != nullptr ? 1.f : ldexpf(1.f, -vjp_scale_exponent(__uint_as_float(__ldg(p.dLdy_amax))));
        float loss_acc = 0.f;                        // loss mode: this thread's share of sum |clamp(sdf) - clamp(target)|
#pragma unroll 1
        for (long long it = -1; it < my_tiles; ++it) {
          if (it >= 0) {
#pragma unroll 1
            for (int ps = 0; ps < 11; ++ps, ++gpass) {
              int layer, half;
              if (ps < 4) { layer = 1 + (ps >> 1); half = ps & 1; }
              else if (ps == 4) { layer = 3; half = 0; }
              else { layer = 4 + ((ps - 5) >> 1); half = (ps - 5) & 1; }
              const float* bias = sbias + (layer - 1) * kHid + half * 256;
              uint32_t* mrow = mbase + ((layer - 1) * 16 + half * 8) * kTileM;
              if (!epi_hidden_pass<FP16, true>(e, bias, q, half * 4, gpass & 1u, wd, nullptr, layer == 3,
                                               half == 0 && layer != 3, mrow))
                goto done;
            }
            float dot = 0.f;
#pragma unroll 1
            for (int hb = 0; hb < 2; ++hb, ++gpass) {   // head halves; each leaves its half of the first backward operand behind
              uint32_t hm[4];
              if (!epi_head_pass<true>(e, sbias + 6 * kHid + 256 * hb, shead + 256 * hb, gpass & 1u, wd, dot, nullptr, hm)) goto done;
              if (!epi_delta7_half<FP16>(e, shead, hm, hb, wd)) goto done;
            }
            e.wphase ^= 0xFFu;
            if (e.set == 1) sdot[row] = dot;
            named_bar_sync(2, kEpiThreads);
            if (e.set == 0) {
              const long long m = row_base + row;
              const float v = tanhf((dot + sdot[row]) + head_b);
              float gsv = 0.f;
              if (m < p.M) {
                if (p.out != nullptr) p.out[m] = v;
                float up;
                if (p.target != nullptr) {           // clamped-L1 fitting loss, formed here (see kernels.h)
                  const float c = p.clamp;
                  const float diff = fminf(fmaxf(v, -c), c) - fminf(fmaxf(__ldg(p.target + m), -c), c);
                  loss_acc += fabsf(diff);
                  up = (v > -c && v < c) ? (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)) : 0.f;
                } else {
                  up = __ldg(p.dLdy + m) * up_scale;
                }
                gsv = up * (1.f - v * v);
              }
              sdot[row] = gsv;                                     // rows past M carry no gradient
            }
            named_bar_sync(2, kEpiThreads);
          }
          const float gs = sdot[row];
#pragma unroll 1
          for (int ps = 13; ps < kPassesBwd; ++ps) {
            if (ps == 25 && it + 1 < my_tiles) {   // first layer of the next tile, behind the last readers of delta1
              const long long next_base = it < 0 ? row_base : row_base + tile_stride;
              if (e.set == 0) {
                const Query qn = load_query(p, next_base + row);
                sxyz[row] = qn.x; sxyz[kTileM + row] = qn.y; sxyz[2 * kTileM + row] = qn.z;
              }
              named_bar_sync(1, kEpiThreads);
              load_l0_weights(wl);
              if (!epi_layer0<FP16, true>(e, warp, wl, sxyz, smem0 + oA, wd, m0_base + ((it + 1) & 1) * (16 * kTileM))) goto done;
              qv = make_float4(sxyz[row], sxyz[kTileM + row], sxyz[2 * kTileM + row], 0.f);
            }
            if (it < 0) continue;
            const uint32_t b = gpass & 1u;
            int layer, half;                       // the layer whose delta this pass produces
            if (ps < 19) { layer = 6 - ((ps - 13) >> 1); half = (ps - 13) & 1; }
            else if (ps == 19) { layer = 3; half = 0; }
            else { layer = 2 - ((ps - 20) >> 1); half = (ps - 20) & 1; }
            uint32_t mw[4];
            if (layer > 0) {
              const uint32_t* mrow = mbase + ((layer - 1) * 16 + half * 8) * kTileM;
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) mw[cc] = mrow[(2 * cc + e.set) * kTileM];
            } else {
              const uint32_t* m0 = m0_base + (it & 1) * (16 * kTileM) + half * (8 * kTileM) + row;
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) mw[cc] = l0_mask_word(m0[(2 * cc) * kTileM], m0[(2 * cc + 1) * kTileM], e.set);
            }
            float t[4] = {0.f, 0.f, 0.f, 0.f};
            const bool sums = layer == 4 || layer == 0;
            if (!epi_bwd_pass<FP16>(e, mw, half * 4, b, wd, sums ? t : nullptr, layer == 6 ? &gs : nullptr,
                                    half == 0 && layer != 3, layer != 0))
              goto done;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (layer == 4) { if (half == 0) cs4[0][i] += t[i]; else cs4[1][i] += t[i]; }
              if (layer == 0) { if (half == 0) cs0[0][i] += t[i]; else cs0[1][i] += t[i]; }
            }
            ++gpass;
          }
          if (it >= 0) row_base += tile_stride;
          q = Query{qv.x, qv.y, qv.z};
        }
        loss_acc_total = loss_acc;
      }
    }
    if constexpr (BWD) {
      if (p.loss_partial != nullptr && e.set == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) loss_acc_total += __shfl_xor_sync(0xffffffffu, loss_acc_total, o);
        if (lane == 0) p.loss_partial[blockIdx.x * 4 + (warp & 3)] = loss_acc_total;
      }
    }
    if constexpr (BWD) {   // this warp's share of the column sums: [CTA][quadrant][delta0 512 | delta4 512]
      float* dst = p.colsum + (static_cast<size_t>(blockIdx.x) * 4 + (warp & 3)) * 1024 + e.set * 32 + lane;
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          dst[256 * h + 64 * cc] = cs0[h][cc];
          dst[512 + 256 * h + 64 * cc] = cs4[h][cc];
        }
    }
  }
done:
  if (p.prof != nullptr && lane == 0 && (warp == 0 || warp >= 8) && (leader || warp != 9)) {
    const int role = warp == 0 ? 0 : warp - 7;            // 0 epilogue, 1 producer, 2 MMA issuer
    long long* dst = p.prof + (static_cast<long long>(blockIdx.x) * 3 + role) * 8;
#pragma unroll
    for (int i = 1; i < 7; ++i) dst[i] = waited[i];
    dst[0] = clock64() - t_start;
    dst[7] = my_tiles;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's shared memory and barriers stay valid until both are done
  if (warp == 9) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<2>(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

cudaError_t fused_decoder_init() {
  const void* fns[4] = {reinterpret_cast<const void*>(fused_decoder_kernel<false, false>),
                        reinterpret_cast<const void*>(fused_decoder_kernel<true, false>),
                        reinterpret_cast<const void*>(fused_decoder_kernel<false, true>),
                        reinterpret_cast<const void*>(fused_decoder_kernel<true, true>)};
  for (const void* f : fns) {
    cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemAlloc));
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// The weight stream (forward blocks, then the backward ones) viewed as a [192*256 rows][64] 16-bit matrix;
// box = 128 rows x 64 = one CTA's half of a block.  The stream already holds swizzled shared-memory images, so no TMA swizzle.
cudaError_t make_wstream_tensor_map(const void* wstream, void* tmap_out) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess) return e;
  if (fn == nullptr || qres != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(kChunkK), static_cast<cuuint64_t>(kBlocksPerTileBwd) * kBlockRows};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(kChunkK) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kChunkK), 128u};
  const cuuint32_t estr[2] = {1u, 1u};
  CUresult r = reinterpret_cast<EncodeTiledFn>(fn)(static_cast<CUtensorMap*>(tmap_out), CU_TENSOR_MAP_DATA_TYPE_UINT16, 2,
                                                   const_cast<void*>(wstream), gdim, gstride, box, estr,
                                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t launch_fused_decoder(const DecodeParams& p, const void* tmap, bool fp16, int num_sms,
                                  cudaStream_t stream) {
  if (p.M <= 0) return cudaSuccess;
  const long long tiles = (p.M + 2 * kTileM - 1) / (2 * kTileM);
  const long long pairs = tiles < num_sms / 2 ? tiles : num_sms / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(2 * pairs));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemAlloc;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const CUtensorMap* tm = static_cast<const CUtensorMap*>(tmap);
  if (p.bwd) {   // forward + backward: needs the mask scratch and the column-sum rows of every CTA
    if (p.mask_scratch == nullptr || p.colsum == nullptr) return cudaErrorInvalidValue;
    if (p.target != nullptr ? p.loss_partial == nullptr : (p.dLdy == nullptr || p.dLdy_amax == nullptr)) return cudaErrorInvalidValue;
    if (fp16) return cudaLaunchKernelEx(&cfg, fused_decoder_kernel<true, true>, p, *tm);
    return cudaLaunchKernelEx(&cfg, fused_decoder_kernel<false, true>, p, *tm);
  }
  if (fp16) return cudaLaunchKernelEx(&cfg, fused_decoder_kernel<true, false>, p, *tm);
  return cudaLaunchKernelEx(&cfg, fused_decoder_kernel<false, false>, p, *tm);
}

}  // namespace sdfb

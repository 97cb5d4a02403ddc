// Thin inline-PTX wrappers for the sm_100a features the kernels use:
// mbarrier, 1-D bulk async copy (TMA unit), tcgen05 (TMEM alloc, UMMA, commit,
// TMEM load, fences) and thread-block-cluster helpers.
//
// Every blocking wait is bounded by a wall-clock watchdog: a protocol bug then
// surfaces as an error code in the context's status word instead of a hung GPU.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sdfb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred px;\n"
      "elect.sync _|px, 0xFFFFFFFF;\n"
      "selp.u32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------- cluster ---
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}

// --------------------------------------------------------------- mbarrier ---
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on a barrier addressed in the shared::cluster window (own or peer CTA)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
// CTA-scope acquire: a cluster-scope acquire makes ptxas emit CCTL.IVALL (an L1 invalidate)
// after every wait, which turns the epilogue's bias loads into L2 round trips.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// Watchdog state shared by all roles of a CTA.
struct Watchdog {
  volatile uint32_t* abort_flag;   // smem word: non-zero => give up everywhere
  unsigned int* status;            // global word: first failure code wins
  uint64_t timeout_ns;
  long long* wait_cycles;          // optional per-thread [8] accumulator of blocked cycles per site class
  unsigned int* status_host;       // optional mirror of `status` in mapped pinned host memory: the host sees a trip at its
                                   // next call on the context without synchronising (plain store; any failing site's code will do)
};

// Records the first failure code in the context's status word and mirrors it to the host.
__device__ __forceinline__ void wd_trip(const Watchdog& wd, uint32_t code) {
  atomicCAS(wd.status, 0u, code);
  if (wd.status_host != nullptr) {
    *reinterpret_cast<volatile unsigned int*>(wd.status_host) = code;
    __threadfence_system();
  }
}

// Bounded wait.  `site` (a multiple of 16) + `idx` identify the wait site in the status word;
// site >> 4 is the class under which blocked cycles are accumulated when profiling.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, const Watchdog& wd,
                                          uint32_t site, uint32_t idx = 0) {
  const long long c0 = wd.wait_cycles != nullptr ? clock64() : 0;
  if (mbar_try_wait(bar, parity)) {   // try_wait itself may block for a while: count that time too
    if (wd.wait_cycles != nullptr) wd.wait_cycles[(site >> 4) & 7u] += clock64() - c0;
    return true;
  }
  uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (true) {
    if (mbar_try_wait(bar, parity)) {
      if (wd.wait_cycles != nullptr) wd.wait_cycles[(site >> 4) & 7u] += clock64() - c0;
      return true;
    }
    if ((++spins & 0xFFu) == 0) {
      if (*wd.abort_flag) return false;
      if (global_timer_ns() - t0 > wd.timeout_ns) {
        *wd.abort_flag = site + idx;
        wd_trip(wd, site + idx);
        return false;
      }
    }
  }
}

// one arrival (count `count`) on the barrier at the same offset in CTA `rank` of the cluster; ordering is supplied by
// the caller's fence_acq_rel_cluster()
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t local_bar, uint32_t rank, uint32_t count) {
  const uint32_t remote = map_to_cta(local_bar, rank);
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(remote), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }

// ------------------------------------------------------------ async proxy ---
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 1-D bulk copy global -> this CTA's shared memory, completion on a local mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// ----------------------------------------------------------------- tcgen05 ---
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
                 "r"(ncols)
                 : "memory");
  else
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
                 "r"(ncols)
                 : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_relinquish() {
  if constexpr (CG == 1)
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  else
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle:
// rows are 128 B (64 x 16-bit) apart, 8-row groups 1024 B apart (SBO), the
// 16-byte chunks of a row XOR-swizzled with (row & 7).  Field layout follows
// the sm_100 descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type [61,64) (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // LBO (ignored for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO
  d |= static_cast<uint64_t>(1) << 46;            // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16: fp32 accumulate, A/B both K-major.
// ab_format: 0 = fp16, 1 = bf16.
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, int ab_format) {
  return (1u << 4)                                  // D format: f32
         | (static_cast<uint32_t>(ab_format) << 7)  // A format
         | (static_cast<uint32_t>(ab_format) << 10) // B format
         | (static_cast<uint32_t>(N >> 3) << 17)    // N
         | (static_cast<uint32_t>(M >> 4) << 24);   // M
}

// D[tmem] (+)= A[smem desc] * B[smem desc]^T ; issued by ONE thread.
template <int CG>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                        uint32_t idesc, uint32_t accumulate) {
  if constexpr (CG == 1)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// All previously issued MMAs of this thread arrive on `bar` when complete.
// CG==2: the arrive is multicast to the same barrier offset in both CTAs of the pair.
// `pair_mask`: the cluster ranks of the pair's two CTAs (3 for a cluster that is just the pair).
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar, uint16_t pair_mask = 3) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
  else
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
        "[%0], %1;" ::"r"(bar),
        "h"(pair_mask)
        : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// TMEM -> registers: 32 lanes x 16 consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// registers -> TMEM: this warp's 32 lanes x 16 consecutive 32-bit columns (used by the TMEM-operand probe, umma_rate.cu)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]^T : the A operand read from tensor memory (128 lanes x K/2 packed columns per CTA)
template <int CG>
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (CG == 1)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- 16-bit packing ------------------------------------------------------------------------
template <bool FP16>
__device__ __forceinline__ uint32_t pack_relu(float lo, float hi) {
  uint32_t d;
  if constexpr (FP16)
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// ---- cluster-scope barrier helpers ---------------------------------------------------------
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// wait used by the MMA issuer: the arrivals come from both CTAs of the pair
__device__ __forceinline__ bool mbar_wait_cluster(uint32_t bar, uint32_t parity, const Watchdog& wd,
                                                  uint32_t site, uint32_t idx = 0) {
  const long long c0 = wd.wait_cycles != nullptr ? clock64() : 0;
  if (mbar_try_wait_cluster(bar, parity)) {   // try_wait itself may block for a while: count that time too
    if (wd.wait_cycles != nullptr) wd.wait_cycles[(site >> 4) & 7u] += clock64() - c0;
    return true;
  }
  uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (true) {
    if (mbar_try_wait_cluster(bar, parity)) {
      if (wd.wait_cycles != nullptr) wd.wait_cycles[(site >> 4) & 7u] += clock64() - c0;
      return true;
    }
    if ((++spins & 0xFFu) == 0) {
      if (*wd.abort_flag) return false;
      if (global_timer_ns() - t0 > wd.timeout_ns) {
        *wd.abort_flag = site + idx;
        wd_trip(wd, site + idx);
        return false;
      }
    }
  }
}
// `count` arrivals on the LEADER CTA's copy of a barrier (local or remote).  Default semantics
// (release at CTA scope), as CUTLASS's ClusterBarrier::arrive(cta_id) does: a cluster-scope release
// costs a MEMBAR of several hundred cycles per arrival (30% of all epilogue stall samples when it
// was tried).  The data being published was made visible to the async proxy by each lane's
// fence.proxy.async and ordered before this arrive by __syncwarp().
__device__ __forceinline__ void arrive_on_leader(uint32_t local_bar, uint32_t count, uint32_t leader_rank = 0) {
  const uint32_t remote = map_to_cta(local_bar, leader_rank);
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0], %1;" ::"r"(remote), "r"(count) : "memory");
}
// Rows [row0, row0 + box rows) of a [rows][64] 16-bit tensor map -> local shared memory; the
// transaction bytes are counted on the LEADER CTA's barrier (.cta_group::2).  (The fused decoder
// loads its half of a weight block with this.)
__device__ __forceinline__ void tma_load_half_block(uint32_t dst_smem, const void* tmap, int row0,
                                                    uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(0), "r"(row0), "r"(leader_bar)
      : "memory");
}

template <bool FP16>
__device__ __forceinline__ uint32_t pack_plain(float lo, float hi) {   // no ReLU: signed coordinates
  uint32_t d;
  if constexpr (FP16)
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}


__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                             uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}

// 256-bit global store / load (sm_100: one instruction per 32-byte sector; the address must be 32-byte aligned)
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
               "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_nc_v8(const void* p, uint32_t* v) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace sdfb

// Hierarchical sparse decode: the field is decoded only where the surface can be, in two
// levels, and what comes out is what the DENSE marching-cubes kernels consume - a res^3 fp32 array that is valid at every
// node of every cell the surface crosses, and the COMPLETE sign bit-planes of the grid (1 bit per node).  No upstream
// source exists (/root/reference/README.md:1); checked against the dense extraction (identical triangle soup).
//
//   level 1: corners of B1^3-cell blocks (B1 = 8)  ->  keep a block if its corners differ in sign or come within
//            tau1 = L1 B1 h sqrt(3)/2 of zero (exact for an L1-Lipschitz field; h = 2 / (res - 1))
//   level 2: inside the kept blocks, the corners of B2^3-cell sub-blocks (B2 = 2; one shared lattice, every node decoded
//            once)  ->  the same test with tau2 = L2 B2 h sqrt(3)/2, L2 estimated on that finer lattice  ->  decode the
//            remaining nodes of the kept sub-blocks (again each node once: the sub-blocks mark a node bitmap, the
//            bitmap is compacted into a query list)
//   signs  : a node that was never decoded lies only in discarded (sub-)blocks, on each of which the sign is constant:
//            it inherits the sign of the lower corner of its sub-block if that corner was decoded, else of its level-1
//            block's corner.
// All of this is HBM-bound integer / byte work (bitmaps, a hand-written three-phase scan for the compaction); the queries
// themselves go through the fused decoder kernel in points mode.
#include "kernels.h"

namespace sdfb {

namespace {

__device__ __forceinline__ int clamp_node(int v, int res) { return v < res - 1 ? v : res - 1; }

__device__ __forceinline__ void set_bit(unsigned int* __restrict__ bits, long long node) {
  atomicOr(bits + (node >> 5), 1u << (node & 31));
}
__device__ __forceinline__ bool get_bit(const unsigned int* __restrict__ bits, long long node) {
  return (bits[node >> 5] >> (node & 31)) & 1u;
}

// node indices of the (nb + 1)^3 level-1 corners (x fastest), clamped to the grid's last node
__global__ void corner_nodes_kernel(int res, int B, int nb, unsigned int* __restrict__ idx) {
  const long long n1 = nb + 1, total = n1 * n1 * n1;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int bx = static_cast<int>(i % n1), by = static_cast<int>((i / n1) % n1), bz = static_cast<int>(i / (n1 * n1));
  const long long x = clamp_node(bx * B, res), y = clamp_node(by * B, res), z = clamp_node(bz * B, res);
  idx[i] = static_cast<unsigned int>((z * res + y) * res + x);
}

// xyz [n][3] of listed nodes (rule A1, bit-exact)
__global__ void node_points_kernel(int res, const unsigned int* __restrict__ idx, long long n, float* __restrict__ xyz) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned int g = idx[i];
  const unsigned int t = g / res;
  const int ix = static_cast<int>(g - t * res), iz = static_cast<int>(t / res), iy = static_cast<int>(t - iz * res);
  const float den = static_cast<float>(res - 1);
  xyz[3 * i] = __fdiv_rn(axis_coord_num(ix, res), den);
  xyz[3 * i + 1] = __fdiv_rn(axis_coord_num(iy, res), den);
  xyz[3 * i + 2] = __fdiv_rn(axis_coord_num(iz, res), den);
}

__global__ void scatter_kernel(const unsigned int* __restrict__ idx, const float* __restrict__ vals, long long n,
                               float* __restrict__ dense) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dense[idx[i]] = vals[i];
}

// max over lattice edges of |f(a) - f(b)| / (nodes apart) -> atomicMax on the bits of a non-negative float
__device__ __forceinline__ void edge_max(float a, float b, int nodes_apart, float& m) {
  if (nodes_apart > 0) m = fmaxf(m, fabsf(a - b) / static_cast<float>(nodes_apart));
}
__device__ __forceinline__ void publish_max(float m, unsigned int* out) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

// level 1: largest difference quotient (per node spacing) along the +x / +y / +z lattice edges
__global__ void corner_lipschitz_kernel(const float* __restrict__ cs, int res, int B, int nb, unsigned int* __restrict__ out) {
  const long long n1 = nb + 1, total = n1 * n1 * n1;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  float m = 0.f;
  if (i < total) {
    const int bx = static_cast<int>(i % n1), by = static_cast<int>((i / n1) % n1), bz = static_cast<int>(i / (n1 * n1));
    const float v = cs[i];
    if (bx < nb) edge_max(v, cs[i + 1], clamp_node((bx + 1) * B, res) - clamp_node(bx * B, res), m);
    if (by < nb) edge_max(v, cs[i + n1], clamp_node((by + 1) * B, res) - clamp_node(by * B, res), m);
    if (bz < nb) edge_max(v, cs[i + n1 * n1], clamp_node((bz + 1) * B, res) - clamp_node(bz * B, res), m);
  }
  publish_max(m, out);
}

struct SubGeom {
  int res, B1, B2, nb1, per;        // per = B1 / B2 sub-blocks per block and axis
};

// sub-block `s` (0 .. per^3) of level-1 block `id`: lower corner node and extent (0 = the sub-block does not exist)
__device__ __forceinline__ bool sub_block(const SubGeom& g, int id, int s, int (&lo)[3], int (&ext)[3]) {
  const int b[3] = {id % g.nb1, (id / g.nb1) % g.nb1, id / (g.nb1 * g.nb1)};
  const int k[3] = {s % g.per, (s / g.per) % g.per, s / (g.per * g.per)};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = b[a] * g.B1 + k[a] * g.B2;
    if (lo[a] >= g.res - 1) return false;
    const int hi = clamp_node(lo[a] + g.B2, g.res);
    ext[a] = hi - lo[a];
  }
  return true;
}

// level 2, step 1: mark the 8 corners of every sub-block of every kept block
__global__ void mark_sub_corners_kernel(SubGeom g, const int* __restrict__ blocks, long long nblk, unsigned int* __restrict__ need) {
  const long long per3 = static_cast<long long>(g.per) * g.per * g.per;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nblk * per3) return;
  int lo[3], ext[3];
  if (!sub_block(g, blocks[i / per3], static_cast<int>(i % per3), lo, ext)) return;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const long long x = lo[0] + (c & 1) * ext[0], y = lo[1] + ((c >> 1) & 1) * ext[1], z = lo[2] + (c >> 2) * ext[2];
    set_bit(need, (z * g.res + y) * g.res + x);
  }
}

// level 2, step 2: largest difference quotient along the 12 edges of every candidate sub-block
// (per_block, optional: the same maximum per level-1 block - the candidates of a block are consecutive)
__global__ void sub_lipschitz_kernel(SubGeom g, const int* __restrict__ blocks, long long nblk, const float* __restrict__ dense,
                                     unsigned int* __restrict__ out, unsigned int* __restrict__ per_block) {
  const long long per3 = static_cast<long long>(g.per) * g.per * g.per;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  float m = 0.f;
  int lo[3], ext[3];
  if (i < nblk * per3 && sub_block(g, blocks[i / per3], static_cast<int>(i % per3), lo, ext)) {
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const long long x = lo[0] + (c & 1) * ext[0], y = lo[1] + ((c >> 1) & 1) * ext[1], z = lo[2] + (c >> 2) * ext[2];
      v[c] = dense[(z * g.res + y) * g.res + x];
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (!(c & 1)) edge_max(v[c], v[c | 1], ext[0], m);
      if (!(c & 2)) edge_max(v[c], v[c | 2], ext[1], m);
      if (!(c & 4)) edge_max(v[c], v[c | 4], ext[2], m);
    }
    if (per_block != nullptr && m > 0.f) atomicMax(per_block + i / per3, __float_as_uint(m));
  }
  publish_max(m, out);
}

// level 2, step 3: keep a sub-block if its corners differ in sign or come within tau (per node spacing: tau_nodes = L2'
// B2 sqrt(3)/2 with L2' in field units per node) of zero; a kept sub-block marks all its nodes.
__global__ void select_sub_blocks_kernel(SubGeom g, const int* __restrict__ blocks, long long nblk, const float* __restrict__ dense,
                                         const unsigned int* __restrict__ lip_bits, float lip_given_per_node, float safety,
                                         const unsigned int* __restrict__ per_block, float local_floor,
                                         unsigned int* __restrict__ need, unsigned long long* __restrict__ kept) {
  const long long per3 = static_cast<long long>(g.per) * g.per * g.per;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nblk * per3) return;
  int lo[3], ext[3];
  if (!sub_block(g, blocks[i / per3], static_cast<int>(i % per3), lo, ext)) return;
  float lip;
  if (lip_given_per_node > 0.f) {
    lip = lip_given_per_node;
  } else {
    const float global = __uint_as_float(max(lip_bits[0], lip_bits[1]));
    // local mode: the block's own largest quotient, but never less than `local_floor` of the global one
    lip = safety * (per_block != nullptr ? fmaxf(__uint_as_float(per_block[i / per3]), local_floor * global) : global);
  }
  const float half_diag = 0.5f * sqrtf(static_cast<float>(ext[0] * ext[0] + ext[1] * ext[1] + ext[2] * ext[2]));
  const float tau = lip * half_diag;
  int n_in = 0;
  float amin = 3.0e38f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const long long x = lo[0] + (c & 1) * ext[0], y = lo[1] + ((c >> 1) & 1) * ext[1], z = lo[2] + (c >> 2) * ext[2];
    const float v = dense[(z * g.res + y) * g.res + x];
    n_in += v < 0.f;
    amin = fminf(amin, fabsf(v));
  }
  if (!((n_in != 0 && n_in != 8) || !(amin > tau))) return;       // NaN counts as "keep"
  atomicAdd(kept, 1ull);
  for (int dz = 0; dz <= ext[2]; ++dz)
    for (int dy = 0; dy <= ext[1]; ++dy)
      for (int dx = 0; dx <= ext[0]; ++dx)
        set_bit(need, (static_cast<long long>(lo[2] + dz) * g.res + lo[1] + dy) * g.res + lo[0] + dx);
}

// level 1 selection as a bitmap over the nb^3 blocks (bit = keep): corners differ in sign or come within tau of zero
__global__ void select_blocks_bits_kernel(const float* __restrict__ cs, int nb, float tau, unsigned int* __restrict__ keep) {
  const long long total = static_cast<long long>(nb) * nb * nb;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int bx = static_cast<int>(i % nb), by = static_cast<int>((i / nb) % nb), bz = static_cast<int>(i / (static_cast<long long>(nb) * nb));
  const long long n1 = nb + 1;
  int n_in = 0;
  float amin = 3.0e38f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float v = cs[((bz + (k >> 2)) * n1 + by + ((k >> 1) & 1)) * n1 + bx + (k & 1)];
    n_in += v < 0.f;
    amin = fminf(amin, fabsf(v));
  }
  if ((n_in != 0 && n_in != 8) || !(amin > tau)) set_bit(keep, i);     // NaN counts as "keep"
}

// need2 &= ~need1 (nodes already decoded are not decoded again)
__global__ void andnot_kernel(unsigned int* __restrict__ a, const unsigned int* __restrict__ b, long long words) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < words) a[i] &= ~b[i];
}

// ---- compaction of a node bitmap into an ascending index list: three-phase scan over the words' popcounts ----
constexpr int kScanThreads = 256, kScanPerThread = 8, kScanTile = kScanThreads * kScanPerThread;   // 2048 words per block

__global__ void __launch_bounds__(kScanThreads) bitmap_tile_counts_kernel(const unsigned int* __restrict__ bits, long long words,
                                                                          unsigned int* __restrict__ tile_sums) {
  __shared__ unsigned int warp_sums[kScanThreads / 32];
  const long long base = static_cast<long long>(blockIdx.x) * kScanTile;
  unsigned int n = 0;
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k) {
    const long long w = base + k * kScanThreads + threadIdx.x;
    if (w < words) n += __popc(bits[w]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int s = 0;
    for (int i = 0; i < kScanThreads / 32; ++i) s += warp_sums[i];
    tile_sums[blockIdx.x] = s;
  }
}

// exclusive scan of the tile sums in place (one block; tiles <= 1 << 20), total -> tile_sums[tiles]
__global__ void __launch_bounds__(1024) scan_tile_sums_kernel(unsigned int* __restrict__ tile_sums, int tiles) {
  __shared__ unsigned int part[1024];
  const int per = (tiles + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = min(lo + per, tiles);
  unsigned int s = 0;
  for (int i = lo; i < hi; ++i) s += tile_sums[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int run = 0;
    for (int i = 0; i < 1024; ++i) { const unsigned int v = part[i]; part[i] = run; run += v; }
    tile_sums[tiles] = run;
  }
  __syncthreads();
  unsigned int run = part[threadIdx.x];
  for (int i = lo; i < hi; ++i) { const unsigned int v = tile_sums[i]; tile_sums[i] = run; run += v; }
}

__global__ void __launch_bounds__(kScanThreads) bitmap_emit_kernel(const unsigned int* __restrict__ bits, long long words,
                                                                   const unsigned int* __restrict__ tile_offsets,
                                                                   unsigned int* __restrict__ out) {
  // a thread owns kScanPerThread CONSECUTIVE words, so the list comes out in ascending node order
  __shared__ unsigned int warp_sums[kScanThreads / 32];
  const long long base = static_cast<long long>(blockIdx.x) * kScanTile + static_cast<long long>(threadIdx.x) * kScanPerThread;
  unsigned int w[kScanPerThread];
  unsigned int n = 0;
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k) {
    w[k] = base + k < words ? bits[base + k] : 0u;
    n += __popc(w[k]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int inc = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  unsigned int off = tile_offsets[blockIdx.x] + inc - n;
  for (int i = 0; i < warp; ++i) off += warp_sums[i];
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k) {
    unsigned int b = w[k];
    while (b) {
      const int bit = __ffs(b) - 1;
      b &= b - 1;
      out[off++] = static_cast<unsigned int>((base + k) * 32 + bit);
    }
  }
}

// the complete sign bit-planes: one thread per word of 32 nodes
__global__ void fill_signs_kernel(SubGeom g, const float* __restrict__ dense, const unsigned int* __restrict__ need1,
                                  const unsigned int* __restrict__ need2, const float* __restrict__ cs, long long nodes,
                                  unsigned int* __restrict__ signs) {
  const long long wi = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long words = (nodes + 31) >> 5;
  if (wi >= words) return;
  const unsigned int known = need1[wi] | need2[wi];
  const int nsb = (g.res - 1 + g.B2 - 1) / g.B2;            // sub-blocks per axis
  const long long n1 = g.nb1 + 1;
  unsigned int out = 0;
  for (int b = 0; b < 32; ++b) {
    const long long node = wi * 32 + b;
    if (node >= nodes) break;
    bool neg;
    if ((known >> b) & 1u) {
      neg = dense[node] < 0.f;
    } else {
      const long long t = node / g.res;
      const int x = static_cast<int>(node - t * g.res), z = static_cast<int>(t / g.res), y = static_cast<int>(t - static_cast<long long>(z) * g.res);
      const int sx = min(x / g.B2, nsb - 1), sy = min(y / g.B2, nsb - 1), sz = min(z / g.B2, nsb - 1);
      const long long e = (static_cast<long long>(sz * g.B2) * g.res + sy * g.B2) * g.res + sx * g.B2;
      if (get_bit(need1, e) || get_bit(need2, e)) {
        neg = dense[e] < 0.f;
      } else {
        const int ax = min(sx * g.B2 / g.B1, g.nb1 - 1), ay = min(sy * g.B2 / g.B1, g.nb1 - 1), az = min(sz * g.B2 / g.B1, g.nb1 - 1);
        neg = cs[(az * n1 + ay) * n1 + ax] < 0.f;
      }
    }
    out |= (neg ? 1u : 0u) << b;
  }
  signs[wi] = out;
}

inline unsigned blocks_for(long long n, int threads = 256) { return static_cast<unsigned>((n + threads - 1) / threads); }

}  // namespace

cudaError_t launch_corner_nodes(int res, int B, int nb, unsigned int* idx, cudaStream_t st) {
  const long long total = static_cast<long long>(nb + 1) * (nb + 1) * (nb + 1);
  corner_nodes_kernel<<<blocks_for(total), 256, 0, st>>>(res, B, nb, idx);
  return cudaGetLastError();
}
cudaError_t launch_node_points(int res, const unsigned int* idx, long long n, float* xyz, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  node_points_kernel<<<blocks_for(n), 256, 0, st>>>(res, idx, n, xyz);
  return cudaGetLastError();
}
cudaError_t launch_scatter(const unsigned int* idx, const float* vals, long long n, float* dense, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  scatter_kernel<<<blocks_for(n), 256, 0, st>>>(idx, vals, n, dense);
  return cudaGetLastError();
}
cudaError_t launch_corner_lipschitz(const float* cs, int res, int B, int nb, unsigned int* out_bits, cudaStream_t st) {
  const long long total = static_cast<long long>(nb + 1) * (nb + 1) * (nb + 1);
  corner_lipschitz_kernel<<<blocks_for(total), 256, 0, st>>>(cs, res, B, nb, out_bits);
  return cudaGetLastError();
}
cudaError_t launch_mark_sub_corners(int res, int B1, int B2, int nb1, const int* blocks, long long nblk, unsigned int* need,
                                    cudaStream_t st) {
  const SubGeom g{res, B1, B2, nb1, B1 / B2};
  const long long total = nblk * g.per * g.per * g.per;
  if (total <= 0) return cudaSuccess;
  mark_sub_corners_kernel<<<blocks_for(total), 256, 0, st>>>(g, blocks, nblk, need);
  return cudaGetLastError();
}
cudaError_t launch_sub_lipschitz(int res, int B1, int B2, int nb1, const int* blocks, long long nblk, const float* dense,
                                 unsigned int* out_bits, unsigned int* per_block, cudaStream_t st) {
  const SubGeom g{res, B1, B2, nb1, B1 / B2};
  const long long total = nblk * g.per * g.per * g.per;
  if (total <= 0) return cudaSuccess;
  sub_lipschitz_kernel<<<blocks_for(total), 256, 0, st>>>(g, blocks, nblk, dense, out_bits, per_block);
  return cudaGetLastError();
}
cudaError_t launch_select_sub_blocks(int res, int B1, int B2, int nb1, const int* blocks, long long nblk, const float* dense,
                                     const unsigned int* lip_bits, float lip_given_per_node, float safety,
                                     const unsigned int* per_block, float local_floor, unsigned int* need,
                                     unsigned long long* kept, cudaStream_t st) {
  const SubGeom g{res, B1, B2, nb1, B1 / B2};
  const long long total = nblk * g.per * g.per * g.per;
  if (total <= 0) return cudaSuccess;
  select_sub_blocks_kernel<<<blocks_for(total), 256, 0, st>>>(g, blocks, nblk, dense, lip_bits, lip_given_per_node, safety, per_block,
                                                               local_floor, need, kept);
  return cudaGetLastError();
}
cudaError_t launch_select_blocks_bits(const float* cs, int nb, float tau, unsigned int* keep, cudaStream_t st) {
  const long long total = static_cast<long long>(nb) * nb * nb;
  select_blocks_bits_kernel<<<blocks_for(total), 256, 0, st>>>(cs, nb, tau, keep);
  return cudaGetLastError();
}
cudaError_t launch_andnot(unsigned int* a, const unsigned int* b, long long words, cudaStream_t st) {
  andnot_kernel<<<blocks_for(words), 256, 0, st>>>(a, b, words);
  return cudaGetLastError();
}
long long bitmap_scan_tiles(long long words) { return (words + kScanTile - 1) / kScanTile; }
// tile_sums: bitmap_scan_tiles(words) + 1 entries; afterwards tile_sums[tiles] holds the number of set bits
cudaError_t launch_bitmap_count(const unsigned int* bits, long long words, unsigned int* tile_sums, cudaStream_t st) {
  const long long tiles = bitmap_scan_tiles(words);
  bitmap_tile_counts_kernel<<<static_cast<unsigned>(tiles), kScanThreads, 0, st>>>(bits, words, tile_sums);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  scan_tile_sums_kernel<<<1, 1024, 0, st>>>(tile_sums, static_cast<int>(tiles));
  return cudaGetLastError();
}
cudaError_t launch_bitmap_emit(const unsigned int* bits, long long words, const unsigned int* tile_offsets, unsigned int* out,
                               cudaStream_t st) {
  const long long tiles = bitmap_scan_tiles(words);
  bitmap_emit_kernel<<<static_cast<unsigned>(tiles), kScanThreads, 0, st>>>(bits, words, tile_offsets, out);
  return cudaGetLastError();
}
cudaError_t launch_fill_signs(int res, int B1, int B2, int nb1, const float* dense, const unsigned int* need1,
                              const unsigned int* need2, const float* cs, unsigned int* signs, cudaStream_t st) {
  const SubGeom g{res, B1, B2, nb1, B1 / B2};
  const long long nodes = static_cast<long long>(res) * res * res;
  fill_signs_kernel<<<blocks_for((nodes + 31) >> 5), 256, 0, st>>>(g, dense, need1, need2, cs, nodes, signs);
  return cudaGetLastError();
}

}  // namespace sdfb

// Small HBM-bound kernels around the tensor-core products of the training steps (oracle:
// oracle/train.py; no upstream source exists, /root/reference/README.md:1): noising + input assembly, residual and loss,
// deterministic column sums (bias gradients), fused Adam with refresh of the 16-bit weight copies.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kernels.h"

namespace sdfb {

namespace {

template <bool FP16>
__device__ __forceinline__ uint16_t to_lowp_bits(float f) {
  if constexpr (FP16) return __half_as_ushort(__float2half_rn(f));
  else return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}
template <bool FP16>
__device__ __forceinline__ float from_lowp_bits(uint16_t b) {
  if constexpr (FP16) return __half2float(__ushort_as_half(b));
  else return __uint_as_float(static_cast<uint32_t>(b) << 16);
}

// one thread per (row, column): columns [0, 256) = x_t, [256, 512) = temb(t_row)
template <bool FP16>
__global__ void ddpm_train_prep_kernel(const float* __restrict__ x0, const float* __restrict__ eps, const int* __restrict__ t,
                                       const float* __restrict__ coef_ab, const float* __restrict__ temb, int n,
                                       uint16_t* __restrict__ in0) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(n) * 512) return;
  const long long r = i >> 9;
  const int c = static_cast<int>(i & 511);
  const int tr = t[r];
  float v;
  if (c < 256) {
    const float a = coef_ab[2 * tr], b = coef_ab[2 * tr + 1];
    v = __fadd_rn(__fmul_rn(a, x0[r * 256 + c]), __fmul_rn(b, eps[r * 256 + c]));
  } else {
    v = temb[tr * 256 + (c - 256)];
  }
  in0[i] = to_lowp_bits<FP16>(v);
}

constexpr int kResBlock = 256, kResPerThread = 16;

template <bool FP16>
__global__ void __launch_bounds__(kResBlock) ddpm_train_residual_kernel(const float* __restrict__ eps_hat, const float* __restrict__ eps,
                                                                        long long count, uint16_t* __restrict__ d_lowp,
                                                                        float* __restrict__ loss_partial) {
  __shared__ float warp_sums[kResBlock / 32];
  const long long base = static_cast<long long>(blockIdx.x) * kResBlock * kResPerThread;
  float s = 0.f;
#pragma unroll 4
  for (int k = 0; k < kResPerThread; ++k) {
    const long long i = base + static_cast<long long>(k) * kResBlock + threadIdx.x;
    if (i < count) {
      const float d = __fsub_rn(eps_hat[i], eps[i]);
      d_lowp[i] = to_lowp_bits<FP16>(d);
      s = fmaf(d, d, s);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < kResBlock / 32; ++i) tot += warp_sums[i];
    loss_partial[blockIdx.x] = tot;
  }
}

__global__ void sum_loss_kernel(const float* __restrict__ partial, int n, float scale, float* __restrict__ out) {
  // one warp, fixed order: lane l adds entries l, l + 32, ...; then a shuffle tree
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) out[0] = s * scale;
}

// column sums of a 16-bit matrix: block = 256 columns (32 threads x 8 columns, one 16-byte load per row) x 8 row-slices; rows in
// order inside a slice, slices added in order (blockIdx.y = row slab of `slab_rows` rows: partial sums [slab][N], added up in
// slab order by colsum_finish_kernel or by the Adam kernel).  ld must be a multiple of 8 (every caller pads it).
template <bool FP16>
__global__ void __launch_bounds__(256) colsum_lowp_kernel(const uint16_t* __restrict__ D, long long M, int ld, int N, long long slab_rows,
                                                          float* __restrict__ partial) {
  __shared__ float part[8][256 + 8];
  const int cg = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + cg * 8;
  const long long s0 = blockIdx.y * slab_rows, s1 = s0 + slab_rows < M ? s0 + slab_rows : M;
  const long long per = (s1 - s0 + 7) / 8;
  const long long r0 = s0 + slice * per, r1 = r0 + per < s1 ? r0 + per : s1;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < ld) {
    const uint16_t* src = D + col;
    for (long long r = r0; r < r1; ++r) {
      const uint4 q = *reinterpret_cast<const uint4*>(src + r * ld);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s[2 * e] += from_lowp_bits<FP16>(static_cast<uint16_t>(w[e] & 0xFFFFu));
        s[2 * e + 1] += from_lowp_bits<FP16>(static_cast<uint16_t>(w[e] >> 16));
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[slice][cg * 8 + e] = s[e];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += part[k][threadIdx.x];
    partial[static_cast<long long>(blockIdx.y) * N + c] = tot;
  }
}

__global__ void colsum_finish_kernel(const float* __restrict__ partial, int slabs, int N, float scale, float* __restrict__ out) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= N) return;
  float tot = 0.f;
  for (int k = 0; k < slabs; ++k) tot += partial[static_cast<long long>(k) * N + col];
  out[col] = tot * scale;
}

// Adam on a [rows][cols] tensor; tile of 32 x 32 per block (a transposed 16-bit copy is written coalesced through smem).
// VEC: a thread owns four consecutive columns of one row (128-bit accesses; needs cols, ld_grad, part_stride % 4 == 0 and
// 16-byte aligned bases - the launcher checks); otherwise one column of four rows.  The arithmetic per element is the same.
template <bool FP16, bool VEC>
__global__ void __launch_bounds__(256) adam_update_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v,
                                                          const float* __restrict__ grad, int nparts, long long part_stride, int ld_grad,
                                                          float scale, int rows, int cols, AdamParams a, uint16_t* __restrict__ w_lowp,
                                                          int ldw, uint16_t* __restrict__ wt_lowp, int ldwt, float* __restrict__ grad_out,
                                                          int apply) {
  __shared__ uint16_t tile[32][34];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  auto one = [&](int r, int c, float g, float& wi, float& mi, float& vi) {
    g *= scale;
    if (grad_out != nullptr) grad_out[static_cast<long long>(r) * cols + c] = g;
    if (apply) {
      mi = a.beta1 * mi + (1.f - a.beta1) * g;
      vi = a.beta2 * vi + (1.f - a.beta2) * g * g;
      wi = wi - a.lr * (mi / a.bias_corr1) / (sqrtf(vi / a.bias_corr2) + a.eps);
    }
  };
  if constexpr (VEC) {
    const int r = r0 + (threadIdx.x >> 3), cq = (threadIdx.x & 7) * 4, c = c0 + cq;
    uint16_t lw[4] = {0, 0, 0, 0};
    if (r < rows && c < cols) {                                   // cols % 4 == 0: the four columns exist together
      const long long i = static_cast<long long>(r) * cols + c;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < nparts; ++s) {
        const float4 q = *reinterpret_cast<const float4*>(grad + s * part_stride + static_cast<long long>(r) * ld_grad + c);
        g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
      }
      float4 wq = *reinterpret_cast<const float4*>(w + i);
      float4 mq = make_float4(0.f, 0.f, 0.f, 0.f), vq = mq;
      if (apply) { mq = *reinterpret_cast<const float4*>(m + i); vq = *reinterpret_cast<const float4*>(v + i); }
      one(r, c, g.x, wq.x, mq.x, vq.x);
      one(r, c + 1, g.y, wq.y, mq.y, vq.y);
      one(r, c + 2, g.z, wq.z, mq.z, vq.z);
      one(r, c + 3, g.w, wq.w, mq.w, vq.w);
      if (apply) {
        *reinterpret_cast<float4*>(m + i) = mq;
        *reinterpret_cast<float4*>(v + i) = vq;
        *reinterpret_cast<float4*>(w + i) = wq;
      }
      lw[0] = to_lowp_bits<FP16>(wq.x); lw[1] = to_lowp_bits<FP16>(wq.y); lw[2] = to_lowp_bits<FP16>(wq.z); lw[3] = to_lowp_bits<FP16>(wq.w);
      if (w_lowp != nullptr) {
        uint16_t* d = w_lowp + static_cast<long long>(r) * ldw + c;
        if ((ldw & 3) == 0) *reinterpret_cast<uint2*>(d) = make_uint2(lw[0] | (static_cast<uint32_t>(lw[1]) << 16), lw[2] | (static_cast<uint32_t>(lw[3]) << 16));
        else { d[0] = lw[0]; d[1] = lw[1]; d[2] = lw[2]; d[3] = lw[3]; }
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) tile[threadIdx.x >> 3][cq + e] = lw[e];
  } else {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + ty + 8 * k, c = c0 + tx;
      uint16_t lw = 0;
      if (r < rows && c < cols) {
        const long long i = static_cast<long long>(r) * cols + c;
        float g = 0.f;
        for (int s = 0; s < nparts; ++s) g += grad[s * part_stride + static_cast<long long>(r) * ld_grad + c];
        float wi = w[i], mi = apply ? m[i] : 0.f, vi = apply ? v[i] : 0.f;
        one(r, c, g, wi, mi, vi);
        if (apply) { m[i] = mi; v[i] = vi; w[i] = wi; }
        lw = to_lowp_bits<FP16>(wi);
        if (w_lowp != nullptr) w_lowp[static_cast<long long>(r) * ldw + c] = lw;
      }
      tile[ty + 8 * k][tx] = lw;
    }
  }
  if (wt_lowp == nullptr) return;
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, r = r0 + tx;         // transposed: row index of W^T = column of W
    if (r < rows && c < cols) wt_lowp[static_cast<long long>(c) * ldwt + r] = tile[tx][ty + 8 * k];
  }
}

template <bool FP16>
__global__ void __launch_bounds__(256) lowp_copies_kernel(const float* __restrict__ w, int ld, int rows, int cols,
                                                          uint16_t* __restrict__ w_lowp, int ldw, uint16_t* __restrict__ wt_lowp, int ldwt) {
  __shared__ uint16_t tile[32][34];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k, c = c0 + tx;
    uint16_t lw = 0;
    if (r < rows && c < cols) {
      lw = to_lowp_bits<FP16>(w[static_cast<long long>(r) * ld + c]);
      if (w_lowp != nullptr) w_lowp[static_cast<long long>(r) * ldw + c] = lw;
    }
    tile[ty + 8 * k][tx] = lw;
  }
  if (wt_lowp == nullptr) return;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, r = r0 + tx;
    if (r < rows && c < cols) wt_lowp[static_cast<long long>(c) * ldwt + r] = tile[tx][ty + 8 * k];
  }
}

// ---- decoder training helpers ------------------------------------------------------------------------------------------
// decoder input rows [M][ld] 16-bit: columns [0, 256) = the row's shape latent, 256..258 = xyz, the rest zero
template <bool FP16>
__global__ void __launch_bounds__(256) dec_train_input_kernel(const float* __restrict__ latents, const float* __restrict__ xyz, long long M,
                                                              long long per_shape, int ld, int col0, int ncols, uint16_t* __restrict__ out) {
  // eight rows per block, a thread per column (and column + 256): ONE division per block instead of one per element (the
  // rows' shape index only steps up inside the block)
  const long long rb = static_cast<long long>(blockIdx.x) * 8;
  long long shape = rb / per_shape, next = (shape + 1) * per_shape;
  for (int k = 0; k < 8; ++k) {
    const long long r = rb + k;
    if (r >= M) return;
    while (r >= next) { ++shape; next += per_shape; }
    const float* z = latents + shape * 256;
    for (int c = threadIdx.x; c < ncols; c += 256) {
      float v = 0.f;
      if (c < 256) v = z[c];
      else if (c < 259) v = xyz[r * 3 + (c - 256)];
      out[r * ld + col0 + c] = to_lowp_bits<FP16>(v);
    }
  }
}

// head forward + loss + first delta: y = tanh(a8 . w8 + b8) (fp32 dot over the 16-bit activations), clamped-L1 loss terms,
// d8 = sign(clamp(y) - clamp(target)) [|y| < clamp] (1 - y^2)  (the 1 / M is applied later), delta7 = d8 w8 [a8 > 0] -> 16-bit.
// One warp per row.
template <bool FP16>
__global__ void __launch_bounds__(256) dec_train_head_kernel(const uint16_t* __restrict__ a8, const float* __restrict__ w8,
                                                             const float* __restrict__ b8, const float* __restrict__ target, float clamp,
                                                             long long M, float* __restrict__ y_out, float* __restrict__ d8_out,
                                                             uint16_t* __restrict__ delta7, float* __restrict__ loss_partial) {
  __shared__ float wsum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long r = static_cast<long long>(blockIdx.x) * 8 + warp;
  float lterm = 0.f;
  if (r < M) {
    float h[16];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      h[k] = from_lowp_bits<FP16>(a8[r * 512 + lane + 32 * k]);
      dot = fmaf(h[k], w8[lane + 32 * k], dot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const float y = tanhf(dot + b8[0]);
    const float c = clamp;
    const float diff = fminf(fmaxf(y, -c), c) - fminf(fmaxf(target[r], -c), c);
    const float up = (y > -c && y < c) ? (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)) : 0.f;
    const float d8 = up * (1.f - y * y);
    if (lane == 0) { y_out[r] = y; d8_out[r] = d8; lterm = fabsf(diff); }
#pragma unroll
    for (int k = 0; k < 16; ++k)
      delta7[r * 512 + lane + 32 * k] = to_lowp_bits<FP16>(h[k] > 0.f ? d8 * w8[lane + 32 * k] : 0.f);
  }
  if (lane == 0) wsum[warp] = lterm;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += wsum[i];
    loss_partial[blockIdx.x] = s;
  }
}

// dW8[k] = scale sum_m d8[m] a8[m][k]; db8 = scale sum_m d8[m]   (same slicing as colsum_lowp_kernel; column 512 = the bias)
template <bool FP16>
__global__ void __launch_bounds__(256) dec_train_head_grad_kernel(const uint16_t* __restrict__ a8, const float* __restrict__ d8, long long M,
                                                                  long long slab_rows, float* __restrict__ partial /* [slabs][513] */) {
  __shared__ float part[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int slice = threadIdx.x >> 5;
  const long long s0 = blockIdx.y * slab_rows, s1 = s0 + slab_rows < M ? s0 + slab_rows : M;
  const long long per = (s1 - s0 + 7) / 8;
  const long long r0 = s0 + slice * per, r1 = r0 + per < s1 ? r0 + per : s1;
  float s = 0.f;
  if (col < 512) {
    for (long long r = r0; r < r1; ++r) s = fmaf(d8[r], from_lowp_bits<FP16>(a8[r * 512 + col]), s);
  } else if (col == 512) {
    for (long long r = r0; r < r1; ++r) s += d8[r];
  }
  part[slice][threadIdx.x & 31] = s;
  __syncthreads();
  if (slice == 0 && col <= 512) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += part[k][threadIdx.x];
    partial[static_cast<long long>(blockIdx.y) * 513 + col] = tot;
  }
}

}  // namespace

cudaError_t launch_ddpm_train_prep(const float* x0, const float* eps, const int* t, const float* coef_ab, const float* temb, int n,
                                   uint16_t* in0, bool fp16, cudaStream_t st) {
  const long long total = static_cast<long long>(n) * 512;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  if (fp16) ddpm_train_prep_kernel<true><<<blocks, 256, 0, st>>>(x0, eps, t, coef_ab, temb, n, in0);
  else ddpm_train_prep_kernel<false><<<blocks, 256, 0, st>>>(x0, eps, t, coef_ab, temb, n, in0);
  return cudaGetLastError();
}

cudaError_t launch_ddpm_train_residual(const float* eps_hat, const float* eps, int n, uint16_t* d_lowp, float* loss_partial,
                                       int* nblocks, bool fp16, cudaStream_t st) {
  const long long count = static_cast<long long>(n) * 256;
  const int blocks = static_cast<int>((count + kResBlock * kResPerThread - 1) / (kResBlock * kResPerThread));
  *nblocks = blocks;
  if (fp16) ddpm_train_residual_kernel<true><<<blocks, kResBlock, 0, st>>>(eps_hat, eps, count, d_lowp, loss_partial);
  else ddpm_train_residual_kernel<false><<<blocks, kResBlock, 0, st>>>(eps_hat, eps, count, d_lowp, loss_partial);
  return cudaGetLastError();
}

cudaError_t launch_sum_loss(const float* partial, int n, float scale, float* loss_out, cudaStream_t st) {
  sum_loss_kernel<<<1, 32, 0, st>>>(partial, n, scale, loss_out);
  return cudaGetLastError();
}

// `scratch`: kColsumSlabs x N floats
cudaError_t launch_colsum_lowp(const uint16_t* D, long long M, int ld, int N, float scale, float* out, float* scratch, bool fp16,
                               cudaStream_t st, int* nslabs_out) {
  // row slabs of at least 64 rows, at most kColsumSlabs of them: a 4096-row batch is 64 slabs x N / 32 blocks (one slab took
  // 108 us per layer of the DDPM training step: 32 blocks walking 4096 rows each)
  const long long want = (M + 63) / 64;
  const long long nsl = want < 1 ? 1 : (want > kColsumSlabs ? kColsumSlabs : want);
  const long long slab_rows = (M + nsl - 1) / nsl > 0 ? (M + nsl - 1) / nsl : 1;
  const int slabs = static_cast<int>((M + slab_rows - 1) / slab_rows);
  if (ld & 7) return cudaErrorInvalidValue;
  const dim3 grid(static_cast<unsigned>((N + 255) / 256), static_cast<unsigned>(slabs > 0 ? slabs : 1));
  if (fp16) colsum_lowp_kernel<true><<<grid, 256, 0, st>>>(D, M, ld, N, slab_rows, scratch);
  else colsum_lowp_kernel<false><<<grid, 256, 0, st>>>(D, M, ld, N, slab_rows, scratch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (nslabs_out != nullptr) *nslabs_out = slabs > 0 ? slabs : 1;
  if (out == nullptr) return cudaSuccess;          // the caller adds the slabs up itself (launch_adam_update does, in slab order)
  colsum_finish_kernel<<<(N + 255) / 256, 256, 0, st>>>(scratch, slabs > 0 ? slabs : 1, N, scale, out);
  return cudaGetLastError();
}

cudaError_t launch_adam_update(float* w, float* m, float* v, const float* grad, int nparts, long long part_stride, int ld_grad,
                               float scale, int rows, int cols, const AdamParams& a, uint16_t* w_lowp, int ldw, uint16_t* wt_lowp,
                               int ldwt, float* grad_out, bool fp16, bool apply, cudaStream_t st) {
  const dim3 grid(static_cast<unsigned>((cols + 31) / 32), static_cast<unsigned>((rows + 31) / 32));
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const bool vec = (cols & 3) == 0 && (ld_grad & 3) == 0 && (part_stride & 3) == 0 && al16(w) && al16(m) && al16(v) && al16(grad) &&
                   (grad_out == nullptr || al16(grad_out)) && (w_lowp == nullptr || (reinterpret_cast<uintptr_t>(w_lowp) & 7u) == 0);
#define SDFB_ADAM(F, V)                                                                                                              \
  adam_update_kernel<F, V><<<grid, 256, 0, st>>>(w, m, v, grad, nparts, part_stride, ld_grad, scale, rows, cols, a, w_lowp, ldw, wt_lowp, \
                                                 ldwt, grad_out, apply ? 1 : 0)
  if (fp16) { if (vec) SDFB_ADAM(true, true); else SDFB_ADAM(true, false); }
  else { if (vec) SDFB_ADAM(false, true); else SDFB_ADAM(false, false); }
#undef SDFB_ADAM
  return cudaGetLastError();
}

cudaError_t launch_lowp_copies(const float* w, int ld, int rows, int cols, uint16_t* w_lowp, int ldw, uint16_t* wt_lowp, int ldwt,
                               bool fp16, cudaStream_t st) {
  const dim3 grid(static_cast<unsigned>((cols + 31) / 32), static_cast<unsigned>((rows + 31) / 32));
  if (fp16) lowp_copies_kernel<true><<<grid, 256, 0, st>>>(w, ld, rows, cols, w_lowp, ldw, wt_lowp, ldwt);
  else lowp_copies_kernel<false><<<grid, 256, 0, st>>>(w, ld, rows, cols, w_lowp, ldw, wt_lowp, ldwt);
  return cudaGetLastError();
}

}  // namespace sdfb

namespace sdfb {
cudaError_t launch_dec_train_input(const float* latents, const float* xyz, long long M, long long per_shape, int ld, int col0, int ncols,
                                   uint16_t* out, bool fp16, cudaStream_t st) {
  if (M <= 0 || ncols <= 0) return cudaSuccess;
  const unsigned blocks = static_cast<unsigned>((M + 7) / 8);
  if (fp16) dec_train_input_kernel<true><<<blocks, 256, 0, st>>>(latents, xyz, M, per_shape, ld, col0, ncols, out);
  else dec_train_input_kernel<false><<<blocks, 256, 0, st>>>(latents, xyz, M, per_shape, ld, col0, ncols, out);
  return cudaGetLastError();
}
cudaError_t launch_dec_train_head(const uint16_t* a8, const float* w8, const float* b8, const float* target, float clamp, long long M,
                                  float* y, float* d8, uint16_t* delta7, float* loss_partial, int* nblocks, bool fp16, cudaStream_t st) {
  const int blocks = static_cast<int>((M + 7) / 8);
  *nblocks = blocks;
  if (fp16) dec_train_head_kernel<true><<<blocks, 256, 0, st>>>(a8, w8, b8, target, clamp, M, y, d8, delta7, loss_partial);
  else dec_train_head_kernel<false><<<blocks, 256, 0, st>>>(a8, w8, b8, target, clamp, M, y, d8, delta7, loss_partial);
  return cudaGetLastError();
}
cudaError_t launch_dec_train_head_grad(const uint16_t* a8, const float* d8, long long M, float scale, float* out, float* scratch, bool fp16,
                                       cudaStream_t st) {
  // row slabs of at least 64 rows, at most kColsumSlabs of them: a 4096-row batch is 64 slabs x N / 32 blocks (one slab took
  // 108 us per layer of the DDPM training step: 32 blocks walking 4096 rows each)
  const long long want = (M + 63) / 64;
  const long long nsl = want < 1 ? 1 : (want > kColsumSlabs ? kColsumSlabs : want);
  const long long slab_rows = (M + nsl - 1) / nsl > 0 ? (M + nsl - 1) / nsl : 1;
  const int slabs = static_cast<int>((M + slab_rows - 1) / slab_rows);
  const dim3 grid(17, static_cast<unsigned>(slabs > 0 ? slabs : 1));
  if (fp16) dec_train_head_grad_kernel<true><<<grid, 256, 0, st>>>(a8, d8, M, slab_rows, scratch);
  else dec_train_head_grad_kernel<false><<<grid, 256, 0, st>>>(a8, d8, M, slab_rows, scratch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  colsum_finish_kernel<<<3, 256, 0, st>>>(scratch, slabs > 0 ? slabs : 1, 513, scale, out);
  return cudaGetLastError();
}

// ---- Adam on a batch of latents (auto-decoder fitting): one block per latent ----
// g = grad + 2 reg z;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g g;  z -= lr (m / c1) / (sqrt(v / c2) + eps);
// loss[b] += reg |z_b|^2 (the latent the loss was evaluated at, i.e. before the update).  Every operation is rounded on its
// own, in the order of the fp32 tensor expression in api.py it replaced - including that expression's division of a tensor by
// a scalar as a multiplication by the scalar's fp32 reciprocal: the moments come out bit-identical to it, the latents within
// an ulp or two per step (tests/test_gpu_train.py).
namespace {
__global__ void __launch_bounds__(256) latent_adam_kernel(float* __restrict__ z, float* __restrict__ m, float* __restrict__ v,
                                                          const float* __restrict__ grad, float* __restrict__ loss, int dim,
                                                          float lr, float reg, float b1, float omb1, float b2, float omb2,
                                                          float c1, float c2, float eps) {
  __shared__ float red[8];
  const long long base = static_cast<long long>(blockIdx.x) * dim;
  float sq = 0.f;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) {
    const float zi = z[base + i];
    sq = __fadd_rn(sq, __fmul_rn(zi, zi));
    const float g = __fadd_rn(grad[base + i], __fmul_rn(__fmul_rn(2.f, reg), zi));
    const float mi = __fadd_rn(__fmul_rn(b1, m[base + i]), __fmul_rn(omb1, g));
    const float vi = __fadd_rn(__fmul_rn(b2, v[base + i]), __fmul_rn(__fmul_rn(omb2, g), g));
    m[base + i] = mi;
    v[base + i] = vi;
    const float num = __fmul_rn(lr, __fmul_rn(mi, c1));          // c1, c2: reciprocals of the bias corrections
    const float den = __fadd_rn(__fsqrt_rn(__fmul_rn(vi, c2)), eps);
    z[base + i] = __fsub_rn(zi, __fdiv_rn(num, den));
  }
  if (loss != nullptr) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sq += __shfl_xor_sync(0xFFFFFFFFu, sq, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) s += red[w];
      loss[blockIdx.x] = __fadd_rn(loss[blockIdx.x], __fmul_rn(reg, s));
    }
  }
}
}  // namespace

cudaError_t launch_latent_adam(float* z, float* m, float* v, const float* grad, float* loss, int batch, int dim, float lr, float reg,
                               double beta1, double beta2, float eps, int step, cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  // 1 - beta^step in double, rounded once, then its fp32 reciprocal
  const float c1 = 1.0f / static_cast<float>(1.0 - pow(beta1, step));
  const float c2 = 1.0f / static_cast<float>(1.0 - pow(beta2, step));
  const float omb1 = static_cast<float>(1.0 - beta1), omb2 = static_cast<float>(1.0 - beta2);
  latent_adam_kernel<<<batch, 256, 0, st>>>(z, m, v, grad, loss, dim, lr, reg, static_cast<float>(beta1), omb1, static_cast<float>(beta2), omb2, c1,
                                            c2, eps);
  return cudaGetLastError();
}

}  // namespace sdfb

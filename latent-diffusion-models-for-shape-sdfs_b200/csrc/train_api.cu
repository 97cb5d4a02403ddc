// C ABI of the training steps (include/sdfb200.h, "training" section; SURVEY.md section 8f row N4, second half): a DDPM
// denoiser training step and the decoder's weight gradients, layer by layer on the general tensor-core product
// (gemm_tc.cu) with every activation and delta kept in 16 bits for the weight-gradient products, fused Adam.
// Oracle: oracle/train.py (fp64 torch autograd + an operand-rounding emulation).  No upstream source exists
// (/root/reference/README.md:1).
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/sdfb200.h"
#include "kernels.h"

namespace sdfb {
int set_error(int code, const char* fmt, ...);
}
using namespace sdfb;

namespace {

#define CU_TRY(expr)                                                                         \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return set_error(SDFB_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

struct DevGuard {
  int prev = -1;
  explicit DevGuard(int dev) { if (cudaGetDevice(&prev) == cudaSuccess) cudaSetDevice(dev); }
  ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int check_sm100(int device, int* num_sms) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return set_error(SDFB_E_DEVICE, "no CUDA device visible: libsdfb200 has no CPU path");
  if (device < 0 || device >= count) return set_error(SDFB_E_INVALID, "device %d out of range (%d devices)", device, count);
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return set_error(SDFB_E_DEVICE, "device %d is sm_%d%d; the kernels are built for sm_100a only", device, prop.major, prop.minor);
  *num_sms = prop.multiProcessorCount;
  return SDFB_OK;
}

int pick_ksplit(int out_rows, int in_cols, int bn, long long k) {
  const int tiles = ((out_rows + 127) / 128) * (in_cols / bn);
  const int ksteps = static_cast<int>((k + 63) / 64);
  int s = 128 / (tiles > 0 ? tiles : 1);
  if (s < 1) s = 1;
  if (s > 32) s = 32;
  if (s > ksteps) s = ksteps;
  if (s < 1) s = 1;
  const int per = (ksteps + s - 1) / s;          // every range non-empty: as many ranges as `per` steps each need
  return (ksteps + per - 1) / per;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ DDPM trainer ----
struct sdfb_ddpm_trainer {
  int device = 0, num_sms = 0;
  bool fp16 = false;
  static constexpr int kL = 5;
  int fin[kL] = {512, 1024, 1024, 1024, 1024}, fout[kL] = {1024, 1024, 1024, 1024, 256};
  long long woff[kL], boff[kL];
  float *params = nullptr, *adam_m = nullptr, *adam_v = nullptr;       // kDdpmParamFloats each
  uint16_t* w_lowp[kL] = {};      // W_l  [out][in]
  uint16_t* wt_lowp[kL] = {};     // W_l^T [in][out]
  float *temb = nullptr, *coef_ab = nullptr;
  // per-batch workspace
  int ws_n = 0;
  uint16_t* act[kL] = {};         // act[0] = [x_t | temb] [n][512]; act[l] = h_l [n][1024]
  uint16_t* delta[kL] = {};       // delta[l] [n][fout[l]]
  float *eps_hat = nullptr, *partial = nullptr, *bias_grad = nullptr, *loss_partial = nullptr, *colsum_scratch = nullptr;
  long long partial_floats = 0;
  unsigned int *status = nullptr, *status_host = nullptr, *status_host_dev = nullptr;
  long long step = 0;
};

extern "C" {

int sdfb_ddpm_trainer_destroy(sdfb_ddpm_trainer* t) {
  if (!t) return SDFB_OK;
  DevGuard g(t->device);
  cudaDeviceSynchronize();
  cudaFree(t->params); cudaFree(t->adam_m); cudaFree(t->adam_v); cudaFree(t->temb); cudaFree(t->coef_ab);
  for (int l = 0; l < sdfb_ddpm_trainer::kL; ++l) { cudaFree(t->w_lowp[l]); cudaFree(t->wt_lowp[l]); cudaFree(t->act[l]); cudaFree(t->delta[l]); }
  cudaFree(t->eps_hat); cudaFree(t->partial); cudaFree(t->bias_grad); cudaFree(t->loss_partial); cudaFree(t->colsum_scratch); cudaFree(t->status);
  if (t->status_host) cudaFreeHost(t->status_host);
  delete t;
  return SDFB_OK;
}

int sdfb_ddpm_trainer_create(const float* params_host, size_t n_floats, int device, int precision, sdfb_ddpm_trainer** out) {
  if (!out || !params_host) return set_error(SDFB_E_INVALID, "null argument");
  *out = nullptr;
  if (n_floats != static_cast<size_t>(kDdpmParamFloats))
    return set_error(SDFB_E_INVALID, "denoiser blob must hold %lld floats, got %zu", kDdpmParamFloats, n_floats);
  if (precision != SDFB_PREC_BF16 && precision != SDFB_PREC_FP16)
    return set_error(SDFB_E_INVALID, "the training step runs its products in bf16 or fp16 (precision %d)", precision);
  int sms = 0;
  int rc = check_sm100(device, &sms);
  if (rc) return rc;
  DevGuard g(device);
  sdfb_ddpm_trainer* t = new (std::nothrow) sdfb_ddpm_trainer();
  if (!t) return set_error(SDFB_E_NOMEM, "out of host memory");
  t->device = device; t->num_sms = sms; t->fp16 = precision == SDFB_PREC_FP16;
  long long o = 0;
  for (int l = 0; l < t->kL; ++l) { t->woff[l] = o; o += static_cast<long long>(t->fin[l]) * t->fout[l]; t->boff[l] = o; o += t->fout[l]; }
  auto bail = [&](int code) { sdfb_ddpm_trainer_destroy(t); return code; };
#define CU_TRY_T(expr)                                                                                        \
  do {                                                                                                        \
    cudaError_t e__ = (expr);                                                                                 \
    if (e__ != cudaSuccess) return bail(set_error(SDFB_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__))); \
  } while (0)
  CU_TRY_T(gemm_tc_init());
  const size_t pbytes = n_floats * sizeof(float);
  CU_TRY_T(cudaMalloc(&t->params, pbytes));
  CU_TRY_T(cudaMalloc(&t->adam_m, pbytes));
  CU_TRY_T(cudaMalloc(&t->adam_v, pbytes));
  CU_TRY_T(cudaMemcpy(t->params, params_host, pbytes, cudaMemcpyHostToDevice));
  CU_TRY_T(cudaMemset(t->adam_m, 0, pbytes));
  CU_TRY_T(cudaMemset(t->adam_v, 0, pbytes));
  for (int l = 0; l < t->kL; ++l) {
    const size_t wb = static_cast<size_t>(t->fin[l]) * t->fout[l] * 2;
    CU_TRY_T(cudaMalloc(&t->w_lowp[l], wb));
    CU_TRY_T(cudaMalloc(&t->wt_lowp[l], wb));
    CU_TRY_T(launch_lowp_copies(t->params + t->woff[l], t->fin[l], t->fout[l], t->fin[l], t->w_lowp[l], t->fin[l], t->wt_lowp[l], t->fout[l],
                                t->fp16, nullptr));
  }
  {   // A5: schedule and time embedding in fp64, cast to fp32 (oracle/ddpm.py ddpm_schedule / time_embedding)
    std::vector<float> ab(2 * kDdpmT), te(static_cast<size_t>(kDdpmT) * kDdpmTemb);
    double abar = 1.0;
    const int half = kDdpmTemb / 2;
    for (int s = 0; s < kDdpmT; ++s) {
      const double beta = 1e-4 + (0.02 - 1e-4) * static_cast<double>(s) / (kDdpmT - 1);
      abar *= 1.0 - beta;
      ab[2 * s] = static_cast<float>(std::sqrt(abar));
      ab[2 * s + 1] = static_cast<float>(std::sqrt(1.0 - abar));
      for (int i = 0; i < half; ++i) {
        const double f = std::exp(-std::log(10000.0) * i / half);
        te[static_cast<size_t>(s) * kDdpmTemb + i] = static_cast<float>(std::sin(s * f));
        te[static_cast<size_t>(s) * kDdpmTemb + half + i] = static_cast<float>(std::cos(s * f));
      }
    }
    CU_TRY_T(cudaMalloc(&t->coef_ab, ab.size() * sizeof(float)));
    CU_TRY_T(cudaMemcpy(t->coef_ab, ab.data(), ab.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU_TRY_T(cudaMalloc(&t->temb, te.size() * sizeof(float)));
    CU_TRY_T(cudaMemcpy(t->temb, te.data(), te.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  CU_TRY_T(cudaMalloc(&t->bias_grad, 1024 * sizeof(float)));
  CU_TRY_T(cudaMalloc(&t->colsum_scratch, static_cast<size_t>(kColsumSlabs) * 1024 * sizeof(float)));
  CU_TRY_T(cudaMalloc(&t->status, sizeof(unsigned int)));
  CU_TRY_T(cudaMemset(t->status, 0, sizeof(unsigned int)));
  CU_TRY_T(cudaHostAlloc(reinterpret_cast<void**>(&t->status_host), sizeof(unsigned int), cudaHostAllocMapped));
  *t->status_host = 0;
  CU_TRY_T(cudaHostGetDevicePointer(reinterpret_cast<void**>(&t->status_host_dev), t->status_host, 0));
  CU_TRY_T(cudaDeviceSynchronize());
#undef CU_TRY_T
  *out = t;
  return SDFB_OK;
}

int sdfb_ddpm_trainer_get_params(sdfb_ddpm_trainer* t, float* params_host) {
  if (!t || !params_host) return set_error(SDFB_E_INVALID, "null argument");
  DevGuard g(t->device);
  CU_TRY(cudaDeviceSynchronize());
  CU_TRY(cudaMemcpy(params_host, t->params, static_cast<size_t>(kDdpmParamFloats) * sizeof(float), cudaMemcpyDeviceToHost));
  unsigned int s = *reinterpret_cast<volatile unsigned int*>(t->status_host);
  if (s != 0) {
    *reinterpret_cast<volatile unsigned int*>(t->status_host) = 0;
    cudaMemset(t->status, 0, sizeof(unsigned int));
    return set_error(SDFB_E_KERNEL, "a training kernel's watchdog tripped at wait site 0x%x", s);
  }
  return SDFB_OK;
}

// One training step on a batch: x_t = sqrt(abar_t) x0 + sqrt(1 - abar_t) eps per row, loss = mean (eps_hat(x_t, t) - eps)^2,
// gradients of every weight and bias, Adam (apply != 0).  grads_dev (optional, blob layout): the gradient itself.
int sdfb_ddpm_trainer_step(sdfb_ddpm_trainer* t, const float* x0_dev, const int32_t* t_dev, const float* eps_dev, int n, float lr,
                           float beta1, float beta2, float adam_eps, int apply, float* loss_dev, float* grads_dev, void* stream) {
  if (!t || !x0_dev || !t_dev || !eps_dev || !loss_dev) return set_error(SDFB_E_INVALID, "null argument");
  if (n <= 0) return set_error(SDFB_E_INVALID, "batch size must be positive");
  DevGuard g(t->device);
  {
    const unsigned int s = *reinterpret_cast<volatile unsigned int*>(t->status_host);
    if (s != 0) {
      *reinterpret_cast<volatile unsigned int*>(t->status_host) = 0;
      cudaMemset(t->status, 0, sizeof(unsigned int));
      return set_error(SDFB_E_KERNEL, "an earlier training kernel's watchdog tripped at wait site 0x%x", s);
    }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  constexpr int L = sdfb_ddpm_trainer::kL;
  if (t->ws_n < n) {
    for (int l = 0; l < L; ++l) { cudaFree(t->act[l]); cudaFree(t->delta[l]); t->act[l] = nullptr; t->delta[l] = nullptr; }
    cudaFree(t->eps_hat); cudaFree(t->loss_partial); t->eps_hat = nullptr; t->loss_partial = nullptr; t->ws_n = 0;
    for (int l = 0; l < L; ++l) {
      CU_TRY(cudaMalloc(&t->act[l], static_cast<size_t>(n) * t->fin[l] * 2));
      CU_TRY(cudaMalloc(&t->delta[l], static_cast<size_t>(n) * t->fout[l] * 2));
    }
    CU_TRY(cudaMalloc(&t->eps_hat, static_cast<size_t>(n) * 256 * sizeof(float)));
    CU_TRY(cudaMalloc(&t->loss_partial, (static_cast<size_t>(n) * 256 / 4096 + 2) * sizeof(float)));
    t->ws_n = n;
  }
  const long long need_partial = 32LL * 1024 * 1024;
  if (t->partial_floats < need_partial) {
    cudaFree(t->partial); t->partial = nullptr; t->partial_floats = 0;
    CU_TRY(cudaMalloc(&t->partial, need_partial * sizeof(float)));
    t->partial_floats = need_partial;
  }
  auto gp = [&]() {
    GemmParams p{};
    p.M = n; p.bn = 256; p.ksplit = 1; p.alpha = 1.f;
    p.status = t->status; p.status_host = t->status_host_dev; p.timeout_ns = 4000000000ull;
    return p;
  };
  // ---- forward ----
  CU_TRY(launch_ddpm_train_prep(x0_dev, eps_dev, t_dev, t->coef_ab, t->temb, n, t->act[0], t->fp16, st));
  for (int l = 0; l < L; ++l) {
    GemmParams p = gp();
    p.N = t->fout[l]; p.K = t->fin[l];
    p.bias = t->params + t->boff[l];
    if (l < L - 1) { p.epi = kGemmEpiBiasReluLowp; p.out_lowp = t->act[l + 1]; p.ldo_lowp = t->fout[l]; }
    else { p.epi = kGemmEpiF32; p.out_f32 = t->eps_hat; p.ldo_f32 = t->fout[l]; }
    CU_TRY(launch_gemm_tc(p, t->act[l], t->fin[l], t->w_lowp[l], t->fin[l], false, t->fp16, t->num_sms, st));
  }
  int nblk = 0;
  CU_TRY(launch_ddpm_train_residual(t->eps_hat, eps_dev, n, t->delta[L - 1], t->loss_partial, &nblk, t->fp16, st));
  const float inv_count = 1.f / (static_cast<float>(n) * 256.f);
  CU_TRY(launch_sum_loss(t->loss_partial, nblk, inv_count, loss_dev, st));
  // ---- backward: delta_{l-1} before W_l changes, then dW_l, db_l, Adam ----
  const float gscale = 2.f * inv_count;          // d loss / d eps_hat = 2 (eps_hat - eps) / (n 256); the deltas carry the bare residual
  if (apply) ++t->step;
  AdamParams ap{lr, beta1, beta2, adam_eps, 1.f - std::pow(beta1, static_cast<float>(t->step > 0 ? t->step : 1)),
                1.f - std::pow(beta2, static_cast<float>(t->step > 0 ? t->step : 1))};
  for (int l = L - 1; l >= 0; --l) {
    if (l > 0) {   // delta_{l-1} = (delta_l W_l) where h_l > 0
      GemmParams p = gp();
      p.N = t->fin[l]; p.K = t->fout[l];
      p.epi = kGemmEpiMaskLowp; p.mask_h = t->act[l]; p.ldh = t->fin[l];
      p.out_lowp = t->delta[l - 1]; p.ldo_lowp = t->fin[l];
      CU_TRY(launch_gemm_tc(p, t->delta[l], t->fout[l], t->wt_lowp[l], t->fout[l], false, t->fp16, t->num_sms, st));
    }
    {              // dW_l [out][in] = delta_l^T h_l: the contraction runs over the batch rows
      GemmParams p = gp();
      p.M = t->fout[l]; p.N = t->fin[l]; p.K = n;
      p.ksplit = pick_ksplit(p.M, p.N, p.bn, n);
      p.epi = kGemmEpiF32; p.out_f32 = t->partial; p.ldo_f32 = t->fin[l];
      p.split_stride = static_cast<long long>(t->fout[l]) * t->fin[l];
      CU_TRY(launch_gemm_tc(p, t->delta[l], t->fout[l], t->act[l], t->fin[l], true, t->fp16, t->num_sms, st));
      CU_TRY(launch_adam_update(t->params + t->woff[l], t->adam_m + t->woff[l], t->adam_v + t->woff[l], t->partial, p.ksplit, p.split_stride,
                                t->fin[l], gscale, t->fout[l], t->fin[l], ap, t->w_lowp[l], t->fin[l], t->wt_lowp[l], t->fout[l],
                                grads_dev ? grads_dev + t->woff[l] : nullptr, t->fp16, apply != 0, st));
    }
    int nsl = 1;               // db_l: per-slab column sums of delta_l, added up (in slab order) by the Adam kernel itself
    CU_TRY(launch_colsum_lowp(t->delta[l], n, t->fout[l], t->fout[l], 1.f, nullptr, t->colsum_scratch, t->fp16, st, &nsl));
    CU_TRY(launch_adam_update(t->params + t->boff[l], t->adam_m + t->boff[l], t->adam_v + t->boff[l], t->colsum_scratch, nsl, t->fout[l], 1, gscale,
                              t->fout[l], 1, ap, nullptr, 0, nullptr, 0, grads_dev ? grads_dev + t->boff[l] : nullptr, t->fp16, apply != 0, st));
  }
  return SDFB_OK;
}

// --------------------------------------------------------------------------------------------- decoder trainer ----
// Weight gradients of the auto-decoder (DeepSDF training step) for a batch of shapes, `points_per_shape` samples each:
// loss = mean over all points of |clamp(sdf) - clamp(target)|.  Layer by layer on the general product; the input rows are
// the dense [z_shape | xyz] (padded to 320 columns), the skip layer's input the dense [h3 (253) | z | xyz] - as the oracle
// computes it (no latent fold: the latents differ per row).  Layer 3's 253 outputs are padded to 256 with zero weights.
}  // extern "C"

struct sdfb_decoder_trainer {
  int device = 0, num_sms = 0;
  bool fp16 = false;
  static constexpr int kL = 8;                           // hidden layers 0..7 on the tensor pipe; layer 8 (the head) in fp32 SIMT
  int fin[kL] = {259, 512, 512, 512, 512, 512, 512, 512}, fout[kL] = {512, 512, 512, 253, 512, 512, 512, 512};
  int fin_p[kL] = {320, 512, 512, 512, 512, 512, 512, 512}, fout_p[kL] = {512, 512, 512, 256, 512, 512, 512, 512};   // padded
  long long woff[9], boff[9];
  float *params = nullptr, *adam_m = nullptr, *adam_v = nullptr;
  uint16_t* w_lowp[kL] = {};      // W_l   [fout_p][fin_p]
  uint16_t* wt_lowp[kL] = {};     // W_l^T [fin_p][fout_p]
  float* b3pad = nullptr;         // layer 3's bias padded to 256
  long long ws_rows = 0;
  uint16_t* act[kL + 1] = {};     // act[l] = input of layer l: act[0] [M][320], act[4] = [h3 | z | xyz] [M][512], others [M][512]; act[8] = h7
  uint16_t* delta[kL] = {};       // delta[l] [M][fout_p[l]]
  float *y = nullptr, *d8 = nullptr, *partial = nullptr, *small_grad = nullptr, *loss_partial = nullptr, *colsum_scratch = nullptr;
  unsigned int *status = nullptr, *status_host = nullptr, *status_host_dev = nullptr;
  long long step = 0;
};

extern "C" {

int sdfb_decoder_trainer_destroy(sdfb_decoder_trainer* t) {
  if (!t) return SDFB_OK;
  DevGuard g(t->device);
  cudaDeviceSynchronize();
  cudaFree(t->params); cudaFree(t->adam_m); cudaFree(t->adam_v); cudaFree(t->b3pad);
  for (int l = 0; l < sdfb_decoder_trainer::kL; ++l) { cudaFree(t->w_lowp[l]); cudaFree(t->wt_lowp[l]); cudaFree(t->delta[l]); }
  for (int l = 0; l <= sdfb_decoder_trainer::kL; ++l) cudaFree(t->act[l]);
  cudaFree(t->y); cudaFree(t->d8); cudaFree(t->partial); cudaFree(t->small_grad); cudaFree(t->loss_partial); cudaFree(t->colsum_scratch); cudaFree(t->status);
  if (t->status_host) cudaFreeHost(t->status_host);
  delete t;
  return SDFB_OK;
}

int sdfb_decoder_trainer_create(const float* params_host, size_t n_floats, int device, int precision, sdfb_decoder_trainer** out) {
  if (!out || !params_host) return set_error(SDFB_E_INVALID, "null argument");
  *out = nullptr;
  if (n_floats != static_cast<size_t>(kDecParamFloats))
    return set_error(SDFB_E_INVALID, "decoder blob must hold %lld floats, got %zu", kDecParamFloats, n_floats);
  if (precision != SDFB_PREC_BF16 && precision != SDFB_PREC_FP16)
    return set_error(SDFB_E_INVALID, "the training step runs its products in bf16 or fp16 (precision %d)", precision);
  int sms = 0;
  int rc = check_sm100(device, &sms);
  if (rc) return rc;
  DevGuard g(device);
  sdfb_decoder_trainer* t = new (std::nothrow) sdfb_decoder_trainer();
  if (!t) return set_error(SDFB_E_NOMEM, "out of host memory");
  t->device = device; t->num_sms = sms; t->fp16 = precision == SDFB_PREC_FP16;
  {
    const int fin9[9] = {259, 512, 512, 512, 512, 512, 512, 512, 512}, fout9[9] = {512, 512, 512, 253, 512, 512, 512, 512, 1};
    long long o = 0;
    for (int l = 0; l < 9; ++l) { t->woff[l] = o; o += static_cast<long long>(fin9[l]) * fout9[l]; t->boff[l] = o; o += fout9[l]; }
  }
  auto bail = [&](int code) { sdfb_decoder_trainer_destroy(t); return code; };
#define CU_TRY_T(expr)                                                                                        \
  do {                                                                                                        \
    cudaError_t e__ = (expr);                                                                                 \
    if (e__ != cudaSuccess) return bail(set_error(SDFB_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__))); \
  } while (0)
  CU_TRY_T(gemm_tc_init());
  const size_t pbytes = n_floats * sizeof(float);
  CU_TRY_T(cudaMalloc(&t->params, pbytes));
  CU_TRY_T(cudaMalloc(&t->adam_m, pbytes));
  CU_TRY_T(cudaMalloc(&t->adam_v, pbytes));
  CU_TRY_T(cudaMemcpy(t->params, params_host, pbytes, cudaMemcpyHostToDevice));
  CU_TRY_T(cudaMemset(t->adam_m, 0, pbytes));
  CU_TRY_T(cudaMemset(t->adam_v, 0, pbytes));
  for (int l = 0; l < t->kL; ++l) {
    const size_t wb = static_cast<size_t>(t->fin_p[l]) * t->fout_p[l] * 2;
    CU_TRY_T(cudaMalloc(&t->w_lowp[l], wb));
    CU_TRY_T(cudaMalloc(&t->wt_lowp[l], wb));
    CU_TRY_T(cudaMemset(t->w_lowp[l], 0, wb));           // the padding rows / columns stay zero for good
    CU_TRY_T(cudaMemset(t->wt_lowp[l], 0, wb));
    CU_TRY_T(launch_lowp_copies(t->params + t->woff[l], t->fin[l], t->fout[l], t->fin[l], t->w_lowp[l], t->fin_p[l], t->wt_lowp[l],
                                t->fout_p[l], t->fp16, nullptr));
  }
  CU_TRY_T(cudaMalloc(&t->b3pad, 256 * sizeof(float)));
  CU_TRY_T(cudaMemset(t->b3pad, 0, 256 * sizeof(float)));
  CU_TRY_T(cudaMalloc(&t->small_grad, 1024 * sizeof(float)));
  CU_TRY_T(cudaMalloc(&t->colsum_scratch, static_cast<size_t>(kColsumSlabs) * 1024 * sizeof(float)));
  CU_TRY_T(cudaMalloc(&t->partial, 32LL * 512 * 512 * sizeof(float)));
  CU_TRY_T(cudaMalloc(&t->status, sizeof(unsigned int)));
  CU_TRY_T(cudaMemset(t->status, 0, sizeof(unsigned int)));
  CU_TRY_T(cudaHostAlloc(reinterpret_cast<void**>(&t->status_host), sizeof(unsigned int), cudaHostAllocMapped));
  *t->status_host = 0;
  CU_TRY_T(cudaHostGetDevicePointer(reinterpret_cast<void**>(&t->status_host_dev), t->status_host, 0));
  CU_TRY_T(cudaDeviceSynchronize());
#undef CU_TRY_T
  *out = t;
  return SDFB_OK;
}

int sdfb_decoder_trainer_get_params(sdfb_decoder_trainer* t, float* params_host) {
  if (!t || !params_host) return set_error(SDFB_E_INVALID, "null argument");
  DevGuard g(t->device);
  CU_TRY(cudaDeviceSynchronize());
  CU_TRY(cudaMemcpy(params_host, t->params, static_cast<size_t>(kDecParamFloats) * sizeof(float), cudaMemcpyDeviceToHost));
  const unsigned int s = *reinterpret_cast<volatile unsigned int*>(t->status_host);
  if (s != 0) {
    *reinterpret_cast<volatile unsigned int*>(t->status_host) = 0;
    cudaMemset(t->status, 0, sizeof(unsigned int));
    return set_error(SDFB_E_KERNEL, "a training kernel's watchdog tripped at wait site 0x%x", s);
  }
  return SDFB_OK;
}

int sdfb_decoder_trainer_step(sdfb_decoder_trainer* t, const float* latents_dev, const float* xyz_dev, const float* target_dev, int batch,
                              int64_t points_per_shape, float clamp_dist, float lr, float beta1, float beta2, float adam_eps, int apply,
                              float* loss_dev, float* grads_dev, float* sdf_dev, void* stream) {
  if (!t || !latents_dev || !xyz_dev || !target_dev || !loss_dev) return set_error(SDFB_E_INVALID, "null argument");
  if (batch <= 0 || points_per_shape <= 0) return set_error(SDFB_E_INVALID, "batch and points per shape must be positive");
  if (!(clamp_dist > 0.f)) return set_error(SDFB_E_INVALID, "clamp distance must be positive");
  const long long M = static_cast<long long>(batch) * points_per_shape;
  if (M > (1LL << 21)) return set_error(SDFB_E_INVALID, "at most 2^21 points per step (the workspace keeps 17 KiB per point)");
  DevGuard g(t->device);
  {
    const unsigned int s = *reinterpret_cast<volatile unsigned int*>(t->status_host);
    if (s != 0) {
      *reinterpret_cast<volatile unsigned int*>(t->status_host) = 0;
      cudaMemset(t->status, 0, sizeof(unsigned int));
      return set_error(SDFB_E_KERNEL, "an earlier training kernel's watchdog tripped at wait site 0x%x", s);
    }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  constexpr int L = sdfb_decoder_trainer::kL;
  if (t->ws_rows < M) {
    for (int l = 0; l <= L; ++l) { cudaFree(t->act[l]); t->act[l] = nullptr; }
    for (int l = 0; l < L; ++l) { cudaFree(t->delta[l]); t->delta[l] = nullptr; }
    cudaFree(t->y); cudaFree(t->d8); cudaFree(t->loss_partial); t->y = t->d8 = t->loss_partial = nullptr; t->ws_rows = 0;
    for (int l = 0; l <= L; ++l) CU_TRY(cudaMalloc(&t->act[l], static_cast<size_t>(M) * (l == 0 ? 320 : 512) * 2));
    for (int l = 0; l < L; ++l) CU_TRY(cudaMalloc(&t->delta[l], static_cast<size_t>(M) * t->fout_p[l] * 2));
    CU_TRY(cudaMalloc(&t->y, static_cast<size_t>(M) * sizeof(float)));
    CU_TRY(cudaMalloc(&t->d8, static_cast<size_t>(M) * sizeof(float)));
    CU_TRY(cudaMalloc(&t->loss_partial, static_cast<size_t>((M + 7) / 8 + 1) * sizeof(float)));
    t->ws_rows = M;
  }
  auto gp = [&]() {
    GemmParams p{};
    p.M = static_cast<int>(M); p.bn = 256; p.ksplit = 1; p.alpha = 1.f;
    p.status = t->status; p.status_host = t->status_host_dev; p.timeout_ns = 4000000000ull;
    return p;
  };
  const float* P = t->params;
  // ---- forward ----
  CU_TRY(launch_dec_train_input(latents_dev, xyz_dev, M, points_per_shape, 320, 0, 320, t->act[0], t->fp16, st));
  CU_TRY(cudaMemcpyAsync(t->b3pad, P + t->boff[3], 253 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  for (int l = 0; l < L; ++l) {
    GemmParams p = gp();
    p.N = t->fout_p[l]; p.K = t->fin_p[l];
    p.bias = l == 3 ? t->b3pad : P + t->boff[l];
    p.epi = kGemmEpiBiasReluLowp;
    p.out_lowp = t->act[l + 1]; p.ldo_lowp = 512;
    CU_TRY(launch_gemm_tc(p, t->act[l], t->fin_p[l], t->w_lowp[l], t->fin_p[l], false, t->fp16, t->num_sms, st));
    if (l == 3)    // the skip concat: columns 253.. of layer 4's input are the decoder input again
      CU_TRY(launch_dec_train_input(latents_dev, xyz_dev, M, points_per_shape, 512, 253, 259, t->act[4], t->fp16, st));
  }
  int nblk = 0;
  CU_TRY(launch_dec_train_head(t->act[8], P + t->woff[8], P + t->boff[8], target_dev, clamp_dist, M, sdf_dev ? sdf_dev : t->y, t->d8,
                               t->delta[7], t->loss_partial, &nblk, t->fp16, st));
  const float inv_m = 1.f / static_cast<float>(M);
  CU_TRY(launch_sum_loss(t->loss_partial, nblk, inv_m, loss_dev, st));
  // ---- backward ----
  if (apply) ++t->step;
  AdamParams ap{lr, beta1, beta2, adam_eps, 1.f - std::pow(beta1, static_cast<float>(t->step > 0 ? t->step : 1)),
                1.f - std::pow(beta2, static_cast<float>(t->step > 0 ? t->step : 1))};
  // head: dW8, db8
  CU_TRY(launch_dec_train_head_grad(t->act[8], t->d8, M, inv_m, t->small_grad, t->colsum_scratch, t->fp16, st));
  CU_TRY(launch_adam_update(t->params + t->woff[8], t->adam_m + t->woff[8], t->adam_v + t->woff[8], t->small_grad, 1, 0, 512, 1.f, 1, 512, ap,
                            nullptr, 0, nullptr, 0, grads_dev ? grads_dev + t->woff[8] : nullptr, t->fp16, apply != 0, st));
  CU_TRY(launch_adam_update(t->params + t->boff[8], t->adam_m + t->boff[8], t->adam_v + t->boff[8], t->small_grad + 512, 1, 0, 1, 1.f, 1, 1, ap,
                            nullptr, 0, nullptr, 0, grads_dev ? grads_dev + t->boff[8] : nullptr, t->fp16, apply != 0, st));
  for (int l = L - 1; l >= 0; --l) {
    if (l > 0) {   // delta_{l-1} = (delta_l W_l)[first fout_p[l-1] input columns] where that input was positive
      GemmParams p = gp();
      p.N = t->fout_p[l - 1]; p.K = t->fout_p[l];
      p.epi = kGemmEpiMaskLowp; p.mask_h = t->act[l]; p.ldh = 512;
      p.out_lowp = t->delta[l - 1]; p.ldo_lowp = t->fout_p[l - 1];
      // W_l^T [fin_p][fout_p]: its first fout_p[l-1] rows are the hidden inputs (for the skip layer: h3's 253 + 3 latent rows,
      // whose products land in padding columns 253-255 and never reach a weight: layer 3's padded weights are zero)
      CU_TRY(launch_gemm_tc(p, t->delta[l], t->fout_p[l], t->wt_lowp[l], t->fout_p[l], false, t->fp16, t->num_sms, st));
    }
    {              // dW_l [out][in] = delta_l^T a_l
      GemmParams p = gp();
      p.M = t->fout_p[l]; p.N = t->fin_p[l]; p.K = static_cast<int>(M);
      p.bn = t->fin_p[l] % 256 == 0 ? 256 : 64;
      p.ksplit = pick_ksplit(p.M, p.N, p.bn, M);
      p.epi = kGemmEpiF32; p.out_f32 = t->partial; p.ldo_f32 = t->fin_p[l];
      p.split_stride = static_cast<long long>(t->fout_p[l]) * t->fin_p[l];
      CU_TRY(launch_gemm_tc(p, t->delta[l], t->fout_p[l], t->act[l], l == 0 ? 320 : 512, true, t->fp16, t->num_sms, st));
      CU_TRY(launch_adam_update(t->params + t->woff[l], t->adam_m + t->woff[l], t->adam_v + t->woff[l], t->partial, p.ksplit, p.split_stride,
                                t->fin_p[l], inv_m, t->fout[l], t->fin[l], ap, t->w_lowp[l], t->fin_p[l], t->wt_lowp[l], t->fout_p[l],
                                grads_dev ? grads_dev + t->woff[l] : nullptr, t->fp16, apply != 0, st));
    }
    int nsl = 1;
    CU_TRY(launch_colsum_lowp(t->delta[l], M, t->fout_p[l], t->fout[l], 1.f, nullptr, t->colsum_scratch, t->fp16, st, &nsl));
    CU_TRY(launch_adam_update(t->params + t->boff[l], t->adam_m + t->boff[l], t->adam_v + t->boff[l], t->colsum_scratch, nsl, t->fout[l], 1, inv_m,
                              t->fout[l], 1, ap, nullptr, 0, nullptr, 0, grads_dev ? grads_dev + t->boff[l] : nullptr, t->fp16, apply != 0, st));
  }
  return SDFB_OK;
}

// Unit-test hook of the general product: out_dev [ksplit][M][N] fp32 = a . b^T (tn = 0: a [M][K], b [N][K]) or a^T . b (tn = 1:
// a [K][M], b [K][N]); row-major 16-bit inputs with leading dimensions lda / ldb.
int sdfb_gemm_selftest(const uint16_t* a_dev, int lda, const uint16_t* b_dev, int ldb, int M, int N, int K, int tn, int ksplit, int bn,
                       int precision, float* out_dev, void* stream) {
  if (!a_dev || !b_dev || !out_dev) return set_error(SDFB_E_INVALID, "null argument");
  if (precision != SDFB_PREC_BF16 && precision != SDFB_PREC_FP16) return set_error(SDFB_E_INVALID, "bf16 or fp16 only");
  CU_TRY(gemm_tc_init());
  int dev = 0, sms = 0;
  CU_TRY(cudaGetDevice(&dev));
  int rc = check_sm100(dev, &sms);
  if (rc) return rc;
  unsigned int* status = nullptr;
  CU_TRY(cudaMalloc(&status, sizeof(unsigned int)));
  CU_TRY(cudaMemset(status, 0, sizeof(unsigned int)));
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.bn = bn; p.ksplit = ksplit; p.epi = kGemmEpiF32; p.alpha = 1.f;
  p.out_f32 = out_dev; p.ldo_f32 = N; p.split_stride = static_cast<long long>(M) * N;
  p.status = status; p.timeout_ns = 2000000000ull;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = launch_gemm_tc(p, a_dev, lda, b_dev, ldb, tn != 0, precision == SDFB_PREC_FP16, sms, st);
  unsigned int s = 0;
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaMemcpy(&s, status, sizeof(s), cudaMemcpyDeviceToHost);
  cudaFree(status);
  if (e != cudaSuccess) return set_error(SDFB_E_CUDA, "gemm selftest: %s", cudaGetErrorString(e));
  if (s != 0) return set_error(SDFB_E_KERNEL, "gemm selftest watchdog tripped (0x%x)", s);
  return SDFB_OK;
}

}  // extern "C"

// metadata.cu — one stage of the reference's MetadataBranch (model.py:27-60):
//     y = Dropout(GELU(BatchNorm1d(x W^T + b)))          (the second stage has no dropout)
// forward and backward, including the batch statistics and the running-stat update of BatchNorm1d in
// training mode. Replaces nn.Linear (cuBLAS), nn.BatchNorm1d (ATen batch_norm kernels), nn.GELU and
// nn.Dropout's multiply of that Sequential — `metadata.enabled: true` is the reference's DEFAULT
// (config.yaml:34-40), so without this the default forward still ran on library kernels (scope row f2).
//
// Shapes are tiny (batch x 13 -> 256 -> 128): the stage is latency-, not bandwidth-bound. What shapes the
// kernel is BatchNorm's reduction over the BATCH: a CTA owns 8 output features for ALL rows, so the batch
// mean / variance (and, backward, dgamma / dbeta and the two correction sums) never leave the CTA — no
// atomics, no second launch, run-to-run deterministic. Thread = (feature tid & 7, row group tid >> 3):
// stores are 32-byte row segments, the weight tile sits transposed in shared memory, a row of x is a
// broadcast load for the 8 threads sharing it.
//
// fp32 throughout (autocast keeps batch_norm in fp32 and these GEMMs are far below tensor-core tile size).
// Variance is two-pass (mean first), like ATen's; normalisation uses the biased variance, the running
// estimate the unbiased one (x B/(B-1)), momentum as nn.BatchNorm1d.
#include "common.cuh"

namespace fv {
namespace {

constexpr int MB_THREADS = 256;
constexpr int MB_FEAT = 8;             // features per CTA
constexpr int MB_GROUPS = MB_THREADS / MB_FEAT;  // 32 row groups
constexpr int MB_MAX_K = 1024;

// sum over the 32 row groups of one value per (feature) thread; result valid in every thread
__device__ __forceinline__ float group_sum(float v, float (*red)[MB_FEAT], int fi, int rg) {
  red[rg][fi] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < MB_GROUPS; ++g) s += red[g][fi];  // fixed order: bit-reproducible
  __syncthreads();
  return s;
}

__global__ void __launch_bounds__(MB_THREADS)
linear_bn_gelu_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                          const float* __restrict__ bias, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float* __restrict__ run_mean,
                          float* __restrict__ run_var, float momentum, float eps, int training,
                          const float* __restrict__ drop_mask, float* __restrict__ y, long long ldy,
                          float* __restrict__ xhat, float* __restrict__ dact, float* __restrict__ rstd_out,
                          int batch, int K, int F) {
  pdl_wait();
  extern __shared__ float sm[];
  float* wt = sm;                                        // [K][MB_FEAT] transposed weight tile
  float(*red)[MB_FEAT] = reinterpret_cast<float(*)[MB_FEAT]>(sm + K * MB_FEAT);
  const int fi = threadIdx.x & (MB_FEAT - 1), rg = threadIdx.x >> 3;
  const int f0 = blockIdx.x * MB_FEAT;
  const int f = f0 + fi;
  const bool live = f < F;
  for (int i = threadIdx.x; i < K * MB_FEAT; i += MB_THREADS) {
    const int k = i / MB_FEAT, j = i % MB_FEAT;
    wt[i] = (f0 + j < F) ? w[static_cast<long long>(f0 + j) * K + k] : 0.f;
  }
  __syncthreads();
  const float b = (live && bias != nullptr) ? bias[f] : 0.f;
  // pass 1: h = x W^T + b, parked in the xhat buffer; batch sum
  float s = 0.f;
  for (int r = rg; r < batch; r += MB_GROUPS) {
    const float* xr = x + static_cast<long long>(r) * ldx;
    float h = b;
    for (int k = 0; k < K; ++k) h = fmaf(__ldg(xr + k), wt[k * MB_FEAT + fi], h);
    if (live) xhat[static_cast<long long>(r) * F + f] = h;
    s += h;
  }
  float mean, var;
  if (training) {
    mean = group_sum(s, red, fi, rg) / batch;
    float q = 0.f;
    if (live)
      for (int r = rg; r < batch; r += MB_GROUPS) {
        const float d = xhat[static_cast<long long>(r) * F + f] - mean;
        q = fmaf(d, d, q);
      }
    var = group_sum(q, red, fi, rg) / batch;  // biased: what the normalisation uses
    if (live && rg == 0) {
      const float unbiased = batch > 1 ? var * (static_cast<float>(batch) / (batch - 1)) : var;
      run_mean[f] = (1.f - momentum) * run_mean[f] + momentum * mean;
      run_var[f] = (1.f - momentum) * run_var[f] + momentum * unbiased;
    }
  } else {
    mean = live ? run_mean[f] : 0.f;
    var = live ? run_var[f] : 1.f;
  }
  const float rstd = rsqrtf(var + eps);
  if (live && rg == 0 && rstd_out != nullptr) rstd_out[f] = rstd;
  if (!live) return;
  const float g = gamma[f], bt = beta[f];
  for (int r = rg; r < batch; r += MB_GROUPS) {
    const long long o = static_cast<long long>(r) * F + f;
    const float xh = (xhat[o] - mean) * rstd;
    const float z = fmaf(xh, g, bt);
    const float m = drop_mask != nullptr ? drop_mask[o] : 1.f;
    xhat[o] = xh;
    y[static_cast<long long>(r) * ldy + f] = gelu_erf(z) * m;
    if (dact != nullptr) dact[o] = gelu_erf_grad(z) * m;
  }
}

// dz = dy * dact;  training: dh = gamma*rstd*(dz - mean_r(dz) - xhat * mean_r(dz * xhat));  eval: dh = gamma*rstd*dz
// dgamma += sum_r dz*xhat, dbeta += sum_r dz, dbias += sum_r dh, dW[f,k] += sum_r dh[r,f] x[r,k]
__global__ void __launch_bounds__(MB_THREADS)
linear_bn_gelu_bwd_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ x, long long ldx,
                          const float* __restrict__ xhat, const float* __restrict__ dact,
                          const float* __restrict__ rstd, const float* __restrict__ gamma, int training,
                          float* __restrict__ dh, float* __restrict__ dw, float* __restrict__ dbias,
                          float* __restrict__ dgamma, float* __restrict__ dbeta, int batch, int K, int F) {
  pdl_wait();
  __shared__ float red[MB_GROUPS][MB_FEAT];
  const int fi = threadIdx.x & (MB_FEAT - 1), rg = threadIdx.x >> 3;
  const int f0 = blockIdx.x * MB_FEAT;
  const int f = f0 + fi;
  const bool live = f < F;
  float s1 = 0.f, s2 = 0.f;
  if (live)
    for (int r = rg; r < batch; r += MB_GROUPS) {
      const long long o = static_cast<long long>(r) * F + f;
      const float dz = dy[static_cast<long long>(r) * lddy + f] * dact[o];
      s1 += dz;
      s2 = fmaf(dz, xhat[o], s2);
    }
  const float sum_dz = group_sum(s1, red, fi, rg);
  const float sum_dzx = group_sum(s2, red, fi, rg);
  const float gr = live ? gamma[f] * rstd[f] : 0.f;
  const float m1 = training ? sum_dz / batch : 0.f, m2 = training ? sum_dzx / batch : 0.f;
  float sb = 0.f;
  if (live)
    for (int r = rg; r < batch; r += MB_GROUPS) {
      const long long o = static_cast<long long>(r) * F + f;
      const float dz = dy[static_cast<long long>(r) * lddy + f] * dact[o];
      const float v = gr * (dz - m1 - xhat[o] * m2);
      dh[o] = v;
      sb += v;
    }
  const float sum_dh = group_sum(sb, red, fi, rg);  // also orders the dh stores before the reads below
  if (live && rg == 0) {
    if (dgamma != nullptr) dgamma[f] += sum_dzx;
    if (dbeta != nullptr) dbeta[f] += sum_dz;
    if (dbias != nullptr) dbias[f] += sum_dh;
  }
  if (dw == nullptr || !live) return;
  // weight gradient of this CTA's 8 features: thread = (feature, k lane); all rows, fixed order
  for (int k = rg; k < K; k += MB_GROUPS) {
    float acc = 0.f;
    for (int r = 0; r < batch; ++r)
      acc = fmaf(dh[static_cast<long long>(r) * F + f], __ldg(x + static_cast<long long>(r) * ldx + k), acc);
    dw[static_cast<long long>(f) * K + k] += acc;
  }
}

}  // namespace
}  // namespace fv

extern "C" int fv_linear_bn_gelu_fwd(const float* x, int64_t ldx, const float* w, const float* bias, const float* gamma,
                                     const float* beta, float* running_mean, float* running_var, float momentum,
                                     float eps, int training, const float* drop_mask, float* y, int64_t ldy,
                                     float* xhat, float* dact, float* rstd, int64_t batch, int64_t in_features,
                                     int64_t out_features, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(x && w && gamma && beta && running_mean && running_var && y && xhat, "fv_linear_bn_gelu_fwd: null pointer");
  FV_CHECK_ARG(batch > 0 && batch < (1 << 24) && in_features > 0 && in_features <= MB_MAX_K && out_features > 0,
               "fv_linear_bn_gelu_fwd: shape out of range (in_features <= %d)", MB_MAX_K);
  FV_CHECK_ARG(!training || batch > 1, "fv_linear_bn_gelu_fwd: BatchNorm needs more than 1 row per feature in training mode");
  FV_CHECK_ARG(ldx >= in_features && ldy >= out_features, "fv_linear_bn_gelu_fwd: leading dimensions");
  const size_t smem = (static_cast<size_t>(in_features) * MB_FEAT + MB_GROUPS * MB_FEAT) * sizeof(float);
  const unsigned grid = static_cast<unsigned>(ceil_div(out_features, MB_FEAT));
  FV_CHECK_CUDA(fv::launch_pdl(linear_bn_gelu_fwd_kernel, dim3(grid), dim3(MB_THREADS), smem,
                               static_cast<cudaStream_t>(stream), x, static_cast<long long>(ldx), w, bias, gamma, beta,
                               running_mean, running_var, momentum, eps, training, drop_mask, y,
                               static_cast<long long>(ldy), xhat, dact, rstd, (int)batch, (int)in_features,
                               (int)out_features));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int fv_linear_bn_gelu_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* xhat,
                                     const float* dact, const float* rstd, const float* gamma, int training,
                                     float* dh, float* dw, float* dbias, float* dgamma, float* dbeta, int64_t batch,
                                     int64_t in_features, int64_t out_features, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(dy && x && xhat && dact && rstd && gamma && dh, "fv_linear_bn_gelu_bwd: null pointer");
  FV_CHECK_ARG(batch > 0 && batch < (1 << 24) && in_features > 0 && out_features > 0,
               "fv_linear_bn_gelu_bwd: shape out of range");
  FV_CHECK_ARG(lddy >= out_features && ldx >= in_features, "fv_linear_bn_gelu_bwd: leading dimensions");
  const unsigned grid = static_cast<unsigned>(ceil_div(out_features, MB_FEAT));
  FV_CHECK_CUDA(fv::launch_pdl(linear_bn_gelu_bwd_kernel, dim3(grid), dim3(MB_THREADS), 0,
                               static_cast<cudaStream_t>(stream), dy, static_cast<long long>(lddy), x,
                               static_cast<long long>(ldx), xhat, dact, rstd, gamma, training, dh, dw, dbias, dgamma,
                               dbeta, (int)batch, (int)in_features, (int)out_features));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

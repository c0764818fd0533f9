// loss.cu — fused softmax + asymmetric-focal loss (forward value and d loss / d logits in one
// pass) and the plain cross-entropy used by the eval loop.
//
// Follows AsymmetricFocalLoss.forward (reference losses.py:41-67) term by term, including the
// clamps — which make the gradient piecewise: a clamp passes gradient only where it is inactive —
// and replaces the ~12 elementwise kernels + autograd graph that forward builds. F.cross_entropy
// (reference utils.py:262) is the second entry point.
//
// One CTA (the batch is at most a few thousand rows of <= 32 classes: latency-, not
// bandwidth-bound); lane c of a warp owns class c of the warp's current row. The row losses are
// reduced in a fixed order so the scalar is bit-reproducible run to run.
#include "common.cuh"

namespace fv {

constexpr int LOSS_WARPS = 8;

template <bool ASL>
__global__ void __launch_bounds__(LOSS_WARPS * 32)
loss_kernel(const float* __restrict__ logits, const long long* __restrict__ targets,
            float* __restrict__ loss, float* __restrict__ dlogits, int batch, int classes,
            float gamma_neg, float gamma_pos, float clip, float eps) {
  pdl_wait();
  __shared__ float warp_loss[LOSS_WARPS];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const bool live = lane < classes;
  const float inv_b = 1.0f / batch;
  float acc = 0.f;
  bool bad = false;
  for (int row = warp; row < batch; row += LOSS_WARPS) {
    const float z = live ? logits[row * classes + lane] : -INFINITY;
    const long long t64 = targets[row];
    // a label outside [0, classes) makes F.one_hot / nll_loss raise in the reference; a kernel cannot
    // raise, so the loss (and with it every gradient downstream) becomes NaN instead of silently
    // training on a row without a positive class
    if (t64 < 0 || t64 >= classes) bad = true;
    const int t = static_cast<int>(t64);
    const float zmax = warp_max(z);
    const float e = live ? expf(z - zmax) : 0.f;
    const float denom = warp_sum(e);
    const float p = e / denom;
    float term = 0.f;   // this class' contribution to the row loss
    float dl_dp = 0.f;  // d(row loss) / d p_c
    if (ASL) {
      if (live) {
        if (lane == t) {
          // -(1-p)^g+ * log(clamp(p, min=eps))
          const float pp = fmaxf(p, eps);
          const float om = fmaxf(1.0f - p, 0.f);
          const float w = powf(om, gamma_pos);
          const float lg = logf(pp);
          term = -w * lg;
          // d/dp (1-p)^g: torch's pow backward returns 0 for exponent 0 (and never forms 0 * inf): with
          // gamma_pos = 0 (the ASL paper's default) and a saturated softmax, powf(0, -1) = inf would
          // turn the product into NaN
          const float dw = (gamma_pos == 0.f || !(1.0f - p >= 0.f)) ? 0.f
                           : (om == 0.f && gamma_pos < 1.0f) ? 0.f
                           : -gamma_pos * powf(om, gamma_pos - 1.0f);
          const float dlg = (p >= eps) ? 1.0f / pp : 0.f;
          dl_dp = -(dw * lg + w * dlg);
        } else {
          // -p^g- * log(1 - p_neg),  p_neg = clamp(clamp(p, max=1-eps) - clip, min=eps)
          const float hi = 1.0f - eps;  // == 1.0f in fp32 for eps=1e-8, as in the reference
          float pn = fminf(p, hi);
          float dpn = (p <= hi) ? 1.0f : 0.f;
          if (clip > 0.f) {
            const float sh = pn - clip;
            if (!(sh >= eps)) dpn = 0.f;
            pn = fmaxf(sh, eps);
          }
          const float pc = fmaxf(p, 0.f);
          const float w = powf(pc, gamma_neg);
          const float lg = logf(1.0f - pn);
          term = -w * lg;
          const float dw = (gamma_neg == 0.f || !(p >= 0.f)) ? 0.f
                           : (pc == 0.f && gamma_neg < 1.0f) ? 0.f
                           : gamma_neg * powf(pc, gamma_neg - 1.0f);
          const float dlg = -dpn / (1.0f - pn);
          dl_dp = -(dw * lg + w * dlg);
        }
      }
      const float row_loss = warp_sum(term);
      acc += row_loss;
      if (dlogits != nullptr) {
        // softmax backward: dz_j = p_j * (g_j - sum_c p_c g_c)
        const float dot = warp_sum(live ? p * dl_dp : 0.f);
        if (live) dlogits[row * classes + lane] = p * (dl_dp - dot) * inv_b;
      }
    } else {
      // cross entropy: -log_softmax(z)[t]
      const float lsm = (z - zmax) - logf(denom);
      acc += warp_sum((live && lane == t) ? -lsm : 0.f);
      if (dlogits != nullptr && live)
        dlogits[row * classes + lane] = (p - (lane == t ? 1.0f : 0.f)) * inv_b;
    }
  }
  if (lane == 0) warp_loss[warp] = bad ? __int_as_float(0x7fc00000) : acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LOSS_WARPS; ++w) s += warp_loss[w];
    *loss = s * inv_b;
  }
}

}  // namespace fv

extern "C" int fv_asl_loss(const float* logits, const int64_t* targets, float* loss, float* dlogits,
                           int64_t batch, int64_t classes, float gamma_neg, float gamma_pos,
                           float clip, float eps, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(logits && targets && loss, "fv_asl_loss: null pointer");
  FV_CHECK_ARG(batch > 0 && batch < (1 << 24), "fv_asl_loss: batch=%lld out of range", (long long)batch);
  FV_CHECK_ARG(classes > 0 && classes <= 32, "fv_asl_loss: classes=%lld must be in 1..32",
               (long long)classes);
  FV_CHECK_CUDA(fv::launch_pdl(loss_kernel<true>, dim3(1), dim3(LOSS_WARPS * 32), 0, static_cast<cudaStream_t>(stream), 
      logits, reinterpret_cast<const long long*>(targets), loss, dlogits, (int)batch, (int)classes,
      gamma_neg, gamma_pos, clip, eps));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int fv_ce_loss(const float* logits, const int64_t* targets, float* loss, float* dlogits,
                          int64_t batch, int64_t classes, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(logits && targets && loss, "fv_ce_loss: null pointer");
  FV_CHECK_ARG(batch > 0 && batch < (1 << 24), "fv_ce_loss: batch=%lld out of range", (long long)batch);
  FV_CHECK_ARG(classes > 0 && classes <= 32, "fv_ce_loss: classes=%lld must be in 1..32",
               (long long)classes);
  FV_CHECK_CUDA(fv::launch_pdl(loss_kernel<false>, dim3(1), dim3(LOSS_WARPS * 32), 0, static_cast<cudaStream_t>(stream), 
      logits, reinterpret_cast<const long long*>(targets), loss, dlogits, (int)batch, (int)classes,
      0.f, 0.f, 0.f, 0.f));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

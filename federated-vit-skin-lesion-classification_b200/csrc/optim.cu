// optim.cu — HBM-bound sweeps over the flat fp32 parameter arena: gradient sum of squares,
// fused clip + AdamW (+ EMA + bf16 weight re-cast), FedAvg fold, casts.
//
// Replaces, in order: torch.nn.utils.clip_grad_norm_ (reference utils.py:192-193, called at
// train.py:157), torch.optim.AdamW.step over the LLRD groups (train.py:158; groups built at
// model.py:228-270), EMA.update (utils.py:76-83). The FedAvg fold has no reference counterpart
// (SURVEY.md F1); it implements SURVEY.md §8.2.
//
// Algorithmic bytes per parameter: sumsq 4; AdamW 28 (read p,g,m,v; write p,m,v), +8 with EMA
// (read+write shadow), +2 with the bf16 copy; FedAvg fold 12 (8 on the first client).
#include "common.cuh"

namespace fv {

constexpr int SWEEP_THREADS = 256;
constexpr int MAX_SEGMENTS = 1024;

static inline unsigned sweep_grid(int64_t nvec, int per_thread = 4) {
  int64_t want = ceil_div(nvec, static_cast<int64_t>(SWEEP_THREADS) * per_thread);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return static_cast<unsigned>(want);
}

__global__ void __launch_bounds__(SWEEP_THREADS)
sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  pdl_wait();
  __shared__ float red[SWEEP_THREADS / 32];
  const long long nvec = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  float s0 = 0.f, s1 = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += stride) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(g) + i);
    s0 += v.x * v.x + v.y * v.y;
    s1 += v.z * v.z + v.w * v.w;
  }
  float s = s0 + s1;
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float t = g[(nvec << 2) + threadIdx.x];
    s += t * t;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < SWEEP_THREADS / 32 ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}

__device__ __forceinline__ float clip_coef(const float* sumsq, float max_norm) {
  if (sumsq == nullptr || max_norm <= 0.f) return 1.0f;
  const float norm = sqrtf(*sumsq);
  return fminf(max_norm / (norm + 1e-6f), 1.0f);
}

template <bool HAS_EMA, bool HAS_LP>
__global__ void __launch_bounds__(SWEEP_THREADS)
adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
             float* __restrict__ v, const long long* __restrict__ seg_end,
             const float* __restrict__ seg_lr, const float* __restrict__ seg_wd, int nseg,
             const float* __restrict__ sumsq, float max_norm, float beta1, float beta2, float eps,
             float bc1, float bc2_sqrt, float* __restrict__ ema, float ema_decay,
             __nv_bfloat16* __restrict__ p_lp, long long nvec, const float* __restrict__ bc_dev,
             int zero_grad) {
  pdl_wait();
  if (bc_dev != nullptr) {  // captured in a CUDA graph: the step-dependent bias corrections live in device memory
    bc1 = __ldg(bc_dev);
    bc2_sqrt = __ldg(bc_dev + 1);
  }
  __shared__ long long s_end[MAX_SEGMENTS];
  __shared__ float s_lr[MAX_SEGMENTS];
  __shared__ float s_wd[MAX_SEGMENTS];
  for (int i = threadIdx.x; i < nseg; i += blockDim.x) {
    s_end[i] = seg_end[i];
    s_lr[i] = seg_lr[i];
    s_wd[i] = seg_wd[i];
  }
  __syncthreads();
  const float coef = clip_coef(sumsq, max_norm);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += stride) {
    const long long e0 = i << 2;
    // first segment whose end is beyond this element (segments are 4-aligned by contract)
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_end[mid] > e0) hi = mid; else lo = mid + 1;
    }
    const float lr = s_lr[lo], wd = s_wd[lo];
    float4 pv = reinterpret_cast<float4*>(p)[i];
    // zero_grad: this sweep is the gradient's last reader, so it leaves the buffer zeroed for the next
    // step's accumulating kernels (optimizer.zero_grad, reference train.py:160) — the separate memset
    // pass over the 345 MB gradient arena disappears. Un-optimised ranges (lr < 0) are zeroed too.
    if (zero_grad && lr < 0.f) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lr >= 0.f) {
      const float4 gv = __ldcs(reinterpret_cast<const float4*>(g) + i);
      float4 mv = reinterpret_cast<float4*>(m)[i];
      float4 vv = reinterpret_cast<float4*>(v)[i];
      const float decay = 1.0f - lr * wd;
      const float step_size = lr / bc1;
#define FV_ADAM_ONE(P, G, M, V)                                 \
  {                                                             \
    const float gg = (G) * coef;                                \
    (P) *= decay;                                               \
    (M) = (M) + (gg - (M)) * (1.0f - beta1);                    \
    (V) = (V) * beta2 + gg * gg * (1.0f - beta2);               \
    const float den = sqrtf(V) / bc2_sqrt + eps;                \
    (P) = (P) - step_size * ((M) / den);                        \
  }
      FV_ADAM_ONE(pv.x, gv.x, mv.x, vv.x)
      FV_ADAM_ONE(pv.y, gv.y, mv.y, vv.y)
      FV_ADAM_ONE(pv.z, gv.z, mv.z, vv.z)
      FV_ADAM_ONE(pv.w, gv.w, mv.w, vv.w)
#undef FV_ADAM_ONE
      reinterpret_cast<float4*>(p)[i] = pv;
      reinterpret_cast<float4*>(m)[i] = mv;
      reinterpret_cast<float4*>(v)[i] = vv;
      // the zeros go out with the other stores: written right behind the gradient load they kept the m / v
      // loads of the iteration behind them and the sweep ran at 1.7 TB/s instead of 6 (1 708 vs 453 us for the
      // ViT-B arena, tools/adamw_probe.py)
      if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (HAS_EMA) {
      float4 sv = reinterpret_cast<float4*>(ema)[i];
      const float w = 1.0f - ema_decay;
      sv.x = sv.x * ema_decay + pv.x * w;
      sv.y = sv.y * ema_decay + pv.y * w;
      sv.z = sv.z * ema_decay + pv.z * w;
      sv.w = sv.w * ema_decay + pv.w * w;
      reinterpret_cast<float4*>(ema)[i] = sv;
    }
    if (HAS_LP) {
      uint2 pk;
      pk.x = pack_bf16(pv.x, pv.y);
      pk.y = pack_bf16(pv.z, pv.w);
      reinterpret_cast<uint2*>(p_lp)[i] = pk;
    }
  }
}

__global__ void __launch_bounds__(SWEEP_THREADS)
scale_kernel(float* __restrict__ x, const float* __restrict__ sumsq, float max_norm, long long nvec) {
  pdl_wait();
  const float coef = clip_coef(sumsq, max_norm);
  if (coef == 1.0f) return;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += stride) {
    float4 v = reinterpret_cast<float4*>(x)[i];
    v.x *= coef; v.y *= coef; v.z *= coef; v.w *= coef;
    reinterpret_cast<float4*>(x)[i] = v;
  }
}

__global__ void __launch_bounds__(SWEEP_THREADS)
ema_kernel(float* __restrict__ s, const float* __restrict__ p, float decay, long long nvec) {
  pdl_wait();
  const float w = 1.0f - decay;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += stride) {
    float4 sv = reinterpret_cast<float4*>(s)[i];
    const float4 pv = __ldcs(reinterpret_cast<const float4*>(p) + i);
    sv.x = sv.x * decay + pv.x * w;
    sv.y = sv.y * decay + pv.y * w;
    sv.z = sv.z * decay + pv.z * w;
    sv.w = sv.w * decay + pv.w * w;
    reinterpret_cast<float4*>(s)[i] = sv;
  }
}

template <bool INIT>
__global__ void __launch_bounds__(SWEEP_THREADS)
fedavg_kernel(float* __restrict__ acc, const float* __restrict__ w, float weight, long long nvec) {
  pdl_wait();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += stride) {
    const float4 wv = __ldcs(reinterpret_cast<const float4*>(w) + i);
    float4 a;
    if (INIT) {
      // multiply only: the fold of the first client is exactly weight * w (no fused add of 0)
      a = make_float4(weight * wv.x, weight * wv.y, weight * wv.z, weight * wv.w);
    } else {
      a = reinterpret_cast<float4*>(acc)[i];
      // separate multiply and add (no FMA contraction) so the oracle's fp32 order is reproduced
      a.x = __fadd_rn(a.x, __fmul_rn(weight, wv.x));
      a.y = __fadd_rn(a.y, __fmul_rn(weight, wv.y));
      a.z = __fadd_rn(a.z, __fmul_rn(weight, wv.z));
      a.w = __fadd_rn(a.w, __fmul_rn(weight, wv.w));
    }
    reinterpret_cast<float4*>(acc)[i] = a;
  }
}

// out = (acc ? acc : 0) + weight * w with the oracle's rounding (one rounded multiply, one rounded add);
// out may alias w or acc. The LAST fold of a round writes straight into the parameter arena — the buffer
// the allreduce then runs on in place — instead of into the accumulator followed by a 345 MB copy, and
// (single rank) emits the bf16 shadow of the new global weights in the same pass.
template <bool HAS_ACC, bool HAS_LP>
__global__ void __launch_bounds__(SWEEP_THREADS)
fedavg_into_kernel(const float* acc, const float* w, float weight, float* out, __nv_bfloat16* out_lp, long long nvec) {
  pdl_wait();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += stride) {
    const float4 wv = reinterpret_cast<const float4*>(w)[i];
    float4 a;
    if (HAS_ACC) {
      a = __ldcs(reinterpret_cast<const float4*>(acc) + i);
      a.x = __fadd_rn(a.x, __fmul_rn(weight, wv.x));
      a.y = __fadd_rn(a.y, __fmul_rn(weight, wv.y));
      a.z = __fadd_rn(a.z, __fmul_rn(weight, wv.z));
      a.w = __fadd_rn(a.w, __fmul_rn(weight, wv.w));
    } else {
      a = make_float4(weight * wv.x, weight * wv.y, weight * wv.z, weight * wv.w);
    }
    reinterpret_cast<float4*>(out)[i] = a;
    if (HAS_LP) {
      uint2 pk;
      pk.x = pack_bf16(a.x, a.y);
      pk.y = pack_bf16(a.z, a.w);
      reinterpret_cast<uint2*>(out_lp)[i] = pk;
    }
  }
}

__global__ void __launch_bounds__(SWEEP_THREADS)
cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long nvec) {
  pdl_wait();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += stride) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(src) + i);
    uint2 pk;
    pk.x = pack_bf16(v.x, v.y);
    pk.y = pack_bf16(v.z, v.w);
    reinterpret_cast<uint2*>(dst)[i] = pk;
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace fv

extern "C" int fv_sumsq(const float* g, int64_t n, float* sumsq, int accumulate, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(g && sumsq && n >= 0, "fv_sumsq: bad argument");
  FV_CHECK_ARG(aligned16(g), "fv_sumsq: g must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!accumulate) FV_CHECK_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(float), st));
  if (n == 0) return FV_OK;
  FV_CHECK_CUDA(fv::launch_pdl(sumsq_kernel, dim3(sweep_grid(n >> 2, 8)), dim3(SWEEP_THREADS), 0, st, g, n, sumsq));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

namespace fv {
static int adamw_flat_impl(float* p, float* g, float* m, float* v, const int64_t* seg_end,
                           const float* seg_lr, const float* seg_wd, int nseg, const float* sumsq,
                           float max_norm, float beta1, float beta2, float eps, int64_t step,
                           const float* bias_corr, float* ema, float ema_decay, void* p_lp, int64_t n,
                           int zero_grad, void* stream) {
  FV_CHECK_ARG(p && g && m && v && seg_end && seg_lr && seg_wd, "fv_adamw_flat: null pointer");
  FV_CHECK_ARG(nseg > 0 && nseg <= MAX_SEGMENTS, "fv_adamw_flat: nseg=%d out of range", nseg);
  FV_CHECK_ARG(n > 0 && n % 4 == 0, "fv_adamw_flat: n=%lld must be a positive multiple of 4", (long long)n);
  FV_CHECK_ARG(bias_corr != nullptr || step >= 1, "fv_adamw_flat: step must be >= 1");
  FV_CHECK_ARG(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v) &&
                   (!ema || aligned16(ema)) && (!p_lp || (reinterpret_cast<uintptr_t>(p_lp) & 7) == 0),
               "fv_adamw_flat: arenas must be 16-byte aligned");
  float fbc1 = 1.f, fbc2s = 1.f;
  if (bias_corr == nullptr) {
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
    fbc1 = static_cast<float>(bc1);
    fbc2s = static_cast<float>(sqrt(bc2));
  }
  const long long nvec = n >> 2;
  const unsigned grid = sweep_grid(nvec, 2);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long* se = reinterpret_cast<const long long*>(seg_end);
  __nv_bfloat16* lp = reinterpret_cast<__nv_bfloat16*>(p_lp);
#define FV_ADAM_LAUNCH(E, L)                                                                      \
  FV_CHECK_CUDA(fv::launch_pdl(adamw_kernel<E, L>, dim3(grid), dim3(SWEEP_THREADS), 0, st, p, g, m, v, se, seg_lr, seg_wd, nseg, sumsq, \
                                                     max_norm, beta1, beta2, eps, fbc1, fbc2s,   \
                                                     ema, ema_decay, lp, nvec, bias_corr, zero_grad))
  if (ema && lp) FV_ADAM_LAUNCH(true, true);
  else if (ema) FV_ADAM_LAUNCH(true, false);
  else if (lp) FV_ADAM_LAUNCH(false, true);
  else FV_ADAM_LAUNCH(false, false);
#undef FV_ADAM_LAUNCH
  FV_LAUNCH_CHECK();
  return FV_OK;
}
}  // namespace fv

extern "C" int fv_adamw_flat(float* p, float* g, float* m, float* v, const int64_t* seg_end,
                             const float* seg_lr, const float* seg_wd, int nseg, const float* sumsq,
                             float max_norm, float beta1, float beta2, float eps, int64_t step,
                             float* ema, float ema_decay, void* p_lp, int64_t n, int zero_grad, void* stream) {
  return fv::adamw_flat_impl(p, g, m, v, seg_end, seg_lr, seg_wd, nseg, sumsq, max_norm, beta1, beta2, eps, step,
                             nullptr, ema, ema_decay, p_lp, n, zero_grad, stream);
}

// Same sweep with the two step-dependent scalars read from device memory — bias_corr[0] = 1 - beta1^t,
// bias_corr[1] = sqrt(1 - beta2^t) — so a launch captured in a CUDA graph stays valid for every step t.
extern "C" int fv_adamw_flat_dev(float* p, float* g, float* m, float* v, const int64_t* seg_end,
                                 const float* seg_lr, const float* seg_wd, int nseg, const float* sumsq,
                                 float max_norm, float beta1, float beta2, float eps, const float* bias_corr,
                                 float* ema, float ema_decay, void* p_lp, int64_t n, int zero_grad, void* stream) {
  FV_CHECK_ARG(bias_corr != nullptr, "fv_adamw_flat_dev: bias_corr is NULL");
  return fv::adamw_flat_impl(p, g, m, v, seg_end, seg_lr, seg_wd, nseg, sumsq, max_norm, beta1, beta2, eps, 0,
                             bias_corr, ema, ema_decay, p_lp, n, zero_grad, stream);
}

// Device-side step counter of the graph-replayed optimiser: *step += 1 and the two bias corrections of
// that step into bias_corr[0..1], one thread. Captured in the CUDA graph right before the sweep, so
// however far the host runs ahead of the device every replay sees exactly its own step (the host-written
// pinned ring this replaces could be overwritten before its copy executed).
namespace fv {
__global__ void adamw_tick_kernel(long long* step, float* bias_corr, float beta1, float beta2) {
  pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const long long t = *step + 1;
    *step = t;
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(t));
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(t));
    bias_corr[0] = static_cast<float>(bc1);
    bias_corr[1] = static_cast<float>(sqrt(bc2));
  }
}
}  // namespace fv

extern "C" int fv_adamw_tick(int64_t* step, float* bias_corr, float beta1, float beta2, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(step && bias_corr, "fv_adamw_tick: null pointer");
  FV_CHECK_CUDA(fv::launch_pdl(adamw_tick_kernel, dim3(1), dim3(32), 0, static_cast<cudaStream_t>(stream),
                               reinterpret_cast<long long*>(step), bias_corr, beta1, beta2));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int fv_scale_inplace(float* x, const float* sumsq, float max_norm, int64_t n, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(x && sumsq && n >= 0 && n % 4 == 0 && aligned16(x), "fv_scale_inplace: bad argument");
  if (n == 0) return FV_OK;
  FV_CHECK_CUDA(fv::launch_pdl(scale_kernel, dim3(sweep_grid(n >> 2)), dim3(SWEEP_THREADS), 0, static_cast<cudaStream_t>(stream), 
      x, sumsq, max_norm, n >> 2));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int fv_ema_update(float* shadow, const float* p, float decay, int64_t n, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(shadow && p && n >= 0 && n % 4 == 0 && aligned16(shadow) && aligned16(p),
               "fv_ema_update: bad argument");
  if (n == 0) return FV_OK;
  FV_CHECK_CUDA(fv::launch_pdl(ema_kernel, dim3(sweep_grid(n >> 2)), dim3(SWEEP_THREADS), 0, static_cast<cudaStream_t>(stream), 
      shadow, p, decay, n >> 2));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int fv_fedavg_accum(float* acc, const float* w, float weight, int init, int64_t n,
                               void* stream) {
  using namespace fv;
  FV_CHECK_ARG(acc && w && n >= 0 && n % 4 == 0 && aligned16(acc) && aligned16(w),
               "fv_fedavg_accum: bad argument");
  if (n == 0) return FV_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (init) FV_CHECK_CUDA(fv::launch_pdl(fedavg_kernel<true>, dim3(sweep_grid(n >> 2)), dim3(SWEEP_THREADS), 0, st, acc, w, weight, n >> 2));
  else FV_CHECK_CUDA(fv::launch_pdl(fedavg_kernel<false>, dim3(sweep_grid(n >> 2)), dim3(SWEEP_THREADS), 0, st, acc, w, weight, n >> 2));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int fv_fedavg_fold_into(const float* acc, const float* w, float weight, float* out, void* out_lp,
                                   int64_t n, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(w && out && n >= 0 && n % 4 == 0 && aligned16(w) && aligned16(out) && (!acc || aligned16(acc)) &&
                   (!out_lp || (reinterpret_cast<uintptr_t>(out_lp) & 7) == 0),
               "fv_fedavg_fold_into: bad argument");
  if (n == 0) return FV_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* lp = reinterpret_cast<__nv_bfloat16*>(out_lp);
  const dim3 grid(sweep_grid(n >> 2)), block(SWEEP_THREADS);
  if (acc && lp) FV_CHECK_CUDA(fv::launch_pdl(fedavg_into_kernel<true, true>, grid, block, 0, st, acc, w, weight, out, lp, n >> 2));
  else if (acc) FV_CHECK_CUDA(fv::launch_pdl(fedavg_into_kernel<true, false>, grid, block, 0, st, acc, w, weight, out, lp, n >> 2));
  else if (lp) FV_CHECK_CUDA(fv::launch_pdl(fedavg_into_kernel<false, true>, grid, block, 0, st, acc, w, weight, out, lp, n >> 2));
  else FV_CHECK_CUDA(fv::launch_pdl(fedavg_into_kernel<false, false>, grid, block, 0, st, acc, w, weight, out, lp, n >> 2));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int fv_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(src && dst && n >= 0 && n % 4 == 0 && aligned16(src) &&
                   (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
               "fv_cast_f32_bf16: bad argument");
  if (n == 0) return FV_OK;
  FV_CHECK_CUDA(fv::launch_pdl(cast_bf16_kernel, dim3(sweep_grid(n >> 2)), dim3(SWEEP_THREADS), 0, static_cast<cudaStream_t>(stream), 
      src, reinterpret_cast<__nv_bfloat16*>(dst), n >> 2));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

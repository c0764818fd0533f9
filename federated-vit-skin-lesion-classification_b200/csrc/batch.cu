// batch.cu — device-side batch assembly (scope row f3): the per-sample tensor work the reference does
// on the host in its Dataset and the batch mixing it does with a handful of ATen kernels.
//
//   fv_assemble_batch  uint8 image (+ uint8 lesion mask) -> normalised fp32 NCHW, 3 or 4 channels
//                      = TF.to_tensor + TF.normalize(IMAGENET_MEAN, IMAGENET_STD) + (mask-0.5)/0.5 +
//                        torch.cat([img, mask], 0)  (reference data.py:148-155, 222-224), so a batch
//                        crosses PCIe as bytes (38.5 MB instead of 154 MB for 256 x 224 x 224 x 3)
//   fv_mix_batch       MixUp / CutMix on an fp32 NCHW batch (reference utils.py:112-150) in one
//                      pass: out = lam*x + (1-lam)*x[perm]  or  x with the box taken from x[perm]
//
// Both are pure HBM streams: one thread moves 4 pixels of one channel plane row (128-bit stores),
// the uint8 side is read as 32-bit words (NCHW) or 12-byte pixel quads (NHWC).
#include "common.cuh"

namespace fv {

struct MixSpec {
  const long long* perm;  // [B] partner sample, or nullptr
  float lam, one_minus_lam;
  int mode;               // 0 none, 1 mixup, 2 cutmix
  int x1, y1, x2, y2;     // cutmix box: rows [x1, x2) of dim 2, columns [y1, y2) of dim 3 (the reference's naming)
};

// mixup arithmetic exactly as torch evaluates lam * a + (1 - lam) * b on fp32 tensors: two rounded
// products and a rounded sum (no fused multiply-add), so the result is bit-identical
__device__ __forceinline__ float mix2(float a, float b, float lam, float oml) {
  return __fadd_rn(__fmul_rn(lam, a), __fmul_rn(oml, b));
}

__global__ void __launch_bounds__(256)
mix_batch_kernel(const float* __restrict__ x, float* __restrict__ out, MixSpec m, int batch, int chans, int height,
                 int width) {
  pdl_wait();
  const int wq = width >> 2;
  const long long total = static_cast<long long>(batch) * chans * height * wq;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(i % wq);
    const long long r = i / wq;  // (b, c, row)
    const int row = static_cast<int>(r % height);
    const long long bc = r / height;
    const int b = static_cast<int>(bc / chans);
    const int c = static_cast<int>(bc - static_cast<long long>(b) * chans);
    const float4 a = __ldcs(reinterpret_cast<const float4*>(x) + i);
    float4 o = a;
    if (m.mode != 0) {
      const long long pb = m.perm[b];
      const float4 p = __ldg(reinterpret_cast<const float4*>(x) + ((pb * chans + c) * height + row) * wq + q);
      if (m.mode == 1) {
        o = make_float4(mix2(a.x, p.x, m.lam, m.one_minus_lam), mix2(a.y, p.y, m.lam, m.one_minus_lam),
                        mix2(a.z, p.z, m.lam, m.one_minus_lam), mix2(a.w, p.w, m.lam, m.one_minus_lam));
      } else if (row >= m.x1 && row < m.x2) {
        const int col = q << 2;
        if (col + 0 >= m.y1 && col + 0 < m.y2) o.x = p.x;
        if (col + 1 >= m.y1 && col + 1 < m.y2) o.y = p.y;
        if (col + 2 >= m.y1 && col + 2 < m.y2) o.z = p.z;
        if (col + 3 >= m.y1 && col + 3 < m.y2) o.w = p.w;
      }
    }
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

// one normalised pixel value: (u/255 - mean) / std evaluated as torch does (to_tensor divides by
// 255 in fp32, normalize subtracts then divides)
__device__ __forceinline__ float norm_px(uint32_t u, float mean, float stdv) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(u), 255.0f), mean), stdv);
}
__device__ __forceinline__ float mask_px(uint32_t u) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(u), 255.0f), 0.5f), 0.5f);
}

struct NormSpec {
  float mean[3], stdv[3];
};

template <bool NHWC>
__device__ __forceinline__ float4 load_norm4(const uint8_t* img, const uint8_t* mask, const NormSpec& n, long long b,
                                             int c, int row, int q, int height, int width) {
  const long long px0 = (b * height + row) * static_cast<long long>(width) + (q << 2);  // pixel index in a plane
  if (c == 3) {
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(mask + px0));
    return make_float4(mask_px(w & 255u), mask_px((w >> 8) & 255u), mask_px((w >> 16) & 255u), mask_px(w >> 24));
  }
  if (NHWC) {
    const uint8_t* p = img + px0 * 3 + c;
    return make_float4(norm_px(p[0], n.mean[c], n.stdv[c]), norm_px(p[3], n.mean[c], n.stdv[c]),
                       norm_px(p[6], n.mean[c], n.stdv[c]), norm_px(p[9], n.mean[c], n.stdv[c]));
  }
  const long long plane = static_cast<long long>(height) * width;
  const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(img + (b * 3 + c) * plane +
                                                              static_cast<long long>(row) * width + (q << 2)));
  return make_float4(norm_px(w & 255u, n.mean[c], n.stdv[c]), norm_px((w >> 8) & 255u, n.mean[c], n.stdv[c]),
                     norm_px((w >> 16) & 255u, n.mean[c], n.stdv[c]), norm_px(w >> 24, n.mean[c], n.stdv[c]));
}

template <bool NHWC>
__global__ void __launch_bounds__(256)
assemble_batch_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask, float* __restrict__ out,
                      NormSpec n, MixSpec m, int batch, int chans, int height, int width) {
  pdl_wait();
  const int wq = width >> 2;
  const long long total = static_cast<long long>(batch) * chans * height * wq;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(i % wq);
    const long long r = i / wq;
    const int row = static_cast<int>(r % height);
    const long long bc = r / height;
    const int b = static_cast<int>(bc / chans);
    const int c = static_cast<int>(bc - static_cast<long long>(b) * chans);
    float4 o = load_norm4<NHWC>(img, mask, n, b, c, row, q, height, width);
    if (m.mode != 0) {
      const float4 p = load_norm4<NHWC>(img, mask, n, m.perm[b], c, row, q, height, width);
      if (m.mode == 1) {
        o = make_float4(mix2(o.x, p.x, m.lam, m.one_minus_lam), mix2(o.y, p.y, m.lam, m.one_minus_lam),
                        mix2(o.z, p.z, m.lam, m.one_minus_lam), mix2(o.w, p.w, m.lam, m.one_minus_lam));
      } else if (row >= m.x1 && row < m.x2) {
        const int col = q << 2;
        if (col + 0 >= m.y1 && col + 0 < m.y2) o.x = p.x;
        if (col + 1 >= m.y1 && col + 1 < m.y2) o.y = p.y;
        if (col + 2 >= m.y1 && col + 2 < m.y2) o.z = p.z;
        if (col + 3 >= m.y1 && col + 3 < m.y2) o.w = p.w;
      }
    }
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

static int check_mix(const int64_t* perm, int mode, const char* who) {
  FV_CHECK_ARG(mode >= 0 && mode <= 2, "%s: mode must be 0 (none), 1 (mixup) or 2 (cutmix)", who);
  FV_CHECK_ARG(mode == 0 || perm != nullptr, "%s: mixing needs the permutation", who);
  return FV_OK;
}

static unsigned stream_grid(long long total) {
  const long long want = ceil_div(total, 256);
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<unsigned>(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace fv

extern "C" int fv_mix_batch(const float* x, const int64_t* perm, float lam, float one_minus_lam, int mode, int x1,
                            int y1, int x2, int y2, float* out, int64_t batch, int64_t chans, int64_t height,
                            int64_t width, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(x && out && batch > 0 && chans > 0 && height > 0 && width > 0 && width % 4 == 0,
               "fv_mix_batch: null pointer, empty batch or width %% 4 != 0");
  FV_CHECK_ARG(x != out || mode == 0, "fv_mix_batch: mixing reads partner samples — it cannot run in place");
  FV_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "fv_mix_batch: pointers must be 16-byte aligned");
  int rc = check_mix(perm, mode, "fv_mix_batch");
  if (rc != FV_OK) return rc;
  MixSpec m{reinterpret_cast<const long long*>(perm), lam, one_minus_lam, mode, x1, y1, x2, y2};
  const long long total = batch * chans * height * (width / 4);
  FV_CHECK_CUDA(fv::launch_pdl(mix_batch_kernel, dim3(stream_grid(total)), dim3(256), 0, static_cast<cudaStream_t>(stream),
                               x, out, m, (int)batch, (int)chans, (int)height, (int)width));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int fv_assemble_batch(const uint8_t* img, int nhwc, const uint8_t* mask, const float* mean, const float* stdv,
                                 const int64_t* perm, float lam, float one_minus_lam, int mode, int x1, int y1, int x2,
                                 int y2, float* out, int64_t batch, int64_t height, int64_t width, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(img && mean && stdv && out && batch > 0 && height > 0 && width > 0 && width % 4 == 0,
               "fv_assemble_batch: null pointer, empty batch or width %% 4 != 0");
  FV_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (nhwc || (reinterpret_cast<uintptr_t>(img) & 3) == 0) &&
                   (mask == nullptr || (reinterpret_cast<uintptr_t>(mask) & 3) == 0),
               "fv_assemble_batch: out must be 16-byte, planar uint8 inputs 4-byte aligned");
  int rc = check_mix(perm, mode, "fv_assemble_batch");
  if (rc != FV_OK) return rc;
  NormSpec n;
  for (int i = 0; i < 3; ++i) {
    n.mean[i] = mean[i];   // host pointers: three floats each, read here
    n.stdv[i] = stdv[i];
    FV_CHECK_ARG(n.stdv[i] > 0.f, "fv_assemble_batch: std must be positive");
  }
  MixSpec m{reinterpret_cast<const long long*>(perm), lam, one_minus_lam, mode, x1, y1, x2, y2};
  const int chans = mask != nullptr ? 4 : 3;
  const long long total = batch * chans * height * (width / 4);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (nhwc)
    FV_CHECK_CUDA(fv::launch_pdl(assemble_batch_kernel<true>, dim3(stream_grid(total)), dim3(256), 0, st, img, mask, out, n,
                                 m, (int)batch, chans, (int)height, (int)width));
  else
    FV_CHECK_CUDA(fv::launch_pdl(assemble_batch_kernel<false>, dim3(stream_grid(total)), dim3(256), 0, st, img, mask, out,
                                 n, m, (int)batch, chans, (int)height, (int)width));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

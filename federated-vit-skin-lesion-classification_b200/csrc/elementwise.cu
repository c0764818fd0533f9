// elementwise.cu — the small HBM-bound helpers around the GEMMs: patch gather, cls/pos rows,
// column sums (bias / pos_embed / cls_token gradients) and the fp32 row softmax of the parity path.
//
// Reference call sites: timm PatchEmbed (Conv2d k=16,s=16 -> flatten -> transpose), the
// cat(cls_token, x) + pos_embed prologue of VisionTransformer.forward, and the bias terms of
// every nn.Linear — all reached from model.py:193; their gradients from train.py:153.
#include "common.cuh"

namespace fv {

// out[(b, ph, pw), (c, py, px)] = img[b, c, ph*16+py, pw*16+px]; one thread moves 4 px.
template <bool OUT_BF16>
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ img, void* __restrict__ out, int batch, int chans,
                int height, int width, int lead) {
  pdl_wait();
  const int gw = width >> 4, gh = height >> 4;
  const int kdim = chans * 256;
  const int rows_per_img = gh * gw + lead;  // `lead` zero rows in front of every image's patches (the cls slot)
  const long long total = static_cast<long long>(batch) * rows_per_img * (kdim >> 2);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kq = static_cast<int>(i % (kdim >> 2));
    const long long row = i / (kdim >> 2);
    const int tok = static_cast<int>(row % rows_per_img) - lead;
    const int b = static_cast<int>(row / rows_per_img);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tok >= 0) {
      const int pw = tok % gw;
      const int ph = tok / gw;
      const int px = (kq & 3) << 2;
      const int py = (kq >> 2) & 15;
      const int c = kq >> 6;
      v = __ldcs(reinterpret_cast<const float4*>(
          img + ((static_cast<long long>(b) * chans + c) * height + (ph * 16 + py)) * width + pw * 16 + px));
    }
    if (OUT_BF16) {
      uint2 pk;
      pk.x = pack_bf16(v.x, v.y);
      pk.y = pack_bf16(v.z, v.w);
      reinterpret_cast<uint2*>(out)[i] = pk;
    } else {
      reinterpret_cast<float4*>(out)[i] = v;
    }
  }
}

// x[b, 0, :] = cls + pos[0, :]
__global__ void __launch_bounds__(256)
cls_pos_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ x,
               int batch, long long tokens, int dim) {
  pdl_wait();
  const int total = batch * dim;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / dim, d = i - b * dim;
    x[static_cast<long long>(b) * tokens * dim + d] = cls[d] + pos[d];
  }
}

// Head of the backward: the loss reaches the backbone through the cls token only (model.py:193 pools
// timm's global_pool='token'), so the gradient entering the last block is dcls in row 0 of every image
// and zero elsewhere. One pass writes both forms the last block's backward reads — the fp32 residual-
// path gradient and its bf16 copy times the block's stochastic-depth factor — instead of
// zeros + slice-assign + clone + cast (four passes over [B*N, D]).
__global__ void __launch_bounds__(256)
cls_grad_rows_kernel(const float* __restrict__ dcls, const float* __restrict__ row_scale, float* __restrict__ dx,
                     void* __restrict__ dy, int dy_bf16, int batch, int tokens, int dim) {
  pdl_wait();
  const int dq = dim >> 2;
  const long long total = static_cast<long long>(batch) * tokens * dq;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % dq);
    const long long row = i / dq;
    const int tok = static_cast<int>(row % tokens);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), s = v;
    if (tok == 0) {
      const int b = static_cast<int>(row / tokens);
      v = reinterpret_cast<const float4*>(dcls)[static_cast<long long>(b) * dq + c];
      const float f = row_scale != nullptr ? row_scale[b] : 1.0f;
      s = make_float4(v.x * f, v.y * f, v.z * f, v.w * f);
    }
    if (dx != nullptr) __stcs(reinterpret_cast<float4*>(dx) + i, v);
    if (dy != nullptr) {
      if (dy_bf16) {
        uint2 pk;
        pk.x = pack_bf16(s.x, s.y);
        pk.y = pack_bf16(s.z, s.w);
        reinterpret_cast<uint2*>(dy)[i] = pk;
      } else {
        reinterpret_cast<float4*>(dy)[i] = s;
      }
    }
  }
}

// out[c] += sum_r a[r, c].  block = 32 x 8; a warp row covers 64 columns (2 per lane).
// VEC2 = false: scalar loads for odd widths / leading dimensions (the 7-class head's bias gradient).
template <bool IN_BF16, bool VEC2 = true>
__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ a, long long lda, float* __restrict__ out, long long rows,
              int cols, int rows_per_block) {
  pdl_wait();
  __shared__ float2 red[8][32];
  const int col = blockIdx.x * 64 + threadIdx.x * 2;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float2 s = make_float2(0.f, 0.f);
  if (col < cols) {
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
      if (!VEC2) {
        const bool two = col + 1 < cols;
        if (IN_BF16) {
          const __nv_bfloat16* row = reinterpret_cast<const __nv_bfloat16*>(a) + r * lda;
          s.x += __bfloat162float(row[col]);
          if (two) s.y += __bfloat162float(row[col + 1]);
        } else {
          const float* row = reinterpret_cast<const float*>(a) + r * lda;
          s.x += row[col];
          if (two) s.y += row[col + 1];
        }
      } else if (IN_BF16) {
        const uint32_t u = __ldcs(reinterpret_cast<const uint32_t*>(
            reinterpret_cast<const __nv_bfloat16*>(a) + r * lda + col));
        const float2 f = unpack_bf16(u);
        s.x += f.x;
        s.y += f.y;
      } else {
        const float2 f = __ldcs(reinterpret_cast<const float2*>(reinterpret_cast<const float*>(a) + r * lda + col));
        s.x += f.x;
        s.y += f.y;
      }
    }
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && col < cols) {
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      s.x += red[j][threadIdx.x].x;
      s.y += red[j][threadIdx.x].y;
    }
    atomicAdd(out + col, s.x);
    if (col + 1 < cols) atomicAdd(out + col + 1, s.y);
  }
}

// p = softmax(scale * s) row-wise; one warp per row (fp32 parity path of the attention core)
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ s, float* __restrict__ p, long long rows, int cols,
                    float scale) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* sr = s + row * cols;
  float* pr = p + row * cols;
  float mx = -INFINITY;
  for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, sr[c] * scale);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < cols; c += 32) sum += expf(sr[c] * scale - mx);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int c = lane; c < cols; c += 32) pr[c] = expf(sr[c] * scale - mx) * inv;
}

// ds = scale * p * (dp - sum(dp * p))
__global__ void __launch_bounds__(256)
softmax_rows_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp,
                        float* __restrict__ ds, long long rows, int cols, float scale) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* pr = p + row * cols;
  const float* dr = dp + row * cols;
  float dot = 0.f;
  for (int c = lane; c < cols; c += 32) dot += pr[c] * dr[c];
  dot = warp_sum(dot);
  for (int c = lane; c < cols; c += 32) ds[row * cols + c] = scale * pr[c] * (dr[c] - dot);
}

}  // namespace fv

extern "C" int fv_patchify(const float* img, void* out, int out_dtype, int64_t batch, int64_t chans,
                           int64_t height, int64_t width, void* stream) {
  return fv_patchify_rows(img, out, out_dtype, batch, chans, height, width, 0, stream);
}

extern "C" int fv_patchify_rows(const float* img, void* out, int out_dtype, int64_t batch, int64_t chans,
                                int64_t height, int64_t width, int64_t lead_rows, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(lead_rows >= 0 && lead_rows <= 16, "fv_patchify_rows: lead_rows out of range");
  FV_CHECK_ARG(img && out, "fv_patchify: null pointer");
  FV_CHECK_ARG(batch > 0 && chans > 0 && height > 0 && width > 0 && height % 16 == 0 && width % 16 == 0,
               "fv_patchify: image %lldx%lldx%lldx%lld must have H, W multiples of 16",
               (long long)batch, (long long)chans, (long long)height, (long long)width);
  FV_CHECK_ARG(out_dtype == FV_F32 || out_dtype == FV_BF16, "fv_patchify: bad out_dtype");
  FV_CHECK_ARG((reinterpret_cast<uintptr_t>(img) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "fv_patchify: pointers must be 16-byte aligned");
  const long long total = batch * ((height / 16) * (width / 16) + lead_rows) * chans * 64;
  long long want = ceil_div(total, 256 * 4);
  const long long cap = static_cast<long long>(num_sms()) * 16;
  const unsigned grid = static_cast<unsigned>(want < cap ? (want < 1 ? 1 : want) : cap);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (out_dtype == FV_BF16)
    FV_CHECK_CUDA(fv::launch_pdl(patchify_kernel<true>, dim3(grid), dim3(256), 0, st, img, out, (int)batch, (int)chans, (int)height, (int)width, (int)lead_rows));
  else
    FV_CHECK_CUDA(fv::launch_pdl(patchify_kernel<false>, dim3(grid), dim3(256), 0, st, img, out, (int)batch, (int)chans, (int)height, (int)width, (int)lead_rows));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int fv_cls_pos_rows(const float* cls, const float* pos, float* x, int64_t batch,
                               int64_t tokens, int64_t dim, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(cls && pos && x && batch > 0 && tokens > 0 && dim > 0 && batch * dim < (1LL << 31),
               "fv_cls_pos_rows: bad argument");
  const unsigned grid = static_cast<unsigned>(ceil_div(batch * dim, 256));
  FV_CHECK_CUDA(fv::launch_pdl(cls_pos_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), cls, pos, x, (int)batch, tokens, (int)dim));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int fv_colsum(const void* a, int a_dtype, int64_t lda, float* out, int accumulate,
                         int64_t rows, int64_t cols, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(a && out && rows >= 0 && cols > 0 && lda >= cols && cols < (1LL << 30), "fv_colsum: bad argument");
  const bool vec2 = cols % 2 == 0 && lda % 2 == 0 && (reinterpret_cast<uintptr_t>(a) & 7) == 0;
  FV_CHECK_ARG(a_dtype == FV_F32 || a_dtype == FV_BF16, "fv_colsum: bad dtype");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!accumulate) FV_CHECK_CUDA(cudaMemsetAsync(out, 0, cols * sizeof(float), st));
  if (rows == 0) return FV_OK;
  const int col_blocks = static_cast<int>(ceil_div(cols, 64));
  // enough row slices to fill the machine, at least 64 rows each
  int64_t slices = ceil_div(static_cast<int64_t>(num_sms()) * 8, col_blocks);
  int64_t rows_per = ceil_div(rows, slices);
  if (rows_per < 64) rows_per = 64;
  slices = ceil_div(rows, rows_per);
  dim3 grid(col_blocks, static_cast<unsigned>(slices));
  dim3 block(32, 8);
  if (a_dtype == FV_BF16 && vec2)
    FV_CHECK_CUDA(fv::launch_pdl(colsum_kernel<true, true>, dim3(grid), dim3(block), 0, st, a, lda, out, rows, (int)cols, (int)rows_per));
  else if (a_dtype == FV_BF16)
    FV_CHECK_CUDA(fv::launch_pdl(colsum_kernel<true, false>, dim3(grid), dim3(block), 0, st, a, lda, out, rows, (int)cols, (int)rows_per));
  else if (vec2)
    FV_CHECK_CUDA(fv::launch_pdl(colsum_kernel<false, true>, dim3(grid), dim3(block), 0, st, a, lda, out, rows, (int)cols, (int)rows_per));
  else
    FV_CHECK_CUDA(fv::launch_pdl(colsum_kernel<false, false>, dim3(grid), dim3(block), 0, st, a, lda, out, rows, (int)cols, (int)rows_per));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int fv_softmax_rows(const float* s, float* p, int64_t rows, int64_t cols, float scale,
                               void* stream) {
  using namespace fv;
  FV_CHECK_ARG(s && p && rows >= 0 && cols > 0 && cols < (1LL << 30), "fv_softmax_rows: bad argument");
  if (rows == 0) return FV_OK;
  FV_CHECK_CUDA(fv::launch_pdl(softmax_rows_kernel, dim3(static_cast<unsigned>(ceil_div(rows, 8))), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      s, p, rows, (int)cols, scale));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int fv_softmax_rows_bwd(const float* p, const float* dp, float* ds, int64_t rows,
                                   int64_t cols, float scale, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(p && dp && ds && rows >= 0 && cols > 0 && cols < (1LL << 30),
               "fv_softmax_rows_bwd: bad argument");
  if (rows == 0) return FV_OK;
  FV_CHECK_CUDA(fv::launch_pdl(softmax_rows_bwd_kernel, dim3(static_cast<unsigned>(ceil_div(rows, 8))), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      p, dp, ds, rows, (int)cols, scale));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int fv_cls_grad_rows(const float* dcls, const float* row_scale, float* dx, void* dy, int dy_dtype,
                                int64_t batch, int64_t tokens, int64_t dim, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(dcls && (dx || dy) && batch > 0 && tokens > 0 && dim > 0 && dim % 4 == 0,
               "fv_cls_grad_rows: bad argument");
  FV_CHECK_ARG(dy_dtype == FV_F32 || dy_dtype == FV_BF16, "fv_cls_grad_rows: dy_dtype");
  const long long total = batch * tokens * (dim >> 2);
  long long blocks = ceil_div(total, 256 * 4);
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  FV_CHECK_CUDA(fv::launch_pdl(cls_grad_rows_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0,
                               static_cast<cudaStream_t>(stream), dcls, row_scale, dx, dy,
                               dy_dtype == FV_BF16 ? 1 : 0, (int)batch, (int)tokens, (int)dim));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

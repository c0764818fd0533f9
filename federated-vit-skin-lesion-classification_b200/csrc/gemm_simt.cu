// gemm_simt.cu — fp32 strided-batched GEMM on the FFMA pipe, same epilogues as gemm_tc.cu.
//
// This is the fp32 arithmetic path of the ViT (north_star gate: logits and gradients within 1e-4
// of the reference in fp32): tensor cores have no fp32-exact mode (tf32 keeps 10 mantissa bits),
// so the parity configuration (ViT-Tiny) runs its nn.Linear / attention
// contractions here. Operands are addressed through (row, col, batch) strides so forward, dgrad,
// wgrad and the per-head QK^T / PV products all map onto the one kernel.
//   C[b][m,n] = alpha * sum_k A[b](m,k) * B[b](n,k)   then the epilogue
// Reference call sites: the same nn.Linear / SDPA calls as gemm_tc.cu (model.py:193, train.py:153).
#include "common.cuh"

namespace fv {

constexpr int SM_BM = 64, SM_BN = 64, SM_BK = 16;

struct SimtParams {
  const float* a; long long a_rs, a_cs, a_bs;
  const float* b; long long b_rs, b_cs, b_bs;
  const float* bias;
  float* c; long long ldc, c_bs;
  float* aux; long long ldaux;
  int m, n, k;
  float alpha;
  int tokens_per_img;
  const float* row_scale;  // RESIDUAL only: per-sample stochastic-depth scale of the branch
  int rows_per_scale;
};

template <int EPI>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const SimtParams p) {
  pdl_wait();
  __shared__ float As[SM_BK][SM_BM + 4];
  __shared__ float Bs[SM_BK][SM_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * SM_BM, n0 = blockIdx.x * SM_BN;
  const float* A = p.a + static_cast<long long>(blockIdx.z) * p.a_bs;
  const float* B = p.b + static_cast<long long>(blockIdx.z) * p.b_bs;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool a_kfast = p.a_cs == 1;
  const bool b_kfast = p.b_cs == 1;
  for (int k0 = 0; k0 < p.k; k0 += SM_BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i;
      int mm, kk;
      if (a_kfast) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < p.m && gk < p.k) ? A[gm * p.a_rs + gk * p.a_cs] : 0.f;
      int nn, kb;
      if (b_kfast) { kb = e & 15; nn = e >> 4; } else { nn = e & 63; kb = e >> 6; }
      const int gn = n0 + nn, gkb = k0 + kb;
      Bs[kb][nn] = (gn < p.n && gkb < p.k) ? B[gn * p.b_rs + gkb * p.b_cs] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SM_BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w};
      const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    __syncthreads();
  }

  float* C = p.c + static_cast<long long>(blockIdx.z) * p.c_bs;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long row = m0 + ty * 4 + i;
    if (row >= p.m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= p.n) continue;
      float v = acc[i][j] * p.alpha;
      if (EPI != FV_EPI_ACCUM && EPI != FV_EPI_DGELU && p.bias != nullptr) v += p.bias[col];
      if (EPI == FV_EPI_NONE) {
        C[row * p.ldc + col] = v;
      } else if (EPI == FV_EPI_RESIDUAL) {
        if (p.row_scale != nullptr) v *= p.row_scale[row / p.rows_per_scale];
        C[row * p.ldc + col] = v + p.aux[row * p.ldaux + col];
      } else if (EPI == FV_EPI_GELU) {
        if (p.aux != nullptr) p.aux[row * p.ldaux + col] = gelu_erf_grad(v);  // kept for the backward instead of v itself
        C[row * p.ldc + col] = gelu_erf(v);
      } else if (EPI == FV_EPI_DGELU) {
        C[row * p.ldc + col] = v * p.aux[row * p.ldaux + col];
      } else if (EPI == FV_EPI_ACCUM) {
        C[row * p.ldc + col] += v;
      } else if (EPI == FV_EPI_PATCH) {
        const long long img = row / p.tokens_per_img;
        const long long tok = row - img * p.tokens_per_img + 1;
        C[(row + img + 1) * p.ldc + col] = v + p.aux[tok * p.ldaux + col];
      }
    }
  }
}

}  // namespace fv

namespace fv {
static thread_local const float* g_row_scale_f32 = nullptr;
static thread_local int g_rows_per_scale_f32 = 0;
}

extern "C" int fv_gemm_f32(const float* a, int64_t a_row_stride, int64_t a_col_stride,
                           int64_t a_batch_stride, const float* b, int64_t b_row_stride,
                           int64_t b_col_stride, int64_t b_batch_stride, const float* bias, float* c,
                           int64_t ldc, int64_t c_batch_stride, float* aux, int64_t ldaux, int64_t m,
                           int64_t n, int64_t k, int64_t batch, float alpha, int epilogue,
                           int tokens_per_img, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(a && b && c, "fv_gemm_f32: null operand");
  FV_CHECK_ARG(m > 0 && n > 0 && k > 0 && batch > 0, "fv_gemm_f32: empty problem");
  FV_CHECK_ARG(m < (1LL << 31) && n < (1LL << 31) && k < (1LL << 31) && batch <= 65535,
               "fv_gemm_f32: size out of range (batch <= 65535)");
  FV_CHECK_ARG(epilogue >= FV_EPI_NONE && epilogue <= FV_EPI_PATCH, "fv_gemm_f32: bad epilogue");
  if (epilogue == FV_EPI_RESIDUAL || epilogue == FV_EPI_DGELU || epilogue == FV_EPI_PATCH)
    FV_CHECK_ARG(aux != nullptr, "fv_gemm_f32: epilogue needs aux");
  if (epilogue == FV_EPI_PATCH) FV_CHECK_ARG(tokens_per_img > 0, "fv_gemm_f32: tokens_per_img");
  FV_CHECK_ARG(ceil_div(m, SM_BM) <= 65535, "fv_gemm_f32: m too large for the grid");
  SimtParams p;
  p.a = a; p.a_rs = a_row_stride; p.a_cs = a_col_stride; p.a_bs = a_batch_stride;
  p.b = b; p.b_rs = b_row_stride; p.b_cs = b_col_stride; p.b_bs = b_batch_stride;
  p.bias = bias;
  p.c = c; p.ldc = ldc; p.c_bs = c_batch_stride;
  p.aux = aux; p.ldaux = ldaux;
  p.m = (int)m; p.n = (int)n; p.k = (int)k;
  p.alpha = alpha;
  p.tokens_per_img = tokens_per_img;
  p.row_scale = g_row_scale_f32;
  p.rows_per_scale = g_rows_per_scale_f32;
  dim3 grid(static_cast<unsigned>(ceil_div(n, SM_BN)), static_cast<unsigned>(ceil_div(m, SM_BM)),
            static_cast<unsigned>(batch));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (epilogue) {
    case FV_EPI_NONE: FV_CHECK_CUDA(fv::launch_pdl(gemm_simt_kernel<FV_EPI_NONE>, dim3(grid), dim3(256), 0, st, p)); break;
    case FV_EPI_RESIDUAL: FV_CHECK_CUDA(fv::launch_pdl(gemm_simt_kernel<FV_EPI_RESIDUAL>, dim3(grid), dim3(256), 0, st, p)); break;
    case FV_EPI_GELU: FV_CHECK_CUDA(fv::launch_pdl(gemm_simt_kernel<FV_EPI_GELU>, dim3(grid), dim3(256), 0, st, p)); break;
    case FV_EPI_DGELU: FV_CHECK_CUDA(fv::launch_pdl(gemm_simt_kernel<FV_EPI_DGELU>, dim3(grid), dim3(256), 0, st, p)); break;
    case FV_EPI_ACCUM: FV_CHECK_CUDA(fv::launch_pdl(gemm_simt_kernel<FV_EPI_ACCUM>, dim3(grid), dim3(256), 0, st, p)); break;
    case FV_EPI_PATCH: FV_CHECK_CUDA(fv::launch_pdl(gemm_simt_kernel<FV_EPI_PATCH>, dim3(grid), dim3(256), 0, st, p)); break;
  }
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int fv_linear_residual_f32(const float* a, int64_t lda, const float* w, int64_t ldw, const float* bias,
                                      const float* residual, int64_t ldr, const float* row_scale,
                                      int64_t rows_per_scale, float* out, int64_t ldo, int64_t m, int64_t n,
                                      int64_t k, void* stream) {
  fv::g_row_scale_f32 = row_scale;
  fv::g_rows_per_scale_f32 = static_cast<int>(rows_per_scale);
  const int rc = fv_gemm_f32(a, lda, 1, 0, w, ldw, 1, 0, bias, out, ldo, 0, const_cast<float*>(residual), ldr, m, n,
                             k, 1, 1.0f, FV_EPI_RESIDUAL, 0, stream);
  fv::g_row_scale_f32 = nullptr;
  fv::g_rows_per_scale_f32 = 0;
  return rc;
}

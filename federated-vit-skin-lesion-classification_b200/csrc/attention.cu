// attention.cu — flash-style multi-head attention core, forward and backward, head_dim 64.
//
// Replaces F.scaled_dot_product_attention inside timm's Attention.forward (reference call site
// model.py:193) and its backward (train.py:153). Scores never touch HBM: S and P live in
// registers, softmax is online (running max / sum in fp32), only O and the per-row log-sum-exp
// are written. The backward recomputes P from the saved LSE:
//     delta = rowsum(dO * O)
//     dQ kernel : per 64-query tile, sweeps the keys      dQ = sum_k dS K
//     dKV kernel: per 64-key tile, sweeps the queries     dV = P^T dO,  dK = dS^T Q
//   with dS = P * (dO V^T - delta) * scale. Two sweeps instead of one keep dQ out of global
//   atomics, so the gradients are bit-reproducible.
// Layouts fold timm's reshape/permute: q/k/v are read straight out of the [B,N,3,H,64] output of
// Attention.qkv and O is written token-major [B,N,H,64] — what Attention.proj consumes.
//
// Sequences are 197 or 577 tokens: ragged against the 64-wide tiles, so tail keys
// are masked to -inf before the softmax and tail rows are zero-filled on load and never stored.
// Tensor-core path: mma.sync m16n8k16 bf16 (ldmatrix-fed). This is round 1's correct baseline;
// the tcgen05/TMEM variant replaces the two score products next.
#include <stdlib.h>

#include "common.cuh"

namespace fv {

constexpr int AT_T = 64;    // tile rows (queries or keys)
constexpr int AT_D = 64;    // head dim
constexpr int AT_LD = 72;   // padded smem row (144 B): conflict-free ldmatrix
constexpr int AT_THREADS = 128;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 -> 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem),
               "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// copy a [64 x 64] bf16 tile (rows row0.., row stride `rs` elements) into padded smem
__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, long long rs,
                                          int row0, int nrows) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = threadIdx.x + AT_THREADS * i;  // 512 chunks of 16 B
    const int r = e >> 3, c = (e & 7) << 3;
    const bool ok = row0 + r < nrows;
    cp_async16(s + r * AT_LD + c, g + static_cast<long long>(ok ? row0 + r : 0) * rs + c, ok);
  }
}

// A-operand fragments of this warp's 16 rows (4 k-steps over the 64 dims)
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[4][4], const __nv_bfloat16* s, int warp,
                                             int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    ldsm_x4(f[ks], s + (warp * 16 + (lane & 15)) * AT_LD + ks * 16 + (lane >> 4) * 8);
}

// acc[16 x 64] += A(regs, 16 x 64dims) * T^T where T is a smem tile [64 rows x 64 dims]
// (contraction over dims; the tile's rows become the accumulator's columns)
__device__ __forceinline__ void mma_a_tileT(float (&acc)[8][4], const uint32_t (&a)[4][4],
                                            const __nv_bfloat16* t, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4(b, t + (np * 16 + (lane & 7) + (lane >> 4) * 8) * AT_LD + ks * 16 + ((lane >> 3) & 1) * 8);
      mma16816(acc[2 * np], a[ks], b[0], b[1]);
      mma16816(acc[2 * np + 1], a[ks], b[2], b[3]);
    }
  }
}

// acc[16 x 64dims] += P(regs, 16 x 64) * T where T is a smem tile [64 rows x 64 dims]
// (contraction over the tile's rows)
__device__ __forceinline__ void mma_p_tile(float (&acc)[8][4], const uint32_t (&p)[4][4],
                                           const __nv_bfloat16* t, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, t + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * AT_LD + np * 16 + (lane >> 4) * 8);
      mma16816(acc[2 * np], p[ks], b[0], b[1]);
      mma16816(acc[2 * np + 1], p[ks], b[2], b[3]);
    }
  }
}

// fp32 accumulator tile (C layout) -> bf16 A-operand fragments for the next product
__device__ __forceinline__ void acc_to_a(uint32_t (&p)[4][4], const float (&s)[8][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    p[ks][0] = pack_bf16(s[2 * ks][0], s[2 * ks][1]);
    p[ks][1] = pack_bf16(s[2 * ks][2], s[2 * ks][3]);
    p[ks][2] = pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
    p[ks][3] = pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
  }
}

// this warp's 16 x 64 fp32 accumulator -> bf16 rows in global memory through a smem staging tile
__device__ __forceinline__ void store_rows(const float (&o)[8][4], __nv_bfloat16* stage,
                                           __nv_bfloat16* g, long long rs, int row0, int nrows,
                                           int warp, int lane) {
  const int gq = lane >> 2, t = lane & 3;
  __nv_bfloat16* w = stage + warp * 16 * AT_LD;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    *reinterpret_cast<uint32_t*>(w + gq * AT_LD + nb * 8 + 2 * t) = pack_bf16(o[nb][0], o[nb][1]);
    *reinterpret_cast<uint32_t*>(w + (gq + 8) * AT_LD + nb * 8 + 2 * t) = pack_bf16(o[nb][2], o[nb][3]);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 4 + (lane >> 3), c = (lane & 7) << 3;
    const int row = row0 + warp * 16 + r;
    if (row < nrows)
      *reinterpret_cast<uint4*>(g + static_cast<long long>(row) * rs + c) =
          *reinterpret_cast<const uint4*>(w + r * AT_LD + c);
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AT_THREADS)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                float* __restrict__ lse, int N, int H, float scale) {
  pdl_wait();
  __shared__ __align__(16) __nv_bfloat16 sQ[AT_T * AT_LD];
  __shared__ __align__(16) __nv_bfloat16 sK[2][AT_T * AT_LD];
  __shared__ __align__(16) __nv_bfloat16 sV[2][AT_T * AT_LD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, t = lane & 3;
  const int b = blockIdx.y / H, h = blockIdx.y % H;
  const int q0 = blockIdx.x * AT_T;
  const long long rs = 3LL * H * AT_D;
  const __nv_bfloat16* qb = qkv + static_cast<long long>(b) * N * rs + h * AT_D;
  const __nv_bfloat16* kb = qb + H * AT_D;
  const __nv_bfloat16* vb = kb + H * AT_D;
  const int nkv = (N + AT_T - 1) / AT_T;
  const float sl2 = scale * LOG2E;

  load_tile(sQ, qb, rs, q0, N);
  load_tile(sK[0], kb, rs, 0, N);
  load_tile(sV[0], vb, rs, 0, N);
  cp_async_commit();

  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  for (int j = 0; j < nkv; ++j) {
    const int buf = j & 1;
    if (j + 1 < nkv) {
      load_tile(sK[buf ^ 1], kb, rs, (j + 1) * AT_T, N);
      load_tile(sV[buf ^ 1], vb, rs, (j + 1) * AT_T, N);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) load_a_frags(qf, sQ, warp, lane);

    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) s[i][jj] = 0.f;
    mma_a_tileT(s, qf, sK[buf], lane);

    if ((j + 1) * AT_T > N) {
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const int key = j * AT_T + nb * 8 + 2 * t;
        if (key >= N) { s[nb][0] = -INFINITY; s[nb][2] = -INFINITY; }
        if (key + 1 >= N) { s[nb][1] = -INFINITY; s[nb][3] = -INFINITY; }
      }
    }
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nb][0], s[nb][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nb][2], s[nb][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    const float a0 = exp2f((m_run[0] - mx[0]) * sl2);
    const float a1 = exp2f((m_run[1] - mx[1]) * sl2);
    m_run[0] = mx[0];
    m_run[1] = mx[1];
    float rsum[2] = {0.f, 0.f};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      s[nb][0] = exp2f((s[nb][0] - mx[0]) * sl2);
      s[nb][1] = exp2f((s[nb][1] - mx[0]) * sl2);
      s[nb][2] = exp2f((s[nb][2] - mx[1]) * sl2);
      s[nb][3] = exp2f((s[nb][3] - mx[1]) * sl2);
      rsum[0] += s[nb][0] + s[nb][1];
      rsum[1] += s[nb][2] + s[nb][3];
      o[nb][0] *= a0; o[nb][1] *= a0;
      o[nb][2] *= a1; o[nb][3] *= a1;
    }
    l_run[0] = l_run[0] * a0 + rsum[0];
    l_run[1] = l_run[1] * a1 + rsum[1];
    uint32_t pf[4][4];
    acc_to_a(pf, s);
    mma_p_tile(o, pf, sV[buf], lane);
    __syncthreads();  // everyone is done with this K/V buffer before it is refilled
  }

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float i0 = 1.0f / l_run[0], i1 = 1.0f / l_run[1];
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    o[nb][0] *= i0; o[nb][1] *= i0;
    o[nb][2] *= i1; o[nb][3] *= i1;
  }
  // sQ is free (Q lives in registers); reuse it as the store staging tile
  __nv_bfloat16* ob = out + static_cast<long long>(b) * N * H * AT_D + h * AT_D;
  store_rows(o, sQ, ob, static_cast<long long>(H) * AT_D, q0, N, warp, lane);
  if (t == 0) {
    const int r0 = q0 + warp * 16 + gq;
    float* lp = lse + (static_cast<long long>(b) * H + h) * N;
    if (r0 < N) lp[r0] = m_run[0] * scale + logf(l_run[0]);
    if (r0 + 8 < N) lp[r0 + 8] = m_run[1] * scale + logf(l_run[1]);
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// delta[b,h,n] = sum_d dO[b,n,h,d] * O[b,n,h,d]; one warp per (b,n,h) row
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                  float* __restrict__ delta, long long rows, int N, int H) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);  // (b, n, h)
  if (row >= rows) return;
  const float2 a = unpack_bf16(reinterpret_cast<const uint32_t*>(o + row * AT_D)[lane]);
  const float2 g = unpack_bf16(reinterpret_cast<const uint32_t*>(d_o + row * AT_D)[lane]);
  const float s = warp_sum(a.x * g.x + a.y * g.y);
  if (lane == 0) {
    const int h = static_cast<int>(row % H);
    const long long bn = row / H;
    const int n = static_cast<int>(bn % N);
    const long long b = bn / N;
    delta[(b * H + h) * N + n] = s;
  }
}

// dQ: one CTA per 64-query tile; Q and dO fragments in registers, K/V tiles streamed through smem
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ d_o,
                   const float* __restrict__ lse, const float* __restrict__ delta,
                   __nv_bfloat16* __restrict__ dqkv, int N, int H, float scale) {
  pdl_wait();
  __shared__ __align__(16) __nv_bfloat16 sK[2][AT_T * AT_LD];
  __shared__ __align__(16) __nv_bfloat16 sV[2][AT_T * AT_LD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, t = lane & 3;
  const int b = blockIdx.y / H, h = blockIdx.y % H;
  const int q0 = blockIdx.x * AT_T;
  const long long rs = 3LL * H * AT_D;
  const long long os = static_cast<long long>(H) * AT_D;
  const __nv_bfloat16* qb = qkv + static_cast<long long>(b) * N * rs + h * AT_D;
  const __nv_bfloat16* kb = qb + H * AT_D;
  const __nv_bfloat16* vb = kb + H * AT_D;
  const __nv_bfloat16* dob = d_o + static_cast<long long>(b) * N * os + h * AT_D;
  const int nkv = (N + AT_T - 1) / AT_T;
  const float sl2 = scale * LOG2E;

  // stage this CTA's Q and dO tiles through buffer 1, lift them into registers below
  load_tile(sK[1], qb, rs, q0, N);
  load_tile(sV[1], dob, os, q0, N);
  cp_async_commit();
  load_tile(sK[0], kb, rs, 0, N);
  load_tile(sV[0], vb, rs, 0, N);
  cp_async_commit();

  // per-row softmax statistics of this thread's two rows
  const int r0 = q0 + warp * 16 + gq;
  const float* lp = lse + (static_cast<long long>(b) * H + h) * N;
  const float* dp = delta + (static_cast<long long>(b) * H + h) * N;
  float l2[2], dl[2];
  l2[0] = r0 < N ? lp[r0] * LOG2E : INFINITY;
  l2[1] = r0 + 8 < N ? lp[r0 + 8] * LOG2E : INFINITY;
  dl[0] = r0 < N ? dp[r0] : 0.f;
  dl[1] = r0 + 8 < N ? dp[r0 + 8] : 0.f;

  uint32_t qf[4][4], dof[4][4];
  float dq[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;
  cp_async_wait<1>();
  __syncthreads();
  load_a_frags(qf, sK[1], warp, lane);
  load_a_frags(dof, sV[1], warp, lane);
  __syncthreads();

  for (int j = 0; j < nkv; ++j) {
    const int buf = j & 1;
    if (j + 1 < nkv) {
      load_tile(sK[buf ^ 1], kb, rs, (j + 1) * AT_T, N);
      load_tile(sV[buf ^ 1], vb, rs, (j + 1) * AT_T, N);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    float s[8][4], pd[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) { s[i][jj] = 0.f; pd[i][jj] = 0.f; }
    mma_a_tileT(s, qf, sK[buf], lane);    // S  = Q K^T
    mma_a_tileT(pd, dof, sV[buf], lane);  // dP = dO V^T
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const int key = j * AT_T + nb * 8 + 2 * t;
      const bool v0 = key < N, v1 = key + 1 < N;
      const float p0 = v0 ? exp2f(s[nb][0] * sl2 - l2[0]) : 0.f;
      const float p1 = v1 ? exp2f(s[nb][1] * sl2 - l2[0]) : 0.f;
      const float p2 = v0 ? exp2f(s[nb][2] * sl2 - l2[1]) : 0.f;
      const float p3 = v1 ? exp2f(s[nb][3] * sl2 - l2[1]) : 0.f;
      s[nb][0] = p0 * (pd[nb][0] - dl[0]) * scale;
      s[nb][1] = p1 * (pd[nb][1] - dl[0]) * scale;
      s[nb][2] = p2 * (pd[nb][2] - dl[1]) * scale;
      s[nb][3] = p3 * (pd[nb][3] - dl[1]) * scale;
    }
    uint32_t dsf[4][4];
    acc_to_a(dsf, s);
    mma_p_tile(dq, dsf, sK[buf], lane);  // dQ += dS K
    __syncthreads();
  }
  __nv_bfloat16* dqb = dqkv + static_cast<long long>(b) * N * rs + h * AT_D;
  store_rows(dq, sK[0], dqb, rs, q0, N, warp, lane);
}

// dK, dV: one CTA per 64-key tile; K and V fragments in registers, Q/dO tiles streamed
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ d_o,
                    const float* __restrict__ lse, const float* __restrict__ delta,
                    __nv_bfloat16* __restrict__ dqkv, int N, int H, float scale) {
  pdl_wait();
  __shared__ __align__(16) __nv_bfloat16 sQ[2][AT_T * AT_LD];
  __shared__ __align__(16) __nv_bfloat16 sdO[2][AT_T * AT_LD];
  __shared__ float sL[2][AT_T];
  __shared__ float sD[2][AT_T];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = lane & 3;
  const int b = blockIdx.y / H, h = blockIdx.y % H;
  const int k0 = blockIdx.x * AT_T;
  const long long rs = 3LL * H * AT_D;
  const long long os = static_cast<long long>(H) * AT_D;
  const __nv_bfloat16* qb = qkv + static_cast<long long>(b) * N * rs + h * AT_D;
  const __nv_bfloat16* kb = qb + H * AT_D;
  const __nv_bfloat16* vb = kb + H * AT_D;
  const __nv_bfloat16* dob = d_o + static_cast<long long>(b) * N * os + h * AT_D;
  const float* lp = lse + (static_cast<long long>(b) * H + h) * N;
  const float* dp = delta + (static_cast<long long>(b) * H + h) * N;
  const int nq = (N + AT_T - 1) / AT_T;
  const float sl2 = scale * LOG2E;

  // stage this CTA's K and V tiles through buffer 1, lift them into registers
  load_tile(sQ[1], kb, rs, k0, N);
  load_tile(sdO[1], vb, rs, k0, N);
  cp_async_commit();
  load_tile(sQ[0], qb, rs, 0, N);
  load_tile(sdO[0], dob, os, 0, N);
  cp_async_commit();
  if (threadIdx.x < AT_T) {
    const int q = threadIdx.x;
    sL[0][q] = q < N ? lp[q] * LOG2E : INFINITY;
    sD[0][q] = q < N ? dp[q] : 0.f;
  }
  cp_async_wait<1>();
  __syncthreads();
  uint32_t kf[4][4], vf[4][4];
  load_a_frags(kf, sQ[1], warp, lane);
  load_a_frags(vf, sdO[1], warp, lane);
  __syncthreads();

  float dk[8][4], dv[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { dk[i][j] = 0.f; dv[i][j] = 0.f; }

  for (int j = 0; j < nq; ++j) {
    const int buf = j & 1;
    if (j + 1 < nq) {
      load_tile(sQ[buf ^ 1], qb, rs, (j + 1) * AT_T, N);
      load_tile(sdO[buf ^ 1], dob, os, (j + 1) * AT_T, N);
      cp_async_commit();
      if (threadIdx.x < AT_T) {
        const int q = (j + 1) * AT_T + threadIdx.x;
        sL[buf ^ 1][threadIdx.x] = q < N ? lp[q] * LOG2E : INFINITY;
        sD[buf ^ 1][threadIdx.x] = q < N ? dp[q] : 0.f;
      }
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    float st[8][4], pd[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) { st[i][jj] = 0.f; pd[i][jj] = 0.f; }
    mma_a_tileT(st, kf, sQ[buf], lane);   // S^T  = K Q^T      [16 keys x 64 queries]
    mma_a_tileT(pd, vf, sdO[buf], lane);  // dP^T = V dO^T
    uint32_t pf[4][4], dsf[4][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const int q = nb * 8 + 2 * t;
      const float la = sL[buf][q], lb = sL[buf][q + 1];
      const float da = sD[buf][q], db = sD[buf][q + 1];
      const float p0 = exp2f(st[nb][0] * sl2 - la);
      const float p1 = exp2f(st[nb][1] * sl2 - lb);
      const float p2 = exp2f(st[nb][2] * sl2 - la);
      const float p3 = exp2f(st[nb][3] * sl2 - lb);
      st[nb][0] = p0; st[nb][1] = p1; st[nb][2] = p2; st[nb][3] = p3;
      pd[nb][0] = p0 * (pd[nb][0] - da) * scale;
      pd[nb][1] = p1 * (pd[nb][1] - db) * scale;
      pd[nb][2] = p2 * (pd[nb][2] - da) * scale;
      pd[nb][3] = p3 * (pd[nb][3] - db) * scale;
    }
    acc_to_a(pf, st);
    acc_to_a(dsf, pd);
    mma_p_tile(dv, pf, sdO[buf], lane);  // dV += P^T dO
    mma_p_tile(dk, dsf, sQ[buf], lane);  // dK += dS^T Q
    __syncthreads();
  }
  __nv_bfloat16* dkb = dqkv + static_cast<long long>(b) * N * rs + H * AT_D + h * AT_D;
  __nv_bfloat16* dvb = dkb + H * AT_D;
  store_rows(dk, sQ[0], dkb, rs, k0, N, warp, lane);
  __syncwarp();
  store_rows(dv, sdO[0], dvb, rs, k0, N, warp, lane);
}

// tcgen05 path for short sequences (attention_tc.cu)
int attention_tc_fwd(const void* qkv, void* out, float* lse, int64_t batch, int64_t tokens, int64_t heads,
                     float scale, cudaStream_t stream);

int attention_tc_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                     int64_t batch, int64_t tokens, int64_t heads, float scale, cudaStream_t stream);

// second-generation forward, two threads per query row (attention_fwd2.cu); FEDVIT_ATTN_FWD=v1 keeps the first
int attention_tc_fwd2(const void* qkv, void* out, float* lse, int64_t batch, int64_t tokens, int64_t heads,
                      float scale, cudaStream_t stream);

// third-generation forward (attention_fwd3.cu): next tile's score MMA behind this tile's PV MMA, pipelined
// tensor-memory loads, early PV start, rotated ragged tile; N <= 208
int attention_tc_fwd3(const void* qkv, void* out, float* lse, int64_t batch, int64_t tokens, int64_t heads,
                      float scale, cudaStream_t stream);

// fourth-generation forward (attention_fwd4.cu): one pass over the scores (no row-maximum pass), softmax and
// read-out on different warps, two score accumulators, two shared-memory stages; N <= 208
int attention_tc_fwd4(const void* qkv, void* out, float* lse, int64_t batch, int64_t tokens, int64_t heads,
                      float scale, cudaStream_t stream);

// second-generation backward, keys on lanes (attention_bwd2.cu); FEDVIT_ATTN_BWD=v1 keeps the first kernel
int attention_tc_bwd2(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                      int64_t batch, int64_t tokens, int64_t heads, float scale, cudaStream_t stream);

int attention_tc_fwd_long(const void* qkv, void* out, float* lse, int64_t batch, int64_t tokens, int64_t heads,
                          float scale, cudaStream_t stream);

int64_t attention_tc_bwd_long_workspace(int64_t batch, int64_t tokens, int64_t heads);
int attention_tc_bwd_long(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv,
                          void* workspace, int64_t batch, int64_t tokens, int64_t heads, float scale,
                          cudaStream_t stream);

static bool attention_forced_legacy() {
  static int forced = -1;  // FEDVIT_ATTN=legacy keeps every length on the mma.sync kernels (A/B checks)
  if (forced < 0) {
    const char* e = getenv("FEDVIT_ATTN");
    forced = (e != nullptr && e[0] == 'l') ? 1 : 0;
  }
  return forced == 1;
}
static bool use_tc_attention_long(int64_t tokens) { return !attention_forced_legacy() && tokens > 256 && tokens <= 768; }

static bool use_tc_attention(int64_t tokens) {
  static int forced = -1;  // FEDVIT_ATTN=legacy keeps every length on the mma.sync kernels (A/B checks)
  if (forced < 0) {
    const char* e = getenv("FEDVIT_ATTN");
    forced = (e != nullptr && e[0] == 'l') ? 1 : 0;
  }
  return forced == 0 && tokens <= 256;
}

}  // namespace fv

// fp32 instances of the attention core are composed on the host from fv_gemm_f32 (strided-batched
// QK^T, PV and their gradients) + fv_softmax_rows; the entry points below are the bf16 flash path.
extern "C" int fv_attention_fwd(const void* qkv, void* out, float* lse, int dtype, int64_t batch,
                                int64_t tokens, int64_t heads, float scale, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(qkv && out && lse, "fv_attention_fwd: null pointer");
  FV_CHECK_ARG(dtype == FV_BF16, "fv_attention_fwd: only FV_BF16 (fp32 is composed from fv_gemm_f32)");
  FV_CHECK_ARG(batch > 0 && tokens > 0 && heads > 0 && batch * heads <= 65535 && tokens < (1 << 20),
               "fv_attention_fwd: shape out of range");
  FV_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "fv_attention_fwd: pointers must be 16-byte aligned");
  if (use_tc_attention(tokens)) {
    // N <= 208 (ViT @ 224: 197): the fourth-generation kernel (attention_fwd4.cu: one pass over the scores, eight
    // softmax warps on 16-lane fragments, read-out on its own warps) — 85 vs 99 us at 256 x 197 x 12 alone,
    // 105 vs 113 us inside a step. FEDVIT_ATTN_FWD = v1 | v2 | v3 selects the earlier generations (v1 also
    // serves 208 < N <= 256); read per call so one process can A/B them
    const char* e = getenv("FEDVIT_ATTN_FWD");
    const char ver = (e != nullptr && e[0] == 'v') ? e[1] : '4';
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (ver == '4' && tokens <= 208 && scale > 0.f) return attention_tc_fwd4(qkv, out, lse, batch, tokens, heads, scale, st);
    if (ver == '3' && tokens <= 208) return attention_tc_fwd3(qkv, out, lse, batch, tokens, heads, scale, st);
    if (ver == '2') return attention_tc_fwd2(qkv, out, lse, batch, tokens, heads, scale, st);
    return attention_tc_fwd(qkv, out, lse, batch, tokens, heads, scale, st);
  }
  if (use_tc_attention_long(tokens))
    return attention_tc_fwd_long(qkv, out, lse, batch, tokens, heads, scale, static_cast<cudaStream_t>(stream));
  dim3 grid(static_cast<unsigned>(ceil_div(tokens, AT_T)), static_cast<unsigned>(batch * heads));
  FV_CHECK_CUDA(fv::launch_pdl(attn_fwd_kernel, dim3(grid), dim3(AT_THREADS), 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), lse,
      (int)tokens, (int)heads, scale));
  count_kernel(FV_KERNEL_ATTN_LEGACY);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int64_t fv_attention_bwd_workspace(int64_t batch, int64_t tokens, int64_t heads) {
  using namespace fv;
  return use_tc_attention_long(tokens) ? attention_tc_bwd_long_workspace(batch, tokens, heads) : 0;
}

extern "C" int fv_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                                float* delta, void* dqkv, int dtype, int64_t batch, int64_t tokens,
                                int64_t heads, float scale, void* workspace, int64_t workspace_bytes,
                                void* stream) {
  using namespace fv;
  FV_CHECK_ARG(qkv && out && dout && lse && delta && dqkv, "fv_attention_bwd: null pointer");
  FV_CHECK_ARG(dtype == FV_BF16, "fv_attention_bwd: only FV_BF16 (fp32 is composed from fv_gemm_f32)");
  FV_CHECK_ARG(batch > 0 && tokens > 0 && heads > 0 && batch * heads <= 65535 && tokens < (1 << 20),
               "fv_attention_bwd: shape out of range");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (use_tc_attention(tokens)) {
    // FEDVIT_ATTN_BWD=v2: the keys-on-lanes kernel (attention_bwd2.cu) — numerically equivalent (P enters
    // dS rounded to bf16: dq / dk error 2.4e-3 -> 2.9e-3 against fp64) and, as measured so far, no faster
    // than the first-generation kernel (208 vs 212 us at 256 x 197 x 12), which therefore stays the
    // default; read per call so one process can A/B them
    const char* e = getenv("FEDVIT_ATTN_BWD");
    const bool v1 = !(e != nullptr && e[0] == 'v' && e[1] == '2');
    return v1 ? attention_tc_bwd(qkv, out, dout, lse, dqkv, batch, tokens, heads, scale, st)
              : attention_tc_bwd2(qkv, out, dout, lse, dqkv, batch, tokens, heads, scale, st);
  }
  const long long rows = batch * tokens * heads;
  FV_CHECK_CUDA(fv::launch_pdl(attn_delta_kernel, dim3(static_cast<unsigned>(ceil_div(rows, 8))), dim3(256), 0, st, 
      reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(dout), delta,
      rows, (int)tokens, (int)heads));
  FV_LAUNCH_CHECK();
  if (use_tc_attention_long(tokens)) {
    FV_CHECK_ARG(workspace != nullptr && workspace_bytes >= attention_tc_bwd_long_workspace(batch, tokens, heads) &&
                     (reinterpret_cast<uintptr_t>(workspace) & 127) == 0,
                 "fv_attention_bwd: needs a 128-byte aligned workspace of fv_attention_bwd_workspace() bytes");
    return attention_tc_bwd_long(qkv, dout, lse, delta, dqkv, workspace, batch, tokens, heads, scale, st);
  }
  dim3 grid(static_cast<unsigned>(ceil_div(tokens, AT_T)), static_cast<unsigned>(batch * heads));
  FV_CHECK_CUDA(fv::launch_pdl(attn_bwd_dq_kernel, dim3(grid), dim3(AT_THREADS), 0, st, 
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(dout), lse,
      delta, reinterpret_cast<__nv_bfloat16*>(dqkv), (int)tokens, (int)heads, scale));
  FV_LAUNCH_CHECK();
  FV_CHECK_CUDA(fv::launch_pdl(attn_bwd_dkv_kernel, dim3(grid), dim3(AT_THREADS), 0, st, 
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(dout), lse,
      delta, reinterpret_cast<__nv_bfloat16*>(dqkv), (int)tokens, (int)heads, scale));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

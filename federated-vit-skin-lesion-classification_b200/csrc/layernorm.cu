// layernorm.cu — fused LayerNorm forward / backward (HBM-bound).
//   forward            one warp per row, registers only
//   backward, default  bulk-copy ring: cp.async.bulk streams whole rows into shared memory, one
//                      consumer warp per row (bf16 dy + residual gradient, cols % 128 == 0)
//   backward, other    register-pipelined, two warps per row (fp32 dy, no residual gradient, odd widths)
//
// Replaces F.layer_norm behind timm's Block.norm1 / norm2 and VisionTransformer.norm (reference
// call site model.py:193) and its autograd backward (train.py:153). Statistics are fp32 whatever
// the output type, as under autocast.
//
// Algorithmic bytes per row of D columns:
//   fwd: 4D (x) + sizeof(y)·D + 8 (mean,rstd)          bwd: sizeof(dy)·D + 4D (x) + 4D (dres)
//                                                           + 4D (dx) [+ 2D dx_lp] + 8
#include <stdlib.h>

#include "common.cuh"

namespace fv {

constexpr int LN_MAX_VEC = 8;     // float4 per lane -> D <= 1024
constexpr int LN_WARPS = 8;       // rows per CTA pass

template <int NV, bool OUT_BF16>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, void* __restrict__ y, float* __restrict__ mean,
                     float* __restrict__ rstd, long long rows, int cols, float eps) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = cols >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * cols);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      v[i] = __ldcs(xr + c);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mu = warp_sum(s) / cols;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const float a = v[i].x - mu, b = v[i].y - mu, d = v[i].z - mu, e = v[i].w - mu;
      q += (a * a + b * b) + (d * d + e * e);
    }
  }
  const float rs = rsqrtf(warp_sum(q) / cols + eps);
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
      float4 o;
      o.x = (v[i].x - mu) * rs * g.x + b.x;
      o.y = (v[i].y - mu) * rs * g.y + b.y;
      o.z = (v[i].z - mu) * rs * g.z + b.z;
      o.w = (v[i].w - mu) * rs * g.w + b.w;
      if (OUT_BF16) {
        uint2 pk;
        pk.x = pack_bf16(o.x, o.y);
        pk.y = pack_bf16(o.z, o.w);
        reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + row * cols)[c] = pk;
      } else {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + row * cols)[c] = o;
      }
    }
  }
}

// Backward. Two warps share a row (each thread owns NV float4 columns: 3 for D = 768), four rows
// per CTA pass, grid-stride over rows. Splitting the row keeps the per-thread state (x-hat, g*dy
// and the running dgamma / dbeta slices) near 80 registers — three CTAs per SM instead of the one
// or two a warp-per-row layout gets — which is what an HBM-bound kernel needs. The two row
// statistics cross the warp pair through shared memory (pair-local named barriers, no CTA-wide
// sync in the row loop); dgamma / dbeta are reduced over the CTA's
// four row slots in shared memory and leave with one atomicAdd per column per CTA.
constexpr int LNB_THREADS = 256;
constexpr int LNB_ROWS = 4;  // rows in flight per CTA (one per warp pair)

template <int NV, bool DY_BF16, int MINB>
__global__ void __launch_bounds__(LNB_THREADS, MINB)
layernorm_bwd_kernel(const void* __restrict__ dy, const float* __restrict__ x,
                     const float* __restrict__ gamma, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ dres,
                     float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_lp,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows,
                     int cols, const float* __restrict__ lp_scale, int rows_per_scale) {
  pdl_wait();
  extern __shared__ float red[];       // [2][LNB_ROWS][cols] for the final dgamma / dbeta reduction
  __shared__ float2 stat[LNB_ROWS][2];  // per warp pair: partial (s1, s2) of each warp
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int slot = warp >> 1;            // row slot of this warp pair
  const int t64 = threadIdx.x & 63;      // thread index inside the pair
  const int nvec = cols >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  float4 dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv_cols = 1.0f / cols;
  const long long stride = static_cast<long long>(gridDim.x) * LNB_ROWS;
  const long long first = static_cast<long long>(blockIdx.x) * LNB_ROWS;
  // each warp pair walks its own rows and only ever synchronises with its partner (named barrier
  // slot+1, 64 threads): no CTA-wide barrier inside the loop, so pairs overlap each other's latency
  for (long long row = first + slot; row < rows; row += stride) {
    const bool live = true;
    float4 xh[NV], gy[NV], rr[NV];
    float s1 = 0.f, s2 = 0.f;
    float mu = 0.f, rs = 0.f;
    if (live) {
      // the residual-stream gradient is only needed after the row reduction, but its load is issued
      // here with the x / dy loads: one round trip to HBM per row instead of two
      if (dres != nullptr) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = t64 + 64 * i;
          if (c < nvec) rr[i] = __ldcs(reinterpret_cast<const float4*>(dres + row * cols) + c);
        }
      }
      mu = mean[row];
      rs = rstd[row];
      const float4* xr = reinterpret_cast<const float4*>(x + row * cols);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = t64 + 64 * i;
        if (c < nvec) {
          const float4 xv = __ldcs(xr + c);
          float4 d;
          if (DY_BF16) {
            const uint2 pk = __ldcs(
                reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy) + row * cols) + c);
            const float2 lo = unpack_bf16(pk.x), hi = unpack_bf16(pk.y);
            d = make_float4(lo.x, lo.y, hi.x, hi.y);
          } else {
            d = __ldcs(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + row * cols) + c);
          }
          const float4 gm = __ldg(g4 + c);
          xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
          gy[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
          dg[i].x += d.x * xh[i].x; dg[i].y += d.y * xh[i].y;
          dg[i].z += d.z * xh[i].z; dg[i].w += d.w * xh[i].w;
          db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
          s1 += (gy[i].x + gy[i].y) + (gy[i].z + gy[i].w);
          s2 += (gy[i].x * xh[i].x + gy[i].y * xh[i].y) + (gy[i].z * xh[i].z + gy[i].w * xh[i].w);
        }
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) stat[slot][warp & 1] = make_float2(s1, s2);
    asm volatile("bar.sync %0, 64;" ::"r"(slot + 1) : "memory");
    const float2 a = stat[slot][0], b2 = stat[slot][1];
    s1 = (a.x + b2.x) * inv_cols;
    s2 = (a.y + b2.y) * inv_cols;
    if (live) {
      // stochastic depth: the bf16 copy feeds the *branch* GEMMs of the preceding sub-layer, whose
      // output was scaled per sample in the forward — scale its gradient the same way
      const float lps = lp_scale != nullptr ? __ldg(lp_scale + row / rows_per_scale) : 1.0f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = t64 + 64 * i;
        if (c < nvec) {
          float4 o;
          o.x = rs * (gy[i].x - s1 - xh[i].x * s2);
          o.y = rs * (gy[i].y - s1 - xh[i].y * s2);
          o.z = rs * (gy[i].z - s1 - xh[i].z * s2);
          o.w = rs * (gy[i].w - s1 - xh[i].w * s2);
          if (dres != nullptr) {
            o.x += rr[i].x; o.y += rr[i].y; o.z += rr[i].z; o.w += rr[i].w;
          }
          reinterpret_cast<float4*>(dx + row * cols)[c] = o;
          if (dx_lp != nullptr) {
            uint2 pk;
            pk.x = pack_bf16(o.x * lps, o.y * lps);
            pk.y = pack_bf16(o.z * lps, o.w * lps);
            reinterpret_cast<uint2*>(dx_lp + row * cols)[c] = pk;
          }
        }
      }
    }
    asm volatile("bar.sync %0, 64;" ::"r"(slot + 1) : "memory");  // stat[] is rewritten by the next pass
  }
  // CTA reduction of the four row slots' dgamma / dbeta slices
  float* red_g = red;
  float* red_b = red + LNB_ROWS * cols;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = t64 + 64 * i;
    if (c < nvec) {
      reinterpret_cast<float4*>(red_g + slot * cols)[c] = dg[i];
      reinterpret_cast<float4*>(red_b + slot * cols)[c] = db[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int w = 0; w < LNB_ROWS; ++w) {
      sg += red_g[w * cols + c];
      sb += red_b[w * cols + c];
    }
    atomicAdd(dgamma + c, sg);
    atomicAdd(dbeta + c, sb);
  }
}


// Software-pipelined variant: the loads of a pair's NEXT row (x, dy, the residual gradient and the
// two statistics) are issued before the current row is reduced and written, so a warp pair always
// has a row in flight; two CTAs per SM (the second row buffer costs ~30 registers).
template <int NV, bool DY_BF16>
struct LnRowRaw {
  float4 x[NV], r[NV];
  uint32_t d[NV * (DY_BF16 ? 2 : 4)];  // four values per column group: two packed words (bf16) or four (fp32)
  float mu, rs;
};
template <int NV, bool DY_BF16>
__device__ __forceinline__ void ln_load_row(LnRowRaw<NV, DY_BF16>& w, long long row, int cols, int nvec, int t64,
                                            const void* dy, const float* x, const float* mean, const float* rstd,
                                            const float* dres) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = t64 + 64 * i;
    if (c < nvec) {
      w.x[i] = __ldcs(reinterpret_cast<const float4*>(x + row * cols) + c);
      if (DY_BF16) {
        const uint2 pk = __ldcs(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy) + row * cols) + c);
        w.d[2 * i] = pk.x;
        w.d[2 * i + 1] = pk.y;
      } else {
        const float4 f = __ldcs(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + row * cols) + c);
        w.d[4 * i] = __float_as_uint(f.x);
        w.d[4 * i + 1] = __float_as_uint(f.y);
        w.d[4 * i + 2] = __float_as_uint(f.z);
        w.d[4 * i + 3] = __float_as_uint(f.w);
      }
      if (dres != nullptr) w.r[i] = __ldcs(reinterpret_cast<const float4*>(dres + row * cols) + c);
    }
  }
  w.mu = __ldg(mean + row);
  w.rs = __ldg(rstd + row);
}

template <int NV, bool DY_BF16>
__global__ void __launch_bounds__(LNB_THREADS, 2)
layernorm_bwd_pipe_kernel(const void* __restrict__ dy, const float* __restrict__ x,
                          const float* __restrict__ gamma, const float* __restrict__ mean,
                          const float* __restrict__ rstd, const float* __restrict__ dres,
                          float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_lp,
                          float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows,
                          int cols, const float* __restrict__ lp_scale, int rows_per_scale) {
  pdl_wait();
  extern __shared__ float red[];
  __shared__ float2 stat[LNB_ROWS][2];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int slot = warp >> 1;
  const int t64 = threadIdx.x & 63;
  const int nvec = cols >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  float4 dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv_cols = 1.0f / cols;
  const long long stride = static_cast<long long>(gridDim.x) * LNB_ROWS;
  long long row = static_cast<long long>(blockIdx.x) * LNB_ROWS + slot;
  LnRowRaw<NV, DY_BF16> cur, nxt;
  if (row < rows) ln_load_row<NV, DY_BF16>(cur, row, cols, nvec, t64, dy, x, mean, rstd, dres);
  for (; row < rows; row += stride) {
    const long long nrow = row + stride;
    if (nrow < rows) ln_load_row<NV, DY_BF16>(nxt, nrow, cols, nvec, t64, dy, x, mean, rstd, dres);
    float4 xh[NV], gy[NV];
    float s1 = 0.f, s2 = 0.f;
    const float mu = cur.mu, rs = cur.rs;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = t64 + 64 * i;
      if (c < nvec) {
        float4 d;
        if (DY_BF16) {
          const float2 lo = unpack_bf16(cur.d[2 * i]), hi = unpack_bf16(cur.d[2 * i + 1]);
          d = make_float4(lo.x, lo.y, hi.x, hi.y);
        } else {
          d = make_float4(__uint_as_float(cur.d[4 * i]), __uint_as_float(cur.d[4 * i + 1]),
                          __uint_as_float(cur.d[4 * i + 2]), __uint_as_float(cur.d[4 * i + 3]));
        }
        const float4 xv = cur.x[i];
        const float4 gm = __ldg(g4 + c);
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        gy[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
        dg[i].x += d.x * xh[i].x; dg[i].y += d.y * xh[i].y;
        dg[i].z += d.z * xh[i].z; dg[i].w += d.w * xh[i].w;
        db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
        s1 += (gy[i].x + gy[i].y) + (gy[i].z + gy[i].w);
        s2 += (gy[i].x * xh[i].x + gy[i].y * xh[i].y) + (gy[i].z * xh[i].z + gy[i].w * xh[i].w);
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) stat[slot][warp & 1] = make_float2(s1, s2);
    asm volatile("bar.sync %0, 64;" ::"r"(slot + 1) : "memory");
    const float2 a = stat[slot][0], b2 = stat[slot][1];
    s1 = (a.x + b2.x) * inv_cols;
    s2 = (a.y + b2.y) * inv_cols;
    const float lps = lp_scale != nullptr ? __ldg(lp_scale + row / rows_per_scale) : 1.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = t64 + 64 * i;
      if (c < nvec) {
        float4 o;
        o.x = rs * (gy[i].x - s1 - xh[i].x * s2);
        o.y = rs * (gy[i].y - s1 - xh[i].y * s2);
        o.z = rs * (gy[i].z - s1 - xh[i].z * s2);
        o.w = rs * (gy[i].w - s1 - xh[i].w * s2);
        if (dres != nullptr) {
          o.x += cur.r[i].x; o.y += cur.r[i].y; o.z += cur.r[i].z; o.w += cur.r[i].w;
        }
        reinterpret_cast<float4*>(dx + row * cols)[c] = o;
        if (dx_lp != nullptr) {
          uint2 pk;
          pk.x = pack_bf16(o.x * lps, o.y * lps);
          pk.y = pack_bf16(o.z * lps, o.w * lps);
          reinterpret_cast<uint2*>(dx_lp + row * cols)[c] = pk;
        }
      }
    }
    asm volatile("bar.sync %0, 64;" ::"r"(slot + 1) : "memory");  // stat[] is rewritten by the next pass
    cur = nxt;
  }
  float* red_g = red;
  float* red_b = red + LNB_ROWS * cols;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = t64 + 64 * i;
    if (c < nvec) {
      reinterpret_cast<float4*>(red_g + slot * cols)[c] = dg[i];
      reinterpret_cast<float4*>(red_b + slot * cols)[c] = db[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int w = 0; w < LNB_ROWS; ++w) {
      sg += red_g[w * cols + c];
      sb += red_b[w * cols + c];
    }
    atomicAdd(dgamma + c, sg);
    atomicAdd(dbeta + c, sb);
  }
}


// ---------------------------------------------------------------------------------------------
// Bulk-copy ring variant (bf16 dy + residual gradient, cols a multiple of 128): one CTA per SM.
// A producer warp streams whole rows — x, the residual gradient and dy, 10 * cols bytes — into a
// shared-memory ring with cp.async.bulk (1-D TMA copies, one mbarrier per group of eight rows);
// eight consumer warps take one row of the group each, pull it into registers, hand the slot back
// and do the arithmetic. With the ring holding 16 - 56 rows per SM (~180 KB in flight instead of
// the ~60 KB eight register-staged rows give) an SM no longer needs 147 peers to cover the HBM
// latency: a fraction of the SMs saturates the memory system, which is what lets this pass share
// the GPU with a tensor-bound kernel.
// ---------------------------------------------------------------------------------------------
constexpr int LNR_ROWS = 8;                          // consumer warps == rows per group
constexpr int LNR_THREADS = (LNR_ROWS + 1) * 32;     // + the producer warp
constexpr int LNR_MAX_DEPTH = 8;
constexpr int LNR_SM_NUM = 73, LNR_SM_DEN = 100;  // share of the SMs the ring kernel runs on (108 of 148)

__device__ __forceinline__ void bulk_load_row(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int NVW /* float4 per lane: cols == 128 * NVW */>
__global__ void __launch_bounds__(LNR_THREADS, 1)
layernorm_bwd_ring_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x,
                          const float* __restrict__ gamma, const float* __restrict__ mean,
                          const float* __restrict__ rstd, const float* __restrict__ dres,
                          float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_lp,
                          float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows,
                          const float* __restrict__ lp_scale, int rows_per_scale, int depth) {
  constexpr int COLS = NVW * 128;
  constexpr uint32_t X_BYTES = COLS * 4, D_BYTES = COLS * 2;
  constexpr uint32_t ROW_BYTES = 2 * X_BYTES + D_BYTES;
  constexpr uint32_t GROUP_BYTES = LNR_ROWS * ROW_BYTES;
  extern __shared__ __align__(128) uint8_t ring_raw[];
  uint8_t* ring = ring_raw + ((128u - (smem_u32(ring_raw) & 127u)) & 127u);
  __shared__ uint64_t full_bar[LNR_MAX_DEPTH], empty_bar[LNR_MAX_DEPTH];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < depth; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], LNR_ROWS);
    }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();
  const long long groups = (rows + LNR_ROWS - 1) / LNR_ROWS;

  if (warp == LNR_ROWS) {
    // ------------------------------- producer ------------------------------------------------
    int slot = 0;
    uint32_t phase = 0;
    const int r = lane / 3, which = lane - 3 * r;  // lanes 0..23: (row of the group, which array)
    for (long long g = blockIdx.x; g < groups; g += gridDim.x) {
      mbar_wait(&empty_bar[slot], phase ^ 1);
      const long long row0 = g * LNR_ROWS;
      const int nrows = rows - row0 < LNR_ROWS ? static_cast<int>(rows - row0) : LNR_ROWS;
      if (lane == 0) mbar_expect_tx(&full_bar[slot], nrows * ROW_BYTES);
      __syncwarp();
      if (lane < 3 * LNR_ROWS && r < nrows) {
        uint8_t* dst = ring + slot * GROUP_BYTES + r * ROW_BYTES;
        const long long row = row0 + r;
        if (which == 0) bulk_load_row(dst, x + row * COLS, X_BYTES, &full_bar[slot]);
        else if (which == 1) bulk_load_row(dst + X_BYTES, dres + row * COLS, X_BYTES, &full_bar[slot]);
        else bulk_load_row(dst + 2 * X_BYTES, dy + row * COLS, D_BYTES, &full_bar[slot]);
      }
      if (++slot == depth) { slot = 0; phase ^= 1; }
    }
    return;
  }

  // --------------------------------- consumers -----------------------------------------------
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  // gamma lives in registers while the budget allows (168 per thread at 288 threads); wider rows
  // re-read it through L1 each row
  constexpr bool GM_REGS = NVW <= 5;
  float4 gm[GM_REGS ? NVW : 1], dg[NVW], db[NVW];
#pragma unroll
  for (int i = 0; i < NVW; ++i) {
    if (GM_REGS) gm[i] = __ldg(g4 + i * 32 + lane);
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  constexpr float inv_cols = 1.0f / COLS;
  int slot = 0;
  uint32_t phase = 0;
  long long row = static_cast<long long>(blockIdx.x) * LNR_ROWS + warp;
  const long long stride = static_cast<long long>(gridDim.x) * LNR_ROWS;
  float mu_n = 0.f, rs_n = 0.f, lps_n = 1.f;
  if (row < rows) {
    mu_n = __ldg(mean + row);
    rs_n = __ldg(rstd + row);
    if (lp_scale != nullptr) lps_n = __ldg(lp_scale + row / rows_per_scale);
  }
  for (long long g = blockIdx.x; g < groups; g += gridDim.x, row += stride) {
    const bool live = row < rows;
    const float mu = mu_n, rs = rs_n, lps = lps_n;
    if (row + stride < rows) {  // the next row's statistics travel while this one is processed
      mu_n = __ldg(mean + row + stride);
      rs_n = __ldg(rstd + row + stride);
      if (lp_scale != nullptr) lps_n = __ldg(lp_scale + (row + stride) / rows_per_scale);
    }
    mbar_wait(&full_bar[slot], phase);
    // Rows up to 768 columns are pulled into registers whole and the slot is handed back at once;
    // wider rows would spill (168 registers per thread), so their residual gradient stays in the
    // ring until the output pass and the slot is released after it.
    constexpr bool LATE = NVW >= 7;
    float4 xh[NVW], rr[LATE ? 1 : NVW];
    uint2 dd[LATE ? 1 : NVW];
    const uint8_t* base = ring + slot * GROUP_BYTES + warp * ROW_BYTES;
    uint64_t* my_empty = &empty_bar[slot];
    if (live && !LATE) {
#pragma unroll
      for (int i = 0; i < NVW; ++i) {
        xh[i] = *reinterpret_cast<const float4*>(base + (i * 32 + lane) * 16);
        rr[i] = *reinterpret_cast<const float4*>(base + X_BYTES + (i * 32 + lane) * 16);
        dd[i] = *reinterpret_cast<const uint2*>(base + 2 * X_BYTES + (i * 32 + lane) * 8);
      }
    }
    if (!LATE || !live) {
      __syncwarp();
      if (lane == 0) mbar_arrive(my_empty);  // the row is in registers: the slot can be refilled
    }
    if (++slot == depth) { slot = 0; phase ^= 1; }
    if (!live) continue;
    float4 gy[NVW];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NVW; ++i) {
      const uint2 dw = LATE ? *reinterpret_cast<const uint2*>(base + 2 * X_BYTES + (i * 32 + lane) * 8)
                            : dd[LATE ? 0 : i];
      const float2 lo = unpack_bf16(dw.x), hi = unpack_bf16(dw.y);
      const float4 d = make_float4(lo.x, lo.y, hi.x, hi.y);
      const float4 xv = LATE ? *reinterpret_cast<const float4*>(base + (i * 32 + lane) * 16) : xh[i];
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      const float4 g = GM_REGS ? gm[GM_REGS ? i : 0] : __ldg(g4 + i * 32 + lane);
      gy[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
      dg[i].x += d.x * xh[i].x; dg[i].y += d.y * xh[i].y;
      dg[i].z += d.z * xh[i].z; dg[i].w += d.w * xh[i].w;
      db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
      s1 += (gy[i].x + gy[i].y) + (gy[i].z + gy[i].w);
      s2 += (gy[i].x * xh[i].x + gy[i].y * xh[i].y) + (gy[i].z * xh[i].z + gy[i].w * xh[i].w);
    }
    s1 = warp_sum(s1) * inv_cols;
    s2 = warp_sum(s2) * inv_cols;
    float4* orow = reinterpret_cast<float4*>(dx + row * COLS);
    uint2* lrow = reinterpret_cast<uint2*>(dx_lp + row * COLS);
#pragma unroll
    for (int i = 0; i < NVW; ++i) {
      const float4 r = LATE ? *reinterpret_cast<const float4*>(base + X_BYTES + (i * 32 + lane) * 16)
                            : rr[LATE ? 0 : i];
      float4 o;
      o.x = rs * (gy[i].x - s1 - xh[i].x * s2) + r.x;
      o.y = rs * (gy[i].y - s1 - xh[i].y * s2) + r.y;
      o.z = rs * (gy[i].z - s1 - xh[i].z * s2) + r.z;
      o.w = rs * (gy[i].w - s1 - xh[i].w * s2) + r.w;
      orow[i * 32 + lane] = o;
      if (dx_lp != nullptr) {
        uint2 pk;
        pk.x = pack_bf16(o.x * lps, o.y * lps);
        pk.y = pack_bf16(o.z * lps, o.w * lps);
        lrow[i * 32 + lane] = pk;
      }
    }
    if (LATE) {
      __syncwarp();
      if (lane == 0) mbar_arrive(my_empty);
    }
  }
  // dgamma / dbeta: the eight warps' register partials meet in the (now idle) ring
  asm volatile("bar.sync 1, %0;" ::"n"(LNR_ROWS * 32) : "memory");
  float* red_g = reinterpret_cast<float*>(ring);
  float* red_b = red_g + LNR_ROWS * COLS;
#pragma unroll
  for (int i = 0; i < NVW; ++i) {
    reinterpret_cast<float4*>(red_g + warp * COLS)[i * 32 + lane] = dg[i];
    reinterpret_cast<float4*>(red_b + warp * COLS)[i * 32 + lane] = db[i];
  }
  asm volatile("bar.sync 1, %0;" ::"n"(LNR_ROWS * 32) : "memory");
  for (int c = threadIdx.x; c < COLS; c += LNR_ROWS * 32) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int w = 0; w < LNR_ROWS; ++w) {
      sg += red_g[w * COLS + c];
      sb += red_b[w * COLS + c];
    }
    atomicAdd(dgamma + c, sg);
    atomicAdd(dbeta + c, sb);
  }
}

}  // namespace fv

extern "C" int fv_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y,
                                int y_dtype, float* mean, float* rstd, int64_t rows, int64_t cols,
                                float eps, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(x && gamma && beta && y && mean && rstd, "fv_layernorm_fwd: null pointer");
  FV_CHECK_ARG(rows >= 0 && cols > 0 && cols % 4 == 0 && cols <= LN_MAX_VEC * 128,
               "fv_layernorm_fwd: cols=%lld must be a multiple of 4 and <= %d", (long long)cols,
               LN_MAX_VEC * 128);
  FV_CHECK_ARG(y_dtype == FV_F32 || y_dtype == FV_BF16, "fv_layernorm_fwd: bad y_dtype");
  if (rows == 0) return FV_OK;
  const unsigned grid = static_cast<unsigned>(ceil_div(rows, LN_WARPS));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define FV_LN_FWD(NV)                                                                              \
  do {                                                                                             \
    if (y_dtype == FV_BF16)                                                                        \
      FV_CHECK_CUDA(fv::launch_pdl(layernorm_fwd_kernel<NV, true>, dim3(grid), dim3(LN_WARPS * 32), 0, st, x, gamma, beta, y, mean, rstd, \
                                                                     rows, (int)cols, eps));        \
    else                                                                                           \
      FV_CHECK_CUDA(fv::launch_pdl(layernorm_fwd_kernel<NV, false>, dim3(grid), dim3(LN_WARPS * 32), 0, st, x, gamma, beta, y, mean, rstd, \
                                                                      rows, (int)cols, eps));       \
  } while (0)
  if (cols <= 256) FV_LN_FWD(2);
  else if (cols <= 512) FV_LN_FWD(4);
  else if (cols <= 768) FV_LN_FWD(6);
  else FV_LN_FWD(8);
#undef FV_LN_FWD
  FV_LAUNCH_CHECK();
  return FV_OK;
}

extern "C" int fv_layernorm_bwd(const void* dy, int dy_dtype, const float* x, const float* gamma,
                                const float* mean, const float* rstd, const float* dres, float* dx,
                                void* dx_lp, float* dgamma, float* dbeta, int64_t rows, int64_t cols,
                                const float* lp_row_scale, int64_t rows_per_scale, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(dy && x && gamma && mean && rstd && dx && dgamma && dbeta,
               "fv_layernorm_bwd: null pointer");
  FV_CHECK_ARG(rows >= 0 && cols > 0 && cols % 4 == 0 && cols <= LN_MAX_VEC * 128,
               "fv_layernorm_bwd: cols=%lld must be a multiple of 4 and <= %d", (long long)cols,
               LN_MAX_VEC * 128);
  FV_CHECK_ARG(dy_dtype == FV_F32 || dy_dtype == FV_BF16, "fv_layernorm_bwd: bad dy_dtype");
  FV_CHECK_ARG(lp_row_scale == nullptr || (rows_per_scale > 0 && dx_lp != nullptr),
               "fv_layernorm_bwd: lp_row_scale needs dx_lp and rows_per_scale > 0");
  if (rows == 0) return FV_OK;
  static int minb = -1;
  if (minb < 0 || getenv("FEDVIT_LN_REREAD") != nullptr) {
    // FEDVIT_LN_MINB (A/B switch). Default: the bulk-copy ring kernel wherever it applies (bf16 dy with a
    // residual gradient, cols % 128 == 0: 107 us for 50432 x 768 on 100 SMs = 5.8 TB/s, 0.89 of the measured
    // copy bandwidth), else the software-pipelined kernel. 0: pipelined kernel everywhere (131 us, 0.72);
    // 3 / 2: one row per warp pair at a time with three / two CTAs/SM.
    const char* e = getenv("FEDVIT_LN_MINB");
    minb = e ? atoi(e) : 9;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* lp = reinterpret_cast<__nv_bfloat16*>(dx_lp);
  // FEDVIT_LN_GRID (measurement switch, read per call): cap on the number of SMs this pass may use
  int sm_cap = num_sms();
  bool capped = false;
  if (const char* e = getenv("FEDVIT_LN_GRID")) {
    const int v = atoi(e);
    if (v > 0 && v < sm_cap) {
      sm_cap = v;
      capped = true;
    }
  }
  const bool ring_ok = dy_dtype == FV_BF16 && dres != nullptr && cols % 128 == 0 && cols <= 1024 &&
                       ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) |
                         reinterpret_cast<uintptr_t>(dres)) & 15) == 0;
  if (minb == 9 && ring_ok) {
    // bulk-copy ring kernel: as many 8-row groups as fit beside the static barriers
    const size_t group = static_cast<size_t>(LNR_ROWS) * 10 * cols;
    int depth = static_cast<int>((220 * 1024) / group);
    if (depth > LNR_MAX_DEPTH) depth = LNR_MAX_DEPTH;
    const size_t smem_ring = depth * group + 128;
    const int64_t groups = ceil_div(rows, LNR_ROWS);
    // ~180 KB in flight per SM saturates HBM from about 100 SMs; more CTAs only add DRAM page
    // conflicts (measured for 50432 x 768: 96 - 120 SMs 105 - 108 us, 148: 107 - 113, 80: 116, 56: 154)
    if (!capped) sm_cap = (sm_cap * LNR_SM_NUM) / LNR_SM_DEN;
    if (sm_cap < 1) sm_cap = 1;
    const unsigned grid_r = static_cast<unsigned>(groups < sm_cap ? groups : sm_cap);
    const __nv_bfloat16* dyb = reinterpret_cast<const __nv_bfloat16*>(dy);
#define FV_LN_RING(NVW)                                                                                      \
  do {                                                                                                       \
    static bool configured = false;                                                                          \
    if (!configured) {                                                                                       \
      FV_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_ring_kernel<NVW>,                                     \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024));          \
      configured = true;                                                                                     \
    }                                                                                                        \
    FV_CHECK_CUDA(fv::launch_pdl(layernorm_bwd_ring_kernel<NVW>, dim3(grid_r), dim3(LNR_THREADS), smem_ring, \
                                 st, dyb, x, gamma, mean, rstd, dres, dx, lp, dgamma, dbeta,                 \
                                 static_cast<long long>(rows), lp_row_scale, (int)rows_per_scale, depth));   \
  } while (0)
    switch (cols / 128) {
      case 1: FV_LN_RING(1); break;
      case 2: FV_LN_RING(2); break;
      case 3: FV_LN_RING(3); break;
      case 4: FV_LN_RING(4); break;
      case 5: FV_LN_RING(5); break;
      case 6: FV_LN_RING(6); break;
      case 7: FV_LN_RING(7); break;
      default: FV_LN_RING(8); break;
    }
#undef FV_LN_RING
    FV_LAUNCH_CHECK();
    return FV_OK;
  }
  int64_t want = ceil_div(rows, LNB_ROWS);
  const int64_t cap = static_cast<int64_t>(sm_cap) * (minb == 0 || minb == 9 ? 2 : minb);
  const unsigned grid = static_cast<unsigned>(want < cap ? want : cap);
  const size_t smem = 2 * LNB_ROWS * cols * sizeof(float);
#define FV_LN_BWD(NV, BF)                                                                        \
  do {                                                                                           \
    if (minb == 0 || minb == 9)                                                                  \
      FV_CHECK_CUDA(fv::launch_pdl(layernorm_bwd_pipe_kernel<NV, BF>, dim3(grid), dim3(LNB_THREADS), smem, st, dy, x, gamma, mean, rstd, dres, dx, \
                                   lp, dgamma, dbeta, rows, (int)cols, lp_row_scale, (int)rows_per_scale)); \
    else if (minb == 2)                                                                          \
      FV_CHECK_CUDA(fv::launch_pdl(layernorm_bwd_kernel<NV, BF, 2>, dim3(grid), dim3(LNB_THREADS), smem, st, dy, x, gamma, mean, rstd, dres, dx, \
                                   lp, dgamma, dbeta, rows, (int)cols, lp_row_scale, (int)rows_per_scale)); \
    else                                                                                         \
      FV_CHECK_CUDA(fv::launch_pdl(layernorm_bwd_kernel<NV, BF, 3>, dim3(grid), dim3(LNB_THREADS), smem, st, dy, x, gamma, mean, rstd, dres, dx, \
                                   lp, dgamma, dbeta, rows, (int)cols, lp_row_scale, (int)rows_per_scale)); \
  } while (0)
#define FV_LN_BWD_NV(NV)                          \
  do {                                            \
    if (dy_dtype == FV_BF16) FV_LN_BWD(NV, true); \
    else FV_LN_BWD(NV, false);                    \
  } while (0)
  if (cols <= 256) FV_LN_BWD_NV(1);
  else if (cols <= 512) FV_LN_BWD_NV(2);
  else if (cols <= 768) FV_LN_BWD_NV(3);
  else FV_LN_BWD_NV(4);
#undef FV_LN_BWD_NV
#undef FV_LN_BWD
  FV_LAUNCH_CHECK();
  return FV_OK;
}

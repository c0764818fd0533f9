// gemm_tc.cu — bf16 GEMM on 5th-gen tensor cores (tcgen05.mma, TMEM accumulators, TMA operands).
//
// Replaces the nn.Linear GEMMs inside timm's Attention / Mlp (reference call site model.py:193)
// and their autograd backward (train.py:153): forward (K-major x K-major), dgrad (K-major x
// MN-major weights) and wgrad (MN-major x MN-major, split-K, fp32 accumulate).
//
// Kernel shape: persistent, one CTA per SM, 256 threads, warp-specialised
//   warp 0      TMA producer   (one thread)            global -> 4-stage smem ring, 128B swizzle
//   warp 1      MMA issuer     (one thread)            128 x 256 x 16 tcgen05.mma, fp32 in TMEM
//   warp 2      TMEM allocator (512 columns = 2 accumulator stages of 256 columns)
//   warps 4..11 epilogue       (two warps per TMEM lane quarter, 128 columns each: one warp per
//                              SMSP left the XU/FMA pipes latency-bound on the GELU epilogues)
//                              tcgen05.ld (thread == accumulator row) -> bias/GELU/residual in
//                              registers -> 128-byte row segments transposed through a swizzled
//                              per-warp smem tile -> fully coalesced 128-bit global accesses
//                              (every store instruction writes 4 complete 128 B lines). Residual /
//                              pre-activation inputs take the same path in reverse and are
//                              prefetched one segment ahead so their latency hides behind the MMAs.
// The epilogue of tile i overlaps the main loop of tile i+1 through the two TMEM stages.
#include <stdlib.h>

#include "common.cuh"

namespace fv {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KiB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int GEMM_THREADS = 384;   // 4 control warps + 8 epilogue warps
constexpr int EPI_WARPS = 8;
constexpr int EPI_STAGE_BYTES = 32 * 128;  // per epilogue warp: 32 rows x one 128-byte segment
constexpr int GEMM_SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * EPI_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*colsum reduce*/;

struct GemmTcParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  int split_k, kb_per_split;
  int a_major, b_major;
  int c_bf16;
  int tokens_per_img;
  const float* bias;
  void* c;
  long long ldc;
  void* aux;
  long long ldaux;
  float* colsum;  // wgrad only: += column sums of op(A) (the bias gradient), or nullptr
  // stochastic depth: RESIDUAL epilogue computes R + row_scale[row / rows_per_scale] * (acc + bias)
  const float* row_scale;
  int rows_per_scale;
  // im2col-free patch embedding (FV_EPI_PATCH with im2col == 1): the A operand is the NCHW fp32
  // image itself, fetched by a 5-D TMA map; an M tile is `ph_per_tile` rows of patches of one image
  int im2col;
  int gw, gh, chans;       // patch grid and input channels
  int ph_per_tile, tiles_per_img;
  int tma_out;  // outputs leave through TMA tensor stores (everything but the PATCH row remap)
  int direct;  // epilogue without smem staging (needs 32-byte aligned rows of c / aux)
  // FEDVIT_GEMM_DBG (measurement / test switches): 1 = epilogue without global stores, 2 = no epilogue work,
  // 4 = direct (unstaged) epilogue, 8 = bias-gradient protocol without the reads, 16 = LSU flush instead of
  // TMA stores, 32 = never use the CTA-pair kernel, 64 = use it for every legal shape (small test problems),
  // 256 = weight gradients stay on the single-CTA kernel
  int dbg;
};

// PATCH epilogue row mapping: local row r of M tile m_blk -> (valid?, global output row, pos row)
__device__ __forceinline__ bool patch_row(const GemmTcParams& p, long long row0, int r, long long& orow,
                                          long long& arow) {
  if (p.im2col) {
    const int m_blk = static_cast<int>(row0 / BM);  // row0 is m_blk*BM + warp quarter offset
    const int lr = static_cast<int>(row0 - static_cast<long long>(m_blk) * BM) + r;
    const int img = m_blk / p.tiles_per_img;
    const int ph0 = (m_blk - img * p.tiles_per_img) * p.ph_per_tile;
    int phs = p.gh - ph0;
    if (phs > p.ph_per_tile) phs = p.ph_per_tile;
    if (lr >= phs * p.gw) return false;
    arow = static_cast<long long>(ph0) * p.gw + lr + 1;            // token index (0 is cls)
    orow = static_cast<long long>(img) * (p.gw * p.gh + 1) + arow;
    return true;
  }
  const long long row = row0 + r;
  if (row >= p.M) return false;
  arow = row % p.tokens_per_img + 1;          // pos_embed row of this token
  orow = row + row / p.tokens_per_img + 1;    // row 0 of each image is cls
  return true;
}

// exact-GELU pieces on packed fp32 pairs. Phi(x) by Abramowitz-Stegun 26.2.17 (|err| <= 7.5e-8):
//   1 - Phi(|x|) = phi(|x|) * (b1 t + ... + b5 t^5),  t = 1 / (1 + 0.2316419 |x|),
// sharing the one exponential phi(x) = exp(-x^2/2)/sqrt(2 pi) between the cdf and the derivative.
// The two transcendental steps are single MUFU instructions (rcp.approx.ftz / ex2.approx.ftz, no
// range-fixup code); everything else is FFMA2 / FMUL2 / FADD2, one FMA-pipe issue per TWO
// elements. History (profiles/): erff()/__expf() 3 MUFU + ~25 FP32 ops per element made the fc1 /
// fc2-dgrad epilogues XU-bound; the scalar A&S form (17 FMA-pipe ops per element) left them
// FMA-pipe-bound at K = 768; the packed form needs ~7 FMA-pipe issues per element.
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t copysign_bits(float mag, float sgn) {
  return (__float_as_uint(mag) & 0x7FFFFFFFu) | (__float_as_uint(sgn) & 0x80000000u);
}
// (u0, u1) -> cdf = Phi(u), pdf = phi(u), both as pairs
__device__ __forceinline__ void gelu_parts2(float u0, float u1, f32x2 u, f32x2& cdf, f32x2& pdf) {
  const f32x2 a = f2_pack(fabsf(u0), fabsf(u1));
  const f32x2 d = f2_fma(a, f2_splat(0.2316419f), f2_splat(1.0f));
  float d0, d1, g0, g1;
  f2_unpack(d, d0, d1);
  const f32x2 t = f2_pack(rcp_ftz(d0), rcp_ftz(d1));
  // phi(u) = 2^(-u^2 * log2(e)/2 + log2(1/sqrt(2 pi)))
  const f32x2 arg = f2_fma(f2_mul(u, u), f2_splat(-0.72134752044448170368f), f2_splat(-1.3257480647361593f));
  f2_unpack(arg, g0, g1);
  pdf = f2_pack(ex2_ftz(g0), ex2_ftz(g1));
  f32x2 poly = f2_fma(f2_splat(1.330274429f), t, f2_splat(-1.821255978f));
  poly = f2_fma(poly, t, f2_splat(1.781477937f));
  poly = f2_fma(poly, t, f2_splat(-0.356563782f));
  poly = f2_fma(poly, t, f2_splat(0.319381530f));
  const f32x2 h = f2_mul(f2_mul(poly, t), pdf);                      // 1 - Phi(|u|), in [0, 0.5]
  const f32x2 g = f2_fma(h, f2_splat(-1.0f), f2_splat(0.5f));        // Phi(|u|) - 0.5
  f2_unpack(g, g0, g1);
  cdf = f2_add(f2_pack_u(copysign_bits(g0, u0), copysign_bits(g1, u1)), f2_splat(0.5f));
}

// ---- per-warp staging tile: 32 rows x 128 B, 16-byte units XOR-swizzled by (row & 7) -----------
// thread-per-row accesses (lane == row) and row-coalesced accesses (8 lanes per row) are both
// bank-conflict free.
__device__ __forceinline__ uint4* stg_unit(uint8_t* stg, int row, int unit) {
  return reinterpret_cast<uint4*>(stg + row * 128 + ((unit ^ (row & 7)) << 4));
}

struct SegGeom {       // one 128-byte output segment of this warp's 32 rows
  long long row0;      // first global row of the warp
  int col0;            // first column of the segment (elements)
  int elem;            // bytes per element of the matrix being moved
};

// Row-coalesced access pattern shared by the aux prefetch and the flush: lane -> (row lane/8 + 4*i,
// physical 16-byte slot lane%8) for i = 0..7. With r = 4*i + lane/8 the swizzled logical unit is
// (lane%8) ^ (r&7), which only depends on the parity of i — so two global pointers per lane (even /
// odd i), each advancing by 8 rows, replace all per-iteration index arithmetic.
struct RowWalk {
  uint8_t* pe;       // global address of (row0 + lr,     unit ue)
  uint8_t* po;       // global address of (row0 + lr + 4, unit uo)
  long long step;    // bytes per 8 rows
  int lr;            // lane / 8
  bool ce, co;       // column predicates of the two units
};
template <int EPI, bool AUX_ROWS>
__device__ __forceinline__ RowWalk make_walk(void* base, long long ld, const GemmTcParams& p, const SegGeom& g,
                                             int lane) {
  RowWalk w;
  const int upe = 16 / g.elem;
  w.lr = lane >> 3;
  const int ue = (lane & 7) ^ w.lr, uo = ue ^ 4;
  w.ce = g.col0 + ue * upe < p.N;
  w.co = g.col0 + uo * upe < p.N;
  w.step = 8 * ld * g.elem;
  // PATCH remaps rows (aux: pos_embed row of the token; out: one cls slot per image) — that
  // epilogue recomputes its addresses per row in the callers below, so only the plain layout here.
  w.pe = reinterpret_cast<uint8_t*>(base) + ((g.row0 + w.lr) * ld + g.col0 + ue * upe) * g.elem;
  w.po = reinterpret_cast<uint8_t*>(base) + ((g.row0 + w.lr + 4) * ld + g.col0 + uo * upe) * g.elem;
  return w;
}

// global -> registers, coalesced
template <int EPI>
__device__ __forceinline__ void aux_prefetch(uint4 (&pre)[8], const GemmTcParams& p, const SegGeom& g,
                                             int lane) {
  if (EPI == FV_EPI_PATCH) {
    const int upe = 16 / g.elem;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + (lane >> 3);
      const int unit = (lane & 7) ^ (r & 7);
      const int col = g.col0 + unit * upe;
      long long orow, arow;
      pre[i] = make_uint4(0u, 0u, 0u, 0u);
      if (patch_row(p, g.row0, r, orow, arow) && col < p.N)
        pre[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.aux) +
                                                      (arow * p.ldaux + col) * g.elem));
    }
    return;
  }
  const RowWalk w = make_walk<EPI, true>(p.aux, p.ldaux, p, g, lane);
  const long long rows_left = p.M - g.row0 - w.lr;  // row (row0+lr+4*i) is valid iff 4*i < rows_left
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    pre[2 * j] = (w.ce && 8 * j < rows_left) ? __ldg(reinterpret_cast<const uint4*>(w.pe + j * w.step))
                                             : make_uint4(0u, 0u, 0u, 0u);
    pre[2 * j + 1] = (w.co && 8 * j + 4 < rows_left) ? __ldg(reinterpret_cast<const uint4*>(w.po + j * w.step))
                                                     : make_uint4(0u, 0u, 0u, 0u);
  }
}
__device__ __forceinline__ void aux_to_stage(const uint4 (&pre)[8], uint8_t* stg, int lane) {
  uint8_t* s0 = stg + (lane >> 3) * 128 + ((lane & 7) << 4);
#pragma unroll
  for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(s0 + i * 512) = pre[i];
}
// staging -> global, coalesced; store, or red.add.f32 for the weight-gradient accumulate
template <int EPI>
__device__ __forceinline__ void stage_flush(uint8_t* stg, void* base, long long ld, const GemmTcParams& p,
                                            const SegGeom& g, int lane) {
  const uint8_t* s0 = stg + (lane >> 3) * 128 + ((lane & 7) << 4);
  if (EPI == FV_EPI_PATCH) {
    const int upe = 16 / g.elem;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + (lane >> 3);
      const int unit = (lane & 7) ^ (r & 7);
      const int col = g.col0 + unit * upe;
      long long orow, arow;
      if (patch_row(p, g.row0, r, orow, arow) && col < p.N)
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(base) + (orow * ld + col) * g.elem) =
            *reinterpret_cast<const uint4*>(s0 + i * 512);
    }
    return;
  }
  const RowWalk w = make_walk<EPI, false>(base, ld, p, g, lane);
  const long long rows_left = p.M - g.row0 - w.lr;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const bool ok = h == 0 ? (w.ce && 8 * j < rows_left) : (w.co && 8 * j + 4 < rows_left);
      if (ok) {
        const uint4 v = *reinterpret_cast<const uint4*>(s0 + (2 * j + h) * 512);
        uint8_t* dst = (h == 0 ? w.pe : w.po) + j * w.step;
        if (EPI == FV_EPI_ACCUM) {
          asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(__uint_as_float(v.x)),
                       "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                       : "memory");
        } else {
          *reinterpret_cast<uint4*>(dst) = v;
        }
      }
    }
  }
}

// One 32-column chunk of this thread's row as 16 packed pairs: acc (+bias) (+aux). `mine` holds
// this row's aux units of the chunk.
template <int EPI, bool BF16>
__device__ __forceinline__ void chunk_math(f32x2 (&v)[16], const GemmTcParams& p, const float4 (&bv)[8],
                                           const uint4* mine /*this chunk's aux units or nullptr*/,
                                           float rscale) {
  if (EPI != FV_EPI_ACCUM && EPI != FV_EPI_DGELU) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      v[q * 2] = f2_add(v[q * 2], f2_pack(bv[q].x, bv[q].y));
      v[q * 2 + 1] = f2_add(v[q * 2 + 1], f2_pack(bv[q].z, bv[q].w));
    }
  }
  if (EPI == FV_EPI_RESIDUAL && p.row_scale != nullptr) {  // drop-path: scale the branch, not the residual
    const f32x2 rs = f2_splat(rscale);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = f2_mul(v[i], rs);
  }
  if (EPI == FV_EPI_RESIDUAL || EPI == FV_EPI_PATCH) {  // fp32 aux, 8 units per chunk
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      v[q * 2] = f2_add(v[q * 2], f2_pack_u(mine[q].x, mine[q].y));
      v[q * 2 + 1] = f2_add(v[q * 2 + 1], f2_pack_u(mine[q].z, mine[q].w));
    }
  } else if (EPI == FV_EPI_DGELU) {
    if (BF16) {  // bf16 gelu'(u) saved by the forward GELU epilogue, 4 units per chunk
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t w[4] = {mine[q].x, mine[q].y, mine[q].z, mine[q].w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          v[q * 4 + j] = f2_mul(v[q * 4 + j], f2_pack_u(w[j] << 16, w[j] & 0xFFFF0000u));
      }
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        v[q * 2] = f2_mul(v[q * 2], f2_pack_u(mine[q].x, mine[q].y));
        v[q * 2 + 1] = f2_mul(v[q * 2 + 1], f2_pack_u(mine[q].z, mine[q].w));
      }
    }
  }
}

__device__ __forceinline__ uint32_t pack_bf16(f32x2 v) {
  float lo, hi;
  f2_unpack(v, lo, hi);
  return pack_bf16(lo, hi);
}

// write this thread's 32 values of a chunk into its staging row, starting at 16-byte unit `u0`
__device__ __forceinline__ void chunk_to_stage(const f32x2 (&v)[16], uint8_t* stg, int lane, int u0,
                                               bool bf16) {
  if (bf16) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *stg_unit(stg, lane, u0 + q) =
          make_uint4(pack_bf16(v[q * 4]), pack_bf16(v[q * 4 + 1]), pack_bf16(v[q * 4 + 2]), pack_bf16(v[q * 4 + 3]));
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint4 w;
      asm("mov.b64 {%0, %1}, %2;" : "=r"(w.x), "=r"(w.y) : "l"(v[q * 2]));
      asm("mov.b64 {%0, %1}, %2;" : "=r"(w.z), "=r"(w.w) : "l"(v[q * 2 + 1]));
      *stg_unit(stg, lane, u0 + q) = w;
    }
  }
}

#ifdef GEMM_TRACE
// measurement build (tools/build_variants.py gemm_tc.cu gtrace:-DGEMM_TRACE, tools/gemm_trace.py): SM-clock
// timestamps of epilogue warp 4 of CTA 0 over its first 24 tiles — 0 arrives at the accumulator wait, 1 accumulator
// there, 2 tile drained; 3 = cycles of that tile spent waiting for a tensor store to release the staging tile,
// 4 = cycles between a tcgen05.ld and its data, 5 = cycles in the chunks' arithmetic (bias, activation), 6 = cycles in
// stage_store (proxy fence, warp sync, tensor-store issue)
__device__ long long g_gemm_trace[24][8];
__device__ long long g_gemm_acq, g_gemm_ldw, g_gemm_math, g_gemm_store;
#define GT_ON (blockIdx.x == 0 && threadIdx.x == 128)
__device__ __forceinline__ long long gt_clock() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
  return t;
}
#endif
// Epilogue of one 128 x 256 accumulator tile for one warp (its 32 TMEM lanes / rows).
// Output path: each warp's staged 32-row x 128-byte segment (already in the TMA 128B-swizzle
// layout) leaves through ONE tensor store issued by lane 0 — cp.async.bulk.tensor (plain outputs)
// or cp.reduce.async.bulk.tensor .add (split-K weight-gradient accumulate) — instead of 8 LDS + 8
// STG/RED per lane: ncu showed the epilogue warps, not the tensor pipe, on the critical path of the
// K = 768 shapes, a third of their stall samples in that flush. `pending` = this warp has a store
// in flight that may still be reading the staging tile.
template <int EPI>
__device__ __forceinline__ void stage_store(const GemmTcParams& p, const CUtensorMap* map, uint8_t* stg, void* base,
                                            long long ld, const SegGeom& g, int lane, bool& pending) {
  if (p.dbg & 1) return;
  if (EPI == FV_EPI_PATCH || !p.tma_out) {
    stage_flush<EPI>(stg, base, ld, p, g, lane);
    return;
  }
#ifdef GEMM_TRACE
  const long long ts0 = gt_clock();
#endif
  fence_proxy_async_smem();  // generic-proxy writes of the tile -> visible to the TMA engine
  __syncwarp();
  if (lane == 0) {
    if (EPI == FV_EPI_ACCUM) {
      tma_reduce_add_2d(map, stg, g.col0, static_cast<int>(g.row0));
    } else {
      tma_store_2d(map, stg, g.col0, static_cast<int>(g.row0));
    }
    tma_store_commit();
  }
  pending = true;
#ifdef GEMM_TRACE
  if (GT_ON) g_gemm_store += gt_clock() - ts0;
#endif
}
__device__ __forceinline__ void stage_acquire(int lane, bool& pending) {
  if (pending) {
#ifdef GEMM_TRACE
    const long long t0 = gt_clock();
#endif
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
    pending = false;
#ifdef GEMM_TRACE
    if (GT_ON) g_gemm_acq += gt_clock() - t0;
#endif
  }
}

template <int EPI, bool BF16>
__device__ __forceinline__ void epilogue_tile(const GemmTcParams& p, const CUtensorMap* map_c,
                                              const CUtensorMap* map_aux, uint8_t* stg, uint32_t taddr,
                                              long long row0, int n0, int width, int lane, bool& pending) {
  constexpr bool HAS_AUX_IN = (EPI == FV_EPI_RESIDUAL || EPI == FV_EPI_DGELU || EPI == FV_EPI_PATCH);
  constexpr bool bf16 = BF16;  // output type (and the type of the DGELU aux input)
  constexpr int seg_cols = bf16 ? 64 : 32;          // columns per 128-byte segment
  constexpr int chunks_per_seg = bf16 ? 2 : 1;
  constexpr int aux_elem = (EPI == FV_EPI_DGELU) ? (bf16 ? 2 : 4) : 4;
  constexpr int aux_units_per_chunk = 32 * aux_elem / 16;
  int nseg = (p.N - n0 + seg_cols - 1) / seg_cols;
  if (nseg > width / seg_cols) nseg = width / seg_cols;

  if (p.dbg & 2) return;
  float rscale = 1.0f;
  if (EPI == FV_EPI_RESIDUAL && p.row_scale != nullptr) {
    const long long row = row0 + lane;  // this thread's accumulator row
    rscale = row < p.M ? __ldg(p.row_scale + row / p.rows_per_scale) : 0.f;
  }
  uint4 pre[8];
  SegGeom ga{row0, n0, aux_elem};
  if (HAS_AUX_IN) aux_prefetch<EPI>(pre, p, ga, lane);

  for (int s = 0; s < nseg; ++s) {
    uint4 mine[8];
    if (HAS_AUX_IN) {
      // bounce the prefetched aux segment through the staging tile so each thread gets its row
      stage_acquire(lane, pending);
      aux_to_stage(pre, stg, lane);
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 8; ++q) mine[q] = *stg_unit(stg, lane, q);
      __syncwarp();
      if (s + 1 < nseg) {
        ga.col0 = n0 + (s + 1) * seg_cols;
        aux_prefetch<EPI>(pre, p, ga, lane);  // in flight while this segment is computed and stored
      }
    }
    const SegGeom go{row0, n0 + s * seg_cols, bf16 ? 2 : 4};
    uint32_t gbuf[32];  // GELU only: the segment's activations, already packed for the store
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (c < chunks_per_seg) {
        const int col0 = go.col0 + c * 32;
        uint32_t r[32];
        tmem_ld_32x32(taddr + (s * seg_cols + c * 32), r);
        // the chunk's 32 bias values (a broadcast read) are fetched while the TMEM load is in flight
        float4 bv[8];
        if (EPI != FV_EPI_ACCUM && EPI != FV_EPI_DGELU) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            bv[q] = (p.bias != nullptr && col0 + q * 4 < p.N)
                        ? __ldg(reinterpret_cast<const float4*>(p.bias + col0) + q)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#ifdef GEMM_TRACE
        const long long tl0 = gt_clock();
#endif
        tmem_ld_wait();
#ifdef GEMM_TRACE
        if (GT_ON) g_gemm_ldw += gt_clock() - tl0;
#endif
        f32x2 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = f2_pack_u(r[2 * i], r[2 * i + 1]);
        chunk_math<EPI, BF16>(v, p, bv, HAS_AUX_IN ? &mine[c * aux_units_per_chunk] : nullptr, rscale);
        if (EPI == FV_EPI_GELU) {
          // Activation AND its derivative from the one cdf/pdf evaluation: out = u*Phi(u) goes to
          // gbuf (stored last), aux = Phi(u) + u*phi(u) replaces the pre-activation as the tensor
          // kept for the backward — the fc2 dgrad epilogue then only multiplies (FV_EPI_DGELU).
          // u is first rounded to the output type, as autocast does (fc1 emits bf16 and nn.GELU
          // runs on that tensor): round by packing the pair (one F2FP per two elements) and
          // unpacking with shifts, not one F2F per element.
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float u0, u1;
            f2_unpack(v[i], u0, u1);
            f32x2 u = v[i];
            if (bf16) {
              const float2 ur = unpack_bf16(pack_bf16(u0, u1));
              u0 = ur.x;
              u1 = ur.y;
              u = f2_pack(u0, u1);
            }
            f32x2 cdf, pdf;
            gelu_parts2(u0, u1, u, cdf, pdf);
            const f32x2 act = f2_mul(u, cdf);
            if (p.aux == nullptr) {  // forward-only: no derivative, the activation takes the staging tile directly
              v[i] = act;
              continue;
            }
            v[i] = f2_fma(u, pdf, cdf);
            if (bf16) {
              gbuf[c * 16 + i] = pack_bf16(act);
            } else {
              asm("mov.b64 {%0, %1}, %2;" : "=r"(gbuf[2 * i]), "=r"(gbuf[2 * i + 1]) : "l"(act));
            }
          }
        }
#ifdef GEMM_TRACE
        if (GT_ON) g_gemm_math += gt_clock() - tl0;
#endif
        if (c == 0) stage_acquire(lane, pending);
        chunk_to_stage(v, stg, lane, c * (bf16 ? 4 : 8), bf16);
      }
    }
    __syncwarp();
    if (EPI == FV_EPI_GELU && p.aux != nullptr) {
      stage_store<EPI>(p, map_aux, stg, p.aux, p.ldaux, go, lane, pending);  // gelu'(u) -> aux
      __syncwarp();
      stage_acquire(lane, pending);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *stg_unit(stg, lane, q) = make_uint4(gbuf[q * 4], gbuf[q * 4 + 1], gbuf[q * 4 + 2], gbuf[q * 4 + 3]);
      __syncwarp();
    }
    stage_store<EPI>(p, map_c, stg, p.c, p.ldc, go, lane, pending);
    __syncwarp();
  }
}

// ---- direct epilogue: no shared-memory staging -------------------------------------------------
// thread == accumulator row; every 32-column chunk leaves (and its aux operand arrives) as 256-bit
// per-thread accesses, each one whole 32-byte sector of the thread's own row. The MMA's operand
// reads keep ~3/4 of the shared-memory bandwidth busy at full rate, so the staged transpose (write
// + read of every output byte through smem) competes with the tensor core; this path does not
// touch shared memory at all.
// store 32 columns (already packed: 16 words bf16 / 32 words fp32) of one row, clipped at N (N % 8 == 0)
template <bool BF16>
__device__ __forceinline__ void store_chunk(void* base, long long ld, long long row, int col0, int N,
                                            const uint32_t* w) {
  if (BF16) {
    uint8_t* dst = reinterpret_cast<uint8_t*>(base) + (row * ld + col0) * 2;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (col0 + j * 16 + 16 <= N) {
        st_v8(dst + j * 32, w + j * 8);
      } else if (col0 + j * 16 + 8 <= N) {
        *reinterpret_cast<uint4*>(dst + j * 32) = make_uint4(w[j * 8], w[j * 8 + 1], w[j * 8 + 2], w[j * 8 + 3]);
      }
    }
  } else {
    uint8_t* dst = reinterpret_cast<uint8_t*>(base) + (row * ld + col0) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (col0 + j * 8 + 8 <= N) st_v8(dst + j * 32, w + j * 8);
  }
}
// load the aux operand of one chunk of one row: AUX_BF16 ? 16 words : 32 words (zero past N / M)
template <bool AUX_BF16>
__device__ __forceinline__ void load_chunk(uint32_t* w, const void* base, long long ld, long long row, bool row_ok,
                                           int col0, int N) {
  constexpr int PIECES = AUX_BF16 ? 2 : 4, CPP = AUX_BF16 ? 16 : 8, ELEM = AUX_BF16 ? 2 : 4;
  const uint8_t* src = reinterpret_cast<const uint8_t*>(base) + (row * ld + col0) * ELEM;
#pragma unroll
  for (int j = 0; j < PIECES; ++j) {
    uint32_t (&dst)[8] = *reinterpret_cast<uint32_t(*)[8]>(w + j * 8);
    if (row_ok && col0 + j * CPP + CPP <= N) {
      ld_v8(dst, src + j * 32);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[i] = 0u;
      if (AUX_BF16 && row_ok && col0 + j * CPP + 8 <= N) {
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(src + j * 32));
        dst[0] = h.x; dst[1] = h.y; dst[2] = h.z; dst[3] = h.w;
      }
    }
  }
}

template <int EPI, bool BF16>
__device__ __forceinline__ void epilogue_tile_direct(const GemmTcParams& p, uint32_t taddr, long long row0, int n0,
                                                     int width, int lane) {
  constexpr bool HAS_AUX_IN = (EPI == FV_EPI_RESIDUAL || EPI == FV_EPI_DGELU);
  constexpr bool AUX_BF16 = (EPI == FV_EPI_DGELU) && BF16;
  constexpr int AUX_WORDS = AUX_BF16 ? 16 : 32;
  const long long row = row0 + lane;
  const bool row_ok = row < p.M;
  int nchunk = (p.N - n0 + 31) / 32;
  if (nchunk > width / 32) nchunk = width / 32;
  float rscale = 1.0f;
  if (EPI == FV_EPI_RESIDUAL && p.row_scale != nullptr) rscale = row_ok ? __ldg(p.row_scale + row / p.rows_per_scale) : 0.f;
  uint32_t pre[HAS_AUX_IN ? AUX_WORDS : 1];
  if (HAS_AUX_IN && nchunk > 0) load_chunk<AUX_BF16>(pre, p.aux, p.ldaux, row, row_ok, n0, p.N);
  for (int c = 0; c < nchunk; ++c) {
    const int col0 = n0 + c * 32;
    uint32_t r[32];
    tmem_ld_32x32(taddr + c * 32, r);
    float4 bv[8];
    if (EPI != FV_EPI_DGELU) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        bv[q] = (p.bias != nullptr && col0 + q * 4 < p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + col0) + q)
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint4 mine[8];
    if (HAS_AUX_IN) {
#pragma unroll
      for (int q = 0; q < AUX_WORDS / 4; ++q) mine[q] = make_uint4(pre[q * 4], pre[q * 4 + 1], pre[q * 4 + 2], pre[q * 4 + 3]);
      if (c + 1 < nchunk) load_chunk<AUX_BF16>(pre, p.aux, p.ldaux, row, row_ok, col0 + 32, p.N);  // next chunk in flight
    }
    tmem_ld_wait();
    f32x2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = f2_pack_u(r[2 * i], r[2 * i + 1]);
    chunk_math<EPI, BF16>(v, p, bv, HAS_AUX_IN ? mine : nullptr, rscale);
    uint32_t w[BF16 ? 16 : 32];
    if (EPI == FV_EPI_GELU) {  // see epilogue_tile: activation -> c, derivative -> aux
      uint32_t g[BF16 ? 16 : 32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float u0, u1;
        f2_unpack(v[i], u0, u1);
        f32x2 u = v[i];
        if (BF16) {
          const float2 ur = unpack_bf16(pack_bf16(u0, u1));
          u0 = ur.x;
          u1 = ur.y;
          u = f2_pack(u0, u1);
        }
        f32x2 cdf, pdf;
        gelu_parts2(u0, u1, u, cdf, pdf);
        const f32x2 grad = f2_fma(u, pdf, cdf), act = f2_mul(u, cdf);
        if (BF16) {
          w[i] = pack_bf16(act);
          g[i] = pack_bf16(grad);
        } else {
          asm("mov.b64 {%0, %1}, %2;" : "=r"(w[2 * i]), "=r"(w[2 * i + 1]) : "l"(act));
          asm("mov.b64 {%0, %1}, %2;" : "=r"(g[2 * i]), "=r"(g[2 * i + 1]) : "l"(grad));
        }
      }
      if (row_ok && !(p.dbg & 1)) store_chunk<BF16>(p.aux, p.ldaux, row, col0, p.N, g);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (BF16) {
          w[i] = pack_bf16(v[i]);
        } else {
          asm("mov.b64 {%0, %1}, %2;" : "=r"(w[2 * i]), "=r"(w[2 * i + 1]) : "l"(v[i]));
        }
      }
    }
    if (row_ok && !(p.dbg & 1)) store_chunk<BF16>(p.c, p.ldc, row, col0, p.N, w);
  }
}

template <int EPI, bool BF16>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_aux,
               const GemmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic on the __shared__ array (not an integer round trip) keeps the address space
  // known to the compiler: LDS / STS instead of generic LD / ST for every staging access
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* epi_stage = smem + STAGES * STAGE_BYTES;  // 8 x 4 KiB, 1024-byte aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + EPI_WARPS * EPI_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (p.tma_out) {
      tma_prefetch_desc(&tmap_c);
      if (EPI == FV_EPI_GELU) tma_prefetch_desc(&tmap_aux);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      // in bias-gradient mode the epilogue warps also read the A tiles: 1 MMA commit + 8 warp arrivals
      // bias-gradient mode: the MMA commit plus the arrival of the stage's column-sum agent warp
      mbar_init(&empty_bar[s], p.colsum != nullptr ? 2 : 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], EPI_WARPS * 32);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // nothing above touches memory another kernel produced

  const int tiles_mn = p.num_m_blocks * p.num_n_blocks;
  const int total_tiles = tiles_mn * p.split_k;

  // Producer and MMA roles: the WHOLE warp walks the loop and one elected lane issues. With
  // warp-uniform control flow the descriptors / coordinates live in uniform registers and every
  // TMA / tcgen05.mma is one predicated instruction; issued from inside an `if (lane == 0)` region
  // the compiler wraps each of them in a vote + broadcast loop (~25 dependent instructions per MMA
  // — as long as the MMA itself runs, which made the single issuing thread the mainloop's limiter).
  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = tile / tiles_mn;
      const int mn = tile - split * tiles_mn;
      const int m_blk = mn / p.num_n_blocks;
      const int n_blk = mn - m_blk * p.num_n_blocks;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        uint8_t* sb = sa + A_STAGE_BYTES;
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage],
                         (EPI == FV_EPI_PATCH && p.im2col) ? 2 * p.gw * p.ph_per_tile * 64 + BN * 128 : STAGE_BYTES);
          if (EPI == FV_EPI_PATCH && p.im2col) {
            // One pixel row of a patch is 16 fp32 = 64 B — the longest contiguous run an NCHW image
            // offers a patch — so an A box is [tokens x 64 B] (5-D map walking px, -, pw, ph,
            // image*channel: the smem rows come out in token order; 64-byte swizzle). A stage holds
            // TWO such pixel rows (k-block pair kb: channel kb/8, rows 2*(kb%8) and +1) in the two 8 KiB
            // halves of the A slot, and the 32 matching weight columns as ONE [256 x 128 B] box
            // (consecutive pixel rows are contiguous in the conv weight; 128-byte swizzle): half the
            // stage hand-shakes and a third fewer 64-byte TMA rows (batch 256, timed alone: 231 -> 224 us
            // with 3 input channels, 251 -> 235 us with 4).
            const int img = m_blk / p.tiles_per_img;
            const int ph0 = (m_blk - img * p.tiles_per_img) * p.ph_per_tile;
            const int py = (kb & 7) * 2, ch = img * p.chans + (kb >> 3);
            tma_load_5d(sa, &tmap_a, &full_bar[stage], 0, py, 0, ph0, ch);
            tma_load_5d(sa + A_STAGE_BYTES / 2, &tmap_a, &full_bar[stage], 0, py + 1, 0, ph0, ch);
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * 32, n_blk * BN);
          } else {
            if (p.a_major == FV_MAJOR_K) {
              tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_2d(sa + j * (64 * BK * 2), &tmap_a, &full_bar[stage], m_blk * BM + j * 64, kb * BK);
            }
            if (p.b_major == FV_MAJOR_K) {
              tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d(sb + j * (64 * BK * 2), &tmap_b, &full_bar[stage], n_blk * BN + j * 64, kb * BK);
            }
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------------------
    const bool tf32 = (EPI == FV_EPI_PATCH) && p.im2col;
    const uint32_t idesc = make_idesc(tf32 ? kFmtTF32 : kFmtBF16, p.a_major, p.b_major, BM, BN);
    // K-major: rows are 128 B apart, 8-row groups 1024 B apart; a K=16 step is 32 B along the row.
    // MN-major: 64-element column blocks are one [BK x 128 B] TMA box (8192 B) apart (LBO),
    //           8-k-row groups 1024 B apart (SBO); a K=16 step is 16 rows = 2048 B.
    const uint32_t a_lbo = p.a_major == FV_MAJOR_K ? 16 : 64 * BK * 2;
    const uint32_t b_lbo = p.b_major == FV_MAJOR_K ? 16 : 64 * BK * 2;
    const uint32_t a_kstep = (p.a_major == FV_MAJOR_K ? 32 : 2048) >> 4;
    const uint32_t b_kstep = (p.b_major == FV_MAJOR_K ? 32 : 2048) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = tile / tiles_mn;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t sb = sa + A_STAGE_BYTES;
        const uint64_t da = make_smem_desc_sw128(sa, a_lbo, 1024);
        const uint64_t db = make_smem_desc_sw128(sb, b_lbo, 1024);
        if (elect_one()) {
          if (tf32) {
            // A: two [tokens x 64 B] halves (64-byte swizzle), two K = 8 steps of 32 bytes in each;
            // B: [256 x 128 B] (128-byte swizzle), the four K = 8 steps 32 bytes apart along the row
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint64_t da64 = make_smem_desc_sw64(sa + j * (A_STAGE_BYTES / 2), 16, 512);
#pragma unroll
              for (int k = 0; k < 2; ++k)
                umma_tf32(tmem_d, da64 + k * 2, db + (j * 2 + k) * 2, idesc, (kb > kb0 || j > 0 || k > 0) ? 1u : 0u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_d, da + k * a_kstep, db + k * b_kstep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if ((warp == 2 || warp == 3) && EPI == FV_EPI_ACCUM && p.colsum != nullptr) {
    // ------------------------------- bias-gradient agents ------------------------------------
    // Fused into the weight-gradient GEMM: while the tensor core consumes the dY tiles (A operand,
    // MN-major: [64 token rows x 128 output columns] per stage, two 128B-swizzled boxes), the two
    // otherwise idle control warps add up their columns. Warp 2 owns ring stages 0 and 2, warp 3
    // stages 1 and 3; an agent takes part in EVERY phase of its stages (a parity wait must never
    // skip a phase) but reads only this CTA's share: the CTAs that work on the same row block (one
    // per n_blk, same dY tiles) split its k-slices round-robin. One stage = 32 conflict-free
    // LDS.128 per lane (8 columns of one box row). The epilogue warps are not involved, so the ring
    // never waits for a tile's accumulator drain. (First version: all 8 epilogue warps polled and
    // arrived on every stage — 12 % slower weight-gradient GEMMs than without the fusion.)
    float* red = reinterpret_cast<float*>(tmem_slot + 8);  // 128 floats after the barriers
    const int agent = warp - 2;
    const int t64 = threadIdx.x - 64;
    int iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = tile / tiles_mn;
      const int mn = tile - split * tiles_mn;
      const int m_blk = mn / p.num_n_blocks;
      const int n_blk = mn - m_blk * p.num_n_blocks;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
      float cs[16];  // columns (lane & 7) * 8 .. + 8 of box 0 and of box 1; token rows = lane >> 3 (mod 4)
#pragma unroll
      for (int i = 0; i < 16; ++i) cs[i] = 0.f;
      for (int kb = kb0; kb < kb1; ++kb, ++iter) {
        const int st = iter % STAGES;
        if ((st & 1) != agent) continue;
        mbar_wait(&full_bar[st], (iter / STAGES) & 1);
        if (kb % p.num_n_blocks == n_blk && !(p.dbg & 8)) {
          const uint8_t* sa = smem + st * STAGE_BYTES;
#pragma unroll 4
          for (int g = 0; g < 16; ++g) {
            const int k = g * 4 + (lane >> 3);
            const int off = k * 128 + (((lane & 7) ^ (k & 7)) << 4);
#pragma unroll
            for (int bx = 0; bx < 2; ++bx) {
              const uint4 w = *reinterpret_cast<const uint4*>(sa + bx * (64 * BK * 2) + off);
              const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                cs[bx * 8 + 2 * j] += __uint_as_float(ww[j] << 16);
                cs[bx * 8 + 2 * j + 1] += __uint_as_float(ww[j] & 0xFFFF0000u);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[st]);
      }
      red[t64] = 0.f;
      red[t64 + 64] = 0.f;
      asm volatile("bar.sync 1, 64;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) {  // token rows (lane >> 3) -> lanes 0..7
        cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 8);
        cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 16);
      }
      if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 16; ++i) atomicAdd(&red[(i >> 3) * 64 + lane * 8 + (i & 7)], cs[i]);
      }
      asm volatile("bar.sync 1, 64;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const long long gcol = static_cast<long long>(m_blk) * BM + t64 + j * 64;
        if (gcol < p.M) atomicAdd(p.colsum + gcol, red[t64 + j * 64]);
      }
      asm volatile("bar.sync 1, 64;" ::: "memory");
    }
  } else if (warp >= 4) {
    // ------------------------------- epilogue ------------------------------------------------
    const int wq = warp & 3;          // TMEM lane quarter this warp may read (== warp % 4)
    const int half = (warp - 4) >> 2;  // which 128 accumulator columns this warp drains
    int acc = 0;
    uint32_t acc_phase = 0;
    bool pending = false;  // a tensor store of this warp may still be reading its staging tile
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = tile / tiles_mn;
      const int mn = tile - split * tiles_mn;
      const int m_blk = mn / p.num_n_blocks;
      const int n_blk = mn - m_blk * p.num_n_blocks;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const long long row0 = static_cast<long long>(m_blk) * BM + wq * 32;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * BN + half * (BN / 2);
      if (EPI != FV_EPI_ACCUM && EPI != FV_EPI_PATCH && p.direct) {
        epilogue_tile_direct<EPI, BF16>(p, taddr, row0, n_blk * BN + half * (BN / 2), BN / 2, lane);
      } else {
        epilogue_tile<EPI, BF16>(p, &tmap_c, &tmap_aux, epi_stage + (warp - 4) * EPI_STAGE_BYTES, taddr, row0,
                                 n_blk * BN + half * (BN / 2), BN / 2, lane, pending);
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (pending && lane == 0) tma_store_wait_all();  // this thread's tensor stores are complete
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair kernel: 256 x 256 output tile per cluster of two CTAs (tcgen05.mma.cta_group::2)
// ---------------------------------------------------------------------------------------------
// Each CTA loads its 128 rows of A and its 128-column HALF of B (32 KiB per k-block instead of the
// 48 KiB a single-CTA 128 x 256 tile needs: one third less L2 -> SM traffic and shared-memory
// operand bandwidth, and a 6-deep instead of 4-deep ring in the same shared memory), the leader
// CTA's MMA thread issues one M = 256 instruction for the pair, and each CTA drains the 128
// accumulator rows that live in its own tensor memory with the epilogue above.
//   full_bar[s]   leader only: 1 arrival (leader's expect_tx of both CTAs' bytes) + complete_tx from both
//   empty_bar[s]  per CTA: tcgen05.commit multicast to both CTAs
//   tmem_full[a]  per CTA: commit multicast;   tmem_empty[a]  leader only: one arrival per epilogue warp of both CTAs
constexpr int STAGES2 = 6;
constexpr int HALF_BN = BN / 2;
constexpr int B2_STAGE_BYTES = HALF_BN * BK * 2;             // 16 KiB
constexpr int STAGE2_BYTES = A_STAGE_BYTES + B2_STAGE_BYTES;  // 32 KiB
static_assert(STAGES2 * STAGE2_BYTES == STAGES * STAGE_BYTES, "both kernels share the shared-memory budget");

template <int EPI, bool BF16>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_aux,
                const GemmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* epi_stage = smem + STAGES2 * STAGE2_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + EPI_WARPS * EPI_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES2;
  uint64_t* tmem_full = empty_bar + STAGES2;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* a_full = tmem_empty + 2;      // bias-gradient mode, per CTA: this CTA's A tile of the stage has landed
  uint64_t* a_ready = a_full + STAGES2;   // bias-gradient mode, leader only: both A tiles landed and summed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + STAGES2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs), 1 = peer
  // weight-gradient mode with the fused bias gradient (column sums of the A operand, see below)
  const bool colsum_mode = (EPI == FV_EPI_ACCUM) && p.colsum != nullptr;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
    if (EPI == FV_EPI_GELU) tma_prefetch_desc(&tmap_aux);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&a_full[s], 1);
      mbar_init(&a_ready[s], 2);  // one agent warp of each CTA
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int num_m2 = (p.M + 2 * BM - 1) / (2 * BM);
  const int tiles_mn = num_m2 * p.num_n_blocks;
  const int total_tiles = tiles_mn * p.split_k;  // split-K (weight gradient): item = (k slice, m2, n_blk)

  if (warp == 0) {
    // ------------------------------- TMA producer (both CTAs) -------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const int split = tile / tiles_mn;
      const int mn = tile - split * tiles_mn;
      const int m2 = mn / p.num_n_blocks;
      const int n_blk = mn - m2 * p.num_n_blocks;
      const int row_a = m2 * 2 * BM + static_cast<int>(rank) * BM;
      const int row_b = n_blk * BN + static_cast<int>(rank) * HALF_BN;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE2_BYTES;
        uint8_t* sb = sa + A_STAGE_BYTES;
        if (elect_one()) {
          // bias-gradient mode: the A tile reports to THIS CTA's a_full (its agent warp sums it, then
          // tells the leader); everything else reports straight to the leader's full barrier
          uint64_t* bar_a = colsum_mode ? &a_full[stage] : &full_bar[stage];
          if (colsum_mode) {
            mbar_expect_tx(&a_full[stage], A_STAGE_BYTES);
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * B2_STAGE_BYTES);
          } else if (rank == 0) {
            mbar_expect_tx(&full_bar[stage], 2 * STAGE2_BYTES);
          }
          if (p.a_major == FV_MAJOR_K) {
            if (colsum_mode) tma_load_2d(sa, &tmap_a, bar_a, kb * BK, row_a);
            else tma_load_2d_pair(sa, &tmap_a, bar_a, kb * BK, row_a);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) {
              if (colsum_mode) tma_load_2d(sa + j * (64 * BK * 2), &tmap_a, bar_a, row_a + j * 64, kb * BK);
              else tma_load_2d_pair(sa + j * (64 * BK * 2), &tmap_a, bar_a, row_a + j * 64, kb * BK);
            }
          }
          if (p.b_major == FV_MAJOR_K) {
            tma_load_2d_pair(sb, &tmap_b, &full_bar[stage], kb * BK, row_b);
          } else {
#pragma unroll
            for (int j = 0; j < HALF_BN / 64; ++j)
              tma_load_2d_pair(sb + j * (64 * BK * 2), &tmap_b, &full_bar[stage], row_b + j * 64, kb * BK);
          }
        }
        __syncwarp();
        if (++stage == STAGES2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (leader CTA only) ---------------------------
    // The WHOLE warp walks the loop and one elected lane issues: with warp-uniform control flow the
    // descriptors live in uniform registers and each tcgen05.mma is a single predicated
    // instruction. (Issued from inside an `if (lane == 0)` region the compiler wraps every MMA in a
    // vote / broadcast loop — ~25 dependent instructions per MMA, as long as the MMA itself runs.)
    if (rank == 0) {
      const uint32_t idesc = make_idesc(kFmtBF16, p.a_major, p.b_major, 2 * BM, BN);
      const uint32_t a_lbo = p.a_major == FV_MAJOR_K ? 16 : 64 * BK * 2;
      const uint32_t b_lbo = p.b_major == FV_MAJOR_K ? 16 : 64 * BK * 2;
      const uint32_t a_kstep = (p.a_major == FV_MAJOR_K ? 32 : 2048) >> 4;
      const uint32_t b_kstep = (p.b_major == FV_MAJOR_K ? 32 : 2048) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int split = tile / tiles_mn;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (colsum_mode) mbar_wait(&a_ready[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE2_BYTES);
          const uint64_t da = make_smem_desc_sw128(sa, a_lbo, 1024);
          const uint64_t db = make_smem_desc_sw128(sa + A_STAGE_BYTES, b_lbo, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_pair(tmem_d, da + k * a_kstep, db + k * b_kstep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage]);  // frees the slot in both CTAs
          }
          __syncwarp();
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit_pair(&tmem_full[acc]);  // accumulator complete -> both CTAs' epilogues
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if ((warp == 2 || warp == 3) && colsum_mode) {
    // ------------------------------- bias-gradient agents (both CTAs) ------------------------
    // Same job as in the single-CTA kernel (column sums of the dY tiles while they sit in shared
    // memory), different hand-shake: here an A tile reports to the a_full barrier of the CTA it
    // lands in; the stage's agent warp (warp 2: even stages, warp 3: odd) sums its share and then
    // arrives on the LEADER's a_ready barrier, which the MMA thread waits for next to the B bytes.
    // The agents run ahead of the MMA by up to the ring depth, so the sums are normally finished
    // long before the tensor core reaches the stage, and the slot is released by the MMA commit
    // alone (no second arrival on the empty barrier, no agent on the ring's critical path).
    float* red = reinterpret_cast<float*>(tmem_slot + 8);  // 128 floats after the barriers
    const int agent = warp - 2;
    const int t64 = threadIdx.x - 64;
    int iter = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const int split = tile / tiles_mn;
      const int mn = tile - split * tiles_mn;
      const int m2 = mn / p.num_n_blocks;
      const int n_blk = mn - m2 * p.num_n_blocks;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
      float cs[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) cs[i] = 0.f;
      for (int kb = kb0; kb < kb1; ++kb, ++iter) {
        const int st = iter % STAGES2;
        if ((st & 1) != agent) continue;
        mbar_wait(&a_full[st], (iter / STAGES2) & 1);
        // the pairs that work on the same row block (one per n_blk) split its k-slices round-robin
        if (kb % p.num_n_blocks == n_blk && !(p.dbg & 8)) {
          const uint8_t* sa = smem + st * STAGE2_BYTES;
#pragma unroll 4
          for (int g = 0; g < 16; ++g) {
            const int k = g * 4 + (lane >> 3);
            const int off = k * 128 + (((lane & 7) ^ (k & 7)) << 4);
#pragma unroll
            for (int bx = 0; bx < 2; ++bx) {
              const uint4 w = *reinterpret_cast<const uint4*>(sa + bx * (64 * BK * 2) + off);
              const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                cs[bx * 8 + 2 * j] += __uint_as_float(ww[j] << 16);
                cs[bx * 8 + 2 * j + 1] += __uint_as_float(ww[j] & 0xFFFF0000u);
              }
            }
          }
        }
        __syncwarp();
        // plain (default-semantics) remote arrive, as for tmem_empty: the tile was written by the TMA
        // engine and is read by the tensor core, both outside this thread's generic-proxy traffic —
        // a .release.cluster arrive here made every stage pay a cluster-scope fence (fc1: 172 -> 256 us)
        if (lane == 0) mbar_arrive_leader(&a_ready[st]);
      }
      red[t64] = 0.f;
      red[t64 + 64] = 0.f;
      asm volatile("bar.sync 1, 64;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) {  // token rows (lane >> 3) -> lanes 0..7
        cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 8);
        cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 16);
      }
      if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 16; ++i) atomicAdd(&red[(i >> 3) * 64 + lane * 8 + (i & 7)], cs[i]);
      }
      asm volatile("bar.sync 1, 64;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const long long gcol = static_cast<long long>(m2) * 2 * BM + rank * BM + t64 + j * 64;
        if (gcol < p.M) atomicAdd(p.colsum + gcol, red[t64 + j * 64]);
      }
      asm volatile("bar.sync 1, 64;" ::: "memory");
    }
  } else if (warp >= 4) {
    // ------------------------------- epilogue (both CTAs, own 128 rows) ----------------------
    const int wq = warp & 3;
    const int half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    bool pending = false;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const int mn = tile % tiles_mn;
      const int m2 = mn / p.num_n_blocks;
      const int n_blk = mn - m2 * p.num_n_blocks;
#ifdef GEMM_TRACE
      const int gti = (tile - cluster_id) / num_clusters;
      if (GT_ON && gti < 24) {
        g_gemm_trace[gti][0] = gt_clock();
        g_gemm_acq = 0;
        g_gemm_ldw = 0;
        g_gemm_math = 0;
        g_gemm_store = 0;
      }
#endif
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#ifdef GEMM_TRACE
      if (GT_ON && gti < 24) g_gemm_trace[gti][1] = gt_clock();
#endif
      const long long row0 = static_cast<long long>(m2) * 2 * BM + rank * BM + wq * 32;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * BN + half * (BN / 2);
      epilogue_tile<EPI, BF16>(p, &tmap_c, &tmap_aux, epi_stage + (warp - 4) * EPI_STAGE_BYTES, taddr, row0,
                               n_blk * BN + half * (BN / 2), BN / 2, lane, pending);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
#ifdef GEMM_TRACE
      if (GT_ON && gti < 24) {
        g_gemm_trace[gti][2] = gt_clock();
        g_gemm_trace[gti][3] = g_gemm_acq;
        g_gemm_trace[gti][4] = g_gemm_ldw;
        g_gemm_trace[gti][5] = g_gemm_math;
        g_gemm_trace[gti][6] = g_gemm_store;
      }
#endif
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (pending && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees tensor memory) while the pair's MMAs / remote arrivals are in flight
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int make_operand_map(CUtensorMap* map, const void* base, int major, int64_t rows, int64_t k,
                            int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return FV_ERR_CUDA;
  }
  cuuint64_t dims[2];
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2];
  cuuint32_t estr[2] = {1, 1};
  if (major == FV_MAJOR_K) {  // stored [rows, k]
    dims[0] = static_cast<cuuint64_t>(k);
    dims[1] = static_cast<cuuint64_t>(rows);
    box[0] = BK;
    box[1] = static_cast<cuuint32_t>(box_rows);
  } else {  // stored [k, rows]
    dims[0] = static_cast<cuuint64_t>(rows);
    dims[1] = static_cast<cuuint64_t>(k);
    box[0] = 64;
    box[1] = BK;
  }
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p major=%d rows=%lld k=%lld ld=%lld",
              static_cast<int>(r), base, major, static_cast<long long>(rows),
              static_cast<long long>(k), static_cast<long long>(ld));
    return FV_ERR_CUDA;
  }
  return FV_OK;
}

// output tensor map: [rows, cols] row-major, box = 32 rows x 128 bytes, 128B swizzle (the layout
// the epilogue warps stage their segments in)
static int make_out_map(CUtensorMap* map, const void* base, bool bf16, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return FV_ERR_CUDA;
  }
  const int elem = bf16 ? 2 : 4;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * elem};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / elem), 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(output) failed (%d): base=%p rows=%lld cols=%lld ld=%lld", static_cast<int>(r),
              base, static_cast<long long>(rows), static_cast<long long>(cols), static_cast<long long>(ld));
    return FV_ERR_CUDA;
  }
  return FV_OK;
}

template <int EPI, bool BF16>
static int launch_gemm_tc(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tx,
                          const GemmTcParams& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    FV_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<EPI, BF16>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       GEMM_SMEM_BYTES));
    configured = true;
  }
  const int total = p.num_m_blocks * p.num_n_blocks * p.split_k;
  int grid = total < num_sms() ? total : num_sms();
  // FEDVIT_GEMM_GRID (measurement switch, read per call): cap on the CTAs (= SMs) of this kernel
  if (const char* e = getenv("FEDVIT_GEMM_GRID")) {
    const int v = atoi(e);
    if (v > 0 && v < grid) grid = v;
  }
  FV_CHECK_CUDA(fv::launch_pdl(gemm_tc_kernel<EPI, BF16>, dim3(grid), dim3(GEMM_THREADS), GEMM_SMEM_BYTES, stream, ta, tb, tc, tx, p));
  count_kernel(FV_KERNEL_GEMM_TC);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

template <int EPI, bool BF16>
static int launch_gemm_tc2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tx,
                           const GemmTcParams& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    FV_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<EPI, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       GEMM_SMEM_BYTES));
    configured = true;
  }
  const int total = static_cast<int>(ceil_div(p.M, 2 * BM)) * p.num_n_blocks * p.split_k;
  int clusters = num_sms() / 2;
  if (clusters > total) clusters = total;
  FV_CHECK_CUDA(fv::launch_pdl_cluster(gemm_tc2_kernel<EPI, BF16>, dim3(2 * clusters), dim3(GEMM_THREADS),
                                       GEMM_SMEM_BYTES, stream, 2u, ta, tb, tc, tx, p));
  count_kernel(FV_KERNEL_GEMM_TC_PAIR);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

}  // namespace fv

namespace fv {
static int gemm_dbg() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FEDVIT_GEMM_DBG");
    v = e ? atoi(e) : 0;
  }
  return v;
}
static thread_local float* g_wgrad_colsum = nullptr;
static thread_local const float* g_row_scale = nullptr;
static thread_local int g_rows_per_scale = 0;
}

extern "C" int fv_gemm_bf16(const void* a, int a_major, int64_t lda, const void* b, int b_major,
                            int64_t ldb, const float* bias, void* c, int c_dtype, int64_t ldc,
                            void* aux, int64_t ldaux, int64_t m, int64_t n, int64_t k, int epilogue,
                            int split_k, int tokens_per_img, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(a && b && c, "fv_gemm_bf16: null operand");
  FV_CHECK_ARG(m > 0 && n > 0 && k > 0, "fv_gemm_bf16: empty problem m=%lld n=%lld k=%lld",
               (long long)m, (long long)n, (long long)k);
  FV_CHECK_ARG(m < (1LL << 31) && n < (1LL << 31) && k < (1LL << 31), "fv_gemm_bf16: size overflow");
  FV_CHECK_ARG((a_major == FV_MAJOR_K || a_major == FV_MAJOR_MN) &&
                   (b_major == FV_MAJOR_K || b_major == FV_MAJOR_MN),
               "fv_gemm_bf16: bad major");
  FV_CHECK_ARG(n % 8 == 0 && ldc % 8 == 0, "fv_gemm_bf16: n and ldc must be multiples of 8");
  FV_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "fv_gemm_bf16: lda/ldb must be multiples of 8");
  FV_CHECK_ARG((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(c) & 15) == 0,
               "fv_gemm_bf16: operands must be 16-byte aligned");
  FV_CHECK_ARG(c_dtype == FV_F32 || c_dtype == FV_BF16, "fv_gemm_bf16: bad c_dtype");
  FV_CHECK_ARG(epilogue >= FV_EPI_NONE && epilogue <= FV_EPI_PATCH, "fv_gemm_bf16: bad epilogue");
  if (epilogue == FV_EPI_RESIDUAL || epilogue == FV_EPI_ACCUM || epilogue == FV_EPI_PATCH)
    FV_CHECK_ARG(c_dtype == FV_F32, "fv_gemm_bf16: this epilogue writes fp32");
  if (epilogue == FV_EPI_RESIDUAL || epilogue == FV_EPI_DGELU || epilogue == FV_EPI_PATCH ||
      (epilogue == FV_EPI_GELU && aux != nullptr))  // GELU: aux NULL == activation only (forward-only use)
    FV_CHECK_ARG(aux != nullptr && ldaux % 8 == 0 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0,
                 "fv_gemm_bf16: epilogue needs a 16-byte aligned aux with ldaux %% 8 == 0");
  if (epilogue == FV_EPI_PATCH) FV_CHECK_ARG(tokens_per_img > 0, "fv_gemm_bf16: tokens_per_img");
  if (split_k < 1) split_k = 1;
  FV_CHECK_ARG(split_k == 1 || epilogue == FV_EPI_ACCUM, "fv_gemm_bf16: split_k needs FV_EPI_ACCUM");

  GemmTcParams p;
  p.M = static_cast<int>(m);
  p.N = static_cast<int>(n);
  p.K = static_cast<int>(k);
  p.num_m_blocks = static_cast<int>(ceil_div(m, BM));
  p.num_n_blocks = static_cast<int>(ceil_div(n, BN));
  p.num_k_blocks = static_cast<int>(ceil_div(k, BK));
  if (split_k > p.num_k_blocks) split_k = p.num_k_blocks;
  p.kb_per_split = static_cast<int>(ceil_div(p.num_k_blocks, split_k));
  p.split_k = static_cast<int>(ceil_div(p.num_k_blocks, p.kb_per_split));  // no empty slice
  p.a_major = a_major;
  p.b_major = b_major;
  p.c_bf16 = c_dtype == FV_BF16;
  p.tokens_per_img = tokens_per_img;
  p.bias = bias;
  p.c = c;
  p.ldc = ldc;
  p.aux = aux;
  p.ldaux = ldaux;
  p.colsum = g_wgrad_colsum;
  p.row_scale = g_row_scale;
  p.rows_per_scale = g_rows_per_scale;
  FV_CHECK_ARG(p.row_scale == nullptr || (epilogue == FV_EPI_RESIDUAL && p.rows_per_scale > 0),
               "fv_gemm_bf16: row scaling needs the residual epilogue");
  p.im2col = 0;
  p.dbg = gemm_dbg();
  {
    const int64_t ce = p.c_bf16 ? 2 : 4;
    const int64_t ae = (epilogue == FV_EPI_RESIDUAL) ? 4 : ce;
    bool ok = (reinterpret_cast<uintptr_t>(c) & 31) == 0 && (ldc * ce) % 32 == 0;
    if (aux != nullptr) ok = ok && (reinterpret_cast<uintptr_t>(aux) & 31) == 0 && (ldaux * ae) % 32 == 0;
    p.direct = (ok && (p.dbg & 4) && !(epilogue == FV_EPI_GELU && aux == nullptr)) ? 1 : 0;
  }
  p.gw = p.gh = p.chans = p.ph_per_tile = p.tiles_per_img = 0;
  FV_CHECK_ARG(p.colsum == nullptr || (epilogue == FV_EPI_ACCUM && a_major == FV_MAJOR_MN),
               "fv_gemm_bf16: bias-gradient fusion needs the weight-gradient mode");

  CUtensorMap ta, tb;
  int rc = make_operand_map(&ta, a, a_major, m, k, lda, BM);
  if (rc != FV_OK) return rc;
  rc = make_operand_map(&tb, b, b_major, n, k, ldb, BN);
  if (rc != FV_OK) return rc;

  CUtensorMap tc = ta, tx = ta;  // placeholders unless the output path uses them
  p.tma_out = (epilogue != FV_EPI_PATCH && !(p.dbg & 16)) ? 1 : 0;
  if (p.tma_out) {
    rc = make_out_map(&tc, c, p.c_bf16 != 0, m, n, ldc);
    if (rc != FV_OK) return rc;
    if (epilogue == FV_EPI_GELU && aux != nullptr) {
      rc = make_out_map(&tx, aux, p.c_bf16 != 0, m, n, ldaux);
      if (rc != FV_OK) return rc;
    }
  }

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // CTA-pair kernel (256 x 256 tiles): forward / dgrad shapes with enough tiles to fill the pairs
  const int64_t pair_items = ceil_div(m, 2 * BM) * p.num_n_blocks * p.split_k;
  const bool pair_fwd = a_major == FV_MAJOR_K && p.split_k == 1 &&
                        (epilogue == FV_EPI_NONE || epilogue == FV_EPI_RESIDUAL || epilogue == FV_EPI_GELU ||
                         epilogue == FV_EPI_DGELU) &&
                        (pair_items >= num_sms() / 2 || (p.dbg & 64));
  // weight gradient (MN-major x MN-major, split-K accumulate) on the pair kernel as well: a third
  // less operand traffic per flop than the 128 x 256 single-CTA tile; FEDVIT_GEMM_DBG & 256 = never
  const bool pair_wgrad = epilogue == FV_EPI_ACCUM && a_major == FV_MAJOR_MN && b_major == FV_MAJOR_MN &&
                          !(p.dbg & 256) && (pair_items >= num_sms() / 4 || (p.dbg & 64));
  const bool pair = p.tma_out && !(p.dbg & 32) && (pair_fwd || pair_wgrad);
  if (pair) {
    // B map box: this CTA's 128-row half (K-major) — the MN-major map already uses 64-column boxes
    if (b_major == FV_MAJOR_K) {
      rc = make_operand_map(&tb, b, b_major, n, k, ldb, HALF_BN);
      if (rc != FV_OK) return rc;
    }
    switch (epilogue) {
      case FV_EPI_NONE:
        return p.c_bf16 ? launch_gemm_tc2<FV_EPI_NONE, true>(ta, tb, tc, tx, p, st)
                        : launch_gemm_tc2<FV_EPI_NONE, false>(ta, tb, tc, tx, p, st);
      case FV_EPI_RESIDUAL: return launch_gemm_tc2<FV_EPI_RESIDUAL, false>(ta, tb, tc, tx, p, st);
      case FV_EPI_GELU:
        return p.c_bf16 ? launch_gemm_tc2<FV_EPI_GELU, true>(ta, tb, tc, tx, p, st)
                        : launch_gemm_tc2<FV_EPI_GELU, false>(ta, tb, tc, tx, p, st);
      case FV_EPI_DGELU:
        return p.c_bf16 ? launch_gemm_tc2<FV_EPI_DGELU, true>(ta, tb, tc, tx, p, st)
                        : launch_gemm_tc2<FV_EPI_DGELU, false>(ta, tb, tc, tx, p, st);
      case FV_EPI_ACCUM: return launch_gemm_tc2<FV_EPI_ACCUM, false>(ta, tb, tc, tx, p, st);
    }
  }
  switch (epilogue) {
    case FV_EPI_NONE:
      return p.c_bf16 ? launch_gemm_tc<FV_EPI_NONE, true>(ta, tb, tc, tx, p, st) : launch_gemm_tc<FV_EPI_NONE, false>(ta, tb, tc, tx, p, st);
    case FV_EPI_RESIDUAL: return launch_gemm_tc<FV_EPI_RESIDUAL, false>(ta, tb, tc, tx, p, st);
    case FV_EPI_GELU:
      return p.c_bf16 ? launch_gemm_tc<FV_EPI_GELU, true>(ta, tb, tc, tx, p, st) : launch_gemm_tc<FV_EPI_GELU, false>(ta, tb, tc, tx, p, st);
    case FV_EPI_DGELU:
      return p.c_bf16 ? launch_gemm_tc<FV_EPI_DGELU, true>(ta, tb, tc, tx, p, st) : launch_gemm_tc<FV_EPI_DGELU, false>(ta, tb, tc, tx, p, st);
    case FV_EPI_ACCUM: return launch_gemm_tc<FV_EPI_ACCUM, false>(ta, tb, tc, tx, p, st);
    case FV_EPI_PATCH: return launch_gemm_tc<FV_EPI_PATCH, false>(ta, tb, tc, tx, p, st);
  }
  return FV_ERR_INVALID_ARG;
}

extern "C" int fv_wgrad_bf16(const void* dy, int64_t lddy, const void* x, int64_t ldx, float* dw, int64_t lddw,
                             float* dbias, int64_t tokens, int64_t out_features, int64_t in_features,
                             int split_k, void* stream) {
  // dW[out,in] += dY^T X ; dbias[out] += column sums of dY — one kernel (see gemm_tc.cu)
  fv::g_wgrad_colsum = dbias;
  const int rc = fv_gemm_bf16(dy, FV_MAJOR_MN, lddy, x, FV_MAJOR_MN, ldx, nullptr, dw, FV_F32, lddw, nullptr, 0,
                              out_features, in_features, tokens, FV_EPI_ACCUM, split_k, 0, stream);
  fv::g_wgrad_colsum = nullptr;
  return rc;
}

// -------------------------------------------------------------------------------------------------
// im2col-free patch embedding: x[b, 1 + t, :] = patch(b, t) . W^T + bias + pos[1 + t]
// -------------------------------------------------------------------------------------------------
extern "C" int fv_patch_embed_tf32(const float* img, const float* weight, const float* bias, const float* pos,
                                   float* x, int64_t batch, int64_t chans, int64_t height, int64_t width,
                                   int64_t dim, void* stream) {
  using namespace fv;
  FV_CHECK_ARG(img && weight && pos && x, "fv_patch_embed_tf32: null pointer");
  FV_CHECK_ARG(batch > 0 && chans > 0 && height % 16 == 0 && width % 16 == 0 && height > 0 && width > 0,
               "fv_patch_embed_tf32: H and W must be positive multiples of 16");
  const int gw = static_cast<int>(width / 16), gh = static_cast<int>(height / 16);
  FV_CHECK_ARG(gw <= 128 && dim % 8 == 0 && dim > 0, "fv_patch_embed_tf32: width <= 2048 px, dim %% 8 == 0");
  FV_CHECK_ARG((reinterpret_cast<uintptr_t>(img) & 15) == 0 && (reinterpret_cast<uintptr_t>(weight) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(pos) & 15) == 0,
               "fv_patch_embed_tf32: pointers must be 16-byte aligned");
  EncodeTiledFn enc = get_encode_tiled();
  FV_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  GemmTcParams p;
  p.im2col = 1;
  p.gw = gw;
  p.gh = gh;
  p.chans = static_cast<int>(chans);
  p.ph_per_tile = BM / gw;
  if (p.ph_per_tile > gh) p.ph_per_tile = gh;
  p.tiles_per_img = (gh + p.ph_per_tile - 1) / p.ph_per_tile;
  p.M = static_cast<int>(batch * gw * gh);  // logical rows (patches)
  p.N = static_cast<int>(dim);
  p.K = static_cast<int>(chans * 256);
  p.num_m_blocks = static_cast<int>(batch) * p.tiles_per_img;
  p.num_n_blocks = static_cast<int>(ceil_div(dim, BN));
  p.num_k_blocks = static_cast<int>(chans) * 8;  // pairs of pixel rows
  p.split_k = 1;
  p.kb_per_split = p.num_k_blocks;
  p.a_major = FV_MAJOR_K;
  p.b_major = FV_MAJOR_K;
  p.c_bf16 = 0;
  p.tokens_per_img = gw * gh;
  p.bias = bias;
  p.c = x;
  p.ldc = dim;
  p.aux = const_cast<float*>(pos);
  p.ldaux = dim;
  p.colsum = nullptr;
  p.row_scale = nullptr;
  p.rows_per_scale = 0;
  p.dbg = 0;
  p.direct = 0;
  p.tma_out = 0;

  CUtensorMap ta, tb;
  {  // image viewed as (px, py, pw, ph, b*c); strides in bytes for dims 1..4
    cuuint64_t dims[5] = {16, 16, static_cast<cuuint64_t>(gw), static_cast<cuuint64_t>(gh),
                          static_cast<cuuint64_t>(batch * chans)};
    cuuint64_t strides[4] = {static_cast<cuuint64_t>(width) * 4, 64, static_cast<cuuint64_t>(width) * 64,
                             static_cast<cuuint64_t>(width) * height * 4};
    cuuint32_t box[5] = {16, 1, static_cast<cuuint32_t>(gw), static_cast<cuuint32_t>(p.ph_per_tile), 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(img), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(image 5-D) failed (%d)", static_cast<int>(r));
      return FV_ERR_CUDA;
    }
  }
  {  // conv weight [dim, chans*256] fp32, K-major; 32 fp32 (two pixel rows) = one 128-byte swizzle row
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(chans * 256), static_cast<cuuint64_t>(dim)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(chans) * 256 * 4};
    cuuint32_t box[2] = {32, BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(weight), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(patch weight) failed (%d)", static_cast<int>(r));
      return FV_ERR_CUDA;
    }
  }
  return launch_gemm_tc<FV_EPI_PATCH, false>(ta, tb, ta, ta, p, static_cast<cudaStream_t>(stream));
}

extern "C" int fv_linear_residual_bf16(const void* a, int64_t lda, const void* w, int64_t ldw, const float* bias,
                                       const float* residual, int64_t ldr, const float* row_scale,
                                       int64_t rows_per_scale, float* out, int64_t ldo, int64_t m, int64_t n,
                                       int64_t k, void* stream) {
  // out = residual + row_scale[row / rows_per_scale] * (a w^T + bias): the branch-plus-residual
  // step of a transformer block with per-sample stochastic depth (row_scale NULL == no drop-path)
  fv::g_row_scale = row_scale;
  fv::g_rows_per_scale = static_cast<int>(rows_per_scale);
  const int rc = fv_gemm_bf16(a, FV_MAJOR_K, lda, w, FV_MAJOR_K, ldw, bias, out, FV_F32, ldo,
                              const_cast<float*>(residual), ldr, m, n, k, FV_EPI_RESIDUAL, 1, 0, stream);
  fv::g_row_scale = nullptr;
  fv::g_rows_per_scale = 0;
  return rc;
}

#ifdef GEMM_TRACE
extern "C" int fv_debug_read_gemm_trace(long long* dst) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(dst, fv::g_gemm_trace, sizeof(fv::g_gemm_trace)) == cudaSuccess ? 0 : -2;
}
#endif

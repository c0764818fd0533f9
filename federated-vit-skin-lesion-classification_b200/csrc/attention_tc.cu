// attention_tc.cu — attention core on 5th-gen tensor cores for the short ViT sequences
// (N <= 256 tokens: 197 at 224 px), head_dim 64. Forward here; see the bottom for dispatch.
//
// Replaces F.scaled_dot_product_attention inside timm's Attention.forward (reference call site
// model.py:193). One CTA per (batch, head): K/V are fetched once and the (one or two) 128-query
// tiles run back to back. The whole key range fits one accumulator tile, so there is no
// online-softmax rescaling at all:
//   TMA      Q [128 x 64], K [kw x 64], V [kw x 64] straight out of timm's [B,N,3,H,64] qkv layout
//            (3-D tensor map, rows past N zero-filled), 128B-swizzled smem
//   MMA 1    S = Q K^T           tcgen05.mma 128 x kw x 64  -> TMEM columns [0, kw)    (fp32)
//   softmax  thread == query row (tcgen05.ld 32x32b): row max, exp2, row sum in registers;
//            P (bf16 pairs) is written back with tcgen05.st into TMEM columns [0, kw/2) — over the
//            part of S this thread has already consumed — and never touches shared memory
//   MMA 2    O = P V             tcgen05.mma, A operand from TMEM, V as an MN-major smem operand
//                                -> TMEM columns [128, 192)
//   epilogue O / rowsum -> bf16 -> swizzled smem transpose -> coalesced stores; LSE saved.
// 256 TMEM columns and 80 KB smem per CTA: two CTAs per SM, so one CTA's softmax (MUFU-bound)
// overlaps the other's MMAs.
#include "common.cuh"

namespace fv {

constexpr int ATC_THREADS = 192;  // warps 0-3 softmax/epilogue, warp 4 MMA issue, warp 5 TMEM alloc + TMA producer
constexpr int ATC_Q = 128;
constexpr int ATC_KV_MAX = 256;
constexpr int ATC_SMEM = 2 * ATC_Q * 128 + 2 * ATC_KV_MAX * 128 + 1024 + 128;
constexpr float ATC_LOG2E = 1.4426950408889634f;

struct AttnTcParams {
  int N, H, kw;  // tokens, heads, keys rounded up to 16
  int items;     // batch * heads
  int kv_box;    // rows of the K/V TMA box
  float scale;
  __nv_bfloat16* out;
  float* lse;
};

#ifdef ATC_TRACE
// measurement build (tools/build_variants.py ... trace:-DATC_TRACE): SM-clock timestamps of the first tiles of
// four CTAs — [cta 0..3][tile 0..31][event 0..7]; events 0-4 softmax warp 0 (S ready, pass 1 done, P written,
// O ready, row stored), 5-6 MMA warp (S issued, PV issued), 7 = smid. Read back with fv_debug_read_trace.
__device__ long long g_atc_trace[4][32][8];
__device__ __forceinline__ void atc_stamp(int gt, int ev) {
  const int c = blockIdx.x == 0 ? 0 : blockIdx.x == 1 ? 1 : blockIdx.x == 148 ? 2 : blockIdx.x == 149 ? 3 : -1;
  if (c >= 0 && gt < 32) {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    g_atc_trace[c][gt][ev] = t;
    if (ev == 0) {
      uint32_t sm;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      g_atc_trace[c][gt][7] = sm;
    }
  }
}
#define ATC_STAMP(gt, ev) do { if (lane == 0) atc_stamp(gt, ev); } while (0)
#else
#define ATC_STAMP(gt, ev) do { } while (0)
#endif

__device__ __forceinline__ float atc_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATC_THREADS, 2)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                   const AttnTcParams p) {
  // Persistent: two co-resident CTAs per SM walk (batch, head) items; barriers, tensor memory and
  // descriptors are set up once. A producer warp refills Q / K as soon as the item's last score
  // MMA has retired and V after its last PV MMA, so the next item's operands arrive while this
  // item's softmax runs; O leaves straight from registers (256-bit stores), so the Q tiles are not
  // needed as staging.
  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic on the __shared__ array (not an integer round trip) keeps the address space
  // known to the compiler: LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                         // 2 x 16 KiB (both query tiles)
  uint8_t* sK = sQ + 2 * ATC_Q * 128;         // 32 KiB
  uint8_t* sV = sK + ATC_KV_MAX * 128;        // 32 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATC_KV_MAX * 128);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint64_t* bar_done = bars + 5;
  uint64_t* bar_qkfree = bars + 6;
  uint64_t* bar_vfree = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = (p.N + ATC_Q - 1) / ATC_Q;  // 1 or 2 query tiles, processed back to back
  const int hd = p.H * 64;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    mbar_init(bar_done, 128);
    mbar_init(bar_qkfree, 1);
    mbar_init(bar_vfree, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();  // nothing above touches memory another kernel produced

  if (warp == 5) {
    // ------------------------------ TMA producer (whole warp, elected lane issues) --------------
    int n = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      const int b = item / p.H, h = item % p.H;
      if (n > 0) mbar_wait(bar_qkfree, (n - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(bar_qk, (nqt * ATC_Q + p.kv_box) * 128);
        for (int t = 0; t < nqt; ++t) tma_load_3d(sQ + t * ATC_Q * 128, &tmap_q, bar_qk, h * 64, t * ATC_Q, b);
        tma_load_3d(sK, &tmap_kv, bar_qk, hd + h * 64, 0, b);
      }
      __syncwarp();
      if (n > 0) mbar_wait(bar_vfree, (n - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(bar_v, p.kv_box * 128);
        tma_load_3d(sV, &tmap_kv, bar_v, 2 * hd + h * 64, 0, b);
      }
      __syncwarp();
    }
  } else if (warp == 4) {
    // ------------------------------ MMA issue (whole warp, elected lane issues) -----------------
    const uint32_t idesc_s = make_idesc(kFmtBF16, 0, 0, ATC_Q, p.kw);
    const uint32_t idesc_o = make_idesc(kFmtBF16, 0, 1, ATC_Q, 64);
    const uint64_t dk = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
    const uint64_t dv = make_smem_desc_sw128(smem_u32(sV), 64 * 128, 1024);
    const int ksteps = p.kw >> 4;
    int n = 0, gt = 0;  // items, query tiles so far
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      mbar_wait(bar_qk, n & 1);
      for (int t = 0; t < nqt; ++t, ++gt) {
        if (gt > 0) mbar_wait(bar_done, (gt - 1) & 1);  // previous tile's O has been read out of TMEM
        tc_fence_after();
        // S = Q K^T : both operands K-major (head dim contiguous), 4 steps of K = 16
        const uint64_t dq = make_smem_desc_sw128(smem_u32(sQ + t * ATC_Q * 128), 16, 1024);
        ATC_STAMP(gt, 5);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem, dq + k * 2, dk + k * 2, idesc_s, k > 0 ? 1u : 0u);
          umma_commit(bar_s);
          if (t == nqt - 1) umma_commit(bar_qkfree);  // Q tiles and K may be refilled when these retire
        }
        __syncwarp();
        mbar_wait(bar_p, gt & 1);
        if (t == 0) mbar_wait(bar_v, n & 1);
        tc_fence_after();
        ATC_STAMP(gt, 6);
        // O = P V : A = P from TMEM (16 keys = 8 packed columns per step), B = V MN-major
        if (elect_one()) {
          for (int k = 0; k < ksteps; ++k)
            umma_bf16_ts(tmem + 128, tmem + k * 8, dv + k * (2048 >> 4), idesc_o, k > 0 ? 1u : 0u);
          umma_commit(bar_o);
          if (t == nqt - 1) umma_commit(bar_vfree);
        }
        __syncwarp();
      }
    }
  } else if (warp < 4) {
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const float sl2 = p.scale * ATC_LOG2E;
    const int nchunks = (p.kw + 31) >> 5;
    int gt = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int b = item / p.H, h = item % p.H;
      for (int t = 0; t < nqt; ++t, ++gt) {
        const uint32_t ph = gt & 1;
        const int q0 = t * ATC_Q;
        const int q = q0 + warp * 32 + lane;
        const bool warp_live = q0 + warp * 32 < p.N;  // warp-uniform
        float mx = -INFINITY, sum = 0.f;
        mbar_wait(bar_s, ph);
        tc_fence_after();
        if (warp == 0) ATC_STAMP(gt, 0);
        if (warp_live) {
          for (int c = 0; c < nchunks; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + c * 32, r);
            tmem_ld_wait();
            if ((c + 1) * 32 <= p.N) {  // whole chunk inside the sequence: no per-element masking
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c * 32 + i < p.N) mx = fmaxf(mx, __uint_as_float(r[i]));
            }
          }
          const float mxs = mx * sl2;
          if (warp == 0) ATC_STAMP(gt, 1);
          for (int c = 0; c < nchunks; ++c) {
            uint32_t r[32], pk[16];
            tmem_ld_32x32(taddr + c * 32, r);
            tmem_ld_wait();
            if ((c + 1) * 32 <= p.N) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float p0 = atc_ex2(fmaf(__uint_as_float(r[i]), sl2, -mxs));
                const float p1 = atc_ex2(fmaf(__uint_as_float(r[i + 1]), sl2, -mxs));
                sum += p0 + p1;  // fp32 row sum (the saved LSE is the exact log-sum-exp)
                pk[i >> 1] = pack_bf16(p0, p1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float p0 = (c * 32 + i < p.N) ? atc_ex2(fmaf(__uint_as_float(r[i]), sl2, -mxs)) : 0.f;
                const float p1 = (c * 32 + i + 1 < p.N) ? atc_ex2(fmaf(__uint_as_float(r[i + 1]), sl2, -mxs)) : 0.f;
                sum += p0 + p1;
                pk[i >> 1] = pack_bf16(p0, p1);
              }
            }
            tmem_st_32x16(taddr + c * 16, pk);  // columns this thread has already consumed
          }
          tmem_st_wait();
        }
        tc_fence_before();
        mbar_arrive(bar_p);
        if (warp == 0) ATC_STAMP(gt, 2);

        mbar_wait(bar_o, ph);
        tc_fence_after();
        if (warp == 0) ATC_STAMP(gt, 3);
        uint32_t o0[32], o1[32];
        if (warp_live) {
          tmem_ld_32x32(taddr + 128, o0);
          tmem_ld_32x32(taddr + 160, o1);
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive(bar_done);  // TMEM may be overwritten by the next tile's S
        if (warp_live && q < p.N) {
          // this thread's row: 64 bf16 = one 128-byte line of out[b, q, h, :], four 256-bit stores
          const float inv = 1.0f / sum;
          uint32_t w[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            w[i] = pack_bf16(__uint_as_float(o0[2 * i]) * inv, __uint_as_float(o0[2 * i + 1]) * inv);
            w[16 + i] = pack_bf16(__uint_as_float(o1[2 * i]) * inv, __uint_as_float(o1[2 * i + 1]) * inv);
          }
          __nv_bfloat16* dst = p.out + ((static_cast<long long>(b) * p.N + q) * p.H + h) * 64;
#pragma unroll
          for (int j = 0; j < 4; ++j) st_v8(dst + j * 16, w + j * 8);
          p.lse[(static_cast<long long>(b) * p.H + h) * p.N + q] = mx * p.scale + logf(sum);
        }
        if (warp == 0) ATC_STAMP(gt, 4);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// 3-D map over qkv [B, N, 3*H*64] bf16: box = 64 columns x `rows` tokens x 1 image
int make_qkv_map(CUtensorMap* map, const void* base, int64_t batch, int64_t tokens, int64_t width,
                 int rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return FV_ERR_CUDA;
  }
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(width), static_cast<cuuint64_t>(tokens),
                        static_cast<cuuint64_t>(batch)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(width) * 2, static_cast<cuuint64_t>(width) * tokens * 2};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(qkv) failed (%d)", static_cast<int>(r));
    return FV_ERR_CUDA;
  }
  return FV_OK;
}

int attention_tc_fwd(const void* qkv, void* out, float* lse, int64_t batch, int64_t tokens, int64_t heads,
                     float scale, cudaStream_t stream) {
  AttnTcParams p;
  p.N = static_cast<int>(tokens);
  p.H = static_cast<int>(heads);
  p.kw = static_cast<int>((tokens + 15) / 16 * 16);
  p.kv_box = p.kw;
  p.items = static_cast<int>(batch * heads);
  p.scale = scale;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  CUtensorMap mq, mkv;
  int rc = make_qkv_map(&mq, qkv, batch, tokens, 3 * heads * 64, ATC_Q);
  if (rc != FV_OK) return rc;
  rc = make_qkv_map(&mkv, qkv, batch, tokens, 3 * heads * 64, p.kv_box);
  if (rc != FV_OK) return rc;
  static bool configured = false;
  if (!configured) {
    FV_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM));
    configured = true;
  }
  const int slots = 2 * num_sms();  // two co-resident CTAs per SM
  const unsigned grid = static_cast<unsigned>(p.items < slots ? p.items : slots);
  FV_CHECK_CUDA(fv::launch_pdl(attn_tc_fwd_kernel, dim3(grid), dim3(ATC_THREADS), ATC_SMEM, stream, mq, mkv, p));
  count_kernel(FV_KERNEL_ATTN_FWD);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

// ---------------------------------------------------------------------------------------------
// backward (N <= 256): one CTA per (batch, head); every product on tcgen05, accumulators in TMEM
// ---------------------------------------------------------------------------------------------
//   per 128-key block kb, per 128-query tile qt:
//     MMA 1   S  = Q_qt K_kb^T  -> TMEM [0,128)        dP = dO_qt V_kb^T -> TMEM [128,256)
//     math    (8 warps; thread == query row, half of the block's keys)
//             P = exp2(S*scale*log2e - LSE*log2e),  dS = P * (dP - delta) * scale
//             -> bf16, 128B-swizzled smem tiles sP / sdS  [128 queries x 128 keys]
//     MMA 2   dV_kb += P^T dO_qt   (A = sP read MN-major)      -> TMEM [320,384)
//             dK_kb += dS^T Q_qt   (A = sdS read MN-major)     -> TMEM [256,320)
//             dQ_qt += dS K_kb     (A = sdS read K-major)      -> TMEM [384,448) / [448,512)
//   dK/dV leave TMEM after the last query tile of a key block, dQ after the last key block.
//   delta = rowsum(dO * O) is computed once per (batch, head) from the O / dO smem tiles (no
//   separate pass over HBM).
// Q / dO tiles double as A operands (K-major) and B operands (MN-major); K / V likewise — every
// tile is loaded once per (batch, head) and nothing is transposed or re-materialised.
constexpr int ATB_THREADS = 384;  // warps 0-7 math, 8 MMA issue, 9 TMEM alloc + TMA producer, 10-11 delta / LSE helpers
constexpr int ATB_TILE = 128 * 128;  // bytes of one [128 x 64] bf16 tile
constexpr int ATB_AUX = 4;  // items of lse / delta the helper warps may run ahead
constexpr int ATB_STAGE = 16 * 128;  // per math warp: 16 output rows x 128 bytes on their way to a TMA store
constexpr int ATB_SMEM = 12 * ATB_TILE + 8 * ATB_STAGE + 1024 + ATB_AUX * 2048 + 256;  // Q(2) dO(2) K(2) V(2) P(2) dS(2) + store staging + {lse, delta}[ATB_AUX][256] + barriers

struct AttnBwdParams {
  int N, H, kw;
  int items;  // batch * heads
  float scale;
  const float* lse;
  const __nv_bfloat16* o;     // forward output  [B, N, H*64]
  const __nv_bfloat16* dout;  // its gradient    [B, N, H*64]
  __nv_bfloat16* dqkv;
};

#ifdef ATC_TRACE
// measurement build: SM-clock timestamps of math warp 0 of CTA 0 over its first 24 (key block, query tile) cells —
// 0 arrives at the score wait, 1 scores there, 2 S / dP of both chunks read (bar_c), 3 math done, 4 previous MMA 2
// retired (sP / sdS free), 5 P / dS stored (bar_p), 6 drains done
__device__ long long g_atb_trace[24][8];
#define ATB_STAMP(it, ev)                                              \
  do {                                                                 \
    if (blockIdx.x == 0 && threadIdx.x == 0 && (it) < 24) {            \
      long long t_;                                                    \
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_));               \
      g_atb_trace[it][ev] = t_;                                        \
    }                                                                  \
  } while (0)
#else
#define ATB_STAMP(it, ev) do { } while (0)
#endif

__global__ void __launch_bounds__(ATB_THREADS, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                   const __grid_constant__ CUtensorMap tmap_dqkv, const AttnBwdParams p) {
  // Persistent: one CTA per SM walks (batch, head) items; barriers, tensor memory and descriptors
  // are set up once, and the operand tiles of item i+1 are fetched (by a dedicated producer thread)
  // as soon as the last MMA that reads the corresponding tile of item i has retired — K0/V0 after
  // the first key block, Q0/dO0 one iteration before the end — so the next item's first MMA never
  // waits for a cold TMA round trip. (ncu on the one-CTA-per-item version: tensor pipe 16 % active,
  // every CTA paying launch + allocation + load latency with nothing to overlap them: 192 KB of
  // shared memory and all 512 TMEM columns allow only one CTA per SM.)
  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic on the __shared__ array (not an integer round trip) keeps the address space
  // known to the compiler: LDS / STS instead of generic LD / ST for every staging access
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                  // 2 tiles
  uint8_t* sdO = sQ + 2 * ATB_TILE;    // 2 tiles
  uint8_t* sK = sdO + 2 * ATB_TILE;    // 2 tiles (both key blocks)
  uint8_t* sV = sK + 2 * ATB_TILE;     // 2 tiles
  uint8_t* sP = sV + 2 * ATB_TILE;     // 2 column blocks of 64 keys
  uint8_t* sdS = sP + 2 * ATB_TILE;    // 2 column blocks
  uint8_t* sStage = sdS + 2 * ATB_TILE;  // 8 x 2 KiB, 1024-byte aligned (TMA 128B-swizzle atoms)
  float* sAux = reinterpret_cast<float*>(sStage + 8 * ATB_STAGE);  // [ATB_AUX items][lse*log2e[256], delta[256]]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sAux + ATB_AUX * 512);
  uint64_t* bar_ld = bars + 0;     // [3] loaded: {Q0 dO0 K0 V0}, {Q1 dO1}, {K1 V1}
  uint64_t* bar_free = bars + 3;   // [4] last reader retired: {K0 V0}, {Q0 dO0}, {K1 V1}, {Q1 dO1}
  uint64_t* bar_s = bars + 7;
  uint64_t* bar_p = bars + 8;
  uint64_t* bar_m2 = bars + 9;
  uint64_t* bar_kvfree = bars + 10;
  uint64_t* bar_dqfree = bars + 11;
  uint64_t* bar_aux = bars + 12;                // [ATB_AUX] lse / delta of item n ready in sAux[n % ATB_AUX]
  uint64_t* bar_auxfree = bar_aux + ATB_AUX;    // [ATB_AUX] ... and read for the last time
  uint64_t* bar_c = bar_auxfree + ATB_AUX;      // S / dP are in the math warps' registers: TMEM may be rewritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_c + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = (p.N + 127) >> 7;  // query tiles == key blocks (1 or 2)
  const int hd = p.H * 64;
  const int nitems = p.items;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_dqkv);
    for (int i = 0; i < 3; ++i) mbar_init(&bar_ld[i], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_free[i], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 256);
    mbar_init(bar_m2, 1);
    mbar_init(bar_kvfree, 256);
    mbar_init(bar_dqfree, 256);
    mbar_init(bar_c, 256);
    for (int i = 0; i < ATB_AUX; ++i) {
      mbar_init(&bar_aux[i], 64);
      mbar_init(&bar_auxfree[i], 256);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();  // nothing above touches memory another kernel produced
  constexpr uint32_t T_S = 0, T_DP = 128, T_DK = 256, T_DV = 320, T_DQ = 384;

  if (warp == 9) {
    // ------------------------------ TMA producer (whole warp, elected lane issues) ----------------
    int n = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++n) {
      const int b = item / p.H, h = item % p.H;
      const uint32_t prev = (n - 1) & 1;
      // a barrier is re-armed only after its previous phase is known complete: the tiles' last
      // readers of item n-1 have retired, so their loads (that phase) finished long ago
      if (n > 0) mbar_wait(&bar_free[0], prev);
      if (elect_one()) {
        mbar_expect_tx(&bar_ld[0], 4 * ATB_TILE);
        tma_load_3d(sK, &tmap_qkv, &bar_ld[0], hd + h * 64, 0, b);
        tma_load_3d(sV, &tmap_qkv, &bar_ld[0], 2 * hd + h * 64, 0, b);
      }
      __syncwarp();
      if (n > 0) mbar_wait(&bar_free[1], prev);
      if (elect_one()) {
        tma_load_3d(sQ, &tmap_qkv, &bar_ld[0], h * 64, 0, b);
        tma_load_3d(sdO, &tmap_do, &bar_ld[0], h * 64, 0, b);
      }
      __syncwarp();
      if (nt > 1) {
        if (n > 0) mbar_wait(&bar_free[3], prev);
        if (elect_one()) {
          mbar_expect_tx(&bar_ld[1], 2 * ATB_TILE);
          tma_load_3d(sQ + ATB_TILE, &tmap_qkv, &bar_ld[1], h * 64, 128, b);
          tma_load_3d(sdO + ATB_TILE, &tmap_do, &bar_ld[1], h * 64, 128, b);
        }
        __syncwarp();
        if (n > 0) mbar_wait(&bar_free[2], prev);
        if (elect_one()) {
          mbar_expect_tx(&bar_ld[2], 2 * ATB_TILE);
          tma_load_3d(sK + ATB_TILE, &tmap_qkv, &bar_ld[2], hd + h * 64, 128, b);
          tma_load_3d(sV + ATB_TILE, &tmap_qkv, &bar_ld[2], 2 * hd + h * 64, 128, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 8) {
    // ------------------------------ MMA issue (whole warp, elected lane issues) -----------------
    // Warp-uniform control flow keeps every descriptor in uniform registers and each tcgen05.mma a
    // single predicated instruction. Issued from inside an `if (lane == 0)` region the compiler
    // wraps every MMA in a vote + broadcast loop of ~25 dependent instructions — longer than these
    // N = 64 MMAs take to execute, so the issuing thread, not the tensor pipe, set the pace.
    const uint32_t id_dvk = make_idesc(kFmtBF16, 1, 1, 128, 64);  // A MN-major (P^T / dS^T), B MN-major
    const uint32_t id_dq = make_idesc(kFmtBF16, 0, 1, 128, 64);   // A K-major (dS), B MN-major (K)
    // MMA 1 of (key block, query tile): scores and dP; contraction over the 64 head dims
    auto issue_mma1 = [&](int kb, int qt) {
      int kwb = p.kw - kb * 128;
      if (kwb > 128) kwb = 128;
      const uint32_t id_s = make_idesc(kFmtBF16, 0, 0, 128, kwb);
      const uint64_t dK_k = make_smem_desc_sw128(smem_u32(sK + kb * ATB_TILE), 16, 1024);
      const uint64_t dV_k = make_smem_desc_sw128(smem_u32(sV + kb * ATB_TILE), 16, 1024);
      const uint64_t dQ_k = make_smem_desc_sw128(smem_u32(sQ + qt * ATB_TILE), 16, 1024);
      const uint64_t dO_k = make_smem_desc_sw128(smem_u32(sdO + qt * ATB_TILE), 16, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + T_S, dQ_k + k * 2, dK_k + k * 2, id_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + T_DP, dO_k + k * 2, dV_k + k * 2, id_s, k > 0 ? 1u : 0u);
        umma_commit(bar_s);
      }
      __syncwarp();
    };
    int it = 0;   // (key block, query tile) iterations so far, over all items
    int kvn = 0;  // key blocks so far
    int n = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++n) {
      const uint32_t ph = n & 1;
      if (n == 0) {
        mbar_wait(&bar_ld[0], 0);
        tc_fence_after();
        issue_mma1(0, 0);
      }
      const bool has_next = item + static_cast<int>(gridDim.x) < nitems;
      for (int kb = 0; kb < nt; ++kb, ++kvn) {
        int kwb = p.kw - kb * 128;
        if (kwb > 128) kwb = 128;
        const uint64_t dK_mn = make_smem_desc_sw128(smem_u32(sK + kb * ATB_TILE), ATB_TILE, 1024);  // MN-major view
        for (int qt = 0; qt < nt; ++qt, ++it) {
          const uint64_t dQ_mn = make_smem_desc_sw128(smem_u32(sQ + qt * ATB_TILE), ATB_TILE, 1024);
          const uint64_t dO_mn = make_smem_desc_sw128(smem_u32(sdO + qt * ATB_TILE), ATB_TILE, 1024);
          // the next (key block, query tile)'s scores are issued as soon as this one's S / dP have
          // been read into registers (bar_c) — half the softmax math and the P / dS stores earlier
          // than "P / dS written" (bar_p); ncu had the math warps waiting on bar_s for a third of
          // their time
          mbar_wait(bar_c, it & 1);
          tc_fence_after();
          if (qt + 1 < nt) {
            if (kb == 0) mbar_wait(&bar_ld[1], ph);
            tc_fence_after();
            issue_mma1(kb, qt + 1);
          } else if (kb + 1 < nt) {
            mbar_wait(&bar_ld[2], ph);
            tc_fence_after();
            issue_mma1(kb + 1, 0);
          }
          mbar_wait(bar_p, it & 1);  // P / dS tiles written
          tc_fence_after();
          if (qt == 0 && kvn > 0) mbar_wait(bar_kvfree, (kvn - 1) & 1);  // previous dK / dV have left TMEM
          if (kb == 0 && qt == 0 && n > 0) mbar_wait(bar_dqfree, (n - 1) & 1);  // previous item's dQ likewise
          tc_fence_after();
          // MMA 2: contraction over the 128 queries (dV, dK) and over the block's keys (dQ)
          const uint64_t dP_mn = make_smem_desc_sw128(smem_u32(sP), ATB_TILE, 1024);
          const uint64_t dS_mn = make_smem_desc_sw128(smem_u32(sdS), ATB_TILE, 1024);
          const uint64_t dS_k0 = make_smem_desc_sw128(smem_u32(sdS), 16, 1024);
          const int ks = kwb >> 4;
          // dV / dK contract over the tile's queries: the ragged last tile (69 of 128 at N = 197) has P = dS = 0
          // beyond the sequence, so only its first ceil(rows / 16) k-steps are issued (5 of 8: 36 KB less
          // shared-memory operand traffic per such cell, and this kernel is bound by that traffic)
          int qrows = p.N - qt * 128;
          if (qrows > 128) qrows = 128;
          const int qs = (qrows + 15) >> 4;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (k < qs) umma_bf16(tmem + T_DV, dP_mn + k * 128, dO_mn + k * 128, id_dvk, (qt > 0 || k > 0) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (k < qs) umma_bf16(tmem + T_DK, dS_mn + k * 128, dQ_mn + k * 128, id_dvk, (qt > 0 || k > 0) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (k < ks) {
                // dS tile k: column block k >> 2 (one 16 KB tile apart), 32 bytes per K = 16 step inside it
                const uint64_t dS_k = dS_k0 + (((k >> 2) * ATB_TILE + (k & 3) * 32) >> 4);
                umma_bf16(tmem + T_DQ + qt * 64, dS_k, dK_mn + k * 128, id_dq, (kb > 0 || k > 0) ? 1u : 0u);
              }
            }
            umma_commit(bar_m2);
            // operand tiles whose last reader was just issued: free them for the next item's loads
            if (kb == 0 && qt == nt - 1) umma_commit(&bar_free[0]);
            if (kb == nt - 1 && qt == 0) umma_commit(&bar_free[1]);
            if (nt > 1 && kb == 1 && qt == nt - 1) umma_commit(&bar_free[2]);
            if (nt > 1 && kb == nt - 1 && qt == 1) umma_commit(&bar_free[3]);
          }
          __syncwarp();
          if (kb == nt - 1 && qt == nt - 1 && has_next) {
            // first scores of the next item (its Q0 / dO0 / K0 / V0 were prefetched)
            mbar_wait(&bar_ld[0], ph ^ 1);
            tc_fence_after();
            issue_mma1(0, 0);
          }
        }
      }
    }
  } else if (warp >= 10) {
    // ------------------------------ delta / LSE helpers ------------------------------------------
    // delta[q] = sum_d dO[q,d] * O[q,d] and lse[q] * log2(e) of the NEXT items, one row per thread
    // pass straight from global memory (O never occupies shared memory), one item ahead of the math
    // warps (up to ATB_AUX - 1 items) — those found the three dependent global round trips at every
    // item start on their critical path (ncu: 22 % of their samples).
    const int t = threadIdx.x - 320;
    int n = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++n) {
      const int b = item / p.H, h = item % p.H;
      const int slot = n % ATB_AUX;
      if (n >= ATB_AUX) mbar_wait(&bar_auxfree[slot], (n / ATB_AUX - 1) & 1);  // previous tenant read out
      float* aux = sAux + slot * 512;
      const float* lse_bh = p.lse + (static_cast<long long>(b) * p.H + h) * p.N;
      // two rows per pass: 32 independent 128-bit loads in flight per thread
      for (int q0 = t; q0 < nt * 128; q0 += 128) {
        uint4 a[2][8], g[2][8];
        float l2[2] = {INFINITY, INFINITY};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int q = q0 + j * 64;
          if (q < p.N) {
            const long long off = (static_cast<long long>(b) * p.N + q) * hd + h * 64;
            const uint4* orow = reinterpret_cast<const uint4*>(p.o + off);
            const uint4* grow = reinterpret_cast<const uint4*>(p.dout + off);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              a[j][u] = __ldg(orow + u);
              g[j][u] = __ldg(grow + u);
            }
            l2[j] = __ldg(lse_bh + q) * ATC_LOG2E;
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) a[j][u] = g[j][u] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int q = q0 + j * 64;
          float acc = 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint32_t aw[4] = {a[j][u].x, a[j][u].y, a[j][u].z, a[j][u].w};
            const uint32_t gw[4] = {g[j][u].x, g[j][u].y, g[j][u].z, g[j][u].w};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const float2 af = unpack_bf16(aw[w]), gf = unpack_bf16(gw[w]);
              acc = fmaf(af.x, gf.x, acc);
              acc = fmaf(af.y, gf.y, acc);
            }
          }
          if (q < nt * 128) {
            aux[q] = l2[j];
            aux[256 + q] = acc;
          }
        }
      }
      mbar_arrive(&bar_aux[slot]);
    }
  } else if (warp < 8) {
    // ------------------------------ math + output warps ----------------------------------------
    const int quarter = warp & 3, hf = warp >> 2;
    const int r = quarter * 32 + lane;  // row inside the 128-row tile (TMEM lane)
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sl2 = p.scale * ATC_LOG2E;
    const long long rs = 3LL * hd;

    // This thread's row of TMEM columns [col, col+64) -> 64 bf16 = one 128-byte line of dqkv slot
    // `slot`. The rows leave through TMA tensor stores, 16 at a time, from a 2 KiB per-warp staging
    // tile in the 128B-swizzle layout (thread-per-row STS.128 is bank-conflict free there): 32 LSU
    // wavefronts per 32 rows. Written straight from registers (four 256-bit stores per thread) every
    // instruction touched 32 different lines — 128 wavefronts per 32 rows — and the drains alone held
    // the kernel's load/store pipe for a fifth of its run time (FEDVIT_ATTN_DBG = 4: 282 -> 223 us).
    // The 3-D tensor map clips rows past the sequence end, so ragged tiles need no predicate.
    uint8_t* stage = sStage + warp * ATB_STAGE;
    bool store_pending = false;  // lane 0: a tensor store of this warp may still be reading `stage`
    auto store_row = [&](uint32_t col, int b, int h, int slot, int tile) {
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(lane_base + col, o0);
      tmem_ld_32x32(lane_base + col + 32, o1);
      tmem_ld_wait();
      const int tok0 = tile * 128 + quarter * 32;  // first row of this warp
      if (tok0 >= p.N) return;  // warp-uniform
      uint32_t w[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        w[i] = pack_bf16(__uint_as_float(o0[2 * i]), __uint_as_float(o0[2 * i + 1]));
        w[16 + i] = pack_bf16(__uint_as_float(o1[2 * i]), __uint_as_float(o1[2 * i + 1]));
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (tok0 + half * 16 >= p.N) break;  // warp-uniform
        if (store_pending) {
          if (lane == 0) tma_store_wait_read();
          store_pending = false;
        }
        __syncwarp();
        if ((lane >> 4) == half) {
          uint8_t* srow = stage + (lane & 15) * 128;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            *reinterpret_cast<uint4*>(srow + ((u ^ (lane & 7)) << 4)) =
                make_uint4(w[u * 4], w[u * 4 + 1], w[u * 4 + 2], w[u * 4 + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmap_dqkv, stage, slot * hd + h * 64, tok0 + half * 16, b);
          tma_store_commit();
        }
        store_pending = true;
      }
    };
    // Drains are deferred by one iteration: dK / dV of a finished key block (and dQ of a finished
    // item) leave tensor memory AFTER this thread has produced the next iteration's P / dS, so the
    // wait for their last MMA 2 overlaps useful work instead of idling the math warps (ncu: 13 % of
    // their samples sat in that wait).
    bool pend_kv = false, pend_dq = false;
    int kv_b = 0, kv_h = 0, kv_kb = 0, dq_b = 0, dq_h = 0;
    auto drain = [&]() {
      if (pend_kv) {
        store_row(hf == 0 ? T_DK : T_DV, kv_b, kv_h, hf == 0 ? 1 : 2, kv_kb);
        tc_fence_before();
        mbar_arrive(bar_kvfree);
        pend_kv = false;
      }
      if (pend_dq) {
        if (hf < nt) store_row(T_DQ + hf * 64, dq_b, dq_h, 0, hf);
        tc_fence_before();
        mbar_arrive(bar_dqfree);
        pend_dq = false;
      }
    };
    int it = 0;
    int n = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++n) {
      const int b = item / p.H, h = item % p.H;
      const float* aux = sAux + (n % ATB_AUX) * 512;
      mbar_wait(&bar_aux[n % ATB_AUX], (n / ATB_AUX) & 1);
      for (int kb = 0; kb < nt; ++kb) {
        for (int qt = 0; qt < nt; ++qt, ++it) {
          const float l2 = aux[qt * 128 + r];
          const float dl = aux[256 + qt * 128 + r];
          ATB_STAMP(it, 0);
          mbar_wait(bar_s, it & 1);
          tc_fence_after();
          ATB_STAMP(it, 1);
          uint32_t pk[2][16], dk[2][16];
          const bool rows_live = qt * 128 + quarter * 32 < p.N;  // warp-uniform
          bool consumed = false;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            // 32-key chunks are dealt out alternately (group 0: chunks 0 and 2, group 1: 1 and 3) so a
            // short last key block (69 keys at N = 197) still splits evenly between the two groups
            const int ch = 2 * c + hf;
            const int key0 = kb * 128 + ch * 32;
            if (!rows_live || key0 >= p.N) {  // nothing but padding here: P = dS = 0, no TMEM traffic
#pragma unroll
              for (int i = 0; i < 16; ++i) { pk[c][i] = 0u; dk[c][i] = 0u; }
              continue;
            }
            uint32_t s[32], d[32];
            tmem_ld_32x32(lane_base + T_S + ch * 32, s);
            tmem_ld_32x32(lane_base + T_DP + ch * 32, d);
            tmem_ld_wait();
            if (c == 1) {  // this thread's last read of S / dP
              tc_fence_before();
              mbar_arrive(bar_c);
              consumed = true;
              ATB_STAMP(it, 2);
            }
            if (key0 + 32 <= p.N) {  // whole chunk inside the sequence: no per-element masking
              const float dls = dl * p.scale;
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float p0 = atc_ex2(fmaf(__uint_as_float(s[i]), sl2, -l2));
                const float p1 = atc_ex2(fmaf(__uint_as_float(s[i + 1]), sl2, -l2));
                const float s0 = p0 * fmaf(__uint_as_float(d[i]), p.scale, -dls);
                const float s1 = p1 * fmaf(__uint_as_float(d[i + 1]), p.scale, -dls);
                pk[c][i >> 1] = pack_bf16(p0, p1);
                dk[c][i >> 1] = pack_bf16(s0, s1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                float p0 = 0.f, p1 = 0.f, s0 = 0.f, s1 = 0.f;
                if (key0 + i < p.N) {
                  p0 = atc_ex2(fmaf(__uint_as_float(s[i]), sl2, -l2));
                  s0 = p0 * (__uint_as_float(d[i]) - dl) * p.scale;
                }
                if (key0 + i + 1 < p.N) {
                  p1 = atc_ex2(fmaf(__uint_as_float(s[i + 1]), sl2, -l2));
                  s1 = p1 * (__uint_as_float(d[i + 1]) - dl) * p.scale;
                }
                pk[c][i >> 1] = pack_bf16(p0, p1);
                dk[c][i >> 1] = pack_bf16(s0, s1);
              }
            }
          }
          if (!consumed) {  // second chunk skipped (padding): nothing left to read
            tc_fence_before();
            mbar_arrive(bar_c);
          }
          ATB_STAMP(it, 3);
          if (it > 0) {
            mbar_wait(bar_m2, (it - 1) & 1);  // previous MMA 2 retired: sP / sdS free, its dK / dV / dQ final
            tc_fence_after();
          }
          ATB_STAMP(it, 4);
          // chunk ch = 2c + hf: column block ch >> 1 == c (64 keys each), 64-byte half ch & 1 == hf
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint8_t* prow = sP + c * ATB_TILE + r * 128;
            uint8_t* srow = sdS + c * ATB_TILE + r * 128;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int unit = (hf * 4 + u) ^ (r & 7);
              *reinterpret_cast<uint4*>(prow + (unit << 4)) =
                  make_uint4(pk[c][u * 4], pk[c][u * 4 + 1], pk[c][u * 4 + 2], pk[c][u * 4 + 3]);
              *reinterpret_cast<uint4*>(srow + (unit << 4)) =
                  make_uint4(dk[c][u * 4], dk[c][u * 4 + 1], dk[c][u * 4 + 2], dk[c][u * 4 + 3]);
            }
          }
          fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async proxy
          tc_fence_before();
          mbar_arrive(bar_p);
          ATB_STAMP(it, 5);
          // What bounds this kernel (profiles/r2_attn_trace_bwd.txt): shared-memory bandwidth. Per cell the five
          // products read 208 KB of operands and the math warps store 64 KB of P / dS — 2 100 cycles at 128 B per
          // cycle; a full cell takes 2 200 - 2 800, and with the tile loads and the drains' staging an item moves
          // ~1.4 MB = 11 000 of its 17 400 cycles. Tried against it in round 2, all measured, none kept: the next
          // item's first scores issued behind this item's last bar_c (the 1 600 - 3 400 cycle wait at every item
          // start shrinks to 1 000, the item takes as long); drains in 8-row steps through two staging halves
          // (242 us), a coalesced LDS + STG flush instead of the tensor store (221 us); dedicated drain warps
          // (16 warps at 128 registers, math warps storing P / dS per chunk: 226 us — the time moves from the
          // drains to the wait for the previous cell's second MMA group, sP / sdS and the accumulators being
          // single-buffered: shared and tensor memory are both full).
          drain();  // whatever finished with the previous iteration's MMA 2
          ATB_STAMP(it, 6);
          if (qt == nt - 1) {
            pend_kv = true;
            kv_b = b; kv_h = h; kv_kb = kb;
          }
        }
      }
      mbar_arrive(&bar_auxfree[n % ATB_AUX]);
      pend_dq = true;
      dq_b = b; dq_h = h;
    }
    if (it > 0) {
      mbar_wait(bar_m2, (it - 1) & 1);
      tc_fence_after();
      drain();
    }
    if (lane == 0) tma_store_wait_all();  // this warp's tensor stores are complete before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

static int make_tok_map(CUtensorMap* map, const void* base, int64_t batch, int64_t tokens, int64_t width) {
  return make_qkv_map(map, base, batch, tokens, width, 128);
}

int attention_tc_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                     int64_t batch, int64_t tokens, int64_t heads, float scale, cudaStream_t stream) {
  AttnBwdParams p;
  p.N = static_cast<int>(tokens);
  p.H = static_cast<int>(heads);
  p.kw = static_cast<int>((tokens + 15) / 16 * 16);
  p.items = static_cast<int>(batch * heads);
  p.scale = scale;
  p.lse = lse;
  p.o = reinterpret_cast<const __nv_bfloat16*>(out);
  p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  CUtensorMap mq, mdo;
  int rc = make_tok_map(&mq, qkv, batch, tokens, 3 * heads * 64);
  if (rc != FV_OK) return rc;
  rc = make_tok_map(&mdo, dout, batch, tokens, heads * 64);
  if (rc != FV_OK) return rc;
  CUtensorMap mdq;  // output: 16-row boxes of one 64-column (slot, head) group; rows past N are clipped
  rc = make_qkv_map(&mdq, dqkv, batch, tokens, 3 * heads * 64, 16);
  if (rc != FV_OK) return rc;
  static bool configured = false;
  if (!configured) {
    FV_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATB_SMEM));
    configured = true;
  }
  const int grid = p.items < num_sms() ? p.items : num_sms();
  FV_CHECK_CUDA(fv::launch_pdl(attn_tc_bwd_kernel, dim3(static_cast<unsigned>(grid)), dim3(ATB_THREADS), ATB_SMEM, stream, mq, mdo, mdq, p));
  count_kernel(FV_KERNEL_ATTN_BWD);
  FV_LAUNCH_CHECK();
  return FV_OK;
}


// ---------------------------------------------------------------------------------------------
// long sequences (256 < N <= 768: ViT-L/16 @ 384 has 577 tokens): forward
// ---------------------------------------------------------------------------------------------
// One persistent CTA per SM walks (batch, head) items. K and V of the item stay resident in shared
// memory as 128-key tiles; the 128-query tiles stream through a two-slot ring. Exact two-pass
// softmax per query tile, no online rescaling:
//   pass A  S_j = Q K_j^T for every key block j  -> row maxima
//   pass B  S_j again -> P_j = exp2(scale*log2e*(S_j - max)) written back into TMEM as packed bf16
//           -> O += P_j V_j  (A operand from TMEM, accumulated in TMEM across the key blocks)
// The score MMA is the cheap part (the softmax's MUFU work sets the pace), so recomputing it costs
// less than the rescale traffic of the online form. S is double-buffered (2 x 128 TMEM columns):
// the MMA of block j+1 runs while the softmax warps work on block j.
constexpr int ATL_THREADS = 192;  // warps 0-3 softmax / epilogue, warp 4 MMA issue, warp 5 TMEM alloc + TMA producer
constexpr int ATL_MAX_KB = 6;     // 768 keys
constexpr int ATL_TILE = 128 * 128;  // one [128 x 64] bf16 tile

struct AttnLongParams {
  int N, H, kw, nkb, nqt, items;
  float scale;
  __nv_bfloat16* out;
  float* lse;
};

__global__ void __launch_bounds__(ATL_THREADS, 1)
attn_tc_fwd_long_kernel(const __grid_constant__ CUtensorMap tmap, const AttnLongParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                        // 2 slots
  uint8_t* sK = sQ + 2 * ATL_TILE;           // nkb tiles
  uint8_t* sV = sK + p.nkb * ATL_TILE;       // nkb tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + p.nkb * ATL_TILE);
  uint64_t* bar_k = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_kfree = bars + 2;
  uint64_t* bar_vfree = bars + 3;
  uint64_t* bar_q = bars + 4;       // [2]
  uint64_t* bar_qfree = bars + 6;   // [2]
  uint64_t* bar_s = bars + 8;       // [2] score buffer full
  uint64_t* bar_free = bars + 10;   // [2] score buffer consumed by the softmax warps
  uint64_t* bar_p = bars + 12;      // [2] P of a pass-B block written; ping-pong: a warp without live rows
                                    // runs one block ahead, its early arrival must not count for this block
  uint64_t* bar_o = bars + 14;
  uint64_t* bar_odone = bars + 15;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hd = p.H * 64;
  const int nkb = p.nkb, nqt = p.nqt;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_kfree, 1);
    mbar_init(bar_vfree, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_q[i], 1);
      mbar_init(&bar_qfree[i], 1);
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_free[i], 128);
    }
    mbar_init(&bar_p[0], 128);
    mbar_init(&bar_p[1], 128);
    mbar_init(bar_o, 1);
    mbar_init(bar_odone, 128);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  constexpr uint32_t T_O = 256;

  if (warp == 5) {
    // ------------------------------ TMA producer ------------------------------------------------
    int n = 0, gt = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      const int b = item / p.H, h = item % p.H;
      if (n > 0) mbar_wait(bar_kfree, (n - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(bar_k, nkb * ATL_TILE);
        for (int j = 0; j < nkb; ++j) tma_load_3d(sK + j * ATL_TILE, &tmap, bar_k, hd + h * 64, j * 128, b);
      }
      __syncwarp();
      if (n > 0) mbar_wait(bar_vfree, (n - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(bar_v, nkb * ATL_TILE);
        for (int j = 0; j < nkb; ++j) tma_load_3d(sV + j * ATL_TILE, &tmap, bar_v, 2 * hd + h * 64, j * 128, b);
      }
      __syncwarp();
      for (int t = 0; t < nqt; ++t, ++gt) {
        const int slot = gt & 1;
        if (gt >= 2) mbar_wait(&bar_qfree[slot], ((gt >> 1) - 1) & 1);
        if (elect_one()) {
          mbar_expect_tx(&bar_q[slot], ATL_TILE);
          tma_load_3d(sQ + slot * ATL_TILE, &tmap, &bar_q[slot], h * 64, t * 128, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 4) {
    // ------------------------------ MMA issue (whole warp, elected lane issues) -----------------
    const uint32_t idesc_o = make_idesc(kFmtBF16, 0, 1, 128, 64);
    int n = 0, gt = 0, u = 0, pb = 0;  // items, query tiles, score-buffer uses, pass-B blocks so far
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      mbar_wait(bar_k, n & 1);
      for (int t = 0; t < nqt; ++t, ++gt) {
        const int slot = gt & 1;
        mbar_wait(&bar_q[slot], (gt >> 1) & 1);
        const uint64_t dq = make_smem_desc_sw128(smem_u32(sQ + slot * ATL_TILE), 16, 1024);
        auto issue_s = [&](int j) {
          const int buf = u & 1;
          mbar_wait(&bar_free[buf], ((u >> 1) & 1) ^ 1);  // the buffer's previous tenant has been read
          tc_fence_after();
          int kwb = p.kw - j * 128;
          if (kwb > 128) kwb = 128;
          const uint32_t idesc_s = make_idesc(kFmtBF16, 0, 0, 128, kwb);
          const uint64_t dk = make_smem_desc_sw128(smem_u32(sK + j * ATL_TILE), 16, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem + buf * 128, dq + k * 2, dk + k * 2, idesc_s, k > 0 ? 1u : 0u);
            umma_commit(&bar_s[buf]);
          }
          __syncwarp();
          ++u;
        };
        auto issue_pv = [&](int j, int buf, bool last) {
          mbar_wait(&bar_p[pb & 1], (pb >> 1) & 1);
          ++pb;
          tc_fence_after();
          int kwb = p.kw - j * 128;
          if (kwb > 128) kwb = 128;
          const uint64_t dv = make_smem_desc_sw128(smem_u32(sV + j * ATL_TILE), 64 * 128, 1024);
          const int ksteps = kwb >> 4;
          if (elect_one()) {
            for (int k = 0; k < ksteps; ++k)
              umma_bf16_ts(tmem + T_O, tmem + buf * 128 + k * 8, dv + k * (2048 >> 4), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
            if (last) {
              umma_commit(bar_o);
              if (t == nqt - 1) umma_commit(bar_vfree);
            }
          }
          __syncwarp();
        };
        for (int j = 0; j < nkb; ++j) issue_s(j);  // pass A: scores for the row maxima
        const int u0 = u;
        for (int j = 0; j < nkb; ++j) {            // pass B: scores again, PV one block behind
          issue_s(j);
          if (j == nkb - 1 && elect_one()) {
            umma_commit(&bar_qfree[slot]);            // this query tile's last reader
            if (t == nqt - 1) umma_commit(bar_kfree);
          }
          __syncwarp();
          if (j == 0) {
            if (t == 0) mbar_wait(bar_v, n & 1);
            if (gt > 0) mbar_wait(bar_odone, (gt - 1) & 1);  // previous tile's O has left TMEM
          } else {
            issue_pv(j - 1, (u0 + j - 1) & 1, false);
          }
        }
        issue_pv(nkb - 1, (u0 + nkb - 1) & 1, true);
      }
    }
  } else if (warp < 4) {
    // ------------------------------ softmax + epilogue -------------------------------------------
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const float sl2 = p.scale * ATC_LOG2E;
    int gt = 0, u = 0, pb = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int b = item / p.H, h = item % p.H;
      for (int t = 0; t < nqt; ++t, ++gt) {
        const int q = t * 128 + warp * 32 + lane;
        const bool warp_live = t * 128 + warp * 32 < p.N;  // warp-uniform
        float mx = -INFINITY, sum = 0.f;
        for (int pass = 0; pass < 2; ++pass) {
          const float mxs = mx * sl2;
          for (int j = 0; j < nkb; ++j, ++u) {
            const int buf = u & 1;
            mbar_wait(&bar_s[buf], (u >> 1) & 1);
            tc_fence_after();
            if (warp_live) {
              int kwb = p.kw - j * 128;
              if (kwb > 128) kwb = 128;
              const int nch = (kwb + 31) >> 5;
              for (int c = 0; c < nch; ++c) {
                const int key0 = j * 128 + c * 32;
                uint32_t r[32];
                tmem_ld_32x32(taddr + buf * 128 + c * 32, r);
                tmem_ld_wait();
                if (pass == 0) {
                  if (key0 + 32 <= p.N) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
                  } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                      if (key0 + i < p.N) mx = fmaxf(mx, __uint_as_float(r[i]));
                  }
                } else {
                  uint32_t pk[16];
                  if (key0 + 32 <= p.N) {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                      const float p0 = atc_ex2(fmaf(__uint_as_float(r[i]), sl2, -mxs));
                      const float p1 = atc_ex2(fmaf(__uint_as_float(r[i + 1]), sl2, -mxs));
                      sum += p0 + p1;
                      pk[i >> 1] = pack_bf16(p0, p1);
                    }
                  } else {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                      const float p0 = (key0 + i < p.N) ? atc_ex2(fmaf(__uint_as_float(r[i]), sl2, -mxs)) : 0.f;
                      const float p1 = (key0 + i + 1 < p.N) ? atc_ex2(fmaf(__uint_as_float(r[i + 1]), sl2, -mxs)) : 0.f;
                      sum += p0 + p1;
                      pk[i >> 1] = pack_bf16(p0, p1);
                    }
                  }
                  tmem_st_32x16(taddr + buf * 128 + c * 16, pk);  // over columns this thread has consumed
                }
              }
              if (pass == 1) tmem_st_wait();
            }
            tc_fence_before();
            if (pass == 1) {
              mbar_arrive(&bar_p[pb & 1]);
              ++pb;
            }
            mbar_arrive(&bar_free[buf]);
          }
        }
        mbar_wait(bar_o, gt & 1);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        if (warp_live) {
          tmem_ld_32x32(taddr + T_O, o0);
          tmem_ld_32x32(taddr + T_O + 32, o1);
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive(bar_odone);
        if (warp_live && q < p.N) {
          const float inv = 1.0f / sum;
          uint32_t w[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            w[i] = pack_bf16(__uint_as_float(o0[2 * i]) * inv, __uint_as_float(o0[2 * i + 1]) * inv);
            w[16 + i] = pack_bf16(__uint_as_float(o1[2 * i]) * inv, __uint_as_float(o1[2 * i + 1]) * inv);
          }
          __nv_bfloat16* dst = p.out + ((static_cast<long long>(b) * p.N + q) * p.H + h) * 64;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) st_v8(dst + jj * 16, w + jj * 8);
          p.lse[(static_cast<long long>(b) * p.H + h) * p.N + q] = mx * p.scale + logf(sum);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int attention_tc_fwd_long(const void* qkv, void* out, float* lse, int64_t batch, int64_t tokens, int64_t heads,
                          float scale, cudaStream_t stream) {
  AttnLongParams p;
  p.N = static_cast<int>(tokens);
  p.H = static_cast<int>(heads);
  p.kw = static_cast<int>((tokens + 15) / 16 * 16);
  p.nkb = (p.kw + 127) / 128;
  p.nqt = (p.N + 127) / 128;
  p.items = static_cast<int>(batch * heads);
  p.scale = scale;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  FV_CHECK_ARG(p.nkb <= ATL_MAX_KB, "attention_tc_fwd_long: at most %d tokens", ATL_MAX_KB * 128);
  CUtensorMap map;
  int rc = make_qkv_map(&map, qkv, batch, tokens, 3 * heads * 64, 128);
  if (rc != FV_OK) return rc;
  const int smem = (2 + 2 * p.nkb) * ATL_TILE + 1024 + 256;
  static int configured = 0;
  if (configured < smem) {
    FV_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_fwd_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  const int grid = p.items < num_sms() ? p.items : num_sms();
  FV_CHECK_CUDA(fv::launch_pdl(attn_tc_fwd_long_kernel, dim3(static_cast<unsigned>(grid)), dim3(ATL_THREADS), smem, stream,
                               map, p));
  count_kernel(FV_KERNEL_ATTN_FWD_LONG);
  FV_LAUNCH_CHECK();
  return FV_OK;
}


// ---------------------------------------------------------------------------------------------
// long sequences (256 < N <= 768): backward
// ---------------------------------------------------------------------------------------------
// Work item = (batch, head, 128-key block). K / V of the block stay resident (double-buffered across
// items), the 128-query tiles with their dO stream through a two-slot ring. Per query tile, exactly
// the iteration of the short kernel: S = Q K^T and dP = dO V^T into TMEM, P / dS through swizzled
// smem tiles, then dV += P^T dO and dK += dS^T Q (accumulated in TMEM over the query tiles, stored
// once per item as bf16) and the block's dQ contribution dS K into one of two TMEM buffers. That
// contribution is a partial sum over key blocks: it leaves as fp32 through a per-warp staging tile
// and a TMA reduce-add (cp.reduce.async.bulk.tensor .add) into an fp32 scratch [B, N, H*64], which
// a small pass converts to bf16 into dqkv afterwards. delta = rowsum(dO * O) comes from the
// attn_delta pre-pass. No recomputation of the probabilities (the two-sweep mma.sync path computes
// them twice), no software atomics.
constexpr int ATLB_THREADS = 320;  // warps 0-7 math, warp 8 MMA issue, warp 9 TMEM alloc + TMA producer

struct AttnLongBwdParams {
  int N, H, kw, nkb, nqt, items;
  float scale;
  const float* lse;
  const float* delta;
  __nv_bfloat16* dqkv;
};

__global__ void __launch_bounds__(ATLB_THREADS, 1)
attn_tc_bwd_long_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                        const __grid_constant__ CUtensorMap tmap_dq, const AttnLongBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sKV = smem;                       // 2 slots x {K, V}
  uint8_t* sQD = sKV + 4 * ATB_TILE;         // 2 slots x {Q, dO}
  uint8_t* sP = sQD + 4 * ATB_TILE;          // 2 column blocks of 64 keys
  uint8_t* sdS = sP + 2 * ATB_TILE;          // 2 column blocks
  uint8_t* sStg = sdS + 2 * ATB_TILE;        // 8 warps x 4 KiB: fp32 dQ tiles on their way to the TMA reduce
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + 8 * 4096);
  uint64_t* bar_kv = bars + 0;       // [2]
  uint64_t* bar_kvfree = bars + 2;   // [2]
  uint64_t* bar_qd = bars + 4;       // [2]
  uint64_t* bar_qdfree = bars + 6;   // [2]
  uint64_t* bar_s = bars + 8;
  uint64_t* bar_p = bars + 9;
  uint64_t* bar_m2 = bars + 10;
  uint64_t* bar_dqfree = bars + 11;  // [2] dQ TMEM buffer drained
  uint64_t* bar_kvdone = bars + 13;  // dK / dV of the finished item drained
  uint64_t* bar_c = bars + 14;       // S / dP are in the math warps' registers: TMEM may be rewritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hd = p.H * 64;
  const int nqt = p.nqt, nkb = p.nkb;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_dq);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_kv[i], 1);
      mbar_init(&bar_kvfree[i], 1);
      mbar_init(&bar_qd[i], 1);
      mbar_init(&bar_qdfree[i], 1);
      mbar_init(&bar_dqfree[i], 256);
    }
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 256);
    mbar_init(bar_m2, 1);
    mbar_init(bar_kvdone, 256);
    mbar_init(bar_c, 256);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  constexpr uint32_t T_S = 0, T_DP = 128, T_DK = 256, T_DV = 320, T_DQ = 384;  // T_DQ: two 64-column buffers

  // item -> (batch, head, key block): key blocks of one (batch, head) are consecutive items, so its
  // Q / dO tiles are re-read from L2
  auto decode = [&](int item, int& b, int& h, int& kb) {
    kb = item % nkb;
    const int bh = item / nkb;
    h = bh % p.H;
    b = bh / p.H;
  };

  if (warp == 9) {
    // ------------------------------ TMA producer ------------------------------------------------
    auto load_kv = [&](int n, int item) {
      int b, h, kb;
      decode(item, b, h, kb);
      const int ks = n & 1;
      if (n >= 2) mbar_wait(&bar_kvfree[ks], ((n >> 1) - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(&bar_kv[ks], 2 * ATB_TILE);
        tma_load_3d(sKV + ks * 2 * ATB_TILE, &tmap_qkv, &bar_kv[ks], hd + h * 64, kb * 128, b);
        tma_load_3d(sKV + ks * 2 * ATB_TILE + ATB_TILE, &tmap_qkv, &bar_kv[ks], 2 * hd + h * 64, kb * 128, b);
      }
      __syncwarp();
    };
    int n = 0, it = 0;
    if (static_cast<int>(blockIdx.x) < p.items) load_kv(0, blockIdx.x);
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      int b, h, kb;
      decode(item, b, h, kb);
      for (int qt = 0; qt < nqt; ++qt, ++it) {
        const int slot = it & 1;
        if (it >= 2) mbar_wait(&bar_qdfree[slot], ((it >> 1) - 1) & 1);
        if (elect_one()) {
          mbar_expect_tx(&bar_qd[slot], 2 * ATB_TILE);
          tma_load_3d(sQD + slot * 2 * ATB_TILE, &tmap_qkv, &bar_qd[slot], h * 64, qt * 128, b);
          tma_load_3d(sQD + slot * 2 * ATB_TILE + ATB_TILE, &tmap_do, &bar_qd[slot], h * 64, qt * 128, b);
        }
        __syncwarp();
        if (qt == 0 && item + static_cast<int>(gridDim.x) < p.items) load_kv(n + 1, item + gridDim.x);  // next item's K / V early
      }
    }
  } else if (warp == 8) {
    // ------------------------------ MMA issue (whole warp, elected lane issues) -----------------
    const uint32_t id_dvk = make_idesc(kFmtBF16, 1, 1, 128, 64);  // A MN-major (P^T / dS^T), B MN-major
    const uint32_t id_dq = make_idesc(kFmtBF16, 0, 1, 128, 64);   // A K-major (dS), B MN-major (K)
    auto kwb_of = [&](int kb) {
      int w = p.kw - kb * 128;
      return w > 128 ? 128 : w;
    };
    auto issue_mma1 = [&](int ks, int slot, int kwb) {
      const uint32_t id_s = make_idesc(kFmtBF16, 0, 0, 128, kwb);
      const uint64_t dK_k = make_smem_desc_sw128(smem_u32(sKV + ks * 2 * ATB_TILE), 16, 1024);
      const uint64_t dV_k = make_smem_desc_sw128(smem_u32(sKV + ks * 2 * ATB_TILE + ATB_TILE), 16, 1024);
      const uint64_t dQ_k = make_smem_desc_sw128(smem_u32(sQD + slot * 2 * ATB_TILE), 16, 1024);
      const uint64_t dO_k = make_smem_desc_sw128(smem_u32(sQD + slot * 2 * ATB_TILE + ATB_TILE), 16, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + T_S, dQ_k + k * 2, dK_k + k * 2, id_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + T_DP, dO_k + k * 2, dV_k + k * 2, id_s, k > 0 ? 1u : 0u);
        umma_commit(bar_s);
      }
      __syncwarp();
    };
    int n = 0, it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      int b, h, kb;
      decode(item, b, h, kb);
      const int ks = n & 1;
      const int kwb = kwb_of(kb);
      const bool has_next = item + static_cast<int>(gridDim.x) < p.items;
      if (n == 0) {
        mbar_wait(&bar_kv[0], 0);
        mbar_wait(&bar_qd[0], 0);
        tc_fence_after();
        issue_mma1(0, 0, kwb);
      }
      const uint64_t dK_mn = make_smem_desc_sw128(smem_u32(sKV + ks * 2 * ATB_TILE), ATB_TILE, 1024);  // MN-major view
      for (int qt = 0; qt < nqt; ++qt, ++it) {
        const int slot = it & 1;
        const uint64_t dQ_mn = make_smem_desc_sw128(smem_u32(sQD + slot * 2 * ATB_TILE), ATB_TILE, 1024);
        const uint64_t dO_mn = make_smem_desc_sw128(smem_u32(sQD + slot * 2 * ATB_TILE + ATB_TILE), ATB_TILE, 1024);
        // the next iteration's scores are issued as soon as this one's S / dP are in registers
        mbar_wait(bar_c, it & 1);
        tc_fence_after();
        if (qt + 1 < nqt) {
          mbar_wait(&bar_qd[(it + 1) & 1], ((it + 1) >> 1) & 1);
          tc_fence_after();
          issue_mma1(ks, (it + 1) & 1, kwb);
        } else if (has_next) {
          int b2, h2, kb2;
          decode(item + gridDim.x, b2, h2, kb2);
          mbar_wait(&bar_kv[(n + 1) & 1], ((n + 1) >> 1) & 1);
          mbar_wait(&bar_qd[(it + 1) & 1], ((it + 1) >> 1) & 1);
          tc_fence_after();
          issue_mma1((n + 1) & 1, (it + 1) & 1, kwb_of(kb2));
        }
        mbar_wait(bar_p, it & 1);  // P / dS tiles written
        tc_fence_after();
        if (qt == 0 && n > 0) mbar_wait(bar_kvdone, (n - 1) & 1);  // previous item's dK / dV have left TMEM
        if (it >= 2) mbar_wait(&bar_dqfree[it & 1], ((it >> 1) - 1) & 1);  // dQ buffer of iteration it-2 drained
        tc_fence_after();
        const uint64_t dP_mn = make_smem_desc_sw128(smem_u32(sP), ATB_TILE, 1024);
        const uint64_t dS_mn = make_smem_desc_sw128(smem_u32(sdS), ATB_TILE, 1024);
        const uint64_t dS_k0 = make_smem_desc_sw128(smem_u32(sdS), 16, 1024);
        const int ks16 = kwb >> 4;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tmem + T_DV, dP_mn + k * 128, dO_mn + k * 128, id_dvk, (qt > 0 || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tmem + T_DK, dS_mn + k * 128, dQ_mn + k * 128, id_dvk, (qt > 0 || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (k < ks16) {
              const uint64_t dS_k = dS_k0 + (((k >> 2) * ATB_TILE + (k & 3) * 32) >> 4);
              umma_bf16(tmem + T_DQ + (it & 1) * 64, dS_k, dK_mn + k * 128, id_dq, k > 0 ? 1u : 0u);
            }
          }
          umma_commit(bar_m2);
          umma_commit(&bar_qdfree[slot]);
          if (qt == nqt - 1) umma_commit(&bar_kvfree[ks]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 8) {
    // ------------------------------ math + output warps ----------------------------------------
    const int quarter = warp & 3, hf = warp >> 2;
    const int r = quarter * 32 + lane;  // row inside the 128-row tile (TMEM lane)
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sl2 = p.scale * ATC_LOG2E;
    const long long rs = 3LL * hd;
    uint8_t* stg = sStg + warp * 4096;
    bool stg_pending = false;

    // dK (group 0) / dV (group 1): this thread's row of 64 bf16 straight from registers
    auto store_kv_row = [&](int b, int h, int kb) {
      uint32_t o0[32], o1[32];
      const uint32_t col = hf == 0 ? T_DK : T_DV;
      tmem_ld_32x32(lane_base + col, o0);
      tmem_ld_32x32(lane_base + col + 32, o1);
      tmem_ld_wait();
      const int tok = kb * 128 + r;
      if (tok < p.N) {
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          w[i] = pack_bf16(__uint_as_float(o0[2 * i]), __uint_as_float(o0[2 * i + 1]));
          w[16 + i] = pack_bf16(__uint_as_float(o1[2 * i]), __uint_as_float(o1[2 * i + 1]));
        }
        __nv_bfloat16* dst = p.dqkv + (static_cast<long long>(b) * p.N + tok) * rs + (hf == 0 ? 1 : 2) * hd + h * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j) st_v8(dst + j * 16, w + j * 8);
      }
    };
    // this warp's 32 rows x 32 fp32 columns of a dQ partial tile -> staging -> TMA reduce-add
    auto reduce_dq = [&](int buf, int b, int h, int qt) {
      uint32_t v[32];
      tmem_ld_32x32(lane_base + T_DQ + buf * 64 + hf * 32, v);
      tmem_ld_wait();
      if (stg_pending) {
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        *reinterpret_cast<uint4*>(stg + lane * 128 + ((u ^ (lane & 7)) << 4)) =
            make_uint4(v[u * 4], v[u * 4 + 1], v[u * 4 + 2], v[u * 4 + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_3d(&tmap_dq, stg, h * 64 + hf * 32, qt * 128 + quarter * 32, b);
        tma_store_commit();
      }
      stg_pending = true;
    };

    int it = 0, n = 0;
    int pb = 0, ph = 0, pqt = 0;              // (batch, head, query tile) of the previous iteration
    bool pend_kv = false;
    int kv_b = 0, kv_h = 0, kv_kb = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      int b, h, kb;
      decode(item, b, h, kb);
      const float* lse_bh = p.lse + (static_cast<long long>(b) * p.H + h) * p.N;
      const float* del_bh = p.delta + (static_cast<long long>(b) * p.H + h) * p.N;
      for (int qt = 0; qt < nqt; ++qt, ++it) {
        const int q = qt * 128 + r;
        const float l2 = q < p.N ? __ldg(lse_bh + q) * ATC_LOG2E : INFINITY;  // in flight during the wait below
        const float dl = q < p.N ? __ldg(del_bh + q) : 0.f;
        mbar_wait(bar_s, it & 1);
        tc_fence_after();
        uint32_t pk[2][16], dk[2][16];
        const bool rows_live = qt * 128 + quarter * 32 < p.N;  // warp-uniform
        bool consumed = false;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int ch = 2 * c + hf;  // 32-key chunks dealt alternately to the two groups
          const int key0 = kb * 128 + ch * 32;
          if (!rows_live || key0 >= p.N) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { pk[c][i] = 0u; dk[c][i] = 0u; }
            continue;
          }
          uint32_t s[32], d[32];
          tmem_ld_32x32(lane_base + T_S + ch * 32, s);
          tmem_ld_32x32(lane_base + T_DP + ch * 32, d);
          tmem_ld_wait();
          if (c == 1) {  // this thread's last read of S / dP
            tc_fence_before();
            mbar_arrive(bar_c);
            consumed = true;
          }
          if (key0 + 32 <= p.N) {
            const float dls = dl * p.scale;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = atc_ex2(fmaf(__uint_as_float(s[i]), sl2, -l2));
              const float p1 = atc_ex2(fmaf(__uint_as_float(s[i + 1]), sl2, -l2));
              const float s0 = p0 * fmaf(__uint_as_float(d[i]), p.scale, -dls);
              const float s1 = p1 * fmaf(__uint_as_float(d[i + 1]), p.scale, -dls);
              pk[c][i >> 1] = pack_bf16(p0, p1);
              dk[c][i >> 1] = pack_bf16(s0, s1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float p0 = 0.f, p1 = 0.f, s0 = 0.f, s1 = 0.f;
              if (key0 + i < p.N) {
                p0 = atc_ex2(fmaf(__uint_as_float(s[i]), sl2, -l2));
                s0 = p0 * (__uint_as_float(d[i]) - dl) * p.scale;
              }
              if (key0 + i + 1 < p.N) {
                p1 = atc_ex2(fmaf(__uint_as_float(s[i + 1]), sl2, -l2));
                s1 = p1 * (__uint_as_float(d[i + 1]) - dl) * p.scale;
              }
              pk[c][i >> 1] = pack_bf16(p0, p1);
              dk[c][i >> 1] = pack_bf16(s0, s1);
            }
          }
        }
        if (!consumed) {  // second chunk skipped (padding): nothing left to read
          tc_fence_before();
          mbar_arrive(bar_c);
        }
        if (it > 0) {
          mbar_wait(bar_m2, (it - 1) & 1);  // previous MMA 2 retired: sP / sdS free, its dQ / dK / dV final
          tc_fence_after();
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint8_t* prow = sP + c * ATB_TILE + r * 128;
          uint8_t* srow = sdS + c * ATB_TILE + r * 128;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int unit = (hf * 4 + u) ^ (r & 7);
            *reinterpret_cast<uint4*>(prow + (unit << 4)) =
                make_uint4(pk[c][u * 4], pk[c][u * 4 + 1], pk[c][u * 4 + 2], pk[c][u * 4 + 3]);
            *reinterpret_cast<uint4*>(srow + (unit << 4)) =
                make_uint4(dk[c][u * 4], dk[c][u * 4 + 1], dk[c][u * 4 + 2], dk[c][u * 4 + 3]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar_p);
        // deferred by one iteration: what the previous MMA 2 produced leaves while this one runs
        if (it > 0) {
          reduce_dq((it - 1) & 1, pb, ph, pqt);
          tc_fence_before();
          mbar_arrive(&bar_dqfree[(it - 1) & 1]);
          if (pend_kv) {
            store_kv_row(kv_b, kv_h, kv_kb);
            tc_fence_before();
            mbar_arrive(bar_kvdone);
            pend_kv = false;
          }
        }
        pb = b; ph = h; pqt = qt;
        if (qt == nqt - 1) {
          pend_kv = true;
          kv_b = b; kv_h = h; kv_kb = kb;
        }
      }
    }
    if (it > 0) {
      mbar_wait(bar_m2, (it - 1) & 1);
      tc_fence_after();
      reduce_dq((it - 1) & 1, pb, ph, pqt);
      if (pend_kv) store_kv_row(kv_b, kv_h, kv_kb);
    }
    if (stg_pending && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// dqkv[b, n, 0, h, :] = bf16(dq32[b, n, h, :]) — the query-gradient slot, after the reduce-adds
__global__ void __launch_bounds__(256)
attn_dq_convert_kernel(const float* __restrict__ dq32, __nv_bfloat16* __restrict__ dqkv, long long rows, int hd) {
  pdl_wait();
  const int vec = hd >> 3;
  const long long total = rows * vec;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / vec;
    const int c = static_cast<int>(i - row * vec) << 3;
    const float4 a = __ldcs(reinterpret_cast<const float4*>(dq32 + row * hd + c));
    const float4 b = __ldcs(reinterpret_cast<const float4*>(dq32 + row * hd + c) + 1);
    uint4 w;
    w.x = pack_bf16(a.x, a.y);
    w.y = pack_bf16(a.z, a.w);
    w.z = pack_bf16(b.x, b.y);
    w.w = pack_bf16(b.z, b.w);
    *reinterpret_cast<uint4*>(dqkv + row * 3 * hd + c) = w;
  }
}

int64_t attention_tc_bwd_long_workspace(int64_t batch, int64_t tokens, int64_t heads) {
  return batch * tokens * heads * 64 * static_cast<int64_t>(sizeof(float));
}

int attention_tc_bwd_long(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv,
                          void* workspace, int64_t batch, int64_t tokens, int64_t heads, float scale,
                          cudaStream_t stream) {
  AttnLongBwdParams p;
  p.N = static_cast<int>(tokens);
  p.H = static_cast<int>(heads);
  p.kw = static_cast<int>((tokens + 15) / 16 * 16);
  p.nkb = (p.kw + 127) / 128;
  p.nqt = (p.N + 127) / 128;
  p.items = static_cast<int>(batch * heads * p.nkb);
  p.scale = scale;
  p.lse = lse;
  p.delta = delta;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  const int64_t hd = heads * 64;
  float* dq32 = reinterpret_cast<float*>(workspace);
  FV_CHECK_CUDA(cudaMemsetAsync(dq32, 0, attention_tc_bwd_long_workspace(batch, tokens, heads), stream));
  CUtensorMap mq, mdo, mdq;
  int rc = make_tok_map(&mq, qkv, batch, tokens, 3 * hd);
  if (rc != FV_OK) return rc;
  rc = make_tok_map(&mdo, dout, batch, tokens, hd);
  if (rc != FV_OK) return rc;
  {  // fp32 scratch [B, N, H*64]: box = 32 columns (128 B) x 32 tokens x 1 image, 128B swizzle
    EncodeTiledFn enc = get_encode_tiled();
    FV_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(hd), static_cast<cuuint64_t>(tokens), static_cast<cuuint64_t>(batch)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(hd) * 4, static_cast<cuuint64_t>(hd) * tokens * 4};
    cuuint32_t box[3] = {32, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&mdq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dq32, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(dq scratch) failed (%d)", static_cast<int>(r));
      return FV_ERR_CUDA;
    }
  }
  const int smem = 12 * ATB_TILE + 8 * 4096 + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    FV_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int grid = p.items < num_sms() ? p.items : num_sms();
  FV_CHECK_CUDA(fv::launch_pdl(attn_tc_bwd_long_kernel, dim3(static_cast<unsigned>(grid)), dim3(ATLB_THREADS), smem, stream,
                               mq, mdo, mdq, p));
  count_kernel(FV_KERNEL_ATTN_BWD_LONG);
  FV_LAUNCH_CHECK();
  const long long rows = batch * tokens;
  const long long total = rows * (hd / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  FV_CHECK_CUDA(fv::launch_pdl(attn_dq_convert_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream,
                               static_cast<const float*>(dq32), p.dqkv, rows, static_cast<int>(hd)));
  FV_LAUNCH_CHECK();
  return FV_OK;
}

}  // namespace fv

#ifdef ATC_TRACE
extern "C" int fv_debug_read_trace_bwd(long long* dst) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(dst, fv::g_atb_trace, sizeof(fv::g_atb_trace)) == cudaSuccess ? 0 : -2;
}
extern "C" int fv_debug_read_trace(long long* dst, int64_t n) {
  cudaDeviceSynchronize();
  const size_t bytes = static_cast<size_t>(n) * sizeof(long long);
  return cudaMemcpyFromSymbol(dst, fv::g_atc_trace, bytes < sizeof(fv::g_atc_trace) ? bytes : sizeof(fv::g_atc_trace)) ==
                 cudaSuccess
             ? 0
             : -2;
}
#endif

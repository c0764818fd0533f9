// attention_fwd2.cu — attention forward for the short ViT sequences (N <= 256: 197 at 224 px), head_dim 64,
// second generation: TWO THREADS PER QUERY ROW.
//
// Replaces F.scaled_dot_product_attention inside timm's Attention.forward (reference call site model.py:193).
//
// The first kernel (attention_tc.cu: attn_tc_fwd_kernel) is bound by its softmax warps, not by the tensor
// core: ncu (profiles/r1_ncu_full_attn_fwd.csv) has the tensor pipe 20 % and the MUFU (ex2) pipe 42 % busy —
// a thread owns a whole query row of 208 scores, walks it twice (row maximum, then exponentials) in seven
// dependent tensor-memory round trips per pass, and with four softmax warps per CTA (two CTAs per SM) every
// scheduler has only two of them to hide those round trips with. Here a row is shared by two threads (same
// TMEM lanes, warps w and w + 4): each walks half of the keys, the halves meet through two shared-memory
// scalars per row (maximum, sum). The per-tile dependency chain is half as long and every scheduler has four
// softmax warps; the MMAs, the TMA producer and the memory traffic are unchanged.
//
//   TMA      Q [128 x 64], K [kw x 64], V [kw x 64] out of timm's [B,N,3,H,64] qkv layout (3-D tensor map)
//   MMA 1    S = Q K^T           tcgen05.mma 128 x kw x 64 -> TMEM columns [0, kw)
//   softmax  thread (row, half): half 0 owns keys [0, kA), half 1 keys [kA, kw), kA = 16 * ceil(kw / 32);
//            P (bf16 pairs) is written back over the scores the SAME thread has already consumed:
//            half 0 -> columns [0, kA/2), half 1 -> columns [kA, kA + (kw-kA)/2) — never another thread's
//   MMA 2    O = P V             A operand from TMEM (two column runs), V MN-major -> TMEM columns [192, 256)
//   epilogue each thread normalises and stores 32 of the row's 64 output columns; LSE saved by half 0.
#include "common.cuh"

namespace fv {

int make_qkv_map(CUtensorMap* map, const void* base, int64_t batch, int64_t tokens, int64_t width, int rows);

namespace {

constexpr int F2_THREADS = 320;  // warps 0-7 softmax / epilogue, warp 8 MMA issue, warp 9 TMEM alloc + TMA producer
constexpr int F2_Q = 128;
constexpr int F2_KV_MAX = 256;
constexpr int F2_SMEM = 2 * F2_Q * 128 + 2 * F2_KV_MAX * 128 + 1024 + 2 * 2 * 128 * 4 + 128;
constexpr float F2_LOG2E = 1.4426950408889634f;
constexpr uint32_t F2_T_O = 192;

struct Fwd2Params {
  int N, H, kw;  // tokens, heads, keys rounded up to 16
  int kA;        // keys of half 0 (multiple of 16)
  int items;     // batch * heads
  float scale;
  __nv_bfloat16* out;
  float* lse;
};

__device__ __forceinline__ float f2_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// the two warps that share a TMEM lane quarter (w and w + 4): 64 threads, named barrier 1 + quarter
__device__ __forceinline__ void pair_sync(int quarter) {
  asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
}

__global__ void __launch_bounds__(F2_THREADS, 2)
attn_tc_fwd2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const Fwd2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                         // 2 x 16 KiB (both query tiles)
  uint8_t* sK = sQ + 2 * F2_Q * 128;          // 32 KiB
  uint8_t* sV = sK + F2_KV_MAX * 128;         // 32 KiB
  float* sMax = reinterpret_cast<float*>(sV + F2_KV_MAX * 128);  // [2 halves][128 rows]
  float* sSum = sMax + 2 * 128;                                  // [2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSum + 2 * 128);
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
  const int nqt = (p.N + F2_Q - 1) / F2_Q;  // 1 or 2 query tiles, processed back to back
  const int hd = p.H * 64;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 256);
    mbar_init(bar_o, 1);
    mbar_init(bar_done, 256);
    mbar_init(bar_qkfree, 1);
    mbar_init(bar_vfree, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();  // nothing above touches memory another kernel produced

  if (warp == 9) {
    // ------------------------------ TMA producer (whole warp, elected lane issues) --------------
    int n = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      const int b = item / p.H, h = item % p.H;
      if (n > 0) mbar_wait(bar_qkfree, (n - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(bar_qk, (nqt * F2_Q + p.kw) * 128);
        for (int t = 0; t < nqt; ++t) tma_load_3d(sQ + t * F2_Q * 128, &tmap_q, bar_qk, h * 64, t * F2_Q, b);
        tma_load_3d(sK, &tmap_kv, bar_qk, hd + h * 64, 0, b);
      }
      __syncwarp();
      if (n > 0) mbar_wait(bar_vfree, (n - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(bar_v, p.kw * 128);
        tma_load_3d(sV, &tmap_kv, bar_v, 2 * hd + h * 64, 0, b);
      }
      __syncwarp();
    }
  } else if (warp == 8) {
    // ------------------------------ MMA issue (whole warp, elected lane issues) -----------------
    const uint32_t idesc_s = make_idesc(kFmtBF16, 0, 0, F2_Q, p.kw);
    const uint32_t idesc_o = make_idesc(kFmtBF16, 0, 1, F2_Q, 64);
    const uint64_t dk = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
    const uint64_t dv = make_smem_desc_sw128(smem_u32(sV), 64 * 128, 1024);
    const int ksteps = p.kw >> 4, kstepsA = p.kA >> 4;
    int n = 0, gt = 0;  // items, query tiles so far
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      mbar_wait(bar_qk, n & 1);
      for (int t = 0; t < nqt; ++t, ++gt) {
        if (gt > 0) mbar_wait(bar_done, (gt - 1) & 1);  // previous tile's O has been read out of TMEM
        tc_fence_after();
        // S = Q K^T : both operands K-major (head dim contiguous), 4 steps of K = 16
        const uint64_t dq = make_smem_desc_sw128(smem_u32(sQ + t * F2_Q * 128), 16, 1024);
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
        // O = P V : A = P from TMEM (16 keys = 8 packed columns per step; half 0's run starts at column 0,
        // half 1's at column kA), B = V MN-major
        if (elect_one()) {
          for (int k = 0; k < ksteps; ++k) {
            const uint32_t pa = k < kstepsA ? static_cast<uint32_t>(k * 8)
                                            : static_cast<uint32_t>(p.kA + (k - kstepsA) * 8);
            umma_bf16_ts(tmem + F2_T_O, tmem + pa, dv + k * (2048 >> 4), idesc_o, k > 0 ? 1u : 0u);
          }
          umma_commit(bar_o);
          if (t == nqt - 1) umma_commit(bar_vfree);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ softmax + epilogue: thread = (query row, key half) -----------
    const int quarter = warp & 3, hf = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t taddr = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sl2 = p.scale * F2_LOG2E;
    const int k0 = hf == 0 ? 0 : p.kA;          // this half's first key / first score column
    const int k1 = hf == 0 ? p.kA : p.kw;       // one past its last (multiples of 16)
    const int nchunks = (k1 - k0 + 31) >> 5;    // 32-column TMEM reads; the last one may be half valid
    int gt = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int b = item / p.H, h = item % p.H;
      for (int t = 0; t < nqt; ++t, ++gt) {
        const uint32_t ph = gt & 1;
        const int q0 = t * F2_Q;
        const int q = q0 + row;
        const bool warp_live = q0 + quarter * 32 < p.N;  // warp-uniform (both warps of the pair agree)
        float mx = -INFINITY, sum = 0.f;
        mbar_wait(bar_s, ph);
        tc_fence_after();
        if (warp_live) {
          // pass 1: maximum over this half's keys (the other three softmax warps of this scheduler cover the
          // tensor-memory round trip; a second buffer does not fit the 102-register budget of 2 x 320 threads)
          for (int c = 0; c < nchunks; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + k0 + c * 32, r);
            tmem_ld_wait();
            const int key0 = k0 + c * 32;
            if (key0 + 32 <= k1 && key0 + 32 <= p.N) {  // whole chunk inside this half and the sequence
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (key0 + i < k1 && key0 + i < p.N) mx = fmaxf(mx, __uint_as_float(r[i]));
            }
          }
          sMax[hf * 128 + row] = mx;
        }
        pair_sync(quarter);
        if (warp_live) {
          mx = fmaxf(sMax[row], sMax[128 + row]);
          const float mxs = mx * sl2;
          // pass 2: exponentials of this half's keys, packed back over the columns just read
          for (int c = 0; c < nchunks; ++c) {
            uint32_t r[32], pk[16];
            tmem_ld_32x32(taddr + k0 + c * 32, r);
            tmem_ld_wait();
            const int key0 = k0 + c * 32;
            if (key0 + 32 <= k1 && key0 + 32 <= p.N) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float p0 = f2_ex2(fmaf(__uint_as_float(r[i]), sl2, -mxs));
                const float p1 = f2_ex2(fmaf(__uint_as_float(r[i + 1]), sl2, -mxs));
                sum += p0 + p1;  // fp32 row sum (the saved LSE is the exact log-sum-exp)
                pk[i >> 1] = pack_bf16(p0, p1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const bool v0 = key0 + i < k1 && key0 + i < p.N, v1 = key0 + i + 1 < k1 && key0 + i + 1 < p.N;
                const float p0 = v0 ? f2_ex2(fmaf(__uint_as_float(r[i]), sl2, -mxs)) : 0.f;
                const float p1 = v1 ? f2_ex2(fmaf(__uint_as_float(r[i + 1]), sl2, -mxs)) : 0.f;
                sum += p0 + p1;
                pk[i >> 1] = pack_bf16(p0, p1);
              }
            }
            // 16 packed columns at k0 + 16 c: inside the score columns this thread has consumed. A chunk
            // that is only half inside this half (k1 - key0 == 16) writes 8 columns: the other 8 would land
            // in the NEXT half's score columns
            if (key0 + 32 <= k1) {
              tmem_st_32x16(taddr + k0 + c * 16, pk);
            } else {
              asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(
                               taddr + k0 + c * 16),
                           "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                           : "memory");
            }
          }
          tmem_st_wait();
          sSum[hf * 128 + row] = sum;
        }
        tc_fence_before();
        mbar_arrive(bar_p);

        mbar_wait(bar_o, ph);
        tc_fence_after();
        uint32_t o[32];
        if (warp_live) {
          tmem_ld_32x32(taddr + F2_T_O + hf * 32, o);
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive(bar_done);  // TMEM may be overwritten by the next tile's S
        pair_sync(quarter);     // the partner's row sum is visible (written before its bar_p arrive, long ago)
        if (warp_live && q < p.N) {
          // this thread's half of the row: 32 bf16 = 64 bytes of out[b, q, h, :], two 256-bit stores
          const float tot = sSum[row] + sSum[128 + row];
          const float inv = 1.0f / tot;
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i)
            w[i] = pack_bf16(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv);
          __nv_bfloat16* dst = p.out + ((static_cast<long long>(b) * p.N + q) * p.H + h) * 64 + hf * 32;
          st_v8(dst, w);
          st_v8(dst + 16, w + 8);
          if (hf == 0) p.lse[(static_cast<long long>(b) * p.H + h) * p.N + q] = mx * p.scale + logf(tot);
        }
        // sMax / sSum of this tile are read (above) before either warp of the pair can overwrite them:
        // the next write happens after the next bar_s wait and the pair barrier orders this read first
        pair_sync(quarter);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

}  // namespace

int attention_tc_fwd2(const void* qkv, void* out, float* lse, int64_t batch, int64_t tokens, int64_t heads,
                      float scale, cudaStream_t stream) {
  Fwd2Params p;
  p.N = static_cast<int>(tokens);
  p.H = static_cast<int>(heads);
  p.kw = static_cast<int>((tokens + 15) / 16 * 16);
  p.kA = (p.kw + 31) / 32 * 16;
  p.items = static_cast<int>(batch * heads);
  p.scale = scale;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  CUtensorMap mq, mkv;
  int rc = make_qkv_map(&mq, qkv, batch, tokens, 3 * heads * 64, F2_Q);
  if (rc != FV_OK) return rc;
  rc = make_qkv_map(&mkv, qkv, batch, tokens, 3 * heads * 64, p.kw);
  if (rc != FV_OK) return rc;
  static bool configured = false;
  if (!configured) {
    FV_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM));
    configured = true;
  }
  const int slots = 2 * num_sms();  // two co-resident CTAs per SM
  const unsigned grid = static_cast<unsigned>(p.items < slots ? p.items : slots);
  FV_CHECK_CUDA(fv::launch_pdl(attn_tc_fwd2_kernel, dim3(grid), dim3(F2_THREADS), F2_SMEM, stream, mq, mkv, p));
  count_kernel(FV_KERNEL_ATTN_FWD);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

}  // namespace fv

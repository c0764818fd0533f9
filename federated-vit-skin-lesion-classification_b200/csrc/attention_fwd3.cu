// attention_fwd3.cu — attention forward for the short ViT sequences (N <= 208 tokens: 197 at 224 px), head_dim 64,
// third generation: the per-tile dependency chain of the first kernel, measured and taken apart.
//
// Replaces F.scaled_dot_product_attention inside timm's Attention.forward (reference call site model.py:193).
//
// SM-clock timestamps of the first kernel (attention_tc.cu: attn_tc_fwd_kernel, tools/attn_trace.py,
// profiles/r2_attn_trace.txt) show one CTA walking a strictly serial chain per 128-query tile — score MMA 1000
// cycles, row-maximum pass 1350, exponential pass 3400, PV MMA 950, read-out 300: 7200 cycles of which the MUFU
// pipe (the unit that bounds the op: 8 cycles per warp-wide ex2) needs 1800 — and two co-resident CTAs that only
// half hide each other's gaps. What changes here, with the tile shape, the operand layouts and the softmax
// arithmetic kept (bit-identical P and O):
//   * the score accumulator sits at TMEM columns [48, 48 + kw) with key j in column 48 + j, O at [0, 64), P over
//     consumed scores from column 64 on. The scores of keys >= 16 (columns >= 64) do not overlap O, so the MMA
//     warp issues the NEXT tile's main score MMA right behind this tile's PV MMA — before the softmax warps have
//     read O out — and only the 16-key sliver that shares columns [48, 64) with O waits for the read-out. The
//     softmax warps never wait for a score MMA again.
//   * both softmax passes are software-pipelined: the tensor-memory load of chunk c + 1 is in flight while
//     chunk c is being reduced / exponentiated (two 32-register buffers), the row maximum uses the 3-input
//     FMNMX3, and the 16-key sliver is kept in registers between the passes.
//   * the PV MMA starts on the first half of the keys while the second half is still being exponentiated
//     (two P barriers), so only half of it is left on the chain after the softmax.
//   * the ragged last query tile (69 of 128 rows at N = 197: one idle and one nearly idle softmax warp, always
//     the same two schedulers) is rotated by (item + CTA) mod 4 warps: the row -> TMEM-lane assignment of that
//     tile is cyclic, so the exponential work is spread evenly over the four schedulers' MUFU pipes.
#include "common.cuh"

namespace fv {

namespace {

#ifndef F3_ROT
#define F3_ROT 1
#endif
#ifndef F3_SPLIT_PV
#define F3_SPLIT_PV 1
#endif

constexpr int F3_THREADS = 192;  // warps 0-3 softmax / epilogue, warp 4 MMA issue, warp 5 TMEM alloc + TMA producer
constexpr int F3_Q = 128;
constexpr int F3_KV_MAX = 208;
constexpr int F3_SMEM = 2 * F3_Q * 128 + 2 * F3_KV_MAX * 128 + 1024 + 128;
constexpr float F3_LOG2E = 1.4426950408889634f;
constexpr uint32_t F3_T_O = 0;    // O accumulator: columns [0, 64)
constexpr uint32_t F3_T_S = 48;   // score of key j: column 48 + j
constexpr uint32_t F3_T_SM = 64;  // = F3_T_S + 16: the main keys' scores (keys >= 16) start here, clear of O
constexpr uint32_t F3_T_P = 64;   // packed bf16 P of the main keys from here on (over consumed scores)

struct Fwd3Params {
  int N, H, kw;  // tokens, heads, keys rounded up to 16 (<= 208)
  int items;     // batch * heads
  float scale;
  __nv_bfloat16* out;
  float* lse;
};

#ifdef ATC_TRACE
// measurement build (tools/build_variants.py attention_fwd3.cu trace:-DATC_TRACE, tools/attn_trace.py): SM-clock
// timestamps of the first tiles of four CTAs — [cta][tile][event]; events 0-4 softmax warp 0 (main scores ready,
// pass 1 done, P written, O ready, row stored), 5-6 MMA warp (score issue, last PV issue), 7 = smid
__device__ long long g_f3_trace[4][32][8];
__device__ __forceinline__ void f3_stamp(int gt, int ev) {
  const int c = blockIdx.x == 0 ? 0 : blockIdx.x == 1 ? 1 : blockIdx.x == 148 ? 2 : blockIdx.x == 149 ? 3 : -1;
  if (c >= 0 && gt < 32) {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    g_f3_trace[c][gt][ev] = t;
    if (ev == 0) {
      uint32_t sm;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      g_f3_trace[c][gt][7] = sm;
    }
  }
}
#define F3_STAMP(gt, ev) do { if (lane == 0) f3_stamp(gt, ev); } while (0)
#else
#define F3_STAMP(gt, ev) do { } while (0)
#endif

__device__ __forceinline__ float f3_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float f3_max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// TMEM -> registers, 16 columns (the 16-key sliver and a 16-wide main tail); fills r[0..16)
__device__ __forceinline__ void f3_ld16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void f3_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// wait for the outstanding tensor-memory loads; the buffer is an in/out operand so that no use of it can be
// scheduled above the wait (the load itself only names the registers, the data lands asynchronously)
__device__ __forceinline__ void f3_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                 "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                 "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// running maximum over a 32-wide chunk (or the 16 entries of a 16-wide one), every entry inside the sequence
template <int W>
__device__ __forceinline__ float f3_max_chunk(const uint32_t (&r)[32], float mx) {
#pragma unroll
  for (int i = 0; i < W; i += 2) mx = f3_max3(mx, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
  return mx;
}
// exponentials of the first W entries -> W / 2 packed bf16 pairs, row sum accumulated in fp32
template <int W>
__device__ __forceinline__ void f3_exp_chunk(const uint32_t (&r)[32], float sl2, float mxs, float& sum, uint32_t* pk) {
#pragma unroll
  for (int i = 0; i < W; i += 2) {
    const float p0 = f3_ex2(fmaf(__uint_as_float(r[i]), sl2, -mxs));
    const float p1 = f3_ex2(fmaf(__uint_as_float(r[i + 1]), sl2, -mxs));
    sum += p0 + p1;  // fp32 row sum (the saved LSE is the exact log-sum-exp)
    pk[i >> 1] = pack_bf16(p0, p1);
  }
}

__global__ void __launch_bounds__(F3_THREADS, 2)
attn_tc_fwd3_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const Fwd3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                       // 2 x 16 KiB (both query tiles)
  uint8_t* sK = sQ + 2 * F3_Q * 128;        // 26 KiB
  uint8_t* sV = sK + F3_KV_MAX * 128;       // 26 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + F3_KV_MAX * 128);
  uint64_t* bar_qk = bars + 0;      // Q tiles + K landed
  uint64_t* bar_v = bars + 1;       // V landed
  uint64_t* bar_sm = bars + 2;      // main scores (keys >= 16) of the tile in TMEM
  uint64_t* bar_sf = bars + 3;      // first 16 keys' scores in TMEM
  uint64_t* bar_p0 = bars + 4;      // P of the first half of the main keys written
  uint64_t* bar_p = bars + 5;       // all of P written
  uint64_t* bar_o = bars + 6;       // O accumulated
  uint64_t* bar_done = bars + 7;    // O read out of TMEM
  uint64_t* bar_qkfree = bars + 8;  // the item's last score MMA has retired: Q tiles and K may be refilled
  uint64_t* bar_vfree = bars + 9;   // the item's last PV MMA has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = (p.N + F3_Q - 1) / F3_Q;  // 1 or 2 query tiles per item
  const int hd = p.H * 64;
  const int km = p.kw - 16;                 // main keys [0, km): all inside the sequence; sliver = keys [km, kw)
  const int nm32 = km >> 5;                 // 32-wide main chunks
  const bool tail16 = (km & 16) != 0;       // + a 16-wide one
  const int cA = F3_SPLIT_PV ? nm32 >> 1 : 0;  // chunks [0, cA) make the first PV batch
  const int kA = cA * 32;                   // keys in it

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_sm, 1);
    mbar_init(bar_sf, 1);
    mbar_init(bar_p0, 128);
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
      const int rot = F3_ROT ? (n + blockIdx.x) & 3 : 0;
      if (n > 0) mbar_wait(bar_qkfree, (n - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(bar_qk, (nqt * F3_Q + p.kw) * 128);
        for (int t = 0; t < nqt; ++t) {
          // 32-row boxes: TMEM lane quarter j of the tile holds query rows t*128 + 32*((j - r) mod 4)
          const int r = t == nqt - 1 ? rot : 0;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            tma_load_3d(sQ + (t * F3_Q + j * 32) * 128, &tmap_q, bar_qk, h * 64, t * F3_Q + 32 * ((j - r) & 3), b);
        }
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
  } else if (warp == 4) {
    // ------------------------------ MMA issue (whole warp, elected lane issues) -----------------
    const uint32_t idesc_sm = make_idesc(kFmtBF16, 0, 0, F3_Q, km > 0 ? km : 16);
    const uint32_t idesc_sf = make_idesc(kFmtBF16, 0, 0, F3_Q, 16);
    const uint32_t idesc_o = make_idesc(kFmtBF16, 0, 1, F3_Q, 64);
    const uint64_t dk = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
    const uint64_t dv = make_smem_desc_sw128(smem_u32(sV), 64 * 128, 1024);
    const int my_items = p.items > static_cast<int>(blockIdx.x) ? (p.items - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int ntiles = my_items * nqt;
    // score MMAs of tile g: the main keys first (they do not touch O of tile g - 1), the 16-key sliver once
    // tile g - 1's O has been read out
    auto issue_scores = [&](int g) {
      const int t = g % nqt;
      if (t == 0) mbar_wait(bar_qk, (g / nqt) & 1);
      tc_fence_after();
      const uint64_t dq = make_smem_desc_sw128(smem_u32(sQ + t * F3_Q * 128), 16, 1024);
      F3_STAMP(g, 5);
      if (elect_one()) {
        if (km > 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + F3_T_SM, dq + k * 2, dk + k * 2, idesc_sm, k > 0 ? 1u : 0u);
        }
        umma_commit(bar_sm);
      }
      __syncwarp();
      if (g > 0) mbar_wait(bar_done, (g - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + F3_T_S, dq + k * 2, dk + km * 8 + k * 2, idesc_sf, k > 0 ? 1u : 0u);
        umma_commit(bar_sf);
        if (t == nqt - 1) umma_commit(bar_qkfree);
      }
      __syncwarp();
    };
    if (ntiles > 0) issue_scores(0);
    for (int g = 0; g < ntiles; ++g) {
      const int t = g % nqt;
      // O = P V : A = P from TMEM (16 keys = 8 packed columns per step), B = V MN-major (16 rows = 2 KiB per step)
      mbar_wait(bar_p0, g & 1);
      if (t == 0) mbar_wait(bar_v, (g / nqt) & 1);
      tc_fence_after();
      if (elect_one()) {
        for (int k = 0; k < (kA >> 4); ++k)
          umma_bf16_ts(tmem + F3_T_O, tmem + F3_T_P + k * 8, dv + 16 * k * 8, idesc_o, k > 0 ? 1u : 0u);
      }
      __syncwarp();
      mbar_wait(bar_p, g & 1);
      tc_fence_after();
      F3_STAMP(g, 6);
      if (elect_one()) {
        for (int k = kA >> 4; k < (km >> 4); ++k)
          umma_bf16_ts(tmem + F3_T_O, tmem + F3_T_P + k * 8, dv + 16 * k * 8, idesc_o, k > 0 ? 1u : 0u);
        umma_bf16_ts(tmem + F3_T_O, tmem + F3_T_P + (km >> 1), dv + km * 8, idesc_o, km > 0 ? 1u : 0u);  // the sliver
        umma_commit(bar_o);
        if (t == nqt - 1) umma_commit(bar_vfree);
      }
      __syncwarp();
      if (g + 1 < ntiles) issue_scores(g + 1);
    }
  } else if (warp < 4) {
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const float sl2 = p.scale * F3_LOG2E;
    int gt = 0, n = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      const int b = item / p.H, h = item % p.H;
      const int rot = F3_ROT ? (n + blockIdx.x) & 3 : 0;
      for (int t = 0; t < nqt; ++t, ++gt) {
        const uint32_t ph = gt & 1;
        const int rb = t == nqt - 1 ? (warp - rot) & 3 : warp;  // this warp's 32-row block of the tile
        const int q = t * F3_Q + rb * 32 + lane;
        const bool warp_live = t * F3_Q + rb * 32 < p.N;  // warp-uniform
        float mx = -INFINITY, sum = 0.f;
        uint32_t pkf[8];  // P of keys [0, 16), parked in registers until the end of the pass
        mbar_wait(bar_sm, ph);
        tc_fence_after();
        if (warp == 0) F3_STAMP(gt, 0);
        if (warp_live) {
          uint32_t a[32], bq[32];
          // ---- pass 1: row maximum over the main chunks with the next load in flight, then the sliver
          if (nm32 > 0) tmem_ld_32x32(trow + F3_T_SM, a);
          for (int c = 0; c < nm32; c += 2) {
            f3_ld_wait(a);
            if (c + 1 < nm32) tmem_ld_32x32(trow + F3_T_SM + (c + 1) * 32, bq);
            mx = f3_max_chunk<32>(a, mx);
            if (c + 1 < nm32) {
              f3_ld_wait(bq);
              if (c + 2 < nm32) tmem_ld_32x32(trow + F3_T_SM + (c + 2) * 32, a);
              mx = f3_max_chunk<32>(bq, mx);
            }
          }
          if (tail16) {
            f3_ld16(trow + F3_T_SM + nm32 * 32, a);
            f3_ld_wait(a);
            mx = f3_max_chunk<16>(a, mx);
          }
          mbar_wait(bar_sf, ph);
          tc_fence_after();
          f3_ld16(trow + F3_T_S, bq);
          if (nm32 > 0) tmem_ld_32x32(trow + F3_T_SM, a);  // first chunk of pass 2 rides along
          f3_ld_wait(bq);
          const int live = p.N - km;  // 1 .. 16 keys of the sliver are inside the sequence
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < live) mx = fmaxf(mx, __uint_as_float(bq[i]));
          const float mxs = mx * sl2;
          if (warp == 0) F3_STAMP(gt, 1);
          // ---- pass 2: exponentials; P goes back into TMEM over scores this thread has already consumed
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float p0 = (i < live) ? f3_ex2(fmaf(__uint_as_float(bq[i]), sl2, -mxs)) : 0.f;
            const float p1 = (i + 1 < live) ? f3_ex2(fmaf(__uint_as_float(bq[i + 1]), sl2, -mxs)) : 0.f;
            sum += p0 + p1;
            pkf[i >> 1] = pack_bf16(p0, p1);
          }
          for (int c = 0; c < nm32; c += 2) {
            uint32_t pk[16];
            f3_ld_wait(a);
            if (c + 1 < nm32) tmem_ld_32x32(trow + F3_T_SM + (c + 1) * 32, bq);
            f3_exp_chunk<32>(a, sl2, mxs, sum, pk);
            tmem_st_32x16(trow + F3_T_P + c * 16, pk);
            if (c + 1 == cA) {
              tmem_st_wait();
              tc_fence_before();
              mbar_arrive(bar_p0);
            }
            if (c + 1 < nm32) {
              f3_ld_wait(bq);
              if (c + 2 < nm32) tmem_ld_32x32(trow + F3_T_SM + (c + 2) * 32, a);
              f3_exp_chunk<32>(bq, sl2, mxs, sum, pk);
              tmem_st_32x16(trow + F3_T_P + (c + 1) * 16, pk);
              if (c + 2 == cA) {
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(bar_p0);
              }
            }
          }
          if (tail16) {
            uint32_t pk[16];
            f3_ld16(trow + F3_T_SM + nm32 * 32, a);
            f3_ld_wait(a);
            f3_exp_chunk<16>(a, sl2, mxs, sum, pk);
            f3_st8(trow + F3_T_P + nm32 * 16, pk);
          }
          f3_st8(trow + F3_T_P + (km >> 1), pkf);
          tmem_st_wait();
        }
        tc_fence_before();
        if (!warp_live || cA == 0) mbar_arrive(bar_p0);
        mbar_arrive(bar_p);
        if (warp == 0) F3_STAMP(gt, 2);

        mbar_wait(bar_o, ph);
        tc_fence_after();
        if (warp == 0) F3_STAMP(gt, 3);
        uint32_t o0[32], o1[32];
        if (warp_live) {
          tmem_ld_32x32(trow + F3_T_O, o0);
          tmem_ld_32x32(trow + F3_T_O + 32, o1);
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive(bar_done);  // columns [48, 64) may take the next tile's 16-key sliver
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
        if (warp == 0) F3_STAMP(gt, 4);
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

}  // namespace

int make_qkv_map(CUtensorMap* map, const void* base, int64_t batch, int64_t tokens, int64_t width, int rows);

int attention_tc_fwd3(const void* qkv, void* out, float* lse, int64_t batch, int64_t tokens, int64_t heads,
                      float scale, cudaStream_t stream) {
  Fwd3Params p;
  p.N = static_cast<int>(tokens);
  p.H = static_cast<int>(heads);
  p.kw = static_cast<int>((tokens + 15) / 16 * 16);
  FV_CHECK_ARG(p.kw <= F3_KV_MAX, "attention_tc_fwd3: at most %d tokens", F3_KV_MAX);
  p.items = static_cast<int>(batch * heads);
  p.scale = scale;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  CUtensorMap mq, mkv;
  int rc = make_qkv_map(&mq, qkv, batch, tokens, 3 * heads * 64, 32);
  if (rc != FV_OK) return rc;
  rc = make_qkv_map(&mkv, qkv, batch, tokens, 3 * heads * 64, p.kw);
  if (rc != FV_OK) return rc;
  static bool configured = false;
  if (!configured) {
    FV_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_fwd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F3_SMEM));
    configured = true;
  }
  const int slots = 2 * num_sms();  // two co-resident CTAs per SM
  const unsigned grid = static_cast<unsigned>(p.items < slots ? p.items : slots);
  FV_CHECK_CUDA(fv::launch_pdl(attn_tc_fwd3_kernel, dim3(grid), dim3(F3_THREADS), F3_SMEM, stream, mq, mkv, p));
  count_kernel(FV_KERNEL_ATTN_FWD);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

}  // namespace fv

#ifdef ATC_TRACE
extern "C" int fv_debug_read_trace3(long long* dst, int64_t n) {
  cudaDeviceSynchronize();
  const size_t bytes = static_cast<size_t>(n) * sizeof(long long);
  return cudaMemcpyFromSymbol(dst, fv::g_f3_trace, bytes < sizeof(fv::g_f3_trace) ? bytes : sizeof(fv::g_f3_trace)) ==
                 cudaSuccess
             ? 0
             : -2;
}
#endif

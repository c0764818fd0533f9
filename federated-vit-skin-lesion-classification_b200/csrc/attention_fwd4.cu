// attention_fwd4.cu — attention forward for the short ViT sequences (N <= 208 tokens: 197 at 224 px), head_dim 64,
// fourth generation: ONE pass over the scores, softmax and read-out on different warps.
//
// Replaces F.scaled_dot_product_attention inside timm's Attention.forward (reference call site model.py:193).
//
// What the measurements behind it say (tools/ubench/tmem_mufu.cu, tools/attn_trace.py; profiles/r2_attn_trace.txt,
// profiles/r2_ubench_tmem_mufu.txt): the op is bound by the MUFU pipe — 8 cycles per warp-wide ex2 per scheduler,
// 1664 cycles for a 128 x 208 tile — and the earlier kernels keep that pipe busy 40 % of the time: per tile one
// CTA waits ~1000 cycles for the score MMA, reads the scores twice (row maximum 1350 cycles, exponentials 3400
// with two CTAs sharing the pipe), waits 950 for the PV MMA and spends 900 on the read-out, and the two
// co-resident CTAs fall into step instead of filling each other's gaps. Here, one CTA per SM:
//   * eight softmax warps do nothing but the exponential pass, two per tensor-memory lane quarter: the pair shares
//     the quarter's 32 rows and splits their keys (32-key chunks alternately), so each scheduler holds two
//     softmax warps whose load latencies, shift checks and stores hide behind each other's exponentials
//     (one warp per scheduler: 560 cycles per chunk where the MUFU pipe needs 256).
//   * the row maximum pass is gone: the shift of the softmax does not have to be the maximum, any m with
//     max - m <= 64 keeps 2^(x - m) far from overflow and leaves the result unchanged (P is rounded to bf16 and
//     accumulated in fp32: relative errors, no absolute ones). Both warps of a pair take m = the maximum of the
//     row's first 32 keys, rounded up to an integer; a later chunk whose maximum (16 three-input max
//     instructions in four chains, next to 32 exponentials) exceeds the warp's m by more than TAU = 64 — a 2^64
//     ratio between keys, not seen outside adversarial inputs — takes a slow path that raises that warp's m and
//     rescales its sum and the P it has written by the exact power of two. At the end of the tile the pair
//     compares shifts through shared memory (one 64-thread named barrier) and the warp with the smaller one
//     rescales once more, so the whole row ends on one m. The saved LSE is m ln2 + log(sum).
//   * two score accumulators in tensor memory: the MMA warp issues tile g + 2's QK^T right behind tile g's PV
//     MMA, so the scores of the next tile are waiting when the softmax warps finish a tile.
//   * four epilogue warps read O out of tensor memory, normalise and store it while the softmax warps are
//     already in the next tile; the row statistics cross through shared memory.
//   * Q, K, V of the next (batch, head) item land in a second shared-memory stage.
//   * the ragged last query tile (69 of 128 rows at N = 197) is rotated over the four lane quarters per item as
//     in the third-generation kernel, and the warps are decoupled (mbarriers only), so an idle pair moves on.
//   * the pass is kept small (one unmasked 32-wide body per register buffer, one 16-wide ragged tail): with
//     every chunk variant inlined at every call site it was 23 KB of straight-line code per loop iteration and
//     ran at 3.5 cycles per instruction.
#include "common.cuh"

namespace fv {

namespace {

constexpr int F4_SW = 8;                          // warps [0, 8) softmax (warp & 3 = lane quarter, warp >> 2 = key half),
                                                  // [F4_SW, F4_SW + 4) epilogue,
constexpr int F4_W_MMA = F4_SW + 4;               // then MMA issue,
constexpr int F4_W_TMA = F4_SW + 5;               // then TMEM alloc + TMA producer
constexpr int F4_THREADS = 32 * (F4_SW + 6);
constexpr int F4_Q = 128;
constexpr int F4_KV_MAX = 208;
constexpr int F4_STAGE = 2 * F4_Q * 128 + 2 * F4_KV_MAX * 128;  // Q tiles, K, V of one item: 84 KiB
constexpr int F4_SMEM = 2 * F4_STAGE + 4 * 256 * 8 + 2 * 512 * 4 + 256 + 1024;
constexpr float F4_LOG2E = 1.4426950408889634f;
constexpr float F4_LN2 = 0.6931471805599453f;
constexpr float F4_TAU = 64.f;  // a chunk maximum more than 2^TAU above the row's shift raises the shift
constexpr uint32_t F4_T_S0 = 0;    // score / P buffer of even tiles: columns [0, 208)
constexpr uint32_t F4_T_S1 = 224;  // odd tiles: [224, 432)
constexpr uint32_t F4_T_O = 448;   // O accumulator: [448, 512)

struct Fwd4Params {
  int N, H, kw;  // tokens, heads, keys rounded up to 16 (<= 208)
  int items;     // batch * heads
  float scale;
  __nv_bfloat16* out;
  float* lse;
};

#ifdef ATC_TRACE
// measurement build (tools/build_variants.py attention_fwd4.cu trace4:-DATC_TRACE, tools/attn_trace.py): SM-clock
// timestamps of the first tiles of four CTAs — [cta][tile][event]; 0-2 softmax warp 0 (scores ready, pass
// started = after the wait, P written), 3-4 epilogue warp 4 (O ready, row stored), 5-6 MMA warp (score issue, PV
// issue), 7 = smid
__device__ long long g_f4_trace[4][32][8];
__device__ __forceinline__ void f4_stamp(int gt, int ev) {
  const int c = blockIdx.x == 0 ? 0 : blockIdx.x == 1 ? 1 : blockIdx.x == 73 ? 2 : blockIdx.x == 147 ? 3 : -1;
  if (c >= 0 && gt < 32) {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    g_f4_trace[c][gt][ev] = t;
    if (ev == 5 && gt < 31) {
      uint32_t sm;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      g_f4_trace[c][gt][7] = sm;
    }
  }
}
#define F4_STAMP(gt, ev) do { if (lane == 0 && (gt) < 31) f4_stamp(gt, ev); } while (0)
// slot 31 of a traced CTA: SM clock and global timer (ns) at kernel entry (events 0, 2) and exit (1, 3)
__device__ __forceinline__ void f4_stamp_kernel(int at_exit) {
  if (threadIdx.x == 0) {
    f4_stamp(31, at_exit);
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const int c = blockIdx.x == 0 ? 0 : blockIdx.x == 1 ? 1 : blockIdx.x == 73 ? 2 : blockIdx.x == 147 ? 3 : -1;
    if (c >= 0) g_f4_trace[c][31][2 + at_exit] = t;
  }
}
#define F4_KSTAMP(e) f4_stamp_kernel(e)
// per-chunk timestamps of softmax warp 0 of CTA 0 on its tile 4: [chunk][after load wait, after check, after the
// next load's issue, after the exponentials, after the P store]
__device__ long long g_f4_chunk[8][5];
#define F4_CSTAMP(c, ev)                                                     \
  do {                                                                       \
    if (blockIdx.x == 0 && g == 4 && threadIdx.x == 0 && (c) < 8) {           \
      long long t_;                                                          \
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_));                     \
      g_f4_chunk[c][ev] = t_;                                                \
    }                                                                        \
  } while (0)
#else
#define F4_STAMP(gt, ev) do { } while (0)
#define F4_CSTAMP(c, ev) do { } while (0)
#define F4_KSTAMP(e) do { } while (0)
#endif

// volatile: keeps the exponentials between the tensor-memory load issued before them and the store after them
// (without it the compiler sinks the next chunk's load below the whole block and its latency is exposed)
__device__ __forceinline__ float f4_ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float f4_max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void f4_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void f4_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// wait for the outstanding tensor-memory loads; the buffer is an in/out operand so that no use of it can be
// scheduled above the wait (the load only names the registers, the data lands asynchronously)
__device__ __forceinline__ void f4_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                 "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                 "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// the softmax state of one row (one thread)
struct F4Row {
  float m;    // the shift, an integer in the log2 domain (x = score * scale * log2e)
  float sum;  // sum of 2^(x - m)
};

// Slow path: scale what a warp has produced for its rows so far — the sum and the packed P of its first `own8`
// groups of 16 keys, which sit at tp + 8 j — by the exact 2^(m - m_new) for the lanes with `raise` set.
// Returns (m', scaled sum).
__device__ __noinline__ float2 f4_rescale(uint32_t tp, int own8, bool raise, float m_new, float m, float sum) {
  m_new = raise ? m_new : m;
  const float r = f4_ex2(m - m_new);  // 1 for the lanes that keep their shift
  tmem_st_wait();                     // the P stores issued so far have to land before they are read back
  for (int j = 0; j < own8; ++j) {
    uint32_t pk[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(pk[0]), "=r"(pk[1]), "=r"(pk[2]), "=r"(pk[3]), "=r"(pk[4]), "=r"(pk[5]), "=r"(pk[6]), "=r"(pk[7])
                 : "r"(tp + j * 8)
                 : "memory");
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 v = unpack_bf16(pk[i]);
      pk[i] = pack_bf16(v.x * r, v.y * r);
    }
    f4_st8(tp + j * 8, pk);
  }
  tmem_st_wait();
  return make_float2(m_new, sum * r);
}

// maximum of W (16 or 32) scores, four independent FMNMX3 chains
template <int W>
__device__ __forceinline__ float f4_max(const uint32_t* r) {
  float c0 = f4_max3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
  float c1 = f4_max3(__uint_as_float(r[3]), __uint_as_float(r[4]), __uint_as_float(r[5]));
  float c2 = f4_max3(__uint_as_float(r[6]), __uint_as_float(r[7]), __uint_as_float(r[8]));
  float c3 = f4_max3(__uint_as_float(r[9]), __uint_as_float(r[10]), __uint_as_float(r[11]));
#pragma unroll
  for (int i = 12; i + 7 < W; i += 8) {
    c0 = f4_max3(c0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
    c1 = f4_max3(c1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
    c2 = f4_max3(c2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
    c3 = f4_max3(c3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
  }
  c0 = f4_max3(c0, __uint_as_float(r[W - 4]), __uint_as_float(r[W - 3]));
  c1 = f4_max3(c1, __uint_as_float(r[W - 2]), __uint_as_float(r[W - 1]));
  return fmaxf(fmaxf(c0, c1), fmaxf(c2, c3));
}
// shift check before a piece's exponentials (cmx = its maximum in the log2 domain; own8 = 16-key groups this warp
// has already written at tp); called with no tensor-memory load in flight — the slow path is a function call, and
// registers an asynchronous load still targets must not be live (or reused) across it
__device__ __forceinline__ void f4_check(float cmx, int own8, uint32_t tp, F4Row& row) {
  if (__any_sync(0xffffffffu, cmx > row.m + F4_TAU)) {
    const float2 ms = f4_rescale(tp, own8, cmx > row.m + F4_TAU, ceilf(cmx), row.m, row.sum);
    row.m = ms.x;
    row.sum = ms.y;
  }
}
// exponentials of W scores with the row's shift: packed bf16 P into pk[0 .. W/2), row sum
template <int W>
__device__ __forceinline__ void f4_exp(const uint32_t* r, float sl2, F4Row& row, uint32_t* pk) {
  const float ms = row.m;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < W; i += 2) {
    const float p0 = f4_ex2(fmaf(__uint_as_float(r[i]), sl2, -ms));
    const float p1 = f4_ex2(fmaf(__uint_as_float(r[i + 1]), sl2, -ms));
    s0 += p0;
    s1 += p1;
    pk[i >> 1] = pack_bf16(p0, p1);
  }
  row.sum += s0 + s1;
}
// maximum of the first nv (1 .. 16) of 16 scores: the last 16 columns of the key range hold the ragged end
__device__ __forceinline__ float f4_max_masked(const uint32_t* r, int nv) {
  float cm = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (i < nv) cm = fmaxf(cm, __uint_as_float(r[i]));
  return cm;
}
// ... and their exponentials: keys outside the sequence get P = 0
__device__ __forceinline__ void f4_exp_masked(const uint32_t* r, int nv, float sl2, F4Row& row, uint32_t* pk) {
  const float ms = row.m;
  float s0 = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const float p0 = i < nv ? f4_ex2(fmaf(__uint_as_float(r[i]), sl2, -ms)) : 0.f;
    const float p1 = i + 1 < nv ? f4_ex2(fmaf(__uint_as_float(r[i + 1]), sl2, -ms)) : 0.f;
    s0 += p0 + p1;
    pk[i >> 1] = pack_bf16(p0, p1);
  }
  row.sum += s0;
}
__device__ __forceinline__ void f4_pair_sync(int quarter) {  // the two softmax warps of a lane quarter
  asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
}

__global__ void __launch_bounds__(F4_THREADS, 1)
attn_tc_fwd4_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const Fwd4Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // stage s: Q tiles (2 x 16 KiB) | K (26 KiB) | V (26 KiB)
  float2* stats = reinterpret_cast<float2*>(smem + 2 * F4_STAGE);  // [4][2][128] (m, sum of the key half), tile g & 3
  float* xm = reinterpret_cast<float*>(smem + 2 * F4_STAGE + 4 * 256 * 8);  // [2][2][2][128] what the pair exchanges, tile g & 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * F4_STAGE + 4 * 256 * 8 + 2 * 512 * 4);
  uint64_t* bar_qk = bars + 0;      // [2] Q tiles + K of the stage landed
  uint64_t* bar_v = bars + 2;       // [2] V landed
  uint64_t* bar_qkfree = bars + 4;  // [2] the item's last score MMA has retired
  uint64_t* bar_vfree = bars + 6;   // [2] the item's last PV MMA has retired
  uint64_t* bar_s = bars + 8;       // [2] scores of the tile in tensor memory (buffer g & 1)
  uint64_t* bar_p = bars + 10;      // [2] P written, row statistics in shared memory
  uint64_t* bar_o = bars + 12;      // O accumulated
  uint64_t* bar_ofree = bars + 13;  // O read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  F4_KSTAMP(0);
  const int nqt = (p.N + F4_Q - 1) / F4_Q;  // 1 or 2 query tiles per item
  const int hd = p.H * 64;
  const int my_items = p.items > static_cast<int>(blockIdx.x) ? (p.items - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const int ntiles = my_items * nqt;

  if (warp == F4_W_MMA && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_qk + i, 1);
      mbar_init(bar_v + i, 1);
      mbar_init(bar_qkfree + i, 1);
      mbar_init(bar_vfree + i, 1);
      mbar_init(bar_s + i, 1);
      mbar_init(bar_p + i, 256);
    }
    mbar_init(bar_o, 1);
    mbar_init(bar_ofree, 128);
    fence_mbar_init();
  }
  if (warp == F4_W_TMA) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();  // nothing above touches memory another kernel produced

  if (warp == F4_W_TMA) {
    // ------------------------------ TMA producer (whole warp, elected lane issues) --------------
    for (int n = 0; n < my_items; ++n) {
      const int item = blockIdx.x + n * gridDim.x;
      const int b = item / p.H, h = item % p.H;
      const int s = n & 1;
      const uint32_t reuse = ((n >> 1) - 1) & 1;  // phase of the stage's previous tenant
      uint8_t* sQ = smem + s * F4_STAGE;
      uint8_t* sK = sQ + 2 * F4_Q * 128;
      uint8_t* sV = sK + F4_KV_MAX * 128;
      const int rot = (n + blockIdx.x) & 3;
      if (n >= 2) mbar_wait(bar_qkfree + s, reuse);
      if (elect_one()) {
        mbar_expect_tx(bar_qk + s, (nqt * F4_Q + p.kw) * 128);
        for (int t = 0; t < nqt; ++t) {
          // 32-row boxes: TMEM lane quarter j of the tile holds query rows t*128 + 32*((j - r) mod 4)
          const int r = t == nqt - 1 ? rot : 0;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            tma_load_3d(sQ + (t * F4_Q + j * 32) * 128, &tmap_q, bar_qk + s, h * 64, t * F4_Q + 32 * ((j - r) & 3), b);
        }
        tma_load_3d(sK, &tmap_kv, bar_qk + s, hd + h * 64, 0, b);
      }
      __syncwarp();
      if (n >= 2) mbar_wait(bar_vfree + s, reuse);
      if (elect_one()) {
        mbar_expect_tx(bar_v + s, p.kw * 128);
        tma_load_3d(sV, &tmap_kv, bar_v + s, 2 * hd + h * 64, 0, b);
      }
      __syncwarp();
    }
  } else if (warp == F4_W_MMA) {
    // ------------------------------ MMA issue (whole warp, elected lane issues) -----------------
    const uint32_t idesc_s = make_idesc(kFmtBF16, 0, 0, F4_Q, p.kw);
    const uint32_t idesc_o = make_idesc(kFmtBF16, 0, 1, F4_Q, 64);
    auto issue_scores = [&](int g) {
      const int n = g / nqt, t = g % nqt, s = n & 1;
      if (t == 0) mbar_wait(bar_qk + s, (n >> 1) & 1);
      tc_fence_after();
      const uint32_t sQ = smem_u32(smem + s * F4_STAGE);
      const uint64_t dq = make_smem_desc_sw128(sQ + t * F4_Q * 128, 16, 1024);
      const uint64_t dk = make_smem_desc_sw128(sQ + 2 * F4_Q * 128, 16, 1024);
      const uint32_t ts = tmem + ((g & 1) ? F4_T_S1 : F4_T_S0);
      F4_STAMP(g, 5);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(ts, dq + k * 2, dk + k * 2, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_s + (g & 1));
        if (t == nqt - 1) umma_commit(bar_qkfree + s);
      }
      __syncwarp();
    };
    if (ntiles > 0) issue_scores(0);
    if (ntiles > 1) issue_scores(1);
    for (int g = 0; g < ntiles; ++g) {
      const int n = g / nqt, t = g % nqt, s = n & 1;
      // O = P V : A = P from TMEM (16 keys = 8 packed columns per step), B = V MN-major (16 rows = 2 KiB per step)
      mbar_wait(bar_p + (g & 1), (g >> 1) & 1);
      if (t == 0) mbar_wait(bar_v + s, (n >> 1) & 1);
      if (g > 0) mbar_wait(bar_ofree, (g - 1) & 1);
      tc_fence_after();
      const uint64_t dv = make_smem_desc_sw128(smem_u32(smem + s * F4_STAGE) + (2 * F4_Q + F4_KV_MAX) * 128, 64 * 128, 1024);
      const uint32_t tp = tmem + ((g & 1) ? F4_T_S1 : F4_T_S0);
      F4_STAMP(g, 6);
      if (elect_one()) {
        // P of key group k: the first warp of a pair packs its groups from column 0, the second from column 16 g0
        const int g0 = p.kw >> 5;
        for (int k = 0; k < (p.kw >> 4); ++k)
          umma_bf16_ts(tmem + F4_T_O, tp + (k < g0 ? k * 8 : g0 * 16 + (k - g0) * 8), dv + 16 * k * 8, idesc_o, k > 0 ? 1u : 0u);
        umma_commit(bar_o);
        if (t == nqt - 1) umma_commit(bar_vfree + s);
      }
      __syncwarp();
      if (g + 2 < ntiles) issue_scores(g + 2);
    }
  } else if (warp < F4_SW) {
    // ------------------------------ softmax: one pass over the scores ---------------------------
    // The pair of a lane quarter splits the keys in the middle (16-key groups [0, g0) and [g0, G)); each warp
    // reads scores and writes P inside its own column range only — P of its group j over the first 8 columns of
    // what it has already read — so no warp ever overwrites scores the other one still needs.
    const int wq = warp & 3;     // TMEM lane quarter
    const int half = warp >> 2;
    const float sl2 = p.scale * F4_LOG2E;
    const int G = p.kw >> 4, g0 = G >> 1;
    const int whole = half ? G - g0 - 1 : g0;  // this warp's 16-key groups that lie inside the sequence entirely;
    const int nch = whole >> 1;                // as 32-wide chunks
    const bool odd16 = (whole & 1) != 0;       // + one 16-wide piece; warp 1 then has the ragged last group
    const int nv_last = p.N - (G - 1) * 16;    // keys of the sequence in it
    for (int g = 0; g < ntiles; ++g) {
      const int n = g / nqt, t = g % nqt;
      const int rot = (n + blockIdx.x) & 3;
      const int rb = t == nqt - 1 ? (wq - rot) & 3 : wq;  // this pair's 32-row block of the tile
      const bool warp_live = t * F4_Q + rb * 32 < p.N;     // warp-uniform, the same for both warps of the pair
      const uint32_t tp = tmem + (static_cast<uint32_t>(wq * 32) << 16) + ((g & 1) ? F4_T_S1 : F4_T_S0) + half * g0 * 16;
      if (warp == 0) F4_STAMP(g, 0);
      mbar_wait(bar_s + (g & 1), (g >> 1) & 1);
      tc_fence_after();
      if (warp == 0) F4_STAMP(g, 1);
      if (warp_live) {
        F4Row row;
        row.sum = 0.f;
        uint32_t a[32], bq[32];
        float* xrow = xm + (g & 1) * 512 + wq * 32 + lane;  // [tile parity][start / end][half][128]
        // ---- the pair's common shift: the maximum over both warps' first pieces, rounded up to an integer
        float first = -INFINITY;
        if (nch > 0) {
          tmem_ld_32x32(tp, a);
          f4_ld_wait(a);
          first = f4_max<32>(a);
        } else if (odd16 || half == 1) {
          f4_ld16(tp, a);
          f4_ld_wait(a);
          first = odd16 ? f4_max<16>(a) : f4_max_masked(a, nv_last);
        }
        xrow[half * 128] = first;
        f4_pair_sync(wq);
        row.m = ceilf(fmaxf(first, xrow[(half ^ 1) * 128]) * sl2);
        // ---- chunks through two register buffers, the load of the next one in flight during the exponentials
        for (int k = 0; k < nch; k += 2) {
          uint32_t pk[16];
          if (k > 0) {
            f4_ld_wait(a);
            F4_CSTAMP(k, 0);
            f4_check(f4_max<32>(a) * sl2, 2 * k, tp, row);
          }
          tmem_ld_32x32(tp + (k + 1) * 32, bq);  // unconditional (a branch here sinks the load below the exponentials)
          F4_CSTAMP(k, 1);
          f4_exp<32>(a, sl2, row, pk);
          F4_CSTAMP(k, 2);
          tmem_st_32x16(tp + k * 16, pk);
          F4_CSTAMP(k, 3);
          if (k + 1 < nch) {
            f4_ld_wait(bq);
            F4_CSTAMP(k + 1, 0);
            f4_check(f4_max<32>(bq) * sl2, 2 * k + 2, tp, row);
            tmem_ld_32x32(tp + (k + 2) * 32, a);
            F4_CSTAMP(k + 1, 1);
            f4_exp<32>(bq, sl2, row, pk);
            F4_CSTAMP(k + 1, 2);
            tmem_st_32x16(tp + (k + 1) * 16, pk);
            F4_CSTAMP(k + 1, 3);
          }
        }
        if (nch > 0) {
          // the look-ahead load of the loop's last iteration holds the columns after the chunks: the 16-wide
          // pieces below (it may reach 16 columns past this warp's range — read, never used)
          if (nch & 1) f4_ld_wait(bq);
          else f4_ld_wait(a);
        }
        {
          uint32_t pk[16];
          uint32_t tl[32];  // static register names: a select per element instead of a run-time choice of array
          const bool from_b = (nch & 1) != 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) tl[i] = from_b ? bq[i] : a[i];
          int own8 = 2 * nch;
          if (odd16) {
            if (nch > 0) f4_check(f4_max<16>(tl) * sl2, own8, tp, row);
            f4_exp<16>(tl, sl2, row, pk);
            f4_st8(tp + own8 * 8, pk);
            ++own8;
          }
          if (half == 1) {
            if (nch == 0 && odd16) {  // the first load was 16 wide: fetch the ragged group
              f4_ld16(tp + 16, tl + 16);
              f4_ld_wait(tl);
            }
            uint32_t tr[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) tr[i] = odd16 ? tl[16 + i] : tl[i];
            if (own8 > 0) f4_check(f4_max_masked(tr, nv_last) * sl2, own8, tp, row);
            f4_exp_masked(tr, nv_last, sl2, row, pk);
            f4_st8(tp + own8 * 8, pk);
            ++own8;
          }
          // ---- the pair settles on one shift per row (they differ only if one of the two raised its own)
          xrow[256 + half * 128] = row.m;
          f4_pair_sync(wq);
          const float m_star = fmaxf(row.m, xrow[256 + (half ^ 1) * 128]);
          if (__any_sync(0xffffffffu, m_star != row.m)) {
            const float2 ms = f4_rescale(tp, own8, m_star != row.m, m_star, row.m, row.sum);
            row.m = ms.x;
            row.sum = ms.y;
          }
        }
        stats[((g & 3) * 2 + half) * 128 + wq * 32 + lane] = make_float2(row.m, row.sum);
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(bar_p + (g & 1));
      if (warp == 0) F4_STAMP(g, 2);
    }
  } else {
    // ------------------------------ epilogue: O / rowsum -> out, LSE ----------------------------
    const int w4 = warp - F4_SW;
    const uint32_t to = tmem + (static_cast<uint32_t>(w4 * 32) << 16) + F4_T_O;
    for (int g = 0; g < ntiles; ++g) {
      const int n = g / nqt, t = g % nqt;
      const int item = blockIdx.x + n * gridDim.x;
      const int b = item / p.H, h = item % p.H;
      const int rot = (n + blockIdx.x) & 3;
      const int rb = t == nqt - 1 ? (w4 - rot) & 3 : w4;
      const int q = t * F4_Q + rb * 32 + lane;
      const bool warp_live = t * F4_Q + rb * 32 < p.N;
      mbar_wait(bar_o, g & 1);
      tc_fence_after();
      if (w4 == 0) F4_STAMP(g, 3);
      uint32_t o0[32], o1[32];
      if (warp_live) {
        tmem_ld_32x32(to, o0);
        tmem_ld_32x32(to + 32, o1);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(bar_ofree);
      if (warp_live && q < p.N) {
        // written before the row's arrival on bar_p, which the PV MMA behind bar_o waited for
        const float2 st0 = stats[((g & 3) * 2 + 0) * 128 + w4 * 32 + lane];
        const float2 st = make_float2(st0.x, st0.y + stats[((g & 3) * 2 + 1) * 128 + w4 * 32 + lane].y);
        const float inv = 1.0f / st.y;
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          w[i] = pack_bf16(__uint_as_float(o0[2 * i]) * inv, __uint_as_float(o0[2 * i + 1]) * inv);
          w[16 + i] = pack_bf16(__uint_as_float(o1[2 * i]) * inv, __uint_as_float(o1[2 * i + 1]) * inv);
        }
        // this thread's row: 64 bf16 = one 128-byte line of out[b, q, h, :], four 256-bit stores
        __nv_bfloat16* dst = p.out + ((static_cast<long long>(b) * p.N + q) * p.H + h) * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j) st_v8(dst + j * 16, w + j * 8);
        p.lse[(static_cast<long long>(b) * p.H + h) * p.N + q] = st.x * F4_LN2 + logf(st.y);
      }
      if (w4 == 0) F4_STAMP(g, 4);
    }
  }
  tc_fence_before();
  __syncthreads();
  F4_KSTAMP(1);
  if (warp == F4_W_TMA) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace

int make_qkv_map(CUtensorMap* map, const void* base, int64_t batch, int64_t tokens, int64_t width, int rows);

int attention_tc_fwd4(const void* qkv, void* out, float* lse, int64_t batch, int64_t tokens, int64_t heads,
                      float scale, cudaStream_t stream) {
  Fwd4Params p;
  p.N = static_cast<int>(tokens);
  p.H = static_cast<int>(heads);
  p.kw = static_cast<int>((tokens + 15) / 16 * 16);
  FV_CHECK_ARG(p.kw <= F4_KV_MAX, "attention_tc_fwd4: at most %d tokens", F4_KV_MAX);
  FV_CHECK_ARG(scale > 0.f, "attention_tc_fwd4: the score scale must be positive");
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
    FV_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_fwd4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F4_SMEM));
    configured = true;
  }
  const int slots = num_sms();  // one CTA per SM (all 512 tensor-memory columns)
  const unsigned grid = static_cast<unsigned>(p.items < slots ? p.items : slots);
  FV_CHECK_CUDA(fv::launch_pdl(attn_tc_fwd4_kernel, dim3(grid), dim3(F4_THREADS), F4_SMEM, stream, mq, mkv, p));
  count_kernel(FV_KERNEL_ATTN_FWD);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

}  // namespace fv

#ifdef ATC_TRACE
extern "C" int fv_debug_read_chunks4(long long* dst) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(dst, fv::g_f4_chunk, sizeof(fv::g_f4_chunk)) == cudaSuccess ? 0 : -2;
}
extern "C" int fv_debug_read_trace4(long long* dst, int64_t n) {
  cudaDeviceSynchronize();
  const size_t bytes = static_cast<size_t>(n) * sizeof(long long);
  return cudaMemcpyFromSymbol(dst, fv::g_f4_trace, bytes < sizeof(fv::g_f4_trace) ? bytes : sizeof(fv::g_f4_trace)) ==
                 cudaSuccess
             ? 0
             : -2;
}
#endif

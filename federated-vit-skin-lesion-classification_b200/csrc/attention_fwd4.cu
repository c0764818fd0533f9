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
//   * eight softmax warps do nothing but the exponential pass, two per tensor-memory lane quarter. Each owns 16
//     whole rows of the tile (tcgen05.ld / st in the 16-lane shapes: a row is spread over four threads as in an
//     MMA accumulator fragment, a thread holds two rows x 16 of every 64 keys), so each scheduler holds two
//     softmax warps whose load latencies, shift checks and stores hide behind each other's exponentials (one
//     warp per scheduler: 560 cycles per 32 exponentials where the MUFU pipe needs 256) and no two warps ever
//     touch the same row: no barrier inside the pass, P goes over scores the warp itself has consumed.
//   * the row maximum pass is gone: the shift of the softmax does not have to be the maximum, any m with
//     max - m <= 64 keeps 2^(x - m) far from overflow and leaves the result unchanged (P is rounded to bf16 and
//     accumulated in fp32: relative errors, no absolute ones). m = the maximum of the row's first 64 keys,
//     rounded up to an integer; a later chunk whose maximum exceeds m by more than TAU = 64 — a 2^64 ratio
//     between keys, not seen outside adversarial inputs — takes a slow path that raises m and rescales the sum
//     and the P already written by the exact power of two. The saved LSE is m ln2 + log(sum).
//   * two score accumulators in tensor memory: the MMA warp issues tile g + 2's QK^T right behind tile g's PV
//     MMA, so the scores of the next tile are waiting when the softmax warps finish a tile.
//   * four epilogue warps read O out of tensor memory, normalise and store it while the softmax warps are
//     already in the next tile; the row statistics cross through shared memory.
//   * Q, K, V of the next (batch, head) item land in a second shared-memory stage.
//   * the ragged last query tile (69 of 128 rows at N = 197) is rotated over the four lane quarters per item as
//     in the third-generation kernel, and the warps are decoupled (mbarriers only), so an idle warp moves on.
//   * the pass is kept small (one 64-key body per register buffer, one 16-key tail body): with every chunk
//     variant inlined at every call site an earlier version was 23 KB of straight-line code per loop iteration
//     and ran at 3.5 cycles per instruction.
#include "common.cuh"

namespace fv {

namespace {

constexpr int F4_SW = 8;                          // warps [0, 8) softmax (warp & 3 = lane quarter, warp >> 2 = its 16-lane half),
                                                  // [F4_SW, F4_SW + 4) epilogue,
constexpr int F4_W_MMA = F4_SW + 4;               // then MMA issue,
constexpr int F4_W_TMA = F4_SW + 5;               // then TMEM alloc + TMA producer
constexpr int F4_THREADS = 32 * (F4_SW + 6);
constexpr int F4_Q = 128;
constexpr int F4_KV_MAX = 208;
constexpr int F4_STAGE = 2 * F4_Q * 128 + 2 * F4_KV_MAX * 128;  // Q tiles, K, V of one item: 84 KiB
constexpr int F4_ONES = 16 * 128;  // one 16-key step of the "ones" column block appended to V (see the PV MMA)
constexpr int F4_SMEM = 2 * F4_STAGE + F4_ONES + 4 * 128 * 8 + 256 + 1024;
constexpr float F4_LOG2E = 1.4426950408889634f;
constexpr float F4_LN2 = 0.6931471805599453f;
constexpr float F4_TAU = 64.f;  // a chunk maximum more than 2^TAU above the row's shift raises the shift
constexpr uint32_t F4_T_S0 = 0;    // score / P buffer of even tiles: columns [0, 208)
constexpr uint32_t F4_T_S1 = 224;  // odd tiles: [224, 432)
constexpr uint32_t F4_T_O = 432;   // O accumulator: [432, 496) + the row sums of the rounded P in column 496

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
// ---- 16-lane tensor-memory shapes: the warp reads / writes lanes [l0, l0 + 16) (l0 = the lane field of the
// address); thread t holds row t / 4 and row t / 4 + 8 of them.
// 16x256b.xN: N x 8 fp32 columns; registers 4 i .. 4 i + 3 = (row A, columns 8 i + 2 (t % 4) + {0, 1}), (row B, same)
__device__ __forceinline__ void f4_ld64(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void f4_ld16(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// 16x128b.xN: N x 4 packed-bf16 columns; registers 2 i, 2 i + 1 = (row A, column 4 i + t % 4), (row B, same) — the
// packed pair of a 16x256b repetition lands exactly there
__device__ __forceinline__ void f4_st64(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void f4_st16(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void f4_ldp16(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
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
__device__ __forceinline__ void f4_ld_wait8(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :
               : "memory");
}
__device__ __forceinline__ float f4_quad_max(float v) {  // over the four threads that share a row
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}

// the softmax state of a thread's two rows (the four threads of a row carry the same shift and partial sums)
struct F4Rows {
  float mA, mB;  // the shifts, integers in the log2 domain (x = score * scale * log2e)
  float sA, sB;  // this thread's share of the sums of 2^(x - m)
};

// Slow path: raise the shifts of the rows whose chunk maximum (cA, cB: this thread's share, log2 domain) is more
// than TAU above them, and scale what the warp has produced for those rows so far — the partial sums and the
// packed P of its first `done16` groups of 16 keys at tp — by the exact 2^(m - m').
__device__ __noinline__ float4 f4_raise(uint32_t tp, int done16, float cA, float cB, float4 st) {  // st = (mA, mB, sA, sB)
  cA = f4_quad_max(cA);
  cB = f4_quad_max(cB);
  const float mA = cA > st.x + F4_TAU ? ceilf(cA) : st.x;
  const float mB = cB > st.y + F4_TAU ? ceilf(cB) : st.y;
  const float rA = f4_ex2(st.x - mA), rB = f4_ex2(st.y - mB);  // 1 for the rows that keep their shift
  tmem_st_wait();  // the P stores issued so far have to land before they are read back
  for (int j = 0; j < done16; ++j) {
    uint32_t pk[4];
    f4_ldp16(tp + j * 8, pk);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 v = unpack_bf16(pk[i]);
      const float r = (i & 1) ? rB : rA;
      pk[i] = pack_bf16(v.x * r, v.y * r);
    }
    f4_st16(tp + j * 8, pk);
  }
  tmem_st_wait();
  return make_float4(mA, mB, st.z * rA, st.w * rB);
}

// this thread's share of the maxima of its two rows over W (64 or 16) keys: registers 4 i + {0, 1} row A, {2, 3} row B
template <int W>
__device__ __forceinline__ void f4_max(const uint32_t* r, float& cA, float& cB) {
  cA = fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1]));
  cB = fmaxf(__uint_as_float(r[2]), __uint_as_float(r[3]));
#pragma unroll
  for (int i = 1; i < W / 8; ++i) {
    cA = f4_max3(cA, __uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]));
    cB = f4_max3(cB, __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
  }
}
// shift check before a chunk's exponentials (done16 = 16-key groups whose P this warp has already written);
// called with no tensor-memory load in flight — the slow path is a function call, and registers an asynchronous
// load still targets must not be live (or reused) across it
__device__ __forceinline__ void f4_check(float cA, float cB, float sl2, int done16, uint32_t tp, F4Rows& rows) {
  cA *= sl2;
  cB *= sl2;
  if (__any_sync(0xffffffffu, cA > rows.mA + F4_TAU || cB > rows.mB + F4_TAU)) {
    const float4 st = f4_raise(tp, done16, cA, cB, make_float4(rows.mA, rows.mB, rows.sA, rows.sB));
    rows.mA = st.x;
    rows.mB = st.y;
    rows.sA = st.z;
    rows.sB = st.w;
  }
}
// exponentials of W keys of the two rows: packed bf16 P into pk[0 .. W/4) in the 16x128b register order, sums
template <int W>
__device__ __forceinline__ void f4_exp(const uint32_t* r, float sl2, F4Rows& rows, uint32_t* pk) {
  const float mA = rows.mA, mB = rows.mB;
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
  for (int i = 0; i < W / 8; ++i) {
    const float p0 = f4_ex2(fmaf(__uint_as_float(r[4 * i]), sl2, -mA));
    const float p1 = f4_ex2(fmaf(__uint_as_float(r[4 * i + 1]), sl2, -mA));
    const float p2 = f4_ex2(fmaf(__uint_as_float(r[4 * i + 2]), sl2, -mB));
    const float p3 = f4_ex2(fmaf(__uint_as_float(r[4 * i + 3]), sl2, -mB));
    a0 += p0;
    a1 += p1;
    b0 += p2;
    b1 += p3;
    pk[2 * i] = pack_bf16(p0, p1);
    pk[2 * i + 1] = pack_bf16(p2, p3);
  }
  rows.sA += a0 + a1;
  rows.sB += b0 + b1;
}

__global__ void __launch_bounds__(F4_THREADS, 1)
attn_tc_fwd4_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const Fwd4Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // stage s: Q tiles (2 x 16 KiB) | K (26 KiB) | V (26 KiB)
  uint8_t* ones = smem + 2 * F4_STAGE;  // 16 keys x 64 columns (MN-major, 128B swizzle): column 0 = 1, the rest 0
  float2* stats = reinterpret_cast<float2*>(ones + F4_ONES);  // [4][128] (m, sum) of the rows of tile g & 3
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones + F4_ONES + 4 * 128 * 8);
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
  if (threadIdx.x < 128) {
    // row r of the block is 128 bytes; the swizzle puts its first 16-byte chunk (columns 0 .. 7) at chunk r % 8
    uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const int r = threadIdx.x >> 3, c = threadIdx.x & 7;
    if (c == (r & 7)) z.x = 0x3f80u;  // bf16 1.0 in column 0
    reinterpret_cast<uint4*>(ones)[threadIdx.x] = z;
    fence_proxy_async_smem();  // written through the generic proxy, read by the tensor core
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
    // O = P [V | 1]: 16 columns more than the head, the first of them all ones — column 64 of the accumulator is
    // the row sum of P exactly as the tensor core saw it (rounded to bf16), and O / that sum is a true convex
    // combination of the value rows (normalising by the fp32 sum of the unrounded exponentials is off by the
    // rounding of every P; with the row maximum no longer pinned to P = 1 a one-key row showed it: 0.4 %)
    const uint32_t idesc_o = make_idesc(kFmtBF16, 0, 1, F4_Q, 80);
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
      const uint32_t sV = smem_u32(smem + s * F4_STAGE) + (2 * F4_Q + F4_KV_MAX) * 128;
      const uint32_t s1 = smem_u32(ones);
      const uint32_t tp = tmem + ((g & 1) ? F4_T_S1 : F4_T_S0);
      F4_STAMP(g, 6);
      if (elect_one()) {
        // per 16-key step: V rows at sV + 2 KiB k; the second 64-column block of the operand ("leading" byte
        // offset) always resolves to the one ones block
        for (int k = 0; k < (p.kw >> 4); ++k)
          umma_bf16_ts(tmem + F4_T_O, tp + k * 8, make_smem_desc_sw128(sV + k * 2048, s1 - (sV + k * 2048), 1024), idesc_o,
                       k > 0 ? 1u : 0u);
        umma_commit(bar_o);
        if (t == nqt - 1) umma_commit(bar_vfree + s);
      }
      __syncwarp();
      if (g + 2 < ntiles) issue_scores(g + 2);
    }
  } else if (warp < F4_SW) {
    // ------------------------------ softmax: one pass over the scores ---------------------------
    const int wq = warp & 3;   // TMEM lane quarter
    const int sub = warp >> 2; // its lanes [16 sub, 16 sub + 16)
    const int qd = lane & 3;   // a row's four threads: thread qd holds columns 8 i + 2 qd + {0, 1}
    const float sl2 = p.scale * F4_LOG2E;
    const int n64 = (p.kw - 16) >> 6;          // 64-key chunks, none of them ragged
    const int n16 = (p.kw - n64 * 64) >> 4;    // + 1 .. 4 groups of 16 keys, the last one holding the ragged end
    const int nv_last = p.N - (p.kw - 16);     // keys of the sequence in it (1 .. 16)
    for (int g = 0; g < ntiles; ++g) {
      const int n = g / nqt, t = g % nqt;
      const int rot = (n + blockIdx.x) & 3;
      const int rb = t == nqt - 1 ? (wq - rot) & 3 : wq;         // this quarter's 32-row block of the tile
      const bool warp_live = t * F4_Q + rb * 32 + sub * 16 < p.N;  // warp-uniform
      const uint32_t tp = tmem + (static_cast<uint32_t>(wq * 32 + sub * 16) << 16) + ((g & 1) ? F4_T_S1 : F4_T_S0);
      if (warp == 0) F4_STAMP(g, 0);
      mbar_wait(bar_s + (g & 1), (g >> 1) & 1);
      tc_fence_after();
      if (warp == 0) F4_STAMP(g, 1);
      if (warp_live) {
        F4Rows rows;
        rows.sA = rows.sB = 0.f;
        uint32_t a[32], bq[32];
        // ---- 64-key chunks through two register buffers, the load of the next one in flight during the exponentials
        if (n64 > 0) {
          f4_ld64(tp, a);
          f4_ld_wait(a);
        }
        for (int k = 0; k < n64; k += 2) {
          uint32_t pk[16];
          float cA, cB;
          if (k > 0) f4_ld_wait(a);
          F4_CSTAMP(k, 0);
          f4_max<64>(a, cA, cB);
          if (k == 0) {  // the first 64 keys set the rows' shifts
            rows.mA = ceilf(f4_quad_max(cA) * sl2);
            rows.mB = ceilf(f4_quad_max(cB) * sl2);
          } else {
            f4_check(cA, cB, sl2, 4 * k, tp, rows);
          }
          f4_ld64(tp + (k + 1) * 64, bq);  // unconditional (a branch here sinks the load below the exponentials)
          F4_CSTAMP(k, 1);
          f4_exp<64>(a, sl2, rows, pk);
          F4_CSTAMP(k, 2);
          f4_st64(tp + k * 32, pk);
          F4_CSTAMP(k, 3);
          if (k + 1 < n64) {
            f4_ld_wait(bq);
            F4_CSTAMP(k + 1, 0);
            f4_max<64>(bq, cA, cB);
            f4_check(cA, cB, sl2, 4 * k + 4, tp, rows);
            f4_ld64(tp + (k + 2) * 64, a);
            F4_CSTAMP(k + 1, 1);
            f4_exp<64>(bq, sl2, rows, pk);
            F4_CSTAMP(k + 1, 2);
            f4_st64(tp + (k + 1) * 32, pk);
            F4_CSTAMP(k + 1, 3);
          }
        }
        tmem_ld_wait();  // retires the loop's last (unused) look-ahead load before any call below
        // ---- the remaining 16-key groups; in the last one the keys outside the sequence get P = 0
        for (int j = 0; j < n16; ++j) {
          uint32_t t8[8], pk[4];
          const int done16 = n64 * 4 + j;
          f4_ld16(tp + done16 * 16, t8);
          f4_ld_wait8(t8);
          if (j == n16 - 1) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int key = 8 * i + 2 * qd;
              if (key >= nv_last) t8[4 * i] = t8[4 * i + 2] = 0xff800000u;  // -inf: exponential 0, no say in the maximum
              if (key + 1 >= nv_last) t8[4 * i + 1] = t8[4 * i + 3] = 0xff800000u;
            }
          }
          float cA, cB;
          f4_max<16>(t8, cA, cB);
          if (done16 == 0) {
            rows.mA = ceilf(f4_quad_max(cA) * sl2);
            rows.mB = ceilf(f4_quad_max(cB) * sl2);
          } else {
            f4_check(cA, cB, sl2, done16, tp, rows);
          }
          f4_exp<16>(t8, sl2, rows, pk);
          f4_st16(tp + done16 * 8, pk);
        }
        // ---- row sums over the four threads of a row
        rows.sA += __shfl_xor_sync(0xffffffffu, rows.sA, 1);
        rows.sB += __shfl_xor_sync(0xffffffffu, rows.sB, 1);
        rows.sA += __shfl_xor_sync(0xffffffffu, rows.sA, 2);
        rows.sB += __shfl_xor_sync(0xffffffffu, rows.sB, 2);
        if (qd == 0) {
          stats[(g & 3) * 128 + wq * 32 + sub * 16 + (lane >> 2)] = make_float2(rows.mA, rows.sA);
          stats[(g & 3) * 128 + wq * 32 + sub * 16 + (lane >> 2) + 8] = make_float2(rows.mB, rows.sB);
        }
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
      uint32_t o0[32], o1[32], psum = 0;
      if (warp_live) {
        tmem_ld_32x32(to, o0);
        tmem_ld_32x32(to + 32, o1);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(psum) : "r"(to + 64) : "memory");
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(bar_ofree);
      if (warp_live && q < p.N) {
        // written before the row's arrival on bar_p, which the PV MMA behind bar_o waited for
        const float2 st = stats[(g & 3) * 128 + w4 * 32 + lane];
        const float inv = 1.0f / __uint_as_float(psum);  // sum of the rounded P (st.y, the exact one, goes into the LSE)
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

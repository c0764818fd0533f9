// attention_bwd2.cu — attention backward for the short ViT sequences (N <= 256: 197 at 224 px),
// head_dim 64, second generation: KEYS ON LANES.
//
// Replaces the autograd backward of F.scaled_dot_product_attention inside timm's Attention.forward
// (reference call site model.py:193, backward from train.py:153).
//
// What the first kernel (attention_tc.cu: attn_tc_bwd_kernel, queries on lanes) was bound by — ncu,
// profiles/r1_ncu_full_attn_bwd.csv: tensor pipe 24 % active, the shared-memory pipe the busiest unit —
// is the traffic of P and dS: both cross shared memory (64 KB of stores per 128 x 128 cell) and are read
// back by three N = 64 shared-memory-operand MMAs (144 KB), which at 6 KB per 32-cycle step run at the
// shared-memory port's speed, not the tensor core's; and its math warps idled 45 % of the time on one long
// dependency chain per cell (scores -> softmax + dS -> dV / dK / dQ).
//
// Here the scores are computed TRANSPOSED, S^T = K Q^T and dP^T = V dO^T, so that a thread of the math
// warps owns a KEY row:
//   * P^T [keys x queries] is written back into tensor memory over the scores it was computed from and
//     is the TMEM A operand of dV = P^T dO  (tcgen05.mma, A from TMEM: no shared-memory read of P at
//     all, only the 2 KB dO slice per step) — P never touches shared memory;
//   * dS^T goes to shared memory ONCE (32 KB per cell, double-buffered) and serves both remaining
//     products from the same bytes: read K-major it is the A operand of dK = dS^T Q, read MN-major it is
//     the A operand of dQ = dS K;
//   * the cell is split into two phases on two TMEM buffers, X (S^T -> P^T) and Y (dP^T): phase A (exp)
//     works on X while the tensor core fills Y, phase B (dS) works on Y while the tensor core computes the
//     next cell's S^T into X and this cell's dV — the math warps and the tensor pipe alternate on two
//     buffers instead of waiting for each other once per cell.
// Per-query statistics (LSE, delta) vary along the COLUMNS in this layout: every thread needs all of them,
// and reads them as broadcast 128-bit shared-memory loads (16 + 16 per cell and thread).
//
// TMEM (512 columns): X [0,128)  Y [128,256)  dV [256,320)  dK [320,384)  dQ tile 0 / 1 [384,448) / [448,512)
// Shared memory: Q, dO, K, V tiles (2 each, 128 KB), dS^T (2 x 32 KB), TMA-store staging (16 KB),
// {lse, delta}[4 items] (8 KB).
//
// Padding needs no masks: key rows past N are zero (TMA fill) -> their dS multiplies zero K rows in dQ
// and their dK / dV rows are clipped by the output tensor map; query columns past N have lse = +inf
// (P = 0) and zero dO rows (dP = delta = 0).
#include "common.cuh"

// A/B build switches (tools/build_variants.py compiles one library per setting). Measured at
// (256, 197, 12), profiles/r2_attn_bwd_variants.txt: coalesced helper loads 212 -> 270 us (8 dependent
// global round trips per item instead of 4), early accumulator release 212 -> 264 us — both off.
#ifndef B2_HELPER_COALESCED
#define B2_HELPER_COALESCED 0
#endif
#ifndef B2_EARLY_RELEASE
#define B2_EARLY_RELEASE 0
#endif

namespace fv {

int make_qkv_map(CUtensorMap* map, const void* base, int64_t batch, int64_t tokens, int64_t width, int rows);

namespace {

constexpr int B2_THREADS = 384;  // warps 0-7 math, 8 MMA issue, 9 TMEM alloc + TMA producer, 10-11 delta / LSE helpers
constexpr int B2_TILE = 128 * 128;   // bytes of one [128 x 64] bf16 tile
constexpr int B2_AUX = 4;            // items of lse / delta the helper warps may run ahead
constexpr int B2_STAGE = 16 * 128;   // per math warp: 16 output rows x 128 bytes on their way to a TMA store
constexpr int B2_SMEM = 8 * B2_TILE + 4 * B2_TILE + 8 * B2_STAGE + 1024 + B2_AUX * 2048 + 512;
constexpr float B2_LOG2E = 1.4426950408889634f;

struct Bwd2Params {
  int N, H, kw;
  int items;  // batch * heads
  float scale;
  const float* lse;
  const __nv_bfloat16* o;     // forward output  [B, N, H*64]
  const __nv_bfloat16* dout;  // its gradient    [B, N, H*64]
};

__device__ __forceinline__ float b2_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(B2_THREADS, 1)
attn_tc_bwd2_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                    const __grid_constant__ CUtensorMap tmap_dqkv, const Bwd2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                    // 2 tiles
  uint8_t* sdO = sQ + 2 * B2_TILE;       // 2 tiles
  uint8_t* sK = sdO + 2 * B2_TILE;       // 2 tiles (both key blocks)
  uint8_t* sV = sK + 2 * B2_TILE;        // 2 tiles
  uint8_t* sdS = sV + 2 * B2_TILE;       // 2 buffers x [2 column blocks of 64 queries][128 key rows x 128 B]
  uint8_t* sStage = sdS + 4 * B2_TILE;   // 8 x 2 KiB, 1024-byte aligned (TMA 128B-swizzle atoms)
  float* sAux = reinterpret_cast<float*>(sStage + 8 * B2_STAGE);  // [B2_AUX items][lse*log2e[256], delta*scale[256]]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sAux + B2_AUX * 512);
  uint64_t* bar_ld = bars + 0;     // [3] loaded: {Q0 dO0 K0 V0}, {Q1 dO1}, {K1 V1}
  uint64_t* bar_free = bars + 3;   // [4] last reader retired: {K0 V0}, {Q0 dO0}, {K1 V1}, {Q1 dO1}
  uint64_t* bar_sx = bars + 7;     // S^T of a cell is in X
  uint64_t* bar_dp = bars + 8;     // dP^T of a cell is in Y
  uint64_t* bar_pt = bars + 9;     // phase A done: X read, P^T written over it
  uint64_t* bar_yfree = bars + 10; // phase B has Y in registers
  uint64_t* bar_dv = bars + 11;    // dV of a cell retired (P^T consumed: X may be overwritten)
  uint64_t* bar_ds = bars + 12;    // [2] dS^T buffer written
  uint64_t* bar_m2 = bars + 14;    // [2] dK / dQ of a cell retired: dS^T buffer free, accumulators final
  uint64_t* bar_kvfree = bars + 16;
  uint64_t* bar_dqfree = bars + 17;
  uint64_t* bar_aux = bars + 18;                // [B2_AUX] lse / delta of item n ready in sAux[n % B2_AUX]
  uint64_t* bar_auxfree = bar_aux + B2_AUX;     // [B2_AUX] ... and read for the last time
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_auxfree + B2_AUX);

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
    mbar_init(bar_sx, 1);
    mbar_init(bar_dp, 1);
    mbar_init(bar_pt, 256);
    mbar_init(bar_yfree, 256);
    mbar_init(bar_dv, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_ds[i], 256);
      mbar_init(&bar_m2[i], 1);
    }
    mbar_init(bar_kvfree, 256);
    mbar_init(bar_dqfree, 256);
    for (int i = 0; i < B2_AUX; ++i) {
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
  constexpr uint32_t T_X = 0, T_Y = 128, T_DV = 256, T_DK = 320, T_DQ = 384;

  // Registers: 384 threads x 168. The math warps keep P between the phases as the packed bf16 pairs the
  // tensor core reads (32 registers, unpacked again in phase B); kept as 64 fp32 values ptxas parked them
  // in local memory and every phase-B multiply waited for an LDL (ncu: 2x slower than the kernel this
  // replaces). setmaxnreg re-partitioning (104 / 200) made ptxas spill MORE here and is not used.

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
        mbar_expect_tx(&bar_ld[0], 4 * B2_TILE);
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
          mbar_expect_tx(&bar_ld[1], 2 * B2_TILE);
          tma_load_3d(sQ + B2_TILE, &tmap_qkv, &bar_ld[1], h * 64, 128, b);
          tma_load_3d(sdO + B2_TILE, &tmap_do, &bar_ld[1], h * 64, 128, b);
        }
        __syncwarp();
        if (n > 0) mbar_wait(&bar_free[2], prev);
        if (elect_one()) {
          mbar_expect_tx(&bar_ld[2], 2 * B2_TILE);
          tma_load_3d(sK + B2_TILE, &tmap_qkv, &bar_ld[2], hd + h * 64, 128, b);
          tma_load_3d(sV + B2_TILE, &tmap_qkv, &bar_ld[2], 2 * hd + h * 64, 128, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 8) {
    // ------------------------------ MMA issue (whole warp, elected lane issues) -----------------
    // Order on the (in-order) tensor pipe per cell i:  dV(i) | S^T(i+1) | dP^T(i+1) | dK(i) dQ(i).
    // S^T(i+1) overwrites X, whose P^T is dV(i)'s A operand: it is issued after dV(i) has RETIRED
    // (bar_dv); dP^T(i+1) overwrites Y once phase B of cell i holds it in registers (bar_yfree).
    const uint32_t id_ts = make_idesc(kFmtBF16, 0, 1, 128, 64);   // A from TMEM (P^T), B MN-major (dO)
    const uint32_t id_dk = make_idesc(kFmtBF16, 0, 1, 128, 64);   // A K-major (dS^T rows), B MN-major (Q)
    const uint32_t id_dq = make_idesc(kFmtBF16, 1, 1, 128, 64);   // A MN-major (dS^T read as dS), B MN-major (K)
    auto q_cols = [&](int qt) {  // queries of tile qt, rounded up to the MMA granule
      int c = p.kw - qt * 128;
      return c > 128 ? 128 : c;
    };
    auto issue_s = [&](int kb, int qt) {
      const uint32_t id_s = make_idesc(kFmtBF16, 0, 0, 128, q_cols(qt));
      const uint64_t dK_k = make_smem_desc_sw128(smem_u32(sK + kb * B2_TILE), 16, 1024);
      const uint64_t dQ_k = make_smem_desc_sw128(smem_u32(sQ + qt * B2_TILE), 16, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + T_X, dK_k + k * 2, dQ_k + k * 2, id_s, k > 0 ? 1u : 0u);
        umma_commit(bar_sx);
      }
      __syncwarp();
    };
    auto issue_dp = [&](int kb, int qt) {
      const uint32_t id_s = make_idesc(kFmtBF16, 0, 0, 128, q_cols(qt));
      const uint64_t dV_k = make_smem_desc_sw128(smem_u32(sV + kb * B2_TILE), 16, 1024);
      const uint64_t dO_k = make_smem_desc_sw128(smem_u32(sdO + qt * B2_TILE), 16, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + T_Y, dV_k + k * 2, dO_k + k * 2, id_s, k > 0 ? 1u : 0u);
        umma_commit(bar_dp);
      }
      __syncwarp();
    };
    // operand tiles of cell (kb, qt) of the item whose load phase is `ph`
    auto wait_tiles = [&](int kb, int qt, uint32_t ph) {
      if (kb == 0 && qt == 0) mbar_wait(&bar_ld[0], ph);
      else if (kb == 0) mbar_wait(&bar_ld[1], ph);
      else if (qt == 0) mbar_wait(&bar_ld[2], ph);
      tc_fence_after();
    };
    int it = 0;   // cells so far, over all items
    int kvn = 0;  // key blocks so far
    int n = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++n) {
      const uint32_t ph = n & 1;
      if (n == 0) {
        wait_tiles(0, 0, 0);
        issue_s(0, 0);
        issue_dp(0, 0);
      }
      const bool has_next = item + static_cast<int>(gridDim.x) < nitems;
      for (int kb = 0; kb < nt; ++kb, ++kvn) {
        int kwb = p.kw - kb * 128;
        if (kwb > 128) kwb = 128;
        const uint64_t dK_mn = make_smem_desc_sw128(smem_u32(sK + kb * B2_TILE), B2_TILE, 1024);  // MN-major view
        for (int qt = 0; qt < nt; ++qt, ++it) {
          const int nq = q_cols(qt);
          const uint64_t dQ_mn = make_smem_desc_sw128(smem_u32(sQ + qt * B2_TILE), B2_TILE, 1024);
          const uint64_t dO_mn = make_smem_desc_sw128(smem_u32(sdO + qt * B2_TILE), B2_TILE, 1024);
          // the cell after this one
          int nkb = kb, nqt = qt + 1;
          bool in_item = true;
          if (nqt == nt) { nqt = 0; nkb = kb + 1; }
          if (nkb == nt) { nkb = 0; in_item = false; }
          // ---- dV(it) = P^T dO: A from tensor memory, 16 queries (8 packed columns) per step ----
          mbar_wait(bar_pt, it & 1);
          if (qt == 0 && kvn > 0) mbar_wait(bar_kvfree, (kvn - 1) & 1);  // previous block's dK / dV have left TMEM
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (k < (nq >> 4))
                umma_bf16_ts(tmem + T_DV, tmem + T_X + 32 * (k >> 1) + 8 * (k & 1), dO_mn + k * 128, id_ts,
                             (qt > 0 || k > 0) ? 1u : 0u);
            umma_commit(bar_dv);
          }
          __syncwarp();
          // ---- next cell's scores as early as its operands allow ----
          bool early = false;
          if (in_item) {
            wait_tiles(nkb, nqt, ph);
            early = true;
          } else if (has_next) {
            // next item: its first tiles are prefetched while this item's last cells run — if they have
            // not landed (always the case for single-cell items, whose tiles are released by THIS
            // cell's last MMA) the scores are issued after dK / dQ instead
            early = __all_sync(0xffffffffu, mbar_try_wait(&bar_ld[0], ph ^ 1) ? 1 : 0) != 0;
            if (early) tc_fence_after();
          }
          if (early) {
            mbar_wait(bar_dv, it & 1);
            tc_fence_after();
            issue_s(nkb, nqt);
            mbar_wait(bar_yfree, it & 1);
            tc_fence_after();
            issue_dp(nkb, nqt);
          }
          // ---- dK(it) = dS^T Q (A K-major), dQ(it) = dS K (the same bytes read MN-major) ----
          mbar_wait(&bar_ds[it & 1], (it >> 1) & 1);
          if (kb == 0 && qt == 0 && n > 0) mbar_wait(bar_dqfree, (n - 1) & 1);  // previous item's dQ has left TMEM
          tc_fence_after();
          uint8_t* ds = sdS + (it & 1) * 2 * B2_TILE;
          const uint64_t dS_k0 = make_smem_desc_sw128(smem_u32(ds), 16, 1024);
          const uint64_t dS_mn = make_smem_desc_sw128(smem_u32(ds), B2_TILE, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (k < (nq >> 4)) {
                // dS^T k-step: query column block k >> 2 (one 16 KB tile apart), 32 bytes per 16 queries inside it
                const uint64_t dS_k = dS_k0 + (((k >> 2) * B2_TILE + (k & 3) * 32) >> 4);
                umma_bf16(tmem + T_DK, dS_k, dQ_mn + k * 128, id_dk, (qt > 0 || k > 0) ? 1u : 0u);
              }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (k < (kwb >> 4))
                umma_bf16(tmem + T_DQ + qt * 64, dS_mn + k * 128, dK_mn + k * 128, id_dq, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&bar_m2[it & 1]);
            // operand tiles whose last reader was just issued: free them for the next item's loads
            if (kb == 0 && qt == nt - 1) umma_commit(&bar_free[0]);
            if (kb == nt - 1 && qt == 0) umma_commit(&bar_free[1]);
            if (nt > 1 && kb == 1 && qt == nt - 1) umma_commit(&bar_free[2]);
            if (nt > 1 && kb == nt - 1 && qt == 1) umma_commit(&bar_free[3]);
          }
          __syncwarp();
          if (!early && (in_item || has_next)) {
            wait_tiles(nkb, nqt, in_item ? ph : (ph ^ 1));
            mbar_wait(bar_dv, it & 1);
            tc_fence_after();
            issue_s(nkb, nqt);
            mbar_wait(bar_yfree, it & 1);
            tc_fence_after();
            issue_dp(nkb, nqt);
          }
        }
      }
    }
  } else if (warp >= 10) {
    // ------------------------------ delta / LSE helpers ------------------------------------------
    // delta[q] * scale = scale * sum_d dO[q,d] * O[q,d] and lse[q] * log2(e) of the NEXT items, straight
    // from global memory (O never occupies shared memory), up to B2_AUX - 1 items ahead of the math warps
    int n = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++n) {
      const int b = item / p.H, h = item % p.H;
      const int slot = n % B2_AUX;
      if (n >= B2_AUX) mbar_wait(&bar_auxfree[slot], (n / B2_AUX - 1) & 1);  // previous tenant read out
      float* aux = sAux + slot * 512;
      const float* lse_bh = p.lse + (static_cast<long long>(b) * p.H + h) * p.N;
#if B2_HELPER_COALESCED
      // Eight lanes per row: a warp-wide 128-bit load covers four whole 128-byte rows (4 wavefronts of the
      // L1 data pipe). One row per THREAD, as the first kernel reads them, touches 32 different lines per
      // instruction — 32 wavefronts — and the two helper warps alone produced half of that pipe's load/store
      // traffic (ncu: ~1000 of ~2000 LSU wavefronts per cell), on the pipe this kernel is bound by.
      const int sub = lane & 7, rsel = lane >> 3;  // 16-byte piece of the row, row within the group of 4
      const int hw = warp - 10;                    // helper warp 0 / 1
      for (int q0 = hw * 4; q0 < nt * 128; q0 += 8 * 4) {
        // 4 passes of 4 rows per iteration: 8 independent 128-bit loads in flight per thread
        float acc[4];
        uint4 a[4], g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int q = q0 + u * 8 + rsel;
          if (q < p.N) {
            const long long off = (static_cast<long long>(b) * p.N + q) * hd + h * 64;
            a[u] = __ldg(reinterpret_cast<const uint4*>(p.o + off) + sub);
            g[u] = __ldg(reinterpret_cast<const uint4*>(p.dout + off) + sub);
          } else {
            a[u] = g[u] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t aw[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
          const uint32_t gw[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
          float t = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const float2 af = unpack_bf16(aw[w]), gf = unpack_bf16(gw[w]);
            t = fmaf(af.x, gf.x, t);
            t = fmaf(af.y, gf.y, t);
          }
          t += __shfl_xor_sync(0xffffffffu, t, 1);
          t += __shfl_xor_sync(0xffffffffu, t, 2);
          t += __shfl_xor_sync(0xffffffffu, t, 4);
          acc[u] = t;
        }
        if (sub < 4) {  // lane `sub` of each group of eight publishes row u = sub
          const int q = q0 + sub * 8 + rsel;
          const float t = sub == 0 ? acc[0] : sub == 1 ? acc[1] : sub == 2 ? acc[2] : acc[3];
          if (q < nt * 128) {
            aux[q] = q < p.N ? __ldg(lse_bh + q) * B2_LOG2E : INFINITY;
            aux[256 + q] = t * p.scale;
          }
        }
      }
#else
      // one row per thread and pass, 16 independent 128-bit loads in flight per thread
      for (int q = threadIdx.x - 320; q < nt * 128; q += 64) {
        uint4 a[8], g[8];
        float l2 = INFINITY;
        if (q < p.N) {
          const long long off = (static_cast<long long>(b) * p.N + q) * hd + h * 64;
          const uint4* orow = reinterpret_cast<const uint4*>(p.o + off);
          const uint4* grow = reinterpret_cast<const uint4*>(p.dout + off);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            a[u] = __ldg(orow + u);
            g[u] = __ldg(grow + u);
          }
          l2 = __ldg(lse_bh + q) * B2_LOG2E;
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) a[u] = g[u] = make_uint4(0u, 0u, 0u, 0u);
        }
        float acc = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t aw[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
          const uint32_t gw[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const float2 af = unpack_bf16(aw[w]), gf = unpack_bf16(gw[w]);
            acc = fmaf(af.x, gf.x, acc);
            acc = fmaf(af.y, gf.y, acc);
          }
        }
        aux[q] = l2;
        aux[256 + q] = acc * p.scale;
      }
#endif
      mbar_arrive(&bar_aux[slot]);
    }
  } else if (warp < 8) {
    // ------------------------------ math + output warps ----------------------------------------
    // thread == key row r of the block (TMEM lane); the two warps of a lane quarter split the query
    // columns in alternating 32-column chunks (chunk j = 2c + hf), so a short query tile (69 queries at
    // N = 197) still splits evenly.
    const int quarter = warp & 3, hf = warp >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sl2 = p.scale * B2_LOG2E;

    // This thread's row of TMEM columns [col, col+64) -> 64 bf16 = one 128-byte line of dqkv slot `slot`;
    // rows leave through TMA tensor stores, 16 at a time, from a 2 KiB per-warp staging tile in the
    // 128B-swizzle layout. The 3-D tensor map clips rows past the sequence end.
    uint8_t* stage = sStage + warp * B2_STAGE;
    bool store_pending = false;
    // part 1: the accumulator row leaves tensor memory (packed to bf16 as it arrives: two passes of 32
    // columns) — after this the MMA warp may overwrite the accumulator
    auto load_row = [&](uint32_t col, uint32_t (&w)[32]) {
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        uint32_t o[32];
        tmem_ld_32x32(lane_base + col + 32 * part, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          w[16 * part + i] = pack_bf16(__uint_as_float(o[2 * i]), __uint_as_float(o[2 * i + 1]));
      }
    };
    // part 2: staging tile -> TMA tensor store, 16 rows at a time
    auto store_row = [&](const uint32_t (&w)[32], int b, int h, int slot, int tile) {
      const int tok0 = tile * 128 + quarter * 32;  // first row of this warp
      if (tok0 >= p.N) return;  // warp-uniform
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
    // Drains are deferred into the next cell, between its two phases: dK / dV of a finished key block
    // (and dQ of a finished item) leave tensor memory while the tensor core works on the next cell.
    bool pend_kv = false, pend_dq = false;
    int kv_b = 0, kv_h = 0, kv_kb = 0, dq_b = 0, dq_h = 0, pend_it = 0;
    auto drain = [&]() {
      if (!(pend_kv || pend_dq)) return;
      mbar_wait(&bar_m2[pend_it & 1], (pend_it >> 1) & 1);  // the cell that completed them has retired
      tc_fence_after();
      // the accumulators are released the moment they are in registers: the MMA warp's next dV (which
      // overwrites them) does not wait for the staging stores and the TMA issue
      if (pend_kv) {
        uint32_t w[32];
        load_row(hf == 0 ? T_DK : T_DV, w);
#if B2_EARLY_RELEASE
        tc_fence_before();
        mbar_arrive(bar_kvfree);
        store_row(w, kv_b, kv_h, hf == 0 ? 1 : 2, kv_kb);
#else
        store_row(w, kv_b, kv_h, hf == 0 ? 1 : 2, kv_kb);
        tc_fence_before();
        mbar_arrive(bar_kvfree);
#endif
        pend_kv = false;
      }
      if (pend_dq) {
        uint32_t w[32];
        if (hf < nt) load_row(T_DQ + hf * 64, w);
#if B2_EARLY_RELEASE
        tc_fence_before();
        mbar_arrive(bar_dqfree);
        if (hf < nt) store_row(w, dq_b, dq_h, 0, hf);
#else
        if (hf < nt) store_row(w, dq_b, dq_h, 0, hf);
        tc_fence_before();
        mbar_arrive(bar_dqfree);
#endif
        pend_dq = false;
      }
    };
    int it = 0;
    int n = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++n) {
      const int b = item / p.H, h = item % p.H;
      const float* aux = sAux + (n % B2_AUX) * 512;
      mbar_wait(&bar_aux[n % B2_AUX], (n / B2_AUX) & 1);
      for (int kb = 0; kb < nt; ++kb) {
        const bool rows_live = kb * 128 + quarter * 32 < p.N;  // warp-uniform: any real key in this warp's rows
        for (int qt = 0; qt < nt; ++qt, ++it) {
          int nq = p.kw - qt * 128;
          if (nq > 128) nq = 128;
          bool valid[2];
#pragma unroll
          for (int c = 0; c < 2; ++c) valid[c] = rows_live && (32 * (2 * c + hf) < nq);  // warp-uniform
          // ================= phase A: P^T = exp2(S^T * scale*log2e - lse*log2e), written over S^T =========
          // P is kept as the packed bf16 pairs the tensor core reads (16 registers per chunk) — phase B
          // unpacks them again (a shift / a mask per value): 32 instead of 64 registers live across the drain
          uint32_t pk[2][16];
          mbar_wait(bar_sx, it & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (valid[c]) {
              uint32_t s[32];
              tmem_ld_32x32(lane_base + T_X + 32 * (2 * c + hf), s);
              tmem_ld_wait();
              const float4* l2v = reinterpret_cast<const float4*>(aux + qt * 128 + 32 * (2 * c + hf));
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 l2 = l2v[i >> 2];
                const float p0 = b2_ex2(fmaf(__uint_as_float(s[i]), sl2, -l2.x));
                const float p1 = b2_ex2(fmaf(__uint_as_float(s[i + 1]), sl2, -l2.y));
                const float p2 = b2_ex2(fmaf(__uint_as_float(s[i + 2]), sl2, -l2.z));
                const float p3 = b2_ex2(fmaf(__uint_as_float(s[i + 3]), sl2, -l2.w));
                pk[c][i >> 1] = pack_bf16(p0, p1);
                pk[c][(i >> 1) + 1] = pack_bf16(p2, p3);
              }
            }
          }
          // chunk j's 16 packed columns land inside chunk j's own 32 score columns, which only this thread
          // reads — and has read: both chunks are in registers before the first P^T column is written
#pragma unroll
          for (int c = 0; c < 2; ++c)
            if (valid[c]) tmem_st_32x16(lane_base + T_X + 32 * (2 * c + hf), pk[c]);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(bar_pt);

          drain();  // whatever the previous cell completed

          // ================= phase B: dS^T = P^T * (dP^T * scale - delta * scale) -> shared memory ==========
          mbar_wait(bar_dp, it & 1);
          tc_fence_after();
          if (it >= 2) {
            mbar_wait(&bar_m2[it & 1], ((it - 2) >> 1) & 1);  // dK / dQ of cell it-2 retired: this dS^T buffer is free
          }
          // chunk j = 2c + hf: query column block c (64 queries each), 64-byte half hf of the 128-byte row
          uint8_t* ds = sdS + (it & 1) * 2 * B2_TILE;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t d[32];
            if (valid[c]) {
              tmem_ld_32x32(lane_base + T_Y + 32 * (2 * c + hf), d);
              tmem_ld_wait();
            }
            if (c == 1) {  // Y is in registers: the next cell's dP^T may overwrite it
              tc_fence_before();
              mbar_arrive(bar_yfree);
            }
            if (valid[c]) {
              const float4* dlv = reinterpret_cast<const float4*>(aux + 256 + qt * 128 + 32 * (2 * c + hf));
              uint32_t dk[16];
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 dl = dlv[i >> 2];
                const float2 pa = unpack_bf16(pk[c][i >> 1]), pb = unpack_bf16(pk[c][(i >> 1) + 1]);
                const float s0 = pa.x * fmaf(__uint_as_float(d[i]), p.scale, -dl.x);
                const float s1 = pa.y * fmaf(__uint_as_float(d[i + 1]), p.scale, -dl.y);
                const float s2 = pb.x * fmaf(__uint_as_float(d[i + 2]), p.scale, -dl.z);
                const float s3 = pb.y * fmaf(__uint_as_float(d[i + 3]), p.scale, -dl.w);
                dk[i >> 1] = pack_bf16(s0, s1);
                dk[(i >> 1) + 1] = pack_bf16(s2, s3);
              }
              uint8_t* srow = ds + c * B2_TILE + r * 128;
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int unit = (hf * 4 + u) ^ (r & 7);
                *reinterpret_cast<uint4*>(srow + (unit << 4)) = make_uint4(dk[u * 4], dk[u * 4 + 1], dk[u * 4 + 2], dk[u * 4 + 3]);
              }
            }
          }
          fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async proxy
          mbar_arrive(&bar_ds[it & 1]);
          if (qt == nt - 1) {
            pend_kv = true;
            kv_b = b; kv_h = h; kv_kb = kb;
            pend_it = it;
          }
        }
      }
      mbar_arrive(&bar_auxfree[n % B2_AUX]);
      pend_dq = true;
      dq_b = b; dq_h = h;
      pend_it = it - 1;
    }
    drain();
    if (lane == 0) tma_store_wait_all();  // this warp's tensor stores are complete before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace

int attention_tc_bwd2(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                      int64_t batch, int64_t tokens, int64_t heads, float scale, cudaStream_t stream) {
  Bwd2Params p;
  p.N = static_cast<int>(tokens);
  p.H = static_cast<int>(heads);
  p.kw = static_cast<int>((tokens + 15) / 16 * 16);
  p.items = static_cast<int>(batch * heads);
  p.scale = scale;
  p.lse = lse;
  p.o = reinterpret_cast<const __nv_bfloat16*>(out);
  p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  CUtensorMap mq, mdo, mdq;
  int rc = make_qkv_map(&mq, qkv, batch, tokens, 3 * heads * 64, 128);
  if (rc != FV_OK) return rc;
  rc = make_qkv_map(&mdo, dout, batch, tokens, heads * 64, 128);
  if (rc != FV_OK) return rc;
  // output: 16-row boxes of one 64-column (slot, head) group; rows past N are clipped
  rc = make_qkv_map(&mdq, dqkv, batch, tokens, 3 * heads * 64, 16);
  if (rc != FV_OK) return rc;
  static bool configured = false;
  if (!configured) {
    FV_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM));
    configured = true;
  }
  const int grid = p.items < num_sms() ? p.items : num_sms();
  FV_CHECK_CUDA(fv::launch_pdl(attn_tc_bwd2_kernel, dim3(static_cast<unsigned>(grid)), dim3(B2_THREADS), B2_SMEM, stream,
                               mq, mdo, mdq, p));
  count_kernel(FV_KERNEL_ATTN_BWD);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

}  // namespace fv

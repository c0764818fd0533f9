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
//   warps 4..7  epilogue       (thread == output row)  tcgen05.ld -> bias/GELU/residual -> global
// The epilogue of tile i overlaps the main loop of tile i+1 through the two TMEM stages.
#include "common.cuh"

namespace fv {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KiB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

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
};

struct Vec8 {
  float v[8];
};

__device__ __forceinline__ Vec8 load8_f32(const float* p) {
  Vec8 r;
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ Vec8 load8_bf16(const __nv_bfloat16* p) {
  Vec8 r;
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 t;
  t = unpack_bf16(u.x); r.v[0] = t.x; r.v[1] = t.y;
  t = unpack_bf16(u.y); r.v[2] = t.x; r.v[3] = t.y;
  t = unpack_bf16(u.z); r.v[4] = t.x; r.v[5] = t.y;
  t = unpack_bf16(u.w); r.v[6] = t.x; r.v[7] = t.y;
  return r;
}
__device__ __forceinline__ void store8_f32(float* p, const Vec8& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ void store8_bf16(__nv_bfloat16* p, const Vec8& r) {
  uint4 u;
  u.x = pack_bf16(r.v[0], r.v[1]);
  u.y = pack_bf16(r.v[2], r.v[3]);
  u.z = pack_bf16(r.v[4], r.v[5]);
  u.w = pack_bf16(r.v[6], r.v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void red_add8_f32(float* p, const Vec8& r) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]),
               "f"(r.v[2]), "f"(r.v[3])
               : "memory");
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + 4), "f"(r.v[4]),
               "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7])
               : "memory");
}

template <int EPI>
__device__ __forceinline__ void epilogue_store8(const GemmTcParams& p, long long row, int col,
                                                Vec8 acc) {
  if (EPI != FV_EPI_ACCUM && EPI != FV_EPI_DGELU && p.bias != nullptr) {
    const Vec8 b = load8_f32(p.bias + col);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] += b.v[i];
  }
  if (EPI == FV_EPI_NONE) {
    if (p.c_bf16) store8_bf16(reinterpret_cast<__nv_bfloat16*>(p.c) + row * p.ldc + col, acc);
    else store8_f32(reinterpret_cast<float*>(p.c) + row * p.ldc + col, acc);
  } else if (EPI == FV_EPI_RESIDUAL) {
    const Vec8 r = load8_f32(reinterpret_cast<const float*>(p.aux) + row * p.ldaux + col);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] += r.v[i];
    store8_f32(reinterpret_cast<float*>(p.c) + row * p.ldc + col, acc);
  } else if (EPI == FV_EPI_GELU) {
    Vec8 g;
    if (p.c_bf16) {
      // the activation is computed from the *rounded* pre-activation, as autocast does
      // (fc1 emits bf16, nn.GELU then runs on that bf16 tensor)
      store8_bf16(reinterpret_cast<__nv_bfloat16*>(p.aux) + row * p.ldaux + col, acc);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        g.v[i] = gelu_erf(__bfloat162float(__float2bfloat16_rn(acc.v[i])));
      store8_bf16(reinterpret_cast<__nv_bfloat16*>(p.c) + row * p.ldc + col, g);
    } else {
      store8_f32(reinterpret_cast<float*>(p.aux) + row * p.ldaux + col, acc);
#pragma unroll
      for (int i = 0; i < 8; ++i) g.v[i] = gelu_erf(acc.v[i]);
      store8_f32(reinterpret_cast<float*>(p.c) + row * p.ldc + col, g);
    }
  } else if (EPI == FV_EPI_DGELU) {
    Vec8 u;
    if (p.c_bf16) u = load8_bf16(reinterpret_cast<const __nv_bfloat16*>(p.aux) + row * p.ldaux + col);
    else u = load8_f32(reinterpret_cast<const float*>(p.aux) + row * p.ldaux + col);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] *= gelu_erf_grad(u.v[i]);
    if (p.c_bf16) store8_bf16(reinterpret_cast<__nv_bfloat16*>(p.c) + row * p.ldc + col, acc);
    else store8_f32(reinterpret_cast<float*>(p.c) + row * p.ldc + col, acc);
  } else if (EPI == FV_EPI_ACCUM) {
    red_add8_f32(reinterpret_cast<float*>(p.c) + row * p.ldc + col, acc);
  } else if (EPI == FV_EPI_PATCH) {
    const long long img = row / p.tokens_per_img;
    const long long tok = row - img * p.tokens_per_img + 1;  // row 0 of every image is cls
    const Vec8 pe = load8_f32(reinterpret_cast<const float*>(p.aux) + tok * p.ldaux + col);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] += pe.v[i];
    store8_f32(reinterpret_cast<float*>(p.c) + (row + img + 1) * p.ldc + col, acc);
  }
}

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_mn = p.num_m_blocks * p.num_n_blocks;
  const int total_tiles = tiles_mn * p.split_k;

  if (warp == 0 && lane == 0) {
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
        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        uint8_t* sb = sa + A_STAGE_BYTES;
        if (p.a_major == FV_MAJOR_K) {
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j)
            tma_load_2d(sa + j * (64 * BK * 2), &tmap_a, &full_bar[stage], m_blk * BM + j * 64,
                        kb * BK);
        }
        if (p.b_major == FV_MAJOR_K) {
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_2d(sb + j * (64 * BK * 2), &tmap_b, &full_bar[stage], n_blk * BN + j * 64,
                        kb * BK);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------- MMA issuer ---------------------------------------------
    const uint32_t idesc = make_idesc(kFmtBF16, p.a_major, p.b_major, BM, BN);
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
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          umma_bf16(tmem_d, da + k * a_kstep, db + k * b_kstep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------- epilogue ------------------------------------------------
    const int wq = warp - 4;  // TMEM lane quarter this warp may read (== warp % 4)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = tile / tiles_mn;
      const int mn = tile - split * tiles_mn;
      const int m_blk = mn / p.num_n_blocks;
      const int n_blk = mn - m_blk * p.num_n_blocks;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const long long row = static_cast<long long>(m_blk) * BM + wq * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        const int col0 = n_blk * BN + chunk * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32(taddr + chunk * 32, r);
        tmem_ld_wait();
        if (row < p.M) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = col0 + g * 8;
            if (col < p.N) {
              Vec8 a;
#pragma unroll
              for (int i = 0; i < 8; ++i) a.v[i] = __uint_as_float(r[g * 8 + i]);
              epilogue_store8<EPI>(p, row, col, a);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
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

template <int EPI>
static int launch_gemm_tc(const CUtensorMap& ta, const CUtensorMap& tb, const GemmTcParams& p,
                          cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    FV_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<EPI>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       GEMM_SMEM_BYTES));
    configured = true;
  }
  const int total = p.num_m_blocks * p.num_n_blocks * p.split_k;
  const int grid = total < num_sms() ? total : num_sms();
  gemm_tc_kernel<EPI><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(ta, tb, p);
  FV_LAUNCH_CHECK();
  return FV_OK;
}

}  // namespace fv

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
  if (epilogue == FV_EPI_RESIDUAL || epilogue == FV_EPI_GELU || epilogue == FV_EPI_DGELU ||
      epilogue == FV_EPI_PATCH)
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

  CUtensorMap ta, tb;
  int rc = make_operand_map(&ta, a, a_major, m, k, lda, BM);
  if (rc != FV_OK) return rc;
  rc = make_operand_map(&tb, b, b_major, n, k, ldb, BN);
  if (rc != FV_OK) return rc;

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (epilogue) {
    case FV_EPI_NONE: return launch_gemm_tc<FV_EPI_NONE>(ta, tb, p, st);
    case FV_EPI_RESIDUAL: return launch_gemm_tc<FV_EPI_RESIDUAL>(ta, tb, p, st);
    case FV_EPI_GELU: return launch_gemm_tc<FV_EPI_GELU>(ta, tb, p, st);
    case FV_EPI_DGELU: return launch_gemm_tc<FV_EPI_DGELU>(ta, tb, p, st);
    case FV_EPI_ACCUM: return launch_gemm_tc<FV_EPI_ACCUM>(ta, tb, p, st);
    case FV_EPI_PATCH: return launch_gemm_tc<FV_EPI_PATCH>(ta, tb, p, st);
  }
  return FV_ERR_INVALID_ARG;
}

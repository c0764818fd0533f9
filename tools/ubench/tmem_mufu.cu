// Microbenchmarks behind the attention kernels' design (B200): tensor-memory read bandwidth per SM, MUFU ex2 rate per
// scheduler, and an FMA-pipe exp2 next to it.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_mufu tmem_mufu.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(slot));
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA / ALU pipes: round-to-nearest split (magic-number add), degree-3 minimax on [-0.5, 0.5], exponent add
__device__ __forceinline__ float ex2_fma(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.05550357f, 0.24022650f);
  p = fmaf(p, f, 0.69314718f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// mode 0: TMEM loads only; 1: ex2 only; 2: ex2_fma only; 3: ex2 + TMEM loads (the softmax pass); 4: st only
template <int mode, int mix>
__global__ void __launch_bounds__(512, 1) bench(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 256);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t trow = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t a[32], b[32];
  float acc = threadIdx.x * 1e-9f;
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0) {
    uint32_t x = 0;
    ld32(trow, a);
    for (int i = 0; i < iters; ++i) {
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      ld32(trow + 32 * ((2 * i + 1) & 7), b);
#pragma unroll
      for (int k = 0; k < 32; ++k) x ^= a[k];
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      ld32(trow + 32 * ((2 * i + 2) & 7), a);
#pragma unroll
      for (int k = 0; k < 32; ++k) x ^= b[k];
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc += __uint_as_float(x & 0xff);
  } else if (mode == 1 || mode == 2) {
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = -0.01f * (k + threadIdx.x);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const bool fma_path = mode == 2 || (mix > 0 && (k % mix) == 0 && mode == 1);
        v[k] = fma_path ? ex2_fma(v[k] - 1.f) : ex2(v[k] - 1.f);
      }
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) acc += v[k];
  } else if (mode == 3) {
    float sum = 0.f;
    ld32(trow, a);
    for (int i = 0; i < iters; ++i) {
      uint32_t pk[16];
      asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(a[0]), "+r"(a[31])::"memory");
      ld32(trow + 32 * ((2 * i + 1) & 7), b);
#pragma unroll
      for (int k = 0; k < 32; k += 2) {
        const bool f0 = mix > 0 && (k % mix) == 0;
        const float p0 = f0 ? ex2_fma(fmaf(__uint_as_float(a[k]), 0.18f, -3.f)) : ex2(fmaf(__uint_as_float(a[k]), 0.18f, -3.f));
        const float p1 = ex2(fmaf(__uint_as_float(a[k + 1]), 0.18f, -3.f));
        sum += p0 + p1;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[k >> 1]) : "f"(p1), "f"(p0));
      }
      st16(trow + 128 + 16 * (i & 3), pk);
      asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(b[0]), "+r"(b[31])::"memory");
      ld32(trow + 32 * ((2 * i + 2) & 7), a);
#pragma unroll
      for (int k = 0; k < 32; k += 2) {
        const bool f0 = mix > 0 && (k % mix) == 0;
        const float p0 = f0 ? ex2_fma(fmaf(__uint_as_float(b[k]), 0.18f, -3.f)) : ex2(fmaf(__uint_as_float(b[k]), 0.18f, -3.f));
        const float p1 = ex2(fmaf(__uint_as_float(b[k + 1]), 0.18f, -3.f));
        sum += p0 + p1;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[k >> 1]) : "f"(p1), "f"(p0));
      }
      st16(trow + 128 + 64 + 16 * (i & 3), pk);
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    acc += sum;
  } else if (mode == 5 || mode == 6 || mode == 7) {
    // 5: cvt.rn.bf16x2.f32 alone; 6: ex2 pair + cvt (the softmax inner step without tensor memory); 7: ex2 pair +
    // integer round-half-up + prmt pack
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = -0.01f * (k + threadIdx.x);
    uint32_t x = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 32; k += 2) {
        float p0 = v[k], p1 = v[k + 1];
        if (mode != 5) {
          p0 = ex2(p0 - 1.f);
          p1 = ex2(p1 - 1.f);
        }
        uint32_t pk;
        if (mode == 7) {
          const uint32_t u0 = __float_as_uint(p0) + 0x8000u, u1 = __float_as_uint(p1) + 0x8000u;
          asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(pk) : "r"(u0), "r"(u1));
        } else {
          asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(p1), "f"(p0));
        }
        x ^= pk;
        v[k] = mode == 5 ? __uint_as_float(pk | 0x3f000000u) * 0.5f : p0;
        v[k + 1] = p1 - 2.f;
      }
    }
    acc += __uint_as_float(x & 0xff);
#pragma unroll
    for (int k = 0; k < 32; ++k) acc += v[k];
  } else if (mode == 8 || mode == 9) {
    // 8: rcp.approx.ftz.f32 alone; 9: FFMA2-only Horner chains (packed fp32 pairs), 32 ops = 16 instructions
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = 1.0f + 0.01f * (k + threadIdx.x);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 32; k += 2) {
        if (mode == 8) {
          asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(v[k]) : "f"(v[k] + 1.f));
          asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(v[k + 1]) : "f"(v[k + 1] + 1.f));
        } else {
          unsigned long long a, b = 0x3f8000003f800000ull, c = 0x3a83126f3a83126full;
          asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(v[k]), "f"(v[k + 1]));
          asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(a) : "l"(a), "l"(b), "l"(c));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(v[k]), "=f"(v[k + 1]) : "l"(a));
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) acc += v[k];
  } else if (mode == 4) {
    uint32_t pk[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) pk[k] = threadIdx.x + k;
    for (int i = 0; i < 2 * iters; ++i) st16(trow + 16 * (i & 15), pk);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 16 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 256);
}

template <int mode, int mix>
void go(int g, int b, int iters, long long* cyc, float* sink) { bench<mode, mix><<<g, b>>>(iters, cyc, sink); }
void launch(int mode, int mix, int g, int b, int iters, long long* cyc, float* sink) {
#define CASE(m, x) if (mode == m && mix == x) return go<m, x>(g, b, iters, cyc, sink);
  CASE(0, 0) CASE(1, 0) CASE(1, 4) CASE(1, 2) CASE(2, 0) CASE(3, 0) CASE(3, 4) CASE(3, 2) CASE(4, 0) CASE(5, 0) CASE(6, 0) CASE(7, 0) CASE(8, 0) CASE(9, 0)
}
int main() {
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, 148 * 2 * 16 * 8);
  cudaMalloc(&sink, 148 * 2 * 512 * 4);
  long long h[148 * 2 * 16];
  const int iters = 256;
  const char* names[] = {"tmem ld x32 (4 KB per warp-load)", "mufu ex2", "fma exp2", "softmax pass: ld + ex2 + cvt + st", "tmem st x16", "cvt.rn.bf16x2 (16 per 32 ops)", "ex2 + cvt.rn.bf16x2", "ex2 + iadd/prmt pack", "mufu rcp", "fma.rn.f32x2 (two elements per op)"};
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int warps : {4, 8, 16}) {
      if (ctas == 2 && warps == 16) continue;
      for (int mode = 0; mode < 10; ++mode)
        for (int mix : {0, 4, 2}) {
          if (mix && mode != 1 && mode != 3) continue;
          for (int rep = 0; rep < 2; ++rep) {
            launch(mode, mix, 148 * ctas, warps * 32, iters, cyc, sink);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          }
          cudaMemcpy(h, cyc, sizeof(long long) * 148 * ctas * 16, cudaMemcpyDeviceToHost);
          double mean = 0;
          for (int c = 0; c < 148 * ctas; ++c)
            for (int w = 0; w < warps; ++w) mean += h[c * 16 + w];
          mean /= 148 * ctas * warps;
          const double per_warp_op = mean / (mode == 0 || mode == 3 || mode == 4 ? 2.0 * iters : 32.0 * iters);
          if (mode == 0 || mode == 3)
            printf("%d CTA/SM x %2d warps  %-36s mix 1/%d: %8.1f cyc per 32-column chunk per warp -> %6.1f B/clk/SM read, %5.2f cyc per warp-wide exp\n",
                   ctas, warps, names[mode], mix, per_warp_op, 4096.0 * warps * ctas / per_warp_op, per_warp_op / 32);
          else if (mode == 4)
            printf("%d CTA/SM x %2d warps  %-36s        : %8.1f cyc per 16-column store per warp -> %6.1f B/clk/SM written\n", ctas, warps,
                   names[mode], per_warp_op, 2048.0 * warps * ctas / per_warp_op);
          else
            printf("%d CTA/SM x %2d warps  %-36s mix 1/%d: %8.2f cyc per warp-wide op (%.2f per scheduler-op with %d warps/scheduler)\n", ctas,
                   warps, names[mode], mix, per_warp_op, per_warp_op / (warps * ctas / 4.0), warps * ctas / 4);
        }
    }
  return 0;
}

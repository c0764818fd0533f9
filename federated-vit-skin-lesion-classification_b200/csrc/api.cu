// api.cu — library-wide state of libfedvit: error string, launch counter, device queries.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace fv {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::atomic<int64_t> g_family[FV_KERNEL_FAMILIES];
void count_kernel(int family) {
  if (family >= 0 && family < FV_KERNEL_FAMILIES) g_family[family].fetch_add(1, std::memory_order_relaxed);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FEDVIT_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

}  // namespace fv

extern "C" int fv_version(void) { return 100; }
extern "C" const char* fv_last_error(void) { return fv::g_error; }
extern "C" int64_t fv_launch_count(void) { return fv::g_launches.load(std::memory_order_relaxed); }
extern "C" int64_t fv_kernel_launches(int family) {
  if (family < 0 || family >= FV_KERNEL_FAMILIES) return -1;
  return fv::g_family[family].load(std::memory_order_relaxed);
}

// Host-side runtime of libuc_b200: error slots, launch accounting, TMA descriptor encoding.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <atomic>

#include "common.cuh"

namespace uc {

static thread_local char g_err[512] = {0};
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return UC_ERR_CUDA;
  }
  return UC_OK;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
  static encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(p);
    else
      (void)cudaGetLastError();
  }
  return fn;
}

int make_tensor_map(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  encode_tiled_fn fn = get_encode();
  UC_REQUIRE(fn != nullptr, UC_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver / GPU)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = fn(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UC_REQUIRE(r == CUDA_SUCCESS, UC_ERR_CUDA,
             "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu] stride0 %llu box [%u,%u,%u] base %p",
             (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0],
             rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, base);
  return UC_OK;
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("UC_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

// Work counters of the persistent kernels with dynamic tile / item queues (pair GEMM, attention): 2048 slots of
// (next item, CTAs or clusters finished), zero at load and reset by the kernels themselves when their last CTA leaves.
// Launches take slots round robin, so two launches that can overlap in time (different streams, programmatic dependent
// launch) never share one, and a captured CUDA graph replays with the slots it was captured with.
__device__ int g_work_slots[4096];
int* work_slot() {
  static int* base = nullptr;
  static unsigned seq = 0;
  if (!base) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, g_work_slots) != cudaSuccess) {
      (void)cudaGetLastError();
      return nullptr;
    }
    base = static_cast<int*>(p);
  }
  return base + 2 * (__atomic_fetch_add(&seq, 1u, __ATOMIC_RELAXED) % 2048u);
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      (void)cudaGetLastError();
      n = 0;
      return 148;
    }
  }
  return n;
}

}  // namespace uc

extern "C" {
int uc_version(void) { return 100; }
size_t uc_last_error(char* buf, size_t cap) {
  size_t n = strlen(uc::g_err);
  if (buf && cap) {
    size_t c = n < cap - 1 ? n : cap - 1;
    memcpy(buf, uc::g_err, c);
    buf[c] = 0;
  }
  return n;
}
uint64_t uc_launch_count(void) { return uc::g_launches.load(); }
}

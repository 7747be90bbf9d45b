// Host plumbing: error slot, launch counter, device query, TMA descriptor encoding.
#include "cs_host.h"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace cs {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return CS_ERR_CUDA;
}
const char* last_error() { return g_err; }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(); }
void reset_launch_count() { g_launches.store(0); }

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides, const uint32_t* box, const uint32_t* estr) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(CS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = estr[i]; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, s,
                  b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[256];
    snprintf(msg, sizeof(msg),
             "cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,..] box=[%u,%u,..]", (int)r,
             rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return set_error(CS_ERR_CUDA, msg);
  }
  return CS_OK;
}

}  // namespace cs

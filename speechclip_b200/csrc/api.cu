// C ABI entry points + host-side plumbing (error strings, tensor-map encoding, device queries).
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"
#include "ops.cuh"

namespace scb {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return SCB_ECUDA;
}

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_16b(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int swizzle128) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    SCB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    SCB_CHECK(p != nullptr && q == cudaDriverEntryPointSuccess, SCB_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  cuuint64_t d[5];
  cuuint64_t s[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  // The element type only matters for OOB fill / arithmetic; both 16-bit formats move as raw 2-byte words.
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SCB_CHECK(r == CUDA_SUCCESS, SCB_ECUDA,
            "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu] strides [%llu,%llu] box [%u,%u,%u] base %p", (int)r,
            rank, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0), (unsigned long long)(rank > 2 ? d[2] : 0),
            (unsigned long long)s[0], (unsigned long long)(rank > 2 ? s[1] : 0), b[0], rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0, base);
  return SCB_OK;
}

}  // namespace scb

extern "C" {

int scb_abi_version(void) { return SCB_ABI_VERSION; }
const char* scb_last_error(void) { return scb::g_err; }
int64_t scb_launch_count(void) { return scb::g_launches.load(); }

int scb_gemm(const scb_gemm_args* args, void* stream) {
  if (!args) {
    scb::set_error("scb_gemm: args is NULL");
    return SCB_EINVAL;
  }
  return scb::gemm(*args, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

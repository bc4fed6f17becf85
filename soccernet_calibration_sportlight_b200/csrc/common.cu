#include <stdlib.h>
// common.cu - error string, ABI version, tensor-map encoding.
#include <stdarg.h>
#include <string.h>

#define CAL_TU "common.cu"
#include "common.cuh"

namespace cal {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// Shared memory the persistent tcgen05 kernels leave free on every SM so that a camera_solve_kernel
// block (one frame, ~39 KB + 64 threads x 168 registers) can be co-resident with them: the solve of
// batch i then runs UNDER the networks of batch i+1 instead of after them (pipeline.py).
// CAL_SMEM_HEADROOM (bytes) or cal_set_smem_headroom(); 0 = the kernels take all they can use.
static int g_smem_headroom = [] { const char* e = getenv("CAL_SMEM_HEADROOM"); return e ? atoi(e) : 0; }();
int smem_headroom() { return g_smem_headroom; }

bool pdl_enabled() {
  // opt-in: measured on B200 at 1.5-3 % per back-to-back 3x3 launch (80 -> 78 us), within run-to-run noise
  static const bool on = [] { const char* e = getenv("CAL_PDL"); return e && e[0] == '1'; }();
  return on;
}

int encode_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estrides,
                    CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return CAL_E_CUDA;
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = estrides ? estrides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d,
                  s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu] box "
              "[%u %u %u %u] estr [%u %u %u %u]",
              (int)r, rank, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0),
              (unsigned long long)(rank > 2 ? d[2] : 0), (unsigned long long)(rank > 3 ? d[3] : 0),
              b[0], rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0, rank > 3 ? b[3] : 0, es[0],
              rank > 1 ? es[1] : 0, rank > 2 ? es[2] : 0, rank > 3 ? es[3] : 0);
    return CAL_E_CUDA;
  }
  return CAL_OK;
}

}  // namespace cal

extern "C" int cal_abi_version(void) { return CAL_ABI_VERSION; }
extern "C" int cal_set_smem_headroom(int bytes) {
  CAL_REQUIRE(bytes >= 0 && bytes <= 96 * 1024, CAL_E_INVALID, "cal_set_smem_headroom: %d bytes (0..98304)", bytes);
  cal::g_smem_headroom = bytes;
  return CAL_OK;
}
extern "C" const char* cal_last_error(void) { return cal::g_err; }

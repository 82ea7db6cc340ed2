#include "host_util.h"

#include <mutex>
#include <stdarg.h>
#include <stdlib.h>

namespace pst3r {

static thread_local char g_err[1024] = {0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_last_error() { return g_err; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return PST3R_ERR_DRIVER;
  }
  if (rank < 2 || rank > 5) {
    set_last_error("encode_tmap: rank %d unsupported", rank);
    return PST3R_ERR_INVALID;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_last_error("encode_tmap: base pointer %p not 16-byte aligned", base);
    return PST3R_ERR_INVALID;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i >= 1) {
      gstr[i - 1] = strides_bytes[i];
      if (strides_bytes[i] % 16 != 0) {
        set_last_error("encode_tmap: stride[%d]=%llu bytes is not a multiple of 16", i,
                       (unsigned long long)strides_bytes[i]);
        return PST3R_ERR_INVALID;
      }
    }
  }
  CUtensorMapDataType dt;
  switch (elem_bytes) {
    case 2: dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; break;
    case 4: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; break;
    case 1: dt = CU_TENSOR_MAP_DATA_TYPE_UINT8; break;
    default: set_last_error("encode_tmap: elem_bytes %d unsupported", elem_bytes); return PST3R_ERR_INVALID;
  }
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
                   (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return PST3R_ERR_DRIVER;
  }
  return PST3R_OK;
}

static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("PST3R_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl == 1;
}
int set_pdl(int on) {
  const int prev = pdl_enabled() ? 1 : 0;
  g_pdl = on ? 1 : 0;
  return prev;
}

static int g_split_k = 0;
bool split_k_enabled() { return g_split_k == 1; }
int set_split_k(int on) {
  const int prev = g_split_k;
  g_split_k = on ? 1 : 0;
  return prev;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
    n = p.multiProcessorCount;
  }
  return n;
}

static int g_sm_budget = 0;
int sm_budget() {
  const int n = num_sms();
  return (g_sm_budget > 0 && g_sm_budget < n) ? g_sm_budget : n;
}
void set_sm_budget(int n) { g_sm_budget = n > 0 ? n : 0; }

}  // namespace pst3r

extern "C" int pst3r_set_sm_budget(int32_t n) {
  const int prev = pst3r::sm_budget();
  pst3r::set_sm_budget(n);
  return prev;
}
extern "C" int pst3r_num_sms(void) { return pst3r::num_sms(); }
extern "C" int pst3r_set_pdl(int32_t on) { return pst3r::set_pdl(on); }
extern "C" int pst3r_set_split_k(int32_t on) { return pst3r::set_split_k(on); }

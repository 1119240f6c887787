// util.cu — error string, device queries, TMA descriptor encoding.
#include "fav_common.cuh"
#include <stdlib.h>

#include <stdarg.h>
#include <mutex>
#include <vector>

namespace fav {

static thread_local char t_err[1024] = "";
unsigned long long g_launch_count = 0;

bool g_prof_on = false;
bool g_pdl_on = false;   // set per API call from the handle (fav_api.cu): on for the torchvision nets, off for I3D (measured)
namespace {
struct ProfRec { int kind; cudaEvent_t e0, e1; double flops, bytes; };
std::vector<ProfRec> g_prof_recs;
cudaEvent_t g_prof_open = nullptr;
}  // namespace
void prof_record(int kind, cudaStream_t s, bool end, double flops, double bytes) {
  if (!end) {
    cudaEventCreate(&g_prof_open);
    cudaEventRecord(g_prof_open, s);
    return;
  }
  ProfRec r;
  r.kind = kind; r.e0 = g_prof_open; r.flops = flops; r.bytes = bytes;
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e1, s);
  g_prof_recs.push_back(r);
  g_prof_open = nullptr;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return t_err; }

int sm_count(int device) {
  static int cached[16] = {0};
  if (device >= 0 && device < 16 && cached[device]) return cached[device];
  int n = 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) n = 148;
  if (device >= 0 && device < 16) cached[device] = n;
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz,
                   const uint32_t* elem_strides) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return FAV_ERR_CUDA;
  }
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                  const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu,%llu] "
              "strides [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u,%u] base %p",
              static_cast<int>(r), rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
              (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0),
              (unsigned long long)(rank > 4 ? gd[4] : 0), (unsigned long long)(rank > 1 ? gs[0] : 0),
              (unsigned long long)(rank > 2 ? gs[1] : 0), (unsigned long long)(rank > 3 ? gs[2] : 0),
              (unsigned long long)(rank > 4 ? gs[3] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0,
              rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0, base);
    return FAV_ERR_CUDA;
  }
  return FAV_OK;
}

}  // namespace fav

extern "C" int fav_profile_begin(void) {
  fav::g_prof_recs.clear();
  fav::g_prof_on = true;
  return 0;
}
// out[kind*4 + {0: ms, 1: launches, 2: algorithmic flops, 3: algorithmic bytes}], kinds in fav.h order
extern "C" int fav_profile_end(double* out, int capacity) {
  fav::g_prof_on = false;
  cudaDeviceSynchronize();
  for (int i = 0; i < capacity; ++i) out[i] = 0.0;
  for (auto& r : fav::g_prof_recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (r.kind * 4 + 3 < capacity) {
      out[r.kind * 4 + 0] += ms;
      out[r.kind * 4 + 1] += 1.0;
      out[r.kind * 4 + 2] += r.flops;
      out[r.kind * 4 + 3] += r.bytes;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  fav::g_prof_recs.clear();
  return 0;
}

extern "C" const char* fav_last_error(void) { return fav::get_error(); }

extern "C" int64_t fav_launch_count(void) { return static_cast<int64_t>(fav::g_launch_count); }

extern "C" const char* fav_build_info(void) { return "sm_100a;" __DATE__ " " __TIME__; }

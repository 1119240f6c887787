// fav_common.cuh — shared helpers: error plumbing, bf16 packing, sm_100a PTX wrappers
// (mbarrier, TMA, tcgen05).  Everything here is internal to libfav.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <utility>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/fav.h"

namespace fav {

// ---- error plumbing ------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();
#define FAV_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      fav::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return FAV_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)
#define FAV_CHECK_ARG(cond, ...)                                                              \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      fav::set_error(__VA_ARGS__);                                                            \
      return FAV_ERR_ARG;                                                                     \
    }                                                                                         \
  } while (0)
#define FAV_TRY(expr)                                                                         \
  do {                                                                                        \
    int _s = (expr);                                                                          \
    if (_s != FAV_OK) return _s;                                                              \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

int sm_count(int device);
// number of kernels this library has launched in this process (bench.py reports it)
extern unsigned long long g_launch_count;
#define FAV_COUNT_LAUNCH() (++fav::g_launch_count)

// ---- per-kernel-family timing with CUDA events on the launching stream (fav_profile_begin / _end) ----------
enum ProfKind {
  PK_APPLY = 0, PK_STEM, PK_CONV_HALO, PK_CONV_TAP, PK_POOL_FWD, PK_POOL_BWD, PK_HEAD_LOSS, PK_STEM_BWD,
  PK_DELTA_UPDATE, PK_OTHER, PK_COUNT
};
extern bool g_prof_on;
void prof_record(int kind, cudaStream_t s, bool end, double flops, double bytes);
struct ProfScope {
  int kind; cudaStream_t s; double flops, bytes;
  ProfScope(int k, cudaStream_t st, double fl = 0.0, double by = 0.0) : kind(k), s(st), flops(fl), bytes(by) {
    if (g_prof_on) prof_record(kind, s, false, 0.0, 0.0);
  }
  ~ProfScope() { if (g_prof_on) prof_record(kind, s, true, flops, bytes); }
};

// ---- TMA descriptor creation (driver entry point fetched at run time: no libcuda link) -------
// rank-5 bf16 tensor map; dims/strides innermost first; strides in BYTES for dims 1..4.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz,
                   const uint32_t* elem_strides = nullptr);

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// Kernels of the step are launched with cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel's CTAs may
// become resident (and run their prologue: barrier init, TMEM allocation, descriptor prefetch) as soon as every CTA of
// the previous kernel has *started* and an SM has room, instead of after the previous grid has drained.  Each such
// kernel calls pdl_sync() — griddepcontrol.launch_dependents + griddepcontrol.wait — before its first global-memory
// access; the wait returns when all prerequisite grids have completed and flushed, so the data flow is unchanged.
// FAV_PDL=0 launches without the attribute (the device instructions are then no-ops).
extern bool g_pdl_on;
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl_on ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

#ifdef __CUDACC__
// ---- device helpers --------------------------------------------------------------------------
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// ---- 16-bit storage formats ------------------------------------------------------------------
// FORWARD activations and forward weights are IEEE fp16 (10 mantissa bits: 8x less rounding noise than bf16, which is
// what decides how many ReLU masks / pooling arg-maxes differ from the reference's fp32 run; tcgen05
// kind::f16 runs fp16 and bf16 operands at the same rate).  GRADIENTS and the data-gradient weights stay bf16 (range).
// fp32 -> fp16 saturates to +-65504 instead of overflowing to inf.
// One F2FP instruction per pair: round to nearest even, saturate to the largest finite value, optionally ReLU first
// (ncu: the 1x1x1 conv launches are bound by the instruction count of their epilogue).
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t pack_f16x2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ float f16_lo(uint32_t v) { return __low2float(*reinterpret_cast<const __half2*>(&v)); }
__device__ __forceinline__ float f16_hi(uint32_t v) { return __high2float(*reinterpret_cast<const __half2*>(&v)); }
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi, bool f16) {
  return f16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
}
__device__ __forceinline__ float lo16(uint32_t v, bool f16) { return f16 ? f16_lo(v) : bf16_lo(v); }
__device__ __forceinline__ float hi16(uint32_t v, bool f16) { return f16 ? f16_hi(v) : bf16_hi(v); }
// ReLU mask of a pair of forward activations (fp16): 0xffff per half whose value is > 0
__device__ __forceinline__ uint32_t relu_mask2(uint32_t act2) {
  return __hgt2_mask(*reinterpret_cast<const __half2*>(&act2), __float2half2_rn(0.0f));
}
constexpr uint32_t kF16One2 = 0x3c003c00u;    // fp16x2 (1, 1): the "no mask" operand
constexpr uint32_t kF16NegInf2 = 0xfc00fc00u; // fp16x2 (-inf, -inf)

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// Global flag set when a bounded wait times out (read back by the host in debug paths).
extern __device__ int g_fav_timeout_flag;

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must surface as an error, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
      g_fav_timeout_flag = 1;
      __threadfence_system();
      __trap();
    }
  }
}

// Spinning wait on test_wait (non-blocking probe): lower wake-up latency than the potentially-suspending try_wait,
// for hand-offs that sit on a kernel's critical path.  Bounded like mbar_wait.
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t it = 0;; ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (it > 200000000u) {
      g_fav_timeout_flag = 1;
      __threadfence_system();
      __trap();
    }
  }
}

// ---- TMA -----------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1)
      : "memory");
}

// 1-D bulk copy global -> shared (16-byte aligned addresses, size a multiple of 16), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- tcgen05 / TMEM --------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]   (kind::f16: bf16 x bf16 or fp16 x fp16 -> f32; the instruction descriptor says which)
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, K-major operand, rows of `row_bytes` (32/64/128) bytes holding
// row_bytes/2 bf16 of K, hardware swizzle == row_bytes, 8-row core groups `sbo` bytes apart.
// Field layout: cute/arch/mma_sm100_desc.hpp (SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) (2=SW128, 4=SW64, 6=SW32).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t row_bytes) {
  uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
  uint64_t sbo = (8u * row_bytes) >> 4;
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3ffff) >> 4);
  d |= 1ull << 16;   // LBO (ignored for swizzled K-major; CUTLASS writes 1)
  d |= sbo << 32;
  d |= 1ull << 46;   // descriptor version (Blackwell)
  d |= layout << 61;
  return d;
}
// Split form for tight issue loops: the high word is constant per layout, the low word is
// (address >> 4) | LBO.  K-advance / row-advance are plain integer adds on the low word.
__device__ __forceinline__ uint32_t umma_desc_hi(uint32_t row_bytes) {
  const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  return ((8u * row_bytes) >> 4) | (1u << 14) | (layout << 29);
}
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr & 0x3ffffu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t make_desc(uint32_t hi, uint32_t lo) {
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
// NO-SWIZZLE K-major descriptor (canonical interleaved layout): element (r, k) of the operand sits at
//   start + 16*(r % 8) + SBO*(r / 8) + 2*(k % 8) + LBO*(k / 8)   bytes.
// With LBO = 16 the eight rows of a group are 16-byte-SHIFTED, overlapping 32-byte windows of the same raw bytes — the
// W-direction im2col of a stride-2 convolution over 8-byte RGBX pixels (output column w reads pixels 2w .. 2w+7) comes
// for free from raw input rows (verified on B200 by tools/ubench/umma_overlap.cu).  high word: SBO, version, layout 0.
__device__ __forceinline__ uint32_t umma_desc_hi_nosw(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14); }

// Same, for an operand whose start is 128-byte aligned but not 1024-byte aligned (a row window of a
// larger swizzled slab): base_offset (bits [49,52)) carries the phase of the swizzle pattern.
__device__ __forceinline__ uint64_t umma_smem_desc_sw128_off(uint32_t saddr, uint32_t use_base_offset) {
  uint64_t d = umma_smem_desc(saddr, 128);
  if (use_base_offset) d |= static_cast<uint64_t>((saddr >> 7) & 7u) << 49;
  return d;
}
// Instruction descriptor for kind::f16, D=f32, both operands K-major (InstrDescriptor): a_format / b_format
// 0 = F16, 1 = BF16 (forward convs run fp16 x fp16, data gradients bf16 x bf16).
__host__ __device__ __forceinline__ uint32_t umma_idesc(int M, int N, bool f16) {
  uint32_t d = 0;
  d |= 1u << 4;                         // c_format = F32
  if (!f16) {
    d |= 1u << 7;                       // a_format = BF16
    d |= 1u << 10;                      // b_format = BF16
  }
  d |= static_cast<uint32_t>(N >> 3) << 17;
  d |= static_cast<uint32_t>(M >> 4) << 24;
  return d;
}
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N) { return umma_idesc(M, N, false); }
#endif  // __CUDACC__

}  // namespace fav

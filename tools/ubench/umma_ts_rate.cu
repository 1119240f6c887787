// Cost of the temporal-sharing stem's MMAs (conv_stem_ts_kernel): 128xNx16, fp16, A = NO-SWIZZLE overlapping row windows
// (LBO = 16 B, SBO = pitch), B = SW64 K-major rows of 64 B, two K halves per (frame, kh) pair, A / B / D addresses
// rotating the way the kernel's loop rotates them.  Compared with the plain SW128 / SW128 case of umma_rate.cu.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=true -I../../flickering_adversarial_video_b200/csrc umma_ts_rate.cu -o umma_ts_rate
#include "fav_common.cuh"
#include <vector>
namespace fav { __device__ int g_fav_timeout_flag = 0; }
using namespace fav;

// mode 0: A no-swizzle windows / B SW64;  1: A SW128 / B SW64 (32-byte K slice);  2: A no-swizzle / B SW128;  3: SW128 / SW128
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int mode, int pitch, int rotate, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc(128, N, true);
    const bool a_nosw = mode == 0 || mode == 2;
    const bool b_sw64 = mode == 0 || mode == 1;
    const uint32_t hi_a = a_nosw ? umma_desc_hi_nosw(pitch) : umma_desc_hi(128);
    const uint32_t hi_b = b_sw64 ? umma_desc_hi(64) : umma_desc_hi(128);
    const uint32_t a0 = umma_desc_lo(smem_u32(smem));            // A region: 64 KB (8 frame slots of 8 KB)
    const uint32_t b0 = umma_desc_lo(smem_u32(smem + 65536));    // B region: 32 KB
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t a_lo = a0 + (rotate ? (static_cast<uint32_t>(i & 7) * 8192u) >> 4 : 0u);
      const uint32_t d = tm + (rotate ? static_cast<uint32_t>((i & 1) * 256) : 0u);
      umma_bf16(d, make_desc(hi_a, a_lo), make_desc(hi_b, b0), idesc, 1u);
      umma_bf16(d, make_desc(hi_a, a_lo + 2), make_desc(hi_b, b0 + 2), idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  const int iters = 2000;
  const char* names[4] = {"A no-swizzle windows / B SW64", "A SW128 / B SW64", "A no-swizzle windows / B SW128", "A SW128 / B SW128"};
  for (int mode = 0; mode < 4; ++mode)
    for (int pitch : {176, 320})
      for (int rotate = 0; rotate < 2; ++rotate) {
        if ((mode == 1 || mode == 3) && pitch != 176) continue;
        printf("%s, pitch %d, %s:", names[mode], pitch, rotate ? "rotating A / D" : "fixed A / D");
        for (int N : {64, 128, 192, 256}) {
          rate_kernel<<<148, 128, 100 * 1024 + 1024>>>(N, iters, mode, pitch, rotate, d);
          cudaError_t e = cudaDeviceSynchronize();
          std::vector<long long> h(148);
          cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
          printf("  N=%d: %.1f clk%s", N, double(h[0]) / (iters * 2.0), e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
        printf("\n");
      }
  return 0;
}

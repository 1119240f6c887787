// Measures the issue rate of tcgen05.mma 128xNx16 (bf16, SS mode, SW128 K-major operands) for several N.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../flickering_adversarial_video_b200/csrc umma_rate.cu -o umma_rate
#include "fav_common.cuh"
#include <vector>
namespace fav { __device__ int g_fav_timeout_flag = 0; }
using namespace fav;

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int swz128, int a_off_rows, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t rb = swz128 ? 128u : 64u;
    const uint32_t hi = umma_desc_hi(rb);
    const uint32_t a_lo = umma_desc_lo(smem_u32(smem) + a_off_rows * rb);
    const uint32_t b_lo = umma_desc_lo(smem_u32(smem + 16384));
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tm, make_desc(hi, a_lo + 2 * (swz128 ? k : (k & 1))), make_desc(hi, b_lo + 2 * (swz128 ? k : (k & 1))), idesc, 1u);
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
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  for (int N = 16; N <= 256; N += 16) {   // N sweep, aligned A, SW128
    rate_kernel<<<148, 128, 50 * 1024 + 1024>>>(N, iters, 1, 0, d);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
    printf("N sweep: N=%3d: %.1f clk per 128xNx16 MMA (N/2 = %d)  [%s]\n", N, double(h[0]) / (iters * 4.0), N / 2, cudaGetErrorString(e));
  }
  for (int off : {0, 1, 4, 8})
  for (int swz = 1; swz >= 0; --swz)
    for (int N : {64, 128, 192, 256}) {
      rate_kernel<<<148, 128, 50 * 1024 + 1024>>>(N, iters, swz, off, d);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<long long> h(148);
      cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
      printf("A row offset %d, swz%d N=%3d: %.1f clk per 128xNx16 MMA (floor N/2 = %d)  [%s]\n", off, swz ? 128 : 64, N,
             double(h[0]) / (iters * 4.0), N / 2, cudaGetErrorString(e));
    }
  return 0;
}

// Which part of a realistic warp-uniform issue loop slows tcgen05.mma below the back-to-back rate?
// FLAGS: 1 commit per batch; 2 tcgen05.fence::after per batch; 4 mbar_wait on an (already complete) barrier per batch;
//        8 wait for the commit issued 4 batches ago (ring of 8); 16 rotate operand buffers
#include "fav_common.cuh"
#include <vector>
namespace fav { __device__ int g_fav_timeout_flag = 0; }
using namespace fav;

template <int FLAGS>
__global__ void __launch_bounds__(128, 1) loop_kernel(int N, int iters, int per_batch, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar, junk[8], ready;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (196608) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1); mbar_init(&ready, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&junk[i], 1);
    mbar_fence_init();
    mbar_arrive(&ready);   // phase 0 complete
  }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t hi = umma_desc_hi(128);
    long long t0 = clock64();
    int jb = 0, rot = 0;
    uint32_t jphase = 0;
    for (int i = 0; i < iters; ++i) {
      if (FLAGS & 4) mbar_wait(&ready, 0);
      if (FLAGS & 2) tc_fence_after();
      const uint32_t base = smem_u32(smem) + ((FLAGS & 16) ? rot * 32768u : 0u);
      const uint32_t a_lo = umma_desc_lo(base), b_lo = umma_desc_lo(base + 16384);
      if (elect_one()) {
        uint32_t d_i = tm;
        for (int m = 0; m < per_batch; ++m) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_i, make_desc(hi, a_lo + 2 * k), make_desc(hi, b_lo + 2 * k), idesc, 1u);
          d_i += N;
        }
        if (FLAGS & 1) umma_commit(&junk[jb]);
      }
      __syncwarp();
      if (FLAGS & 8) {
        const int wb = (jb + 4) & 7;
        if (i >= 4) mbar_wait(&junk[wb], (wb > jb ? jphase ^ 1 : jphase));
      }
      if (++jb == 8) { jb = 0; jphase ^= 1; }
      if (++rot == 6) rot = 0;
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int FLAGS>
void run(int N, int pb, long long* d) {
  const int iters = 2000;
  cudaFuncSetAttribute(loop_kernel<FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  loop_kernel<FLAGS><<<148, 128, 196608 + 1024>>>(N, iters, pb, d);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
  printf("N=%3d batch=%d flags=%2d: %.1f clk/MMA [%s]\n", N, pb * 4, FLAGS, double(h[0]) / (iters * 4.0 * pb), cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  for (int N : {64, 128, 192})
    for (int pb : {1, 2}) {
      if (pb * N > 512) continue;
      run<0>(N, pb, d); run<1>(N, pb, d); run<2>(N, pb, d); run<4>(N, pb, d); run<7>(N, pb, d); run<15>(N, pb, d);
      run<16>(N, pb, d); run<31>(N, pb, d);
    }
  return 0;
}

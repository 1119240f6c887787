// L2 -> shared memory fill rate per SM with all SMs pulling the same L2-resident weight set:
//   (a) cp.async.bulk.tensor.2d boxes of `rows` x 128 B (SWIZZLE_128B) — how the conv kernels fetch weight tiles,
//   (b) cp.async.bulk (1-D, contiguous) of the same number of bytes — what a host-side pre-swizzled weight image allows.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=true -I../../flickering_adversarial_video_b200/csrc tma_rate.cu ../../flickering_adversarial_video_b200/csrc/util.cu -o tma_rate
#include "fav_common.cuh"
#include <vector>
#include <cstdio>
#include <algorithm>
namespace fav { __device__ int g_fav_timeout_flag = 0; }
using namespace fav;

constexpr int kRingMax = 32;

__global__ void __launch_bounds__(128, 1) rate_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* flat, int rows,
                                                      int inner_bytes, int ntiles_total, int iters, int mode, int ring, int nissuers, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t full_all[4 * kRingMax];
  const int tile_bytes = rows * inner_bytes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4 * kRingMax; ++s) mbar_init(&full_all[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && wid < nissuers) {
    smem += wid * ring * tile_bytes;
    uint64_t* full = full_all + wid * kRingMax;
    long long t0 = clock64();
    long long t_issue = 0, t_wait = 0;
    uint32_t phase_bits = 0;
    int issued = 0, done = 0;
    int tile = blockIdx.x * 7 + wid * 13;
    while (done < iters) {
      while (issued < iters && issued - done < ring) {
        const int s = issued % ring;
        const long long c0 = clock64();
        mbar_expect_tx(&full[s], static_cast<uint32_t>(tile_bytes));
        tile = (tile + 1) % ntiles_total;
        if (mode == 0) tma_load_2d(smem + s * tile_bytes, &tm, &full[s], 0, tile * rows);
        else bulk_load_1d(smem + s * tile_bytes, flat + static_cast<size_t>(tile) * tile_bytes, tile_bytes, &full[s]);
        ++issued;
        t_issue += clock64() - c0;
      }
      const int s = done % ring;
      const long long c1 = clock64();
      mbar_wait(&full[s], (phase_bits >> s) & 1u);
      t_wait += clock64() - c1;
      phase_bits ^= 1u << s;
      ++done;
    }
    if (wid == 0) { out[blockIdx.x] = clock64() - t0; out[148 + blockIdx.x] = t_issue; out[296 + blockIdx.x] = t_wait; }
  }
}

int main() {
  const int total_rows = 16384;            // 2 MB of 128-byte rows: L2 resident
  uint8_t* d;
  cudaMalloc(&d, static_cast<size_t>(total_rows) * 128);
  cudaMemset(d, 1, static_cast<size_t>(total_rows) * 128);
  long long* dout;
  cudaMalloc(&dout, 3 * 148 * 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  for (int inner : {128})
    for (int rows : {32, 96, 192}) for (int nissuers : {1, 2, 4}) {
      for (int mode = 0; mode < 2; ++mode) {
        CUtensorMap tm;
        uint64_t dims[2] = {static_cast<uint64_t>(inner / 2), static_cast<uint64_t>(total_rows)};
        uint64_t strides[1] = {128};
        uint32_t box[2] = {static_cast<uint32_t>(inner / 2), static_cast<uint32_t>(rows)};
        if (make_tmap_bf16(&tm, d, 2, dims, strides, box,
                           inner == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (inner == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B)) != 0) {
          printf("tensor map failed: %s\n", "see fav_last_error");
          return 1;
        }
        const int tile_bytes = rows * inner;
        const int ring = std::max(2, std::min(kRingMax, (190 * 1024) / (tile_bytes * nissuers)));
        rate_kernel<<<148, 128, nissuers * ring * tile_bytes + 2048>>>(tm, d, rows, inner, total_rows * 128 / tile_bytes - 1, iters, mode, ring, nissuers, dout);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<long long> h(3 * 148);
        cudaMemcpy(h.data(), dout, 3 * 148 * 8, cudaMemcpyDeviceToHost);
        double mean = 0, mi = 0, mw = 0;
        for (int i = 0; i < 148; ++i) { mean += static_cast<double>(h[i]); mi += h[148 + i]; mw += h[296 + i]; }
        mean /= 148.0; mi /= 148.0; mw /= 148.0;
        printf("%-14s issuers %d (issue %.0f / wait %.0f clk per op) ring %2d rows %3d x %3d B = %5d B/tile: %7.1f clk/tile, %5.1f B/clk/SM, %5.2f clk/row  [%s]\n",
               mode == 0 ? "tensor 2-D box" : "bulk 1-D", nissuers, mi / iters, mw / iters, ring, rows, inner, tile_bytes, mean / iters, tile_bytes * iters * nissuers / mean,
               mean / iters / rows, cudaGetErrorString(e));
      }
    }
  return 0;
}

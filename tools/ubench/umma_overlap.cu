// Does a NO-SWIZZLE K-major UMMA shared-memory descriptor with LBO = 16 bytes address OVERLAPPING rows?
// Canonical interleaved layout: element (r, k) of the operand sits at
//     start + 16*(r % 8) + SBO*(r / 8) + 2*(k % 8) + LBO*(k / 8)        (bytes, 16-bit elements)
// With LBO = 16 and SBO = P, row r of an 8-row group reads the 32 bytes starting at 16*r of a raw buffer of pitch P:
// consecutive rows are 16-byte-shifted windows of the same bytes — exactly the W-direction im2col of a stride-2, 7-tap
// convolution over RGBX pixels (8 bytes each): output column wo reads pixels 2*wo .. 2*wo+7.  If the hardware accepts
// it, the stem conv can feed its A operand from RAW input rows (no 4x expanded im2col tile through L2 / TMA).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../flickering_adversarial_video_b200/csrc umma_overlap.cu -o umma_overlap
#include "fav_common.cuh"
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
namespace fav { __device__ int g_fav_timeout_flag = 0; }
using namespace fav;

constexpr int kN = 32;          // B rows
constexpr int kABytes = 16 * 256 + 1024;   // 16 row groups of pitch <= 256 + slack
constexpr int kBBytes = kN * 64;           // kN rows x 32 K elements

__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3ffff) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= 1ull << 46;
  return d;   // layout_type 0 = no swizzle
}

// ksteps MMAs of K = 16 each; A start advances by a_kstep bytes per step, B by 2 core matrices
__global__ void __launch_bounds__(128, 1) overlap_kernel(const uint16_t* a_raw, const uint16_t* b_can, float* out, int P,
                                                         int lbo_a, int ksteps, int a_kstep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  uint16_t* sa = reinterpret_cast<uint16_t*>(smem);
  uint16_t* sb = reinterpret_cast<uint16_t*>(smem + 8192);
  for (int i = threadIdx.x; i < kABytes / 2; i += blockDim.x) sa[i] = a_raw[i];
  for (int i = threadIdx.x; i < kBBytes / 2; i += blockDim.x) sb[i] = b_can[i];
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 32);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc(128, kN, true);
    for (int k = 0; k < ksteps; ++k) {
      // B canonical: (n, k) at 16*(n%8) + 256*ksteps... dense: K core matrices adjacent (LBO 128), row groups SBO = 128*2*ksteps
      const uint64_t ad = desc_nosw(smem_u32(smem) + k * a_kstep, lbo_a, P);
      const uint64_t bd = desc_nosw(smem_u32(smem + 8192) + k * 256, 128, 128 * 2 * ksteps);
      umma_bf16(tm, ad, bd, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  tc_fence_after();
  uint32_t r[32];
  tmem_ld_32x32(tm + ((threadIdx.x & ~31u) << 16), r);
  tmem_ld_wait();
  for (int j = 0; j < kN; ++j) out[threadIdx.x * kN + j] = __uint_as_float(r[j]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 32);
}

static float h2f(uint16_t h) { __half x; memcpy(&x, &h, 2); return __half2float(x); }
static uint16_t f2h(float f) { __half x = __float2half(f); uint16_t h; memcpy(&h, &x, 2); return h; }

int main() {
  std::vector<uint16_t> a(kABytes / 2), b(kBBytes / 2);
  srand(1);
  for (auto& v : a) v = f2h(static_cast<float>(rand() % 15 - 7));
  uint16_t *da, *db;
  float* dout;
  cudaMalloc(&da, kABytes); cudaMalloc(&db, kBBytes); cudaMalloc(&dout, 128 * kN * 4);
  cudaMemcpy(da, a.data(), kABytes, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(overlap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  int bad_total = 0;
  for (int P : {128, 192, 256})
    for (int ksteps : {1, 2}) {
      const int K = 16 * ksteps;
      // logical B [kN][K] -> canonical: (n, k) at element offset 8*(n%8) + (128*2*ksteps/2)*(n/8) + (k%8) + 64*(k/8)
      std::vector<float> bl(kN * K);
      for (auto& v : bl) v = static_cast<float>(rand() % 9 - 4);
      std::fill(b.begin(), b.end(), 0);
      for (int n = 0; n < kN; ++n)
        for (int k = 0; k < K; ++k) b[8 * (n % 8) + 128 * ksteps * (n / 8) + (k % 8) + 64 * (k / 8)] = f2h(bl[n * K + k]);
      cudaMemcpy(db, b.data(), kBBytes, cudaMemcpyHostToDevice);
      overlap_kernel<<<1, 128, 24 * 1024>>>(da, db, dout, P, 16, ksteps, 32);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> out(128 * kN);
      cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < kN; ++n) {
          float ref = 0;
          for (int k = 0; k < K; ++k) {
            // overlapped window: row r reads raw bytes at 16*(r%8) + P*(r/8) + 2*k  (k = 0..K-1 contiguous)
            const int byte = 16 * (r % 8) + P * (r / 8) + 2 * k;
            ref += h2f(a[byte / 2]) * bl[n * K + k];
          }
          if (ref != out[r * kN + n]) ++bad;
        }
      printf("P=%d K=%d: %s, %d of %d outputs differ from the overlapped-window model\n", P, K, cudaGetErrorString(e), bad,
             128 * kN);
      bad_total += bad;
    }
  printf(bad_total ? "OVERLAP_DESCRIPTOR_MISMATCH\n" : "OVERLAP_DESCRIPTOR_OK\n");
  return 0;
}

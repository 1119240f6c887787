// Torch-free check of fav_op_resize_crop on a GPU box: random uint8 frames, the same float32 arithmetic on the host
// (compile with -ffp-contract=off), bitwise comparison of the uint8 and the normalised fp32 outputs.
//   g++ -O2 -ffp-contract=off -I/usr/local/cuda/include -o loader_selftest loader_selftest.cpp \
//       -L../../flickering_adversarial_video_b200 -lfav -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN/../../flickering_adversarial_video_b200'
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../include/fav.h"

struct Tap { int i0, i1; float l0, l1; };
static Tap tap(int dst, int n_in, int n_out, float ratio) {
  Tap a;
  if (n_in == n_out) { a.i0 = a.i1 = dst; a.l0 = 1.f; a.l1 = 0.f; return a; }
  float s = ratio * ((float)dst + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  int i0 = (int)floorf(s); if (i0 > n_in - 1) i0 = n_in - 1;
  float l1 = s - (float)i0; l1 = fminf(fmaxf(l1, 0.f), 1.f);
  a.i0 = i0; a.i1 = i0 + (i0 < n_in - 1); a.l1 = l1; a.l0 = 1.f - l1;
  return a;
}

static int run_case(int N, int H, int W, int size, int crop, int fpc) {
  double scale = (double)size / (H < W ? H : W);
  int rh = (int)floor(H * scale), rw = (int)floor(W * scale);
  float ratio = (float)(1.0 / scale);
  int ci = (int)nearbyint((rh - crop) / 2.0), cj = (int)nearbyint((rw - crop) / 2.0);
  fav_norm_params nrm = {{0.43216f, 0.394666f, 0.37645f}, {0.22803f, 0.22145f, 0.216989f}, 0.f, 0.f};
  std::vector<uint8_t> src((size_t)N * H * W * 3), ref8((size_t)N * crop * crop * 3), got8(ref8.size());
  std::vector<float> reff(ref8.size()), gotf(ref8.size());
  uint32_t r = 12345u + H * 31 + W;
  for (auto& b : src) { r = r * 1664525u + 1013904223u; b = (r >> 24); if ((r & 0xf00) == 0) b = (r & 1) ? 255 : 0; }
  for (int f = 0; f < N; ++f)
    for (int y = 0; y < crop; ++y)
      for (int x = 0; x < crop; ++x) {
        Tap ty = tap(y + ci, H, rh, ratio), tx = tap(x + cj, W, rw, ratio);
        for (int c = 0; c < 3; ++c) {
          auto px = [&](int yy, int xx) { return (float)src[(((size_t)f * H + yy) * W + xx) * 3 + c] / 255.f; };
          float top = px(ty.i0, tx.i0) * tx.l0 + px(ty.i0, tx.i1) * tx.l1;
          float bot = px(ty.i1, tx.i0) * tx.l0 + px(ty.i1, tx.i1) * tx.l1;
          float v = top * ty.l0 + bot * ty.l1;
          long q = lrintf(v * 255.f); q = q < 0 ? 0 : (q > 255 ? 255 : q);
          size_t p = (size_t)y * crop + x;
          ref8[((size_t)f * crop * crop + p) * 3 + c] = (uint8_t)q;
          int b = f / fpc, t = f % fpc;
          reff[(((size_t)b * 3 + c) * fpc + t) * crop * crop + p] = (v - nrm.mean[c]) / nrm.std[c];
        }
      }
  uint8_t *d_src, *d_u8; float* d_f;
  if (cudaMalloc(&d_src, src.size()) || cudaMalloc(&d_u8, got8.size()) || cudaMalloc(&d_f, gotf.size() * 4)) return -1;
  cudaMemcpy(d_src, src.data(), src.size(), cudaMemcpyHostToDevice);
  cudaMemset(d_u8, 0x5a, got8.size());
  int st = fav_op_resize_crop(0, d_src, N, H, W, rh, rw, ratio, ratio, ci, cj, crop, crop, fpc, &nrm, d_u8, d_f, 0);
  if (st != 0) { printf("fav_op_resize_crop -> %d: %s\n", st, fav_last_error()); return -1; }
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("sync failed: %s\n", cudaGetErrorString(cudaGetLastError())); return -1; }
  cudaMemcpy(got8.data(), d_u8, got8.size(), cudaMemcpyDeviceToHost);
  cudaMemcpy(gotf.data(), d_f, gotf.size() * 4, cudaMemcpyDeviceToHost);
  size_t bad8 = 0, badf = 0; int max8 = 0; float maxf = 0.f;
  for (size_t i = 0; i < got8.size(); ++i) {
    int d = abs((int)got8[i] - (int)ref8[i]); bad8 += d != 0; if (d > max8) max8 = d;
    float e = fabsf(gotf[i] - reff[i]); badf += gotf[i] != reff[i]; if (e > maxf) maxf = e;
  }
  // timing: 20 launches
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaMemset(d_u8, 0x5a, got8.size());
  cudaEventRecord(e0, 0);
  for (int i = 0; i < 20; ++i)
    fav_op_resize_crop(0, d_src, N, H, W, rh, rw, ratio, ratio, ci, cj, crop, crop, fpc, nullptr, d_u8, nullptr, 0);
  cudaEventRecord(e1, 0); cudaEventSynchronize(e1);
  float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
  // the uint8-only launches of the timing loop must have produced the same bytes
  std::vector<uint8_t> again(got8.size());
  cudaMemcpy(again.data(), d_u8, again.size(), cudaMemcpyDeviceToHost);
  size_t bad_again = 0;
  for (size_t i = 0; i < again.size(); ++i) bad_again += again[i] != ref8[i];
  printf("  uint8-only launch: %zu mismatches\n", bad_again);
  bad8 += bad_again;
  double bytes = (double)N * (3.0 * H * W + 3.0 * crop * crop);
  printf("resize_crop N=%d %dx%d -> %dx%d crop %d@(%d,%d): uint8 mismatches %zu/%zu (max %d), fp32 mismatches %zu (max abs %.3e), "
         "%.1f us/launch, %.0f GB/s algorithmic\n", N, H, W, rh, rw, crop, ci, cj, bad8, got8.size(), max8, badf, maxf,
         ms / 20 * 1e3, bytes / (ms / 20 * 1e-3) / 1e9);
  cudaFree(d_src); cudaFree(d_u8); cudaFree(d_f);
  return (max8 <= 1 && bad8 * 1000 <= got8.size() && maxf <= 1e-6f) ? 0 : 1;
}

int main() {
  printf("%s\n", fav_build_info());
  int rc = 0;
  rc |= run_case(16, 256, 340, 128, 112, 16);
  rc |= run_case(32, 240, 320, 128, 112, 16);
  rc |= run_case(8, 360, 270, 128, 112, 4);
  rc |= run_case(4, 128, 171, 128, 112, 2);
  rc |= run_case(6, 90, 64, 32, 28, 3);
  rc |= run_case(256, 256, 340, 128, 112, 16);
  printf(rc == 0 ? "SELFTEST PASS\n" : "SELFTEST FAIL\n");
  return rc;
}

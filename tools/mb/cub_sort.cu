// Reference point for the sort stage: cub::DeviceRadixSort::SortPairs (CUDA 12.9's CUB, onesweep) against
// gs_sort_pairs_device on the same 6,131,954 (uint32 key, uint32 payload) pairs, same box, CUDA events.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/mb/cub_sort.cu -Iinclude \
//        -Lunitygaussiansplatting_b200 -lgsplat_b200 -Xlinker -rpath=$PWD/unitygaussiansplatting_b200 -o /tmp/cub_sort
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include "gsplat_b200.h"

static uint32_t rng(uint32_t &s) { s = s * 747796405u + 2891336453u; uint32_t w = ((s >> ((s >> 28) + 4)) ^ s) * 277803737u; return (w >> 22) ^ w; }
static uint32_t sortable(float f) { uint32_t u; memcpy(&u, &f, 4); return u ^ ((uint32_t)(-(int32_t)(u >> 31)) | 0x80000000u); }

int main(int argc, char **argv) {
  const uint32_t n = argc > 1 ? (uint32_t)atoi(argv[1]) : 6131954u;
  const int reps = 30;
  std::vector<uint32_t> hk(n), hv(n);
  uint32_t *dk, *dv, *dk2, *dv2, *k0;
  cudaMalloc(&dk, n * 4); cudaMalloc(&dv, n * 4); cudaMalloc(&dk2, n * 4); cudaMalloc(&dv2, n * 4); cudaMalloc(&k0, n * 4);
  GsContext *ctx = nullptr;
  if (gs_create(0, nullptr, &ctx) != 0) { printf("gs_create failed\n"); return 1; }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int kind = 0; kind < 2; ++kind) {
    uint32_t s = 12345u + kind;
    for (uint32_t i = 0; i < n; ++i) {
      if (kind == 0) hk[i] = rng(s);
      else {  // depth-like keys: sortable-uint of view-space z, clustered scene seen from z = -6
        const float u1 = (rng(s) >> 8) * (1.0f / 16777216.0f), u2 = (rng(s) >> 8) * (1.0f / 16777216.0f);
        hk[i] = sortable(6.0f + 8.0f * (2.0f * u1 - 1.0f) + 0.5f * (u2 - 0.5f));
      }
      hv[i] = i;
    }
    cudaMemcpy(k0, hk.data(), n * 4, cudaMemcpyHostToDevice);
    // --- CUB ---
    size_t tmp_bytes = 0; void *tmp = nullptr;
    { cub::DoubleBuffer<uint32_t> K(dk, dk2), V(dv, dv2); cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, K, V, (int)n); }
    cudaMalloc(&tmp, tmp_bytes);
    std::vector<float> t_cub, t_gs;
    std::vector<uint32_t> cub_v(n), gs_v(n);
    for (int r = 0; r < reps + 3; ++r) {
      cudaMemcpy(dk, k0, n * 4, cudaMemcpyDeviceToDevice); cudaMemcpy(dv, hv.data(), n * 4, cudaMemcpyHostToDevice);
      cub::DoubleBuffer<uint32_t> K(dk, dk2), V(dv, dv2);
      cudaDeviceSynchronize();
      cudaEventRecord(a); cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, K, V, (int)n); cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); if (r >= 3) t_cub.push_back(ms);
      if (r == reps + 2) cudaMemcpy(cub_v.data(), V.Current(), n * 4, cudaMemcpyDeviceToHost);
    }
    // --- ours, through the C ABI (includes its histogram kernel and the count upload) ---
    cudaStream_t gs_stream = (cudaStream_t)gs_context_stream(ctx);
    for (int r = 0; r < reps + 3; ++r) {
      cudaMemcpy(dk, k0, n * 4, cudaMemcpyDeviceToDevice); cudaMemcpy(dv, hv.data(), n * 4, cudaMemcpyHostToDevice);
      cudaDeviceSynchronize();
      cudaEventRecord(a, gs_stream); if (gs_sort_pairs_device(ctx, dk, dv, n) != 0) { printf("gs_sort failed: %s\n", gs_last_error(ctx)); return 1; }
      cudaEventRecord(b, gs_stream); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); if (r >= 3) t_gs.push_back(ms);
      if (r == reps + 2) cudaMemcpy(gs_v.data(), dv, n * 4, cudaMemcpyDeviceToHost);
    }
    std::sort(t_cub.begin(), t_cub.end()); std::sort(t_gs.begin(), t_gs.end());
    const double bytes = 68.0 * n;
    printf("%s keys, n=%u: CUB %.1f us (%.2f TB/s at 68 B/pair) | gs_sort_pairs_device %.1f us (%.2f TB/s) | payloads %s\n",
           kind ? "depth-like" : "uniform", n, t_cub[reps / 2] * 1e3, bytes / t_cub[reps / 2] / 1e9, t_gs[reps / 2] * 1e3,
           bytes / t_gs[reps / 2] / 1e9, cub_v == gs_v ? "identical" : "DIFFER");
    cudaFree(tmp);
  }
  gs_destroy(ctx);
  return 0;
}

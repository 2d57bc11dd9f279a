// micro-benchmark: warp match of 8-bit digits, MATCH.ANY vs 8 ballots (B200)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_match(const uint32_t* in, uint32_t* out, int iters) {
  uint32_t d = in[threadIdx.x + blockIdx.x * blockDim.x] & 255u, acc = 0;
  for (int i = 0; i < iters; ++i) { uint32_t m = __match_any_sync(0xffffffffu, d); acc += __popc(m); d = (d * 1664525u + acc + 1013904223u) & 255u; }
  out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}
__global__ void k_ballot(const uint32_t* in, uint32_t* out, int iters) {
  uint32_t d = in[threadIdx.x + blockIdx.x * blockDim.x] & 255u, acc = 0;
  for (int i = 0; i < iters; ++i) {
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; ++b) { bool bit = (d >> b) & 1u; uint32_t bal = __ballot_sync(0xffffffffu, bit); m &= bit ? bal : ~bal; }
    acc += __popc(m); d = (d * 1664525u + acc + 1013904223u) & 255u; }
  out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}
int main() {
  const int blocks = 148 * 8, threads = 256, iters = 2000;
  uint32_t *in, *out; cudaMalloc(&in, blocks * threads * 4); cudaMalloc(&out, blocks * threads * 4);
  uint32_t* h = new uint32_t[blocks * threads]; for (int i = 0; i < blocks * threads; ++i) h[i] = rand();
  cudaMemcpy(in, h, blocks * threads * 4, cudaMemcpyHostToDevice);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(a); k_match<<<blocks, threads>>>(in, out, iters); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("match.any : %.3f ms  -> %.2f ns per warp-op per SM-resident-warp-set; %.1f Gops(lane)/s\n", ms, ms * 1e6 / iters, (double)blocks * threads * iters / ms / 1e6);
    cudaEventRecord(a); k_ballot<<<blocks, threads>>>(in, out, iters); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("8 ballots : %.3f ms  -> %.2f ns; %.1f Gops(lane)/s\n", ms, ms * 1e6 / iters, (double)blocks * threads * iters / ms / 1e6);
  }
  return 0;
}

// gs_sort.cu -- stable LSD radix sort of (uint32 key, uint32 payload) pairs for sm_100a.
//
// Replaces GpuSorting.Dispatch (R/GpuSorting.cs:142-198) and the DeviceRadixSort kernels
// (S/DeviceRadixSort.hlsl: Init/Upsweep/Scan/Downsweep; 13 dispatches, 80 B/pair of traffic).
// Same contract -- ascending, stable, 8-bit digits, 4 passes, key+payload, result back in the
// input buffers -- different algorithm: a single-pass-per-digit "onesweep" with decoupled
// look-back, so a sort is 4 data passes + one histogram read (68 B/pair) instead of
// reduce-then-scan's two reads per pass:
//   * digit histograms for all passes come from one read of the keys (or for free from
//     k_calc_distances, which wrote the keys in the first place);
//   * each CTA takes a 4096-pair tile by atomic ticket (so a tile's predecessors are
//     always resident: look-back cannot deadlock), ranks its keys with warp ballots
//     (8 ballots per key build the match mask of equal digits, popc of the lower lanes is
//     the stable rank, one lane bumps the warp-private shared histogram), publishes the
//     tile's per-digit count with a LOCAL flag, walks back over predecessors' status words
//     until it meets an INCLUSIVE one, then publishes its own inclusive prefix;
//   * keys and payloads are first scattered inside shared memory into digit order, then
//     written out in runs, so global stores are coalesced per digit run.
// No tensor-core path: there is no contraction here, only byte/integer traffic.
#include "gs_kernels.cuh"

namespace gs {

constexpr int kSortThreads = 256;
constexpr int kSortKPT = 16;
static_assert(kSortThreads * kSortKPT == (int)kSortTileItems, "tile size");
constexpr int kSortWarps = kSortThreads / 32;
constexpr uint32_t kFlagLocal = 1u << 30, kFlagIncl = 2u << 30, kValMask = (1u << 30) - 1u;

size_t sort_lookback_words(uint32_t capacity, int passes) {
  size_t tiles = ((size_t)capacity + kSortTileItems - 1) / kSortTileItems;
  return tiles * 256 * (size_t)passes;
}

// ---- digit histograms of all passes in one read -------------------------------------------
template <int PASSES>
__global__ void __launch_bounds__(256) k_sort_hist(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ d_count,
                                                   uint32_t *__restrict__ ghist) {
  __shared__ uint32_t sh[PASSES * 256];
  for (int i = threadIdx.x; i < PASSES * 256; i += 256) sh[i] = 0;
  __syncthreads();
  const uint32_t n = *d_count;
  const uint32_t nvec = n >> 2;
  const uint4 *k4 = reinterpret_cast<const uint4 *>(keys);
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < nvec; i += gridDim.x * 256) {
    uint4 v = __ldg(k4 + i);
    const uint32_t kk[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int p = 0; p < PASSES; ++p) atomicAdd(&sh[p * 256 + ((kk[q] >> (8 * p)) & 255u)], 1u);
  }
  if (blockIdx.x == 0) {
    for (uint32_t i = (nvec << 2) + threadIdx.x; i < n; i += 256) {
      uint32_t k = keys[i];
#pragma unroll
      for (int p = 0; p < PASSES; ++p) atomicAdd(&sh[p * 256 + ((k >> (8 * p)) & 255u)], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PASSES * 256; i += 256) {
    uint32_t c = sh[i];
    if (c) atomicAdd(&ghist[i], c);
  }
}

// ---- block-wide exclusive scan of one value per thread (256 threads) ------------------------
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t *s_warp /*8*/) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t wsum = (lane < kSortWarps) ? s_warp[lane] : 0u;
  uint32_t winc = wsum;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
    if (lane >= (uint32_t)o) winc += t;
  }
  uint32_t wexcl = __shfl_sync(0xffffffffu, winc - wsum, warp);
  __syncthreads();  // s_warp reusable afterwards
  return wexcl + inc - v;
}

// ---- one digit pass -------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads)
k_onesweep(const uint32_t *__restrict__ src_k, const uint32_t *__restrict__ src_v, uint32_t *__restrict__ dst_k,
           uint32_t *__restrict__ dst_v, const uint32_t *__restrict__ d_count, int shift, const uint32_t *__restrict__ ghist,
           volatile uint32_t *lookback, uint32_t *ticket) {
  __shared__ uint32_t s_keys[kSortTileItems];
  __shared__ uint32_t s_vals[kSortTileItems];
  __shared__ uint32_t s_whist[kSortWarps][256];
  __shared__ uint32_t s_dig_start[256];
  __shared__ uint32_t s_off[256];
  __shared__ uint32_t s_scan[kSortWarps];
  __shared__ uint32_t s_tile;

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t n = *d_count;
  const uint32_t num_tiles = (n + kSortTileItems - 1) / kSortTileItems;
  if (tile >= num_tiles) return;
  const uint32_t tile_base = tile * kSortTileItems;

  // warp-striped load: warp w owns 512 consecutive pairs, item i of lane l is base + i*32 + l
  uint32_t key[kSortKPT], val[kSortKPT];
  const uint32_t wbase = tile_base + warp * (32 * kSortKPT) + lane;
#pragma unroll
  for (int i = 0; i < kSortKPT; ++i) {
    uint32_t idx = wbase + i * 32;
    key[i] = (idx < n) ? __ldg(src_k + idx) : 0xFFFFFFFFu;  // pads sort last (S/SortCommon.hlsl:244-247 does the same)
  }
#pragma unroll
  for (int i = 0; i < kSortKPT; ++i) {
    uint32_t idx = wbase + i * 32;
    val[i] = (idx < n) ? __ldg(src_v + idx) : 0u;
  }
#pragma unroll
  for (int d = lane; d < 256; d += 32) s_whist[warp][d] = 0;
  __syncwarp();

  // stable in-warp ranking by ballot match
  uint32_t rank[kSortKPT];
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < kSortKPT; ++i) {
    const uint32_t d = (key[i] >> shift) & 255u;
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const bool bit = (d >> b) & 1u;
      const uint32_t bal = __ballot_sync(0xffffffffu, bit);
      m &= bit ? bal : ~bal;
    }
    const uint32_t leader = __ffs(m) - 1;
    uint32_t prev = 0;
    if (lane == leader) {
      prev = s_whist[warp][d];
      s_whist[warp][d] = prev + __popc(m);
    }
    prev = __shfl_sync(0xffffffffu, prev, leader);
    rank[i] = prev + __popc(m & lt_mask);
    __syncwarp();
  }
  __syncthreads();

  // thread `tid` now owns digit `tid`: exclusive prefix over warps, tile total
  uint32_t run = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) {
    uint32_t c = s_whist[w][tid];
    s_whist[w][tid] = run;
    run += c;
  }
  const uint32_t tile_total = run;  // includes pads (only digit 255 of the last tile)
  uint32_t pub = tile_total;
  if (tid == 255 && tile_base + kSortTileItems > n) pub -= (tile_base + kSortTileItems - n);

  // publish early so successors can start their look-back while we scan
  volatile uint32_t *lb = lookback + (size_t)tile * 256 + tid;
  if (tile == 0) *lb = kFlagIncl | pub; else *lb = kFlagLocal | pub;

  const uint32_t dig_start = block_excl_scan_256(tile_total, s_scan);
  const uint32_t gbase = block_excl_scan_256(__ldg(ghist + tid), s_scan);

  uint32_t prefix = 0;
  if (tile > 0) {
    int t = (int)tile - 1;
    while (true) {
      uint32_t v = lookback[(size_t)t * 256 + tid];
      if (v == 0) continue;  // predecessor not published yet
      prefix += v & kValMask;
      if (v & kFlagIncl) break;
      --t;
    }
    *lb = kFlagIncl | (prefix + pub);
  }
  s_dig_start[tid] = dig_start;
  s_off[tid] = gbase + prefix - dig_start;
  __syncthreads();

  // scatter into digit order inside shared memory
#pragma unroll
  for (int i = 0; i < kSortKPT; ++i) {
    const uint32_t d = (key[i] >> shift) & 255u;
    const uint32_t pos = s_dig_start[d] + s_whist[warp][d] + rank[i];
    s_keys[pos] = key[i];
    s_vals[pos] = val[i];
  }
  __syncthreads();

  const uint32_t valid = min(kSortTileItems, n - tile_base);
#pragma unroll 4
  for (uint32_t j = tid; j < valid; j += kSortThreads) {
    const uint32_t k = s_keys[j];
    const uint32_t dst = j + s_off[(k >> shift) & 255u];
    dst_k[dst] = k;
    dst_v[dst] = s_vals[j];
  }
}

void launch_sort_pairs(uint32_t *keys, uint32_t *vals, const uint32_t *d_count, uint32_t capacity, int passes, bool hist_ready,
                       const SortScratch &sc, cudaStream_t s, cudaEvent_t *pass_events) {
  if (capacity == 0) return;
  const uint32_t tiles = (capacity + kSortTileItems - 1) / kSortTileItems;
  cudaMemsetAsync(sc.lookback, 0, (size_t)tiles * 256 * passes * sizeof(uint32_t), s);
  cudaMemsetAsync(sc.tickets, 0, 4 * sizeof(uint32_t), s);
  if (!hist_ready) {
    cudaMemsetAsync(sc.ghist, 0, 4 * 256 * sizeof(uint32_t), s);
    const uint32_t grid = min(tiles, 148u * 8u);
    if (passes == 4) k_sort_hist<4><<<grid, 256, 0, s>>>(keys, d_count, sc.ghist);
    else k_sort_hist<2><<<grid, 256, 0, s>>>(keys, d_count, sc.ghist);
  }
  uint32_t *sk = keys, *sv = vals, *dk = sc.alt_keys, *dv = sc.alt_vals;
  for (int p = 0; p < passes; ++p) {
    if (pass_events) cudaEventRecord(pass_events[p], s);
    k_onesweep<<<tiles, kSortThreads, 0, s>>>(sk, sv, dk, dv, d_count, 8 * p, sc.ghist + 256 * p,
                                               sc.lookback + (size_t)p * tiles * 256, sc.tickets + p);
    uint32_t *t = sk; sk = dk; dk = t;
    t = sv; sv = dv; dv = t;
  }
  if (pass_events) cudaEventRecord(pass_events[passes], s);
}

}  // namespace gs

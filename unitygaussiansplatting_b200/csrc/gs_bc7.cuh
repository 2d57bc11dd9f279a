// gs_bc7.cuh -- single-texel BC7 (BPTC) decode for ColorFormat.BC7 assets (R/GaussianSplatAsset.cs:56,169).
//
// The reference samples a RGBA_BC7_UNorm texture: the decode is the texture unit's, i.e. the published block format.
// CSCalcViewData needs ONE texel per splat (S/GaussianSplatting.hlsl:423-426), so instead of expanding whole blocks this
// reads only the fields that texel depends on: its subset's two endpoints, their p-bits and its own index bits.
// Plain C++ (no intrinsics) so that tests/bc7_texel_test.cpp can compile the very same code with g++ and check it
// against the Pillow-decoded golden blocks without a GPU.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define GS_BC7_HD __device__ __forceinline__
#define GS_BC7_TABLE static __device__ const
#else
#define GS_BC7_HD static inline
#define GS_BC7_TABLE static const
#endif

namespace gs {
namespace bc7 {

#include "bc7_tables.h"

// per mode: subsets, partition bits, rotation bits, index-selection bit, colour bits, alpha bits, per-endpoint p-bit,
// shared (per-subset) p-bit, index bits, secondary index bits -- one nibble each, mode m in kModes[m]
GS_BC7_TABLE uint64_t kModes[8] = {
    // ib2 ib spb epb ab cb isb rb pb ns   (nibbles, most significant first)
    0x0301040043ull, 0x0310060062ull, 0x0200050063ull, 0x0201070062ull,
    0x3200651201ull, 0x2200870201ull, 0x0401770001ull, 0x0201550062ull};
GS_BC7_TABLE uint8_t kW2[4] = {0, 21, 43, 64};
GS_BC7_TABLE uint8_t kW3[8] = {0, 9, 18, 27, 37, 46, 55, 64};
GS_BC7_TABLE uint8_t kW4[16] = {0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64};

// n (0..8) bits starting at bit `pos` of the 128-bit block (w[0] = least significant word)
GS_BC7_HD uint32_t bits(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t pos, uint32_t n) {
  const uint32_t wi = pos >> 5, sh = pos & 31u;
  const uint32_t lo = wi == 0 ? w0 : wi == 1 ? w1 : wi == 2 ? w2 : w3;
  const uint32_t hi = wi == 0 ? w1 : wi == 1 ? w2 : wi == 2 ? w3 : 0u;
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)(v >> sh) & ((1u << n) - 1u);
}

GS_BC7_HD uint32_t weight(uint32_t nbits, uint32_t i) { return nbits == 2 ? kW2[i] : nbits == 3 ? kW3[i] : kW4[i]; }

// RGBA8 of texel `pix` (0..15, raster order) of one block, packed r | g<<8 | b<<16 | a<<24
GS_BC7_HD uint32_t decode_texel(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t pix) {
  const uint32_t m8 = w0 & 0xffu;
  if (m8 == 0) return 0u;  // reserved mode: all channels 0
  uint32_t mode = 0;
  while (!((m8 >> mode) & 1u)) ++mode;
  const uint64_t mi = kModes[mode];
  const uint32_t ns = (uint32_t)mi & 15u, pb = (uint32_t)(mi >> 4) & 15u, rb = (uint32_t)(mi >> 8) & 15u, isbn = (uint32_t)(mi >> 12) & 15u,
                 cb = (uint32_t)(mi >> 16) & 15u, ab = (uint32_t)(mi >> 20) & 15u, epb = (uint32_t)(mi >> 24) & 15u,
                 spb = (uint32_t)(mi >> 28) & 15u, ib = (uint32_t)(mi >> 32) & 15u, ib2 = (uint32_t)(mi >> 36) & 15u;
  uint32_t pos = mode + 1;
  const uint32_t part = bits(w0, w1, w2, w3, pos, pb); pos += pb;
  const uint32_t rot = bits(w0, w1, w2, w3, pos, rb); pos += rb;
  const uint32_t isb = bits(w0, w1, w2, w3, pos, isbn); pos += isbn;
  uint32_t subset = 0, a1 = 0, a2 = 0;
  if (ns == 2) { subset = (kBc7Part2[part] >> pix) & 1u; a1 = kBc7Anchor2[part]; }
  if (ns == 3) { subset = (kBc7Part3[part] >> (2 * pix)) & 3u; a1 = kBc7Anchor3a[part]; a2 = kBc7Anchor3b[part]; }
  const uint32_t ne = ns * 2, e0 = subset * 2;
  const uint32_t pos_a = pos + 3 * ne * cb, pos_p = pos_a + ne * ab, pos_i = pos_p + (epb ? ne : spb ? ns : 0u);
  uint32_t lo[4], hi[4];
  for (uint32_t c = 0; c < 3; ++c) {
    lo[c] = bits(w0, w1, w2, w3, pos + (c * ne + e0) * cb, cb);
    hi[c] = bits(w0, w1, w2, w3, pos + (c * ne + e0 + 1) * cb, cb);
  }
  lo[3] = bits(w0, w1, w2, w3, pos_a + e0 * ab, ab);
  hi[3] = bits(w0, w1, w2, w3, pos_a + (e0 + 1) * ab, ab);
  uint32_t cbt = cb, abt = ab;
  if (epb | spb) {
    const uint32_t p0 = bits(w0, w1, w2, w3, pos_p + (epb ? e0 : subset), 1), p1 = epb ? bits(w0, w1, w2, w3, pos_p + e0 + 1, 1) : p0;
    for (uint32_t c = 0; c < 3; ++c) { lo[c] = (lo[c] << 1) | p0; hi[c] = (hi[c] << 1) | p1; }
    cbt += 1;
    if (ab && epb) { lo[3] = (lo[3] << 1) | p0; hi[3] = (hi[3] << 1) | p1; abt += 1; }
  }
  for (uint32_t c = 0; c < 3; ++c) {
    uint32_t x = lo[c] << (8 - cbt); lo[c] = x | (x >> cbt);
    x = hi[c] << (8 - cbt); hi[c] = x | (x >> cbt);
  }
  if (ab) {
    uint32_t x = lo[3] << (8 - abt); lo[3] = x | (x >> abt);
    x = hi[3] << (8 - abt); hi[3] = x | (x >> abt);
  } else {
    lo[3] = hi[3] = 255u;
  }
  // primary index: anchors (pixel 0, a1, a2) are stored one bit short, so each anchor before `pix` shifts it by one
  const uint32_t anchor = subset == 0 ? 0u : subset == 1 ? a1 : a2;
  const uint32_t before = (pix > 0 ? 1u : 0u) + ((ns >= 2 && a1 < pix) ? 1u : 0u) + ((ns == 3 && a2 < pix) ? 1u : 0u);
  const uint32_t i1 = bits(w0, w1, w2, w3, pos_i + pix * ib - before, ib - (pix == anchor ? 1u : 0u));
  uint32_t ci = i1, cbits = ib, ai = i1, abits = ib;
  if (ib2) {
    const uint32_t pos_i2 = pos_i + 16 * ib - ns;
    const uint32_t i2 = bits(w0, w1, w2, w3, pos_i2 + pix * ib2 - (pix > 0 ? 1u : 0u), ib2 - (pix == 0 ? 1u : 0u));
    if (!isb) { ai = i2; abits = ib2; } else { ci = i2; cbits = ib2; }
  }
  const uint32_t wc = weight(cbits, ci), wa = weight(abits, ai);
  uint32_t px[4];
  for (uint32_t c = 0; c < 3; ++c) px[c] = ((64 - wc) * lo[c] + wc * hi[c] + 32) >> 6;
  px[3] = ((64 - wa) * lo[3] + wa * hi[3] + 32) >> 6;
  if (rot == 1) { const uint32_t t = px[3]; px[3] = px[0]; px[0] = t; }  // channel rotation (modes 4, 5): A <-> R / G / B
  if (rot == 2) { const uint32_t t = px[3]; px[3] = px[1]; px[1] = t; }
  if (rot == 3) { const uint32_t t = px[3]; px[3] = px[2]; px[2] = t; }
  return px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24);
}

}  // namespace bc7
}  // namespace gs

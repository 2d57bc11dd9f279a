// gs_export.cu -- CSExportData (S/SplatUtilities.compute:616-669): every splat of an uploaded asset decoded back to the
// INRIA .ply attribute record (the 62-float InputSplatData layout, E/Utils/GaussianFileReader.cs:17-26), which is what
// GaussianSplatRendererEditor.ExportPlyFile (E/GaussianSplatRendererEditor.cs:394-445) writes after dropping deleted and
// cut splats.  Runs once per export, not per frame: a plain one-thread-per-splat full decode (LoadSplatData,
// S/GaussianSplatting.hlsl:428-608) with none of the render path's deferred loads.
#include "gs_kernels.cuh"
#include "gs_bc7.cuh"

namespace gs {

namespace {

__device__ __forceinline__ float chunk_lerp(uint32_t packed, float t) { return lerpf(f16lo(packed), f16hi(packed), t); }

__device__ bool export_is_cut(uint32_t count, const GsCutout *__restrict__ cut, float3 p) {  // IsSplatCut, :164-187
  bool finalCut = false;
  for (uint32_t i = 0; i < count; ++i) {
    const GsCutout &c = cut[i];
    const uint32_t type = c.type_and_flags & 0xFFu;
    if (type == 0xFFu) continue;
    const bool invert = (c.type_and_flags & 0xFF00u) != 0;
    const float *m = c.mat;
    const float cx = fmaf(m[8], p.z, fmaf(m[4], p.y, fmaf(m[0], p.x, m[12])));
    const float cy = fmaf(m[9], p.z, fmaf(m[5], p.y, fmaf(m[1], p.x, m[13])));
    const float cz = fmaf(m[10], p.z, fmaf(m[6], p.y, fmaf(m[2], p.x, m[14])));
    if (type == 0) { if (cx * cx + cy * cy + cz * cz <= 1.0f) return invert; }
    if (type == 1) { if (fabsf(cx) <= 1.0f && fabsf(cy) <= 1.0f && fabsf(cz) <= 1.0f) return invert; }
    finalCut |= !invert;
  }
  return finalCut;
}

__global__ void __launch_bounds__(256) k_export_data(AssetView a, uint32_t cutoutCount, const GsCutout *__restrict__ cutouts,
                                                      float *__restrict__ out) {
  const uint32_t idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= a.n) return;
  const bool clustered = a.shFmt > 3;
  const uint32_t otherStride = 4 + vec_stride(a.scaleFmt) + (clustered ? 2u : 0u);
  const uint64_t otherAddr = (uint64_t)idx * otherStride;
  float3 pos = load_vector(a.pos, (uint64_t)idx * vec_stride(a.posFmt), a.posFmt);
  const uint32_t rq = (otherAddr & 3) == 0 ? ld_u32(a.other + otherAddr) : ld_u32_a2(a.other + otherAddr);
  float3 scale = load_vector(a.other, otherAddr + 4, a.scaleFmt);
  // DecodeRotation(DecodePacked_10_10_10_2), :219-229,:293-300
  float4 rot;
  {
    const float px = (float)(rq & 1023) * GS_INV(1023.0f), py = (float)((rq >> 10) & 1023) * GS_INV(1023.0f),
                pz = (float)((rq >> 20) & 1023) * GS_INV(1023.0f);
    const uint32_t qi = rq >> 30;
    const float kSqrt2 = 1.41421354f, kInvSqrt2 = 0.707106769f;
    const float x = fmaf(px, kSqrt2, -kInvSqrt2), y = fmaf(py, kSqrt2, -kInvSqrt2), z = fmaf(pz, kSqrt2, -kInvSqrt2);
    const float w = sqrtf(1.0f - satf(fmaf(z, z, fmaf(y, y, x * x))));
    rot = make_float4(x, y, z, w);
    if (qi == 0) rot = make_float4(w, x, y, z);
    if (qi == 1) rot = make_float4(x, w, y, z);
    if (qi == 2) rot = make_float4(x, y, w, z);
  }
  // colour texel
  float4 col;
  {
    const uint32_t ti = splat_index_to_texel(idx);
    if (a.colFmt == 0) {
      col = __ldg(reinterpret_cast<const float4 *>(a.color) + ti);
    } else if (a.colFmt == 1) {
      const uint2 e = __ldg(reinterpret_cast<const uint2 *>(a.color) + ti);
      col = make_float4(f16lo(e.x), f16hi(e.x), f16lo(e.y), f16hi(e.y));
    } else {
      uint32_t e;
      if (a.colFmt == 2) {
        e = __ldg(reinterpret_cast<const uint32_t *>(a.color) + ti);
      } else {
        const uint32_t x = ti & (kTexWidth - 1), y = ti / kTexWidth;
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(a.color) + ((uint64_t)(y >> 2) * (kTexWidth / 4) + (x >> 2)));
        e = bc7::decode_texel(b.x, b.y, b.z, b.w, (y & 3u) * 4u + (x & 3u));
      }
      col = make_float4(__fdiv_rn((float)(e & 255u), 255.0f), __fdiv_rn((float)((e >> 8) & 255u), 255.0f),
                        __fdiv_rn((float)((e >> 16) & 255u), 255.0f), __fdiv_rn((float)(e >> 24), 255.0f));
    }
  }
  // SH, :467-562
  float sh[45];
  {
    const uint32_t shIdx = clustered ? ld_u16(a.other + otherAddr + otherStride - 2) : idx;
    const uint32_t stride = a.shFmt == 0 ? 192u : (a.shFmt == 1 || clustered) ? 96u : a.shFmt == 2 ? 60u : 32u;
    const uint8_t *p = a.sh + (uint64_t)shIdx * stride;
    for (int j = 0; j < 15; ++j) {
      float3 v;
      if (a.shFmt == 0) {
        v = make_float3(__uint_as_float(ld_u32(p + (j * 3) * 4)), __uint_as_float(ld_u32(p + (j * 3 + 1) * 4)), __uint_as_float(ld_u32(p + (j * 3 + 2) * 4)));
      } else if (a.shFmt == 1 || clustered) {
        v = make_float3(f16lo(ld_u16(p + (j * 3) * 2)), f16lo(ld_u16(p + (j * 3 + 1) * 2)), f16lo(ld_u16(p + (j * 3 + 2) * 2)));
      } else if (a.shFmt == 2) {
        v = dec_11_10_11(ld_u32(p + j * 4));
      } else {
        v = dec_5_6_5(ld_u16(p + j * 2));
      }
      sh[j * 3] = v.x; sh[j * 3 + 1] = v.y; sh[j * 3 + 2] = v.z;
    }
  }
  const uint32_t ci = idx / kChunkSize;
  if (ci < a.chunkCount) {  // :565-603
    const Chunk c = a.chunks[ci];
    pos.x = lerpf(c.posX.x, c.posX.y, pos.x); pos.y = lerpf(c.posY.x, c.posY.y, pos.y); pos.z = lerpf(c.posZ.x, c.posZ.y, pos.z);
    scale.x = chunk_lerp(c.sclX, scale.x); scale.y = chunk_lerp(c.sclY, scale.y); scale.z = chunk_lerp(c.sclZ, scale.z);
    scale.x *= scale.x; scale.x *= scale.x; scale.x *= scale.x;
    scale.y *= scale.y; scale.y *= scale.y; scale.y *= scale.y;
    scale.z *= scale.z; scale.z *= scale.z; scale.z *= scale.z;
    col.x = chunk_lerp(c.colR, col.x); col.y = chunk_lerp(c.colG, col.y); col.z = chunk_lerp(c.colB, col.z); col.w = chunk_lerp(c.colA, col.w);
    float x = col.w - 0.5f;
    x *= 0.5f;
    const float sg = (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f);
    col.w = sqrtf(fabsf(x)) * sg + 0.5f;
    if (a.shFmt > 0 && a.shFmt <= 3) {
      for (int j = 0; j < 15; ++j) {
        sh[j * 3] = chunk_lerp(c.shR, sh[j * 3]); sh[j * 3 + 1] = chunk_lerp(c.shG, sh[j * 3 + 1]); sh[j * 3 + 2] = chunk_lerp(c.shB, sh[j * 3 + 2]);
      }
    }
  }
  const bool isCut = cutoutCount && export_is_cut(cutoutCount, cutouts, pos);
  // ExportSplatData, :645-667
  float *d = out + (size_t)idx * 62;
  d[0] = pos.x; d[1] = pos.y; d[2] = pos.z;
  const float nor = isCut ? 1.0f : 0.0f;  // "mark as skipped for export"
  d[3] = nor; d[4] = nor; d[5] = nor;
  d[6] = __fdiv_rn(col.x - 0.5f, 0.2820948f); d[7] = __fdiv_rn(col.y - 0.5f, 0.2820948f); d[8] = __fdiv_rn(col.z - 0.5f, 0.2820948f);  // ColorToSH0 :537-540
  for (int ch = 0; ch < 3; ++ch)
    for (int j = 0; j < 15; ++j) d[9 + ch * 15 + j] = sh[j * 3 + ch];   // shR14 R58 R9C RDF, then G, then B: channel-major like f_rest_*
  d[54] = logf(__fdiv_rn(col.w, fmaxf(1.0f - col.w, 1.0e-6f)));       // InvSigmoid :541-544
  d[55] = logf(scale.x); d[56] = logf(scale.y); d[57] = logf(scale.z);
  d[58] = rot.w; d[59] = rot.x; d[60] = rot.y; d[61] = rot.z;          // src.rot.wxyz
}

}  // namespace

void launch_export_data(const AssetView &a, uint32_t cutoutCount, const GsCutout *cutouts, float *out, cudaStream_t s) {
  if (!a.n) return;
  k_export_data<<<(a.n + 255) / 256, 256, 0, s>>>(a, cutoutCount, cutouts, out);
}

}  // namespace gs

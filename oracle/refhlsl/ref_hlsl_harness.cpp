// ref_hlsl_harness.cpp -- C entry points around the reference's own shader code compiled for the CPU (see build_ref_hlsl.py,
// hlsl_shim.hpp).  TEST INFRASTRUCTURE ONLY.  The host side below is what package/Runtime/GaussianSplatRenderer.cs does before a
// dispatch: bind buffers, pack _SplatFormat (:502), and multiply the matrices C# multiplies (:586-606, :617-631) with
// Unity's Matrix4x4 operator* (row times column, summed left to right).  UNITY_MATRIX_VP / _P are engine globals in
// the reference: here P is the GPU projection the caller passes and VP = P * V.
#include "hlsl_shim.hpp"

#include "../gs_oracle.h"   // GsoAsset / GsoFrame / GsoView: plain-C structs shared with the oracle's Python binding

namespace hlsl {

static float4x4 from_colmajor(const float *m) {   // UnityEngine.Matrix4x4 memory order -> row-major element access
  float4x4 r;
  for (int row = 0; row < 4; ++row) for (int col = 0; col < 4; ++col) r.m[row][col] = m[col * 4 + row];
  return r;
}
static float4x4 mat_mul(const float4x4 &a, const float4x4 &b) {   // Matrix4x4.operator*
  float4x4 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] + a.m[i][3] * b.m[3][j];
  return r;
}

namespace refcs {
static float4x4 unity_MatrixVP, glstate_matrix_projection;
#define UNITY_MATRIX_VP unity_MatrixVP
#define UNITY_MATRIX_P glstate_matrix_projection
#include "ref_cs.inc"
#undef UNITY_MATRIX_VP
#undef UNITY_MATRIX_P
}  // namespace refcs

#undef SplatBufferDataType
namespace refps {
static bool g_discarded;
static float4 _ScreenParams;
#include "ref_ps.inc"
}  // namespace refps

// what the texture unit does for the colour image (GraphicsFormat by ColorFormat, R/GaussianSplatAsset.cs:161-172)
static const uint8_t *g_color;
static uint g_color_format;
static float4 fetch_color(uint x, uint y) {
  const uint64_t ti = (uint64_t)y * 2048u + x;
  float4 r;
  if (g_color_format == 0) std::memcpy(&r.x, g_color + ti * 16, 16);
  else if (g_color_format == 1) { uint16_t h[4]; std::memcpy(h, g_color + ti * 8, 8); for (int k = 0; k < 4; ++k) r[k] = f16tof32(h[k]); }
  else { for (int k = 0; k < 4; ++k) r[k] = (float)g_color[ti * 4 + k] / 255.0f; }
  return r;
}

static int bind_asset(const GsoAsset *a) {
  using namespace refcs;
  if (a->color_format > 2) return -1;   // BC7: the block decode belongs to the texture unit, not to the reference's code
  _SplatPos.p = (const uint8_t *)a->pos;
  _SplatOther.p = (const uint8_t *)a->other;
  _SplatSH.p = (const uint8_t *)a->sh;
  g_color = (const uint8_t *)a->color;
  g_color_format = a->color_format;
  _SplatColor.fetch = fetch_color;
  _SplatChunks.p = (SplatChunkInfo *)a->chunks;
  _SplatChunkCount = a->chunks ? (uint)(a->chunk_bytes / 64) : 0u;
  _SplatFormat = a->pos_format | (a->scale_format << 8) | (a->sh_format << 16);   // R/GaussianSplatRenderer.cs:502
  _SplatCount = a->splat_count;
  return 0;
}

}  // namespace hlsl

using namespace hlsl;

extern "C" {

#define REF_API __attribute__((visibility("default")))

// SortPoints (R/GaussianSplatRenderer.cs:612-639) -> CSCalcDistances
REF_API int refhlsl_calc_distances(const GsoAsset *a, const GsoFrame *f, uint32_t *order, uint32_t *keys) {
  using namespace refcs;
  if (bind_asset(a)) return -1;
  float4x4 w2c = from_colmajor(f->mat_view);
  w2c.m[2][0] *= -1; w2c.m[2][1] *= -1; w2c.m[2][2] *= -1;                     // :621-623
  _MatrixMV = mat_mul(w2c, from_colmajor(f->mat_object_to_world));           // :629
  _SplatSortKeys.p = order;
  _SplatSortDistances.p = keys;
  for (uint i = 0; i < a->splat_count; ++i) CSCalcDistances(uint3(i, 0, 0));
  return 0;
}

// CalcViewData (:579-610) -> CSCalcViewData
REF_API int refhlsl_calc_view(const GsoAsset *a, const GsoFrame *f, GsoView *out) {
  using namespace refcs;
  if (bind_asset(a)) return -1;
  const float4x4 V = from_colmajor(f->mat_view), O2W = from_colmajor(f->mat_object_to_world);
  _MatrixObjectToWorld = O2W;
  _MatrixWorldToObject = from_colmajor(f->mat_world_to_object);
  _MatrixMV = mat_mul(V, O2W);                                               // :594
  glstate_matrix_projection = from_colmajor(f->mat_proj_gpu);
  unity_MatrixVP = mat_mul(glstate_matrix_projection, V);
  _VecScreenParams = float4(f->screen_w, f->screen_h, 0, 0);
  _VecWorldSpaceCameraPos = float4(f->cam_pos_world[0], f->cam_pos_world[1], f->cam_pos_world[2], 0);
  _SplatScale = f->splat_scale;
  _SplatOpacityScale = f->opacity_scale;
  _SHOrder = f->sh_order;
  _SHOnly = f->sh_only;
  static_assert(sizeof(GaussianCutoutShaderData) == 68 && sizeof(SplatViewData) == 40, "layouts the C# side relies on");
  // GsoCutout.mat is column-major like Matrix4x4; the shader reads a row-major-indexed float4x4: convert a copy
  static GaussianCutoutShaderData cuts[64];
  const uint nc = f->cutouts ? (f->cutout_count < 64 ? f->cutout_count : 64) : 0;
  for (uint i = 0; i < nc; ++i) { cuts[i].mat = from_colmajor(f->cutouts[i].mat); cuts[i].typeAndFlags = f->cutouts[i].type_and_flags; }
  _SplatCutouts.p = cuts;
  _SplatCutoutsCount = nc;
  _SplatDeletedBits.p = (const uint8_t *)f->deleted_bits;
  _SplatBitsValid = f->deleted_bits ? 1u : 0u;
  _SplatViewData.p = (SplatViewData *)out;
  for (uint i = 0; i < a->splat_count; ++i) CSCalcViewData(uint3(i, 0, 0));
  return 0;
}

// EditExportData (:936-958) -> CSExportData; bake != 0 also exercises RotateSH / CalcSHRotMatrix
REF_API int refhlsl_export(const GsoAsset *a, const GsoFrame *f, int bake, const float rot_xyzw[4], const float scale[3], float *out62) {
  using namespace refcs;
  if (bind_asset(a)) return -1;
  static_assert(sizeof(ExportSplatData) == 248, "InputSplatData layout");
  _MatrixObjectToWorld = from_colmajor(f->mat_object_to_world);
  _ExportTransformFlags = bake ? 1u : 0u;
  if (bake) {
    _ExportTransformRotation = float4(rot_xyzw[0], rot_xyzw[1], rot_xyzw[2], rot_xyzw[3]);
    _ExportTransformScale = float3(scale[0], scale[1], scale[2]);
  }
  static GaussianCutoutShaderData cuts[64];
  const uint nc = f->cutouts ? (f->cutout_count < 64 ? f->cutout_count : 64) : 0;
  for (uint i = 0; i < nc; ++i) { cuts[i].mat = from_colmajor(f->cutouts[i].mat); cuts[i].typeAndFlags = f->cutouts[i].type_and_flags; }
  _SplatCutouts.p = cuts;
  _SplatCutoutsCount = nc;
  _ExportBuffer.p = (ExportSplatData *)out62;
  for (uint i = 0; i < a->splat_count; ++i) CSExportData(uint3(i, 0, 0));
  return 0;
}

// The draw's programmable stages for ONE splat: vert for the 4 quad corners, frag at a given interpolated quad position.
// (Which pixels a quad covers and the blend are fixed-function in the reference.)
REF_API void refhlsl_vert(const GsoView *views, uint32_t *order, uint32_t inst, float screen_w, float screen_h, float out_clip[4][4],
                          float out_pos[4][2], float out_col[4]) {
  using namespace refps;
  _OrderBuffer.p = order;
  _SplatViewData.p = (SplatViewData *)views;
  _SplatBitsValid = 0;
  _ScreenParams = float4(screen_w, screen_h, 0, 0);
  _CameraTargetTexture_TexelSize = float4(0, 0, 0, 0);
  for (uint v = 0; v < 4; ++v) {
    v2f o = vert(v, inst);
    for (int k = 0; k < 4; ++k) out_clip[v][k] = o.vertex[k];
    out_pos[v][0] = o.pos.x; out_pos[v][1] = o.pos.y;
    for (int k = 0; k < 4; ++k) out_col[k] = o.col[k];
  }
}
// vert with an edit selection bound: _SplatBitsValid = 1 and _SplatSelectedBits = bits (S/RenderGaussianSplats.shader:63-73)
REF_API void refhlsl_vert_sel(const GsoView *views, uint32_t *order, uint32_t inst, float screen_w, float screen_h, const uint32_t *selected_bits,
                              float out_col[4]) {
  using namespace refps;
  _OrderBuffer.p = order;
  _SplatViewData.p = (SplatViewData *)views;
  _SplatSelectedBits.p = (const uint8_t *)selected_bits;
  _SplatBitsValid = selected_bits ? 1u : 0u;
  _ScreenParams = float4(screen_w, screen_h, 0, 0);
  _CameraTargetTexture_TexelSize = float4(0, 0, 0, 0);
  v2f o = vert(0, inst);
  for (int k = 0; k < 4; ++k) out_col[k] = o.col[k];
  _SplatBitsValid = 0;
}
// PackSmallest3Rotation + EncodeQuatToNorm10 (S/GaussianSplatting.hlsl:231-259,301-304): the HLSL twins of the importer's
// C# helpers (R/GaussianUtils.cs:46-76, E/GaussianSplatAssetCreator.cs:717-725)
REF_API uint32_t refhlsl_pack_rotation(const float q_xyzw[4], float packed[4]) {
  using namespace refcs;
  const float4 p = PackSmallest3Rotation(float4(q_xyzw[0], q_xyzw[1], q_xyzw[2], q_xyzw[3]));
  for (int k = 0; k < 4; ++k) packed[k] = p[k];
  return EncodeQuatToNorm10(p);
}
REF_API void refhlsl_decode_rotation(uint32_t enc, float q_xyzw[4]) {
  using namespace refcs;
  const float4 q = DecodeRotation(DecodePacked_10_10_10_2(enc));
  for (int k = 0; k < 4; ++k) q_xyzw[k] = q[k];
}
REF_API int refhlsl_frag(const float col[4], float pos_x, float pos_y, float out_rgba[4]) {
  using namespace refps;
  v2f i = v2f();
  i.col = float4(col[0], col[1], col[2], col[3]);
  i.pos = float2(pos_x, pos_y);
  g_discarded = false;
  const float4 r = frag(i);
  for (int k = 0; k < 4; ++k) out_rgba[k] = r[k];
  return g_discarded ? 1 : 0;
}

}  // extern "C"

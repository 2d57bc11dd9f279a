/* gsplat_asset.h -- C ABI of the host-side asset packer and synthetic-scene generator.
 *
 * Host (CPU) code, no CUDA.  It reproduces the byte layout written by the reference's
 * importer (package/Editor/GaussianSplatAssetCreator.cs) so that
 * synthetic scenes enter the render path in exactly the format a Unity-made
 * GaussianSplatAsset would.  Citations: E/ = package/Editor, R/ = package/Runtime.
 */
#ifndef GSPLAT_ASSET_H
#define GSPLAT_ASSET_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSA_API __attribute__((visibility("default")))

/* E/Utils/GaussianFileReader.cs:17-26 -- 62 floats / 248 bytes, AFTER LinearizeData
 * (:210-232): rot = PackSmallest3Rotation output (xyz in 0..1, w = index/3),
 * scale linear, dc0 = colour, opacity = sigmoid. */
typedef struct GsaInputSplat {
  float pos[3];
  float nor[3];
  float dc0[3];
  float sh[45]; /* sh1..shF, each rgb */
  float opacity;
  float scale[3];
  float rot[4];
} GsaInputSplat;

typedef struct GsaSizes {
  uint64_t pos_bytes, other_bytes, color_bytes, sh_bytes, chunk_bytes;
  uint32_t tex_width, tex_height;
} GsaSizes;

enum GsaSceneKind {
  GSA_SCENE_LATTICE = 0,   /* BASELINE cfg1: axis-aligned lattice with deliberate depth ties */
  GSA_SCENE_CLUSTERED = 1, /* cfg2-4 "scene-like": clusters + background shell */
  GSA_SCENE_UNIFORM = 2    /* cfg5 "random gaussians" */
};

/* Deterministic synthetic splats.  Writes n GsaInputSplat records. */
GSA_API int gsa_generate(uint32_t kind, uint32_t n, uint32_t seed, GsaInputSplat *out);

/* R/GaussianSplatAsset.cs:152-203 size formulas (+ the 8-byte padding of pos/other,
 * E/GaussianSplatAssetCreator.cs:815,842). */
GSA_API int gsa_calc_sizes(uint32_t n, uint32_t pos_fmt, uint32_t scale_fmt, uint32_t color_fmt,
                           uint32_t sh_fmt, GsaSizes *out);

/* CreateAsset (E/GaussianSplatAssetCreator.cs:248-330): bounds, Morton reorder
 * (:387-429), chunk min/max + normalise (:520-639) when any stream is lossy (:54-58),
 * then the five blobs (:705-1066).  `splats` is reordered and modified in place, exactly
 * like the reference's NativeArray.  Output buffers must have the sizes gsa_calc_sizes
 * reports; `chunks` may be NULL when the format set is fully float32.
 * color_fmt 3 (BC7) and sh_fmt 4..8 (Cluster64k..4k) are the VeryLow / Low presets' formats (E/...:195-206).
 * bounds_out: 6 floats (min xyz, max xyz) or NULL. */
GSA_API int gsa_create_asset(GsaInputSplat *splats, uint32_t n, uint32_t pos_fmt, uint32_t scale_fmt,
                             uint32_t color_fmt, uint32_t sh_fmt, void *pos, void *other, void *color,
                             void *sh, void *chunks, float *bounds_out);

/* INRIA .ply input: header parse + attribute mapping (E/Utils/PLYFileReader.cs:24-67,
 * E/Utils/GaussianFileReader.cs:45-169), SH re-interleave (:183-205) and LinearizeData (:207-232).
 * gsa_ply_vertex_count returns the vertex count (< 0: not a binary-little-endian gaussian-splat ply).
 * gsa_ply_read fills `capacity` >= count records (already linearised, ready for gsa_create_asset). */
GSA_API int64_t gsa_ply_vertex_count(const char *path);
GSA_API int gsa_ply_read(const char *path, GsaInputSplat *out, uint32_t capacity);

/* .ply output (E/GaussianSplatRendererEditor.cs:394-445, ExportPlyFile): `records` = n x 62 raw attribute values (what
 * gs_export_splats returns); records with a deleted bit (may be NULL) or a non-zero normal (cut marker) are dropped.
 * Returns the number of vertices written, < 0 on I/O error. */
GSA_API int64_t gsa_ply_write(const char *path, const float *records, uint32_t n, const uint32_t *deleted_bits);

/* "Bake transform" of an export (the _ExportTransformFlags branch of CSExportData, S/SplatUtilities.compute:626-643), as a
 * host pass over gs_export_splats' records (n x 62 raw attribute values, in place): position by `o2w` (column-major 4x4 =
 * tr.localToWorldMatrix), orientation by `rot_xyzw` (tr.localRotation; axis flips for negative scale as in the
 * reference), log-scale by |scale| (tr.localScale), SH bands 1..3 rotated by o2w's normalised 3x3 (CalcSHRotMatrix). */
GSA_API int gsa_bake_transform(float *records, uint32_t n, const float o2w[16], const float rot_xyzw[4], const float scale[3]);

/* Niantic/Scaniverse .spz input (gzip stream, version 2; E/Utils/SPZFileReader.cs:20-195): same contract as the
 * ply pair above.  Unlike PLY the records need no LinearizeData pass: the unpack already yields linear values. */
GSA_API int64_t gsa_spz_vertex_count(const char *path);
GSA_API int gsa_spz_read(const char *path, GsaInputSplat *out, uint32_t capacity);

/* Mini-batch k-means with k-means++ seeding as the importer runs it for the Cluster4k..64k SH formats
 * (E/Utils/KMeansClustering.cs:29-136, call site E/GaussianSplatAssetCreator.cs:476-518: dim 45, batch 2048,
 * passes 0.3..1.2).  data: data_size x dim floats; out_means: k x dim; out_labels: data_size.  Deterministic. */
GSA_API int gsa_kmeans(uint32_t dim, const float *data, uint32_t data_size, uint32_t batch_size, float passes_over_data,
                       float *out_means, uint32_t k, int32_t *out_labels);

/* One 4x4 block of float RGBA in [0,1] (raster order) -> 16 bytes of BC7 (mode 6).  Stands in for Unity's
 * EditorUtility.CompressTexture (E/GaussianSplatAssetCreator.cs:901-912), which is closed source: lossy and not
 * bit-identical to it, but any conformant BC7 decoder (the texture unit, gs_bc7.cuh) reads it. */
GSA_API void gsa_bc7_encode_block(const float rgba[64], uint8_t out[16]);

/* Individual pieces, exposed for tests. */
GSA_API uint64_t gsa_morton_encode3(uint32_t x, uint32_t y, uint32_t z);       /* R/GaussianUtils.cs:81-95 */
GSA_API uint32_t gsa_splat_index_to_texture_index(uint32_t idx);              /* E/GaussianSplatAssetCreator.cs:863-871 */
GSA_API void gsa_pack_smallest3(const float q_xyzw[4], float out[4]);         /* R/GaussianUtils.cs:46-76 */
GSA_API uint32_t gsa_f32tof16(float v);                                       /* Unity.Mathematics math.f32tof16 */

#ifdef __cplusplus
}
#endif
#endif

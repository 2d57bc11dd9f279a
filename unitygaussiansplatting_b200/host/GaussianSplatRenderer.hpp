// GaussianSplatRenderer.hpp -- header-only C++ host mirror of the reference's C# component
// (package/Runtime/GaussianSplatRenderer.cs) over the C ABI in include/gsplat_b200.h.
// The reference's host language (C# / Unity) has no toolchain in this image; C++ is the compiled
// host language used instead.  Same member names and call order as the C# class for the hot path;
// errors never throw: like the reference (:361-369,:655) a failed call logs and returns false.
#pragma once
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/gsplat_b200.h"

namespace GaussianSplatting {

struct Matrix4x4 { float m[16]; };  // column-major, like UnityEngine.Matrix4x4

struct CameraState {               // what the C# code reads from UnityEngine.Camera
  Matrix4x4 worldToCameraMatrix;   // :586
  Matrix4x4 gpuProjectionMatrix;   // GL.GetGPUProjectionMatrix(cam.projectionMatrix, true)
  float pixelWidth, pixelHeight;   // :589-591
  float position[3];               // :592
};

class GaussianSplatRenderer {
 public:
  // serialized knobs, :225-251
  float m_SplatScale = 1.0f, m_OpacityScale = 1.0f;
  int m_SHOrder = 3;
  bool m_SHOnly = false;
  int m_SortNthFrame = 1;
  Matrix4x4 localToWorldMatrix{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}};
  Matrix4x4 worldToLocalMatrix{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}};
  std::vector<GsCutout> m_Cutouts;
  GsRenderOptions options{};

  ~GaussianSplatRenderer() { OnDisable(); }

  // OnEnable -> CreateResourcesForAsset, :373-445,:475
  bool OnEnable(const GsAssetDesc &asset, int cudaDevice = 0, void *stream = nullptr) {
    OnDisable();
    if (!ok(gs_create(cudaDevice, stream, &m_Ctx))) return false;
    if (!ok(gs_asset_upload(m_Ctx, &asset, &m_Asset))) return false;
    m_FrameCounter = 0;
    return true;
  }
  void OnDisable() {  // :533-577
    if (m_Asset) gs_asset_destroy(m_Asset);
    if (m_Ctx) gs_destroy(m_Ctx);
    m_Asset = nullptr;
    m_Ctx = nullptr;
  }
  bool HasValidRenderSetup() const { return m_Ctx && m_Asset; }                           // :369
  bool SortPoints(const CameraState &cam) { GsFrameParams fp = params(cam); return ok(gs_sort(m_Ctx, m_Asset, &fp)); }          // :612
  bool CalcViewData(const CameraState &cam) { GsFrameParams fp = params(cam); return ok(gs_calc_view(m_Ctx, m_Asset, &fp)); }   // :579
  bool DrawSplats(const CameraState &cam, GsImage *rt) { GsFrameParams fp = params(cam); return ok(gs_render(m_Ctx, m_Asset, &fp, &options, rt)); }  // :165
  bool Composite(const GsImage *rt, GsImage *cameraTarget) { return ok(gs_composite(m_Ctx, rt, cameraTarget)); }               // :206-210
  // one renderer's share of GaussianSplatRenderSystem.SortAndRenderSplats, :108-169
  bool SortAndRenderSplats(const CameraState &cam, GsImage *rt, GsImage *cameraTarget = nullptr) {
    const int doSort = (m_FrameCounter % (m_SortNthFrame > 0 ? m_SortNthFrame : 1)) == 0;  // :120-121
    ++m_FrameCounter;
    GsFrameParams fp = params(cam);
    return ok(gs_frame(m_Ctx, m_Asset, &fp, &options, doSort, rt, cameraTarget));
  }
  bool Sync() { return ok(gs_sync(m_Ctx)); }
  // EditExportData, R/GaussianSplatRenderer.cs:936-958: n x 62 floats (the .ply attribute record), cut splats marked by nor = 1
  bool EditExportData(float *dstRecords, bool bakeTransform = false) {
    return ok(gs_export_splats(m_Ctx, m_Asset, m_Cutouts.empty() ? nullptr : m_Cutouts.data(), (uint32_t)m_Cutouts.size(), bakeTransform ? 1u : 0u,
                               dstRecords));
  }
  GsContext *context() const { return m_Ctx; }
  GsAsset *asset() const { return m_Asset; }

 private:
  GsFrameParams params(const CameraState &cam) const {  // CalcViewData / SortPoints uniform binding, :586-606,:617-631
    GsFrameParams fp;
    std::memset(&fp, 0, sizeof(fp));
    std::memcpy(fp.mat_object_to_world, localToWorldMatrix.m, 64);
    std::memcpy(fp.mat_world_to_object, worldToLocalMatrix.m, 64);
    std::memcpy(fp.mat_view, cam.worldToCameraMatrix.m, 64);
    std::memcpy(fp.mat_proj_gpu, cam.gpuProjectionMatrix.m, 64);
    fp.screen_w = cam.pixelWidth;
    fp.screen_h = cam.pixelHeight;
    std::memcpy(fp.cam_pos_world, cam.position, 12);
    fp.splat_scale = m_SplatScale;
    fp.opacity_scale = m_OpacityScale;
    fp.sh_order = (uint32_t)m_SHOrder;
    fp.sh_only = m_SHOnly ? 1u : 0u;
    fp.cutouts = m_Cutouts.empty() ? nullptr : m_Cutouts.data();
    fp.cutout_count = (uint32_t)m_Cutouts.size();
    return fp;
  }
  bool ok(int rc) const {
    if (rc == GS_OK) return true;
    std::fprintf(stderr, "GaussianSplatRenderer: %s (%d)\n", gs_last_error(m_Ctx), rc);  // Debug.LogError + skip
    return false;
  }
  GsContext *m_Ctx = nullptr;
  GsAsset *m_Asset = nullptr;
  int m_FrameCounter = 0;
};

// The same renderer on several GPUs of one box, one process driving all of them (gs_group_create / gs_group_frame):
// identical knobs; every non-null image receives the complete frame, bit-identical to what one GPU renders.
class GaussianSplatRendererGroup {
 public:
  float m_SplatScale = 1.0f, m_OpacityScale = 1.0f;
  int m_SHOrder = 3;
  bool m_SHOnly = false;
  int m_SortNthFrame = 1;
  Matrix4x4 localToWorldMatrix{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}};
  Matrix4x4 worldToLocalMatrix{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}};
  GsRenderOptions options{};

  ~GaussianSplatRendererGroup() { OnDisable(); }

  bool OnEnable(const GsAssetDesc &asset, const std::vector<int> &cudaDevices, uint32_t flags = 0) {
    OnDisable();
    if (!ok(gs_group_create(cudaDevices.data(), (uint32_t)cudaDevices.size(), flags, &m_Group))) return false;
    m_Assets.assign(gs_group_local_count(m_Group), nullptr);
    if (!ok(gs_group_asset_upload(m_Group, &asset, m_Assets.data()))) return false;
    m_FrameCounter = 0;
    return true;
  }
  void OnDisable() {
    for (GsAsset *a : m_Assets) if (a) gs_asset_destroy(a);   // the library owns the contexts: assets go first
    m_Assets.clear();
    if (m_Group) gs_group_destroy(m_Group);
    m_Group = nullptr;
  }
  // rts[i]: the image of local member i (device or host memory) or nullptr
  bool SortAndRenderSplats(const CameraState &cam, GsImage *const *rts) {
    const int doSort = (m_FrameCounter % (m_SortNthFrame > 0 ? m_SortNthFrame : 1)) == 0;
    ++m_FrameCounter;
    GsFrameParams fp;
    std::memset(&fp, 0, sizeof(fp));
    std::memcpy(fp.mat_object_to_world, localToWorldMatrix.m, 64);
    std::memcpy(fp.mat_world_to_object, worldToLocalMatrix.m, 64);
    std::memcpy(fp.mat_view, cam.worldToCameraMatrix.m, 64);
    std::memcpy(fp.mat_proj_gpu, cam.gpuProjectionMatrix.m, 64);
    fp.screen_w = cam.pixelWidth;
    fp.screen_h = cam.pixelHeight;
    std::memcpy(fp.cam_pos_world, cam.position, 12);
    fp.splat_scale = m_SplatScale;
    fp.opacity_scale = m_OpacityScale;
    fp.sh_order = (uint32_t)m_SHOrder;
    fp.sh_only = m_SHOnly ? 1u : 0u;
    return ok(gs_group_frame(m_Group, m_Assets.data(), &fp, &options, doSort, rts));
  }
  bool Sync() { return ok(gs_group_sync(m_Group)); }
  GsGroup *group() const { return m_Group; }
  GsAsset *asset(size_t i) const { return m_Assets[i]; }

 private:
  bool ok(int rc) const {
    if (rc == GS_OK) return true;
    std::fprintf(stderr, "GaussianSplatRendererGroup: %s (%d)\n", gs_last_error(m_Group ? gs_group_context(m_Group, 0) : nullptr), rc);
    return false;
  }
  GsGroup *m_Group = nullptr;
  std::vector<GsAsset *> m_Assets;
  int m_FrameCounter = 0;
};

}  // namespace GaussianSplatting

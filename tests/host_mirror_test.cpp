// Compiles the C++ host mirror against the C ABI and drives one tiny frame.  Exit codes:
// 0 = frame rendered, 3 = no CUDA device (expected on the CPU box: the library must fail loudly), 1 = anything else.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../unitygaussiansplatting_b200/host/GaussianSplatRenderer.hpp"

int main() {
  using namespace GaussianSplatting;
  const uint32_t n = 256;
  // one VeryHigh (all float32, no chunks) asset: splats on a ring in front of the camera
  std::vector<float> pos(n * 3), other(n * 4), color(2048 * 16 * 4, 0.0f), sh(n * 48, 0.0f);
  for (uint32_t i = 0; i < n; ++i) {
    float a = 6.2831853f * i / n;
    pos[i * 3] = 0.8f * std::cos(a); pos[i * 3 + 1] = 0.8f * std::sin(a); pos[i * 3 + 2] = 0.0f;
    uint32_t q = 511u | (511u << 10) | (511u << 20) | (3u << 30);   // ~identity rotation, 10.10.10.2
    std::memcpy(&other[i * 4], &q, 4);
    other[i * 4 + 1] = other[i * 4 + 2] = other[i * 4 + 3] = 0.05f;
    uint32_t t = (i & 0xFF) | ((i & 0xFE) << 7); t &= 0x5555; t = (t ^ (t >> 1)) & 0x3333; t = (t ^ (t >> 2)) & 0x0f0f;
    uint32_t tx = t & 0xF, ty = t >> 8;
    float *c = &color[(ty * 2048 + tx) * 4];
    c[0] = 1.0f; c[1] = 0.5f; c[2] = 0.25f; c[3] = 0.9f;
  }
  GsAssetDesc d{};
  d.splat_count = n;
  d.pos = pos.data(); d.pos_bytes = pos.size() * 4;
  d.other = other.data(); d.other_bytes = other.size() * 4;
  d.color = color.data(); d.color_bytes = color.size() * 4;
  d.sh = sh.data(); d.sh_bytes = sh.size() * 4;
  GaussianSplatRenderer r;
  if (!r.OnEnable(d)) return r.context() == nullptr ? 3 : 1;
  CameraState cam{};
  const float f = 1.0f / std::tan(0.5f * 0.8f), zn = 0.3f, zf = 100.0f;
  float view[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, -3, 1};                 // camera at (0,0,-3) looking +z; row 2 negated
  float proj[16] = {f, 0, 0, 0, 0, -f, 0, 0, 0, 0, zn / (zf - zn), -1, 0, 0, zf * zn / (zf - zn), 0};   // reversed-z, y flipped
  std::memcpy(cam.worldToCameraMatrix.m, view, 64);
  std::memcpy(cam.gpuProjectionMatrix.m, proj, 64);
  cam.pixelWidth = 128; cam.pixelHeight = 128; cam.position[2] = -3.0f;
  std::vector<uint16_t> rt(128 * 128 * 4);
  GsImage im{rt.data(), 128, 128, 0, GS_PIX_RGBA16F, GS_MEM_HOST};
  if (!r.SortAndRenderSplats(cam, &im)) return 1;
  size_t lit = 0;
  for (size_t i = 3; i < rt.size(); i += 4) lit += rt[i] != 0;
  std::printf("host mirror: %zu lit pixels\n", lit);
  if (lit <= 100) return 1;
  // the same frame through the group API: three contexts emulated on this device; every member's image must equal the
  // single-GPU frame byte for byte
  GaussianSplatRendererGroup grp;
  if (!grp.OnEnable(d, std::vector<int>{0, 0, 0}, GS_GROUP_EMULATE)) return 1;
  std::vector<uint16_t> g0(rt.size()), g1(rt.size()), g2(rt.size());
  GsImage i0{g0.data(), 128, 128, 0, GS_PIX_RGBA16F, GS_MEM_HOST}, i1{g1.data(), 128, 128, 0, GS_PIX_RGBA16F, GS_MEM_HOST},
      i2{g2.data(), 128, 128, 0, GS_PIX_RGBA16F, GS_MEM_HOST};
  GsImage *rts[3] = {&i0, &i1, &i2};
  if (!grp.SortAndRenderSplats(cam, rts)) return 1;
  if (g0 != rt || g1 != rt || g2 != rt) { std::fprintf(stderr, "group frame differs from the single-GPU frame\n"); return 1; }
  std::printf("host mirror: group of 3 equals one GPU\n");
  return 0;
}

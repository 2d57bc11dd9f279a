// Host build of the product's single-texel BC7 decoder (csrc/gs_bc7.cuh is plain C++), exported for tests/test_bc7.py.
#include "../unitygaussiansplatting_b200/csrc/gs_bc7.cuh"
#include <cstring>
extern "C" __attribute__((visibility("default"))) void bc7_texels(const uint8_t *blocks, uint32_t nblocks, uint8_t *out_rgba) {
  for (uint32_t b = 0; b < nblocks; ++b) {
    uint32_t w[4];
    std::memcpy(w, blocks + (size_t)b * 16, 16);
    for (uint32_t p = 0; p < 16; ++p) {
      const uint32_t v = gs::bc7::decode_texel(w[0], w[1], w[2], w[3], p);
      std::memcpy(out_rgba + ((size_t)b * 16 + p) * 4, &v, 4);
    }
  }
}

// Bit-sliced uint8 template layout shared by view_templates.cu and sharded.cu (see the layout comment in
// view_templates.cu, "Bit-sliced ("packed") uint8 library").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// one group of 32 templates: uint4 planes[32 rows][2 halves][32 lanes] + uint4 rowsum[4][32 lanes]
constexpr int kVtGroupU4 = 32 * 2 * 32 + 4 * 32;  // 2176 uint4 = 34816 bytes

// Thread t (0..31) packs row t of the row-major template `tpl` into slot `ti` of the packed library
// (ViewTemplates' append, view_templates.py:68-71).
__device__ __forceinline__ void vt_pack_row(const uint8_t* __restrict__ tpl, uint4* __restrict__ packed, long long ti, int t) {
  const uint32_t* row = reinterpret_cast<const uint32_t*>(tpl + t * 32);
  uint32_t pl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t sum = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const uint32_t v = row[w];
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const uint32_t px = (v >> (8 * bb)) & 0xffu;
      sum += px;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) pl[kk] |= ((px >> kk) & 1u) << (w * 4 + bb);
    }
  }
  uint4* grp = packed + (size_t)(ti >> 5) * kVtGroupU4;
  const int lane = (int)(ti & 31);
  grp[(t * 2 + 0) * 32 + lane] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  grp[(t * 2 + 1) * 32 + lane] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
  uint16_t* rs = reinterpret_cast<uint16_t*>(grp + 32 * 2 * 32);
  const int w = t >> 1;
  rs[(((w >> 2) * 32 + lane) * 4 + (w & 3)) * 2 + (t & 1)] = (uint16_t)sum;
}

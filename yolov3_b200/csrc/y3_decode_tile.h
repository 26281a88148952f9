// yolov3_b200 — tile arithmetic of the staged Detect decode (head_decode2_kernel, y3_detect.cu).  Plain functions of the
// tile / element index so that the SAME code is compiled into the kernel and into the host-side index test
// (tests/emul/decode_tile_emul.cpp, built with g++: no GPU needed to check which element lands where).
//
// A tile = kTileCells consecutive cells of one image and one pyramid level: kTileCells * head_ld contiguous floats of the
// head conv's fp32 pixel-major output [bs*ny*nx, head_ld] (column a*no + k), decoded into na chunks of kTileCells * no
// contiguous floats of z [bs, rows, no] (row = row_off_l + a*plane + cell, models/yolo.py:100-110).
#pragma once
#include <stdint.h>

#include "../../include/yolov3_b200.h"

#ifdef __CUDACC__
#define Y3_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define Y3_HD inline
#endif

namespace y3 {

constexpr int kTileCells = 16;

struct HeadDecodeArgs {
  const float* head[Y3_MAX_LEVELS];
  float* raw[Y3_MAX_LEVELS];
  int head_ld[Y3_MAX_LEVELS];
  int ny[Y3_MAX_LEVELS], nx[Y3_MAX_LEVELS];
  int row_off[Y3_MAX_LEVELS + 1];
  float stride[Y3_MAX_LEVELS];
  float anchor_w[Y3_MAX_LEVELS][Y3_MAX_ANCHORS], anchor_h[Y3_MAX_LEVELS][Y3_MAX_ANCHORS];  // pixels
  int nl, bs, na, no;
  float* z;
};

// staged kernel only: all levels share head_ld = 4 << ld4_shift; plane_l % kTileCells == 0
struct HeadDecode2Args {
  HeadDecodeArgs a;
  int tile_off[Y3_MAX_LEVELS + 1];  // first tile of each level; [nl] = number of tiles
  int ld4_shift;                    // log2(head_ld / 4)
};

struct DecodeTile {
  int l, b, cell0;
  long long src_f4;  // float4 index of the tile's first element in head[l]
  long long zrow0;   // z row (over the whole batch) of (anchor 0, cell0)
};

Y3_HD DecodeTile decode_tile(const HeadDecode2Args& p, int t) {
  DecodeTile ti;
  int l = 0;
  while (l + 1 < p.a.nl && t >= p.tile_off[l + 1]) ++l;
  const int plane = p.a.ny[l] * p.a.nx[l];
  const long long pix0 = static_cast<long long>(t - p.tile_off[l]) * kTileCells;  // pixel of the level's [bs*plane] list
  ti.l = l;
  ti.b = static_cast<int>(pix0 / plane);
  ti.cell0 = static_cast<int>(pix0 - static_cast<long long>(ti.b) * plane);
  ti.src_f4 = pix0 << p.ld4_shift;
  ti.zrow0 = static_cast<long long>(ti.b) * p.a.row_off[p.a.nl] + p.a.row_off[l] + ti.cell0;
  return ti;
}

// sigmoid through the fast exp/divide units (relative error ~1e-6, inside the 1e-5 decode tolerance); the host build of the
// index test uses the libm forms
Y3_HD float decode_sigmoid(float x) {
#ifdef __CUDA_ARCH__
  return __fdividef(1.0f, 1.0f + __expf(-x));
#else
  return 1.0f / (1.0f + expf(-x));
#endif
}

// one output element: k = field of the row, (x, y) = grid cell.  Operation order of models/yolo.py:104-109
// (xy * 2 + grid) * stride with grid = index - 0.5, (wh * 2) ** 2 * anchor_grid — compiled without FMA contraction.
Y3_HD float decode_value(float v, int k, int x, int y, float stride, float aw, float ah) {
  const float s = decode_sigmoid(v);
  if (k >= 4) return s;
  const float t2 = s * 2.0f;
  if (k < 2) return (t2 + (static_cast<float>(k == 0 ? x : y) - 0.5f)) * stride;
  return (t2 * t2) * (k == 2 ? aw : ah);
}

// A thread's four channels.  float4 number f of a tile (cell-major, head_ld/4 per cell) holds channels c0 .. c0+3 of cell
// cl = f >> ld4_shift with c0 = (f & (head_ld/4 - 1)) * 4; a block of 256 threads walks f = tid, tid + 256, ... and head_ld/4
// divides 256 (launcher), so c0 — and with it (anchor, field) of each of the four values — is fixed per thread.
struct DecodeLane {
  int base[4];  // staging index of channel c0+i for cell 0 of the tile: a*kTileCells*no + k; -1: padding channel (c >= na*no)
  int k[4];     // field of the row (0..3 = box), or 4 for every class / objectness field
  int a[4];
  int anybox;   // some k[i] < 4: this thread needs the grid position of its cells
};
Y3_HD DecodeLane decode_lane(const HeadDecode2Args& p, int tid) {
  DecodeLane ln;
  ln.anybox = 0;
  const int c0 = (tid & ((1 << p.ld4_shift) - 1)) * 4;
  const int no = p.a.no, nch = p.a.na * no;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + i;
    const int a = c / no, k = c - a * no;
    ln.a[i] = a < p.a.na ? a : 0;
    ln.k[i] = k < 4 ? k : 4;
    ln.base[i] = c < nch ? a * kTileCells * no + k : -1;
    if (c < nch && k < 4) ln.anybox = 1;
  }
  return ln;
}

// decode the four values of float4 number f and put them at their z position inside the staging buffer
// ([na][kTileCells][no] floats).  Returns the number of values written (test hook).
Y3_HD int decode_stage4(const HeadDecode2Args& p, const DecodeTile& ti, const DecodeLane& ln, int f, float v0, float v1,
                        float v2, float v3, float* staging) {
  const int cl = f >> p.ld4_shift;
  const int no = p.a.no;
  const float v[4] = {v0, v1, v2, v3};
  int x = 0, y = 0;
  if (ln.anybox) {  // box fields are 4 of the no channels of an anchor: few threads work out the grid position
    const int nx = p.a.nx[ti.l];
    const int cell = ti.cell0 + cl;
    y = cell / nx;
    x = cell - y * nx;
  }
  int written = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int i = 0; i < 4; ++i) {
    if (ln.base[i] < 0) continue;
    const float o = ln.k[i] >= 4 ? decode_sigmoid(v[i])
                                 : decode_value(v[i], ln.k[i], x, y, p.a.stride[ti.l], p.a.anchor_w[ti.l][ln.a[i]],
                                                p.a.anchor_h[ti.l][ln.a[i]]);
    staging[ln.base[i] + cl * no] = o;
    ++written;
  }
  return written;
}

// float4 number idx of the staging buffer ([na][kTileCells][no] floats, read linearly) -> float4 index in z
Y3_HD long long decode_out4(const HeadDecode2Args& p, const DecodeTile& ti, int idx) {
  const int per_a4 = kTileCells / 4 * p.a.no;  // float4 per anchor chunk
  const int a = idx / per_a4, q = idx - a * per_a4;
  const int plane = p.a.ny[ti.l] * p.a.nx[ti.l];
  // (zrow0 + a*plane) is a multiple of 4 (checked by the launcher), so the chunk starts on a 16-byte boundary
  return ((ti.zrow0 + static_cast<long long>(a) * plane) >> 2) * p.a.no + q;
}

}  // namespace y3

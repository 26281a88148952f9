// Test infrastructure (CPU): runs the index arithmetic of the staged Detect decode kernel (yolov3_b200/csrc/y3_decode_tile.h,
// the same header head_decode2_kernel compiles) with the kernel's own loop structure — tiles over blocks, f = tid, tid+256, ...
// — on the host, so that tests/test_emul_cpu.py can compare z with the oracle decode and check that every staging slot and
// every z vector is written exactly once.  Built with g++ by the test; never part of the product library.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../yolov3_b200/csrc/y3_decode_tile.h"

struct F4 {
  float x, y, z, w;
};

// returns 0 on success; > 0: a staging slot or z vector was written twice / never (the code says which check failed)
extern "C" int decode2_emul(int nl, int bs, int na, int no, int ld, const int* ny, const int* nx, const float* stride,
                            const float* anchor_wh /*[nl][na][2] pixels*/, const float* const* head, float* z, int grid) {
  y3::HeadDecode2Args p;
  memset(&p, 0, sizeof(p));
  p.a.nl = nl;
  p.a.bs = bs;
  p.a.na = na;
  p.a.no = no;
  p.a.z = z;
  int off = 0, tiles = 0;
  int sh = 0;
  while ((4 << sh) < ld) ++sh;
  if ((4 << sh) != ld || ld > 1024) return 100;
  for (int l = 0; l < nl; ++l) {
    p.a.head[l] = head[l];
    p.a.head_ld[l] = ld;
    p.a.ny[l] = ny[l];
    p.a.nx[l] = nx[l];
    p.a.stride[l] = stride[l];
    p.a.row_off[l] = off;
    off += na * ny[l] * nx[l];
    for (int j = 0; j < na; ++j) {
      p.a.anchor_w[l][j] = anchor_wh[(l * na + j) * 2];
      p.a.anchor_h[l][j] = anchor_wh[(l * na + j) * 2 + 1];
    }
    if ((ny[l] * nx[l]) % y3::kTileCells) return 101;
    p.tile_off[l] = tiles;
    tiles += bs * ny[l] * nx[l] / y3::kTileCells;
  }
  p.a.row_off[nl] = off;
  p.tile_off[nl] = tiles;
  p.ld4_shift = sh;
  const int nf4 = y3::kTileCells << sh;
  const int nout4 = na * (y3::kTileCells / 4) * no;
  std::vector<float> staging(static_cast<size_t>(nout4) * 4);
  std::vector<int> hits(staging.size());
  std::vector<unsigned char> zhits(static_cast<size_t>(off) * bs * no / 4, 0);
  F4* z4 = reinterpret_cast<F4*>(z);
  for (int block = 0; block < grid; ++block) {
    for (int t = block; t < tiles; t += grid) {
      const y3::DecodeTile ti = y3::decode_tile(p, t);
      const F4* src = reinterpret_cast<const F4*>(p.a.head[ti.l]) + ti.src_f4;
      std::fill(hits.begin(), hits.end(), 0);
      std::fill(staging.begin(), staging.end(), -12345.f);
      for (int tid = 0; tid < 256; ++tid) {
        const y3::DecodeLane ln = y3::decode_lane(p, tid);
        for (int f = tid; f < nf4; f += 256) {
          const F4 v = src[f];
          // count the writes through a shadow pass: every slot the call fills differs from the sentinel afterwards
          std::vector<float> before;
          for (int i = 0; i < 4; ++i)
            if (ln.base[i] >= 0) before.push_back(staging[ln.base[i] + (f >> sh) * no]);
          const int w = y3::decode_stage4(p, ti, ln, f, v.x, v.y, v.z, v.w, staging.data());
          if (w != static_cast<int>(before.size())) return 1;
          for (float b : before)
            if (b != -12345.f) return 2;  // slot written twice
          for (int i = 0; i < 4; ++i)
            if (ln.base[i] >= 0) ++hits[ln.base[i] + (f >> sh) * no];
        }
      }
      for (size_t i = 0; i < hits.size(); ++i)
        if (hits[i] != 1) return 3;  // slot never written
      for (int tid = 0; tid < 256; ++tid)
        for (int idx = tid; idx < nout4; idx += 256) {
          const long long o = y3::decode_out4(p, ti, idx);
          if (o < 0 || o >= static_cast<long long>(zhits.size())) return 4;
          if (zhits[o]++) return 5;  // z vector written twice
          z4[o] = reinterpret_cast<const F4*>(staging.data())[idx];
        }
    }
  }
  for (unsigned char h : zhits)
    if (h != 1) return 6;  // z vector never written
  return 0;
}

// yolov3_b200 — Detect-head decode: sigmoid + anchor-grid transform of all pyramid levels in ONE coalesced pass.
// Replaces Detect.forward's eval branch and _make_grid (reference models/yolo.py:100-123): no grid / anchor_grid
// tensors, no split/cat; grid offsets come from the element index.  raw level l is fp32 [bs, na, ny, nx, no]
// (= the reference's x[i]); z is fp32 [bs, sum_l na*ny*nx, no] with row = off_l + (a*ny + y)*nx + x, i.e. exactly
// torch.cat(z, 1) of models/yolo.py:110.  Compiled without fast-math/FMA contraction: (2*s + g) * stride is evaluated
// with the reference's operation order and rounding.
#include "y3_common.cuh"
#include "y3_internal.h"

namespace y3 {
namespace {

struct DecodeArgs {
  const float* raw[Y3_MAX_LEVELS];
  int ny[Y3_MAX_LEVELS], nx[Y3_MAX_LEVELS];
  int row_off[Y3_MAX_LEVELS + 1];  // first z row of each level; [nl] = total rows
  float stride[Y3_MAX_LEVELS];
  float anchor_w[Y3_MAX_LEVELS][Y3_MAX_ANCHORS], anchor_h[Y3_MAX_LEVELS][Y3_MAX_ANCHORS];  // pixels
  int nl, bs, na, no;
  float* z;
};

__global__ void __launch_bounds__(256) decode_kernel(const DecodeArgs p) {
  const long long per_img = static_cast<long long>(p.row_off[p.nl]) * p.no;
  const long long total = per_img * p.bs;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / per_img);
    const long long e = i - b * per_img;
    const int row = static_cast<int>(e / p.no);
    const int k = static_cast<int>(e - static_cast<long long>(row) * p.no);
    int l = 0;
    while (l + 1 < p.nl && row >= p.row_off[l + 1]) ++l;
    const int r = row - p.row_off[l];
    const int plane = p.ny[l] * p.nx[l];
    const float v = p.raw[l][(static_cast<long long>(b) * p.na * plane + r) * p.no + k];
    const float s = 1.0f / (1.0f + expf(-v));
    float o = s;
    if (k < 4) {
      const int a = r / plane, cell = r - a * plane;
      const int y = cell / p.nx[l], x = cell - y * p.nx[l];
      if (k == 0)
        o = (s * 2.0f + (static_cast<float>(x) - 0.5f)) * p.stride[l];
      else if (k == 1)
        o = (s * 2.0f + (static_cast<float>(y) - 0.5f)) * p.stride[l];
      else {
        const float t = s * 2.0f;
        o = (t * t) * (k == 2 ? p.anchor_w[l][a] : p.anchor_h[l][a]);
      }
    }
    p.z[i] = o;
  }
}

// ---- fused variant used by the graph executor: reads the head convs' fp32 pixel-major output [bs*ny*nx, ld]
// (column a*no + k), writes z AND (optionally) the reference-layout logits raw_l[bs, na, ny, nx, no].
// One warp per (image, cell, anchor): 340-byte contiguous reads and writes.
struct HeadDecodeArgs {
  const float* head[Y3_MAX_LEVELS];
  float* raw[Y3_MAX_LEVELS];
  int head_ld[Y3_MAX_LEVELS];
  int ny[Y3_MAX_LEVELS], nx[Y3_MAX_LEVELS];
  int row_off[Y3_MAX_LEVELS + 1];
  float stride[Y3_MAX_LEVELS];
  float anchor_w[Y3_MAX_LEVELS][Y3_MAX_ANCHORS], anchor_h[Y3_MAX_LEVELS][Y3_MAX_ANCHORS];
  int nl, bs, na, no;
  float* z;
};

constexpr int kDecodeRows = 4;  // consecutive z rows per warp iteration (they are contiguous in z and in the logits)

// sigmoid through the fast exp/divide units: relative error ~1e-6, far inside the 1e-5 decode tolerance (the bit-exact
// requirement applies to NMS on a given z, not to z itself); the IEEE expf/division sequence cost ~1/3 of this kernel.
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

__global__ void __launch_bounds__(256) head_decode_kernel(const HeadDecodeArgs p) {
  const int lane = threadIdx.x & 31;
  const int rows_per_img = p.row_off[p.nl];
  const int total = rows_per_img * p.bs;  // < 2^31 (checked by the launcher)
  const int groups = (total + kDecodeRows - 1) / kDecodeRows;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int no = p.no;
  for (int gidx = warp0; gidx < groups; gidx += nwarps) {
    // per-row descriptors, computed by lanes 0..3 and broadcast (saves every lane redoing the integer divisions)
    const int w = gidx * kDecodeRows + (lane & 3);
    long long src_off = -1, raw_off = 0;
    int lxy = 0, lla = 0;
    if (lane < kDecodeRows && w < total) {
      const int b = w / rows_per_img;
      const int row = w - b * rows_per_img;
      int l = 0;
      while (l + 1 < p.nl && row >= p.row_off[l + 1]) ++l;
      const int r = row - p.row_off[l];
      const int plane = p.ny[l] * p.nx[l];
      const int a = r / plane, cell = r - a * plane;
      const int y = cell / p.nx[l], x = cell - y * p.nx[l];
      src_off = (static_cast<long long>(b) * plane + cell) * p.head_ld[l] + a * no;
      raw_off = ((static_cast<long long>(b) * p.na + a) * plane + cell) * no;
      lxy = (y << 16) | x;
      lla = (l << 8) | a;
    }
    const long long z_base = static_cast<long long>(gidx) * kDecodeRows * no;  // rows of a group are contiguous in z
    const int n_el = min(kDecodeRows, total - gidx * kDecodeRows) * no;
    // the group's kDecodeRows*no outputs form one contiguous range of z: lanes stride over it -> fully coalesced stores
#pragma unroll 1
    for (int e0 = 0; e0 < n_el; e0 += 128) {
      float v[4];
      int qq[4], kk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = e0 + j * 32 + lane;
        const int q = e / no;
        qq[j] = q;
        kk[j] = e - q * no;
        const long long so = __shfl_sync(0xffffffffu, src_off, q & 3);
        const int l = __shfl_sync(0xffffffffu, lla, q & 3) >> 8;
        v[j] = (e < n_el) ? __ldg(p.head[l] + so + kk[j]) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = e0 + j * 32 + lane;
        const int q = qq[j], k = kk[j];
        const long long ro = __shfl_sync(0xffffffffu, raw_off, q & 3);
        const int la = __shfl_sync(0xffffffffu, lla, q & 3);
        const int xy = __shfl_sync(0xffffffffu, lxy, q & 3);
        if (e >= n_el) continue;
        const int l = la >> 8, a = la & 255;
        const float x = v[j];
        if (p.raw[l]) p.raw[l][ro + k] = x;
        if (p.z) {
          const float s = sigmoid_fast(x);
          float o = s;
          if (k == 0)
            o = (s * 2.0f + (static_cast<float>(xy & 0xFFFF) - 0.5f)) * p.stride[l];
          else if (k == 1)
            o = (s * 2.0f + (static_cast<float>(xy >> 16) - 0.5f)) * p.stride[l];
          else if (k < 4) {
            const float t = s * 2.0f;
            o = (t * t) * (k == 2 ? p.anchor_w[l][a] : p.anchor_h[l][a]);
          }
          p.z[z_base + e] = o;
        }
      }
    }
  }
}

}  // namespace
}  // namespace y3

extern "C" int y3_detect_head_decode_fwd(const y3_decode_desc* d, y3_stream_t stream) {
  Y3_REQUIRE(d && d->nl >= 1 && d->nl <= Y3_MAX_LEVELS && d->na >= 1 && d->na <= Y3_MAX_ANCHORS && d->bs > 0 && d->no >= 5,
             "head_decode: bad arguments");
  y3::HeadDecodeArgs a{};
  a.nl = d->nl;
  a.bs = d->bs;
  a.na = d->na;
  a.no = d->no;
  a.z = d->z;
  int off = 0;
  for (int l = 0; l < d->nl; ++l) {
    const y3_detect_level& lv = d->levels[l];
    Y3_REQUIRE(lv.head && lv.ny > 0 && lv.nx > 0 && lv.head_ld >= d->na * d->no, "head_decode: bad level %d", l);
    a.head[l] = lv.head;
    a.head_ld[l] = lv.head_ld;
    a.raw[l] = lv.raw_out;
    a.ny[l] = lv.ny;
    a.nx[l] = lv.nx;
    a.stride[l] = lv.stride;
    a.row_off[l] = off;
    off += d->na * lv.ny * lv.nx;
    for (int j = 0; j < d->na; ++j) {
      a.anchor_w[l][j] = lv.anchor_w[j];
      a.anchor_h[l][j] = lv.anchor_h[j];
    }
  }
  a.row_off[d->nl] = off;
  Y3_REQUIRE(static_cast<long long>(off) * d->bs < (1ll << 31), "head_decode: too many rows");
  const long long warps = (static_cast<long long>(off) * d->bs + y3::kDecodeRows - 1) / y3::kDecodeRows;
  long long blocks = (warps + 7) / 8;
  const long long cap = static_cast<long long>(y3::num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  y3::head_decode_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  Y3_CHECK_CUDA(cudaGetLastError());
  return Y3_OK;
}

extern "C" int y3_detect_decode_fwd(const y3_detect_level* levels, int32_t nl, int32_t bs, int32_t na, int32_t no,
                                    float* z, y3_stream_t stream) {
  Y3_REQUIRE(levels && z && nl >= 1 && nl <= Y3_MAX_LEVELS && na >= 1 && na <= Y3_MAX_ANCHORS && bs > 0 && no >= 5,
             "decode: bad arguments (nl=%d na=%d bs=%d no=%d)", nl, na, bs, no);
  y3::DecodeArgs a{};
  a.nl = nl;
  a.bs = bs;
  a.na = na;
  a.no = no;
  a.z = z;
  int off = 0;
  for (int l = 0; l < nl; ++l) {
    Y3_REQUIRE(levels[l].raw && levels[l].ny > 0 && levels[l].nx > 0, "decode: bad level %d", l);
    a.raw[l] = levels[l].raw;
    a.ny[l] = levels[l].ny;
    a.nx[l] = levels[l].nx;
    a.stride[l] = levels[l].stride;
    a.row_off[l] = off;
    off += na * levels[l].ny * levels[l].nx;
    for (int j = 0; j < na; ++j) {
      a.anchor_w[l][j] = levels[l].anchor_w[j];
      a.anchor_h[l][j] = levels[l].anchor_h[j];
    }
  }
  a.row_off[nl] = off;
  const long long total = static_cast<long long>(off) * no * bs;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(y3::num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  y3::decode_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  Y3_CHECK_CUDA(cudaGetLastError());
  return Y3_OK;
}

// yolov3_b200 — test-time augmentation pieces: the reference's Model._forward_augment
// (models/yolo.py:239-280) runs the network on three views of the batch — scales 1 / 0.83 / 0.67, the middle one flipped
// left-right — built by scale_img (ultralytics: F.interpolate(bilinear, align_corners=False) to int(h*r) x int(w*r), padded
// right/bottom with 0.447 up to a stride multiple), then de-scales / de-flips the decoded rows, drops the P5 rows of the
// full-size view and the P3 rows of the smallest one (_clip_augmented) and concatenates.  Two kernels:
//   scale_img   fp32 NCHW -> fp32 NCHW view (optional left-right flip of the SOURCE, bilinear resample, constant pad)
//   tta_merge   rows [row_begin, row_end) of one view's z -> their place in the merged output, xywh /= scale, x = W - x when
//               flipped (the in-place arithmetic of _descale_pred; the clip and the concat cost no extra pass)
#include "y3_common.cuh"
#include "y3_internal.h"

namespace y3 {
namespace {

// PyTorch upsample_bilinear2d (align_corners = false): src = scale * (dst + 0.5) - 0.5 clamped at 0, scale = in / out (float)
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float s = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = static_cast<int>(s);
  i0 = i0 < in_size - 1 ? i0 : in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - static_cast<float>(i0);
  l0 = 1.0f - l1;
}

__global__ void __launch_bounds__(256) scale_img_kernel(const float* __restrict__ in, int planes, int h, int w, int rh, int rw,
                                                        int oh, int ow, int flip_lr, float pad, float* __restrict__ out) {
  const float sh = static_cast<float>(h) / static_cast<float>(rh), sw = static_cast<float>(w) / static_cast<float>(rw);
  const long long total = static_cast<long long>(planes) * oh * ow;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % ow);
    const long long t = i / ow;
    const int y = static_cast<int>(t % oh);
    const long long pl = t / oh;
    float v = pad;
    if (y < rh && x < rw) {
      int y0, y1, x0, x1;
      float ly0, ly1, lx0, lx1;
      src_index(y, sh, h, y0, y1, ly0, ly1);
      src_index(x, sw, w, x0, x1, lx0, lx1);
      if (flip_lr) {  // the resample reads x.flip(3)
        x0 = w - 1 - x0;
        x1 = w - 1 - x1;
      }
      const float* p = in + pl * h * w;
      v = ly0 * (lx0 * __ldg(p + y0 * w + x0) + lx1 * __ldg(p + y0 * w + x1)) +
          ly1 * (lx0 * __ldg(p + y1 * w + x0) + lx1 * __ldg(p + y1 * w + x1));
    }
    out[i] = v;
  }
}

__global__ void __launch_bounds__(256) tta_merge_kernel(const float* __restrict__ z, int bs, int rows, int no, int row_begin,
                                                        int row_end, float scale, int flip_lr, float img_w,
                                                        float* __restrict__ out, int out_rows, int out_row_off) {
  const int keep = row_end - row_begin;
  const long long total = static_cast<long long>(bs) * keep * no;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % no);
    const long long t = i / no;
    const int r = static_cast<int>(t % keep);
    const int b = static_cast<int>(t / keep);
    float v = z[(static_cast<long long>(b) * rows + row_begin + r) * no + c];
    if (c < 4) {
      v = __fdiv_rn(v, scale);                      // p[..., :4] /= scale
      if (c == 0 && flip_lr) v = __fsub_rn(img_w, v);  // p[..., 0] = img_size[1] - p[..., 0]
    }
    out[(static_cast<long long>(b) * out_rows + out_row_off + r) * no + c] = v;
  }
}

int blocks_for(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = 16ll * num_sms();
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace
}  // namespace y3

extern "C" int y3_scale_img_f32(const float* in, int32_t n, int32_t c, int32_t h, int32_t w, int32_t rh, int32_t rw, int32_t oh,
                                int32_t ow, int32_t flip_lr, float pad_value, float* out, y3_stream_t stream) {
  Y3_REQUIRE(in && out && n > 0 && c > 0 && h > 0 && w > 0 && rh > 0 && rw > 0 && oh >= rh && ow >= rw, "scale_img: bad arguments");
  const long long total = static_cast<long long>(n) * c * oh * ow;
  y3::scale_img_kernel<<<y3::blocks_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, n * c, h, w, rh, rw, oh, ow,
                                                                                             flip_lr, pad_value, out);
  Y3_CHECK_CUDA(cudaGetLastError());
  return Y3_OK;
}

extern "C" int y3_tta_merge(const float* z, int32_t bs, int32_t rows, int32_t no, int32_t row_begin, int32_t row_end, float scale,
                            int32_t flip_lr, float img_w, float* out, int32_t out_rows, int32_t out_row_off, y3_stream_t stream) {
  Y3_REQUIRE(z && out && bs > 0 && rows > 0 && no >= 5 && row_begin >= 0 && row_end <= rows && row_begin <= row_end && scale > 0.f,
             "tta_merge: bad arguments");
  Y3_REQUIRE(out_row_off >= 0 && out_row_off + (row_end - row_begin) <= out_rows, "tta_merge: rows do not fit the output");
  if (row_end == row_begin) return Y3_OK;
  const long long total = static_cast<long long>(bs) * (row_end - row_begin) * no;
  y3::tta_merge_kernel<<<y3::blocks_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      z, bs, rows, no, row_begin, row_end, scale, flip_lr, img_w, out, out_rows, out_row_off);
  Y3_CHECK_CUDA(cudaGetLastError());
  return Y3_OK;
}

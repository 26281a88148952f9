// yolov3_b200 — pairwise box IoU.  Replaces ultralytics box_iou as re-exported by the reference
// (utils/metrics.py:10; callers val.py:176 process_batch, utils/general.py:737): inter / (a1 + a2 - inter + eps),
// boxes xyxy, out [N, M].  Separately rounded fp32 operations in the reference's order (no FMA contraction).
#include "y3_common.cuh"
#include "y3_internal.h"

namespace y3 {
namespace {
__global__ void __launch_bounds__(256) box_iou_kernel(const float4* __restrict__ b1, int n, const float4* __restrict__ b2,
                                                      int m, float eps, float* __restrict__ out) {
  const long long total = static_cast<long long>(n) * m;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / m), c = static_cast<int>(i - static_cast<long long>(r) * m);
    const float4 a = __ldg(b1 + r), b = __ldg(b2 + c);
    const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
    const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
    const float inter = __fmul_rn(w, h);
    const float a1 = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float a2 = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    out[i] = __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(a1, a2), inter), eps));
  }
}

// scale_boxes / clip_boxes (reference utils/general.py:613-626 + ultralytics clip_boxes): in place on columns 0..3 of
// every row: x = clamp((x - pad_x) / gain, 0, max_x), y likewise; separately rounded fp32 ops in the reference's order
// (sub, div, clamp); torch.clamp propagates NaN.
__device__ __forceinline__ float clamp_like_torch(float v, float hi) { return v != v ? v : fminf(fmaxf(v, 0.0f), hi); }

__global__ void __launch_bounds__(256) scale_boxes_kernel(float* __restrict__ boxes, long long n, int row_stride, float pad_x,
                                                          float pad_y, float gain, float max_x, float max_y) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float* b = boxes + i * row_stride;
    b[0] = clamp_like_torch(__fdiv_rn(__fsub_rn(b[0], pad_x), gain), max_x);
    b[1] = clamp_like_torch(__fdiv_rn(__fsub_rn(b[1], pad_y), gain), max_y);
    b[2] = clamp_like_torch(__fdiv_rn(__fsub_rn(b[2], pad_x), gain), max_x);
    b[3] = clamp_like_torch(__fdiv_rn(__fsub_rn(b[3], pad_y), gain), max_y);
  }
}
}  // namespace
}  // namespace y3

extern "C" int y3_scale_boxes(float* boxes, int64_t n, int32_t row_stride, float pad_x, float pad_y, float gain, float max_x,
                              float max_y, y3_stream_t stream) {
  Y3_REQUIRE(n >= 0 && row_stride >= 4, "scale_boxes: bad shape (n=%lld, row stride %d)", static_cast<long long>(n), row_stride);
  if (n == 0) return Y3_OK;
  Y3_REQUIRE(boxes != nullptr, "scale_boxes: null pointer");
  long long blocks = (n + 255) / 256;
  const long long cap = static_cast<long long>(y3::num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  y3::scale_boxes_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      boxes, n, row_stride, pad_x, pad_y, gain, max_x, max_y);
  Y3_CHECK_CUDA(cudaGetLastError());
  return Y3_OK;
}

extern "C" int y3_box_iou(const float* box1, int32_t n, const float* box2, int32_t m, float eps, float* out,
                          y3_stream_t stream) {
  Y3_REQUIRE(n >= 0 && m >= 0, "box_iou: negative size");
  if (n == 0 || m == 0) return Y3_OK;
  Y3_REQUIRE(box1 && box2 && out, "box_iou: null pointer");
  Y3_REQUIRE((reinterpret_cast<uintptr_t>(box1) & 15) == 0 && (reinterpret_cast<uintptr_t>(box2) & 15) == 0,
             "box_iou: boxes must be 16-byte aligned [k,4] fp32");
  const long long total = static_cast<long long>(n) * m;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(y3::num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  y3::box_iou_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(box1), n, reinterpret_cast<const float4*>(box2), m, eps, out);
  Y3_CHECK_CUDA(cudaGetLastError());
  return Y3_OK;
}

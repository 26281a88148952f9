// yolov3_b200 — layer-0 convolution (c_in = 3): 3x3 stride-1 pad-1 on the fp32 NCHW image with folded BN + SiLU,
// fused with the NCHW-fp32 -> padded-NHWC-bf16 conversion.  K = 27 is too thin for an MMA tile, and the layer is
// bandwidth-bound (24.7 FLOP/B): one thread per output pixel, fp32 FMAs against smem-broadcast weights,
// one contiguous 2*c_out-byte store per pixel.  Also the test-only layout converters.
// Replaces Conv.forward_fuse for model.0 (reference models/common.py:77-81, models/yolov3.yaml:18).
#include "y3_common.cuh"
#include "y3_internal.h"

namespace y3 {
namespace {

template <typename TIN>
__device__ __forceinline__ float load_px(const TIN* p, float div);
template <>
__device__ __forceinline__ float load_px<float>(const float* p, float div) {
  const float v = __ldg(p);
  return div > 0.f ? __fdiv_rn(v, div) : v;
}
template <>
__device__ __forceinline__ float load_px<uint8_t>(const uint8_t* p, float div) {
  const float v = static_cast<float>(__ldg(p));
  return div > 0.f ? __fdiv_rn(v, div) : v;
}

template <int COUT, typename TIN>
__global__ void __launch_bounds__(128) conv_first_kernel(const TIN* __restrict__ in, float in_div, int H, int W,
                                                         const float* __restrict__ wgt, const float* __restrict__ bias,
                                                         __nv_bfloat16* __restrict__ out, int out_ld, int out_coff) {
  __shared__ __align__(16) float sw[27 * COUT];
  __shared__ __align__(16) float sb[COUT];
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) sw[i] = wgt[i];
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) sb[i] = bias[i];
  __syncthreads();
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int h = blockIdx.y, n = blockIdx.z;
  if (w >= W) return;

  float x[27];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const TIN* plane = in + (static_cast<size_t>(n) * 3 + c) * H * W;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hh = h + kh - 1;
      const bool row_ok = hh >= 0 && hh < H;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ww = w + kw - 1;
        x[(c * 3 + kh) * 3 + kw] = (row_ok && ww >= 0 && ww < W) ? load_px<TIN>(plane + static_cast<size_t>(hh) * W + ww, in_div) : 0.f;
      }
    }
  }
  float acc[COUT];
#pragma unroll
  for (int co = 0; co < COUT; ++co) acc[co] = sb[co];
#pragma unroll
  for (int k = 0; k < 27; ++k) {
#pragma unroll
    for (int co = 0; co < COUT; co += 4) {
      const float4 wv = *reinterpret_cast<const float4*>(&sw[k * COUT + co]);
      acc[co + 0] = fmaf(x[k], wv.x, acc[co + 0]);
      acc[co + 1] = fmaf(x[k], wv.y, acc[co + 1]);
      acc[co + 2] = fmaf(x[k], wv.z, acc[co + 2]);
      acc[co + 3] = fmaf(x[k], wv.w, acc[co + 3]);
    }
  }
  const size_t row = (static_cast<size_t>(n) * (H + 2) + h + 1) * (W + 2) + w + 1;
  uint4* dst = reinterpret_cast<uint4*>(out + row * out_ld + out_coff);
#pragma unroll
  for (int q = 0; q < COUT / 8; ++q) {
    uint4 o;
    o.x = pack_bf16x2(silu_f(acc[q * 8 + 0]), silu_f(acc[q * 8 + 1]));
    o.y = pack_bf16x2(silu_f(acc[q * 8 + 2]), silu_f(acc[q * 8 + 3]));
    o.z = pack_bf16x2(silu_f(acc[q * 8 + 4]), silu_f(acc[q * 8 + 5]));
    o.w = pack_bf16x2(silu_f(acc[q * 8 + 6]), silu_f(acc[q * 8 + 7]));
    dst[q] = o;
  }
}

__global__ void nchw_to_padded_kernel(const float* __restrict__ src, int C, int H, int W, __nv_bfloat16* __restrict__ dst,
                                      int ld, int coff, size_t total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int w = i % W;
  const int h = (i / W) % H;
  const int c = (i / (static_cast<size_t>(W) * H)) % C;
  const int n = i / (static_cast<size_t>(W) * H * C);
  dst[((static_cast<size_t>(n) * (H + 2) + h + 1) * (W + 2) + w + 1) * ld + coff + c] = __float2bfloat16(src[i]);
}

__global__ void padded_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, int ld, int coff, int C, int H, int W,
                                      float* __restrict__ dst, size_t total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int w = i % W;
  const int h = (i / W) % H;
  const int c = (i / (static_cast<size_t>(W) * H)) % C;
  const int n = i / (static_cast<size_t>(W) * H * C);
  dst[i] = __bfloat162float(src[((static_cast<size_t>(n) * (H + 2) + h + 1) * (W + 2) + w + 1) * ld + coff + c]);
}

}  // namespace
}  // namespace y3

extern "C" int y3_conv_first_fwd(const y3_first_desc* d, y3_stream_t stream) {
  Y3_REQUIRE(d && d->in && d->weight && d->bias && d->out, "conv_first: null pointer");
  Y3_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0 && d->n <= 65535 && d->h <= 65535, "conv_first: bad shape");
  Y3_REQUIRE(d->out_ld % 8 == 0 && d->out_coff % 8 == 0 && d->out_coff + d->c_out <= d->out_ld &&
                 (reinterpret_cast<uintptr_t>(d->out) & 15) == 0,
             "conv_first: bad output slice");
  Y3_REQUIRE(d->c_out == 16 || d->c_out == 32, "conv_first: c_out=%d unsupported (16 or 32)", d->c_out);
  Y3_REQUIRE(d->in_dtype == Y3_IN_F32 || d->in_dtype == Y3_IN_U8, "conv_first: bad input dtype %d", d->in_dtype);
  const dim3 grid((d->w + 127) / 128, d->h, d->n), block(128);
  auto* o = static_cast<__nv_bfloat16*>(d->out);
  auto s = static_cast<cudaStream_t>(stream);
#define Y3_FIRST(CO, T)                                                                                            \
  y3::conv_first_kernel<CO, T><<<grid, block, 0, s>>>(static_cast<const T*>(d->in), d->in_div, d->h, d->w, d->weight, \
                                                      d->bias, o, d->out_ld, d->out_coff)
  if (d->c_out == 32) {
    if (d->in_dtype == Y3_IN_F32) Y3_FIRST(32, float); else Y3_FIRST(32, uint8_t);
  } else {
    if (d->in_dtype == Y3_IN_F32) Y3_FIRST(16, float); else Y3_FIRST(16, uint8_t);
  }
#undef Y3_FIRST
  Y3_CHECK_CUDA(cudaGetLastError());
  return Y3_OK;
}

extern "C" int y3_nchw_to_padded_nhwc(const float* src, int32_t n, int32_t c, int32_t h, int32_t w, void* dst,
                                      int32_t dst_ld, int32_t dst_coff, y3_stream_t stream) {
  Y3_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0 && dst_coff + c <= dst_ld, "nchw_to_padded: bad args");
  const size_t total = static_cast<size_t>(n) * c * h * w;
  y3::nchw_to_padded_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, c, h, w, static_cast<__nv_bfloat16*>(dst), dst_ld, dst_coff, total);
  Y3_CHECK_CUDA(cudaGetLastError());
  return Y3_OK;
}

extern "C" int y3_padded_nhwc_to_nchw(const void* src, int32_t src_ld, int32_t src_coff, int32_t n, int32_t c, int32_t h,
                                      int32_t w, float* dst, y3_stream_t stream) {
  Y3_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0 && src_coff + c <= src_ld, "padded_to_nchw: bad args");
  const size_t total = static_cast<size_t>(n) * c * h * w;
  y3::padded_to_nchw_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), src_ld, src_coff, c, h, w, dst, total);
  Y3_CHECK_CUDA(cudaGetLastError());
  return Y3_OK;
}

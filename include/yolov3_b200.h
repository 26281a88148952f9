/* yolov3_b200 — C ABI of the Blackwell-native (sm_100a) YOLOv3 detection hot path.
 *
 * The reference (ultralytics/yolov3, pure Python) has no FFI: its seams for this path are Python call signatures
 * (SURVEY.md §8b).  This header is the drop-in boundary a binding for those seams attaches to; each entry point cites
 * the reference interface it replaces.  Conventions:
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers owned by the caller (PyTorch);
 *   - every function returns Y3_OK (0) or a negative Y3_ERR_*; text via y3_last_error(); nothing throws;
 *   - nothing allocates device memory, nothing synchronises the stream (the only host sync is y3_model_create's
 *     one-off capability probe); all launches go to the caller's stream and are CUDA-graph capturable;
 *   - re-entrant: no global mutable state besides the last-error string (thread-local) and a mutex-guarded
 *     per-process driver-entry-point cache.  A y3_model is immutable after create.
 * Activation layout ("padded NHWC"): bf16 [n, h+2, w+2, ld] with a one-pixel all-zero halo; a tensor may be a channel
 * slice [coff, coff+c) of a wider buffer (zero-copy Concat, models/common.py:424-428).
 */
#ifndef YOLOV3_B200_H
#define YOLOV3_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define Y3_OK 0
#define Y3_ERR_BAD_ARG (-1)     /* shape/alignment/enum not supported by this path */
#define Y3_ERR_CUDA (-2)        /* a CUDA runtime/driver call failed (message in y3_last_error) */
#define Y3_ERR_UNSUPPORTED (-3) /* device is not sm_100 (no fallback path exists by design) */

#define Y3_ACT_NONE 0
#define Y3_ACT_SILU 1

typedef void* y3_stream_t; /* cudaStream_t */

int y3_version(void);
/* Copies the calling thread's last error text into buf (NUL-terminated); returns its length. */
int y3_last_error(char* buf, size_t n);
/* Y3_OK iff the current CUDA device is compute capability 10.x. */
int y3_device_check(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Conv + folded-BN + SiLU (+ residual add, + nearest-2x upsample, + concat-offset store, or Detect raw store).
 * Replaces Conv.forward_fuse (models/common.py:77-81) after BaseModel.fuse (models/yolo.py:163-172), the shortcut add
 * of Bottleneck.forward (common.py:163-165), nn.Upsample+Concat (models/yolov3.yaml:43-44,51-52) and Detect.m[i]
 * (models/yolo.py:96-98).  tcgen05 implicit GEMM; c_in % 16 == 0, c_out_pad = c_out rounded up to the tile N.
 * ksize 1|3 with stride 1, or ksize 3 with stride 2 (h, w even); pad = ksize/2.
 */
typedef struct y3_conv_desc {
  int32_t n, h, w;       /* batch, UNPADDED input height/width */
  int32_t c_in, c_out;   /* logical channels */
  int32_t ksize, stride;
  int32_t act;           /* Y3_ACT_* */
  const void* in;        /* padded NHWC bf16 [n, h+2, w+2, in_ld]; the conv reads channels [in_coff, in_coff+c_in) */
  int32_t in_ld, in_coff;
  const void* weight;    /* bf16 [c_out_pad, ksize*ksize*c_in], k index = (kh*ksize+kw)*c_in + c, BN folded */
  const float* bias;     /* fp32 [c_out_pad] */
  void* out;             /* padded NHWC bf16 [n, ho*u+2, wo*u+2, out_ld], u = 1+upsample; written at [out_coff, +c_out) */
  int32_t out_ld, out_coff;
  const void* res;       /* optional residual (NULL = none): padded NHWC bf16 with the conv-output geometry */
  int32_t res_ld, res_coff;
  int32_t upsample;      /* 1: replicate every output pixel 2x2 into `out` */
  float* raw;            /* Detect head (NULL = off): fp32 [n, na, ho, wo, no] logits, c_out == na*no; `out` unused */
  int32_t na, no;
  int32_t* err;          /* optional device int32 error word written by the in-kernel watchdog */
} y3_conv_desc;
int y3_conv_bn_act_fwd(const y3_conv_desc* d, y3_stream_t stream);
/* Tile N the kernel will use for c_out (weights/bias must be padded to a multiple of it). */
int y3_conv_cout_pad(int32_t c_out);

/* First layer: 3x3 stride-1 conv on the fp32 NCHW image (c_in = 3), folded BN + SiLU, writing padded NHWC bf16.
 * Replaces Conv.forward_fuse for layer 0 together with the NCHW->NHWC/bf16 conversion.  c_out in {16, 32}.
 * weight: fp32 [27, c_out] (k = (c*3+kh)*3+kw), bias fp32 [c_out]. */
int y3_conv_first_fwd(const float* in_nchw, int32_t n, int32_t h, int32_t w, const float* weight, const float* bias,
                      int32_t c_out, void* out, int32_t out_ld, int32_t out_coff, y3_stream_t stream);

/* Layout helpers (tests / feeding intermediate tensors): NCHW fp32 <-> padded NHWC bf16 channel slice. */
int y3_nchw_to_padded_nhwc(const float* src, int32_t n, int32_t c, int32_t h, int32_t w, void* dst, int32_t dst_ld,
                           int32_t dst_coff, y3_stream_t stream);
int y3_padded_nhwc_to_nchw(const void* src, int32_t src_ld, int32_t src_coff, int32_t n, int32_t c, int32_t h,
                           int32_t w, float* dst, y3_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* YOLOV3_B200_H */

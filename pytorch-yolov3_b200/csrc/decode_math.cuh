// decode_math.cuh — scalar pieces of the YOLO decode shared by decode.cu and the fused head epilogue
// of conv_umma.cu: every fp32 step keeps the reference's operation order (yolov3/darknet.py:79-108,
// yolov3/inference.py:342-353) with explicit round-to-nearest intrinsics, so no FMA contraction
// changes a rounding.
#pragma once
#include "common.cuh"

namespace y3 {

__device__ __forceinline__ float sigmoidf_ref(float v) {
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v)));
}

__device__ __forceinline__ int f2i_trunc(float v) {
  // numpy astype(int) truncates toward zero; values stay far inside int32 for any sane logit
  return __float2int_rz(v);
}

// One thresholded box -> candidate record (x1, y1, x2, y2 | prob, cls, box, 0): pixel scaling by the
// original image size in fp32, truncation, cxywh_to_tlbr (c -/+ wh // 2, wh >= 0).
__device__ __forceinline__ void make_cand(float tx, float ty, float tw, float th, float prob, int cls, int box,
                                          int row, int col, int g_h, int g_w, float anchor_w, float anchor_h,
                                          float train_w, float train_h, float oh, float ow, uint4& lo, uint4& hi) {
  const float x = __fdiv_rn(__fadd_rn(sigmoidf_ref(tx), (float)col), (float)g_w);
  const float y = __fdiv_rn(__fadd_rn(sigmoidf_ref(ty), (float)row), (float)g_h);
  const float w = __fdiv_rn(__fmul_rn(expf(tw), anchor_w), train_w);
  const float h = __fdiv_rn(__fmul_rn(expf(th), anchor_h), train_h);
  const int cx = f2i_trunc(__fmul_rn(x, ow));  // inference.py:351-353
  const int cy = f2i_trunc(__fmul_rn(y, oh));
  const int bw = f2i_trunc(__fmul_rn(w, ow));
  const int bh = f2i_trunc(__fmul_rn(h, oh));
  const int hw = bw >> 1, hh = bh >> 1;
  lo = make_uint4((uint32_t)(cx - hw), (uint32_t)(cy - hh), (uint32_t)(cx + hw), (uint32_t)(cy + hh));
  hi = make_uint4(__float_as_uint(prob), (uint32_t)cls, (uint32_t)box, 0u);
}

}  // namespace y3

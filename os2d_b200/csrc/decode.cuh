// Box decoding arithmetic shared by the stand-alone decode kernel (postproc.cu) and the fused per-label
// decode + NMS kernel (detect.cu), so that both produce bit-identical boxes:
//   torchvision BoxCoder.decode_single (weights 10,10,5,5; dw,dh clamped at ln(1000/16)), clip_boxes_to_image,
//   empty test, BoxList.resize to the original image        os2d/modeling/box_coder.py:319-330, 490-520,
//   anchors                                                 os2d/modeling/box_coder.py:42-59
// All float arithmetic uses explicit round-to-nearest intrinsics so that no FMA contraction changes a decision
// relative to the scalar CPU order of operations.
#pragma once
#include <cuda_runtime.h>

namespace os2d {


struct Decoded {
  float4 box;       // clipped box, rescaled to the original image
  float4 anchor;    // anchor, rescaled
  bool empty;       // degenerate after clipping (before rescaling)
};

// anchor n of a feature map of width fm_w; loc values l0..l3; clip window img_w x img_h; rescale sx / sy
template <typename AnchorGrid>   // any struct with stride_w, stride_h, box_w, box_h
__device__ __forceinline__ Decoded decode_one(const AnchorGrid& G, int n, int fm_w, float l0, float l1, float l2, float l3,
                                              float img_w, float img_h, float sx, float sy) {
  const int y = n / fm_w, x = n - y * fm_w;
  // anchor (box_coder.py:42-59): centre (i + 0.5) * stride, xyxy = c -/+ size / 2
  const float cx = __fmul_rn(static_cast<float>(x) + 0.5f, G.stride_w), cy = __fmul_rn(static_cast<float>(y) + 0.5f, G.stride_h);
  const float hw = G.box_w / 2.0f, hh = G.box_h / 2.0f;
  const float ax1 = __fsub_rn(cx, hw), ay1 = __fsub_rn(cy, hh), ax2 = __fadd_rn(cx, hw), ay2 = __fadd_rn(cy, hh);
  const float aw = __fsub_rn(ax2, ax1), ah = __fsub_rn(ay2, ay1);
  const float actr_x = __fadd_rn(ax1, __fmul_rn(0.5f, aw)), actr_y = __fadd_rn(ay1, __fmul_rn(0.5f, ah));
  const float dx = __fdiv_rn(l0, 10.0f), dy = __fdiv_rn(l1, 10.0f);
  const float kClip = 4.135166556742356f;   // log(1000 / 16)
  const float dw = fminf(__fdiv_rn(l2, 5.0f), kClip);
  const float dh = fminf(__fdiv_rn(l3, 5.0f), kClip);
  const float pcx = __fadd_rn(__fmul_rn(dx, aw), actr_x), pcy = __fadd_rn(__fmul_rn(dy, ah), actr_y);
  const float pw = __fmul_rn(expf(dw), aw), ph = __fmul_rn(expf(dh), ah);
  const float cw = __fmul_rn(0.5f, pw), chh = __fmul_rn(0.5f, ph);
  float x1 = __fsub_rn(pcx, cw), y1 = __fsub_rn(pcy, chh), x2 = __fadd_rn(pcx, cw), y2 = __fadd_rn(pcy, chh);
  x1 = fminf(fmaxf(x1, 0.f), img_w); x2 = fminf(fmaxf(x2, 0.f), img_w);
  y1 = fminf(fmaxf(y1, 0.f), img_h); y2 = fminf(fmaxf(y2, 0.f), img_h);
  Decoded d;
  d.empty = (y2 <= y1) || (x2 <= x1);
  d.box = make_float4(__fmul_rn(x1, sx), __fmul_rn(y1, sy), __fmul_rn(x2, sx), __fmul_rn(y2, sy));
  d.anchor = make_float4(__fmul_rn(ax1, sx), __fmul_rn(ay1, sy), __fmul_rn(ax2, sx), __fmul_rn(ay2, sy));
  return d;
}

// torchvision greedy-NMS overlap test (suppress IoU > thr, areas without +1, comparison in double as in the CPU kernel)
__device__ __forceinline__ bool iou_exceeds(const float4& bi, float iarea, const float4& bj, double thr) {
  const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y), xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
  const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  const float jarea = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter));
  return static_cast<double>(ovr) > thr;
}
// Same decision with an early exit for boxes that do not intersect (the common case among NMS candidates): then
// inter = +0, so ovr is 0 / (iarea + jarea) = +0 - or NaN when both areas are 0 - and `ovr > thr` is false for every
// thr >= 0.  `nonneg_thr` (thr >= 0) is hoisted by the caller; with a negative threshold the full test always runs.
__device__ __forceinline__ bool iou_exceeds_fast(const float4& bi, float iarea, const float4& bj, float jarea, double thr,
                                                 bool nonneg_thr) {
  const float xx1 = fmaxf(bi.x, bj.x), xx2 = fminf(bi.z, bj.z);
  const float w = __fsub_rn(xx2, xx1);
  const float yy1 = fmaxf(bi.y, bj.y), yy2 = fminf(bi.w, bj.w);
  const float h = __fsub_rn(yy2, yy1);
  if (nonneg_thr && (!(w > 0.f) || !(h > 0.f))) return false;
  const float inter = __fmul_rn(fmaxf(0.f, w), fmaxf(0.f, h));
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter));
  return static_cast<double>(ovr) > thr;
}
__device__ __forceinline__ float box_area(const float4& b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }

}  // namespace os2d

// Detection evaluation, device side: matching of detections to ground-truth boxes,
// os2d/data/voc_eval.py:109-126 - for every detection the ground-truth box OF THE SAME IMAGE AND LABEL with the largest
// IoU (first one among equals, numpy argmax), or -1 when that IoU is below the threshold.  IoU as boxlist_iou /
// torchvision.ops.box_iou in fp32 after +1 on (x2, y2) of both boxes ("integer typed bounding boxes").  The greedy
// true/false-positive flags, precision / recall and AP follow from these indices with sorts and scans (os2d_b200/voc_eval.py).
#include "common.cuh"
#include "kernels.h"

namespace os2d {

// thread per detection; the ground truth of one image is a handful of boxes (gt_offsets[img] .. gt_offsets[img + 1])
__global__ void __launch_bounds__(256) voc_match_kernel(const float4* __restrict__ det_boxes, const int* __restrict__ det_img,
                                                         const int* __restrict__ det_label, const float4* __restrict__ gt_boxes,
                                                         const int* __restrict__ gt_label, const int* __restrict__ gt_offsets,
                                                         int n_det, float thr, int* __restrict__ gt_index) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_det) return;
  const float4 b = det_boxes[i];
  const float bx2 = __fadd_rn(b.z, 1.0f), by2 = __fadd_rn(b.w, 1.0f);
  const float area_p = __fmul_rn(__fsub_rn(bx2, b.x), __fsub_rn(by2, b.y));
  const int img = det_img[i], label = det_label[i];
  float best = -1.0f;
  int idx = -1;
  for (int j = gt_offsets[img]; j < gt_offsets[img + 1]; ++j) {
    if (gt_label[j] != label) continue;
    const float4 g = gt_boxes[j];
    const float gx2 = __fadd_rn(g.z, 1.0f), gy2 = __fadd_rn(g.w, 1.0f);
    const float area_g = __fmul_rn(__fsub_rn(gx2, g.x), __fsub_rn(gy2, g.y));
    const float w = fmaxf(__fsub_rn(fminf(bx2, gx2), fmaxf(b.x, g.x)), 0.0f);
    const float h = fmaxf(__fsub_rn(fminf(by2, gy2), fmaxf(b.y, g.y)), 0.0f);
    const float inter = __fmul_rn(w, h);
    const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_p, area_g), inter));
    if (iou > best) { best = iou; idx = j; }
  }
  if (idx >= 0 && best < thr) idx = -1;
  gt_index[i] = idx;
}

int launch_voc_match(const float* det_boxes, const int* det_img, const int* det_label, const float* gt_boxes,
                     const int* gt_label, const int* gt_offsets, int n_det, float iou_thr, int* gt_index, cudaStream_t st) {
  if (n_det < 0) return kErrBadArg;
  if (n_det == 0) return kOk;
  voc_match_kernel<<<(n_det + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float4*>(det_boxes), det_img, det_label,
                                                       reinterpret_cast<const float4*>(gt_boxes), gt_label, gt_offsets, n_det,
                                                       iou_thr, gt_index);
  OS2D_AFTER_LAUNCH();
  return kOk;
}

}  // namespace os2d

// K4 / K5: box decode + filtering for all classes of a pyramid level in one launch, and batched greedy NMS
// (one CTA per segment, boxes resident in shared memory).
//   decode: torchvision BoxCoder.decode_single (weights 10,10,5,5; dw,dh clamped at ln(1000/16)), clip_boxes_to_image,
//           empty / score filter, BoxList.resize to the original image      os2d/modeling/box_coder.py:319-330, 490-520
//   nms:    torchvision greedy NMS semantics (suppress IoU > thr, areas without +1, comparison in double as in the
//           CPU kernel), boxes visited in the caller-provided score order  os2d/structures/bounding_box.py:344-387
// All float arithmetic uses explicit round-to-nearest intrinsics so that no FMA contraction changes a decision
// relative to the scalar CPU order of operations.
#include "common.cuh"
#include "decode.cuh"
#include "kernels.h"

namespace os2d {

__global__ void __launch_bounds__(256) decode_kernel(DecodeArgs A, const float* __restrict__ loc,
                                                      const float* __restrict__ score, const float* __restrict__ corners,
                                                      float4* __restrict__ boxes, float4* __restrict__ anchors_out,
                                                      float* __restrict__ corners_out, uint8_t* __restrict__ valid) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (n >= A.N) return;
  const DecodeArgs& G = A;
  const float* l = loc + static_cast<size_t>(c) * 4 * A.N + n;
  const float sx = A.scale_x, sy = A.same_scale ? A.scale_x : A.scale_y;
  const Decoded d = decode_one(G, n, A.fm_w, l[0], l[A.N], l[2 * static_cast<size_t>(A.N)], l[3 * static_cast<size_t>(A.N)],
                               A.img_w, A.img_h, sx, sy);
  const float s = score[static_cast<size_t>(c) * A.N + n];
  valid[static_cast<size_t>(c) * A.N + n] = (s > A.score_thr) && !d.empty;
  boxes[static_cast<size_t>(c) * A.N + n] = d.box;
  if (c == 0) anchors_out[n] = d.anchor;
  if (corners != nullptr) {
    const float* co = corners + static_cast<size_t>(c) * 8 * A.N + n;
    float* dst = corners_out + (static_cast<size_t>(c) * A.N + n) * 8;
#pragma unroll
    for (int q = 0; q < 8; ++q) dst[q] = __fmul_rn(co[static_cast<size_t>(q) * A.N], (q & 1) ? sy : sx);
  }
}

int launch_decode(const DecodeArgs& A, const float* loc, const float* score, const float* corners, float* boxes,
                  float* anchors_out, float* corners_out, uint8_t* valid, cudaStream_t st) {
  if (A.C <= 0 || A.N <= 0 || A.fm_w <= 0) return kErrBadArg;
  decode_kernel<<<dim3((A.N + 255) / 256, A.C), 256, 0, st>>>(A, loc, score, corners, reinterpret_cast<float4*>(boxes),
                                                            reinterpret_cast<float4*>(anchors_out), corners_out, valid);
  OS2D_AFTER_LAUNCH();
  return kOk;
}

constexpr int kNmsMaxSeg = 10000;   // nms_max_batch of the reference (bounding_box.py:344)
constexpr int kNmsThreads = 1024;

__global__ void __launch_bounds__(kNmsThreads, 1) nms_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ order,
                                                              const int32_t* __restrict__ seg_off, double thr,
                                                              uint8_t* __restrict__ keep) {
  extern __shared__ __align__(16) uint8_t nms_smem[];
  const int seg = blockIdx.x;
  const int beg = seg_off[seg], n = seg_off[seg + 1] - beg;
  if (n <= 0) return;
  float4* sb = reinterpret_cast<float4*>(nms_smem);
  uint8_t* supp = nms_smem + static_cast<size_t>(kNmsMaxSeg) * sizeof(float4);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    sb[i] = boxes[order[beg + i]];
    supp[i] = 0;
  }
  __syncthreads();
  __shared__ int s_next;
  int start = 0;
  while (true) {
    // warp 0 finds the next box that is still alive (32 flags per step), everybody else waits at the barrier
    if (threadIdx.x < 32) {
      int found = -1;
      for (int j = start; j < n; j += 32) {
        const bool alive = (j + static_cast<int>(threadIdx.x) < n) && !supp[j + threadIdx.x];
        const unsigned m = __ballot_sync(0xffffffffu, alive);
        if (m) { found = j + __ffs(m) - 1; break; }
      }
      if (threadIdx.x == 0) s_next = found;
    }
    __syncthreads();
    const int i = s_next;
    if (i < 0) break;
    start = i + 1;
    const float4 bi = sb[i];
    const float iarea = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
    for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
      if (supp[j]) continue;
      if (iou_exceeds(bi, iarea, sb[j], thr)) supp[j] = 1;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) keep[beg + i] = supp[i] ? 0 : 1;
}

int launch_nms(const float* boxes, const int32_t* order, const int32_t* seg_offsets, int num_segs, double iou_thr,
               uint8_t* keep, cudaStream_t st) {
  if (num_segs <= 0) return kOk;
  const size_t smem = static_cast<size_t>(kNmsMaxSeg) * (sizeof(float4) + 1);
  OS2D_SET_MAX_DYN_SMEM(nms_kernel, smem);
  nms_kernel<<<num_segs, kNmsThreads, smem, st>>>(reinterpret_cast<const float4*>(boxes), order, seg_offsets, iou_thr, keep);
  OS2D_AFTER_LAUNCH();
  return kOk;
}

}  // namespace os2d

// extern "C" shim of include/os2d_b200.h + host utilities (error string, tensor-map encoder).
#include <cudaTypedefs.h>

#include <atomic>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/os2d_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace os2d {

static thread_local char g_err[512] = "";

void set_last_error(const char* what, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}
void set_last_error_msg(const char* what) { snprintf(g_err, sizeof(g_err), "%s", what); }

bool pdl_enabled() {
  static const bool on = getenv("OS2D_B200_NO_PDL") == nullptr;
  return on;
}

static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
    set_last_error_msg("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed");
    return nullptr;
  }
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  return fn;
}

int encode_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                      const uint32_t* box, CUtensorMapSwizzle swizzle, CUtensorMapL2promotion promo) {
  auto fn = get_encode_fn();
  if (!fn) return kErrDriver;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr,
                  bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu, box %u %u %u)",
             static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
    return kErrDriver;
  }
  return kOk;
}

static int num_sms_cached() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (dev < 0 || dev >= 64) return -1;
  if (cached[dev] == 0) {
    int n = 0, major = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
    if (major != 10) {
      set_last_error_msg("os2d_b200 requires a compute capability 10.x device (sm_100a kernels)");
      return -1;
    }
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace os2d

using namespace os2d;

extern "C" {

int os2d_b200_abi_version(void) { return 1; }
const char* os2d_b200_last_error(void) { return os2d::g_err; }
int os2d_b200_num_sms(void) { return num_sms_cached(); }
unsigned long long os2d_b200_launch_count(void) { return os2d::g_launches.load(std::memory_order_relaxed); }

int os2d_pack_class_features(const float* maps, int C, int D, int h, int w, int normalize, float* cf32, void* packed,
                             void* stream) {
  if (!maps || !cf32 || !packed) return kErrBadArg;
  return launch_pack_class(maps, C, D, h, w, normalize, cf32, packed, static_cast<cudaStream_t>(stream));
}

int os2d_pack_class_features_ragged(const float* const* map_ptrs, const int* hw, int C, int D, int normalize, float* cf32,
                                    void* packed, void* stream) {
  if (!map_ptrs || !hw || !cf32 || !packed) return kErrBadArg;
  return launch_pack_class_ragged(map_ptrs, hw, C, D, normalize, cf32, packed, static_cast<cudaStream_t>(stream));
}

int os2d_pack_image_features(const float* fm, int B, int D, int N, float* inv_ws, void* packed, void* stream) {
  if (!fm || !inv_ws || !packed) return kErrBadArg;
  return launch_pack_image(fm, B, D, N, inv_ws, packed, static_cast<cudaStream_t>(stream));
}

int os2d_pack_image_features_nhwc(const void* a, const void* b, int is_half, int relu, long long rows, int D, void* packed,
                                  void* stream) {
  if (!a || !packed) return kErrBadArg;
  return launch_pack_image_nhwc(a, b, is_half, relu, rows, D, packed, static_cast<cudaStream_t>(stream));
}

int os2d_correlate(const void* img_packed, const void* cls_packed, int B, int C, int D, int H, int W, void* zvol,
                   void* rawvol, void* stream) {
  if (!img_packed || !cls_packed || !zvol || !rawvol) return kErrBadArg;
  const int sms = num_sms_cached();
  if (sms <= 0) return kErrUnsupported;
  return launch_corr(img_packed, cls_packed, B, C, D, H, W, zvol, rawvol, sms, static_cast<cudaStream_t>(stream));
}

// K1 and conv1 side by side: the correlation kernel runs on `corr_sms` SMs (stream `aux_stream`) and publishes every finished
// plane; conv1 runs on the remaining SMs (stream `stream`) and consumes the planes as they complete, so the z volume is read
// while it is still in L2 and K1 - off the critical path - is no longer bound by the chip-wide L2 -> SM delivery rate.
int os2d_correlate_conv1_concurrent(const void* img_packed, const void* cls_packed, int B, int C, int D, int H, int W, void* zvol,
                                    void* rawvol, const void* w1blob, const float* alpha1, const float* beta1, void* h1,
                                    unsigned int* plane_flags, int corr_sms, void* stream, void* aux_stream) {
  if (!img_packed || !cls_packed || !zvol || !rawvol || !w1blob || !alpha1 || !beta1 || !h1 || !plane_flags || !aux_stream)
    return kErrBadArg;
  const int sms = num_sms_cached();
  if (sms <= 0) return kErrUnsupported;
  if (corr_sms < 2 || (corr_sms & 1) || corr_sms > sms - 2) return kErrBadArg;
  cudaStream_t st = static_cast<cudaStream_t>(stream), aux = static_cast<cudaStream_t>(aux_stream);
  int dev = 0;
  OS2D_CUDA_TRY(cudaGetDevice(&dev));
  static cudaEvent_t ev_fork[kMaxDevices] = {}, ev_join[kMaxDevices] = {};
  if (!ev_fork[dev]) {
    OS2D_CUDA_TRY(cudaEventCreateWithFlags(&ev_fork[dev], cudaEventDisableTiming));
    OS2D_CUDA_TRY(cudaEventCreateWithFlags(&ev_join[dev], cudaEventDisableTiming));
  }
  const int planes = B * C;
  OS2D_CUDA_TRY(cudaMemsetAsync(plane_flags, 0, sizeof(unsigned int) * planes, st));
  OS2D_CUDA_TRY(cudaEventRecord(ev_fork[dev], st));
  OS2D_CUDA_TRY(cudaStreamWaitEvent(aux, ev_fork[dev], 0));
  unsigned int target = 0;
  int rc = launch_corr(img_packed, cls_packed, B, C, D, H, W, zvol, rawvol, corr_sms, aux, plane_flags, &target);
  if (rc != kOk) return rc;
  OS2D_CUDA_TRY(cudaEventRecord(ev_join[dev], aux));
  ConvLayerDesc L;
  L.lo_scale = 1.0f / 2048.0f; L.ksize = 7; L.in_chunks16 = kCorrPad / 16; L.out_real = 128; L.mode = 0;
  rc = launch_conv(L, zvol, w1blob, alpha1, beta1, h1, planes, H, W, sms - corr_sms, st, plane_flags, target);
  if (rc != kOk) return rc;
  OS2D_CUDA_TRY(cudaStreamWaitEvent(st, ev_join[dev], 0));     // raw volume (K3's input) and K1 itself are complete
  return kOk;
}

size_t os2d_conv_weight_blob_bytes(int ksize, int in_chunks16) { return conv_weight_blob_bytes(ksize, in_chunks16); }
size_t os2d_conv3_weight_blob_bytes(int P) { return conv3s_weight_blob_bytes(P); }

int os2d_transform_conv(int layer, int out_real, const void* in_vol, const void* wblob, const float* alpha,
                        const float* beta, void* out, int planes, int H, int W, void* stream) {
  if (!in_vol || !wblob || !alpha || !beta || !out) return kErrBadArg;
  const int sms = num_sms_cached();
  if (sms <= 0) return kErrUnsupported;
  ConvLayerDesc L;
  L.lo_scale = 1.0f / 2048.0f;
  if (layer == 1) { L.ksize = 7; L.in_chunks16 = kCorrPad / 16; L.out_real = 128; L.mode = 0; }
  else if (layer == 2) { L.ksize = 5; L.in_chunks16 = 8; L.out_real = 64; L.mode = 1; }
  else if (layer == 3) {
    // scatter-form kernel: wblob = conv3s blob, alpha[0] = 1 / weight scale, beta[0..P) = bias
    return launch_conv3s(in_vol, wblob, beta, alpha, reinterpret_cast<float*>(out), planes, out_real, H, W, sms,
                         static_cast<cudaStream_t>(stream));
  } else return kErrBadArg;
  return launch_conv(L, in_vol, wblob, alpha, beta, out, planes, H, W, sms, static_cast<cudaStream_t>(stream));
}

int os2d_resample_boxes(const void* rawvol, const float* params, int planes, int P, int H, int W, int inverse,
                        float stride_w, float stride_h, float box_w, float box_h, float* score, float* loc,
                        float* corners, long long score_plane_stride, long long loc_plane_stride,
                        long long corners_plane_stride, void* stream) {
  if (!rawvol || !params || !score || !loc || !corners) return kErrBadArg;
  return launch_resample(rawvol, params, planes, P, H, W, inverse, stride_w, stride_h, box_w, box_h, score, loc, corners,
                         score_plane_stride, loc_plane_stride, corners_plane_stride, static_cast<cudaStream_t>(stream));
}

int os2d_resample_boxes_p2p(const void* rawvol, const float* params, int planes, int P, int H, int W, int inverse,
                            float stride_w, float stride_h, float box_w, float box_h, const void* const* peer_bases, int n_peers,
                            long long score_off, long long loc_off, long long corners_off, long long plane_stride,
                            void* stream) {
  if (!rawvol || !params || !peer_bases) return kErrBadArg;
  return launch_resample_p2p(rawvol, params, planes, P, H, W, inverse, stride_w, stride_h, box_w, box_h,
                             reinterpret_cast<float* const*>(const_cast<void* const*>(peer_bases)), n_peers, score_off, loc_off,
                             corners_off, plane_stride, static_cast<cudaStream_t>(stream));
}

int os2d_pack_corr_maps(const float* corr, int planes, int H, int W, void* zvol, void* rawvol, void* stream) {
  if (!corr || !zvol || !rawvol) return kErrBadArg;
  return launch_pack_corr(corr, planes, H * W, zvol, rawvol, static_cast<cudaStream_t>(stream));
}

int os2d_affine_grids(const float* params, int planes, int P, int H, int W, int inverse, float* grids, void* stream) {
  if (!params || !grids) return kErrBadArg;
  return launch_affine_grids(params, planes, P, H * W, inverse, grids, static_cast<cudaStream_t>(stream));
}

int os2d_resample_with_grid(const float* corr, const float* grids, const float* mask, int planes, int C, int H, int W,
                            float* out, void* stream) {
  if (!corr || !grids || !mask || !out) return kErrBadArg;
  return launch_resample_grid(corr, grids, mask, planes, C, H, W, out, static_cast<cudaStream_t>(stream));
}

int os2d_decode_boxes(int C, int N, int fm_w, float stride_w, float stride_h, float box_w, float box_h, float img_w,
                      float img_h, float score_thr, float scale_x, float scale_y, int same_scale, const float* loc,
                      const float* score, const float* corners, float* boxes, float* anchors, float* corners_out,
                      uint8_t* valid, void* stream) {
  if (!loc || !score || !boxes || !anchors || !valid) return kErrBadArg;
  if (corners && !corners_out) return kErrBadArg;
  DecodeArgs A;
  A.C = C; A.N = N; A.fm_w = fm_w;
  A.stride_w = stride_w; A.stride_h = stride_h; A.box_w = box_w; A.box_h = box_h;
  A.img_w = img_w; A.img_h = img_h; A.score_thr = score_thr;
  A.scale_x = scale_x; A.scale_y = scale_y; A.same_scale = same_scale;
  return launch_decode(A, loc, score, corners, boxes, anchors, corners_out, valid, static_cast<cudaStream_t>(stream));
}

int os2d_nms_segments(const float* boxes, const int32_t* order, const int32_t* seg_offsets, int num_segs,
                      double iou_threshold, uint8_t* keep, void* stream) {
  if (num_segs < 0) return kErrBadArg;
  if (num_segs == 0) return kOk;
  if (!boxes || !order || !seg_offsets || !keep) return kErrBadArg;
  return launch_nms(boxes, order, seg_offsets, num_segs, iou_threshold, keep, static_cast<cudaStream_t>(stream));
}

static int fill_detect_args(DetectArgs& A, const os2d_pyramid_level* levels, int num_levels, int num_views, int n_labels,
                            float stride_w, float stride_h, float box_w, float box_h) {
  if (!levels || num_levels <= 0 || num_levels > kMaxPyramidLevels || num_views <= 0 || n_labels <= 0) return kErrBadArg;
  A.L = num_levels; A.C = num_views; A.n_labels = n_labels;
  long long sum = 0;
  for (int l = 0; l < num_levels; ++l) {
    const os2d_pyramid_level& s = levels[l];
    if (!s.loc || !s.score || s.num_anchors <= 0 || s.fm_w <= 0) return kErrBadArg;
    A.base[l] = static_cast<long long>(num_views) * sum;
    A.lv[l].loc = s.loc; A.lv[l].score = s.score; A.lv[l].corners = s.corners;
    A.lv[l].N = s.num_anchors; A.lv[l].fm_w = s.fm_w;
    A.lv[l].img_w = s.img_w; A.lv[l].img_h = s.img_h; A.lv[l].scale_x = s.scale_x; A.lv[l].scale_y = s.scale_y;
    A.lv[l].same_scale = s.same_scale;
    sum += s.num_anchors;
  }
  for (int l = num_levels; l <= kMaxPyramidLevels; ++l) A.base[l] = static_cast<long long>(num_views) * sum;
  A.sumN = sum;
  if (static_cast<long long>(num_views) * sum >= (1ll << 31)) {
    set_last_error_msg("os2d_detect_pyramid: more than 2^31 (class view, anchor) pairs in one call");
    return kErrUnsupported;
  }
  A.grid.stride_w = stride_w; A.grid.stride_h = stride_h; A.grid.box_w = box_w; A.grid.box_h = box_h;
  return kOk;
}

int os2d_detect_pyramid(const os2d_pyramid_level* levels, int num_levels, int num_views, const int32_t* view_offsets,
                        const int32_t* view_ids, int n_labels, int max_views_per_label, float stride_w, float stride_h,
                        float box_w, float box_h, float score_thr, double iou_threshold, int32_t* cand_ws, uint64_t* key_ws, int32_t* out_ids,
                        int32_t* counts, int32_t* offsets, uint32_t* done_counter, void* stream) {
  if (!view_offsets || !view_ids || !cand_ws || !out_ids || !counts || !offsets || !done_counter || max_views_per_label <= 0)
    return kErrBadArg;
  DetectArgs A;
  const int rc = fill_detect_args(A, levels, num_levels, num_views, n_labels, stride_w, stride_h, box_w, box_h);
  if (rc != kOk) return rc;
  if (!key_ws && static_cast<long long>(max_views_per_label) * A.sumN > 10000) {
    set_last_error_msg("os2d_detect_pyramid: key_ws is required when a label can hold more than one NMS chunk of candidates");
    return kErrBadArg;
  }
  A.score_thr = score_thr; A.iou_thr = iou_threshold;
  A.view_off = view_offsets; A.view_ids = view_ids;
  A.cand = cand_ws; A.keys = reinterpret_cast<unsigned long long*>(key_ws); A.out_ids = out_ids;
  A.counts = counts; A.offsets = offsets; A.done = done_counter;
  return launch_label_nms(A, static_cast<cudaStream_t>(stream));
}

int os2d_gather_detections(const os2d_pyramid_level* levels, int num_levels, int num_views, const int32_t* view_offsets,
                           int n_labels, float stride_w, float stride_h, float box_w, float box_h, const int32_t* out_ids,
                           const int32_t* counts, const int32_t* offsets, const int64_t* label_values, float* boxes,
                           float* scores, int64_t* labels, float* anchors, float* corners, void* stream) {
  if (!view_offsets || !out_ids || !counts || !offsets || !label_values || !boxes || !scores || !labels || !anchors)
    return kErrBadArg;
  DetectArgs A;
  const int rc = fill_detect_args(A, levels, num_levels, num_views, n_labels, stride_w, stride_h, box_w, box_h);
  if (rc != kOk) return rc;
  if (corners) for (int l = 0; l < num_levels; ++l) if (!levels[l].corners) return kErrBadArg;
  A.score_thr = 0.f; A.iou_thr = 0.0;
  A.view_off = view_offsets; A.view_ids = nullptr;
  A.cand = nullptr; A.keys = nullptr; A.out_ids = const_cast<int32_t*>(out_ids);
  A.counts = const_cast<int32_t*>(counts); A.offsets = const_cast<int32_t*>(offsets); A.done = nullptr;
  return launch_gather_detections(A, reinterpret_cast<const long long*>(label_values), boxes, scores,
                                  reinterpret_cast<long long*>(labels), anchors, corners, static_cast<cudaStream_t>(stream));
}

int os2d_voc_match(const float* det_boxes, const int* det_img, const int* det_label, const float* gt_boxes,
                   const int* gt_label, const int* gt_offsets, int n_det, float iou_thr, int* gt_index, void* stream) {
  if (n_det > 0 && (!det_boxes || !det_img || !det_label || !gt_offsets || !gt_index)) return kErrBadArg;
  return launch_voc_match(det_boxes, det_img, det_label, gt_boxes, gt_label, gt_offsets, n_det, iou_thr, gt_index,
                          static_cast<cudaStream_t>(stream));
}

int os2d_resize_level(const uint8_t* img_hwc, int H, int W, int out_h, int out_w, const int* xbounds, const int* xcoeffs,
                      int xksize, const int* ybounds, const int* ycoeffs, int yksize, const float* mean3, const float* std3,
                      uint8_t* tmp, float* out_chw, uint8_t* out_u8_hwc, void* stream) {
  if (!img_hwc || !xbounds || !xcoeffs || !ybounds || !ycoeffs || !mean3 || !std3 || !tmp || !out_chw) return kErrBadArg;
  return launch_resize_level(img_hwc, H, W, out_h, out_w, xbounds, xcoeffs, xksize, ybounds, ycoeffs, yksize, mean3, std3, tmp,
                             out_chw, out_u8_hwc, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

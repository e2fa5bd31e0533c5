/* os2d_b200 - C ABI of the B200-native OS2D dense correlation-and-alignment head.
 *
 * The reference (aosokin/os2d) has no FFI: the hot path sits behind Python classes
 * (os2d/modeling/head.py Os2dHead.forward :308-435, os2d/modeling/box_coder.py decode_pyramid :448-536).
 * This header is the boundary a binding for that path uses; os2d_b200/_cabi.py is the ctypes binding that
 * the drop-in Python classes (os2d_b200/head.py, box_coder.py) call.  Each entry point cites the reference
 * code whose arithmetic it replaces.
 *
 * Conventions: plain device pointers and sizes, no allocation inside (the caller owns outputs and
 * workspaces), stream-ordered on `stream` (a cudaStream_t passed as void*), re-entrant, return 0 on success
 * or a negative error code; os2d_b200_last_error() gives a thread-local message for the last failure.
 * All entry points require an sm_100 device (tcgen05 / TMA); there is no CPU or library fallback.
 */
#ifndef OS2D_B200_H_
#define OS2D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OS2D_B200_OK 0
#define OS2D_B200_ERR_BAD_ARG (-1)
#define OS2D_B200_ERR_CUDA (-2)
#define OS2D_B200_ERR_UNSUPPORTED (-3)
#define OS2D_B200_ERR_DRIVER (-4)

/* geometry constants of the packed layouts */
#define OS2D_B200_GRID 15          /* template grid (head.py:66-69) */
#define OS2D_B200_CORR_CH 225
#define OS2D_B200_CORR_PAD 240     /* rows of the packed class operand, channels of the z volume */

int os2d_b200_abi_version(void);
const char* os2d_b200_last_error(void);
/* number of SMs of the current device (persistent-grid size) */
int os2d_b200_num_sms(void);
/* number of kernels this library has launched in this process so far (every launcher counts; bench.py reports the delta over
 * its timed region as `gpu_launches`) */
unsigned long long os2d_b200_launch_count(void);

/* ---- K0: operand preparation -------------------------------------------------------------------------
 * Class side (head.py:241-259 resize to 15x15, :293 L2 norm, :342-344 transposed channel order):
 *   maps   [C, D, h, w] fp32 (all classes of one call share h, w)
 *   cf32   [C, D, 15, 15] fp32  - the normalised maps the reference keeps in Os2dHead.class_feature_maps
 *   packed [C, 240, D] fp16     - GEMM operand, row k = tx*15 + ty, rows 225..239 zero, values * 32
 *   normalize = 0 returns the resized maps without the L2 normalisation (head.py:241-259 alone). */
int os2d_pack_class_features(const float* maps, int C, int D, int h, int w, int normalize, float* cf32, void* packed,
                             void* stream);
/* Same for a set of differently sized class maps in ONE launch (the class branch of the
 * backbone runs once per image size, model.py:80-88; nothing is concatenated): map_ptrs = device array of C pointers to
 * [D, h_c, w_c] fp32 maps, hw = device array [C][2] of (h_c, w_c).  Bit-identical to the uniform entry point. */
int os2d_pack_class_features_ragged(const float* const* map_ptrs, const int* hw, int C, int D, int normalize, float* cf32,
                                    void* packed, void* stream);
/* Image side (head.py:339): fm [B, D, N] fp32 -> packed [B, N, D] fp16 (values * 32 / (norm + 1e-5)).
 * inv_ws: workspace of B*N floats. */
int os2d_pack_image_features(const float* fm, int B, int D, int N, float* inv_ws, void* packed, void* stream);
/* Channels-last producer side (os2d/modeling/feature_extractor.py:23-72 + head.py:339): the
 * backbone runs in NHWC, so its output rows [rows = B*H*W][D] are already in operand order.  x = a (+ b when b != NULL)
 * (ReLU when relu != 0: the residual add + ReLU that ends layer3's last bottleneck is fused here), then the L2
 * normalisation over D -> packed [rows, D] fp16 (x 32).  a / b: fp32 (is_half = 0) or fp16 (is_half = 1). */
int os2d_pack_image_features_nhwc(const void* a, const void* b, int is_half, int relu, long long rows, int D, void* packed,
                                  void* stream);

/* ---- K1: correlation + ReLU/L2-norm epilogue (head.py:342-350, :650) ---------------------------------
 *   zvol   [B*C, 30, H*W, 8] fp16  centred/normalised correlation + DC side channels (conv1 operand)
 *   rawvol [B*C, 225, H*W]  fp16  correlation values (sampler input) */
int os2d_correlate(const void* img_packed, const void* cls_packed, int B, int C, int D, int H, int W, void* zvol,
                   void* rawvol, void* stream);

/* K1 and the first TransformNet layer side by side (stage-concurrent form of os2d_correlate + os2d_transform_conv(1)):
 * the correlation kernel runs on `corr_sms` SMs (even, launched on aux_stream) and releases a per-plane completion count
 * after every tile; conv1 runs on the other SMs (launched on `stream`) and acquires the count of a plane before its first TMA
 * load of it, so it consumes the z volume while it is still in L2.  plane_flags: workspace of B*C uint32 (zeroed inside,
 * stream-ordered).  On return `stream` also waits for the correlation kernel (raw volume complete).  Same results bit for
 * bit as the two separate calls. */
int os2d_correlate_conv1_concurrent(const void* img_packed, const void* cls_packed, int B, int C, int D, int H, int W, void* zvol,
                                    void* rawvol, const void* w1blob, const float* alpha1, const float* beta1, void* h1,
                                    unsigned int* plane_flags, int corr_sms, void* stream, void* aux_stream);

/* ---- K2: TransformNet convolution layers (head.py:604-655) -------------------------------------------
 * layer 1: 225(+DC)->128 k7, BN+ReLU -> h1 [planes,16,H*W,8] fp16;
 * layer 2: 128->64 k5 (hi/lo weight rows), BN+ReLU -> h2 [planes,16,H*W,8] fp16 (chunks 0..7 value, 8..15 residual);
 * layer 3: 64->P k5 in scatter form (csrc/conv3s.cu) -> fp32 [planes, P, H*W]; alpha[0] = 1/weight scale, beta = bias.
 * wblob: weights packed by os2d_b200.head.pack_transform_net (shared-memory images, see csrc/conv.cu, conv3s.cu);
 * alpha/beta: 128 floats each (folded BN scale/shift incl. operand pre-scales). */
size_t os2d_conv_weight_blob_bytes(int ksize, int in_chunks16);
size_t os2d_conv3_weight_blob_bytes(int P);
int os2d_transform_conv(int layer, int out_real, const void* in_vol, const void* wblob, const float* alpha,
                        const float* beta, void* out, int planes, int H, int W, void* stream);

/* ---- K3: affine grid + bilinear resample/pool + box regression (head.py:81-193, 371-433, 439-520) -----
 *   params [planes, P, H*W] fp32 (P = 6 affine, 4 simplified), inverse: use_inverse_geom_model
 *   score [planes, H*W], loc [planes, 4, H*W], corners [planes, 8, H*W] with caller-given plane strides
 *   (in floats) so the outputs can live inside one gather buffer. */
int os2d_resample_boxes(const void* rawvol, const float* params, int planes, int P, int H, int W, int inverse,
                        float stride_w, float stride_h, float box_w, float box_h, float* score, float* loc,
                        float* corners, long long score_plane_stride, long long loc_plane_stride,
                        long long corners_plane_stride, void* stream);

/* K3 fused with the all-gather of the class-sharded multi-GPU path ("K3's epilogue writes directly
 * into the rank's slice of the gather buffer ... stretch: P2P"): same computation as os2d_resample_boxes, but the 13 outputs of
 * every location are stored into the gather buffer of EVERY rank.  peer_bases: DEVICE array of n_peers pointers to the
 * (peer-mapped, symmetric) fp32 gather buffers; *_off: float offsets of this call's first plane inside a buffer;
 * plane_stride: floats between consecutive planes (13 * H * W for the [G,B,C/G,13,N] layout).  A cross-rank barrier must
 * follow before the buffers are read (os2d_b200/dist.py).  Same kernel as os2d_resample_boxes, instantiated with a peer
 * sink (csrc/resample.cu); validated on 2 and 8 GPUs (tests/test_gpu_dist.py). */
int os2d_resample_boxes_p2p(const void* rawvol, const float* params, int planes, int P, int H, int W, int inverse,
                            float stride_w, float stride_h, float box_w, float box_h, const void* const* peer_bases, int n_peers,
                            long long score_off, long long loc_off, long long corners_off, long long plane_stride,
                            void* stream);

/* ---- secondary entry points: the public methods of the reference classes that take / return the large tensors ----
 * os2d_pack_corr_maps:     fp32 correlation maps [planes,225,H*W] -> zvol / rawvol (TransformationNet.forward input side,
 *                          head.py:648-650); then os2d_transform_conv 1..3 give the parameters.
 * os2d_affine_grids:       params [planes,P,H*W] -> grids [planes,H*W,15,15,2] fp32 in local coordinates
 *                          (Os2dAlignment.forward, head.py:155-193).
 * os2d_resample_with_grid: Os2dHead.resample_of_correlation_map_fast/_simple (head.py:439-594): corr fp32
 *                          [planes,225,H*W], grids [planes,H*W,15,15,2] in [-1,1] unit coordinates of the feature map,
 *                          mask [C,225] -> pooled [planes,H*W] (plane = image * C + class). */
int os2d_pack_corr_maps(const float* corr, int planes, int H, int W, void* zvol, void* rawvol, void* stream);
int os2d_affine_grids(const float* params, int planes, int P, int H, int W, int inverse, float* grids, void* stream);
int os2d_resample_with_grid(const float* corr, const float* grids, const float* mask, int planes, int C, int H, int W,
                            float* out, void* stream);

/* ---- K4: decode + filter, all classes of one pyramid level (box_coder.py:490-520) ---------------------
 *   loc [C,4,N], score [C,N], corners [C,8,N] or NULL -> boxes [C,N,4], anchors [N,4], corners_out [C,N,8],
 *   valid [C,N] (score > thr and non-empty after clipping); boxes/anchors/corners are rescaled to the
 *   original image (BoxList.resize, bounding_box.py:138-163). */
int os2d_decode_boxes(int C, int N, int fm_w, float stride_w, float stride_h, float box_w, float box_h, float img_w,
                      float img_h, float score_thr, float scale_x, float scale_y, int same_scale, const float* loc,
                      const float* score, const float* corners, float* boxes, float* anchors, float* corners_out,
                      uint8_t* valid, void* stream);

/* ---- K5: batched greedy NMS (bounding_box.py:344-387 per chunk, torchvision nms semantics) -------------
 *   boxes [M,4]; order [total] candidate indices, each segment sorted by score descending;
 *   seg_offsets [num_segs+1]; every segment <= 10000 boxes; keep [total] (1 = survives) in `order` positions. */
int os2d_nms_segments(const float* boxes, const int32_t* order, const int32_t* seg_offsets, int num_segs,
                      double iou_threshold, uint8_t* keep, void* stream);

/* ---- K4+K5 fused: decode_pyramid in two launches (box_coder.py:448-536, :424-437; bounding_box.py:344-387) -----------
 * os2d_detect_pyramid: ONE launch, one CTA per real label: decode (torchvision BoxCoder.decode_single, clip, empty / score
 *   filter, BoxList.resize to the original image), candidate list in the reference's concatenation order (class view,
 *   level, anchor), NMS in consecutive chunks of 10000 iterated to the fixpoint, final score-descending order.
 *   levels: HOST array of num_levels (<= 12) descriptors with DEVICE tensors; views of label i are
 *   view_ids[view_offsets[i] .. view_offsets[i+1]) (device int32 arrays; duplicated class ids merge, box_coder.py:483-488).
 *   Candidates / results are flat indices f = C * sum_{k<l} N_k + view * N_l + anchor.
 *   Workspaces (device): cand_ws, out_ids int32 [C * sum_l N_l]; key_ws uint64 [C * sum_l N_l], may be NULL when no label
 *   can exceed one chunk (max_views_per_label * sum_l N_l <= 10000); counts [n_labels]; offsets [n_labels + 1] (exclusive scan
 *   of counts, offsets[n_labels] = number of detections); done_counter: one zero-initialised uint32, reset by the kernel.
 * os2d_gather_detections: second launch, writes the detections label by label (label order = order of view_offsets),
 *   score-descending inside a label: boxes [T,4], scores [T], labels [T] int64 (= label_values[label]), anchors [T,4]
 *   ("default_boxes"), corners [T,8] (or NULL), T = offsets[n_labels] read back by the caller between the two launches. */
typedef struct os2d_pyramid_level {
  const float* loc;      /* [C,4,N] */
  const float* score;    /* [C,N] */
  const float* corners;  /* [C,8,N] or NULL */
  int num_anchors;       /* N = fm_h * fm_w */
  int fm_w;
  float img_w, img_h;    /* clip window: image size of this pyramid level */
  float scale_x, scale_y;/* level -> original image (BoxList.resize), 1 without inverse transforms */
  int same_scale;        /* ratio_w == ratio_h branch of BoxList.resize: single multiply by scale_x */
} os2d_pyramid_level;
int os2d_detect_pyramid(const os2d_pyramid_level* levels, int num_levels, int num_views, const int32_t* view_offsets,
                        const int32_t* view_ids, int n_labels, int max_views_per_label, float stride_w, float stride_h,
                        float box_w, float box_h, float score_thr, double iou_threshold, int32_t* cand_ws, uint64_t* key_ws, int32_t* out_ids,
                        int32_t* counts, int32_t* offsets, uint32_t* done_counter, void* stream);
int os2d_gather_detections(const os2d_pyramid_level* levels, int num_levels, int num_views, const int32_t* view_offsets,
                           int n_labels, float stride_w, float stride_h, float box_w, float box_h, const int32_t* out_ids,
                           const int32_t* counts, const int32_t* offsets, const int64_t* label_values, float* boxes,
                           float* scores, int64_t* labels, float* anchors, float* corners, void* stream);

/* ---- image pyramid level: PIL Image.resize(BILINEAR) + ToTensor + Normalize (transforms.py:72, dataloader.py:322-341) ----
 *   img_hwc [H,W,3] uint8 (device) -> out_chw [3,out_h,out_w] fp32 = (resized byte / 255 - mean) / std, bit-identical to
 *   Pillow's ImagingResample (horizontal pass, uint8 intermediate `tmp` [H,out_w,3], vertical pass; 22-bit fixed point).
 *   xbounds [out_w,2] / ybounds [out_h,2] = (first tap, tap count), xcoeffs [out_w,xksize] / ycoeffs [out_h,yksize] int32
 *   (device), computed like Pillow's precompute_coeffs (os2d_b200/pyramid.py); mean3 / std3: 3 HOST floats each;
 *   out_u8_hwc (optional, may be NULL): the resized bytes [out_h,out_w,3]. */
int os2d_resize_level(const uint8_t* img_hwc, int H, int W, int out_h, int out_w, const int* xbounds, const int* xcoeffs,
                      int xksize, const int* ybounds, const int* ycoeffs, int yksize, const float* mean3, const float* std3,
                      uint8_t* tmp, float* out_chw, uint8_t* out_u8_hwc, void* stream);

/* ---- detection evaluation: matching step of calc_detection_voc_prec_rec (os2d/data/voc_eval.py:109-126) ----
 *   det_boxes [n_det,4] xyxy fp32 (already resized to the ground-truth image size), det_img / det_label [n_det] int32;
 *   ground truth concatenated image by image: gt_boxes [n_gt,4], gt_label [n_gt], gt_offsets [n_images+1];
 *   gt_index [n_det]: index of the ground-truth box of the same image and label with the largest IoU (+1 on x2,y2, fp32,
 *   first among equals), -1 when there is none or that IoU < iou_thr. */
int os2d_voc_match(const float* det_boxes, const int* det_img, const int* det_label, const float* gt_boxes,
                   const int* gt_label, const int* gt_offsets, int n_det, float iou_thr, int* gt_index, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OS2D_B200_H_ */

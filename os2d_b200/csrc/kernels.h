// Internal launcher declarations + layout constants shared by the kernels and the C ABI shim.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace os2d {

// geometry of the OS2D head (os2d/modeling/head.py:66-69)
constexpr int kGrid = 15;                 // template grid 15 x 15
constexpr int kCorrCh = kGrid * kGrid;    // 225 correlation channels
constexpr int kCorrPad = 240;             // padded to a multiple of 16 (MMA N / K granularity)
constexpr int kZChunks = kCorrPad / 8;    // 30 chunks of 8 channels in the z volume
constexpr int kDcCh = 225;                // first of the 3 DC side channels (mean hi, mean hi, mean lo)

// power-of-two operand pre-scales (exact in fp16/fp32; undone in the epilogues)
constexpr float kScaleFeat = 32.0f;       // L2-normalised features -> fp16
constexpr float kScaleZ = 64.0f;          // centred normalised correlation (z - mean) -> fp16
constexpr float kScaleMean = 8.0f;        // per-pixel mean of z -> fp16 hi/lo pair

// ---- conv implicit-GEMM tile geometry (conv.cu) ----
constexpr int kStripW = 8;                // a strip is 8 pixels wide
constexpr int kStripsPerTile = 2;         // two adjacent strips share one halo box
constexpr int kMaxTileRows = 32;          // 8 * 32 = 256 = max MMA N
constexpr int kMinTileRows = 20;          // N >= 160 keeps the per-MMA shared-memory read rate under 128 B/clk

int launch_pack_class(const float* maps, int C, int D, int h, int w, int normalize, float* cf32, void* packed,
                      cudaStream_t st);
int launch_pack_class_ragged(const float* const* map_ptrs, const int* hw, int C, int D, int normalize, float* cf32,
                             void* packed, cudaStream_t st);
int launch_pack_image(const float* fm, int B, int D, int N, float* inv_ws, void* packed, cudaStream_t st);
int launch_pack_image_nhwc(const void* a, const void* b, int is_half, int relu, long long rows, int D, void* packed,
                           cudaStream_t st);

// correlation GEMM + ReLU/L2norm/centering epilogue
// plane_done (optional): per-plane completion counters for a concurrently running conv1 (see os2d_correlate_conv1_concurrent);
// signals_per_plane (optional, host): how many increments complete a plane
int launch_corr(const void* img_packed, const void* cls_packed, int B, int C, int D, int H, int W, void* zvol,
                void* rawvol, int num_sms, cudaStream_t st, unsigned int* plane_done = nullptr,
                unsigned int* signals_per_plane = nullptr);

// implicit-GEMM convolution layer of the TransformNet.  layer: 1, 2, 3
struct ConvLayerDesc {
  int ksize;          // 7 or 5
  int in_chunks16;    // input channels / 16 (15, 8, 4)
  int out_real;       // real output channels (128, 64, P)
  int mode;           // 0: relu(alpha*acc+beta) -> fp16 chunk8 volume (128 rows)
                      // 1: combine hi/lo rows (c, c+64), relu(alpha*acc+beta) -> fp16 chunk8 volume (64 ch)
                      //    as two planes sets: chunks 0..7 = fp16 value, chunks 8..15 = fp16 residual
  float lo_scale;     // factor applied to the lo-row accumulator before adding (2^-11) in modes 1/2
};
// wait_flags (optional): the producer of a tile polls wait_flags[plane] >= wait_target (acquire, gpu scope) before its first
// load of that plane - the input volume is being written by a kernel running concurrently on other SMs
int launch_conv(const ConvLayerDesc& L, const void* in_vol, const void* wblob, const float* alpha, const float* beta,
                void* out, int planes, int H, int W, int num_sms, cudaStream_t st, const unsigned int* wait_flags = nullptr,
                unsigned int wait_target = 0);

int launch_resample(const void* rawvol, const float* params, int planes, int P, int H, int W, int inverse,
                    float stride_w, float stride_h, float box_w, float box_h, float* score, float* loc,
                    float* corners, long long score_ps, long long loc_ps, long long corners_ps, cudaStream_t st);

// K3 with the all-gather fused in: outputs stored into every rank's gather buffer through peer pointers (resample_p2p.cu)
int launch_resample_p2p(const void* rawvol, const float* params, int planes, int P, int H, int W, int inverse,
                        float stride_w, float stride_h, float box_w, float box_h, float* const* peers, int n_peers,
                        long long score_off, long long loc_off, long long corners_off, long long plane_stride, cudaStream_t st);

// secondary entry points (aux.cu): stand-alone TransformationNet / Os2dAlignment / resample_of_correlation_map_* methods
int launch_pack_corr(const float* corr, int planes, int N, void* zvol, void* rawvol, cudaStream_t st);
int launch_affine_grids(const float* params, int planes, int P, int N, int inverse, float* grid, cudaStream_t st);
int launch_resample_grid(const float* corr, const float* grid, const float* mask, int planes, int C, int H, int W, float* out,
                         cudaStream_t st);

// post-processing
struct DecodeArgs {
  int C, N, fm_w;                 // classes, anchors (= fm_h * fm_w), feature-map width
  float stride_w, stride_h, box_w, box_h;   // anchor grid
  float img_w, img_h;             // clip window (pyramid-level image size)
  float score_thr;                // keep score > thr (strict)
  float scale_x, scale_y;         // level -> original image (BoxList.resize), 1 when no inverse transform
  int same_scale;                 // ratio_w == ratio_h branch of BoxList.resize (single multiply)
};
int launch_decode(const DecodeArgs& A, const float* loc, const float* score, const float* corners, float* boxes,
                  float* anchors_out, float* corners_out, uint8_t* valid, cudaStream_t st);
int launch_nms(const float* boxes, const int32_t* order, const int32_t* seg_offsets, int num_segs, double iou_thr,
               uint8_t* keep, cudaStream_t st);

// fused decode + per-label NMS over a pyramid (detect.cu)
constexpr int kMaxPyramidLevels = 12;
struct DetectLevel {
  const float* loc;       // [C,4,N]
  const float* score;     // [C,N]
  const float* corners;   // [C,8,N] or nullptr
  int N, fm_w;
  float img_w, img_h, scale_x, scale_y;
  int same_scale;
};
struct DetectArgs {
  int L, C, n_labels;
  long long sumN;                          // anchors of all levels
  long long base[kMaxPyramidLevels + 1];   // flat index base of level l = C * sum_{k<l} N_k
  DetectLevel lv[kMaxPyramidLevels];
  struct { float stride_w, stride_h, box_w, box_h; } grid;
  float score_thr;
  double iou_thr;
  const int* view_off;     // device [n_labels + 1]: class views of label i = view_ids[view_off[i] .. view_off[i+1])
  const int* view_ids;     // device [C]
  int* cand;               // workspace [C * sumN]
  unsigned long long* keys;   // workspace [C * sumN] (only touched when a label needs more than one NMS chunk)
  int* out_ids;            // [C * sumN]: survivors of label i at view_off[i] * sumN, score-descending
  int* counts;             // [n_labels]
  int* offsets;            // [n_labels + 1] exclusive scan of counts (written by the last CTA)
  unsigned int* done;      // zero-initialised counter, reset by the kernel
};
int launch_label_nms(const DetectArgs& A, cudaStream_t st);
int launch_gather_detections(const DetectArgs& A, const long long* label_values, float* boxes, float* scores, long long* labels,
                             float* anchors, float* corners, cudaStream_t st);

// image pyramid level (pyramid.cu): Pillow-exact bilinear resize + ToTensor + Normalize
int launch_resize_level(const uint8_t* img, int H, int W, int out_h, int out_w, const int* xbounds, const int* xcoeffs, int xk,
                        const int* ybounds, const int* ycoeffs, int yk, const float* mean, const float* stdv, uint8_t* tmp,
                        float* out, uint8_t* out_u8, cudaStream_t st);

// detection evaluation (voc.cu)
int launch_voc_match(const float* det_boxes, const int* det_img, const int* det_label, const float* gt_boxes,
                     const int* gt_label, const int* gt_offsets, int n_det, float iou_thr, int* gt_index, cudaStream_t st);

size_t conv_weight_blob_bytes(int ksize, int in_chunks16);

// scatter-form last layer (conv3s.cu): h2 volume [planes][16 chunk8][H*W][8] -> params fp32 [planes][P][H*W]
size_t conv3s_weight_blob_bytes(int P);
int launch_conv3s(const void* h2_vol, const void* wblob, const float* bias, const float* inv_scale, float* out, int planes,
                  int P, int H, int W, int num_sms, cudaStream_t st);

}  // namespace os2d

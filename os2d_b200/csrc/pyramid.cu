// Image pyramid level on the device: the evaluation data path of the reference resizes the
// uint8 RGB image once per pyramid level with PIL (`img.resize((w, h), Image.BILINEAR)`, os2d/structures/transforms.py:72
// from os2d/data/dataloader.py:322-334) and then applies ToTensor + Normalize (dataloader.py:336-341) on the CPU.
// These two kernels reproduce Pillow's ImagingResample bit for bit - separable passes, horizontal first, uint8
// intermediate, 22-bit fixed-point coefficients, clip8((2^21 + sum pixel * coeff) >> 22) - and fuse the fp32
// (byte / 255 - mean) / std epilogue and the HWC -> CHW transposition into the vertical pass.  The coefficient tables
// (first tap, tap count, weights per output coordinate) are computed on the host in double precision like Pillow's
// precompute_coeffs (os2d_b200/pyramid.py).  Byte work, HBM-bound: one thread per output pixel, 3 channels each.
#include "common.cuh"
#include "kernels.h"

namespace os2d {

constexpr int kResizeBits = 32 - 8 - 2;   // Pillow PRECISION_BITS

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kResizeBits;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// [H][W][3] u8 -> [H][out_w][3] u8; bounds [out_w][2] = (first tap, taps), coeffs [out_w][ksize]
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ src, int H, int W, int out_w,
                                                        const int* __restrict__ bounds, const int* __restrict__ coeffs,
                                                        int ksize, uint8_t* __restrict__ dst) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (ox >= out_w) return;
  const int first = bounds[2 * ox], n = bounds[2 * ox + 1];
  const int* k = coeffs + static_cast<size_t>(ox) * ksize;
  const uint8_t* row = src + (static_cast<size_t>(y) * W + first) * 3;
  int a0 = 1 << (kResizeBits - 1), a1 = a0, a2 = a0;
  for (int t = 0; t < n; ++t) {
    const int c = k[t];
    a0 += row[3 * t] * c; a1 += row[3 * t + 1] * c; a2 += row[3 * t + 2] * c;
  }
  uint8_t* o = dst + (static_cast<size_t>(y) * out_w + ox) * 3;
  o[0] = clip8(a0); o[1] = clip8(a1); o[2] = clip8(a2);
}

// [H][w][3] u8 -> [3][out_h][w] fp32 = (clip8(...) / 255 - mean) / std   (torchvision ToTensor + Normalize, fp32, IEEE)
__global__ void __launch_bounds__(256) resize_v_norm_kernel(const uint8_t* __restrict__ src, int H, int w, int out_h,
                                                             const int* __restrict__ bounds, const int* __restrict__ coeffs,
                                                             int ksize, float m0, float m1, float m2, float s0, float s1,
                                                             float s2, float* __restrict__ dst, uint8_t* __restrict__ dst_u8) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y;
  if (x >= w) return;
  const int first = bounds[2 * oy], n = bounds[2 * oy + 1];
  const int* k = coeffs + static_cast<size_t>(oy) * ksize;
  const uint8_t* col = src + (static_cast<size_t>(first) * w + x) * 3;
  int a0 = 1 << (kResizeBits - 1), a1 = a0, a2 = a0;
  for (int t = 0; t < n; ++t) {
    const int c = k[t];
    const uint8_t* p = col + static_cast<size_t>(t) * w * 3;
    a0 += p[0] * c; a1 += p[1] * c; a2 += p[2] * c;
  }
  const uint8_t b0 = clip8(a0), b1 = clip8(a1), b2 = clip8(a2);
  const size_t plane = static_cast<size_t>(out_h) * w, o = static_cast<size_t>(oy) * w + x;
  dst[o] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(b0), 255.0f), m0), s0);
  dst[plane + o] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(b1), 255.0f), m1), s1);
  dst[2 * plane + o] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(b2), 255.0f), m2), s2);
  if (dst_u8 != nullptr) { uint8_t* q = dst_u8 + o * 3; q[0] = b0; q[1] = b1; q[2] = b2; }
}

int launch_resize_level(const uint8_t* img, int H, int W, int out_h, int out_w, const int* xbounds, const int* xcoeffs, int xk,
                        const int* ybounds, const int* ycoeffs, int yk, const float* mean, const float* stdv, uint8_t* tmp,
                        float* out, uint8_t* out_u8, cudaStream_t st) {
  if (H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0 || xk <= 0 || yk <= 0 || H > 65535 || out_h > 65535) return kErrBadArg;
  resize_h_kernel<<<dim3((out_w + 255) / 256, H), 256, 0, st>>>(img, H, W, out_w, xbounds, xcoeffs, xk, tmp);
  OS2D_AFTER_LAUNCH();
  resize_v_norm_kernel<<<dim3((out_w + 255) / 256, out_h), 256, 0, st>>>(tmp, H, out_w, out_h, ybounds, ycoeffs, yk, mean[0],
                                                                        mean[1], mean[2], stdv[0], stdv[1], stdv[2], out,
                                                                        out_u8);
  OS2D_AFTER_LAUNCH();
  return kOk;
}

}  // namespace os2d

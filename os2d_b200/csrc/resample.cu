// K3: affine grid generation + bilinear resample/pool of the correlation volume + box regression, one thread
// per (image, class, location); the [.,H,W,15,15,2] grid tensors of the reference are never materialised.
//   theta assembly / inverse     os2d/modeling/head.py:81-153
//   grid = theta * (x_j, y_i, 1) os2d/modeling/head.py:184   (F.affine_grid, align_corners=True)
//   local -> feature-map coords  os2d/modeling/head.py:18-40, 371-384  (px = 7.5 gx + x + 0.5, clamped)
//   resample + masked mean       os2d/modeling/head.py:439-520 (inner 11x11 points, channel k = j*15 + i)
//   box / corners / loc          os2d/modeling/head.py:404-433, box_coder.py:306-317, bounding_box.py:267-277
// Lanes run along x so that the 4 taps of a grid point are near-coalesced reads of one channel plane.
//
// The kernel is a template on its output sink: LocalSink writes the 13 output planes of a location into local tensors (or
// this rank's slice of a gather buffer); PeerSink is K3 fused with its collective (multi-GPU, class-axis sharding): every
// thread stores its 13 outputs into this rank's slice of the [G,B,C/G,13,N] gather buffer of EVERY rank through
// peer-mapped pointers (symmetric memory over NVLink 5 / NVSwitch), 13 coalesced 128 B row stores per warp and peer, and a
// device-side barrier across the ranks (os2d_b200/dist.py) replaces the all-gather.
#include "common.cuh"
#include "kernels.h"

namespace os2d {

struct LocalSink {
  float* score; float* loc; float* corners;
  long long score_ps, loc_ps, corners_ps;          // plane strides (floats)
  __device__ __forceinline__ void store(int plane, int N, int pix, float sc, const float (&l)[4], const float (&X)[4],
                                        const float (&Y)[4]) const {
    score[static_cast<size_t>(plane) * score_ps + pix] = sc;
    float* lo = loc + static_cast<size_t>(plane) * loc_ps + pix;
#pragma unroll
    for (int q = 0; q < 4; ++q) lo[static_cast<size_t>(q) * N] = l[q];
    float* co = corners + static_cast<size_t>(plane) * corners_ps + pix;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      co[static_cast<size_t>(2 * q) * N] = X[q];
      co[static_cast<size_t>(2 * q + 1) * N] = Y[q];
    }
  }
};

struct PeerSink {
  float* const* peers; int n_peers;                 // base of every rank's gather buffer (peer-mapped)
  long long score_off, loc_off, corners_off, plane_stride;
  __device__ __forceinline__ void store(int plane, int N, int pix, float sc, const float (&l)[4], const float (&X)[4],
                                        const float (&Y)[4]) const {
    const size_t po = static_cast<size_t>(plane) * plane_stride + pix;
    for (int p = 0; p < n_peers; ++p) {
      const LocalSink s{peers[p] + score_off + po, peers[p] + loc_off + po, peers[p] + corners_off + po, 0, 0, 0};
      s.store(0, N, 0, sc, l, X, Y);
    }
  }
};

__device__ __forceinline__ float lin15(int i) {
  const float step = 2.0f / 14.0f;
  return (i < 7) ? (-1.0f + step * i) : (1.0f - step * (14 - i));
}

template <typename Sink>
__global__ void __launch_bounds__(128, 12) resample_kernel(const __half* __restrict__ raw, const float* __restrict__ params,
                                                        int P, int H, int W, int inverse, float stride_w,
                                                        float stride_h, float box_w, float box_h, const Sink sink) {
  pdl_launch_dependents();
  pdl_wait();          // params / raw are the previous kernels' outputs
  const int N = H * W;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int plane = blockIdx.y;
  if (pix >= N) return;
  const int y = pix / W, x = pix - y * W;
  const float* pp = params + static_cast<size_t>(plane) * P * N + pix;
  float a, b, tx, c, d, ty;
  if (P == 6) {
    a = pp[0]; b = pp[N]; tx = pp[2 * N]; c = pp[3 * static_cast<size_t>(N)]; d = pp[4 * static_cast<size_t>(N)];
    ty = pp[5 * static_cast<size_t>(N)];
  } else {
    a = pp[0]; b = 0.f; tx = pp[N]; c = 0.f; d = pp[2 * N]; ty = pp[3 * static_cast<size_t>(N)];
  }
  if (inverse) invert_affine(a, b, tx, c, d, ty);

  // ---- score: mean of bilinear samples at the inner 11 x 11 grid points ----
  const __half* rplane = raw + static_cast<size_t>(plane) * kCorrCh * N;
  const float cx0 = x + 0.5f, cy0 = y + 0.5f;
  const float xmax = static_cast<float>(W - 1), ymax = static_cast<float>(H - 1);
  float acc = 0.f;
  const unsigned short* rp16 = reinterpret_cast<const unsigned short*>(rplane);
#pragma unroll 1
  for (int j = 2; j <= 12; ++j) {          // template x index
    const float xj = lin15(j);
    const float gx_j = fmaf(a, xj, tx), gy_j = fmaf(c, xj, ty);
    // phase 1: addresses and weights of the 11 points of this template column; phase 2: all 44 taps in flight;
    // phase 3: interpolate (keeps ~44 independent L2 gathers outstanding per thread instead of 4)
    int o00[11], ox1[11], oy1[11];
    float wxs[11], wys[11];
#pragma unroll
    for (int i = 2; i <= 12; ++i) {        // template y index
      const float yi = lin15(i);
      const float gx = fmaf(b, yi, gx_j), gy = fmaf(d, yi, gy_j);
      const float px = fminf(fmaxf(fmaf(gx, 7.5f, cx0), 0.f), xmax);
      const float py = fminf(fmaxf(fmaf(gy, 7.5f, cy0), 0.f), ymax);
      const float fx0 = floorf(px), fy0 = floorf(py);
      wxs[i - 2] = px - fx0;
      wys[i - 2] = py - fy0;
      const int x0 = static_cast<int>(fx0), y0 = static_cast<int>(fy0);
      o00[i - 2] = (j * kGrid + i) * N + y0 * W + x0;
      ox1[i - 2] = (x0 + 1 < W) ? 1 : 0;
      oy1[i - 2] = (y0 + 1 < H) ? W : 0;
    }
    unsigned short r00[11], r01[11], r10[11], r11[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) {
      r00[i] = __ldg(rp16 + o00[i]);
      r01[i] = __ldg(rp16 + o00[i] + ox1[i]);
      r10[i] = __ldg(rp16 + o00[i] + oy1[i]);
      r11[i] = __ldg(rp16 + o00[i] + oy1[i] + ox1[i]);
    }
#pragma unroll
    for (int i = 0; i < 11; ++i) {
      const float v00 = __half2float(__ushort_as_half(r00[i])), v01 = __half2float(__ushort_as_half(r01[i]));
      const float v10 = __half2float(__ushort_as_half(r10[i])), v11 = __half2float(__ushort_as_half(r11[i]));
      const float top = fmaf(wxs[i], v01 - v00, v00), bot = fmaf(wxs[i], v11 - v10, v10);
      acc += fmaf(wys[i], bot - top, top);
    }
  }
  const float sc = acc * (1.0f / 121.0f);

  // ---- box = bbox of the transformed grid (extremes are at the 4 corner points), corners, loc ----
  // anchor: centre (x + 0.5) * stride, size box (240 = 16 * 14 + 16 for the ResNet C4 backbones, head.py:216-238)
  const float acx = (x + 0.5f) * stride_w, acy = (y + 0.5f) * stride_h;
  const float hbw = 0.5f * box_w, hbh = 0.5f * box_h;
  float X[4], Y[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float yi = (q & 2) ? 1.0f : -1.0f;   // i = 0 / 14
    const float xj = (q & 1) ? 1.0f : -1.0f;   // j = 0 / 14
    const float gx = a * xj + b * yi + tx, gy = c * xj + d * yi + ty;
    X[q] = fmaf(gx, hbw, acx);
    Y[q] = fmaf(gy, hbh, acy);
  }
  const float x1 = fminf(fminf(X[0], X[1]), fminf(X[2], X[3])), y1 = fminf(fminf(Y[0], Y[1]), fminf(Y[2], Y[3]));
  float x2 = fmaxf(fmaxf(X[0], X[1]), fmaxf(X[2], X[3])), y2 = fmaxf(fmaxf(Y[0], Y[1]), fmaxf(Y[2], Y[3]));
  if (x1 + 1.0f > x2) x2 = x1 + 1.0f;
  if (y1 + 1.0f > y2) y2 = y1 + 1.0f;
  const float gw = x2 - x1, gh = y2 - y1;
  const float gcx = x1 + 0.5f * gw, gcy = y1 + 0.5f * gh;
  const float l[4] = {10.0f * (gcx - acx) / box_w, 10.0f * (gcy - acy) / box_h, 5.0f * logf(gw / box_w),
                      5.0f * logf(gh / box_h)};
  sink.store(plane, N, pix, sc, l, X, Y);
}

int launch_resample(const void* rawvol, const float* params, int planes, int P, int H, int W, int inverse,
                    float stride_w, float stride_h, float box_w, float box_h, float* score, float* loc,
                    float* corners, long long score_ps, long long loc_ps, long long corners_ps, cudaStream_t st) {
  if (planes <= 0 || (P != 4 && P != 6) || H < 2 || W < 2) return kErrBadArg;
  const int N = H * W;
  const LocalSink sink{score, loc, corners, score_ps, loc_ps, corners_ps};
  OS2D_CUDA_TRY(launch_pdl(resample_kernel<LocalSink>, dim3((N + 127) / 128, planes), dim3(128), 0, st, 1,
                           reinterpret_cast<const __half*>(rawvol), params, P, H, W, inverse, stride_w, stride_h, box_w, box_h,
                           sink));
  os2d::note_launch();
  return kOk;
}

int launch_resample_p2p(const void* rawvol, const float* params, int planes, int P, int H, int W, int inverse,
                        float stride_w, float stride_h, float box_w, float box_h, float* const* peers, int n_peers,
                        long long score_off, long long loc_off, long long corners_off, long long plane_stride, cudaStream_t st) {
  if (planes <= 0 || (P != 4 && P != 6) || H < 2 || W < 2 || n_peers <= 0 || !peers) return kErrBadArg;
  const int N = H * W;
  const PeerSink sink{peers, n_peers, score_off, loc_off, corners_off, plane_stride};
  OS2D_CUDA_TRY(launch_pdl(resample_kernel<PeerSink>, dim3((N + 127) / 128, planes), dim3(128), 0, st, 1,
                           reinterpret_cast<const __half*>(rawvol), params, P, H, W, inverse, stride_w, stride_h, box_w, box_h,
                           sink));
  os2d::note_launch();
  return kOk;
}

}  // namespace os2d

// Secondary entry points of the reference head that take / return the big tensors the fused path never materialises.
// They exist so that the public methods of the drop-in classes work stand-alone (API completeness), not for speed:
//   pack_corr_maps      fp32 correlation maps [planes,225,H,W] -> z volume + raw volume of the conv / sampler kernels
//                       (TransformationNet.forward input side: ReLU + L2 norm, os2d/modeling/head.py:648-650)
//   affine_grids        regressed parameters -> grid of transformed points [planes,H,W,15,15,2] in local coordinates
//                       (Os2dAlignment.forward: head.py:81-193, F.affine_grid align_corners=True)
//   resample_with_grid  Os2dHead.resample_of_correlation_map_fast / _simple (head.py:439-594): bilinear samples of channel
//                       k = tx*15 + ty at caller-given unit coordinates (border clamp), weighted by the pool mask, summed
#include "common.cuh"
#include "kernels.h"

namespace os2d {

__device__ __forceinline__ float aux_lin15(int i) {
  const float step = 2.0f / 14.0f;
  return (i < 7) ? (-1.0f + step * i) : (1.0f - step * (14 - i));
}

// thread per (plane, pixel): coalesced over pixels, two passes over the 225 channels (the second one hits L1/L2)
__global__ void __launch_bounds__(128) pack_corr_kernel(const float* __restrict__ corr, int N, __half* __restrict__ zvol,
                                                         __half* __restrict__ rawvol) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int plane = blockIdx.y;
  if (pix >= N) return;
  const float* src = corr + static_cast<size_t>(plane) * kCorrCh * N + pix;
  float s1 = 0.f, s2 = 0.f;
  for (int k = 0; k < kCorrCh; ++k) {
    const float v = fmaxf(src[static_cast<size_t>(k) * N], 0.f);
    s1 += v;
    s2 = fmaf(v, v, s2);
  }
  const float inv = 1.0f / (sqrtf(s2) + 1e-6f);
  const float mean = s1 * inv * (1.0f / kCorrCh);
  __half* zbase = zvol + (static_cast<size_t>(plane) * kZChunks * N + pix) * 8;
  __half* rbase = rawvol + static_cast<size_t>(plane) * kCorrCh * N + pix;
  for (int ch = 0; ch < kZChunks; ++ch) {
    __align__(16) __half hz[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = ch * 8 + j;
      float z = 0.f;
      if (k < kCorrCh) {
        const float a = src[static_cast<size_t>(k) * N];
        rbase[static_cast<size_t>(k) * N] = __float2half(a);
        z = (fmaxf(a, 0.f) * inv - mean) * kScaleZ;
      }
      hz[j] = __float2half(z);
    }
    if (ch == kDcCh / 8) {   // chunk 28: channel 224 real, 225/226 = fp16(8 mean), 227 = residual
      const float m8 = mean * kScaleMean;
      const __half mh = __float2half(m8);
      hz[1] = mh; hz[2] = mh; hz[3] = __float2half(m8 - __half2float(mh));
    }
    *reinterpret_cast<uint4*>(zbase + static_cast<size_t>(ch) * N * 8) = *reinterpret_cast<const uint4*>(hz);
  }
}

// thread per (plane, pixel): writes the 225 x 2 grid values of its location (order i (y) major, j (x), then (gx, gy))
__global__ void __launch_bounds__(128) affine_grid_kernel(const float* __restrict__ params, int P, int N, int inverse,
                                                           float* __restrict__ grid) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int plane = blockIdx.y;
  if (pix >= N) return;
  const float* pp = params + static_cast<size_t>(plane) * P * N + pix;
  float a, b, tx, c, d, ty;
  if (P == 6) { a = pp[0]; b = pp[N]; tx = pp[2 * static_cast<size_t>(N)]; c = pp[3 * static_cast<size_t>(N)]; d = pp[4 * static_cast<size_t>(N)]; ty = pp[5 * static_cast<size_t>(N)]; }
  else { a = pp[0]; b = 0.f; tx = pp[N]; c = 0.f; d = pp[2 * static_cast<size_t>(N)]; ty = pp[3 * static_cast<size_t>(N)]; }
  if (inverse) invert_affine(a, b, tx, c, d, ty);
  float2* g = reinterpret_cast<float2*>(grid) + (static_cast<size_t>(plane) * N + pix) * kCorrCh;
  for (int i = 0; i < kGrid; ++i) {
    const float yi = aux_lin15(i);
    for (int j = 0; j < kGrid; ++j) {
      const float xj = aux_lin15(j);
      g[i * kGrid + j] = make_float2(a * xj + b * yi + tx, c * xj + d * yi + ty);
    }
  }
}

// thread per (plane, pixel); grid [plane][pixel][i][j][2] unit coordinates in [-1, 1]; mask [C][225] indexed i*15+j
__global__ void __launch_bounds__(128) resample_grid_kernel(const float* __restrict__ corr, const float* __restrict__ grid,
                                                             const float* __restrict__ mask, int C, int H, int W,
                                                             float* __restrict__ out) {
  const int N = H * W;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int plane = blockIdx.y;
  if (pix >= N) return;
  const float* cplane = corr + static_cast<size_t>(plane) * kCorrCh * N;
  const float2* g = reinterpret_cast<const float2*>(grid) + (static_cast<size_t>(plane) * N + pix) * kCorrCh;
  const float* m = mask + static_cast<size_t>(plane % C) * kCorrCh;
  float acc = 0.f;
  for (int i = 0; i < kGrid; ++i) {
    for (int j = 0; j < kGrid; ++j) {
      const float wgt = m[i * kGrid + j];
      if (wgt == 0.f) continue;
      const float2 u = g[i * kGrid + j];
      const float ux = fminf(fmaxf(u.x, -1.f), 1.f), uy = fminf(fmaxf(u.y, -1.f), 1.f);
      const float px = (ux + 1.f) * 0.5f * (W - 1), py = (uy + 1.f) * 0.5f * (H - 1);
      const float fx0 = fminf(floorf(px), static_cast<float>(W - 1)), fy0 = fminf(floorf(py), static_cast<float>(H - 1));
      const float wx = px - fx0, wy = py - fy0;
      const int x0 = static_cast<int>(fx0), y0 = static_cast<int>(fy0);
      const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
      const float* ch = cplane + static_cast<size_t>(j * kGrid + i) * N;     // transposed channel order (head.py:480)
      const float v00 = ch[y0 * W + x0], v01 = ch[y0 * W + x1], v10 = ch[y1 * W + x0], v11 = ch[y1 * W + x1];
      const float top = fmaf(wx, v01 - v00, v00), bot = fmaf(wx, v11 - v10, v10);
      acc = fmaf(wgt, fmaf(wy, bot - top, top), acc);
    }
  }
  out[static_cast<size_t>(plane) * N + pix] = acc;
}

int launch_pack_corr(const float* corr, int planes, int N, void* zvol, void* rawvol, cudaStream_t st) {
  if (planes <= 0 || N <= 0) return kErrBadArg;
  pack_corr_kernel<<<dim3((N + 127) / 128, planes), 128, 0, st>>>(corr, N, reinterpret_cast<__half*>(zvol),
                                                                 reinterpret_cast<__half*>(rawvol));
  OS2D_AFTER_LAUNCH();
  return kOk;
}

int launch_affine_grids(const float* params, int planes, int P, int N, int inverse, float* grid, cudaStream_t st) {
  if (planes <= 0 || N <= 0 || (P != 4 && P != 6)) return kErrBadArg;
  affine_grid_kernel<<<dim3((N + 127) / 128, planes), 128, 0, st>>>(params, P, N, inverse, grid);
  OS2D_AFTER_LAUNCH();
  return kOk;
}

int launch_resample_grid(const float* corr, const float* grid, const float* mask, int planes, int C, int H, int W, float* out,
                         cudaStream_t st) {
  if (planes <= 0 || C <= 0 || H < 2 || W < 2) return kErrBadArg;
  resample_grid_kernel<<<dim3((H * W + 127) / 128, planes), 128, 0, st>>>(corr, grid, mask, C, H, W, out);
  OS2D_AFTER_LAUNCH();
  return kOk;
}

}  // namespace os2d

// K0: operand preparation.
//  pack_class_features : bilinear resize of class feature maps to 15x15 (align_corners, zero pad),
//                        L2-normalise over D, emit fp32 [C,D,15,15] (API attribute) and the fp16
//                        K-major GEMM operand [C][240][D] with channel order k = tx*15 + ty.
//                        (reference: os2d/modeling/head.py:241-259, 293, 342-344)
//  pack_image_features : L2-normalise the image feature map over D and transpose to fp16 [B][N][D].
//                        (reference: os2d/modeling/head.py:339, 597-601)
#include "common.cuh"
#include "kernels.h"

namespace os2d {

__device__ __forceinline__ float linspace15(int i) {
  // torch.linspace(-1, 1, 15): symmetric evaluation, step 2/14
  const float step = 2.0f / 14.0f;
  return (i < 15 / 2) ? (-1.0f + step * i) : (1.0f - step * (14 - i));
}

// grid (240, C), block 256.  Rows 225..239 of the packed operand are zero padding.
__global__ void __launch_bounds__(256) pack_class_kernel(const float* __restrict__ maps, int D, int h, int w,
                                                          int normalize, float* __restrict__ cf32,
                                                          __half* __restrict__ packed) {
  const int k = blockIdx.x;  // packed row: k = tx*15 + ty
  const int c = blockIdx.y;
  __half* prow = packed + (static_cast<size_t>(c) * kCorrPad + k) * D;
  if (k >= kCorrCh) {
    for (int d = threadIdx.x; d < D; d += blockDim.x) prow[d] = __float2half(0.f);
    return;
  }
  const int tx = k / kGrid, ty = k % kGrid;
  const float ys = (linspace15(ty) + 1.0f) * 0.5f * (h - 1);
  const float xs = (linspace15(tx) + 1.0f) * 0.5f * (w - 1);
  const float y0f = floorf(ys), x0f = floorf(xs);
  const float wy1 = ys - y0f, wx1 = xs - x0f, wy0 = 1.f - wy1, wx0 = 1.f - wx1;
  const int y0 = static_cast<int>(y0f), x0 = static_cast<int>(x0f), y1 = y0 + 1, x1 = x0 + 1;
  const bool vy0 = y0 >= 0 && y0 < h, vy1 = y1 >= 0 && y1 < h, vx0 = x0 >= 0 && x0 < w, vx1 = x1 >= 0 && x1 < w;
  const float* base = maps + static_cast<size_t>(c) * D * h * w;

  extern __shared__ float vals[];  // D floats
  float ss = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float* pl = base + static_cast<size_t>(d) * h * w;
    float v00 = (vy0 && vx0) ? pl[y0 * w + x0] : 0.f;
    float v01 = (vy0 && vx1) ? pl[y0 * w + x1] : 0.f;
    float v10 = (vy1 && vx0) ? pl[y1 * w + x0] : 0.f;
    float v11 = (vy1 && vx1) ? pl[y1 * w + x1] : 0.f;
    float v = v00 * (wy0 * wx0) + v01 * (wy0 * wx1) + v10 * (wy1 * wx0) + v11 * (wy1 * wx1);
    vals[d] = v;
    ss += v * v;
  }
  __shared__ float red[8];
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) tot += red[i];
  const float inv = normalize ? 1.0f / (sqrtf(tot) + 1e-5f) : 1.0f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float v = vals[d] * inv;
    cf32[((static_cast<size_t>(c) * D + d) * kGrid + ty) * kGrid + tx] = v;
    prow[d] = __float2half(v * kScaleFeat);
  }
}

// L2-normalise over D and transpose [B][D][N] fp32 -> [B][N][D] fp16 (x 32), one block per 32 pixels.
// Phase 1 accumulates the squared norms (coalesced 128 B rows, 8 warps striding the channels); phase 2 re-reads the
// same 128 KB (L2 hits) in 64-channel tiles, transposes through shared memory and writes 128 B per pixel row.
// grid (ceil(N/32), B), block 256
__global__ void __launch_bounds__(256) image_pack_kernel(const float* __restrict__ fm, int D, int N,
                                                          float* __restrict__ inv_out, __half* __restrict__ out) {
  __shared__ float red[8][33];
  __shared__ float invs[32];
  __shared__ float tile[64][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.y, p0 = blockIdx.x * 32, p = p0 + lane;
  const float* src = fm + static_cast<size_t>(b) * D * N;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (p < N) {
    int d = w;
    for (; d + 24 < D; d += 32) {
      const float a0 = src[static_cast<size_t>(d) * N + p], a1 = src[static_cast<size_t>(d + 8) * N + p];
      const float a2 = src[static_cast<size_t>(d + 16) * N + p], a3 = src[static_cast<size_t>(d + 24) * N + p];
      s0 = fmaf(a0, a0, s0); s1 = fmaf(a1, a1, s1); s2 = fmaf(a2, a2, s2); s3 = fmaf(a3, a3, s3);
    }
    for (; d < D; d += 8) { const float a = src[static_cast<size_t>(d) * N + p]; s0 = fmaf(a, a, s0); }
  }
  red[w][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (w == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][lane];
    const float inv = kScaleFeat / (sqrtf(t) + 1e-5f);
    invs[lane] = inv;
    if (p < N) inv_out[static_cast<size_t>(b) * N + p] = inv;
  }
  __syncthreads();
  for (int d0 = 0; d0 < D; d0 += 64) {
#pragma unroll
    for (int r = w; r < 64; r += 8) tile[r][lane] = (p < N) ? src[static_cast<size_t>(d0 + r) * N + p] : 0.f;
    __syncthreads();
#pragma unroll
    for (int pr = w; pr < 32; pr += 8) {
      if (p0 + pr < N) {
        const float sc = invs[pr];
        const __half2 v = __floats2half2_rn(tile[2 * lane][pr] * sc, tile[2 * lane + 1][pr] * sc);
        *reinterpret_cast<__half2*>(out + (static_cast<size_t>(b) * N + p0 + pr) * D + d0 + 2 * lane) = v;
      }
    }
    __syncthreads();
  }
}

int launch_pack_class(const float* maps, int C, int D, int h, int w, int normalize, float* cf32, void* packed,
                      cudaStream_t st) {
  if (C <= 0 || D <= 0 || h <= 0 || w <= 0 || D > 12288) return kErrBadArg;
  dim3 grid(kCorrPad, C);
  pack_class_kernel<<<grid, 256, D * sizeof(float), st>>>(maps, D, h, w, normalize, cf32,
                                                           reinterpret_cast<__half*>(packed));
  OS2D_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int launch_pack_image(const float* fm, int B, int D, int N, float* inv_ws, void* packed, cudaStream_t st) {
  if (B <= 0 || D <= 0 || N <= 0 || (D % 64) != 0) return kErrBadArg;
  image_pack_kernel<<<dim3((N + 31) / 32, B), 256, 0, st>>>(fm, D, N, inv_ws, reinterpret_cast<__half*>(packed));
  OS2D_CUDA_TRY(cudaGetLastError());
  return kOk;
}

}  // namespace os2d

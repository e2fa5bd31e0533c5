// K0: operand preparation.
//  pack_class_features : bilinear resize of class feature maps to 15x15 (align_corners, zero pad),
//                        L2-normalise over D, emit fp32 [C,D,15,15] (API attribute) and the fp16
//                        K-major GEMM operand [C][240][D] with channel order k = tx*15 + ty; uniform
//                        ([C,D,h,w]) or ragged (per-class pointer + size) input, one launch either way.
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

// One thread-block CLUSTER of 8 CTAs per class, CTA r owning channels [r D/8, (r+1) D/8):
//   pass 1  bilinear resize of every channel plane to 15 x 15 (thread = one output point kk = ty*15 + tx for half of the
//           CTA's channels: the 4 taps of neighbouring points share sectors of the same ~1 KB plane; the unnormalised
//           values go out as coalesced 900 B rows) + squared norms over the CTA's channels;
//   exchange the 225 partial sums of the 8 CTAs meet through distributed shared memory (fixed rank order: deterministic);
//   pass 2  re-read the CTA's own rows (L2 hits), scale, and emit the fp16 operand transposed to [k = tx*15 + ty][D]
//           through a 64-channel shared-memory tile (128 B row segments).
// HBM traffic = input + 3 x cf32 + operand.  History (100 classes, D = 1024): one block per output row with 4-byte
// accesses at 900 B strides 0.61 ms; one 1024-thread block per class 0.63 ms (latency-bound: 256 dependent iterations);
// this version: profiles/r01_class_pipeline.json.
// Uniform call: maps [C,D,h,w] contiguous, map_ptrs == nullptr.  Ragged call: per-class pointer and (h, w).
constexpr int kPackCluster = 8;
constexpr int kPackThreads = 512;
constexpr int kPackTileD = 64;
constexpr int kPackTileRS = 227;     // halfs per tile row, odd => the transposed reads are bank-conflict free

__device__ __forceinline__ uint32_t pack_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void pack_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float pack_ld_dsmem(const float* local, uint32_t rank) {
  uint32_t remote;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
  return v;
}

__global__ void __launch_bounds__(kPackThreads) pack_class_kernel(const float* __restrict__ maps,
                                                                   const float* const* __restrict__ map_ptrs,
                                                                   const int* __restrict__ hw, int h_u, int w_u, int D,
                                                                   int normalize, float* __restrict__ cf32,
                                                                   __half* __restrict__ packed) {
  __shared__ float red[2][256];
  __shared__ float part[256];
  __shared__ float inv_s[256];
  __shared__ __half tile[kPackTileD * kPackTileRS];
  const uint32_t rank = pack_cluster_rank();
  const int c = blockIdx.x / kPackCluster;
  int h = h_u, w = w_u;
  const float* base;
  if (map_ptrs != nullptr) { base = map_ptrs[c]; h = hw[2 * c]; w = hw[2 * c + 1]; }
  else base = maps + static_cast<size_t>(c) * D * h * w;
  const int t = threadIdx.x, kk = t & 255, dl = t >> 8, lane = t & 31, warp = t >> 5;
  // channel range of this CTA: multiples of the tile depth
  const int per = ((D + kPackCluster - 1) / kPackCluster + kPackTileD - 1) / kPackTileD * kPackTileD;
  const int d_begin = min(static_cast<int>(rank) * per, D), d_end = min(d_begin + per, D);
  float* cfc = cf32 + static_cast<size_t>(c) * D * kCorrCh;
  __half* pc = packed + static_cast<size_t>(c) * kCorrPad * D;

  // ---- pass 1: bilinear resize (align_corners, zero pad) + squared norms ----
  float ss = 0.f;
  if (kk < kCorrCh) {
    const int ty = kk / kGrid, tx = kk - ty * kGrid;
    const float ys = (linspace15(ty) + 1.0f) * 0.5f * (h - 1);
    const float xs = (linspace15(tx) + 1.0f) * 0.5f * (w - 1);
    const float y0f = floorf(ys), x0f = floorf(xs);
    const float wy1 = ys - y0f, wx1 = xs - x0f, wy0 = 1.f - wy1, wx0 = 1.f - wx1;
    const int y0 = static_cast<int>(y0f), x0 = static_cast<int>(x0f), y1 = y0 + 1, x1 = x0 + 1;
    const bool vy0 = y0 >= 0 && y0 < h, vy1 = y1 >= 0 && y1 < h, vx0 = x0 >= 0 && x0 < w, vx1 = x1 >= 0 && x1 < w;
    // an invalid tap reads the (always valid) element 0 with weight 0: no divergent loads
    const float w00 = (vy0 && vx0) ? wy0 * wx0 : 0.f, w01 = (vy0 && vx1) ? wy0 * wx1 : 0.f;
    const float w10 = (vy1 && vx0) ? wy1 * wx0 : 0.f, w11 = (vy1 && vx1) ? wy1 * wx1 : 0.f;
    const int o00 = (vy0 && vx0) ? y0 * w + x0 : 0, o01 = (vy0 && vx1) ? y0 * w + x1 : 0;
    const int o10 = (vy1 && vx0) ? y1 * w + x0 : 0, o11 = (vy1 && vx1) ? y1 * w + x1 : 0;
    const size_t plane = static_cast<size_t>(h) * w;
#pragma unroll 4
    for (int d = d_begin + dl; d < d_end; d += 2) {
      const float* pl = base + static_cast<size_t>(d) * plane;
      const float v = pl[o00] * w00 + pl[o01] * w01 + pl[o10] * w10 + pl[o11] * w11;
      cfc[static_cast<size_t>(d) * kCorrCh + kk] = v;
      ss += v * v;
    }
  }
  red[dl][kk] = ss;
  __syncthreads();
  if (t < 256) part[t] = red[0][t] + red[1][t];
  pack_cluster_sync();                         // every CTA's partial sums are visible cluster-wide
  if (t < kCorrCh) {
    float tot = 0.f;
#pragma unroll
    for (uint32_t r = 0; r < kPackCluster; ++r) tot += pack_ld_dsmem(&part[t], r);
    inv_s[t] = normalize ? 1.0f / (sqrtf(tot) + 1e-5f) : 1.0f;
  }
  __syncthreads();   // also orders this block's global writes of pass 1 before its reads below

  // ---- pass 2: normalise, write cf32 (API attribute) and the transposed fp16 operand ----
  for (int d0 = d_begin; d0 < d_end; d0 += kPackTileD) {
#pragma unroll 4
    for (int e = t; e < kPackTileD * kCorrCh; e += kPackThreads) {
      const int dd = e / kCorrCh, k2 = e - dd * kCorrCh;
      const int d = d0 + dd;
      float v = 0.f;
      if (d < d_end) {
        float* q = cfc + static_cast<size_t>(d) * kCorrCh + k2;
        v = *q * inv_s[k2];
        *q = v;
      }
      tile[dd * kPackTileRS + k2] = __float2half(v * kScaleFeat);
    }
    __syncthreads();
    const int d = d0 + 2 * lane;
    if (d < d_end) {
      for (int kp = warp; kp < kCorrCh; kp += kPackThreads / 32) {   // operand row kp = tx*15 + ty <-> cf32 point ty*15 + tx
        const int tx = kp / kGrid, ty = kp - tx * kGrid;
        const int k2 = ty * kGrid + tx;
        const __half2 v = __halves2half2(tile[(2 * lane) * kPackTileRS + k2], tile[(2 * lane + 1) * kPackTileRS + k2]);
        *reinterpret_cast<__half2*>(pc + static_cast<size_t>(kp) * D + d) = v;
      }
    }
    __syncthreads();
  }
  // rows 225..239 of the operand are zero padding (this CTA's channel range)
  const int nd2 = (d_end - d_begin) / 2;
  for (int e = t; e < (kCorrPad - kCorrCh) * nd2; e += kPackThreads) {
    const int row = e / nd2, j = e - row * nd2;
    *reinterpret_cast<__half2*>(pc + static_cast<size_t>(kCorrCh + row) * D + d_begin + 2 * j) = __floats2half2_rn(0.f, 0.f);
  }
  pack_cluster_sync();                         // nobody exits while a peer may still read its partial sums
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
  pdl_launch_dependents();
  pdl_wait();          // the packed operand / norm workspace may still be read by the previous image's kernels
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

// Channels-last variant ("backbone whose last layer writes the L2-normalised, MMA-ready feature
// layout"): the backbone runs in torch.channels_last, so its output rows [B*N][D] already ARE the operand layout and no
// transpose is needed.  One warp per location: x = a (+ b, ReLU: the residual add + ReLU that ends the last bottleneck of
// layer3, os2d/modeling/feature_extractor.py:23-72, fused here), squared norm by shuffles, fp16 row = x * 32 / (|x| + 1e-5).
// The whole row lives in registers (D <= 1024 per pass of 32 lanes x 32 values); a / b are fp32 or fp16.
template <typename T>
__device__ __forceinline__ float load_as_float(const T* p);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p) { return __half2float(*p); }

template <typename T>
__global__ void __launch_bounds__(256) image_pack_nhwc_kernel(const T* __restrict__ a, const T* __restrict__ b, int relu,
                                                               long long rows, int D, __half* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* pa = a + row * D;
  const T* pb = b ? b + row * D : nullptr;
  __half* po = out + row * D;
  constexpr int kMaxPer = 32;                       // D <= 32 lanes * 2 * 32 = 2048
  float2 v[kMaxPer];
  float ss = 0.f;
  const int pairs = D >> 1;                         // lane handles element pairs lane, lane + 32, ...
#pragma unroll
  for (int i = 0; i < kMaxPer; ++i) {
    const int e = lane + 32 * i;
    float x0 = 0.f, x1 = 0.f;
    if (e < pairs) {
      x0 = load_as_float(pa + 2 * e); x1 = load_as_float(pa + 2 * e + 1);
      if (pb) { x0 += load_as_float(pb + 2 * e); x1 += load_as_float(pb + 2 * e + 1); }
      if (relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
    }
    v[i] = make_float2(x0, x1);
    ss = fmaf(x0, x0, fmaf(x1, x1, ss));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float sc = kScaleFeat / (sqrtf(ss) + 1e-5f);
#pragma unroll
  for (int i = 0; i < kMaxPer; ++i) {
    const int e = lane + 32 * i;
    if (e < pairs) *reinterpret_cast<__half2*>(po + 2 * e) = __floats2half2_rn(v[i].x * sc, v[i].y * sc);
  }
}

int launch_pack_image_nhwc(const void* a, const void* b, int is_half, int relu, long long rows, int D, void* packed,
                           cudaStream_t st) {
  if (rows <= 0 || D <= 0 || (D % 64) != 0 || D > 2048) return kErrBadArg;
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  if (is_half) {
    OS2D_CUDA_TRY(launch_pdl(image_pack_nhwc_kernel<__half>, dim3(grid), dim3(256), 0, st, 1, reinterpret_cast<const __half*>(a),
                             reinterpret_cast<const __half*>(b), relu, rows, D, reinterpret_cast<__half*>(packed)));
  } else {
    OS2D_CUDA_TRY(launch_pdl(image_pack_nhwc_kernel<float>, dim3(grid), dim3(256), 0, st, 1, reinterpret_cast<const float*>(a),
                             reinterpret_cast<const float*>(b), relu, rows, D, reinterpret_cast<__half*>(packed)));
  }
  os2d::note_launch();
  return kOk;
}

static int launch_pack_class_any(const float* maps, const float* const* map_ptrs, const int* hw, int h, int w, int C,
                                 int D, int normalize, float* cf32, void* packed, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPackCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3(static_cast<unsigned>(C) * kPackCluster);
  cfg.blockDim = dim3(kPackThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cfg.attrs = attr; cfg.numAttrs = 1;
  OS2D_CUDA_TRY(cudaLaunchKernelEx(&cfg, pack_class_kernel, maps, map_ptrs, hw, h, w, D, normalize, cf32,
                                   reinterpret_cast<__half*>(packed)));
  os2d::note_launch();
  return kOk;
}

int launch_pack_class(const float* maps, int C, int D, int h, int w, int normalize, float* cf32, void* packed,
                      cudaStream_t st) {
  if (C <= 0 || D <= 0 || h <= 0 || w <= 0 || (D & 1)) return kErrBadArg;
  return launch_pack_class_any(maps, nullptr, nullptr, h, w, C, D, normalize, cf32, packed, st);
}

int launch_pack_class_ragged(const float* const* map_ptrs, const int* hw, int C, int D, int normalize, float* cf32,
                             void* packed, cudaStream_t st) {
  if (C <= 0 || D <= 0 || (D & 1)) return kErrBadArg;
  return launch_pack_class_any(nullptr, map_ptrs, hw, 0, 0, C, D, normalize, cf32, packed, st);
}

int launch_pack_image(const float* fm, int B, int D, int N, float* inv_ws, void* packed, cudaStream_t st) {
  if (B <= 0 || D <= 0 || N <= 0 || (D % 64) != 0) return kErrBadArg;
  OS2D_CUDA_TRY(launch_pdl(image_pack_kernel, dim3((N + 31) / 32, B), dim3(256), 0, st, 1, fm, D, N, inv_ws,
                           reinterpret_cast<__half*>(packed)));
  os2d::note_launch();
  return kOk;
}

}  // namespace os2d

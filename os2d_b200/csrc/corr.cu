// K1: dense correlation as a tcgen05 GEMM with a fused ReLU / L2-norm / centering epilogue.
//
//   corr[b*C+c, k, p] = sum_d F[b, p, d] * Cf[c, k, d]        (reference: os2d/modeling/head.py:342-350)
//   z = relu(corr) / (||relu(corr)||_225 + 1e-6)              (reference: head.py:650, 597-601)
//
// One CTA tile = 128 image locations (MMA M, one TMEM lane per location) x one class's 240 padded
// correlation channels (MMA N), K = D in 64-element TMA boxes (128B swizzle) through a 4-stage
// mbarrier ring.  Warp roles: 0 = TMA producer, 1 = MMA issuer (one thread), 2 = TMEM allocator,
// 4..19 = epilogue: 4 lane quadrants x 4 column groups; a thread owns one location and 64 channels, the
// 225-channel sums are combined over the 4 column groups through shared memory (one named barrier per tile).
// The accumulator is double buffered in TMEM (2 x 256 columns) so the epilogue of tile t overlaps
// the MMAs of tile t+1.
//
// Outputs (never the fp32 [C,225,H,W] volume of the reference):
//   zvol  fp16 [plane][30 chunks][N][8 ch]  = (z - mean_k z) * 64 for k < 225, DC side channels
//         225/226 = fp16(8*mean), 227 = fp16 residual of 8*mean, rest 0   (conv1 B operand)
//   rawvol fp16 [plane][225][N]             = corr                        (sampler input)
#include "common.cuh"
#include "kernels.h"

namespace os2d {
namespace corr {

constexpr int BM = 128, BN = kCorrPad, BK = 64, STAGES = 4;
constexpr uint32_t A_BYTES = BM * BK * 2;          // 16384
constexpr uint32_t B_BYTES = BN * BK * 2;          // 30720
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;  // 47104 = 46 * 1024
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2 * 4 * 128 * 8 /*partial sums*/;
constexpr int THREADS = 128 + 512;   // 4 control warps + 16 epilogue warps
constexpr uint32_t TMEM_COLS = 512, ACC_COLS = 256;

struct Params {
  int B, C, D, N, MT;       // MT = m-tiles per plane
  int total_tiles;
  __half* zvol;
  __half* rawvol;
};

__global__ void __launch_bounds__(THREADS, 1)
corr_kernel(const __grid_constant__ CUtensorMap map_img, const __grid_constant__ CUtensorMap map_cls, Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float2* part = reinterpret_cast<float2*>(smem + STAGES * STAGE_BYTES + 256);   // [2 acc stages][4 col groups][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = P.D / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_img);
    tma_prefetch_desc(&map_cls);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 16); }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
        const int plane = t / P.MT, mt = t - plane * P.MT;
        const int b = plane / P.C, c = plane - b * P.C;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          tma_load_3d(sa, &map_img, &full[stage], kb * BK, mt * BM, b);
          tma_load_3d(sa + A_BYTES, &map_cls, &full[stage], kb * BK, 0, c);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * ACC_COLS;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // 128B-swizzled K-major tiles: 8-row groups 1024 B apart, K advance = 32 B inside the atom
            const uint64_t da = umma_smem_desc(sa + k * 32, 16, 1024, 2);
            const uint64_t db = umma_smem_desc(sb + k * 32, 16, 1024, 2);
            umma_f16(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue (16 warps) ------------------------------
    // warp -> TMEM lane quadrant q = warp % 4 (hardware restriction) and column group cg = (warp - 4) / 4:
    // columns [64 cg, 64 cg + 64) (the last group holds 48: channels 192..224, the DC side channels and padding).
    const int q = warp & 3, cg = (warp - 4) >> 2;
    const int ncb = (cg == 3) ? 3 : 4;                      // 16-column blocks owned by this warp
    int as = 0; uint32_t aphase = 0;
    const float inv_n = 1.0f / static_cast<float>(kCorrCh);
    const float corr_scale = 1.0f / (kScaleFeat * kScaleFeat);
    for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
      const int plane = t / P.MT, mt = t - plane * P.MT;
      const int row = q * 32 + lane;
      const int pix = mt * BM + row;
      const bool valid = pix < P.N;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + as * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16) + cg * 64;

      // pass 1: partial sums over this warp's columns (accumulator = 1024 * corr), exchanged through shared memory
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int cb = 0; cb < ncb; ++cb) {
        uint32_t r[16];
        tmem_ld16(tbase + cb * 16, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float v = fmaxf(__uint_as_float(r[j]), 0.f);
          s1 += v;
          s2 = fmaf(v, v, s2);
        }
      }
      float2* pbuf = part + as * (4 * BM);
      pbuf[cg * BM + row] = make_float2(s1, s2);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
      s1 = 0.f; s2 = 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) { const float2 v = pbuf[g * BM + row]; s1 += v.x; s2 += v.y; }

      // z_k = relu(acc_k) / (sqrt(s2) + 1024 * 1e-6)
      const float inv = 1.0f / (sqrtf(s2) + (kScaleFeat * kScaleFeat) * 1e-6f);
      const float mean = s1 * inv * inv_n;
      const float zs = inv * kScaleZ, zo = -mean * kScaleZ;

      __half* zbase = P.zvol + ((static_cast<size_t>(plane) * kZChunks + cg * 8) * P.N + pix) * 8;
      __half* rptr = P.rawvol + (static_cast<size_t>(plane) * kCorrCh + cg * 64) * P.N + pix;
      // pass 2: z (centred, fp16, chunk8 layout) and raw correlation (fp16, channel-major)
#pragma unroll 1
      for (int cb = 0; cb < ncb; ++cb) {
        uint32_t r[16];
        tmem_ld16(tbase + cb * 16, r);
        tmem_ld_wait();
        uint32_t zq[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float a0 = __uint_as_float(r[j]), a1 = __uint_as_float(r[j + 1]);
          const __half2 hz = __floats2half2_rn(fmaf(fmaxf(a0, 0.f), zs, zo), fmaf(fmaxf(a1, 0.f), zs, zo));
          zq[j >> 1] = *reinterpret_cast<const uint32_t*>(&hz);
          const __half2 hr = __floats2half2_rn(a0 * corr_scale, a1 * corr_scale);
          const int k = cg * 64 + cb * 16 + j;
          if (valid && k < kCorrCh) rptr[0] = __low2half(hr);
          if (valid && k + 1 < kCorrCh) rptr[P.N] = __high2half(hr);
          rptr += 2 * static_cast<size_t>(P.N);
        }
        if (cg == 3 && cb == 2) {
          // channels 224..239: 224 real, 225/226 = fp16(8 * mean), 227 = residual, 228.. zero
          const float m8 = mean * kScaleMean;
          const __half mh = __float2half(m8);
          const __half ml = __float2half(m8 - __half2float(mh));
          const __half2 p0 = __halves2half2(__low2half(*reinterpret_cast<const __half2*>(&zq[0])), mh);
          const __half2 p1 = __halves2half2(mh, ml);
          zq[0] = *reinterpret_cast<const uint32_t*>(&p0);
          zq[1] = *reinterpret_cast<const uint32_t*>(&p1);
          zq[2] = zq[3] = zq[4] = zq[5] = zq[6] = zq[7] = 0u;
        }
        if (valid) {
          *reinterpret_cast<uint4*>(zbase + static_cast<size_t>(2 * cb) * P.N * 8) = make_uint4(zq[0], zq[1], zq[2], zq[3]);
          *reinterpret_cast<uint4*>(zbase + static_cast<size_t>(2 * cb + 1) * P.N * 8) = make_uint4(zq[4], zq[5], zq[6], zq[7]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace corr

int launch_corr(const void* img_packed, const void* cls_packed, int B, int C, int D, int H, int W, void* zvol,
                void* rawvol, int num_sms, cudaStream_t st) {
  using namespace corr;
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || D % BK != 0) return kErrBadArg;
  const int N = H * W;
  CUtensorMap map_img, map_cls;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(D), static_cast<uint64_t>(N), static_cast<uint64_t>(B)};
    uint64_t strides[2] = {static_cast<uint64_t>(D) * 2, static_cast<uint64_t>(D) * 2 * N};
    uint32_t box[3] = {BK, BM, 1};
    int rc = encode_tensor_map(&map_img, img_packed, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc != kOk) return rc;
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(D), static_cast<uint64_t>(BN), static_cast<uint64_t>(C)};
    uint64_t strides[2] = {static_cast<uint64_t>(D) * 2, static_cast<uint64_t>(D) * 2 * BN};
    uint32_t box[3] = {BK, BN, 1};
    int rc = encode_tensor_map(&map_cls, cls_packed, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc != kOk) return rc;
  }
  Params P;
  P.B = B; P.C = C; P.D = D; P.N = N;
  P.MT = (N + BM - 1) / BM;
  P.total_tiles = B * C * P.MT;
  P.zvol = reinterpret_cast<__half*>(zvol);
  P.rawvol = reinterpret_cast<__half*>(rawvol);
  static bool attr_set = false;
  if (!attr_set) {
    OS2D_CUDA_TRY(cudaFuncSetAttribute(corr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  const int grid = P.total_tiles < num_sms ? P.total_tiles : num_sms;
  corr_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(map_img, map_cls, P);
  OS2D_CUDA_TRY(cudaGetLastError());
  return kOk;
}

}  // namespace os2d

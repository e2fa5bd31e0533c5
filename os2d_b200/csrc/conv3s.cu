// K2c: last TransformNet layer (conv 64 -> P, k5 p2, reference head.py:629 `linear`) in *scatter form*.
//
// With only P = 6 (or 4) output channels the direct implicit GEMM wastes the 128-row MMA on 12 rows.  Here the GEMM runs
// over the taps instead:   E[px', tap*P + co] = sum_ci h2[px', ci] * W[tap][co][ci]     (one GEMM, K = 64, N = 25 P)
// for every pixel px' of a 24 x 16 input halo (3 M-blocks of 128 TMEM lanes), and the epilogue gathers
//                          out[co][p] = sum_tap E[p + tap, tap*P + co]                  (col2im through shared memory).
// FLOPs equal the direct convolution (no padding of P), the MMA work per tile is 36 instructions.
// Precision: h2 arrives as fp16 hi + fp16 residual (16 chunk8 planes), the weights as fp16 hi + fp16 residual; the three
// significant products hi*hi, hi*lo, lo*hi accumulate in the same fp32 TMEM columns (K = 192) => fp32-grade result.
#include "common.cuh"
#include "kernels.h"

namespace os2d {
namespace conv3s {

constexpr int THREADS = 128 + 384;       // control warps 0..3, epilogue warps 4..15 (one thread per halo pixel)
constexpr int HX = 24, HY = 16;          // input halo box
constexpr int OX = HX - 4, OY = HY - 4;  // output tile 20 x 12
constexpr int HPIX = HX * HY;            // 384 = 3 * 128
constexpr int IN_CHUNKS = 16;            // 8 hi + 8 lo chunk8 planes of h2
constexpr int KSTEPS = IN_CHUNKS / 2;      // 8 K steps of 16 channels per tile: 0..3 = h2 value, 4..7 = h2 residual
constexpr int STAGES = 10;                // ring of K-step stages: the loads of tile t+1 stream in under the MMAs and the
                                          // epilogue of tile t
constexpr uint32_t STAGE_BYTES = 2 * HPIX * 16;       // 12288: [2 chunk8][HY][HX][8] fp16
constexpr uint32_t A_BYTES = STAGES * STAGE_BYTES;    // 122880
constexpr int NPAD_MAX = 160;
constexpr uint32_t B_BYTES_MAX = 16 * NPAD_MAX * 16;  // w_hi (8 chunks) + w_lo (8 chunks)
constexpr int RS_MAX = 31;                            // staging row stride (floats), odd => conflict free
constexpr uint32_t S_BYTES = HPIX * RS_MAX * 4;
constexpr uint32_t SMEM_BYTES = A_BYTES + B_BYTES_MAX + S_BYTES + 512 + 128;

struct Params {
  int planes, H, W, P;
  int NPAD;            // round_up(25 P, 16)
  int TX, TY, total_tiles;
  const uint8_t* wblob;   // [16 chunks][NPAD][8] fp16
  const float* inv_scale; // device scalar 1 / s3
  const float* bias;      // P floats
  float* out;             // [planes][P][H*W]
};

__global__ void __launch_bounds__(THREADS, 1) conv3s_kernel(const __grid_constant__ CUtensorMap map_in, Params P) {
  extern __shared__ uint8_t smem_raw[];
  // aligned by offsetting the __shared__ pointer itself (integer round-trips of the pointer lose the address space and
  // turn every staging access into a generic LD/ST)
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint8_t* sa = smem;
  uint8_t* sb = sa + A_BYTES;
  float* S = reinterpret_cast<float*>(sb + B_BYTES_MAX);
  uint64_t* afull = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(S) + S_BYTES);   // [STAGES]
  uint64_t* aempty = afull + STAGES;                                                          // [STAGES]
  uint64_t* tfull = aempty + STAGES;
  uint64_t* tempty = tfull + 1;
  uint64_t* wfull = tempty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) tma_prefetch_desc(&map_in);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(afull + i, 1); mbar_init(aempty + i, 1); }
    mbar_init(tfull, 1); mbar_init(tempty, 12); mbar_init(wfull, 1);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_plane = P.TX * P.TY;
  const uint32_t b_bytes = 16u * P.NPAD * 16u;
  pdl_wait();          // h2 is the previous kernel's output

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wfull, b_bytes);
      bulk_load_1d(sb, P.wblob, b_bytes, wfull);
      uint32_t st = 0, ph = 0;
      for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
        const int plane = t / tiles_per_plane, rem = t - plane * tiles_per_plane;
        const int ty = rem / P.TX, tx = rem - ty * P.TX;
        for (int ks = 0; ks < KSTEPS; ++ks) {             // chunk pair 2 ks (chunks 0..7 value, 8..15 residual)
          mbar_wait(aempty + st, ph ^ 1);
          mbar_expect_tx(afull + st, STAGE_BYTES);
          tma_load_4d(sa + st * STAGE_BYTES, &map_in, afull + st, 8 * (tx * OX - 2), ty * OY - 2, 2 * ks, plane);
          if (++st == STAGES) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // whole warp stays converged, one elected lane issues (a divergent `lane == 0` issue path costs ~10 SASS instr / MMA)
    const uint32_t idesc = umma_idesc_f16(128, P.NPAD);
    const uint32_t a0 = smem_u32(sa), b0 = smem_u32(sb);
    const uint32_t a_lbo = HPIX * 16, b_lbo = P.NPAD * 16;
    mbar_wait(wfull, 0);
    uint32_t st = 0, ph = 0, tph = 0;
    for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
      mbar_wait(tempty, tph ^ 1);
      for (int ks = 0; ks < KSTEPS; ++ks) {
        mbar_wait(afull + st, ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_st = a0 + st * STAGE_BYTES;
          const int kk = ks & 3;
          const uint64_t db_hi = umma_smem_desc(b0 + kk * 2 * b_lbo, b_lbo, 128, 0);
          const uint64_t db_lo = umma_smem_desc(b0 + (8 + kk * 2) * b_lbo, b_lbo, 128, 0);
#pragma unroll
          for (int mb = 0; mb < 3; ++mb) {
            const uint64_t da = umma_smem_desc(a_st + mb * 128 * 16, a_lbo, 128, 0);
            if (ks < 4) {                                  // h2 value x (w hi, w lo)
              umma_f16(tmem_base + mb * NPAD_MAX, da, db_hi, idesc, ks != 0);
              umma_f16(tmem_base + mb * NPAD_MAX, da, db_lo, idesc, 1u);
            } else {                                       // h2 residual x w hi
              umma_f16(tmem_base + mb * NPAD_MAX, da, db_hi, idesc, 1u);
            }
          }
          umma_commit(aempty + st);
        }
        __syncwarp();
        if (++st == STAGES) { st = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(tfull);
      __syncwarp();
      tph ^= 1;
    }
  } else if (warp >= 4) {
    const int we = warp - 4;
    const int q = warp & 3, mb = we >> 2;
    const int hp = mb * 128 + q * 32 + lane;           // halo pixel owned for the TMEM -> shared staging
    const int o = threadIdx.x - 128;                   // output pixel owned for the gather (o < 240)
    const int oy = o / OX, ox = o - oy * OX;
    const int PP = P.P, G = 5 * PP, RS = G + 1;        // columns per kernel row, staging row stride
    const size_t HW = static_cast<size_t>(P.H) * P.W;
    const float inv_scale = P.inv_scale[0];
    float bias[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) bias[c] = (c < PP) ? P.bias[c] : 0.f;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
      const int plane = t / tiles_per_plane, rem = t - plane * tiles_per_plane;
      const int ty = rem / P.TX, tx = rem - ty * P.TX;
      float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      mbar_wait(tfull, ph);
      tc_fence_after();
      const uint32_t tbase = tmem_base + mb * NPAD_MAX + (static_cast<uint32_t>(q * 32) << 16);
      for (int dy = 0; dy < 5; ++dy) {
        uint32_t r[32];
        tmem_ld32(tbase + dy * G, r);
        tmem_ld_wait();
        if (dy == 4) {                                   // last TMEM read of the tile: release the accumulators
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty);
        }
        float* srow = S + hp * RS;
#pragma unroll
        for (int j = 0; j < 30; ++j) if (j < G) srow[j] = __uint_as_float(r[j]);
        asm volatile("bar.sync 1, 384;" ::: "memory");
        if (o < OX * OY) {
          const float* g = S + ((oy + dy) * HX + ox) * RS;
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) {
#pragma unroll
            for (int c = 0; c < 6; ++c) if (c < PP) acc[c] += g[dx * RS + dx * PP + c];
          }
        }
        asm volatile("bar.sync 1, 384;" ::: "memory");
      }
      const int y = ty * OY + oy, x = tx * OX + ox;
      if (o < OX * OY && y < P.H && x < P.W) {
        float* dst = P.out + static_cast<size_t>(plane) * PP * HW + static_cast<size_t>(y) * P.W + x;
#pragma unroll
        for (int c = 0; c < 6; ++c) if (c < PP) dst[c * HW] = fmaf(acc[c], inv_scale, bias[c]);
      }
      ph ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace conv3s

size_t conv3s_weight_blob_bytes(int P) { return static_cast<size_t>(16) * ((25 * P + 15) / 16 * 16) * 16; }

int launch_conv3s(const void* h2_vol, const void* wblob, const float* bias, const float* inv_scale, float* out, int planes,
                  int P, int H, int W, int num_sms, cudaStream_t st) {
  using namespace conv3s;
  if (planes <= 0 || H <= 0 || W <= 0 || (P != 4 && P != 6)) return kErrBadArg;
  CUtensorMap map_in;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(8) * W, static_cast<uint64_t>(H), static_cast<uint64_t>(IN_CHUNKS),
                        static_cast<uint64_t>(planes)};
    uint64_t strides[3] = {static_cast<uint64_t>(16) * W, static_cast<uint64_t>(16) * W * H,
                           static_cast<uint64_t>(16) * W * H * IN_CHUNKS};
    uint32_t box[4] = {8 * HX, HY, 2, 1};
    int rc = encode_tensor_map(&map_in, h2_vol, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc != kOk) return rc;
  }
  Params Pm;
  Pm.planes = planes; Pm.H = H; Pm.W = W; Pm.P = P;
  Pm.NPAD = (25 * P + 15) / 16 * 16;
  Pm.TX = (W + OX - 1) / OX;
  Pm.TY = (H + OY - 1) / OY;
  Pm.total_tiles = planes * Pm.TX * Pm.TY;
  Pm.wblob = reinterpret_cast<const uint8_t*>(wblob);
  Pm.inv_scale = inv_scale;
  Pm.bias = bias;
  Pm.out = out;
  OS2D_SET_MAX_DYN_SMEM(conv3s_kernel, SMEM_BYTES);
  const int grid = Pm.total_tiles < num_sms ? Pm.total_tiles : num_sms;
  OS2D_CUDA_TRY(launch_pdl(conv3s_kernel, dim3(grid), dim3(THREADS), SMEM_BYTES, st, 1, map_in, Pm));
  os2d::note_launch();
  return kOk;
}

}  // namespace os2d

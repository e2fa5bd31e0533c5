// K1: dense correlation as a tcgen05 GEMM with a fused ReLU / L2-norm / centering epilogue.
//
//   corr[b*C+c, k, p] = sum_d F[b, p, d] * Cf[c, k, d]        (reference: os2d/modeling/head.py:342-350)
//   z = relu(corr) / (||relu(corr)||_225 + 1e-6)              (reference: head.py:650, 597-601)
//
// One CTA tile = 128 image locations (MMA M, one TMEM lane per location) x one class's 240 padded correlation
// channels (MMA N), K = D in 64-element TMA boxes (128B swizzle) through an mbarrier ring.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (one thread), 2 = TMEM allocator, 4..19 = epilogue: 4 lane
// quadrants x 4 column groups; a thread owns one location and 64 channels, the 225-channel sums are combined over
// the 4 column groups through shared memory (one named barrier per tile).  The accumulator is double buffered in
// TMEM (2 x 256 columns) so the epilogue of tile t overlaps the MMAs of tile t+1.
//
// Two variants of the main loop (template kTwoCta):
//   * 1-CTA: tcgen05.mma.cta_group::1, every CTA streams its A tile (16 KB) and the whole class tile (30 KB) per K block:
//     96 B per MMA clock and SM; with 184 KB of ring the kernel is bounded by bytes in flight (measured 0.70 of peak).
//   * 2-CTA (default): a cluster of two CTAs works on two neighbouring location tiles of the SAME class with
//     tcgen05.mma.cta_group::2 (M = 256 over the SM pair): each CTA loads its own A tile and only HALF of the class
//     tile (15 KB), 31 KB per stage => 6 stages, 65 B per MMA clock and SM.  The leader CTA issues the MMAs; TMA
//     transactions of both CTAs complete on the leader's barrier; tcgen05.commit multicasts to both CTAs.
//
// Outputs (never the fp32 [C,225,H,W] volume of the reference):
//   zvol  fp16 [plane][30 chunks][N][8 ch]  = (z - mean_k z) * 64 for k < 225, DC side channels
//         225/226 = fp16(8*mean), 227 = fp16 residual of 8*mean, rest 0   (conv1 B operand)
//   rawvol fp16 [plane][225][N]             = corr                        (sampler input)
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace os2d {
namespace corr {

constexpr int BM = 128, BN = kCorrPad, BK = 64;
constexpr uint32_t A_BYTES = BM * BK * 2;            // 16384
constexpr int THREADS = 128 + 512;                   // 4 control warps + 16 epilogue warps
constexpr uint32_t TMEM_COLS = 512, ACC_COLS = 256;
constexpr uint32_t TAIL_BYTES = 1024 /*align*/ + 256 /*barriers*/ + 2 * 4 * 128 * 8 /*partial sums*/;

template <bool kTwoCta>
struct Cfg {
  static constexpr int STAGES = kTwoCta ? 6 : 4;
  static constexpr int B_ROWS = kTwoCta ? BN / 2 : BN;             // class rows held by one CTA
  static constexpr uint32_t B_BYTES = B_ROWS * BK * 2;             // 15360 / 30720
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;       // 31744 / 47104 (multiples of 1024)
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + TAIL_BYTES;
};

struct Params {
  int B, C, D, N, MT;       // MT = m-tiles per plane
  int total_tiles;          // 1-CTA: planes * MT tiles; 2-CTA: planes * ceil(MT / 2) tile pairs
  __half* zvol;
  __half* rawvol;
  unsigned int* plane_done;   // optional: += 1 per finished (CTA, tile) of a plane, released at gpu scope (stage-concurrent conv1)
};

// ---- cluster helpers (2-CTA variant) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a local shared-memory pointer) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(const void* p, uint32_t rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(smem_u32(p)), "r"(rank));
  return out;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-pair bit of a shared address => the leader's copy
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, uint64_t* leader_bar, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

template <bool kTwoCta>
__global__ void __launch_bounds__(THREADS, 1)
corr_kernel(const __grid_constant__ CUtensorMap map_img, const __grid_constant__ CUtensorMap map_cls, Params P) {
  using K = Cfg<kTwoCta>;
  constexpr int STAGES = K::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * K::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint32_t* tile_cnt = tmem_slot + 2;      // [2]: epilogue warps that have finished the tile in accumulator stage 0 / 1
  float2* part = reinterpret_cast<float2*>(smem + STAGES * K::STAGE_BYTES + 256);   // [2 acc stages][4 col groups][128 rows]

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = P.D / BK;
  const uint32_t rank = kTwoCta ? cluster_ctarank() : 0u;      // 0 = leader of the CTA pair
  const int unit = kTwoCta ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);   // work-list walker
  const int nunits = kTwoCta ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int MTU = kTwoCta ? (P.MT + 1) / 2 : P.MT;             // work items per plane

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_img);
    tma_prefetch_desc(&map_cls);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kTwoCta ? 32 : 16); }
    tile_cnt[0] = 0u; tile_cnt[1] = 0u;
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (kTwoCta) {
      tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kTwoCta) cluster_sync_all();     // barriers of both CTAs initialised before any remote arrive / TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();          // operands / output volumes may still be in use by the previous kernel of the stream

  if (warp == 0) {
    // ------------------------------ TMA producer (every CTA: its A tile + its part of the class tile) ----------
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int u = unit; u < P.total_tiles; u += nunits) {
        const int plane = u / MTU, mu = u - plane * MTU;
        const int mt = kTwoCta ? 2 * mu + static_cast<int>(rank) : mu;
        const int b = plane / P.C, c = plane - b * P.C;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * K::STAGE_BYTES;
          if constexpr (kTwoCta) {
            if (rank == 0) mbar_expect_tx(&full[stage], 2 * K::STAGE_BYTES);   // bytes of both CTAs land on the leader
            tma_load_3d_2sm(sa, &map_img, &full[stage], kb * BK, mt * BM, b);
            tma_load_3d_2sm(sa + A_BYTES, &map_cls, &full[stage], kb * BK, static_cast<int>(rank) * K::B_ROWS, c);
          } else {
            mbar_expect_tx(&full[stage], K::STAGE_BYTES);
            tma_load_3d(sa, &map_img, &full[stage], kb * BK, mt * BM, b);
            tma_load_3d(sa + A_BYTES, &map_cls, &full[stage], kb * BK, 0, c);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only in the 2-CTA variant) ------------------------------
    // The warp stays converged (warp-uniform control flow); one elected lane issues the tcgen05 instructions - issuing from
    // a divergent `lane == 0` branch makes ptxas wrap every UMMA in a uniform-register election loop.
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kTwoCta ? 2 * BM : BM, BN);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      for (int u = unit; u < P.total_tiles; u += nunits) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * ACC_COLS;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          // 128B-swizzled K-major tiles: 8-row groups 1024 B apart, K advance = 32 B (+2 in the address field) inside the atom
          const uint64_t da0 = umma_smem_desc(smem_u32(smem + stage * K::STAGE_BYTES), 16, 1024, 2);
          const uint64_t db0 = umma_smem_desc(smem_u32(smem + stage * K::STAGE_BYTES) + A_BYTES, 16, 1024, 2);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint32_t acc = (k == 0) ? (kb != 0 ? 1u : 0u) : 1u;
              if constexpr (kTwoCta) umma_f16_2sm(d_tmem, da0 + 2 * k, db0 + 2 * k, idesc, acc);
              else umma_f16(d_tmem, da0 + 2 * k, db0 + 2 * k, idesc, acc);
            }
            if constexpr (kTwoCta) umma_commit_2sm(&empty[stage]); else umma_commit(&empty[stage]);
            if (kb == KB - 1) {
              if constexpr (kTwoCta) umma_commit_2sm(&tfull[as]); else umma_commit(&tfull[as]);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
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
    uint32_t tempty_remote[2] = {0u, 0u};
    if constexpr (kTwoCta) {
      tempty_remote[0] = map_to_rank(&tempty[0], 0);
      tempty_remote[1] = map_to_rank(&tempty[1], 0);
    }
    for (int u = unit; u < P.total_tiles; u += nunits) {
      const int plane = u / MTU, mu = u - plane * MTU;
      const int mt = kTwoCta ? 2 * mu + static_cast<int>(rank) : mu;
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
      if (P.plane_done != nullptr) __threadfence();     // this thread's z / raw stores are visible at gpu scope
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (P.plane_done != nullptr) {
          // the 16th epilogue warp of this CTA to finish the tile publishes it (fence + relaxed RMW = release at gpu scope,
          // cumulative over the other warps' fenced stores through the shared-memory counter)
          if (atomicAdd(&tile_cnt[as], 1u) == 15u) {
            tile_cnt[as] = 0u;
            __threadfence();
            atomicAdd(P.plane_done + plane, 1u);
          }
        }
        if constexpr (kTwoCta) mbar_arrive_remote(tempty_remote[as]);   // the leader's barrier collects both CTAs
        else mbar_arrive(&tempty[as]);
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (kTwoCta) {
    cluster_sync_all();          // nobody frees TMEM / exits while the peer may still signal or read
    if (warp == 2) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  } else {
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <bool kTwoCta>
static int launch_t(const void* img_packed, const void* cls_packed, int B, int C, int D, int N, void* zvol, void* rawvol,
                    int num_sms, unsigned int* plane_done, unsigned int* signals_per_plane, cudaStream_t st) {
  using K = Cfg<kTwoCta>;
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
    uint32_t box[3] = {BK, static_cast<uint32_t>(K::B_ROWS), 1};
    int rc = encode_tensor_map(&map_cls, cls_packed, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc != kOk) return rc;
  }
  Params P;
  P.B = B; P.C = C; P.D = D; P.N = N;
  P.MT = (N + BM - 1) / BM;
  P.total_tiles = B * C * (kTwoCta ? (P.MT + 1) / 2 : P.MT);
  P.zvol = reinterpret_cast<__half*>(zvol);
  P.rawvol = reinterpret_cast<__half*>(rawvol);
  P.plane_done = plane_done;
  if (signals_per_plane) *signals_per_plane = static_cast<unsigned int>(kTwoCta ? 2 * ((P.MT + 1) / 2) : P.MT);
  OS2D_SET_MAX_DYN_SMEM(corr_kernel<kTwoCta>, K::SMEM_BYTES);
  unsigned grid;
  if (kTwoCta) {
    int pairs = num_sms / 2;
    if (pairs > P.total_tiles) pairs = P.total_tiles;
    grid = static_cast<unsigned>(2 * pairs);
  } else {
    grid = static_cast<unsigned>(P.total_tiles < num_sms ? P.total_tiles : num_sms);
  }
  OS2D_CUDA_TRY(launch_pdl(corr_kernel<kTwoCta>, dim3(grid), dim3(THREADS), K::SMEM_BYTES, st, kTwoCta ? 2 : 1, map_img, map_cls, P));
  os2d::note_launch();
  return kOk;
}

}  // namespace corr

int launch_corr(const void* img_packed, const void* cls_packed, int B, int C, int D, int H, int W, void* zvol,
                void* rawvol, int num_sms, cudaStream_t st, unsigned int* plane_done, unsigned int* signals_per_plane) {
  using namespace corr;
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || D % BK != 0) return kErrBadArg;
  const int N = H * W;
  // variant selection: 2-CTA unless OS2D_B200_CORR_1CTA is set (A/B switch; both are tested)
  static const bool one_cta = getenv("OS2D_B200_CORR_1CTA") != nullptr;
  if (one_cta || num_sms < 2)
    return launch_t<false>(img_packed, cls_packed, B, C, D, N, zvol, rawvol, num_sms, plane_done, signals_per_plane, st);
  return launch_t<true>(img_packed, cls_packed, B, C, D, N, zvol, rawvol, num_sms, plane_done, signals_per_plane, st);
}

}  // namespace os2d

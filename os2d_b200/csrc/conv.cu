// K2: the TransformNet convolutions as tcgen05 implicit GEMMs (reference: os2d/modeling/head.py:604-655,
// conv 225->128 k7 p3 + BN + ReLU, conv 128->64 k5 p2 + BN + ReLU, conv 64->P k5 p2).
//
// Formulation:  D[cout (M = 128 TMEM lanes), pixel (N <= 256 columns)] += W[tap][cout, ci] * X[pixel + tap, ci]
//   * A operand (weights): packed on the host into the exact shared-memory image
//     [sub-chunk of 16 ci][dy][dx][2 x 8-channel K groups][128 rows][8 ci] fp16, streamed with 1-D bulk copies
//     (one kernel row = KS taps per ring stage).  No-swizzle K-major core matrices (8 rows x 16 B).
//   * B operand (activations): volume layout [plane][chunk8][H][W][8 ci] fp16.  One TMA box per 16-channel
//     sub-chunk brings the halo (16 + 2 pad) x (32 + 2 pad) pixels of a 2-strip tile into shared memory as
//     [2][rows][cols][16 B]; zero padding of the convolution = TMA out-of-bounds fill.  Every tap is then a
//     *shifted view* of the same halo: the UMMA descriptor start address moves by (dy * cols + dx) * 16 B,
//     SBO = cols * 16 B (one image row of an 8-pixel-wide strip per 8-row core-matrix group).
//   * a CTA tile = 2 strips (8 px wide, up to 32 rows) => two N<=256 accumulators = all 512 TMEM columns;
//     every weight byte fetched from L2 is used for 512 pixels.
// Epilogue modes: see ConvLayerDesc in kernels.h (BN folded to fp32 alpha/beta, hi/lo weight rows combined).
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace os2d {
namespace conv {

constexpr int THREADS = 128 + 256;   // 4 control warps + 8 epilogue warps (one group of 4 per strip)
constexpr int NUM_H = 2;   // halo ring depth
constexpr int NUM_W = 4;   // weight ring depth
constexpr uint32_t TAP_BYTES = 2 * 128 * 16;   // one tap, 16 input channels: [2][128 rows][16 B]
constexpr uint32_t TRANS_STRIDE = 32 * 16 + 16;  // per-warp transpose staging: chunk stride (bank skew)
constexpr uint32_t TRANS_BYTES = 4 * TRANS_STRIDE;  // per warp

template <int KS>
struct Geo {
  static constexpr int PAD = KS / 2;
  static constexpr int HWX = kStripW * kStripsPerTile + 2 * PAD;   // halo columns
  static constexpr int HWY = kMaxTileRows + 2 * PAD;               // halo rows (box height)
  static constexpr uint32_t HALO_BYTES = 2u * HWY * HWX * 16u;
  static constexpr uint32_t WSTAGE_BYTES = KS * TAP_BYTES;
  static constexpr uint32_t SMEM_BYTES = NUM_H * HALO_BYTES + NUM_W * WSTAGE_BYTES + 8 * TRANS_BYTES +
                                         256 /*barriers*/ + 128 /*align*/;
};

struct Params {
  int planes, H, W;
  int T;              // rows per tile (even, <= 32)
  int TY, TXP;        // tiles along y, strip pairs along x
  int total_tiles;
  int nsub;           // input channels / 16
  int mode;
  int out_real;
  float lo_scale;
  const uint8_t* wblob;
  const float* alpha;
  const float* beta;
  void* out;
  int n_double;       // tiles [0, n_double) are 2-strip tiles; the rest are the last partial wave split in 1-strip tiles
  const unsigned int* wait_flags;   // optional: per-plane completion counters of the kernel that is WRITING the input volume
  unsigned int wait_target;         //           concurrently on other SMs (stage-concurrent K1 -> conv1); plane ready at >= target
};

struct TileInfo { int plane, y0, x0, rows, nstrips; };

// Work list: 2-strip tiles (plane, row tile, strip pair) in order; when the last wave of the persistent grid would be
// less than half full, its 2-strip tiles are split into 1-strip tiles so that the tail costs half a tile per CTA.
__device__ __forceinline__ TileInfo decode_tile(const Params& P, int t) {
  int d, s0, ns;
  if (t < P.n_double) { d = t; s0 = 0; ns = 2; }
  else { const int u = t - P.n_double; d = P.n_double + (u >> 1); s0 = u & 1; ns = 1; }
  const int tiles_per_plane = P.TY * P.TXP;
  TileInfo ti;
  ti.plane = d / tiles_per_plane;
  const int rem = d - ti.plane * tiles_per_plane;
  const int ty = rem / P.TXP, txp = rem - ty * P.TXP;
  ti.y0 = ty * P.T;
  ti.x0 = (kStripsPerTile * txp + s0) * kStripW;
  ti.rows = min(P.T, P.H - ti.y0);
  const int strips_left = (P.W - ti.x0 + kStripW - 1) / kStripW;
  ti.nstrips = max(0, min(ns, strips_left));
  return ti;
}

// (Negative result, kept out of the code: staging the weight tile of a tap in TMEM with tcgen05.cp and running the MMAs in
// TS mode is correct but slower - 1.19 vs 1.12 ms for conv1 - because the copy serialises with the MMAs in the tensor pipe.)

template <int KS>
__global__ void __launch_bounds__(THREADS, 1) conv_kernel(const __grid_constant__ CUtensorMap map_in, Params P) {
  using G = Geo<KS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* halo = smem;
  uint8_t* wst = halo + NUM_H * G::HALO_BYTES;
  uint8_t* trans = wst + NUM_W * G::WSTAGE_BYTES;
  uint64_t* hfull = reinterpret_cast<uint64_t*>(trans + 8 * TRANS_BYTES);
  uint64_t* hempty = hfull + NUM_H;
  uint64_t* wfull = hempty + NUM_H;
  uint64_t* wempty = wfull + NUM_W;
  uint64_t* tfull = wempty + NUM_W;
  uint64_t* tempty = tfull + 2;            // one full/empty pair per strip accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&map_in);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NUM_H; ++i) { mbar_init(&hfull[i], 1); mbar_init(&hempty[i], 1); }
    for (int i = 0; i < NUM_W; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();          // the input volume is the previous kernel's output

  // Strip staggering: with 2 strips the first and the last 16-channel sub-chunk are issued strip by strip (their weight rows
  // are streamed twice), every other sub-chunk for both strips per weight load.  Strip 0's accumulator is therefore complete
  // KS*KS MMAs before strip 1's, and each strip has its own TMEM full/empty barrier pair: the epilogue of strip 0 overlaps the
  // last MMAs of strip 1, and the epilogue of strip 1 overlaps the first MMAs (strip 0) of the next tile.
  if (warp == 0) {
    // ------------------------------ producer: halo boxes (TMA) + weight rows (bulk) ------------------------------
    if (lane == 0) {
      int hs = 0; uint32_t hph = 0;
      int ws = 0; uint32_t wph = 0;
      int ready_plane = -1;
      for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
        const TileInfo ti = decode_tile(P, t);
        if (ti.nstrips == 0) continue;
        const int plane = ti.plane, y0 = ti.y0, x0 = ti.x0;
        const bool stag = ti.nstrips == 2 && P.nsub >= 2;
        if (P.wait_flags != nullptr && plane > ready_plane) {
          // the producing kernel runs concurrently: acquire the plane's completion count, then order the generic-proxy
          // acquire before the async-proxy (TMA) reads.  Bounded wait (~4 s) so that a scheduling accident cannot hang the GPU.
          const unsigned int* f = P.wait_flags + plane;
          const long long t0 = clock64();
          while (ld_acquire_gpu_u32(f) < P.wait_target) {
            __nanosleep(200);
            if (clock64() - t0 > (1ll << 33)) break;
          }
          asm volatile("fence.proxy.async;" ::: "memory");
          ready_plane = plane;
        }
        for (int sc = 0; sc < P.nsub; ++sc) {
          mbar_wait(&hempty[hs], hph ^ 1);
          mbar_expect_tx(&hfull[hs], G::HALO_BYTES);
          tma_load_4d(halo + hs * G::HALO_BYTES, &map_in, &hfull[hs], 8 * (x0 - G::PAD), y0 - G::PAD, 2 * sc, plane);
          if (++hs == NUM_H) { hs = 0; hph ^= 1; }
          const int reps = (stag && (sc == 0 || sc == P.nsub - 1)) ? 2 : 1;
          for (int rep = 0; rep < reps; ++rep) {
            for (int dy = 0; dy < KS; ++dy) {
              mbar_wait(&wempty[ws], wph ^ 1);
              mbar_expect_tx(&wfull[ws], G::WSTAGE_BYTES);
              bulk_load_1d(wst + ws * G::WSTAGE_BYTES, P.wblob + static_cast<size_t>(sc * KS + dy) * G::WSTAGE_BYTES,
                           G::WSTAGE_BYTES, &wfull[ws]);
              if (++ws == NUM_W) { ws = 0; wph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    // The whole warp runs the (warp-uniform) control flow converged; one elected lane issues the tcgen05 instructions.
    // (Issuing from a divergent `lane == 0` branch makes ptxas wrap every UMMA in a uniform-register election loop.)
    {
      int hs = 0; uint32_t hph = 0;
      int ws = 0; uint32_t wph = 0;
      uint32_t tph[2] = {0u, 0u};
      for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
        const TileInfo ti = decode_tile(P, t);
        if (ti.nstrips == 0) continue;
        const int rows = (ti.rows + 1) & ~1;
        const int nstrips = ti.nstrips;
        const uint32_t idesc = umma_idesc_f16(128, rows * kStripW);
        const bool stag = nstrips == 2 && P.nsub >= 2;
        uint64_t db0 = 0;

        // One kernel row of weights per ring stage; strips [S0, S1) (compile time) of the resident halo; `fresh`: first
        // sub-chunk.  The issuing thread is the critical resource of this kernel (one UMMA every ~112 clocks): descriptors
        // are formed by adding small constants to per-stage bases (the 14-bit address fields cannot carry for shared
        // memory addresses < 256 KB), everything else is unrolled.
        auto run_rows = [&](auto s0c, auto s1c, bool fresh) {
          constexpr int S0 = decltype(s0c)::value, S1 = decltype(s1c)::value;
          for (int dy = 0; dy < KS; ++dy) {
            mbar_wait(&wfull[ws], wph);
            tc_fence_after();
            const uint64_t da0 = umma_smem_desc(smem_u32(wst + ws * G::WSTAGE_BYTES), /*LBO (K group)*/ 128 * 16, /*SBO*/ 128, 0);
            const uint64_t dbr = db0 + static_cast<uint64_t>(dy * G::HWX);
            const uint32_t acc0 = (fresh && dy == 0) ? 0u : 1u;
            if (elect_one()) {
#pragma unroll
            for (int dx = 0; dx < KS; ++dx) {
              const uint64_t da = da0 + static_cast<uint64_t>(dx * (TAP_BYTES >> 4));
              const uint32_t acc = (dx == 0) ? acc0 : 1u;
#pragma unroll
              for (int s = S0; s < S1; ++s) {
                const uint64_t db = dbr + static_cast<uint64_t>(dx + kStripW * s);
                umma_f16(tmem_base + s * 256, da, db, idesc, acc);
              }
            }
            umma_commit(&wempty[ws]);
            }
            __syncwarp();
            if (++ws == NUM_W) { ws = 0; wph ^= 1; }
          }
        };
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        using I2 = std::integral_constant<int, 2>;
        auto wait_halo = [&]() {
          mbar_wait(&hfull[hs], hph);
          tc_fence_after();
          db0 = umma_smem_desc(smem_u32(halo + hs * G::HALO_BYTES), /*LBO*/ G::HWY * G::HWX * 16, /*SBO*/ G::HWX * 16, 0);
        };
        auto release_halo = [&]() {
          if (elect_one()) umma_commit(&hempty[hs]);
          __syncwarp();
          if (++hs == NUM_H) { hs = 0; hph ^= 1; }
        };
        auto wait_acc = [&](int s) {      // epilogue has drained strip s of the previous tile
          mbar_wait(&tempty[s], tph[s] ^ 1);
          tc_fence_after();
        };

        if (!stag) {
          for (int s = 0; s < nstrips; ++s) wait_acc(s);
          for (int sc = 0; sc < P.nsub; ++sc) {
            wait_halo();
            if (nstrips == 2) run_rows(I0{}, I2{}, sc == 0); else run_rows(I0{}, I1{}, sc == 0);
            release_halo();
          }
          for (int s = 0; s < nstrips; ++s) { if (elect_one()) umma_commit(&tfull[s]); __syncwarp(); tph[s] ^= 1; }
        } else {
          wait_halo();
          wait_acc(0); run_rows(I0{}, I1{}, true);
          wait_acc(1); run_rows(I1{}, I2{}, true);
          release_halo();
          for (int sc = 1; sc + 1 < P.nsub; ++sc) {
            wait_halo();
            run_rows(I0{}, I2{}, false);
            release_halo();
          }
          wait_halo();
          run_rows(I0{}, I1{}, false);
          if (elect_one()) umma_commit(&tfull[0]);
          __syncwarp(); tph[0] ^= 1;
          run_rows(I1{}, I2{}, false);
          release_halo();
          if (elect_one()) umma_commit(&tfull[1]);
          __syncwarp(); tph[1] ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ------------------------------
    // 8 warps: e = TMEM lane quadrant (warp % 4), sg = strip group; each group of 4 warps drains one strip's accumulator
    const int e = warp & 3, sg = (warp - 4) >> 2;
    const int row = e * 32 + lane;                 // accumulator row = output channel (or hi/lo row)
    uint8_t* my_trans = trans + (warp - 4) * TRANS_BYTES;
    float alpha = 0.f, beta = 0.f;
    alpha = P.alpha[row]; beta = P.beta[row];      // 128 entries, indexed by accumulator row
    uint32_t tph = 0;
    const size_t HW = static_cast<size_t>(P.H) * P.W;
    const int out_chunks = 16;   // conv1: 128 channels; conv2: 64 hi + 64 residual
    for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
      const TileInfo ti = decode_tile(P, t);
      if (ti.nstrips == 0) continue;
      const int plane = ti.plane, y0 = ti.y0, x0 = ti.x0, rows = ti.rows, nstrips = ti.nstrips;
      const int nquads = (rows + 3) >> 2;

      if (sg >= nstrips) continue;               // this strip group has no accumulator in a 1-strip tile
      mbar_wait(&tfull[sg], tph);
      tc_fence_after();
      for (int s = sg; s < nstrips; s += 2) {
        for (int rq = 0; rq < nquads; ++rq) {
          uint32_t r[32];
          tmem_ld32(tmem_base + s * 256 + rq * 32 + (static_cast<uint32_t>(e * 32) << 16), r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);

          // pixel j of this block: image row y0 + 4*rq + j/8, column x0 + 8*s + j%8
          const int y = y0 + 4 * rq + (lane >> 3), x = x0 + kStripW * s + (lane & 7);
          const bool ok = (4 * rq + (lane >> 3)) < rows && x < P.W;
          if (P.mode == 0) {
            // conv1: BN + ReLU -> fp16, transpose 32 channels x 32 pixels through shared memory
            const uint32_t cl = lane >> 3, pos = lane & 7;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const __half hv = __float2half(fmaxf(fmaf(alpha, v[j], beta), 0.f));
              *reinterpret_cast<__half*>(my_trans + cl * TRANS_STRIDE + j * 16 + pos * 2) = hv;
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 val = *reinterpret_cast<const uint4*>(my_trans + q * TRANS_STRIDE + lane * 16);
              if (ok) {
                __half* o = reinterpret_cast<__half*>(P.out) +
                            ((static_cast<size_t>(plane) * out_chunks + (e * 4 + q)) * HW + static_cast<size_t>(y) * P.W + x) * 8;
                *reinterpret_cast<uint4*>(o) = val;
              }
            }
            __syncwarp();
          } else {
            // conv2: row 32e + l holds output channel 16e + (l & 15), weight part l >> 4 (0 = fp16 hi, 1 = residual * 2048).
            // Lanes l and l ^ 16 exchange halves: l < 16 finishes pixels 0..15, l >= 16 pixels 16..31 of its channel.
            const bool upper = lane >= 16;
            float w[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float other = __shfl_xor_sync(0xffffffffu, upper ? v[j] : v[j + 16], 16);
              const float hi = upper ? other : v[j];
              const float lo = upper ? v[j + 16] : other;
              w[j] = fmaxf(fmaf(alpha, fmaf(lo, P.lo_scale, hi), beta), 0.f);
            }
            // output: fp16 value plane (chunks 0..7) + fp16 residual plane (chunks 8..15), 2 chunk8 per warp
            const uint32_t cl = (lane & 15) >> 3, pos = lane & 7, pbase = upper ? 16u : 0u;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const __half hi = __float2half(w[j]);
                const __half hv = (pass == 0) ? hi : __float2half(w[j] - __half2float(hi));
                *reinterpret_cast<__half*>(my_trans + cl * TRANS_STRIDE + (pbase + j) * 16 + pos * 2) = hv;
              }
              __syncwarp();
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const uint4 val = *reinterpret_cast<const uint4*>(my_trans + q * TRANS_STRIDE + lane * 16);
                if (ok) {
                  __half* o = reinterpret_cast<__half*>(P.out) +
                              ((static_cast<size_t>(plane) * out_chunks + (pass * 8 + e * 2 + q)) * HW + static_cast<size_t>(y) * P.W + x) * 8;
                  *reinterpret_cast<uint4*>(o) = val;
                }
              }
              __syncwarp();
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[sg]);
      tph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int KS>
static int launch_t(const ConvLayerDesc& L, const void* in_vol, const void* wblob, const float* alpha,
                    const float* beta, void* out, int planes, int H, int W, int num_sms, cudaStream_t st,
                    const unsigned int* wait_flags, unsigned int wait_target) {
  using G = Geo<KS>;
  const int in_chunks8 = L.in_chunks16 * 2;
  CUtensorMap map_in;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(8) * W, static_cast<uint64_t>(H), static_cast<uint64_t>(in_chunks8),
                        static_cast<uint64_t>(planes)};
    uint64_t strides[3] = {static_cast<uint64_t>(16) * W, static_cast<uint64_t>(16) * W * H,
                           static_cast<uint64_t>(16) * W * H * in_chunks8};
    uint32_t box[4] = {static_cast<uint32_t>(8 * G::HWX), static_cast<uint32_t>(G::HWY), 2, 1};
    int rc = encode_tensor_map(&map_in, in_vol, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc != kOk) return rc;
  }
  Params P;
  P.planes = planes; P.H = H; P.W = W;
  P.TXP = (W + kStripW * kStripsPerTile - 1) / (kStripW * kStripsPerTile);
  // Tile height: the fewest row tiles that fit 32 rows, split evenly (even T).  Measured on B200 at 80x80, C = 100:
  // T = 28 (3 row tiles) 1.24 ms vs T = 20..24 (4 row tiles, better wave balance on paper) 1.41 ms - taller tiles win
  // because the per-tile epilogue is not overlapped and N >= 224 keeps the MMA shared-memory read rate low.
  const int nty = (H + kMaxTileRows - 1) / kMaxTileRows;
  int T = (H + nty - 1) / nty;
  T = (T + 1) & ~1;
  // Small problems (fewer 2-strip tiles than SMs, e.g. 512 px x 4 classes = 8 tiles): fill the machine instead - every
  // tile becomes a 1-strip tile and the tile height shrinks (not below 8 rows) until there is a tile per SM.
  const int strips = (W + kStripW - 1) / kStripW;
  bool all_single = false;
  if (static_cast<long>(planes) * ((H + T - 1) / T) * P.TXP < num_sms) {
    all_single = true;
    while (T > 8 && static_cast<long>(planes) * ((H + T - 1) / T) * strips < num_sms) T -= 2;
  }
  if (const char* ev = getenv("OS2D_B200_CONV_TILE_ROWS")) {   // tuning override (even, 2..32)
    const int v = atoi(ev);
    if (v >= 2 && v <= kMaxTileRows && (v & 1) == 0) T = v;
  }
  P.T = T;
  P.TY = (H + T - 1) / T;
  {
    const int doubles = planes * P.TY * P.TXP;
    const int grid_sz = doubles < num_sms ? doubles : num_sms;
    const int rem = doubles % grid_sz;
    // split the last wave into 1-strip tiles when that halves its cost (it does iff 2 * rem <= grid)
    const bool split = rem > 0 && 2 * rem <= grid_sz;
    P.n_double = split ? doubles - rem : doubles;
    P.total_tiles = split ? doubles + rem : doubles;
    if (all_single) { P.n_double = 0; P.total_tiles = 2 * doubles; }
    if (getenv("OS2D_B200_CONV_NO_TAIL_SPLIT")) { P.n_double = doubles; P.total_tiles = doubles; }
  }
  P.nsub = L.in_chunks16;
  P.mode = L.mode;
  P.out_real = L.out_real;
  P.lo_scale = L.lo_scale;
  P.wblob = reinterpret_cast<const uint8_t*>(wblob);
  P.alpha = alpha;
  P.beta = beta;
  P.out = out;
  P.wait_flags = wait_flags;
  P.wait_target = wait_target;
  OS2D_SET_MAX_DYN_SMEM(conv_kernel<KS>, G::SMEM_BYTES);
  const int grid = P.total_tiles < num_sms ? P.total_tiles : num_sms;
  OS2D_CUDA_TRY(launch_pdl(conv_kernel<KS>, dim3(grid), dim3(THREADS), G::SMEM_BYTES, st, 1, map_in, P));
  os2d::note_launch();
  return kOk;
}

}  // namespace conv

size_t conv_weight_blob_bytes(int ksize, int in_chunks16) {
  return static_cast<size_t>(in_chunks16) * ksize * ksize * conv::TAP_BYTES;
}

int launch_conv(const ConvLayerDesc& L, const void* in_vol, const void* wblob, const float* alpha, const float* beta,
                void* out, int planes, int H, int W, int num_sms, cudaStream_t st, const unsigned int* wait_flags,
                unsigned int wait_target) {
  if (planes <= 0 || H <= 0 || W <= 0 || L.in_chunks16 <= 0 || L.mode < 0 || L.mode > 1) return kErrBadArg;
  if ((static_cast<uint64_t>(16) * W) % 16 != 0) return kErrBadArg;
  if (L.ksize == 7) return conv::launch_t<7>(L, in_vol, wblob, alpha, beta, out, planes, H, W, num_sms, st, wait_flags, wait_target);
  if (L.ksize == 5) return conv::launch_t<5>(L, in_vol, wblob, alpha, beta, out, planes, H, W, num_sms, st, wait_flags, wait_target);
  return kErrUnsupported;
}

}  // namespace os2d

// K4+K5 fused: decode + score/empty filter + per-label NMS across pyramid levels in ONE launch, one CTA per real label, and
// a second launch that writes the surviving detections.  No intermediate [C,N,4] box tensors, no ATen index / sort / scan
// kernels (the round-1 path ran ~25 of them around two small hand kernels).
//
//   reference: Os2dBoxCoder.decode_pyramid           os2d/modeling/box_coder.py:448-536
//              _nms_box_lists                        os2d/modeling/box_coder.py:424-437
//              nms (chunks of 10000, fixpoint loop)  os2d/structures/bounding_box.py:344-387
//              torchvision nms (greedy, IoU > thr, stable score order), BoxCoder.decode_single, clip_boxes_to_image
//
// Per label (CTA of 1024 threads; candidate ids are FLAT indices f = base_l + view * N_l + anchor into the pyramid):
//   1. candidates = valid anchors in the reference's concatenation order (class view, level, anchor), found by decoding
//      on the fly (ordered block compaction through ballots) -> workspace `cand`
//   2. pass loop of bounding_box.py:356-374: consecutive chunks of <= 10000 candidates; per chunk a stable
//      score-descending order (bitonic network in shared memory on 64-bit keys (monotone score bits, ~position)), boxes
//      re-decoded into shared memory in that order, greedy NMS, survivors written back in place; repeat until there was
//      at most one chunk or nothing was removed
//   3. final score-descending order (already true after a single-chunk pass; otherwise one stable sort of the survivors in
//      global memory), count; the last CTA to finish turns the per-label counts into output offsets.
// Greedy NMS in batches: warp 0 collects the next <= 32 boxes that are still alive, resolves the greedy decisions among
// them from their pairwise overlap bits, then all threads test the remaining boxes against the <= 32 boxes just kept.
// This is the serial greedy algorithm exactly (a box is examined only after every earlier kept box has been applied to it)
// with ~32x fewer block-wide barriers than one box per step.
#include <cub/block/block_radix_sort.cuh>

#include "common.cuh"
#include "decode.cuh"
#include "kernels.h"

namespace os2d {
namespace detect {

constexpr int kThreads = 1024;
constexpr int kChunk = 10000;                 // nms_max_batch of the reference (bounding_box.py:344)
constexpr size_t kBoxBytes = static_cast<size_t>(kChunk) * sizeof(float4);     // 160000 (sort keys alias the first 80000)
constexpr size_t kOrdBytes = static_cast<size_t>(kChunk) * sizeof(int);        // 40000
constexpr size_t kSuppBytes = 10016;
constexpr size_t kSmemBytes = kBoxBytes + kOrdBytes + kSuppBytes;

// per-chunk order: stable descending radix sort on the 32 monotone score bits (CUB block primitive inside this kernel;
// 10 items per thread cover a chunk of 10000), candidate positions as values - equal scores keep their list order
constexpr int kItems = (kChunk + kThreads - 1) / kThreads;                      // 10
using ChunkSort = cub::BlockRadixSort<unsigned int, kThreads, kItems, int>;
static_assert(sizeof(ChunkSort::TempStorage) <= kBoxBytes, "sort scratch aliases the box tile");

struct Split { int level; long long rem; };   // rem = view * N_l + anchor

__device__ __forceinline__ Split split_flat(const DetectArgs& A, int f) {
  int l = 0;
  while (l + 1 < A.L && static_cast<long long>(f) >= A.base[l + 1]) ++l;
  return {l, static_cast<long long>(f) - A.base[l]};
}
__device__ __forceinline__ float score_of(const DetectArgs& A, int f) {
  const Split s = split_flat(A, f);
  return A.lv[s.level].score[s.rem];
}
__device__ __forceinline__ Decoded decode_flat(const DetectArgs& A, int f) {
  const Split s = split_flat(A, f);
  const DetectLevel& lv = A.lv[s.level];
  const int view = static_cast<int>(s.rem / lv.N), a = static_cast<int>(s.rem - static_cast<long long>(view) * lv.N);
  const float* l = lv.loc + static_cast<size_t>(view) * 4 * lv.N + a;
  return decode_one(A.grid, a, lv.fm_w, l[0], l[lv.N], l[2 * static_cast<size_t>(lv.N)], l[3 * static_cast<size_t>(lv.N)],
                    lv.img_w, lv.img_h, lv.scale_x, lv.same_scale ? lv.scale_x : lv.scale_y);
}
// monotone float -> uint32 map (larger score => larger key); -0.0 ties with +0.0 like a float comparison
__device__ __forceinline__ uint32_t mono_bits(float s) {
  uint32_t b = __float_as_uint(s);
  if (s == 0.f) b = 0u;
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ unsigned long long make_key(float s, int pos) {
  return (static_cast<unsigned long long>(mono_bits(s)) << 32) | static_cast<unsigned long long>(~static_cast<uint32_t>(pos));
}
__device__ __forceinline__ int key_pos(unsigned long long k) { return static_cast<int>(~static_cast<uint32_t>(k)); }

// ordered compaction: rank of this thread's flag among the block's flags (thread order), block total
__device__ __forceinline__ int block_rank(bool flag, int* wsum, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  int r = __popc(m & ((1u << lane) - 1u));
  if (lane == 0) wsum[warp] = __popc(m);
  __syncthreads();
  if (warp == 0) {
    const int v = wsum[lane];
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    wsum[lane] = inc - v;
    if (lane == 31) wsum[32] = inc;
  }
  __syncthreads();
  r += wsum[warp];
  total = wsum[32];
  __syncthreads();
  return r;
}

// descending bitonic sorting network for arbitrary n (flip / disperse form: every comparator puts the larger key first, so
// the virtual padding keys "-inf" at positions >= n never move and need no storage)
__device__ __forceinline__ void cmpx(unsigned long long* k, int i, int l) {
  const unsigned long long a = k[i], b = k[l];
  if (a < b) { k[i] = b; k[l] = a; }
}
__device__ __forceinline__ void sort_desc(unsigned long long* k, int n) {
  if (n < 2) return;
  int p2 = 2;
  while (p2 < n) p2 <<= 1;
  const int half = p2 >> 1;
  for (int sz = 2, lg = 1; sz <= p2; sz <<= 1, ++lg) {
    const int hs = sz >> 1;
    for (int t = threadIdx.x; t < half; t += kThreads) {
      const int blk = t >> (lg - 1), off = t & (hs - 1);
      const int i = blk * sz + off, l = blk * sz + (sz - 1 - off);
      if (l < n) cmpx(k, i, l);
    }
    __syncthreads();
    for (int j = hs >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < half; t += kThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i + j;
        if (l < n) cmpx(k, i, l);
      }
      __syncthreads();
    }
  }
}

// greedy NMS of boxes sb[0..m) given in score-descending order; supp[i] = 1 for suppressed boxes.
// Per batch: (1) warp 0 collects the next <= 32 alive boxes; (2) warp j tests batch box j against the later batch boxes
// (one overlap test per lane) and publishes the result as a ballot; (3) every thread resolves the greedy decisions of the
// batch from the 32 ballots (a box is kept iff no earlier KEPT box of the batch overlaps it) and tests its own remaining
// boxes against the kept ones.  Three block barriers per batch, no serial overlap loop.
__device__ __forceinline__ void greedy_nms(const float4* sb, uint8_t* supp, int m, double thr, unsigned* ballots, int* sh,
                                           int* bidx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool nonneg = thr >= 0.0;
  int start = 0;
  while (start < m) {
    if (warp == 0) {
      int cnt = 0, pos = start;
      while (cnt < 32 && pos < m) {
        const int idx = pos + lane;
        const bool alive = idx < m && !supp[idx];
        const unsigned mask = __ballot_sync(0xffffffffu, alive);
        const int room = 32 - cnt;
        const int before = __popc(mask & ((1u << lane) - 1u));
        if (alive && before < room) bidx[cnt + before] = idx;
        const int pc = __popc(mask);
        if (pc > room) {                       // the window holds more alive boxes than the batch takes: stop after the last taken
          const unsigned sel = __ballot_sync(0xffffffffu, alive && before == room - 1);
          pos += __ffs(sel);
          cnt = 32;
        } else {
          cnt += pc;
          pos += 32;
        }
      }
      if (lane == 0) { sh[0] = cnt; sh[1] = min(pos, m); }
    }
    __syncthreads();
    const int k = sh[0], end = sh[1];
    if (warp < k) {                            // row `warp` of the batch's overlap matrix: bit i = box warp overlaps box i (i > warp)
      const float4 bj = sb[bidx[warp]];
      bool ov = false;
      if (lane > warp && lane < k) {
        const float4 bi = sb[bidx[lane]];
        ov = iou_exceeds_fast(bj, box_area(bj), bi, box_area(bi), thr, nonneg);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, ov);
      if (lane == 0) ballots[warp] = bal;
    }
    __syncthreads();
    unsigned kept = 0u, dead = 0u;
    for (int j = 0; j < k; ++j) {
      if (!((dead >> j) & 1u)) { kept |= 1u << j; dead |= ballots[j]; }
    }
    if (threadIdx.x < k && !((kept >> threadIdx.x) & 1u)) supp[bidx[threadIdx.x]] = 1;      // batch members that lost
    for (int j = end + static_cast<int>(threadIdx.x); j < m; j += kThreads) {
      if (supp[j]) continue;
      const float4 bj = sb[j];
      const float aj = box_area(bj);
      for (unsigned rest = kept; rest; rest &= rest - 1u) {
        const float4 bq = sb[bidx[__ffs(rest) - 1]];
        if (iou_exceeds_fast(bq, box_area(bq), bj, aj, thr, nonneg)) { supp[j] = 1; break; }
      }
    }
    __syncthreads();
    start = end;
  }
}

__global__ void __launch_bounds__(kThreads, 1) label_nms_kernel(const DetectArgs A) {
  extern __shared__ __align__(16) uint8_t dsm[];
  float4* sb = reinterpret_cast<float4*>(dsm);
  int* ord = reinterpret_cast<int*>(dsm + kBoxBytes);
  uint8_t* supp = dsm + kBoxBytes + kOrdBytes;
  __shared__ int wsum[33];
  __shared__ unsigned ballots[32];
  __shared__ int sh[2];
  __shared__ int bidx[32];
  __shared__ int s_last;

  const int label = blockIdx.x;
  const int tid = threadIdx.x;
  const int v0 = A.view_off[label], v1 = A.view_off[label + 1];
  const long long region = static_cast<long long>(v0) * A.sumN;
  int* cand = A.cand + region;
  int* outp = A.out_ids + region;

  // ---- 1. candidates in (class view, level, anchor) order ----
  int n = 0;
  for (int v = v0; v < v1; ++v) {
    const int c = A.view_ids[v];
    for (int l = 0; l < A.L; ++l) {
      const DetectLevel& lv = A.lv[l];
      const float* sc = lv.score + static_cast<size_t>(c) * lv.N;
      const float* lo = lv.loc + static_cast<size_t>(c) * 4 * lv.N;
      const float sy = lv.same_scale ? lv.scale_x : lv.scale_y;
      for (int a0 = 0; a0 < lv.N; a0 += kThreads) {
        const int a = a0 + tid;
        bool ok = false;
        if (a < lv.N && sc[a] > A.score_thr) {
          const Decoded d = decode_one(A.grid, a, lv.fm_w, lo[a], lo[lv.N + a], lo[2 * static_cast<size_t>(lv.N) + a],
                                       lo[3 * static_cast<size_t>(lv.N) + a], lv.img_w, lv.img_h, lv.scale_x, sy);
          ok = !d.empty;
        }
        int total;
        const int r = block_rank(ok, wsum, total);
        if (ok) cand[n + r] = static_cast<int>(A.base[l] + static_cast<long long>(c) * lv.N + a);
        n += total;
      }
    }
  }
  __syncthreads();

  // ---- 2. chunked passes to the fixpoint (bounding_box.py:356-374) ----
  bool sorted = true;
  while (true) {
    const int nb = (n + kChunk - 1) / kChunk;
    int w = 0;
    for (int s = 0; s < n; s += kChunk) {
      const int m = min(kChunk, n - s);
      {
        unsigned int keys[kItems];
        int vals[kItems];
#pragma unroll
        for (int it = 0; it < kItems; ++it) {
          const int i = tid * kItems + it;                      // blocked arrangement = list order (stability)
          keys[it] = i < m ? mono_bits(score_of(A, cand[s + i])) : 0u;      // 0 sorts behind every real score
          vals[it] = i;
        }
        ChunkSort(*reinterpret_cast<ChunkSort::TempStorage*>(dsm)).SortDescendingBlockedToStriped(keys, vals);
#pragma unroll
        for (int it = 0; it < kItems; ++it) {
          const int r = it * kThreads + tid;                    // striped arrangement: rank r
          if (r < m) ord[r] = cand[s + vals[it]];
        }
      }
      __syncthreads();
      for (int i = tid; i < m; i += kThreads) {
        sb[i] = decode_flat(A, ord[i]).box;
        supp[i] = 0;
      }
      __syncthreads();
      greedy_nms(sb, supp, m, A.iou_thr, ballots, sh, bidx);
      for (int i0 = 0; i0 < m; i0 += kThreads) {
        const int i = i0 + tid;
        const bool alive = i < m && !supp[i];
        int total;
        const int r = block_rank(alive, wsum, total);
        if (alive) cand[w + r] = ord[i];       // w + r <= s + i: the chunk itself already lives in shared memory
        w += total;
      }
      __syncthreads();
    }
    const bool unchanged = (w == n);
    n = w;
    if (nb <= 1 || unchanged) { sorted = nb <= 1; break; }
  }

  // ---- 3. final order: score descending, ties by list position (box_coder.py:431-435) ----
  if (sorted) {
    for (int i = tid; i < n; i += kThreads) outp[i] = cand[i];
  } else {
    unsigned long long* gk = A.keys + region;
    for (int i = tid; i < n; i += kThreads) gk[i] = make_key(score_of(A, cand[i]), i);
    __syncthreads();
    sort_desc(gk, n);
    for (int i = tid; i < n; i += kThreads) outp[i] = cand[key_pos(gk[i])];
  }
  if (tid == 0) {
    A.counts[label] = n;
    __threadfence();
    s_last = (atomicAdd(A.done, 1u) == static_cast<unsigned>(A.n_labels) - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (s_last && tid == 0) {
    __threadfence();
    int acc = 0;
    for (int i = 0; i < A.n_labels; ++i) {
      A.offsets[i] = acc;
      acc += reinterpret_cast<volatile int*>(A.counts)[i];
    }
    A.offsets[A.n_labels] = acc;
    *A.done = 0u;                                 // ready for the next launch
  }
}

__global__ void __launch_bounds__(256) gather_detections_kernel(const DetectArgs A, const long long* __restrict__ label_values,
                                                                 float4* __restrict__ boxes, float* __restrict__ scores,
                                                                 long long* __restrict__ labels, float4* __restrict__ anchors,
                                                                 float* __restrict__ corners) {
  const int label = blockIdx.x;
  const int n = A.counts[label], off = A.offsets[label];
  const int* ids = A.out_ids + static_cast<long long>(A.view_off[label]) * A.sumN;
  const long long lab = label_values[label];
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
    const int f = ids[i];
    const Split s = split_flat(A, f);
    const DetectLevel& lv = A.lv[s.level];
    const Decoded d = decode_flat(A, f);
    const int o = off + i;
    boxes[o] = d.box;
    anchors[o] = d.anchor;
    scores[o] = lv.score[s.rem];
    labels[o] = lab;
    if (corners != nullptr) {
      const int view = static_cast<int>(s.rem / lv.N), a = static_cast<int>(s.rem - static_cast<long long>(view) * lv.N);
      const float* co = lv.corners + static_cast<size_t>(view) * 8 * lv.N + a;
      const float sx = lv.scale_x, sy = lv.same_scale ? lv.scale_x : lv.scale_y;
      float* dst = corners + static_cast<size_t>(o) * 8;
#pragma unroll
      for (int q = 0; q < 8; ++q) dst[q] = __fmul_rn(co[static_cast<size_t>(q) * lv.N], (q & 1) ? sy : sx);
    }
  }
}

}  // namespace detect

int launch_label_nms(const DetectArgs& A, cudaStream_t st) {
  using namespace detect;
  if (A.L <= 0 || A.L > kMaxPyramidLevels || A.n_labels <= 0 || A.C <= 0) return kErrBadArg;
  OS2D_SET_MAX_DYN_SMEM(label_nms_kernel, kSmemBytes);
  label_nms_kernel<<<A.n_labels, kThreads, kSmemBytes, st>>>(A);
  OS2D_AFTER_LAUNCH();
  return kOk;
}

int launch_gather_detections(const DetectArgs& A, const long long* label_values, float* boxes, float* scores, long long* labels,
                             float* anchors, float* corners, cudaStream_t st) {
  using namespace detect;
  if (A.L <= 0 || A.L > kMaxPyramidLevels || A.n_labels <= 0) return kErrBadArg;
  gather_detections_kernel<<<dim3(A.n_labels, 4), 256, 0, st>>>(A, label_values, reinterpret_cast<float4*>(boxes), scores, labels,
                                                                reinterpret_cast<float4*>(anchors), corners);
  OS2D_AFTER_LAUNCH();
  return kOk;
}

}  // namespace os2d

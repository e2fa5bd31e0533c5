// Shared device helpers for the sm_100a kernels: mbarrier, TMA (tensor + 1-D bulk), tcgen05
// (alloc / mma / commit / ld) wrappers written as raw PTX, UMMA descriptor builders, and the
// host-side tensor-map encoder (driver entry point fetched through the runtime, no libcuda link).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace os2d {

// ----------------------------------------------------------------------------------------------
// error codes of the C ABI (include/os2d_b200.h)
// ----------------------------------------------------------------------------------------------
enum : int { kOk = 0, kErrBadArg = -1, kErrCuda = -2, kErrUnsupported = -3, kErrDriver = -4 };

#define OS2D_CUDA_TRY(expr)                                  \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) { os2d::set_last_error(#expr, _e); return os2d::kErrCuda; } \
  } while (0)

void set_last_error(const char* what, cudaError_t e);
void set_last_error_msg(const char* what);
// kernel-launch counter of the library (os2d_b200_launch_count): every launcher reports its launches
void note_launch();
#define OS2D_AFTER_LAUNCH()                    \
  do {                                         \
    OS2D_CUDA_TRY(cudaGetLastError());         \
    os2d::note_launch();                       \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: set it once per (kernel, device), keyed
// on the current device like num_sms_cached() in c_api.cu (a per-process flag breaks the first launch on a second GPU).
constexpr int kMaxDevices = 64;
#define OS2D_SET_MAX_DYN_SMEM(kernel, bytes)                                                                   \
  do {                                                                                                         \
    static bool _done[os2d::kMaxDevices] = {};                                                                 \
    int _dev = 0;                                                                                              \
    OS2D_CUDA_TRY(cudaGetDevice(&_dev));                                                                       \
    if (_dev < 0 || _dev >= os2d::kMaxDevices) { os2d::set_last_error_msg("device index out of range"); return os2d::kErrUnsupported; } \
    if (!_done[_dev]) {                                                                                        \
      OS2D_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes))); \
      _done[_dev] = true;                                                                                      \
    }                                                                                                          \
  } while (0)

// ----------------------------------------------------------------------------------------------
// small PTX wrappers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------------
// Every kernel of the head chain (image pack -> K1 -> conv1..3 -> K3) triggers its dependents first thing and waits for its
// prerequisite right after its own set-up (barrier init, TMEM allocation, descriptor prefetch): the next kernel's CTAs are
// placed and set up while the tail of the previous kernel is still running; no global memory is touched before the wait.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---- TMA -------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
// 1-D bulk copy global -> shared (size multiple of 16 B, both addresses 16 B aligned)
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- tcgen05 / TMEM --------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, fp16 operands, fp32 accumulate.  Issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- inverse of the affine model (head.py:100-153) ---------------------------------------------------------------------
// [[a,b,tx],[c,d,ty],[0,0,1]] -> its inverse, closed form.  Exactly singular matrices (fp32 a*d - b*c == 0, evaluated
// without FMA contraction like the elementwise torch expression of the oracle) get the reference's failure handling
// (robust_inverse, head.py:123-134): 1e-5 is added to the diagonal of the 3x3, homogeneous 1 included, and the inverse is
// taken of that matrix - here in fp64 (rare branch).  The reference regularises the whole chunk of 65535 matrices that
// contains a singular one; the other matrices of such a chunk move by ~1e-5 relative, which this code does not reproduce.
__device__ __forceinline__ void invert_affine(float& a, float& b, float& tx, float& c, float& d, float& ty) {
  const float det = __fsub_rn(__fmul_rn(a, d), __fmul_rn(b, c));
  if (det != 0.f) {
    const float ia = d / det, ib = -b / det, ic = -c / det, id = a / det;
    const float itx = -(ia * tx + ib * ty), ity = -(ic * tx + id * ty);
    a = ia; b = ib; c = ic; d = id; tx = itx; ty = ity;
  } else {
    const double e = 1e-5;
    const double ar = static_cast<double>(a) + e, dr = static_cast<double>(d) + e, br = b, cr = c;
    const double detr = ar * dr - br * cr;
    const double ia = dr / detr, ib = -br / detr, ic = -cr / detr, id = ar / detr;
    const double itx = -(ia * tx + ib * ty) / (1.0 + e), ity = -(ic * tx + id * ty) / (1.0 + e);
    a = static_cast<float>(ia); b = static_cast<float>(ib); c = static_cast<float>(ic); d = static_cast<float>(id);
    tx = static_cast<float>(itx); ty = static_cast<float>(ity);
  }
}

// ---- UMMA descriptors ------------------------------------------------------------------------
// Shared-memory matrix descriptor (sm_100 layout): start[0,14) LBO[16,30) SBO[32,46) version[46,48)=1
// base_offset[49,52) layout[61,64) (0 none, 2 = 128B swizzle).  All byte quantities >> 4.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout & 7) << 61;
  return d;
}
// Instruction descriptor, kind::f16: fp16 A/B (K-major both), fp32 D, shape M x N.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4)                                   // D format = F32
         | (0u << 7) | (0u << 10)                    // A, B format = F16
         | (0u << 15) | (0u << 16)                   // A, B K-major
         | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// ----------------------------------------------------------------------------------------------
// host: tensor-map encoding through the driver entry point
// ----------------------------------------------------------------------------------------------
// rank <= 5; dims/strides fastest-first; strides[i] = byte stride of dim i+1 (rank-1 entries).
// Launch with the programmatic-stream-serialization attribute (+ an optional cluster dimension).  OS2D_B200_NO_PDL=1 in
// the environment falls back to plain stream order (A/B switch).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster_x); attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cfg.attrs = n ? attr : nullptr; cfg.numAttrs = static_cast<unsigned>(n);
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int encode_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                      const uint32_t* box, CUtensorMapSwizzle swizzle, CUtensorMapL2promotion promo);

}  // namespace os2d

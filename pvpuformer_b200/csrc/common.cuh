// Common device helpers for the sm_100a kernels of pvpuformer_b200: error plumbing, bf16
// packing, mbarrier / TMA / tcgen05 PTX wrappers.  Hand-written PTX; no CUTLASS dependency.
#pragma once
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vpu {

// ---- host-side error plumbing (thread-local message read back by vpu_last_error) -----------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);        // every kernel launch of this library is counted (vpu_launch_count)
// Development knobs (A/B variants of kernels, ablations) are read from the environment only in -DVPU_DEBUG builds
// (python -m pvpuformer_b200.build --debug); the shipped library has one path per operation and ignores them.
#ifdef VPU_DEBUG
inline const char* vpu_debug_env(const char* name) { return getenv(name); }
#else
inline const char* vpu_debug_env(const char*) { return nullptr; }
#endif
unsigned long long launch_count();
#define VPU_CHECK_CUDA(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            vpu::set_error("%s:%d CUDA error %d (%s) in %s", __FILE__, __LINE__, (int)_e, \
                           cudaGetErrorString(_e), #expr);                                \
            return 2;                                                                     \
        }                                                                                 \
    } while (0)
#define VPU_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            vpu::set_error(__VA_ARGS__);  \
            return 1;                     \
        }                                 \
    } while (0)

bool pdl_enabled();      // programmatic dependent launch: VPU_PDL=1/0 forces it, default = the forward's choice (gemm.cu)
void pdl_set_auto(bool on);   // per-thread default used when VPU_PDL is not set (run_forward: on for small batches)
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// ---- small math / packing -----------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
// erf-GELU (reference nn.GELU(), models_vit.py:14-27) evaluated as x * Phi(x) with
// Phi(x) = 0.5 (1 + tanh(u)) = 1 / (1 + exp(-2u)), u = x (a + b x^2 + c x^4), (a, b, c) a minimax fit of the erf form:
// |gelu(x) - x Phi(x)| <= 2.6e-5 for every x (the fit is exact to O(x^3) at 0).  ex2.approx + rcp.approx keep the
// evaluation error at ~1e-6 (tanh.approx's 2^-11 relative error moved the stress-weights parity test), so the total
// stays below the half-ulp of the bf16 this epilogue stores.  7 FMA-pipe instructions + two MUFUs per element: the
// previous Abramowitz-Stegun erf (19 instructions + MUFU) made the fc1 epilogue (32768 elements per CTA tile) longer
// than the tile's MMA time (A/B on B200, fc1 M=50176: 243 us -> 218 us).
#ifdef VPU_GELU_AS   // A/B build only: the previous Abramowitz-Stegun 7.1.28 erf (|error| <= 3e-7, 19 instructions + MUFU)
__device__ __forceinline__ float gelu_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    float p = 0.0000430638f;
    p = fmaf(p, z, 0.0002765672f); p = fmaf(p, z, 0.0001520143f); p = fmaf(p, z, 0.0092705272f);
    p = fmaf(p, z, 0.0422820123f); p = fmaf(p, z, 0.0705230784f); p = fmaf(p, z, 1.0f);
    p = p * p; p = p * p; p = p * p; p = p * p;
    float rp;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rp) : "f"(p));
    const float h = 0.5f * x;
    return fmaf(h, copysignf(1.0f - rp, x), h);
}
#else
__device__ __forceinline__ float gelu_fast(float x) {
    const float x2 = fminf(x * x, 64.0f);   // the quartic turns over at |x| = 11; Phi is saturated (|u| >= 13.8) from |x| = 8
    // -2 log2(e) * (a, b, c): v = -2 u log2(e)
    float t = fmaf(1.014263054e-3f, x2, -1.067757239e-1f);
    t = fmaf(t, x2, -2.301121339f);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * t));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));     // e = +inf (x << 0) -> r = 0
    return x * r;
}
#endif
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (-> cudaErrorLaunchFailure reported through the C-ABI)
// instead of hanging the device.  try_wait sleeps in hardware, so the bound costs nothing.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------
// With VPU_PDL=1 every kernel of this library is launched with cudaLaunchAttributeProgrammaticStreamSerialization (launch_pdl),
// calls pdl_launch_dependents() first thing -- so the next kernel of the stream may be scheduled while this one runs --
// and pdl_wait() before it reads or writes global memory: griddepcontrol.wait returns once the preceding kernel has
// completed and its writes are visible, so the semantics stay those of a serialised stream, minus the launch latency
// and the prologue (barrier init, TMEM allocation, descriptor prefetch) of ~185 kernels per forward.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// One lane of the (converged) warp.  Issue code for TMA / tcgen05 belongs under `if (elect_one())` in a warp that runs its
// loop converged, NOT under `if (lane == 0)`: there nvcc cannot prove single-lane execution and wraps every uniform-datapath
// instruction in an ELECT / R2UR / BRA.U.ANY loop.
__device__ __forceinline__ bool elect_one() {   // one lane of the (converged) warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- TMA ----------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// ---- tcgen05 / TMEM -------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp as alloc
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T ; bf16 inputs, fp32 accumulate; one thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp gets lane (base_lane + i), columns c..c+31.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// ---- 2-CTA (cta_group::2) variants: a CTA pair on one TPC shares one UMMA of M = 256 ------------------
// shared::cluster addresses carry the CTA rank of the pair in bit 24; clearing it addresses the leader
// (even) CTA's copy of the same shared-memory offset.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* tm, uint64_t* leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
// same, multicast: the box lands at this shared-memory offset in every CTA of `mask`, each destination's pair leader is signalled
__device__ __forceinline__ void tma_load_2d_2sm_mc(void* smem_dst, const CUtensorMap* tm, uint64_t* leader_bar, int c0, int c1,
                                                   uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1),
        "h"(mask)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {   // arrive on the leader CTA's barrier
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {   // one warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive (once all previously issued MMAs have completed) on the barrier at this offset in every CTA of `mask`.
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

// UMMA shared-memory matrix descriptor: K-major operand tile, 128-byte swizzle (the layout a
// TMA load with CU_TENSOR_MAP_SWIZZLE_128B and a 64-element bf16 inner box produces):
// rows of 128 B, 8-row swizzle atoms 1024 B apart (SBO), LBO unused, descriptor version 1.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                       // LBO (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // SBO = 1024 B
    d |= (uint64_t)1 << 46;                       // version = 1 (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}
// Instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, dense.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
#endif  // __CUDACC__

}  // namespace vpu

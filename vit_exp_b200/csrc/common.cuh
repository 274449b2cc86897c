// Shared device/host helpers for the ctk (CT-CLIP kernels) library: error plumbing,
// mbarrier / TMA / tcgen05 PTX wrappers for sm_100a, small math utilities.
#pragma once
#include <stdlib.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ctk.h"

// ----------------------------------------------------------------------------------------------
// host-side error plumbing
// ----------------------------------------------------------------------------------------------
void ctk_set_error(const char* fmt, ...);
void ctk_count_launch();  // diagnostic counter behind ctk_launch_count() (bench.py's gpu_launches)
int ctk_check_device();   // CTK_OK or CTK_ERR_ARCH (no CPU / non-sm_100 fallback)
// cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link dependency).
// dims/box innermost first; strides_bytes has rank-1 entries (dims 1..rank-1). swizzle: 0 none, 1 32B, 2 64B, 3 128B.
int ctk_make_tmap(CUtensorMap* m, const void* ptr, bool f32, int rank, const unsigned long long* dims,
                  const unsigned long long* strides_bytes, const unsigned int* box, int swizzle);

// tcgen05/TMEM spatial attention (attention_tc.cu); 24x24-token slices, head dim 32, bias table required
int ctk_attn_fwd_tc(const void* qkv, const float* table, void* out, float* lse, int nseq, int heads, cudaStream_t stream);
// dq / dk / dv / dtable of the same (attention_tc_bwd.cu); delta = rowsum(dO o O) must already be in `delta`
int ctk_attn_bwd_tc(const void* qkv, const float* table, const void* dout, const float* lse, const float* delta,
                    void* dqkv, float* dtable, int dtable_tc, int nseq, int heads, cudaStream_t stream);

// temporal stack (24-token sequences, no bias): TMA ring + warp-level MMA (attention_seq24.cu)
bool ctk_attn_seq24_supported(int L, int heads);
int ctk_attn_seq24_fwd(const void* qkv, void* out, float* lse, int nseq, int heads, cudaStream_t stream);
int ctk_attn_seq24_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int nseq,
                       int heads, cudaStream_t stream);

#define CTK_REQUIRE(cond, code, ...)                                                   \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            ctk_set_error(__VA_ARGS__);                                                \
            return (code);                                                             \
        }                                                                              \
    } while (0)

#define CTK_CUDA(call)                                                                 \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            ctk_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                 \
                          cudaGetErrorString(e__));                                    \
            return CTK_ERR_CUDA;                                                       \
        }                                                                              \
    } while (0)

#define CTK_LAUNCH_CHECK()                                                             \
    do {                                                                               \
        ctk_count_launch();                                                            \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            ctk_set_error("%s:%d launch -> %s", __FILE__, __LINE__,                    \
                          cudaGetErrorString(e__));                                    \
            return CTK_ERR_CUDA;                                                       \
        }                                                                              \
    } while (0)

#define CTK_ALIGNED(p, a) ((reinterpret_cast<uintptr_t>(p) % (a)) == 0)

// Per-DEVICE caches (a process may touch several GPUs; attributes set with cudaFuncSetAttribute and the SM count are
// properties of the device, not of the process).  Plain int / byte stores of an idempotent value: safe across threads.
constexpr int CTK_MAX_DEVICES = 64;
static inline int ctk_current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev;
}
static inline int ctk_num_sms() {
    static int n[CTK_MAX_DEVICES] = {};
    const int dev = ctk_current_device();
    const int slot = (dev >= 0 && dev < CTK_MAX_DEVICES) ? dev : 0;
    if (n[slot] == 0) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        n[slot] = v > 0 ? v : 148;
    }
    return n[slot];
}
// SMs the persistent kernels may fill.  CTK_GEMM_RESERVE_SMS=n leaves n SMs (rounded to CTA pairs) to co-running work:
// the tcgen05 GEMM partitions its tiles statically over its CTAs, so a CTA that cannot start because NCCL's kernels
// hold its SM delays the whole launch by that CTA's full share (multi-GPU runs: gradient all-reduces overlap the
// encoder's backward).
static inline int ctk_gemm_sms() {
    static int reserve = -1;
    if (reserve < 0) {
        const char* e = getenv("CTK_GEMM_RESERVE_SMS");
        reserve = e ? atoi(e) : 0;
        if (reserve < 0 || reserve > 64) reserve = 0;
    }
    int n = ctk_num_sms() - reserve;
    n &= ~1;
    return n < 2 ? 2 : n;
}
// opt a kernel into `bytes` of dynamic shared memory, once per device
#define CTK_SET_MAX_SMEM(kern, bytes)                                                                          \
    do {                                                                                                       \
        static volatile unsigned char _ctk_done[CTK_MAX_DEVICES] = {};                                         \
        const int _ctk_dev = ctk_current_device();                                                             \
        const bool _ctk_in = _ctk_dev >= 0 && _ctk_dev < CTK_MAX_DEVICES;                                      \
        if (!_ctk_in || !_ctk_done[_ctk_dev]) {                                                                \
            CTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));   \
            if (_ctk_in) _ctk_done[_ctk_dev] = 1;                                                              \
        }                                                                                                      \
    } while (0)

// ----------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

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

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}

// 2^x on the SFU (ex2.approx): inputs here are <= 0 after max subtraction, -inf -> +0
__device__ __forceinline__ float fast_exp2(float x) {
#ifdef CTK_NO_FAST_EXP2
    return exp2f(x);
#else
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}

// Standard normal cdf and pdf of g in ~18 instructions (2 SFU ops), sharing exp(-g^2/2):
// erf(x) = 1 - (a1 t + ... + a5 t^5) exp(-x^2), t = 1/(1 + p x)   (Abramowitz-Stegun 7.1.26,
// |error| <= 1.5e-7: below fp32 resolution of the bf16-rounded GEGLU outputs it feeds).
__device__ __forceinline__ void normal_cdf_pdf(float g, float& cdf, float& pdf) {
    const float ax = fabsf(g) * 0.70710678118654752f;
    float e, t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-ax * ax * 1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float half_tail = 0.5f * poly * t * e;                 // 0.5 * (1 - erf(|x|))
    cdf = g >= 0.f ? 1.0f - half_tail : half_tail;
    pdf = 0.39894228040143268f * e;
}

// ---- packed fp32 (two lanes per instruction: fma.rn.f32x2 issues one FFMA2 for two FMAs) ----------
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 f2_set(float v) { return make_float2(v, v); }
// normal_cdf_pdf for two values at once: the polynomial and the scalings run as packed fp32 instructions, the two
// ex2 / rcp stay scalar SFU operations.  Same formula, same accuracy.
__device__ __forceinline__ void normal_cdf_pdf2(float2 g, float2& cdf, float2& pdf) {
    const float2 ax = f2_mul(make_float2(fabsf(g.x), fabsf(g.y)), f2_set(0.70710678118654752f));
    const float2 arg = f2_mul(f2_mul(ax, ax), f2_set(-1.4426950408889634f));
    const float2 den = f2_fma(f2_set(0.3275911f), ax, f2_set(1.0f));
    float2 e, t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(arg.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(arg.y));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
    float2 poly = f2_fma(f2_set(1.061405429f), t, f2_set(-1.453152027f));
    poly = f2_fma(poly, t, f2_set(1.421413741f));
    poly = f2_fma(poly, t, f2_set(-0.284496736f));
    poly = f2_fma(poly, t, f2_set(0.254829592f));
    const float2 ht = f2_mul(f2_mul(poly, t), f2_mul(e, f2_set(0.5f)));         // 0.5 * (1 - erf(|x|))
    cdf.x = g.x >= 0.f ? 1.0f - ht.x : ht.x;
    cdf.y = g.y >= 0.f ? 1.0f - ht.y : ht.y;
    pdf = f2_mul(e, f2_set(0.39894228040143268f));
}

__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}
// d/dx gelu(x) = Phi(x) + x * phi(x)
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// ---- mbarrier --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Spin guard: a wedged pipeline traps (-> CUDA error on the host) instead of hanging the GPU.
#ifndef CTK_SPIN_LIMIT
#define CTK_SPIN_LIMIT (1u << 26)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > CTK_SPIN_LIMIT) {
            printf("ctk: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x,
                   (int)threadIdx.x);
            __trap();
        }
    }
}

// ---- TMA -------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
        : "memory");
}

// 1-D bulk copy global -> shared (16-byte aligned addresses and size), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- tcgen05 / TMEM --------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread.
__device__ __forceinline__ void tc_mma_f16(uint32_t taddr_d, uint64_t adesc, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(taddr_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives once every tcgen05 op issued so far by this thread has completed.
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
            smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread l of the warp gets TMEM lane (base_lane + l).
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// ---- CTA-pair (cta_group::2) variants: two SMs of one TPC cooperate on a 256-row MMA -----------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair; completion bytes are posted on the barrier at
// `bar_cluster_addr` (the leader CTA's barrier).
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                 int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t taddr_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(taddr_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at the same shared-memory offset in every CTA of `cta_mask` once all
// tcgen05 ops issued so far by this thread have completed
__device__ __forceinline__ void tc_commit_pair_mc(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(cta_mask)
        : "memory");
}

// UMMA shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
// Field layout follows cute::UMMA::SmemDescriptor (bits: addr[0,14) lbo[16,30) sbo[32,46)
// version[46,48) layout_type[61,64)); SWIZZLE_128B == 2.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                    uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, bf16 A/B.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, int a_mn_major,
                                                       int b_mn_major) {
    return (1u << 4)                                  // c_format  = F32
           | (1u << 7)                                // a_format  = BF16
           | (1u << 10)                               // b_format  = BF16
           | (static_cast<uint32_t>(a_mn_major) << 15)
           | (static_cast<uint32_t>(b_mn_major) << 16)
           | (static_cast<uint32_t>(n >> 3) << 17)
           | (static_cast<uint32_t>(m >> 4) << 24);
}

#endif  // __CUDACC__

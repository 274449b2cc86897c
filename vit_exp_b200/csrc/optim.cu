// Optimizer tail of the CT-CLIP train step: global-norm gradient clipping
// (CTCLIPTrainer.py:711-712, torch.nn.utils.clip_grad_norm_) folded into Adam / AdamW
// (optimizer.py:14-24) over all parameter tensors in two launches:
//   1. ctk_multi_sqnorm : sum of squares of every gradient element -> one fp32 scalar
//   2. ctk_multi_adam   : g *= min(1, max_norm / (norm + 1e-6)); moments; parameter update
// Tensors are described by a device table (one row per tensor: p, g, m, v pointers, element count,
// first chunk); a CTA owns one 64 K-element chunk and finds its tensor by binary search, so ~400
// tensors of 135 M parameters take two launches instead of torch's ~30.  HBM-bound: 4 reads +
// 3 writes of 4 B per parameter (the clipped gradient is not written back).
#include "common.cuh"

namespace {

constexpr int OPT_CHUNK = 65536;      // elements per CTA
constexpr int OPT_THREADS = 512;

struct __align__(8) OptRow {
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;
    long long chunk0;                 // index of the tensor's first chunk
};

__device__ __forceinline__ int find_tensor(const OptRow* __restrict__ rows, int ntensors, long long chunk) {
    int lo = 0, hi = ntensors - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (rows[mid].chunk0 <= chunk) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(OPT_THREADS)
multi_sqnorm_kernel(const OptRow* __restrict__ rows, int ntensors, float* __restrict__ out_sq) {
    __shared__ float red[OPT_THREADS / 32];
    const int t = find_tensor(rows, ntensors, blockIdx.x);
    const OptRow r = rows[t];
    const long long base = ((long long)blockIdx.x - r.chunk0) * OPT_CHUNK;
    const long long end = min(r.n, base + OPT_CHUNK);
    const float* g = r.g;
    float acc = 0.f;
    if ((reinterpret_cast<uintptr_t>(g + base) & 15) == 0) {
        const long long n4 = (end - base) >> 2;
        const float4* g4 = reinterpret_cast<const float4*>(g + base);
        for (long long i = threadIdx.x; i < n4; i += OPT_THREADS) {
            const float4 x = g4[i];
            acc = fmaf(x.x, x.x, acc); acc = fmaf(x.y, x.y, acc); acc = fmaf(x.z, x.z, acc); acc = fmaf(x.w, x.w, acc);
        }
        for (long long i = base + (n4 << 2) + threadIdx.x; i < end; i += OPT_THREADS) acc = fmaf(g[i], g[i], acc);
    } else {
        for (long long i = base + threadIdx.x; i < end; i += OPT_THREADS) acc = fmaf(g[i], g[i], acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < OPT_THREADS / 32 ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(out_sq, v);
    }
}

struct AdamArgs {
    float lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, max_norm;
    int adamw, write_grad;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& a) {
    if (a.weight_decay != 0.f) {
        if (a.adamw) p -= a.lr * a.weight_decay * p;       // decoupled decay (torch.optim.AdamW)
        else g = fmaf(a.weight_decay, p, g);               // L2 term in the gradient (torch.optim.Adam)
    }
    m = fmaf(a.beta1, m, (1.f - a.beta1) * g);
    v = fmaf(a.beta2, v, (1.f - a.beta2) * g * g);
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p -= (a.lr / a.bc1) * (m / denom);
}

__global__ void __launch_bounds__(OPT_THREADS)
multi_adam_kernel(const OptRow* __restrict__ rows, int ntensors, const float* __restrict__ sq, AdamArgs a) {
    const int t = find_tensor(rows, ntensors, blockIdx.x);
    const OptRow r = rows[t];
    const long long base = ((long long)blockIdx.x - r.chunk0) * OPT_CHUNK;
    const long long end = min(r.n, base + OPT_CHUNK);
    float coef = 1.f;
    if (a.max_norm > 0.f) coef = fminf(1.f, a.max_norm / (sqrtf(__ldg(sq)) + 1e-6f));   // clip_grad_norm_
    const bool vec = ((reinterpret_cast<uintptr_t>(r.p + base) | reinterpret_cast<uintptr_t>(r.g + base) |
                       reinterpret_cast<uintptr_t>(r.m + base) | reinterpret_cast<uintptr_t>(r.v + base)) & 15) == 0;
    long long scalar_from = base;
    if (vec) {
        const long long n4 = (end - base) >> 2;
        float4* p4 = reinterpret_cast<float4*>(r.p + base);
        const float4* g4 = reinterpret_cast<const float4*>(r.g + base);
        float4* m4 = reinterpret_cast<float4*>(r.m + base);
        float4* v4 = reinterpret_cast<float4*>(r.v + base);
        for (long long i = threadIdx.x; i < n4; i += OPT_THREADS) {
            float4 p = p4[i], g = g4[i], m = m4[i], v = v4[i];
            g.x *= coef; g.y *= coef; g.z *= coef; g.w *= coef;
            adam_one(p.x, g.x, m.x, v.x, a); adam_one(p.y, g.y, m.y, v.y, a);
            adam_one(p.z, g.z, m.z, v.z, a); adam_one(p.w, g.w, m.w, v.w, a);
            p4[i] = p; m4[i] = m; v4[i] = v;
            if (a.write_grad) const_cast<float4*>(g4)[i] = g;
        }
        scalar_from = base + (n4 << 2);
    }
    for (long long i = scalar_from + threadIdx.x; i < end; i += OPT_THREADS) {
        float p = r.p[i], g = r.g[i] * coef, m = r.m[i], v = r.v[i];
        adam_one(p, g, m, v, a);
        r.p[i] = p; r.m[i] = m; r.v[i] = v;
        if (a.write_grad) const_cast<float*>(r.g)[i] = g;
    }
}

// table upload by loads over PCIe instead of a DMA copy: a copy-engine transfer would queue behind the
// multi-GB input batch that the training loop keeps in flight on another stream
__global__ void copy_pinned_kernel(int4* __restrict__ dst, const int4* __restrict__ src, int n16) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // namespace

extern "C" int ctk_opt_chunk_elems(void) { return OPT_CHUNK; }

// dst (device) <- src (pinned host memory, device-accessible through UVA); bytes % 16 == 0, both 16-byte aligned
extern "C" int ctk_copy_from_pinned(void* dst, const void* src_pinned, long long bytes, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(dst && src_pinned && bytes > 0 && bytes % 16 == 0 && CTK_ALIGNED(dst, 16) && CTK_ALIGNED(src_pinned, 16),
                CTK_ERR_ALIGN, "copy_from_pinned: 16-byte aligned buffers and size");
    const int n16 = (int)(bytes / 16);
    copy_pinned_kernel<<<(n16 + 255) / 256 > 8 ? 8 : (n16 + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
        reinterpret_cast<int4*>(dst), reinterpret_cast<const int4*>(src_pinned), n16);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

// rows: device table [ntensors] of {p, g, m, v, n, chunk0} (6 x 8 bytes each); out_sq is overwritten.
extern "C" int ctk_multi_sqnorm(const void* rows, int ntensors, long long nchunks, float* out_sq, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(rows && out_sq && ntensors > 0 && nchunks > 0 && nchunks < (1LL << 31), CTK_ERR_SHAPE, "multi_sqnorm: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    CTK_CUDA(cudaMemsetAsync(out_sq, 0, sizeof(float), s));
    multi_sqnorm_kernel<<<(unsigned)nchunks, OPT_THREADS, 0, s>>>(reinterpret_cast<const OptRow*>(rows), ntensors, out_sq);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

// sq: device scalar written by ctk_multi_sqnorm (ignored when max_norm <= 0).  step >= 1.
extern "C" int ctk_multi_adam(const void* rows, int ntensors, long long nchunks, const float* sq, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int adamw, long long step, float max_norm,
                              int write_clipped_grad, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(rows && ntensors > 0 && nchunks > 0 && nchunks < (1LL << 31) && step >= 1 && (sq || max_norm <= 0.f),
                CTK_ERR_SHAPE, "multi_adam: bad args");
    AdamArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
    a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    a.max_norm = max_norm; a.adamw = adamw; a.write_grad = write_clipped_grad;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    multi_adam_kernel<<<(unsigned)nchunks, OPT_THREADS, 0, s>>>(reinterpret_cast<const OptRow*>(rows), ntensors, sq, a);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

// Library plumbing: error strings, device gate, small utility kernels (fill, casts, packing).
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void ctk_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int ctk_check_device() {
    // cached per device; the only supported target is compute capability 10.x (sm_100a cubin).
    static int ok_dev[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        ctk_set_error("no CUDA device: libctk has no CPU path");
        return CTK_ERR_ARCH;
    }
    if (dev < 64 && ok_dev[dev] == 1) return CTK_OK;
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) {
        ctk_set_error("device %d has compute capability %d.x; libctk is sm_100a only", dev, major);
        return CTK_ERR_ARCH;
    }
    if (dev < 64) ok_dev[dev] = 1;
    return CTK_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int ctk_make_tmap(CUtensorMap* m, const void* ptr, bool f32, int rank, const unsigned long long* dims,
                  const unsigned long long* strides_bytes, const unsigned int* box, int swizzle) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    CTK_REQUIRE(fn != nullptr, CTK_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    CTK_REQUIRE(rank >= 1 && rank <= 5, CTK_ERR_SHAPE, "tensor map rank %d", rank);
    cuuint64_t d[5], st[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
    const CUtensorMapSwizzle sw = swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
                    const_cast<void*>(ptr), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CTK_REQUIRE(r == CUDA_SUCCESS, CTK_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d), rank %d", (int)r, rank);
    return CTK_OK;
}

#include <atomic>
static std::atomic<unsigned long long> g_launches{0};
void ctk_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" unsigned long long ctk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" const char* ctk_last_error(void) { return g_err; }
extern "C" int ctk_version(void) { return 100; }
extern "C" int ctk_device_ok(void) { return ctk_check_device(); }

namespace {

__global__ void fill_kernel(float* p, float v, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

__global__ void cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                            long long rows, long long cols, long long ld,
                            const float* __restrict__ col_scale) {
    const long long total = rows * ld;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const long long r = i / ld, c = i % ld;
        float v = 0.f;
        if (c < cols) {
            v = src[r * cols + c];
            if (col_scale) v *= col_scale[c];
        }
        dst[i] = __float2bfloat16(v);
    }
}

// dst[c, r] = src[r, c]; 32x32 tiles through shared memory.
__global__ void transpose_cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                      long long rows, long long cols, long long ld, long long wlimit) {
    __shared__ float tile[32][33];
    const long long c0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long r = r0 + j, c = c0 + threadIdx.x;
        tile[j][threadIdx.x] = (r < rows && c < cols) ? src[r * cols + c] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long c = c0 + j, r = r0 + threadIdx.x;     // dst row = c, dst col = r
        if (c < cols && r < wlimit) dst[c * ld + r] = __float2bfloat16(tile[threadIdx.x][j]);
    }
}

// zero the pad columns [rows, ld) of a transposed copy
__global__ void zero_pad_cols_kernel(__nv_bfloat16* dst, long long nrows, long long c_begin,
                                     long long ld) {
    const long long w = ld - c_begin;
    const long long total = nrows * w;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < total; i += (long long)gridDim.x * blockDim.x)
        dst[(i / w) * ld + c_begin + (i % w)] = __float2bfloat16(0.f);
}

// interleaved row r' of the packed W1: block = r'/256, within = r'%256;
// within < 128 -> value unit block*128+within ; else gate unit block*128+within-128.
__global__ void pack_w1_kernel(const float* __restrict__ w1, __nv_bfloat16* __restrict__ dst,
                               __nv_bfloat16* __restrict__ dst_t, int* __restrict__ row_map,
                               int inner, int inner_pad, int dim) {
    const int rp = blockIdx.x;                       // packed row
    const int block = rp / 256, within = rp % 256;
    const int unit = block * 128 + (within % 128);
    const bool gate = within >= 128;
    const int src_row = unit < inner ? (gate ? inner + unit : unit) : -1;
    if (threadIdx.x == 0 && row_map) row_map[rp] = src_row;
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
        const float v = src_row >= 0 ? w1[(long long)src_row * dim + c] : 0.f;
        const __nv_bfloat16 b = __float2bfloat16(v);
        dst[(long long)rp * dim + c] = b;
        if (dst_t) dst_t[(long long)c * (2 * inner_pad) + rp] = b;
    }
}

__global__ void patch_affine_bwd_kernel(const float* __restrict__ P, const float* __restrict__ W,
                                        const float* __restrict__ gamma,
                                        const float* __restrict__ beta,
                                        const float* __restrict__ db, float* __restrict__ dW,
                                        float* __restrict__ dgamma, float* __restrict__ dbeta,
                                        int n, int k) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= k) return;
    const float g = gamma[col], b = beta[col];
    float dg = 0.f, dbt = 0.f;
    for (int r = 0; r < n; ++r) {
        const float p = P[(long long)r * k + col];
        const float w = W[(long long)r * k + col];
        dW[(long long)r * k + col] = g * p + b * db[r];
        dg += w * p;
        dbt += w * db[r];
    }
    dgamma[col] = dg;
    dbeta[col] = dbt;
}

// column sums: block handles 32 columns x a slab of rows, atomics at the end
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ dy, float* __restrict__ out, long long rows,
                              int cols, long long rows_per_block) {
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > rows) r1 = rows;
    float acc = 0.f;
    if (c < cols)
        for (long long r = r0 + threadIdx.y; r < r1; r += 8) acc += (float)dy[r * cols + c];
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += red[j][threadIdx.x];
        atomicAdd(out + c, s);
    }
}

}  // namespace

extern "C" int ctk_fill_f32(float* p, float v, long long n, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    if (n <= 0) return CTK_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    long long blocks = (n + 255) / 256;
    if (blocks > 4 * 148 * 8) blocks = 4 * 148 * 8;
    fill_kernel<<<(int)blocks, 256, 0, s>>>(p, v, n);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

extern "C" int ctk_cast_bf16(const float* src, void* dst, long long rows, long long cols,
                             long long ld_dst, const float* col_scale, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(src && dst && rows > 0 && cols > 0 && ld_dst >= cols, CTK_ERR_SHAPE, "cast: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    long long blocks = (rows * ld_dst + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cast_kernel<<<(int)blocks, 256, 0, s>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), rows, cols,
                                           ld_dst, col_scale);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

extern "C" int ctk_transpose_cast_bf16(const float* src, void* dst, long long rows, long long cols,
                                       long long ld_dst, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(src && dst && rows > 0 && cols > 0 && ld_dst >= rows, CTK_ERR_SHAPE,
                "transpose_cast: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    transpose_cast_kernel<<<grid, dim3(32, 8), 0, s>>>(src, reinterpret_cast<__nv_bfloat16*>(dst),
                                                       rows, cols, ld_dst, ld_dst);
    CTK_LAUNCH_CHECK();
    const long long rows_cov = ((rows + 31) / 32) * 32;   // columns already written (zeros past rows)
    if (ld_dst > rows_cov) {
        zero_pad_cols_kernel<<<64, 256, 0, s>>>(reinterpret_cast<__nv_bfloat16*>(dst), cols,
                                                rows_cov, ld_dst);
        CTK_LAUNCH_CHECK();
    }
    return CTK_OK;
}

extern "C" int ctk_transpose_cast_bf16_slice(const float* src, void* dst, long long rows, long long cols,
                                             long long ld_dst, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(src && dst && rows > 0 && cols > 0 && ld_dst >= rows, CTK_ERR_SHAPE,
                "transpose_cast_slice: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    transpose_cast_kernel<<<grid, dim3(32, 8), 0, s>>>(src, reinterpret_cast<__nv_bfloat16*>(dst),
                                                       rows, cols, ld_dst, rows);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

extern "C" int ctk_pack_ff_w1(const float* w1, void* dst, void* dst_t, int* row_map, int inner,
                              int inner_pad, int dim, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(w1 && dst && inner > 0 && inner_pad >= inner && inner_pad % 128 == 0 && dim > 0,
                CTK_ERR_SHAPE, "pack_ff_w1: inner_pad must be a multiple of 128 and >= inner");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    pack_w1_kernel<<<2 * inner_pad, 128, 0, s>>>(w1, reinterpret_cast<__nv_bfloat16*>(dst),
                                                 reinterpret_cast<__nv_bfloat16*>(dst_t), row_map,
                                                 inner, inner_pad, dim);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

extern "C" int ctk_patch_affine_bwd(const float* P, const float* W, const float* gamma,
                                    const float* beta, const float* db, float* dW, float* dgamma,
                                    float* dbeta, int n, int k, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(P && W && gamma && beta && db && dW && dgamma && dbeta && n > 0 && k > 0,
                CTK_ERR_SHAPE, "patch_affine_bwd: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    patch_affine_bwd_kernel<<<(k + 127) / 128, 128, 0, s>>>(P, W, gamma, beta, db, dW, dgamma,
                                                            dbeta, n, k);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

extern "C" int ctk_colsum(const void* dy_bf16, const float* dy_f32, float* out, long long rows,
                          int cols, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE((dy_bf16 || dy_f32) && out && rows > 0 && cols > 0, CTK_ERR_SHAPE, "colsum: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    const long long rpb = 2048;
    dim3 grid((cols + 31) / 32, (unsigned)((rows + rpb - 1) / rpb));
    if (dy_bf16)
        colsum_kernel<__nv_bfloat16><<<grid, dim3(32, 8), 0, s>>>(
            reinterpret_cast<const __nv_bfloat16*>(dy_bf16), out, rows, cols, rpb);
    else
        colsum_kernel<float><<<grid, dim3(32, 8), 0, s>>>(dy_f32, out, rows, cols, rpb);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

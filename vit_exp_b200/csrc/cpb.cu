// ContinuousPositionBias (attention.py:335-382) evaluated on the (2gh-1)(2gw-1) DISTINCT relative
// offsets instead of all (gh*gw)^2 pairs (177 GF -> 1.2 GF per forward, bit-identical
// table gather). fp32 throughout, as the reference forces (.float(), attention.py:377-380).
//   in[n] = (sign*log(|dy|+1), sign*log(|dx|+1)),  n = (dy+gh-1)*(2gw-1) + (dx+gw-1)
//   h0 = lrelu(in W0^T + b0); h1 = lrelu(h0 W1^T + b1); table[h][n] = h1 W2[h] + b2[h]
#include "common.cuh"
#include "sgemm.cuh"

namespace {

__device__ __forceinline__ float slog(int d) {
    const float a = logf(fabsf((float)d) + 1.0f);
    return d > 0 ? a : (d < 0 ? -a : 0.f);
}
__device__ __forceinline__ float lrelu(float z) { return z > 0.f ? z : 0.1f * z; }
__device__ __forceinline__ float lrelu_grad_from_out(float h) { return h > 0.f ? 1.f : 0.1f; }

__global__ void cpb_l0_kernel(const float* __restrict__ w0, const float* __restrict__ b0,
                              float* __restrict__ h0, int gh, int gw, int dim) {
    const int n = blockIdx.x;
    const int ww = 2 * gw - 1;
    const float fy = slog(n / ww - (gh - 1)), fx = slog(n % ww - (gw - 1));
    for (int c = threadIdx.x; c < dim; c += blockDim.x)
        h0[(long long)n * dim + c] = lrelu(fmaf(fy, w0[2 * c], fmaf(fx, w0[2 * c + 1], b0[c])));
}

// C = epilogue(A * B^T + bias) or A * B (NN) ... generic fp32 tile kernel with a small epilogue menu
//  MODE 0: C = lrelu(A B^T + bias[col])                       (layer 1 forward)
//  MODE 1: C = (A B) * lrelu'(aux[row, col])                  (dh0 -> dz0)
//  MODE 2: C = A^T B                                          (dW1 = dz1^T h0)
template <int MODE>
__global__ void __launch_bounds__(256)
cpb_gemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                const float* __restrict__ bias, const float* __restrict__ aux, int M, int N, int K,
                int lda, int ldb) {
    __shared__ float As[16][68];
    __shared__ float Bs[16][68];
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4];
    if (MODE == 0) sgemm_tile_64x64<true, false>(A, lda, B, ldb, m0, n0, M, N, K, acc, As, Bs);
    else if (MODE == 1) sgemm_tile_64x64<false, false>(A, lda, B, ldb, m0, n0, M, N, K, acc, As, Bs);
    else sgemm_tile_64x64<false, true>(A, lda, B, ldb, m0, n0, M, N, K, acc, As, Bs);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = m0 + ty * 4 + i;
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + tx * 4 + j;
            if (c >= N) continue;
            float v = acc[i][j];
            if (MODE == 0) v = lrelu(v + bias[c]);
            if (MODE == 1) v *= lrelu_grad_from_out(aux[(long long)r * N + c]);
            C[(long long)r * N + c] = v;
        }
    }
}

// table[h][n] = <h1[n], w2[h]> + b2[h]; one warp per offset n
__global__ void cpb_out_kernel(const float* __restrict__ h1, const float* __restrict__ w2,
                               const float* __restrict__ b2, float* __restrict__ table, int n_off,
                               int dim, int heads) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= n_off) return;
    for (int h = 0; h < heads; ++h) {
        float a = 0.f;
        for (int c = lane; c < dim; c += 32) a = fmaf(h1[(long long)n * dim + c], __ldg(w2 + (long long)h * dim + c), a);
        a = warp_sum(a);
        if (lane == 0) table[(long long)h * n_off + n] = a + b2[h];
    }
}

// dz1[n, c] = (sum_h dtable[h][n] w2[h][c]) * lrelu'(h1[n, c])
__global__ void cpb_dz1_kernel(const float* __restrict__ dtable, const float* __restrict__ w2,
                               const float* __restrict__ h1, float* __restrict__ dz1, int n_off,
                               int dim, int heads) {
    const int n = blockIdx.x;
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
        float a = 0.f;
        for (int h = 0; h < heads; ++h) a = fmaf(dtable[(long long)h * n_off + n], w2[(long long)h * dim + c], a);
        dz1[(long long)n * dim + c] = a * lrelu_grad_from_out(h1[(long long)n * dim + c]);
    }
}

// dw2[h][c] += sum_{n in chunk} dtable[h][n] h1[n][c];  db2[h] += sum dtable[h][n]    grid (heads, chunks)
__global__ void cpb_dw2_kernel(const float* __restrict__ dtable, const float* __restrict__ h1,
                               float* __restrict__ dw2, float* __restrict__ db2, int n_off, int dim, int per) {
    const int h = blockIdx.x;
    const int n0 = blockIdx.y * per, n1 = min(n_off, n0 + per);
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
        float a = 0.f;
        for (int n = n0; n < n1; ++n) a = fmaf(dtable[(long long)h * n_off + n], h1[(long long)n * dim + c], a);
        atomicAdd(dw2 + (long long)h * dim + c, a);
    }
    if (threadIdx.x < 32) {
        float a = 0.f;
        for (int n = n0 + threadIdx.x; n < n1; n += 32) a += dtable[(long long)h * n_off + n];
        a = warp_sum(a);
        if (threadIdx.x == 0) atomicAdd(db2 + h, a);
    }
}

// column sums of dz [n_off, dim] (-> db); optionally also dw0[c][0..1] = sum_n dz[n][c] * in[n][.]
// grid (dim / 128, chunks), atomics into pre-zeroed outputs
__global__ void cpb_colred_kernel(const float* __restrict__ dz, float* __restrict__ db,
                                  float* __restrict__ dw0, int n_off, int dim, int gh, int gw, int per) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= dim) return;
    const int n0 = blockIdx.y * per, n1 = min(n_off, n0 + per);
    const int ww = 2 * gw - 1;
    float s = 0.f, sy = 0.f, sx = 0.f;
    for (int n = n0; n < n1; ++n) {
        const float v = dz[(long long)n * dim + c];
        s += v;
        if (dw0) {
            sy = fmaf(v, slog(n / ww - (gh - 1)), sy);
            sx = fmaf(v, slog(n % ww - (gw - 1)), sx);
        }
    }
    atomicAdd(db + c, s);
    if (dw0) { atomicAdd(dw0 + 2 * c, sy); atomicAdd(dw0 + 2 * c + 1, sx); }
}

}  // namespace

extern "C" int ctk_cpb_fwd(const float* w0, const float* b0, const float* w1, const float* b1,
                           const float* w2, const float* b2, float* h0, float* h1, float* table,
                           int gh, int gw, int dim, int heads, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(w0 && b0 && w1 && b1 && w2 && b2 && h0 && h1 && table, CTK_ERR_SHAPE, "cpb_fwd: null pointer");
    CTK_REQUIRE(gh > 0 && gw > 0 && dim > 0 && heads > 0, CTK_ERR_SHAPE, "cpb_fwd: bad shape");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    const int n_off = (2 * gh - 1) * (2 * gw - 1);
    cpb_l0_kernel<<<n_off, 128, 0, s>>>(w0, b0, h0, gh, gw, dim);
    CTK_LAUNCH_CHECK();
    cpb_gemm_kernel<0><<<dim3((dim + 63) / 64, (n_off + 63) / 64), 256, 0, s>>>(h0, w1, h1, b1, nullptr, n_off, dim, dim, dim, dim);
    CTK_LAUNCH_CHECK();
    cpb_out_kernel<<<(n_off + 7) / 8, 256, 0, s>>>(h1, w2, b2, table, n_off, dim, heads);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

extern "C" int ctk_cpb_bwd(const float* dtable, const float* w0, const float* w1, const float* w2,
                           const float* h0, const float* h1, float* dw0, float* db0, float* dw1,
                           float* db1, float* dw2, float* db2, float* ws, int gh, int gw, int dim,
                           int heads, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(dtable && w0 && w1 && w2 && h0 && h1 && dw0 && db0 && dw1 && db1 && dw2 && db2 && ws,
                CTK_ERR_SHAPE, "cpb_bwd: null pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    const int n_off = (2 * gh - 1) * (2 * gw - 1);
    float* dz1 = ws;
    float* dz0 = ws + (size_t)n_off * dim;
    const int chunks = 32;
    const int per = (n_off + chunks - 1) / chunks;
    CTK_CUDA(cudaMemsetAsync(dw2, 0, sizeof(float) * (size_t)heads * dim, s));
    CTK_CUDA(cudaMemsetAsync(db2, 0, sizeof(float) * heads, s));
    CTK_CUDA(cudaMemsetAsync(db1, 0, sizeof(float) * dim, s));
    CTK_CUDA(cudaMemsetAsync(db0, 0, sizeof(float) * dim, s));
    CTK_CUDA(cudaMemsetAsync(dw0, 0, sizeof(float) * 2 * dim, s));
    cpb_dw2_kernel<<<dim3(heads, chunks), 256, 0, s>>>(dtable, h1, dw2, db2, n_off, dim, per);
    CTK_LAUNCH_CHECK();
    cpb_dz1_kernel<<<n_off, 128, 0, s>>>(dtable, w2, h1, dz1, n_off, dim, heads);
    CTK_LAUNCH_CHECK();
    // dW1[o][i] = sum_n dz1[n][o] h0[n][i]
    cpb_gemm_kernel<2><<<dim3((dim + 63) / 64, (dim + 63) / 64), 256, 0, s>>>(dz1, h0, dw1, nullptr, nullptr, dim, dim, n_off, dim, dim);
    CTK_LAUNCH_CHECK();
    cpb_colred_kernel<<<dim3((dim + 127) / 128, chunks), 128, 0, s>>>(dz1, db1, nullptr, n_off, dim, gh, gw, per);
    CTK_LAUNCH_CHECK();
    // dz0 = (dz1 W1) * lrelu'(h0)
    cpb_gemm_kernel<1><<<dim3((dim + 63) / 64, (n_off + 63) / 64), 256, 0, s>>>(dz1, w1, dz0, nullptr, h0, n_off, dim, dim, dim, dim);
    CTK_LAUNCH_CHECK();
    cpb_colred_kernel<<<dim3((dim + 127) / 128, chunks), 128, 0, s>>>(dz0, db0, dw0, n_off, dim, gh, gw, per);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

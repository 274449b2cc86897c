// Contrastive head: token mean-pool, latent projection + l2norm, and the fused symmetric
// InfoNCE forward/backward (ct_clip.py:1280-1388; closed form in SURVEY.md appendix B).
//
// Loss pipeline (every rank evaluates the full N x N matrix redundantly, as the reference does):
//   pass 1  clip_tile_stats : 64x64 logit tiles in fp32; per-tile online (max, sum e^x, sum x e^x)
//                             for rows and columns; diagonal. Logits stay in registers/smem.
//   pass 2  clip_reduce     : merge tile partials -> row/col log-sum-exp, loss, d(log temp).
//   pass 3  clip_grad_tiles : recompute logits for the rank's rows / columns only and emit
//                             G = dloss/dS stripes [b_local, N];  pass 4: dT = s G I, dI = s G^T T.
// For N >= 1024 (clip_loss_tc below; CTK_CLIP_LOSS_TC=0 disables) the same passes run with the contractions on the
// tcgen05 GEMM (split-bf16 operands, fused LSE / gradient epilogues).
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "sgemm.cuh"

namespace {

constexpr int CLIP_TC_MIN_N = 1024;     // below this the loss is launch-latency bound: fp32 SIMT tiles
constexpr int CLIP_TC_MAX_D = 512;      // latent width the tensor-core path's workspace is sized for

// ---------------------------------------------------------------- mean pool
__global__ void mean_pool_kernel(const float* __restrict__ x, float* __restrict__ pooled,
                                 long long n, int dim, long long tok_per_block) {
    const int b = blockIdx.y;
    const long long t0 = (long long)blockIdx.x * tok_per_block;
    long long t1 = t0 + tok_per_block;
    if (t1 > n) t1 = n;
    const float inv = 1.0f / (float)n;
    for (int c = threadIdx.x * 4; c < dim; c += blockDim.x * 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* p = x + ((long long)b * n + t0) * dim + c;
        for (long long t = t0; t < t1; ++t, p += dim) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float* o = pooled + (long long)b * dim + c;
        atomicAdd(o + 0, acc.x * inv);
        atomicAdd(o + 1, acc.y * inv);
        atomicAdd(o + 2, acc.z * inv);
        atomicAdd(o + 3, acc.w * inv);
    }
}

// Short axis, wide rows - the legacy pooling of CTCLIP.forward_old (ct_clip.py:1549: mean over the n = 24 frames of
// rows of h*w*C = 294 912 floats): one thread per float4 column, frames summed in index order.  Coalesced, no atomics,
// dim/1024 x B CTAs (mean_pool_kernel above would put B CTAs on it).
__global__ void __launch_bounds__(256)
mean_frames_kernel(const float* __restrict__ x, float* __restrict__ pooled, int n, long long dim4) {
    const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
    if (c >= dim4) return;
    const int b = blockIdx.y;
    const float4* p = reinterpret_cast<const float4*>(x) + (long long)b * n * dim4 + c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < n; ++t, p += dim4) {
        const float4 v = __ldg(p);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const float inv = 1.0f / (float)n;
    reinterpret_cast<float4*>(pooled)[(long long)b * dim4 + c] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
}

// ---------------------------------------------------------------- latent projection + l2norm
// raw[b, j] = <W[j, :], x[b, :]>.  One warp per output row j holds W[j, :] in registers (read once, coalesced) and
// sweeps the B input rows staged in shared memory; grid = dl / 8 CTAs (a B-CTA grid left 140 SMs idle and made every
// CTA stream the whole weight matrix).  din <= 1024.
constexpr int LAT_MAX_B = 8;               // input rows staged per pass: 8 x 1024 floats = 32 KB
__global__ void __launch_bounds__(256)
latent_raw_kernel(const float* __restrict__ x, long long x_stride, const float* __restrict__ W,
                  float* __restrict__ raw, int B, int din, int dl) {
    extern __shared__ float sm[];              // [min(B, LAT_MAX_B)][din]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + warp;
    float w[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const int i = lane + 32 * k;
        w[k] = (j < dl && i < din) ? __ldg(W + (long long)j * din + i) : 0.f;
    }
    for (int b0 = 0; b0 < B; b0 += LAT_MAX_B) {
        const int nb = min(LAT_MAX_B, B - b0);
        __syncthreads();
        for (int t = threadIdx.x; t < nb * din; t += blockDim.x)
            sm[t] = x[(long long)(b0 + t / din) * x_stride + t % din];
        __syncthreads();
        for (int b = 0; b < nb; ++b) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const int i = lane + 32 * k;
                if (i < din) acc = fmaf(w[k], sm[b * din + i], acc);
            }
            acc = warp_sum(acc);
            if (lane == 0 && j < dl) raw[(long long)(b0 + b) * dl + j] = acc;
        }
    }
}
// din > 1024: the to_visual_latent of the original CT-CLIP checkpoints, Linear(294 912 -> 512) on the legacy pooling
// (ct_clip.py:1614; scripts/run_zero_shot_latent.py:26-31) - 604 MB of fp32 weights against <= 8 input rows, bound by
// the one pass over W.  CTA = 8 warps = LATW_ROWS output rows x 2 halves of a LATW_KC-float slice of K; per slice the
// weight quads are requested first (8 x 16 B per lane in flight), then the input rows' slice is staged in shared
// memory and swept.  Each output is summed in a fixed order (no split-K atomics): bitwise repeatable.
constexpr int LATW_KC = 2048;
constexpr int LATW_ROWS = 4;
constexpr int LATW_Q = LATW_KC / 4 / 2 / 32;          // float4 per lane per slice = 8
constexpr int LATW_XV = LAT_MAX_B * (LATW_KC / 4) / 256;                     // staged float4 per thread per slice = 16
constexpr size_t LATW_SMEM = sizeof(float4) * LAT_MAX_B * (LATW_KC / 4);     // 64 KB
__global__ void __launch_bounds__(256)
latent_raw_wide_kernel(const float* __restrict__ x, long long x_stride, const float* __restrict__ W,
                       float* __restrict__ raw, int B, int din, int dl) {
    extern __shared__ float4 xs[];                    // [LAT_MAX_B][LATW_KC / 4]
    __shared__ float part[LATW_ROWS][2][LAT_MAX_B];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = warp >> 1, half = warp & 1;
    const int j = blockIdx.x * LATW_ROWS + r;
    constexpr int QS = LATW_KC / 4;                   // quads per slice
    for (int b0 = 0; b0 < B; b0 += LAT_MAX_B) {
        const int nb = min(LAT_MAX_B, B - b0);
        float acc[LAT_MAX_B];
#pragma unroll
        for (int bb = 0; bb < LAT_MAX_B; ++bb) acc[bb] = 0.f;
        // One slice of K in registers: this warp's weight quads and this thread's share of the input rows' slice
        // (all-zero past the end of K).  The requests for slice k+1 are issued before slice k is swept.
        float4 w[LATW_Q], xv[LATW_XV];
        auto request = [&](int k0, float4 (&wq)[LATW_Q]) {
            const int nq = min(LATW_KC, din - k0) >> 2;           // valid quads of the slice (din % 4 == 0); <= 0 past the end
            const float4* wrow = reinterpret_cast<const float4*>(W + (long long)(j < dl ? j : 0) * din + k0);
#pragma unroll
            for (int u = 0; u < LATW_Q; ++u) {
                const int q = half * (QS / 2) + u * 32 + lane;
                wq[u] = (j < dl && q < nq) ? __ldg(wrow + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < LATW_XV; ++i) {
                const int t = i * 256 + threadIdx.x, bb = t / QS, q = t % QS;
                xv[i] = (bb < nb && q < nq)
                            ? *reinterpret_cast<const float4*>(x + (long long)(b0 + bb) * x_stride + k0 + 4 * q)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        request(0, w);
        for (int k0 = 0; k0 < din; k0 += LATW_KC) {
            __syncthreads();                                       // previous slice fully consumed
#pragma unroll
            for (int i = 0; i < LATW_XV; ++i) xs[i * 256 + threadIdx.x] = xv[i];
            __syncthreads();
            float4 wn[LATW_Q];
            request(k0 + LATW_KC, wn);
#pragma unroll
            for (int bb = 0; bb < LAT_MAX_B; ++bb) {
                if (bb < nb) {
                    float a = acc[bb];
#pragma unroll
                    for (int u = 0; u < LATW_Q; ++u) {
                        const float4 xr = xs[bb * QS + half * (QS / 2) + u * 32 + lane];
                        a = fmaf(w[u].x, xr.x, a); a = fmaf(w[u].y, xr.y, a);
                        a = fmaf(w[u].z, xr.z, a); a = fmaf(w[u].w, xr.w, a);
                    }
                    acc[bb] = a;
                }
            }
#pragma unroll
            for (int u = 0; u < LATW_Q; ++u) w[u] = wn[u];
        }
#pragma unroll
        for (int bb = 0; bb < LAT_MAX_B; ++bb) {
            const float v = warp_sum(acc[bb]);
            if (lane == 0) part[r][half][bb] = v;
        }
        __syncthreads();
        if (threadIdx.x < LATW_ROWS * LAT_MAX_B) {
            const int r2 = threadIdx.x / LAT_MAX_B, bb = threadIdx.x % LAT_MAX_B;
            const int j2 = blockIdx.x * LATW_ROWS + r2;
            if (j2 < dl && bb < nb) raw[(long long)(b0 + bb) * dl + j2] = part[r2][0][bb] + part[r2][1][bb];
        }
        __syncthreads();
    }
}
// latent[b, :] = raw[b, :] / max(|raw[b, :]|, 1e-12) in place; rnorm[b] = the reciprocal (F.normalize, ct_clip.py:70-71)
__global__ void __launch_bounds__(256)
latent_norm_kernel(float* __restrict__ latent, float* __restrict__ rnorm, int dl) {
    __shared__ float red[8];
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    float ss = 0.f;
    for (int j = threadIdx.x; j < dl; j += blockDim.x) { const float v = latent[(long long)b * dl + j]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < nwarp; ++w) tot += red[w];
    const float rn = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
    for (int j = threadIdx.x; j < dl; j += blockDim.x) latent[(long long)b * dl + j] *= rn;
    if (threadIdx.x == 0) rnorm[b] = rn;
}

// draw[b] = rnorm_b * (dlat_b - lat_b * <lat_b, dlat_b>) ; dx[b, i] = sum_j draw[b, j] W[j, i].
// grid (ceil(din / 64), B); 256 threads = 64 columns i x 4 slices of j, reduced through shared memory.
__global__ void __launch_bounds__(256)
latent_bwd_dx_kernel(const float* __restrict__ dlat, const float* __restrict__ lat,
                     const float* __restrict__ rnorm, const float* __restrict__ W,
                     float* __restrict__ dx, long long dx_stride, int din, int dl) {
    extern __shared__ float sm[];
    float* draw = sm;          // [dl]
    __shared__ float red[8];
    __shared__ float part[4][64];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    float dot = 0.f;
    for (int j = threadIdx.x; j < dl; j += blockDim.x)
        dot += lat[(long long)b * dl + j] * dlat[(long long)b * dl + j];
    dot = warp_sum(dot);
    if (lane == 0) red[warp] = dot;
    __syncthreads();
    float coef = 0.f;
    for (int w = 0; w < nwarp; ++w) coef += red[w];
    const float rn = rnorm[b];
    for (int j = threadIdx.x; j < dl; j += blockDim.x)
        draw[j] = rn * (dlat[(long long)b * dl + j] - lat[(long long)b * dl + j] * coef);
    __syncthreads();
    const int ic = threadIdx.x & 63, q = threadIdx.x >> 6;
    const int i = blockIdx.x * 64 + ic;
    float acc = 0.f;
    if (i < din)
        for (int j = q; j < dl; j += 4) acc = fmaf(draw[j], __ldg(W + (long long)j * din + i), acc);
    part[q][ic] = acc;
    __syncthreads();
    if (q == 0 && i < din) dx[(long long)b * dx_stride + i] = (part[0][ic] + part[1][ic]) + (part[2][ic] + part[3][ic]);
}

// dW[j, i] = sum_b draw[b, j] x[b, i]; one CTA per 8 output rows j and `cols` columns i (grid.y; one slab when
// din <= LAT_DW_COLS, 72 of them for the 294 912-wide projection of forward_old)
constexpr int LAT_DW_COLS = 4096;
__global__ void latent_bwd_dw_kernel(const float* __restrict__ dlat, const float* __restrict__ lat,
                                     const float* __restrict__ rnorm, const float* __restrict__ x,
                                     long long x_stride, float* __restrict__ dW, int B, int din,
                                     int dl, int cols) {
    extern __shared__ float sm[];
    float* coef = sm;                 // [B]
    float* draw = sm + B;             // [8][B]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int b = warp; b < B; b += nwarp) {
        float dot = 0.f;
        for (int j = lane; j < dl; j += 32)
            dot += lat[(long long)b * dl + j] * dlat[(long long)b * dl + j];
        dot = warp_sum(dot);
        if (lane == 0) coef[b] = dot;
    }
    __syncthreads();
    const int j0 = blockIdx.x * 8;
    for (int t = threadIdx.x; t < 8 * B; t += blockDim.x) {
        const int jj = t / B, b = t % B;
        const int j = j0 + jj;
        draw[jj * B + b] = j < dl ? rnorm[b] * (dlat[(long long)b * dl + j] -
                                               lat[(long long)b * dl + j] * coef[b])
                                  : 0.f;
    }
    __syncthreads();
    const int i_end = min(din, (int)(blockIdx.y + 1) * cols);
    for (int i = blockIdx.y * cols + threadIdx.x; i < i_end; i += blockDim.x) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int b = 0; b < B; ++b) {
            const float xv = x[(long long)b * x_stride + i];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) acc[jj] = fmaf(draw[jj * B + b], xv, acc[jj]);
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
            if (j0 + jj < dl) dW[(long long)(j0 + jj) * din + i] = acc[jj];
    }
}

// ---------------------------------------------------------------- contrastive loss
struct Stat { float m, l, w; };     // online max, sum exp(x-m), sum x exp(x-m)

// pass 1: grid (col blocks, row blocks); S = s * T I^T
__global__ void __launch_bounds__(256)
clip_tile_stats_kernel(const float* __restrict__ T, const float* __restrict__ I,
                       const float* __restrict__ log_temp, float* __restrict__ rowpart,
                       float* __restrict__ colpart, float* __restrict__ diag, int N, int d) {
    __shared__ float As[16][68];
    __shared__ float Bs[16][68];
    __shared__ float St[64][65];
    const int cb = blockIdx.x, rb = blockIdx.y;
    const int m0 = rb * 64, n0 = cb * 64;
    float acc[4][4];
    sgemm_tile_64x64<true>(T, d, I, d, m0, n0, N, N, d, acc, As, Bs);
    const float s = expf(*log_temp);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) St[ty * 4 + i][tx * 4 + j] = acc[i][j] * s;
    __syncthreads();
    const int t = threadIdx.x;
    if (t < 128) {
        const bool is_row = t < 64;
        const int idx = t & 63;
        const int g = (is_row ? m0 : n0) + idx;
        if (g < N) {
            const int lim = min(64, N - (is_row ? n0 : m0));
            float m = -INFINITY;
            for (int q = 0; q < lim; ++q) m = fmaxf(m, is_row ? St[idx][q] : St[q][idx]);
            float l = 0.f, w = 0.f;
            for (int q = 0; q < lim; ++q) {
                const float v = is_row ? St[idx][q] : St[q][idx];
                const float e = expf(v - m);
                l += e;
                w = fmaf(e, v, w);
            }
            float* dst = is_row ? rowpart + ((long long)cb * N + g) * 3
                                : colpart + ((long long)rb * N + g) * 3;
            dst[0] = m; dst[1] = l; dst[2] = w;
            if (is_row && rb == cb) diag[g] = St[idx][idx];
        }
    }
}

// pass 2: thread per index; merges partials; accumulates loss and dlog_temp into out[0..1]
__global__ void clip_reduce_kernel(const float* __restrict__ rowpart,
                                   const float* __restrict__ colpart,
                                   const float* __restrict__ diag, float* __restrict__ row_lse,
                                   float* __restrict__ col_lse, float* __restrict__ out, int N,
                                   int nblk, float inv_2nb) {
    __shared__ float red[2][8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float loss = 0.f, dtemp = 0.f;
    if (i < N) {
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const float* part = side == 0 ? rowpart : colpart;
            float m = -INFINITY;
            for (int b = 0; b < nblk; ++b) m = fmaxf(m, part[((long long)b * N + i) * 3]);
            float l = 0.f, w = 0.f;
            for (int b = 0; b < nblk; ++b) {
                const float* p = part + ((long long)b * N + i) * 3;
                const float sc = expf(p[0] - m);
                l = fmaf(p[1], sc, l);
                w = fmaf(p[2], sc, w);
            }
            const float lse = m + logf(l);
            (side == 0 ? row_lse : col_lse)[i] = lse;
            loss += lse - diag[i];
            dtemp += w / l - diag[i];
        }
    }
    loss = warp_sum(loss);
    dtemp = warp_sum(dtemp);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = loss; red[1][warp] = dtemp; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[0][w]; b += red[1][w]; }
        atomicAdd(out + 0, a * inv_2nb);
        atomicAdd(out + 1, b * inv_2nb);
    }
}

// pass 3: grid (N/64, b_local/64, 2). z = 0: rows = local text rows, cols = all images,
// Gout[(r-row0), c] = G[r, c].  z = 1: rows = local image cols, cols = all texts,
// Gout[(c-row0), i] = G[i, c]  (second stripe lives at Gout + b_local*N).
__global__ void __launch_bounds__(256)
clip_grad_tiles_kernel(const float* __restrict__ T, const float* __restrict__ I,
                       const float* __restrict__ log_temp, const float* __restrict__ row_lse,
                       const float* __restrict__ col_lse, float* __restrict__ Gout, int N, int d,
                       int b_local, int row0, float inv_2nb) {
    __shared__ float As[16][68];
    __shared__ float Bs[16][68];
    const int z = blockIdx.z;
    const float* A = (z == 0 ? T : I) + (long long)row0 * d;
    const float* B = z == 0 ? I : T;
    const float* lse_a = (z == 0 ? row_lse : col_lse) + row0;   // lse of the stripe's own index
    const float* lse_b = z == 0 ? col_lse : row_lse;            // lse along the other index
    float* G = Gout + (long long)z * b_local * N;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4];
    sgemm_tile_64x64<true>(A, d, B, d, m0, n0, b_local, N, d, acc, As, Bs);
    const float s = expf(*log_temp);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = m0 + ty * 4 + i;
        if (r >= b_local) continue;
        const float la = lse_a[r];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + tx * 4 + j;
            if (c >= N) continue;
            const float v = acc[i][j] * s;
            float g = expf(v - la) + expf(v - lse_b[c]);
            if (c == row0 + r) g -= 2.f;
            G[(long long)r * N + c] = g * inv_2nb;
        }
    }
}

// pass 4: C[M, Nc] = alpha * A[M, K] * B[K, Nc]   (fp32, NN)
__global__ void __launch_bounds__(256)
sgemm_nn_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                const float* __restrict__ log_alpha, int M, int Nc, int K) {
    __shared__ float As[16][68];
    __shared__ float Bs[16][68];
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4];
    sgemm_tile_64x64<false>(A, K, B, Nc, m0, n0, M, Nc, K, acc, As, Bs);
    const float alpha = log_alpha ? expf(*log_alpha) : 1.f;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = m0 + ty * 4 + i;
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + tx * 4 + j;
            if (c < Nc) C[(long long)r * Nc + c] = acc[i][j] * alpha;
        }
    }
}

__global__ void pair_logits_kernel(const float* __restrict__ tl, const float* __restrict__ il,
                                   const float* __restrict__ log_temp, float* __restrict__ out,
                                   int P, int d) {
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= P) return;
    float acc = 0.f;
    for (int i = lane; i < d; i += 32) acc = fmaf(tl[(long long)p * d + i], il[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[p] = acc * expf(*log_temp);
}

// Tensor-core path (N >= 1024): fp32 latents as bf16 pairs x = hi + lo (16 mantissa bits).  The three products a
// split-precision dot needs are laid side by side along K so that ONE bf16 GEMM with K = 3d evaluates
// hi.hi + hi.lo + lo.hi:   text rows [hi | hi | lo],  image rows [hi | lo | hi].
__global__ void clip_split_kernel(const float* __restrict__ X, __nv_bfloat16* __restrict__ cat, long long n, int d,
                                  int image_pattern) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * d) return;
    const long long r = idx / d;
    const int c = (int)(idx % d);
    const float x = X[idx];
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    __nv_bfloat16* row = cat + r * 3 * d;
    row[c] = hi;
    row[d + c] = image_pattern ? lo : hi;
    row[2 * d + c] = image_pattern ? hi : lo;
}

inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

}  // namespace

extern "C" int ctk_mean_pool_fwd(const float* x, float* pooled, int B, long long n, int dim,
                                 void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(x && pooled && B > 0 && n > 0 && dim > 0 && dim % 4 == 0, CTK_ERR_SHAPE,
                "mean_pool: bad args");
    CTK_REQUIRE(CTK_ALIGNED(x, 16), CTK_ERR_ALIGN, "mean_pool: x must be 16-byte aligned");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (n <= 256 && dim >= 2048 && CTK_ALIGNED(pooled, 16)) {      // short axis, wide rows: CTCLIP.forward_old's pooling
        const long long dim4 = dim / 4;
        mean_frames_kernel<<<dim3((unsigned)((dim4 + 255) / 256), B), 256, 0, s>>>(x, pooled, (int)n, dim4);
        CTK_LAUNCH_CHECK();
        return CTK_OK;
    }
    CTK_CUDA(cudaMemsetAsync(pooled, 0, sizeof(float) * (size_t)B * dim, s));
    const long long tpb = 128;
    dim3 grid((unsigned)((n + tpb - 1) / tpb), B);
    mean_pool_kernel<<<grid, 128, 0, s>>>(x, pooled, n, dim, tpb);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

extern "C" int ctk_latent_fwd(const float* x, long long x_stride, const float* W, float* latent,
                              float* rnorm, int B, int din, int dl, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(x && W && latent && rnorm && B > 0 && din > 0 && dl > 0, CTK_ERR_SHAPE,
                "latent_fwd: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (din <= 1024) {
        const size_t sm = sizeof(float) * (size_t)din * (size_t)(B < LAT_MAX_B ? B : LAT_MAX_B);
        latent_raw_kernel<<<(dl + 7) / 8, 256, sm, s>>>(x, x_stride, W, latent, B, din, dl);
    } else {
        CTK_REQUIRE(din % 4 == 0 && x_stride % 4 == 0 && CTK_ALIGNED(x, 16) && CTK_ALIGNED(W, 16), CTK_ERR_ALIGN,
                    "latent_fwd: din %d > 1024 needs din and the row stride to be multiples of 4 and 16-byte aligned x, W", din);
        CTK_SET_MAX_SMEM(latent_raw_wide_kernel, LATW_SMEM);
        latent_raw_wide_kernel<<<(dl + LATW_ROWS - 1) / LATW_ROWS, 256, LATW_SMEM, s>>>(x, x_stride, W, latent, B, din, dl);
    }
    CTK_LAUNCH_CHECK();
    latent_norm_kernel<<<B, 256, 0, s>>>(latent, rnorm, dl);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

extern "C" int ctk_latent_bwd(const float* dlatent, const float* latent, const float* rnorm,
                              const float* x, long long x_stride, const float* W, float* dW,
                              float* dx, long long dx_stride, int B, int din, int dl,
                              void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(dlatent && latent && rnorm && x && W && B > 0 && din > 0 && dl > 0, CTK_ERR_SHAPE,
                "latent_bwd: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (dx) {
        latent_bwd_dx_kernel<<<dim3((din + 63) / 64, B), 256, sizeof(float) * dl, s>>>(dlatent, latent, rnorm, W, dx,
                                                                                       dx_stride, din, dl);
        CTK_LAUNCH_CHECK();
    }
    if (dW) {
        const size_t sm = sizeof(float) * (size_t)(9 * B);
        CTK_REQUIRE(sm <= 48 * 1024, CTK_ERR_SHAPE, "latent_bwd: batch too large");
        const int cols = din <= LAT_DW_COLS ? din : LAT_DW_COLS;
        latent_bwd_dw_kernel<<<dim3((dl + 7) / 8, (din + cols - 1) / cols), 256, sm, s>>>(dlatent, latent, rnorm, x,
                                                                                          x_stride, dW, B, din, dl, cols);
        CTK_LAUNCH_CHECK();
    }
    return CTK_OK;
}

extern "C" size_t ctk_clip_loss_ws_bytes(int N, int b_local) {
    if (N <= 0 || b_local <= 0) return 0;
    const size_t nblk = (size_t)(N + 63) / 64;
    size_t bytes = 0;
    bytes += align256(sizeof(float) * nblk * N * 3) * 2;     // row / col partials
    bytes += align256(sizeof(float) * N) * 3;                // diag, row_lse, col_lse
    bytes += align256(sizeof(float) * (size_t)b_local * N * 2);   // G stripes (fp32, or bf16 hi/lo pairs)
    if (N >= CLIP_TC_MIN_N)
        bytes += align256((size_t)N * 3 * CLIP_TC_MAX_D * 2) * 2;     // tensor-core path: split latents [N, 3d] bf16 x 2
    return bytes;
}

namespace {

// tcgen05 split-bf16 path for N >= 1024 (validated on hardware in round 2: tests/test_clip_loss_tc_gpu.py; 0.60 ms
// against 3.58 ms for the fp32 SIMT tiles at N = 4096); CTK_CLIP_LOSS_TC=0 forces the SIMT tiles for every N
bool clip_tc_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("CTK_CLIP_LOSS_TC");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// Same four passes as the SIMT path with the three contractions on ctk_gemm_bf16:
//   pass 1  two GEMMs (rows = texts, rows = images), K = 3d, epilogue LSE_PART: per-row online statistics over
//           128-column blocks + the diagonal; the logits never leave the SM.
//   pass 2  clip_reduce_kernel (shared with the SIMT path).
//   pass 3  two GEMMs [N, b_local] with epilogue CLIP_GRAD: the stripes of s * dloss/dS this rank needs, stored
//           k-index-major as bf16 hi/lo pairs.
//   pass 4  dT_local = G0^T I, dI_local = G1^T T: three MN-major split-K products each (hi.hi, hi.lo, lo.hi)
//           accumulated in fp32 (the weight-gradient form of ctk_gemm_bf16).
int clip_loss_tc(const float* T, const float* I, const float* log_temp, float* out, float* d_local, uint8_t* p, int N,
                 int d, int b_local, int row0, cudaStream_t s) {
    void* stream = reinterpret_cast<void*>(s);
    const int nblk = (N + 127) / 128;
    const size_t nblk64 = (size_t)(N + 63) / 64;             // the workspace is carved exactly as ctk_clip_loss_ws_bytes sizes it
    float* rowpart = reinterpret_cast<float*>(p); p += align256(sizeof(float) * nblk64 * N * 3);
    float* colpart = reinterpret_cast<float*>(p); p += align256(sizeof(float) * nblk64 * N * 3);
    float* diag = reinterpret_cast<float*>(p);    p += align256(sizeof(float) * N);
    float* row_lse = reinterpret_cast<float*>(p); p += align256(sizeof(float) * N);
    float* col_lse = reinterpret_cast<float*>(p); p += align256(sizeof(float) * N);
    __nv_bfloat16* G = reinterpret_cast<__nv_bfloat16*>(p); p += align256(sizeof(float) * (size_t)b_local * N * 2);
    __nv_bfloat16* catT = reinterpret_cast<__nv_bfloat16*>(p); p += align256((size_t)N * 3 * CLIP_TC_MAX_D * 2);
    __nv_bfloat16* catI = reinterpret_cast<__nv_bfloat16*>(p);
    const long long stripe = (long long)N * b_local;         // elements of one [N, b_local] bf16 stripe
    __nv_bfloat16 *g0_hi = G, *g0_lo = G + stripe, *g1_hi = G + 2 * stripe, *g1_lo = G + 3 * stripe;

    const float inv_2nb = 1.0f / (2.0f * (float)N * (float)b_local);
    CTK_CUDA(cudaMemsetAsync(out, 0, 2 * sizeof(float), s));
    const long long nel = (long long)N * d;
    clip_split_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, s>>>(T, catT, N, d, 0);
    CTK_LAUNCH_CHECK();
    clip_split_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, s>>>(I, catI, N, d, 1);
    CTK_LAUNCH_CHECK();

    ctk_gemm_epilogue_t e;
    memset(&e, 0, sizeof(e));
    e.alpha = 1.f;
    e.vec1 = log_temp;
    // ---- pass 1
    e.C = rowpart; e.aux0 = diag;
    int rc = ctk_gemm_bf16(catT, 3LL * d, 0, catI, 3LL * d, 0, N, N, 3 * d, CTK_EPI_LSE_PART, &e, 0, stream);
    if (rc) return rc;
    e.C = colpart; e.aux0 = nullptr;
    rc = ctk_gemm_bf16(catI, 3LL * d, 0, catT, 3LL * d, 0, N, N, 3 * d, CTK_EPI_LSE_PART, &e, 0, stream);
    if (rc) return rc;
    // ---- pass 2
    clip_reduce_kernel<<<(N + 127) / 128, 128, 0, s>>>(rowpart, colpart, diag, row_lse, col_lse, out, N, nblk, inv_2nb);
    CTK_LAUNCH_CHECK();
    if (!d_local) return CTK_OK;
    // ---- pass 3: G0[c, r] = g(text row0+r, image c);  G1[i, c] = g(text i, image row0+c);  both [N, b_local]
    e.alpha = inv_2nb;
    e.i0 = 0; e.i1 = row0;                                   // diagonal: gemm row == gemm col + row0
    e.ldc = b_local; e.ld_aux0 = b_local;
    e.C = g0_hi; e.aux0 = g0_lo; e.vec0 = col_lse; e.bias = row_lse + row0;
    rc = ctk_gemm_bf16(catI, 3LL * d, 0, catT + (long long)row0 * 3 * d, 3LL * d, 0, N, b_local, 3 * d,
                       CTK_EPI_CLIP_GRAD, &e, 0, stream);
    if (rc) return rc;
    e.C = g1_hi; e.aux0 = g1_lo; e.vec0 = row_lse; e.bias = col_lse + row0;
    rc = ctk_gemm_bf16(catT, 3LL * d, 0, catI + (long long)row0 * 3 * d, 3LL * d, 0, N, b_local, 3 * d,
                       CTK_EPI_CLIP_GRAD, &e, 0, stream);
    if (rc) return rc;
    // ---- pass 4: contraction over the N gathered samples, operands [K = N, rows] (MN-major), fp32 atomics
    CTK_CUDA(cudaMemsetAsync(d_local, 0, sizeof(float) * 2 * (size_t)b_local * d, s));
    memset(&e, 0, sizeof(e));
    e.alpha = 1.f;
    e.ldc = d;
    const __nv_bfloat16 *i_hi = catI, *i_lo = catI + d, *t_hi = catT, *t_lo = catT + 2 * d;
    const __nv_bfloat16* ga[2][2] = {{g0_hi, g0_lo}, {g1_hi, g1_lo}};
    const __nv_bfloat16* xb[2][2] = {{i_hi, i_lo}, {t_hi, t_lo}};
    for (int z = 0; z < 2; ++z) {
        e.C = d_local + (long long)z * b_local * d;
        const int term[3][2] = {{0, 0}, {0, 1}, {1, 0}};     // (G part, latent part): hi.hi, hi.lo, lo.hi
        for (int t = 0; t < 3; ++t) {
            rc = ctk_gemm_bf16(ga[z][term[t][0]], b_local, 1, xb[z][term[t][1]], 3LL * d, 1, b_local, d, N,
                               CTK_EPI_ATOMIC_F32, &e, 0, stream);
            if (rc) return rc;
        }
    }
    return CTK_OK;
}

}  // namespace

extern "C" int ctk_clip_loss_fwd_bwd(const float* T, const float* I, const float* log_temp,
                                     float* out, float* d_local, void* ws, size_t ws_bytes, int N,
                                     int d, int b_local, int row0, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(T && I && log_temp && out && ws, CTK_ERR_SHAPE, "clip_loss: null pointer");
    CTK_REQUIRE(N > 0 && d > 0 && b_local > 0 && row0 >= 0 && row0 + b_local <= N, CTK_ERR_SHAPE,
                "clip_loss: bad N=%d d=%d b_local=%d row0=%d", N, d, b_local, row0);
    CTK_REQUIRE(ws_bytes >= ctk_clip_loss_ws_bytes(N, b_local), CTK_ERR_SHAPE,
                "clip_loss: workspace too small");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (clip_tc_enabled() && N >= CLIP_TC_MIN_N && N % 128 == 0 && d % 64 == 0 && d <= CLIP_TC_MAX_D && b_local % 32 == 0 &&
        CTK_ALIGNED(ws, 256))
        return clip_loss_tc(T, I, log_temp, out, d_local, reinterpret_cast<uint8_t*>(ws), N, d, b_local, row0, s);
    const int nblk = (N + 63) / 64;
    uint8_t* p = reinterpret_cast<uint8_t*>(ws);
    float* rowpart = reinterpret_cast<float*>(p); p += align256(sizeof(float) * (size_t)nblk * N * 3);
    float* colpart = reinterpret_cast<float*>(p); p += align256(sizeof(float) * (size_t)nblk * N * 3);
    float* diag = reinterpret_cast<float*>(p);    p += align256(sizeof(float) * N);
    float* row_lse = reinterpret_cast<float*>(p); p += align256(sizeof(float) * N);
    float* col_lse = reinterpret_cast<float*>(p); p += align256(sizeof(float) * N);
    float* G = reinterpret_cast<float*>(p);

    // reference: / 2 (two directions) / bs_single_gpu (ct_clip.py:1379), means over N rows
    const float inv_2nb = 1.0f / (2.0f * (float)N * (float)b_local);
    CTK_CUDA(cudaMemsetAsync(out, 0, 2 * sizeof(float), s));
    clip_tile_stats_kernel<<<dim3(nblk, nblk), 256, 0, s>>>(T, I, log_temp, rowpart, colpart, diag,
                                                           N, d);
    CTK_LAUNCH_CHECK();
    clip_reduce_kernel<<<(N + 127) / 128, 128, 0, s>>>(rowpart, colpart, diag, row_lse, col_lse,
                                                      out, N, nblk, inv_2nb);
    CTK_LAUNCH_CHECK();
    if (d_local) {
        clip_grad_tiles_kernel<<<dim3(nblk, (b_local + 63) / 64, 2), 256, 0, s>>>(
            T, I, log_temp, row_lse, col_lse, G, N, d, b_local, row0, inv_2nb);
        CTK_LAUNCH_CHECK();
        dim3 g2((d + 63) / 64, (b_local + 63) / 64);
        sgemm_nn_kernel<<<g2, 256, 0, s>>>(G, I, d_local, log_temp, b_local, d, N);
        CTK_LAUNCH_CHECK();
        sgemm_nn_kernel<<<g2, 256, 0, s>>>(G + (long long)b_local * N, T,
                                           d_local + (long long)b_local * d, log_temp, b_local, d, N);
        CTK_LAUNCH_CHECK();
    }
    return CTK_OK;
}

extern "C" int ctk_pair_logits(const float* text_lat, const float* image_lat, const float* log_temp,
                               float* out, int P, int d, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(text_lat && image_lat && log_temp && out && P > 0 && d > 0, CTK_ERR_SHAPE,
                "pair_logits: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    pair_logits_kernel<<<(P + 3) / 4, 128, 0, s>>>(text_lat, image_lat, log_temp, out, P, d);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

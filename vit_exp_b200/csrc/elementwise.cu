// HBM-bound kernels of the CTViT encoder: tubelet patch gather + LayerNorm statistics,
// LayerNorm forward/backward (with the spatial<->temporal token transposition folded into the
// row index), PEG depthwise 3x3x3 conv forward/backward, row l2norm, VQ gather / EMA update.
#include "common.cuh"
#include <stdlib.h>
#include <type_traits>

namespace {

// =============================================================================================
// patch gather + LayerNorm(pt*p1*p2) statistics   (ctvit.py:170-172)
// One CTA = G consecutive patches along W: (pt*p1) contiguous runs of G*p2 floats.
// =============================================================================================
template <int G>
__global__ void __launch_bounds__(256)
patch_norm_kernel(const float* __restrict__ video, __nv_bfloat16* __restrict__ xhat, long long ld,
                  float* __restrict__ mean_out, float* __restrict__ rstd_out, int D, int H, int W,
                  int pt, int p1, int p2, float eps) {
    extern __shared__ float sm[];                 // [G][K]
    __shared__ float red[G][8];
    __shared__ float stat[G][2];
    const int K = pt * p1 * p2;
    const int T = D / pt, Hp = H / p1, Wp = W / p2;
    const int wg = Wp / G;
    long long cta = blockIdx.x;
    const int wgi = (int)(cta % wg); cta /= wg;
    const int hp = (int)(cta % Hp);  cta /= Hp;
    const int tp = (int)(cta % T);   cta /= T;
    const int b = (int)cta;
    const int runw = G * p2;                      // floats per contiguous run
    const int nrun = pt * p1;
    const float* base = video + (((long long)b * D + (long long)tp * pt) * H + (long long)hp * p1) * W +
                        (long long)wgi * runw;
    const bool vec = (p2 % 4 == 0) && (W % 4 == 0);
    if (vec) {
        const int q_per_run = runw / 4;
        for (int i = threadIdx.x; i < nrun * q_per_run; i += blockDim.x) {
            const int r = i / q_per_run, q = i % q_per_run;
            const int dt = r / p1, dy = r % p1;
            const float4 v = __ldg(reinterpret_cast<const float4*>(base + ((long long)dt * H + dy) * W) + q);
            const int x = q * 4;
            const int g = x / p2, dx = x % p2;
            float* dst = sm + g * K + r * p2 + dx;
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
    } else {
        for (int i = threadIdx.x; i < nrun * runw; i += blockDim.x) {
            const int r = i / runw, x = i % runw;
            const int dt = r / p1, dy = r % p1;
            sm[(x / p2) * K + r * p2 + (x % p2)] = __ldg(base + ((long long)dt * H + dy) * W + x);
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // pass 1: means
    float s[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        float a = 0.f;
        for (int k = threadIdx.x; k < K; k += blockDim.x) a += sm[g * K + k];
        s[g] = warp_sum(a);
    }
    if (lane == 0)
#pragma unroll
        for (int g = 0; g < G; ++g) red[g][warp] = s[g];
    __syncthreads();
    if (threadIdx.x < G) {
        float a = 0.f;
        for (int w = 0; w < 8; ++w) a += red[threadIdx.x][w];
        stat[threadIdx.x][0] = a / (float)K;
    }
    __syncthreads();
    // pass 2: centred variance (exact two-pass: padding voxels are -1 while data is in [0,1])
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const float mu = stat[g][0];
        float a = 0.f;
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const float d = sm[g * K + k] - mu;
            a = fmaf(d, d, a);
        }
        s[g] = warp_sum(a);
    }
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int g = 0; g < G; ++g) red[g][warp] = s[g];
    __syncthreads();
    const long long row0 = (((long long)b * T + tp) * Hp + hp) * Wp + (long long)wgi * G;
    if (threadIdx.x < G) {
        float a = 0.f;
        for (int w = 0; w < 8; ++w) a += red[threadIdx.x][w];
        const float rs = rsqrtf(a / (float)K + eps);
        stat[threadIdx.x][1] = rs;
        mean_out[row0 + threadIdx.x] = stat[threadIdx.x][0];
        rstd_out[row0 + threadIdx.x] = rs;
    }
    __syncthreads();
    // write xhat (bf16), 8 elements (16 B) per thread-iteration; pad columns [K, ld) zeroed
    const int k8 = (int)(ld / 8);
    for (int i = threadIdx.x; i < G * k8; i += blockDim.x) {
        const int g = i / k8, j = (i % k8) * 8;
        const float mu = stat[g][0], rs = stat[g][1];
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (j + e < K) ? (sm[g * K + j + e] - mu) * rs : 0.f;
        uint4 u;
        u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
        u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(xhat + (row0 + g) * ld + j) = u;
    }
}

// TMA variant for the production patch geometry: the G patches of a CTA are ONE 3-D TMA box
// {G*P2, P1, PT} of the fp32 volume (64 KB for 4 patches of 20x20x10), landed in shared memory in
// exactly the (pt p1 p2) K-order of ctvit.py:171; statistics and the bf16 write-out then run out
// of shared memory with compile-time index arithmetic.
template <int PT, int P1, int P2, int G>
__global__ void __launch_bounds__(256)
patch_norm_tma_kernel(const __grid_constant__ CUtensorMap tmap, __nv_bfloat16* __restrict__ xhat, long long ld,
                      float* __restrict__ mean_out, float* __restrict__ rstd_out, int D, int H, int W, float eps) {
    constexpr int K = PT * P1 * P2, RUNW = G * P2, NRUN = PT * P1;
    extern __shared__ __align__(128) float sm[];          // [NRUN][RUNW]
    __shared__ __align__(8) uint64_t bar;
    __shared__ float red[8][G];
    __shared__ float stat[G][2];
    const int T = D / PT, Hp = H / P1, Wp = W / P2;
    const int wg = Wp / G;
    long long cta = blockIdx.x;
    const int wgi = (int)(cta % wg); cta /= wg;
    const int hp = (int)(cta % Hp);  cta /= Hp;
    const int tp = (int)(cta % T);   cta /= T;
    const int b = (int)cta;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, NRUN * RUNW * 4);
        tma_load_3d(sm, &tmap, &bar, wgi * RUNW, hp * P1, b * D + tp * PT);
    }
    mbar_wait(&bar, 0);
    // thread <-> column x of the box (fixed patch g = x / P2), strided over the NRUN rows
    constexpr int TPR = 256 / RUNW;                         // row lanes per column
    const int x = tid % RUNW, r0 = tid / RUNW;
    const bool act = tid < TPR * RUNW;
    const int g = x / P2;
    for (int pass = 0; pass < 2; ++pass) {
        if (tid < 8 * G) red[tid / G][tid % G] = 0.f;
        __syncthreads();
        if (act) {
            float a = 0.f;
            if (pass == 0) {
                for (int r = r0; r < NRUN; r += TPR) a += sm[r * RUNW + x];
            } else {
                const float mu = stat[g][0];
                for (int r = r0; r < NRUN; r += TPR) { const float d = sm[r * RUNW + x] - mu; a = fmaf(d, d, a); }
            }
            atomicAdd(&red[warp][g], a);
        }
        __syncthreads();
        if (tid < G) {
            float a = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) a += red[w][tid];
            if (pass == 0) stat[tid][0] = a / (float)K;
            else stat[tid][1] = rsqrtf(a / (float)K + eps);
        }
        __syncthreads();
    }
    const long long row0 = (((long long)b * T + tp) * Hp + hp) * Wp + (long long)wgi * G;
    if (tid < G) { mean_out[row0 + tid] = stat[tid][0]; rstd_out[row0 + tid] = stat[tid][1]; }
    // write xhat (bf16): 4 elements (8 B) per thread-iteration, consecutive threads -> consecutive k
    constexpr int Q = K / 4;
    static_assert(P2 % 4 == 0, "P2 must be a multiple of 4");
    for (int i = tid; i < G * Q; i += 256) {
        const int gg = i / Q, q = i % Q;
        const int r = q / (P2 / 4), dx = (q % (P2 / 4)) * 4;
        const float4 v = *reinterpret_cast<const float4*>(sm + r * RUNW + gg * P2 + dx);
        const float mu = stat[gg][0], rs = stat[gg][1];
        uint2 u;
        u.x = pack_bf16x2((v.x - mu) * rs, (v.y - mu) * rs);
        u.y = pack_bf16x2((v.z - mu) * rs, (v.w - mu) * rs);
        *reinterpret_cast<uint2*>(xhat + (row0 + gg) * ld + 4 * q) = u;
    }
    // pad columns [K, ld)
    for (int i = tid; i < G * (int)(ld - K); i += 256)
        xhat[(row0 + i / (int)(ld - K)) * ld + K + i % (int)(ld - K)] = __float2bfloat16(0.f);
}

// =============================================================================================
// LayerNorm forward: one warp per row, NV float2 pairs per lane (dim = 64 * NV)
// =============================================================================================
__device__ __forceinline__ long long perm_row(long long row, int outer, int inner) {
    if (inner <= 0) return row;
    const long long i = row % inner;
    const long long t = row / inner;
    const long long o = t % outer, g = t / outer;
    return (g * inner + i) * outer + o;
}

template <int NV>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, __nv_bfloat16* __restrict__ out_bf16,
                     float* __restrict__ out_f32, __nv_bfloat16* __restrict__ xraw_bf16,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out, long long rows,
                     float eps, int perm_outer, int perm_inner) {
    constexpr int DIM = NV * 64;
    const int lane = threadIdx.x & 31;
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    float2 gm[NV], bt[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        gm[j] = __ldg(reinterpret_cast<const float2*>(gamma) + j * 32 + lane);
        bt[j] = beta ? __ldg(reinterpret_cast<const float2*>(beta) + j * 32 + lane)
                     : make_float2(0.f, 0.f);
    }
    for (long long row = warp_global; row < rows; row += nwarps) {
        const float2* xr = reinterpret_cast<const float2*>(x + row * DIM);
        float2 v[NV];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            v[j] = xr[j * 32 + lane];
            s += v[j].x + v[j].y;
        }
        const float mu = warp_sum(s) * (1.0f / DIM);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const float a = v[j].x - mu, b = v[j].y - mu;
            q = fmaf(a, a, q);
            q = fmaf(b, b, q);
        }
        const float rs = rsqrtf(warp_sum(q) * (1.0f / DIM) + eps);
        if (lane == 0) {
            if (mean_out) mean_out[row] = mu;
            if (rstd_out) rstd_out[row] = rs;
        }
        const long long orow = perm_row(row, perm_outer, perm_inner);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const float a = (v[j].x - mu) * rs * gm[j].x + bt[j].x;
            const float b = (v[j].y - mu) * rs * gm[j].y + bt[j].y;
            if (out_bf16)
                reinterpret_cast<uint32_t*>(out_bf16 + orow * DIM)[j * 32 + lane] = pack_bf16x2(a, b);
            if (out_f32)
                reinterpret_cast<float2*>(out_f32 + orow * DIM)[j * 32 + lane] = make_float2(a, b);
            if (xraw_bf16)
                reinterpret_cast<uint32_t*>(xraw_bf16 + row * DIM)[j * 32 + lane] =
                    pack_bf16x2(v[j].x, v[j].y);
        }
    }
}

// LayerNorm backward. dgamma/dbeta: per-warp register partials -> smem -> one atomic per column
// per CTA.
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy_bf16, const float* __restrict__ dy_f32,
                     const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_bf16, int dx_accum,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows,
                     int perm_outer, int perm_inner, long long bcast_rows, float dy_scale) {
    constexpr int DIM = NV * 64;
    __shared__ float sg[8][DIM + 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    float2 gm[NV], dg[NV], db[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        gm[j] = __ldg(reinterpret_cast<const float2*>(gamma) + j * 32 + lane);
        dg[j] = make_float2(0.f, 0.f);
        db[j] = make_float2(0.f, 0.f);
    }
    for (long long row = warp_global; row < rows; row += nwarps) {
        const long long drow = bcast_rows > 0 ? row / bcast_rows : perm_row(row, perm_outer, perm_inner);
        const float mu = mean[row], rs = rstd[row];
        const float2* xr = reinterpret_cast<const float2*>(x + row * DIM);
        float2 xh[NV], g[NV], prev[NV];
        float2* dxr = reinterpret_cast<float2*>(dx + row * DIM);
        // the accumulation target is fetched with the inputs, not after the row reductions
#pragma unroll
        for (int j = 0; j < NV; ++j) prev[j] = dx_accum ? dxr[j * 32 + lane] : make_float2(0.f, 0.f);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            float2 d;
            if (dy_bf16) d = unpack_bf16x2(reinterpret_cast<const uint32_t*>(dy_bf16 + drow * DIM)[j * 32 + lane]);
            else d = reinterpret_cast<const float2*>(dy_f32 + drow * DIM)[j * 32 + lane];
            d.x *= dy_scale; d.y *= dy_scale;
            const float2 xv = xr[j * 32 + lane];
            xh[j] = make_float2((xv.x - mu) * rs, (xv.y - mu) * rs);
            dg[j].x = fmaf(d.x, xh[j].x, dg[j].x); dg[j].y = fmaf(d.y, xh[j].y, dg[j].y);
            db[j].x += d.x; db[j].y += d.y;
            g[j] = make_float2(d.x * gm[j].x, d.y * gm[j].y);
            s1 += g[j].x + g[j].y;
            s2 = fmaf(g[j].x, xh[j].x, s2);
            s2 = fmaf(g[j].y, xh[j].y, s2);
        }
        s1 = warp_sum(s1) * (1.0f / DIM);
        s2 = warp_sum(s2) * (1.0f / DIM);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const float2 o = make_float2(rs * (g[j].x - s1 - xh[j].x * s2) + prev[j].x,
                                         rs * (g[j].y - s1 - xh[j].y * s2) + prev[j].y);
            dxr[j * 32 + lane] = o;
            if (dx_bf16)
                reinterpret_cast<uint32_t*>(dx_bf16 + row * DIM)[j * 32 + lane] = pack_bf16x2(o.x, o.y);
        }
    }
    // reduce dgamma / dbeta across the CTA's warps
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1 && !dbeta) break;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const float2 v = pass == 0 ? dg[j] : db[j];
            sg[warp][(j * 32 + lane) * 2] = v.x;
            sg[warp][(j * 32 + lane) * 2 + 1] = v.y;
        }
        __syncthreads();
        float* dst = pass == 0 ? dgamma : dbeta;
        for (int c = threadIdx.x; c < DIM; c += blockDim.x) {
            float a = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += sg[w][c];
            atomicAdd(dst + c, a);
        }
    }
}

// =============================================================================================
// PEG: depthwise causal 3x3x3 conv + bias + residual on [B, n0, n1, n2, dim] (attention.py:62-90,443)
//
// CTA = (batch b, tile of PEG_T1 rows along axis 1, slab of 32 channels). It walks axis 0 keeping a
// ring of 5 input planes in shared memory: 3 feed the current output plane while the TMA engine
// fills the other two for the next steps. Each plane is ONE 5-D TMA box {32 ch, n2+2, PEG_T1+2, 1, 1}
// whose out-of-range coordinates (halo rows/columns, planes before the causal start) are
// zero-filled by the hardware, so every input element is fetched once per CTA (1.5x halo
// overhead) instead of 9-27 times and there is no index arithmetic on the load path.
// lane <-> channel (a warp reads 128 contiguous bytes per token, bank-conflict free in smem),
// warp <-> (output row, half of axis 2); a 3-wide register window slides along axis 2.
// MODE 0: y = conv(x) + b + x          MODE 1: dx = conv^T(dy) + dy (flipped taps, planes a0..a0+2)
// MODE 2: dw, db accumulation (x planes in smem, dy read directly)
// =============================================================================================
constexpr int PEG_T1 = 4;
constexpr int PEG_RING = 5;            // 3 planes in use + 2 in flight
constexpr int PEG_CS = 32;

template <int MODE>
__global__ void __launch_bounds__(256, 2)
peg_tile_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ w,
                const float* __restrict__ b, const float* __restrict__ dy, float* __restrict__ y,
                __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ dw, float* __restrict__ db, int B,
                int n0, int n1, int n2, int dim) {
    extern __shared__ __align__(128) float psm[];
    __shared__ __align__(8) uint64_t full_bar[PEG_RING];
    const int W2 = n2 + 2;
    const int plane_floats = (PEG_T1 + 2) * W2 * PEG_CS;
    const uint32_t plane_bytes = (uint32_t)plane_floats * 4u;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles1 = (n1 + PEG_T1 - 1) / PEG_T1;
    const int slab = blockIdx.x;                                  // slabs of a tile run together (L2 reuse)
    const int bb = blockIdx.y / tiles1;
    const int r_lo = (blockIdx.y % tiles1) * PEG_T1;
    const int c0 = slab * PEG_CS;
    const int c = c0 + lane;
    const int lrow = warp % PEG_T1, half = warp / PEG_T1;        // 8 warps = 4 rows x 2 halves of axis 2
    const int a1 = r_lo + lrow;                                   // this warp's output row
    const bool row_ok = a1 < n1;
    const int p_lo = half * ((n2 + 1) / 2), p_hi = half == 0 ? min(n2, (n2 + 1) / 2) : n2;
    constexpr bool REV = MODE == 1;
    constexpr int shift = REV ? 0 : -2;                          // load n holds plane n + shift
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < PEG_RING; ++i) mbar_init(&full_bar[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int n) {                                     // thread 0 only
        const int sl = n % PEG_RING;
        mbar_expect_tx(&full_bar[sl], plane_bytes);
        tma_load_5d(psm + sl * plane_floats, &tmap, &full_bar[sl], c0, -1, r_lo - 1, n + shift, bb);
    };
    float wt[27];
    if (MODE != 2) {
#pragma unroll
        for (int t = 0; t < 27; ++t) wt[t] = __ldg(w + (long long)c * 27 + (REV ? 26 - t : t));
    }
    const float bias = (MODE == 0 && b) ? __ldg(b + c) : 0.f;
    float acc_w[27];
    float acc_b = 0.f;
    if (MODE == 2) {
#pragma unroll
        for (int t = 0; t < 27; ++t) acc_w[t] = 0.f;
    }
    if (tid == 0) {               // n0 + 2 loads in total; never leave one un-awaited
        issue(0); issue(1); issue(2);
        if (n0 >= 2) issue(3);
    }
    for (int a0 = 0; a0 < n0; ++a0) {
        // loads a0, a0+1, a0+2 feed this step; the first two were awaited by earlier steps
        if (a0 == 0) {
            mbar_wait(&full_bar[0], 0);
            mbar_wait(&full_bar[1], 0);
        }
        mbar_wait(&full_bar[(a0 + 2) % PEG_RING], ((a0 + 2) / PEG_RING) & 1);
        __syncthreads();          // every warp finished step a0-1 -> slot of load a0-1 is free
        if (tid == 0 && a0 + 2 < n0) issue(a0 + 4);
        if (!row_ok || p_lo >= p_hi) continue;
        // smem line (k0, k1): load a0+k0, local row lrow + k1   (local row 0 = a1 - 1)
        const float* ln[9];
#pragma unroll
        for (int k0 = 0; k0 < 3; ++k0) {
            const float* pl = psm + ((a0 + k0) % PEG_RING) * plane_floats;
#pragma unroll
            for (int k1 = 0; k1 < 3; ++k1) ln[k0 * 3 + k1] = pl + ((lrow + k1) * W2 + p_lo) * PEG_CS + lane;
        }
        float win[9][3];
#pragma unroll
        for (int l = 0; l < 9; ++l) { win[l][0] = ln[l][0]; win[l][1] = ln[l][PEG_CS]; }
        constexpr int centre = REV ? 1 : 7;                      // (k0,k1) of the un-shifted line
        const long long obase = ((((long long)bb * n0 + a0) * n1 + a1) * n2) * dim + c;
        // MODE 2: the warp's output gradients of this step are fetched up front (16 positions cover
        // n2 <= 32), so the walk below never waits on global memory
        constexpr int GMAX = 16;
        float gv[GMAX];
        if (MODE == 2) {
#pragma unroll
            for (int q = 0; q < GMAX; ++q) gv[q] = p_lo + q < p_hi ? dy[obase + (long long)(p_lo + q) * dim] : 0.f;
        }
        if (MODE == 2) {
            // fully unrolled walk: the 3-wide window rotates at compile time (no register moves)
#pragma unroll
            for (int q = 0; q < GMAX; ++q) {
                if (p_lo + q < p_hi) {
#pragma unroll
                    for (int l = 0; l < 9; ++l) win[l][(q + 2) % 3] = ln[l][(q + 2) * PEG_CS];
                    const float g = gv[q];
                    acc_b += g;
#pragma unroll
                    for (int l = 0; l < 9; ++l)
#pragma unroll
                        for (int k2 = 0; k2 < 3; ++k2)
                            acc_w[l * 3 + k2] = fmaf(g, win[l][(q + k2) % 3], acc_w[l * 3 + k2]);
                }
            }
            for (int a2 = p_lo + GMAX; a2 < p_hi; ++a2) {          // n2 > 32 only
                const float g = dy[obase + (long long)a2 * dim];
                acc_b += g;
#pragma unroll
                for (int l = 0; l < 9; ++l)
#pragma unroll
                    for (int k2 = 0; k2 < 3; ++k2)
                        acc_w[l * 3 + k2] = fmaf(g, ln[l][(a2 - p_lo + k2) * PEG_CS], acc_w[l * 3 + k2]);
            }
        } else {
            for (int a2 = p_lo; a2 < p_hi; ++a2) {
#pragma unroll
                for (int l = 0; l < 9; ++l) win[l][2] = ln[l][(a2 - p_lo + 2) * PEG_CS];
                float acc = bias + win[centre][1];               // + residual
#pragma unroll
                for (int l = 0; l < 9; ++l)
#pragma unroll
                    for (int k2 = 0; k2 < 3; ++k2) acc = fmaf(wt[l * 3 + k2], win[l][k2], acc);
                y[obase + (long long)a2 * dim] = acc;
                if (y_bf16) y_bf16[obase + (long long)a2 * dim] = __float2bfloat16(acc);
#pragma unroll
                for (int l = 0; l < 9; ++l) { win[l][0] = win[l][1]; win[l][1] = win[l][2]; }
            }
        }
    }
    if (MODE == 2) {
        // reduce the 8 warps of the CTA through shared memory, then one atomic per (c, tap)
        __syncthreads();
        float* red = psm;                                        // [8][28][32]
        const bool contributes = row_ok && p_lo < p_hi;
#pragma unroll
        for (int t = 0; t < 27; ++t) red[(warp * 28 + t) * 32 + lane] = contributes ? acc_w[t] : 0.f;
        red[(warp * 28 + 27) * 32 + lane] = contributes ? acc_b : 0.f;
        __syncthreads();
        for (int i = tid; i < 28 * 32; i += 256) {
            const int t = i / 32, l = i % 32;
            float sacc = 0.f;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) sacc += red[(ww * 28 + t) * 32 + l];
            if (t < 27) atomicAdd(dw + (long long)(c0 + l) * 27 + t, sacc);
            else atomicAdd(db + c0 + l, sacc);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// PEG forward / input gradient with packed fp32 math.  Same CTA tile, TMA plane ring and shared
// memory layout as peg_tile_kernel, different thread mapping: the kernel above is bound by
// instruction issue (27 FFMA + 9 LDS + stores per output), so here a thread owns TWO adjacent
// channels and uses fma.rn.f32x2 (FFMA2: two fp32 FMAs per issue slot) on 8-byte shared-memory
// loads.  lane = (row of a row pair, channel pair); warp = (row pair, quarter of axis 2).  The
// three input planes are walked one after the other over the warp's positions, partial sums in
// registers, so only a 3x3 window of float2 is live and the 27 weights are re-read per plane from
// a float2 copy in shared memory.
// MODE 0: y = conv(x) + b + x          MODE 1: dx = conv^T(dy) + dy
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

constexpr int PEG_QMAX = 8;            // positions per warp walk (ceil(n2 / 4) <= 8, i.e. n2 <= 32)

template <int MODE>
__global__ void __launch_bounds__(256, 2)
peg_pair_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ w, const float* __restrict__ b,
                float* __restrict__ y, __nv_bfloat16* __restrict__ y_bf16, int B, int n0, int n1, int n2, int dim) {
    extern __shared__ __align__(128) float psm[];
    __shared__ __align__(8) uint64_t full_bar[PEG_RING];
    __shared__ __align__(8) float2 swt[27][16];                  // weights of the slab, [tap][channel pair]
    const int W2 = n2 + 2;
    const int plane_floats = (PEG_T1 + 2) * W2 * PEG_CS;
    const uint32_t plane_bytes = (uint32_t)plane_floats * 4u;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles1 = (n1 + PEG_T1 - 1) / PEG_T1;
    const int slab = blockIdx.x;
    const int bb = blockIdx.y / tiles1;
    const int r_lo = (blockIdx.y % tiles1) * PEG_T1;
    const int c0 = slab * PEG_CS;
    const int cp = lane & 15;
    const int lrow = 2 * (warp & 1) + (lane >> 4);               // row of the 4-row tile
    const int a1 = r_lo + lrow;
    const bool row_ok = a1 < n1;
    const int qlen = (n2 + 3) / 4;
    const int p_lo = (warp >> 1) * qlen, p_hi = min(n2, p_lo + qlen);
    constexpr bool REV = MODE == 1;
    constexpr int shift = REV ? 0 : -2;
    constexpr int centre_k0 = REV ? 0 : 2;
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < PEG_RING; ++i) mbar_init(&full_bar[i], 1);
        mbar_fence_init();
    }
    for (int i = tid; i < 27 * 16; i += 256) {
        const int t = i / 16, q = i % 16;
        const int tt = REV ? 26 - t : t;
        swt[t][q] = make_float2(__ldg(w + (long long)(c0 + 2 * q) * 27 + tt), __ldg(w + (long long)(c0 + 2 * q + 1) * 27 + tt));
    }
    float2 bias = make_float2(0.f, 0.f);
    if (MODE == 0 && b) bias = make_float2(__ldg(b + c0 + 2 * cp), __ldg(b + c0 + 2 * cp + 1));
    __syncthreads();
    auto issue = [&](int n) {
        const int sl = n % PEG_RING;
        mbar_expect_tx(&full_bar[sl], plane_bytes);
        tma_load_5d(psm + sl * plane_floats, &tmap, &full_bar[sl], c0, -1, r_lo - 1, n + shift, bb);
    };
    if (tid == 0) {
        issue(0); issue(1); issue(2);
        if (n0 >= 2) issue(3);
    }
    for (int a0 = 0; a0 < n0; ++a0) {
        if (a0 == 0) {
            mbar_wait(&full_bar[0], 0);
            mbar_wait(&full_bar[1], 0);
        }
        mbar_wait(&full_bar[(a0 + 2) % PEG_RING], ((a0 + 2) / PEG_RING) & 1);
        __syncthreads();
        if (tid == 0 && a0 + 2 < n0) issue(a0 + 4);
        if (!row_ok || p_lo >= p_hi) continue;
        float2 acc[PEG_QMAX];
#pragma unroll
        for (int q = 0; q < PEG_QMAX; ++q) acc[q] = bias;
#pragma unroll
        for (int k0 = 0; k0 < 3; ++k0) {
            const float* pl = psm + ((a0 + k0) % PEG_RING) * plane_floats + (lrow * W2 + p_lo) * PEG_CS + 2 * cp;
            float2 wk[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) wk[t] = swt[k0 * 9 + t][cp];
            float2 win[3][3];
#pragma unroll
            for (int k1 = 0; k1 < 3; ++k1) {
                win[k1][0] = *reinterpret_cast<const float2*>(pl + (k1 * W2) * PEG_CS);
                win[k1][1] = *reinterpret_cast<const float2*>(pl + (k1 * W2 + 1) * PEG_CS);
            }
#pragma unroll
            for (int q = 0; q < PEG_QMAX; ++q) {
                if (p_lo + q < p_hi) {
#pragma unroll
                    for (int k1 = 0; k1 < 3; ++k1)
                        win[k1][(q + 2) % 3] = *reinterpret_cast<const float2*>(pl + (k1 * W2 + q + 2) * PEG_CS);
                    if (k0 == centre_k0) {                        // + residual (the un-shifted input)
                        const float2 ctr = win[1][(q + 1) % 3];
                        acc[q].x += ctr.x;
                        acc[q].y += ctr.y;
                    }
#pragma unroll
                    for (int k1 = 0; k1 < 3; ++k1)
#pragma unroll
                        for (int k2 = 0; k2 < 3; ++k2) acc[q] = ffma2(wk[k1 * 3 + k2], win[k1][(q + k2) % 3], acc[q]);
                }
            }
        }
        const long long obase = ((((long long)bb * n0 + a0) * n1 + a1) * n2 + p_lo) * dim + c0 + 2 * cp;
#pragma unroll
        for (int q = 0; q < PEG_QMAX; ++q) {
            if (p_lo + q < p_hi) {
                *reinterpret_cast<float2*>(y + obase + (long long)q * dim) = acc[q];
                if (y_bf16) *reinterpret_cast<uint32_t*>(y_bf16 + obase + (long long)q * dim) = pack_bf16x2(acc[q].x, acc[q].y);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// PEG forward / input gradient, plane-scatter form.  The gather form above reads 9 shared-memory
// values per output (3 planes x 3 rows, window along axis 2) and is bound by the shared-memory
// crossbar.  Here every input plane is read ONCE per thread (3 rows x (q + 2) columns for q output
// positions) and scattered into the partial sums of the three output planes it contributes to,
// which live in registers (3 x q float2) and rotate at compile time (the plane loop is unrolled by
// 3): 3 shared-memory reads feed 27 packed FMAs.  Planes are consumed once, so the ring only holds
// the current plane and the prefetched ones, and no out-of-range planes are ever loaded.
// MODE 0: y = conv(x) + b + x   (plane p feeds outputs p, p+1, p+2)
// MODE 1: dx = conv^T(dy) + dy  (plane p feeds outputs p, p-1, p-2; two outputs are flushed at the end)
// ---------------------------------------------------------------------------------------------
constexpr int PEGS_RING = 4;

template <int MODE, int QL, int R>
__device__ __forceinline__ void peg_scatter_plane(const float* pl, int W2, int ncol, const float2 (&wk)[27], float2 (&acc)[3][QL],
                                                  int p, int n0) {
    // output plane of tap k0 and its register slot: MODE 0: p + 2 - k0 -> (R + 2 - k0) % 3; MODE 1: p - k0 -> (R + 3 - k0) % 3
    bool use[3];
#pragma unroll
    for (int k0 = 0; k0 < 3; ++k0) use[k0] = MODE == 0 ? (p + 2 - k0 < n0) : (p - k0 >= 0);
#pragma unroll
    for (int j = 0; j < QL + 2; ++j) {
        if (j < ncol) {
            float2 col[3];
#pragma unroll
            for (int k1 = 0; k1 < 3; ++k1) col[k1] = *reinterpret_cast<const float2*>(pl + (k1 * W2 + j) * PEG_CS);
            if (j >= 1 && j - 1 < QL) {                         // + residual: the un-shifted input goes to output plane p
                acc[R][j - 1].x += col[1].x;
                acc[R][j - 1].y += col[1].y;
            }
#pragma unroll
            for (int k0 = 0; k0 < 3; ++k0) {
                const int slot = MODE == 0 ? (R + 2 - k0) % 3 : (R + 3 - k0) % 3;
                if (use[k0]) {
#pragma unroll
                    for (int k1 = 0; k1 < 3; ++k1)
#pragma unroll
                        for (int k2 = 0; k2 < 3; ++k2)
                            if (j - k2 >= 0 && j - k2 < QL) acc[slot][j - k2] = ffma2(wk[(k0 * 3 + k1) * 3 + k2], col[k1], acc[slot][j - k2]);
                }
            }
        }
    }
}

template <int MODE, int QL>
__global__ void __launch_bounds__(256, 2)
peg_scatter_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ w, const float* __restrict__ b,
                   float* __restrict__ y, __nv_bfloat16* __restrict__ y_bf16, int B, int n0, int n1, int n2, int dim) {
    extern __shared__ __align__(128) float psm[];
    __shared__ __align__(8) uint64_t full_bar[PEGS_RING];
    __shared__ __align__(8) float2 swt[27][16];                  // weights of the slab, [tap][channel pair]
    const int W2 = n2 + 2;
    const int plane_floats = (PEG_T1 + 2) * W2 * PEG_CS;
    const uint32_t plane_bytes = (uint32_t)plane_floats * 4u;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles1 = (n1 + PEG_T1 - 1) / PEG_T1;
    const int bb = blockIdx.y / tiles1;
    const int r_lo = (blockIdx.y % tiles1) * PEG_T1;
    const int c0 = blockIdx.x * PEG_CS;
    const int cp = lane & 15;
    const int lrow = 2 * (warp & 1) + (lane >> 4);               // row of the 4-row tile
    const int a1 = r_lo + lrow;
    const int qlen = (n2 + 3) / 4;                               // <= QL
    const int p_lo = (warp >> 1) * qlen, p_hi = min(n2, p_lo + qlen);
    const bool work = a1 < n1 && p_lo < p_hi;
    const int ncol = p_hi - p_lo + 2;
    constexpr bool REV = MODE == 1;
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < PEGS_RING; ++i) mbar_init(&full_bar[i], 1);
        mbar_fence_init();
    }
    for (int i = tid; i < 27 * 16; i += 256) {
        const int t = i / 16, q = i % 16;
        const int tt = REV ? 26 - t : t;
        swt[t][q] = make_float2(__ldg(w + (long long)(c0 + 2 * q) * 27 + tt), __ldg(w + (long long)(c0 + 2 * q + 1) * 27 + tt));
    }
    float2 bias = make_float2(0.f, 0.f);
    if (MODE == 0 && b) bias = make_float2(__ldg(b + c0 + 2 * cp), __ldg(b + c0 + 2 * cp + 1));
    __syncthreads();
    float2 wk[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) wk[t] = swt[t][cp];
    auto issue = [&](int n) {                                     // plane n -> slot n % PEGS_RING (thread 0 only)
        const int sl = n % PEGS_RING;
        mbar_expect_tx(&full_bar[sl], plane_bytes);
        tma_load_5d(psm + sl * plane_floats, &tmap, &full_bar[sl], c0, -1, r_lo - 1, n, bb);
    };
    if (tid == 0)
        for (int n = 0; n < PEGS_RING - 1 && n < n0; ++n) issue(n);
    float2 acc[3][QL];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int q = 0; q < QL; ++q) acc[s][q] = bias;
    auto store = [&](int o, float2 (&a)[QL]) {                    // finished output plane o; the slot restarts from the bias
        const long long obase = ((((long long)bb * n0 + o) * n1 + a1) * n2 + p_lo) * dim + c0 + 2 * cp;
#pragma unroll
        for (int q = 0; q < QL; ++q) {
            if (p_lo + q < p_hi) {
                *reinterpret_cast<float2*>(y + obase + (long long)q * dim) = a[q];
                if (y_bf16) *reinterpret_cast<uint32_t*>(y_bf16 + obase + (long long)q * dim) = pack_bf16x2(a[q].x, a[q].y);
            }
            a[q] = bias;
        }
    };
    auto step = [&](int p, auto rtag) {
        constexpr int R = decltype(rtag)::value;
        mbar_wait(&full_bar[p % PEGS_RING], (p / PEGS_RING) & 1);
        __syncthreads();                                          // every warp finished plane p-1: its slot can be refilled
        if (tid == 0 && p + PEGS_RING - 1 < n0) issue(p + PEGS_RING - 1);
        if (!work) return;
        const float* pl = psm + (p % PEGS_RING) * plane_floats + (lrow * W2 + p_lo) * PEG_CS + 2 * cp;
        peg_scatter_plane<MODE, QL, R>(pl, W2, ncol, wk, acc, p, n0);
        if (MODE == 0) store(p, acc[R]);
        else if (p >= 2) store(p - 2, acc[(R + 1) % 3]);
    };
    for (int p = 0; p < n0; p += 3) {
        step(p, std::integral_constant<int, 0>{});
        if (p + 1 < n0) step(p + 1, std::integral_constant<int, 1>{});
        if (p + 2 < n0) step(p + 2, std::integral_constant<int, 2>{});
    }
    if (MODE == 1 && work) {                                      // outputs n0-2 and n0-1 received their last plane
        const int r_last = (n0 - 1) % 3;
        auto flush = [&](int o) {
            if (o < 0) return;
            if (o % 3 == 0) store(o, acc[0]);
            else if (o % 3 == 1) store(o, acc[1]);
            else store(o, acc[2]);
        };
        (void)r_last;
        flush(n0 - 2);
        flush(n0 - 1);
    }
}

// ---------------------------------------------------------------------------------------------
// PEG weight / bias gradient, plane-scatter form on channel pairs: dw[k0][k1][k2] += dy[o] x[o-2+k0]
// with x plane p staged in shared memory (read once per thread) and the three dy rows it meets
// (output planes p, p+1, p+2: rows of 6 positions, 8-byte coalesced global loads) held in
// registers and rotated at compile time.  27 float2 accumulators per thread, reduced over the
// CTA's warps through shared memory, one atomic per (channel, tap) per CTA.
// ---------------------------------------------------------------------------------------------
template <int QL, int R>
__device__ __forceinline__ void peg_wgrad_plane(const float* pl, int W2, int ncol, float2 (&acc)[27], const float2 (&g)[3][QL],
                                                int p, int n0) {
    // tap k0 pairs x plane p with output plane p + 2 - k0, whose gradient row sits in slot (R + 2 - k0) % 3
    bool use[3];
#pragma unroll
    for (int k0 = 0; k0 < 3; ++k0) use[k0] = p + 2 - k0 < n0;
#pragma unroll
    for (int j = 0; j < QL + 2; ++j) {
        if (j < ncol) {
            float2 col[3];
#pragma unroll
            for (int k1 = 0; k1 < 3; ++k1) col[k1] = *reinterpret_cast<const float2*>(pl + (k1 * W2 + j) * PEG_CS);
#pragma unroll
            for (int k0 = 0; k0 < 3; ++k0) {
                const int slot = (R + 2 - k0) % 3;
                if (use[k0]) {
#pragma unroll
                    for (int k1 = 0; k1 < 3; ++k1)
#pragma unroll
                        for (int k2 = 0; k2 < 3; ++k2)
                            if (j - k2 >= 0 && j - k2 < QL)
                                acc[(k0 * 3 + k1) * 3 + k2] = ffma2(g[slot][j - k2], col[k1], acc[(k0 * 3 + k1) * 3 + k2]);
                }
            }
        }
    }
}

template <int QL>
__global__ void __launch_bounds__(256, 2)
peg_wgrad_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ dy, float* __restrict__ dw,
                 float* __restrict__ db, int B, int n0, int n1, int n2, int dim) {
    extern __shared__ __align__(128) float psm[];
    __shared__ __align__(8) uint64_t full_bar[PEGS_RING];
    const int W2 = n2 + 2;
    const int plane_floats = (PEG_T1 + 2) * W2 * PEG_CS;
    const uint32_t plane_bytes = (uint32_t)plane_floats * 4u;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles1 = (n1 + PEG_T1 - 1) / PEG_T1;
    const int bb = blockIdx.y / tiles1;
    const int r_lo = (blockIdx.y % tiles1) * PEG_T1;
    const int c0 = blockIdx.x * PEG_CS;
    const int cp = lane & 15;
    const int lrow = 2 * (warp & 1) + (lane >> 4);
    const int a1 = r_lo + lrow;
    const int qlen = (n2 + 3) / 4;
    const int p_lo = (warp >> 1) * qlen, p_hi = min(n2, p_lo + qlen);
    const bool work = a1 < n1 && p_lo < p_hi;
    const int ncol = p_hi - p_lo + 2;
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < PEGS_RING; ++i) mbar_init(&full_bar[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int n) {
        const int sl = n % PEGS_RING;
        mbar_expect_tx(&full_bar[sl], plane_bytes);
        tma_load_5d(psm + sl * plane_floats, &tmap, &full_bar[sl], c0, -1, r_lo - 1, n, bb);
    };
    if (tid == 0)
        for (int n = 0; n < PEGS_RING - 1 && n < n0; ++n) issue(n);
    float2 acc[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) acc[t] = make_float2(0.f, 0.f);
    float2 accb = make_float2(0.f, 0.f);
    float2 g[3][QL];
    auto load_g = [&](int o, float2 (&row)[QL]) {                 // gradient row of output plane o (zeros past the end)
        const long long gbase = ((((long long)bb * n0 + o) * n1 + a1) * n2 + p_lo) * dim + c0 + 2 * cp;
#pragma unroll
        for (int q = 0; q < QL; ++q) {
            row[q] = make_float2(0.f, 0.f);
            if (work && o < n0 && p_lo + q < p_hi) row[q] = *reinterpret_cast<const float2*>(dy + gbase + (long long)q * dim);
            accb.x += row[q].x;
            accb.y += row[q].y;
        }
    };
    // before plane p: slots (p % 3, (p+1) % 3) hold output planes p, p+1; plane p+2 is fetched into the third slot
    load_g(0, g[0]);
    load_g(1, g[1]);
    auto step = [&](int p, auto rtag) {
        constexpr int R = decltype(rtag)::value;
        load_g(p + 2, g[(R + 2) % 3]);                            // issued before the wait: in flight while the plane lands
        mbar_wait(&full_bar[p % PEGS_RING], (p / PEGS_RING) & 1);
        __syncthreads();
        if (tid == 0 && p + PEGS_RING - 1 < n0) issue(p + PEGS_RING - 1);
        if (!work) return;
        const float* pl = psm + (p % PEGS_RING) * plane_floats + (lrow * W2 + p_lo) * PEG_CS + 2 * cp;
        peg_wgrad_plane<QL, R>(pl, W2, ncol, acc, g, p, n0);
    };
    for (int p = 0; p < n0; p += 3) {
        step(p, std::integral_constant<int, 0>{});
        if (p + 1 < n0) step(p + 1, std::integral_constant<int, 1>{});
        if (p + 2 < n0) step(p + 2, std::integral_constant<int, 2>{});
    }
    // reduce over the CTA (16 threads share a channel pair): [16 half-warps][28][32 channels] in the ring's memory
    __syncthreads();
    float* red = psm;
    const int hw = warp * 2 + (lane >> 4);
#pragma unroll
    for (int t = 0; t < 27; ++t) {
        red[(hw * 28 + t) * 32 + 2 * cp] = work ? acc[t].x : 0.f;
        red[(hw * 28 + t) * 32 + 2 * cp + 1] = work ? acc[t].y : 0.f;
    }
    red[(hw * 28 + 27) * 32 + 2 * cp] = work ? accb.x : 0.f;
    red[(hw * 28 + 27) * 32 + 2 * cp + 1] = work ? accb.y : 0.f;
    __syncthreads();
    for (int i = tid; i < 28 * 32; i += 256) {
        const int t = i / 32, l = i % 32;
        float sacc = 0.f;
#pragma unroll
        for (int h = 0; h < 16; ++h) sacc += red[(h * 28 + t) * 32 + l];
        if (t < 27) atomicAdd(dw + (long long)(c0 + l) * 27 + t, sacc);
        else atomicAdd(db + c0 + l, sacc);
    }
}

// =============================================================================================
// VQ helpers
// =============================================================================================
__global__ void l2norm_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xn_bf16,
                                   float* __restrict__ xn_f32, long long rows, int dim) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * dim;
    float ss = 0.f;
    for (int c = lane; c < dim; c += 32) { const float v = xr[c]; ss = fmaf(v, v, ss); }
    const float rn = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
    for (int c = lane; c < dim; c += 32) {
        const float v = xr[c] * rn;
        if (xn_bf16) xn_bf16[row * dim + c] = __float2bfloat16(v);
        if (xn_f32) xn_f32[row * dim + c] = v;
    }
}

__global__ void vq_gather_kernel(const unsigned long long* __restrict__ best,
                                 const float* __restrict__ embed, long long* __restrict__ ind,
                                 float* __restrict__ quant, long long rows, int dim, int C) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    long long idx = (long long)(0xffffffffu - (uint32_t)(best[row] & 0xffffffffull));
    if (idx < 0 || idx >= C) idx = 0;
    if (lane == 0) ind[row] = idx;
    const float* e = embed + idx * dim;
    for (int c = lane; c < dim; c += 32) quant[row * dim + c] = __ldg(e + c);
}

// Exact VQ code selection (pass 2; pass 1 = the CTK_EPI_ARGMAX_PART GEMM epilogue on bf16-rounded unit vectors).
// CTA = 32 rows: the per-block (best key, second value) pairs are staged in shared memory with coalesced loads,
// then each warp resolves 4 rows.  A row whose bf16 maximum is unique within `margin` keeps that code; otherwise
// every code that may be the fp32 maximum - the best of each block within the margin, or all 128 codes of a block
// whose second best is within it too - is re-scored as an fp32 dot product and the largest (lowest index on ties) wins.
__device__ __forceinline__ float vq_key_value(unsigned long long k) {
    uint32_t u = (uint32_t)(k >> 32);
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}

__global__ void __launch_bounds__(256)
vq_select_kernel(const unsigned long long* __restrict__ pkey, const float* __restrict__ psec, int nblk,
                 const float* __restrict__ x, const float* __restrict__ en, const float* __restrict__ embed,
                 long long* __restrict__ ind, float* __restrict__ quant, long long rows, int dim, int C, float margin) {
    extern __shared__ __align__(16) uint8_t vq_sm[];
    unsigned long long* sk = reinterpret_cast<unsigned long long*>(vq_sm);
    float* ss = reinterpret_cast<float*>(sk + (size_t)nblk * 32);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long row0 = (long long)blockIdx.x * 32;
    for (int i = tid; i < nblk * 32; i += 256) {
        const int b = i >> 5, r = i & 31;
        const bool ok = row0 + r < rows;
        sk[i] = ok ? pkey[(long long)b * rows + row0 + r] : 0ull;
        ss[i] = ok ? psec[(long long)b * rows + row0 + r] : -INFINITY;
    }
    __syncthreads();
    const int nv = dim >> 5;
    for (int rr = 0; rr < 4; ++rr) {
        const int r = warp * 4 + rr;
        const long long row = row0 + r;
        if (row >= rows) break;                                   // warp-uniform
        unsigned long long kb = 0ull;
        for (int b = lane; b < nblk; b += 32) { const unsigned long long k = sk[b * 32 + r]; kb = k > kb ? k : kb; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, kb, o);
            kb = t > kb ? t : kb;
        }
        const float thr = vq_key_value(kb) - margin;
        // candidate blocks as ballots over groups of 32 blocks: best within the margin / second best within the margin
        auto ballots = [&](int b0, unsigned& cand, unsigned& full) {
            const int b = b0 + lane;
            const bool in = b < nblk;
            cand = __ballot_sync(0xffffffffu, in && vq_key_value(sk[(in ? b : 0) * 32 + r]) >= thr);
            full = __ballot_sync(0xffffffffu, in && ss[(in ? b : 0) * 32 + r] >= thr);
        };
        int ncand = 0;
        unsigned anyfull = 0u;
        for (int b0 = 0; b0 < nblk; b0 += 32) {
            unsigned c, f;
            ballots(b0, c, f);
            ncand += __popc(c);
            anyfull |= f;
        }
        long long best_i = (long long)(0xffffffffu - (uint32_t)(kb & 0xffffffffull));
        if (ncand > 1 || anyfull != 0u) {
            // fp32 re-score: xn = x * (1 / max(|x|, 1e-12)) exactly as ctk_l2norm_rows forms it
            float xv[32];
            const float* xr = x + row * dim;
            float sq = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                xv[j] = j < nv ? xr[lane + 32 * j] : 0.f;
                sq = fmaf(xv[j], xv[j], sq);
            }
            const float rn = 1.0f / fmaxf(sqrtf(warp_sum(sq)), 1e-12f);
#pragma unroll
            for (int j = 0; j < 32; ++j) xv[j] *= rn;
            float best_v = -INFINITY;
            best_i = 0x7fffffffffffffffLL;
            auto score = [&](long long c) {
                const float* e = en + c * dim;
                float d = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < nv) d = fmaf(xv[j], __ldg(e + lane + 32 * j), d);
                d = warp_sum(d);
                if (d > best_v || (d == best_v && c < best_i)) { best_v = d; best_i = c; }
            };
            for (int b0 = 0; b0 < nblk; b0 += 32) {
                unsigned cand, full;
                ballots(b0, cand, full);
                for (unsigned m = cand; m; m &= m - 1) {               // warp-uniform walk over the set bits
                    const int bit = __ffs(m) - 1, b = b0 + bit;
                    if ((full >> bit) & 1u) {
                        const int c1 = min(C, (b + 1) * 128);
                        for (int c = b * 128; c < c1; ++c) score(c);
                    } else {
                        score((long long)(0xffffffffu - (uint32_t)(sk[b * 32 + r] & 0xffffffffull)));
                    }
                }
            }
        }
        if (best_i < 0 || best_i >= C) best_i = 0;
        if (lane == 0) ind[row] = best_i;
        const float4* e4 = reinterpret_cast<const float4*>(embed + best_i * dim);
        float4* q4 = reinterpret_cast<float4*>(quant + row * dim);
        for (int c = lane; c < dim / 4; c += 32) q4[c] = __ldg(e4 + c);
    }
}

// embed_sum[code] += xn[row]; bins[code] += 1
__global__ void vq_scatter_kernel(const float* __restrict__ xn, const long long* __restrict__ ind,
                                  float* __restrict__ esum, float* __restrict__ bins, long long rows,
                                  int dim) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const long long code = ind[row];
    if (lane == 0) atomicAdd(bins + code, 1.0f);
    for (int c = lane; c < dim; c += 32) atomicAdd(esum + code * dim + c, xn[row * dim + c]);
}

// one warp per code: cluster_size EMA; embed <- decay*embed + (1-decay)*l2norm(esum/bins) (codes
// without hits are pulled towards their own l2-normalised value, as in the cosine-sim codebook)
__global__ void vq_ema_kernel(float* __restrict__ cluster_size, float* __restrict__ embed,
                              const float* __restrict__ esum, const float* __restrict__ bins, int C,
                              int dim, float decay) {
    const int lane = threadIdx.x & 31;
    const int code = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (code >= C) return;
    const float n = bins[code];
    if (lane == 0) cluster_size[code] = cluster_size[code] * decay + n * (1.0f - decay);
    const float* src = n > 0.f ? esum + (long long)code * dim : embed + (long long)code * dim;
    const float inv = n > 0.f ? 1.0f / n : 1.0f;
    float ss = 0.f;
    for (int c = lane; c < dim; c += 32) { const float v = src[c] * inv; ss = fmaf(v, v, ss); }
    const float rn = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
    for (int c = lane; c < dim; c += 32) {
        const float v = src[c] * inv * rn;
        float* e = embed + (long long)code * dim + c;
        *e = *e * decay + v * (1.0f - decay);
    }
}

template <int NV>
int launch_ln_fwd(const float* x, const float* gamma, const float* beta, void* out_bf16, float* out_f32,
                  void* xraw, float* mean, float* rstd, long long rows, float eps, int po, int pi,
                  cudaStream_t s) {
    long long blocks = (rows + 7) / 8;
    const long long cap = (long long)ctk_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    layernorm_fwd_kernel<NV><<<(int)blocks, 256, 0, s>>>(
        x, gamma, beta, reinterpret_cast<__nv_bfloat16*>(out_bf16), out_f32,
        reinterpret_cast<__nv_bfloat16*>(xraw), mean, rstd, rows, eps, po, pi);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

template <int NV>
int launch_ln_bwd(const void* dyb, const float* dyf, const float* x, const float* gamma, const float* mean,
                  const float* rstd, float* dx, void* dxb, int accum, float* dgamma, float* dbeta, long long rows,
                  int po, int pi, long long bc, float sc, cudaStream_t s) {
    long long blocks = (rows + 7) / 8;
    const long long cap = (long long)ctk_num_sms() * 4;
    if (blocks > cap) blocks = cap;
    layernorm_bwd_kernel<NV><<<(int)blocks, 256, 0, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(dyb), dyf, x, gamma, mean, rstd, dx,
        reinterpret_cast<__nv_bfloat16*>(dxb), accum, dgamma, dbeta, rows, po, pi, bc, sc);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

}  // namespace

extern "C" int ctk_patch_norm_fwd(const float* video, void* xhat, long long ld, float* mean,
                                  float* rstd, int B, int D, int H, int W, int pt, int p1, int p2,
                                  float eps, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(video && xhat && mean && rstd, CTK_ERR_SHAPE, "patch_norm: null pointer");
    CTK_REQUIRE(B > 0 && pt > 0 && p1 > 0 && p2 > 0 && D % pt == 0 && H % p1 == 0 && W % p2 == 0,
                CTK_ERR_SHAPE, "patch_norm: volume %dx%dx%d not divisible by patch %dx%dx%d", D, H, W, pt, p1, p2);
    const int K = pt * p1 * p2;
    CTK_REQUIRE(ld >= K && ld % 8 == 0 && CTK_ALIGNED(xhat, 16) && CTK_ALIGNED(video, 16), CTK_ERR_ALIGN,
                "patch_norm: xhat pitch must be a multiple of 8 and >= %d", K);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    const int Wp = W / p2;
    if (pt == 10 && p1 == 20 && p2 == 20 && Wp % 4 == 0 && W % 4 == 0) {
        // production geometry (run_train.py:56-66): TMA-staged boxes
        CUtensorMap tm;
        const unsigned long long dims[3] = {(unsigned long long)W, (unsigned long long)H, (unsigned long long)B * D};
        const unsigned long long st[2] = {(unsigned long long)W * 4, (unsigned long long)W * H * 4};
        const unsigned int box[3] = {80, 20, 10};
        rc = ctk_make_tmap(&tm, video, true, 3, dims, st, box, 0);
        if (rc) return rc;
        const size_t smem = (size_t)4 * K * sizeof(float);
        auto kern = patch_norm_tma_kernel<10, 20, 20, 4>;
        CTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const long long ctas = (long long)B * (D / pt) * (H / p1) * (Wp / 4);
        kern<<<(unsigned)ctas, 256, smem, s>>>(tm, reinterpret_cast<__nv_bfloat16*>(xhat), ld, mean, rstd, D, H, W, eps);
        CTK_LAUNCH_CHECK();
        return CTK_OK;
    }
    const int G = (Wp % 4 == 0 && (size_t)4 * K * 4 <= 96 * 1024) ? 4 : (Wp % 2 == 0 && (size_t)2 * K * 4 <= 96 * 1024) ? 2 : 1;
    const size_t smem = (size_t)G * K * sizeof(float);
    CTK_REQUIRE(smem <= 200 * 1024, CTK_ERR_SHAPE, "patch_norm: patch of %d voxels too large", K);
    const long long ctas = (long long)B * (D / pt) * (H / p1) * (Wp / G);
#define CTK_PN(GV)                                                                                 \
    {                                                                                              \
        CTK_CUDA(cudaFuncSetAttribute(patch_norm_kernel<GV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        patch_norm_kernel<GV><<<(unsigned)ctas, 256, smem, s>>>(                                   \
            video, reinterpret_cast<__nv_bfloat16*>(xhat), ld, mean, rstd, D, H, W, pt, p1, p2, eps); \
    }
    if (G == 4) CTK_PN(4) else if (G == 2) CTK_PN(2) else CTK_PN(1)
#undef CTK_PN
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

extern "C" int ctk_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* out_bf16,
                                 float* out_f32, void* xraw_bf16, float* mean, float* rstd,
                                 long long rows, int dim, float eps, int perm_outer, int perm_inner,
                                 void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(x && gamma && rows > 0, CTK_ERR_SHAPE, "layernorm_fwd: bad args");
    CTK_REQUIRE(dim % 64 == 0 && dim >= 64 && dim <= 1024, CTK_ERR_SHAPE,
                "layernorm: dim %d must be a multiple of 64 in [64, 1024]", dim);
    CTK_REQUIRE(perm_inner <= 0 || (perm_outer > 0 && rows % ((long long)perm_outer * perm_inner) == 0),
                CTK_ERR_SHAPE, "layernorm: rows not divisible by the permutation block");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    switch (dim / 64) {
#define CTK_LN(NV) case NV: return launch_ln_fwd<NV>(x, gamma, beta, out_bf16, out_f32, xraw_bf16, mean, rstd, rows, eps, perm_outer, perm_inner, s);
        CTK_LN(1) CTK_LN(2) CTK_LN(4) CTK_LN(8) CTK_LN(12) CTK_LN(16)
#undef CTK_LN
        default: break;
    }
    ctk_set_error("layernorm: dim %d not instantiated (64,128,256,512,768,1024)", dim);
    return CTK_ERR_SHAPE;
}

extern "C" int ctk_layernorm_bwd(const void* dy_bf16, const float* dy_f32, const float* x,
                                 const float* gamma, const float* mean, const float* rstd, float* dx,
                                 void* dx_bf16, int dx_accum, float* dgamma, float* dbeta, long long rows, int dim,
                                 int perm_outer, int perm_inner, long long dy_bcast_rows, float dy_scale,
                                 void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE((dy_bf16 || dy_f32) && x && gamma && mean && rstd && dx && dgamma && rows > 0,
                CTK_ERR_SHAPE, "layernorm_bwd: bad args");
    CTK_REQUIRE(dim % 64 == 0 && dim >= 64 && dim <= 1024, CTK_ERR_SHAPE, "layernorm_bwd: dim %d", dim);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    switch (dim / 64) {
#define CTK_LN(NV) case NV: return launch_ln_bwd<NV>(dy_bf16, dy_f32, x, gamma, mean, rstd, dx, dx_bf16, dx_accum, dgamma, dbeta, rows, perm_outer, perm_inner, dy_bcast_rows, dy_scale, s);
        CTK_LN(1) CTK_LN(2) CTK_LN(4) CTK_LN(8) CTK_LN(12) CTK_LN(16)
#undef CTK_LN
        default: break;
    }
    ctk_set_error("layernorm_bwd: dim %d not instantiated", dim);
    return CTK_ERR_SHAPE;
}

static size_t peg_smem(int n2) {
    const size_t ring = (size_t)PEG_RING * (PEG_T1 + 2) * (n2 + 2) * PEG_CS * sizeof(float);
    const size_t red = (size_t)8 * 28 * 32 * sizeof(float);          // MODE 2 cross-warp reduction
    return ring > red ? ring : red;
}

template <int MODE>
static int peg_launch(const float* staged, const float* w, const float* b, const float* dy, float* y, void* y_bf16,
                      float* dw, float* db, int B, int n0, int n1, int n2, int dim, cudaStream_t s) {
    const size_t sm = peg_smem(n2);
    CTK_REQUIRE(sm <= 200 * 1024 && n2 + 2 <= 256, CTK_ERR_SHAPE, "peg: axis-2 extent %d does not fit in shared memory", n2);
    // 5-D view {dim, n2, n1, n0, B} of the tensor that is staged through shared memory
    CUtensorMap tm;
    const unsigned long long dims[5] = {(unsigned long long)dim, (unsigned long long)n2, (unsigned long long)n1,
                                        (unsigned long long)n0, (unsigned long long)B};
    const unsigned long long st[4] = {(unsigned long long)dim * 4, (unsigned long long)dim * n2 * 4,
                                      (unsigned long long)dim * n2 * n1 * 4, (unsigned long long)dim * n2 * n1 * n0 * 4};
    const unsigned int box[5] = {PEG_CS, (unsigned)(n2 + 2), PEG_T1 + 2, 1, 1};
    int rc = ctk_make_tmap(&tm, staged, true, 5, dims, st, box, 0);
    if (rc) return rc;
    const int tiles1 = (n1 + PEG_T1 - 1) / PEG_T1;
    if constexpr (MODE != 2) {
        static const bool gather = [] { const char* e = getenv("CTK_PEG_GATHER"); return e && e[0] == '1'; }();
        if (n2 <= 24 && !gather) {                                   // plane-scatter form, 6 positions per thread
            const size_t sm_s = (size_t)PEGS_RING * (PEG_T1 + 2) * (n2 + 2) * PEG_CS * 4;
            CTK_CUDA(cudaFuncSetAttribute(peg_scatter_kernel<MODE, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_s));
            peg_scatter_kernel<MODE, 6><<<dim3(dim / PEG_CS, B * tiles1), 256, sm_s, s>>>(
                tm, w, b, y, reinterpret_cast<__nv_bfloat16*>(y_bf16), B, n0, n1, n2, dim);
            CTK_LAUNCH_CHECK();
            return CTK_OK;
        }
        if (n2 <= 4 * PEG_QMAX) {
            CTK_CUDA(cudaFuncSetAttribute(peg_pair_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            peg_pair_kernel<MODE><<<dim3(dim / PEG_CS, B * tiles1), 256, sm, s>>>(
                tm, w, b, y, reinterpret_cast<__nv_bfloat16*>(y_bf16), B, n0, n1, n2, dim);
            CTK_LAUNCH_CHECK();
            return CTK_OK;
        }
    }
    if constexpr (MODE == 2) {
        static const bool gather = [] { const char* e = getenv("CTK_PEG_GATHER"); return e && e[0] == '1'; }();
        const size_t sm_s = (size_t)PEGS_RING * (PEG_T1 + 2) * (n2 + 2) * PEG_CS * 4;
        if (n2 <= 24 && !gather && sm_s >= (size_t)16 * 28 * 32 * 4) {
            CTK_CUDA(cudaFuncSetAttribute(peg_wgrad_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_s));
            peg_wgrad_kernel<6><<<dim3(dim / PEG_CS, B * tiles1), 256, sm_s, s>>>(tm, dy, dw, db, B, n0, n1, n2, dim);
            CTK_LAUNCH_CHECK();
            return CTK_OK;
        }
    }
    CTK_CUDA(cudaFuncSetAttribute(peg_tile_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    peg_tile_kernel<MODE><<<dim3(dim / PEG_CS, B * tiles1), 256, sm, s>>>(
        tm, w, b, dy, y, reinterpret_cast<__nv_bfloat16*>(y_bf16), dw, db, B, n0, n1, n2, dim);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

extern "C" int ctk_peg_fwd(const float* x, const float* w, const float* b, float* y, int B, int n0,
                           int n1, int n2, int dim, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(x && w && b && y && B > 0 && n0 > 0 && n1 > 0 && n2 > 0 && dim > 0, CTK_ERR_SHAPE, "peg_fwd: bad args");
    CTK_REQUIRE(dim % PEG_CS == 0 && CTK_ALIGNED(x, 16), CTK_ERR_ALIGN, "peg: dim must be a multiple of 32, x 16-byte aligned");
    CTK_REQUIRE(x != y, CTK_ERR_SHAPE, "peg_fwd: in-place not supported");
    return peg_launch<0>(x, w, b, nullptr, y, nullptr, nullptr, nullptr, B, n0, n1, n2, dim,
                         reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int ctk_peg_bwd(const float* dy, const float* x, const float* w, float* dx, void* dx_bf16,
                           float* dw, float* db, int B, int n0, int n1, int n2, int dim, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(dy && x && w && dx && dw && db && B > 0 && dim > 0, CTK_ERR_SHAPE, "peg_bwd: bad args");
    CTK_REQUIRE(dim % PEG_CS == 0 && CTK_ALIGNED(x, 16) && CTK_ALIGNED(dy, 16), CTK_ERR_ALIGN,
                "peg: dim must be a multiple of 32, tensors 16-byte aligned");
    CTK_REQUIRE(dy != dx, CTK_ERR_SHAPE, "peg_bwd: in-place not supported");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    rc = peg_launch<1>(dy, w, nullptr, dy, dx, dx_bf16, nullptr, nullptr, B, n0, n1, n2, dim, s);
    if (rc) return rc;
    return peg_launch<2>(x, nullptr, nullptr, dy, nullptr, nullptr, dw, db, B, n0, n1, n2, dim, s);
}

extern "C" int ctk_l2norm_rows(const float* x, void* xn_bf16, float* xn_f32, long long rows, int dim,
                               void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(x && (xn_bf16 || xn_f32) && rows > 0 && dim > 0, CTK_ERR_SHAPE, "l2norm_rows: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    l2norm_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(x, reinterpret_cast<__nv_bfloat16*>(xn_bf16), xn_f32, rows, dim);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

extern "C" int ctk_vq_select(const void* part_key, const float* part_second, int nblk, const float* x, const float* en,
                             const float* embed, long long* ind, float* quant, long long rows, int dim,
                             int codebook_size, float margin, void* stream) {
    int rc = ctk_check_device();
    if (rc != CTK_OK) return rc;
    CTK_REQUIRE(part_key && part_second && x && en && embed && ind && quant && rows > 0, CTK_ERR_SHAPE, "vq_select: bad args");
    CTK_REQUIRE(dim > 0 && dim % 32 == 0 && dim <= 1024 && codebook_size > 0 && nblk == (codebook_size + 127) / 128,
                CTK_ERR_SHAPE, "vq_select: dim %d (need %% 32 == 0, <= 1024), codebook %d, nblk %d", dim, codebook_size, nblk);
    CTK_REQUIRE(margin > 0.f, CTK_ERR_SHAPE, "vq_select: margin must be positive");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)nblk * 32 * 12;
    CTK_REQUIRE(smem <= 200 * 1024, CTK_ERR_SHAPE, "vq_select: codebook too large (%d blocks)", nblk);
    if (smem > 48 * 1024)
        CTK_CUDA(cudaFuncSetAttribute(vq_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vq_select_kernel<<<(unsigned)((rows + 31) / 32), 256, smem, s>>>(
        reinterpret_cast<const unsigned long long*>(part_key), part_second, nblk, x, en, embed, ind, quant, rows, dim,
        codebook_size, margin);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

extern "C" int ctk_vq_gather(const void* best, const float* embed, long long* ind, float* quant,
                             long long rows, int dim, int codebook_size, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(best && embed && ind && quant && rows > 0 && dim > 0 && codebook_size > 0, CTK_ERR_SHAPE, "vq_gather: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    vq_gather_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(
        reinterpret_cast<const unsigned long long*>(best), embed, ind, quant, rows, dim, codebook_size);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

extern "C" int ctk_vq_ema_update(const float* xn_f32, const long long* ind, float* cluster_size,
                                 float* embed, float* ws, long long rows, int dim, int codebook_size,
                                 float decay, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(xn_f32 && ind && cluster_size && embed && ws && rows > 0, CTK_ERR_SHAPE, "vq_ema_update: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    float* esum = ws;
    float* bins = ws + (size_t)codebook_size * dim;
    CTK_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * ((size_t)codebook_size * dim + codebook_size), s));
    vq_scatter_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(xn_f32, ind, esum, bins, rows, dim);
    CTK_LAUNCH_CHECK();
    vq_ema_kernel<<<(codebook_size + 7) / 8, 256, 0, s>>>(cluster_size, embed, esum, bins, codebook_size, dim, decay);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

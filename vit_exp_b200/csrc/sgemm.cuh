// fp32 SIMT 64x64 tile product used where the reference's arithmetic is fp32 by contract and the
// problem is small: the N x N contrastive logits (ct_clip.py:1347), their gradients, and the
// ContinuousPositionBias MLP (attention.py:377-380 forces .float()).
#pragma once
#include "common.cuh"

// 256 threads; thread (ty, tx) = (tid / 16, tid % 16) owns rows m0+ty*4..+3, cols n0+tx*4..+3.
// A row-major [M, K]; B row-major [N, K] when BT (C = A * B^T) else row-major [K, N] (C = A * B).
// smem: As, Bs = float[16][68] each.
// AT: A is stored transposed, row-major [K, M] (C = A^T * B).
template <bool BT, bool AT = false>
__device__ __forceinline__ void sgemm_tile_64x64(const float* __restrict__ A, long long lda,
                                                 const float* __restrict__ B, long long ldb,
                                                 int m0, int n0, int M, int N, int K,
                                                 float (&acc)[4][4], float (*As)[68],
                                                 float (*Bs)[68]) {
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += 16) {
        // A tile: 64 rows x 16 k -> As[k][row]
        if (AT) {
            const int kk = tid >> 4, c = (tid & 15) * 4;
            const int gk = k0 + kk;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gm = m0 + c + q;
                As[kk][c + q] = (gk < K && gm < M) ? A[(long long)gk * lda + gm] : 0.f;
            }
        } else {
            const int r = tid >> 2, kk = (tid & 3) * 4;
            const int gr = m0 + r;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gk = k0 + kk + q;
                As[kk + q][r] = (gr < M && gk < K) ? A[(long long)gr * lda + gk] : 0.f;
            }
        }
        if (BT) {
            const int r = tid >> 2, kk = (tid & 3) * 4;
            const int gr = n0 + r;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gk = k0 + kk + q;
                Bs[kk + q][r] = (gr < N && gk < K) ? B[(long long)gr * ldb + gk] : 0.f;
            }
        } else {
            const int kk = tid >> 4, c = (tid & 15) * 4;
            const int gk = k0 + kk;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gc = n0 + c + q;
                Bs[kk][c + q] = (gk < K && gc < N) ? B[(long long)gk * ldb + gc] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
}

// Warp-level tensor-core helpers (mma.sync m16n8k16 / m16n8k8, ldmatrix) on [rows][32 bf16] tiles
// with 64-byte rows whose 16-byte chunks are XOR-swizzled by (row >> 1) & 3 - the layout a TMA box
// {32 elements, rows} with CU_TENSOR_MAP_SWIZZLE_64B produces.  Shared by attention.cu (legacy
// long-sequence kernels) and attention_seq24.cu (temporal stack).
#pragma once
#include "common.cuh"

namespace attn_mma {

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// [rows][32 bf16] tile, 64-byte rows, 16-byte chunks XOR-swizzled by (row>>1)&3: conflict-free ldmatrix
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
    return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}

// copy `nrows` rows (32 bf16 each, global pitch ld elements) into a swizzled tile; rows >= valid are zero
__device__ __forceinline__ void load_tile(uint8_t* dst, const __nv_bfloat16* src, long long ld, int nrows,
                                          int valid, int tid, int nthreads) {
    for (int i = tid; i < nrows * 4; i += nthreads) {
        const int row = i >> 2, c = i & 3;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (row < valid) v = __ldg(reinterpret_cast<const uint4*>(src + (long long)row * ld) + c);
        *reinterpret_cast<uint4*>(dst + tile_off(row, c)) = v;
    }
}

// A fragments (16 rows x 32 k) of a tile for the warp's rows r0..r0+15 : a[kstep][4]
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[2][4], uint32_t tile, int r0, int lane) {
    const int m = lane >> 3;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
        ldsm4(a[ks], tile + tile_off(r0 + (m & 1) * 8 + (lane & 7), 2 * ks + (m >> 1)));
}
// B fragments for an operand stored [n][k] (k = 32 channels): n-tile n0..n0+7 -> {b0,b1 (k 0-15), b0,b1 (k 16-31)}
__device__ __forceinline__ void load_b_nk(uint32_t (&b)[4], uint32_t tile, int n0, int lane) {
    ldsm4(b, tile + tile_off(n0 + (lane & 7), lane >> 3));
}
// B fragments for an operand stored [k][n] (n = 32 channels): k-block k0..k0+15, channel n-tiles 2cp, 2cp+1
__device__ __forceinline__ void load_b_kn(uint32_t (&b)[4], uint32_t tile, int k0, int cp, int lane) {
    const int m = lane >> 3;
    ldsm4t(b, tile + tile_off(k0 + (m & 1) * 8 + (lane & 7), 2 * cp + (m >> 1)));
}

// acc[8][4] (16 x 64) = A(16x32) * B^T where B is a [64][32] block of `tile` starting at row n_base
__device__ __forceinline__ void mma_16x64(float (&acc)[8][4], const uint32_t (&a)[2][4], uint32_t tile,
                                          int n_base, int lane) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        uint32_t b[4];
        load_b_nk(b, tile, n_base + nt * 8, lane);
        mma16816(acc[nt], a[0], b[0], b[1]);
        mma16816(acc[nt], a[1], b[2], b[3]);
    }
}
// out[4][4] (16 x 32) += P(16 x 64, bf16 A-fragments from an accumulator) * B where B = [64][32] rows k_base..
__device__ __forceinline__ void mma_acc_16x32(float (&out)[4][4], const float (&p)[8][4], uint32_t tile,
                                              int k_base, int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[4];
        a[0] = pack_bf16x2(p[2 * kk][0], p[2 * kk][1]);
        a[1] = pack_bf16x2(p[2 * kk][2], p[2 * kk][3]);
        a[2] = pack_bf16x2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
        a[3] = pack_bf16x2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
        for (int cp = 0; cp < 2; ++cp) {
            uint32_t b[4];
            load_b_kn(b, tile, k_base + kk * 16, cp, lane);
            mma16816(out[2 * cp], a, b[0], b[1]);
            mma16816(out[2 * cp + 1], a, b[2], b[3]);
        }
    }
}

__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}


// m16n8k8: D += A(16x8, {a0: row g, a1: row g+8; k = 2t, 2t+1}) * B(8x8, b0: k = 2t, 2t+1; n = g)
__device__ __forceinline__ void mma1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(b0));
}

}  // namespace attn_mma

// Cosine attention core (attention.py:162-184) on the packed qkv buffer produced by the
// CTK_EPI_QKV GEMM epilogue.  q is already l2-normalised, scaled by q_scale and by the constant 8;
// k is l2-normalised and scaled by k_scale, so logits = q.k (+ relative position bias).
//
// Two regimes (hard parts 2-3):
//  * "long" sequences (spatial stack, L = 576 tokens per slice): flash-style kernels - the L x L
//    logits live only in registers, K/V of one (sequence, head) are staged once in shared memory
//    with a 16-byte XOR swizzle, the products run on the tensor cores (bf16 m16n8k16, fp32
//    accumulate), softmax is online in fp32 with exp2.  The bias is gathered from the
//    (2gh-1)x(2gw-1) table in shared memory.  Backward = dq kernel + dk/dv kernel (transposed
//    formulation, no atomics) + bias-table gradient kernel (tile fixed, loops over sequences).
//  * "short" sequences (temporal stack, L = 24): one CTA per sequence, one warp per head, one lane
//    per query/key, fp32 SIMT - a 24x24 problem cannot fill an MMA tile.
//
// NOTE: this round the long-sequence path uses the legacy mma.sync tensor-core path; moving it to
// tcgen05/TMEM is tracked in DESIGN.md (attention is ~5% of the encoder FLOPs).
#include "attention_mma.cuh"
#include <stdlib.h>

using namespace attn_mma;

namespace {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

// Relative-position bias lookup.  index(i, j) = (yi-yj)*ww + (xi-xj) + off = pos[i] + off - pos[j]
// with pos[n] = (n / gw) * ww + n % gw precomputed per CTA in shared memory (no divisions inside
// the tile loops).
struct BiasIdx {
    const float* tab;      // shared memory, [(2gh-1)*(2gw-1)] for this head (nullptr = no bias)
    const int* pos;        // shared memory, [Lp]
    int off;
    __device__ __forceinline__ int index(int i, int j) const { return pos[i] + off - pos[j]; }
};

__device__ __forceinline__ void fill_pos(int* pos, int Lp, int L, int gw, int tid, int nthreads) {
    const int ww = 2 * gw - 1;
    for (int n = tid; n < Lp; n += nthreads) {
        const int m = n < L ? n : L - 1;
        pos[n] = gw > 0 ? (m / gw) * ww + (m % gw) : 0;
    }
}
__device__ __forceinline__ BiasIdx make_bias(const float* tab_smem, const int* pos_smem, int gh, int gw) {
    BiasIdx b;
    b.tab = tab_smem; b.pos = pos_smem; b.off = (gh - 1) * (2 * gw - 1) + (gw - 1);
    return b;
}

// =============================================================================================
// forward: grid (q blocks of 16*NW, heads, nseq), 32*NW threads
// smem: K [Lp][32] | V [Lp][32] | Q [16 NW][32] | table | pos
// =============================================================================================
template <int NW>
__global__ void __launch_bounds__(32 * NW)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ table,
                __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int L, int heads, int gh, int gw) {
    constexpr int QB = 16 * NW, NT = 32 * NW;
    extern __shared__ __align__(128) uint8_t smem[];
    const int Lp = (L + 63) & ~63;
    uint8_t* sK = smem;
    uint8_t* sV = sK + Lp * 64;
    uint8_t* sQ = sV + Lp * 64;
    int* sP = reinterpret_cast<int*>(sQ + QB * 64);
    float* sT = reinterpret_cast<float*>(sP + Lp);
    const int qb = blockIdx.x, h = blockIdx.y, s = blockIdx.z;
    const int inner = heads * 32;
    const long long ld = 3LL * inner;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const __nv_bfloat16* base = qkv + (long long)s * L * ld + h * 32;
    load_tile(sK, base + inner, ld, Lp, L, tid, NT);
    load_tile(sV, base + 2 * inner, ld, Lp, L, tid, NT);
    load_tile(sQ, base + (long long)qb * QB * ld, ld, QB, L - qb * QB, tid, NT);
    const int n_off = (2 * gh - 1) * (2 * gw - 1);
    if (table)
        for (int i = tid; i < n_off; i += NT) sT[i] = table[(long long)h * n_off + i] * LOG2E;
    fill_pos(sP, Lp, L, gw, tid, NT);
    __syncthreads();
    const BiasIdx bias = make_bias(table ? sT : nullptr, sP, gh, gw);
    const uint32_t tK = smem_u32(sK), tV = smem_u32(sV), tQ = smem_u32(sQ);
    uint32_t qa[2][4];
    load_a_frags(qa, tQ, warp * 16, lane);
    const int g = lane >> 2, t = lane & 3;
    const int i0 = qb * QB + warp * 16 + g, i1 = i0 + 8;      // this thread's two query rows
    const int rb0 = sP[min(i0, Lp - 1)] + bias.off, rb1 = sP[min(i1, Lp - 1)] + bias.off;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) o[a][0] = o[a][1] = o[a][2] = o[a][3] = 0.f;
    const bool has_bias = table != nullptr;

    for (int kb = 0; kb < Lp; kb += 64) {
        float sc[8][4];
        mma_16x64(sc, qa, tK, kb, lane);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int j = kb + nt * 8 + 2 * t;
            float b00 = 0.f, b01 = 0.f, b10 = 0.f, b11 = 0.f;
            if (has_bias) {
                const int2 pj = *reinterpret_cast<const int2*>(sP + j);
                b00 = sT[rb0 - pj.x]; b01 = sT[rb0 - pj.y]; b10 = sT[rb1 - pj.x]; b11 = sT[rb1 - pj.y];
            }
            sc[nt][0] = fmaf(sc[nt][0], LOG2E, b00); sc[nt][1] = fmaf(sc[nt][1], LOG2E, b01);
            sc[nt][2] = fmaf(sc[nt][2], LOG2E, b10); sc[nt][3] = fmaf(sc[nt][3], LOG2E, b11);
            if (kb + 64 > L) {                   // only the last key block can hold padding
                if (j >= L) { sc[nt][0] = -INFINITY; sc[nt][2] = -INFINITY; }
                if (j + 1 >= L) { sc[nt][1] = -INFINITY; sc[nt][3] = -INFINITY; }
            }
            mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
        }
        mx0 = quad_max(mx0); mx1 = quad_max(mx1);
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float r0 = exp2f(m0 - mn0), r1 = exp2f(m1 - mn1);
        m0 = mn0; m1 = mn1;
        float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            sc[nt][0] = exp2f(sc[nt][0] - mn0); sc[nt][1] = exp2f(sc[nt][1] - mn0);
            sc[nt][2] = exp2f(sc[nt][2] - mn1); sc[nt][3] = exp2f(sc[nt][3] - mn1);
            ps0 += sc[nt][0] + sc[nt][1];
            ps1 += sc[nt][2] + sc[nt][3];
        }
        l0 = l0 * r0 + ps0; l1 = l1 * r1 + ps1;
#pragma unroll
        for (int a = 0; a < 4; ++a) { o[a][0] *= r0; o[a][1] *= r0; o[a][2] *= r1; o[a][3] *= r1; }
        mma_acc_16x32(o, sc, tV, kb, lane);
    }
    l0 = quad_sum(l0); l1 = quad_sum(l1);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    __nv_bfloat16* ob = out + (long long)s * L * inner + h * 32;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int d = a * 8 + 2 * t;
        if (i0 < L) *reinterpret_cast<uint32_t*>(ob + (long long)i0 * inner + d) = pack_bf16x2(o[a][0] * inv0, o[a][1] * inv0);
        if (i1 < L) *reinterpret_cast<uint32_t*>(ob + (long long)i1 * inner + d) = pack_bf16x2(o[a][2] * inv1, o[a][3] * inv1);
    }
    if (t == 0) {
        float* lp = lse + ((long long)s * heads + h) * L;
        if (i0 < L) lp[i0] = m0 * LN2 + logf(l0);
        if (i1 < L) lp[i1] = m1 * LN2 + logf(l1);
    }
}

// delta[s,h,i] = sum_d dO O     thread per (row, head)
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                                  float* __restrict__ delta, long long rows, int L, int heads) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * heads) return;
    const long long row = idx / heads;
    const int h = (int)(idx % heads);
    const uint4* a = reinterpret_cast<const uint4*>(out + row * heads * 32 + h * 32);
    const uint4* b = reinterpret_cast<const uint4*>(dout + row * heads * 32 + h * 32);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint4 x = a[c], y = b[c];
        const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 u = unpack_bf16x2(xs[e]), v = unpack_bf16x2(ys[e]);
            acc = fmaf(u.x, v.x, acc);
            acc = fmaf(u.y, v.y, acc);
        }
    }
    const long long sq = row / L;
    delta[(sq * heads + h) * L + (row % L)] = acc;
}

// dS for a 16x64 accumulator pair: sc <- p * (dp - delta), p = exp2(s*log2e + bias - lse)
// (rows = the thread's fixed index pair, columns = the looped index). ROWS_ARE_QUERIES selects
// which of (row, column) is the query for the bias / lse / delta lookups. MASKED adds the validity
// tests needed when the sequence or the row block is padded.
template <bool ROWS_ARE_QUERIES, bool MASKED>
__device__ __forceinline__ void ds_from_scores(float (&sc)[8][4], const float (&dpv)[8][4], const BiasIdx& bias,
                                               bool has_bias, int c0, int t, int r0, int r1, int L,
                                               float rstat0, float rstat1, float rdel0, float rdel1,
                                               const float* sL, const float* sD, bool keep_p, float (&pout)[8][4]) {
    // rows-are-queries: rstat = lse*log2e of the two rows, rdel = delta of the two rows
    // rows-are-keys   : per-column lse/delta come from sL/sD (shared memory)
    const int pr0 = bias.pos[min(r0, L - 1)], pr1 = bias.pos[min(r1, L - 1)];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int c = c0 + nt * 8 + 2 * t;
        float b[4] = {0.f, 0.f, 0.f, 0.f};
        if (has_bias) {
            const int2 pc = *reinterpret_cast<const int2*>(bias.pos + c);
            if (ROWS_ARE_QUERIES) {
                b[0] = bias.tab[pr0 + bias.off - pc.x]; b[1] = bias.tab[pr0 + bias.off - pc.y];
                b[2] = bias.tab[pr1 + bias.off - pc.x]; b[3] = bias.tab[pr1 + bias.off - pc.y];
            } else {
                b[0] = bias.tab[pc.x + bias.off - pr0]; b[1] = bias.tab[pc.y + bias.off - pr0];
                b[2] = bias.tab[pc.x + bias.off - pr1]; b[3] = bias.tab[pc.y + bias.off - pr1];
            }
        }
        float ls[4], dl[4];
        if (ROWS_ARE_QUERIES) {
            ls[0] = ls[1] = rstat0; ls[2] = ls[3] = rstat1;
            dl[0] = dl[1] = rdel0; dl[2] = dl[3] = rdel1;
        } else {
            const float2 l2 = *reinterpret_cast<const float2*>(sL + c);
            const float2 d2 = *reinterpret_cast<const float2*>(sD + c);
            ls[0] = ls[2] = l2.x; ls[1] = ls[3] = l2.y;
            dl[0] = dl[2] = d2.x; dl[1] = dl[3] = d2.y;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float p = fast_exp2(fmaf(sc[nt][e], LOG2E, b[e]) - ls[e]);
            if (MASKED) {
                const int cc = c + (e & 1);
                const int rr = e < 2 ? r0 : r1;
                if (!(cc < L && rr < L)) p = 0.f;
            }
            if (keep_p) pout[nt][e] = p;
            sc[nt][e] = p * (dpv[nt][e] - dl[e]);
        }
    }
}

// =============================================================================================
// backward, dq: grid (q blocks of 16 NW, heads, nseq). smem: K | V | Q tile | dO tile | pos | table
// =============================================================================================
template <int NW>
__global__ void __launch_bounds__(32 * NW, 2)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ table,
                   const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                   const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int L, int heads,
                   int gh, int gw) {
    constexpr int QB = 16 * NW, NT = 32 * NW;
    extern __shared__ __align__(128) uint8_t smem[];
    const int Lp = (L + 63) & ~63;
    uint8_t* sK = smem;
    uint8_t* sV = sK + Lp * 64;
    uint8_t* sQ = sV + Lp * 64;
    uint8_t* sdO = sQ + QB * 64;
    int* sP = reinterpret_cast<int*>(sdO + QB * 64);
    float* sT = reinterpret_cast<float*>(sP + Lp);
    const int qb = blockIdx.x, h = blockIdx.y, s = blockIdx.z;
    const int inner = heads * 32;
    const long long ld = 3LL * inner;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const __nv_bfloat16* base = qkv + (long long)s * L * ld + h * 32;
    load_tile(sK, base + inner, ld, Lp, L, tid, NT);
    load_tile(sV, base + 2 * inner, ld, Lp, L, tid, NT);
    load_tile(sQ, base + (long long)qb * QB * ld, ld, QB, L - qb * QB, tid, NT);
    load_tile(sdO, dout + ((long long)s * L + qb * QB) * inner + h * 32, inner, QB, L - qb * QB, tid, NT);
    const int n_off = (2 * gh - 1) * (2 * gw - 1);
    if (table)
        for (int i = tid; i < n_off; i += NT) sT[i] = table[(long long)h * n_off + i] * LOG2E;
    fill_pos(sP, Lp, L, gw, tid, NT);
    __syncthreads();
    const BiasIdx bias = make_bias(table ? sT : nullptr, sP, gh, gw);
    const uint32_t tK = smem_u32(sK), tV = smem_u32(sV);
    uint32_t qa[2][4], da[2][4];
    load_a_frags(qa, smem_u32(sQ), warp * 16, lane);
    load_a_frags(da, smem_u32(sdO), warp * 16, lane);
    const int g = lane >> 2, t = lane & 3;
    const int i0 = qb * QB + warp * 16 + g, i1 = i0 + 8;
    const float* lp = lse + ((long long)s * heads + h) * L;
    const float* dp = delta + ((long long)s * heads + h) * L;
    const float lse0 = (i0 < L ? lp[i0] : 0.f) * LOG2E, lse1 = (i1 < L ? lp[i1] : 0.f) * LOG2E;
    const float dl0 = i0 < L ? dp[i0] : 0.f, dl1 = i1 < L ? dp[i1] : 0.f;
    float dq[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) dq[a][0] = dq[a][1] = dq[a][2] = dq[a][3] = 0.f;
    for (int kb = 0; kb < Lp; kb += 64) {
        float sc[8][4], dpv[8][4];
        mma_16x64(sc, qa, tK, kb, lane);
        mma_16x64(dpv, da, tV, kb, lane);
        // padded key columns must contribute nothing; padded query rows only feed their own (unstored) dq
        if (Lp != L) ds_from_scores<true, true>(sc, dpv, bias, table != nullptr, kb, t, i0, i1, L, lse0, lse1, dl0, dl1,
                                                nullptr, nullptr, false, dpv);
        else ds_from_scores<true, false>(sc, dpv, bias, table != nullptr, kb, t, i0, i1, L, lse0, lse1, dl0, dl1,
                                         nullptr, nullptr, false, dpv);
        mma_acc_16x32(dq, sc, tK, kb, lane);
    }
    __nv_bfloat16* ob = dqkv + (long long)s * L * ld + h * 32;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int d = a * 8 + 2 * t;
        if (i0 < L) *reinterpret_cast<uint32_t*>(ob + (long long)i0 * ld + d) = pack_bf16x2(dq[a][0], dq[a][1]);
        if (i1 < L) *reinterpret_cast<uint32_t*>(ob + (long long)i1 * ld + d) = pack_bf16x2(dq[a][2], dq[a][3]);
    }
}

// =============================================================================================
// backward, dk & dv (transposed tiles: rows = keys): grid (kv blocks of 16 NW, heads, nseq).
// smem: Q [Lp][32] | dO [Lp][32] | K tile | V tile | lse [Lp] | delta [Lp] | pos | table
// =============================================================================================
template <int NW>
__global__ void __launch_bounds__(32 * NW, 2)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ table,
                    const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                    const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int L, int heads,
                    int gh, int gw) {
    constexpr int KB = 16 * NW, NT = 32 * NW;
    extern __shared__ __align__(128) uint8_t smem[];
    const int Lp = (L + 63) & ~63;
    uint8_t* sQ = smem;
    uint8_t* sdO = sQ + Lp * 64;
    uint8_t* sK = sdO + Lp * 64;
    uint8_t* sV = sK + KB * 64;
    float* sL = reinterpret_cast<float*>(sV + KB * 64);
    float* sD = sL + Lp;
    int* sP = reinterpret_cast<int*>(sD + Lp);
    float* sT = reinterpret_cast<float*>(sP + Lp);
    const int jb = blockIdx.x, h = blockIdx.y, s = blockIdx.z;
    const int inner = heads * 32;
    const long long ld = 3LL * inner;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const __nv_bfloat16* base = qkv + (long long)s * L * ld + h * 32;
    load_tile(sQ, base, ld, Lp, L, tid, NT);
    load_tile(sdO, dout + (long long)s * L * inner + h * 32, inner, Lp, L, tid, NT);
    load_tile(sK, base + inner + (long long)jb * KB * ld, ld, KB, L - jb * KB, tid, NT);
    load_tile(sV, base + 2 * inner + (long long)jb * KB * ld, ld, KB, L - jb * KB, tid, NT);
    const float* lp = lse + ((long long)s * heads + h) * L;
    const float* dp = delta + ((long long)s * heads + h) * L;
    for (int i = tid; i < Lp; i += NT) {
        sL[i] = i < L ? lp[i] * LOG2E : 0.f;
        sD[i] = i < L ? dp[i] : 0.f;
    }
    const int n_off = (2 * gh - 1) * (2 * gw - 1);
    if (table)
        for (int i = tid; i < n_off; i += NT) sT[i] = table[(long long)h * n_off + i] * LOG2E;
    fill_pos(sP, Lp, L, gw, tid, NT);
    __syncthreads();
    const BiasIdx bias = make_bias(table ? sT : nullptr, sP, gh, gw);
    const uint32_t tQ = smem_u32(sQ), tdO = smem_u32(sdO);
    uint32_t ka[2][4], va[2][4];
    load_a_frags(ka, smem_u32(sK), warp * 16, lane);
    load_a_frags(va, smem_u32(sV), warp * 16, lane);
    const int g = lane >> 2, t = lane & 3;
    const int j0 = jb * KB + warp * 16 + g, j1 = j0 + 8;       // this thread's two key rows
    float dk[4][4], dv[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        dk[a][0] = dk[a][1] = dk[a][2] = dk[a][3] = 0.f;
        dv[a][0] = dv[a][1] = dv[a][2] = dv[a][3] = 0.f;
    }
    for (int ib = 0; ib < Lp; ib += 64) {
        float st[8][4], dpt[8][4], pt[8][4];
        mma_16x64(st, ka, tQ, ib, lane);          // S^T[key, query] = K Q^T
        mma_16x64(dpt, va, tdO, ib, lane);        // dP^T[key, query] = V dO^T
        // st <- dS^T, pt <- P^T. Padded queries (columns) or keys (rows) must contribute nothing.
        if (Lp != L || (L % KB) != 0)
            ds_from_scores<false, true>(st, dpt, bias, table != nullptr, ib, t, j0, j1, L, 0.f, 0.f, 0.f, 0.f, sL, sD, true, pt);
        else
            ds_from_scores<false, false>(st, dpt, bias, table != nullptr, ib, t, j0, j1, L, 0.f, 0.f, 0.f, 0.f, sL, sD, true, pt);
        mma_acc_16x32(dv, pt, tdO, ib, lane);     // dV += P^T dO
        mma_acc_16x32(dk, st, tQ, ib, lane);      // dK += dS^T Q
    }
    __nv_bfloat16* ob = dqkv + (long long)s * L * ld + h * 32;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int d = a * 8 + 2 * t;
        if (j0 < L) {
            *reinterpret_cast<uint32_t*>(ob + (long long)j0 * ld + inner + d) = pack_bf16x2(dk[a][0], dk[a][1]);
            *reinterpret_cast<uint32_t*>(ob + (long long)j0 * ld + 2 * inner + d) = pack_bf16x2(dv[a][0], dv[a][1]);
        }
        if (j1 < L) {
            *reinterpret_cast<uint32_t*>(ob + (long long)j1 * ld + inner + d) = pack_bf16x2(dk[a][2], dk[a][3]);
            *reinterpret_cast<uint32_t*>(ob + (long long)j1 * ld + 2 * inner + d) = pack_bf16x2(dv[a][2], dv[a][3]);
        }
    }
}

// =============================================================================================
// backward, bias table: grid (kv blocks, q blocks, heads * nchunk); each CTA owns one 64x64 tile of
// dS, walks its chunk of the sequences with a 2-stage cp.async ring, accumulates dS in registers,
// folds it into a shared-memory copy of the (2gh-1)(2gw-1) table and flushes the touched entries.
// smem: 2 x {Q, dO, K, V tiles, lse[64], delta[64]} | pos | bias table | accumulation table
// =============================================================================================
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void load_tile_async(uint32_t dst, const __nv_bfloat16* src, long long ld, int valid,
                                                int tid) {
    for (int i = tid; i < 64 * 4; i += 128) {
        const int row = i >> 2, c = i & 3;
        const bool ok = row < valid;
        cp_async16(dst + tile_off(row, c), src + (long long)(ok ? row : 0) * ld + c * 8, ok);
    }
}

constexpr int DB_STAGE = 4 * 64 * 64 + 2 * 64 * 4;     // bytes per stage

__global__ void __launch_bounds__(128, 3)
attn_bwd_dbias_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ table,
                      const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                      const float* __restrict__ delta, float* __restrict__ dtable, int nseq, int L,
                      int heads, int gh, int gw, int nchunk) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int Lp = (L + 63) & ~63;
    const int n_off = (2 * gh - 1) * (2 * gw - 1);
    uint8_t* stage0 = smem;
    int* sP = reinterpret_cast<int*>(smem + 2 * DB_STAGE);
    float* sT = reinterpret_cast<float*>(sP + Lp);
    float* sA = sT + n_off;
    const int jb = blockIdx.x, qb = blockIdx.y;
    const int h = blockIdx.z / nchunk, chunk = blockIdx.z % nchunk;
    const int inner = heads * 32;
    const long long ld = 3LL * inner;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n_off; i += 128) { sT[i] = table[(long long)h * n_off + i] * LOG2E; sA[i] = 0.f; }
    fill_pos(sP, Lp, L, gw, tid, 128);
    const int per = (nseq + nchunk - 1) / nchunk;
    const int s_begin = chunk * per, s_end = min(nseq, s_begin + per);
    const int q_valid = L - qb * 64, k_valid = L - jb * 64;

    auto issue = [&](int s, int st) {
        const uint32_t sb = smem_u32(stage0 + st * DB_STAGE);
        const __nv_bfloat16* base = qkv + (long long)s * L * ld + h * 32;
        load_tile_async(sb, base + (long long)qb * 64 * ld, ld, q_valid, tid);
        load_tile_async(sb + 4096, dout + ((long long)s * L + qb * 64) * inner + h * 32, inner, q_valid, tid);
        load_tile_async(sb + 8192, base + inner + (long long)jb * 64 * ld, ld, k_valid, tid);
        load_tile_async(sb + 12288, base + 2 * inner + (long long)jb * 64 * ld, ld, k_valid, tid);
        if (tid < 64) {
            const int i = min(qb * 64 + tid, L - 1);
            cp_async4(sb + 16384 + tid * 4, lse + ((long long)s * heads + h) * L + i);
            cp_async4(sb + 16384 + 256 + tid * 4, delta + ((long long)s * heads + h) * L + i);
        }
        cp_async_commit();
    };

    const int g = lane >> 2, t = lane & 3;
    const int r0 = warp * 16 + g, r1 = r0 + 8;                 // tile-local query rows
    const int i0 = qb * 64 + r0, i1 = qb * 64 + r1;
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    if (s_begin < s_end) issue(s_begin, 0);
    for (int s = s_begin; s < s_end; ++s) {
        const int st = (s - s_begin) & 1;
        cp_async_wait<0>();
        __syncthreads();                          // stage `st` landed; everyone is done with the other stage
        if (s + 1 < s_end) issue(s + 1, st ^ 1);
        const uint8_t* sb = stage0 + st * DB_STAGE;
        const uint32_t tb = smem_u32(sb);
        const float* sLs = reinterpret_cast<const float*>(sb + 16384);
        const float* sDs = sLs + 64;
        uint32_t qa[2][4], da[2][4];
        load_a_frags(qa, tb, warp * 16, lane);
        load_a_frags(da, tb + 4096, warp * 16, lane);
        float sc[8][4], dpv[8][4];
        mma_16x64(sc, qa, tb + 8192, 0, lane);
        mma_16x64(dpv, da, tb + 12288, 0, lane);
        const BiasIdx bias = make_bias(sT, sP, gh, gw);
        // columns are addressed globally (jb*64 + ...) for the bias / validity tests
        const float lse0 = sLs[r0] * LOG2E, lse1 = sLs[r1] * LOG2E;
        if (Lp != L) ds_from_scores<true, true>(sc, dpv, bias, true, jb * 64, t, i0, i1, L, lse0, lse1, sDs[r0], sDs[r1],
                                                nullptr, nullptr, false, dpv);
        else ds_from_scores<true, false>(sc, dpv, bias, true, jb * 64, t, i0, i1, L, lse0, lse1, sDs[r0], sDs[r1],
                                         nullptr, nullptr, false, dpv);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nt][e] += sc[nt][e];
    }
    __syncthreads();
    const BiasIdx bias = make_bias(sT, sP, gh, gw);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = jb * 64 + nt * 8 + 2 * t + (e & 1);
            const int i = (e < 2) ? i0 : i1;
            if (i < L && j < L) atomicAdd(sA + bias.index(i, j), acc[nt][e]);
        }
    __syncthreads();
    float* dt = dtable + (long long)h * n_off;
    for (int i = tid; i < n_off; i += 128) {
        const float v = sA[i];
        if (v != 0.f) atomicAdd(dt + i, v);
    }
}

// =============================================================================================
// short sequences (L <= 32): CTA per sequence, warp per head, lane per query (fwd, dq) / key (dk, dv)
// smem: the sequence's packed qkv block [L][3*inner] bf16
// =============================================================================================
__device__ __forceinline__ void load_row32(const __nv_bfloat16* p, float (&v)[32]) {
    const uint4* s4 = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint4 u = s4[c];
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_bf16x2(w[e]);
            v[c * 8 + e * 2] = f.x;
            v[c * 8 + e * 2 + 1] = f.y;
        }
    }
}
__device__ __forceinline__ void store_row32(__nv_bfloat16* p, const float (&v)[32]) {
    uint4* d4 = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int c = 0; c < 4; ++c)
        d4[c] = make_uint4(pack_bf16x2(v[c * 8], v[c * 8 + 1]), pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]),
                           pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]), pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]));
}
__device__ __forceinline__ float dot32(const float (&a)[32], const float (&b)[32]) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) s = fmaf(a[d], b[d], s);
    return s;
}

__device__ __forceinline__ void load_row32f(const float* p, float (&v)[32]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 f = reinterpret_cast<const float4*>(p)[c];
        v[4 * c] = f.x; v[4 * c + 1] = f.y; v[4 * c + 2] = f.z; v[4 * c + 3] = f.w;
    }
}
__device__ __forceinline__ void stage_bf16_as_f32(float* dst, const __nv_bfloat16* src, int n8, int tid, int nthr) {
    for (int i = tid; i < n8; i += nthr) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + i);
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        reinterpret_cast<float4*>(dst)[2 * i] = make_float4(a.x, a.y, b.x, b.y);
        reinterpret_cast<float4*>(dst)[2 * i + 1] = make_float4(c.x, c.y, d.x, d.y);
    }
}

__global__ void __launch_bounds__(256)
attn_short_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                      float* __restrict__ lse, int L, int heads) {
    extern __shared__ __align__(16) uint8_t smem[];
    float* sq = reinterpret_cast<float*>(smem);             // fp32 copy of the sequence's qkv block
    const int s = blockIdx.x;
    const int inner = heads * 32, ld = 3 * inner;
    stage_bf16_as_f32(sq, qkv + (long long)s * L * ld, L * ld / 8, threadIdx.x, blockDim.x);
    __syncthreads();
    const int h = threadIdx.x >> 5, i = threadIdx.x & 31;
    if (h >= heads || i >= L) return;
    float q[32], o[32];
    load_row32f(sq + i * ld + h * 32, q);
#pragma unroll
    for (int d = 0; d < 32; ++d) o[d] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < L; ++j) {
        float kv[32];
        load_row32f(sq + j * ld + inner + h * 32, kv);
        const float x = dot32(q, kv) * LOG2E;
        const float mn = fmaxf(m, x);
        const float r = fast_exp2(m - mn), p = fast_exp2(x - mn);
        m = mn;
        l = l * r + p;
        load_row32f(sq + j * ld + 2 * inner + h * 32, kv);
#pragma unroll
        for (int d = 0; d < 32; ++d) o[d] = fmaf(p, kv[d], o[d] * r);
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < 32; ++d) o[d] *= inv;
    store_row32(out + ((long long)s * L + i) * inner + h * 32, o);
    lse[((long long)s * heads + h) * L + i] = m * LN2 + logf(l);
}

// smem (all fp32, converted once while staging): qkv block [L][3 inner] | dO block [L][inner] |
// P [heads][L][L+1] | dP [heads][L][L+1] | delta [heads][32]
// 3*heads warps. Phase 1: group 0 (lane = query) fills P = softmax probabilities, group 1 fills
// dP = dO v^T, group 2 computes delta. Phase 2: group 0 -> dq, group 1 -> dk, group 2 -> dv, each
// thread owning one output row with dS_ij = P_ij (dP_ij - delta_i) formed on the fly.
// One CTA = (sequence, group of HG <= 4 heads): 3*HG warps, ~70 KB of shared memory -> 2-3 CTAs / SM so
// the staging of one CTA overlaps the arithmetic of the others.
__global__ void __launch_bounds__(384, 2)
attn_short_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                      const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                      __nv_bfloat16* __restrict__ dqkv, int L, int heads, int HG) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int inner = heads * 32, ld = 3 * inner;
    const int gw = HG * 32;                                  // channels of the head group
    const int lds = 3 * gw;                                  // smem row: [q | k | v] of the group
    const int PS = L + 1;
    float* sq = reinterpret_cast<float*>(smem);
    float* sdo = sq + L * lds;
    float* sP = sdo + L * gw;
    float* sdP = sP + HG * L * PS;
    float* sD = sdP + HG * L * PS;
    const int ngroups = heads / HG;
    const int s = blockIdx.x / ngroups, hg = blockIdx.x % ngroups;
    const int h0 = hg * HG;
    // stage the group's q, k, v and dO as fp32: 8-element (16-byte) chunks
    {
        const int cpr = gw / 8;                              // chunks per (row, part)
        for (int i = threadIdx.x; i < L * 3 * cpr; i += blockDim.x) {
            const int c8 = i % cpr, part = (i / cpr) % 3, row = i / (3 * cpr);
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(
                qkv + ((long long)s * L + row) * ld + part * inner + h0 * 32 + c8 * 8));
            const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
            float4* dst = reinterpret_cast<float4*>(sq + row * lds + part * gw + c8 * 8);
            dst[0] = make_float4(a.x, a.y, b.x, b.y);
            dst[1] = make_float4(c.x, c.y, d.x, d.y);
        }
        for (int i = threadIdx.x; i < L * cpr; i += blockDim.x) {
            const int c8 = i % cpr, row = i / cpr;
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(dout + ((long long)s * L + row) * inner + h0 * 32 + c8 * 8));
            const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
            float4* dst = reinterpret_cast<float4*>(sdo + row * gw + c8 * 8);
            dst[0] = make_float4(a.x, a.y, b.x, b.y);
            dst[1] = make_float4(c.x, c.y, d.x, d.y);
        }
    }
    const int wg = (threadIdx.x >> 5) / HG;                 // warp group 0..2
    const int hl = (threadIdx.x >> 5) % HG, i = threadIdx.x & 31;
    const int h = h0 + hl;
    const bool active = i < L;
    float* Ph = sP + hl * L * PS;
    float* dPh = sdP + hl * L * PS;
    const float* qrow = sq + hl * 32;                       // + row * lds
    const float* krow = sq + gw + hl * 32;
    const float* vrow = sq + 2 * gw + hl * 32;
    const float* dorow = sdo + hl * 32;                     // + row * gw
    __syncthreads();
    if (active) {
        if (wg == 0) {
            float q[32];
            load_row32f(qrow + i * lds, q);
            const float li = lse[((long long)s * heads + h) * L + i] * LOG2E;
            for (int j = 0; j < L; ++j) {
                float kk[32];
                load_row32f(krow + j * lds, kk);
                Ph[i * PS + j] = fast_exp2(dot32(q, kk) * LOG2E - li);
            }
        } else if (wg == 1) {
            float dO[32];
            load_row32f(dorow + i * gw, dO);
            for (int j = 0; j < L; ++j) {
                float vv[32];
                load_row32f(vrow + j * lds, vv);
                dPh[i * PS + j] = dot32(dO, vv);
            }
        } else {
            float a[32], b[32];
            load_row32(out + ((long long)s * L + i) * inner + h * 32, a);
            load_row32f(dorow + i * gw, b);
            sD[hl * 32 + i] = dot32(a, b);
        }
    }
    __syncthreads();
    if (!active) return;
    float acc[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) acc[d] = 0.f;
    if (wg == 0) {                                           // dq_i = sum_j dS_ij k_j
        const float di = sD[hl * 32 + i];
        for (int j = 0; j < L; ++j) {
            float kk[32];
            load_row32f(krow + j * lds, kk);
            const float ds = Ph[i * PS + j] * (dPh[i * PS + j] - di);
#pragma unroll
            for (int d = 0; d < 32; ++d) acc[d] = fmaf(ds, kk[d], acc[d]);
        }
        store_row32(dqkv + ((long long)s * L + i) * ld + h * 32, acc);
    } else if (wg == 1) {                                    // dk_j = sum_i dS_ij q_i   (lane = j)
        for (int qi = 0; qi < L; ++qi) {
            float q[32];
            load_row32f(qrow + qi * lds, q);
            const float ds = Ph[qi * PS + i] * (dPh[qi * PS + i] - sD[hl * 32 + qi]);
#pragma unroll
            for (int d = 0; d < 32; ++d) acc[d] = fmaf(ds, q[d], acc[d]);
        }
        store_row32(dqkv + ((long long)s * L + i) * ld + inner + h * 32, acc);
    } else {                                                 // dv_j = sum_i P_ij dO_i   (lane = j)
        for (int qi = 0; qi < L; ++qi) {
            float dO[32];
            load_row32f(dorow + qi * gw, dO);
            const float pv = Ph[qi * PS + i];
#pragma unroll
            for (int d = 0; d < 32; ++d) acc[d] = fmaf(pv, dO[d], acc[d]);
        }
        store_row32(dqkv + ((long long)s * L + i) * ld + 2 * inner + h * 32, acc);
    }
}

// =============================================================================================
// backward of the qk l2norm + per-channel scale (epilogue CTK_EPI_QKV). Warp per (row, q|k head),
// lane = channel. In place on dqkv columns [0, 2*inner).
//   y = xhat * sc, sc = alpha*q_scale (q) or k_scale (k);  g = dy*sc;  dx = rnorm (g - xhat <xhat, g>)
// =============================================================================================
__global__ void __launch_bounds__(256)
qknorm_bwd_kernel(__nv_bfloat16* __restrict__ dqkv, const __nv_bfloat16* __restrict__ qkv,
                  const float* __restrict__ rnorm, const float* __restrict__ q_scale,
                  const float* __restrict__ k_scale, float alpha, float* __restrict__ dq_scale,
                  float* __restrict__ dk_scale, long long rows, int heads) {
    // 4 threads per (row, q|k head), 8 channels (16 bytes) each: fully coalesced 16-byte accesses,
    // the head-wide dot product is two shuffles. A thread's (slot type, channel group) is fixed
    // across its grid-stride loop (the stride is a multiple of 8*heads), so the per-channel scale
    // gradients accumulate in 8 registers and are folded through shared memory at the end.
    __shared__ float red[2][32];
    if (threadIdx.x < 64) red[threadIdx.x >> 5][threadIdx.x & 31] = 0.f;
    __syncthreads();
    const int inner = heads * 32;
    const long long ld = 3LL * inner;
    const int nslot = 2 * heads;
    const long long total = rows * nslot * 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long idx0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int grp = (int)(idx0 & 3);
    const int slot = (int)((idx0 >> 2) % nslot);
    const bool is_q = slot < heads;
    const float* scp = (is_q ? q_scale : k_scale) + grp * 8;
    const float a = is_q ? alpha : 1.f;
    float sc[8], acc[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) { sc[d] = __ldg(scp + d) * a; acc[d] = 0.f; }
    for (long long idx = idx0; idx < total; idx += stride) {
        const long long row = (idx >> 2) / nslot;
        const long long off = row * ld + (long long)slot * 32 + grp * 8;
        const uint4 yu = *reinterpret_cast<const uint4*>(qkv + off);
        const uint4 du = *reinterpret_cast<const uint4*>(dqkv + off);
        const uint32_t yw[4] = {yu.x, yu.y, yu.z, yu.w}, dw_[4] = {du.x, du.y, du.z, du.w};
        float xh[8], g[8];
        float dotv = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 yy = unpack_bf16x2(yw[e]), dd = unpack_bf16x2(dw_[e]);
            const float y2[2] = {yy.x, yy.y}, d2[2] = {dd.x, dd.y};
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int d = 2 * e + q;
                xh[d] = sc[d] != 0.f ? __fdividef(y2[q], sc[d]) : 0.f;
                acc[d] = fmaf(d2[q], xh[d], acc[d]);
                g[d] = d2[q] * sc[d];                          // d loss / d xhat
                dotv = fmaf(xh[d], g[d], dotv);
            }
        }
        dotv += __shfl_xor_sync(0xffffffffu, dotv, 1);
        dotv += __shfl_xor_sync(0xffffffffu, dotv, 2);
        const float rn = rnorm[row * nslot + slot];
        uint4 o;
        o.x = pack_bf16x2(rn * (g[0] - xh[0] * dotv), rn * (g[1] - xh[1] * dotv));
        o.y = pack_bf16x2(rn * (g[2] - xh[2] * dotv), rn * (g[3] - xh[3] * dotv));
        o.z = pack_bf16x2(rn * (g[4] - xh[4] * dotv), rn * (g[5] - xh[5] * dotv));
        o.w = pack_bf16x2(rn * (g[6] - xh[6] * dotv), rn * (g[7] - xh[7] * dotv));
        *reinterpret_cast<uint4*>(dqkv + off) = o;
    }
#pragma unroll
    for (int d = 0; d < 8; ++d) atomicAdd(&red[is_q ? 0 : 1][grp * 8 + d], acc[d] * a);
    __syncthreads();
    if (threadIdx.x < 32) atomicAdd(dq_scale + threadIdx.x, red[0][threadIdx.x]);
    else if (threadIdx.x < 64) atomicAdd(dk_scale + threadIdx.x - 32, red[1][threadIdx.x - 32]);
}

}  // namespace

static size_t long_smem(int L, int gh, int gw, bool bias, int row_block, bool bwd_dkv) {
    const int Lp = (L + 63) & ~63;
    size_t b = (size_t)Lp * 128 + (size_t)Lp * 4 + (bias ? (size_t)(2 * gh - 1) * (2 * gw - 1) * 4 : 0);
    b += bwd_dkv ? (size_t)row_block * 128 + (size_t)Lp * 8 : (size_t)row_block * 128;   // 2 tiles (+lse, delta)
    return b;
}

// CTK_ATTN_LEGACY=1 keeps the 24x24 spatial stack on the mma.sync kernels (A/B measurements)
static bool attn_legacy() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("CTK_ATTN_LEGACY"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

template <typename K>
static int set_smem(K kern, size_t bytes) {
    CTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return CTK_OK;
}

extern "C" int ctk_attn_fwd(const void* qkv, const float* table, void* out, float* lse, int nseq, int L,
                            int heads, int gh, int gw, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(qkv && out && lse && nseq > 0 && L > 0 && heads > 0 && heads <= 8, CTK_ERR_SHAPE, "attn_fwd: bad args");
    CTK_REQUIRE(!table || gh * gw == L, CTK_ERR_SHAPE, "attn_fwd: bias table needs L == gh*gw");
    CTK_REQUIRE(CTK_ALIGNED(qkv, 16) && CTK_ALIGNED(out, 16), CTK_ERR_ALIGN, "attn_fwd: alignment");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    auto q = reinterpret_cast<const __nv_bfloat16*>(qkv);
    auto o = reinterpret_cast<__nv_bfloat16*>(out);
    if (!table && ctk_attn_seq24_supported(L, heads) && !attn_legacy()) return ctk_attn_seq24_fwd(qkv, out, lse, nseq, heads, s);
    if (!table && L <= 32) {
        const size_t sm = (size_t)L * 3 * heads * 32 * 4;
        CTK_CUDA(cudaFuncSetAttribute(attn_short_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        attn_short_fwd_kernel<<<nseq, heads * 32, sm, s>>>(q, o, lse, L, heads);
        CTK_LAUNCH_CHECK();
        return CTK_OK;
    }
    if (table && L == 576 && gh == 24 && gw == 24 && !attn_legacy()) return ctk_attn_fwd_tc(qkv, table, out, lse, nseq, heads, s);
    // legacy mma.sync path for other shapes. 9 warps (144 queries) per CTA when the sequence is long
    // enough: 576 = 4 x 144, 2 CTAs / SM
    if (L >= 144) {
        const size_t sm = long_smem(L, gh, gw, table != nullptr, 144, false) - 144 * 64;   // one tile (Q) only
        CTK_REQUIRE(sm <= 220 * 1024, CTK_ERR_SHAPE, "attn_fwd: sequence of %d tokens does not fit in shared memory", L);
        if ((rc = set_smem(attn_fwd_kernel<9>, sm))) return rc;
        attn_fwd_kernel<9><<<dim3((L + 143) / 144, heads, nseq), 288, sm, s>>>(q, table, o, lse, L, heads, gh, gw);
    } else {
        const size_t sm = long_smem(L, gh, gw, table != nullptr, 64, false) - 64 * 64;
        if ((rc = set_smem(attn_fwd_kernel<4>, sm))) return rc;
        attn_fwd_kernel<4><<<dim3((L + 63) / 64, heads, nseq), 128, sm, s>>>(q, table, o, lse, L, heads, gh, gw);
    }
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

extern "C" int ctk_attn_bwd(const void* qkv, const float* table, const void* out, const void* dout,
                            const float* lse, float* delta, void* dqkv, float* dtable, int nseq, int L,
                            int heads, int gh, int gw, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(qkv && out && dout && lse && delta && dqkv && nseq > 0 && L > 0 && heads > 0 && heads <= 8,
                CTK_ERR_SHAPE, "attn_bwd: bad args");
    CTK_REQUIRE(!table || (gh * gw == L && dtable), CTK_ERR_SHAPE, "attn_bwd: bias table needs L == gh*gw and dtable");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    auto q = reinterpret_cast<const __nv_bfloat16*>(qkv);
    auto o = reinterpret_cast<const __nv_bfloat16*>(out);
    auto d_o = reinterpret_cast<const __nv_bfloat16*>(dout);
    auto dq = reinterpret_cast<__nv_bfloat16*>(dqkv);
    const int inner = heads * 32;
    if (!table && ctk_attn_seq24_supported(L, heads) && !attn_legacy())
        return ctk_attn_seq24_bwd(qkv, out, dout, lse, dqkv, nseq, heads, s);
    if (!table && L <= 32) {
        const int HG = heads % 4 == 0 ? 4 : (heads % 2 == 0 ? 2 : 1);
        const size_t sm = (size_t)L * 3 * HG * 32 * 4 + (size_t)L * HG * 32 * 4 + 2 * (size_t)HG * L * (L + 1) * 4 +
                          (size_t)HG * 32 * 4;
        CTK_CUDA(cudaFuncSetAttribute(attn_short_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        attn_short_bwd_kernel<<<nseq * (heads / HG), HG * 96, sm, s>>>(q, o, d_o, lse, dq, L, heads, HG);
        CTK_LAUNCH_CHECK();
        return CTK_OK;
    }
    const long long rows = (long long)nseq * L;
    attn_delta_kernel<<<(unsigned)((rows * heads + 255) / 256), 256, 0, s>>>(o, d_o, delta, rows, L, heads);
    CTK_LAUNCH_CHECK();
    const bool hb = table != nullptr;
    const bool tc_path = hb && L == 576 && gh == 24 && gw == 24 && !attn_legacy();
    static int dbias_tc = -1;
    if (dbias_tc < 0) { const char* e = getenv("CTK_DBIAS_TC"); dbias_tc = (e && e[0] == '0') ? 0 : 1; }
    if (tc_path) {
        if ((rc = ctk_attn_bwd_tc(qkv, table, dout, lse, delta, dqkv, dtable, dbias_tc, nseq, heads, s))) return rc;
        if (dbias_tc) return CTK_OK;
    } else if (L >= 128) {
        // dq: 8 warps (128 queries), dk/dv: 6 warps (96 keys; 576 = 6 x 96); 2 CTAs / SM each
        const size_t sm_dq = long_smem(L, gh, gw, hb, 128, false), sm_dkv = long_smem(L, gh, gw, hb, 96, true);
        CTK_REQUIRE(sm_dq <= 220 * 1024 && sm_dkv <= 220 * 1024, CTK_ERR_SHAPE,
                    "attn_bwd: sequence of %d tokens does not fit in shared memory", L);
        if ((rc = set_smem(attn_bwd_dq_kernel<8>, sm_dq))) return rc;
        if ((rc = set_smem(attn_bwd_dkv_kernel<6>, sm_dkv))) return rc;
        attn_bwd_dq_kernel<8><<<dim3((L + 127) / 128, heads, nseq), 256, sm_dq, s>>>(q, table, d_o, lse, delta, dq, L, heads, gh, gw);
        CTK_LAUNCH_CHECK();
        attn_bwd_dkv_kernel<6><<<dim3((L + 95) / 96, heads, nseq), 192, sm_dkv, s>>>(q, table, d_o, lse, delta, dq, L, heads, gh, gw);
        CTK_LAUNCH_CHECK();
    } else {
        const size_t sm_dq = long_smem(L, gh, gw, hb, 64, false), sm_dkv = long_smem(L, gh, gw, hb, 64, true);
        if ((rc = set_smem(attn_bwd_dq_kernel<4>, sm_dq))) return rc;
        if ((rc = set_smem(attn_bwd_dkv_kernel<4>, sm_dkv))) return rc;
        const dim3 grid((L + 63) / 64, heads, nseq);
        attn_bwd_dq_kernel<4><<<grid, 128, sm_dq, s>>>(q, table, d_o, lse, delta, dq, L, heads, gh, gw);
        CTK_LAUNCH_CHECK();
        attn_bwd_dkv_kernel<4><<<grid, 128, sm_dkv, s>>>(q, table, d_o, lse, delta, dq, L, heads, gh, gw);
        CTK_LAUNCH_CHECK();
    }
    if (table) {
        const int nb = (L + 63) / 64;
        const int Lp = nb * 64;
        int nchunk = nseq / 16;
        if (nchunk < 1) nchunk = 1;
        if (nchunk > 8) nchunk = 8;
        const size_t sm = 2 * (size_t)DB_STAGE + (size_t)Lp * 4 + 2 * (size_t)(2 * gh - 1) * (2 * gw - 1) * 4;
        if ((rc = set_smem(attn_bwd_dbias_kernel, sm))) return rc;
        attn_bwd_dbias_kernel<<<dim3(nb, nb, heads * nchunk), 128, sm, s>>>(q, table, d_o, lse, delta, dtable, nseq, L,
                                                                           heads, gh, gw, nchunk);
        CTK_LAUNCH_CHECK();
    }
    return CTK_OK;
}

extern "C" int ctk_qknorm_bwd(void* dqkv, const void* qkv, const float* rnorm, const float* q_scale,
                              const float* k_scale, float alpha, float* dq_scale, float* dk_scale,
                              long long rows, int heads, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(dqkv && qkv && rnorm && q_scale && k_scale && dq_scale && dk_scale && rows > 0 && heads > 0,
                CTK_ERR_SHAPE, "qknorm_bwd: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    CTK_REQUIRE(256 % (8 * heads) == 0, CTK_ERR_SHAPE, "qknorm_bwd: heads must divide 32");
    long long blocks = (rows * 8 * heads + 255) / 256;
    const long long cap = (long long)ctk_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    qknorm_bwd_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<__nv_bfloat16*>(dqkv),
                                                       reinterpret_cast<const __nv_bfloat16*>(qkv), rnorm,
                                                       q_scale, k_scale, alpha, dq_scale, dk_scale, rows, heads);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

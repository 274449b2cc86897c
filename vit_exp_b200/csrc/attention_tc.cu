// Spatial cosine attention (attention.py:162-184) on the 5th-generation tensor cores: tcgen05.mma with
// TMEM accumulators, operands staged by TMA, for the CT-CLIP spatial stack (24 x 24 = 576 tokens
// per slice, head dim 32, continuous position bias).  Other shapes stay on attention.cu.
//
// Work unit = one 128-query tile of one (slice, head); a CTA owns a contiguous range of units so
// K/V (72 KB, three TMA boxes of 192 keys each) are staged once per (slice, head).  Two CTAs per SM
// (100 KB shared memory, 256 TMEM columns each): while one CTA waits for its MMAs the other one
// keeps the SFU busy - the kernel is bound by one exp2 per logit (head dim 32 gives only 128
// tensor FLOPs per exponential), not by the tensor pipe.
//
// Per tile and 96-key chunk c (= 4 rows of the token grid), logits double-buffered in TMEM:
//   control thread : S = Q K_c^T        tcgen05.mma SS, M128 N96 K32 (two k-steps), D = TMEM[96*(c&1), +96)
//   8 softmax warps: p = exp2(s*log2e + bias - M_i), written back as bf16 pairs over the columns the
//                    logits were read from (tcgen05.ld / tcgen05.st; warp w owns TMEM lanes 32*(w%4),
//                    warps 1-4 the first 48 keys of the chunk, warps 5-8 the last 48)
//   control thread : O += P V_c         tcgen05.mma TS (A = P in TMEM), M128 N32 K96, D = TMEM[192,224)
// S of chunk c+2 is issued right behind PV of chunk c (the tensor pipe executes in order), across
// tile boundaries too, so the softmax warps find their next chunk waiting.
// M_i = |q_i| max_j|k_j| + max(bias) bounds every logit of the row (Cauchy-Schwarz), so there is no
// running-max pass and no rescaling of O; rows whose bound could be more than 2^64 above the true
// maximum (never at the scales CT-CLIP trains with) take an exact SIMT maximum instead.
// The bias index of (i, j) is pos(i) - pos(j) + const with pos(n) = (n/24)*56 + n%24 (table rows are
// padded from 47 to 56 = 24 mod 32 words so that 32 consecutive tokens hit 32 different banks); with
// 48 keys = 2 grid rows per warp pass, pos(j) is a compile-time immediate of the unrolled loop.
#include "attention_tc.cuh"

using namespace attn_tc;

namespace {

// exact max_j (q_i.k_j log2e + bias') by SIMT dot products out of the swizzled tiles (rare path)
__device__ __noinline__ float exact_row_max(const uint8_t* sQrow, int qrow, const uint8_t* sK, uint32_t bias_addr) {
    float q[32];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint4 u = *reinterpret_cast<const uint4*>(sQrow + ((c ^ ((qrow >> 1) & 3)) << 4));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_bf16x2(w[e]);
            q[c * 8 + 2 * e] = f.x;
            q[c * 8 + 2 * e + 1] = f.y;
        }
    }
    float m = -INFINITY;
    for (int j = 0; j < TL; ++j) {
        const int r = j % CH;
        const uint8_t* kr = sK + (j / CH) * CH_BYTES + r * 64;
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint4 u = *reinterpret_cast<const uint4*>(kr + ((c ^ ((r >> 1) & 3)) << 4));
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = unpack_bf16x2(w[e]);
                dot = fmaf(q[c * 8 + 2 * e], f.x, dot);
                dot = fmaf(q[c * 8 + 2 * e + 1], f.y, dot);
            }
        }
        m = fmaxf(m, fmaf(dot, LOG2E, lds_f32(bias_addr - 4u * (uint32_t)((j / TGW) * TPW + j % TGW))));
    }
    return m;
}

// NE logits (local keys J0 .. J0+NE-1 of the warp's 48) -> probabilities: row sum and bf16 pairs
template <int J0, int NE>
__device__ __forceinline__ void softmax_block(const uint32_t* r, uint32_t* pk, uint32_t bias_addr, float negM, float& l) {
#pragma unroll
    for (int e = 0; e < NE; e += 2) {
        const int j0 = J0 + e, j1 = J0 + e + 1;
        const float b0 = lds_f32(bias_addr - 4u * (uint32_t)((j0 / TGW) * TPW + j0 % TGW));
        const float b1 = lds_f32(bias_addr - 4u * (uint32_t)((j1 / TGW) * TPW + j1 % TGW));
        const float p0 = fast_exp2(fmaf(__uint_as_float(r[e]), LOG2E, negM) + b0);
        const float p1 = fast_exp2(fmaf(__uint_as_float(r[e + 1]), LOG2E, negM) + b1);
        l += p0 + p1;
        pk[e >> 1] = pack_bf16x2(p0, p1);
    }
}

__global__ void __launch_bounds__(NTHR, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                   const float* __restrict__ table, __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                   int nseq, int heads) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sK = smem + OFF_K;
    uint8_t* sV = smem + OFF_V;
    uint8_t* sQ = smem + OFF_Q;
    float* sT = reinterpret_cast<float*>(smem + OFF_T);
    float* sLp = reinterpret_cast<float*>(smem + OFF_LP);
    float* sRed = reinterpret_cast<float*>(smem + OFF_RED);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* bar_K = bars;            // [3] TMA -> MMA / softmax
    uint64_t* bar_V = bars + 3;        // [3] TMA -> MMA
    uint64_t* bar_Q = bars + 6;        // [2] TMA -> MMA / softmax
    uint64_t* bar_S = bars + 8;        // [2] MMA -> softmax: logits of a chunk are in TMEM buffer b
    uint64_t* bar_P = bars + 10;       // [2] softmax -> MMA: probabilities of a chunk are in TMEM buffer b
    uint64_t* bar_O = bars + 12;       // MMA -> softmax: the tile's O is complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_SLOT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int inner = heads * 32;
    // contiguous range of tiles per CTA, balanced by work: a full tile weighs 2, the 64-query tail tile 1
    constexpr int ITEM_W = 2 * (NQT - 1) + 1;
    const long long W = (long long)nseq * heads * ITEM_W;
    auto tile_at = [&](long long w) { return (int)((w / ITEM_W) * NQT + (w % ITEM_W + 1) / 2); };
    const int g0 = tile_at(W * blockIdx.x / gridDim.x), g1 = tile_at(W * (blockIdx.x + 1) / gridDim.x);

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&tmap_q);
            tma_prefetch_desc(&tmap_kv);
            for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
            mbar_init(&bar_S[0], 1);
            mbar_init(&bar_S[1], 1);
            mbar_init(&bar_P[0], NSOFT);
            mbar_init(&bar_P[1], NSOFT);
            mbar_init(bar_O, 1);
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ control: TMA + MMA issue ================================
        if (lane == 0 && g0 < g1) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(QT, SC, 0, 0);
            constexpr uint32_t idesc_pv = umma_idesc_bf16(QT, 32, 0, 1);
            auto load_k = [&](int item) {
                const int s = item % nseq, h = item / nseq;
                for (int c = 0; c < NCH; ++c) {
                    mbar_expect_tx(&bar_K[c], CH_BYTES);
                    tma_load_2d(sK + c * CH_BYTES, &tmap_kv, &bar_K[c], inner + h * 32, s * TL + c * CH);
                }
            };
            auto load_v = [&](int item) {
                const int s = item % nseq, h = item / nseq;
                for (int c = 0; c < NCH; ++c) {
                    mbar_expect_tx(&bar_V[c], CH_BYTES);
                    tma_load_2d(sV + c * CH_BYTES, &tmap_kv, &bar_V[c], 2 * inner + h * 32, s * TL + c * CH);
                }
            };
            auto load_q = [&](int g, int buf) {
                const int item = g / NQT, t = g % NQT;
                const int s = item % nseq, h = item / nseq;
                mbar_expect_tx(&bar_Q[buf], Q_BYTES);
                tma_load_2d(sQ + buf * Q_BYTES, &tmap_q, &bar_Q[buf], h * 32, s * TL + t * QT);
            };
            const uint32_t sK_a = smem_u32(sK), sV_a = smem_u32(sV), sQ_a = smem_u32(sQ);
            const int ntiles = g1 - g0;
            // S(n, c): logits of chunk c of local tile n into TMEM buffer c & 1.  K parity: item ordinal.
            auto issue_s = [&](int n, int c) {
                const int g = g0 + n, buf = n & 1;
                const int item_ord = g / NQT - g0 / NQT;
                if (c == 0) mbar_wait(&bar_Q[buf], (uint32_t)(n >> 1) & 1u);
                mbar_wait(&bar_K[c >> 1], (uint32_t)item_ord & 1u);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    tc_mma_f16(tmem_base + COL_S + (c & 1) * SC, umma_desc(sQ_a + buf * Q_BYTES + k * 32, 16, 512, SW64),
                               umma_desc(sK_a + c * (SC * 64) + k * 32, 16, 512, SW64), idesc_s, k);
                tc_commit(&bar_S[c & 1]);
            };
            load_k(g0 / NQT);
            load_q(g0, 0);
            load_v(g0 / NQT);
            if (ntiles > 1) load_q(g0 + 1, 1);
            uint32_t p_par = 0, o_par = 0;
            bool restart = true;                          // the S pipeline is empty (start, or after an item switch)
            for (int n = 0; n < ntiles; ++n) {
                const int g = g0 + n, item = g / NQT;
                const int item_ord = item - g0 / NQT;
                const bool last_of_item = (n + 1 == ntiles) || ((g + 1) / NQT != item);
                if (restart) { issue_s(n, 0); issue_s(n, 1); restart = false; }
#pragma unroll 1
                for (int c = 0; c < NSC; ++c) {
                    mbar_wait(&bar_P[c & 1], (p_par >> (c & 1)) & 1u);
                    p_par ^= 1u << (c & 1);
                    mbar_wait(&bar_V[c >> 1], (uint32_t)item_ord & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < SC / 16; ++k) {
                        // P: keys [0,48) of the chunk in columns [0,24) of its buffer, keys [48,96) in columns [48,72)
                        const uint32_t a_col = COL_S + (c & 1) * SC + (k < 3 ? k * 8 : 48 + (k - 3) * 8);
                        tc_mma_f16_ts(tmem_base + COL_O, tmem_base + a_col,
                                      umma_desc(sV_a + c * (SC * 64) + k * 1024, 512, 512, SW64), idesc_pv,
                                      (c > 0 || k > 0) ? 1u : 0u);
                    }
                    if (c == NSC - 1) {
                        tc_commit(bar_O);
                        // every S MMA of this tile has retired (its last chunk was consumed): the Q buffer
                        // and, after the last tile of the item, the K boxes can be refilled
                        if (n + 2 < ntiles) load_q(g + 2, n & 1);
                        if (last_of_item && n + 1 < ntiles) load_k(item + 1);
                    }
                    // logits two chunks ahead, behind this PV in the in-order tensor pipe
                    if (c + 2 < NSC) issue_s(n, c + 2);
                    else if (!last_of_item) issue_s(n + 1, c + 2 - NSC);
                }
                if (last_of_item && n + 1 < ntiles) {
                    mbar_wait(bar_O, o_par);        // the item's last PV has retired: V can be refilled
                    load_v(item + 1);
                    restart = true;
                }
                o_par ^= 1;
            }
        }
    } else {
        // ================================ softmax warps ================================
        const int sw = warp - 1;
        const int quarter = warp & 3;               // TMEM lane quarter this warp may access
        const int hsel = sw >> 2;                   // which 48 keys of a chunk
        const int row = quarter * 32 + lane;
        const int st = threadIdx.x - 32;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t sT_a = smem_u32(sT);
        uint32_t kv_par = 1, s_par = 0, o_par = 0, q_par = 0;        // bit b of s_par / q_par: parity of buffer b
        int cur_item = -1, cur_head = -1, n = 0;
        float kmax = 0.f, bmax = 0.f, bmin = 0.f;
        for (int g = g0; g < g1; ++g, ++n) {
            const int item = g / NQT, t = g % NQT, buf = n & 1;
            const int s = item % nseq, h = item / nseq;         // items are head-major: the table rarely changes
            if (item != cur_item) {
                cur_item = item;
                kv_par ^= 1;
                soft_sync();                         // everyone is done with the previous item (table, K)
                float mx = bmax, mn = bmin;
                if (h != cur_head) {
                    cur_head = h;
                    mx = -INFINITY; mn = INFINITY;
                    const float* tg = table + (long long)h * TNOFF;
                    constexpr int NLD = (TNOFF + 32 * NSOFT - 1) / (32 * NSOFT);
                    float tv[NLD];
#pragma unroll
                    for (int u = 0; u < NLD; ++u) {
                        const int i = st + u * 32 * NSOFT;
                        tv[u] = i < TNOFF ? __ldg(tg + i) * LOG2E : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < NLD; ++u) {
                        const int i = st + u * 32 * NSOFT;
                        if (i < TNOFF) {
                            sT[(i / TWW) * TPW + i % TWW] = tv[u];
                            mx = fmaxf(mx, tv[u]);
                            mn = fminf(mn, tv[u]);
                        }
                    }
                }
                for (int c = 0; c < NCH; ++c) mbar_wait(&bar_K[c], kv_par);
                float k2 = 0.f;
                for (int r = st; r < TL; r += 32 * NSOFT) {
                    const uint4* kr = reinterpret_cast<const uint4*>(sK + (r / CH) * CH_BYTES + (r % CH) * 64);
                    float a = 0.f;
#pragma unroll
                    for (int c = 0; c < 4; ++c) a = sumsq16(kr[c], a);
                    k2 = fmaxf(k2, a);
                }
                mx = warp_max(mx);
                mn = -warp_max(-mn);
                k2 = warp_max(k2);
                if (lane == 0) { sRed[sw] = mx; sRed[NSOFT + sw] = mn; sRed[2 * NSOFT + sw] = k2; }
                soft_sync();
                mx = sRed[0]; mn = sRed[NSOFT]; k2 = sRed[2 * NSOFT];
#pragma unroll
                for (int w = 1; w < NSOFT; ++w) {
                    mx = fmaxf(mx, sRed[w]);
                    mn = fminf(mn, sRed[NSOFT + w]);
                    k2 = fmaxf(k2, sRed[2 * NSOFT + w]);
                }
                bmax = mx; bmin = mn; kmax = sqrtf(k2);
            }
            mbar_wait(&bar_Q[buf], (q_par >> buf) & 1u);
            q_par ^= 1u << buf;
            const int i = t * QT + row;
            const bool active = t * QT + quarter * 32 < TL;          // warp-uniform
            float Mi = 0.f, l = 0.f;
            uint32_t bias_row = sT_a;
            if (active) {
                const uint8_t* qrow = sQ + buf * Q_BYTES + row * 64;
                float q2 = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) q2 = sumsq16(reinterpret_cast<const uint4*>(qrow)[c], q2);
                const float qk = sqrtf(q2) * kmax * LOG2E;
                Mi = qk + bmax;
                bias_row = sT_a + 4u * (uint32_t)((i / TGW) * TPW + (i % TGW) + TOFF);
                if (__any_sync(0xffffffffu, 2.f * qk + (bmax - bmin) > 64.f)) Mi = exact_row_max(qrow, row, sK, bias_row);
            }
            const float negM = -Mi;
#pragma unroll 1
            for (int c = 0; c < NSC; ++c) {
                mbar_wait(&bar_S[c & 1], (s_par >> (c & 1)) & 1u);
                s_par ^= 1u << (c & 1);
                tc_fence_after();
                if (active) {
                    const uint32_t bb = bias_row - 4u * (uint32_t)((c * 4 + hsel * 2) * TPW);
                    const uint32_t t_s = t_lane + COL_S + (c & 1) * SC + hsel * 48;
                    uint32_t r[32], r2[16], pk[16];
                    tc_ld_32x32(t_s, r);
                    tc_ld_32x32_x16(t_s + 32, r2);
                    tc_wait_ld();
                    softmax_block<0, 32>(r, pk, bb, negM, l);
                    tc_st_32x32_x16(t_s, pk);
                    softmax_block<32, 16>(r2, pk, bb, negM, l);
                    tc_st_32x32_x8(t_s + 16, pk);
                    tc_wait_st();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_P[c & 1]);
            }
            if (active) sLp[hsel * QT + row] = l;
            soft_sync();
            mbar_wait(bar_O, o_par);
            o_par ^= 1;
            tc_fence_after();
            if (active) {
                const float lt = sLp[row] + sLp[QT + row];
                const float inv = 1.f / lt;
                uint32_t o[16];
                tc_ld_32x32_x16(t_lane + COL_O + hsel * 16, o);
                tc_wait_ld();
                uint4 w0, w1;
                w0.x = pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
                w0.y = pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
                w0.z = pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
                w0.w = pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
                w1.x = pack_bf16x2(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
                w1.y = pack_bf16x2(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
                w1.z = pack_bf16x2(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
                w1.w = pack_bf16x2(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
                uint4* op = reinterpret_cast<uint4*>(out + ((long long)s * TL + i) * inner + h * 32 + hsel * 16);
                op[0] = w0;
                op[1] = w1;
                if (hsel == 0) lse[((long long)s * heads + h) * TL + i] = Mi * LN2 + __logf(lt);
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

// Forward for the 24x24 spatial stack; the caller (ctk_attn_fwd) has validated the arguments.
int ctk_attn_fwd_tc(const void* qkv, const float* table, void* out, float* lse, int nseq, int heads, cudaStream_t stream) {
    const int inner = heads * 32;
    const unsigned long long rows = (unsigned long long)nseq * TL;
    CUtensorMap tq, tkv;
    const unsigned long long dims[2] = {(unsigned long long)(3 * inner), rows};
    const unsigned long long strides[1] = {(unsigned long long)(3 * inner) * 2};
    const unsigned int box_q[2] = {32, QT}, box_kv[2] = {32, CH};
    int rc;
    if ((rc = ctk_make_tmap(&tq, qkv, false, 2, dims, strides, box_q, 2))) return rc;
    if ((rc = ctk_make_tmap(&tkv, qkv, false, 2, dims, strides, box_kv, 2))) return rc;
    CTK_SET_MAX_SMEM(attn_fwd_tc_kernel, SMEM_BYTES);
    const long long G = (long long)nseq * heads * NQT;
    long long grid = 2LL * ctk_num_sms();
    if (grid > G) grid = G;
    attn_fwd_tc_kernel<<<(unsigned)grid, NTHR, SMEM_BYTES, stream>>>(tq, tkv, table, reinterpret_cast<__nv_bfloat16*>(out), lse,
                                                                     nseq, heads);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

// Backward of the spatial cosine attention (attention.py:162-184) on tcgen05/TMEM, 24 x 24 token
// slices, head dim 32.  One kernel template, two launches:
//   DKV = false : rows = queries.  dQ_t   = sum_c dS K_c             (dS = P o (dP - delta))
//   DKV = true  : rows = keys.     dV_t   = sum_c P^T dO_c,  dK_t = sum_c dS^T Q_c
// with P = exp(S + bias - lse) recomputed from the saved log-sum-exp, S = Q K^T, dP = dO V^T.
// Both have the structure of the forward kernel (attention_tc.cu): per 128-row tile and 96-column
// chunk c, the control thread issues X1 = T1 R1_c^T and X2 = T2 R2_c^T (tcgen05.mma SS, K = 32) into
// TMEM, the 8 softmax warps turn them into bf16 P / dS written over the columns they were read from,
// and the control thread feeds those back as the TMEM A operand of the accumulating products
// (tcgen05.mma TS, M128 N32 K96).  "Resident" operands R1, R2 (576 rows each: K, V for dQ; Q, dO
// for dK/dV) are staged once per (slice, head) by TMA, the tile operands T1, T2 once per tile.
//   DKV = false : T1 = Q_t, T2 = dO_t, R1 = K, R2 = V;   acc0 += dS R1_c
//   DKV = true  : T1 = K_t, T2 = V_t,  R1 = Q, R2 = dO;  acc0 += dS^T R1_c (dK), acc1 += P^T R2_c (dV)
// TMEM columns: X1 [0,96)  X2 [96,192)  acc0 [192,224)  acc1 [224,256); two CTAs per SM.
// The bias-table gradient is attn_dbias_tc_kernel below (CTK_DBIAS_TC=0 selects attn_bwd_dbias_kernel of attention.cu).
#include "attention_tc.cuh"

using namespace attn_tc;

namespace {

constexpr int B_OFF_R1 = 0;
constexpr int B_OFF_R2 = B_OFF_R1 + NCH * CH_BYTES;
constexpr int B_OFF_T1 = B_OFF_R2 + NCH * CH_BYTES;
constexpr int B_OFF_T2 = B_OFF_T1 + Q_BYTES;
constexpr int B_OFF_TAB = B_OFF_T2 + Q_BYTES;
constexpr int B_OFF_L = B_OFF_TAB + TWW * TPW * 4;       // lse * log2e of the item's queries (DKV)
constexpr int B_OFF_D = B_OFF_L + TL * 4;                // delta of the item's queries (DKV)
constexpr int B_OFF_BAR = B_OFF_D + TL * 4;
constexpr int B_NBAR = 3 + 1 + 3;
constexpr int B_OFF_SLOT = B_OFF_BAR + B_NBAR * 8;
constexpr int B_SMEM_BYTES = B_OFF_SLOT + 16 + 1024;
constexpr uint32_t COL_X1 = 0, COL_X2 = 96, COL_A0 = 192, COL_A1 = 224;

__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// 16 columns (local column index J0 .. J0+15 of the warp's 48) of X1 / X2 -> bf16 pairs of P and dS
template <bool DKV, int J0>
__device__ __forceinline__ void bwd_block(const uint32_t (&x1)[16], const uint32_t (&x2)[16], uint32_t (&pp)[8],
                                          uint32_t (&ds)[8], uint32_t bias_addr, uint32_t stat_addr, float lse2,
                                          float delta) {
    float ls[16], dl[16];
    if (DKV) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 a = lds_f32x4(stat_addr + 4u * (uint32_t)(J0 + 4 * q));
            const float4 b = lds_f32x4(stat_addr + 4u * (uint32_t)(TL + J0 + 4 * q));
            ls[4 * q] = a.x; ls[4 * q + 1] = a.y; ls[4 * q + 2] = a.z; ls[4 * q + 3] = a.w;
            dl[4 * q] = b.x; dl[4 * q + 1] = b.y; dl[4 * q + 2] = b.z; dl[4 * q + 3] = b.w;
        }
    }
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
        float p[2], d[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = J0 + e + u;
            const uint32_t pos = 4u * (uint32_t)((j / TGW) * TPW + j % TGW);
            const float b = lds_f32(DKV ? bias_addr + pos : bias_addr - pos);
            // DKV: ls = natural-log lse of the column's query (staged raw by the bulk copy); else lse2 = lse * log2e
            p[u] = DKV ? fast_exp2(fmaf(__uint_as_float(x1[e + u]) - ls[e + u], LOG2E, b))
                       : fast_exp2(fmaf(__uint_as_float(x1[e + u]), LOG2E, b) - lse2);
            d[u] = p[u] * (__uint_as_float(x2[e + u]) - (DKV ? dl[e + u] : delta));
        }
        if (DKV) pp[e >> 1] = pack_bf16x2(p[0], p[1]);
        ds[e >> 1] = pack_bf16x2(d[0], d[1]);
    }
}

template <bool DKV>
__global__ void __launch_bounds__(NTHR, 2)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv_box, const __grid_constant__ CUtensorMap tmap_qkv_tile,
                   const __grid_constant__ CUtensorMap tmap_do_box, const __grid_constant__ CUtensorMap tmap_do_tile,
                   const float* __restrict__ table, const float* __restrict__ lse, const float* __restrict__ delta,
                   __nv_bfloat16* __restrict__ dqkv, int nseq, int heads) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sR1 = smem + B_OFF_R1;
    uint8_t* sR2 = smem + B_OFF_R2;
    uint8_t* sT1 = smem + B_OFF_T1;
    uint8_t* sT2 = smem + B_OFF_T2;
    float* sTab = reinterpret_cast<float*>(smem + B_OFF_TAB);
    float* sL = reinterpret_cast<float*>(smem + B_OFF_L);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_OFF_BAR);
    uint64_t* bar_R = bars;            // [3] TMA -> MMA: box c of both resident operands
    uint64_t* bar_T = bars + 3;        // TMA -> MMA (/ softmax): both tile operands
    uint64_t* bar_S = bars + 4;        // MMA -> softmax: X1, X2 of a chunk are in TMEM
    uint64_t* bar_P = bars + 5;        // softmax -> MMA: P / dS of a chunk are in TMEM
    uint64_t* bar_O = bars + 6;        // MMA -> softmax: the tile's accumulators are complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + B_OFF_SLOT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int inner = heads * 32;
    constexpr int ITEM_W = 2 * (NQT - 1) + 1;
    const long long W = (long long)nseq * heads * ITEM_W;
    auto tile_at = [&](long long w) { return (int)((w / ITEM_W) * NQT + (w % ITEM_W + 1) / 2); };
    const int g0 = tile_at(W * blockIdx.x / gridDim.x), g1 = tile_at(W * (blockIdx.x + 1) / gridDim.x);

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&tmap_qkv_box);
            tma_prefetch_desc(&tmap_qkv_tile);
            tma_prefetch_desc(&tmap_do_box);
            tma_prefetch_desc(&tmap_do_tile);
            for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
            mbar_init(bar_S, 1);
            mbar_init(bar_P, NSOFT);
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
            constexpr uint32_t idesc_x = umma_idesc_bf16(QT, SC, 0, 0);
            constexpr uint32_t idesc_acc = umma_idesc_bf16(QT, 32, 0, 1);
            auto load_r = [&](int item) {
                const int s = item % nseq, h = item / nseq;
                for (int c = 0; c < NCH; ++c) {
                    mbar_expect_tx(&bar_R[c], 2 * CH_BYTES + ((DKV && c == 0) ? 2 * TL * 4 : 0));
                    if (DKV && c == 0) {          // lse / delta of the item's 576 queries travel with the first box
                        const long long so = ((long long)s * heads + h) * TL;
                        bulk_load_1d(sL, lse + so, TL * 4, &bar_R[0]);
                        bulk_load_1d(sL + TL, delta + so, TL * 4, &bar_R[0]);
                    }
                    if (DKV) {
                        tma_load_2d(sR1 + c * CH_BYTES, &tmap_qkv_box, &bar_R[c], h * 32, s * TL + c * CH);
                        tma_load_2d(sR2 + c * CH_BYTES, &tmap_do_box, &bar_R[c], h * 32, s * TL + c * CH);
                    } else {
                        tma_load_2d(sR1 + c * CH_BYTES, &tmap_qkv_box, &bar_R[c], inner + h * 32, s * TL + c * CH);
                        tma_load_2d(sR2 + c * CH_BYTES, &tmap_qkv_box, &bar_R[c], 2 * inner + h * 32, s * TL + c * CH);
                    }
                }
            };
            auto load_t = [&](int g) {
                const int item = g / NQT, t = g % NQT;
                const int s = item % nseq, h = item / nseq;
                const uint32_t stat_bytes = DKV ? 0u : (uint32_t)min(QT, TL - t * QT) * 4u;
                mbar_expect_tx(bar_T, 2 * Q_BYTES + 2 * stat_bytes);
                if (!DKV) {                       // lse / delta of the tile's query rows travel with the tile
                    const long long so = ((long long)s * heads + h) * TL + t * QT;
                    bulk_load_1d(sL, lse + so, stat_bytes, bar_T);
                    bulk_load_1d(sL + TL, delta + so, stat_bytes, bar_T);
                }
                if (DKV) {
                    tma_load_2d(sT1, &tmap_qkv_tile, bar_T, inner + h * 32, s * TL + t * QT);
                    tma_load_2d(sT2, &tmap_qkv_tile, bar_T, 2 * inner + h * 32, s * TL + t * QT);
                } else {
                    tma_load_2d(sT1, &tmap_qkv_tile, bar_T, h * 32, s * TL + t * QT);
                    tma_load_2d(sT2, &tmap_do_tile, bar_T, h * 32, s * TL + t * QT);
                }
            };
            const uint32_t sR1_a = smem_u32(sR1), sR2_a = smem_u32(sR2), sT1_a = smem_u32(sT1), sT2_a = smem_u32(sT2);
            const int ntiles = g1 - g0;
            load_r(g0 / NQT);
            load_t(g0);
            uint32_t p_par = 0, o_par = 0;
            for (int n = 0; n < ntiles; ++n) {
                const int g = g0 + n, item = g / NQT;
                const uint32_t r_par = (uint32_t)(item - g0 / NQT) & 1u;
                const bool last_of_item = (n + 1 == ntiles) || ((g + 1) / NQT != item);
                mbar_wait(bar_T, (uint32_t)n & 1u);
#pragma unroll 1
                for (int c = 0; c < NSC; ++c) {
                    mbar_wait(&bar_R[c >> 1], r_par);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        tc_mma_f16(tmem_base + COL_X1, umma_desc(sT1_a + k * 32, 16, 512, SW64),
                                   umma_desc(sR1_a + c * (SC * 64) + k * 32, 16, 512, SW64), idesc_x, k);
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        tc_mma_f16(tmem_base + COL_X2, umma_desc(sT2_a + k * 32, 16, 512, SW64),
                                   umma_desc(sR2_a + c * (SC * 64) + k * 32, 16, 512, SW64), idesc_x, k);
                    tc_commit(bar_S);
                    mbar_wait(bar_P, p_par);
                    p_par ^= 1;
                    tc_fence_after();
                    // X MMAs of this chunk have retired; after the tile's last chunk the tile operands are free
                    if (c == NSC - 1 && n + 1 < ntiles && !last_of_item) load_t(g + 1);
#pragma unroll
                    for (int k = 0; k < SC / 16; ++k) {
                        const uint32_t a_col = (k < 3 ? k * 8 : 48 + (k - 3) * 8);
                        tc_mma_f16_ts(tmem_base + COL_A0, tmem_base + COL_X2 + a_col,
                                      umma_desc(sR1_a + c * (SC * 64) + k * 1024, 512, 512, SW64), idesc_acc,
                                      (c > 0 || k > 0) ? 1u : 0u);
                        if (DKV)
                            tc_mma_f16_ts(tmem_base + COL_A1, tmem_base + COL_X1 + a_col,
                                          umma_desc(sR2_a + c * (SC * 64) + k * 1024, 512, 512, SW64), idesc_acc,
                                          (c > 0 || k > 0) ? 1u : 0u);
                    }
                }
                tc_commit(bar_O);
                if (last_of_item && n + 1 < ntiles) {
                    mbar_wait(bar_O, o_par);        // every MMA of the item has retired: refill the resident operands
                    load_r(item + 1);
                    load_t(g + 1);
                }
                o_par ^= 1;
            }
        }
    } else {
        // ================================ softmax warps ================================
        const int sw = warp - 1;
        const int quarter = warp & 3;
        const int hsel = sw >> 2;
        const int row = quarter * 32 + lane;
        const int st = threadIdx.x - 32;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t sTab_a = smem_u32(sTab), sL_a = smem_u32(sL);
        uint32_t s_par = 0, o_par = 0;
        int cur_item = -1, cur_head = -1, n = 0;
        for (int g = g0; g < g1; ++g, ++n) {
            const int item = g / NQT, t = g % NQT;
            const int s = item % nseq, h = item / nseq;
            if (item != cur_item) {
                cur_item = item;
                if (h != cur_head) {                             // rare: items are head-major
                    soft_sync();                                 // everyone is done with the previous head's table
                    cur_head = h;
                    const float* tg = table + (long long)h * TNOFF;
                    for (int i = st; i < TNOFF; i += 32 * NSOFT) sTab[(i / TWW) * TPW + i % TWW] = __ldg(tg + i) * LOG2E;
                    soft_sync();
                }
                // DKV: the item's lse / delta arrive with the first resident box (bulk copies, no global loads here)
                if (DKV) mbar_wait(&bar_R[0], (uint32_t)(item - g0 / NQT) & 1u);
            }
            const int i = t * QT + row;                            // query (dQ) or key (dK/dV) of this thread
            const bool active = t * QT + quarter * 32 < TL;        // warp-uniform
            float lse2 = 0.f, dlt = 0.f;
            uint32_t bias_row = sTab_a;
            if (active) {
                const int pos = (i / TGW) * TPW + (i % TGW);
                if (DKV) {
                    bias_row = sTab_a + 4u * (uint32_t)(TOFF - pos);
                } else {
                    bias_row = sTab_a + 4u * (uint32_t)(TOFF + pos);
                    mbar_wait(bar_T, (uint32_t)n & 1u);          // the tile's statistics came with its operands
                    lse2 = sL[row] * LOG2E;
                    dlt = sL[TL + row];
                }
            }
#pragma unroll 1
            for (int c = 0; c < NSC; ++c) {
                mbar_wait(bar_S, s_par);
                s_par ^= 1;
                tc_fence_after();
                if (active) {
                    const int cb = (c * 4 + hsel * 2) * TPW;        // first grid row of the warp's 48 columns
                    const uint32_t bb = DKV ? bias_row + 4u * (uint32_t)cb : bias_row - 4u * (uint32_t)cb;
                    const uint32_t stat = sL_a + 4u * (uint32_t)(c * SC + hsel * 48);
                    const uint32_t t_x1 = t_lane + COL_X1 + hsel * 48, t_x2 = t_lane + COL_X2 + hsel * 48;
                    uint32_t x1[16], x2[16], pp[8], ds[8];
                    tc_ld_32x32_x16(t_x1, x1);
                    tc_ld_32x32_x16(t_x2, x2);
                    tc_wait_ld();
                    bwd_block<DKV, 0>(x1, x2, pp, ds, bb, stat, lse2, dlt);
                    tc_ld_32x32_x16(t_x1 + 16, x1);
                    tc_ld_32x32_x16(t_x2 + 16, x2);
                    tc_wait_ld();
                    tc_st_32x32_x8(t_x2, ds);
                    if (DKV) tc_st_32x32_x8(t_x1, pp);
                    bwd_block<DKV, 16>(x1, x2, pp, ds, bb, stat, lse2, dlt);
                    tc_ld_32x32_x16(t_x1 + 32, x1);
                    tc_ld_32x32_x16(t_x2 + 32, x2);
                    tc_wait_ld();
                    tc_st_32x32_x8(t_x2 + 8, ds);
                    if (DKV) tc_st_32x32_x8(t_x1 + 8, pp);
                    bwd_block<DKV, 32>(x1, x2, pp, ds, bb, stat, lse2, dlt);
                    tc_st_32x32_x8(t_x2 + 16, ds);
                    if (DKV) tc_st_32x32_x8(t_x1 + 16, pp);
                    tc_wait_st();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_P);
            }
            mbar_wait(bar_O, o_par);
            o_par ^= 1;
            tc_fence_after();
            if (active) {
                __nv_bfloat16* orow = dqkv + ((long long)s * TL + i) * (3LL * inner) + h * 32 + hsel * 16;
#pragma unroll
                for (int a = 0; a < (DKV ? 2 : 1); ++a) {
                    uint32_t o[16];
                    tc_ld_32x32_x16(t_lane + (a == 0 ? COL_A0 : COL_A1) + hsel * 16, o);
                    tc_wait_ld();
                    uint4 w0, w1;
                    w0.x = pack_bf16x2(__uint_as_float(o[0]), __uint_as_float(o[1]));
                    w0.y = pack_bf16x2(__uint_as_float(o[2]), __uint_as_float(o[3]));
                    w0.z = pack_bf16x2(__uint_as_float(o[4]), __uint_as_float(o[5]));
                    w0.w = pack_bf16x2(__uint_as_float(o[6]), __uint_as_float(o[7]));
                    w1.x = pack_bf16x2(__uint_as_float(o[8]), __uint_as_float(o[9]));
                    w1.y = pack_bf16x2(__uint_as_float(o[10]), __uint_as_float(o[11]));
                    w1.z = pack_bf16x2(__uint_as_float(o[12]), __uint_as_float(o[13]));
                    w1.w = pack_bf16x2(__uint_as_float(o[14]), __uint_as_float(o[15]));
                    // dQ -> q slot; dK (acc0) -> k slot, dV (acc1) -> v slot of the packed gradient
                    uint4* op = reinterpret_cast<uint4*>(orow + (DKV ? (a + 1) * inner : 0));
                    op[0] = w0;
                    op[1] = w1;
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Bias-table gradient: dtable[h][rel(i, j)] += sum over slices of dS[s, h, i, j].
// "Stationary" formulation: a CTA owns one (head, 128-query tile, 64-key chunk) block and walks a
// range of slices; the block's dS (bf16, as for dK) is summed over the slices by the tensor core
// itself - G += dS * I_64 (tcgen05.mma TS against an identity tile in shared memory, fp32
// accumulation in TMEM) - so the softmax warps spend no instruction on the reduction.  After the
// last slice G is scattered once into a shared-memory copy of the table and flushed with one
// atomic per touched entry.  Operands (and the rows' lse / delta) of the next two slices are in flight in
// a 3-stage TMA ring.  Two CTAs per SM; TMEM columns: X1 = Q K^T [0,64)  X2 = dO V^T [64,128)  G [128,192).
// ---------------------------------------------------------------------------------------------
constexpr int DC = 64;                      // keys per block
constexpr int NDC = TL / DC;                // 9
constexpr int D_STAGES = 3;
constexpr int D_NSOFT = 8;                 // softmax warps: 2 per TMEM lane quarter, 32 keys of the chunk each
constexpr int D_NTHR = 32 * (1 + D_NSOFT);
constexpr int D_STAT_OFF = 2 * Q_BYTES + 2 * DC * 64;          // lse[128] | delta[128] of the tile's rows for this slice
constexpr int D_STAGE_BYTES = D_STAT_OFF + 2 * QT * 4;
constexpr int D_OFF_ID = D_STAGES * D_STAGE_BYTES;
constexpr int D_OFF_TAB = D_OFF_ID + 2 * DC * 64;
constexpr int D_OFF_ACC = D_OFF_TAB + TWW * TPW * 4;
constexpr int D_OFF_BAR = D_OFF_ACC + TWW * TPW * 4;
constexpr int D_NBAR = 2 * D_STAGES + 3;
constexpr int D_OFF_SLOT = D_OFF_BAR + D_NBAR * 8;
constexpr int D_SMEM_BYTES = D_OFF_SLOT + 16 + 1024;
constexpr uint32_t D_COL_X1 = 0, D_COL_X2 = 64, D_COL_G = 128, D_TMEM_COLS = 256;

// 16 columns (keys base + J0 .. base + J0 + 15, base = PHASE mod 24 inside its grid row) -> bf16 pairs of dS
__device__ __forceinline__ void dsoft_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * D_NSOFT) : "memory"); }

template <int PHASE, int J0>
__device__ __forceinline__ void dbias_block(const uint32_t (&x1)[16], const uint32_t (&x2)[16], uint32_t (&ds)[8],
                                            uint32_t bias_addr, float lse2, float delta) {
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
        float d[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = PHASE + J0 + e + u;
            const float b = lds_f32(bias_addr - 4u * (uint32_t)((j / TGW) * TPW + j % TGW));
            const float p = fast_exp2(fmaf(__uint_as_float(x1[e + u]), LOG2E, b) - lse2);
            d[u] = p * (__uint_as_float(x2[e + u]) - delta);
        }
        ds[e >> 1] = pack_bf16x2(d[0], d[1]);
    }
}
template <int PHASE>
__device__ __forceinline__ void dbias_chunk(uint32_t t_x1, uint32_t t_x2, uint32_t bias_addr, float lse2, float delta) {
    uint32_t x1[16], x2[16], ds[8];
    tc_ld_32x32_x16(t_x1, x1);
    tc_ld_32x32_x16(t_x2, x2);
    tc_wait_ld();
    dbias_block<PHASE, 0>(x1, x2, ds, bias_addr, lse2, delta);
    tc_ld_32x32_x16(t_x1 + 16, x1);
    tc_ld_32x32_x16(t_x2 + 16, x2);
    tc_wait_ld();
    tc_st_32x32_x8(t_x2, ds);
    dbias_block<PHASE, 16>(x1, x2, ds, bias_addr, lse2, delta);
    tc_st_32x32_x8(t_x2 + 8, ds);
    tc_wait_st();
}

__global__ void __launch_bounds__(D_NTHR, 2)
attn_dbias_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv_tile, const __grid_constant__ CUtensorMap tmap_do_tile,
                     const __grid_constant__ CUtensorMap tmap_qkv_chunk, const float* __restrict__ table,
                     const float* __restrict__ lse, const float* __restrict__ delta, float* __restrict__ dtable, int nseq,
                     int heads, int nsplit) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sId = smem + D_OFF_ID;
    float* sTab = reinterpret_cast<float*>(smem + D_OFF_TAB);
    float* sAcc = reinterpret_cast<float*>(smem + D_OFF_ACC);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + D_OFF_BAR);
    uint64_t* full = bars;                       // [3] TMA -> MMA
    uint64_t* empty = bars + D_STAGES;           // [3] MMA -> TMA (X products of the stage have retired)
    uint64_t* bar_S = bars + 2 * D_STAGES;       // MMA -> softmax: the logit products of a slice are in TMEM
    uint64_t* bar_P = bar_S + 1;                 // softmax -> MMA: dS of the slice is in TMEM
    uint64_t* bar_O = bar_S + 2;                 // MMA -> softmax: G of the item is complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + D_OFF_SLOT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int inner = heads * 32;
    // item = (split of the slices, head, query tile, key chunk); items of one CTA: blockIdx.x, + gridDim.x, ...
    const int items = nsplit * heads * NQT * NDC;
    auto decode = [&](int item, int& sp, int& h, int& t, int& c) {
        c = item % NDC; item /= NDC;
        t = item % NQT; item /= NQT;
        h = item % heads; sp = item / heads;
    };
    auto s_begin = [&](int sp) { return (int)((long long)nseq * sp / nsplit); };

    // identity tile (B operand of the accumulation): [64 n][64 k] bf16 as two 32-wide K blocks, 64 B rows, SW64
    for (int i = threadIdx.x; i < 2 * DC * 64 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sId)[i] = 0u;
    __syncthreads();
    if (threadIdx.x < DC) {
        const int n = threadIdx.x;
        const int chunk = ((n % 32) / 8) ^ ((n >> 1) & 3);
        *reinterpret_cast<__nv_bfloat16*>(sId + (n / 32) * (DC * 64) + n * 64 + chunk * 16 + (n % 8) * 2) = __float2bfloat16(1.0f);
    }
    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&tmap_qkv_tile);
            tma_prefetch_desc(&tmap_do_tile);
            tma_prefetch_desc(&tmap_qkv_chunk);
            for (int i = 0; i < 2 * D_STAGES; ++i) mbar_init(&bars[i], 1);
            mbar_init(bar_S, 1);
            mbar_init(bar_P, D_NSOFT);
            mbar_init(bar_O, 1);
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, D_TMEM_COLS);
        tmem_relinquish();
    }
    // generic-proxy writes of the identity tile must be visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ control: TMA + MMA issue ================================
        if (lane == 0 && (int)blockIdx.x < items) {
            constexpr uint32_t idesc_x = umma_idesc_bf16(QT, DC, 0, 0);
            constexpr uint32_t idesc_g = umma_idesc_bf16(QT, DC, 0, 0);
            const uint32_t ring_a = smem_u32(smem), sId_a = smem_u32(sId);
            // flattened (item, slice) stream for the loads
            int l_item = blockIdx.x, l_sp, l_h, l_t, l_c, l_s, l_send, l_n = 0;
            decode(l_item, l_sp, l_h, l_t, l_c);
            l_s = s_begin(l_sp); l_send = s_begin(l_sp + 1);
            auto load_next = [&]() {                              // returns false when the stream is exhausted
                while (l_s >= l_send) {
                    l_item += gridDim.x;
                    if (l_item >= items) return false;
                    decode(l_item, l_sp, l_h, l_t, l_c);
                    l_s = s_begin(l_sp); l_send = s_begin(l_sp + 1);
                }
                const int st = l_n % D_STAGES;
                if (l_n >= D_STAGES) mbar_wait(&empty[st], ((uint32_t)(l_n / D_STAGES) & 1u) ^ 1u);
                uint8_t* sb = smem + st * D_STAGE_BYTES;
                const uint32_t stat_bytes = (uint32_t)min(QT, TL - l_t * QT) * 4u;       // the last tile holds 64 rows
                mbar_expect_tx(&full[st], D_STAT_OFF + 2 * stat_bytes);
                const long long so = ((long long)l_s * heads + l_h) * TL + l_t * QT;
                bulk_load_1d(sb + D_STAT_OFF, lse + so, stat_bytes, &full[st]);
                bulk_load_1d(sb + D_STAT_OFF + QT * 4, delta + so, stat_bytes, &full[st]);
                tma_load_2d(sb, &tmap_qkv_tile, &full[st], l_h * 32, l_s * TL + l_t * QT);
                tma_load_2d(sb + Q_BYTES, &tmap_do_tile, &full[st], l_h * 32, l_s * TL + l_t * QT);
                tma_load_2d(sb + 2 * Q_BYTES, &tmap_qkv_chunk, &full[st], inner + l_h * 32, l_s * TL + l_c * DC);
                tma_load_2d(sb + 2 * Q_BYTES + DC * 64, &tmap_qkv_chunk, &full[st], 2 * inner + l_h * 32, l_s * TL + l_c * DC);
                ++l_s; ++l_n;
                return true;
            };
            for (int i = 0; i < D_STAGES - 1; ++i) load_next();
            uint32_t p_par = 0;
            int n = 0;                                            // consumed (item, slice) pairs
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                int sp, h, t, c;
                decode(item, sp, h, t, c);
                const int s0 = s_begin(sp), s1 = s_begin(sp + 1);
                for (int s = s0; s < s1; ++s, ++n) {
                    load_next();                                  // D_STAGES - 1 pairs ahead
                    const int st = n % D_STAGES;
                    mbar_wait(&full[st], (uint32_t)(n / D_STAGES) & 1u);
                    tc_fence_after();
                    const uint32_t sb = ring_a + st * D_STAGE_BYTES;
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        tc_mma_f16(tmem_base + D_COL_X1, umma_desc(sb + k * 32, 16, 512, SW64),
                                   umma_desc(sb + 2 * Q_BYTES + k * 32, 16, 512, SW64), idesc_x, k);
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        tc_mma_f16(tmem_base + D_COL_X2, umma_desc(sb + Q_BYTES + k * 32, 16, 512, SW64),
                                   umma_desc(sb + 2 * Q_BYTES + DC * 64 + k * 32, 16, 512, SW64), idesc_x, k);
                    tc_commit(bar_S);
                    tc_commit(&empty[st]);
                    mbar_wait(bar_P, p_par);
                    p_par ^= 1;
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < DC / 16; ++k)               // G += dS * I (keys of warp half k/2 at columns 32 (k/2))
                        tc_mma_f16_ts(tmem_base + D_COL_G, tmem_base + D_COL_X2 + (k >> 1) * 32 + (k & 1) * 8,
                                      umma_desc(sId_a + (k >> 1) * (DC * 64) + (k & 1) * 32, 16, 512, SW64), idesc_g,
                                      (s > s0 || k > 0) ? 1u : 0u);
                }
                tc_commit(bar_O);
            }
        }
    } else {
        // ================================ softmax warps ================================
        const int sw = warp - 1;
        const int quarter = warp & 3;
        const int hsel = sw >> 2;                                  // which 32 keys of the chunk
        const int row = quarter * 32 + lane;
        const int st_id = threadIdx.x - 32;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t sTab_a = smem_u32(sTab), sAcc_a = smem_u32(sAcc);
        uint32_t s_par = 0, o_par = 0;
        int cur_head = -1, n = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            int sp, h, t, c;
            decode(item, sp, h, t, c);
            const int s0 = s_begin(sp), s1 = s_begin(sp + 1);
            dsoft_sync();                                          // previous item's flush is done
            if (h != cur_head) {
                cur_head = h;
                const float* tg = table + (long long)h * TNOFF;
                for (int i = st_id; i < TNOFF; i += 32 * D_NSOFT) sTab[(i / TWW) * TPW + i % TWW] = __ldg(tg + i) * LOG2E;
            }
            for (int i = st_id; i < TWW * TPW; i += 32 * D_NSOFT) sAcc[i] = 0.f;
            dsoft_sync();
            const int i = t * QT + row;
            const bool active = t * QT + quarter * 32 < TL;
            const int kbase = c * DC + hsel * 32;                  // first key of this warp's 32
            const int phase = kbase % TGW;                         // 0, 8 or 16
            // address of bias(i, key at the start of kbase's grid row)
            const uint32_t rel = (uint32_t)(TOFF + (i / TGW) * TPW + (i % TGW) - (kbase / TGW) * TPW);
            const uint32_t bias_addr = sTab_a + 4u * rel;
            for (int s = s0; s < s1; ++s, ++n) {
                const uint32_t t_x1 = t_lane + D_COL_X1 + hsel * 32, t_x2 = t_lane + D_COL_X2 + hsel * 32;
                mbar_wait(bar_S, s_par);
                s_par ^= 1;
                tc_fence_after();
                if (active) {
                    // row statistics of this slice came with the stage (acquire its bulk copies through the stage's own
                    // barrier; it completed before the X products that signalled bar_S were even issued)
                    mbar_wait(&full[n % D_STAGES], (uint32_t)(n / D_STAGES) & 1u);
                    const float* stats = reinterpret_cast<const float*>(smem + (n % D_STAGES) * D_STAGE_BYTES + D_STAT_OFF);
                    const float lse2 = stats[row] * LOG2E, dlt = stats[QT + row];
                    if (phase == 0) dbias_chunk<0>(t_x1, t_x2, bias_addr, lse2, dlt);
                    else if (phase == 8) dbias_chunk<8>(t_x1, t_x2, bias_addr, lse2, dlt);
                    else dbias_chunk<16>(t_x1, t_x2, bias_addr, lse2, dlt);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_P);
            }
            mbar_wait(bar_O, o_par);
            o_par ^= 1;
            tc_fence_after();
            if (active) {
                uint32_t g[32];
                tc_ld_32x32(t_lane + D_COL_G + hsel * 32, g);
                tc_wait_ld();
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int j = phase + e;
                    const uint32_t a = sAcc_a + 4u * (rel - (uint32_t)((j / TGW) * TPW + j % TGW));
                    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(__uint_as_float(g[e])) : "memory");
                }
            }
            tc_fence_before();
            dsoft_sync();
            float* dt = dtable + (long long)h * TNOFF;
            for (int k = st_id; k < TNOFF; k += 32 * D_NSOFT) {
                const float v = sAcc[(k / TWW) * TPW + k % TWW];
                if (v != 0.f) atomicAdd(dt + k, v);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, D_TMEM_COLS);
}

template <bool DKV>
int launch_bwd(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& td, const float* table,
               const float* lse, const float* delta, __nv_bfloat16* dqkv, int nseq, int heads, cudaStream_t stream) {
    auto kern = attn_bwd_tc_kernel<DKV>;
    CTK_SET_MAX_SMEM(kern, B_SMEM_BYTES);
    const long long G = (long long)nseq * heads * NQT;
    long long grid = 2LL * ctk_num_sms();
    if (grid > G) grid = G;
    kern<<<(unsigned)grid, NTHR, B_SMEM_BYTES, stream>>>(ta, tb, tc, td, table, lse, delta, dqkv, nseq, heads);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

}  // namespace

// dq, dk, dv of the 24x24 spatial stack (delta = rowsum(dO o O) already computed); validated by ctk_attn_bwd.
int ctk_attn_bwd_tc(const void* qkv, const float* table, const void* dout, const float* lse, const float* delta,
                    void* dqkv, float* dtable, int dtable_tc, int nseq, int heads, cudaStream_t stream) {
    const int inner = heads * 32;
    const unsigned long long rows = (unsigned long long)nseq * TL;
    CUtensorMap ta, tb, tc, td;
    const unsigned long long dims_qkv[2] = {(unsigned long long)(3 * inner), rows};
    const unsigned long long str_qkv[1] = {(unsigned long long)(3 * inner) * 2};
    const unsigned long long dims_do[2] = {(unsigned long long)inner, rows};
    const unsigned long long str_do[1] = {(unsigned long long)inner * 2};
    const unsigned int box_tile[2] = {32, QT}, box_box[2] = {32, CH};
    int rc;
    if ((rc = ctk_make_tmap(&ta, qkv, false, 2, dims_qkv, str_qkv, box_box, 2))) return rc;
    if ((rc = ctk_make_tmap(&tb, qkv, false, 2, dims_qkv, str_qkv, box_tile, 2))) return rc;
    if ((rc = ctk_make_tmap(&tc, dout, false, 2, dims_do, str_do, box_box, 2))) return rc;
    if ((rc = ctk_make_tmap(&td, dout, false, 2, dims_do, str_do, box_tile, 2))) return rc;
    auto g = reinterpret_cast<__nv_bfloat16*>(dqkv);
    if ((rc = launch_bwd<false>(ta, tb, tc, td, table, lse, delta, g, nseq, heads, stream))) return rc;
    if ((rc = launch_bwd<true>(ta, tb, tc, td, table, lse, delta, g, nseq, heads, stream))) return rc;
    // bias-table gradient: tcgen05 formulation below (0.50 ms at B=8) unless dtable_tc == 0, in which case the caller
    // runs the mma.sync kernel of attention.cu (fp32 dS, 0.62 ms)
    if (!dtable_tc) return CTK_OK;
    CUtensorMap te;
    const unsigned int box_chunk[2] = {32, DC};
    if ((rc = ctk_make_tmap(&te, qkv, false, 2, dims_qkv, str_qkv, box_chunk, 2))) return rc;
    CTK_SET_MAX_SMEM(attn_dbias_tc_kernel, D_SMEM_BYTES);
    // two CTAs per SM (256 TMEM columns each): one CTA's control / MMA round trip is covered by the other's
    // softmax; split the slices so that there are ~5 items per CTA
    const int blocks = heads * NQT * NDC;
    int nsplit = (10 * ctk_num_sms() + blocks - 1) / blocks;
    if (nsplit > nseq) nsplit = nseq;
    if (nsplit < 1) nsplit = 1;
    long long grid = 2LL * ctk_num_sms();
    if (grid > (long long)blocks * nsplit) grid = (long long)blocks * nsplit;
    attn_dbias_tc_kernel<<<(unsigned)grid, D_NTHR, D_SMEM_BYTES, stream>>>(tb, td, te, table, lse, delta, dtable, nseq, heads, nsplit);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

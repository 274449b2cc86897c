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
// Per tile and 192-key chunk c:
//   control thread : S = Q K_c^T        tcgen05.mma SS, M128 N192 K32 (two k-steps), D = TMEM[0,192)
//   8 softmax warps: p = exp2(s*log2e + bias - M_i), written back as bf16 pairs over the columns the
//                    logits were read from (tcgen05.ld / tcgen05.st; warp w owns TMEM lanes 32*(w%4),
//                    warps 1-4 the first 96 keys of the chunk, warps 5-8 the last 96)
//   control thread : O += P V_c         tcgen05.mma TS (A = P in TMEM), M128 N32 K192, D = TMEM[192,224)
// M_i = |q_i| max_j|k_j| + max(bias) bounds every logit of the row (Cauchy-Schwarz), so there is no
// running-max pass and no rescaling of O; rows whose bound could be more than 2^64 above the true
// maximum (never at the scales CT-CLIP trains with) take an exact SIMT maximum instead.
// The bias index of (i, j) is pos(i) - pos(j) + const with pos(n) = (n/24)*47 + n%24; with 96 keys =
// 4 grid rows per warp pass, pos(j) is a compile-time immediate of the unrolled loop.
#include "common.cuh"

namespace {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

constexpr int TL = 576;                 // tokens per sequence
constexpr int TGW = 24;                 // token grid width (and height)
constexpr int TWW = 2 * TGW - 1;        // 47 distinct offsets per axis
constexpr int TNOFF = TWW * TWW;        // 2209 table entries per head
constexpr int TOFF = (TGW - 1) * TWW + (TGW - 1);
constexpr int QT = 128;                 // queries per tile
constexpr int NQT = (TL + QT - 1) / QT; // 5 (the last one holds 64 queries)
constexpr int CH = 192;                 // keys per chunk
constexpr int NCH = TL / CH;            // 3
constexpr int NSOFT = 8;                // softmax warps
constexpr int NTHR = 32 * (1 + NSOFT);

constexpr int CH_BYTES = CH * 64;       // 12288
constexpr int Q_BYTES = QT * 64;        // 8192
constexpr int OFF_K = 0;
constexpr int OFF_V = OFF_K + NCH * CH_BYTES;
constexpr int OFF_Q = OFF_V + NCH * CH_BYTES;
constexpr int OFF_T = OFF_Q + 2 * Q_BYTES;
constexpr int OFF_LP = OFF_T + ((TNOFF * 4 + 15) & ~15);
constexpr int OFF_RED = OFF_LP + 2 * QT * 4;
constexpr int OFF_BAR = OFF_RED + 3 * NSOFT * 4;
constexpr int NBAR = 3 + 3 + 2 + 3;
constexpr int OFF_SLOT = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_SLOT + 16 + 1024;      // + alignment slack

constexpr uint32_t COL_S = 0, COL_O = 192, TMEM_COLS = 256;

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout_type) << 61;
    return d;
}
constexpr uint32_t SW64 = 4;            // cute::UMMA::LayoutType::SWIZZLE_64B

// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t taddr_d, uint32_t taddr_a, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(taddr_d),
        "r"(taddr_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld_32x32_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_st_32x32_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void soft_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * NSOFT) : "memory"); }

__device__ __forceinline__ float sumsq16(const uint4 u, float acc) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(w[e]);
        acc = fmaf(f.x, f.x, acc);
        acc = fmaf(f.y, f.y, acc);
    }
    return acc;
}

// exact max_j (q_i.k_j log2e + bias') by SIMT dot products out of the swizzled tiles (rare path)
__device__ __noinline__ float exact_row_max(const uint8_t* sQrow, int qrow, const uint8_t* sK, const float* bias_base) {
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
        m = fmaxf(m, fmaf(dot, LOG2E, bias_base[-((j / TGW) * TWW + j % TGW)]));
    }
    return m;
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
    uint64_t* bar_S = bars + 8;        // MMA -> softmax: logits of a chunk are in TMEM
    uint64_t* bar_P = bars + 9;        // softmax -> MMA: probabilities of a chunk are in TMEM
    uint64_t* bar_O = bars + 10;       // MMA -> softmax: the tile's O is complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_SLOT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int inner = heads * 32;
    const long long G = (long long)nseq * heads * NQT;
    const int g0 = (int)(G * blockIdx.x / gridDim.x), g1 = (int)(G * (blockIdx.x + 1) / gridDim.x);

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&tmap_q);
            tma_prefetch_desc(&tmap_kv);
            for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
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
            constexpr uint32_t idesc_s = umma_idesc_bf16(QT, CH, 0, 0);
            constexpr uint32_t idesc_pv = umma_idesc_bf16(QT, 32, 0, 1);
            auto load_k = [&](int item) {
                const int s = item / heads, h = item % heads;
                for (int c = 0; c < NCH; ++c) {
                    mbar_expect_tx(&bar_K[c], CH_BYTES);
                    tma_load_2d(sK + c * CH_BYTES, &tmap_kv, &bar_K[c], inner + h * 32, s * TL + c * CH);
                }
            };
            auto load_v = [&](int item) {
                const int s = item / heads, h = item % heads;
                for (int c = 0; c < NCH; ++c) {
                    mbar_expect_tx(&bar_V[c], CH_BYTES);
                    tma_load_2d(sV + c * CH_BYTES, &tmap_kv, &bar_V[c], 2 * inner + h * 32, s * TL + c * CH);
                }
            };
            auto load_q = [&](int g, int buf) {
                const int item = g / NQT, t = g % NQT;
                const int s = item / heads, h = item % heads;
                mbar_expect_tx(&bar_Q[buf], Q_BYTES);
                tma_load_2d(sQ + buf * Q_BYTES, &tmap_q, &bar_Q[buf], h * 32, s * TL + t * QT);
            };
            load_k(g0 / NQT);
            load_q(g0, 0);
            load_v(g0 / NQT);
            if (g0 + 1 < g1) load_q(g0 + 1, 1);
            uint32_t kv_par = 0, p_par = 0, o_par = 0, q_par[2] = {0, 0};
            int cur_item = g0 / NQT, n = 0;
            bool first = true;
            const uint32_t sK_a = smem_u32(sK), sV_a = smem_u32(sV), sQ_a = smem_u32(sQ);
            for (int g = g0; g < g1; ++g, ++n) {
                const int item = g / NQT, buf = n & 1;
                if (item != cur_item) { cur_item = item; kv_par ^= 1; first = true; }
                const bool last_of_item = (g + 1 == g1) || ((g + 1) / NQT != item);
                mbar_wait(&bar_Q[buf], q_par[buf]);
                q_par[buf] ^= 1;
                for (int c = 0; c < NCH; ++c) {
                    if (first) mbar_wait(&bar_K[c], kv_par);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        tc_mma_f16(tmem_base + COL_S, umma_desc(sQ_a + buf * Q_BYTES + k * 32, 16, 512, SW64),
                                   umma_desc(sK_a + c * CH_BYTES + k * 32, 16, 512, SW64), idesc_s, k);
                    tc_commit(bar_S);
                    mbar_wait(bar_P, p_par);
                    p_par ^= 1;
                    tc_fence_after();
                    if (c == NCH - 1) {
                        // every S MMA of this tile has retired: its Q buffer (and, after the last tile of
                        // the item, the K chunks) can be refilled
                        if (g + 2 < g1) load_q(g + 2, buf);
                        if (last_of_item && g + 1 < g1) load_k(item + 1);
                    }
                    if (first) mbar_wait(&bar_V[c], kv_par);
#pragma unroll
                    for (int k = 0; k < CH / 16; ++k) {
                        // P: keys [0,96) of the chunk in columns [0,48), keys [96,192) in columns [96,144)
                        const uint32_t a_col = COL_S + (k < 6 ? k * 8 : 96 + (k - 6) * 8);
                        tc_mma_f16_ts(tmem_base + COL_O, tmem_base + a_col,
                                      umma_desc(sV_a + c * CH_BYTES + k * 1024, 512, 512, SW64), idesc_pv,
                                      (c > 0 || k > 0) ? 1u : 0u);
                    }
                    if (c == NCH - 1) tc_commit(bar_O);
                }
                first = false;
                if (last_of_item && g + 1 < g1) {
                    mbar_wait(bar_O, o_par);        // the item's last PV has retired: V can be refilled
                    load_v(item + 1);
                }
                o_par ^= 1;
            }
        }
    } else {
        // ================================ softmax warps ================================
        const int sw = warp - 1;
        const int quarter = warp & 3;               // TMEM lane quarter this warp may access
        const int hsel = sw >> 2;                   // which 96 keys of a chunk
        const int row = quarter * 32 + lane;
        const int st = threadIdx.x - 32;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        uint32_t kv_par = 1, s_par = 0, o_par = 0, q_par[2] = {0, 0};
        int cur_item = -1, n = 0;
        float kmax = 0.f, bmax = 0.f, bmin = 0.f;
        for (int g = g0; g < g1; ++g, ++n) {
            const int item = g / NQT, t = g % NQT, buf = n & 1;
            const int s = item / heads, h = item % heads;
            if (item != cur_item) {
                cur_item = item;
                kv_par ^= 1;
                soft_sync();                         // everyone is done with the previous head's table
                float mx = -INFINITY, mn = INFINITY;
                const float* tg = table + (long long)h * TNOFF;
                for (int i = st; i < TNOFF; i += 32 * NSOFT) {
                    const float v = __ldg(tg + i) * LOG2E;
                    sT[i] = v;
                    mx = fmaxf(mx, v);
                    mn = fminf(mn, v);
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
            mbar_wait(&bar_Q[buf], q_par[buf]);
            q_par[buf] ^= 1;
            const int i = t * QT + row;
            const bool active = t * QT + quarter * 32 < TL;          // warp-uniform
            float Mi = 0.f, l = 0.f;
            const float* bias_row = sT;
            if (active) {
                const uint8_t* qrow = sQ + buf * Q_BYTES + row * 64;
                float q2 = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) q2 = sumsq16(reinterpret_cast<const uint4*>(qrow)[c], q2);
                const float qk = sqrtf(q2) * kmax * LOG2E;
                Mi = qk + bmax;
                bias_row = sT + (i / TGW) * TWW + (i % TGW) + TOFF;
                if (__any_sync(0xffffffffu, 2.f * qk + (bmax - bmin) > 64.f)) Mi = exact_row_max(qrow, row, sK, bias_row);
            }
            const float negM = -Mi;
            for (int c = 0; c < NCH; ++c) {
                mbar_wait(bar_S, s_par);
                s_par ^= 1;
                tc_fence_after();
                if (active) {
                    const float* bb = bias_row - (c * 8 + hsel * 4) * TWW;
                    const uint32_t t_s = t_lane + COL_S + hsel * 96;
#pragma unroll
                    for (int blk = 0; blk < 3; ++blk) {
                        uint32_t r[32];
                        tc_ld_32x32(t_s + blk * 32, r);
                        tc_wait_ld();
                        uint32_t pk[16];
#pragma unroll
                        for (int e = 0; e < 32; e += 2) {
                            const int j0 = blk * 32 + e, j1 = j0 + 1;
                            const float p0 = fast_exp2(fmaf(__uint_as_float(r[e]), LOG2E, negM) + bb[-((j0 / TGW) * TWW + j0 % TGW)]);
                            const float p1 = fast_exp2(fmaf(__uint_as_float(r[e + 1]), LOG2E, negM) + bb[-((j1 / TGW) * TWW + j1 % TGW)]);
                            l += p0 + p1;
                            pk[e >> 1] = pack_bf16x2(p0, p1);
                        }
                        tc_st_32x32_x16(t_s + blk * 16, pk);
                    }
                    tc_wait_st();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_P);
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
    static bool configured = false;
    if (!configured) {
        CTK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const long long G = (long long)nseq * heads * NQT;
    long long grid = 2LL * ctk_num_sms();
    if (grid > G) grid = G;
    attn_fwd_tc_kernel<<<(unsigned)grid, NTHR, SMEM_BYTES, stream>>>(tq, tkv, table, reinterpret_cast<__nv_bfloat16*>(out), lse,
                                                                     nseq, heads);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

// Cosine attention core for the temporal stack (attention.py:162-184): sequences of 24 tokens, no
// position bias, head dim 32.  The kernels are HBM-bound (a sequence is a contiguous 36 KB block of
// the packed qkv buffer and needs 0.6 MFLOP), so the design is about streaming:
//   * persistent CTAs; a producer warp keeps a 3-stage ring of sequences in flight with TMA
//     (one {32 x 24} 64B-swizzled box per (q|k|v, head): conflict-free ldmatrix, no repacking);
//   * one consumer warp per head works entirely in registers: bf16 mma.sync (m16n8k16 / m16n8k8,
//     24 = 16 + 8 along the contraction, so no padding is ever multiplied), fp32 softmax;
//   * no cross-warp synchronisation besides the ring's full / empty mbarriers.
// A 24 x 24 problem would leave 81 % of a 128-row tcgen05 tile empty, and its latency chain
// (TMEM round trips per sequence) cannot be amortised; warp-level MMA is the right grain here.
// Backward recomputes P = exp(S - lse) in both orientations (rows = queries for dQ, rows = keys
// for dK / dV) so that every product takes its left operand straight from accumulator registers.
#include "attention_mma.cuh"

using namespace attn_mma;

namespace {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int SL = 24;                    // tokens per sequence
constexpr int TILE = SL * 64;             // bytes of one (part, head) tile
constexpr int STAGES = 3;

__device__ __forceinline__ void mul_bf16x2_sum(uint32_t a, uint32_t b, float& acc) {
    const float2 x = unpack_bf16x2(a), y = unpack_bf16x2(b);
    acc = fmaf(x.x, y.x, acc);
    acc = fmaf(x.y, y.y, acc);
}

// acc[3][4] (16 x 24) = A(16 x 32) * B^T, B = rows 0..23 of `tile`
__device__ __forceinline__ void mma_16x24(float (&acc)[3][4], const uint32_t (&a)[2][4], uint32_t tile, int lane) {
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
        acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        uint32_t b[4];
        load_b_nk(b, tile, nt * 8, lane);
        mma16816(acc[nt], a[0], b[0], b[1]);
        mma16816(acc[nt], a[1], b[2], b[3]);
    }
}
// out[4][4] (16 x 32) = P(16 x 24, fp32 accumulator layout) * B, B = rows 0..23 of `tile` ([k][n = 32 channels])
__device__ __forceinline__ void mma_p_16x32(float (&out)[4][4], const float (&p)[3][4], uint32_t tile, int lane) {
    uint32_t a[4];
    a[0] = pack_bf16x2(p[0][0], p[0][1]);
    a[1] = pack_bf16x2(p[0][2], p[0][3]);
    a[2] = pack_bf16x2(p[1][0], p[1][1]);
    a[3] = pack_bf16x2(p[1][2], p[1][3]);
    const uint32_t a8_0 = pack_bf16x2(p[2][0], p[2][1]), a8_1 = pack_bf16x2(p[2][2], p[2][3]);
#pragma unroll
    for (int cp = 0; cp < 2; ++cp) {
        uint32_t b[4];
        load_b_kn(b, tile, 0, cp, lane);
        out[2 * cp][0] = out[2 * cp][1] = out[2 * cp][2] = out[2 * cp][3] = 0.f;
        out[2 * cp + 1][0] = out[2 * cp + 1][1] = out[2 * cp + 1][2] = out[2 * cp + 1][3] = 0.f;
        mma16816(out[2 * cp], a, b[0], b[1]);
        mma16816(out[2 * cp + 1], a, b[2], b[3]);
        load_b_kn(b, tile, 16, cp, lane);              // rows 16..23 in b[0], b[2] (rows 24..31 unused)
        mma1688(out[2 * cp], a8_0, a8_1, b[0]);
        mma1688(out[2 * cp + 1], a8_0, a8_1, b[2]);
    }
}

// 16 x 32 accumulator tile -> bf16 rows of a [rows][ld] matrix (rows r0 = first row of lane group g)
__device__ __forceinline__ void store_16x32(__nv_bfloat16* base, long long ld, int row0, int lane, const float (&o)[4][4],
                                            float s0, float s1) {
    const int g = lane >> 2, t = lane & 3;
    const int r0 = row0 + g, r1 = r0 + 8;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int d = a * 8 + 2 * t;
        if (r0 < SL) *reinterpret_cast<uint32_t*>(base + (long long)r0 * ld + d) = pack_bf16x2(o[a][0] * s0, o[a][1] * s0);
        if (r1 < SL) *reinterpret_cast<uint32_t*>(base + (long long)r1 * ld + d) = pack_bf16x2(o[a][2] * s1, o[a][3] * s1);
    }
}

// ---------------------------------------------------------------------------------------------
// producer: lane l loads tile l of the stage (tiles [0, n_qkv) from the packed qkv buffer at column
// 32 l, then n_x tiles per extra tensor)
// ---------------------------------------------------------------------------------------------
template <int NH, int NEXTRA>
__device__ __forceinline__ void produce(uint8_t* ring, uint64_t* full, uint64_t* empty, const CUtensorMap* tq,
                                        const CUtensorMap* tx0, const CUtensorMap* tx1, int nseq, int lane) {
    constexpr int NT = (3 + NEXTRA) * NH;
    constexpr uint32_t STAGE_BYTES = NT * TILE;
    int it = 0;
    for (int s = blockIdx.x; s < nseq; s += gridDim.x, ++it) {
        const int st = it % STAGES;
        mbar_wait(&empty[st], ((uint32_t)(it / STAGES) & 1u) ^ 1u);
        if (lane == 0) mbar_expect_tx(&full[st], STAGE_BYTES);
        __syncwarp();
        uint8_t* stage = ring + (size_t)st * STAGE_BYTES;
        for (int l = lane; l < NT; l += 32) {
            if (l < 3 * NH) tma_load_2d(stage + l * TILE, tq, &full[st], l * 32, s * SL);
            else if (l < 4 * NH) tma_load_2d(stage + l * TILE, tx0, &full[st], (l - 3 * NH) * 32, s * SL);
            else tma_load_2d(stage + l * TILE, tx1, &full[st], (l - 4 * NH) * 32, s * SL);
        }
    }
}

template <int NH>
__global__ void __launch_bounds__(32 * NH + 32)
attn_seq24_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, __nv_bfloat16* __restrict__ out,
                      float* __restrict__ lse, int nseq) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    constexpr uint32_t STAGE_BYTES = 3 * NH * TILE;
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + STAGES * STAGE_BYTES + 512);   // 512 B pad: ldmatrix over-read
    uint64_t* empty = full + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_qkv);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NH); }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == NH) {
        produce<NH, 0>(ring, full, empty, &tmap_qkv, nullptr, nullptr, nseq, lane);
        return;
    }
    const int h = warp, inner = NH * 32;
    const int g = lane >> 2, t = lane & 3;
    int it = 0;
    for (int s = blockIdx.x; s < nseq; s += gridDim.x, ++it) {
        const int st = it % STAGES;
        mbar_wait(&full[st], (uint32_t)(it / STAGES) & 1u);
        const uint32_t stage = smem_u32(ring + (size_t)st * STAGE_BYTES);
        const uint32_t tQ = stage + h * TILE, tK = stage + (NH + h) * TILE, tV = stage + (2 * NH + h) * TILE;
        __nv_bfloat16* ob = out + (long long)s * SL * inner + h * 32;
        float* lp = lse + ((long long)s * NH + h) * SL;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            uint32_t qa[2][4];
            load_a_frags(qa, tQ, mt * 16, lane);
            float sc[3][4];
            mma_16x24(sc, qa, tK, lane);
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
#pragma unroll
                for (int e = 0; e < 4; ++e) sc[nt][e] *= LOG2E;
                m0 = fmaxf(m0, fmaxf(sc[nt][0], sc[nt][1]));
                m1 = fmaxf(m1, fmaxf(sc[nt][2], sc[nt][3]));
            }
            m0 = quad_max(m0); m1 = quad_max(m1);
            float l0 = 0.f, l1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                sc[nt][0] = fast_exp2(sc[nt][0] - m0); sc[nt][1] = fast_exp2(sc[nt][1] - m0);
                sc[nt][2] = fast_exp2(sc[nt][2] - m1); sc[nt][3] = fast_exp2(sc[nt][3] - m1);
                l0 += sc[nt][0] + sc[nt][1];
                l1 += sc[nt][2] + sc[nt][3];
            }
            l0 = quad_sum(l0); l1 = quad_sum(l1);
            float o[4][4];
            mma_p_16x32(o, sc, tV, lane);
            store_16x32(ob, inner, mt * 16, lane, o, 1.f / l0, 1.f / l1);
            if (t == 0) {
                const int r0 = mt * 16 + g, r1 = r0 + 8;
                if (r0 < SL) lp[r0] = m0 * LN2 + __logf(l0);
                if (r1 < SL) lp[r1] = m1 * LN2 + __logf(l1);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
    }
}

// stage tiles: [q | k | v] x NH, dO x NH, O x NH
template <int NH>
__global__ void __launch_bounds__(32 * NH + 32)
attn_seq24_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                      const __grid_constant__ CUtensorMap tmap_o, const float* __restrict__ lse,
                      __nv_bfloat16* __restrict__ dqkv, int nseq) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    constexpr uint32_t STAGE_BYTES = 5 * NH * TILE;
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + STAGES * STAGE_BYTES + 512);
    uint64_t* empty = full + STAGES;
    float* sStat = reinterpret_cast<float*>(empty + STAGES);        // [NH][2][32]: lse*log2e, delta per query
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_qkv);
        tma_prefetch_desc(&tmap_do);
        tma_prefetch_desc(&tmap_o);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NH); }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == NH) {
        produce<NH, 2>(ring, full, empty, &tmap_qkv, &tmap_do, &tmap_o, nseq, lane);
        return;
    }
    const int h = warp, inner = NH * 32;
    const long long ld = 3LL * inner;
    const int g = lane >> 2, t = lane & 3;
    float* myL = sStat + h * 64;
    float* myD = myL + 32;
    int it = 0;
    for (int s = blockIdx.x; s < nseq; s += gridDim.x, ++it) {
        const int st = it % STAGES;
        mbar_wait(&full[st], (uint32_t)(it / STAGES) & 1u);
        const uint32_t stage = smem_u32(ring + (size_t)st * STAGE_BYTES);
        const uint32_t tQ = stage + h * TILE, tK = stage + (NH + h) * TILE, tV = stage + (2 * NH + h) * TILE;
        const uint32_t tdO = stage + (3 * NH + h) * TILE, tO = stage + (4 * NH + h) * TILE;
        __nv_bfloat16* gb = dqkv + (long long)s * SL * ld + h * 32;
        if (lane < SL) myL[lane] = __ldg(lse + ((long long)s * NH + h) * SL + lane) * LOG2E;
        __syncwarp();
        // ---- rows = queries: delta, dQ = dS K
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            uint32_t qa[2][4], da[2][4], oa[2][4];
            load_a_frags(qa, tQ, mt * 16, lane);
            load_a_frags(da, tdO, mt * 16, lane);
            load_a_frags(oa, tO, mt * 16, lane);
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                mul_bf16x2_sum(da[ks][0], oa[ks][0], d0); mul_bf16x2_sum(da[ks][2], oa[ks][2], d0);
                mul_bf16x2_sum(da[ks][1], oa[ks][1], d1); mul_bf16x2_sum(da[ks][3], oa[ks][3], d1);
            }
            d0 = quad_sum(d0); d1 = quad_sum(d1);
            const int r0 = mt * 16 + g, r1 = r0 + 8;
            if (t == 0) { myD[r0] = d0; myD[r1] = d1; }
            const float ls0 = myL[min(r0, SL - 1)], ls1 = myL[min(r1, SL - 1)];
            float sc[3][4], dp[3][4];
            mma_16x24(sc, qa, tK, lane);
            mma_16x24(dp, da, tV, lane);
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                const float p0 = fast_exp2(fmaf(sc[nt][0], LOG2E, -ls0)), p1 = fast_exp2(fmaf(sc[nt][1], LOG2E, -ls0));
                const float p2 = fast_exp2(fmaf(sc[nt][2], LOG2E, -ls1)), p3 = fast_exp2(fmaf(sc[nt][3], LOG2E, -ls1));
                sc[nt][0] = p0 * (dp[nt][0] - d0); sc[nt][1] = p1 * (dp[nt][1] - d0);
                sc[nt][2] = p2 * (dp[nt][2] - d1); sc[nt][3] = p3 * (dp[nt][3] - d1);
            }
            float dq[4][4];
            mma_p_16x32(dq, sc, tK, lane);
            store_16x32(gb, ld, mt * 16, lane, dq, 1.f, 1.f);
        }
        __syncwarp();                                   // delta of every query is in myD
        // ---- rows = keys: dV = P^T dO, dK = dS^T Q
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            uint32_t ka[2][4], va[2][4];
            load_a_frags(ka, tK, mt * 16, lane);
            load_a_frags(va, tV, mt * 16, lane);
            float stt[3][4], dpt[3][4];
            mma_16x24(stt, ka, tQ, lane);               // S^T[key, query]
            mma_16x24(dpt, va, tdO, lane);              // dP^T[key, query]
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                const int c = nt * 8 + 2 * t;
                const float2 lc = *reinterpret_cast<const float2*>(myL + c);
                const float2 dc = *reinterpret_cast<const float2*>(myD + c);
                const float p0 = fast_exp2(fmaf(stt[nt][0], LOG2E, -lc.x)), p1 = fast_exp2(fmaf(stt[nt][1], LOG2E, -lc.y));
                const float p2 = fast_exp2(fmaf(stt[nt][2], LOG2E, -lc.x)), p3 = fast_exp2(fmaf(stt[nt][3], LOG2E, -lc.y));
                stt[nt][0] = p0; stt[nt][1] = p1; stt[nt][2] = p2; stt[nt][3] = p3;
                dpt[nt][0] = p0 * (dpt[nt][0] - dc.x); dpt[nt][1] = p1 * (dpt[nt][1] - dc.y);
                dpt[nt][2] = p2 * (dpt[nt][2] - dc.x); dpt[nt][3] = p3 * (dpt[nt][3] - dc.y);
            }
            float acc[4][4];
            mma_p_16x32(acc, stt, tdO, lane);           // dV
            store_16x32(gb + 2 * inner, ld, mt * 16, lane, acc, 1.f, 1.f);
            mma_p_16x32(acc, dpt, tQ, lane);            // dK
            store_16x32(gb + inner, ld, mt * 16, lane, acc, 1.f, 1.f);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
    }
}

template <int NH>
int launch_fwd(const CUtensorMap& tq, __nv_bfloat16* out, float* lse, int nseq, cudaStream_t s) {
    auto kern = attn_seq24_fwd_kernel<NH>;
    constexpr int smem = STAGES * 3 * NH * TILE + 512 + 2 * STAGES * 8 + 1024;
    CTK_SET_MAX_SMEM(kern, smem);
    const int per_sm = (227 * 1024) / (smem + 1024) > 0 ? (227 * 1024) / (smem + 1024) : 1;
    int grid = ctk_num_sms() * (per_sm > 4 ? 4 : per_sm);
    if (grid > nseq) grid = nseq;
    kern<<<grid, 32 * NH + 32, smem, s>>>(tq, out, lse, nseq);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}
template <int NH>
int launch_bwd(const CUtensorMap& tq, const CUtensorMap& tdo, const CUtensorMap& to, const float* lse, __nv_bfloat16* dqkv,
               int nseq, cudaStream_t s) {
    auto kern = attn_seq24_bwd_kernel<NH>;
    constexpr int smem = STAGES * 5 * NH * TILE + 512 + 2 * STAGES * 8 + NH * 64 * 4 + 1024;
    CTK_SET_MAX_SMEM(kern, smem);
    const int per_sm = (227 * 1024) / (smem + 1024) > 0 ? (227 * 1024) / (smem + 1024) : 1;
    int grid = ctk_num_sms() * (per_sm > 4 ? 4 : per_sm);
    if (grid > nseq) grid = nseq;
    kern<<<grid, 32 * NH + 32, smem, s>>>(tq, tdo, to, lse, dqkv, nseq);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

int make_maps(CUtensorMap* m, const void* p, int cols, unsigned long long rows) {
    const unsigned long long dims[2] = {(unsigned long long)cols, rows};
    const unsigned long long strides[1] = {(unsigned long long)cols * 2};
    const unsigned int box[2] = {32, SL};
    return ctk_make_tmap(m, p, false, 2, dims, strides, box, 2);
}

}  // namespace

bool ctk_attn_seq24_supported(int L, int heads) { return L == SL && (heads == 8 || heads == 4 || heads == 2); }

int ctk_attn_seq24_fwd(const void* qkv, void* out, float* lse, int nseq, int heads, cudaStream_t s) {
    CUtensorMap tq;
    int rc;
    if ((rc = make_maps(&tq, qkv, 3 * heads * 32, (unsigned long long)nseq * SL))) return rc;
    auto o = reinterpret_cast<__nv_bfloat16*>(out);
    if (heads == 8) return launch_fwd<8>(tq, o, lse, nseq, s);
    if (heads == 4) return launch_fwd<4>(tq, o, lse, nseq, s);
    return launch_fwd<2>(tq, o, lse, nseq, s);
}

int ctk_attn_seq24_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int nseq,
                       int heads, cudaStream_t s) {
    CUtensorMap tq, tdo, to;
    int rc;
    const unsigned long long rows = (unsigned long long)nseq * SL;
    if ((rc = make_maps(&tq, qkv, 3 * heads * 32, rows))) return rc;
    if ((rc = make_maps(&tdo, dout, heads * 32, rows))) return rc;
    if ((rc = make_maps(&to, out, heads * 32, rows))) return rc;
    auto g = reinterpret_cast<__nv_bfloat16*>(dqkv);
    if (heads == 8) return launch_bwd<8>(tq, tdo, to, lse, g, nseq, s);
    if (heads == 4) return launch_bwd<4>(tq, tdo, to, lse, g, nseq, s);
    return launch_bwd<2>(tq, tdo, to, lse, g, nseq, s);
}

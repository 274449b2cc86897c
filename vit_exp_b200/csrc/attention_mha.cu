// Multi-head softmax attention on packed bf16 projections, flash-style, for the two places of the reference that use a
// plain softmax(q k^T * scale + key mask) v core:
//   * head dim 64: HF BertSelfAttention of the text tower (CT_CLIP/ct_clip/ct_clip.py:1271; reports padded to 512 tokens,
//     scripts/CTCLIPTrainer.py:562): key-padding mask, dropout on the probabilities;
//   * head dim 32: `FlashAttention` of CTViT3D (transformer_maskgit/attention.py:189-284: SDPA over 13 824 tokens plus
//     2 learned null key/value pairs that every query also attends to, attention.py:240-248).
// Input: qkv [B*L, 3H] (q | k | v, heads contiguous inside each third); optional null pairs null_k / null_v
// [heads, n_null, DH] shared by all sequences of the batch, processed as one extra key block.
//
// Warp-level tensor-core path (mma.sync m16n8k16, bf16 in / fp32 accumulate, ldmatrix from XOR-swizzled shared memory,
// cp.async double buffering): the L x L probabilities only ever live in registers.
//   forward : CTA = 64 queries (4 warps x 16), streams 64-key blocks; online softmax; out bf16 + lse fp32
//   backward: delta = rowsum(dO . O); dQ kernel (CTA = 64 queries, streams keys); dK/dV kernel (CTA = 64 keys, streams
//             queries, transposed formulation; one more CTA per (head, sequence) for the null pairs) - no atomics,
//             deterministic.
// Why not tcgen05 here: the text tower's problem is 0.14 % of the train step's tensor work in 768 items of 64 x 512 x 64
// (a 128-row TMEM tile cannot amortise its round trip over 8 key blocks); the CTViT3D core is exponent-bound at head dim
// 32 (one ex2 per 128 tensor FLOPs) and is the "next" tier of the scope table - see DESIGN.md.
// Dropout of the probabilities (training mode, CXR-BERT ships 0.1) uses a counter-based hash of (seed, report*head,
// query, key), regenerated in the backward kernels; masks are statistically, not bitwise, torch's.
#include "attention_mma.cuh"
#include "mha_dropout.cuh"

using namespace attn_mma;

namespace {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int BQ = 64;             // rows per CTA (queries; keys in the dK/dV kernel)
constexpr int BK = 64;             // streamed block
template <int DH> constexpr int tile_bytes() { return 64 * DH * 2; }     // one [64][DH] bf16 tile

// [rows][DH bf16] tile; 16-byte chunks XOR-swizzled so that ldmatrix (8 rows x 16 bytes) is conflict-free:
// DH = 64: 128-byte rows, chunk ^= row & 7;   DH = 32: 64-byte rows, chunk ^= (row >> 1) & 3
template <int DH> __device__ __forceinline__ uint32_t toff(int row, int chunk) {
    if constexpr (DH == 64) return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
    else return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;            // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// rows [row0, row0+64) of a [*, ld] bf16 matrix (DH columns starting at `src`) -> swizzled tile; rows >= nrows are zero
template <int DH>
__device__ __forceinline__ void load_tile64(uint32_t dst, const __nv_bfloat16* src, long long ld, int row0, int nrows, int tid) {
    constexpr int CPR = DH / 8;                  // 16-byte chunks per row
#pragma unroll
    for (int i = tid; i < 64 * CPR; i += 128) {
        const int r = i / CPR, c = i % CPR;
        const bool ok = row0 + r < nrows;
        const __nv_bfloat16* p = src + (long long)(ok ? row0 + r : 0) * ld + c * 8;
        cp_async16(dst + toff<DH>(r, c), p, ok);
    }
}

// A fragments (16 rows x DH k) of the warp's rows r0..r0+15
template <int DH>
__device__ __forceinline__ void load_a64(uint32_t (&a)[DH / 16][4], uint32_t tile, int r0, int lane) {
    const int m = lane >> 3;
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) ldsm4(a[ks], tile + toff<DH>(r0 + (m & 1) * 8 + (lane & 7), 2 * ks + (m >> 1)));
}
// acc (16 x 64) = A(16 x DH) * T^T, T = [64 n][DH k] tile (operand stored [n][k])
template <int DH>
__device__ __forceinline__ void mma_nk(float (&acc)[8][4], const uint32_t (&a)[DH / 16][4], uint32_t tile, int lane) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
        for (int half = 0; half < DH / 32; ++half) {
            uint32_t b[4];
            ldsm4(b, tile + toff<DH>(nt * 8 + (lane & 7), half * 4 + (lane >> 3)));
            mma16816(acc[nt], a[2 * half], b[0], b[1]);
            mma16816(acc[nt], a[2 * half + 1], b[2], b[3]);
        }
    }
}
// out (16 x DH) += P(16 x 64, fp32 accumulator layout, rounded to bf16) * T, T = [64 k][DH n] tile (operand stored [k][n])
template <int DH>
__device__ __forceinline__ void mma_kn(float (&out)[DH / 8][4], const float (&p)[8][4], uint32_t tile, int lane) {
    const int m = lane >> 3;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[4];
        a[0] = pack_bf16x2(p[2 * kk][0], p[2 * kk][1]);
        a[1] = pack_bf16x2(p[2 * kk][2], p[2 * kk][3]);
        a[2] = pack_bf16x2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
        a[3] = pack_bf16x2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
        for (int cp = 0; cp < DH / 16; ++cp) {
            uint32_t b[4];
            ldsm4t(b, tile + toff<DH>(kk * 16 + (m & 1) * 8 + (lane & 7), 2 * cp + (m >> 1)));
            mma16816(out[2 * cp], a, b[0], b[1]);
            mma16816(out[2 * cp + 1], a, b[2], b[3]);
        }
    }
}
// store a 16 x DH fp32 accumulator (times `mul`) as rows of a [*, ld] matrix (bf16 or fp32); rows >= nrows are skipped
template <int DH>
__device__ __forceinline__ void store_acc_f32(const float (&o)[DH / 8][4], float mul, float* dst, long long ld, int row_a,
                                              int nrows, int lane) {
    const int t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) {
        if (row_a < nrows) *reinterpret_cast<float2*>(dst + (long long)row_a * ld + nt * 8 + 2 * t) = make_float2(o[nt][0] * mul, o[nt][1] * mul);
        if (row_a + 8 < nrows)
            *reinterpret_cast<float2*>(dst + (long long)(row_a + 8) * ld + nt * 8 + 2 * t) = make_float2(o[nt][2] * mul, o[nt][3] * mul);
    }
}
template <int DH>
__device__ __forceinline__ void store_acc(const float (&o)[DH / 8][4], float mul0, float mul1, __nv_bfloat16* dst, long long ld,
                                          int row_a, int nrows, int lane) {
    const int t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) {
        if (row_a < nrows)
            *reinterpret_cast<uint32_t*>(dst + (long long)row_a * ld + nt * 8 + 2 * t) = pack_bf16x2(o[nt][0] * mul0, o[nt][1] * mul0);
        if (row_a + 8 < nrows)
            *reinterpret_cast<uint32_t*>(dst + (long long)(row_a + 8) * ld + nt * 8 + 2 * t) = pack_bf16x2(o[nt][2] * mul1, o[nt][3] * mul1);
    }
}


// Key blocks: 0 .. nkb-1 are rows of the packed buffer; block nkb (present when n_null > 0) holds the learned null pairs
// [heads, n_null, DH], which every query of every sequence attends to and which no key mask touches.
struct KeySrc {
    const __nv_bfloat16* k;       // keys of this (sequence, head): row pitch ld
    const __nv_bfloat16* v;
    long long ld;
    const __nv_bfloat16* nk;      // null keys / values of this head (row pitch DH) or nullptr
    const __nv_bfloat16* nv;
    const unsigned char* mask;    // key mask of this sequence [L] or nullptr
    int L, n_null, nkb;
};
template <int DH>
__device__ __forceinline__ void load_kv_block(const KeySrc& ks, int kb, uint32_t sK, uint32_t sV, float* sM, int tid) {
    if (kb < ks.nkb) {
        load_tile64<DH>(sK, ks.k, ks.ld, kb * BK, ks.L, tid);
        load_tile64<DH>(sV, ks.v, ks.ld, kb * BK, ks.L, tid);
        if (tid < BK) {
            const int j = kb * BK + tid;
            sM[tid] = (j < ks.L && (ks.mask == nullptr || ks.mask[j] != 0)) ? 0.f : -INFINITY;
        }
    } else {
        load_tile64<DH>(sK, ks.nk, DH, 0, ks.n_null, tid);
        load_tile64<DH>(sV, ks.nv, DH, 0, ks.n_null, tid);
        if (tid < BK) sM[tid] = tid < ks.n_null ? 0.f : -INFINITY;
    }
}

// ------------------------------------------------------------------------------------------------
// forward.  grid (ceil(L/64), heads, B), 128 threads.  smem: Q | K[2] | V[2] | additive key mask [2][64]
// ------------------------------------------------------------------------------------------------
template <int DH, bool DROP>
__global__ void __launch_bounds__(128)
mha_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const unsigned char* __restrict__ key_mask,
               const __nv_bfloat16* __restrict__ null_k, const __nv_bfloat16* __restrict__ null_v, int n_null,
               __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int L, int heads, float scale_log2,
               uint32_t drop_thresh, float inv_keep, const unsigned long long* __restrict__ seed_ptr, unsigned long long seed_off) {
    constexpr int TILE = tile_bytes<DH>();
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sQ = smem_u32(smem), sK = sQ + TILE, sV = sK + 2 * TILE;
    float* sM = reinterpret_cast<float*>(smem + 5 * TILE);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
    const int H = heads * DH;
    const long long ld = 3LL * H;
    const __nv_bfloat16* base = qkv + (long long)b * L * ld + h * DH;
    KeySrc ks;
    ks.k = base + H; ks.v = base + 2 * H; ks.ld = ld;
    ks.nk = n_null ? null_k + (long long)h * n_null * DH : nullptr;
    ks.nv = n_null ? null_v + (long long)h * n_null * DH : nullptr;
    ks.mask = key_mask ? key_mask + (long long)b * L : nullptr;
    ks.L = L; ks.n_null = n_null; ks.nkb = (L + BK - 1) / BK;
    const int nblk = ks.nkb + (n_null ? 1 : 0);

    load_tile64<DH>(sQ, base, ld, q0, L, tid);
    load_kv_block<DH>(ks, 0, sK, sV, sM, tid);
    cp_commit();

    uint32_t qa[DH / 16][4];
    float o[DH / 8][4];
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const int row_a = q0 + warp * 16 + g;               // this lane's rows: row_a and row_a + 8
    const uint32_t bh = (uint32_t)(b * heads + h);
    const unsigned long long seed = mha_seed(DROP && seed_ptr ? *seed_ptr : 0ull, seed_off);

    for (int kb = 0; kb < nblk; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nblk) load_kv_block<DH>(ks, kb + 1, sK + (buf ^ 1) * TILE, sV + (buf ^ 1) * TILE, sM + (buf ^ 1) * BK, tid);
        cp_commit();
        cp_wait<1>();
        __syncthreads();
        if (kb == 0) load_a64<DH>(qa, sQ, warp * 16, lane);
        float s[8][4];
        mma_nk<DH>(s, qa, sK + buf * TILE, lane);
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float ma = sM[buf * BK + nt * 8 + 2 * t], mb = sM[buf * BK + nt * 8 + 2 * t + 1];
            s[nt][0] = fmaf(s[nt][0], scale_log2, ma);
            s[nt][1] = fmaf(s[nt][1], scale_log2, mb);
            s[nt][2] = fmaf(s[nt][2], scale_log2, ma);
            s[nt][3] = fmaf(s[nt][3], scale_log2, mb);
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = quad_max(mx0);
        mx1 = quad_max(mx1);
        const float ms0 = mx0 == -INFINITY ? 0.f : mx0, ms1 = mx1 == -INFINITY ? 0.f : mx1;   // all keys masked so far
        const float al0 = fast_exp2(m0 - ms0), al1 = fast_exp2(m1 - ms1);
        m0 = mx0; m1 = mx1;
        float r0 = 0.f, r1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = fast_exp2(s[nt][0] - ms0);
            s[nt][1] = fast_exp2(s[nt][1] - ms0);
            s[nt][2] = fast_exp2(s[nt][2] - ms1);
            s[nt][3] = fast_exp2(s[nt][3] - ms1);
            r0 += s[nt][0] + s[nt][1];
            r1 += s[nt][2] + s[nt][3];
            if constexpr (DROP) {
                const uint32_t j = (uint32_t)(kb * BK + nt * 8 + 2 * t);
                s[nt][0] = mha_keep(seed, bh, (uint32_t)row_a, j, drop_thresh) ? s[nt][0] * inv_keep : 0.f;
                s[nt][1] = mha_keep(seed, bh, (uint32_t)row_a, j + 1, drop_thresh) ? s[nt][1] * inv_keep : 0.f;
                s[nt][2] = mha_keep(seed, bh, (uint32_t)row_a + 8, j, drop_thresh) ? s[nt][2] * inv_keep : 0.f;
                s[nt][3] = mha_keep(seed, bh, (uint32_t)row_a + 8, j + 1, drop_thresh) ? s[nt][3] * inv_keep : 0.f;
            }
        }
        l0 = l0 * al0 + quad_sum(r0);
        l1 = l1 * al1 + quad_sum(r1);
#pragma unroll
        for (int nt = 0; nt < DH / 8; ++nt) {
            o[nt][0] *= al0; o[nt][1] *= al0;
            o[nt][2] *= al1; o[nt][3] *= al1;
        }
        mma_kn<DH>(o, s, sV + buf * TILE, lane);
        __syncthreads();
    }
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
    store_acc<DH>(o, i0, i1, out + (long long)b * L * H + h * DH, H, row_a, L, lane);
    if (t == 0) {
        float* lp = lse + ((long long)b * heads + h) * L;
        if (row_a < L) lp[row_a] = (m0 == -INFINITY ? 0.f : m0) * LN2 + __logf(fmaxf(l0, 1e-30f));
        if (row_a + 8 < L) lp[row_a + 8] = (m1 == -INFINITY ? 0.f : m1) * LN2 + __logf(fmaxf(l1, 1e-30f));
    }
}

// delta[b, h, i] = sum_d dO[i, h, d] * O[i, h, d].  DH / 8 lanes per (token, head).
template <int DH>
__global__ void __launch_bounds__(256)
mha_delta_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout, float* __restrict__ delta,
                 int B, int L, int heads) {
    constexpr int LPI = DH / 8;                                                     // lanes per item
    const long long item = ((long long)blockIdx.x * 256 + threadIdx.x) / LPI;      // (b*L + i) * heads + h
    const int sub = threadIdx.x % LPI;
    const long long total = (long long)B * L * heads;
    float acc = 0.f;
    if (item < total) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(out + item * DH) + sub);
        const uint4 d = __ldg(reinterpret_cast<const uint4*>(dout + item * DH) + sub);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 x = unpack_bf16x2(aw[e]), y = unpack_bf16x2(dw[e]);
            acc = fmaf(x.x, y.x, acc);
            acc = fmaf(x.y, y.y, acc);
        }
    }
#pragma unroll
    for (int o = 1; o < LPI; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (item < total && sub == 0) {
        const long long tok = item / heads;
        const int h = (int)(item % heads);
        const long long b = tok / L, i = tok % L;
        delta[(b * heads + h) * L + i] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// backward, dQ.  grid (ceil(L/64), heads, B).  smem: Q | dO | K[2] | V[2] | key mask [2][64]
// ------------------------------------------------------------------------------------------------
template <int DH, bool DROP>
__global__ void __launch_bounds__(128)
mha_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const unsigned char* __restrict__ key_mask,
                  const __nv_bfloat16* __restrict__ null_k, const __nv_bfloat16* __restrict__ null_v, int n_null,
                  const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse, const float* __restrict__ delta,
                  __nv_bfloat16* __restrict__ dqkv, int L, int heads, float scale, uint32_t drop_thresh, float inv_keep,
                  const unsigned long long* __restrict__ seed_ptr, unsigned long long seed_off) {
    constexpr int TILE = tile_bytes<DH>();
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sQ = smem_u32(smem), sDO = sQ + TILE, sK = sDO + TILE, sV = sK + 2 * TILE;
    float* sM = reinterpret_cast<float*>(smem + 6 * TILE);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
    const int H = heads * DH;
    const long long ld = 3LL * H;
    const __nv_bfloat16* base = qkv + (long long)b * L * ld + h * DH;
    const float scale_log2 = scale * LOG2E;
    KeySrc ks;
    ks.k = base + H; ks.v = base + 2 * H; ks.ld = ld;
    ks.nk = n_null ? null_k + (long long)h * n_null * DH : nullptr;
    ks.nv = n_null ? null_v + (long long)h * n_null * DH : nullptr;
    ks.mask = key_mask ? key_mask + (long long)b * L : nullptr;
    ks.L = L; ks.n_null = n_null; ks.nkb = (L + BK - 1) / BK;
    const int nblk = ks.nkb + (n_null ? 1 : 0);

    load_tile64<DH>(sQ, base, ld, q0, L, tid);
    load_tile64<DH>(sDO, dout + (long long)b * L * H + h * DH, H, q0, L, tid);
    load_kv_block<DH>(ks, 0, sK, sV, sM, tid);
    cp_commit();

    const int row_a = q0 + warp * 16 + g;
    const float* lp = lse + ((long long)b * heads + h) * L;
    const float* dp = delta + ((long long)b * heads + h) * L;
    // rows beyond L: lse = +inf makes every probability of the row 0
    const float ls0 = row_a < L ? lp[row_a] * LOG2E : INFINITY, ls1 = row_a + 8 < L ? lp[row_a + 8] * LOG2E : INFINITY;
    const float de0 = row_a < L ? dp[row_a] : 0.f, de1 = row_a + 8 < L ? dp[row_a + 8] : 0.f;
    const uint32_t bh = (uint32_t)(b * heads + h);
    const unsigned long long seed = mha_seed(DROP && seed_ptr ? *seed_ptr : 0ull, seed_off);

    uint32_t qa[DH / 16][4], da[DH / 16][4];
    float dq[DH / 8][4];
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;

    for (int kb = 0; kb < nblk; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nblk) load_kv_block<DH>(ks, kb + 1, sK + (buf ^ 1) * TILE, sV + (buf ^ 1) * TILE, sM + (buf ^ 1) * BK, tid);
        cp_commit();
        cp_wait<1>();
        __syncthreads();
        if (kb == 0) {
            load_a64<DH>(qa, sQ, warp * 16, lane);
            load_a64<DH>(da, sDO, warp * 16, lane);
        }
        float s[8][4], dpv[8][4];
        mma_nk<DH>(s, qa, sK + buf * TILE, lane);            // logits
        mma_nk<DH>(dpv, da, sV + buf * TILE, lane);          // dP = dO V^T
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float ma = sM[buf * BK + nt * 8 + 2 * t], mb = sM[buf * BK + nt * 8 + 2 * t + 1];
            const float p0 = fast_exp2(fmaf(s[nt][0], scale_log2, ma) - ls0);
            const float p1 = fast_exp2(fmaf(s[nt][1], scale_log2, mb) - ls0);
            const float p2 = fast_exp2(fmaf(s[nt][2], scale_log2, ma) - ls1);
            const float p3 = fast_exp2(fmaf(s[nt][3], scale_log2, mb) - ls1);
            float d0 = dpv[nt][0], d1 = dpv[nt][1], d2 = dpv[nt][2], d3 = dpv[nt][3];
            if constexpr (DROP) {
                const uint32_t j = (uint32_t)(kb * BK + nt * 8 + 2 * t);
                d0 = mha_keep(seed, bh, (uint32_t)row_a, j, drop_thresh) ? d0 * inv_keep : 0.f;
                d1 = mha_keep(seed, bh, (uint32_t)row_a, j + 1, drop_thresh) ? d1 * inv_keep : 0.f;
                d2 = mha_keep(seed, bh, (uint32_t)row_a + 8, j, drop_thresh) ? d2 * inv_keep : 0.f;
                d3 = mha_keep(seed, bh, (uint32_t)row_a + 8, j + 1, drop_thresh) ? d3 * inv_keep : 0.f;
            }
            s[nt][0] = p0 * (d0 - de0);                  // dS
            s[nt][1] = p1 * (d1 - de0);
            s[nt][2] = p2 * (d2 - de1);
            s[nt][3] = p3 * (d3 - de1);
        }
        mma_kn<DH>(dq, s, sK + buf * TILE, lane);            // dQ += dS K
        __syncthreads();
    }
    store_acc<DH>(dq, scale, scale, dqkv + (long long)b * L * ld + h * DH, ld, row_a, L, lane);
}

// ------------------------------------------------------------------------------------------------
// backward, dK and dV (transposed: rows = keys).  grid (ceil(L/64) [+ 1 for the null pairs], heads, B).
// smem: K | V | Q[2] | dO[2] | lse*log2e [2][64] | delta [2][64]
// The extra CTA owns the null pairs: its gradients go to dnull_k / dnull_v [B, heads, n_null, DH] fp32 (one slab per
// sequence; the caller sums over the batch).
// ------------------------------------------------------------------------------------------------
template <int DH, bool DROP>
__global__ void __launch_bounds__(128)
mha_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const unsigned char* __restrict__ key_mask,
                   const __nv_bfloat16* __restrict__ null_k, const __nv_bfloat16* __restrict__ null_v, int n_null,
                   const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse, const float* __restrict__ delta,
                   __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dnull_k, float* __restrict__ dnull_v, int L, int heads,
                   float scale, uint32_t drop_thresh, float inv_keep, const unsigned long long* __restrict__ seed_ptr,
                   unsigned long long seed_off) {
    constexpr int TILE = tile_bytes<DH>();
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sK = smem_u32(smem), sV = sK + TILE, sQ = sV + TILE, sDO = sQ + 2 * TILE;
    float* sL = reinterpret_cast<float*>(smem + 6 * TILE);
    float* sD = sL + 2 * BK;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int h = blockIdx.y, b = blockIdx.z;
    const int nkb = (L + BK - 1) / BK;
    const bool is_null = (int)blockIdx.x >= nkb;         // CTA-uniform
    const int k0 = is_null ? 0 : blockIdx.x * BQ;
    const int H = heads * DH;
    const long long ld = 3LL * H;
    const __nv_bfloat16* base = qkv + (long long)b * L * ld + h * DH;
    const __nv_bfloat16* dob = dout + (long long)b * L * H + h * DH;
    const float* lp = lse + ((long long)b * heads + h) * L;
    const float* dp = delta + ((long long)b * heads + h) * L;
    const int nqb = (L + BK - 1) / BK;
    const float scale_log2 = scale * LOG2E;

    auto load_q = [&](int qb, int buf) {
        load_tile64<DH>(sQ + buf * TILE, base, ld, qb * BK, L, tid);
        load_tile64<DH>(sDO + buf * TILE, dob, H, qb * BK, L, tid);
        if (tid < BK) {
            const int i = qb * BK + tid;
            sL[buf * BK + tid] = i < L ? lp[i] * LOG2E : INFINITY;      // rows beyond L: probability 0
            sD[buf * BK + tid] = i < L ? dp[i] : 0.f;
        }
    };
    if (is_null) {
        load_tile64<DH>(sK, null_k + (long long)h * n_null * DH, DH, 0, n_null, tid);
        load_tile64<DH>(sV, null_v + (long long)h * n_null * DH, DH, 0, n_null, tid);
    } else {
        load_tile64<DH>(sK, base + H, ld, k0, L, tid);
        load_tile64<DH>(sV, base + 2 * H, ld, k0, L, tid);
    }
    load_q(0, 0);
    cp_commit();

    const int row_a = k0 + warp * 16 + g;                // this lane's keys: row_a and row_a + 8
    const int nkeys = is_null ? n_null : L;
    const unsigned char* mk = (!is_null && key_mask) ? key_mask + (long long)b * L : nullptr;
    const bool ok0 = row_a < nkeys && (mk == nullptr || mk[row_a] != 0);
    const bool ok1 = row_a + 8 < nkeys && (mk == nullptr || mk[row_a + 8] != 0);
    const float km0 = ok0 ? 0.f : -INFINITY, km1 = ok1 ? 0.f : -INFINITY;
    const uint32_t bh = (uint32_t)(b * heads + h);
    const uint32_t jcol = (uint32_t)(is_null ? nkb * BK : 0) + (uint32_t)row_a;      // key index in the dropout hash
    const unsigned long long seed = mha_seed(DROP && seed_ptr ? *seed_ptr : 0ull, seed_off);

    uint32_t ka[DH / 16][4], va[DH / 16][4];
    float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) {
        dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f;
        dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f;
    }
    for (int qb = 0; qb < nqb; ++qb) {
        const int buf = qb & 1;
        if (qb + 1 < nqb) load_q(qb + 1, buf ^ 1);
        cp_commit();
        cp_wait<1>();
        __syncthreads();
        if (qb == 0) {
            load_a64<DH>(ka, sK, warp * 16, lane);
            load_a64<DH>(va, sV, warp * 16, lane);
        }
        float st[8][4], dpt[8][4];
        mma_nk<DH>(st, ka, sQ + buf * TILE, lane);           // S^T  [key][query]
        mma_nk<DH>(dpt, va, sDO + buf * TILE, lane);         // dP^T = V dO^T
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int c = buf * BK + nt * 8 + 2 * t;
            const float la = sL[c], lb = sL[c + 1], da_ = sD[c], db_ = sD[c + 1];
            const float p0 = fast_exp2(fmaf(st[nt][0], scale_log2, km0) - la);
            const float p1 = fast_exp2(fmaf(st[nt][1], scale_log2, km0) - lb);
            const float p2 = fast_exp2(fmaf(st[nt][2], scale_log2, km1) - la);
            const float p3 = fast_exp2(fmaf(st[nt][3], scale_log2, km1) - lb);
            float w0 = 1.f, w1 = 1.f, w2 = 1.f, w3 = 1.f;          // dropout keep / (1 - p)
            if constexpr (DROP) {
                const uint32_t i = (uint32_t)(qb * BK + nt * 8 + 2 * t);
                w0 = mha_keep(seed, bh, i, jcol, drop_thresh) ? inv_keep : 0.f;
                w1 = mha_keep(seed, bh, i + 1, jcol, drop_thresh) ? inv_keep : 0.f;
                w2 = mha_keep(seed, bh, i, jcol + 8, drop_thresh) ? inv_keep : 0.f;
                w3 = mha_keep(seed, bh, i + 1, jcol + 8, drop_thresh) ? inv_keep : 0.f;
            }
            st[nt][0] = p0 * w0; st[nt][1] = p1 * w1; st[nt][2] = p2 * w2; st[nt][3] = p3 * w3;       // dropped P^T
            dpt[nt][0] = p0 * (dpt[nt][0] * w0 - da_);                                                  // dS^T
            dpt[nt][1] = p1 * (dpt[nt][1] * w1 - db_);
            dpt[nt][2] = p2 * (dpt[nt][2] * w2 - da_);
            dpt[nt][3] = p3 * (dpt[nt][3] * w3 - db_);
        }
        mma_kn<DH>(dv, st, sDO + buf * TILE, lane);          // dV += P^T dO
        mma_kn<DH>(dk, dpt, sQ + buf * TILE, lane);          // dK += dS^T Q
        __syncthreads();
    }
    if (is_null) {
        const long long slab = ((long long)b * heads + h) * n_null * DH;
        store_acc_f32<DH>(dk, scale, dnull_k + slab, DH, row_a, n_null, lane);
        store_acc_f32<DH>(dv, 1.f, dnull_v + slab, DH, row_a, n_null, lane);
    } else {
        store_acc<DH>(dk, scale, scale, dqkv + (long long)b * L * ld + H + h * DH, ld, row_a, L, lane);
        store_acc<DH>(dv, 1.f, 1.f, dqkv + (long long)b * L * ld + 2 * H + h * DH, ld, row_a, L, lane);
    }
}

int check_args(int B, int L, int heads, int dh, float p_drop, const unsigned long long* seed_ptr, const void* null_k,
               const void* null_v, int n_null) {
    CTK_REQUIRE(B > 0 && L > 0 && heads > 0, CTK_ERR_SHAPE, "mha: bad shape B %d L %d heads %d", B, L, heads);
    CTK_REQUIRE(dh == 64 || dh == 32, CTK_ERR_SHAPE, "mha: head dim %d (32 and 64 are instantiated)", dh);
    CTK_REQUIRE(p_drop >= 0.f && p_drop < 1.f, CTK_ERR_SHAPE, "mha: dropout probability %f", (double)p_drop);
    CTK_REQUIRE(B <= 65535 && heads <= 65535, CTK_ERR_SHAPE, "mha: grid limits");
    CTK_REQUIRE(p_drop == 0.f || seed_ptr != nullptr, CTK_ERR_SHAPE, "mha: dropout needs the device seed pointer");
    CTK_REQUIRE(n_null >= 0 && n_null <= BK && (n_null == 0 || (null_k && null_v && CTK_ALIGNED(null_k, 16) && CTK_ALIGNED(null_v, 16))),
                CTK_ERR_SHAPE, "mha: null key/value pairs: 0 <= n_null <= 64 with 16-byte aligned buffers");
    return CTK_OK;
}

template <int DH>
int launch_fwd(const __nv_bfloat16* qkv, const unsigned char* key_mask, const __nv_bfloat16* nk, const __nv_bfloat16* nv, int n_null,
               __nv_bfloat16* out, float* lse, int B, int L, int heads, float scale, float p_drop,
               const unsigned long long* seed_ptr, unsigned long long seed_off, cudaStream_t s) {
    const dim3 grid((L + BQ - 1) / BQ, heads, B);
    constexpr int smem = 5 * tile_bytes<DH>() + 2 * BK * 4;
    const float scale_log2 = scale * LOG2E;
    if (p_drop > 0.f) {
        CTK_SET_MAX_SMEM((mha_fwd_kernel<DH, true>), smem);
        mha_fwd_kernel<DH, true><<<grid, 128, smem, s>>>(qkv, key_mask, nk, nv, n_null, out, lse, L, heads, scale_log2,
                                                         mha_drop_threshold(p_drop), 1.f / (1.f - p_drop), seed_ptr, seed_off);
    } else {
        CTK_SET_MAX_SMEM((mha_fwd_kernel<DH, false>), smem);
        mha_fwd_kernel<DH, false><<<grid, 128, smem, s>>>(qkv, key_mask, nk, nv, n_null, out, lse, L, heads, scale_log2, 0u, 1.f,
                                                          nullptr, 0ull);
    }
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

template <int DH>
int launch_bwd(const __nv_bfloat16* q, const unsigned char* key_mask, const __nv_bfloat16* nk, const __nv_bfloat16* nv, int n_null,
               const __nv_bfloat16* o, const __nv_bfloat16* d_o, const float* lse, float* delta, __nv_bfloat16* dq, float* dnk,
               float* dnv, int B, int L, int heads, float scale, float p_drop, const unsigned long long* seed_ptr,
               unsigned long long seed_off, cudaStream_t s) {
    const long long items = (long long)B * L * heads;
    mha_delta_kernel<DH><<<(unsigned)((items * (DH / 8) + 255) / 256), 256, 0, s>>>(o, d_o, delta, B, L, heads);
    CTK_LAUNCH_CHECK();
    const int nkb = (L + BQ - 1) / BQ;
    const dim3 grid_q(nkb, heads, B), grid_k(nkb + (n_null ? 1 : 0), heads, B);
    constexpr int smem = 6 * tile_bytes<DH>() + 4 * BK * 4;
    const uint32_t thr = p_drop > 0.f ? mha_drop_threshold(p_drop) : 0u;
    const float inv_keep = 1.f / (1.f - p_drop);
    if (p_drop > 0.f) {
        CTK_SET_MAX_SMEM((mha_bwd_dq_kernel<DH, true>), smem);
        CTK_SET_MAX_SMEM((mha_bwd_dkv_kernel<DH, true>), smem);
        mha_bwd_dq_kernel<DH, true><<<grid_q, 128, smem, s>>>(q, key_mask, nk, nv, n_null, d_o, lse, delta, dq, L, heads, scale, thr,
                                                              inv_keep, seed_ptr, seed_off);
        CTK_LAUNCH_CHECK();
        mha_bwd_dkv_kernel<DH, true><<<grid_k, 128, smem, s>>>(q, key_mask, nk, nv, n_null, d_o, lse, delta, dq, dnk, dnv, L, heads,
                                                               scale, thr, inv_keep, seed_ptr, seed_off);
    } else {
        CTK_SET_MAX_SMEM((mha_bwd_dq_kernel<DH, false>), smem);
        CTK_SET_MAX_SMEM((mha_bwd_dkv_kernel<DH, false>), smem);
        mha_bwd_dq_kernel<DH, false><<<grid_q, 128, smem, s>>>(q, key_mask, nk, nv, n_null, d_o, lse, delta, dq, L, heads, scale, thr,
                                                               inv_keep, seed_ptr, seed_off);
        CTK_LAUNCH_CHECK();
        mha_bwd_dkv_kernel<DH, false><<<grid_k, 128, smem, s>>>(q, key_mask, nk, nv, n_null, d_o, lse, delta, dq, dnk, dnv, L, heads,
                                                                scale, thr, inv_keep, seed_ptr, seed_off);
    }
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

}  // namespace

extern "C" int ctk_mha_fwd(const void* qkv, const unsigned char* key_mask, const void* null_k, const void* null_v, int n_null,
                           void* out, float* lse, int B, int L, int heads, int dh, float scale, float p_drop,
                           const unsigned long long* seed_ptr, unsigned long long seed_off, void* stream) {
    int rc = ctk_check_device();
    if (rc != CTK_OK) return rc;
    if ((rc = check_args(B, L, heads, dh, p_drop, seed_ptr, null_k, null_v, n_null))) return rc;
    CTK_REQUIRE(qkv && out && lse && CTK_ALIGNED(qkv, 16) && CTK_ALIGNED(out, 16), CTK_ERR_ALIGN, "mha_fwd: null / unaligned pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
    const __nv_bfloat16* nk = reinterpret_cast<const __nv_bfloat16*>(null_k);
    const __nv_bfloat16* nv = reinterpret_cast<const __nv_bfloat16*>(null_v);
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    if (dh == 64) return launch_fwd<64>(q, key_mask, nk, nv, n_null, o, lse, B, L, heads, scale, p_drop, seed_ptr, seed_off, s);
    return launch_fwd<32>(q, key_mask, nk, nv, n_null, o, lse, B, L, heads, scale, p_drop, seed_ptr, seed_off, s);
}

extern "C" int ctk_mha_bwd(const void* qkv, const unsigned char* key_mask, const void* null_k, const void* null_v, int n_null,
                           const void* out, const void* dout, const float* lse, float* delta, void* dqkv, float* dnull_k,
                           float* dnull_v, int B, int L, int heads, int dh, float scale, float p_drop,
                           const unsigned long long* seed_ptr, unsigned long long seed_off, void* stream) {
    int rc = ctk_check_device();
    if (rc != CTK_OK) return rc;
    if ((rc = check_args(B, L, heads, dh, p_drop, seed_ptr, null_k, null_v, n_null))) return rc;
    CTK_REQUIRE(qkv && out && dout && lse && delta && dqkv && CTK_ALIGNED(qkv, 16) && CTK_ALIGNED(out, 16) &&
                CTK_ALIGNED(dout, 16) && CTK_ALIGNED(dqkv, 16), CTK_ERR_ALIGN, "mha_bwd: null / unaligned pointer");
    CTK_REQUIRE(n_null == 0 || (dnull_k && dnull_v && CTK_ALIGNED(dnull_k, 8) && CTK_ALIGNED(dnull_v, 8)), CTK_ERR_SHAPE,
                "mha_bwd: null pairs need their gradient buffers");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
    const __nv_bfloat16* nk = reinterpret_cast<const __nv_bfloat16*>(null_k);
    const __nv_bfloat16* nv = reinterpret_cast<const __nv_bfloat16*>(null_v);
    const __nv_bfloat16* o = reinterpret_cast<const __nv_bfloat16*>(out);
    const __nv_bfloat16* d_o = reinterpret_cast<const __nv_bfloat16*>(dout);
    __nv_bfloat16* dq = reinterpret_cast<__nv_bfloat16*>(dqkv);
    if (dh == 64)
        return launch_bwd<64>(q, key_mask, nk, nv, n_null, o, d_o, lse, delta, dq, dnull_k, dnull_v, B, L, heads, scale, p_drop,
                              seed_ptr, seed_off, s);
    return launch_bwd<32>(q, key_mask, nk, nv, n_null, o, d_o, lse, delta, dq, dnull_k, dnull_v, B, L, heads, scale, p_drop,
                          seed_ptr, seed_off, s);
}

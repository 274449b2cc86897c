// ctk_gemm_bf16: persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] (fp32, TMEM) = A[M,K] * B[N,K]^T      A, B bf16 in HBM, staged by TMA (128B swizzle)
//
// Operand majors: "K-major" = K contiguous in memory (row-major [rows,K]); "MN-major" = the M/N
// index contiguous (row-major [K,rows]) - used by the weight-gradient products dW = dY^T X whose
// contraction runs over tokens. Both are fed straight to the tensor core through UMMA shared
// memory descriptors; nothing is transposed in HBM.
//
// Roles per CTA (384 threads, 1 CTA / SM, grid = #SMs, static round-robin tile schedule):
//   warp 0    : TMA producer (one elected lane)      3-stage smem ring, full/empty mbarriers
//   warp 1    : MMA issuer   (one elected lane)      tcgen05.mma 128x256x16, 2 TMEM accumulators
//   warp 2    : TMEM allocator (512 columns)
//   warps 4-11: epilogue. Warp e owns TMEM lanes 32*(e%4).. and tile columns 128*(e/4)..: it pulls 32
//               columns at a time with tcgen05.ld, applies the fused epilogue, writes its 32x32 block
//               into a swizzled shared-memory staging slot and hands it to the TMA engine
//               (cp.async.bulk.tensor store) - global writes are full-line bursts and edge clipping is
//               done by the tensor map. Epilogue inputs (residual stream, GEGLU pre-activations) come
//               in the same way through TMA loads.
//
// Fused epilogues replace the elementwise passes the reference runs as separate ATen kernels
// (attention.py:45-58 GEGLU, :158-162 qk l2norm+scale, :443-450 residual adds).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "clip_epilogue_math.cuh"

namespace {

constexpr int BM = 128;          // UMMA M (cta_group::1: TMEM lane == output row)
constexpr int BN = 256;          // UMMA N
constexpr int BK = 64;           // 64 bf16 = 128 B = one swizzle row
constexpr int UK = 16;           // UMMA K for 16-bit inputs
constexpr int A_BYTES = BM * BK * 2;   // 16 KB
// single-CTA tiles: B = 256 rows (32 KB), 3 stages. CTA pairs (cta_group::2): each CTA stages its own
// 128 rows of A and HALF of the 256-row B tile (16 KB) -> 4 stages and 1/3 less L2->SM traffic.
template <bool PAIR> struct Cfg {
    static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = PAIR ? 4 : 3;
};
constexpr int NEPI = 8;                // epilogue warps
constexpr int SLOT_BYTES = 2048;       // one 32x32 bf16 block; an fp32 block takes two slots
constexpr int SLOTS_PER_WARP = 4;
constexpr int STAGING_BYTES = NEPI * SLOTS_PER_WARP * SLOT_BYTES;   // 64 KB
constexpr int SMEM_BYTES = 3 * (A_BYTES + BN * BK * 2) + STAGING_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
static_assert(4 * (A_BYTES + (BN / 2) * BK * 2) <= 3 * (A_BYTES + BN * BK * 2), "pair ring must fit the same budget");
constexpr int NTHREADS = 32 * (4 + NEPI);

struct EpiParams {
    void* C;
    long long ldc;
    const float* bias;
    const float* resid;
    long long ldr;
    void* aux0;
    long long ld_aux0;
    const float* vec0;
    const float* vec1;
    const int* row_map;
    float alpha;
    int i0, i1;
};

__device__ __forceinline__ void ld_acc(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    tc_ld_32x32(taddr, r);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- staging slots: 32 rows, swizzled exactly like the TMA tensor maps expect -------------------
// bf16 block: 64-byte rows, CU_TENSOR_MAP_SWIZZLE_64B  (16-byte chunk ^= (row >> 1) & 3)
// fp32 block: 128-byte rows, CU_TENSOR_MAP_SWIZZLE_128B (16-byte chunk ^= row & 7)
__device__ __forceinline__ void slot_write_bf16(uint8_t* slot, int row, const float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
        u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
        u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
        u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        *reinterpret_cast<uint4*>(slot + row * 64 + ((j ^ ((row >> 1) & 3)) << 4)) = u;
    }
}
__device__ __forceinline__ void slot_read_bf16(const uint8_t* slot, int row, float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint4 u = *reinterpret_cast<const uint4*>(slot + row * 64 + ((j ^ ((row >> 1) & 3)) << 4));
        float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        v[8 * j + 0] = a.x; v[8 * j + 1] = a.y; v[8 * j + 2] = b.x; v[8 * j + 3] = b.y;
        v[8 * j + 4] = c.x; v[8 * j + 5] = c.y; v[8 * j + 6] = d.x; v[8 * j + 7] = d.y;
    }
}
__device__ __forceinline__ void slot_write_f32(uint8_t* slot, int row, const float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(slot + row * 128 + ((j ^ (row & 7)) << 4)) =
            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void slot_read_f32(const uint8_t* slot, int row, float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 f = *reinterpret_cast<const float4*>(slot + row * 128 + ((j ^ (row & 7)) << 4));
        v[4 * j] = f.x; v[4 * j + 1] = f.y; v[4 * j + 2] = f.z; v[4 * j + 3] = f.w;
    }
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Per-warp staging context. All lanes call the methods; lane 0 talks to the TMA engine.
struct Stager {
    uint8_t* base;        // SLOTS_PER_WARP * SLOT_BYTES, 1024-byte aligned
    uint64_t* bar;        // this warp's two mbarriers for TMA loads (double-buffered prefetch)
    uint32_t phase[2];
    int lane;
    // make the slots reusable: all earlier stores have finished reading shared memory
    __device__ __forceinline__ void begin() {
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
    }
    // start TMA loads of 1 or 2 blocks (`bytes_each` each) into slot_a / slot_b, tracked by barrier `which`
    __device__ __forceinline__ void issue_load(int which, const CUtensorMap* m, int slot_a, int ca, int slot_b, int cb,
                                               int row0, int bytes_each, int nblocks) {
        if (lane == 0) {
            mbar_expect_tx(&bar[which], bytes_each * nblocks);
            tma_load_2d(base + slot_a * SLOT_BYTES, m, &bar[which], ca, row0);
            if (nblocks > 1) tma_load_2d(base + slot_b * SLOT_BYTES, m, &bar[which], cb, row0);
        }
    }
    __device__ __forceinline__ void wait_load(int which) {
        mbar_wait(&bar[which], phase[which]);
        phase[which] ^= 1;
    }
    // publish the generic-proxy writes of the whole warp, then store slot -> global
    __device__ __forceinline__ void fence() {
        fence_async_smem();
        __syncwarp();
    }
    __device__ __forceinline__ void store(const CUtensorMap* m, int slot, int c0, int row0) {
        if (lane == 0) tma_store_2d(m, base + slot * SLOT_BYTES, c0, row0);
    }
    __device__ __forceinline__ void commit() {
        if (lane == 0) bulk_commit();
    }
};

// Epilogue of one epilogue warp for one 128x256 accumulator. `t_row` = TMEM address of (this warp's
// lane quarter, accumulator column 0); `row0` = first global row of the warp's 32 rows; `hf` = which
// half of the tile's columns this warp owns; n0 = first global column of the tile.
template <int EPI>
__device__ __forceinline__ void run_epilogue(const EpiParams& p, const CUtensorMap* mc0, const CUtensorMap* mc1,
                                             Stager& sg, uint32_t t_row, int row0, int hf, int n0, int M, int N) {
    const int lane = sg.lane;
    const long long row = (long long)row0 + lane;
    const bool row_ok = row < M;
    if constexpr (EPI == CTK_EPI_BF16 || EPI == CTK_EPI_QKV) {
        // bf16 out through mc0; QKV: per-head l2norm + scale on the first i0 columns, column offset i1
        float* rn = reinterpret_cast<float*>(p.aux0);
#pragma unroll 1
        for (int cc = 0; cc < 128; cc += 64) {
            sg.begin();
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int c = hf * 128 + cc + h2 * 32;
                const int col = n0 + c;
                if (col >= N) break;                     // warp-uniform
                float v[32];
                ld_acc(t_row + c, v);
                if constexpr (EPI == CTK_EPI_BF16) {
                    if (p.bias) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + col + i);
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
                } else {
                    if (col < p.i0) {
                        float ss = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; ++i) ss += v[i] * v[i];
                        const float rnorm = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = v[i] * rnorm * __ldg(p.vec0 + i) * p.alpha;
                        if (row_ok) rn[row * p.ld_aux0 + (col + p.i1) / 32] = rnorm;
                    }
                }
                slot_write_bf16(sg.base + h2 * SLOT_BYTES, lane, v);
                sg.fence();
                sg.store(mc0, h2, col + (EPI == CTK_EPI_QKV ? p.i1 : 0), row0);
            }
            sg.commit();
        }
    } else if constexpr (EPI == CTK_EPI_F32 || EPI == CTK_EPI_RESID_F32) {
        // four 32-column fp32 blocks per warp; block i lives in slots 2*(i&1), 2*(i&1)+1 (4 KB) and the
        // residual block i+1 is prefetched by TMA while block i is processed
        constexpr bool RES = EPI == CTK_EPI_RESID_F32;
        const int cbase = n0 + hf * 128;
        int nvalid = (N - cbase + 31) / 32;
        nvalid = nvalid < 0 ? 0 : (nvalid > 4 ? 4 : nvalid);
        if (nvalid > 0) {
            sg.begin();
            if (RES) sg.issue_load(0, mc1, 0, cbase, 0, 0, row0, 4096, 1);
#pragma unroll 1
            for (int i = 0; i < nvalid; ++i) {
                const int par = i & 1;
                const int col = cbase + i * 32;
                uint8_t* slot = sg.base + par * 2 * SLOT_BYTES;
                if (i + 1 < nvalid) {
                    if (i >= 1) sg.begin();                 // block i-1's store has left the other buffer
                    if (RES) sg.issue_load(par ^ 1, mc1, (par ^ 1) * 2, col + 32, 0, 0, row0, 4096, 1);
                } else if (!RES && i >= 2) {
                    sg.begin();
                }
                float v[32];
                ld_acc(t_row + hf * 128 + i * 32, v);
                if (p.bias) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] += __ldg(p.bias + col + q);
                }
                if (RES) {
                    sg.wait_load(par);
                    float r[32];
                    slot_read_f32(slot, lane, r);
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] += r[q];
                }
                slot_write_f32(slot, lane, v);
                sg.fence();
                sg.store(mc0, par * 2, col, row0);
                sg.commit();
            }
        }
    } else if constexpr (EPI == CTK_EPI_GEGLU) {
        // tile columns [0,128) = value rows, [128,256) = gate rows of 128 hidden units (weights were
        // interleaved by ctk_pack_ff_w1). U (mc0) keeps the pre-activations for the backward pass,
        // H (mc1) = gelu(gate) * value (attention.py:45-48). This warp: units hf*64 .. hf*64+63.
        const int tile = n0 / BN;
#pragma unroll 1
        for (int cc = 0; cc < 64; cc += 32) {
            sg.begin();
            const int c = hf * 64 + cc;
            float val[32], gate[32];
            ld_acc(t_row + c, val);
            ld_acc(t_row + 128 + c, gate);
            slot_write_bf16(sg.base, lane, val);
            slot_write_bf16(sg.base + SLOT_BYTES, lane, gate);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {                    // packed fp32: two hidden units per instruction
                const float2 g2 = make_float2(gate[i], gate[i + 1]);
                float2 cdf, pdf;
                normal_cdf_pdf2(g2, cdf, pdf);
                const float2 h2 = f2_mul(f2_mul(g2, cdf), make_float2(val[i], val[i + 1]));   // gelu(gate) * value
                val[i] = h2.x;
                val[i + 1] = h2.y;
            }
            slot_write_bf16(sg.base + 2 * SLOT_BYTES, lane, val);
            sg.fence();
            sg.store(mc0, 0, n0 + c, row0);
            sg.store(mc0, 1, n0 + 128 + c, row0);
            sg.store(mc1, 2, tile * 128 + c, row0);
            sg.commit();
        }
    } else if constexpr (EPI == CTK_EPI_GEGLU_BWD) {
        // accumulator = dH for hidden units [n0, n0+256); mc1 = U (value|gate interleaved per 128
        // units); mc0 = dU in the same interleaved layout. Chunk i (32 units) uses slots 2*(i&1),
        // 2*(i&1)+1; the U blocks of chunk i+1 are prefetched while chunk i is processed.
        const int ubase = n0 + hf * 128;
        int nvalid = (N - ubase + 31) / 32;
        nvalid = nvalid < 0 ? 0 : (nvalid > 4 ? 4 : nvalid);
        auto ucol_of = [&](int i) { const int unit = ubase + i * 32; return (unit / 128) * 256 + (unit % 128); };
        if (nvalid > 0) {
            sg.begin();
            sg.issue_load(0, mc1, 0, ucol_of(0), 1, ucol_of(0) + 128, row0, SLOT_BYTES, 2);
#pragma unroll 1
            for (int i = 0; i < nvalid; ++i) {
                const int par = i & 1;
                const int ucol = ucol_of(i);
                if (i + 1 < nvalid) {
                    if (i >= 1) sg.begin();                 // chunk i-1's stores have left the other buffer
                    sg.issue_load(par ^ 1, mc1, (par ^ 1) * 2, ucol_of(i + 1), (par ^ 1) * 2 + 1, ucol_of(i + 1) + 128,
                                  row0, SLOT_BYTES, 2);
                }
                float dh[32], val[32], gate[32];
                ld_acc(t_row + hf * 128 + i * 32, dh);
                sg.wait_load(par);
                uint8_t* sv = sg.base + (par * 2) * SLOT_BYTES;
                uint8_t* sgt = sv + SLOT_BYTES;
                slot_read_bf16(sv, lane, val);
                slot_read_bf16(sgt, lane, gate);
#pragma unroll
                for (int q = 0; q < 32; q += 2) {               // packed fp32: two hidden units per instruction
                    const float2 g2 = make_float2(gate[q], gate[q + 1]), v2 = make_float2(val[q], val[q + 1]);
                    const float2 d2 = make_float2(dh[q], dh[q + 1]);
                    float2 cdf, pdf;
                    normal_cdf_pdf2(g2, cdf, pdf);
                    const float2 dv = f2_mul(f2_mul(d2, g2), cdf);                   // dh * gelu(gate)
                    const float2 dg = f2_mul(f2_mul(d2, v2), f2_fma(g2, pdf, cdf));  // dh * value * gelu'(gate)
                    val[q] = dv.x; val[q + 1] = dv.y;
                    gate[q] = dg.x; gate[q + 1] = dg.y;
                }
                slot_write_bf16(sv, lane, val);
                slot_write_bf16(sgt, lane, gate);
                sg.fence();
                sg.store(mc0, par * 2, ucol, row0);
                sg.store(mc0, par * 2 + 1, ucol + 128, row0);
                sg.commit();
            }
        }
    } else if constexpr (EPI == CTK_EPI_GELU) {
        // U (mc0) = acc + bias, kept in bf16 for the backward pass; G (mc1) = gelu(U), exact erf form
        // (HF BertIntermediate, hidden_act "gelu").  Block h2 of a pass uses slots 2*h2 (U) and 2*h2+1 (G).
#pragma unroll 1
        for (int cc = 0; cc < 128; cc += 64) {
            sg.begin();
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int c = hf * 128 + cc + h2 * 32;
                const int col = n0 + c;
                if (col >= N) break;                     // warp-uniform
                float v[32];
                ld_acc(t_row + c, v);
                if (p.bias) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + col + i);
                }
                uint8_t* su = sg.base + (2 * h2) * SLOT_BYTES;
                slot_write_bf16(su, lane, v);
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float2 u2 = make_float2(v[i], v[i + 1]);
                    float2 cdf, pdf;
                    normal_cdf_pdf2(u2, cdf, pdf);
                    const float2 g2 = f2_mul(u2, cdf);
                    v[i] = g2.x;
                    v[i + 1] = g2.y;
                }
                slot_write_bf16(su + SLOT_BYTES, lane, v);
                sg.fence();
                sg.store(mc0, 2 * h2, col, row0);
                sg.store(mc1, 2 * h2 + 1, col, row0);
            }
            sg.commit();
        }
    } else if constexpr (EPI == CTK_EPI_GELU_BWD) {
        // accumulator = dG; mc1 = U (bf16 pre-activations); mc0 = dU = dG * gelu'(U).  Chunk i (32 columns)
        // lives in slot 2*(i&1); the U block of chunk i+1 is prefetched while chunk i is processed.
        const int cbase = n0 + hf * 128;
        int nvalid = (N - cbase + 31) / 32;
        nvalid = nvalid < 0 ? 0 : (nvalid > 4 ? 4 : nvalid);
        if (nvalid > 0) {
            sg.begin();
            sg.issue_load(0, mc1, 0, cbase, 0, 0, row0, SLOT_BYTES, 1);
#pragma unroll 1
            for (int i = 0; i < nvalid; ++i) {
                const int par = i & 1;
                const int col = cbase + i * 32;
                if (i + 1 < nvalid) {
                    if (i >= 1) sg.begin();                 // chunk i-1's store has left the other buffer
                    sg.issue_load(par ^ 1, mc1, (par ^ 1) * 2, col + 32, 0, 0, row0, SLOT_BYTES, 1);
                }
                float dg[32], u[32];
                ld_acc(t_row + hf * 128 + i * 32, dg);
                sg.wait_load(par);
                uint8_t* slot = sg.base + (par * 2) * SLOT_BYTES;
                slot_read_bf16(slot, lane, u);
#pragma unroll
                for (int q = 0; q < 32; q += 2) {
                    const float2 u2 = make_float2(u[q], u[q + 1]);
                    float2 cdf, pdf;
                    normal_cdf_pdf2(u2, cdf, pdf);
                    const float2 r2 = f2_mul(make_float2(dg[q], dg[q + 1]), f2_fma(u2, pdf, cdf));   // gelu'(u) = Phi(u) + u phi(u)
                    dg[q] = r2.x;
                    dg[q + 1] = r2.y;
                }
                slot_write_bf16(slot, lane, dg);
                sg.fence();
                sg.store(mc0, par * 2, col, row0);
                sg.commit();
            }
        }
    } else if constexpr (EPI == CTK_EPI_LSE_PART) {
        // Contrastive logits x = exp(*vec1) * acc (ct_clip.py:1347).  This warp reduces its 32 rows over its
        // (up to) 128 columns to the online statistics (max, sum e^(x-max), sum x e^(x-max)) and writes them to
        // part[(column block, row)][3]; column block = (n0 + hf*128) / 128.  The diagonal logit (row + i0 == col)
        // goes to aux0[row] when aux0 != NULL.  Nothing else leaves the SM.
        const int cbase = n0 + hf * 128;
        if (cbase < N) {
            const float scale = expf(__ldg(p.vec1));
            float* diag = reinterpret_cast<float*>(p.aux0);
            const long long drow = row + p.i0;
            clipepi::RowStat st = {-INFINITY, 0.f, 0.f};
#pragma unroll 1
            for (int cc = 0; cc < 128; cc += 32) {
                const int col = cbase + cc;
                if (col >= N) break;                     // warp-uniform
                float v[32];
                ld_acc(t_row + hf * 128 + cc, v);
                clipepi::lse_chunk(st, v, scale, col, N, drow, (diag && row_ok) ? diag + row : nullptr);
            }
            const float m = st.m, l = st.l, w = st.w;
            if (row_ok) {
                float* dst = reinterpret_cast<float*>(p.C) + ((long long)(cbase / 128) * M + row) * 3;
                dst[0] = m; dst[1] = l; dst[2] = w;
            }
        }
    } else if constexpr (EPI == CTK_EPI_CLIP_GRAD) {
        // dloss/dS of the symmetric InfoNCE for x = exp(*vec1) * acc:
        //   g = alpha * exp(*vec1) * (e^(x - vec0[row]) + e^(x - bias[col]) - 2 [row + i0 == col + i1])
        // (the logit scale is folded in so that the latent gradients are plain products with g), written as a
        // bf16 pair hi (mc0) + lo (mc1), hi + lo = g to 16 mantissa bits.
        const float scale = expf(__ldg(p.vec1));
        const float la = row_ok ? __ldg(p.vec0 + row) : 0.f;
        const float gs = p.alpha * scale;
        const long long drow = row + p.i0;
#pragma unroll 1
        for (int cc = 0; cc < 128; cc += 64) {
            sg.begin();
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int c = hf * 128 + cc + h2 * 32;
                const int col = n0 + c;
                if (col >= N) break;                     // warp-uniform
                float v[32], lo[32];
                ld_acc(t_row + c, v);
                clipepi::clip_grad_chunk(v, lo, scale, gs, la, p.bias + col, drow, col, p.i1);
                uint8_t* sh = sg.base + (2 * h2) * SLOT_BYTES;
                slot_write_bf16(sh, lane, v);
                slot_write_bf16(sh + SLOT_BYTES, lane, lo);
                sg.fence();
                sg.store(mc0, 2 * h2, col, row0);
                sg.store(mc1, 2 * h2 + 1, col, row0);
            }
            sg.commit();
        }
    } else if constexpr (EPI == CTK_EPI_ATOMIC_F32) {
        float* C = reinterpret_cast<float*>(p.C);
        const long long orow = row_ok ? (p.row_map ? (long long)p.row_map[row] : row) : -1;
#pragma unroll 1
        for (int cc = 0; cc < 128; cc += 32) {
            const int c = hf * 128 + cc;
            const int col = n0 + c;
            if (col >= N) break;
            float v[32];
            ld_acc(t_row + c, v);
            if (orow < 0) continue;
            float* dst = C + orow * p.ldc + col;
            if ((p.ldc & 3) == 0 && col + 32 <= N) {
                // 16-byte vector reductions: 8 instead of 32 atomics per thread and chunk
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i),
                                 "f"(v[i] * p.alpha), "f"(v[i + 1] * p.alpha), "f"(v[i + 2] * p.alpha),
                                 "f"(v[i + 3] * p.alpha)
                                 : "memory");
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (col + i < N) atomicAdd(dst + i, v[i] * p.alpha);
            }
        }
    } else if constexpr (EPI == CTK_EPI_ARGMAX) {
        // per-row running arg-max over all N columns: 64-bit atomicMax of
        // (orderable(value) << 32) | (0xffffffff - column)  -> ties resolve to the lowest column.
        unsigned long long* best = reinterpret_cast<unsigned long long*>(p.C);
        float bv = -INFINITY;
        int bi = 0;
#pragma unroll 1
        for (int cc = 0; cc < 128; cc += 32) {
            const int c = hf * 128 + cc;
            const int col = n0 + c;
            if (col >= N) break;
            float v[32];
            ld_acc(t_row + c, v);
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (col + i < N && v[i] > bv) { bv = v[i]; bi = col + i; }
        }
        if (row_ok && bv > -INFINITY) {
            uint32_t u = __float_as_uint(bv);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
            const unsigned long long key =
                (static_cast<unsigned long long>(u) << 32) | (0xffffffffu - (uint32_t)bi);
            atomicMax(best + row, key);
        }
    } else if constexpr (EPI == CTK_EPI_ARGMAX_PART) {
        // best (value, column) and second-best value of this warp's 32 rows over its block of 128 columns:
        // key[blk][row], second[blk][row] with blk = column / 128.  No atomics; ctk_vq_select merges the blocks.
        // Branch-free: four independent (best, index, second) trackers over the columns = 0..3 mod 4 (short
        // dependency chains), merged at the end; equal values keep the lower column and count as "second == best".
        const int cbase = n0 + hf * 128;
        if (cbase < N) {
            float b1[4], b2[4];
            int bi[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { b1[q] = -INFINITY; b2[q] = -INFINITY; bi[q] = cbase + q; }
#pragma unroll 1
            for (int cc = 0; cc < 128; cc += 32) {
                const int col = cbase + cc;
                if (col >= N) break;
                float v[32];
                ld_acc(t_row + hf * 128 + cc, v);
                if (col + 32 > N) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (col + i >= N) v[i] = -INFINITY;
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int q = i & 3;
                    const bool gt = v[i] > b1[q];
                    b2[q] = fmaxf(b2[q], fminf(v[i], b1[q]));
                    b1[q] = gt ? v[i] : b1[q];
                    bi[q] = gt ? col + i : bi[q];
                }
            }
            auto merge = [](float& a1, int& ai, float& a2, float c1, int ci, float c2) {
                const bool a_wins = a1 > c1 || (a1 == c1 && ai < ci);
                const float lose = a_wins ? c1 : a1;
                a2 = fmaxf(fmaxf(a2, c2), lose);
                a1 = a_wins ? a1 : c1;
                ai = a_wins ? ai : ci;
            };
            merge(b1[0], bi[0], b2[0], b1[1], bi[1], b2[1]);
            merge(b1[2], bi[2], b2[2], b1[3], bi[3], b2[3]);
            merge(b1[0], bi[0], b2[0], b1[2], bi[2], b2[2]);
            if (row_ok) {
                uint32_t u = __float_as_uint(b1[0]);
                u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
                const long long slot = (long long)(cbase / 128) * M + row;
                reinterpret_cast<unsigned long long*>(p.C)[slot] =
                    (static_cast<unsigned long long>(u) << 32) | (0xffffffffu - (uint32_t)bi[0]);
                reinterpret_cast<float*>(p.aux0)[slot] = b2[0];
            }
        }
    }
}

template <int EPI, bool A_MN, bool B_MN, bool PAIR>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_c0, const __grid_constant__ CUtensorMap tmap_c1,
            int M, int N, int K, int splits, EpiParams ep) {
    constexpr int STAGES = Cfg<PAIR>::STAGES;
    constexpr int STAGE_BYTES = Cfg<PAIR>::STAGE_BYTES;
    constexpr int BN_LOCAL = PAIR ? BN / 2 : BN;                 // B rows staged by this CTA
    constexpr int ROWS_PER_ITEM = PAIR ? 2 * BM : BM;            // output rows per work item
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* staging = smem + 3 * (A_BYTES + BN * BK * 2);                // 1024-byte aligned
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + STAGING_BYTES);
    uint64_t* full_bar = bars;                    // [4]  TMA -> MMA        (pair: leader's are used)
    uint64_t* empty_bar = bars + 4;               // [4]  MMA -> TMA        (pair: multicast commit)
    uint64_t* tfull_bar = bars + 8;               // [2]  MMA -> epilogue   (pair: multicast commit)
    uint64_t* tempty_bar = bars + 10;             // [2]  epilogue -> MMA   (pair: leader's, both CTAs arrive)
    uint64_t* epi_bar = bars + 12;                // [2 * NEPI] TMA loads of the epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12 + 2 * NEPI);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;          // 0 = leader of the pair
    const int unit = PAIR ? blockIdx.x >> 1 : blockIdx.x;        // scheduling unit (CTA or CTA pair)
    const int nunits = PAIR ? gridDim.x >> 1 : gridDim.x;

    const int m_tiles = (M + ROWS_PER_ITEM - 1) / ROWS_PER_ITEM;
    const int n_tiles = (N + BN - 1) / BN;
    const long long mn_tiles = (long long)m_tiles * n_tiles;
    const int kb_total = (K + BK - 1) / BK;
    const int kb_per = (kb_total + splits - 1) / splits;
    const long long total_work = mn_tiles * splits;
    // work item w -> tile = w % mn_tiles (n fastest), split = w / mn_tiles: units that run together
    // share the k-range (operands hit L2) and cover different output tiles.

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        tma_prefetch_desc(&tmap_c0);
        tma_prefetch_desc(&tmap_c1);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], PAIR ? 2 * NEPI : NEPI);
        }
        for (int i = 0; i < 2 * NEPI; ++i) mbar_init(&epi_bar[i], 1);
        mbar_fence_init();
    }
    if (warp == 2) {
        if constexpr (PAIR) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
        else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs of a pair load their own halves) ==========
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long w = unit; w < total_work; w += nunits) {
                const int split = (int)(w / mn_tiles);
                const long long t = w % mn_tiles;
                const int n_blk = (int)(t % n_tiles), m_blk = (int)(t / n_tiles);
                const int m0 = m_blk * ROWS_PER_ITEM + (int)rank * BM;
                const int n0 = n_blk * BN + (int)rank * BN_LOCAL;
                const int kb0 = split * kb_per;
                const int kb1 = min(kb0 + kb_per, kb_total);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    if constexpr (PAIR) {
                        // bytes of both CTAs land on the leader's barrier
                        const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                        if constexpr (!A_MN) {
                            tma_load_2d_pair(sa, &tmap_a, fb, kb * BK, m0);
                        } else {
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j)
                                tma_load_2d_pair(sa + j * (BK * 128), &tmap_a, fb, m0 + j * 64, kb * BK);
                        }
                        if constexpr (!B_MN) {
                            tma_load_2d_pair(sb, &tmap_b, fb, kb * BK, n0);
                        } else {
#pragma unroll
                            for (int j = 0; j < BN_LOCAL / 64; ++j)
                                tma_load_2d_pair(sb + j * (BK * 128), &tmap_b, fb, n0 + j * 64, kb * BK);
                        }
                    } else {
                        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                        if constexpr (!A_MN) {
                            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m0);
                        } else {
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j)
                                tma_load_2d(sa + j * (BK * 128), &tmap_a, &full_bar[stage], m0 + j * 64, kb * BK);
                        }
                        if constexpr (!B_MN) {
                            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n0);
                        } else {
#pragma unroll
                            for (int j = 0; j < BN_LOCAL / 64; ++j)
                                tma_load_2d(sb + j * (BK * 128), &tmap_b, &full_bar[stage], n0 + j * 64, kb * BK);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only for a pair) =====================
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * BM : BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (long long w = unit; w < total_work; w += nunits) {
                const int split = (int)(w / mn_tiles);
                const int kb0 = split * kb_per;
                const int kb1 = min(kb0 + kb_per, kb_total);
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t sb = sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k) {
                        // K-major: advance 16 elements (32 B) inside the 128 B swizzle row.
                        // MN-major: advance 16 k-rows of 128 B; chunks of 64 m/n are LBO apart.
                        const uint64_t adesc =
                            A_MN ? umma_desc_sw128(sa + k * (UK * 128), BK * 128, 1024)
                                 : umma_desc_sw128(sa + k * (UK * 2), 16, 1024);
                        const uint64_t bdesc =
                            B_MN ? umma_desc_sw128(sb + k * (UK * 128), BK * 128, 1024)
                                 : umma_desc_sw128(sb + k * (UK * 2), 16, 1024);
                        if constexpr (PAIR) tc_mma_f16_pair(tacc, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        else tc_mma_f16(tacc, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    // frees the smem slot (in both CTAs of a pair) when the MMAs retire
                    if constexpr (PAIR) tc_commit_pair_mc(&empty_bar[stage], 3); else tc_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                // accumulator complete -> epilogue warps (of both CTAs)
                if constexpr (PAIR) tc_commit_pair_mc(&tfull_bar[acc], 3); else tc_commit(&tfull_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4;
        const int quarter = ew & 3;                     // == warp % 4 -> TMEM lane quarter
        const int hf = ew >> 2;                         // column half of the tile
        Stager sg;
        sg.base = staging + ew * (SLOTS_PER_WARP * SLOT_BYTES);
        sg.bar = &epi_bar[2 * ew];
        sg.phase[0] = sg.phase[1] = 0;
        sg.lane = lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (long long w = unit; w < total_work; w += nunits) {
            const long long t = w % mn_tiles;
            const int n_blk = (int)(t % n_tiles), m_blk = (int)(t / n_tiles);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16);
            run_epilogue<EPI>(ep, &tmap_c0, &tmap_c1, sg, t_row,
                              m_blk * ROWS_PER_ITEM + (int)rank * BM + quarter * 32, hf, n_blk * BN, M, N);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
                else mbar_arrive(&tempty_bar[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) bulk_wait_all();                 // all TMA stores of this warp have landed
    }

    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D tensor map: dim0 = contiguous extent, dim1 = rows with `ld` elements pitch.
int make_tmap(CUtensorMap* m, const void* ptr, bool f32, long long dim0, long long dim1, long long ld,
              int box0, int box1, CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = get_encode_fn();
    CTK_REQUIRE(fn != nullptr, CTK_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    const int esz = f32 ? 4 : 2;
    cuuint64_t dims[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
    cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
    cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                    const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CTK_REQUIRE(r == CUDA_SUCCESS, CTK_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d): ptr %p dims %lld x %lld ld %lld box %d x %d", (int)r, ptr,
                dim0, dim1, ld, box0, box1);
    return CTK_OK;
}
// 32x32 epilogue block maps
int make_block_tmap(CUtensorMap* m, const void* ptr, bool f32, long long cols, long long rows, long long ld) {
    CTK_REQUIRE(ptr && CTK_ALIGNED(ptr, 16) && (ld * (f32 ? 4 : 2)) % 16 == 0, CTK_ERR_ALIGN,
                "gemm: epilogue tensor needs a 16-byte aligned base and pitch");
    return make_tmap(m, ptr, f32, cols, rows, ld, 32, 32, f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

template <int EPI, bool A_MN, bool B_MN, bool PAIR>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc0, const CUtensorMap& tc1, int M,
           int N, int K, int splits, const EpiParams& ep, cudaStream_t stream) {
    auto kern = gemm_kernel<EPI, A_MN, B_MN, PAIR>;
    CTK_SET_MAX_SMEM(kern, SMEM_BYTES);
    const int rows = PAIR ? 2 * BM : BM;
    const long long work = (long long)((M + rows - 1) / rows) * ((N + BN - 1) / BN) * splits;
    const int sms = ctk_gemm_sms();
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    if (PAIR) {
        const long long units = work < sms / 2 ? work : sms / 2;
        cfg.gridDim = dim3((unsigned)(2 * units));
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    } else {
        cfg.gridDim = dim3((unsigned)(work < sms ? work : sms));
    }
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    CTK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tc0, tc1, M, N, K, splits, ep));
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

// split-K factor for the atomic (weight gradient) epilogue: fill whole waves of SMs
int pick_splits(long long tiles, int kb_total, int sms) {
    int max_splits = kb_total / 4;
    if (max_splits < 1) max_splits = 1;
    if (max_splits > 64) max_splits = 64;
    int best = 1;
    double best_score = -1.0;
    for (int sp = 1; sp <= max_splits; ++sp) {
        const long long items = tiles * sp;
        const long long waves = (items + sms - 1) / sms;
        double eff = (double)items / (double)(waves * sms);          // wave quantisation
        if (items < sms) eff *= 0.999;                                // prefer filling the machine
        const double score = eff - 0.002 * sp;                        // mild bias against extra atomics
        if (score > best_score) { best_score = score; best = sp; }
    }
    return best;
}

}  // namespace

extern "C" int ctk_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B,
                             long long ldb, int b_mn_major, int M, int N, int K, int epilogue,
                             const ctk_gemm_epilogue_t* e, int split_k, void* stream_) {
    int rc = ctk_check_device();
    if (rc != CTK_OK) return rc;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CTK_REQUIRE(A && B && e && e->C, CTK_ERR_SHAPE, "gemm: null pointer");
    CTK_REQUIRE(M > 0 && N > 0 && K > 0, CTK_ERR_SHAPE, "gemm: bad shape %d %d %d", M, N, K);
    CTK_REQUIRE(CTK_ALIGNED(A, 16) && CTK_ALIGNED(B, 16) && lda % 8 == 0 && ldb % 8 == 0,
                CTK_ERR_ALIGN, "gemm: operands need 16-byte aligned base and pitch");
    CTK_REQUIRE(a_mn_major == b_mn_major || (!a_mn_major && b_mn_major), CTK_ERR_SHAPE,
                "gemm: an MN-major A with a K-major B is not instantiated");
    // a K-major A with an MN-major B = the input-gradient products dX = dY W with W [out, in] used as stored
    // (no transposed weight copy); instantiated for the epilogues those products use
    const bool mixed = !a_mn_major && b_mn_major;
    if (mixed)
        CTK_REQUIRE(epilogue == CTK_EPI_BF16 || epilogue == CTK_EPI_RESID_F32 || epilogue == CTK_EPI_GEGLU_BWD ||
                    epilogue == CTK_EPI_GELU_BWD, CTK_ERR_SHAPE, "gemm: epilogue %d has no MN-major-B instantiation", epilogue);

    EpiParams ep;
    ep.C = e->C; ep.ldc = e->ldc; ep.bias = e->bias; ep.resid = e->resid; ep.ldr = e->ldr;
    ep.aux0 = e->aux0; ep.ld_aux0 = e->ld_aux0; ep.vec0 = e->vec0; ep.vec1 = e->vec1;
    ep.row_map = e->row_map; ep.alpha = e->alpha; ep.i0 = e->i0; ep.i1 = e->i1;

    const int kb_total = (K + BK - 1) / BK;
    int splits = split_k;
    if (epilogue != CTK_EPI_ATOMIC_F32) {
        splits = 1;
    } else if (splits <= 0) {
        const int rows_item = (M > BM) ? 2 * BM : BM;     // CTA pairs own 256 rows
        splits = pick_splits((long long)((M + rows_item - 1) / rows_item) * ((N + BN - 1) / BN), kb_total,
                             M > BM ? ctk_gemm_sms() / 2 : ctk_gemm_sms());
    }
    if (splits > kb_total) splits = kb_total;
    {   // every split must own at least one k-block
        int per = (kb_total + splits - 1) / splits;
        splits = (kb_total + per - 1) / per;
    }
    if (epilogue == CTK_EPI_BF16 || epilogue == CTK_EPI_F32 || epilogue == CTK_EPI_RESID_F32 ||
        epilogue == CTK_EPI_QKV || epilogue == CTK_EPI_GEGLU_BWD || epilogue == CTK_EPI_GELU ||
        epilogue == CTK_EPI_GELU_BWD || epilogue == CTK_EPI_CLIP_GRAD)
        CTK_REQUIRE(N % 32 == 0, CTK_ERR_SHAPE, "gemm: N %% 32 != 0 for a block epilogue");
    if (epilogue == CTK_EPI_GELU || epilogue == CTK_EPI_GELU_BWD)
        CTK_REQUIRE(e->aux0 && !a_mn_major, CTK_ERR_SHAPE, "gemm: GELU epilogues need the aux0 buffer and a K-major A");
    if (epilogue == CTK_EPI_LSE_PART)
        CTK_REQUIRE(e->vec1 && !a_mn_major, CTK_ERR_SHAPE, "gemm: LSE_PART needs the log-scale pointer and K-major operands");
    if (epilogue == CTK_EPI_CLIP_GRAD)
        CTK_REQUIRE(e->vec0 && e->vec1 && e->bias && e->aux0 && !a_mn_major, CTK_ERR_SHAPE,
                    "gemm: CLIP_GRAD needs both lse vectors, the log-scale pointer, the lo buffer and K-major operands");
    if (epilogue == CTK_EPI_ARGMAX_PART)
        CTK_REQUIRE(e->aux0 && !a_mn_major, CTK_ERR_SHAPE, "gemm: ARGMAX_PART needs the second-best buffer and K-major operands");
    if (epilogue == CTK_EPI_GEGLU)
        CTK_REQUIRE(N % BN == 0 && e->aux0, CTK_ERR_SHAPE, "gemm: GEGLU needs N %% 256 == 0 and an H buffer");
    if (epilogue == CTK_EPI_GEGLU_BWD)
        CTK_REQUIRE(N % 128 == 0 && e->aux0, CTK_ERR_SHAPE, "gemm: GEGLU_BWD needs N %% 128 == 0");
    if (epilogue == CTK_EPI_QKV)
        CTK_REQUIRE(e->aux0 && e->vec0 && e->i0 % 32 == 0 && e->i1 % 32 == 0 && e->i0 <= N,
                    CTK_ERR_SHAPE, "gemm: QKV epilogue needs rnorm buffer, scale vector, 32-aligned i0/i1");

    // CTA pairs (cta_group::2) unless disabled (CTK_GEMM_PAIR=0) or the problem is a single tile row
    static int pair_env = -1;
    if (pair_env < 0) {
        const char* v = getenv("CTK_GEMM_PAIR");
        pair_env = (v && v[0] == '0') ? 0 : 1;
    }
    const bool pair = pair_env == 1 && M > BM;
    CUtensorMap ta, tb, tc0, tc1;
    if (!a_mn_major) rc = make_tmap(&ta, A, false, K, M, lda, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B);   // [M rows][K]
    else rc = make_tmap(&ta, A, false, M, K, lda, 64, BK, CU_TENSOR_MAP_SWIZZLE_128B);                // [K rows][M]
    if (rc) return rc;
    if (!b_mn_major) rc = make_tmap(&tb, B, false, K, N, ldb, BK, pair ? BN / 2 : BN, CU_TENSOR_MAP_SWIZZLE_128B);   // [N rows][K]
    else rc = make_tmap(&tb, B, false, N, K, ldb, 64, BK, CU_TENSOR_MAP_SWIZZLE_128B);                              // [K rows][N]
    if (rc) return rc;
    // epilogue tensor maps (32x32 blocks). Unused ones alias the A map so the kernel always gets
    // valid descriptors.
    tc0 = ta;
    tc1 = ta;
    switch (epilogue) {
        case CTK_EPI_BF16:
            rc = make_block_tmap(&tc0, e->C, false, N, M, e->ldc);
            break;
        case CTK_EPI_QKV:
            rc = make_block_tmap(&tc0, e->C, false, (long long)N + e->i1, M, e->ldc);
            break;
        case CTK_EPI_F32:
            rc = make_block_tmap(&tc0, e->C, true, N, M, e->ldc);
            break;
        case CTK_EPI_RESID_F32:
            CTK_REQUIRE(e->resid, CTK_ERR_SHAPE, "gemm: residual pointer missing");
            rc = make_block_tmap(&tc0, e->C, true, N, M, e->ldc);
            if (!rc) rc = make_block_tmap(&tc1, e->resid, true, N, M, e->ldr);
            break;
        case CTK_EPI_GEGLU:
            rc = make_block_tmap(&tc0, e->C, false, N, M, e->ldc);
            if (!rc) rc = make_block_tmap(&tc1, e->aux0, false, N / 2, M, e->ld_aux0);
            break;
        case CTK_EPI_GEGLU_BWD:
            rc = make_block_tmap(&tc0, e->C, false, 2LL * N, M, e->ldc);
            if (!rc) rc = make_block_tmap(&tc1, e->aux0, false, 2LL * N, M, e->ld_aux0);
            break;
        case CTK_EPI_GELU:
        case CTK_EPI_GELU_BWD:
        case CTK_EPI_CLIP_GRAD:
            rc = make_block_tmap(&tc0, e->C, false, N, M, e->ldc);
            if (!rc) rc = make_block_tmap(&tc1, e->aux0, false, N, M, e->ld_aux0);
            break;
        default:
            break;
    }
    if (rc) return rc;

#define CTK_GEMM_CASE_MIXED(E)                                                                             \
    case E:                                                                                                \
        if (pair) return launch<E, false, true, true>(ta, tb, tc0, tc1, M, N, K, splits, ep, stream);      \
        return launch<E, false, true, false>(ta, tb, tc0, tc1, M, N, K, splits, ep, stream);
    if (mixed) {
        switch (epilogue) {
            CTK_GEMM_CASE_MIXED(CTK_EPI_BF16)
            CTK_GEMM_CASE_MIXED(CTK_EPI_RESID_F32)
            CTK_GEMM_CASE_MIXED(CTK_EPI_GEGLU_BWD)
            CTK_GEMM_CASE_MIXED(CTK_EPI_GELU_BWD)
            default:
                break;
        }
    }
#undef CTK_GEMM_CASE_MIXED
#define CTK_GEMM_CASE(E)                                                                                   \
    case E:                                                                                                \
        if (pair)                                                                                          \
            return a_mn_major ? launch<E, true, true, true>(ta, tb, tc0, tc1, M, N, K, splits, ep, stream)   \
                              : launch<E, false, false, true>(ta, tb, tc0, tc1, M, N, K, splits, ep, stream); \
        return a_mn_major ? launch<E, true, true, false>(ta, tb, tc0, tc1, M, N, K, splits, ep, stream)      \
                          : launch<E, false, false, false>(ta, tb, tc0, tc1, M, N, K, splits, ep, stream);
#define CTK_GEMM_CASE_KMAJOR(E)                                                                            \
    case E:                                                                                                \
        if (pair) return launch<E, false, false, true>(ta, tb, tc0, tc1, M, N, K, splits, ep, stream);     \
        return launch<E, false, false, false>(ta, tb, tc0, tc1, M, N, K, splits, ep, stream);
    switch (epilogue) {
        CTK_GEMM_CASE_KMAJOR(CTK_EPI_GELU)
        CTK_GEMM_CASE_KMAJOR(CTK_EPI_GELU_BWD)
        CTK_GEMM_CASE_KMAJOR(CTK_EPI_LSE_PART)
        CTK_GEMM_CASE_KMAJOR(CTK_EPI_CLIP_GRAD)
        CTK_GEMM_CASE_KMAJOR(CTK_EPI_ARGMAX_PART)
        CTK_GEMM_CASE(CTK_EPI_BF16)
        CTK_GEMM_CASE(CTK_EPI_F32)
        CTK_GEMM_CASE(CTK_EPI_RESID_F32)
        CTK_GEMM_CASE(CTK_EPI_GEGLU)
        CTK_GEMM_CASE(CTK_EPI_GEGLU_BWD)
        CTK_GEMM_CASE(CTK_EPI_QKV)
        CTK_GEMM_CASE(CTK_EPI_ATOMIC_F32)
        CTK_GEMM_CASE(CTK_EPI_ARGMAX)
        default:
            break;
    }
#undef CTK_GEMM_CASE
#undef CTK_GEMM_CASE_KMAJOR
    ctk_set_error("gemm: unknown epilogue %d", epilogue);
    return CTK_ERR_SHAPE;
}

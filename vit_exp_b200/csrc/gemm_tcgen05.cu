// ctk_gemm_bf16: persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] (fp32, TMEM) = A[M,K] * B[N,K]^T      A, B bf16 in HBM, staged by TMA (128B swizzle)
//
// Operand majors: "K-major" = K contiguous in memory (row-major [rows,K]); "MN-major" = the M/N
// index contiguous (row-major [K,rows]) - used by the weight-gradient products dW = dY^T X whose
// contraction runs over tokens. Both are fed straight to the tensor core through UMMA shared
// memory descriptors; nothing is transposed in HBM.
//
// Roles per CTA (256 threads, 1 CTA / SM, grid = #SMs, static round-robin tile schedule):
//   warp 0   : TMA producer (one elected lane)      4-stage smem ring, full/empty mbarriers
//   warp 1   : MMA issuer   (one elected lane)      tcgen05.mma 128x256x16, 2 TMEM accumulators
//   warp 2   : TMEM allocator (512 columns)
//   warps 4-7: epilogue - tcgen05.ld 32 columns at a time, fused epilogue, direct global stores
//
// Fused epilogues replace the elementwise passes the reference runs as separate ATen kernels
// (attention.py:45-58 GEGLU, :158-162 qk l2norm+scale, :443-450 residual adds).
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace {

constexpr int BM = 128;          // UMMA M (cta_group::1: TMEM lane == output row)
constexpr int BN = 256;          // UMMA N
constexpr int BK = 64;           // 64 bf16 = 128 B = one swizzle row
constexpr int UK = 16;           // UMMA K for 16-bit inputs
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_BYTES = BN * BK * 2;   // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int NTHREADS = 256;

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

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32]) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
        u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
        u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
        u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
        d4[i] = u;
    }
}
__device__ __forceinline__ void load_bf16x32(const __nv_bfloat16* src, float (&v)[32]) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 u = s4[i];
        float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z),
               d = unpack_bf16x2(u.w);
        v[8 * i + 0] = a.x; v[8 * i + 1] = a.y; v[8 * i + 2] = b.x; v[8 * i + 3] = b.y;
        v[8 * i + 4] = c.x; v[8 * i + 5] = c.y; v[8 * i + 6] = d.x; v[8 * i + 7] = d.y;
    }
}

__device__ __forceinline__ void ld_acc(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    tc_ld_32x32(taddr, r);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Epilogue for one 128x256 accumulator: `row` is this thread's global output row, `t_row` the
// TMEM address of (its lane, column 0 of the accumulator), `n0` the tile's first output column.
template <int EPI>
__device__ __forceinline__ void run_epilogue(const EpiParams& p, uint32_t t_row, long long row,
                                             int n0, int M, int N) {
    const bool row_ok = row < M;
    if constexpr (EPI == CTK_EPI_BF16 || EPI == CTK_EPI_F32 || EPI == CTK_EPI_RESID_F32) {
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            const int col = n0 + c;
            if (col >= N) break;                       // warp-uniform
            float v[32];
            ld_acc(t_row + c, v);
            if (!row_ok) continue;
            if (p.bias) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + col + i);
            }
            if constexpr (EPI == CTK_EPI_BF16) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
                store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.ldc + col, v);
            } else {
                float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) +
                                                        row * p.ldc + col);
                if constexpr (EPI == CTK_EPI_RESID_F32) {
                    const float4* rs = reinterpret_cast<const float4*>(p.resid + row * p.ldr + col);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float4 r = rs[i];
                        dst[i] = make_float4(v[4 * i] + r.x, v[4 * i + 1] + r.y,
                                             v[4 * i + 2] + r.z, v[4 * i + 3] + r.w);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
        }
    } else if constexpr (EPI == CTK_EPI_GEGLU) {
        // tile columns [0,128) = value rows, [128,256) = gate rows of 128 hidden units
        // (weights were interleaved by ctk_pack_ff_w1). U (C) keeps the pre-activations for
        // the backward pass, H (aux0) = gelu(gate) * value (attention.py:45-48).
        const int tile = n0 / BN;
        __nv_bfloat16* U = reinterpret_cast<__nv_bfloat16*>(p.C);
        __nv_bfloat16* H = reinterpret_cast<__nv_bfloat16*>(p.aux0);
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
            float val[32], gate[32];
            ld_acc(t_row + c, val);
            ld_acc(t_row + 128 + c, gate);
            if (!row_ok) continue;
            store_bf16x32(U + row * p.ldc + n0 + c, val);
            store_bf16x32(U + row * p.ldc + n0 + 128 + c, gate);
#pragma unroll
            for (int i = 0; i < 32; ++i) val[i] = gelu_erf(gate[i]) * val[i];
            store_bf16x32(H + row * p.ld_aux0 + tile * 128 + c, val);
        }
    } else if constexpr (EPI == CTK_EPI_GEGLU_BWD) {
        // accumulator = dH for hidden units [n0, n0+256); aux0 = U (value|gate interleaved per
        // 128 units); C = dU in the same interleaved layout.
        const __nv_bfloat16* U = reinterpret_cast<const __nv_bfloat16*>(p.aux0);
        __nv_bfloat16* dU = reinterpret_cast<__nv_bfloat16*>(p.C);
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            const int unit = n0 + c;
            if (unit >= N) break;
            float dh[32];
            ld_acc(t_row + c, dh);
            if (!row_ok) continue;
            const long long ucol = (long long)(unit / 128) * 256 + (unit % 128);
            float val[32], gate[32];
            load_bf16x32(U + row * p.ld_aux0 + ucol, val);
            load_bf16x32(U + row * p.ld_aux0 + ucol + 128, gate);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float dv = dh[i] * gelu_erf(gate[i]);
                const float dg = dh[i] * val[i] * gelu_erf_grad(gate[i]);
                val[i] = dv;
                gate[i] = dg;
            }
            store_bf16x32(dU + row * p.ldc + ucol, val);
            store_bf16x32(dU + row * p.ldc + ucol + 128, gate);
        }
    } else if constexpr (EPI == CTK_EPI_QKV) {
        // Every 32-column chunk is one head (dim_head 32). The first i0 columns are l2-normalised
        // per head (eps 1e-12) and scaled per channel (attention.py:158-160; the constant logit
        // scale rides in alpha for q); remaining columns (v) pass through. Output lands in the
        // packed [M, 3*inner] buffer at column offset i1.
        __nv_bfloat16* C = reinterpret_cast<__nv_bfloat16*>(p.C);
        float* rn = reinterpret_cast<float*>(p.aux0);
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            const int col = n0 + c;
            if (col >= N) break;
            float v[32];
            ld_acc(t_row + c, v);
            if (!row_ok) continue;
            if (col < p.i0) {
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) ss += v[i] * v[i];
                const float rnorm = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = v[i] * rnorm * __ldg(p.vec0 + i) * p.alpha;
                rn[row * p.ld_aux0 + (col + p.i1) / 32] = rnorm;
            }
            store_bf16x32(C + row * p.ldc + col + p.i1, v);
        }
    } else if constexpr (EPI == CTK_EPI_ATOMIC_F32) {
        float* C = reinterpret_cast<float*>(p.C);
        const long long orow = row_ok ? (p.row_map ? (long long)p.row_map[row] : row) : -1;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            const int col = n0 + c;
            if (col >= N) break;
            float v[32];
            ld_acc(t_row + c, v);
            if (orow < 0) continue;
            float* dst = C + orow * p.ldc + col;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (col + i < N) atomicAdd(dst + i, v[i] * p.alpha);
        }
    } else if constexpr (EPI == CTK_EPI_ARGMAX) {
        // per-row running arg-max over all N columns: 64-bit atomicMax of
        // (orderable(value) << 32) | (0xffffffff - column)  -> ties resolve to the lowest column.
        unsigned long long* best = reinterpret_cast<unsigned long long*>(p.C);
        float bv = -INFINITY;
        int bi = 0;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            const int col = n0 + c;
            if (col >= N) break;
            float v[32];
            ld_acc(t_row + c, v);
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (col + i < N && v[i] > bv) { bv = v[i]; bi = col + i; }
        }
        if (row_ok) {
            uint32_t u = __float_as_uint(bv);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
            const unsigned long long key =
                (static_cast<unsigned long long>(u) << 32) | (0xffffffffu - (uint32_t)bi);
            atomicMax(best + row, key);
        }
    }
}

template <int EPI, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            int M, int N, int K, int splits, EpiParams ep) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full_bar = bars;                    // [STAGES]  TMA -> MMA
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]  MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]       MMA -> epilogue
    uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]       epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int m_tiles = (M + BM - 1) / BM;
    const int n_tiles = (N + BN - 1) / BN;
    const int kb_total = (K + BK - 1) / BK;
    const int kb_per = (kb_total + splits - 1) / splits;
    const long long total_work = (long long)m_tiles * n_tiles * splits;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 4);
        }
        mbar_fence_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long w = blockIdx.x; w < total_work; w += gridDim.x) {
                const int split = (int)(w % splits);
                const long long t = w / splits;
                const int n_blk = (int)(t % n_tiles), m_blk = (int)(t / n_tiles);
                const int kb0 = split * kb_per;
                const int kb1 = min(kb0 + kb_per, kb_total);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                    if constexpr (!A_MN) {
                        tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j)
                            tma_load_2d(sa + j * (BK * 128), &tmap_a, &full_bar[stage],
                                        m_blk * BM + j * 64, kb * BK);
                    }
                    if constexpr (!B_MN) {
                        tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN);
                    } else {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j)
                            tma_load_2d(sb + j * (BK * 128), &tmap_b, &full_bar[stage],
                                        n_blk * BN + j * 64, kb * BK);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (long long w = blockIdx.x; w < total_work; w += gridDim.x) {
                const int split = (int)(w % splits);
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
                        tc_mma_f16(tacc, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(&empty_bar[stage]);      // frees the smem slot when MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tfull_bar[acc]);            // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4;                        // == warp % 4 -> TMEM lane quarter
        int acc = 0;
        uint32_t acc_phase = 0;
        for (long long w = blockIdx.x; w < total_work; w += gridDim.x) {
            const long long t = w / splits;
            const int n_blk = (int)(t % n_tiles), m_blk = (int)(t / n_tiles);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16);
            const long long row = (long long)m_blk * BM + ew * 32 + lane;
            run_epilogue<EPI>(ep, t_row, row, n_blk * BN, M, N);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
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

// 2-D bf16 tensor map: dim0 = contiguous extent, dim1 = rows with `ld` elements pitch.
int make_tmap(CUtensorMap* m, const void* ptr, long long dim0, long long dim1, long long ld,
              int box0, int box1) {
    EncodeTiledFn fn = get_encode_fn();
    CTK_REQUIRE(fn != nullptr, CTK_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CTK_REQUIRE(r == CUDA_SUCCESS, CTK_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CTK_OK;
}

template <int EPI, bool A_MN, bool B_MN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, int splits,
           const EpiParams& ep, cudaStream_t stream) {
    auto kern = gemm_kernel<EPI, A_MN, B_MN>;
    static bool configured = false;
    if (!configured) {
        CTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const long long work = (long long)((M + BM - 1) / BM) * ((N + BN - 1) / BN) * splits;
    const int grid = (int)(work < ctk_num_sms() ? work : ctk_num_sms());
    kern<<<grid, NTHREADS, SMEM_BYTES, stream>>>(ta, tb, M, N, K, splits, ep);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
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
    CTK_REQUIRE(a_mn_major == b_mn_major, CTK_ERR_SHAPE,
                "gemm: mixed operand majors are not instantiated");

    EpiParams ep;
    ep.C = e->C; ep.ldc = e->ldc; ep.bias = e->bias; ep.resid = e->resid; ep.ldr = e->ldr;
    ep.aux0 = e->aux0; ep.ld_aux0 = e->ld_aux0; ep.vec0 = e->vec0; ep.vec1 = e->vec1;
    ep.row_map = e->row_map; ep.alpha = e->alpha; ep.i0 = e->i0; ep.i1 = e->i1;

    const int kb_total = (K + BK - 1) / BK;
    int splits = split_k;
    if (epilogue != CTK_EPI_ATOMIC_F32) {
        splits = 1;
    } else if (splits <= 0) {
        // enough (tile, split) work items to cover the SMs ~2x, each with >= 8 k-blocks
        const long long tiles = (long long)((M + BM - 1) / BM) * ((N + BN - 1) / BN);
        long long want = (2LL * ctk_num_sms() + tiles - 1) / tiles;
        long long cap = kb_total / 8 > 0 ? kb_total / 8 : 1;
        splits = (int)(want < cap ? want : cap);
        if (splits < 1) splits = 1;
    }
    if (splits > kb_total) splits = kb_total;
    {   // every split must own at least one k-block
        int per = (kb_total + splits - 1) / splits;
        splits = (kb_total + per - 1) / per;
    }
    if (epilogue == CTK_EPI_BF16 || epilogue == CTK_EPI_F32 || epilogue == CTK_EPI_RESID_F32 ||
        epilogue == CTK_EPI_QKV || epilogue == CTK_EPI_GEGLU_BWD) {
        CTK_REQUIRE(N % 32 == 0, CTK_ERR_SHAPE, "gemm: N %% 32 != 0 for a vector epilogue");
        CTK_REQUIRE(CTK_ALIGNED(e->C, 16) && e->ldc % 8 == 0, CTK_ERR_ALIGN, "gemm: C alignment");
    }
    if (epilogue == CTK_EPI_RESID_F32)
        CTK_REQUIRE(e->resid && CTK_ALIGNED(e->resid, 16) && e->ldr % 4 == 0, CTK_ERR_ALIGN,
                    "gemm: residual alignment");
    if (epilogue == CTK_EPI_GEGLU)
        CTK_REQUIRE(N % BN == 0 && e->aux0 && e->ldc % 8 == 0 && e->ld_aux0 % 8 == 0,
                    CTK_ERR_SHAPE, "gemm: GEGLU needs N %% 256 == 0 and an H buffer");
    if (epilogue == CTK_EPI_GEGLU_BWD)
        CTK_REQUIRE(N % 128 == 0 && e->aux0, CTK_ERR_SHAPE, "gemm: GEGLU_BWD needs N %% 128 == 0");
    if (epilogue == CTK_EPI_QKV)
        CTK_REQUIRE(e->aux0 && e->vec0 && e->i0 % 32 == 0 && e->i1 % 32 == 0 && e->i0 <= N,
                    CTK_ERR_SHAPE, "gemm: QKV epilogue needs rnorm buffer, scale vector, 32-aligned i0/i1");

    CUtensorMap ta, tb;
    if (!a_mn_major) {
        rc = make_tmap(&ta, A, K, M, lda, BK, BM);            // [M rows][K]
        if (rc) return rc;
        rc = make_tmap(&tb, B, K, N, ldb, BK, BN);            // [N rows][K]
        if (rc) return rc;
    } else {
        rc = make_tmap(&ta, A, M, K, lda, 64, BK);            // [K rows][M]
        if (rc) return rc;
        rc = make_tmap(&tb, B, N, K, ldb, 64, BK);            // [K rows][N]
        if (rc) return rc;
    }

#define CTK_GEMM_CASE(E)                                                                     \
    case E:                                                                                  \
        return a_mn_major ? launch<E, true, true>(ta, tb, M, N, K, splits, ep, stream)       \
                          : launch<E, false, false>(ta, tb, M, N, K, splits, ep, stream);
    switch (epilogue) {
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
    ctk_set_error("gemm: unknown epilogue %d", epilogue);
    return CTK_ERR_SHAPE;
}

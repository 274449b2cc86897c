// Shared pieces of the tcgen05 spatial-attention kernels (attention_tc.cu forward, attention_tc_bwd.cu
// backward): geometry of the 24x24 token grid, shared-memory bias table layout, PTX wrappers.
#pragma once
#include "common.cuh"

namespace attn_tc {


constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

constexpr int TL = 576;                 // tokens per sequence
constexpr int TGW = 24;                 // token grid width (and height)
constexpr int TWW = 2 * TGW - 1;        // 47 distinct offsets per axis
constexpr int TNOFF = TWW * TWW;        // 2209 table entries per head
constexpr int TPW = 56;                 // padded row pitch of the table in shared memory (words)
constexpr int TOFF = (TGW - 1) * TPW + (TGW - 1);
constexpr int QT = 128;                 // queries per tile
constexpr int NQT = (TL + QT - 1) / QT; // 5 (the last one holds 64 queries)
constexpr int CH = 192;                 // keys per TMA box of K / V
constexpr int NCH = TL / CH;            // 3
constexpr int SC = 96;                  // keys per logits chunk
constexpr int NSC = TL / SC;            // 6
constexpr int NSOFT = 8;                // softmax warps
constexpr int NTHR = 32 * (1 + NSOFT);

constexpr int CH_BYTES = CH * 64;       // 12288
constexpr int Q_BYTES = QT * 64;        // 8192
// forward shared-memory map
constexpr int OFF_K = 0;
constexpr int OFF_V = OFF_K + NCH * CH_BYTES;
constexpr int OFF_Q = OFF_V + NCH * CH_BYTES;
constexpr int OFF_T = OFF_Q + 2 * Q_BYTES;
constexpr int OFF_LP = OFF_T + TWW * TPW * 4;
constexpr int OFF_RED = OFF_LP + 2 * QT * 4;
constexpr int OFF_BAR = OFF_RED + 3 * NSOFT * 4;
constexpr int NBAR = 3 + 3 + 2 + 5;
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

__device__ __forceinline__ void tc_st_32x32_x8(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
        : "memory");
}
// one word of shared memory by 32-bit shared-space address (ptxas folds constant offsets into the
// LDS immediate); volatile: ordered against the barriers around the per-head table refill
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}


}  // namespace attn_tc

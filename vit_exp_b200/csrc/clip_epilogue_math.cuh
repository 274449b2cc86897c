// Per-row arithmetic of the contrastive-loss GEMM epilogues (CTK_EPI_LSE_PART, CTK_EPI_CLIP_GRAD in
// gemm_tcgen05.cu), shared with tests/host_checks.cu, which runs it on the CPU over a host-computed accumulator so
// that the online-statistics update, the masking of ragged edges, the diagonal rule and the hi/lo split are checked
// against the oracle without a GPU.  `v` is one thread's 32 consecutive accumulator columns of one row.
#pragma once
#include <cuda_bf16.h>
#include <math.h>

#if defined(__CUDACC__)
#define CTK_EPI_HD __host__ __device__ __forceinline__
#else
#define CTK_EPI_HD inline
#endif

namespace clipepi {

struct RowStat { float m, l, w; };            // running max, sum e^(x-m), sum x e^(x-m)

// Fold logits x = scale * v[i] of columns [col, col+32) ∩ [0, N) into the row's running statistics; the logit with
// col + i == drow (the positive pair) is written to *diag_out when diag_out != nullptr.
CTK_EPI_HD void lse_chunk(RowStat& st, float (&v)[32], float scale, int col, int N, long long drow, float* diag_out) {
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        v[i] = col + i < N ? v[i] * scale : -INFINITY;
        mx = fmaxf(mx, v[i]);
    }
    if (mx > st.m) {                            // rescale the running sums to the new maximum
        const float r = expf(st.m - mx);        // st.m = -inf on the first chunk: r = 0 and l = w = 0 anyway
        st.l *= r;
        st.w *= r;
        st.m = mx;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if (col + i < N) {
            const float e = expf(v[i] - st.m);
            st.l += e;
            st.w = fmaf(e, v[i], st.w);
            if (diag_out && drow == col + i) *diag_out = v[i];
        }
    }
}

// g = gs * (e^(x - la) + e^(x - lb[i]) - 2 [drow == col + i + i1]) for x = scale * v[i]; returns the bf16 pair
// hi (in v) + lo (in lo): hi + lo = g to 16 mantissa bits.  lb points at the column statistics of column `col`.
CTK_EPI_HD void clip_grad_chunk(float (&v)[32], float (&lo)[32], float scale, float gs, float la, const float* lb,
                                long long drow, int col, int i1) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float x = v[i] * scale;
        float g = expf(x - la) + expf(x - lb[i]);
        if (drow == (long long)col + i + i1) g -= 2.f;
        g *= gs;
        const float hi = __bfloat162float(__float2bfloat16_rn(g));
        v[i] = hi;
        lo[i] = g - hi;
    }
}

}  // namespace clipepi

// TEST INFRASTRUCTURE ONLY (built by tests/test_host_checks_cpu.py into tests/_build/, never linked into libctk.so):
// runs the __host__ __device__ per-element code of the CUDA kernels on the CPU so that their arithmetic and
// indexing can be compared with the oracle on a machine without a GPU.
#include "../vit_exp_b200/csrc/volume_prep_math.cuh"

extern "C" void hostcheck_volume_prep(const void* src, int src_is_f16, int D, int H, int W, float* dst, int Dt, int Ht,
                                      int Wt) {
    using namespace volprep;
    const AxisPlan pz = plan_axis(D, Dt), py = plan_axis(H, Ht), px = plan_axis(W, Wt);
    const long long nvec = (long long)Dt * Ht * (Wt / 4);
    for (long long i = 0; i < nvec; ++i) {
        float v[4];
        if (src_is_f16) prep_vec4<true>(i, src, H, W, pz, py, px, Ht, Wt, v);
        else prep_vec4<false>(i, src, H, W, pz, py, px, Ht, Wt, v);
        for (int j = 0; j < 4; ++j) dst[4 * i + j] = v[j];
    }
}

// ---- contrastive-loss GEMM epilogues (gemm_tcgen05.cu: CTK_EPI_LSE_PART, CTK_EPI_CLIP_GRAD) over a host accumulator
#include "../vit_exp_b200/csrc/clip_epilogue_math.cuh"

// acc fp32 [M, N] row-major = the GEMM accumulator; part fp32 [ceil(N/128), M, 3]; diag fp32 [M] or NULL.
// Same loop structure as the epilogue warp: per row and 128-column block, chunks of 32 columns.
extern "C" void hostcheck_lse_part(const float* acc, int M, int N, float log_scale, int i0, float* part, float* diag) {
    const float scale = expf(log_scale);
    for (int row = 0; row < M; ++row)
        for (int cbase = 0; cbase < N; cbase += 128) {
            clipepi::RowStat st = {-INFINITY, 0.f, 0.f};
            for (int cc = 0; cc < 128; cc += 32) {
                const int col = cbase + cc;
                if (col >= N) break;
                float v[32];
                for (int i = 0; i < 32; ++i) v[i] = col + i < N ? acc[(long long)row * N + col + i] : 12345.f;   // garbage beyond N
                clipepi::lse_chunk(st, v, scale, col, N, (long long)row + i0, diag ? diag + row : nullptr);
            }
            float* dst = part + ((long long)(cbase / 128) * M + row) * 3;
            dst[0] = st.m; dst[1] = st.l; dst[2] = st.w;
        }
}

// hi / lo fp32 [M, N] receive the bf16-representable halves of the gradient stripe (N % 32 == 0 as the GEMM requires)
extern "C" void hostcheck_clip_grad(const float* acc, int M, int N, float log_scale, float alpha, const float* vec0,
                                    const float* bias, int i0, int i1, float* hi, float* lo) {
    const float scale = expf(log_scale);
    const float gs = alpha * scale;
    for (int row = 0; row < M; ++row)
        for (int col = 0; col < N; col += 32) {
            float v[32], l[32];
            for (int i = 0; i < 32; ++i) v[i] = acc[(long long)row * N + col + i];
            clipepi::clip_grad_chunk(v, l, scale, gs, vec0[row], bias + col, (long long)row + i0, col, i1);
            for (int i = 0; i < 32; ++i) {
                hi[(long long)row * N + col + i] = v[i];
                lo[(long long)row * N + col + i] = l[i];
            }
        }
}

// ---- attention-probability dropout mask of the text tower (csrc/mha_dropout.cuh, used by attention_mha64.cu) ----
#include "../vit_exp_b200/csrc/mha_dropout.cuh"
// keep[bh][i][j] (uint8) for bh < nbh, i, j < L; seed = mha_seed(base, offset)
extern "C" void hostcheck_mha_keep(unsigned long long base, unsigned long long offset, int nbh, int L, float p,
                                   unsigned char* keep) {
    const unsigned long long seed = mha_seed(base, offset);
    const uint32_t thr = mha_drop_threshold(p);
    for (int bh = 0; bh < nbh; ++bh)
        for (int i = 0; i < L; ++i)
            for (int j = 0; j < L; ++j)
                keep[((long long)bh * L + i) * L + j] = mha_keep(seed, (uint32_t)bh, (uint32_t)i, (uint32_t)j, thr) ? 1 : 0;
}

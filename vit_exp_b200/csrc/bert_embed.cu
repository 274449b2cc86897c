// BertEmbeddings of the text tower (HF BertModel as the reference calls it, CT_CLIP/ct_clip/ct_clip.py:1271):
//   e[m, :] = word[ids[m]] + token_type[tt[m] or 0] + position[m % L]        (fp32; LayerNorm follows in ctk_layernorm_fwd)
// and its backward: scatter of de into the three tables.  HBM-bound row gathers / scatters, one warp per token row.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
bert_embed_fwd_kernel(const long long* __restrict__ ids, const long long* __restrict__ tt, const float* __restrict__ word,
                      const float* __restrict__ pos, const float* __restrict__ typ, float* __restrict__ out,
                      long long M, int L, int H) {
    const int lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (m >= M) return;
    const float4* w = reinterpret_cast<const float4*>(word + ids[m] * H);
    const float4* p = reinterpret_cast<const float4*>(pos + (m % L) * H);
    const float4* t = reinterpret_cast<const float4*>(typ + (tt ? tt[m] : 0) * H);
    float4* o = reinterpret_cast<float4*>(out + m * H);
    for (int c = lane; c < H / 4; c += 32) {
        const float4 a = __ldg(w + c), b = __ldg(t + c), d = __ldg(p + c);
        o[c] = make_float4((a.x + b.x) + d.x, (a.y + b.y) + d.y, (a.z + b.z) + d.z, (a.w + b.w) + d.w);
    }
}

// dword[ids[m]] += de[m] (rows equal to pad_idx are skipped: nn.Embedding never updates its padding row);
// dtyp[tt[m]] += de[m] when token types are given (otherwise the caller takes the column sum of de).
__global__ void __launch_bounds__(256)
bert_embed_scatter_kernel(const float* __restrict__ de, const long long* __restrict__ ids, const long long* __restrict__ tt,
                          float* __restrict__ dword, float* __restrict__ dtyp, long long M, int H, long long pad_idx) {
    const int lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (m >= M) return;
    const long long id = ids[m];
    const float4* g = reinterpret_cast<const float4*>(de + m * H);
    float* dw = id == pad_idx ? nullptr : dword + id * H;
    float* dt = tt ? dtyp + tt[m] * H : nullptr;
    for (int c = lane; c < H / 4; c += 32) {
        const float4 v = __ldg(g + c);
        if (dw) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dw + 4 * c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        if (dt) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dt + 4 * c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
}

// dpos[l, c] = sum_b de[b * L + l, c]   (no atomics: one thread per (l, 4 columns))
__global__ void __launch_bounds__(256)
bert_embed_dpos_kernel(const float* __restrict__ de, float* __restrict__ dpos, int B, int L, int H) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const int h4 = H / 4;
    if (i >= (long long)L * h4) return;
    const long long l = i / h4;
    const int c = (int)(i % h4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(de + ((long long)b * L + l) * H) + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(dpos + l * H)[c] = acc;
}

}  // namespace

extern "C" int ctk_bert_embed_fwd(const long long* ids, const long long* token_type, const float* word, const float* pos,
                                  const float* typ, float* out, long long M, int L, int H, void* stream) {
    int rc = ctk_check_device();
    if (rc != CTK_OK) return rc;
    CTK_REQUIRE(ids && word && pos && typ && out && M > 0 && L > 0 && H > 0 && H % 4 == 0 && M % L == 0, CTK_ERR_SHAPE,
                "bert_embed_fwd: bad args (M %lld, L %d, H %d)", M, L, H);
    bert_embed_fwd_kernel<<<(unsigned)((M + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        ids, token_type, word, pos, typ, out, M, L, H);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

extern "C" int ctk_bert_embed_bwd(const float* de, const long long* ids, const long long* token_type, float* dword,
                                  float* dpos, float* dtyp, long long M, int L, int H, long long pad_idx, void* stream) {
    int rc = ctk_check_device();
    if (rc != CTK_OK) return rc;
    CTK_REQUIRE(de && ids && dword && dpos && M > 0 && L > 0 && H > 0 && H % 4 == 0 && M % L == 0 && (!token_type || dtyp),
                CTK_ERR_SHAPE, "bert_embed_bwd: bad args (M %lld, L %d, H %d)", M, L, H);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    bert_embed_scatter_kernel<<<(unsigned)((M + 7) / 8), 256, 0, s>>>(de, ids, token_type, dword, dtyp, M, H, pad_idx);
    CTK_LAUNCH_CHECK();
    const long long n = (long long)L * (H / 4);
    bert_embed_dpos_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(de, dpos, (int)(M / L), L, H);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

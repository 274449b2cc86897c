// Per-voxel arithmetic and index plan of ctk_volume_prep (scripts/data.py:49-111), shared by the CUDA kernel
// (volume_prep.cu) and by the host-side check tests/host_checks.cu, which runs exactly this code on the CPU against
// the oracle - so the arithmetic and the crop / pad indexing are verified without a GPU; only the launch is not.
#pragma once
#include <cuda_fp16.h>
#include <math.h>

#if defined(__CUDACC__)
#define CTK_HD __host__ __device__ __forceinline__
#else
#define CTK_HD inline
#endif

namespace volprep {

struct AxisPlan { int start, len, pad; };     // source offset, copied length, leading pad (data.py:77-98)

inline AxisPlan plan_axis(int n, int t) {
    AxisPlan p;
    p.start = (n - t) / 2 > 0 ? (n - t) / 2 : 0;
    const int end = p.start + t < n ? p.start + t : n;
    p.len = end - p.start;
    p.pad = (t - p.len) / 2;
    return p;
}

// IEEE single-precision add / multiply without contraction into an FMA (device: explicit intrinsics; host: the
// compiler never contracts across these separate statements at the default -ffp-contract of nvcc's host pass for
// volatile-free code, and (c + 1) * 0.5 is exact in the second step anyway)
CTK_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}
CTK_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}

CTK_HD float prep_f32(float x) {
    if (x != x) return x;                                        // np.clip propagates NaN
    const float c = fminf(fmaxf(x, -1.f), 1.f);
    return mul_rn(add_rn(c, 1.f), 0.5f);
}
// numpy evaluates float16 ufuncs as float32 operations rounded back to float16 after each step
CTK_HD float prep_f16(__half h) {
    const float x = __half2float(h);
    if (x != x) return x;
    const float c = fminf(fmaxf(x, -1.f), 1.f);
    const __half s = __float2half_rn(add_rn(c, 1.f));
    const __half r = __float2half_rn(mul_rn(__half2float(s), 0.5f));
    return __half2float(r);
}

// output vector i (4 consecutive voxels along W) of the (Dt, Ht, Wt) target
template <bool HALF>
CTK_HD void prep_vec4(long long i, const void* src_, int H, int W, AxisPlan pz, AxisPlan py, AxisPlan px, int Ht,
                      int Wt, float (&v)[4]) {
    const int wv = Wt / 4;
    const int xq = (int)(i % wv);
    const long long zy = i / wv;
    const int y = (int)(zy % Ht), z = (int)(zy / Ht);
    const int sz = z - pz.pad, sy = y - py.pad;
    v[0] = v[1] = v[2] = v[3] = -1.f;                            // pad value (data.py:100)
    if (sz >= 0 && sz < pz.len && sy >= 0 && sy < py.len) {
        const long long row = ((long long)(pz.start + sz) * H + (py.start + sy)) * W;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int sx = xq * 4 + j - px.pad;
            if (sx >= 0 && sx < px.len) {
                const long long idx = row + px.start + sx;
                if (HALF) v[j] = prep_f16(reinterpret_cast<const __half*>(src_)[idx]);
                else v[j] = prep_f32(reinterpret_cast<const float*>(src_)[idx]);
            }
        }
    }
}

}  // namespace volprep

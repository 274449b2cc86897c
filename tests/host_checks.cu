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

// Loader-side volume preparation on the GPU: scripts/data.py:49-111 `npz_to_tensor`.
//
//   stored array (D, H, W), float32 or float16   ->   float32 (Dt, Ht, Wt) = (240, 480, 480)
//   v = (clip(x, -1, 1) + 1) / 2 evaluated in the STORED dtype (numpy keeps float16 arithmetic for float16
//   arrays, data.py:59-61), centre crop (data.py:77-85), centre pad with the constant -1 (data.py:87-98).
//   The reference's transpose (data.py:53) and permute (data.py:104) cancel: the output is (D, H, W) ordered.
//
// HBM-bound: every stored voxel is read at most once, every output voxel written once (16-byte stores); the host
// ships the volume in its stored dtype (2 bytes per voxel for the fp16 datasets) instead of the fp32 result, which
// is what bounds the end-to-end step over PCIe.  Algorithmic bytes per volume: Dt*Ht*Wt*(src_bytes + 4).
// The per-voxel code lives in volume_prep_math.cuh so that tests/host_checks.cu can run it on the CPU.
#include "common.cuh"
#include "volume_prep_math.cuh"

namespace {

using volprep::AxisPlan;

template <bool HALF>
__global__ void __launch_bounds__(256)
volume_prep_kernel(const void* __restrict__ src, float* __restrict__ dst, int H, int W, AxisPlan pz, AxisPlan py,
                   AxisPlan px, int Dt, int Ht, int Wt) {
    const long long nvec = (long long)Dt * Ht * (Wt / 4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
         i += (long long)gridDim.x * blockDim.x) {
        float v[4];
        volprep::prep_vec4<HALF>(i, src, H, W, pz, py, px, Ht, Wt, v);
        reinterpret_cast<float4*>(dst)[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

}  // namespace

extern "C" int ctk_volume_prep(const void* src, int src_is_f16, int D, int H, int W, float* dst, int Dt, int Ht,
                               int Wt, void* stream_) {
    int rc = ctk_check_device();
    if (rc) return rc;
    CTK_REQUIRE(src && dst && D > 0 && H > 0 && W > 0 && Dt > 0 && Ht > 0 && Wt > 0, CTK_ERR_SHAPE,
                "volume_prep: bad arguments");
    CTK_REQUIRE(Wt % 4 == 0 && CTK_ALIGNED(dst, 16), CTK_ERR_ALIGN, "volume_prep: output rows need 16-byte alignment");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    const AxisPlan pz = volprep::plan_axis(D, Dt), py = volprep::plan_axis(H, Ht), px = volprep::plan_axis(W, Wt);
    const long long nvec = (long long)Dt * Ht * (Wt / 4);
    long long blocks = (nvec + 255) / 256;
    const long long cap = (long long)ctk_num_sms() * 16;         // grid-stride: a whole number of waves
    if (blocks > cap) blocks = cap;
    if (src_is_f16)
        volume_prep_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(src, dst, H, W, pz, py, px, Dt, Ht, Wt);
    else
        volume_prep_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(src, dst, H, W, pz, py, px, Dt, Ht, Wt);
    CTK_LAUNCH_CHECK();
    return CTK_OK;
}

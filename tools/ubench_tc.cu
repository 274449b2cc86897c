// Micro-benchmark of the tcgen05 / TMEM / mbarrier round trips that bound the attention kernels
// (one CTA, clock64 timestamps).  Build + run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I vit_exp_b200/csrc tools/ubench_tc.cu -o /tmp/ubench_tc && /tmp/ubench_tc
// Reports cycles (SM clock) for
//   * a chain of n dependent tcgen05.mma (same TMEM accumulator) from first issue to the mbarrier the commit arrives on,
//     for the tile shapes the attention kernels use (SS M128 N96 K16, TS M128 N32 K16, SS M128 N192 K16);
//   * tcgen05.ld 32x32b.x32 + wait::ld, tcgen05.st .x16 + wait::st;
//   * an mbarrier ping-pong between two warps (arrive -> try_wait wake-up, both directions).
#include "common.cuh"
#include <cstdio>
#include <vector>
#include <algorithm>

__device__ __forceinline__ uint64_t desc_sw64(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                 "r"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}

constexpr int NREP = 32;

// out[test][rep]
__global__ void __launch_bounds__(160, 1) ubench(long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
        mbar_fence_init();
    }
    if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_slot;
    const uint32_t sA = smem_u32(smem), sB = sA + 16384;
    int test = 0;
    // ---- MMA chains: {SS N96, TS N32, SS N192} x n in {1, 2, 6, 12}
    const int shapes[3][2] = {{96, 0}, {32, 1}, {192, 0}};
    const int chain[4] = {1, 2, 6, 12};
    uint32_t par = 0;
    for (int sh = 0; sh < 3; ++sh)
        for (int ci = 0; ci < 4; ++ci, ++test) {
            const int N = shapes[sh][0], ts = shapes[sh][1], n = chain[ci];
            const uint32_t idesc = umma_idesc_bf16(128, N, 0, ts ? 1 : 0);
            for (int rep = 0; rep < NREP; ++rep) {
                if (threadIdx.x == 0) {
                    const long long t0 = clock64();
                    for (int k = 0; k < n; ++k) {
                        if (ts) mma_ts(tb + 256, tb + (k & 7) * 8, desc_sw64(sB + (k & 7) * 1024, 512, 512), idesc, k > 0);
                        else tc_mma_f16(tb, desc_sw64(sA + (k & 1) * 32, 16, 512), desc_sw64(sB + (k & 1) * 32, 16, 512), idesc, k > 0);
                    }
                    tc_commit(&bar[0]);
                    mbar_wait(&bar[0], par);
                    out[test * NREP + rep] = clock64() - t0;
                }
                par ^= 1;
                __syncthreads();
            }
        }
    // ---- 12 MMAs spread round-robin over 4 independent accumulators (TS N32, then SS N96)
    for (int v = 0; v < 2; ++v, ++test) {
        const int N = v == 0 ? 32 : 96;
        const uint32_t idesc = umma_idesc_bf16(128, N, 0, v == 0 ? 1 : 0);
        for (int rep = 0; rep < NREP; ++rep) {
            if (threadIdx.x == 0) {
                const long long t0 = clock64();
                for (int k = 0; k < 12; ++k) {
                    const uint32_t dcol = 256 + (k & 3) * (v == 0 ? 32 : 0) + (v == 1 ? (k & 3) * 0 : 0);
                    if (v == 0) mma_ts(tb + dcol, tb + (k & 7) * 8, desc_sw64(sB + (k & 7) * 1024, 512, 512), idesc, k > 3);
                    else tc_mma_f16(tb + (k & 3) * 96, desc_sw64(sA + (k & 1) * 32, 16, 512), desc_sw64(sB + (k & 1) * 32, 16, 512), idesc, k > 3);
                }
                tc_commit(&bar[0]);
                mbar_wait(&bar[0], par);
                out[test * NREP + rep] = clock64() - t0;
            }
            par ^= 1;
            __syncthreads();
        }
    }
    // ---- two threads (warps 0 and 2) issue 6 TS N32 MMAs each into their own accumulator at the same time
    {
        const uint32_t idesc = umma_idesc_bf16(128, 32, 0, 1);
        uint32_t p3 = 0;
        for (int rep = 0; rep < NREP; ++rep) {
            if (threadIdx.x == 0 || threadIdx.x == 64) {
                const int w = threadIdx.x == 0 ? 0 : 1;
                const long long t0 = clock64();
                for (int k = 0; k < 6; ++k)
                    mma_ts(tb + 256 + w * 32, tb + w * 128 + (k & 7) * 8, desc_sw64(sB + (k & 7) * 1024, 512, 512), idesc, k > 0);
                tc_commit(&bar[w == 0 ? 0 : 3]);
                mbar_wait(&bar[w == 0 ? 0 : 3], w == 0 ? par : p3);
                if (w == 0) out[test * NREP + rep] = clock64() - t0;
            }
            par ^= 1; p3 ^= 1;
            __syncthreads();
        }
        ++test;
    }
    // ---- TMEM load / store round trips (warp 1)
    for (int rep = 0; rep < NREP; ++rep) {
        if (warp == 1) {
            uint32_t r[32];
            __syncwarp();
            const long long t0 = clock64();
            tc_ld_32x32(tb + (32u << 16), r);
            tc_wait_ld();
            const long long t1 = clock64();
            uint32_t acc = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) acc ^= r[i];
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i] = acc + i;
            const long long t2 = clock64();
            st16(tb + (32u << 16) + 64, w);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            const long long t3 = clock64();
            if (lane == 0) { out[test * NREP + rep] = t1 - t0; out[(test + 1) * NREP + rep] = t3 - t2; }
        }
        __syncthreads();
    }
    test += 2;
    // ---- mbarrier ping-pong: warp 1 lane 0 arrives on bar[1]; thread 0 waits, arrives on bar[2]; warp 1 waits
    uint32_t pp = 0;
    for (int rep = 0; rep < NREP; ++rep) {
        if (warp == 1) {
            const long long t0 = clock64();
            if (lane == 0) mbar_arrive(&bar[1]);
            mbar_wait(&bar[2], pp);
            if (lane == 0) out[test * NREP + rep] = clock64() - t0;
        } else if (threadIdx.x == 0) {
            mbar_wait(&bar[1], pp);
            mbar_arrive(&bar[2]);
        }
        pp ^= 1;
        __syncthreads();
    }
    ++test;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

void ctk_set_error(const char*, ...) {}
void ctk_count_launch() {}

int main() {
    const int ntests = 3 * 4 + 3 + 3;
    long long* d;
    cudaMalloc(&d, ntests * NREP * sizeof(long long));
    cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    ubench<<<1, 160, 64 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<long long> h(ntests * NREP);
    cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    const char* names[] = {"SS M128 N96 K16 x1", "SS M128 N96 K16 x2", "SS M128 N96 K16 x6", "SS M128 N96 K16 x12",
                           "TS M128 N32 K16 x1", "TS M128 N32 K16 x2", "TS M128 N32 K16 x6", "TS M128 N32 K16 x12",
                           "SS M128 N192 K16 x1", "SS M128 N192 K16 x2", "SS M128 N192 K16 x6", "SS M128 N192 K16 x12",
                           "TS M128 N32 K16 x12 over 4 accumulators", "SS M128 N96 K16 x12 over 4 accumulators",
                           "TS M128 N32 K16 x6, two threads issuing concurrently (each)",
                           "tcgen05.ld 32x32b.x32 + wait::ld", "tcgen05.st 32x32b.x16 + wait::st",
                           "mbarrier ping-pong (arrive -> peer wait -> peer arrive -> wait)"};
    printf("%-66s %8s %8s\n", "round trip (issue .. mbarrier / wait), SM cycles", "median", "min");
    for (int t = 0; t < ntests; ++t) {
        std::vector<long long> v(h.begin() + t * NREP + 4, h.begin() + (t + 1) * NREP);    // skip warm-up reps
        std::sort(v.begin(), v.end());
        printf("%-66s %8lld %8lld\n", names[t], v[v.size() / 2], v[0]);
    }
    return 0;
}

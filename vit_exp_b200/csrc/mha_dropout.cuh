// Counter-based dropout mask of the text tower's attention probabilities (attention_mha64.cu): element (report*head,
// query i, key j) is KEPT iff hash(seed, bh, i, j) >= p * 2^32.  __host__ __device__ so that the CPU host checks and the
// tests' torch restatement (tests/emulated_ops.py: mha_keep_mask) reproduce the exact mask the kernels apply.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MHA_HD __host__ __device__ __forceinline__
#else
#define MHA_HD inline
#endif

MHA_HD uint32_t mha_mix32(uint32_t x) {          // "lowbias32" finaliser
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}
MHA_HD uint32_t mha_hash(unsigned long long seed, uint32_t bh, uint32_t i, uint32_t j) {
    uint32_t x = (i * 0x9E3779B1u) ^ (j * 0x85EBCA77u) ^ (bh * 0xC2B2AE3Du) ^ (uint32_t)(seed & 0xFFFFFFFFull);
    x = mha_mix32(x);
    x += (uint32_t)(seed >> 32) + j * 0x27D4EB2Fu;
    return mha_mix32(x);
}
// per-call seed: the device counter the caller keeps (graph replays read it from memory) + a per-layer offset
MHA_HD unsigned long long mha_seed(unsigned long long base, unsigned long long offset) {
    return base + offset * 0x9E3779B97F4A7C15ull;
}
MHA_HD uint32_t mha_drop_threshold(float p) {
    const double t = (double)p * 4294967296.0;
    return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}
MHA_HD bool mha_keep(unsigned long long seed, uint32_t bh, uint32_t i, uint32_t j, uint32_t thresh) {
    return mha_hash(seed, bh, i, j) >= thresh;
}

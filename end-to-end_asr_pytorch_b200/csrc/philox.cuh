// philox.cuh - the counter-based random bits of every dropout in this library (attention probabilities: mha.cu; the
// dropout in front of a residual + LayerNorm: ln.cu).  Forward and backward regenerate the same bits from an element's
// coordinates and the call's seed, so no mask is ever stored.
#pragma once
#include <cstdint>

namespace asr {

// seed of a call = host seed, or (CUDA-graph replays) *seed_dev + the per-call constant handed out while the step was traced
__device__ __forceinline__ void effective_seed(const uint64_t* seed_dev, uint32_t& lo, uint32_t& hi) {
    if (seed_dev != nullptr) {
        const uint64_t s = __ldg(reinterpret_cast<const unsigned long long*>(seed_dev)) + (((uint64_t)hi << 32) | lo);
        lo = (uint32_t)s;
        hi = (uint32_t)(s >> 32);
    }
}

// Philox4x32-7 (counter-based: forward and backward regenerate the same bits from the element's
// coordinates).  One call yields the 16 random bytes of keys [k16*16, k16*16+16) of row q of head bh.
__device__ __forceinline__ uint4 philox16(uint32_t k16, uint32_t q, uint32_t bh, uint32_t seed_lo, uint32_t seed_hi) {
    uint32_t c0 = k16, c1 = q, c2 = bh, c3 = 0x2545F491u;
    uint32_t k0 = seed_lo, k1 = seed_hi;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
// random byte of key (k & 15) out of a philox16 result
__device__ __forceinline__ uint32_t philox_byte(const uint4& r, int k) {
    const uint32_t w = (k & 8) ? ((k & 4) ? r.w : r.z) : ((k & 4) ? r.y : r.x);
    return (w >> ((k & 3) * 8)) & 0xffu;
}

static inline uint32_t drop_threshold(float p_drop) {      // keep an element when its random byte >= threshold (p in 1/256 steps)
    if (!(p_drop > 0.0f)) return 0;
    int t = (int)(p_drop * 256.0f + 0.5f);
    return (uint32_t)(t < 0 ? 0 : (t > 255 ? 255 : t));
}

}  // namespace asr

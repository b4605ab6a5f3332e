// common.cuh - shared helpers for libasr_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/asr_sm100.h"

namespace asr {

// ---- host-side error plumbing ------------------------------------------------
void set_error(const char* fmt, ...);
int get_opt(const char* key);           // returns current option value (0 if unset)
void count_launch(int n = 1);

#define ASR_CHECK_CUDA(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            asr::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

#define ASR_REQUIRE(cond, ...)                                                            \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            asr::set_error(__VA_ARGS__);                                                  \
            return 2;                                                                     \
        }                                                                                 \
    } while (0)

#define ASR_LAUNCH_CHECK()                                                                \
    do {                                                                                  \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            asr::set_error("%s:%d: kernel launch failed -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return 1;                                                                     \
        }                                                                                 \
        asr::count_launch();                                                              \
    } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int num_sms();

// TMA descriptor for a row-major 2-D fp32/bf16 matrix (driver entry point fetched
// at run time: no -lcuda link dependency).
int make_tmap_2d(CUtensorMap* map, CUtensorMapDataType dtype, size_t elem_bytes, const void* base,
                 uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                 uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swizzle);
int make_tmap_nd(CUtensorMap* map, CUtensorMapDataType dtype, size_t elem_bytes, const void* base, int rank,
                 const uint64_t* dims, const uint64_t* strides_bytes /*rank-1*/, const uint32_t* box,
                 CUtensorMapSwizzle swizzle);

// ---- device-side PTX wrappers ------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the
// hint expires) instead of spinning, so waiting warps do not eat the issue slots of working warps.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(200000u)
        : "memory");
    return ok != 0;
}
// Bounded wait: a transfer that never lands (bad descriptor) traps instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 22)) __trap();
    }
}

// One arrival per warp: every lane's earlier writes are ordered before lane 0's arrive by the warp
// barrier.  512 per-thread arrivals on one mbarrier are 512 serialised shared-memory atomics.
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}

// Wait flavours for latency experiments: 0 = suspend-hinted try_wait (mbar_wait), 1 = plain try_wait
// (short hardware time limit), 2 = non-blocking test_wait polling.
__device__ __forceinline__ void mbar_wait_mode(uint64_t* bar, uint32_t parity, int mode) {
    if (mode == 0) {
        mbar_wait(bar, parity);
        return;
    }
    uint32_t ok = 0, spins = 0;
    while (true) {
        if (mode == 1) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        } else {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        }
        if (ok) return;
        if (++spins > (1u << 26)) __trap();
    }
}

// 2-D TMA tile load: global (tensor map) -> shared, completion on an mbarrier.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

#endif  // __CUDACC__

}  // namespace asr

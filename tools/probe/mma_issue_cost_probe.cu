// What does a tcgen05.mma issuer cost the busy warps of its own sub-partition?
// Warp 0 runs the attention exponential loop body (FFMA2 + 2 MUFU + FADD2 + F2FP per pair of scores).  One elected
// thread of warp `iw` issues groups of `g` tcgen05.mma (M128 N128 K16, bf16, garbage operands) + one commit and waits
// for the group before the next one.  iw = 4: same sub-partition as the worker; iw = 5: another one.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t cvt2(float lo, float hi) { uint32_t w; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(hi), "f"(lo)); return w; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() { uint32_t pred; asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred)); return pred != 0; }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFFu);
    d |= (uint64_t)(1) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
constexpr int NP = 16;
__global__ void probe(float* out, long long* cyc, int iw, int g, int groups, int iters, float cc, float mm, int wait_mode) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (w == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    long long issued_cyc = 0;
    if (w == iw && iw != 0) {
        if (elect_one()) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 16384);
            long long t0 = clock64();
            for (int it = 0; it < groups; ++it) {
                for (int k = 0; k < g; ++k) {
                    const uint64_t ad = desc_sw128(a_addr + (k & 3) * 32), bd = desc_sw128(b_addr + (k & 3) * 32);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(k > 0 ? 1u : 0u) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                uint32_t ok = 0;
                while (!ok) {
                    if (wait_mode == 0)
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(it & 1), "r"(200000u) : "memory");
                    else if (wait_mode == 1)
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(it & 1) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(it & 1) : "memory");
                }
            }
            issued_cyc = clock64() - t0;
            cyc[1] = issued_cyc;
        }
    }
    if (w == 0) {
        float x[2 * NP];
#pragma unroll
        for (int k = 0; k < 2 * NP; ++k) x[k] = -(lane * 0.01f + k * 0.1f);
        uint64_t c2 = pack2(cc, cc), m2 = pack2(mm, mm);
        uint64_t acc = pack2(0.f, 0.f);
        uint32_t pk = 0;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                float a = x[2 * k], b = x[2 * k + 1];
                float y0, y1; unpack2(fma2(pack2(a, b), c2, m2), y0, y1); float p0 = ex2f(y0), p1 = ex2f(y1); acc = add2(acc, pack2(p0, p1)); pk ^= cvt2(p0, p1); x[2 * k] = y0; x[2 * k + 1] = y1;
            }
        }
        long long t1 = clock64();
        float s = 0, a0, a1;
        unpack2(acc, a0, a1);
#pragma unroll
        for (int k = 0; k < 2 * NP; ++k) s += x[k];
        out[threadIdx.x] = s + a0 + a1 + __uint_as_float(pk);
        if (lane == 0) cyc[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
    }
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int iters = 1024;   // worker: 1024 * 16 pairs * ~21 cycles = 344 k cycles
    struct { int iw, g, groups; const char* name; int wm; } cases[] = {
        {0, 0, 0, "no issuer", 0},
        {4, 4, 1500, "issuer on the worker's sub-partition, groups of 4", 0},
        {5, 4, 1500, "issuer on another sub-partition, groups of 4", 0},
        {4, 12, 500, "same sub-partition, groups of 12", 0},
        {5, 12, 500, "other sub-partition, groups of 12", 0},
        {4, 1, 4000, "same sub-partition, groups of 1", 0},
        {4, 1, 4000, "groups of 1, plain try_wait", 1},
        {4, 1, 4000, "groups of 1, test_wait spin", 2},
        {5, 1, 4000, "groups of 1, test_wait spin, other sub-partition", 2},
    };
    for (auto& c : cases) {
        long long h[2] = {0, 0};
        for (int rep = 0; rep < 2; ++rep) { cudaMemset(cyc, 0, 64); probe<<<1, 256, 65536>>>(out, cyc, c.iw, c.g, c.groups, iters, 0.18f, -0.5f, c.wm); cudaDeviceSynchronize(); }
        cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
        printf("%-52s worker cycles per pair = %.2f   issuer: %lld cycles for %d MMAs (%.1f per MMA)\n", c.name, (double)h[0] / (iters * (double)NP), h[1], c.g * c.groups, c.g ? (double)h[1] / (c.g * c.groups) : 0.0);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

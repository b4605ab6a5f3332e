// Does a warp that sleeps in mbarrier.try_wait cost issue slots of a busy warp on the same sub-partition?
// Warp 0 runs the attention exponential loop body; warp 4 (same sub-partition, higher warp id = higher issue
// priority) waits on an mbarrier that warp 0 completes when it is done.  Modes of the waiter:
//   0 absent   1 one elected lane, try_wait with a 200 us suspend hint   2 one elected lane, plain try_wait
//   3 whole warp, hinted try_wait   4 one lane, test_wait spin   5 one lane, hinted try_wait, waiter on warp 5 (other sub-partition)
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
constexpr int NP = 16;
__global__ void probe(float* out, long long* cyc, int mode, int iters, float cc, float mm) {
    __shared__ uint64_t bar;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    const int waiter_warp = (mode == 5) ? 5 : 4;
    if (w == waiter_warp && mode != 0) {
        const bool me = (mode == 3) ? true : elect_one();
        if (me) {
            uint32_t ok = 0;
            while (!ok) {
                if (mode == 1 || mode == 3 || mode == 5)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0), "r"(200000u) : "memory");
                else if (mode == 2)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
            }
        }
        return;
    }
    if (w != 0) return;
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
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
    float s = 0, a0, a1;
    unpack2(acc, a0, a1);
#pragma unroll
    for (int k = 0; k < 2 * NP; ++k) s += x[k];
    out[threadIdx.x] = s + a0 + a1 + __uint_as_float(pk);
    if (lane == 0) cyc[0] = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
    const int iters = 2048;
    const char* names[] = {"no waiter", "1 lane, hinted try_wait", "1 lane, plain try_wait", "32 lanes, hinted try_wait", "1 lane, test_wait spin", "1 lane hinted, other sub-partition"};
    for (int mode = 0; mode < 6; ++mode) {
        for (int rep = 0; rep < 2; ++rep) { probe<<<1, 256>>>(out, cyc, mode, iters, 0.18f, -0.5f); cudaDeviceSynchronize(); }
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-36s worker cycles per pair = %.2f\n", names[mode], (double)h / (iters * (double)NP));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

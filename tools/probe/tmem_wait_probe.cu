// Fixed costs of the tensor-memory hand-over instructions for one warp (cycles per iteration, 256 iterations):
//   0: LDTM.x32 + wait::ld          1: 4 x LDTM.x32 + one wait::ld      2: STTM.x16 + wait::st     3: 4 x STTM.x16 + wait::st
//   4: fence::before + __syncwarp + mbarrier.arrive (lane 0)   5: try_wait on a completed phase + fence::after   6: vote.any + branch
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define LD32(taddr, r) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory")
#define ST16(taddr, r) asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" \
    ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory")
template <int KIND>
__global__ void probe(uint32_t* out, long long* cyc, int iters) {
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (w == 0) {
        uint32_t r[4][32];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int i = 0; i < 32; ++i) r[q][i] = lane + i + q;
        uint32_t acc = 0;
        if (KIND == 5 && lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");   // phase 0 complete
        __syncwarp();
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (KIND == 0) { LD32(tmem, r[0]); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += r[0][it & 31]; }
            if (KIND == 1) { LD32(tmem, r[0]); LD32(tmem + 32, r[1]); LD32(tmem + 64, r[2]); LD32(tmem + 96, r[3]); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += r[0][it & 31] + r[3][it & 31]; }
            if (KIND == 2) { ST16(tmem + 128, r[0]); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
            if (KIND == 3) { ST16(tmem + 128, r[0]); ST16(tmem + 144, r[1]); ST16(tmem + 160, r[2]); ST16(tmem + 176, r[3]); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
            if (KIND == 4) { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncwarp(); if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory"); }
            if (KIND == 5) {
                uint32_t ok = 0;
                while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0), "r"(200000u) : "memory");
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            if (KIND == 6) { if (__any_sync(0xffffffffu, (acc + it) == 0x7fffffffu)) acc += 17; acc += it; }
        }
        long long t1 = clock64();
        out[lane] = acc;
        if (lane == 0) cyc[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
    }
}
int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
    const int iters = 256;
    const char* names[] = {"LDTM.x32 + wait::ld", "4 x LDTM.x32 + wait::ld", "STTM.x16 + wait::st", "4 x STTM.x16 + wait::st", "fence::before + syncwarp + mbarrier.arrive", "try_wait (complete) + fence::after", "vote.any + branch"};
    for (int kind = 0; kind < 7; ++kind) {
        for (int rep = 0; rep < 2; ++rep) {
#define GO(K) if (kind == K) probe<K><<<1, 128>>>(out, cyc, iters)
            GO(0); GO(1); GO(2); GO(3); GO(4); GO(5); GO(6);
            cudaDeviceSynchronize();
        }
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-46s %.1f cycles per iteration\n", names[kind], (double)h / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

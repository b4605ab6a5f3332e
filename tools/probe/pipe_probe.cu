// Measures per-warp issue cost (cycles per warp instruction) of MUFU.EX2, SHFL and LDS when 1..8 warps of one CTA
// run the same independent-chain loop.  Answers: is the unit per sub-partition or shared by the SM?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int KIND>
__global__ void probe(float* out, long long* cyc, unsigned active_mask, int iters) {
    __shared__ float sm[8 * 32 * 8];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8 * 32 * 8; i += blockDim.x) sm[i] = i * 1e-3f;
    __syncthreads();
    if (!((active_mask >> w) & 1)) return;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = lane * 0.01f + k;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (KIND == 0) v[k] = ex2f(v[k]);
            if (KIND == 1) v[k] = __shfl_up_sync(0xffffffffu, v[k], 1);
            if (KIND == 2) { v[k] = sm[(w * 8 + k) * 32 + ((lane + __float_as_int(v[k])) & 31)]; }
            if (KIND == 3) { asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(v[k])); }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (lane == 0) cyc[blockIdx.x * 8 + w] = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8 * 8);
    const int iters = 4096;
    const unsigned masks[] = {0x1, 0x3, 0x11, 0x5, 0xF, 0xFF};
    const char* names[] = {"ex2", "shfl", "lds", "lg2"};
    for (int kind = 0; kind < 4; ++kind)
        for (unsigned m : masks) {
            for (int rep = 0; rep < 2; ++rep) {
                if (kind == 0) probe<0><<<1, 256>>>(out, cyc, m, iters);
                if (kind == 1) probe<1><<<1, 256>>>(out, cyc, m, iters);
                if (kind == 2) probe<2><<<1, 256>>>(out, cyc, m, iters);
                if (kind == 3) probe<3><<<1, 256>>>(out, cyc, m, iters);
                cudaDeviceSynchronize();
            }
            long long h[8]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            int w0 = 0; while (!((m >> w0) & 1)) ++w0;
            printf("%s warps_mask=0x%02x cycles/warp-instr=%.2f\n", names[kind], m, (double)h[w0] / (iters * 8.0));
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

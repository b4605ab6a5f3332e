// Which pipe do the instructions of the attention exponential loop use, and what does the loop cost per pair of
// scores when one or two warps of a sub-partition run it?  cycles per loop body (one pair = 2 scores):
//   kind 0: 2 MUFU.EX2            1: F2FP.BF16.F32.PACK_AB      2: 2 MUFU + F2FP     3: FFMA2 + 2 MUFU + FADD2 + F2FP
//   kind 4: FFMA2                 5: FADD2                       6: 2 MUFU + FADD2    7: FMNMX3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t cvt2(float lo, float hi) { uint32_t w; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(hi), "f"(lo)); return w; }
constexpr int NP = 16;   // pairs per loop iteration (independent)
template <int KIND>
__global__ void probe(float* out, long long* cyc, unsigned active_mask, int iters, float cc, float mm) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (!((active_mask >> w) & 1)) return;
    float x[2 * NP];
#pragma unroll
    for (int k = 0; k < 2 * NP; ++k) x[k] = -(lane * 0.01f + k * 0.1f);
    uint64_t c2 = pack2(cc, cc), m2 = pack2(mm, mm);
    uint64_t acc = pack2(0.f, 0.f);
    uint32_t pk = 0;
    float mx = -1e30f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            float a = x[2 * k], b = x[2 * k + 1];
            if (KIND == 0) { x[2 * k] = ex2f(a); x[2 * k + 1] = ex2f(b); }
            if (KIND == 1) { uint32_t wv = cvt2(a, b); x[2 * k] = __uint_as_float(wv); }
            if (KIND == 2) { float p0 = ex2f(a), p1 = ex2f(b); pk ^= cvt2(p0, p1); x[2 * k] = p0; x[2 * k + 1] = p1; }
            if (KIND == 3) { float y0, y1; unpack2(fma2(pack2(a, b), c2, m2), y0, y1); float p0 = ex2f(y0), p1 = ex2f(y1); acc = add2(acc, pack2(p0, p1)); pk ^= cvt2(p0, p1); x[2 * k] = y0; x[2 * k + 1] = y1; }
            if (KIND == 4) { float y0, y1; unpack2(fma2(pack2(a, b), c2, m2), y0, y1); x[2 * k] = y0; x[2 * k + 1] = y1; }
            if (KIND == 5) { acc = add2(acc, pack2(a, b)); }
            if (KIND == 6) { float p0 = ex2f(a), p1 = ex2f(b); acc = add2(acc, pack2(p0, p1)); x[2 * k] = p0; x[2 * k + 1] = p1; }
            if (KIND == 7) { mx = fmaxf(fmaxf(mx, a), b); x[2 * k] = mx; }
        }
    }
    long long t1 = clock64();
    float s = 0, a0, a1;
    unpack2(acc, a0, a1);
#pragma unroll
    for (int k = 0; k < 2 * NP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + a0 + a1 + __uint_as_float(pk) + mx;
    if (lane == 0) cyc[blockIdx.x * 8 + w] = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8 * 8);
    const int iters = 2048;
    const unsigned masks[] = {0x1, 0x11, 0x3, 0xFF};
    const char* names[] = {"2xMUFU", "F2FP", "2xMUFU+F2FP", "FFMA2+2xMUFU+FADD2+F2FP", "FFMA2", "FADD2", "2xMUFU+FADD2", "FMNMX3"};
    for (int kind = 0; kind < 8; ++kind)
        for (unsigned m : masks) {
            for (int rep = 0; rep < 2; ++rep) {
#define GO(K) if (kind == K) probe<K><<<1, 256>>>(out, cyc, m, iters, 0.18f, -0.5f)
                GO(0); GO(1); GO(2); GO(3); GO(4); GO(5); GO(6); GO(7);
                cudaDeviceSynchronize();
            }
            long long h[8]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            printf("%-26s warps_mask=0x%02x cycles per pair per warp = %.2f\n", names[kind], m, (double)h[0] / (iters * (double)NP));
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

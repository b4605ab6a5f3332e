// SpecAugment on the device (second half): /root/reference/src/utils/utils.py:168-194.
//
// The reference fills frequency bands with the per-frame mean and time spans with the per-bin mean of the
// ORIGINAL batch (both means are taken before any mask is applied, utils.py:171-173), mask by mask in a Python
// loop over the batch.  Because every frequency mask writes the same value to a cell (freq_mean[b,t]) and every
// time mask writes the same value (time_mean[b,v]), and all time masks come after all frequency masks, the
// result is independent of the order inside each family:
//     out[b,t,v] = time_mean[b,v]  if t lies in any time span of utterance b
//                  freq_mean[b,t]  else if v lies in any frequency band of utterance b
//                  in[b,t,v]       else
// Pass 1 reads the batch once (row means, column sums per 64-frame chunk, fixed summation order: results are
// bit-reproducible run to run), pass 2 folds the chunk sums, pass 3 touches only the masked cells in place.
// The draws (torch.rand, utils.py:178-181,186-189) stay in torch: the host mirror makes the same calls in the
// same order and hands the integer bands / spans to this library.
#include "common.cuh"

namespace asr {

constexpr int kSaRows = 64;      // frames per CTA of pass 1
constexpr int kSaMaxV = 1024;    // feature bins (80 raw, 320 after 4-frame stacking)

__global__ void __launch_bounds__(256) specaug_means_kernel(const float* __restrict__ x, int T, int V, int nchunk,
                                                            float* __restrict__ freq_mean, float* __restrict__ part) {
    __shared__ float s_col[8][kSaMaxV];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.x, b = blockIdx.y;
    const int nv = (V + 31) >> 5;
    float acc[kSaMaxV / 32];
#pragma unroll
    for (int c = 0; c < kSaMaxV / 32; ++c) acc[c] = 0.0f;
    const float inv_v = 1.0f / (float)V;
    for (int i = 0; i < kSaRows / 8; ++i) {
        const int t = chunk * kSaRows + i * 8 + warp;      // warp-uniform
        if (t >= T) break;
        const float* row = x + ((size_t)b * T + t) * V;
        float rs = 0.0f;
#pragma unroll
        for (int c = 0; c < kSaMaxV / 32; ++c) {
            if (c < nv) {
                const int v = c * 32 + lane;
                const float val = v < V ? __ldg(row + v) : 0.0f;
                rs += val;
                acc[c] += val;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
        if (lane == 0) freq_mean[(size_t)b * T + t] = rs * inv_v;
    }
#pragma unroll
    for (int c = 0; c < kSaMaxV / 32; ++c)
        if (c < nv) s_col[warp][c * 32 + lane] = acc[c];
    __syncthreads();
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += s_col[w][v];
        part[((size_t)b * nchunk + chunk) * V + v] = s;
    }
}

// Same pass for V % 4 == 0 and 16-byte aligned rows: float4 loads, four rows of a warp in flight at once
// (NV4 = float4 per lane and row = ceil(V / 128)): 4 * NV4 independent 16-byte loads per lane.
template <int NV4>
__global__ void __launch_bounds__(256) specaug_means_vec_kernel(const float* __restrict__ x, int T, int V, int nchunk,
                                                                float* __restrict__ freq_mean, float* __restrict__ part) {
    __shared__ float4 s_col[8][kSaMaxV / 4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.x, b = blockIdx.y;
    const int v4n = V >> 2;
    float4 acc[NV4];
#pragma unroll
    for (int c = 0; c < NV4; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float inv_v = 1.0f / (float)V;
    // warp w owns rows chunk*64 + w*8 .. +7 (consecutive rows: one 8-row slab per warp)
    const int t_base = chunk * kSaRows + warp * 8;
#pragma unroll
    for (int i0 = 0; i0 < 8; i0 += 4) {
        float4 val[4][NV4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int t = t_base + i0 + i;
            const float4* row = reinterpret_cast<const float4*>(x + ((size_t)b * T + min(t, T - 1)) * V);
#pragma unroll
            for (int c = 0; c < NV4; ++c) {
                const int v4 = c * 32 + lane;
                val[i][c] = (t < T && v4 < v4n) ? __ldg(row + v4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float rs = 0.0f;
#pragma unroll
            for (int c = 0; c < NV4; ++c) {
                rs += (val[i][c].x + val[i][c].y) + (val[i][c].z + val[i][c].w);
                acc[c].x += val[i][c].x; acc[c].y += val[i][c].y; acc[c].z += val[i][c].z; acc[c].w += val[i][c].w;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
            const int t = t_base + i0 + i;
            if (lane == 0 && t < T) freq_mean[(size_t)b * T + t] = rs * inv_v;
        }
    }
#pragma unroll
    for (int c = 0; c < NV4; ++c)
        if (c * 32 + lane < v4n) s_col[warp][c * 32 + lane] = acc[c];
    __syncthreads();
    for (int v4 = threadIdx.x; v4 < v4n; v4 += blockDim.x) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const float4 a = s_col[w][v4];
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
        reinterpret_cast<float4*>(part + ((size_t)b * nchunk + chunk) * V)[v4] = s;
    }
}

__global__ void __launch_bounds__(256) specaug_time_mean_kernel(const float* __restrict__ part, const int* __restrict__ len,
                                                                int V, int nchunk, float* __restrict__ time_mean) {
    const int b = blockIdx.y;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    float s = 0.0f;
    for (int c = 0; c < nchunk; ++c) s += part[((size_t)b * nchunk + c) * V + v];
    time_mean[(size_t)b * V + v] = s / (float)__ldg(len + b);     // utils.py:172-173 (a zero length divides by zero there too)
}

// blockIdx.x = mask (R frequency bands, then R time spans), blockIdx.y = utterance, blockIdx.z = slab of frames
constexpr int kSaMaxR = 16;
constexpr int kSaSlabs = 8;
__global__ void __launch_bounds__(256) specaug_apply_kernel(float* __restrict__ x, const int* __restrict__ f0, const int* __restrict__ fw,
                                                            const int* __restrict__ t0, const int* __restrict__ tw, int R, int B,
                                                            int T, int V, const float* __restrict__ freq_mean,
                                                            const float* __restrict__ time_mean) {
    __shared__ int s_a[kSaMaxR], s_b[kSaMaxR];       // the utterance's time spans [a, b)
    const int b = blockIdx.y;
    const int m = blockIdx.x;
    const int z = blockIdx.z, nz = gridDim.z;
    float* xb = x + (size_t)b * T * V;
    if (m < R) {
        if (threadIdx.x < R) {
            const int a0 = __ldg(t0 + threadIdx.x * B + b);
            s_a[threadIdx.x] = a0;
            s_b[threadIdx.x] = a0 + __ldg(tw + threadIdx.x * B + b);
        }
        __syncthreads();
        const int lo = max(__ldg(f0 + m * B + b), 0);
        const int hi = min(__ldg(f0 + m * B + b) + __ldg(fw + m * B + b), V);
        const int w = hi - lo;
        if (w <= 0) return;
        const int per = (T + nz - 1) / nz;
        const int t_lo = z * per, t_hi = min(T, t_lo + per);
        const float* fm = freq_mean + (size_t)b * T;
        for (int idx = threadIdx.x; idx < (t_hi - t_lo) * w; idx += blockDim.x) {
            const int dt = idx / w;
            const int t = t_lo + dt;
            const int v = lo + (idx - dt * w);
            bool in_time = false;      // a time span overwrites the band there (utils.py:185-192 run after :176-183)
            for (int r = 0; r < R; ++r) in_time = in_time || (t >= s_a[r] && t < s_b[r]);
            if (!in_time) xb[(size_t)t * V + v] = fm[t];
        }
    } else {
        const int r = m - R;
        const int lo = max(__ldg(t0 + r * B + b), 0);
        const int hi = min(__ldg(t0 + r * B + b) + __ldg(tw + r * B + b), T);
        if (hi <= lo) return;
        const int per = (hi - lo + nz - 1) / nz;
        const int r_lo = lo + z * per, r_hi = min(hi, r_lo + per);
        const float* tm = time_mean + (size_t)b * V;
        for (int idx = threadIdx.x; idx < (r_hi - r_lo) * V; idx += blockDim.x) {
            const int dt = idx / V;
            const int v = idx - dt * V;
            xb[(size_t)(r_lo + dt) * V + v] = tm[v];
        }
    }
}

}  // namespace asr

using namespace asr;

extern "C" size_t asr_spec_aug_workspace_bytes(int B, int T, int V) {
    if (B <= 0 || T <= 0 || V <= 0) return 0;
    const size_t nchunk = (size_t)(T + kSaRows - 1) / kSaRows;
    // three sections, each starting on a 16-byte boundary
    return ((((size_t)B * T + 3) & ~(size_t)3) + (((size_t)B * V + 3) & ~(size_t)3) + (size_t)B * nchunk * V) * sizeof(float) + 256;
}

extern "C" int asr_spec_aug_f32(float* feats, const int* lens, const int* f0, const int* fw, const int* t0, const int* tw,
                                int R, int B, int T, int V, void* ws, size_t ws_bytes, void* stream) {
    ASR_REQUIRE(feats && lens && ws, "asr_spec_aug_f32: null pointer");
    ASR_REQUIRE(B > 0 && T > 0 && V > 0 && R >= 0, "asr_spec_aug_f32: bad shape B=%d T=%d V=%d R=%d", B, T, V, R);
    ASR_REQUIRE(V <= kSaMaxV, "asr_spec_aug_f32: V=%d exceeds %d feature bins", V, kSaMaxV);
    ASR_REQUIRE(B <= 65535, "asr_spec_aug_f32: B exceeds the grid limit");
    ASR_REQUIRE(R <= kSaMaxR, "asr_spec_aug_f32: more than %d masks per family", kSaMaxR);
    ASR_REQUIRE(R == 0 || (f0 && fw && t0 && tw), "asr_spec_aug_f32: null mask arrays");
    ASR_REQUIRE(ws_bytes >= asr_spec_aug_workspace_bytes(B, T, V), "asr_spec_aug_f32: workspace too small");
    if (asr_device_ok() != 0) return 3;
    if (R == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nchunk = (T + kSaRows - 1) / kSaRows;
    uintptr_t w = (reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255;
    float* freq_mean = reinterpret_cast<float*>(w);
    float* time_mean = freq_mean + (((size_t)B * T + 3) & ~(size_t)3);
    float* part = time_mean + (((size_t)B * V + 3) & ~(size_t)3);
    if ((V & 3) == 0 && aligned16(feats)) {
        const dim3 grid(nchunk, B);
        switch ((V + 127) / 128) {
            case 1: specaug_means_vec_kernel<1><<<grid, 256, 0, st>>>(feats, T, V, nchunk, freq_mean, part); break;
            case 2: specaug_means_vec_kernel<2><<<grid, 256, 0, st>>>(feats, T, V, nchunk, freq_mean, part); break;
            case 3: specaug_means_vec_kernel<3><<<grid, 256, 0, st>>>(feats, T, V, nchunk, freq_mean, part); break;
            case 4: specaug_means_vec_kernel<4><<<grid, 256, 0, st>>>(feats, T, V, nchunk, freq_mean, part); break;
            default: specaug_means_vec_kernel<8><<<grid, 256, 0, st>>>(feats, T, V, nchunk, freq_mean, part); break;
        }
    } else {
        specaug_means_kernel<<<dim3(nchunk, B), 256, 0, st>>>(feats, T, V, nchunk, freq_mean, part);
    }
    ASR_LAUNCH_CHECK();
    specaug_time_mean_kernel<<<dim3((V + 255) / 256, B), 256, 0, st>>>(part, lens, V, nchunk, time_mean);
    ASR_LAUNCH_CHECK();
    specaug_apply_kernel<<<dim3(2 * R, B, kSaSlabs), 256, 0, st>>>(feats, f0, fw, t0, tw, R, B, T, V, freq_mean, time_mean);
    ASR_LAUNCH_CHECK();
    return 0;
}

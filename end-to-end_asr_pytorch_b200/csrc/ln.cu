// ln.cu - dropout + residual + LayerNorm of the training step in one pass each way (sm_100a, HBM-bound).
//
// Every sub-layer of the reference's encoder / decoder ends in  LayerNorm(dropout(y) + residual)
// (/root/reference/src/transformer/module.py:48-53 - PositionwiseFeedForward; attention.py:59-60 - the projection after
// the heads; encoder.py:49 - LayerNorm of the input projection, no dropout / residual).  As torch kernels that is
// dropout (read y, write y' and a mask), add (read y', residual; write z), LayerNorm (read z; write out) and, backwards,
// the input gradient (read g, z), the gamma / beta gradients (read g, z again - a kernel that at 15 030 rows x 512 ran
// at a fifth of the memory bandwidth), and the dropout backward (read dz and the mask; write dy): 13 passes over a
// [rows x d_model] tensor.  Here:
//
//   forward   read y, residual;  write z = dropout(y) + residual (saved for backward) and out;  mean / rstd per row;
//             out may be scaled per row (the non-pad mask every layer multiplies its sub-layers' outputs with)
//   backward  read g_out, z;     write dz (= the residual's gradient) and dy = dz * keep / p_keep;  gamma / beta gradients
//             accumulated in registers over the rows a warp walks, added across the CTA's warps in a fixed order, one
//             partial row per CTA, and a column sum over the partial rows (gemm2.cu's asr_colsum): deterministic
//
// The dropout mask is never stored: both directions regenerate it from (seed, row, column) with the Philox stream the
// attention kernels use (philox.cuh); a CUDA-graph replay reads the seed from device memory.
// One warp per row; a lane owns the float4 groups lane, lane + 32, ... of the row (128-byte-coalesced warp accesses),
// so d_model must be a multiple of 128 (256 / 512 / 1024 are instantiated).  fp32 arithmetic; y / dy fp32 or bf16.
#include "common.cuh"
#include "philox.cuh"

#include <cuda_bf16.h>

#include <algorithm>

extern "C" int asr_colsum(const void* x, int is_bf16, int M, int N, int ld, float* out, void* stream);

namespace asr {

constexpr int kLnWarps = 8;
constexpr uint32_t kLnTag = 0x4c4e0001u;      // third Philox counter word: keeps these streams apart from the attention's (head index there)

struct LnArgs {
    const void* y;            // [M, D] fp32 / bf16 (forward);  unused (backward)
    const float* residual;    // [M, D] or null
    const float* gamma;
    const float* beta;
    const float* row_scale;   // [M] or null: out = LayerNorm(...) * row_scale[row] (the layers' non-pad mask)
    float* z;                 // forward: out (nullable);  backward: in
    float* out;               // forward: LayerNorm output;  backward: g_out (in)
    float* mean;              // [M]
    float* rstd;              // [M]
    __nv_bfloat16* out_bf16;  // forward: a bf16 copy of out for the next layer's GEMM (or null);  backward: its gradient (in, or null)
    float* g_z;               // backward: [M, D] or null
    void* g_y;                // backward: [M, D] fp32 / bf16 or null
    float* partial;           // backward: [gridDim.x, 2, D]
    const uint64_t* seed_dev;
    uint32_t seed_lo, seed_hi;
    uint32_t thresh;          // 0: no dropout
    float inv_keep;
    float eps;
    int M;
};

template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void store4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float4& v) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}

// keep factors (inv_keep or 0) of the 4 * NJ elements a lane owns in `row`: one Philox call per four float4 groups
template <int NJ>
__device__ __forceinline__ void keep_factors(float (&k)[NJ][4], int row, int lane, uint32_t thresh, float inv_keep, uint32_t lo, uint32_t hi) {
#pragma unroll
    for (int jq = 0; jq < (NJ + 3) / 4; ++jq) {
        const uint4 rnd = philox16((uint32_t)(lane + 32 * jq), (uint32_t)row, kLnTag, lo, hi);
#pragma unroll
        for (int j = 4 * jq; j < NJ && j < 4 * jq + 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) k[j][e] = philox_byte(rnd, 4 * (j & 3) + e) >= thresh ? inv_keep : 0.0f;
    }
}

template <typename TY, int NJ>
__global__ void __launch_bounds__(kLnWarps * 32) ln_fwd_kernel(const LnArgs a) {
    constexpr int D = 128 * NJ;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t lo = a.seed_lo, hi = a.seed_hi;
    effective_seed(a.seed_dev, lo, hi);
    float gam[NJ][4], bet[NJ][4];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma) + lane + 32 * j);
        const float4 b = __ldg(reinterpret_cast<const float4*>(a.beta) + lane + 32 * j);
        gam[j][0] = g.x; gam[j][1] = g.y; gam[j][2] = g.z; gam[j][3] = g.w;
        bet[j][0] = b.x; bet[j][1] = b.y; bet[j][2] = b.z; bet[j][3] = b.w;
    }
    const TY* y = static_cast<const TY*>(a.y);
    for (int row = blockIdx.x * kLnWarps + warp; row < a.M; row += gridDim.x * kLnWarps) {
        const size_t base = (size_t)row * D;
        float v[NJ][4];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const float4 t = load4<TY>(y + base + 4 * (lane + 32 * j));
            v[j][0] = t.x; v[j][1] = t.y; v[j][2] = t.z; v[j][3] = t.w;
        }
        if (a.thresh != 0) {
            float k[NJ][4];
            keep_factors<NJ>(k, row, lane, a.thresh, a.inv_keep, lo, hi);
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) v[j][e] *= k[j][e];
        }
        if (a.residual != nullptr) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const float4 r = *reinterpret_cast<const float4*>(a.residual + base + 4 * (lane + 32 * j));
                v[j][0] += r.x; v[j][1] += r.y; v[j][2] += r.z; v[j][3] += r.w;
            }
        }
        if (a.z != nullptr) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) store4(a.z + base + 4 * (lane + 32 * j), make_float4(v[j][0], v[j][1], v[j][2], v[j][3]));
        }
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) s += (v[j][0] + v[j][1]) + (v[j][2] + v[j][3]);
        const float mean = warp_sum(s) * (1.0f / D);
        float q = 0.0f;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float d = v[j][e] - mean;
                q += d * d;
            }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + a.eps);
        if (lane == 0) {
            a.mean[row] = mean;
            a.rstd[row] = rstd;
        }
        const float rs = a.row_scale != nullptr ? __ldg(a.row_scale + row) : 1.0f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            float4 o;
            o.x = ((v[j][0] - mean) * rstd * gam[j][0] + bet[j][0]) * rs;
            o.y = ((v[j][1] - mean) * rstd * gam[j][1] + bet[j][1]) * rs;
            o.z = ((v[j][2] - mean) * rstd * gam[j][2] + bet[j][2]) * rs;
            o.w = ((v[j][3] - mean) * rstd * gam[j][3] + bet[j][3]) * rs;
            store4(a.out + base + 4 * (lane + 32 * j), o);
            if (a.out_bf16 != nullptr) store4(a.out_bf16 + base + 4 * (lane + 32 * j), o);
        }
    }
}

template <typename TY, int NJ>
__global__ void __launch_bounds__(kLnWarps * 32) ln_bwd_kernel(const LnArgs a) {
    constexpr int D = 128 * NJ;
    __shared__ float red[2][D];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t lo = a.seed_lo, hi = a.seed_hi;
    effective_seed(a.seed_dev, lo, hi);
    float gam[NJ][4], dgam[NJ][4], dbet[NJ][4];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma) + lane + 32 * j);
        gam[j][0] = g.x; gam[j][1] = g.y; gam[j][2] = g.z; gam[j][3] = g.w;
#pragma unroll
        for (int e = 0; e < 4; ++e) dgam[j][e] = dbet[j][e] = 0.0f;
    }
    TY* g_y = static_cast<TY*>(a.g_y);
    for (int row = blockIdx.x * kLnWarps + warp; row < a.M; row += gridDim.x * kLnWarps) {
        const size_t base = (size_t)row * D;
        const float mean = a.mean[row], rstd = a.rstd[row];
        const float rs = a.row_scale != nullptr ? __ldg(a.row_scale + row) : 1.0f;
        float go[NJ][4], xh[NJ][4];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.out != nullptr) g = *reinterpret_cast<const float4*>(a.out + base + 4 * (lane + 32 * j));
            if (a.out_bf16 != nullptr) {         // the gradient that arrived through the bf16 copy
                const float4 h = load4<__nv_bfloat16>(a.out_bf16 + base + 4 * (lane + 32 * j));
                g.x += h.x; g.y += h.y; g.z += h.z; g.w += h.w;
            }
            const float4 z = *reinterpret_cast<const float4*>(a.z + base + 4 * (lane + 32 * j));
            go[j][0] = g.x * rs; go[j][1] = g.y * rs; go[j][2] = g.z * rs; go[j][3] = g.w * rs;
            xh[j][0] = (z.x - mean) * rstd; xh[j][1] = (z.y - mean) * rstd; xh[j][2] = (z.z - mean) * rstd; xh[j][3] = (z.w - mean) * rstd;
        }
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                dgam[j][e] += go[j][e] * xh[j][e];
                dbet[j][e] += go[j][e];
                const float g = go[j][e] * gam[j][e];
                go[j][e] = g;
                s1 += g;
                s2 += g * xh[j][e];
            }
        const float c1 = warp_sum(s1) * (1.0f / D), c2 = warp_sum(s2) * (1.0f / D);
        float k[NJ][4];
        if (a.thresh != 0 && g_y != nullptr) keep_factors<NJ>(k, row, lane, a.thresh, a.inv_keep, lo, hi);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            float4 dz;
            dz.x = rstd * (go[j][0] - c1 - xh[j][0] * c2);
            dz.y = rstd * (go[j][1] - c1 - xh[j][1] * c2);
            dz.z = rstd * (go[j][2] - c1 - xh[j][2] * c2);
            dz.w = rstd * (go[j][3] - c1 - xh[j][3] * c2);
            if (a.g_z != nullptr) store4(a.g_z + base + 4 * (lane + 32 * j), dz);
            if (g_y != nullptr) {
                if (a.thresh != 0) { dz.x *= k[j][0]; dz.y *= k[j][1]; dz.z *= k[j][2]; dz.w *= k[j][3]; }
                store4(g_y + base + 4 * (lane + 32 * j), dz);
            }
        }
    }
    // the CTA's warps add their column sums one after the other (fixed order), then one partial row per CTA
    for (int w = 0; w < kLnWarps; ++w) {
        if (warp == w) {
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = 4 * (lane + 32 * j) + e;
                    red[0][c] = (w == 0 ? 0.0f : red[0][c]) + dgam[j][e];
                    red[1][c] = (w == 0 ? 0.0f : red[1][c]) + dbet[j][e];
                }
        }
        __syncthreads();
    }
    float* dst = a.partial + (size_t)blockIdx.x * 2 * D;
    for (int i = threadIdx.x; i < 2 * D; i += kLnWarps * 32) dst[i] = (&red[0][0])[i];
}

// Evaluation flavour: out = LayerNorm(y + residual) * gamma + beta with bf16 y / residual / out (the model in bf16, no
// gradient): three 2-byte passes over the rows, nothing saved.
template <int NJ>
__global__ void __launch_bounds__(kLnWarps * 32) ln_eval_bf16_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ residual,
                                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                    __nv_bfloat16* __restrict__ out, int M, float eps) {
    constexpr int D = 128 * NJ;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float gam[NJ][4], bet[NJ][4];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * j);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * j);
        gam[j][0] = g.x; gam[j][1] = g.y; gam[j][2] = g.z; gam[j][3] = g.w;
        bet[j][0] = b.x; bet[j][1] = b.y; bet[j][2] = b.z; bet[j][3] = b.w;
    }
    for (int row = blockIdx.x * kLnWarps + warp; row < M; row += gridDim.x * kLnWarps) {
        const size_t base = (size_t)row * D;
        float v[NJ][4];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const float4 t = load4<__nv_bfloat16>(y + base + 4 * (lane + 32 * j));
            v[j][0] = t.x; v[j][1] = t.y; v[j][2] = t.z; v[j][3] = t.w;
        }
        if (residual != nullptr) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const float4 r = load4<__nv_bfloat16>(residual + base + 4 * (lane + 32 * j));
                v[j][0] += r.x; v[j][1] += r.y; v[j][2] += r.z; v[j][3] += r.w;
            }
        }
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) s += (v[j][0] + v[j][1]) + (v[j][2] + v[j][3]);
        const float mean = warp_sum(s) * (1.0f / D);
        float q = 0.0f;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float d = v[j][e] - mean;
                q += d * d;
            }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            float4 o;
            o.x = (v[j][0] - mean) * rstd * gam[j][0] + bet[j][0];
            o.y = (v[j][1] - mean) * rstd * gam[j][1] + bet[j][1];
            o.z = (v[j][2] - mean) * rstd * gam[j][2] + bet[j][2];
            o.w = (v[j][3] - mean) * rstd * gam[j][3] + bet[j][3];
            store4(out + base + 4 * (lane + 32 * j), o);
        }
    }
}

// keep[row, col] = 1 where the dropout keeps the element (the bits the kernels regenerate): for tests
__global__ void __launch_bounds__(256) ln_dropout_keep_kernel(uint8_t* keep, int M, int D, uint32_t thresh, uint32_t lo, uint32_t hi) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // one float4 group each
    const int groups = D / 4;
    if (idx >= (long long)M * groups) return;
    const int row = (int)(idx / groups), grp = (int)(idx % groups);
    const int lane = grp & 31, j = grp >> 5;
    const uint4 rnd = philox16((uint32_t)(lane + 32 * (j >> 2)), (uint32_t)row, kLnTag, lo, hi);
    for (int e = 0; e < 4; ++e) keep[(size_t)row * D + 4 * grp + e] = philox_byte(rnd, 4 * (j & 3) + e) >= thresh ? 1 : 0;
}

// CTAs per SM: what the registers allow (forward 80 -> 3, backward 128 -> 2); an HBM-bound kernel wants the loads of all of them in flight
// out = (y > 0) ? gy : 0 on bf16 vectors of 8: the ReLU backward of the feed-forward block's first layer (the mask comes
// from the saved OUTPUT, module.py:50), one pass instead of torch's compare + cast + multiply
__global__ void __launch_bounds__(256) relu_bwd_bf16_kernel(const uint4* __restrict__ gy, const uint4* __restrict__ y, uint4* __restrict__ out,
                                                           size_t n8, const __nv_bfloat16* gy_tail, const __nv_bfloat16* y_tail,
                                                           __nv_bfloat16* out_tail, int tail) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const uint4 g = gy[i], v = y[i];
        uint4 o;
        // y > 0 as torch evaluates it (false for zero, negative and NaN activations)
        auto pick = [](uint32_t gw, uint32_t yw) {
            const __nv_bfloat162 yy = *reinterpret_cast<const __nv_bfloat162*>(&yw);
            const float2 f = __bfloat1622float2(yy);
            return (f.x > 0.0f ? (gw & 0x0000ffffu) : 0u) | (f.y > 0.0f ? (gw & 0xffff0000u) : 0u);
        };
        o.x = pick(g.x, v.x); o.y = pick(g.y, v.y); o.z = pick(g.z, v.z); o.w = pick(g.w, v.w);
        out[i] = o;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < tail)
        out_tail[threadIdx.x] = __bfloat162float(y_tail[threadIdx.x]) > 0.0f ? gy_tail[threadIdx.x] : __float2bfloat16_rn(0.0f);
}

static int ln_grid(int M, int per_sm) { return std::max(1, std::min((M + kLnWarps - 1) / kLnWarps, per_sm * num_sms())); }

}  // namespace asr

using namespace asr;

static int ln_check(const char* who, int M, int D, float p_drop) {
    ASR_REQUIRE(M > 0, "%s: no rows", who);
    ASR_REQUIRE(D == 256 || D == 512 || D == 1024, "%s: width %d not supported (256, 512, 1024)", who, D);
    ASR_REQUIRE(p_drop >= 0.0f && p_drop < 1.0f, "%s: p_drop %f outside [0, 1)", who, (double)p_drop);
    return 0;
}

#define ASR_LN_DISPATCH(KERNEL, BF16, D, ...)                                             \
    do {                                                                                  \
        if (BF16) {                                                                       \
            if (D == 256) KERNEL<__nv_bfloat16, 2> __VA_ARGS__;                           \
            else if (D == 512) KERNEL<__nv_bfloat16, 4> __VA_ARGS__;                      \
            else KERNEL<__nv_bfloat16, 8> __VA_ARGS__;                                    \
        } else {                                                                          \
            if (D == 256) KERNEL<float, 2> __VA_ARGS__;                                   \
            else if (D == 512) KERNEL<float, 4> __VA_ARGS__;                              \
            else KERNEL<float, 8> __VA_ARGS__;                                            \
        }                                                                                 \
    } while (0)

extern "C" float asr_ln_dropout_keep_prob(float p_drop) { return (256.0f - (float)drop_threshold(p_drop)) / 256.0f; }

extern "C" int asr_ln_fwd(const void* y, int y_bf16, const float* residual, const float* gamma, const float* beta,
                          const float* row_scale, int M, int D, float eps, float p_drop, uint64_t seed, const uint64_t* seed_dev, float* z,
                          float* out, void* out_bf16, float* mean, float* rstd, void* stream) {
    if (ln_check("asr_ln_fwd", M, D, p_drop)) return 2;
    ASR_REQUIRE(y && gamma && beta && out && mean && rstd, "asr_ln_fwd: null pointer");
    ASR_REQUIRE(aligned16(y) && aligned16(residual) && aligned16(gamma) && aligned16(beta) && aligned16(z) && aligned16(out) && aligned16(out_bf16),
                "asr_ln_fwd: pointers must be 16-byte aligned");
    if (asr_device_ok() != 0) return 3;
    LnArgs a = {};
    a.y = y; a.residual = residual; a.gamma = gamma; a.beta = beta; a.row_scale = row_scale; a.z = z; a.out = out; a.out_bf16 = static_cast<__nv_bfloat16*>(out_bf16); a.mean = mean; a.rstd = rstd;
    a.seed_dev = seed_dev; a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
    a.thresh = drop_threshold(p_drop);
    a.inv_keep = 256.0f / (256.0f - (float)a.thresh);
    a.eps = eps; a.M = M;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = ln_grid(M, 3);
    ASR_LN_DISPATCH(ln_fwd_kernel, y_bf16, D, <<<grid, kLnWarps * 32, 0, st>>>(a));
    ASR_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t asr_ln_bwd_workspace_bytes(int M, int D) {
    if (M <= 0 || D <= 0) return 0;
    return (size_t)ln_grid(M, 2) * 2 * (size_t)D * sizeof(float);
}

extern "C" int asr_ln_bwd(const float* g_out, const void* g_out_bf16, const float* z, const float* mean, const float* rstd, const float* gamma,
                          const float* row_scale, int M, int D, float p_drop, uint64_t seed, const uint64_t* seed_dev, float* g_z, void* g_y, int y_bf16,
                          float* g_gamma_beta, void* ws, size_t ws_bytes, void* stream) {
    if (ln_check("asr_ln_bwd", M, D, p_drop)) return 2;
    ASR_REQUIRE((g_out || g_out_bf16) && z && mean && rstd && gamma && g_gamma_beta && ws, "asr_ln_bwd: null pointer");
    ASR_REQUIRE(aligned16(g_out_bf16), "asr_ln_bwd: pointers must be 16-byte aligned");
    ASR_REQUIRE(g_z || g_y, "asr_ln_bwd: neither input gradient requested");
    ASR_REQUIRE(aligned16(g_out) && aligned16(z) && aligned16(gamma) && aligned16(g_z) && aligned16(g_y) && aligned16(ws),
                "asr_ln_bwd: pointers must be 16-byte aligned");
    ASR_REQUIRE(ws_bytes >= asr_ln_bwd_workspace_bytes(M, D), "asr_ln_bwd: workspace too small");
    if (asr_device_ok() != 0) return 3;
    LnArgs a = {};
    a.gamma = gamma; a.row_scale = row_scale; a.z = const_cast<float*>(z); a.out = const_cast<float*>(g_out);
    a.out_bf16 = static_cast<__nv_bfloat16*>(const_cast<void*>(g_out_bf16));
    a.mean = const_cast<float*>(mean); a.rstd = const_cast<float*>(rstd);
    a.g_z = g_z; a.g_y = g_y; a.partial = static_cast<float*>(ws);
    a.seed_dev = seed_dev; a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
    a.thresh = drop_threshold(p_drop);
    a.inv_keep = 256.0f / (256.0f - (float)a.thresh);
    a.M = M;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = ln_grid(M, 2);
    ASR_LN_DISPATCH(ln_bwd_kernel, y_bf16, D, <<<grid, kLnWarps * 32, 0, st>>>(a));
    ASR_LAUNCH_CHECK();
    return asr_colsum(ws, 0, grid, 2 * D, 2 * D, g_gamma_beta, stream);
}

extern "C" int asr_ln_dropout_keep(uint8_t* keep, int M, int D, float p_drop, uint64_t seed, void* stream) {
    if (ln_check("asr_ln_dropout_keep", M, D, p_drop)) return 2;
    ASR_REQUIRE(keep, "asr_ln_dropout_keep: null pointer");
    if (asr_device_ok() != 0) return 3;
    const long long total = (long long)M * (D / 4);
    ln_dropout_keep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        keep, M, D, drop_threshold(p_drop), (uint32_t)seed, (uint32_t)(seed >> 32));
    ASR_LAUNCH_CHECK();
    return 0;
}

extern "C" int asr_relu_bwd_bf16(const void* gy, const void* y, void* out, size_t n, void* stream) {
    ASR_REQUIRE(gy && y && out, "asr_relu_bwd_bf16: null pointer");
    ASR_REQUIRE(aligned16(gy) && aligned16(y) && aligned16(out), "asr_relu_bwd_bf16: pointers must be 16-byte aligned");
    if (asr_device_ok() != 0) return 3;
    if (n == 0) return 0;
    const size_t n8 = n / 8;
    const int tail = (int)(n - n8 * 8);
    const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((n8 + 255) / 256, (size_t)num_sms() * 8));
    const __nv_bfloat16 *g16 = static_cast<const __nv_bfloat16*>(gy), *y16 = static_cast<const __nv_bfloat16*>(y);
    relu_bwd_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(gy), static_cast<const uint4*>(y), static_cast<uint4*>(out), n8, g16 + n8 * 8, y16 + n8 * 8,
        static_cast<__nv_bfloat16*>(out) + n8 * 8, tail);
    ASR_LAUNCH_CHECK();
    return 0;
}

extern "C" int asr_ln_eval_bf16(const void* y, const void* residual, const float* gamma, const float* beta, int M, int D, float eps,
                                void* out, void* stream) {
    if (ln_check("asr_ln_eval_bf16", M, D, 0.0f)) return 2;
    ASR_REQUIRE(y && gamma && beta && out, "asr_ln_eval_bf16: null pointer");
    ASR_REQUIRE(aligned16(y) && aligned16(residual) && aligned16(gamma) && aligned16(beta) && aligned16(out),
                "asr_ln_eval_bf16: pointers must be 16-byte aligned");
    if (asr_device_ok() != 0) return 3;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = ln_grid(M, 4);
    const __nv_bfloat16 *y16 = static_cast<const __nv_bfloat16*>(y), *r16 = static_cast<const __nv_bfloat16*>(residual);
    __nv_bfloat16* o16 = static_cast<__nv_bfloat16*>(out);
    if (D == 256) ln_eval_bf16_kernel<2><<<grid, kLnWarps * 32, 0, st>>>(y16, r16, gamma, beta, o16, M, eps);
    else if (D == 512) ln_eval_bf16_kernel<4><<<grid, kLnWarps * 32, 0, st>>>(y16, r16, gamma, beta, o16, M, eps);
    else ln_eval_bf16_kernel<8><<<grid, kLnWarps * 32, 0, st>>>(y16, r16, gamma, beta, o16, M, eps);
    ASR_LAUNCH_CHECK();
    return 0;
}

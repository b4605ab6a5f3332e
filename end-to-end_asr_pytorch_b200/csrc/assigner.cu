// The CIF weight producer: tail of the attention assigner fused with the scaling glue of
// CIF_Model.forward  (SURVEY.md 8(f2)).
//
// Reference:
//   /root/reference/src/transformer/attentionAssigner.py:36-40
//       alphas = sigmoid(linear(x).squeeze(-1)) * sequence_mask(input_lengths)
//   /root/reference/src/transformer/cif_model.py:43-48
//       _num = alpha.sum(-1);  alpha *= (num_noise / _num)[:, None]
// (num_noise = #targets + U[0,1) - 0.5 is drawn by the caller, cif_model.py:46-47.)
//
//   forward   K1 assigner_rows     one warp per frame row: <x_bt, w> + b -> sigmoid -> pad mask -> a_raw
//                                  (padded rows are not read); HBM-bound: reads x once
//             K2 assigner_scale    one CTA per utterance: _num = sum_t a_raw in a fixed order,
//                                  alpha = a_raw * num_noise / _num
//   backward  K3 assigner_dz       one CTA per utterance: through the scaling (alpha and _num both
//                                  depend on every a_raw of the utterance), the mask and the sigmoid
//             K4 assigner_grads    g_x = dz * w (one write of [B,T,D]) and the per-CTA partial of
//                                  g_w = sum dz * x (one read of x), fixed row -> CTA assignment
//             K5 assigner_reduce   partials -> g_w, g_bias in a fixed order
// Every reduction has a fixed order: results are reproducible run to run (no atomics).
#include "common.cuh"

#include <cstdint>

namespace asr {

struct AssignerArgs {
    const float* x;          // [B,T,D]
    const float* w;          // [D]
    const float* bias;       // [1]
    const int* len;          // [B]
    const float* num_noise;  // [B] or null (no scaling: alpha = a_raw)
    int B, T, D;
    float* alpha;            // [B,T]
    float* a_raw;            // [B,T]
    float* num_raw;          // [B]   _num
    // backward
    const float* g_alpha;    // [B,T]
    const float* g_num;      // [B] or null
    float* dz;               // [B,T]     workspace
    float* part;             // [nCTA, D + 1] workspace: partial g_w, last = partial g_bias
    float* g_x;              // [B,T,D]
    float* g_w;              // [D]
    float* g_bias;           // [1]
    int vec4;                // rows are 16-byte aligned and D % 4 == 0
};

__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    float r = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += red[i];
    return r;
}

// ---- K1: a_raw ---------------------------------------------------------------------
constexpr int kRowWarps = 8;

__global__ void __launch_bounds__(kRowWarps * 32) assigner_rows_kernel(const AssignerArgs a) {
    extern __shared__ __align__(16) float sw[];   // [D]
    for (int i = threadIdx.x; i < a.D; i += blockDim.x) sw[i] = __ldg(a.w + i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const float bias = __ldg(a.bias);
    const long long rows = (long long)a.B * a.T;
    const long long stride = (long long)gridDim.x * kRowWarps;
    constexpr int R = 4;                             // rows per warp iteration: R * (D/128) 16-byte loads in flight per lane
    for (long long row0 = (long long)blockIdx.x * kRowWarps + (threadIdx.x >> 5); row0 < rows; row0 += R * stride) {
        float acc[R];
        bool live[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long row = row0 + r * stride;
            acc[r] = 0.0f;
            live[r] = false;
            if (row < rows) {
                const int b = (int)(row / a.T);
                const int t = (int)(row - (long long)b * a.T);
                live[r] = t < min(max(__ldg(a.len + b), 0), a.T);      // warp-uniform
            }
        }
        if (a.vec4) {
            const float4* wv = reinterpret_cast<const float4*>(sw);
            for (int i = lane; i < a.D / 4; i += 32) {
                float4 xx[R];
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (live[r]) xx[r] = __ldg(reinterpret_cast<const float4*>(a.x + (size_t)(row0 + r * stride) * a.D) + i);
                const float4 ww = wv[i];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (live[r]) {
                        acc[r] = fmaf(xx[r].x, ww.x, acc[r]);
                        acc[r] = fmaf(xx[r].y, ww.y, acc[r]);
                        acc[r] = fmaf(xx[r].z, ww.z, acc[r]);
                        acc[r] = fmaf(xx[r].w, ww.w, acc[r]);
                    }
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (live[r]) {
                    const float* xr = a.x + (size_t)(row0 + r * stride) * a.D;
                    for (int i = lane; i < a.D; i += 32) acc[r] = fmaf(__ldg(xr + i), sw[i], acc[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long row = row0 + r * stride;
            if (row < rows) {
                float out = 0.0f;
                if (live[r]) {
                    const float z = warp_sum(acc[r]) + bias;
                    out = 1.0f / (1.0f + expf(-z));  // torch.sigmoid
                }
                if (lane == 0) a.a_raw[row] = out;
            }
        }
    }
}

// ---- K2: _num and the scaled weights ------------------------------------------------
__global__ void __launch_bounds__(256) assigner_scale_kernel(const AssignerArgs a) {
    __shared__ float red[8];
    const int b = blockIdx.x;
    const float* ar = a.a_raw + (size_t)b * a.T;
    float s = 0.0f;
    for (int t = threadIdx.x; t < a.T; t += 256) s += ar[t];
    const float num = block_sum_256(s, red);
    const float r = a.num_noise ? __ldg(a.num_noise + b) / num : 1.0f;
    if (threadIdx.x == 0) a.num_raw[b] = num;
    for (int t = threadIdx.x; t < a.T; t += 256) a.alpha[(size_t)b * a.T + t] = ar[t] * r;
}

// ---- K3: gradient at the pre-sigmoid activations ------------------------------------
// alpha_t = a_t * r, r = n / N, N = sum_t a_t:
//   d a_t = g_alpha_t * r + c,   c = g_N - (r / N) * sum_t' g_alpha_t' a_t'
//   d z_t = d a_t * mask_t * a_t (1 - a_t)           (a_t is the sigmoid where the mask is 1)
__global__ void __launch_bounds__(256) assigner_dz_kernel(const AssignerArgs a) {
    __shared__ float red[8];
    const int b = blockIdx.x;
    const float* ar = a.a_raw + (size_t)b * a.T;
    const float* ga = a.g_alpha + (size_t)b * a.T;
    const int len = min(max(__ldg(a.len + b), 0), a.T);
    float dot = 0.0f;
    for (int t = threadIdx.x; t < a.T; t += 256) dot = fmaf(ga[t], ar[t], dot);
    dot = block_sum_256(dot, red);
    const float num = a.num_raw[b];
    float r = 1.0f, c = a.g_num ? __ldg(a.g_num + b) : 0.0f;
    if (a.num_noise) {
        r = __ldg(a.num_noise + b) / num;
        c -= (r / num) * dot;
    }
    for (int t = threadIdx.x; t < a.T; t += 256) {
        const float s = ar[t];
        a.dz[(size_t)b * a.T + t] = (t < len) ? (ga[t] * r + c) * s * (1.0f - s) : 0.0f;
    }
}

// ---- K4: g_x and the partial g_w -----------------------------------------------------
constexpr int kGradWarps = 8;

__global__ void __launch_bounds__(kGradWarps * 32) assigner_grads_kernel(const AssignerArgs a) {
    extern __shared__ __align__(16) float smem[];     // [D] w | [kGradWarps][D] partial g_w
    float* sw = smem;
    float* pw = smem + a.D;
    for (int i = threadIdx.x; i < a.D; i += blockDim.x) sw[i] = __ldg(a.w + i);
    for (int i = threadIdx.x; i < kGradWarps * a.D; i += blockDim.x) pw[i] = 0.0f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* myw = pw + (size_t)warp * a.D;
    float bsum = 0.0f;
    const long long rows = (long long)a.B * a.T;
    const long long stride = (long long)gridDim.x * kGradWarps;
    for (long long row = (long long)blockIdx.x * kGradWarps + warp; row < rows; row += stride) {
        const float d = a.dz[row];                    // warp-uniform
        float* gx = a.g_x + (size_t)row * a.D;
        const float* xr = a.x + (size_t)row * a.D;
        if (a.vec4) {
            float4* gv = reinterpret_cast<float4*>(gx);
            if (d == 0.0f) {                          // padded frame (or a saturated sigmoid): zero row, x is not read
                for (int i = lane; i < a.D / 4; i += 32) __stcs(gv + i, make_float4(0.f, 0.f, 0.f, 0.f));
                continue;
            }
            const float4* xv = reinterpret_cast<const float4*>(xr);
            const float4* wv = reinterpret_cast<const float4*>(sw);
            float4* mv = reinterpret_cast<float4*>(myw);
            for (int i = lane; i < a.D / 4; i += 32) {
                const float4 ww = wv[i];
                __stcs(gv + i, make_float4(d * ww.x, d * ww.y, d * ww.z, d * ww.w));
                const float4 xx = __ldg(xv + i);
                float4 m = mv[i];
                m.x = fmaf(d, xx.x, m.x);
                m.y = fmaf(d, xx.y, m.y);
                m.z = fmaf(d, xx.z, m.z);
                m.w = fmaf(d, xx.w, m.w);
                mv[i] = m;
            }
        } else {
            if (d == 0.0f) {
                for (int i = lane; i < a.D; i += 32) gx[i] = 0.0f;
                continue;
            }
            for (int i = lane; i < a.D; i += 32) {
                gx[i] = d * sw[i];
                myw[i] = fmaf(d, __ldg(xr + i), myw[i]);
            }
        }
        bsum += d;                                    // same value in every lane
    }
    __syncthreads();
    float* out = a.part + (size_t)blockIdx.x * (a.D + 1);
    for (int i = threadIdx.x; i < a.D; i += blockDim.x) {
        float s = 0.0f;
#pragma unroll
        for (int w2 = 0; w2 < kGradWarps; ++w2) s += pw[(size_t)w2 * a.D + i];
        out[i] = s;
    }
    __shared__ float bred[kGradWarps];
    if (lane == 0) bred[warp] = bsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
#pragma unroll
        for (int w2 = 0; w2 < kGradWarps; ++w2) s += bred[w2];
        out[a.D] = s;
    }
}

// ---- K5: partials -> g_w, g_bias -------------------------------------------------------
__global__ void __launch_bounds__(256) assigner_reduce_kernel(const AssignerArgs a, int nparts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > a.D) return;
    float s = 0.0f;
    for (int p = 0; p < nparts; ++p) s += a.part[(size_t)p * (a.D + 1) + i];
    if (i < a.D)
        a.g_w[i] = s;
    else
        a.g_bias[0] = s;
}

static int grads_ctas() { return 2 * num_sms(); }

}  // namespace asr

using namespace asr;

extern "C" int asr_cif_alpha_fwd_f32(const float* x, const float* w, const float* bias, const int* len,
                                     const float* num_noise, int B, int T, int D, float* alpha, float* a_raw,
                                     float* num_raw, void* stream) {
    ASR_REQUIRE(B > 0 && T > 0 && D > 0, "asr_cif_alpha_fwd_f32: bad shape B=%d T=%d D=%d", B, T, D);
    ASR_REQUIRE(x && w && bias && len && alpha && a_raw && num_raw, "asr_cif_alpha_fwd_f32: null pointer");
    ASR_REQUIRE(D <= 12288, "asr_cif_alpha_fwd_f32: D=%d > 12288 not supported", D);
    if (asr_device_ok() != 0) return 3;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    AssignerArgs a{};
    a.x = x; a.w = w; a.bias = bias; a.len = len; a.num_noise = num_noise;
    a.B = B; a.T = T; a.D = D;
    a.alpha = alpha; a.a_raw = a_raw; a.num_raw = num_raw;
    a.vec4 = (D % 4 == 0) && aligned16(x);
    const long long rows = (long long)B * T;
    long long blocks = (rows + kRowWarps - 1) / kRowWarps;
    if (blocks > 8ll * num_sms()) blocks = 8ll * num_sms();
    assigner_rows_kernel<<<(unsigned)blocks, kRowWarps * 32, (size_t)D * 4, st>>>(a);
    ASR_LAUNCH_CHECK();
    assigner_scale_kernel<<<B, 256, 0, st>>>(a);
    ASR_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t asr_cif_alpha_bwd_workspace_bytes(int B, int T, int D) {
    if (B <= 0 || T <= 0 || D <= 0) return 0;
    return (size_t)B * T * sizeof(float) + (size_t)grads_ctas() * (D + 1) * sizeof(float) + 256;
}

extern "C" int asr_cif_alpha_bwd_f32(const float* x, const float* w, const int* len, const float* num_noise,
                                     const float* a_raw, const float* num_raw, const float* g_alpha, const float* g_num,
                                     int B, int T, int D, float* g_x, float* g_w, float* g_bias, void* ws,
                                     size_t ws_bytes, void* stream) {
    ASR_REQUIRE(B > 0 && T > 0 && D > 0, "asr_cif_alpha_bwd_f32: bad shape B=%d T=%d D=%d", B, T, D);
    ASR_REQUIRE(x && w && len && a_raw && num_raw && g_alpha && g_x && g_w && g_bias && ws, "asr_cif_alpha_bwd_f32: null pointer");
    ASR_REQUIRE(ws_bytes >= asr_cif_alpha_bwd_workspace_bytes(B, T, D), "asr_cif_alpha_bwd_f32: workspace too small");
    const size_t smem = (size_t)(1 + kGradWarps) * D * 4;
    ASR_REQUIRE(smem <= 200 * 1024, "asr_cif_alpha_bwd_f32: D=%d needs %zu bytes of shared memory", D, smem);
    if (asr_device_ok() != 0) return 3;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    AssignerArgs a{};
    a.x = x; a.w = w; a.len = len; a.num_noise = num_noise;
    a.B = B; a.T = T; a.D = D;
    a.a_raw = const_cast<float*>(a_raw);
    a.num_raw = const_cast<float*>(num_raw);
    a.g_alpha = g_alpha; a.g_num = g_num;
    uintptr_t p = (reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255;
    a.dz = reinterpret_cast<float*>(p);
    a.part = a.dz + (size_t)B * T;
    a.g_x = g_x; a.g_w = g_w; a.g_bias = g_bias;
    a.vec4 = (D % 4 == 0) && aligned16(x) && aligned16(g_x);
    assigner_dz_kernel<<<B, 256, 0, st>>>(a);
    ASR_LAUNCH_CHECK();
    const int nparts = grads_ctas();
    ASR_CHECK_CUDA(cudaFuncSetAttribute(assigner_grads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    assigner_grads_kernel<<<nparts, kGradWarps * 32, smem, st>>>(a);
    ASR_LAUNCH_CHECK();
    assigner_reduce_kernel<<<(D + 1 + 255) / 256, 256, 0, st>>>(a, nparts);
    ASR_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------
// Input side of the path: low-frame-rate stacking of a padded batch.
// Reference: build_LFR_features, /root/reference/src/utils/data.py:191-218 (numpy, one
// utterance at a time in the data loader): output frame i is the concatenation of input
// frames i*n .. i*n+m-1, the last ones repeated beyond the end of the utterance.
//   out[b, i, j*D + d] = in[b, min(i*n + j, len_b - 1), d]   for i < ceil(len_b / n), else 0
// Pure copy: bit-exact.  One warp per output frame, 16-byte lanes when the rows allow it.
// ---------------------------------------------------------------------------------
namespace asr {

__global__ void __launch_bounds__(256) lfr_kernel(const float* in, const int* len, int B, int T, int D, int m, int n,
                                                  int To, float* out, int* out_len, int vec4) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= (long long)B * To) return;
    const int b = (int)(row / To);
    const int i = (int)(row - (long long)b * To);
    const int lb = min(max(__ldg(len + b), 0), T);
    const int lo = (lb + n - 1) / n;
    if (i == 0 && lane == 0) out_len[b] = lo;
    float* orow = out + (size_t)row * m * D;
    const float* ib = in + (size_t)b * T * D;
    if (vec4) {
        const int dv = D / 4;
        float4* ov = reinterpret_cast<float4*>(orow);
        for (int c = lane; c < m * dv; c += 32) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < lo) {
                const int j = c / dv;
                const int t = min(i * n + j, lb - 1);
                v = __ldg(reinterpret_cast<const float4*>(ib + (size_t)t * D) + (c - j * dv));
            }
            ov[c] = v;
        }
    } else {
        for (int c = lane; c < m * D; c += 32) {
            float v = 0.0f;
            if (i < lo) {
                const int j = c / D;
                const int t = min(i * n + j, lb - 1);
                v = __ldg(ib + (size_t)t * D + (c - j * D));
            }
            orow[c] = v;
        }
    }
}

}  // namespace asr

extern "C" int asr_lfr_f32(const float* in, const int* len, int B, int T, int D, int m, int n, float* out, int* out_len,
                           void* stream) {
    ASR_REQUIRE(in && len && out && out_len, "asr_lfr_f32: null pointer");
    ASR_REQUIRE(B > 0 && T > 0 && D > 0 && m > 0 && n > 0, "asr_lfr_f32: bad arguments B=%d T=%d D=%d m=%d n=%d", B, T, D, m, n);
    if (asr_device_ok() != 0) return 3;
    const int To = (T + n - 1) / n;
    const int vec4 = (D % 4 == 0) && aligned16(in) && aligned16(out);
    const long long rows = (long long)B * To;
    lfr_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, len, B, T, D, m, n, To, out,
                                                                                         out_len, vec4);
    ASR_LAUNCH_CHECK();
    return 0;
}

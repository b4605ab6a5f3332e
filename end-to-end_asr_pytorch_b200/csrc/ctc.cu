// ctc.cu - CTC loss with fused log-softmax, alpha-beta recursion and gradient (sm_100a).
//
// Replaces log_softmax + F.ctc_loss as the reference calls it
// (/root/reference/src/transformer/loss.py:39-43, src/ctcModel/loss.py:7-11):
// blank = V-1, 0-padded [B,S] int64 targets, reduction='mean', zero_infinity=False.
//
// Two kernels per call, logits are read ONCE and the gradient written ONCE:
//
//  K1 ctc_rows    one CTA per frame row (b,t): the row lives in registers; one
//                 pass gives the row log-sum-exp, the <= S+1 log-probs the lattice
//                 needs (gathered into a compact [B,T,S+1] table) and - when a
//                 gradient is wanted - the dense part of it, softmax * 1/(B*len).
//                 Pure streaming: 4V bytes read + 4V bytes written per frame.
//  K2 ctc_lattice one warp per utterance: log-space alpha sweep with warp-level
//                 logsumexp (each lane owns NS consecutive lattice states, one
//                 shuffle per step), checkpoints every K frames in shared memory,
//                 then a backward beta sweep that recomputes alpha block by block
//                 from the checkpoints.  The T x (2S+1) lattice never exists in
//                 HBM.  The sparse part of the gradient, -occupancy(t,c)/(B*len),
//                 is applied in place with fire-and-forget RED.ADD (one add per
//                 address: deterministic).
#include "common.cuh"

namespace asr {

__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }

struct CtcArgs {
    const float* logits;
    const int64_t* targets;
    const int* in_len;
    const int* tgt_len;
    int B, T, V, S, blank, SP;
    float* nll;
    float* g;     // may be null
    float* glp;   // [B,T,SP]: [0] = blank, [1+j] = label j
};

// ---------------------------------------------------------------------------------
// K1: one CTA per row.
// ---------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int i = 1; i < NT / 32; ++i) r = fmaxf(r, red[i]);
    __syncthreads();
    return r;
}
template <int NT>
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int i = 1; i < NT / 32; ++i) r += red[i];
    __syncthreads();
    return r;
}

template <int NT, int VPT, bool GRAD>
__global__ void __launch_bounds__(NT) ctc_rows_kernel(const CtcArgs a) {
    __shared__ float red[NT / 32];
    const int tid = threadIdx.x;
    const long long row = blockIdx.x;
    const int b = (int)(row / a.T);
    const int t = (int)(row - (long long)b * a.T);
    const int Tb = min(max(__ldg(a.in_len + b), 0), a.T);
    const int V = a.V;
    const float* x = a.logits + (size_t)row * V;
    float* g = GRAD ? a.g + (size_t)row * V : nullptr;

    // 16-byte alignment peel (rows are only 4-byte aligned when V is odd)
    int lead = (int)((4u - ((reinterpret_cast<uintptr_t>(x) >> 2) & 3u)) & 3u);
    if (lead > V) lead = V;
    const int nvec = (V - lead) >> 2;
    const int tail0 = lead + (nvec << 2);
    const int nscal = lead + (V - tail0);   // <= 6 scalar elements
    const int sidx = (tid < lead) ? tid : tail0 + (tid - lead);

    if (t >= Tb) {   // padded frame: gradient is exactly zero, nothing else to do
        if (GRAD) {
            float4* gv = reinterpret_cast<float4*>(g + lead);
            for (int i = tid; i < nvec; i += NT) gv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tid < nscal) g[sidx] = 0.0f;
        }
        return;
    }

    const float4* xv = reinterpret_cast<const float4*>(x + lead);
    float4 v[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        const int i = tid + j * NT;
        v[j] = (i < nvec) ? __ldg(xv + i) : make_float4(neg_inf(), neg_inf(), neg_inf(), neg_inf());
    }
    float xs = (tid < nscal) ? __ldg(x + sidx) : neg_inf();

    float m = xs;
#pragma unroll
    for (int j = 0; j < VPT; ++j) m = fmaxf(m, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
    const float M = block_max<NT>(m, red);
    const float Ms = (M == neg_inf()) ? 0.0f : M;

    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        v[j].x = __expf(v[j].x - Ms);
        v[j].y = __expf(v[j].y - Ms);
        v[j].z = __expf(v[j].z - Ms);
        v[j].w = __expf(v[j].w - Ms);
        s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    xs = __expf(xs - Ms);
    s += xs;
    const float Ssum = block_sum<NT>(s, red);
    const float lse = Ms + __logf(Ssum);

    // gather the log-probs the lattice needs: blank + this utterance's labels
    const int Sb = min(max(__ldg(a.tgt_len + b), 0), a.S);
    float* glp = a.glp + (size_t)row * a.SP;
    for (int j = tid; j <= Sb; j += NT) {
        int c = (j == 0) ? a.blank : (int)__ldg(a.targets + (size_t)b * a.S + (j - 1));
        c = min(max(c, 0), V - 1);
        glp[j] = __ldg(x + c) - lse;
    }

    if (GRAD) {
        const float coef = 1.0f / (Ssum * (float)a.B * (float)max(Sb, 1));
        float4* gv = reinterpret_cast<float4*>(g + lead);
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int i = tid + j * NT;
            if (i < nvec) __stcs(gv + i, make_float4(v[j].x * coef, v[j].y * coef, v[j].z * coef, v[j].w * coef));
        }
        if (tid < nscal) g[sidx] = xs * coef;
    }
}

// ---------------------------------------------------------------------------------
// K2: one warp per utterance.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float lse2(float a, float b) {
    const float m = fmaxf(a, b);
    const float ms = (m == neg_inf()) ? 0.0f : m;
    return ms + __logf(__expf(a - ms) + __expf(b - ms));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
    const float m = fmaxf(fmaxf(a, b), c);
    const float ms = (m == neg_inf()) ? 0.0f : m;
    return ms + __logf(__expf(a - ms) + __expf(b - ms) + __expf(c - ms));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int NS>
struct Lattice {
    static constexpr int NH = NS / 2;
    int lab[NH];
    int li[NH];     // index of this label's log-prob in a gathered row (0 when invalid)
    bool vl[NH];    // label state valid
    bool vb[NH];    // blank state valid
    bool skp[NH];   // label state may be entered from s-2
    bool skf[NH];   // label state may jump to s+2

    // alpha_t from alpha_{t-1}; row = gathered log-probs of frame t
    __device__ __forceinline__ void alpha_step(float (&al)[NS], const float* row, int lane) const {
        float x = __shfl_up_sync(0xffffffffu, al[NS - 1], 1);
        if (lane == 0) x = neg_inf();
        const float lpb = row[0];
        float nw[NS];
        nw[0] = vb[0] ? lse2(al[0], x) + lpb : neg_inf();
        nw[1] = vl[0] ? lse3(al[1], al[0], skp[0] ? x : neg_inf()) + row[li[0]] : neg_inf();
#pragma unroll
        for (int q = 1; q < NH; ++q) {
            nw[2 * q] = vb[q] ? lse2(al[2 * q], al[2 * q - 1]) + lpb : neg_inf();
            nw[2 * q + 1] = vl[q] ? lse3(al[2 * q + 1], al[2 * q], skp[q] ? al[2 * q - 1] : neg_inf()) + row[li[q]]
                                  : neg_inf();
        }
#pragma unroll
        for (int r = 0; r < NS; ++r) al[r] = nw[r];
    }
    __device__ __forceinline__ void alpha_init(float (&al)[NS], const float* row, int lane) const {
#pragma unroll
        for (int r = 0; r < NS; ++r) al[r] = neg_inf();
        if (lane == 0) {
            al[0] = row[0];
            if (vl[0]) al[1] = row[1];
        }
    }
    // beta_t from beta_{t+1}; row = gathered log-probs of frame t
    __device__ __forceinline__ void beta_step(float (&be)[NS], const float* row, int lane) const {
        float y0 = __shfl_down_sync(0xffffffffu, be[0], 1);
        float y1 = __shfl_down_sync(0xffffffffu, be[1], 1);
        if (lane == 31) {
            y0 = neg_inf();
            y1 = neg_inf();
        }
        const float lpb = row[0];
        float nw[NS];
#pragma unroll
        for (int q = 0; q < NH; ++q) {
            nw[2 * q] = vb[q] ? lse2(be[2 * q], be[2 * q + 1]) + lpb : neg_inf();
            const float n1 = (q < NH - 1) ? be[(2 * q + 2) % NS] : y0;
            const float n2 = (q < NH - 1) ? be[(2 * q + 3) % NS] : y1;
            nw[2 * q + 1] = vl[q] ? lse3(be[2 * q + 1], n1, skf[q] ? n2 : neg_inf()) + row[li[q]] : neg_inf();
        }
#pragma unroll
        for (int r = 0; r < NS; ++r) be[r] = nw[r];
    }
};

template <int NS>
__global__ void __launch_bounds__(32) ctc_lattice_kernel(const CtcArgs a, int K) {
    constexpr int NH = NS / 2;
    constexpr int NSL = 32 * NS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int b = blockIdx.x;
    const int T = a.T, SP = a.SP;
    const int Tb = min(max(__ldg(a.in_len + b), 0), T);
    const int Sb = min(max(__ldg(a.tgt_len + b), 0), a.S);
    const int nc_max = (T + K - 1) / K;

    // shared memory carve-up
    float* lpbuf0 = reinterpret_cast<float*>(smem_raw);
    float* lpbuf1 = lpbuf0 + (size_t)K * SP;
    float* ckpt = lpbuf1 + (size_t)K * SP;            // [nc_max][NSL]
    float* blk = ckpt + (size_t)nc_max * NSL;         // [K][NSL]
    float* blpart = blk + (size_t)K * NSL;            // [K][33]
    float* occ = blpart + (size_t)K * 33;             // [32*NH]
    int* dupn = reinterpret_cast<int*>(occ + 32 * NH);   // [32*NH]
    int* tgt = dupn + 32 * NH;                        // [32*NH]

    // ---- per-lane lattice description ------------------------------------------
    for (int j = lane; j < 32 * NH; j += 32) tgt[j] = (j < Sb) ? (int)__ldg(a.targets + (size_t)b * a.S + j) : -1;
    __syncwarp();
    Lattice<NS> lat;
    bool leader[NH];
    int any_dup = 0;
#pragma unroll
    for (int q = 0; q < NH; ++q) {
        const int j = lane * NH + q;
        lat.vl[q] = j < Sb;
        lat.vb[q] = j <= Sb;
        lat.lab[q] = lat.vl[q] ? min(max(tgt[j], 0), a.V - 1) : 0;
        lat.li[q] = lat.vl[q] ? 1 + j : 0;
        lat.skp[q] = lat.vl[q] && j > 0 && tgt[j] != tgt[j - 1];
        lat.skf[q] = lat.vl[q] && (j + 1 < Sb) && tgt[j + 1] != tgt[j];
        leader[q] = lat.vl[q];
        int nxt = -1;
        if (lat.vl[q]) {
            for (int jj = 0; jj < j; ++jj)
                if (tgt[jj] == tgt[j]) leader[q] = false;
            for (int jj = Sb - 1; jj > j; --jj)
                if (tgt[jj] == tgt[j]) nxt = jj;
        }
        dupn[j] = nxt;
        if (nxt >= 0) any_dup = 1;
    }
    any_dup = __any_sync(0xffffffffu, any_dup);
    __syncwarp();

    if (Tb == 0) {   // no frames: nll = 0 for an empty target, +inf otherwise (ATen)
        if (lane == 0) a.nll[b] = (Sb == 0) ? 0.0f : -neg_inf();
        return;
    }

    const float* glp_b = a.glp + (size_t)b * T * SP;
    const int nc = (Tb + K - 1) / K;

    auto load_chunk = [&](int c, float* dst) {
        const int t0 = c * K;
        const int n = min(K, Tb - t0);
        const float* src = glp_b + (size_t)t0 * SP;
        const int pieces = (n * SP) >> 2;
        for (int i = lane; i < pieces; i += 32) cp_async16(dst + 4 * i, src + 4 * i);
        cp_async_commit();
    };

    // ---- sweep 1: alpha, checkpoint at the end of every chunk --------------------
    float al[NS];
    load_chunk(0, lpbuf0);
    for (int c = 0; c < nc; ++c) {
        float* cur = (c & 1) ? lpbuf1 : lpbuf0;
        if (c + 1 < nc) {
            load_chunk(c + 1, (c & 1) ? lpbuf0 : lpbuf1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const int n = min(K, Tb - c * K);
        for (int i = 0; i < n; ++i) {
            const float* row = cur + i * SP;
            if (c == 0 && i == 0)
                lat.alpha_init(al, row, lane);
            else
                lat.alpha_step(al, row, lane);
        }
#pragma unroll
        for (int r = 0; r < NS; ++r) ckpt[(size_t)c * NSL + lane * NS + r] = al[r];
        __syncwarp();
    }
    // nll = -LSE(alpha_{T-1}(2S), alpha_{T-1}(2S-1))
    float nll;
    {
        const float* fin = ckpt + (size_t)(nc - 1) * NSL;
        const float a_end = fin[2 * Sb];
        const float a_lab = (Sb > 0) ? fin[2 * Sb - 1] : neg_inf();
        nll = -lse2(a_end, a_lab);
        if (lane == 0) a.nll[b] = nll;
    }
    if (a.g == nullptr) return;

    float* g_b = a.g + (size_t)b * T * a.V;
    if (!(nll < -neg_inf())) {
        // infeasible alignment (or NaN input): the reference's gradient is NaN on every
        // valid frame row (log_softmax backward spreads the NaN), zero_infinity=False
        const float qnan = __int_as_float(0x7fc00000);
        const size_t n = (size_t)Tb * a.V;
        for (size_t i = lane; i < n; i += 32) g_b[i] = qnan;
        return;
    }
    const float scale = 1.0f / ((float)a.B * (float)max(Sb, 1));

    // ---- sweep 2: beta backwards, alpha recomputed per chunk ----------------------
    float be[NS];
#pragma unroll
    for (int r = 0; r < NS; ++r) be[r] = neg_inf();
    load_chunk(nc - 1, ((nc - 1) & 1) ? lpbuf1 : lpbuf0);
    for (int c = nc - 1; c >= 0; --c) {
        float* cur = (c & 1) ? lpbuf1 : lpbuf0;
        if (c > 0) {
            load_chunk(c - 1, ((c - 1) & 1) ? lpbuf1 : lpbuf0);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const int t0 = c * K;
        const int n = min(K, Tb - t0);
        // recompute alpha for the chunk
        if (c > 0) {
#pragma unroll
            for (int r = 0; r < NS; ++r) al[r] = ckpt[(size_t)(c - 1) * NSL + lane * NS + r];
        }
        for (int i = 0; i < n; ++i) {
            const float* row = cur + i * SP;
            if (c == 0 && i == 0)
                lat.alpha_init(al, row, lane);
            else
                lat.alpha_step(al, row, lane);
#pragma unroll
            for (int r = 0; r < NS; ++r) blk[(size_t)i * NSL + lane * NS + r] = al[r];
        }
        // beta + occupancy
        for (int i = n - 1; i >= 0; --i) {
            const int t = t0 + i;
            const float* row = cur + i * SP;
            if (t == Tb - 1) {
#pragma unroll
                for (int q = 0; q < NH; ++q) {
                    const int j = lane * NH + q;
                    be[2 * q] = (j == Sb) ? row[0] : neg_inf();
                    be[2 * q + 1] = (j == Sb - 1) ? row[lat.li[q]] : neg_inf();
                }
            } else {
                lat.beta_step(be, row, lane);
            }
            const float lpb = row[0];
            float bsum = 0.0f;
            float ov[NH];
#pragma unroll
            for (int q = 0; q < NH; ++q) {
                const float ab = blk[(size_t)i * NSL + lane * NS + 2 * q] + be[2 * q];
                if (lat.vb[q]) bsum += __expf(ab - lpb + nll);
                const float al_l = blk[(size_t)i * NSL + lane * NS + 2 * q + 1] + be[2 * q + 1];
                ov[q] = lat.vl[q] ? __expf(al_l - row[lat.li[q]] + nll) : 0.0f;
            }
            blpart[i * 33 + lane] = bsum;
            float* grow = g_b + (size_t)t * a.V;
            if (any_dup) {
#pragma unroll
                for (int q = 0; q < NH; ++q) occ[lane * NH + q] = ov[q];
                __syncwarp();
#pragma unroll
                for (int q = 0; q < NH; ++q) {
                    if (leader[q]) {
                        float vsum = ov[q];
                        for (int jj = dupn[lane * NH + q]; jj >= 0; jj = dupn[jj]) vsum += occ[jj];
                        atomicAdd(grow + lat.lab[q], -vsum * scale);
                    }
                }
                __syncwarp();
            } else {
#pragma unroll
                for (int q = 0; q < NH; ++q)
                    if (lat.vl[q]) atomicAdd(grow + lat.lab[q], -ov[q] * scale);
            }
        }
        __syncwarp();
        // blank column: one lane per frame of the chunk sums the 32 partials
        for (int i = lane; i < n; i += 32) {
            float sacc = 0.0f;
#pragma unroll 8
            for (int l = 0; l < 32; ++l) sacc += blpart[i * 33 + l];
            atomicAdd(g_b + (size_t)(t0 + i) * a.V + a.blank, -sacc * scale);
        }
        __syncwarp();
    }
}

static size_t lattice_smem_bytes(int NS, int K, int T, int SP) {
    const int NH = NS / 2, NSL = 32 * NS;
    const size_t nc = (size_t)(T + K - 1) / K;
    size_t f = 2 * (size_t)K * SP + nc * NSL + (size_t)K * NSL + (size_t)K * 33 + 32 * NH;
    return f * 4 + 2 * (size_t)32 * NH * 4;
}

// ---------------------------------------------------------------------------------
// g *= *scale (skipped on the device when the scale is exactly 1)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scale_inplace_kernel(float* g, size_t n, const float* scale_dev) {
    const float s = __ldg(scale_dev);
    if (s == 1.0f) return;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) g[i] *= s;
}

}  // namespace asr

using namespace asr;

static inline int round_up4(int x) { return (x + 3) & ~3; }

extern "C" size_t asr_ctc_workspace_bytes(int B, int T, int V, int S) {
    (void)V;
    if (B <= 0 || T <= 0 || S < 0) return 0;
    return (size_t)B * T * round_up4(S + 1) * sizeof(float) + 256;
}

template <int NT, bool GRAD>
static int launch_rows(const CtcArgs& a, int vpt, long long rows, cudaStream_t st) {
    const unsigned grid = (unsigned)rows;
    if (vpt <= 1)
        ctc_rows_kernel<NT, 1, GRAD><<<grid, NT, 0, st>>>(a);
    else if (vpt <= 3)
        ctc_rows_kernel<NT, 3, GRAD><<<grid, NT, 0, st>>>(a);
    else if (vpt <= 9)
        ctc_rows_kernel<NT, 9, GRAD><<<grid, NT, 0, st>>>(a);
    else
        ctc_rows_kernel<NT, 16, GRAD><<<grid, NT, 0, st>>>(a);
    ASR_LAUNCH_CHECK();
    return 0;
}

template <bool GRAD>
static int launch_rows_nt(const CtcArgs& a, long long rows, cudaStream_t st) {
    const int nvec_max = a.V / 4 + 1;
    int nt = 128;
    while (nt < 1024 && nvec_max > nt * 16) nt <<= 1;
    const int vpt = (nvec_max + nt - 1) / nt;
    switch (nt) {
        case 128: return launch_rows<128, GRAD>(a, vpt, rows, st);
        case 256: return launch_rows<256, GRAD>(a, vpt, rows, st);
        case 512: return launch_rows<512, GRAD>(a, vpt, rows, st);
        default: return launch_rows<1024, GRAD>(a, vpt, rows, st);
    }
}

template <int NS>
static int launch_lattice(const CtcArgs& a, cudaStream_t st) {
    int K = 32;
    size_t smem = lattice_smem_bytes(NS, K, a.T, a.SP);
    while (smem > 200 * 1024 && K < 256) {
        K <<= 1;
        smem = lattice_smem_bytes(NS, K, a.T, a.SP);
    }
    ASR_REQUIRE(smem <= 227 * 1024, "asr_ctc: T=%d S=%d needs %zu bytes of shared memory for the lattice (max 232448)",
                a.T, a.S, smem);
    ASR_CHECK_CUDA(cudaFuncSetAttribute(ctc_lattice_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctc_lattice_kernel<NS><<<a.B, 32, smem, st>>>(a, K);
    ASR_LAUNCH_CHECK();
    return 0;
}

extern "C" int asr_ctc_fwd_bwd_f32(const float* logits, const int64_t* targets, const int* in_len, const int* tgt_len,
                                   int B, int T, int V, int S, int blank, float* nll, float* g_logits, void* ws,
                                   size_t ws_bytes, void* stream) {
    ASR_REQUIRE(B > 0 && T > 0 && V > 1 && S >= 0, "asr_ctc_fwd_bwd_f32: bad shape B=%d T=%d V=%d S=%d", B, T, V, S);
    ASR_REQUIRE(logits && in_len && tgt_len && nll && ws && (S == 0 || targets), "asr_ctc_fwd_bwd_f32: null pointer");
    ASR_REQUIRE(blank >= 0 && blank < V, "asr_ctc_fwd_bwd_f32: blank %d out of range", blank);
    ASR_REQUIRE(ws_bytes >= asr_ctc_workspace_bytes(B, T, V, S), "asr_ctc_fwd_bwd_f32: workspace too small (%zu < %zu)",
                ws_bytes, asr_ctc_workspace_bytes(B, T, V, S));
    ASR_REQUIRE(V <= 65000, "asr_ctc_fwd_bwd_f32: V=%d > 65000 not supported", V);
    ASR_REQUIRE(2 * S + 1 <= 32 * 16, "asr_ctc_fwd_bwd_f32: S=%d > 255 labels not supported", S);
    ASR_REQUIRE((long long)B * T < (1ll << 31) - 1, "asr_ctc_fwd_bwd_f32: B*T too large");
    if (asr_device_ok() != 0) return 3;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    CtcArgs a;
    a.logits = logits;
    a.targets = targets;
    a.in_len = in_len;
    a.tgt_len = tgt_len;
    a.B = B; a.T = T; a.V = V; a.S = S; a.blank = blank;
    a.SP = round_up4(S + 1);
    a.nll = nll;
    a.g = g_logits;
    uintptr_t w = (reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255;
    a.glp = reinterpret_cast<float*>(w);

    const long long rows = (long long)B * T;
    int rc = g_logits ? launch_rows_nt<true>(a, rows, st) : launch_rows_nt<false>(a, rows, st);
    if (rc != 0) return rc;

    const int states = 2 * S + 1;
    if (states <= 64) return launch_lattice<2>(a, st);
    if (states <= 128) return launch_lattice<4>(a, st);
    if (states <= 192) return launch_lattice<6>(a, st);
    if (states <= 256) return launch_lattice<8>(a, st);
    if (states <= 384) return launch_lattice<12>(a, st);
    return launch_lattice<16>(a, st);
}

extern "C" int asr_scale_inplace_f32(float* g, size_t n, const float* scale_dev, void* stream) {
    ASR_REQUIRE(g && scale_dev, "asr_scale_inplace_f32: null pointer");
    if (n == 0) return 0;
    if (asr_device_ok() != 0) return 3;
    size_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    const size_t cap = (size_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    scale_inplace_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(g, n, scale_dev);
    ASR_LAUNCH_CHECK();
    return 0;
}

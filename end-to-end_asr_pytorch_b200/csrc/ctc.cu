// ctc.cu - CTC loss with fused log-softmax, alpha-beta recursion and gradient (sm_100a).
//
// Replaces log_softmax + F.ctc_loss as the reference calls it
// (/root/reference/src/transformer/loss.py:39-43, src/ctcModel/loss.py:7-11):
// blank = V-1, 0-padded [B,S] int64 targets, reduction='mean', zero_infinity=False.
//
// Three kernels per call; logits are read ONCE and the gradient written ONCE:
//
//  K1 ctc_rows     one CTA per frame row (b,t): the row lives in registers; one pass
//                  gives the row log-sum-exp, the <= S+1 log-probs the lattice needs
//                  (gathered into a compact [B,T,S+3] table, log2 domain, relative to the frame's blank) and - when a
//                  gradient is wanted - the dense part of it, softmax * 1/(B*len).
//                  Pure streaming: 4V bytes read + 4V bytes written per frame.
//  K2 ctc_lattice  one warp per utterance: log-space (base 2) alpha sweep with
//                  warp-level logsumexp - each lane owns NS consecutive lattice
//                  states, one shuffle per step, 2-3 MUFU ops per state, no branches
//                  - with checkpoints every K frames in shared memory, then a
//                  backward beta sweep that recomputes alpha block by block from the
//                  checkpoints.  The T x (2S+1) lattice never exists in HBM.  The
//                  per-frame label occupancies overwrite the gathered table in place.
//  K3 ctc_apply    one warp per frame row: g[t,c] -= occupancy(t,c)/(B*len) for the
//                  <= S+1 classes of the utterance (repeated labels merged in a fixed
//                  order): every address is touched once, deterministic, no atomics.
//
// Design notes from measurement on B200: a linear-domain lattice in
// fp64 ran 2.7x slower than this MUFU version (dependent DADD/DMUL chains), and an
// fp32 linear-domain lattice with per-frame power-of-two rescaling was faster but
// loses probability mass: the dynamic range ACROSS states at one frame reaches
// 2^-400 for T=1600, S=80, far beyond fp32 exponents.  Log space is the robust choice.
#include "common.cuh"

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>

namespace asr {

__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }

// Large finite stand-in for log(0): every lattice step stays branch-free (max / min /
// MUFU), unreachable states simply sit near kNeg and exp2 of anything that far down is 0.
constexpr float kNeg = -1.0e30f;

struct CtcArgs {
    const float* logits;
    const int64_t* targets;
    const int* in_len;
    const int* tgt_len;
    int B, T, V, S, blank, SP;
    int ld;       // row stride of logits / g in floats (>= V)
    int Bn;       // batch size the mean loss is normalised by (>= B when this launch covers a slice of the batch)
    float* nll;
    float* g;     // may be null
    float* glp;   // [B,T,SP] log2-probabilities: [0] = blank, [1+j] = label j, [S+1] = kNeg
    int* dlink;   // [B,S] repeated-label links: (next occurrence + 1) | (has earlier occurrence << 30)
    int fuse_apply;   // 1: the lattice kernel applies the sparse update itself (RED.ADD), K3 is not launched
    float* ckpt;      // [B][ckpt_stride] checkpoints of the bidirectional lattice (lane-private slots)
    size_t ckpt_stride;
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
    const float* x = a.logits + (size_t)row * a.ld;
    float* g = GRAD ? a.g + (size_t)row * a.ld : nullptr;

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
    // The lattice only needs each frame's probabilities up to a common factor (it cancels in
    // alpha*beta/(p*likelihood)), so the row is stored relative to the frame's blank: the blank
    // entry is exactly 0 and the recursions stay small in magnitude - their fp32 rounding error
    // scales with |alpha|, which otherwise grows by log2(V) bits per frame.  The factor itself
    // goes to its own slot and is added back to the log-likelihood by the lattice kernel.
    const float xb = __ldg(x + a.blank);
    const bool rel_ok = (xb - lse) >= -1.0e4f;       // a (nearly) impossible blank: fall back to a fixed offset
    const float off = rel_ok ? (xb - lse) : -1.0e4f;
    for (int j = tid; j <= Sb; j += NT) {
        int c = (j == 0) ? a.blank : (int)__ldg(a.targets + (size_t)b * a.S + (j - 1));
        c = min(max(c, 0), V - 1);
        const float xc = __ldg(x + c);
        glp[j] = (rel_ok ? (xc - xb) : ((xc - lse) + 1.0e4f)) * 1.4426950408889634f;   // log2 domain for the lattice
    }
    if (tid == NT - 1) {
        glp[a.S + 1] = kNeg;                          // "impossible" slot read by out-of-range lattice states
        glp[a.S + 2] = off * 1.4426950408889634f;     // log2 of the frame's common factor
    }

    if (GRAD) {
        // gradient written over the logits (asr_ctc_fwd_bwd_ld_f32 with g == logits): the gather above reads scattered
        // elements of the row, which another thread of the CTA must not have overwritten yet
        if (a.g == a.logits) __syncthreads();
        const float coef = 1.0f / (Ssum * (float)a.Bn * (float)max(Sb, 1));
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
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2f(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// log2(2^a + 2^b): one EX2 + one LG2
__device__ __forceinline__ float lse2(float a, float b) {
    const float m = fmaxf(a, b);
    const float d = fminf(a, b) - m;
    return m + lg2f(1.0f + ex2f(d));
}
// log2(2^a + 2^b + 2^c): two EX2 + one LG2 (the largest term contributes exactly 1)
__device__ __forceinline__ float lse3(float a, float b, float c) {
    const float hi = fmaxf(a, b), lo2 = fminf(a, b);
    const float m = fmaxf(hi, c);
    const float lo = fminf(lo2, c);
    const float mid = fmaxf(lo2, fminf(hi, c));
    return m + lg2f((1.0f + ex2f(mid - m)) + ex2f(lo - m));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Lane l owns states s = l*NS .. l*NS+NS-1 (even s = blank, odd s = label (s-1)/2).
template <int NS>
struct Lattice {
    static constexpr int NH = NS / 2;
    int li[NH];     // index of this label's log-prob in a gathered row (the kNeg slot when invalid)
    int bi[NH];     // index of the blank log-prob (the kNeg slot when the blank state is invalid)
    bool skp[NH];   // label state may be entered from s-2
    bool skf[NH];   // label state may jump to s+2

    // alpha_t from alpha_{t-1}; row = gathered log2-probs of frame t
    __device__ __forceinline__ void alpha_step(float (&al)[NS], const float* row, int lane) const {
        float x = __shfl_up_sync(0xffffffffu, al[NS - 1], 1);
        if (lane == 0) x = kNeg;
        float nw[NS];
        nw[0] = lse2(al[0], x) + row[bi[0]];
        nw[1] = lse3(al[1], al[0], skp[0] ? x : kNeg) + row[li[0]];
#pragma unroll
        for (int q = 1; q < NH; ++q) {
            nw[2 * q] = lse2(al[2 * q], al[2 * q - 1]) + row[bi[q]];
            nw[2 * q + 1] = lse3(al[2 * q + 1], al[2 * q], skp[q] ? al[2 * q - 1] : kNeg) + row[li[q]];
        }
#pragma unroll
        for (int r = 0; r < NS; ++r) al[r] = nw[r];
    }
    __device__ __forceinline__ void alpha_init(float (&al)[NS], const float* row, int lane) const {
#pragma unroll
        for (int r = 0; r < NS; ++r) al[r] = kNeg;
        if (lane == 0) {
            al[0] = row[bi[0]];
            al[1] = row[li[0]];
        }
    }
    // beta_t from beta_{t+1}; row = gathered log2-probs of frame t
    __device__ __forceinline__ void beta_step(float (&be)[NS], const float* row, int lane) const {
        float y0 = __shfl_down_sync(0xffffffffu, be[0], 1);
        float y1 = __shfl_down_sync(0xffffffffu, be[1], 1);
        if (lane == 31) {
            y0 = kNeg;
            y1 = kNeg;
        }
        float nw[NS];
#pragma unroll
        for (int q = 0; q < NH; ++q) {
            nw[2 * q] = lse2(be[2 * q], be[2 * q + 1]) + row[bi[q]];
            const float n1 = (q < NH - 1) ? be[(2 * q + 2) % NS] : y0;
            const float n2 = (q < NH - 1) ? be[(2 * q + 3) % NS] : y1;
            nw[2 * q + 1] = lse3(be[2 * q + 1], n1, skf[q] ? n2 : kNeg) + row[li[q]];
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
    const int ZI = a.S + 1;   // slot of every row that holds kNeg
    const int Tb = min(max(__ldg(a.in_len + b), 0), T);
    const int Sb = min(max(__ldg(a.tgt_len + b), 0), a.S);
    const int nc_max = (T + K - 1) / K;

    // shared memory carve-up
    float* lpbuf0 = reinterpret_cast<float*>(smem_raw);   // [K][SP] chunk of the gathered table
    float* lpbuf1 = lpbuf0 + (size_t)K * SP;
    float* ckpt = lpbuf1 + (size_t)K * SP;                // [nc_max][NSL] alpha checkpoints
    float* blk = ckpt + (size_t)nc_max * NSL;             // [K][NSL] alpha of the block
    float* blpart = blk + (size_t)K * NSL;                // [K][33] per-lane blank occupancy partials
    int* tgt = reinterpret_cast<int*>(blpart + (size_t)K * 33);   // [32*NH]
    int* dupn = tgt + 32 * NH;                            // [32*NH] next occurrence of the same class, -1 = none
    float* occ = reinterpret_cast<float*>(dupn + 32 * NH);   // [32*NH] per-frame label occupancies (repeat merging)

    // ---- per-lane lattice description ------------------------------------------
    // A label outside [0,V) or equal to the blank has no lattice of its own (its gradient entry would collide
    // with the blank's in K3): the utterance is reported as NaN loss / NaN gradient instead of a silently wrong one.
    int bad_label = 0;
    for (int j = lane; j < 32 * NH; j += 32) {
        const long long c = (j < Sb) ? (long long)__ldg(a.targets + (size_t)b * a.S + j) : -1;
        if (j < Sb && (c < 0 || c >= a.V || c == a.blank)) bad_label = 1;
        tgt[j] = (int)c;
    }
    bad_label = __any_sync(0xffffffffu, bad_label);
    __syncwarp();
    Lattice<NS> lat;
    bool vl[NH], vb[NH], leader[NH];
    int lab[NH];
    int any_dup = 0;
#pragma unroll
    for (int q = 0; q < NH; ++q) {
        const int j = lane * NH + q;
        vl[q] = j < Sb;
        vb[q] = j <= Sb;
        lat.li[q] = vl[q] ? 1 + j : ZI;
        lat.bi[q] = vb[q] ? 0 : ZI;
        lat.skp[q] = vl[q] && j > 0 && tgt[j] != tgt[j - 1];
        lat.skf[q] = vl[q] && (j + 1 < Sb) && tgt[j + 1] != tgt[j];
        lab[q] = vl[q] ? min(max(tgt[j], 0), a.V - 1) : 0;
        leader[q] = vl[q];
        int nxt = -1;
        if (vl[q] && a.g != nullptr && a.fuse_apply) {
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
    // repeated-label links for K3 (it merges the occupancies of a class that occurs more than once)
    if (a.g != nullptr && !a.fuse_apply) {
        for (int j = lane; j < Sb; j += 32) {
            int nxt = -1, earlier = 0;
            for (int jj = 0; jj < j; ++jj) earlier |= (tgt[jj] == tgt[j]);
            for (int jj = Sb - 1; jj > j; --jj)
                if (tgt[jj] == tgt[j]) nxt = jj;
            a.dlink[(size_t)b * a.S + j] = (nxt + 1) | (earlier << 30);
        }
    }

    if (Tb == 0) {   // no frames: nll = 0 for an empty target, +inf otherwise (ATen)
        if (lane == 0) a.nll[b] = (Sb == 0) ? 0.0f : -neg_inf();
        return;
    }

    float* glp_b = a.glp + (size_t)b * T * SP;
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
    float off_acc = 0.0f;
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
        for (int i = lane; i < n; i += 32) off_acc += cur[i * SP + ZI + 1];   // the frames' common factors
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
    // nll = -LSE(alpha_{T-1}(2S), alpha_{T-1}(2S-1));  nll2 is the same in log2 units
    float nll2;
    bool feasible;
    {
        const float* fin = ckpt + (size_t)(nc - 1) * NSL;
        const float a_end = fin[2 * Sb];
        const float a_lab = (Sb > 0) ? fin[2 * Sb - 1] : kNeg;
        const float ll2 = lse2(a_end, a_lab);       // of the rows relative to their blanks
        feasible = ll2 > -1.0e29f && !bad_label;
        nll2 = -ll2;
        const float off = warp_sum(off_acc);
        if (lane == 0) a.nll[b] = bad_label ? __int_as_float(0x7fc00000) : feasible ? -(ll2 + off) * 0.6931471805599453f : -neg_inf();
    }
    if (a.g == nullptr) return;

    if (!feasible) {
        // infeasible alignment (or NaN input): the reference's gradient is NaN on every
        // valid frame row (log_softmax backward spreads the NaN), zero_infinity=False
        float* g_b = a.g + (size_t)b * T * a.ld;
        const float qnan = __int_as_float(0x7fc00000);
        for (int t = 0; t < Tb; ++t)
            for (int i = lane; i < a.V; i += 32) g_b[(size_t)t * a.ld + i] = qnan;
        for (size_t i = lane; i < (size_t)Tb * SP; i += 32) glp_b[i] = 0.0f;   // nothing for K3 to apply
        return;
    }

    // ---- sweep 2: beta backwards, alpha recomputed per chunk, occupancy out -------
    float be[NS];
#pragma unroll
    for (int r = 0; r < NS; ++r) be[r] = kNeg;
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
        // beta (sequential): alpha + beta replaces alpha in the block buffer; the occupancies are
        // computed afterwards in a loop whose iterations are independent (off the critical path)
        for (int i = n - 1; i >= 0; --i) {
            const int t = t0 + i;
            const float* row = cur + i * SP;
            if (t == Tb - 1) {
#pragma unroll
                for (int q = 0; q < NH; ++q) {
                    const int j = lane * NH + q;
                    be[2 * q] = (j == Sb) ? row[0] : kNeg;
                    be[2 * q + 1] = (j == Sb - 1) ? row[lat.li[q]] : kNeg;
                }
            } else {
                lat.beta_step(be, row, lane);
            }
            float* ab = blk + (size_t)i * NSL + lane * NS;
#pragma unroll
            for (int r = 0; r < NS; ++r) ab[r] += be[r];
        }
        // occupancy(t,s) = exp2(alpha + beta - lp + nll); both alpha and beta include lp_t
        if (a.fuse_apply) {
            // apply g[t,c] -= occupancy/(B*len) right here: one fire-and-forget RED.ADD per touched
            // class and frame (repeats merged by their first occurrence -> one addend per address,
            // deterministic); occupancies below 1e-12 cannot change the fp32 gradient and are skipped
            const float scale = 1.0f / ((float)a.Bn * (float)max(Sb, 1));
            float* g_b = a.g + (size_t)b * T * a.ld;
            if (!any_dup) {
#pragma unroll 4
                for (int i = 0; i < n; ++i) {
                    const float* row = cur + i * SP;
                    const float* ab = blk + (size_t)i * NSL + lane * NS;
                    const float lpb = row[0];
                    float bsum = 0.0f;
                    float* grow = g_b + (size_t)(t0 + i) * a.ld;
#pragma unroll
                    for (int q = 0; q < NH; ++q) {
                        bsum += vb[q] ? ex2f((ab[2 * q] - lpb) + nll2) : 0.0f;
                        const float ov = ex2f((ab[2 * q + 1] - row[lat.li[q]]) + nll2);
                        if (vl[q] && !(ov < 1.0e-12f)) atomicAdd(grow + lab[q], -ov * scale);
                    }
                    blpart[i * 33 + lane] = bsum;
                }
            } else {
                for (int i = 0; i < n; ++i) {
                    const float* row = cur + i * SP;
                    const float* ab = blk + (size_t)i * NSL + lane * NS;
                    const float lpb = row[0];
                    float bsum = 0.0f;
                    float ov[NH];
#pragma unroll
                    for (int q = 0; q < NH; ++q) {
                        bsum += vb[q] ? ex2f((ab[2 * q] - lpb) + nll2) : 0.0f;
                        ov[q] = vl[q] ? ex2f((ab[2 * q + 1] - row[lat.li[q]]) + nll2) : 0.0f;
                        occ[lane * NH + q] = ov[q];
                    }
                    blpart[i * 33 + lane] = bsum;
                    __syncwarp();
                    float* grow = g_b + (size_t)(t0 + i) * a.ld;
#pragma unroll
                    for (int q = 0; q < NH; ++q) {
                        if (leader[q]) {
                            float vsum = ov[q];
                            for (int jj = dupn[lane * NH + q]; jj >= 0; jj = dupn[jj]) vsum += occ[jj];
                            if (!(vsum < 1.0e-12f)) atomicAdd(grow + lab[q], -vsum * scale);
                        }
                    }
                    __syncwarp();
                }
            }
            __syncwarp();
            for (int i = lane; i < n; i += 32) {
                float sacc = 0.0f;
#pragma unroll 8
                for (int l = 0; l < 32; ++l) sacc += blpart[i * 33 + l];
                if (!(sacc < 1.0e-12f)) atomicAdd(g_b + (size_t)(t0 + i) * a.ld + a.blank, -sacc * scale);
            }
            __syncwarp();
        } else {
#pragma unroll 4
            for (int i = 0; i < n; ++i) {
                const float* row = cur + i * SP;
                const float* ab = blk + (size_t)i * NSL + lane * NS;
                const float lpb = row[0];
                float bsum = 0.0f;
                float* orow = glp_b + (size_t)(t0 + i) * SP;
#pragma unroll
                for (int q = 0; q < NH; ++q) {
                    bsum += vb[q] ? ex2f((ab[2 * q] - lpb) + nll2) : 0.0f;
                    const float ov = ex2f((ab[2 * q + 1] - row[lat.li[q]]) + nll2);
                    if (vl[q]) orow[1 + lane * NH + q] = ov;   // per label position; K3 merges repeats
                }
                blpart[i * 33 + lane] = bsum;
            }
            __syncwarp();
            // blank column: one lane per frame of the chunk sums the 32 partials
            for (int i = lane; i < n; i += 32) {
                float sacc = 0.0f;
#pragma unroll 8
                for (int l = 0; l < 32; ++l) sacc += blpart[i * 33 + l];
                glp_b[(size_t)(t0 + i) * SP] = sacc;
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------
// K2, bidirectional variant: the two recursions meet in the middle, four warps per utterance.
//   phase 1   warp 0: alpha over frames [0,Tm)        warp 1: beta over frames [Tb-1 .. Tm]
//             (both keep a checkpoint per K frames in the global workspace)
//   midpoint  alpha(Tm-1) and beta(Tm) are exchanged through shared memory; the likelihood is
//             sum_s alpha(Tm-1,s) beta(Tm-1,s) / p(Tm-1,s)  (any frame's cut gives it)
//   phase 2   warp 0: beta continues down through [0,Tm)    warp 2: recomputes alpha, one block ahead
//             warp 1: alpha continues up through [Tm,Tb)     warp 3: recomputes beta, one block ahead
// The dependent chain is half as long as in the one-warp kernel, and in phase 2 the
// recomputation runs concurrently on its own warp: the hand-over of a block of K frames is a
// hardware named barrier (bar.arrive / bar.sync, nobody spins), the blocks live in a two-slot
// ring, the gathered rows in a three-slot ring filled with cp.async by the recomputing warp.
// Boundary conditions are "virtual" frames so that every frame is a regular step:
//   alpha(-1) = delta(s = 0),  beta(Tb) = delta(s = 2*Sb)   (log2 domain: 0 and kNeg).
// ---------------------------------------------------------------------------------
struct LatticeMid {
    float nll2;
    int feasible;
    float off_b;   // sum of the common factors of the frames of the second half
};
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <int NS>
__global__ void __launch_bounds__(128) ctc_lattice_mitm_kernel(const CtcArgs a, int K) {
    constexpr int NH = NS / 2;
    constexpr int NSL = 32 * NS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int half = warp & 1;        // 0: frames [0,Tm), 1: frames [Tm,Tb)
    const int helper = warp >> 1;     // 1: the recomputing warp of the half (phase 2 only)
    const int b = blockIdx.x;
    const int T = a.T, SP = a.SP;
    const int ZI = a.S + 1;
    const int Tb = min(max(__ldg(a.in_len + b), 0), T);
    const int Sb = min(max(__ldg(a.tgt_len + b), 0), a.S);

    const size_t per_half = (size_t)3 * K * SP + (size_t)2 * K * NSL + (size_t)3 * NSL + (size_t)K * 33;
    float* mine = reinterpret_cast<float*>(smem_raw) + half * per_half;
    float* rows_ring = mine;                               // [3][K][SP] gathered rows, block c in slot c % 3
    float* blk_ring = rows_ring + (size_t)3 * K * SP;      // [2][K][NSL] recomputed recursion, block c in slot c & 1
    float* ck_ring = blk_ring + (size_t)2 * K * NSL;       // [3][NSL] checkpoint that starts the recomputation of block c
    float* blpart = ck_ring + (size_t)3 * NSL;             // [K][33]
    float* xchg = reinterpret_cast<float*>(smem_raw) + 2 * per_half;   // [2][NSL]: alpha(Tm-1), beta(Tm)
    int* tgt = reinterpret_cast<int*>(xchg + 2 * NSL);     // [32*NH]
    LatticeMid* mid = reinterpret_cast<LatticeMid*>(tgt + 32 * NH);
    auto rows_of = [&](int c) { return rows_ring + (size_t)(c % 3) * K * SP; };
    auto blk_of = [&](int c) { return blk_ring + (size_t)(c & 1) * K * NSL; };
    auto ck_of = [&](int c) { return ck_ring + (size_t)(c % 3) * NSL; };
    // named barriers 1..8 (0 is __syncthreads): per half, "block slot filled" and "block slot free"
    auto bar_full = [&](int c) { return 1 + 4 * half + (c & 1); };
    auto bar_free = [&](int c) { return 3 + 4 * half + (c & 1); };

    int bad_label = 0;      // see ctc_lattice_kernel
    for (int j = threadIdx.x; j < 32 * NH; j += 128) {
        const long long c = (j < Sb) ? (long long)__ldg(a.targets + (size_t)b * a.S + j) : -1;
        if (j < Sb && (c < 0 || c >= a.V || c == a.blank)) bad_label = 1;
        tgt[j] = (int)c;
    }
    bad_label = __syncthreads_or(bad_label);
    Lattice<NS> lat;
    bool vl[NH], vb[NH];
#pragma unroll
    for (int q = 0; q < NH; ++q) {
        const int j = lane * NH + q;
        vl[q] = j < Sb;
        vb[q] = j <= Sb;
        lat.li[q] = vl[q] ? 1 + j : ZI;
        lat.bi[q] = vb[q] ? 0 : ZI;
        lat.skp[q] = vl[q] && j > 0 && tgt[j] != tgt[j - 1];
        lat.skf[q] = vl[q] && (j + 1 < Sb) && tgt[j + 1] != tgt[j];
    }
    if (warp == 3 && a.g != nullptr) {   // repeated-label links for K3
        for (int j = lane; j < Sb; j += 32) {
            int nxt = -1, earlier = 0;
            for (int jj = 0; jj < j; ++jj) earlier |= (tgt[jj] == tgt[j]);
            for (int jj = Sb - 1; jj > j; --jj)
                if (tgt[jj] == tgt[j]) nxt = jj;
            a.dlink[(size_t)b * a.S + j] = (nxt + 1) | (earlier << 30);
        }
    }
    if (Tb == 0) {
        if (threadIdx.x == 0) a.nll[b] = (Sb == 0) ? 0.0f : -neg_inf();
        return;
    }
    // frames [0,Tm) belong to half 0, [Tm,Tb) to half 1; Tm is a multiple of K (or Tb itself)
    const int Tm = min(Tb, (((Tb + 1) >> 1) + K - 1) / K * K);
    const int nB = Tb - Tm;
    const int ncA = (Tm + K - 1) / K, ncB = (nB + K - 1) / K;
    float* glp_b = a.glp + (size_t)b * T * SP;
    // this half's frames start at frame `base`; its checkpoints at `ck` ([block][NSL])
    const int base = (half == 0) ? 0 : Tm;
    const int nfr = (half == 0) ? Tm : nB;
    const int nblk = (half == 0) ? ncA : ncB;
    float* ck = a.ckpt + (size_t)b * a.ckpt_stride + (size_t)(half == 0 ? 0 : ncA) * NSL;
    auto blk_len = [&](int c) { return min(K, nfr - c * K); };

    auto load_rows = [&](int c) {                 // rows of this half's block c -> ring (no commit)
        const float* src = glp_b + (size_t)(base + c * K) * SP;
        float* dst = rows_of(c);
        const int pieces = (blk_len(c) * SP) >> 2;
        for (int i = lane; i < pieces; i += 32) cp_async16(dst + 4 * i, src + 4 * i);
    };
    auto load_ck = [&](int from_block, int for_block) {   // checkpoint `from_block` -> ring slot of `for_block`
        const float* src = ck + (size_t)from_block * NSL;
        float* dst = ck_of(for_block);
        for (int i = lane; i < NSL / 4; i += 32) cp_async16(dst + 4 * i, src + 4 * i);
    };
    auto alpha_virtual = [&](float (&v)[NS]) {    // alpha(-1)
#pragma unroll
        for (int r = 0; r < NS; ++r) v[r] = (lane == 0 && r == 0) ? 0.0f : kNeg;
    };
    auto beta_virtual = [&](float (&v)[NS]) {     // beta(Tb)
#pragma unroll
        for (int r = 0; r < NS; ++r) v[r] = (lane * NS + r == 2 * Sb) ? 0.0f : kNeg;
    };

    // ---- phase 1 (warps 0 and 1) -------------------------------------------------------------
    float st[NS];   // the recursion a main warp carries: alpha (half 0) or beta (half 1)
    float off_acc = 0.0f;
    if (warp == 0) {
        alpha_virtual(st);
        load_rows(0);
        cp_async_commit();
        for (int c = 0; c < nblk; ++c) {
            if (c + 1 < nblk) {
                load_rows(c + 1);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();
            const float* cur = rows_of(c);
            const int n = blk_len(c);
            for (int i = lane; i < n; i += 32) off_acc += cur[i * SP + ZI + 1];   // the frames' common factors
            for (int i = 0; i < n; ++i) lat.alpha_step(st, cur + i * SP, lane);
#pragma unroll
            for (int r = 0; r < NS; ++r) ck[(size_t)c * NSL + lane * NS + r] = st[r];   // alpha at the last frame of block c
            __syncwarp();
        }
#pragma unroll
        for (int r = 0; r < NS; ++r) xchg[lane * NS + r] = st[r];
    } else if (warp == 1 && nB > 0) {
        beta_virtual(st);
        load_rows(nblk - 1);
        cp_async_commit();
        for (int c = nblk - 1; c >= 0; --c) {
            if (c > 0) {
                load_rows(c - 1);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();
            const float* cur = rows_of(c);
            const int n = blk_len(c);
            for (int i = lane; i < n; i += 32) off_acc += cur[i * SP + ZI + 1];
            for (int i = n - 1; i >= 0; --i) lat.beta_step(st, cur + i * SP, lane);
#pragma unroll
            for (int r = 0; r < NS; ++r) ck[(size_t)c * NSL + lane * NS + r] = st[r];   // beta at the first frame of block c
            __syncwarp();
        }
#pragma unroll
        for (int r = 0; r < NS; ++r) xchg[NSL + lane * NS + r] = st[r];
    }
    if (warp == 1) {
        const float off = warp_sum(off_acc);
        if (lane == 0) mid->off_b = off;
    }
    __syncthreads();

    // ---- midpoint: likelihood from the cut at frame Tm-1 (warp 0) ---------------------------------
    if (warp == 0) {
        float be[NS];
        if (nB > 0) {
#pragma unroll
            for (int r = 0; r < NS; ++r) be[r] = xchg[NSL + lane * NS + r];
        } else {
            beta_virtual(be);
        }
        const float* row = rows_of(ncA - 1) + (size_t)(Tm - 1 - (ncA - 1) * K) * SP;
        lat.beta_step(be, row, lane);
        float v[NS];
        float m = kNeg;
#pragma unroll
        for (int q = 0; q < NH; ++q) {
            v[2 * q] = vb[q] ? (st[2 * q] + be[2 * q]) - row[0] : kNeg;
            v[2 * q + 1] = vl[q] ? (st[2 * q + 1] + be[2 * q + 1]) - row[lat.li[q]] : kNeg;
            m = fmaxf(m, fmaxf(v[2 * q], v[2 * q + 1]));
        }
        m = warp_max(m);
        float sum = 0.0f;
#pragma unroll
        for (int r = 0; r < NS; ++r) sum += ex2f(v[r] - m);
        sum = warp_sum(sum);
        const float ll2 = m + lg2f(sum);            // of the rows relative to their blanks
        const float off = warp_sum(off_acc) + mid->off_b;
        if (lane == 0) {
            const bool ok = ll2 > -1.0e29f && !bad_label;
            mid->feasible = ok;
            mid->nll2 = -ll2;
            a.nll[b] = bad_label ? __int_as_float(0x7fc00000) : ok ? -(ll2 + off) * 0.6931471805599453f : -neg_inf();
        }
    }
    __syncthreads();
    if (a.g == nullptr) return;
    const float nll2 = mid->nll2;
    if (!mid->feasible) {
        float* g_b = a.g + (size_t)b * T * a.ld;
        const float qnan = __int_as_float(0x7fc00000);
        for (int t = 0; t < Tb; ++t)
            for (int i = threadIdx.x; i < a.V; i += 128) g_b[(size_t)t * a.ld + i] = qnan;
        for (size_t i = threadIdx.x; i < (size_t)Tb * SP; i += 128) glp_b[i] = 0.0f;
        return;
    }
    if (nfr == 0) return;

    // ---- phase 2 ---------------------------------------------------------------------------
    // Blocks are visited in the direction of the continuing recursion: half 0 from its last block
    // down to 0, half 1 from 0 up.  `seq` counts visits, blkid(seq) is the block index.
    const int dir = (half == 0) ? -1 : 1;
    auto blkid = [&](int seq) { return (half == 0) ? nblk - 1 - seq : seq; };

    if (helper) {
        // ===== recomputing warp: alpha (half 0) or beta (half 1) of each block from its checkpoint =====
        // block c starts from the checkpoint of the block before it (half 0: alpha at the end of
        // c-1) or after it (half 1: beta at the start of c+1); the virtual frame at the boundary
        auto stage_block = [&](int seq) {
            if (seq < nblk) {
                const int c = blkid(seq);
                load_rows(c);
                const int src = c + dir;
                if (src >= 0 && src < nblk) load_ck(src, c);
            }
            cp_async_commit();
        };
        float rc[NS];
        stage_block(0);
        for (int seq = 0; seq < nblk; ++seq) {
            const int c = blkid(seq);
            // the slots of visit seq+1 (rows) and seq (block) were last used by visit seq-2
            if (seq >= 2) named_bar_sync(bar_free(c), 64);
            stage_block(seq + 1);
            cp_async_wait<1>();
            __syncwarp();
            const int n = blk_len(c);
            const float* rows = rows_of(c);
            float* blk = blk_of(c);
            const int src = c + dir;
            if (src >= 0 && src < nblk) {
                const float* p = ck_of(c) + lane * NS;
#pragma unroll
                for (int r = 0; r < NS; ++r) rc[r] = p[r];
            } else if (half == 0) {
                alpha_virtual(rc);
            } else {
                beta_virtual(rc);
            }
            if (half == 0) {
                for (int i = 0; i < n; ++i) {
                    lat.alpha_step(rc, rows + i * SP, lane);
#pragma unroll
                    for (int r = 0; r < NS; ++r) blk[(size_t)i * NSL + lane * NS + r] = rc[r];
                }
            } else {
                for (int i = n - 1; i >= 0; --i) {
                    lat.beta_step(rc, rows + i * SP, lane);
#pragma unroll
                    for (int r = 0; r < NS; ++r) blk[(size_t)i * NSL + lane * NS + r] = rc[r];
                }
            }
            named_bar_arrive(bar_full(c), 64);
        }
        return;
    }

    // ===== main warp: continues its recursion through the half, adds it to the recomputed one and
    //       turns the sums into occupancies ===================================================
    // st switches role: half 0 continues beta from beta(Tm) (or the boundary), half 1 alpha from alpha(Tm-1)
    if (half == 0) {
        if (nB > 0) {
#pragma unroll
            for (int r = 0; r < NS; ++r) st[r] = xchg[NSL + lane * NS + r];
        } else {
            beta_virtual(st);
        }
    } else {
#pragma unroll
        for (int r = 0; r < NS; ++r) st[r] = xchg[lane * NS + r];
    }
    for (int seq = 0; seq < nblk; ++seq) {
        const int c = blkid(seq);
        const int n = blk_len(c);
        const int t0 = base + c * K;
        const float* rows = rows_of(c);
        float* blk = blk_of(c);
        named_bar_sync(bar_full(c), 64);
        if (half == 0) {
            for (int f = n - 1; f >= 0; --f) {
                lat.beta_step(st, rows + f * SP, lane);
                float* ab = blk + (size_t)f * NSL + lane * NS;
#pragma unroll
                for (int r = 0; r < NS; ++r) ab[r] += st[r];
            }
        } else {
            for (int i = 0; i < n; ++i) {
                lat.alpha_step(st, rows + i * SP, lane);
                float* ab = blk + (size_t)i * NSL + lane * NS;
#pragma unroll
                for (int r = 0; r < NS; ++r) ab[r] += st[r];
            }
        }
        // occupancy(t,s) = exp2(alpha + beta - lp + nll); both alpha and beta include lp_t
#pragma unroll 4
        for (int i = 0; i < n; ++i) {
            const float* row = rows + i * SP;
            const float* ab = blk + (size_t)i * NSL + lane * NS;
            const float lpb = row[0];
            float bsum = 0.0f;
            float* orow = glp_b + (size_t)(t0 + i) * SP;
#pragma unroll
            for (int q = 0; q < NH; ++q) {
                bsum += vb[q] ? ex2f((ab[2 * q] - lpb) + nll2) : 0.0f;
                const float ov = ex2f((ab[2 * q + 1] - row[lat.li[q]]) + nll2);
                if (vl[q]) orow[1 + lane * NH + q] = ov;   // per label position; K3 merges repeats
            }
            blpart[i * 33 + lane] = bsum;
        }
        __syncwarp();
        if (seq + 2 < nblk) named_bar_arrive(bar_free(c), 64);   // rows and block slot may be refilled
        // blank column: one lane per frame sums the 32 partials
        for (int i = lane; i < n; i += 32) {
            float sacc = 0.0f;
#pragma unroll 8
            for (int l = 0; l < 32; ++l) sacc += blpart[i * 33 + l];
            glp_b[(size_t)(t0 + i) * SP] = sacc;
        }
        __syncwarp();
    }
}

static size_t lattice_mitm_smem_bytes(int NS, int K, int SP) {
    const int NH = NS / 2, NSL = 32 * NS;
    const size_t per_half = (size_t)3 * K * SP + (size_t)2 * K * NSL + (size_t)3 * NSL + (size_t)K * 33;
    return (2 * per_half + 2 * (size_t)NSL + 32 * NH) * 4 + sizeof(LatticeMid) + 16;
}

static size_t lattice_smem_bytes(int NS, int K, int T, int SP) {
    const int NH = NS / 2, NSL = 32 * NS;
    const size_t nc = (size_t)(T + K - 1) / K;
    size_t f = 2 * (size_t)K * SP + nc * NSL + (size_t)K * NSL + (size_t)K * 33 + 3 * 32 * NH;
    return f * 4;
}

// ---------------------------------------------------------------------------------
// K3: g[b,t,c] -= occupancy(b,t,c) / (B * len_b)  for the <= S+1 classes of the utterance.
// One warp per valid frame row.  A class that occurs several times in the target is
// handled by its first occurrence, which walks the link chain and sums the others in a
// fixed order, so every address is touched exactly once (deterministic, no atomics).
// Occupancies below 1e-12 are skipped: they cannot change an fp32 gradient whose dense
// part is softmax/(B*len) by more than 1e-12 of the gradient scale.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ctc_apply_kernel(const CtcArgs a) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= (long long)a.B * a.T) return;
    const int b = (int)(row / a.T);
    const int t = (int)(row - (long long)b * a.T);
    const int Tb = min(max(__ldg(a.in_len + b), 0), a.T);
    if (t >= Tb) return;
    const int Sb = min(max(__ldg(a.tgt_len + b), 0), a.S);
    const float scale = 1.0f / ((float)a.Bn * (float)max(Sb, 1));
    const float* __restrict__ orow = a.glp + (size_t)row * a.SP;
    const int* __restrict__ dl = a.dlink + (size_t)b * a.S;
    float* __restrict__ grow = a.g + (size_t)row * a.ld;
    // Up to four entries per lane and pass (S + 1 <= 128 in one pass), in three phases so that a lane's independent
    // loads are in flight together: occupancies / links / classes, then the gradient words, then the stores.  Every
    // address of the row is touched by exactly one (lane, entry), so the order of the read-modify-writes is free.
    constexpr int kE = 4;
    for (int j0 = 0; j0 <= Sb; j0 += 32 * kE) {
        float v[kE], g[kE];
        int c[kE], link[kE];
        bool ok[kE];
#pragma unroll
        for (int e = 0; e < kE; ++e) {
            const int j = j0 + lane + 32 * e;
            ok[e] = j <= Sb;
            v[e] = ok[e] ? orow[j] : 0.0f;
            link[e] = (ok[e] && j > 0) ? __ldg(dl + j - 1) : 0;
            c[e] = (ok[e] && j > 0) ? (int)__ldg(a.targets + (size_t)b * a.S + (j - 1)) : a.blank;
        }
#pragma unroll
        for (int e = 0; e < kE; ++e) {
            const int j = j0 + lane + 32 * e;
            if (ok[e] && j > 0) {
                if (link[e] >> 30) {
                    ok[e] = false;      // a repeat: its first occurrence carries the sum
                } else {
                    for (int nx = (link[e] & 0x3fffffff) - 1; nx >= 0; nx = (__ldg(dl + nx) & 0x3fffffff) - 1) v[e] += orow[1 + nx];
                    c[e] = min(max(c[e], 0), a.V - 1);
                }
            }
            ok[e] = ok[e] && !(fabsf(v[e]) < 1.0e-12f);      // NaN goes through
        }
#pragma unroll
        for (int e = 0; e < kE; ++e) g[e] = ok[e] ? grow[c[e]] : 0.0f;
#pragma unroll
        for (int e = 0; e < kE; ++e)
            if (ok[e]) grow[c[e]] = g[e] - v[e] * scale;
    }
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
// row stride of the gathered table: blank + S labels + the kNeg slot + the frame's common factor, 16-byte rows
static inline int table_stride(int S) { return round_up4(S + 3); }

// lattice states per lane for a target of S labels (2S+1 states over 32 lanes)
static inline int lattice_ns(int S) {
    const int states = 2 * S + 1;
    return states <= 64 ? 2 : states <= 128 ? 4 : states <= 192 ? 6 : states <= 256 ? 8 : states <= 384 ? 12 : 16;
}
// checkpoint floats per utterance of the bidirectional lattice (blocks of >= 16 frames, both halves)
static inline size_t ckpt_stride_floats(int T, int S) { return (size_t)(T / 16 + 3) * 32 * lattice_ns(S); }

extern "C" size_t asr_ctc_workspace_bytes(int B, int T, int V, int S) {
    (void)V;
    if (B <= 0 || T <= 0 || S < 0) return 0;
    return (size_t)B * T * table_stride(S) * sizeof(float) + (size_t)B * (S + 1) * sizeof(int) +
           (size_t)B * ckpt_stride_floats(T, S) * sizeof(float) + 1024;
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
static int launch_lattice(const CtcArgs& a, int stages, cudaStream_t st) {
    int K = 32;
    size_t smem = lattice_smem_bytes(NS, K, a.T, a.SP);
    while (smem > 200 * 1024 && K < 256) {
        K <<= 1;
        smem = lattice_smem_bytes(NS, K, a.T, a.SP);
    }
    if (stages & 2) {
        const int variant = get_opt("ctc_lattice_variant");
        // bidirectional lattice: blocks of 32 frames, or 16 when that lets two CTAs share an SM
        int Km = 32;
        if (lattice_mitm_smem_bytes(NS, 32, a.SP) > 113 * 1024 && lattice_mitm_smem_bytes(NS, 16, a.SP) <= 113 * 1024) Km = 16;
        if (lattice_mitm_smem_bytes(NS, Km, a.SP) > 227 * 1024) Km = 16;
        const size_t smem3 = lattice_mitm_smem_bytes(NS, Km, a.SP);
        if (variant != 1 && !a.fuse_apply && smem3 <= 227 * 1024) {
            ASR_CHECK_CUDA(cudaFuncSetAttribute(ctc_lattice_mitm_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
            ctc_lattice_mitm_kernel<NS><<<a.B, 128, smem3, st>>>(a, Km);
        } else {
            ASR_REQUIRE(smem <= 227 * 1024, "asr_ctc: T=%d S=%d needs %zu bytes of shared memory for the lattice (max 232448)",
                        a.T, a.S, smem);
            ASR_CHECK_CUDA(cudaFuncSetAttribute(ctc_lattice_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ctc_lattice_kernel<NS><<<a.B, 32, smem, st>>>(a, K);
        }
        ASR_LAUNCH_CHECK();
    }
    if ((stages & 4) && a.g != nullptr && !a.fuse_apply) {
        const long long rows = (long long)a.B * a.T;
        ctc_apply_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(a);
        ASR_LAUNCH_CHECK();
    }
    return 0;
}

// One K1 / K2 / K3 sequence over the utterances described by `a`, on one stream.
static int ctc_run(const CtcArgs& a, int stages, cudaStream_t st) {
    if (stages & 1) {
        const long long rows = (long long)a.B * a.T;
        int rc = a.g ? launch_rows_nt<true>(a, rows, st) : launch_rows_nt<false>(a, rows, st);
        if (rc != 0) return rc;
    }
    if ((stages & 6) == 0) return 0;
    const int states = 2 * a.S + 1;
    if (states <= 64) return launch_lattice<2>(a, stages, st);
    if (states <= 128) return launch_lattice<4>(a, stages, st);
    if (states <= 192) return launch_lattice<6>(a, stages, st);
    if (states <= 256) return launch_lattice<8>(a, stages, st);
    if (states <= 384) return launch_lattice<12>(a, stages, st);
    return launch_lattice<16>(a, stages, st);
}

// Streams and events of the sliced pipeline, one set per device, created on first use.
constexpr int kMaxChunks = 8;     // slices per call
constexpr int kTickets = 4;       // calls in flight between begin and finish
constexpr int kLatStreams = 8;
struct CtcPipe {
    cudaStream_t lat[kLatStreams] = {};
    cudaEvent_t k1[kTickets][kMaxChunks] = {}, done[kTickets][kMaxChunks] = {};
    std::atomic<int> busy[kTickets] = {};      // 1 between a begin and its finish
};
static CtcPipe* ctc_pipe() {
    static std::mutex mu;
    static std::map<int, CtcPipe*> pipes;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("asr_ctc: cudaGetDevice failed"); return nullptr; }
    std::lock_guard<std::mutex> lock(mu);
    auto it = pipes.find(dev);
    if (it != pipes.end()) return it->second;
    CtcPipe* p = new CtcPipe();
    int lo = 0, hi = 0;
    bool ok = cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess;
    // the lattices get the highest priority: a few warps that must start the moment their rows are ready
    for (int c = 0; c < kLatStreams && ok; ++c)
        ok = ok && cudaStreamCreateWithPriority(&p->lat[c], cudaStreamNonBlocking, hi) == cudaSuccess;
    for (int t = 0; t < kTickets && ok; ++t)
        for (int c = 0; c < kMaxChunks && ok; ++c) {
            ok = ok && cudaEventCreateWithFlags(&p->k1[t][c], cudaEventDisableTiming) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&p->done[t][c], cudaEventDisableTiming) == cudaSuccess;
        }
    if (!ok) { set_error("asr_ctc: could not create the pipeline streams: %s", cudaGetErrorString(cudaGetLastError())); delete p; return nullptr; }
    pipes[dev] = p;
    return p;
}

static int ctc_make_args(CtcArgs& a, const float* logits, const int64_t* targets, const int* in_len, const int* tgt_len,
                         int B, int T, int V, int S, int blank, float* nll, float* g_logits, void* ws, size_t ws_bytes, int ld = 0) {
    if (ld == 0) ld = V;
    ASR_REQUIRE(ld >= V, "asr_ctc: row stride %d smaller than V=%d", ld, V);
    ASR_REQUIRE((long long)B * T * (long long)ld < (1ll << 40), "asr_ctc: tensor too large");
    ASR_REQUIRE(B > 0 && T > 0 && V > 1 && S >= 0, "asr_ctc: bad shape B=%d T=%d V=%d S=%d", B, T, V, S);
    ASR_REQUIRE(logits && in_len && tgt_len && nll && ws && (S == 0 || targets), "asr_ctc: null pointer");
    ASR_REQUIRE(blank >= 0 && blank < V, "asr_ctc: blank %d out of range", blank);
    ASR_REQUIRE(ws_bytes >= asr_ctc_workspace_bytes(B, T, V, S), "asr_ctc: workspace too small (%zu < %zu)",
                ws_bytes, asr_ctc_workspace_bytes(B, T, V, S));
    ASR_REQUIRE(V <= 65000, "asr_ctc: V=%d > 65000 not supported", V);
    ASR_REQUIRE(2 * S + 1 <= 32 * 16, "asr_ctc: S=%d > 255 labels not supported", S);
    ASR_REQUIRE((long long)B * T < (1ll << 31) - 1, "asr_ctc: B*T too large");
    if (asr_device_ok() != 0) return 3;
    a.logits = logits;
    a.targets = targets;
    a.in_len = in_len;
    a.tgt_len = tgt_len;
    a.B = B; a.T = T; a.V = V; a.S = S; a.blank = blank;
    a.ld = ld;
    a.Bn = B;
    a.SP = table_stride(S);
    a.nll = nll;
    a.g = g_logits;
    uintptr_t w = (reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255;
    a.glp = reinterpret_cast<float*>(w);
    a.dlink = reinterpret_cast<int*>(a.glp + (size_t)B * T * a.SP);
    a.ckpt_stride = ckpt_stride_floats(T, S);
    a.ckpt = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(a.dlink + (size_t)B * (S + 1)) + 255) & ~(uintptr_t)255);
    a.fuse_apply = get_opt("ctc_fuse_apply") == 1 ? 1 : 0;   // measured on B200: the separate K3 pass is 22% faster end to end
    return 0;
}

// utterances [b0, b0+n) of `a`
static CtcArgs ctc_slice(const CtcArgs& a, int b0, int n) {
    CtcArgs ac = a;
    ac.B = n;
    ac.logits = a.logits + (size_t)b0 * a.T * a.ld;
    ac.targets = a.targets ? a.targets + (size_t)b0 * a.S : nullptr;
    ac.in_len = a.in_len + b0;
    ac.tgt_len = a.tgt_len + b0;
    ac.nll = a.nll + b0;
    ac.g = a.g ? a.g + (size_t)b0 * a.T * a.ld : nullptr;
    ac.glp = a.glp + (size_t)b0 * a.T * a.SP;
    ac.dlink = a.dlink + (size_t)b0 * a.S;
    ac.ckpt = a.ckpt + (size_t)b0 * a.ckpt_stride;
    return ac;
}

// The lattice kernel is a latency-bound chain over T (its duration does not depend on B), the
// row and apply kernels are HBM-bound - and two HBM-bound kernels side by side only slow each
// other down (measured).  So the HBM-bound kernels stay in order on the caller's stream and only
// the lattices leave it: the batch is cut into slices, each slice's lattice runs on a
// library-owned high-priority stream right after the slice's row kernel and overlaps whatever
// the caller's stream does next (the next slice's rows, or the caller's own kernels between
// begin and finish).  finish waits for the lattices and applies the sparse update.
static int ctc_slices(int B, int T) {
    int nchunk = get_opt("ctc_chunks");
    if (nchunk <= 0) nchunk = (int)std::min<long long>(4, std::max<long long>(1, (long long)B * T / 65536));   // measured at B*T = 410k: 4 beats 1, 2 and 8
    return std::min(std::min(nchunk, kMaxChunks), B);
}

// Slice c of nchunk: the last slice is a quarter of the others.  Its lattice is the one nothing
// hides (it starts when the last row kernel ends), so at least the apply pass that has to wait
// for it is short.
static void ctc_slice_bounds(int B, int nchunk, int c, int& b0, int& n) {
    if (nchunk <= 1) {
        b0 = 0;
        n = B;
        return;
    }
    int last = B / (4 * (nchunk - 1) + 1);
    if (last < 1) last = (B + nchunk - 1) / nchunk;
    const int rest = B - last;
    const int per = (rest + nchunk - 2) / (nchunk - 1);
    if (c < nchunk - 1) {
        b0 = c * per;
        n = std::max(0, std::min(per, rest - b0));
    } else {
        b0 = rest;
        n = last;
    }
}

extern "C" int asr_ctc_begin_f32(const float* logits, const int64_t* targets, const int* in_len, const int* tgt_len,
                                 int B, int T, int V, int S, int blank, float* nll, float* g_logits, void* ws,
                                 size_t ws_bytes, void* stream, int* ticket) {
    ASR_REQUIRE(ticket != nullptr, "asr_ctc_begin_f32: ticket is null");
    CtcArgs a;
    int rc = ctc_make_args(a, logits, targets, in_len, tgt_len, B, T, V, S, blank, nll, g_logits, ws, ws_bytes);
    if (rc != 0) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CtcPipe* p = ctc_pipe();
    if (p == nullptr) return 3;
    int tk = -1;
    for (int t = 0; t < kTickets && tk < 0; ++t) {
        int expected = 0;
        if (p->busy[t].compare_exchange_strong(expected, 1)) tk = t;
    }
    ASR_REQUIRE(tk >= 0, "asr_ctc_begin_f32: %d calls are already between begin and finish on this device (finish one first)", kTickets);
    struct Guard {      // a begin that fails half way gives its ticket back
        std::atomic<int>& f;
        bool keep = false;
        ~Guard() { if (!keep) f.store(0); }
    } guard{p->busy[tk]};
    const int nchunk = ctc_slices(B, T);
    for (int c = 0; c < nchunk; ++c) {
        int b0, n;
        ctc_slice_bounds(B, nchunk, c, b0, n);
        if (n <= 0) continue;
        const CtcArgs ac = ctc_slice(a, b0, n);
        rc = ctc_run(ac, 1, st);
        if (rc != 0) return rc;
        cudaStream_t ls = p->lat[(tk * kMaxChunks + c) % kLatStreams];
        ASR_CHECK_CUDA(cudaEventRecord(p->k1[tk][c], st));
        ASR_CHECK_CUDA(cudaStreamWaitEvent(ls, p->k1[tk][c], 0));
        rc = ctc_run(ac, 2, ls);
        if (rc != 0) return rc;
        ASR_CHECK_CUDA(cudaEventRecord(p->done[tk][c], ls));
    }
    guard.keep = true;
    *ticket = tk;
    return 0;
}

// per_slice: apply each slice as soon as its lattice is done (right for begin immediately followed by
// finish: the early slices' apply passes run next to the last lattices).  Otherwise wait for every
// lattice and apply the whole batch in one launch: when the caller has queued its own HBM-bound work
// in between, the lattices have had that time, and the apply kernel and a lattice slow each other
// down 2.5-2.8x when they do run side by side (tools/ctc_k2_contention.py).
static int ctc_finish_impl(const float* logits, const int64_t* targets, const int* in_len, const int* tgt_len,
                           int B, int T, int V, int S, int blank, float* nll, float* g_logits, void* ws,
                           size_t ws_bytes, void* stream, int ticket, bool per_slice) {
    ASR_REQUIRE(ticket >= 0 && ticket < kTickets, "asr_ctc_finish_f32: bad ticket %d", ticket);
    CtcArgs a;
    int rc = ctc_make_args(a, logits, targets, in_len, tgt_len, B, T, V, S, blank, nll, g_logits, ws, ws_bytes);
    if (rc != 0) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CtcPipe* p = ctc_pipe();
    if (p == nullptr) return 3;
    ASR_REQUIRE(p->busy[ticket].load() == 1, "asr_ctc_finish_f32: ticket %d is not open (finish called twice, or without begin)", ticket);
    struct Release {      // the ticket is free again once the waits below are queued (stream order protects the events)
        std::atomic<int>& f;
        ~Release() { f.store(0); }
    } release{p->busy[ticket]};
    const int nchunk = ctc_slices(B, T);
    for (int c = 0; c < nchunk; ++c) {
        int b0, n;
        ctc_slice_bounds(B, nchunk, c, b0, n);
        if (n <= 0) continue;
        ASR_CHECK_CUDA(cudaStreamWaitEvent(st, p->done[ticket][c], 0));
        if (per_slice) {
            rc = ctc_run(ctc_slice(a, b0, n), 4, st);
            if (rc != 0) return rc;
        }
    }
    return per_slice ? 0 : ctc_run(a, 4, st);
}

extern "C" int asr_ctc_finish_f32(const float* logits, const int64_t* targets, const int* in_len, const int* tgt_len,
                                  int B, int T, int V, int S, int blank, float* nll, float* g_logits, void* ws,
                                  size_t ws_bytes, void* stream, int ticket) {
    return ctc_finish_impl(logits, targets, in_len, tgt_len, B, T, V, S, blank, nll, g_logits, ws, ws_bytes, stream, ticket,
                           get_opt("ctc_finish_per_slice") == 1);
}

extern "C" int asr_ctc_stages_f32(const float* logits, const int64_t* targets, const int* in_len, const int* tgt_len,
                                  int B, int T, int V, int S, int blank, float* nll, float* g_logits, void* ws,
                                  size_t ws_bytes, int stages, void* stream) {
    CtcArgs a;
    int rc = ctc_make_args(a, logits, targets, in_len, tgt_len, B, T, V, S, blank, nll, g_logits, ws, ws_bytes);
    if (rc != 0) return rc;
    return ctc_run(a, stages, static_cast<cudaStream_t>(stream));
}

extern "C" int asr_ctc_fwd_bwd_f32(const float* logits, const int64_t* targets, const int* in_len, const int* tgt_len,
                                   int B, int T, int V, int S, int blank, float* nll, float* g_logits, void* ws,
                                   size_t ws_bytes, void* stream) {
    if (ctc_slices(B, T) <= 1)
        return asr_ctc_stages_f32(logits, targets, in_len, tgt_len, B, T, V, S, blank, nll, g_logits, ws, ws_bytes, 7, stream);
    int ticket = 0;
    int rc = asr_ctc_begin_f32(logits, targets, in_len, tgt_len, B, T, V, S, blank, nll, g_logits, ws, ws_bytes, stream, &ticket);
    if (rc != 0) return rc;
    return ctc_finish_impl(logits, targets, in_len, tgt_len, B, T, V, S, blank, nll, g_logits, ws, ws_bytes, stream, ticket, true);
}

extern "C" int asr_ctc_fwd_bwd_ld_f32(const float* logits, const int64_t* targets, const int* in_len, const int* tgt_len,
                                      int B, int T, int V, int ld, int S, int blank, float* nll, float* g_logits, void* ws,
                                      size_t ws_bytes, void* stream) {
    CtcArgs a;
    int rc = ctc_make_args(a, logits, targets, in_len, tgt_len, B, T, V, S, blank, nll, g_logits, ws, ws_bytes, ld);
    if (rc != 0) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nchunk = ctc_slices(B, T);
    if (nchunk <= 1) return ctc_run(a, 7, st);
    // same sliced schedule as asr_ctc_fwd_bwd_f32: rows on the caller's stream, lattices on the library's streams
    CtcPipe* p = ctc_pipe();
    if (p == nullptr) return 3;
    int tk = -1;
    for (int t = 0; t < kTickets && tk < 0; ++t) {
        int expected = 0;
        if (p->busy[t].compare_exchange_strong(expected, 1)) tk = t;
    }
    ASR_REQUIRE(tk >= 0, "asr_ctc_fwd_bwd_ld_f32: %d calls are already between begin and finish on this device", kTickets);
    struct Release {
        std::atomic<int>& f;
        ~Release() { f.store(0); }
    } release{p->busy[tk]};
    for (int c = 0; c < nchunk; ++c) {
        int b0, n;
        ctc_slice_bounds(B, nchunk, c, b0, n);
        if (n <= 0) continue;
        const CtcArgs ac = ctc_slice(a, b0, n);
        rc = ctc_run(ac, 1, st);
        if (rc != 0) return rc;
        cudaStream_t ls = p->lat[(tk * kMaxChunks + c) % kLatStreams];
        ASR_CHECK_CUDA(cudaEventRecord(p->k1[tk][c], st));
        ASR_CHECK_CUDA(cudaStreamWaitEvent(ls, p->k1[tk][c], 0));
        rc = ctc_run(ac, 2, ls);
        if (rc != 0) return rc;
        ASR_CHECK_CUDA(cudaEventRecord(p->done[tk][c], ls));
    }
    for (int c = 0; c < nchunk; ++c) {
        int b0, n;
        ctc_slice_bounds(B, nchunk, c, b0, n);
        if (n <= 0) continue;
        ASR_CHECK_CUDA(cudaStreamWaitEvent(st, p->done[tk][c], 0));
        rc = ctc_run(ctc_slice(a, b0, n), 4, st);
        if (rc != 0) return rc;
    }
    return 0;
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

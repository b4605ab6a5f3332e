// cif.cu - continuous integrate-and-fire, forward + analytic backward (sm_100a).
//
// Replaces CIF_Model.cif (/root/reference/src/transformer/cif_model.py:57-106).
//
// Forward: ONE WARP per (utterance, hidden slice).  The scalar recurrence
// (integrate / fire / cur / rem) is evaluated redundantly by every lane in the
// reference's exact fp32 operation order (no FMA contraction, no reassociation),
// so fire positions are bit-identical; the [T, slice] stream of encoder frames is
// the only real traffic.  Two variants, same arithmetic:
//   * tma   : 2-D TMA tiles [32 rows x W floats] into a multi-stage shared-memory
//             ring with mbarrier completion (needs H % 4 == 0);
//   * plain : vectorised global loads with an 8-row register prefetch (any H).
// Backward: the carried frame gradient is constant inside a fire segment, so
// every frame row is independent given the saved schedule: one warp per row
// streams hidden/g_out and writes g_hidden plus two dot products; a tiny second
// kernel turns the dot products into g_alpha with a reverse (suffix) scan.
#include "common.cuh"

namespace asr {

// ---- vector helpers -----------------------------------------------------------
template <int VEC>
struct VecT;
template <>
struct VecT<1> {
    using type = float;
};
template <>
struct VecT<2> {
    using type = float2;
};
template <>
struct VecT<4> {
    using type = float4;
};

template <int VEC>
__device__ __forceinline__ void vload(float (&d)[VEC], const float* p) {
    if constexpr (VEC == 4) {
        float4 v = *reinterpret_cast<const float4*>(p);
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    } else if constexpr (VEC == 2) {
        float2 v = *reinterpret_cast<const float2*>(p);
        d[0] = v.x; d[1] = v.y;
    } else {
        d[0] = *p;
    }
}
template <int VEC>
__device__ __forceinline__ void vload_nc(float (&d)[VEC], const float* p) {
    if constexpr (VEC == 4) {
        float4 v = __ldg(reinterpret_cast<const float4*>(p));
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    } else if constexpr (VEC == 2) {
        float2 v = __ldg(reinterpret_cast<const float2*>(p));
        d[0] = v.x; d[1] = v.y;
    } else {
        d[0] = __ldg(p);
    }
}
template <int VEC>
__device__ __forceinline__ void vstore(float* p, const float (&s)[VEC]) {
    if constexpr (VEC == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(s[0], s[1], s[2], s[3]);
    } else if constexpr (VEC == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(s[0], s[1]);
    } else {
        *p = s[0];
    }
}

// One step of the scalar recurrence, cif_model.py:69-81, in the reference's
// operation order.  __f*_rn intrinsics are never contracted into FMAs.
__device__ __forceinline__ void cif_chain_step(float a, float thr, float& integ, bool& fire, float& cur, float& rem) {
    const float dc = __fsub_rn(1.0f, integ);   // distribution_completion   (:69)
    const float s = __fadd_rn(integ, a);        // integrate += alpha        (:71)
    fire = s > thr;                             // fire_place                (:74)
    integ = fire ? __fsub_rn(s, 1.0f) : s;      //                           (:75-77)
    cur = fire ? dc : a;                        //                           (:78-80)
    rem = __fsub_rn(a, cur);                    // remainds                  (:81)
}

struct CifFwdArgs {
    const float* hidden;
    const float* alphas;
    float thr;
    int B, T, H, L;
    float* out;
    int* fire_t;
    int* n_fired;
    float* cur;
    float* rem;
    int* sched;
    float* alpha_sum;
    const float* target_num;
    float* qua_term;
};

// Phase 1 of a 32-frame chunk: the scalar recurrence only (no memory traffic).
// The sequential loop carries nothing but `integrate` (4 dependent instructions per
// frame); lane r snapshots the value of integrate BEFORE frame r, and cur / rem /
// fire are then derived lane-parallel with the reference's own operations, so they
// are bit-identical to a frame-by-frame evaluation.
template <bool FULL>
__device__ __forceinline__ void cif_chain_chunk(float my_alpha, int nrow, float thr, int lane, float& integ,
                                                float& asum, float& my_cur, float& my_rem, unsigned& fire_mask) {
    float my_prev = 0.0f;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        if (FULL || r < nrow) {   // warp-uniform
            const float al = __shfl_sync(0xffffffffu, my_alpha, r);
            if (lane == r) my_prev = integ;
            const float s = __fadd_rn(integ, al);              // integrate += alpha        (:71)
            integ = (s > thr) ? __fsub_rn(s, 1.0f) : s;        // fire -> integrate - 1     (:74-77)
        }
    }
    const float dc = __fsub_rn(1.0f, my_prev);                 // distribution_completion   (:69)
    const float s = __fadd_rn(my_prev, my_alpha);
    const bool fire = (lane < nrow) && (s > thr);
    my_cur = fire ? dc : my_alpha;                             //                           (:78-80)
    my_rem = __fsub_rn(my_alpha, my_cur);                      // remainds                  (:81)
    fire_mask = __ballot_sync(0xffffffffu, fire);
    asum += warp_sum(my_alpha);                                // lanes >= nrow hold 0
}

// Saved-for-backward schedule and fire positions of a chunk (slice-0 warps only).
__device__ __forceinline__ void cif_record_chunk(const CifFwdArgs& a, int b, int tt, int lane, int k, float my_cur,
                                                 float my_rem, unsigned fire_mask) {
    if (tt < a.T) {
        const int before = k + __popc(fire_mask & ((1u << lane) - 1u));
        const bool fired = (fire_mask >> lane) & 1u;
        a.cur[(size_t)b * a.T + tt] = my_cur;
        a.rem[(size_t)b * a.T + tt] = my_rem;
        a.sched[(size_t)b * a.T + tt] = (before << 1) | (fired ? 1 : 0);
        if (fired && before < a.L) a.fire_t[(size_t)b * a.L + before] = tt;
    }
}

// Shared epilogue: zero rows k..L-1 of this warp's slice and publish counters.
template <int VEC>
__device__ __forceinline__ void cif_fwd_finish(const CifFwdArgs& a, int b, int slice, int lane, int col, bool col_ok,
                                               int k, float asum) {
    float z[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) z[i] = 0.0f;
    if (col_ok) {
        for (int kk = k; kk < a.L; ++kk) vstore<VEC>(a.out + ((size_t)b * a.L + kk) * a.H + col, z);
    }
    if (slice == 0) {
        for (int kk = k + lane; kk < a.L; kk += 32) a.fire_t[(size_t)b * a.L + kk] = -1;
        if (lane == 0) {
            a.n_fired[b] = k;
            a.alpha_sum[b] = asum;
            if (a.target_num != nullptr && a.qua_term != nullptr) {
                const float d = __fsub_rn(asum, a.target_num[b]);
                a.qua_term[b] = __fmul_rn(d, d);
            }
        }
    }
}

// Predicated vector store (no branch): if (pred) *p = v.
template <int VEC>
__device__ __forceinline__ void vstore_if(float* p, const float (&v)[VEC], bool pred) {
    const unsigned pr = pred ? 1u : 0u;
    if constexpr (VEC == 4) {
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q st.global.v4.f32 [%1], {%2, %3, %4, %5};\n\t}"
                     ::"r"(pr), "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    } else if constexpr (VEC == 2) {
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q st.global.v2.f32 [%1], {%2, %3};\n\t}"
                     ::"r"(pr), "l"(p), "f"(v[0]), "f"(v[1]) : "memory");
    } else {
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q st.global.f32 [%1], %2;\n\t}"
                     ::"r"(pr), "l"(p), "f"(v[0]) : "memory");
    }
}

// One frame of phase 2: frame += cur*h (:83); on a fire emit the frame and restart
// it from rem*h (:85-87).  `fired` is warp-uniform.  Written with selects, a predicated
// store and a running output pointer, so the unrolled loop contains no branch at all.
// `optr` points at this lane's slot of output row k; `room` = rows still free (L - k).
template <int VEC>
__device__ __forceinline__ void cif_frame_step(const float (&h)[VEC], float c, float rm, bool fired, bool col_ok,
                                               size_t row_stride, float (&frame)[VEC], float*& optr, int& room) {
    float pre[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) pre[i] = __fadd_rn(frame[i], __fmul_rn(c, h[i]));
    vstore_if<VEC>(optr, pre, fired && col_ok && room > 0);
#pragma unroll
    for (int i = 0; i < VEC; ++i) frame[i] = fired ? __fmul_rn(rm, h[i]) : pre[i];
    optr += fired ? row_stride : 0;
    room -= fired ? 1 : 0;
}

// Phase 2 over a full 32-row tile in shared memory: loads are issued 8 rows at a
// time ahead of the arithmetic (tile rows + the broadcast (cur, rem) pairs).
template <int VEC>
__device__ __forceinline__ void cif_tile_full(bool col_ok, size_t out_stride, const float* tile, int row_stride,
                                              const float2* cr, unsigned mask, float (&frame)[VEC], float*& optr,
                                              int& room) {
#pragma unroll 1
    for (int r0 = 0; r0 < 32; r0 += 8) {
        float h[8][VEC];
        float2 w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            vload<VEC>(h[u], tile + (r0 + u) * row_stride);
            w[u] = cr[r0 + u];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            cif_frame_step<VEC>(h[u], w[u].x, w[u].y, (mask >> (r0 + u)) & 1u, col_ok, out_stride, frame, optr, room);
    }
}
template <int VEC>
__device__ __forceinline__ void cif_tile_partial(bool col_ok, size_t out_stride, const float* tile, int row_stride,
                                                 const float2* cr, unsigned mask, int nrow, float (&frame)[VEC],
                                                 float*& optr, int& room) {
    for (int r = 0; r < nrow; ++r) {
        float h[VEC];
        vload<VEC>(h, tile + r * row_stride);
        const float2 w = cr[r];
        cif_frame_step<VEC>(h, w.x, w.y, (mask >> r) & 1u, col_ok, out_stride, frame, optr, room);
    }
}

// ---- forward, plain loads -----------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(128) cif_fwd_plain_kernel(const CifFwdArgs a, int nslices) {
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (warp >= a.B * nslices) return;
    const int b = warp / nslices;
    const int slice = warp - b * nslices;
    const int col = slice * (32 * VEC) + lane * VEC;
    const bool col_ok = col < a.H;
    const bool rec = (slice == 0);
    const float* hrow = a.hidden + (size_t)b * a.T * a.H + (col_ok ? col : 0);
    const float* arow = a.alphas + (size_t)b * a.T;

    float frame[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) frame[i] = 0.0f;
    float integ = 0.0f, asum = 0.0f;
    int k = 0;                 // fires so far (schedule side)
    int room = a.L;            // output rows still free (data side)
    float* optr = a.out + (size_t)b * a.L * a.H + (col_ok ? col : 0);
    constexpr int U = 8;

    for (int t0 = 0; t0 < a.T; t0 += 32) {
        const int tt = t0 + lane;
        const float my_alpha = (tt < a.T) ? __ldg(arow + tt) : 0.0f;
        const int nrow = min(32, a.T - t0);
        float my_cur, my_rem;
        unsigned fire_mask;
        if (nrow == 32)
            cif_chain_chunk<true>(my_alpha, nrow, a.thr, lane, integ, asum, my_cur, my_rem, fire_mask);
        else
            cif_chain_chunk<false>(my_alpha, nrow, a.thr, lane, integ, asum, my_cur, my_rem, fire_mask);
        if (rec) cif_record_chunk(a, b, tt, lane, k, my_cur, my_rem, fire_mask);
        k += __popc(fire_mask);

        for (int r0 = 0; r0 < nrow; r0 += U) {
            float h[U][VEC];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (r0 + u < nrow && col_ok) {
                    vload_nc<VEC>(h[u], hrow + (size_t)(t0 + r0 + u) * a.H);
                } else {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) h[u][i] = 0.0f;
                }
            }
            float cc[U], rr[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                cc[u] = __shfl_sync(0xffffffffu, my_cur, (r0 + u) & 31);
                rr[u] = __shfl_sync(0xffffffffu, my_rem, (r0 + u) & 31);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int r = r0 + u;
                if (r < nrow)   // warp-uniform
                    cif_frame_step<VEC>(h[u], cc[u], rr[u], (fire_mask >> r) & 1u, col_ok, (size_t)a.H, frame, optr, room);
            }
        }
    }
    cif_fwd_finish<VEC>(a, b, slice, lane, col, col_ok, k, asum);
}

// ---- forward, schedule + segment-parallel rows ---------------------------------
// The fire schedule only depends on alphas: one warp per utterance runs the exact recurrence
// once (cur / rem / sched / fire_t / n_fired: what backward needs anyway).  With the fire
// positions known, output row l is an ordered weighted sum over the frames of one segment,
//     out[l] = rem[t0] h[t0] + sum_{t0 < t <= t1} cur[t] h[t],   t0 = fire_t[l-1], t1 = fire_t[l]
// (row 0: 0 + sum_{t <= t1}), so rows are independent: one warp per (row, column slice), every
// warp streams ~T/L frames with 8 loads in flight.  Same operations in the same order as the
// frame-by-frame loop, hence the same bits; the boundary frame of each segment is read twice
// (+L/T of the traffic), and the grid no longer depends on B*H/width filling the machine.
__global__ void __launch_bounds__(128) cif_schedule_kernel(const CifFwdArgs a) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= a.B) return;
    const float* arow = a.alphas + (size_t)b * a.T;
    float integ = 0.0f, asum = 0.0f;
    int k = 0;
    for (int t0 = 0; t0 < a.T; t0 += 32) {
        const int tt = t0 + lane;
        const float my_alpha = (tt < a.T) ? __ldg(arow + tt) : 0.0f;
        const int nrow = min(32, a.T - t0);
        float my_cur, my_rem;
        unsigned fire_mask;
        if (nrow == 32)
            cif_chain_chunk<true>(my_alpha, nrow, a.thr, lane, integ, asum, my_cur, my_rem, fire_mask);
        else
            cif_chain_chunk<false>(my_alpha, nrow, a.thr, lane, integ, asum, my_cur, my_rem, fire_mask);
        cif_record_chunk(a, b, tt, lane, k, my_cur, my_rem, fire_mask);
        k += __popc(fire_mask);
    }
    for (int kk = k + lane; kk < a.L; kk += 32) a.fire_t[(size_t)b * a.L + kk] = -1;
    if (lane == 0) {
        a.n_fired[b] = k;
        a.alpha_sum[b] = asum;
        if (a.target_num != nullptr && a.qua_term != nullptr) {
            const float d = __fsub_rn(asum, a.target_num[b]);
            a.qua_term[b] = __fmul_rn(d, d);
        }
    }
}

template <int VEC>
__global__ void __launch_bounds__(128) cif_rows_kernel(const CifFwdArgs a, int nslices) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= (long long)a.B * a.L * nslices) return;
    const int slice = (int)(w % nslices);
    const long long bl = w / nslices;
    const int l = (int)(bl % a.L);
    const int b = (int)(bl / a.L);
    const int col = slice * (32 * VEC) + lane * VEC;
    if (col >= a.H) return;
    float* orow = a.out + ((size_t)b * a.L + l) * a.H + col;
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.0f;
    if (l >= __ldg(a.n_fired + b)) {     // beyond the last fire: zero row
        vstore<VEC>(orow, acc);
        return;
    }
    const int t1 = __ldg(a.fire_t + (size_t)b * a.L + l);
    const int t0 = (l > 0) ? __ldg(a.fire_t + (size_t)b * a.L + l - 1) : -1;
    const float* hrow = a.hidden + (size_t)b * a.T * a.H + col;
    const float* crow = a.cur + (size_t)b * a.T;
    if (l > 0) {                         // the fire frame of the previous row opens this one: frame = rem * h   (:87)
        float h[VEC];
        vload_nc<VEC>(h, hrow + (size_t)t0 * a.H);
        const float rm = a.rem[(size_t)b * a.T + t0];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = __fmul_rn(rm, h[i]);
    }
    constexpr int U = 8;
    for (int t = t0 + 1; t <= t1; t += U) {
        float h[U][VEC];
        float c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int tu = min(t + u, t1);   // clamped loads (valid addresses), unused beyond t1
            vload_nc<VEC>(h[u], hrow + (size_t)tu * a.H);
            c[u] = crow[tu];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (t + u <= t1) {               // warp-uniform
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] = __fadd_rn(acc[i], __fmul_rn(c[u], h[u][i]));   // frame += cur * h   (:83)
            }
        }
    }
    vstore<VEC>(orow, acc);
}

// ---- forward, TMA pipeline ----------------------------------------------------
// One warp per CTA.  Tile = 32 rows x (32*VEC) floats; NSTAGE tiles in flight.
constexpr int kCifRows = 32;
constexpr int kCifMaxStages = 12;

template <int VEC>
__global__ void __launch_bounds__(32) cif_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmap, const CifFwdArgs a,
                                                         int nstage) {
    constexpr int W = 32 * VEC;
    constexpr uint32_t kTileBytes = kCifRows * W * sizeof(float);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[kCifMaxStages];
    __shared__ __align__(16) float2 s_cr[kCifRows];

    const int lane = threadIdx.x;
    const int slice = blockIdx.x;
    const int b = blockIdx.y;
    const int col = slice * W + lane * VEC;
    const bool col_ok = col < a.H;
    const bool rec = (slice == 0);
    const float* arow = a.alphas + (size_t)b * a.T;
    const int nchunk = (a.T + kCifRows - 1) / kCifRows;
    const int row0 = b * a.T;

    if (lane == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < nstage; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (lane == 0) {
        for (int s = 0; s < nstage && s < nchunk; ++s) {
            mbar_arrive_expect_tx(&bars[s], kTileBytes);
            tma_load_2d(smem_raw + (size_t)s * kTileBytes, &tmap, slice * W, row0 + s * kCifRows, &bars[s]);
        }
    }

    float frame[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) frame[i] = 0.0f;
    float integ = 0.0f, asum = 0.0f;
    int k = 0;                 // fires so far (schedule side)
    int room = a.L;            // output rows still free (data side)
    float* optr = a.out + (size_t)b * a.L * a.H + (col_ok ? col : 0);
    int stage = 0;
    uint32_t phase = 0;
    float next_alpha = (lane < a.T) ? __ldg(arow + lane) : 0.0f;

    for (int c = 0; c < nchunk; ++c) {
        const int t0 = c * kCifRows;
        const int tt = t0 + lane;
        const float my_alpha = next_alpha;
        {
            const int tn = tt + kCifRows;
            next_alpha = (tn < a.T) ? __ldg(arow + tn) : 0.0f;
        }
        const int nrow = min(kCifRows, a.T - t0);

        // phase 1: recurrence (overlaps the TMA transfer of this stage)
        float my_cur, my_rem;
        unsigned fire_mask;
        if (nrow == kCifRows)
            cif_chain_chunk<true>(my_alpha, nrow, a.thr, lane, integ, asum, my_cur, my_rem, fire_mask);
        else
            cif_chain_chunk<false>(my_alpha, nrow, a.thr, lane, integ, asum, my_cur, my_rem, fire_mask);
        if (rec) cif_record_chunk(a, b, tt, lane, k, my_cur, my_rem, fire_mask);

        // phase 2: weighted accumulation of the staged tile
        __syncwarp();
        s_cr[lane] = make_float2(my_cur, my_rem);
        __syncwarp();
        mbar_wait(&bars[stage], phase);
        const float* tile = reinterpret_cast<const float*>(smem_raw + (size_t)stage * kTileBytes) + lane * VEC;
        k += __popc(fire_mask);
        if (nrow == kCifRows)
            cif_tile_full<VEC>(col_ok, (size_t)a.H, tile, W, s_cr, fire_mask, frame, optr, room);
        else
            cif_tile_partial<VEC>(col_ok, (size_t)a.H, tile, W, s_cr, fire_mask, nrow, frame, optr, room);
        // every lane is done reading this stage -> refill it
        __syncwarp();
        const int cn = c + nstage;
        if (lane == 0 && cn < nchunk) {
            mbar_arrive_expect_tx(&bars[stage], kTileBytes);
            tma_load_2d(smem_raw + (size_t)stage * kTileBytes, &tmap, slice * W, row0 + cn * kCifRows, &bars[stage]);
        }
        if (++stage == nstage) {
            stage = 0;
            phase ^= 1u;
        }
    }
    cif_fwd_finish<VEC>(a, b, slice, lane, col, col_ok, k, asum);
}

// ---- forward, warp-specialised TMA pipeline ---------------------------------------
// CTA = NW data warps + 1 schedule warp, working on (utterance b, a group of
// NW*32*VEC hidden columns).  The schedule warp runs the scalar recurrence once per
// CTA (not once per slice), publishes each chunk's cur/rem/fire-mask in shared
// memory and issues the TMA tile loads; data warps only do the weighted
// accumulation, branch-free, from shared memory.  full[] = tile landed AND schedule
// published; empty[] = all data warps are done with the stage.
struct CifSched {
    float2 cr[kCifRows];   // (cur, rem) per frame of the chunk
    unsigned mask;         // fire bits
    int pad[3];
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

template <int VEC>
__global__ void __launch_bounds__(160) cif_fwd_ws_kernel(const __grid_constant__ CUtensorMap tmap, const CifFwdArgs a,
                                                         int nstage, int nw) {
    constexpr int W = 32 * VEC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[kCifMaxStages];
    __shared__ __align__(8) uint64_t empty[kCifMaxStages];
    __shared__ __align__(16) CifSched sched[kCifMaxStages];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int CW = nw * W;   // columns per CTA
    const uint32_t tile_bytes = (uint32_t)(kCifRows * CW * sizeof(float));
    const int cgroup = blockIdx.x;
    const int b = blockIdx.y;
    const int nchunk = (a.T + kCifRows - 1) / kCifRows;
    const int row0 = b * a.T;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < nstage; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], nw);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == nw) {
        // ===== schedule + producer warp =====
        const float* arow = a.alphas + (size_t)b * a.T;
        const bool rec = (cgroup == 0);
        float integ = 0.0f, asum = 0.0f;
        int k = 0, stage = 0;
        uint32_t phase = 0;
        float next_alpha = (lane < a.T) ? __ldg(arow + lane) : 0.0f;
        for (int c = 0; c < nchunk; ++c) {
            const int t0 = c * kCifRows;
            const int tt = t0 + lane;
            const float my_alpha = next_alpha;
            {
                const int tn = tt + kCifRows;
                next_alpha = (tn < a.T) ? __ldg(arow + tn) : 0.0f;
            }
            const int nrow = min(kCifRows, a.T - t0);
            if (c >= nstage) mbar_wait(&empty[stage], phase ^ 1u);
            if (lane == 0) {
                mbar_expect_tx(&full[stage], tile_bytes);
                tma_load_2d(smem_raw + (size_t)stage * tile_bytes, &tmap, cgroup * CW, row0 + t0, &full[stage]);
            }
            float my_cur, my_rem;
            unsigned mask;
            if (nrow == kCifRows)
                cif_chain_chunk<true>(my_alpha, nrow, a.thr, lane, integ, asum, my_cur, my_rem, mask);
            else
                cif_chain_chunk<false>(my_alpha, nrow, a.thr, lane, integ, asum, my_cur, my_rem, mask);
            sched[stage].cr[lane] = make_float2(my_cur, my_rem);
            if (lane == 0) sched[stage].mask = mask;
            __syncwarp();
            if (rec && tt < a.T) {
                const float2 cr = make_float2(my_cur, my_rem);
                const int before = k + __popc(mask & ((1u << lane) - 1u));
                const bool fired = (mask >> lane) & 1u;
                a.cur[(size_t)b * a.T + tt] = cr.x;
                a.rem[(size_t)b * a.T + tt] = cr.y;
                a.sched[(size_t)b * a.T + tt] = (before << 1) | (fired ? 1 : 0);
                if (fired && before < a.L) a.fire_t[(size_t)b * a.L + before] = tt;
            }
            k += __popc(mask);
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);   // release: schedule visible, tile bytes pending
            if (++stage == nstage) {
                stage = 0;
                phase ^= 1u;
            }
        }
        if (rec) {
            for (int kk = k + lane; kk < a.L; kk += 32) a.fire_t[(size_t)b * a.L + kk] = -1;
            if (lane == 0) {
                a.n_fired[b] = k;
                a.alpha_sum[b] = asum;
                if (a.target_num != nullptr && a.qua_term != nullptr) {
                    const float d = __fsub_rn(asum, a.target_num[b]);
                    a.qua_term[b] = __fmul_rn(d, d);
                }
            }
        }
    } else {
        // ===== data warps =====
        const int col = cgroup * CW + warp * W + lane * VEC;
        const bool col_ok = col < a.H;
        float frame[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) frame[i] = 0.0f;
        int k = 0, stage = 0;
        int room = a.L;
        float* optr = a.out + (size_t)b * a.L * a.H + (col_ok ? col : 0);
        uint32_t phase = 0;
        for (int c = 0; c < nchunk; ++c) {
            const int nrow = min(kCifRows, a.T - c * kCifRows);
            mbar_wait(&full[stage], phase);
            const float* tile = reinterpret_cast<const float*>(smem_raw + (size_t)stage * tile_bytes) + warp * W + lane * VEC;
            const float2* cr = sched[stage].cr;
            const unsigned mask = sched[stage].mask;
            k += __popc(mask);
            if (nrow == kCifRows)
                cif_tile_full<VEC>(col_ok, (size_t)a.H, tile, CW, cr, mask, frame, optr, room);
            else
                cif_tile_partial<VEC>(col_ok, (size_t)a.H, tile, CW, cr, mask, nrow, frame, optr, room);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == nstage) {
                stage = 0;
                phase ^= 1u;
            }
        }
        float z[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) z[i] = 0.0f;
        if (col_ok)
            for (int kk = k; kk < a.L; ++kk) vstore<VEC>(a.out + ((size_t)b * a.L + kk) * a.H + col, z);
    }
}

// ---- backward -----------------------------------------------------------------
struct CifBwdArgs {
    const float* hidden;
    const float* g_out;
    const int* n_fired;
    const float* cur;
    const float* rem;
    const int* sched;
    int B, T, H, L;
    float* g_hidden;
    float* part;   // == g_alphas, finalised by the scan kernel
    float* gcf;    // workspace [B*T]: fire ? gcur : 0
};

constexpr int kBwdRowsPerCta = 32;

template <int VEC>
__global__ void __launch_bounds__(256) cif_bwd_rows_kernel(const CifBwdArgs a) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp = blockDim.x >> 5;
    const long long total = (long long)a.B * a.T;
    const long long base = (long long)blockIdx.x * kBwdRowsPerCta;
    for (int i = warp; i < kBwdRowsPerCta; i += nwarp) {
        const long long row = base + i;
        if (row >= total) break;
        const int b = (int)(row / a.T);
        const int sc = __ldg(a.sched + row);
        const bool fire = (sc & 1) != 0;
        const int seg = sc >> 1;
        const int nf = min(__ldg(a.n_fired + b), a.L);
        const float c = __ldg(a.cur + row);
        const float r = __ldg(a.rem + row);
        const float* gp = (seg < nf) ? a.g_out + ((size_t)b * a.L + seg) * a.H : nullptr;
        const float* gG = (fire && seg + 1 < nf) ? a.g_out + ((size_t)b * a.L + seg + 1) * a.H : nullptr;
        const float* h = a.hidden + (size_t)row * a.H;
        float* gh = a.g_hidden + (size_t)row * a.H;
        float d1 = 0.0f, d2 = 0.0f;
        for (int col = lane * VEC; col < a.H; col += 32 * VEC) {
            float hv[VEC], pv[VEC], Gv[VEC], o[VEC];
            vload_nc<VEC>(hv, h + col);
            if (gp != nullptr) {
                vload<VEC>(pv, gp + col);
            } else {
#pragma unroll
                for (int j = 0; j < VEC; ++j) pv[j] = 0.0f;
            }
            if (gG != nullptr) {
                vload<VEC>(Gv, gG + col);
            } else {
#pragma unroll
                for (int j = 0; j < VEC; ++j) Gv[j] = 0.0f;
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                o[j] = __fmul_rn(c, pv[j]);
                d1 = fmaf(pv[j], hv[j], d1);
                if (fire) {
                    o[j] = __fadd_rn(o[j], __fmul_rn(r, Gv[j]));
                    d2 = fmaf(Gv[j], hv[j], d2);
                }
            }
            vstore<VEC>(gh + col, o);
        }
        d1 = warp_sum(d1);
        d2 = warp_sum(d2);
        if (lane == 0) {
            a.part[row] = fire ? d2 : d1;
            a.gcf[row] = fire ? (d1 - d2) : 0.0f;
        }
    }
}

// g_alpha[t] = part[t] - sum_{s>t} gcf[s]    (one CTA per utterance; tiles of
// 256 x 4 frames processed from the end, block-wide exclusive suffix scan per tile)
constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
__global__ void __launch_bounds__(kScanThreads) cif_bwd_scan_kernel(float* g_alpha, const float* gcf, int B, int T) {
    __shared__ float wsum[kScanThreads / 32];
    __shared__ float tile_total;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* ga = g_alpha + (size_t)b * T;
    const float* gc = gcf + (size_t)b * T;
    constexpr int TILE = kScanThreads * kScanItems;
    float carry = 0.0f;
    for (int base = ((T - 1) / TILE) * TILE; base >= 0; base -= TILE) {
        float v[kScanItems], p[kScanItems];
        const int t0 = base + tid * kScanItems;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            const int t = t0 + i;
            v[i] = (t < T) ? gc[t] : 0.0f;
            p[i] = (t < T) ? ga[t] : 0.0f;
        }
        // thread-local suffix sums: s[i] = sum_{j>i} v[j]
        float s[kScanItems];
        float run = 0.0f;
#pragma unroll
        for (int i = kScanItems - 1; i >= 0; --i) {
            s[i] = run;
            run += v[i];
        }
        // warp-level exclusive suffix over thread totals
        float incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float y = __shfl_down_sync(0xffffffffu, incl, o);
            if (lane + o < 32) incl += y;
        }
        if (lane == 0) wsum[warp] = incl;
        __syncthreads();
        float after = incl - run;   // later threads of this warp
        for (int w = warp + 1; w < kScanThreads / 32; ++w) after += wsum[w];
        if (tid == 0) {
            float tot = 0.0f;
            for (int w = 0; w < kScanThreads / 32; ++w) tot += wsum[w];
            tile_total = tot;
        }
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            const int t = t0 + i;
            if (t < T) ga[t] = p[i] - (carry + after + s[i]);
        }
        __syncthreads();
        carry += tile_total;
        __syncthreads();
    }
}

}  // namespace asr

using namespace asr;

extern "C" int asr_cif_fwd_f32(const float* hidden, const float* alphas, float threshold, int B, int T, int H, int L,
                               float* out, int* fire_t, int* n_fired, float* cur, float* rem, int* sched,
                               float* alpha_sum, const float* target_num, float* qua_term, void* stream) {
    return asr_cif_fwd_hint_f32(hidden, alphas, threshold, B, T, H, L, out, fire_t, n_fired, cur, rem, sched, alpha_sum,
                                target_num, qua_term, 0, stream);
}

extern "C" int asr_cif_fwd_hint_f32(const float* hidden, const float* alphas, float threshold, int B, int T, int H, int L,
                                    float* out, int* fire_t, int* n_fired, float* cur, float* rem, int* sched,
                                    float* alpha_sum, const float* target_num, float* qua_term, int kernel_hint, void* stream) {
    ASR_REQUIRE(B > 0 && T > 0 && H > 0 && L >= 0, "asr_cif_fwd_f32: bad shape B=%d T=%d H=%d L=%d", B, T, H, L);
    ASR_REQUIRE(hidden && alphas && n_fired && cur && rem && sched && alpha_sum, "asr_cif_fwd_f32: null pointer");
    ASR_REQUIRE(L == 0 || (out && fire_t), "asr_cif_fwd_f32: null output with L=%d", L);
    ASR_REQUIRE((long long)B * T < (1ll << 30), "asr_cif_fwd_f32: B*T too large");
    if (asr_device_ok() != 0) return 3;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CifFwdArgs a{hidden, alphas, threshold, B, T, H, L, out, fire_t, n_fired, cur, rem, sched, alpha_sum, target_num, qua_term};

    const bool vec4_ok = (H % 4 == 0) && aligned16(hidden) && (L == 0 || aligned16(out));
    int variant = kernel_hint > 0 ? kernel_hint : get_opt("cif_fwd_variant");      // per-call hint first, then the process-wide option
    bool auto_v2 = false;
    if (variant == 0) {
        // measured on B200 (tools/gpu_probe.py cif): with enough (utterance, 128-column) slices to keep
        // 4 one-warp CTAs per SM busy for several waves, the one-warp TMA pipeline (variant 2, width 128,
        // 3 stages = 48 KB) streams at 5.4 TB/s (B=256, T=1600, H=512: 164 us vs 198 us); below that the
        // warp-specialised kernel (variant 3) wins (B=64, T=3000: 105-111 us vs 124-137 us; B=128: 88 us).
        variant = (vec4_ok && T >= 64) ? 3 : 1;
        if (variant == 3 && (long long)B * ((H + 127) / 128) >= 6ll * num_sms()) {
            variant = 2;
            auto_v2 = true;
        }
    }
    if ((variant == 2 || variant == 3) && !vec4_ok) variant = 1;

    // slice width: widest that still yields >= 2 warps per SM
    int width = get_opt("cif_fwd_width");
    if (auto_v2) width = 128;
    if (width != 32 && width != 64 && width != 128) {
        const int want = 2 * num_sms();
        width = 128;
        while (width > 32 && (long long)B * ((H + width - 1) / width) < want) width >>= 1;
    }
    if (!vec4_ok) {
        // generic path: scalar or float2 lanes
        width = (H % 2 == 0 && (reinterpret_cast<uintptr_t>(hidden) & 7u) == 0 && (L == 0 || (reinterpret_cast<uintptr_t>(out) & 7u) == 0)) ? 64 : 32;
        if (width == 64 && (long long)B * ((H + 63) / 64) < 2 * num_sms()) width = 32;
    }
    const int nslices = (H + width - 1) / width;

    if (variant == 4) {
        cif_schedule_kernel<<<(B + 3) / 4, 128, 0, st>>>(a);
        ASR_LAUNCH_CHECK();
        if (L > 0) {
            const int rw = vec4_ok ? 128 : width;          // float4 lanes whenever the rows allow it
            const int rs = (H + rw - 1) / rw;
            const long long nw = (long long)B * L * rs;
            const unsigned blocks = (unsigned)((nw + 3) / 4);
            if (rw == 128) {
                cif_rows_kernel<4><<<blocks, 128, 0, st>>>(a, rs);
            } else if (rw == 64) {
                cif_rows_kernel<2><<<blocks, 128, 0, st>>>(a, rs);
            } else {
                cif_rows_kernel<1><<<blocks, 128, 0, st>>>(a, rs);
            }
            ASR_LAUNCH_CHECK();
        }
        return 0;
    }

    if (variant == 3) {
        // warp-specialised: CTA columns = nw * width, at most 256 (TMA box limit)
        int nw = get_opt("cif_fwd_rows");   // reused knob: data warps per CTA (0 = auto)
        int wv = get_opt("cif_fwd_width");
        if (wv != 32 && wv != 64 && wv != 128) wv = 64;   // float2 per lane: best measured (cfg 4)
        if (nw <= 0) nw = 4;
        while (nw * wv > 256) nw >>= 1;
        while (nw > 1 && (nw - 1) * wv >= H) --nw;   // no idle data warps for narrow H
        if (nw > 4) nw = 4;
        const int cw = nw * wv;
        const int ngroups = (H + cw - 1) / cw;
        CUtensorMap tmap;
        if (make_tmap_2d(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, hidden, (uint64_t)B * T, (uint64_t)H,
                         (uint64_t)H * 4, kCifRows, (uint32_t)cw, CU_TENSOR_MAP_SWIZZLE_NONE) != 0)
            return 4;
        int nstage = get_opt("cif_fwd_stages");
        if (nstage <= 0) nstage = (cw >= 256) ? 2 : 3;   // <= 64-96 KB per CTA: 2-3 CTAs (10-15 warps) per SM
        if (nstage > kCifMaxStages) nstage = kCifMaxStages;
        const size_t smem = (size_t)nstage * kCifRows * cw * 4;
        ASR_REQUIRE(B <= 65535, "asr_cif_fwd_f32: B=%d exceeds grid.y", B);
        dim3 grid(ngroups, B);
        const int threads = (nw + 1) * 32;
        if (wv == 128) {
            ASR_CHECK_CUDA(cudaFuncSetAttribute(cif_fwd_ws_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cif_fwd_ws_kernel<4><<<grid, threads, smem, st>>>(tmap, a, nstage, nw);
        } else if (wv == 64) {
            ASR_CHECK_CUDA(cudaFuncSetAttribute(cif_fwd_ws_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cif_fwd_ws_kernel<2><<<grid, threads, smem, st>>>(tmap, a, nstage, nw);
        } else {
            ASR_CHECK_CUDA(cudaFuncSetAttribute(cif_fwd_ws_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cif_fwd_ws_kernel<1><<<grid, threads, smem, st>>>(tmap, a, nstage, nw);
        }
        ASR_LAUNCH_CHECK();
        return 0;
    }

    if (variant == 2) {
        CUtensorMap tmap;
        if (make_tmap_2d(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, hidden, (uint64_t)B * T, (uint64_t)H,
                         (uint64_t)H * 4, kCifRows, (uint32_t)width, CU_TENSOR_MAP_SWIZZLE_NONE) != 0)
            return 4;
        int nstage = get_opt("cif_fwd_stages");
        if (nstage <= 0) nstage = auto_v2 ? 3 : 6;
        if (nstage > kCifMaxStages) nstage = kCifMaxStages;
        const size_t smem = (size_t)nstage * kCifRows * width * 4;
        dim3 grid(nslices, B);
        ASR_REQUIRE(B <= 65535, "asr_cif_fwd_f32: B=%d exceeds grid.y", B);
        if (width == 128) {
            ASR_CHECK_CUDA(cudaFuncSetAttribute(cif_fwd_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cif_fwd_tma_kernel<4><<<grid, 32, smem, st>>>(tmap, a, nstage);
        } else if (width == 64) {
            ASR_CHECK_CUDA(cudaFuncSetAttribute(cif_fwd_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cif_fwd_tma_kernel<2><<<grid, 32, smem, st>>>(tmap, a, nstage);
        } else {
            ASR_CHECK_CUDA(cudaFuncSetAttribute(cif_fwd_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cif_fwd_tma_kernel<1><<<grid, 32, smem, st>>>(tmap, a, nstage);
        }
        ASR_LAUNCH_CHECK();
        return 0;
    }

    const long long nwarps = (long long)B * nslices;
    const int blocks = (int)((nwarps + 3) / 4);
    if (width == 128) {
        cif_fwd_plain_kernel<4><<<blocks, 128, 0, st>>>(a, nslices);
    } else if (width == 64) {
        cif_fwd_plain_kernel<2><<<blocks, 128, 0, st>>>(a, nslices);
    } else {
        cif_fwd_plain_kernel<1><<<blocks, 128, 0, st>>>(a, nslices);
    }
    ASR_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t asr_cif_bwd_workspace_bytes(int B, int T) {
    return (size_t)(B > 0 ? B : 0) * (size_t)(T > 0 ? T : 0) * sizeof(float);
}

extern "C" int asr_cif_bwd_f32(const float* hidden, const float* g_out, const int* n_fired, const float* cur,
                               const float* rem, const int* sched, int B, int T, int H, int L, float* g_hidden,
                               float* g_alphas, void* ws, size_t ws_bytes, void* stream) {
    ASR_REQUIRE(B > 0 && T > 0 && H > 0 && L >= 0, "asr_cif_bwd_f32: bad shape B=%d T=%d H=%d L=%d", B, T, H, L);
    ASR_REQUIRE(hidden && n_fired && cur && rem && sched && g_hidden && g_alphas && ws, "asr_cif_bwd_f32: null pointer");
    ASR_REQUIRE(L == 0 || g_out, "asr_cif_bwd_f32: null g_out with L=%d", L);
    ASR_REQUIRE(ws_bytes >= asr_cif_bwd_workspace_bytes(B, T), "asr_cif_bwd_f32: workspace too small (%zu < %zu)",
                ws_bytes, asr_cif_bwd_workspace_bytes(B, T));
    if (asr_device_ok() != 0) return 3;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CifBwdArgs a{hidden, g_out, n_fired, cur, rem, sched, B, T, H, L, g_hidden, g_alphas, static_cast<float*>(ws)};
    const long long rows = (long long)B * T;
    const int blocks = (int)((rows + kBwdRowsPerCta - 1) / kBwdRowsPerCta);
    const bool v4 = (H % 4 == 0) && aligned16(hidden) && aligned16(g_hidden) && (L == 0 || aligned16(g_out));
    const bool v2 = (H % 2 == 0) && ((reinterpret_cast<uintptr_t>(hidden) | reinterpret_cast<uintptr_t>(g_hidden) |
                                      reinterpret_cast<uintptr_t>(g_out)) & 7u) == 0;
    if (v4) {
        cif_bwd_rows_kernel<4><<<blocks, 256, 0, st>>>(a);
    } else if (v2) {
        cif_bwd_rows_kernel<2><<<blocks, 256, 0, st>>>(a);
    } else {
        cif_bwd_rows_kernel<1><<<blocks, 256, 0, st>>>(a);
    }
    ASR_LAUNCH_CHECK();
    cif_bwd_scan_kernel<<<B, kScanThreads, 0, st>>>(g_alphas, static_cast<const float*>(ws), B, T);
    ASR_LAUNCH_CHECK();
    return 0;
}

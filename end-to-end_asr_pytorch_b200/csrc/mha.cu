// mha.cu - scaled-dot-product attention core on tcgen05 / TMEM, fed by TMA (sm_100a).
//
// Replaces ScaledDotProductAttention.forward
// (/root/reference/src/transformer/attention.py:74-86: bmm, / sqrt(d_k), masked_fill(-inf),
// softmax(dim=2), dropout, bmm) and the head split / merge copies around it
// (attention.py:47-49, 56-57).  bf16 operands, fp32 accumulation and softmax statistics.
//
// Layout: q [B,Lq,Hh,64], k,v [B,Lk,Hh,64] bf16 - the natural layout of the projection
// outputs - addressed through 4-D TMA descriptors, so no head-major copy is ever made;
// out [B,Lq,Hh,64] bf16 is exactly what the `fc` projection consumes.
//
// Forward kernel: one CTA per (b, head, 128-query tile), 6 warps:
//   warps 0-3  softmax + epilogue (thread t owns query row t = TMEM lane t)
//   warp  4    TMA producer (Q once, K/V tiles into a 2-stage ring, 128-byte swizzle)
//   warp  5    TMEM allocator + single-thread tcgen05.mma issuer
// Per 128-key block:  S = Q K^T  (tcgen05.mma kind::f16, M=128 N=128 K=64, SMEM x SMEM -> TMEM)
//                     softmax warps read S from TMEM (tcgen05.ld), apply the mask, online
//                     softmax, write P (bf16) into shared memory in the canonical K-major
//                     128B-swizzled layout
//                     PV = P V   (M=128 N=64 K=128, V consumed MN-major straight from its TMA tile)
//                     softmax warps read PV from TMEM and fold it into the fp32 O registers.
#include "common.cuh"
#include "tcgen05.cuh"
#include "philox.cuh"

#include <cuda_bf16.h>
#include <type_traits>

#include <algorithm>

namespace asr {


constexpr int kD = 64;          // head dim (d_k = d_v = 64 in every reference recipe)
constexpr int kBM = 128;        // query rows per CTA
constexpr int kBN = 128;        // keys per block
constexpr int kTileBytes = kBM * kD * 2;   // 16 KB: one [128 x 64] bf16 tile, 128-byte rows

struct MhaFwdArgs {
    const int* kv_len;          // [B] or null
    const uint8_t* dense_mask;  // [B,Lq,Lk] or null (non-zero = masked)
    int causal;
    int B, Hh, Lq, Lk;
    float scale_log2;           // softmax scale * log2(e)
    __nv_bfloat16* out;         // [B,Lq,Hh,64]
    float* lse;                 // [B,Hh,Lq]   natural-log LSE of the scaled scores
    // dropout on the probabilities (attention.py:83): element (b,h,q,k) is dropped when its random
    // byte is < drop_thresh; kept elements are scaled by inv_keep = 256 / (256 - drop_thresh)
    uint32_t drop_thresh;       // 0 = no dropout
    float inv_keep;
    uint32_t seed_lo, seed_hi;
    const uint64_t* seed_dev;   // not null: the seed is *seed_dev + (seed_hi:seed_lo), read on the device (a step captured in
                                // a CUDA graph gets fresh masks on every replay by bumping one device word)
};

// Keep masks of a packed bf16 pair (0xffff per kept lane) from two random bytes of `word` (bytes 2p, 2p + 1): a byte permute
// turns each byte into the fp16 number 1 + byte / 1024 (0x3c00 | byte), and ONE packed fp16 compare against 1 + thresh / 1024
// (`thr_h2` = 0x3c00 | thresh in both halves) yields the mask - two instructions per pair where extract / compare / select
// per key were six.
__device__ __forceinline__ uint32_t keep_mask2(uint32_t word, int p, uint32_t thr_h2) {
    const uint32_t x = __byte_perm(word, 0x3c3c3c3cu, p ? 0x4342 : 0x4140);
    uint32_t m;
    asm("set.ge.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(x), "r"(thr_h2));
    return m;
}

struct __align__(8) MhaBarriers {
    uint64_t q_full;
    uint64_t kv_full[2];
    uint64_t kv_empty[2];
    uint64_t s_full;
    uint64_t s_free;
    uint64_t p_full;
    uint64_t pv_full;
    uint32_t tmem_base;
    uint32_t pad;
};

// Q + 2x(K,V) + P + barriers = 112.1 KB, so that two CTAs (and their 2 x 256 TMEM columns) share one SM
constexpr int kFwdSmem = kTileBytes /*Q*/ + 4 * kTileBytes /*K,V x2*/ + 2 * kTileBytes /*P*/ + 128 /*barriers*/ + 128 /*ones*/;

// ---- forward, eight softmax warps per tile --------------------------------------------------
// 10 warps: softmax warps 0-7 (warp w: TMEM lane quarter w % 4 = query rows, key half w / 4 of
// every 128-key block), TMA producer (warp 8), MMA issuer (warp 9); two CTAs per SM.
// A query row is shared by two threads: each takes the maximum over its 64 scores, the two
// halves meet through shared memory and a 256-thread named barrier, each exponentiates its
// half and owns 32 of the 64 output columns.  Four softmax warps per SM sub-partition (two
// CTAs) hide the dependent-instruction latency that one warp per sub-partition leaves exposed.
// The tensor pipe is fed out of order with respect to the tiles: S of block j+1 is issued as
// soon as S of block j has been read (before P V of block j), and O is updated with P V of
// block j-1 while block j is in flight, so neither product is waited for right after its issue.
constexpr int kFwd3Threads = 320;
constexpr int kFwd3Smem = kFwdSmem + 2 * 128 * 2 /*row maxima of the two halves, bf16*/;
static_assert(2 * (kFwd3Smem + 1024) <= 233472, "two CTAs of the forward kernel must fit one SM");

template <bool DROP>
__global__ void __launch_bounds__(kFwd3Threads, 2)
mha_fwd3_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const MhaFwdArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint32_t seed_lo = a.seed_lo, seed_hi = a.seed_hi;
    if (DROP) effective_seed(a.seed_dev, seed_lo, seed_hi);
    unsigned char* sQ = smem;
    unsigned char* sK = sQ + kTileBytes;          // 2 stages
    unsigned char* sV = sK + 2 * kTileBytes;      // 2 stages
    unsigned char* sP = sV + 2 * kTileBytes;      // [2 key halves][128 rows][128 B]
    MhaBarriers* bars = reinterpret_cast<MhaBarriers*>(sP + 2 * kTileBytes);
    uint32_t* sOnes = reinterpret_cast<uint32_t*>(sP + 2 * kTileBytes + 128);   // 128 bytes of bf16 1.0
    __nv_bfloat16* sMax = reinterpret_cast<__nv_bfloat16*>(sP + 2 * kTileBytes + 256);   // [2 halves][128 rows]

    // Row sums: each softmax thread adds up the (bf16-rounded) probabilities it produces; the tensor-core
    // alternative (P x ones, eight more MMAs per block that re-read the P tile) measured 14 % slower.
    constexpr bool kUseOnes = false;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) sOnes[threadIdx.x] = 0x3F803F80u;
    fence_proxy_async();
    const int q0 = blockIdx.x * kBM;
    const int h = blockIdx.y;
    const int b = blockIdx.z;
    const int kvlen = a.kv_len ? min(max(__ldg(a.kv_len + b), 0), a.Lk) : a.Lk;
    int k_end = a.causal ? min(kvlen, q0 + kBM) : kvlen;
    if (a.dense_mask) k_end = a.Lk;
    const int nblk = max(1, (k_end + kBN - 1) / kBN);

    if (threadIdx.x == 0) {
        mbar_init(&bars->q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->kv_full[s], 1);
            mbar_init(&bars->kv_empty[s], 1);
        }
        mbar_init(&bars->s_full, 1);
        mbar_init(&bars->s_free, 8);      // one arrival per softmax warp
        mbar_init(&bars->p_full, 8);
        mbar_init(&bars->pv_full, 1);
        fence_mbar_init();
    }
    if (warp == 9) {
        tmem_alloc(&bars->tmem_base, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    const uint32_t tmem_s = tmem;          // 128 columns: S
    const uint32_t tmem_pv = tmem + 128;   // 64 columns: P V
    const uint32_t tmem_l = tmem + 192;    // 16 columns: P x ones (row sums)

    if (warp == 8) {
        // ===== TMA producer =====
        if (elect_one_sync()) {
            tma_prefetch_desc(&tm_q);
            tma_prefetch_desc(&tm_k);
            tma_prefetch_desc(&tm_v);
            mbar_arrive_expect_tx(&bars->q_full, kTileBytes);
            tma_load_4d(sQ, &tm_q, 0, h, q0, b, &bars->q_full);
            for (int j = 0; j < nblk; ++j) {
                const int s = j & 1;
                if (j >= 2) mbar_wait(&bars->kv_empty[s], ((j >> 1) - 1) & 1);
                mbar_arrive_expect_tx(&bars->kv_full[s], 2 * kTileBytes);
                tma_load_4d(sK + s * kTileBytes, &tm_k, 0, h, j * kBN, b, &bars->kv_full[s]);
                tma_load_4d(sV + s * kTileBytes, &tm_v, 0, h, j * kBN, b, &bars->kv_full[s]);
            }
        }
    } else if (warp == 9) {
        // ===== MMA issuer (one thread) =====
        if (elect_one_sync()) {
            constexpr uint32_t idesc_s = make_idesc(kBM, kBN, 0, 0);    // S = Q K^T : A, B K-major
            constexpr uint32_t idesc_pv = make_idesc(kBM, kD, 0, 1);    // PV = P V  : A K-major, B MN-major
            constexpr uint32_t idesc_l = make_idesc(kBM, 16, 0, 0);     // row sums = P x ones
            const uint32_t q_addr = smem_u32(sQ);
            const uint32_t p_addr = smem_u32(sP);
            const uint64_t ones_desc = smem_desc_ones(smem_u32(sOnes));
            fence_proxy_async();
            auto issue_s = [&](int j) {
                const int s = j & 1;
                mbar_wait(&bars->kv_full[s], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sK + s * kTileBytes);
#pragma unroll
                for (int kk = 0; kk < kD / 16; ++kk)
                    umma_bf16(tmem_s, smem_desc_sw128(q_addr + kk * 32, 16, 1024), smem_desc_sw128(k_addr + kk * 32, 16, 1024),
                              idesc_s, kk > 0 ? 1u : 0u);
                tc_commit(&bars->s_full);
            };
            mbar_wait(&bars->q_full, 0);
            issue_s(0);
            for (int j = 0; j < nblk; ++j) {
                const int s = j & 1;
                const uint32_t v_addr = smem_u32(sV + s * kTileBytes);
                if (j + 1 < nblk) {
                    mbar_wait(&bars->s_free, j & 1);      // S of block j has been read
                    issue_s(j + 1);
                }
                mbar_wait(&bars->p_full, j & 1);          // P of block j is in shared memory, P V of block j-1 has been read
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < kBN / 16; ++kk) {
                    const uint64_t ad = smem_desc_sw128(p_addr + (kk >> 2) * kTileBytes + (kk & 3) * 32, 16, 1024);
                    const uint64_t bd = smem_desc_sw128(v_addr + kk * 2048, kTileBytes, 1024);
                    umma_bf16(tmem_pv, ad, bd, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);   // O accumulates over the blocks
                    if (kUseOnes) umma_bf16(tmem_l, ad, ones_desc, idesc_l, (j > 0 || kk > 0) ? 1u : 0u);
                }
                tc_commit(&bars->pv_full);
                tc_commit(&bars->kv_empty[s]);
            }
        }
    } else {
        // ===== softmax + epilogue: thread = (query row, key half) =====
        // O and the row sums accumulate in TMEM across the key blocks (P V and P x ones issued with
        // accumulate).  They are scaled with m_used, the row maximum at the last rescale; a new
        // maximum only forces a rescale (TMEM load, multiply, TMEM store) when it exceeds m_used by
        // more than 8 in the exponent - otherwise the probabilities simply run up to 2^8, which
        // bf16 P and the fp32 accumulators hold without loss.
        const int row = (warp & 3) * 32 + lane;
        const int half = warp >> 2;
        const int qi = q0 + row;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        float m_used = -INFINITY;
        float l_part = 0.0f;         // DROP: sum of this thread's (undropped) probabilities, scaled like O
        const float c = a.scale_log2;
        const uint8_t* mrow = a.dense_mask ? a.dense_mask + ((size_t)b * a.Lq + min(qi, a.Lq - 1)) * a.Lk : nullptr;
        for (int j = 0; j < nblk; ++j) {
            const int key0 = j * kBN + half * 64;
            int lim = kvlen;
            if (a.causal) lim = min(lim, qi + 1);
            const bool need_mask = (j * kBN + kBN > lim) || (mrow != nullptr);
            // dead keys among this thread's 64 as two 32-bit masks (length / causal limit by arithmetic, a dense mask by a
            // rolled byte loop: unrolled per key it was most of the kernel's code)
            uint32_t dead_lo = 0, dead_hi = 0;
            if (__builtin_expect(need_mask, 0)) {
                const int live = lim - key0;
                dead_lo = live >= 32 ? 0u : live <= 0 ? 0xffffffffu : (0xffffffffu << live);
                dead_hi = live >= 64 ? 0u : live <= 32 ? 0xffffffffu : (0xffffffffu << (live - 32));
                if (mrow != nullptr) {
                    const int kmax = min(64, a.Lk - key0);
#pragma unroll 1
                    for (int kk = 0; kk < kmax; ++kk) {
                        const uint32_t bit = (mrow[key0 + kk] != 0) ? (1u << (kk & 31)) : 0u;
                        dead_lo |= (kk < 32) ? bit : 0u;
                        dead_hi |= (kk < 32) ? 0u : bit;
                    }
                }
            }
            auto apply_mask = [&](uint32_t (&r)[32], int cc) {
                const uint32_t dead = cc ? dead_hi : dead_lo;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if ((dead >> i) & 1u) r[i] = 0xff800000u;   // -inf
            };
            mbar_wait(&bars->s_full, j & 1);
            tc_fence_after();
            // pass 1: maximum over this thread's 64 scores, then over both halves of the row
            float m_half = -INFINITY;
            {
                uint32_t ra[32], rb[32];
                tmem_ld32_issue(tmem_s + lane_base + half * 64, ra);
                tmem_ld32_issue(tmem_s + lane_base + half * 64 + 32, rb);
                tmem_ld_wait();
                if (need_mask) {
                    apply_mask(ra, 0);
                    apply_mask(rb, 32);
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) m_half = fmaxf(m_half, fmaxf(__uint_as_float(ra[i]), __uint_as_float(rb[i])));
            }
            // (any common reference works for the exponent, so the halves trade bf16-rounded maxima:
            // 512 bytes instead of 1 KB keeps two CTAs on one SM; the next write is ordered behind
            // this read by the s_free -> s_full chain)
            const __nv_bfloat16 m_half_r = __float2bfloat16_rn(m_half);
            sMax[half * 128 + row] = m_half_r;
            bar_sync_named(1, 256);
            const float m_new = fmaxf(m_used, fmaxf(__bfloat162float(m_half_r), __bfloat162float(sMax[(half ^ 1) * 128 + row])));
            // P V of the previous block must be complete before its P tile is overwritten (and before O is rescaled)
            if (j > 0) {
                mbar_wait(&bars->pv_full, (j - 1) & 1);
                tc_fence_after();
            }
            const bool grow = (j > 0) && ((m_new - m_used) * c > 8.0f);   // same answer in both threads of the row
            if (j == 0) {
                m_used = m_new;
            } else if (__any_sync(0xffffffffu, grow)) {          // TMEM accesses are warp-collective: all lanes go
                const float f = grow ? ex2_approx((m_used - m_new) * c) : 1.0f;   // m_used = -inf -> 0
                uint32_t r[32];
                tmem_ld32_issue(tmem_pv + lane_base + half * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
                tmem_st32(tmem_pv + lane_base + half * 32, r);
                if (!kUseOnes) {
                    l_part *= f;
                } else if (half == 0) {
                    tmem_st1(tmem_l + lane_base, tmem_ld1(tmem_l + lane_base) * f);
                }
                tmem_st_wait();
                if (grow) m_used = m_new;
            }
            const float mc = ((m_used == -INFINITY) ? 0.0f : m_used) * c;   // fully masked so far: keep exp2 finite
            // pass 2: probabilities -> bf16 -> shared memory (K-major, 128B swizzle)
            unsigned char* prow = sP + half * kTileBytes + row * 128;
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                uint32_t r[32];
                tmem_ld32_issue(tmem_s + lane_base + half * 64 + part * 32, r);
                tmem_ld_wait();
                if (need_mask) apply_mask(r, part * 32);
                uint32_t pk[16];
                if (DROP) {
                    // fp32 exponentials (packed f32x2 arithmetic around them): their sum is the softmax normaliser, the
                    // dropped and rescaled copy goes to the P tile
                    const uint64_t c2 = pack_f32x2(c, c), mc2 = pack_f32x2(-mc, -mc);
                    uint64_t l2 = pack_f32x2(0.0f, 0.0f);
#pragma unroll
                    const uint32_t thr_h2 = (0x3c00u | a.drop_thresh) * 0x00010001u;
                    for (int g = 0; g < 2; ++g) {
                        const uint4 rnd = philox16((uint32_t)(key0 + part * 32 + g * 16) >> 4, (uint32_t)qi, (uint32_t)(b * a.Hh + h),
                                                   seed_lo, seed_hi);
                        // the keep mask of a packed pair (keep_mask2) is ANDed into the packed probabilities; the 1 / keep scale
                        // is applied once, in the epilogue
                        const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            float x0, x1;
                            unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[g * 16 + i]), __uint_as_float(r[g * 16 + i + 1])), c2, mc2), x0, x1);
                            const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
                            l2 = add_f32x2(l2, pack_f32x2(p0, p1));
                            pk[(g * 16 + i) >> 1] = cvt_bf16x2(p0, p1) & keep_mask2(rw[i >> 2], (i >> 1) & 1, thr_h2);
                        }
                    }
                    float la, lb;
                    unpack_f32x2(l2, la, lb);
                    l_part += la + lb;
                } else {
                    float ls = 0.0f;
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const uint32_t w = ex2_bf16x2(fmaf(__uint_as_float(r[i]), c, -mc), fmaf(__uint_as_float(r[i + 1]), c, -mc));
                        pk[i >> 1] = w;
                        // the normaliser is the sum of the probabilities P V really uses: the bf16 values
                        if (!kUseOnes) ls += __uint_as_float(w << 16) + __uint_as_float(w & 0xffff0000u);
                    }
                    l_part += ls;
                }
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int chunk = (part * 4 + q4) ^ (row & 7);
                    *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
                }
            }
            tc_fence_before();
            mbar_arrive_warp(&bars->s_free);       // S may be overwritten by the next Q K^T
            fence_proxy_async();              // P (generic proxy) -> visible to the tensor core (async proxy)
            mbar_arrive_warp(&bars->p_full);
        }
        mbar_wait(&bars->pv_full, (nblk - 1) & 1);
        tc_fence_after();
        // epilogue: O / l -> bf16 -> out[b, qi, h, half*32 ..]; a fully masked row is 0/0 = NaN like the reference
        if (!kUseOnes) {
            // the two halves of a row add their normaliser parts through two spare accumulator columns
            tmem_st1(tmem_l + lane_base + half, l_part);
            tmem_st_wait();
            tc_fence_before();
            bar_sync_named(1, 256);
            tc_fence_after();
        }
        {
            uint32_t r[32];
            tmem_ld32_issue(tmem_pv + lane_base + half * 32, r);
            float l_run = tmem_ld1(tmem_l + lane_base);
            if (!kUseOnes) l_run += tmem_ld1(tmem_l + lane_base + 1);
            tc_fence_before();
            if (qi < a.Lq) {
                const float inv = (DROP ? a.inv_keep : 1.0f) / l_run;      // kept probabilities went into P V unscaled
                __nv_bfloat16* dst = a.out + (((size_t)b * a.Lq + qi) * a.Hh + h) * kD + half * 32;
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint32_t w[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const __nv_bfloat162 v2 = __floats2bfloat162_rn(__uint_as_float(r[i + 2 * u]) * inv, __uint_as_float(r[i + 2 * u + 1]) * inv);
                        w[u] = *reinterpret_cast<const uint32_t*>(&v2);
                    }
                    *reinterpret_cast<uint4*>(dst + i) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                if (a.lse != nullptr && half == 0)
                    a.lse[((size_t)b * a.Hh + h) * a.Lq + qi] = (m_used * c + log2f(l_run)) * 0.6931471805599453f;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

#ifdef ASR_MHA_TRACE
// clock64 stamps of CTA (0,0,0): [tile][block < 16][point < 20], read back with asr_debug_mha_trace (tools/mha_trace.py)
__device__ long long g_mha_trace[2 * 16 * 20];
// per-CTA timeline of mha_fwd8_kernel (tools/mha_cta_timeline.py): [linear CTA id < 2048][sm id, entry, softmax role start,
// last block done, exit]
__device__ long long g_mha_cta[2048 * 5];
__device__ __forceinline__ void mha_cta_stamp(int slot) {
    const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    if (cta < 2048) {
        if (slot == 1) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            g_mha_cta[cta * 5] = smid;
        }
        g_mha_cta[cta * 5 + slot] = clock64();
    }
}
#define MHA_CTA_STAMP(slot) do { if (threadIdx.x == 0) mha_cta_stamp(slot); } while (0)
#define MHA_TRACE(t, j, k)                                                                    \
    do {                                                                                      \
        if (lane == 0 && (warp & 3) == 0 && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && (j) < 16) \
            g_mha_trace[((t) * 16 + (j)) * 20 + (k)] = clock64();                             \
    } while (0)
#define MHA_TRACE_WARP(t, j, k)                                                               \
    do {                                                                                      \
        if (lane == 0 && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && (j) < 16)             \
            g_mha_trace[((t) * 16 + (j)) * 20 + (k) + (warp & 3)] = clock64();                \
    } while (0)
#define MHA_TRACE_MMA(t, j, k)                                                                \
    do {                                                                                      \
        if ((blockIdx.x | blockIdx.y | blockIdx.z) == 0 && (j) < 16) g_mha_trace[((t) * 16 + (j)) * 20 + (k)] = clock64(); \
    } while (0)
#else
#define MHA_CTA_STAMP(slot) do { } while (0)
#define MHA_TRACE(t, j, k) do { } while (0)
#define MHA_TRACE_MMA(t, j, k) do { } while (0)
#define MHA_TRACE_WARP(t, j, k) do { } while (0)
#endif

constexpr int kFwd6Stages = 4;   // K and V rings: a TMA load takes longer than a block of work, so it is issued three blocks ahead

// ---- forward, two query tiles per CTA, one thread per query row, scores read from TMEM once ---------------
// Tensor memory has ONE read port of 64 B/clk per SM (measured with clock64 traces): the softmax warps'
// tcgen05.ld and the A-operand reads of TMEM-sourced MMAs share it, so a 128 x 128 fp32 score tile costs
// 1024 cycles per read.  Here the only TMEM reader is ONE load of the scores: a softmax thread owns a whole
// query row, its 128 scores go to registers once (setmaxnreg moves registers from the TMA / MMA warpgroup
// to the softmax warpgroups: 224 per thread), the row maximum needs no exchange with other threads, P goes
// to shared memory (K-major, 128-byte swizzle) and the row sums are accumulated by the row's own thread
// from the packed bf16 probabilities it has just produced (no P x ones product).  The scores sit in
// registers ~50 cycles after S is ready, so S is handed back at once and Q K^T of the next block runs under
// the current block's exponentials.  Two tiles per CTA (warps 0-3 and 4-7) share every K/V tile (4-deep TMA
// rings for K and for V: a load takes longer than a block of work) and run in opposite phases.  O
// accumulates in TMEM with the lazy rescale of mha_fwd3_kernel.
// TMEM (512 columns): S_A 0-127 | S_B 128-255 | O_A 256-319 | O_B 320-383 | P_A 384-447 | P_B 448-511.
struct __align__(8) MhaBarriers8 {
    uint64_t q_full;
    uint64_t k_full[kFwd6Stages];
    uint64_t k_empty[kFwd6Stages];
    uint64_t v_full[kFwd6Stages];
    uint64_t v_empty[kFwd6Stages];
    uint64_t s_full[2];     // per tile
    uint64_t s_free[2];
    uint64_t p_full[2];
    uint64_t pv_full[2];
    uint32_t tmem_base;
    uint32_t pad;
};
constexpr int kFwd8Threads = 384;   // 8 softmax warps + one utility warpgroup (TMA, MMA, two idle warps)
// Q x2 + K ring + V ring + barriers (P lives in tensor memory)
constexpr int kFwd8Smem = (2 + 2 * kFwd6Stages) * kTileBytes + 256;
static_assert(kFwd8Smem <= 232448, "shared memory of the forward kernel");

// The two tiles take turns on the XU pipe (named-barrier token), which keeps them in opposite phases: one loads its scores
// and finds the row maxima while the other runs its exponentials.  Exponentials in fp32 with packed f32x2 arithmetic
// (FFMA2 / FADD2), row sums from the unrounded fp32 values; P stays in tensor memory (own columns) as the A operand of
// P V; one MMA issuer per tile.  (Round 2 measured the alternatives - P through shared memory, one issuer for both tiles,
// bf16x2 exponentials, no token: DESIGN.md 4 keeps the numbers - and removed them.)
template <bool DROP>
__global__ void __launch_bounds__(kFwd8Threads, 1)
mha_fwd8_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const MhaFwdArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    MHA_CTA_STAMP(1);
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint32_t seed_lo = a.seed_lo, seed_hi = a.seed_hi;
    if (DROP) effective_seed(a.seed_dev, seed_lo, seed_hi);
    unsigned char* sQ = smem;                              // 2 tiles
    unsigned char* sK = sQ + 2 * kTileBytes;               // kFwd6Stages tiles
    unsigned char* sV = sK + kFwd6Stages * kTileBytes;     // kFwd6Stages tiles
    MhaBarriers8* bars = reinterpret_cast<MhaBarriers8*>(sV + kFwd6Stages * kTileBytes);
    static_assert(sizeof(MhaBarriers8) <= 256, "barrier block");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 2 * kBM;
    const int h = blockIdx.y;
    const int b = blockIdx.z;
    const int kvlen = a.kv_len ? min(max(__ldg(a.kv_len + b), 0), a.Lk) : a.Lk;
    const int ntile = (q0 + kBM < a.Lq) ? 2 : 1;           // the second tile may lie entirely beyond Lq
    // key blocks each tile needs: up to kvlen, and up to its last query row when causal
    int nb[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        int k_end = a.causal ? min(kvlen, q0 + t * kBM + kBM) : kvlen;
        if (a.dense_mask) k_end = a.Lk;
        nb[t] = (t < ntile) ? max(1, (k_end + kBN - 1) / kBN) : 0;
    }
    const int nblk = max(nb[0], nb[1]);                    // nb[1] >= nb[0] whenever tile 1 exists

    if (threadIdx.x == 0) {
        mbar_init(&bars->q_full, 1);
        for (int s = 0; s < kFwd6Stages; ++s) {
            mbar_init(&bars->k_full[s], 1);
            // one MMA issuer per tile, each of them releases every stage once
            mbar_init(&bars->k_empty[s], ntile);
            mbar_init(&bars->v_full[s], 1);
            mbar_init(&bars->v_empty[s], ntile);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&bars->s_full[t], 1);
            mbar_init(&bars->s_free[t], 4);       // one arrival per softmax warp of the tile
            mbar_init(&bars->p_full[t], 4);
            mbar_init(&bars->pv_full[t], 1);
        }
        fence_mbar_init();
    }
    constexpr int kTmaWarp = 8, kMmaWarp = 9;
    // (register reallocation between warpgroups at the head of every role: the softmax threads hold a
    // whole row of scores, the utility warpgroup needs next to nothing)
    if (warp == kMmaWarp) {
        tmem_alloc(&bars->tmem_base, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == kTmaWarp) {
        // ===== TMA producer =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (elect_one_sync()) {
            tma_prefetch_desc(&tm_q);
            tma_prefetch_desc(&tm_k);
            tma_prefetch_desc(&tm_v);
            mbar_arrive_expect_tx(&bars->q_full, ntile * kTileBytes);
            for (int t = 0; t < ntile; ++t) tma_load_4d(sQ + t * kTileBytes, &tm_q, 0, h, q0 + t * kBM, b, &bars->q_full);
            for (int j = 0; j < nblk; ++j) {
                const int s = j % kFwd6Stages;
                const int use = j / kFwd6Stages;
                if (use > 0) mbar_wait(&bars->k_empty[s], (use - 1) & 1);    // freed by the last Q K^T of block j - stages
                mbar_arrive_expect_tx(&bars->k_full[s], kTileBytes);
                tma_load_4d(sK + s * kTileBytes, &tm_k, 0, h, j * kBN, b, &bars->k_full[s]);
                if (use > 0) mbar_wait(&bars->v_empty[s], (use - 1) & 1);    // freed by the last P V of block j - stages
                mbar_arrive_expect_tx(&bars->v_full[s], kTileBytes);
                tma_load_4d(sV + s * kTileBytes, &tm_v, 0, h, j * kBN, b, &bars->v_full[s]);
            }
        }
    } else if (warp == kMmaWarp || warp == kMmaWarp + 1) {
        // ===== MMA issuers, one elected thread per tile =====
        // A single issuer walking S(0) S(1) PV(0) PV(1) in order spends ~100 cycles per Q K^T MMA, ~70 per P V
        // MMA and 100-200 per (already complete) mbarrier wait - its sub-partition is shared with two busy
        // softmax warps - which adds up to the whole period of a key block.  Two issuers on two sub-partitions
        // halve that chain, and a tile's P V no longer queues behind the other tile's waits.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        const int t = warp - kMmaWarp;
        if (t < ntile && elect_one_sync()) {
            constexpr uint32_t idesc_s = make_idesc(kBM, kBN, 0, 0);    // S = Q K^T : A, B K-major
            constexpr uint32_t idesc_pv = make_idesc(kBM, kD, 0, 1);    // PV = P V  : A K-major, B MN-major
            const uint32_t q_addr = smem_u32(sQ + t * kTileBytes);
            const uint32_t tmem_s = tmem + 128 * t;
            const uint32_t tmem_o = tmem + 256 + 64 * t;
            const uint32_t tmem_p = tmem + 384 + 64 * t;
            const int nbt = nb[t];
            // S of block j, or just the release of its K stage when this tile does not need the block
            auto issue_s = [&](int j) {
                const int ks = j % kFwd6Stages;
                mbar_wait(&bars->k_full[ks], (j / kFwd6Stages) & 1);
                if (j >= nbt) {
                    mbar_arrive(&bars->k_empty[ks]);
                    return;
                }
                if (j > 0) mbar_wait(&bars->s_free[t], (j - 1) & 1);      // scores of block j-1 are in registers
                tc_fence_after();
                MHA_TRACE_MMA(t, j, 10);
                const uint32_t k_addr = smem_u32(sK + ks * kTileBytes);
#pragma unroll
                for (int kk = 0; kk < kD / 16; ++kk)
                    umma_bf16(tmem_s, smem_desc_sw128(q_addr + kk * 32, 16, 1024), smem_desc_sw128(k_addr + kk * 32, 16, 1024),
                              idesc_s, kk > 0 ? 1u : 0u);
                tc_commit(&bars->s_full[t]);
                tc_commit(&bars->k_empty[ks]);
                MHA_TRACE_MMA(t, j, 11);
            };
            mbar_wait(&bars->q_full, 0);
            issue_s(0);
            for (int j = 0; j < nblk; ++j) {
                const int vs = j % kFwd6Stages;
                if (j + 1 < nblk) issue_s(j + 1);
                MHA_TRACE_MMA(t, j, 16);
                mbar_wait(&bars->v_full[vs], (j / kFwd6Stages) & 1);
                MHA_TRACE_MMA(t, j, 17);
                if (j >= nbt) {
                    mbar_arrive(&bars->v_empty[vs]);
                    continue;
                }
                mbar_wait(&bars->p_full[t], j & 1);
                tc_fence_after();
                MHA_TRACE_MMA(t, j, 8);
                const uint32_t v_addr = smem_u32(sV + vs * kTileBytes);
#pragma unroll
                for (int kk = 0; kk < kBN / 16; ++kk) {
                    const uint64_t bd = smem_desc_sw128(v_addr + kk * 2048, kTileBytes, 1024);
                    umma_bf16_ts(tmem_o, tmem_p + 8 * kk, bd, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
                }
                tc_commit(&bars->pv_full[t]);
                tc_commit(&bars->v_empty[vs]);
                MHA_TRACE_MMA(t, j, 9);
            }
        }
    } else if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");   // the two idle warps of the utility warpgroup
    } else {
        // ===== softmax + epilogue: warps 0-3 own tile 0, warps 4-7 tile 1; thread = query row =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        const int t = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        if (t < ntile) {
            const int qi = q0 + t * kBM + row;
            const uint8_t* mrow = a.dense_mask ? a.dense_mask + ((size_t)b * a.Lq + min(qi, a.Lq - 1)) * a.Lk : nullptr;
            const uint32_t tmem_s = tmem + 128 * t + lane_base;
            const uint32_t tmem_o = tmem + 256 + 64 * t + lane_base;
            float m_used = -INFINITY;
            float l_run = 0.0f;          // sum of the row's (undropped) probabilities, scaled like O
            const float c = a.scale_log2;
            const int nbt = nb[t];
            // token ring of the two tiles: tile 0 waits on barrier 1, tile 1 on barrier 2; a tile holds
            // the token while it runs its exponentials and then hands it to the other tile, so the two tiles
            // cannot fall into lockstep on the XU pipe.  Both tiles walk nblk turns (a tile that has run out of
            // key blocks just passes the token on); tile 1 grants the first turn.
            const bool pp = ntile == 2;
            const int nturn = pp ? nblk : nbt;
            MHA_CTA_STAMP(2);
            if (pp && t == 1) bar_arrive_named(1, 256);
            for (int j = 0; j < nturn; ++j) {
                if (j >= nbt) {
                    bar_sync_named(1 + t, 256);
                    if (t == 0 || j + 1 < nturn) bar_arrive_named(2 - t, 256);
                    continue;
                }
                const int key0 = j * kBN;
                int lim = kvlen;
                if (a.causal) lim = min(lim, qi + 1);
                const bool need_mask = (key0 + kBN > lim) || (mrow != nullptr);
                MHA_TRACE(t, j, 0);
                mbar_wait(&bars->s_full[t], j & 1);
                tc_fence_after();
                MHA_TRACE(t, j, 1);
                uint32_t r[4][32];           // the row's 128 scores: read from TMEM exactly once
#pragma unroll
                for (int q = 0; q < 4; ++q) tmem_ld32_issue(tmem_s + 32 * q, r[q]);
                tmem_ld_wait();
#ifdef ASR_MHA_TRACE
                if ((r[0][0] ^ r[1][0] ^ r[2][0] ^ r[3][31]) == 0x7fc54321u) __trap();   // the stamp waits for the data
#endif
                MHA_TRACE(t, j, 2);
                tc_fence_before();
                mbar_arrive_warp(&bars->s_free[t]);      // the scores are in registers: S may be overwritten
                float m_q[4];
                if (__builtin_expect(need_mask, 0)) {
                    // dead keys of this row as four 32-bit masks: the length / causal limit by arithmetic, a dense mask by a
                    // ROLLED byte loop (unrolled it is 128 x (address, LDG.U8, compare, select): a quarter of the kernel's code)
                    uint32_t dead[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int live = lim - (key0 + 32 * q);                 // keys [0, live) of this chunk are inside the limit
                        dead[q] = live >= 32 ? 0u : live <= 0 ? 0xffffffffu : (0xffffffffu << live);
                    }
                    if (mrow != nullptr) {
                        const int kmax = min(kBN, a.Lk - key0);
#pragma unroll 1
                        for (int kk = 0; kk < kmax; ++kk) {
                            const uint32_t bit = (mrow[key0 + kk] != 0) ? (1u << (kk & 31)) : 0u;
                            const int q = kk >> 5;
                            dead[0] |= (q == 0) ? bit : 0u;
                            dead[1] |= (q == 1) ? bit : 0u;
                            dead[2] |= (q == 2) ? bit : 0u;
                            dead[3] |= (q == 3) ? bit : 0u;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if ((dead[q] >> i) & 1u) r[q][i] = 0xff800000u;   // -inf
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    m_q[q] = -INFINITY;          // four independent chains
#pragma unroll
                    for (int i = 0; i < 32; ++i) m_q[q] = fmaxf(m_q[q], __uint_as_float(r[q][i]));
                }
                const float m_blk = fmaxf(fmaxf(m_q[0], m_q[1]), fmaxf(m_q[2], m_q[3]));
                const float m_new = fmaxf(m_used, m_blk);
                const bool grow = (j > 0) && ((m_new - m_used) * c > 8.0f);
                bool pv_waited = false;
                MHA_TRACE(t, j, 3);
                if (j == 0) {
                    m_used = m_new;
                } else if (__builtin_expect(__any_sync(0xffffffffu, grow), 0)) {      // TMEM accesses are warp-collective: all lanes go
                    // O is being accumulated by P V of this tile's previous block: wait for it before touching O
                    mbar_wait(&bars->pv_full[t], (j - 1) & 1);
                    tc_fence_after();
                    pv_waited = true;
                    const float f = grow ? ex2_approx((m_used - m_new) * c) : 1.0f;   // m_used = -inf -> 0
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint32_t o32[32];
                        tmem_ld32_issue(tmem_o + 32 * hh, o32);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o32[i] = __float_as_uint(__uint_as_float(o32[i]) * f);
                        tmem_st32(tmem_o + 32 * hh, o32);
                    }
                    l_run *= f;
                    tmem_st_wait();
                    if (grow) m_used = m_new;
                }
                const float mc = ((m_used == -INFINITY) ? 0.0f : m_used) * c;   // fully masked so far: keep exp2 finite
                MHA_TRACE(t, j, 4);
                // The P tile is free once P V of the previous block has read it.
                if (j > 0 && !pv_waited) mbar_wait(&bars->pv_full[t], (j - 1) & 1);
                float lsum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                uint32_t pk_all[4][16];
                // 32 keys = 16 packed columns of the row's TMEM lane
                auto store_chunk = [&](int q) { tmem_st16(tmem + 384 + 64 * t + lane_base + 16 * q, pk_all[q]); };
                if (!DROP) {
                    // x = s * scale - max in place, packed; no XU work yet
                    const uint64_t c2 = pack_f32x2(c, c);
                    const uint64_t mc2 = pack_f32x2(-mc, -mc);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            float x0, x1;
                            unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[q][i]), __uint_as_float(r[q][i + 1])), c2, mc2), x0, x1);
                            r[q][i] = __float_as_uint(x0);
                            r[q][i + 1] = __float_as_uint(x1);
                        }
                    }
                    if (pp) bar_sync_named(1 + t, 256);          // this tile's turn on the XU pipe
                    MHA_TRACE(t, j, 5);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint64_t ls2 = pack_f32x2(0.0f, 0.0f);
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const float p0 = ex2_approx(__uint_as_float(r[q][i]));
                            const float p1 = ex2_approx(__uint_as_float(r[q][i + 1]));
                            ls2 = add_f32x2(ls2, pack_f32x2(p0, p1));
                            pk_all[q][i >> 1] = cvt_bf16x2(p0, p1);
                        }
                        float l0, l1;
                        unpack_f32x2(ls2, l0, l1);
                        lsum[q] = l0 + l1;
                        store_chunk(q);
                    }
                } else {
                    if (pp) bar_sync_named(1 + t, 256);
                    MHA_TRACE(t, j, 5);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t (&pk)[16] = pk_all[q];
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            const uint4 rnd = philox16((uint32_t)(key0 + 32 * q + g * 16) >> 4, (uint32_t)qi, (uint32_t)(b * a.Hh + h),
                                                       seed_lo, seed_hi);
#pragma unroll
                            for (int i = 0; i < 16; i += 2) {
                                float p0 = ex2_approx(fmaf(__uint_as_float(r[q][g * 16 + i]), c, -mc));
                                float p1 = ex2_approx(fmaf(__uint_as_float(r[q][g * 16 + i + 1]), c, -mc));
                                lsum[q] += p0 + p1;
                                p0 = (philox_byte(rnd, i) < a.drop_thresh) ? 0.0f : p0 * a.inv_keep;
                                p1 = (philox_byte(rnd, i + 1) < a.drop_thresh) ? 0.0f : p1 * a.inv_keep;
                                const __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
                                pk[(g * 16 + i) >> 1] = *reinterpret_cast<const uint32_t*>(&pb);
                            }
                        }
                        store_chunk(q);
                    }
                }
                if (pp && (t == 0 || j + 1 < nturn)) bar_arrive_named(2 - t, 256);
                l_run += (lsum[0] + lsum[1]) + (lsum[2] + lsum[3]);
                MHA_TRACE(t, j, 6);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive_warp(&bars->p_full[t]);
                MHA_TRACE(t, j, 7);
                MHA_TRACE_WARP(t, j, 12);
            }
            // epilogue
            MHA_CTA_STAMP(3);
            mbar_wait(&bars->pv_full[t], (nbt - 1) & 1);
            tc_fence_after();
            const float inv = 1.0f / l_run;
            __nv_bfloat16* dst = a.out + (((size_t)b * a.Lq + min(qi, a.Lq - 1)) * a.Hh + h) * kD;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t o32[32];
                tmem_ld32_issue(tmem_o + 32 * hh, o32);
                tmem_ld_wait();
                if (qi < a.Lq) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint32_t w[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const __nv_bfloat162 v2 = __floats2bfloat162_rn(__uint_as_float(o32[i + 2 * u]) * inv,
                                                                             __uint_as_float(o32[i + 2 * u + 1]) * inv);
                            w[u] = *reinterpret_cast<const uint32_t*>(&v2);
                        }
                        *reinterpret_cast<uint4*>(dst + 32 * hh + i) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
            tc_fence_before();
            if (qi < a.Lq && a.lse != nullptr)
                a.lse[((size_t)b * a.Hh + h) * a.Lq + qi] = (m_used * c + log2f(l_run)) * 0.6931471805599453f;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
    MHA_CTA_STAMP(4);
}

// ---- forward, persistent: the default two-tile kernel with the work items looped inside the CTA --------------
// Per-CTA timeline of mha_fwd8_kernel (tools/mha_cta_timeline.py, L = 2048): 51 100 cycles per work item on an SM, of
// which 41 900 are the key-block loop; the rest is per-CTA cost - 910 cycles of barrier / tensor-memory set-up, ~3400
// until the first S is ready (tensor-map fetch, Q and K loads, the first product), ~3350 from the last block to the
// exit (last P V, normalise, store, dealloc) and ~1540 in which the SM waits for its next CTA: 18 % of the time at
// L = 2048, 10 % at L = 4096.  Here one CTA per SM walks the (batch, head, 256-query) items i = blockIdx.x, + gridDim.x,
// ...; the roles keep running across items:
//   * the K / V rings and every per-block barrier continue with CTA-wide block counters (stage = block % 4);
//   * Q is double-buffered: the producer fetches the next item's Q (and its first K / V blocks) while the current
//     item is still in its key-block loop (q_free: both issuers have finished the item's last Q K^T);
//   * an issuer starts the next item's S = Q K^T as soon as the last scores of the current item are in registers, i.e.
//     under the last block's exponentials and the epilogue; the first P V of an item overwrites O, so it waits for o_free
//     (the epilogue has read the accumulator);
//   * tensor memory is allocated once.
// Inner structure, TMEM layout and numerics are those of mha_fwd8_kernel<DROP> (outputs are bit-identical).
struct __align__(8) MhaBarriersP {
    uint64_t q_full[2];     // per Q buffer (items alternate)
    uint64_t q_free[2];
    uint64_t k_full[kFwd6Stages];
    uint64_t k_empty[kFwd6Stages];
    uint64_t v_full[kFwd6Stages];
    uint64_t v_empty[kFwd6Stages];
    uint64_t s_full[2];     // per tile
    uint64_t s_free[2];
    uint64_t p_full[2];
    uint64_t pv_full[2];
    uint64_t o_free[2];
    uint32_t tmem_base;
    uint32_t pad;
};
static_assert(sizeof(MhaBarriersP) <= 256, "barrier block");
constexpr int kFwdPSmem = (4 + 2 * kFwd6Stages) * kTileBytes + 256;     // Q x2 x2 + K ring + V ring + barriers
static_assert(kFwdPSmem <= 232448, "shared memory of the persistent forward kernel");

struct MhaItem {
    int b, h, q0, ntile, nb0, nb1, nblk;      // (no array: a runtime tile index would push the struct into local memory)
    __device__ __forceinline__ int nb(int t) const { return t == 0 ? nb0 : nb1; }
};
// The items of a CTA: item = blockIdx.x, + gridDim.x, ...; item -> (batch, head, 256-query tile), tiles fastest.  The walk
// keeps (batch * heads + head, tile) and advances them by additions (three divisions once per thread instead of per item).
struct MhaItemWalk {
    int bh, qt, dbh, dqt, nq2, left;
    int k, n_bh;        // causal walk: visit number, batch * heads
    bool causal;
    __device__ __forceinline__ MhaItemWalk(const MhaFwdArgs& a, int n_items, int nq2_) : nq2(nq2_) {
        const int first = blockIdx.x, step = gridDim.x;
        bh = first / nq2;
        qt = first - bh * nq2;
        dbh = step / nq2;
        dqt = step - dbh * nq2;
        left = first < n_items ? (n_items - 1 - first) / step + 1 : 0;
        k = 0;
        n_bh = a.B * a.Hh;
        causal = a.causal != 0 && a.dense_mask == nullptr;
    }
    __device__ __forceinline__ bool valid() const { return left > 0; }
    __device__ __forceinline__ void next() {
        --left;
        ++k;
        bh += dbh;
        qt += dqt;
        if (qt >= nq2) {
            qt -= nq2;
            ++bh;
        }
    }
    __device__ __forceinline__ MhaItem get(const MhaFwdArgs& a) const {
        MhaItem it;
        int ibh = bh, iq = qt;
        if (causal) {
            // Causal items cost 2 (qt + 1) key blocks: they are dealt longest first and in snake order (visit k: position
            // k G + c on even visits, k G + G - 1 - c on odd ones, in the list sorted by falling cost), which evens out the
            // CTAs' totals without a work counter.  The price is one division per item and K / V tiles that are shared by
            // CTAs of different visits instead of neighbours.
            const int G = gridDim.x, c = blockIdx.x;
            const int p = k * G + ((k & 1) ? G - 1 - c : c);
            const int total = n_bh * nq2;
            // the last visit is ragged: positions beyond the list belong to nobody; the plain walk counted `left` for
            // position k G + c, so map the tail straight
            const int pp = p < total ? p : k * G + c;
            iq = nq2 - 1 - pp / n_bh;
            ibh = pp - (pp / n_bh) * n_bh;
        }
        it.b = ibh / a.Hh;
        it.h = ibh - it.b * a.Hh;
        it.q0 = iq * 2 * kBM;
        const int kvlen = a.kv_len ? min(max(__ldg(a.kv_len + it.b), 0), a.Lk) : a.Lk;
        it.ntile = (it.q0 + kBM < a.Lq) ? 2 : 1;
        int k_end0 = a.causal ? min(kvlen, it.q0 + kBM) : kvlen;
        int k_end1 = a.causal ? min(kvlen, it.q0 + 2 * kBM) : kvlen;
        if (a.dense_mask) k_end0 = k_end1 = a.Lk;
        it.nb0 = max(1, (k_end0 + kBN - 1) / kBN);
        it.nb1 = (it.ntile == 2) ? max(1, (k_end1 + kBN - 1) / kBN) : 0;
        it.nblk = max(it.nb0, it.nb1);
        return it;
    }
};

template <bool DROP>
__global__ void __launch_bounds__(kFwd8Threads, 1)
mha_fwdp_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const MhaFwdArgs a, const int n_items, const int nq2) {
    extern __shared__ __align__(1024) unsigned char smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint32_t seed_lo = a.seed_lo, seed_hi = a.seed_hi;
    if (DROP) effective_seed(a.seed_dev, seed_lo, seed_hi);
    unsigned char* sQ = smem;                              // [2 buffers][2 tiles]
    unsigned char* sK = sQ + 4 * kTileBytes;               // kFwd6Stages tiles
    unsigned char* sV = sK + kFwd6Stages * kTileBytes;     // kFwd6Stages tiles
    MhaBarriersP* bars = reinterpret_cast<MhaBarriersP*>(sV + kFwd6Stages * kTileBytes);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->q_full[i], 1);
            mbar_init(&bars->q_free[i], 2);               // one arrival per tile issuer
        }
        for (int s = 0; s < kFwd6Stages; ++s) {
            mbar_init(&bars->k_full[s], 1);
            mbar_init(&bars->k_empty[s], 2);              // both issuers release every stage (a tile that does not need
            mbar_init(&bars->v_full[s], 1);               // the block arrives without a product)
            mbar_init(&bars->v_empty[s], 2);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&bars->s_full[t], 1);
            mbar_init(&bars->s_free[t], 4);               // one arrival per softmax warp of the tile
            mbar_init(&bars->p_full[t], 4);
            mbar_init(&bars->pv_full[t], 1);
            mbar_init(&bars->o_free[t], 4);
        }
        fence_mbar_init();
    }
    constexpr int kTmaWarp = 8, kMmaWarp = 9;
    if (warp == kMmaWarp) {
        tmem_alloc(&bars->tmem_base, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == kTmaWarp) {
        // ===== TMA producer =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        if (elect_one_sync()) {
            tma_prefetch_desc(&tm_q);
            tma_prefetch_desc(&tm_k);
            tma_prefetch_desc(&tm_v);
            int g = 0;                                    // blocks loaded by this CTA so far
            int k = 0;                                    // items
            for (MhaItemWalk walk(a, n_items, nq2); walk.valid(); walk.next(), ++k) {
                const MhaItem it = walk.get(a);
                const int qb = k & 1;
                if (k >= 2) mbar_wait(&bars->q_free[qb], ((k >> 1) - 1) & 1);     // the item two back is done with this Q buffer
                mbar_arrive_expect_tx(&bars->q_full[qb], it.ntile * kTileBytes);
                for (int t = 0; t < it.ntile; ++t)
                    tma_load_4d(sQ + (qb * 2 + t) * kTileBytes, &tm_q, 0, it.h, it.q0 + t * kBM, it.b, &bars->q_full[qb]);
                for (int j = 0; j < it.nblk; ++j, ++g) {
                    const int s = g % kFwd6Stages;
                    const int use = g / kFwd6Stages;
                    if (use > 0) mbar_wait(&bars->k_empty[s], (use - 1) & 1);
                    mbar_arrive_expect_tx(&bars->k_full[s], kTileBytes);
                    tma_load_4d(sK + s * kTileBytes, &tm_k, 0, it.h, j * kBN, it.b, &bars->k_full[s]);
                    if (use > 0) mbar_wait(&bars->v_empty[s], (use - 1) & 1);
                    mbar_arrive_expect_tx(&bars->v_full[s], kTileBytes);
                    tma_load_4d(sV + s * kTileBytes, &tm_v, 0, it.h, j * kBN, it.b, &bars->v_full[s]);
                }
            }
        }
    } else if (warp == kMmaWarp || warp == kMmaWarp + 1) {
        // ===== MMA issuers, one elected thread per tile =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        const int t = warp - kMmaWarp;
        if (elect_one_sync()) {
            constexpr uint32_t idesc_s = make_idesc(kBM, kBN, 0, 0);    // S = Q K^T : A, B K-major
            constexpr uint32_t idesc_pv = make_idesc(kBM, kD, 0, 1);    // PV = P V  : A from tensor memory, B MN-major
            const uint32_t tmem_s = tmem + 128 * t;
            const uint32_t tmem_o = tmem + 256 + 64 * t;
            const uint32_t tmem_p = tmem + 384 + 64 * t;
            int g = 0;        // K / V blocks seen by this CTA before the current item
            int n = 0;        // blocks this tile has issued before the current item (phases of s_*, p_*, pv_*)
            int act = 0;      // items in which this tile was active (phase of o_free)
            int k = 0;
            // S of block j of item `it` (its first block is the CTA's block g0 and this tile's block n0; Q buffer k0 & 1), or
            // just the release of the K stage when this tile does not need the block
            auto issue_s = [&](const MhaItem& it, int g0, int n0, int k0, int j) {
                const int nbt = it.nb(t);
                const int qb = k0 & 1;
                const int ks = (g0 + j) % kFwd6Stages;
                if (j == 0) mbar_wait(&bars->q_full[qb], (k0 >> 1) & 1);
                mbar_wait(&bars->k_full[ks], ((g0 + j) / kFwd6Stages) & 1);
                if (j >= nbt) {
                    mbar_arrive(&bars->k_empty[ks]);
                    if (nbt == 0 && j == 0) mbar_arrive(&bars->q_free[qb]);
                    return;
                }
                if (n0 + j > 0) mbar_wait(&bars->s_free[t], (n0 + j - 1) & 1);      // the previous scores are in registers
                tc_fence_after();
                const uint32_t q_addr = smem_u32(sQ + (qb * 2 + t) * kTileBytes);
                const uint32_t k_addr = smem_u32(sK + ks * kTileBytes);
#pragma unroll
                for (int kk = 0; kk < kD / 16; ++kk)
                    umma_bf16(tmem_s, smem_desc_sw128(q_addr + kk * 32, 16, 1024), smem_desc_sw128(k_addr + kk * 32, 16, 1024),
                              idesc_s, kk > 0 ? 1u : 0u);
                tc_commit(&bars->s_full[t]);
                tc_commit(&bars->k_empty[ks]);
                if (j + 1 == nbt) tc_commit(&bars->q_free[qb]);     // this tile's last Q K^T of the item: Q is free behind it
            };
            MhaItemWalk walk(a, n_items, nq2);
            MhaItem it = walk.get(a);
            if (walk.valid()) issue_s(it, 0, 0, 0, 0);
            while (walk.valid()) {
                walk.next();
                const bool has_next = walk.valid();
                MhaItem nx = it;
                if (has_next) nx = walk.get(a);
                const int nbt = it.nb(t);
                for (int j = 0; j < it.nblk; ++j) {
                    const int vs = (g + j) % kFwd6Stages;
                    // the next block's scores first: they only need s_free, which arrives long before p_full.  Behind an
                    // item's last block that is the FIRST block of the next item: its product runs under this item's last
                    // exponentials and epilogue.
                    if (j + 1 < it.nblk) issue_s(it, g, n, k, j + 1);
                    else if (has_next) {
                        MHA_TRACE_MMA(t, k + 1, 10);
                        issue_s(nx, g + it.nblk, n + nbt, k + 1, 0);
                        MHA_TRACE_MMA(t, k + 1, 11);
                    }
                    mbar_wait(&bars->v_full[vs], ((g + j) / kFwd6Stages) & 1);
                    if (j >= nbt) {
                        mbar_arrive(&bars->v_empty[vs]);
                        continue;
                    }
                    mbar_wait(&bars->p_full[t], (n + j) & 1);
                    if (j == 0 && act > 0) mbar_wait(&bars->o_free[t], (act - 1) & 1);   // the last epilogue has read O
                    tc_fence_after();
                    const uint32_t v_addr = smem_u32(sV + vs * kTileBytes);
#pragma unroll
                    for (int kk = 0; kk < kBN / 16; ++kk)
                        umma_bf16_ts(tmem_o, tmem_p + 8 * kk, smem_desc_sw128(v_addr + kk * 2048, kTileBytes, 1024), idesc_pv,
                                     (j > 0 || kk > 0) ? 1u : 0u);
                    tc_commit(&bars->pv_full[t]);
                    tc_commit(&bars->v_empty[vs]);
                    if (j == 0) MHA_TRACE_MMA(t, k, 12);
                    if (j + 1 == it.nblk) MHA_TRACE_MMA(t, k, 13);
                }
                g += it.nblk;
                n += nbt;
                if (nbt > 0) ++act;
                ++k;
                it = nx;
            }
        }
    } else if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");   // the idle warp of the utility warpgroup
    } else {
        // ===== softmax + epilogue: warps 0-3 own tile 0, warps 4-7 tile 1; thread = query row =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");    // 8 x 32 x 216 + 4 x 32 x 72 = 384 x 168, the CTA's pool
        const int t = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tmem_s = tmem + 128 * t + lane_base;
        const uint32_t tmem_o = tmem + 256 + 64 * t + lane_base;
        const uint32_t tmem_p = tmem + 384 + 64 * t + lane_base;
        const float c = a.scale_log2;
        int n = 0;            // blocks this tile has processed before the current item
        // The epilogue of an item is DEFERRED behind the first key block of the tile's next item: the wait for the item's
        // last P V (~850 cycles from the hand-over to the visible barrier) then costs nothing, and the normalise / store
        // runs while the other tile holds the XU pipe.  `ep_*` is the finished item's state.
        bool ep_pending = false;
        float ep_m = 0.0f, ep_l = 0.0f;
        int ep_qi = 0, ep_b = 0, ep_h = 0;
        auto epilogue = [&]() {
            // (its last P V completed long ago when this runs deferred: the next item's first block has waited for it
            // before storing its probabilities)
            mbar_wait(&bars->pv_full[t], (n - 1) & 1);
            tc_fence_after();
            uint32_t o32[2][32];
            tmem_ld32_issue(tmem_o, o32[0]);
            tmem_ld32_issue(tmem_o + 32, o32[1]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive_warp(&bars->o_free[t]);        // the next item's first P V may overwrite O
            const float inv = (DROP ? a.inv_keep : 1.0f) / ep_l;      // with dropout the kept probabilities went into P V unscaled
            if (ep_qi < a.Lq) {
                __nv_bfloat16* dst = a.out + (((size_t)ep_b * a.Lq + ep_qi) * a.Hh + ep_h) * kD;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint32_t w[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const __nv_bfloat162 v2 = __floats2bfloat162_rn(__uint_as_float(o32[hh][i + 2 * u]) * inv,
                                                                             __uint_as_float(o32[hh][i + 2 * u + 1]) * inv);
                            w[u] = *reinterpret_cast<const uint32_t*>(&v2);
                        }
                        *reinterpret_cast<uint4*>(dst + 32 * hh + i) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
                if (a.lse != nullptr)
                    a.lse[((size_t)ep_b * a.Hh + ep_h) * a.Lq + ep_qi] = (ep_m * c + log2f(ep_l)) * 0.6931471805599453f;
            }
            ep_pending = false;
        };
        // Token ring of the two tiles: tile 0 waits on named barrier 1, tile 1 on barrier 2; a tile holds the token while it
        // runs its exponentials and then hands it to the other tile, which keeps the tiles in opposite phases.  The ring runs
        // on across the items (tile 1 hands the token back after its last block too, so tile 0 can start the next item's
        // exponentials at once); tile 1 grants the very first turn.  Items with one tile do not touch it.
        bool ring_started = false;
        int kk_item = -1;
        for (MhaItemWalk walk(a, n_items, nq2); walk.valid(); walk.next()) {
            const MhaItem it = walk.get(a);
            ++kk_item;
            if (t >= it.ntile) continue;
            const int b = it.b, h = it.h;
            const int kvlen = a.kv_len ? min(max(__ldg(a.kv_len + b), 0), a.Lk) : a.Lk;
            const int qi = it.q0 + t * kBM + row;
            const uint8_t* mrow = a.dense_mask ? a.dense_mask + ((size_t)b * a.Lq + min(qi, a.Lq - 1)) * a.Lk : nullptr;
            float m_used = -INFINITY;
            float l_run = 0.0f;          // sum of the row's (undropped) probabilities, scaled like O
            const int nbt = it.nb(t);
            const bool pp = it.ntile == 2;
            const int nturn = pp ? it.nblk : nbt;
            MHA_TRACE(t, kk_item, 0);
            if (pp && t == 1 && !ring_started) bar_arrive_named(1, 256);
            if (pp) ring_started = true;
            for (int j = 0; j < nturn; ++j) {
                if (j >= nbt) {
                    bar_sync_named(1 + t, 256);
                    bar_arrive_named(2 - t, 256);
                    continue;
                }
                const int key0 = j * kBN;
                int lim = kvlen;
                if (a.causal) lim = min(lim, qi + 1);
                const bool need_mask = (key0 + kBN > lim) || (mrow != nullptr);
                if (j == 0) MHA_TRACE(t, kk_item, 8);
                mbar_wait(&bars->s_full[t], (n + j) & 1);
                tc_fence_after();
                if (j == 0) MHA_TRACE(t, kk_item, 9);
                uint32_t r[4][32];           // the row's 128 scores: read from TMEM exactly once
#pragma unroll
                for (int q = 0; q < 4; ++q) tmem_ld32_issue(tmem_s + 32 * q, r[q]);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive_warp(&bars->s_free[t]);      // the scores are in registers: S may be overwritten
                if (j == 0) MHA_TRACE(t, kk_item, 1);
                if (j == 1) MHA_TRACE(t, kk_item, 6);
                if (j == 2) MHA_TRACE(t, kk_item, 7);
                float m_q[4];
                if (__builtin_expect(need_mask, 0)) {
                    // dead keys of this row as four 32-bit masks: the length / causal limit by arithmetic, a dense mask by a
                    // ROLLED byte loop (unrolled it is 128 x (address, LDG.U8, compare, select): a quarter of the kernel's
                    // code, which the persistent loop pays for in instruction-cache misses)
                    uint32_t dead[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int live = lim - (key0 + 32 * q);                 // keys [0, live) of this chunk are inside the limit
                        dead[q] = live >= 32 ? 0u : live <= 0 ? 0xffffffffu : (0xffffffffu << live);
                    }
                    if (mrow != nullptr) {
                        const int kmax = min(kBN, a.Lk - key0);
#pragma unroll 1
                        for (int kk = 0; kk < kmax; ++kk) {
                            const uint32_t bit = (mrow[key0 + kk] != 0) ? (1u << (kk & 31)) : 0u;
                            const int q = kk >> 5;
                            dead[0] |= (q == 0) ? bit : 0u;
                            dead[1] |= (q == 1) ? bit : 0u;
                            dead[2] |= (q == 2) ? bit : 0u;
                            dead[3] |= (q == 3) ? bit : 0u;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if ((dead[q] >> i) & 1u) r[q][i] = 0xff800000u;   // -inf
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    m_q[q] = -INFINITY;          // four independent chains
#pragma unroll
                    for (int i = 0; i < 32; ++i) m_q[q] = fmaxf(m_q[q], __uint_as_float(r[q][i]));
                }
                const float m_blk = fmaxf(fmaxf(m_q[0], m_q[1]), fmaxf(m_q[2], m_q[3]));
                const float m_new = fmaxf(m_used, m_blk);
                const bool grow = (j > 0) && ((m_new - m_used) * c > 8.0f);
                bool pv_waited = false;
                if (j == 0) {
                    m_used = m_new;
                } else if (__builtin_expect(__any_sync(0xffffffffu, grow), 0)) {      // TMEM accesses are warp-collective: all lanes go
                    // O is being accumulated by P V of this tile's previous block: wait for it before touching O
                    mbar_wait(&bars->pv_full[t], (n + j - 1) & 1);
                    tc_fence_after();
                    pv_waited = true;
                    const float f = grow ? ex2_approx((m_used - m_new) * c) : 1.0f;   // m_used = -inf -> 0
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint32_t o32[32];
                        tmem_ld32_issue(tmem_o + 32 * hh, o32);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o32[i] = __float_as_uint(__uint_as_float(o32[i]) * f);
                        tmem_st32(tmem_o + 32 * hh, o32);
                    }
                    l_run *= f;
                    tmem_st_wait();
                    if (grow) m_used = m_new;
                }
                const float mc = ((m_used == -INFINITY) ? 0.0f : m_used) * c;   // fully masked so far: keep exp2 finite
                // The P columns are free once P V of the previous block - of the previous ITEM at an item's first block - has
                // read them.
                if (n + j > 0 && !pv_waited) mbar_wait(&bars->pv_full[t], (n + j - 1) & 1);
                float lsum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                uint32_t pk_all[4][16];
                if (!DROP) {
                    // x = s * scale - max in place, packed; no XU work yet
                    const uint64_t c2 = pack_f32x2(c, c);
                    const uint64_t mc2 = pack_f32x2(-mc, -mc);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            float x0, x1;
                            unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[q][i]), __uint_as_float(r[q][i + 1])), c2, mc2), x0, x1);
                            r[q][i] = __float_as_uint(x0);
                            r[q][i + 1] = __float_as_uint(x1);
                        }
                    }
                    if (pp) {
                        bar_sync_named(1 + t, 256);          // this tile's turn on the XU pipe
                        if (j == 0) MHA_TRACE(t, kk_item, 2);
                        // the exponentials stay behind the barrier (they are pure: ptxas would hoist most of them above it, into
                        // the other tile's turn - two warps of a sub-partition feeding the XU pipe at once are slower than one
                        // after the other: 2980 vs 2630 cycles per key block)
#pragma unroll
                        for (int q = 0; q < 4; ++q) reg_tie32(r[q]);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint64_t ls2 = pack_f32x2(0.0f, 0.0f);
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const float p0 = ex2_approx(__uint_as_float(r[q][i]));
                            const float p1 = ex2_approx(__uint_as_float(r[q][i + 1]));
                            ls2 = add_f32x2(ls2, pack_f32x2(p0, p1));
                            pk_all[q][i >> 1] = cvt_bf16x2(p0, p1);
                        }
                        float l0, l1;
                        unpack_f32x2(ls2, l0, l1);
                        lsum[q] = l0 + l1;
                        tmem_st16(tmem_p + 16 * q, pk_all[q]);
                    }
                } else {
                    // with dropout: the same packed pre-scaling pass, then exponentials, the row sum of the UNDROPPED
                    // probabilities and the keep masks (keep_mask2: one byte permute + one packed fp16 compare per pair, ANDed
                    // into the packed probabilities; the 1 / keep scale is applied in the epilogue)
                    const uint64_t c2 = pack_f32x2(c, c);
                    const uint64_t mc2 = pack_f32x2(-mc, -mc);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            float x0, x1;
                            unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[q][i]), __uint_as_float(r[q][i + 1])), c2, mc2), x0, x1);
                            r[q][i] = __float_as_uint(x0);
                            r[q][i + 1] = __float_as_uint(x1);
                        }
                    }
                    const uint32_t thr_h2 = (0x3c00u | a.drop_thresh) * 0x00010001u;
                    if (pp) bar_sync_named(1 + t, 256);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t (&pk)[16] = pk_all[q];
                        uint64_t ls2 = pack_f32x2(0.0f, 0.0f);
#pragma unroll
                        for (int g2 = 0; g2 < 2; ++g2) {
                            const uint4 rnd = philox16((uint32_t)(key0 + 32 * q + g2 * 16) >> 4, (uint32_t)qi, (uint32_t)(b * a.Hh + h),
                                                       seed_lo, seed_hi);
                            const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
                            for (int i = 0; i < 16; i += 2) {
                                const float p0 = ex2_approx(__uint_as_float(r[q][g2 * 16 + i]));
                                const float p1 = ex2_approx(__uint_as_float(r[q][g2 * 16 + i + 1]));
                                ls2 = add_f32x2(ls2, pack_f32x2(p0, p1));
                                pk[(g2 * 16 + i) >> 1] = cvt_bf16x2(p0, p1) & keep_mask2(rw[i >> 2], (i >> 1) & 1, thr_h2);
                            }
                        }
                        float l0, l1;
                        unpack_f32x2(ls2, l0, l1);
                        lsum[q] = l0 + l1;
                        tmem_st16(tmem_p + 16 * q, pk);
                    }
                }
                if (pp) bar_arrive_named(2 - t, 256);
                l_run += (lsum[0] + lsum[1]) + (lsum[2] + lsum[3]);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive_warp(&bars->p_full[t]);
                if (j == 0) MHA_TRACE(t, kk_item, 3);
                if (j == 0 && ep_pending) epilogue();        // of the previous item (n still counts its blocks only)
                if (j == 0) MHA_TRACE(t, kk_item, 4);
                if (j + 1 == nturn) MHA_TRACE(t, kk_item, 5);
            }
            n += nbt;
            ep_pending = true;
            ep_m = m_used;
            ep_l = l_run;
            ep_qi = qi;
            ep_b = b;
            ep_h = h;
        }
        if (ep_pending) epilogue();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// =====================================================================================
// Backward.  One CTA per (b, head, 128-key tile j); loop over the 128-query tiles i.
//   S  = Q_i K_j^T            dP = dO_i V_j^T                (both into TMEM)
//   P  = exp2(S*c - lse_i)    dS = P o (dP - delta_i) * scale (softmax warps, masks as in forward)
//   dV_j += P^T dO_i          dK_j += dS^T Q_i               (accumulate in TMEM over i)
//   dQ_i  = dS K_j            -> fp32 red.add into the dQ workspace (summed over the key tiles)
// P and dS are written once to shared memory as [2 key halves][128 q][64 keys] bf16 with the
// 128-byte swizzle; the same bytes serve as the K-major A operand of dS K and, read
// MN-major, as the transposed A operand of P^T dO and dS^T Q.  Q_i, dO_i, K_j, V_j are
// consumed straight from their TMA tiles, K-major or MN-major as each product needs.
// TMEM columns: S 0-127 | dP 128-255 | dV 256-319 | dK 320-383 | dQ 384-447  (512 allocated).
// =====================================================================================
struct MhaBwdArgs {
    const int* kv_len;
    const uint8_t* dense_mask;
    int causal;
    int B, Hh, Lq, Lk;
    float scale;
    float scale_log2;
    const float* lse;     // [B,Hh,Lq]
    const float* delta;   // [B,Hh,Lq]  rowsum(dO o O)
    float* dq_acc;        // [B,Lq,Hh,64] fp32, zeroed
    __nv_bfloat16* g_k;   // [B,Lk,Hh,64]
    __nv_bfloat16* g_v;
    uint32_t drop_thresh; // dropout on the probabilities, same meaning as in MhaFwdArgs
    float inv_keep;
    uint32_t seed_lo, seed_hi;
    const uint64_t* seed_dev;
};

struct __align__(8) MhaBwdBarriers {
    uint64_t kv_full;
    uint64_t qdo_full[2];
    uint64_t qdo_empty[2];
    uint64_t sdp_full;
    uint64_t pds_full;
    uint64_t dq_full;
    uint64_t dq_free;      // DQW: the dQ warps have read the dQ accumulator of a tile
    uint32_t tmem_base;
    uint32_t pad;
    uint32_t seed[2];      // effective dropout seed (kept in shared memory: the 16-warp instance has no register to spare)
};
static_assert(sizeof(MhaBwdBarriers) <= 128, "barrier block of the backward kernel");

constexpr int kBwdSmem = 2 * kTileBytes /*K,V*/ + 4 * kTileBytes /*Q,dO x2*/ + 2 * kTileBytes /*P*/ + 2 * kTileBytes /*dS*/ +
                         2 * kTileBytes /*dQ staging: two [128 x 32] fp32 boxes*/ + 128;

// Bulk reduce-add of a shared-memory box into a global fp32 tensor (out-of-range rows are clipped).
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(smem_src))
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// CG = column groups: the softmax-backward work of a 128 x 128 tile is spread over 4*CG warps
// (warp w: TMEM lane quarter w % 4 = query rows, column group w / 4).  One warp per SM
// sub-partition (CG = 1) leaves every dependent instruction's latency exposed; 4 per
// sub-partition hide it.
// DQW: four more warps (one per TMEM lane quarter) take dQ out of the CTA: they wait for a tile's dQ = dS K, read it from
// tensor memory, stage it in shared memory and issue the bulk reduce-add, while the softmax-backward warps are already on the
// next tile.  On those warps' own timeline (clock64: ~4900 cycles per 128 x 128 tile = softmax 1600-2000 + P / dS store +
// hand-over ~700 + dQ load and flush 900-1500 + loop top ~800) the flush was a quarter; the tensor pipe needs 1280 cycles
// per tile.  24 warps (16 softmax, TMA, MMA, two idle, four dQ) at 80 registers, reallocated with setmaxnreg to 96 / 40 / 56.
template <int CG, bool DROP, bool DQW = false>
__global__ void __launch_bounds__(DQW ? 768 : 128 * CG + 64, 1)
mha_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
               const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
               const __grid_constant__ CUtensorMap tm_dqacc, const MhaBwdArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    unsigned char* sK = smem;
    unsigned char* sV = sK + kTileBytes;
    unsigned char* sQ = sV + kTileBytes;            // 2 stages
    unsigned char* sDO = sQ + 2 * kTileBytes;       // 2 stages
    unsigned char* sP = sDO + 2 * kTileBytes;       // [2][128][128B]
    unsigned char* sDS = sP + 2 * kTileBytes;       // [2][128][128B]
    unsigned char* sDQ = sDS + 2 * kTileBytes;      // [2][128][128B] fp32 staging of one dQ tile (columns 0-31 | 32-63)
    MhaBwdBarriers* bars = reinterpret_cast<MhaBwdBarriers*>(sDQ + 2 * kTileBytes);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int key0 = blockIdx.x * kBN;
    const int h = blockIdx.y;
    const int b = blockIdx.z;
    const int kvlen = a.kv_len ? min(max(__ldg(a.kv_len + b), 0), a.Lk) : a.Lk;
    const int nq = (a.Lq + kBM - 1) / kBM;
    int i_start = a.causal ? (key0 / kBM) : 0;
    // a key tile that is entirely padding (and no dense mask decides otherwise) gets zero gradients
    const bool dead_tile = (a.dense_mask == nullptr) && (key0 >= kvlen);
    if (dead_tile) i_start = nq;
    const int nsteps = max(0, nq - i_start);

    if (threadIdx.x == 0) {
        mbar_init(&bars->kv_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->qdo_full[s], 1);
            mbar_init(&bars->qdo_empty[s], 1);
        }
        mbar_init(&bars->sdp_full, 1);
        mbar_init(&bars->pds_full, 4 * CG);   // one arrival per softmax warp
        mbar_init(&bars->dq_full, 1);
        mbar_init(&bars->dq_free, 4);
        fence_mbar_init();
        if (DROP) {
            uint32_t lo = a.seed_lo, hi = a.seed_hi;
            effective_seed(a.seed_dev, lo, hi);
            bars->seed[0] = lo;
            bars->seed[1] = hi;
        }
    }
    constexpr int kTmaWarp = 4 * CG, kMmaWarp = 4 * CG + 1;
    if (warp == kMmaWarp) {
        tmem_alloc(&bars->tmem_base, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    const uint32_t tm_s = tmem, tm_dp = tmem + 128, tm_dv = tmem + 256, tm_dk = tmem + 320, tm_dq = tmem + 384;

    static_assert(!DQW || CG == 4, "the dQ warps come with the 16-warp instance");
    if (DQW && warp >= 4 * CG && warp < 4 * CG + 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");   // TMA, MMA, two idle warps
    if (DQW && warp >= 4 * CG + 4) {
        // ===== dQ warps: thread = query row of the tile =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
        const bool lead = (warp == 4 * CG + 4) && lane == 0;
        for (int it = 0; it < nsteps; ++it) {
            mbar_wait(&bars->dq_full, it & 1);
            tc_fence_after();
            if (lead) bulk_wait_read0();                      // the previous tile's reduce has read the staging
            bar_sync_named(1, 128);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {                  // columns 0-31 | 32-63: one [128 x 32] fp32 box each
                float dq[32];
                tmem_ld32(tm_dq + lane_base + 32 * hf, dq);
                unsigned char* dst = sDQ + hf * kTileBytes + row * 128;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const int chunk = (i >> 2) ^ (row & 7);
                    *reinterpret_cast<float4*>(dst + chunk * 16) = make_float4(dq[i], dq[i + 1], dq[i + 2], dq[i + 3]);
                }
            }
            tc_fence_before();
            mbar_arrive_warp(&bars->dq_free);                 // the accumulator may take the next tile's dS K
            fence_proxy_async();
            bar_sync_named(2, 128);
            if (lead) {
                tma_reduce_add_4d(&tm_dqacc, sDQ, 0, h, (i_start + it) * kBM, b);
                tma_reduce_add_4d(&tm_dqacc, sDQ + kTileBytes, 32, h, (i_start + it) * kBM, b);
                bulk_commit();
            }
        }
        if (lead && nsteps > 0) bulk_wait0();                 // the reduces must have landed before the kernel ends
    } else
    if (warp == kTmaWarp) {
        // ===== TMA producer =====
        if (nsteps > 0 && elect_one_sync()) {
            tma_prefetch_desc(&tm_q);
            tma_prefetch_desc(&tm_do);
            mbar_arrive_expect_tx(&bars->kv_full, 2 * kTileBytes);
            tma_load_4d(sK, &tm_k, 0, h, key0, b, &bars->kv_full);
            tma_load_4d(sV, &tm_v, 0, h, key0, b, &bars->kv_full);
            for (int it = 0; it < nsteps; ++it) {
                const int s = it & 1;
                if (it >= 2) mbar_wait(&bars->qdo_empty[s], ((it >> 1) - 1) & 1);
                mbar_arrive_expect_tx(&bars->qdo_full[s], 2 * kTileBytes);
                tma_load_4d(sQ + s * kTileBytes, &tm_q, 0, h, (i_start + it) * kBM, b, &bars->qdo_full[s]);
                tma_load_4d(sDO + s * kTileBytes, &tm_do, 0, h, (i_start + it) * kBM, b, &bars->qdo_full[s]);
            }
        }
    } else if (warp == kMmaWarp) {
        // ===== MMA issuer =====
        // Order on the tensor pipe: S/dP of tile it+1 go BEFORE dV/dK/dQ of tile it, so the softmax
        // warps can start on tile it+1 while the three accumulating products of tile it run.
        if (nsteps > 0 && elect_one_sync()) {
            constexpr uint32_t id_s = make_idesc(kBM, kBN, 0, 0);     // S  = Q K^T   / dP = dO V^T
            constexpr uint32_t id_t = make_idesc(kBN, kD, 1, 1);      // dV = P^T dO  / dK = dS^T Q  (A and B MN-major)
            constexpr uint32_t id_q = make_idesc(kBM, kD, 0, 1);      // dQ = dS K    (A K-major, B MN-major)
            const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
            const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sDS);
            auto issue_sdp = [&](int it) {
                const int s = it & 1;
                const uint32_t q_addr = smem_u32(sQ + s * kTileBytes);
                const uint32_t do_addr = smem_u32(sDO + s * kTileBytes);
                mbar_wait(&bars->qdo_full[s], (it >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < kD / 16; ++kk)
                    umma_bf16(tm_s, smem_desc_sw128(q_addr + kk * 32, 16, 1024), smem_desc_sw128(k_addr + kk * 32, 16, 1024),
                              id_s, kk > 0 ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < kD / 16; ++kk)
                    umma_bf16(tm_dp, smem_desc_sw128(do_addr + kk * 32, 16, 1024), smem_desc_sw128(v_addr + kk * 32, 16, 1024),
                              id_s, kk > 0 ? 1u : 0u);
                tc_commit(&bars->sdp_full);
            };
            mbar_wait(&bars->kv_full, 0);
            issue_sdp(0);
            for (int it = 0; it < nsteps; ++it) {
                const int s = it & 1;
                const uint32_t q_addr = smem_u32(sQ + s * kTileBytes);
                const uint32_t do_addr = smem_u32(sDO + s * kTileBytes);
                // P/dS of tile it are in shared memory; S/dP (TMEM) and dQ of tile it-1 (TMEM) have been read
                mbar_wait(&bars->pds_full, it & 1);
                tc_fence_after();
                if (it + 1 < nsteps) issue_sdp(it + 1);
                // contraction over the 128 queries of the tile: 8 steps of 16 rows (2048 bytes)
#pragma unroll
                for (int kk = 0; kk < kBM / 16; ++kk) {
                    umma_bf16(tm_dv, smem_desc_sw128(p_addr + kk * 2048, kTileBytes, 1024),
                              smem_desc_sw128(do_addr + kk * 2048, kTileBytes, 1024), id_t, (it > 0 || kk > 0) ? 1u : 0u);
                }
#pragma unroll
                for (int kk = 0; kk < kBM / 16; ++kk) {
                    umma_bf16(tm_dk, smem_desc_sw128(ds_addr + kk * 2048, kTileBytes, 1024),
                              smem_desc_sw128(q_addr + kk * 2048, kTileBytes, 1024), id_t, (it > 0 || kk > 0) ? 1u : 0u);
                }
                // contraction over the 128 keys: dS K-major (two 64-key halves), K tile MN-major
                if (DQW && it > 0) {
                    mbar_wait(&bars->dq_free, (it - 1) & 1);      // the dQ warps have read the previous tile's dQ
                    tc_fence_after();
                }
#pragma unroll
                for (int kk = 0; kk < kBN / 16; ++kk) {
                    umma_bf16(tm_dq, smem_desc_sw128(ds_addr + (kk >> 2) * kTileBytes + (kk & 3) * 32, 16, 1024),
                              smem_desc_sw128(k_addr + kk * 2048, kTileBytes, 1024), id_q, kk > 0 ? 1u : 0u);
                }
                tc_commit(&bars->dq_full);
                tc_commit(&bars->qdo_empty[s]);
            }
        }
    } else if (warp >= 4 * CG) {
        // (DQW: the two idle warps of the utility warpgroup)
    } else {
        // ===== softmax-backward warps: thread = (query row of the tile, column group); key row in the epilogue =====
        if (DQW) asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        const int row = (warp & 3) * 32 + lane;
        const int cg = warp >> 2;
        constexpr int kColsS = kBN / CG;    // S / dP columns of this thread
        constexpr int kColsD = kD / CG;     // dQ / dK / dV columns of this thread
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const float c = a.scale_log2;
        // dQ of a tile leaves through shared memory and ONE bulk reduce-add per 32-column box (TMA,
        // asynchronous, rows beyond Lq clipped by the tensor map) instead of per-thread RED.ADDs
        const bool dq_issuer = (warp == 0 && lane == 0);
        auto flush_dq = [&](const float (&dq)[kColsD], int q_tile) {
            if (dq_issuer) bulk_wait_read0();                 // the previous tile's reduce has read the staging
            bar_sync_named(1, 128 * CG);
            const int col0 = cg * kColsD;                     // first of this thread's dQ columns
            unsigned char* dst = sDQ + (col0 >> 5) * kTileBytes + row * 128;
#pragma unroll
            for (int i = 0; i < kColsD; i += 4) {
                const int chunk = (((col0 & 31) + i) >> 2) ^ (row & 7);
                *reinterpret_cast<float4*>(dst + chunk * 16) = make_float4(dq[i], dq[i + 1], dq[i + 2], dq[i + 3]);
            }
            fence_proxy_async();
            bar_sync_named(2, 128 * CG);
            if (dq_issuer) {
                tma_reduce_add_4d(&tm_dqacc, sDQ, 0, h, q_tile * kBM, b);
                tma_reduce_add_4d(&tm_dqacc, sDQ + kTileBytes, 32, h, q_tile * kBM, b);
                bulk_commit();
            }
        };
        auto load_dq = [&](float (&dq)[kColsD]) {
            if constexpr (kColsD == 32) {
                tmem_ld32(tm_dq + lane_base + cg * kColsD, dq);
            } else {
                tmem_ld16(tm_dq + lane_base + cg * kColsD, dq);
            }
        };
        // row statistics of the next tile are fetched one tile ahead
        auto load_stats = [&](int it, float& lse2_o, float& dlt_o) {
            const int qi = (i_start + it) * kBM + row;
            const bool ok = it < nsteps && qi < a.Lq;
            const size_t stat = ((size_t)b * a.Hh + h) * a.Lq + (ok ? qi : 0);
            lse2_o = ok ? __ldg(a.lse + stat) * 1.4426950408889634f : INFINITY;
            dlt_o = ok ? __ldg(a.delta + stat) : 0.0f;
        };
        float lse2_n, dlt_n;
        load_stats(0, lse2_n, dlt_n);
        for (int it = 0; it < nsteps; ++it) {
            const int qi = (i_start + it) * kBM + row;
            const float lse2 = lse2_n, dlt = dlt_n;
            load_stats(it + 1, lse2_n, dlt_n);
            int lim = kvlen;
            if (a.causal) lim = min(lim, qi + 1);
            const uint8_t* mrow = a.dense_mask ? a.dense_mask + ((size_t)b * a.Lq + min(qi, a.Lq - 1)) * a.Lk : nullptr;
            const bool need_mask = (key0 + kBN > lim) || (mrow != nullptr);
            // dead keys among this thread's kColsS columns as a bit mask: the length / causal limit by arithmetic, a dense
            // mask by a rolled byte loop (unrolled into the packed loop below it was two guarded byte loads per pair)
            using dead_t = typename std::conditional<(kColsS <= 32), uint32_t, uint64_t>::type;
            dead_t dead = 0;
            if (__builtin_expect(need_mask, 0)) {
                const int first = key0 + cg * kColsS;
                const int live = lim - first;                     // columns [0, live) are inside the limit
                dead = live >= kColsS ? (dead_t)0 : live <= 0 ? (dead_t)~(dead_t)0 : (dead_t)((dead_t)~(dead_t)0 << live);
                if (mrow != nullptr) {
                    const int kmax = min(kColsS, a.Lk - first);
#pragma unroll 1
                    for (int kk = 0; kk < kmax; ++kk) dead |= (dead_t)(mrow[first + kk] != 0) << kk;
                }
            }

            // P and dS of this thread's columns, packed bf16, kept in registers until the products of
            // the previous tile have released the shared-memory tiles
            uint32_t pk[kColsS / 2], dk[kColsS / 2];
            mbar_wait(&bars->sdp_full, it & 1);
            tc_fence_after();
            // 16 columns at a time: S and dP of a 32-column pass would hold 64 registers next to pk / dk / dq, and
            // the 16-warp instance (96 registers per thread) spilled
#pragma unroll
            for (int c0 = 0; c0 < kColsS; c0 += 16) {
                const int cc = cg * kColsS + c0;
                float sv[16], dp[16];
                tmem_ld16(tm_s + lane_base + cc, sv);
                tmem_ld16(tm_dp + lane_base + cc, dp);
                uint4 rnd[1];
                if (DROP) rnd[0] = philox16((uint32_t)(key0 + cc) >> 4, (uint32_t)qi, (uint32_t)(b * a.Hh + h), bars->seed[0], bars->seed[1]);
                {
                    // packed f32x2 arithmetic: P = 2^(S c - lse), dS = P (dP' scale - delta scale); with dropout
                    // (mask M, keep probability k) dP' = dP o M / k and the P that dV sees is P o M / k
                    const uint64_t c2 = pack_f32x2(c, c), nl2 = pack_f32x2(-lse2, -lse2);
                    const uint64_t sc2 = pack_f32x2(a.scale, a.scale), nd2 = pack_f32x2(-dlt * a.scale, -dlt * a.scale);
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        float x0, x1;
                        unpack_f32x2(fma_f32x2(pack_f32x2(sv[i], sv[i + 1]), c2, nl2), x0, x1);
                        float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
                        if ((dead >> (c0 + i)) & 1u) p0 = 0.0f;
                        if ((dead >> (c0 + i + 1)) & 1u) p1 = 0.0f;
                        const uint64_t p2 = pack_f32x2(p0, p1);
                        uint64_t g2 = pack_f32x2(dp[i], dp[i + 1]);      // d(loss)/d(dropped, rescaled probability)
                        uint64_t pd2 = p2;                               // the probabilities P V was computed with
                        if (DROP) {
                            const float k0 = philox_byte(rnd[0], i) >= a.drop_thresh ? a.inv_keep : 0.0f;
                            const float k1 = philox_byte(rnd[0], i + 1) >= a.drop_thresh ? a.inv_keep : 0.0f;
                            const uint64_t kf2 = pack_f32x2(k0, k1);
                            g2 = mul_f32x2(g2, kf2);
                            pd2 = mul_f32x2(p2, kf2);
                        }
                        float s0, s1, q0, q1;
                        unpack_f32x2(mul_f32x2(p2, fma_f32x2(g2, sc2, nd2)), s0, s1);
                        unpack_f32x2(pd2, q0, q1);
                        pk[(c0 + i) >> 1] = cvt_bf16x2(q0, q1);
                        dk[(c0 + i) >> 1] = cvt_bf16x2(s0, s1);
                    }
                }
            }
            float dq[kColsD];
            if (it > 0) {
                // dV/dK/dQ of the previous tile are done: its P/dS tiles are free and its dQ is in TMEM
                mbar_wait(&bars->dq_full, (it - 1) & 1);
                tc_fence_after();
                if (!DQW) load_dq(dq);
            }
#pragma unroll
            for (int c0 = 0; c0 < kColsS; c0 += 32) {
                const int cc = cg * kColsS + c0;
                const int off = (cc >> 6) * kTileBytes + row * 128;
                const int chunk0 = (cc & 63) >> 3;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int chunk = (chunk0 + q4) ^ (row & 7);
                    const int w0 = (c0 >> 1) + 4 * q4;
                    *reinterpret_cast<uint4*>(sP + off + chunk * 16) = make_uint4(pk[w0], pk[w0 + 1], pk[w0 + 2], pk[w0 + 3]);
                    *reinterpret_cast<uint4*>(sDS + off + chunk * 16) = make_uint4(dk[w0], dk[w0 + 1], dk[w0 + 2], dk[w0 + 3]);
                }
            }
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive_warp(&bars->pds_full);
            if (!DQW && it > 0) flush_dq(dq, i_start + it - 1);
        }
        if (nsteps > 0) {
            mbar_wait(&bars->dq_full, (nsteps - 1) & 1);      // every product of the key tile is done: dK, dV are final
            tc_fence_after();
            if (!DQW) {
                float dq[kColsD];
                load_dq(dq);
                flush_dq(dq, i_start + nsteps - 1);
                if (dq_issuer) bulk_wait0();    // the reduce must have landed before the kernel ends
            }
        }
        // epilogue: dK_j, dV_j (thread = key row, kColsD columns).  The TMEM loads are warp-collective,
        // so every thread issues them; only rows inside the sequence store.
        const int key = key0 + row;
        const bool key_ok = key < a.Lk;
        const size_t kv_off = (((size_t)b * a.Lk + (key_ok ? key : 0)) * a.Hh + h) * kD + cg * kColsD;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            const uint32_t src = (part == 0 ? tm_dk : tm_dv) + lane_base + cg * kColsD;
            __nv_bfloat16* dst = (part == 0 ? a.g_k : a.g_v) + kv_off;
#pragma unroll
            for (int cc = 0; cc < kColsD; cc += 16) {
                float acc[16];
                if (nsteps > 0) {   // CTA-uniform
                    tmem_ld16(src + cc, acc);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
                }
                if (key_ok) {
#pragma unroll
                    for (int i = 0; i < 16; i += 8) {
                        uint32_t w[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const __nv_bfloat162 v2 = __floats2bfloat162_rn(acc[i + 2 * u], acc[i + 2 * u + 1]);
                            w[u] = *reinterpret_cast<const uint32_t*>(&v2);
                        }
                        *reinterpret_cast<uint4*>(dst + cc + i) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ---- backward, persistent: the 16 + 4-warp kernel above with the (batch, head, key tile) items looped inside one CTA per SM
// Per CTA the one-item kernel pays ~9000 cycles around its query-tile loop (barrier / tensor-memory set-up, the first K, V, Q,
// dO loads and products, the dK / dV epilogue, the SM's turnaround to the next CTA): 11 % at L = 2048, most of the call at the
// model's shapes (two query tiles per key tile).  Here the roles keep running across items: the Q / dO ring and the per-tile
// barriers continue on CTA-wide tile counters; K / V of the next item are fetched as soon as the last products of the current
// item have read them (kv_free), i.e. under the epilogue; the next item's first S / dP products go out as soon as K / V have
// landed, and its first dV / dK products - which overwrite the accumulators - wait for acc_free (the epilogue has read dK, dV).
// Causal items cost (number of query tiles - key tile index) tiles: they are dealt longest first in snake order.
struct __align__(8) MhaBwdBarriersP {
    uint64_t kv_full, kv_free;
    uint64_t qdo_full[2];
    uint64_t qdo_empty[2];
    uint64_t sdp_full;
    uint64_t pds_full;
    uint64_t dq_full;
    uint64_t dq_free;
    uint64_t acc_free;
    uint32_t tmem_base;
    uint32_t pad;
    uint32_t seed[2];
};
static_assert(sizeof(MhaBwdBarriersP) <= 128, "barrier block of the persistent backward kernel");

struct MhaBwdItem {
    int b, h, key0, i_start, nsteps;
};
struct MhaBwdWalk {
    int k, left, n_bh, nkt;
    bool causal;
    __device__ __forceinline__ MhaBwdWalk(const MhaBwdArgs& a, int n_items, int nkt_) : nkt(nkt_) {
        const int first = blockIdx.x, step = gridDim.x;
        left = first < n_items ? (n_items - 1 - first) / step + 1 : 0;
        k = 0;
        n_bh = a.B * a.Hh;
        causal = a.causal != 0 && a.dense_mask == nullptr;
    }
    __device__ __forceinline__ bool valid() const { return left > 0; }
    __device__ __forceinline__ void next() { --left; ++k; }
    __device__ __forceinline__ MhaBwdItem get(const MhaBwdArgs& a) const {
        const int G = gridDim.x, c = blockIdx.x;
        int jt, bh;
        if (causal) {
            // longest first (key tile 0 sees every query tile), snake order over the CTAs; a ragged last visit maps straight
            const int total = n_bh * nkt;
            const int p = k * G + ((k & 1) ? G - 1 - c : c);
            const int pp = p < total ? p : k * G + c;
            jt = pp / n_bh;
            bh = pp - jt * n_bh;
        } else {
            const int item = k * G + c;            // key tiles of one (batch, head) next to each other: Q / dO stay in L2
            bh = item / nkt;
            jt = item - bh * nkt;
        }
        MhaBwdItem it;
        it.b = bh / a.Hh;
        it.h = bh - it.b * a.Hh;
        it.key0 = jt * kBN;
        const int kvlen = a.kv_len ? min(max(__ldg(a.kv_len + it.b), 0), a.Lk) : a.Lk;
        const int nq = (a.Lq + kBM - 1) / kBM;
        it.i_start = a.causal ? (it.key0 / kBM) : 0;
        if (a.dense_mask == nullptr && it.key0 >= kvlen) it.i_start = nq;      // a key tile that is entirely padding
        it.nsteps = max(0, nq - it.i_start);
        return it;
    }
};

template <bool DROP>
__global__ void __launch_bounds__(768, 1)
mha_bwdp_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_dqacc, const MhaBwdArgs a, const int n_items, const int nkt) {
    constexpr int CG = 4;
    extern __shared__ __align__(1024) unsigned char smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    unsigned char* sK = smem;
    unsigned char* sV = sK + kTileBytes;
    unsigned char* sQ = sV + kTileBytes;            // 2 stages
    unsigned char* sDO = sQ + 2 * kTileBytes;       // 2 stages
    unsigned char* sP = sDO + 2 * kTileBytes;       // [2][128][128B]
    unsigned char* sDS = sP + 2 * kTileBytes;       // [2][128][128B]
    unsigned char* sDQ = sDS + 2 * kTileBytes;      // [2][128][128B] fp32 staging of one dQ tile (columns 0-31 | 32-63)
    MhaBwdBarriersP* bars = reinterpret_cast<MhaBwdBarriersP*>(sDQ + 2 * kTileBytes);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(&bars->kv_full, 1);
        mbar_init(&bars->kv_free, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->qdo_full[s], 1);
            mbar_init(&bars->qdo_empty[s], 1);
        }
        mbar_init(&bars->sdp_full, 1);
        mbar_init(&bars->pds_full, 4 * CG);   // one arrival per softmax warp
        mbar_init(&bars->dq_full, 1);
        mbar_init(&bars->dq_free, 4);
        mbar_init(&bars->acc_free, 4 * CG);
        fence_mbar_init();
        if (DROP) {
            uint32_t lo = a.seed_lo, hi = a.seed_hi;
            effective_seed(a.seed_dev, lo, hi);
            bars->seed[0] = lo;
            bars->seed[1] = hi;
        }
    }
    constexpr int kTmaWarp = 4 * CG, kMmaWarp = 4 * CG + 1;
    if (warp == kMmaWarp) {
        tmem_alloc(&bars->tmem_base, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    const uint32_t tm_s = tmem, tm_dp = tmem + 128, tm_dv = tmem + 256, tm_dk = tmem + 320, tm_dq = tmem + 384;

    // CTA-wide counters, kept alike by every role: g = query tiles processed before the current item, m = items with at
    // least one tile before the current one (phases of kv_full / kv_free / acc_free)
    if (warp >= 4 * CG && warp < 4 * CG + 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");   // TMA, MMA, two idle warps
    if (warp >= 4 * CG + 4) {
        // ===== dQ warps: thread = query row of the tile =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
        const bool lead = (warp == 4 * CG + 4) && lane == 0;
        int g = 0;
        for (MhaBwdWalk walk(a, n_items, nkt); walk.valid(); walk.next()) {
            const MhaBwdItem item = walk.get(a);
            for (int it = 0; it < item.nsteps; ++it, ++g) {
                mbar_wait(&bars->dq_full, g & 1);
                tc_fence_after();
                if (lead) bulk_wait_read0();                      // the previous tile's reduce has read the staging
                bar_sync_named(1, 128);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {                  // columns 0-31 | 32-63: one [128 x 32] fp32 box each
                    float dq[32];
                    tmem_ld32(tm_dq + lane_base + 32 * hf, dq);
                    unsigned char* dst = sDQ + hf * kTileBytes + row * 128;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const int chunk = (i >> 2) ^ (row & 7);
                        *reinterpret_cast<float4*>(dst + chunk * 16) = make_float4(dq[i], dq[i + 1], dq[i + 2], dq[i + 3]);
                    }
                }
                tc_fence_before();
                mbar_arrive_warp(&bars->dq_free);                 // the accumulator may take the next tile's dS K
                fence_proxy_async();
                bar_sync_named(2, 128);
                if (lead) {
                    tma_reduce_add_4d(&tm_dqacc, sDQ, 0, item.h, (item.i_start + it) * kBM, item.b);
                    tma_reduce_add_4d(&tm_dqacc, sDQ + kTileBytes, 32, item.h, (item.i_start + it) * kBM, item.b);
                    bulk_commit();
                }
            }
        }
        if (lead) bulk_wait0();                                   // the reduces must have landed before the kernel ends
    } else if (warp == kTmaWarp) {
        // ===== TMA producer =====
        if (elect_one_sync()) {
            tma_prefetch_desc(&tm_q);
            tma_prefetch_desc(&tm_do);
            tma_prefetch_desc(&tm_k);
            tma_prefetch_desc(&tm_v);
            int g = 0, m = 0;
            for (MhaBwdWalk walk(a, n_items, nkt); walk.valid(); walk.next()) {
                const MhaBwdItem item = walk.get(a);
                if (item.nsteps == 0) continue;
                if (m > 0) mbar_wait(&bars->kv_free, (m - 1) & 1);     // the previous item's last products have read K, V
                mbar_arrive_expect_tx(&bars->kv_full, 2 * kTileBytes);
                tma_load_4d(sK, &tm_k, 0, item.h, item.key0, item.b, &bars->kv_full);
                tma_load_4d(sV, &tm_v, 0, item.h, item.key0, item.b, &bars->kv_full);
                for (int it = 0; it < item.nsteps; ++it, ++g) {
                    const int s = g & 1;
                    if (g >= 2) mbar_wait(&bars->qdo_empty[s], ((g >> 1) - 1) & 1);
                    mbar_arrive_expect_tx(&bars->qdo_full[s], 2 * kTileBytes);
                    tma_load_4d(sQ + s * kTileBytes, &tm_q, 0, item.h, (item.i_start + it) * kBM, item.b, &bars->qdo_full[s]);
                    tma_load_4d(sDO + s * kTileBytes, &tm_do, 0, item.h, (item.i_start + it) * kBM, item.b, &bars->qdo_full[s]);
                }
                ++m;
            }
        }
    } else if (warp == kMmaWarp) {
        // ===== MMA issuer =====
        // Order on the tensor pipe: S/dP of tile t+1 go BEFORE dV/dK/dQ of tile t, so the softmax warps can start on tile
        // t+1 while the three accumulating products of tile t run.  Across an item boundary "tile t+1" is the first tile of
        // the next item, whose S / dP need the next item's K, V: they follow the current item's last dV / dK / dQ instead.
        if (elect_one_sync()) {
            constexpr uint32_t id_s = make_idesc(kBM, kBN, 0, 0);     // S  = Q K^T   / dP = dO V^T
            constexpr uint32_t id_t = make_idesc(kBN, kD, 1, 1);      // dV = P^T dO  / dK = dS^T Q  (A and B MN-major)
            constexpr uint32_t id_q = make_idesc(kBM, kD, 0, 1);      // dQ = dS K    (A K-major, B MN-major)
            const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
            const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sDS);
            auto issue_sdp = [&](int gt) {
                const int s = gt & 1;
                const uint32_t q_addr = smem_u32(sQ + s * kTileBytes);
                const uint32_t do_addr = smem_u32(sDO + s * kTileBytes);
                mbar_wait(&bars->qdo_full[s], (gt >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < kD / 16; ++kk)
                    umma_bf16(tm_s, smem_desc_sw128(q_addr + kk * 32, 16, 1024), smem_desc_sw128(k_addr + kk * 32, 16, 1024),
                              id_s, kk > 0 ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < kD / 16; ++kk)
                    umma_bf16(tm_dp, smem_desc_sw128(do_addr + kk * 32, 16, 1024), smem_desc_sw128(v_addr + kk * 32, 16, 1024),
                              id_s, kk > 0 ? 1u : 0u);
                tc_commit(&bars->sdp_full);
            };
            int g = 0, m = 0;
            for (MhaBwdWalk walk(a, n_items, nkt); walk.valid(); walk.next()) {
                const MhaBwdItem item = walk.get(a);
                if (item.nsteps == 0) continue;
                mbar_wait(&bars->kv_full, m & 1);
                // S / dP of the item's first tile: the S / dP columns are free (the softmax warps arrived on pds_full of the
                // previous tile after reading them; that barrier was waited for before the previous tile's products)
                issue_sdp(g);
                for (int it = 0; it < item.nsteps; ++it, ++g) {
                    const int s = g & 1;
                    const uint32_t q_addr = smem_u32(sQ + s * kTileBytes);
                    const uint32_t do_addr = smem_u32(sDO + s * kTileBytes);
                    // P/dS of this tile are in shared memory; S/dP (TMEM) have been read
                    mbar_wait(&bars->pds_full, g & 1);
                    tc_fence_after();
                    if (it + 1 < item.nsteps) issue_sdp(g + 1);
                    if (it == 0 && m > 0) {
                        mbar_wait(&bars->acc_free, (m - 1) & 1);      // the previous item's epilogue has read dK, dV
                        tc_fence_after();
                    }
                    // contraction over the 128 queries of the tile: 8 steps of 16 rows (2048 bytes)
#pragma unroll
                    for (int kk = 0; kk < kBM / 16; ++kk) {
                        umma_bf16(tm_dv, smem_desc_sw128(p_addr + kk * 2048, kTileBytes, 1024),
                                  smem_desc_sw128(do_addr + kk * 2048, kTileBytes, 1024), id_t, (it > 0 || kk > 0) ? 1u : 0u);
                    }
#pragma unroll
                    for (int kk = 0; kk < kBM / 16; ++kk) {
                        umma_bf16(tm_dk, smem_desc_sw128(ds_addr + kk * 2048, kTileBytes, 1024),
                                  smem_desc_sw128(q_addr + kk * 2048, kTileBytes, 1024), id_t, (it > 0 || kk > 0) ? 1u : 0u);
                    }
                    // contraction over the 128 keys: dS K-major (two 64-key halves), K tile MN-major
                    if (g > 0) {
                        mbar_wait(&bars->dq_free, (g - 1) & 1);       // the dQ warps have read the previous tile's dQ
                        tc_fence_after();
                    }
#pragma unroll
                    for (int kk = 0; kk < kBN / 16; ++kk) {
                        umma_bf16(tm_dq, smem_desc_sw128(ds_addr + (kk >> 2) * kTileBytes + (kk & 3) * 32, 16, 1024),
                                  smem_desc_sw128(k_addr + kk * 2048, kTileBytes, 1024), id_q, kk > 0 ? 1u : 0u);
                    }
                    tc_commit(&bars->dq_full);
                    tc_commit(&bars->qdo_empty[s]);
                    if (it + 1 == item.nsteps) tc_commit(&bars->kv_free);     // K, V may be overwritten behind these products
                }
                ++m;
            }
        }
    } else if (warp >= 4 * CG) {
        // (the two idle warps of the utility warpgroup)
    } else {
        // ===== softmax-backward warps: thread = (query row of the tile, column group); key row in the epilogue =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        const int row = (warp & 3) * 32 + lane;
        const int cg = warp >> 2;
        constexpr int kColsS = kBN / CG;    // S / dP columns of this thread
        constexpr int kColsD = kD / CG;     // dK / dV columns of this thread
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const float c = a.scale_log2;
        int g = 0;
        for (MhaBwdWalk walk(a, n_items, nkt); walk.valid(); walk.next()) {
            const MhaBwdItem item = walk.get(a);
            const int b = item.b, h = item.h, key0 = item.key0, i_start = item.i_start, nsteps = item.nsteps;
            const int kvlen = a.kv_len ? min(max(__ldg(a.kv_len + b), 0), a.Lk) : a.Lk;
            // row statistics of the next tile are fetched one tile ahead
            auto load_stats = [&](int it, float& lse2_o, float& dlt_o) {
                const int qi = (i_start + it) * kBM + row;
                const bool ok = it < nsteps && qi < a.Lq;
                const size_t stat = ((size_t)b * a.Hh + h) * a.Lq + (ok ? qi : 0);
                lse2_o = ok ? __ldg(a.lse + stat) * 1.4426950408889634f : INFINITY;
                dlt_o = ok ? __ldg(a.delta + stat) : 0.0f;
            };
            float lse2_n, dlt_n;
            load_stats(0, lse2_n, dlt_n);
            for (int it = 0; it < nsteps; ++it, ++g) {
                const int qi = (i_start + it) * kBM + row;
                const float lse2 = lse2_n, dlt = dlt_n;
                load_stats(it + 1, lse2_n, dlt_n);
                int lim = kvlen;
                if (a.causal) lim = min(lim, qi + 1);
                const uint8_t* mrow = a.dense_mask ? a.dense_mask + ((size_t)b * a.Lq + min(qi, a.Lq - 1)) * a.Lk : nullptr;
                const bool need_mask = (key0 + kBN > lim) || (mrow != nullptr);
                uint32_t dead = 0;               // dead keys among this thread's 32 columns
                if (__builtin_expect(need_mask, 0)) {
                    const int first = key0 + cg * kColsS;
                    const int live = lim - first;
                    dead = live >= kColsS ? 0u : live <= 0 ? 0xffffffffu : (0xffffffffu << live);
                    if (mrow != nullptr) {
                        const int kmax = min(kColsS, a.Lk - first);
#pragma unroll 1
                        for (int kk = 0; kk < kmax; ++kk) dead |= (uint32_t)(mrow[first + kk] != 0) << kk;
                    }
                }
                uint32_t pk[kColsS / 2], dk[kColsS / 2];
                mbar_wait(&bars->sdp_full, g & 1);
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < kColsS; c0 += 16) {
                    const int cc = cg * kColsS + c0;
                    float sv[16], dp[16];
                    tmem_ld16(tm_s + lane_base + cc, sv);
                    tmem_ld16(tm_dp + lane_base + cc, dp);
                    uint4 rnd[1];
                    if (DROP) rnd[0] = philox16((uint32_t)(key0 + cc) >> 4, (uint32_t)qi, (uint32_t)(b * a.Hh + h), bars->seed[0], bars->seed[1]);
                    const uint64_t c2 = pack_f32x2(c, c), nl2 = pack_f32x2(-lse2, -lse2);
                    const uint64_t sc2 = pack_f32x2(a.scale, a.scale), nd2 = pack_f32x2(-dlt * a.scale, -dlt * a.scale);
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        float x0, x1;
                        unpack_f32x2(fma_f32x2(pack_f32x2(sv[i], sv[i + 1]), c2, nl2), x0, x1);
                        float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
                        if ((dead >> (c0 + i)) & 1u) p0 = 0.0f;
                        if ((dead >> (c0 + i + 1)) & 1u) p1 = 0.0f;
                        const uint64_t p2 = pack_f32x2(p0, p1);
                        uint64_t g2 = pack_f32x2(dp[i], dp[i + 1]);      // d(loss)/d(dropped, rescaled probability)
                        uint64_t pd2 = p2;                               // the probabilities P V was computed with
                        if (DROP) {
                            const float k0 = philox_byte(rnd[0], i) >= a.drop_thresh ? a.inv_keep : 0.0f;
                            const float k1 = philox_byte(rnd[0], i + 1) >= a.drop_thresh ? a.inv_keep : 0.0f;
                            const uint64_t kf2 = pack_f32x2(k0, k1);
                            g2 = mul_f32x2(g2, kf2);
                            pd2 = mul_f32x2(p2, kf2);
                        }
                        float s0, s1, q0, q1;
                        unpack_f32x2(mul_f32x2(p2, fma_f32x2(g2, sc2, nd2)), s0, s1);
                        unpack_f32x2(pd2, q0, q1);
                        pk[(c0 + i) >> 1] = cvt_bf16x2(q0, q1);
                        dk[(c0 + i) >> 1] = cvt_bf16x2(s0, s1);
                    }
                }
                if (g > 0) {
                    // dV/dK/dQ of the previous tile (of this item or the one before) are done: the P / dS tiles are free
                    mbar_wait(&bars->dq_full, (g - 1) & 1);
                    tc_fence_after();
                }
#pragma unroll
                for (int c0 = 0; c0 < kColsS; c0 += 32) {
                    const int cc = cg * kColsS + c0;
                    const int off = (cc >> 6) * kTileBytes + row * 128;
                    const int chunk0 = (cc & 63) >> 3;
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const int chunk = (chunk0 + q4) ^ (row & 7);
                        const int w0 = (c0 >> 1) + 4 * q4;
                        *reinterpret_cast<uint4*>(sP + off + chunk * 16) = make_uint4(pk[w0], pk[w0 + 1], pk[w0 + 2], pk[w0 + 3]);
                        *reinterpret_cast<uint4*>(sDS + off + chunk * 16) = make_uint4(dk[w0], dk[w0 + 1], dk[w0 + 2], dk[w0 + 3]);
                    }
                }
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive_warp(&bars->pds_full);
            }
            // epilogue: dK_j, dV_j (thread = key row, kColsD columns)
            float acc[2][kColsD];
            if (nsteps > 0) {
                mbar_wait(&bars->dq_full, (g - 1) & 1);      // every product of the key tile is done: dK, dV are final
                tc_fence_after();
                tmem_ld16(tm_dk + lane_base + cg * kColsD, acc[0]);
                tmem_ld16(tm_dv + lane_base + cg * kColsD, acc[1]);
                tc_fence_before();
                mbar_arrive_warp(&bars->acc_free);           // the next item's first dV / dK products may overwrite them
            } else {
#pragma unroll
                for (int i = 0; i < kColsD; ++i) acc[0][i] = acc[1][i] = 0.0f;
            }
            const int key = key0 + row;
            if (key < a.Lk) {
                const size_t kv_off = (((size_t)b * a.Lk + key) * a.Hh + h) * kD + cg * kColsD;
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    __nv_bfloat16* dst = (part == 0 ? a.g_k : a.g_v) + kv_off;
#pragma unroll
                    for (int i = 0; i < kColsD; i += 8) {
                        uint32_t w[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) w[u] = cvt_bf16x2(acc[part][i + 2 * u], acc[part][i + 2 * u + 1]);
                        *reinterpret_cast<uint4*>(dst + i) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// delta[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d]   (one warp per row of 64)
// (also clears the row's 64 floats of the dQ accumulator [B,Lq,Hh,64], which has the same row order: one launch
// instead of a memset + this kernel)
__global__ void __launch_bounds__(256) mha_delta_kernel(const __nv_bfloat16* o, const __nv_bfloat16* d_o, float* delta,
                                                        float* dq_acc, int B, int Hh, int Lq) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // (b, q, h)
    if (row >= (long long)B * Lq * Hh) return;
    reinterpret_cast<float2*>(dq_acc + row * kD)[lane] = make_float2(0.0f, 0.0f);
    const __nv_bfloat162 ov = reinterpret_cast<const __nv_bfloat162*>(o + row * kD)[lane];
    const __nv_bfloat162 gv = reinterpret_cast<const __nv_bfloat162*>(d_o + row * kD)[lane];
    float s = __low2float(ov) * __low2float(gv) + __high2float(ov) * __high2float(gv);
    s = warp_sum(s);
    if (lane == 0) {
        const int hh = (int)(row % Hh);
        const long long bq = row / Hh;
        const int qi = (int)(bq % Lq);
        const int bb = (int)(bq / Lq);
        delta[((size_t)bb * Hh + hh) * Lq + qi] = s;
    }
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* src, __nv_bfloat16* dst, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        const __nv_bfloat162 a2 = __floats2bfloat162_rn(v.x, v.y), b2 = __floats2bfloat162_rn(v.z, v.w);
        reinterpret_cast<uint2*>(dst)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&a2), *reinterpret_cast<const uint32_t*>(&b2));
    }
}

// ---- attention probabilities (only when a caller asks for `attn`) ----------------------
// attn[(h*B + b), q, k] f32, the reference's head-major order.  One warp per (b,h,q) row; CUDA cores.
__global__ void __launch_bounds__(128) mha_probs_kernel(const __nv_bfloat16* q, const __nv_bfloat16* k, const int* kv_len,
                                                        const uint8_t* dense_mask, int causal, int B, int Hh, int Lq,
                                                        int Lk, float scale, float* attn) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= (long long)B * Hh * Lq) return;
    const int qi = (int)(row % Lq);
    const int h = (int)((row / Lq) % Hh);
    const int b = (int)(row / ((long long)Lq * Hh));
    const int kvlen = kv_len ? min(max(kv_len[b], 0), Lk) : Lk;
    const __nv_bfloat16* qr = q + (((size_t)b * Lq + qi) * Hh + h) * kD;
    float* dst = attn + (((size_t)h * B + b) * Lq + qi) * Lk;
    const float q0 = __bfloat162float(qr[lane]), q1 = __bfloat162float(qr[lane + 32]);
    float mx = -INFINITY;
    for (int key = 0; key < Lk; ++key) {
        const __nv_bfloat16* kr = k + (((size_t)b * Lk + key) * Hh + h) * kD;
        float s = warp_sum(q0 * __bfloat162float(kr[lane]) + q1 * __bfloat162float(kr[lane + 32])) * scale;
        bool dead = key >= kvlen || (causal && key > qi);
        if (dense_mask) dead = dead || dense_mask[((size_t)b * Lq + qi) * Lk + key] != 0;
        if (dead) s = -INFINITY;
        if (lane == 0) dst[key] = s;
        mx = fmaxf(mx, s);
    }
    __syncwarp();
    float sum = 0.0f;
    for (int key = lane; key < Lk; key += 32) sum += expf(dst[key] - mx);
    sum = warp_sum(sum);
    for (int key = lane; key < Lk; key += 32) dst[key] = expf(dst[key] - mx) / sum;
}

// keep[b,h,q,k] = 1 when the dropout of (seed, threshold) keeps that probability - the same bits the
// attention kernels regenerate; for tests and for callers that want to inspect a mask
__global__ void __launch_bounds__(256) mha_dropout_keep_kernel(uint8_t* keep, int B, int Hh, int Lq, int Lk, uint32_t thresh,
                                                              uint32_t seed_lo, uint32_t seed_hi) {
    const int k16n = (Lk + 15) >> 4;
    const long long total = (long long)B * Hh * Lq * k16n;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int k16 = (int)(idx % k16n);
    const long long rq = idx / k16n;
    const int q = (int)(rq % Lq);
    const int bh = (int)(rq / Lq);
    const uint4 rnd = philox16((uint32_t)k16, (uint32_t)q, (uint32_t)bh, seed_lo, seed_hi);
    uint8_t* dst = keep + ((size_t)bh * Lq + q) * Lk + (size_t)k16 * 16;
    for (int i = 0; i < 16 && k16 * 16 + i < Lk; ++i) dst[i] = philox_byte(rnd, i) >= thresh ? 1 : 0;
}

static int make_qkv_map(CUtensorMap* map, const void* base, int B, int L, int Hh) {
    const uint64_t dims[4] = {(uint64_t)kD, (uint64_t)Hh, (uint64_t)L, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)kD * 2, (uint64_t)Hh * kD * 2, (uint64_t)L * Hh * kD * 2};
    const uint32_t box[4] = {(uint32_t)kD, 1u, (uint32_t)kBM, 1u};
    return make_tmap_nd(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace asr

using namespace asr;

static int mha_fwd_impl(const void* q, const void* k, const void* v, const int* kv_len, const uint8_t* dense_mask,
                        int causal, int B, int Hh, int Lq, int Lk, int D, float scale, void* out, float* lse,
                        float p_drop, uint64_t seed, const uint64_t* seed_dev, void* stream) {
    ASR_REQUIRE(q && k && v && out, "asr_mha_fwd_bf16: null pointer");
    ASR_REQUIRE(D == kD, "asr_mha_fwd_bf16: head dim %d not supported (64 only)", D);
    ASR_REQUIRE(B > 0 && Hh > 0 && Lq > 0 && Lk > 0, "asr_mha_fwd_bf16: bad shape B=%d Hh=%d Lq=%d Lk=%d", B, Hh, Lq, Lk);
    ASR_REQUIRE(B <= 65535 && Hh <= 65535, "asr_mha_fwd_bf16: B/Hh exceed the grid limits");
    ASR_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(out), "asr_mha_fwd_bf16: pointers must be 16-byte aligned");
    ASR_REQUIRE(p_drop >= 0.0f && p_drop < 1.0f, "asr_mha_fwd_dropout_bf16: p_drop %f outside [0, 1)", (double)p_drop);
    if (asr_device_ok() != 0) return 3;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUtensorMap tq, tk, tv;
    if (make_qkv_map(&tq, q, B, Lq, Hh) || make_qkv_map(&tk, k, B, Lk, Hh) || make_qkv_map(&tv, v, B, Lk, Hh)) return 4;
    MhaFwdArgs a;
    a.kv_len = kv_len;
    a.dense_mask = dense_mask;
    a.causal = causal;
    a.B = B; a.Hh = Hh; a.Lq = Lq; a.Lk = Lk;
    a.scale_log2 = scale * 1.4426950408889634f;
    a.out = static_cast<__nv_bfloat16*>(out);
    a.lse = lse;
    a.drop_thresh = drop_threshold(p_drop);
    a.inv_keep = 256.0f / (256.0f - (float)a.drop_thresh);
    a.seed_lo = (uint32_t)seed;
    a.seed_hi = (uint32_t)(seed >> 32);
    a.seed_dev = seed_dev;
    // Three kernels (the earlier generations are gone, DESIGN.md 4 keeps their numbers):
    //   mha_fwdp_kernel<DROP>      PERSISTENT: one CTA per SM walks the (batch, head, 256-query) items; per item two 128-query
    //                              tiles share every K/V tile, one thread per query row, P in tensor memory.  Every call with
    //                              more than 128 queries, with and without dropout (L = 2048: 722 TFLOP/s, 537 with dropout)
    //   mha_fwd8_kernel<DROP>      the same inner structure with one CTA per item (L = 2048: 707 TFLOP/s): the reference point
    //                              of the persistent kernel (mha_variant = 21; outputs bit-identical without dropout)
    //   mha_fwd3_kernel<DROP>      CTA = one 128-query tile, two threads per row, two CTAs per SM: short query sequences
    //                              (Lq <= 128: the second tile of a 256-query item would be empty; model shapes are
    //                              L = 21 .. 167, U <= 15); with dropout at L = 2048: 481 TFLOP/s
    // "mha_variant": 0 = that rule, 3 = always mha_fwd3_kernel, 21 = always mha_fwd8_kernel, 40 = always mha_fwdp_kernel.
    const int variant = get_opt("mha_variant");
    const bool drop = a.drop_thresh > 0;
    const bool one_tile = variant == 3 || (variant != 21 && variant != 40 && Lq <= kBM);
    const bool persistent = !one_tile && (variant == 40 || variant == 0);
    static bool attr_done[4] = {false, false, false, false};      // cudaFuncSetAttribute once per kernel, not per call
    if (one_tile) {
        dim3 grid((Lq + kBM - 1) / kBM, Hh, B);
        if (drop) {
            if (!attr_done[0]) { ASR_CHECK_CUDA(cudaFuncSetAttribute(mha_fwd3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwd3Smem)); attr_done[0] = true; }
            mha_fwd3_kernel<true><<<grid, kFwd3Threads, kFwd3Smem, st>>>(tq, tk, tv, a);
        } else {
            if (!attr_done[1]) { ASR_CHECK_CUDA(cudaFuncSetAttribute(mha_fwd3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwd3Smem)); attr_done[1] = true; }
            mha_fwd3_kernel<false><<<grid, kFwd3Threads, kFwd3Smem, st>>>(tq, tk, tv, a);
        }
    } else {
        dim3 grid((Lq + 2 * kBM - 1) / (2 * kBM), Hh, B);
        if (persistent) {
            const int nq2 = (Lq + 2 * kBM - 1) / (2 * kBM);
            const int n_items = B * Hh * nq2;
            const dim3 pgrid(std::min(n_items, num_sms()));
            if (drop) {
                static bool done = false;
                if (!done) { ASR_CHECK_CUDA(cudaFuncSetAttribute(mha_fwdp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdPSmem)); done = true; }
                mha_fwdp_kernel<true><<<pgrid, kFwd8Threads, kFwdPSmem, st>>>(tq, tk, tv, a, n_items, nq2);
            } else {
                static bool done = false;
                if (!done) { ASR_CHECK_CUDA(cudaFuncSetAttribute(mha_fwdp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdPSmem)); done = true; }
                mha_fwdp_kernel<false><<<pgrid, kFwd8Threads, kFwdPSmem, st>>>(tq, tk, tv, a, n_items, nq2);
            }
        } else
        if (drop) {
            if (!attr_done[2]) { ASR_CHECK_CUDA(cudaFuncSetAttribute(mha_fwd8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwd8Smem)); attr_done[2] = true; }
            mha_fwd8_kernel<true><<<grid, kFwd8Threads, kFwd8Smem, st>>>(tq, tk, tv, a);
        } else {
            if (!attr_done[3]) { ASR_CHECK_CUDA(cudaFuncSetAttribute(mha_fwd8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwd8Smem)); attr_done[3] = true; }
            mha_fwd8_kernel<false><<<grid, kFwd8Threads, kFwd8Smem, st>>>(tq, tk, tv, a);
        }
    }
    ASR_LAUNCH_CHECK();
    return 0;
}

extern "C" int asr_mha_fwd_bf16(const void* q, const void* k, const void* v, const int* kv_len, const uint8_t* dense_mask,
                                int causal, int B, int Hh, int Lq, int Lk, int D, float scale, void* out, float* lse,
                                void* stream) {
    return mha_fwd_impl(q, k, v, kv_len, dense_mask, causal, B, Hh, Lq, Lk, D, scale, out, lse, 0.0f, 0, nullptr, stream);
}

extern "C" int asr_mha_fwd_dropout_bf16(const void* q, const void* k, const void* v, const int* kv_len,
                                        const uint8_t* dense_mask, int causal, int B, int Hh, int Lq, int Lk, int D,
                                        float scale, float p_drop, uint64_t seed, void* out, float* lse, void* stream) {
    return mha_fwd_impl(q, k, v, kv_len, dense_mask, causal, B, Hh, Lq, Lk, D, scale, out, lse, p_drop, seed, nullptr, stream);
}

extern "C" int asr_mha_fwd_dropout_dev_bf16(const void* q, const void* k, const void* v, const int* kv_len,
                                            const uint8_t* dense_mask, int causal, int B, int Hh, int Lq, int Lk, int D,
                                            float scale, float p_drop, const uint64_t* seed_dev, uint64_t seed_add, void* out,
                                            float* lse, void* stream) {
    ASR_REQUIRE(seed_dev != nullptr, "asr_mha_fwd_dropout_dev_bf16: seed_dev is null");
    return mha_fwd_impl(q, k, v, kv_len, dense_mask, causal, B, Hh, Lq, Lk, D, scale, out, lse, p_drop, seed_add, seed_dev, stream);
}

#ifdef ASR_MHA_TRACE
extern "C" int asr_debug_mha_trace(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, asr::g_mha_trace, sizeof(long long) * 2 * 16 * 20);
}
extern "C" int asr_debug_mha_cta(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, asr::g_mha_cta, sizeof(long long) * 2048 * 5);
}
#endif

extern "C" float asr_mha_dropout_keep_prob(float p_drop) { return (256.0f - (float)drop_threshold(p_drop)) / 256.0f; }

extern "C" int asr_mha_dropout_keep_u8(int B, int Hh, int Lq, int Lk, float p_drop, uint64_t seed, uint8_t* keep, void* stream) {
    ASR_REQUIRE(keep != nullptr && B > 0 && Hh > 0 && Lq > 0 && Lk > 0, "asr_mha_dropout_keep_u8: bad arguments");
    if (asr_device_ok() != 0) return 3;
    const long long total = (long long)B * Hh * Lq * ((Lk + 15) >> 4);
    mha_dropout_keep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        keep, B, Hh, Lq, Lk, drop_threshold(p_drop), (uint32_t)seed, (uint32_t)(seed >> 32));
    ASR_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t asr_mha_bwd_workspace_bytes(int B, int Hh, int Lq, int Lk, int D) {
    (void)Lk;
    if (B <= 0 || Hh <= 0 || Lq <= 0 || D <= 0) return 0;
    // fp32 dQ accumulator [B,Lq,Hh,D] + delta [B,Hh,Lq]
    return (size_t)B * Lq * Hh * D * sizeof(float) + (size_t)B * Hh * Lq * sizeof(float) + 256;
}

static int mha_bwd_impl(const void* q, const void* k, const void* v, const void* out, const void* g_out,
                        const float* lse, const int* kv_len, const uint8_t* dense_mask, int causal, int B, int Hh,
                        int Lq, int Lk, int D, float scale, float p_drop, uint64_t seed, const uint64_t* seed_dev, void* g_q,
                        void* g_k, void* g_v, void* ws, size_t ws_bytes, void* stream) {
    ASR_REQUIRE(q && k && v && out && g_out && lse && g_q && g_k && g_v && ws, "asr_mha_bwd_bf16: null pointer");
    ASR_REQUIRE(D == kD, "asr_mha_bwd_bf16: head dim %d not supported (64 only)", D);
    ASR_REQUIRE(B > 0 && Hh > 0 && Lq > 0 && Lk > 0, "asr_mha_bwd_bf16: bad shape B=%d Hh=%d Lq=%d Lk=%d", B, Hh, Lq, Lk);
    ASR_REQUIRE(B <= 65535 && Hh <= 65535, "asr_mha_bwd_bf16: B/Hh exceed the grid limits");
    ASR_REQUIRE(ws_bytes >= asr_mha_bwd_workspace_bytes(B, Hh, Lq, Lk, D), "asr_mha_bwd_bf16: workspace too small");
    ASR_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(g_out) && aligned16(g_q) && aligned16(g_k) &&
                    aligned16(g_v), "asr_mha_bwd_bf16: pointers must be 16-byte aligned");
    if (asr_device_ok() != 0) return 3;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uintptr_t w = (reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255;
    float* dq_acc = reinterpret_cast<float*>(w);
    const size_t nq_elems = (size_t)B * Lq * Hh * kD;
    float* delta = dq_acc + nq_elems;
    const long long rows = (long long)B * Lq * Hh;
    mha_delta_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(out),
                                                                 static_cast<const __nv_bfloat16*>(g_out), delta, dq_acc, B, Hh, Lq);
    ASR_LAUNCH_CHECK();
    CUtensorMap tq, tk, tv, tdo, tdq;
    if (make_qkv_map(&tq, q, B, Lq, Hh) || make_qkv_map(&tk, k, B, Lk, Hh) || make_qkv_map(&tv, v, B, Lk, Hh) ||
        make_qkv_map(&tdo, g_out, B, Lq, Hh))
        return 4;
    {   // fp32 dQ accumulator [B,Lq,Hh,64]: boxes of [128 rows x 32 columns] (128-byte rows, swizzled)
        const uint64_t dims[4] = {(uint64_t)kD, (uint64_t)Hh, (uint64_t)Lq, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)kD * 4, (uint64_t)Hh * kD * 4, (uint64_t)Lq * Hh * kD * 4};
        const uint32_t box[4] = {32u, 1u, (uint32_t)kBM, 1u};
        if (make_tmap_nd(&tdq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dq_acc, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
            return 4;
    }
    MhaBwdArgs a;
    a.kv_len = kv_len;
    a.dense_mask = dense_mask;
    a.causal = causal;
    a.B = B; a.Hh = Hh; a.Lq = Lq; a.Lk = Lk;
    a.scale = scale;
    a.scale_log2 = scale * 1.4426950408889634f;
    a.lse = lse;
    a.delta = delta;
    a.dq_acc = dq_acc;
    a.g_k = static_cast<__nv_bfloat16*>(g_k);
    a.g_v = static_cast<__nv_bfloat16*>(g_v);
    ASR_REQUIRE(p_drop >= 0.0f && p_drop < 1.0f, "asr_mha_bwd_dropout_bf16: p_drop %f outside [0, 1)", (double)p_drop);
    a.drop_thresh = drop_threshold(p_drop);
    a.inv_keep = 256.0f / (256.0f - (float)a.drop_thresh);
    a.seed_lo = (uint32_t)seed;
    a.seed_hi = (uint32_t)(seed >> 32);
    a.seed_dev = seed_dev;
    const bool drop = a.drop_thresh > 0;
    dim3 grid((Lk + kBN - 1) / kBN, Hh, B);
    const int cgo = get_opt("mha_bwd_groups");   // 0 / 7 = persistent kernel (16 softmax-backward warps + 4 dQ warps; default),
                                                 // 5 = the same with one CTA per item, 4 = 16 warps that also flush dQ, 2 = 8 warps
#define ASR_LAUNCH_BWD_DQW(DR)                                                                                              \
    do {                                                                                                                   \
        static bool attr_done = false;                                                                                     \
        if (!attr_done) {                                                                                                  \
            ASR_CHECK_CUDA(cudaFuncSetAttribute(mha_bwd_kernel<4, DR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem)); \
            attr_done = true;                                                                                              \
        }                                                                                                                  \
        mha_bwd_kernel<4, DR, true><<<grid, 768, kBwdSmem, st>>>(tq, tk, tv, tdo, tdq, a);                                  \
    } while (0)
#define ASR_LAUNCH_BWD(CGV, DR)                                                                                              \
    do {                                                                                                                   \
        static bool attr_done = false;                                                                                     \
        if (!attr_done) {                                                                                                  \
            ASR_CHECK_CUDA(cudaFuncSetAttribute(mha_bwd_kernel<CGV, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem)); \
            attr_done = true;                                                                                              \
        }                                                                                                                  \
        mha_bwd_kernel<CGV, DR><<<grid, 128 * CGV + 64, kBwdSmem, st>>>(tq, tk, tv, tdo, tdq, a);                          \
    } while (0)
    if (cgo == 0 || cgo == 7) {          // default, persistent: one CTA per SM walks the (batch, head, key tile) items
        const int nkt = (Lk + kBN - 1) / kBN;
        const int n_items = B * Hh * nkt;
        const dim3 pgrid(std::min(n_items, num_sms()));
        if (drop) {
            static bool done = false;
            if (!done) { ASR_CHECK_CUDA(cudaFuncSetAttribute(mha_bwdp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem)); done = true; }
            mha_bwdp_kernel<true><<<pgrid, 768, kBwdSmem, st>>>(tq, tk, tv, tdo, tdq, a, n_items, nkt);
        } else {
            static bool done = false;
            if (!done) { ASR_CHECK_CUDA(cudaFuncSetAttribute(mha_bwdp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem)); done = true; }
            mha_bwdp_kernel<false><<<pgrid, 768, kBwdSmem, st>>>(tq, tk, tv, tdo, tdq, a, n_items, nkt);
        }
    } else if (cgo == 5) {               // one CTA per item: the 16-warp instance with dedicated dQ warps
        if (drop) ASR_LAUNCH_BWD_DQW(true); else ASR_LAUNCH_BWD_DQW(false);
    } else if (cgo == 2) {
        if (drop) ASR_LAUNCH_BWD(2, true); else ASR_LAUNCH_BWD(2, false);
    } else {
        if (drop) ASR_LAUNCH_BWD(4, true); else ASR_LAUNCH_BWD(4, false);
    }
#undef ASR_LAUNCH_BWD
#undef ASR_LAUNCH_BWD_DQW
    ASR_LAUNCH_CHECK();
    size_t blocks = (nq_elems / 4 + 255) / 256;
    if (blocks > (size_t)num_sms() * 16) blocks = (size_t)num_sms() * 16;
    f32_to_bf16_kernel<<<(unsigned)blocks, 256, 0, st>>>(dq_acc, static_cast<__nv_bfloat16*>(g_q), nq_elems / 4);
    ASR_LAUNCH_CHECK();
    return 0;
}

extern "C" int asr_mha_bwd_bf16(const void* q, const void* k, const void* v, const void* out, const void* g_out,
                                const float* lse, const int* kv_len, const uint8_t* dense_mask, int causal, int B, int Hh,
                                int Lq, int Lk, int D, float scale, void* g_q, void* g_k, void* g_v, void* ws,
                                size_t ws_bytes, void* stream) {
    return mha_bwd_impl(q, k, v, out, g_out, lse, kv_len, dense_mask, causal, B, Hh, Lq, Lk, D, scale, 0.0f, 0, nullptr, g_q, g_k,
                        g_v, ws, ws_bytes, stream);
}

extern "C" int asr_mha_bwd_dropout_bf16(const void* q, const void* k, const void* v, const void* out, const void* g_out,
                                        const float* lse, const int* kv_len, const uint8_t* dense_mask, int causal, int B,
                                        int Hh, int Lq, int Lk, int D, float scale, float p_drop, uint64_t seed, void* g_q,
                                        void* g_k, void* g_v, void* ws, size_t ws_bytes, void* stream) {
    return mha_bwd_impl(q, k, v, out, g_out, lse, kv_len, dense_mask, causal, B, Hh, Lq, Lk, D, scale, p_drop, seed, nullptr, g_q,
                        g_k, g_v, ws, ws_bytes, stream);
}

extern "C" int asr_mha_bwd_dropout_dev_bf16(const void* q, const void* k, const void* v, const void* out, const void* g_out,
                                            const float* lse, const int* kv_len, const uint8_t* dense_mask, int causal, int B,
                                            int Hh, int Lq, int Lk, int D, float scale, float p_drop, const uint64_t* seed_dev,
                                            uint64_t seed_add, void* g_q, void* g_k, void* g_v, void* ws, size_t ws_bytes,
                                            void* stream) {
    ASR_REQUIRE(seed_dev != nullptr, "asr_mha_bwd_dropout_dev_bf16: seed_dev is null");
    return mha_bwd_impl(q, k, v, out, g_out, lse, kv_len, dense_mask, causal, B, Hh, Lq, Lk, D, scale, p_drop, seed_add, seed_dev,
                        g_q, g_k, g_v, ws, ws_bytes, stream);
}

extern "C" int asr_mha_probs_f32(const void* q, const void* k, const int* kv_len, const uint8_t* dense_mask, int causal,
                                 int B, int Hh, int Lq, int Lk, int D, float scale, float* attn, void* stream) {
    ASR_REQUIRE(q && k && attn, "asr_mha_probs_f32: null pointer");
    ASR_REQUIRE(D == kD, "asr_mha_probs_f32: head dim %d not supported (64 only)", D);
    if (asr_device_ok() != 0) return 3;
    const long long rows = (long long)B * Hh * Lq;
    mha_probs_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k), kv_len, dense_mask, causal, B, Hh, Lq,
        Lk, scale, attn);
    ASR_LAUNCH_CHECK();
    return 0;
}

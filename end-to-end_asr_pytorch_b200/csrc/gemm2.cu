// gemm2.cu - general tcgen05 GEMMs for the TRAINING half of the linear layers and for the vocabulary
// projection fused with the CTC loss (8(f1)): C[M,N] = A * B with either operand stored K-major or MN-major, so the
// three products of a Linear layer y = x W^T run on the tensors as torch stores them - no transposed copies:
//
//     forward   y  = x  W^T     A = x  [M,K] K-major          B = W  [N,K]  K-major
//     dX        gx = gy W       A = gy [M,N] K-major          B = W  [N,K]  MN-major  (contraction over N)
//     dW        gW = gy^T x     A = gy [M,N] MN-major         B = x  [M,K]  MN-major  (contraction over the rows M)
//
// (/root/reference/src/transformer/module.py:46-53, attention.py:40-45,59-60, cif_model.py:38 and what autograd makes of them.)
//
// Two element types:
//   fp32 "3xTF32"  x = hi + lo (hi = the top 19 bits, what the tensor core reads of an fp32 word; lo = x - hi, exact in
//                  fp32), A B ~= lo_a hi_b + hi_a lo_b + hi_a hi_b as three kind::tf32 MMAs per K step into one fp32
//                  accumulator: fp32-level accuracy (~5e-6 relative at K = 512) on the tensor pipe.  The lo tiles are
//                  produced in shared memory by the four epilogue warps - an elementwise pass, so it is layout-agnostic.
//   bf16           one kind::f16 MMA per K step, fp32 accumulation, bf16 or fp32 output.
//
// Tiles: CTA = 128 x BN output tile (BN = 128 or 256), K step = 128 bytes (32 fp32 / 64 bf16).  A K-major tile is one TMA
// box [rows x 128 bytes]; an MN-major tile is a row of boxes [K step rows x 128 bytes], one per 128-byte chunk of the
// M / N extent - exactly the canonical 128-byte-swizzled MN-major layout of the tcgen05 shared-memory descriptor
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: LBO = one box, SBO = 1024 bytes (bf16); MN-major tf32 operands exist only
// in the 32-byte-atom flavour of that swizzle (TMA SWIZZLE_128B_ATOM_32B, descriptor layout 1, 4-row K groups, SBO = 512),
// which is what the fp32 kernel uses for them.  Out-of-range rows / columns are
// zero-filled by TMA, so odd sizes (the 4233-wide vocabulary) need no padding copies, only 16-byte-aligned row strides.
// Split-K over gridDim.z (dW of a small layer has few output tiles and a long contraction; at the model's shapes nearly
// every product has fewer tiles than the GPU has SMs).  The CTAs of one output tile form a thread-block CLUSTER along z:
// each leaves its partial tile in its own shared memory (the operand ring is free by then), and after a cluster barrier
// CTA z adds rows [z * 128 / S, ...) of all S partial tiles through distributed shared memory in a fixed order (bias
// first, then split 0, 1, ...: deterministic, no atomics) and stores them with row-contiguous 16-byte stores - no
// workspace, no second launch.  The older flavour (partial tiles to a workspace + a reduce kernel) remains for 256-wide
// tiles, whose partial tile does not fit next to nothing else in shared memory, and behind option gemm_split_mode = 1.
#include "common.cuh"
#include "tcgen05.cuh"

#include <cuda_bf16.h>

#include <algorithm>

namespace asr {

constexpr int kG2M = 128;          // rows of C per CTA
constexpr int kG2RowBytes = 128;   // bytes of one shared-memory row = one swizzle span

struct __align__(8) Gemm2Barriers {
    uint64_t full[4];
    uint64_t split[4];
    uint64_t empty[4];
    uint64_t acc_full;
    uint32_t tmem_base;
    uint32_t pad;
};

// ---- split-K inside a cluster -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// alone: a cluster of one CTA (or no cluster at all).  exit_only: the barrier only keeps a CTA's shared memory alive until its
// peers have read it - no data is handed over, so the arrival needs no release (which costs a MEMBAR.ALL.GPU behind the
// epilogue's global stores) and the wait no acquire (a CCTL.IVALL)
__device__ __forceinline__ void cluster_sync_all(bool alone, bool exit_only = false) {
    if (alone) __syncthreads();
    else if (exit_only) asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
    else asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_addr, uint32_t cta) {
    uint32_t remote;
    float4 v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(cta));
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote) : "memory");
    return v;
}
constexpr int kG2MaxSplits = 8;                               // = the portable cluster size
template <int BN>
struct G2Part {                                               // one CTA's partial tile in its shared memory
    static constexpr int kLd = BN + 4;                        // row stride in floats: a warp's 32 row-strided float4 stores hit 8 distinct bank groups per phase
    static constexpr int kBytes = kG2M * kLd * 4;
};
// thread `row` of the tile leaves 32 accumulator columns in the partial tile
template <int BN>
__device__ __forceinline__ void part_store32(unsigned char* smem, int row, int cc, const float (&v)[32]) {
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(smem) + (size_t)row * G2Part<BN>::kLd + cc);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
// After the cluster barrier: this CTA's share of the tile's rows, 128 threads, one float4 (4 columns of one row) per
// thread and pass, two passes in flight.  init(r, c) -> the float4 the sum starts from (bias), emit(r, c, sum).
template <int BN, typename Init, typename Emit>
__device__ __forceinline__ void cluster_splitk_reduce(unsigned char* smem, int tid128, int n_cols, Init init, Emit emit) {
    constexpr int kF4 = BN / 4;
    const int S = (int)cluster_nctarank(), z = (int)cluster_ctarank();
    const int rows_per = (kG2M + S - 1) / S;
    const int r_lo = z * rows_per, r_hi = min(kG2M, r_lo + rows_per);
    const int total = max(0, r_hi - r_lo) * kF4;
    const uint32_t base = smem_u32(smem);
    for (int idx0 = tid128; idx0 < total; idx0 += 256) {
        float4 part[2][kG2MaxSplits];
        int r[2], c[2];
        bool live[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int idx = idx0 + 128 * u;
            r[u] = r_lo + idx / kF4;
            c[u] = (idx % kF4) * 4;
            live[u] = idx < total && c[u] < n_cols;
            const uint32_t off = base + (uint32_t)(r[u] * G2Part<BN>::kLd + c[u]) * 4u;
#pragma unroll
            for (int s = 0; s < kG2MaxSplits; ++s)
                if (live[u] && s < S) part[u][s] = ld_dsmem_f4(off, (uint32_t)s);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!live[u]) continue;
            float4 acc = init(r[u], c[u]);
#pragma unroll
            for (int s = 0; s < kG2MaxSplits; ++s)
                if (s < S) { acc.x += part[u][s].x; acc.y += part[u][s].y; acc.z += part[u][s].z; acc.w += part[u][s].w; }
            emit(r[u], c[u], acc);
        }
    }
}

// One operand tile: ROWS = extent along M (A) or N (B), ELEM = bytes per element.
template <bool MN, int ROWS, int ELEM>
struct G2Tile {
    static constexpr int kChunk = kG2RowBytes / ELEM;       // elements per 128-byte row: 32 fp32 / 64 bf16
    static constexpr int kBK = kChunk;                        // K elements per stage
    static constexpr int kBytes = ROWS * kG2RowBytes;
    static constexpr int kBoxBytes = kBK * kG2RowBytes;       // MN-major: one box = kBK k-rows x 128 bytes
    static constexpr int kUmmaK = 32 / ELEM;                  // K of one MMA: 8 (tf32) / 16 (bf16)
    // mn0 = first row (A) / column (B) of the tile in the M / N extent, kc = first K element of the stage
    __device__ static __forceinline__ void load(unsigned char* tile, const CUtensorMap* map, int kc, int mn0, uint64_t* bar) {
        if (!MN) {
#pragma unroll
            for (int part = 0; part < (ROWS + 255) / 256; ++part)      // a TMA box has at most 256 rows
                tma_load_2d(tile + part * 256 * kG2RowBytes, map, kc, mn0 + part * 256, bar);
        } else {
#pragma unroll
            for (int c = 0; c < ROWS / kChunk; ++c) tma_load_2d(tile + c * kBoxBytes, map, mn0 + c * kChunk, kc, bar);
        }
    }
    __device__ static __forceinline__ uint64_t desc(uint32_t tile_addr, int kk) {
        if (!MN) return smem_desc_sw128(tile_addr + kk * 32, 16, 1024);
        // MN-major: 16-bit operands use the plain 128-byte swizzle (atoms of 8 K rows, SBO = 1024); 32-bit (tf32) operands
        // must use the 32-byte-atom flavour (atoms of 4 K rows, SBO = 512) - the tile was loaded with the matching TMA swizzle
        if (ELEM == 4) return smem_desc_sw128_base32(tile_addr + kk * kUmmaK * kG2RowBytes, kBoxBytes, 512);
        return smem_desc_sw128(tile_addr + kk * kUmmaK * kG2RowBytes, kBoxBytes, 1024);
    }
};

// ------------------------------------------------------------------------------------------------------------
// fp32, three TF32 products per K step
// ------------------------------------------------------------------------------------------------------------
template <int BN>
struct G2F32Cfg {
    static constexpr int kStages = BN == 256 ? 2 : 3;
    static constexpr int kATile = kG2M * kG2RowBytes;            // 16 KB
    static constexpr int kBTile = BN * kG2RowBytes;
    static constexpr int kStage = 2 * kATile + 2 * kBTile;       // [A | A lo | B | B lo]
    static constexpr int kSmem = kStages * kStage + 256;
};

__device__ __forceinline__ void split_lo_tile(const unsigned char* hi_tile, unsigned char* lo_tile, int bytes, int tid128) {
    const uint4* hi = reinterpret_cast<const uint4*>(hi_tile);
    float4* lo = reinterpret_cast<float4*>(lo_tile);
    for (int idx = tid128; idx < bytes / 16; idx += 128) {
        const uint4 v = hi[idx];
        lo[idx] = make_float4(__uint_as_float(v.x) - __uint_as_float(v.x & 0xffffe000u), __uint_as_float(v.y) - __uint_as_float(v.y & 0xffffe000u),
                              __uint_as_float(v.z) - __uint_as_float(v.z & 0xffffe000u), __uint_as_float(v.w) - __uint_as_float(v.w & 0xffffe000u));
    }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(192, 1)
gemm_f32x3_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const float* __restrict__ bias,
                  float* __restrict__ c, int M, int N, int K, int ldc, int stages_per_split, size_t split_stride,
                  const int* __restrict__ row_len, int group_rows, int dead_mode, int stage_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    using Cfg = G2F32Cfg<BN>;
    using TA = G2Tile<A_MN, kG2M, 4>;
    using TB = G2Tile<B_MN, BN, 4>;
    constexpr int kStages = Cfg::kStages;
    Gemm2Barriers* bars = reinterpret_cast<Gemm2Barriers*>(smem + kStages * Cfg::kStage);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * kG2M;
    const int nk_total = (K + TA::kBK - 1) / TA::kBK;          // a K tail is zero-filled by TMA
    const int k_first = blockIdx.z * stages_per_split;
    int nk = max(0, min(stages_per_split, nk_total - k_first));
    // split-K inside a cluster: partial tiles meet in shared memory.  stage_out: an unsplit tile takes the same way out
    // (accumulator -> shared memory -> row-contiguous stores) instead of one strided 16-byte store per thread and row
    const bool alone = cluster_nctarank() == 1;
    const bool clustered = !alone || (stage_out != 0 && gridDim.z == 1);
    // Ragged rows (padded utterances): the "row" dimension - M when A is K-major, the contraction when A is MN-major - is
    // made of groups of `group_rows` rows of which only the first row_len[g] are valid; the others are known to be zero
    // (gradient rows of padded frames) or never read (their logits).  A row tile / K step that lies entirely in padding
    // is skipped: dead_mode 1 = A K-major, dead tiles are not written at all; 2 = A K-major, dead tiles are written as
    // zeros; 3 = contraction over the rows, dead K steps contribute nothing.
    auto rows_dead = [&](int r0, int n) {       // rows [r0, r0 + n) all padding?
        const int g0 = r0 / group_rows, g1 = (r0 + n - 1) / group_rows;
        for (int g = g0; g <= g1; ++g) {
            const int lo = max(r0, g * group_rows) - g * group_rows;      // first row of the range inside group g
            if (lo < __ldg(row_len + g)) return false;
        }
        return true;
    };
    if (row_len != nullptr && (dead_mode == 1 || dead_mode == 2) && rows_dead(m0, min(kG2M, M - m0))) {
        if (dead_mode == 1) return;      // uniform for the CTA, before any barrier / allocation
        nk = 0;
    }
    const bool skip_k = row_len != nullptr && dead_mode == 3;
    auto stage_dead = [&](int k) { return skip_k && rows_dead((k_first + k) * TA::kBK, min(TA::kBK, K - (k_first + k) * TA::kBK)); };
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->split[s], 4);
            mbar_init(&bars->empty[s], 1);
        }
        mbar_init(&bars->acc_full, 1);
        fence_mbar_init();
    }
    // Two accumulators [128 x BN]: the head products hi_a hi_b go to the first, the two correction products to the second.
    // The tensor core's fp32 accumulation loses up to an ulp OF THE ACCUMULATOR per addition; kept apart, the main sum
    // takes one addition per K step instead of three and the corrections (2^-11 of the magnitude) round among themselves.
    if (warp == 5) {
        tmem_alloc(&bars->tmem_base, 2 * BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    if (warp == 4) {
        if (elect_one_sync()) {
            tma_prefetch_desc(&tm_a);
            tma_prefetch_desc(&tm_b);
            int it = 0;                                  // live stages only: every role counts them the same way
            for (int k = 0; k < nk; ++k) {
                if (stage_dead(k)) continue;
                const int s = it % kStages;
                if (it >= kStages) mbar_wait(&bars->empty[s], ((it / kStages) - 1) & 1);
                unsigned char* st = smem + s * Cfg::kStage;
                mbar_arrive_expect_tx(&bars->full[s], Cfg::kATile + Cfg::kBTile);
                const int kc = (k_first + k) * TA::kBK;
                TA::load(st, &tm_a, kc, m0, &bars->full[s]);
                TB::load(st + 2 * Cfg::kATile, &tm_b, kc, n0, &bars->full[s]);
                ++it;
            }
        }
    } else if (warp == 5) {
        if (elect_one_sync()) {
            constexpr uint32_t idesc = make_idesc_tf32(kG2M, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            int it = 0;
            for (int k = 0; k < nk; ++k) {
                if (stage_dead(k)) continue;
                const int s = it % kStages;
                mbar_wait(&bars->split[s], (it / kStages) & 1);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + s * Cfg::kStage);
                const uint32_t a_hi = base, a_lo = base + Cfg::kATile, b_hi = base + 2 * Cfg::kATile, b_lo = b_hi + Cfg::kBTile;
#pragma unroll
                for (int kk = 0; kk < TA::kBK / TA::kUmmaK; ++kk) {
                    const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
                    umma_tf32(tmem + BN, TA::desc(a_lo, kk), TB::desc(b_hi, kk), idesc, acc);
                    umma_tf32(tmem + BN, TA::desc(a_hi, kk), TB::desc(b_lo, kk), idesc, 1u);
                    umma_tf32(tmem, TA::desc(a_hi, kk), TB::desc(b_hi, kk), idesc, acc);
                }
                tc_commit(&bars->empty[s]);
                ++it;
            }
            if (it > 0) tc_commit(&bars->acc_full);
        }
    } else {
        // splitter: lo = x - (x with the 13 low mantissa bits cleared); the raw tile serves as the TF32 head
        int live = 0;
        for (int k = 0; k < nk; ++k) {
            if (stage_dead(k)) continue;
            const int s = live % kStages;
            mbar_wait(&bars->full[s], (live / kStages) & 1);
            unsigned char* st = smem + s * Cfg::kStage;
            split_lo_tile(st, st + Cfg::kATile, Cfg::kATile, threadIdx.x);
            split_lo_tile(st + 2 * Cfg::kATile, st + 2 * Cfg::kATile + Cfg::kBTile, Cfg::kBTile, threadIdx.x);
            fence_proxy_async();          // generic-proxy writes -> visible to the tensor core's operand fetch
            mbar_arrive_warp(&bars->split[s]);
            ++live;
        }
        nk = live;                        // the epilogue below: zeros when no stage was live
        // epilogue: thread = row of the tile (TMEM lane), 32 columns at a time
        const int row = m0 + warp * 32 + lane;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        float* dst = c + (size_t)blockIdx.z * split_stride + (size_t)min(row, M - 1) * ldc + n0;
        const bool add_bias = bias != nullptr && gridDim.z == 1;
        // float4 stores when rows are 16-byte aligned; a float4 that starts inside N may run into the row's padding
        // (ldc >= N rounded up to 4), where the accumulator holds exact zeros (TMA zero fill)
        const bool vec = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(c) & 15u) == 0;
        const int n_store = vec ? min(ldc, (N + 3) & ~3) : N;
        if (nk > 0) {
            mbar_wait(&bars->acc_full, 0);
            tc_fence_after();
        }
#pragma unroll 1
        for (int cc = 0; cc < BN; cc += 32) {
            if (n0 + cc >= n_store) break;
            float v[32];
            if (nk > 0) {
                float corr[32];
                tmem_ld32(tmem + lane_base + cc, v);
                tmem_ld32(tmem + BN + lane_base + cc, corr);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += corr[i];
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.0f;
            }
            if (clustered) {          // the operand ring is free: every MMA that read it has completed (acc_full)
                part_store32<BN>(smem, warp * 32 + lane, cc, v);
                continue;
            }
            if (vec) {
                // A thread owns a ROW of the tile: stored directly, each of its 16-byte pieces lands in another 128-byte line
                // (32 lines per instruction; at the 4233-wide projection the epilogue was ~30 % of a tile's time).  The warp
                // parks its 32 x 32 chunk in the operand ring (free: every MMA has completed) and writes it back four full
                // 128-byte rows per instruction.
                if (add_bias) {
                    if (n0 + cc + 32 <= N && (reinterpret_cast<uintptr_t>(bias) & 15u) == 0) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + cc + i));
                            v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] += (n0 + cc + i < N) ? __ldg(bias + n0 + cc + i) : 0.0f;
                    }
                }
                constexpr int kParkRow = 36;                       // floats per parked row: 32 + 4 of padding (conflict-free)
                float* park = reinterpret_cast<float*>(smem) + warp * 32 * kParkRow;
                float4* mine = reinterpret_cast<float4*>(park + lane * kParkRow);
#pragma unroll
                for (int i = 0; i < 8; ++i) mine[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                __syncwarp();
                const int piece = lane & 7;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int r = (lane >> 3) + 4 * j;
                    const int grow = m0 + warp * 32 + r, n = n0 + cc + piece * 4;
                    if (grow < M && n < n_store)
                        *reinterpret_cast<float4*>(c + (size_t)blockIdx.z * split_stride + (size_t)grow * ldc + n) =
                            *reinterpret_cast<const float4*>(park + r * kParkRow + piece * 4);
                }
                __syncwarp();
            } else if (row < M) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int n = n0 + cc + i;
                    if (n < N) dst[cc + i] = v[i] + (add_bias ? __ldg(bias + n) : 0.0f);
                }
            }
        }
        tc_fence_before();
    }
    if (clustered) {
        cluster_sync_all(alone);
        if (warp < 4) {
            const bool vec = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(c) & 15u) == 0;
            const int n_store = vec ? min(ldc, (N + 3) & ~3) : N;
            auto bias4 = [&](int, int col) {
                const int n = n0 + col;
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bias != nullptr) {
                    o.x = (n < N) ? __ldg(bias + n) : 0.0f;
                    o.y = (n + 1 < N) ? __ldg(bias + n + 1) : 0.0f;
                    o.z = (n + 2 < N) ? __ldg(bias + n + 2) : 0.0f;
                    o.w = (n + 3 < N) ? __ldg(bias + n + 3) : 0.0f;
                }
                return o;
            };
            auto emit = [&](int r, int col, const float4& o) {
                const int gr = m0 + r, n = n0 + col;
                if (gr >= M) return;
                float* out = c + (size_t)gr * ldc + n;
                if (vec) {
                    *reinterpret_cast<float4*>(out) = o;
                } else {
                    if (n < N) out[0] = o.x;
                    if (n + 1 < N) out[1] = o.y;
                    if (n + 2 < N) out[2] = o.z;
                    if (n + 3 < N) out[3] = o.w;
                }
            };
            cluster_splitk_reduce<BN>(smem, threadIdx.x, min(BN, n_store - n0), bias4, emit);
        }
        cluster_sync_all(alone, true);  // nobody leaves while a peer still reads its partial tile
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem, 2 * BN);
    }
}

// out[i] = sum_z part[z][i] (+ bias[i % N] on the valid columns): fixed order, deterministic
__global__ void __launch_bounds__(256) splitk_reduce_f32_kernel(const float* __restrict__ part, int splits, size_t split_stride, int M,
                                                               int N, int ldp, const float* __restrict__ bias, float* __restrict__ out,
                                                               int ldc) {
    const size_t total = (size_t)M * N;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int r = (int)(i / N), n = (int)(i - (size_t)r * N);
        float acc = bias != nullptr ? __ldg(bias + n) : 0.0f;
        for (int z = 0; z < splits; ++z) acc += part[(size_t)z * split_stride + (size_t)r * ldp + n];
        out[(size_t)r * ldc + n] = acc;
    }
}

// ------------------------------------------------------------------------------------------------------------
// bf16
// ------------------------------------------------------------------------------------------------------------
template <int BN>
struct G2B16Cfg {
    static constexpr int kStages = BN == 256 ? 2 : 3;
    static constexpr int kATile = kG2M * kG2RowBytes;
    static constexpr int kBTile = BN * kG2RowBytes;
    static constexpr int kStage = kATile + kBTile;
    static constexpr int kSmem = kStages * kStage + 256;
};

// EPI bit 0: ReLU; OUT_F32: fp32 output (weight gradients for fp32 master weights, split-K partials) instead of bf16
template <int BN, bool A_MN, bool B_MN, bool OUT_F32, bool RELU>
__global__ void __launch_bounds__(192, 2)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const float* __restrict__ bias,
                 void* __restrict__ c_, int M, int N, int K, int ldc, int stages_per_split, size_t split_stride, int stage_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    using Cfg = G2B16Cfg<BN>;
    using TA = G2Tile<A_MN, kG2M, 2>;
    using TB = G2Tile<B_MN, BN, 2>;
    constexpr int kStages = Cfg::kStages;
    Gemm2Barriers* bars = reinterpret_cast<Gemm2Barriers*>(smem + kStages * Cfg::kStage);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * kG2M;
    const int nk_total = (K + TA::kBK - 1) / TA::kBK;
    const int k_first = blockIdx.z * stages_per_split;
    const int nk = max(0, min(stages_per_split, nk_total - k_first));
    constexpr bool kCanCluster = G2Part<BN>::kBytes <= kStages * Cfg::kStage;      // the partial tile must fit in the operand ring
    const bool alone = cluster_nctarank() == 1;
    const bool clustered = kCanCluster && (!alone || (stage_out != 0 && gridDim.z == 1));
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        mbar_init(&bars->acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 5) {
        tmem_alloc(&bars->tmem_base, BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    if (warp == 4) {
        if (elect_one_sync()) {
            tma_prefetch_desc(&tm_a);
            tma_prefetch_desc(&tm_b);
            for (int k = 0; k < nk; ++k) {
                const int s = k % kStages;
                if (k >= kStages) mbar_wait(&bars->empty[s], ((k / kStages) - 1) & 1);
                unsigned char* st = smem + s * Cfg::kStage;
                mbar_arrive_expect_tx(&bars->full[s], Cfg::kATile + Cfg::kBTile);
                const int kc = (k_first + k) * TA::kBK;
                TA::load(st, &tm_a, kc, m0, &bars->full[s]);
                TB::load(st + Cfg::kATile, &tm_b, kc, n0, &bars->full[s]);
            }
        }
    } else if (warp == 5) {
        if (elect_one_sync()) {
            constexpr uint32_t idesc = make_idesc(kG2M, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            for (int k = 0; k < nk; ++k) {
                const int s = k % kStages;
                mbar_wait(&bars->full[s], (k / kStages) & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * Cfg::kStage);
                const uint32_t b_addr = a_addr + Cfg::kATile;
#pragma unroll
                for (int kk = 0; kk < TA::kBK / TA::kUmmaK; ++kk)
                    umma_bf16(tmem, TA::desc(a_addr, kk), TB::desc(b_addr, kk), idesc, (k > 0 || kk > 0) ? 1u : 0u);
                tc_commit(&bars->empty[s]);
            }
            tc_commit(&bars->acc_full);
        }
    } else {
        const int row = m0 + warp * 32 + lane;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        const bool add_bias = bias != nullptr && gridDim.z == 1;
        if (nk > 0) {
            mbar_wait(&bars->acc_full, 0);
            tc_fence_after();
        }
        const size_t row_off = (size_t)blockIdx.z * split_stride + (size_t)min(row, M - 1) * ldc + n0;
        float* dst_f = static_cast<float*>(c_) + row_off;
        __nv_bfloat16* dst_h = static_cast<__nv_bfloat16*>(c_) + row_off;
        constexpr int kVecElems = OUT_F32 ? 4 : 8;                       // elements per 16-byte store
        const bool vec = (ldc % kVecElems) == 0 && (reinterpret_cast<uintptr_t>(c_) & 15u) == 0;
        const int n_store = vec ? min(ldc, (N + kVecElems - 1) / kVecElems * kVecElems) : N;
#pragma unroll 1
        for (int cc = 0; cc < BN; cc += 32) {
            if (n0 + cc >= n_store) break;
            float v[32];
            if (nk > 0) {
                tmem_ld32(tmem + lane_base + cc, v);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.0f;
            }
            if (kCanCluster && clustered) {
                part_store32<BN>(smem, warp * 32 + lane, cc, v);
                continue;
            }
            if (row < M) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int n = n0 + cc + i;
                    float x = v[i] + ((add_bias && n < N) ? __ldg(bias + n) : 0.0f);
                    if (RELU) x = fmaxf(x, 0.0f);
                    v[i] = x;
                }
#pragma unroll
                for (int i = 0; i < 32; i += kVecElems) {
                    const int n = n0 + cc + i;
                    if (vec) {
                        if (n < n_store) {
                            if (OUT_F32) {
                                *reinterpret_cast<float4*>(dst_f + cc + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                            } else {
                                *reinterpret_cast<uint4*>(dst_h + cc + i) =
                                    make_uint4(cvt_bf16x2(v[i], v[i + 1]), cvt_bf16x2(v[i + 2], v[i + 3]),
                                               cvt_bf16x2(v[i + 4], v[i + 5]), cvt_bf16x2(v[i + 6], v[i + 7]));
                            }
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < kVecElems; ++u) {
                            if (n + u < N) {
                                if (OUT_F32) dst_f[cc + i + u] = v[i + u];
                                else dst_h[cc + i + u] = __float2bfloat16_rn(v[i + u]);
                            }
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    if (kCanCluster && clustered) {
        cluster_sync_all(alone);
        if (warp < 4) {
            constexpr int kAlign = OUT_F32 ? 16 : 8;             // four outputs per store
            const bool vec = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(c_) & (kAlign - 1)) == 0;
            const int n_store = vec ? min(ldc, (N + 3) & ~3) : N;
            auto bias4 = [&](int, int col) {
                const int n = n0 + col;
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bias != nullptr) {
                    o.x = (n < N) ? __ldg(bias + n) : 0.0f;
                    o.y = (n + 1 < N) ? __ldg(bias + n + 1) : 0.0f;
                    o.z = (n + 2 < N) ? __ldg(bias + n + 2) : 0.0f;
                    o.w = (n + 3 < N) ? __ldg(bias + n + 3) : 0.0f;
                }
                return o;
            };
            auto emit = [&](int r, int col, float4 o) {
                const int gr = m0 + r, n = n0 + col;
                if (gr >= M) return;
                if (RELU) { o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f); }
                const size_t at = (size_t)gr * ldc + n;
                if (OUT_F32) {
                    float* out = static_cast<float*>(c_) + at;
                    if (vec) {
                        *reinterpret_cast<float4*>(out) = o;
                    } else {
                        if (n < N) out[0] = o.x;
                        if (n + 1 < N) out[1] = o.y;
                        if (n + 2 < N) out[2] = o.z;
                        if (n + 3 < N) out[3] = o.w;
                    }
                } else {
                    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(c_) + at;
                    if (vec) {
                        *reinterpret_cast<uint2*>(out) = make_uint2(cvt_bf16x2(o.x, o.y), cvt_bf16x2(o.z, o.w));
                    } else {
                        if (n < N) out[0] = __float2bfloat16_rn(o.x);
                        if (n + 1 < N) out[1] = __float2bfloat16_rn(o.y);
                        if (n + 2 < N) out[2] = __float2bfloat16_rn(o.z);
                        if (n + 3 < N) out[3] = __float2bfloat16_rn(o.w);
                    }
                }
            };
            cluster_splitk_reduce<BN>(smem, threadIdx.x, min(BN, n_store - n0), bias4, emit);
        }
        cluster_sync_all(alone, true);
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem, BN);
    }
}

// ------------------------------------------------------------------------------------------------------------
// bf16, persistent: one CTA walks the output tiles  blockIdx.x, + gridDim.x, ...  (n fastest: neighbours share A rows in L2).
// At the model's shapes a tile is SHORT (K = 512: eight K steps, ~1.8 us of tensor work) and the one-tile kernel spends
// most of a CTA's life on set-up and tear-down (TMEM allocation, barrier initialisation, the first loads' latency, the
// epilogue with nothing behind it).  Here the roles keep running across tiles: the TMA ring never drains (the producer
// is already loading tile i + 1 while tile i's products run), the accumulator is double-buffered in tensor memory
// (2 x BN columns: the epilogue of tile i reads buffer i & 1 while the products of tile i + 1 fill the other one), and the
// fixed costs are paid once per CTA.  Two CTAs per SM as before (2 x 256 TMEM columns, 2 x 96 KB).
// ------------------------------------------------------------------------------------------------------------
struct __align__(8) Gemm2BarriersP {
    uint64_t full[4];
    uint64_t empty[4];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

template <int BN>
struct G2B16PCfg {      // 128-wide tiles: two CTAs per SM (2 x 256 TMEM columns, 2 x 96 KB); 256-wide: one CTA owns the SM (512 columns, 4 stages)
    static constexpr int kStages = BN == 256 ? 4 : 3;
    static constexpr int kCtasPerSm = BN == 256 ? 1 : 2;
    static constexpr int kOutRow = 80;                      // bytes per staged output row: 32 bf16 + 16 bytes of padding (conflict-free 16-byte accesses)
    static constexpr int kOutStage = 4 * 32 * kOutRow;      // one 32 x 32 bf16 chunk per epilogue warp
    static constexpr int kBiasStage = 2 * BN * 4;           // the tile's bias values, double-buffered over tiles
    static constexpr int kSmem = kStages * G2B16Cfg<BN>::kStage + 256 + kOutStage + kBiasStage;
};

template <int BN, bool A_MN, bool B_MN, bool OUT_F32, bool RELU>
__global__ void __launch_bounds__(192, G2B16PCfg<BN>::kCtasPerSm)
gemm_bf16_persistent_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                            const float* __restrict__ bias, void* __restrict__ c_, int M, int N, int K, int ldc, int tiles_n,
                            int n_tiles, int debug) {
    extern __shared__ __align__(1024) unsigned char smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    using Cfg = G2B16Cfg<BN>;
    using TA = G2Tile<A_MN, kG2M, 2>;
    using TB = G2Tile<B_MN, BN, 2>;
    constexpr int kStages = G2B16PCfg<BN>::kStages;
    Gemm2BarriersP* bars = reinterpret_cast<Gemm2BarriersP*>(smem + kStages * Cfg::kStage);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = (K + TA::kBK - 1) / TA::kBK;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->acc_full[b], 1);
            mbar_init(&bars->acc_empty[b], 4);        // one arrival per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 5) {
        tmem_alloc(&bars->tmem_base, 2 * BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    if (warp == 4) {
        if (elect_one_sync()) {
            tma_prefetch_desc(&tm_a);
            tma_prefetch_desc(&tm_b);
            int it = 0;                                   // K steps since the kernel started: the ring does not care about tiles
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * kG2M, n0 = (tile % tiles_n) * BN;
                for (int k = 0; k < nk; ++k, ++it) {
                    const int s = it % kStages;
                    if (it >= kStages) mbar_wait(&bars->empty[s], ((it / kStages) - 1) & 1);
                    unsigned char* st = smem + s * Cfg::kStage;
                    mbar_arrive_expect_tx(&bars->full[s], Cfg::kATile + Cfg::kBTile);
                    TA::load(st, &tm_a, k * TA::kBK, m0, &bars->full[s]);
                    TB::load(st + Cfg::kATile, &tm_b, k * TA::kBK, n0, &bars->full[s]);
                }
            }
        }
    } else if (warp == 5) {
        if (elect_one_sync()) {
            constexpr uint32_t idesc = make_idesc(kG2M, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            int it = 0, lt = 0;                           // lt: tiles of this CTA so far
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
                const int buf = lt & 1;
                if (lt >= 2) {                            // the epilogue must have emptied this buffer (tile lt - 2)
                    mbar_wait(&bars->acc_empty[buf], ((lt >> 1) - 1) & 1);
                    tc_fence_after();
                }
                const uint32_t acc = tmem + (uint32_t)(buf * BN);
                for (int k = 0; k < nk; ++k, ++it) {
                    const int s = it % kStages;
                    mbar_wait(&bars->full[s], (it / kStages) & 1);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * Cfg::kStage);
                    const uint32_t b_addr = a_addr + Cfg::kATile;
#pragma unroll
                    for (int kk = 0; kk < TA::kBK / TA::kUmmaK; ++kk)
                        umma_bf16(acc, TA::desc(a_addr, kk), TB::desc(b_addr, kk), idesc, (k > 0 || kk > 0) ? 1u : 0u);
                    tc_commit(&bars->empty[s]);
                }
                tc_commit(&bars->acc_full[buf]);
            }
        }
    } else {
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        constexpr int kVecElems = OUT_F32 ? 4 : 8;                       // elements per 16-byte store
        const bool vec = (ldc % kVecElems) == 0 && (reinterpret_cast<uintptr_t>(c_) & 15u) == 0;
        const int n_store = vec ? min(ldc, (N + kVecElems - 1) / kVecElems * kVecElems) : N;
        int lt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
            const int buf = lt & 1;
            const int m0 = (tile / tiles_n) * kG2M, n0 = (tile % tiles_n) * BN;
            const int row = m0 + warp * 32 + lane;
            // the tile's bias values go to shared memory while the products are still running (one word per thread and
            // 128 columns), read back below as broadcast 16-byte loads; two buffers, one 128-thread barrier per tile
            float* bias_s = reinterpret_cast<float*>(smem + kStages * Cfg::kStage + 256 + G2B16PCfg<BN>::kOutStage) + buf * BN;
            if (bias != nullptr) {
                for (int i = threadIdx.x; i < BN; i += 128) bias_s[i] = (n0 + i < N) ? __ldg(bias + n0 + i) : 0.0f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            mbar_wait(&bars->acc_full[buf], (lt >> 1) & 1);
            tc_fence_after();
            const size_t row_off = (size_t)min(row, M - 1) * ldc + n0;
            float* dst_f = static_cast<float*>(c_) + row_off;
            __nv_bfloat16* dst_h = static_cast<__nv_bfloat16*>(c_) + row_off;
            const uint32_t acc = tmem + (uint32_t)(buf * BN) + lane_base;
#pragma unroll 1
            for (int cc = 0; cc < BN; cc += 32) {
                if (n0 + cc >= n_store) break;
                float v[32];
                tmem_ld32(acc + cc, v);
                if (cc + 32 >= BN || n0 + cc + 32 >= n_store) {       // last read of this buffer: hand it back before the stores
                    tc_fence_before();
                    mbar_arrive_warp(&bars->acc_empty[buf]);
                }
                if (debug & 1) continue;
                const bool full_chunk = n0 + cc + 32 <= N;
                if (bias != nullptr) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + cc + i);
                        v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                    }
                }
                if (RELU) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
                }
                if (!OUT_F32 && vec && full_chunk) {
                    // bf16 rows leave through shared memory: a thread owns a ROW of the tile, and stored directly its 16-byte
                    // pieces land in 32 different 128-byte lines per instruction (measured: bias + stores were 45 % of the
                    // kernel).  The warp parks its 32 x 32 chunk (64 bytes per row) and writes it back eight rows per
                    // instruction, four lanes per row.
                    unsigned char* park = smem + kStages * Cfg::kStage + 256 + warp * 32 * G2B16PCfg<BN>::kOutRow;
                    uint4* mine = reinterpret_cast<uint4*>(park + lane * G2B16PCfg<BN>::kOutRow);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        mine[i] = make_uint4(cvt_bf16x2(v[8 * i], v[8 * i + 1]), cvt_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                             cvt_bf16x2(v[8 * i + 4], v[8 * i + 5]), cvt_bf16x2(v[8 * i + 6], v[8 * i + 7]));
                    __syncwarp();
                    const int piece = lane & 3;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = (lane >> 2) + 8 * j;
                        const int grow = m0 + warp * 32 + r;
                        if (grow < M)
                            *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(c_) + (size_t)grow * ldc + n0 + cc + piece * 8) =
                                *reinterpret_cast<const uint4*>(park + r * G2B16PCfg<BN>::kOutRow + piece * 16);
                    }
                    __syncwarp();
                } else if (row < M) {
#pragma unroll
                    for (int i = 0; i < 32; i += kVecElems) {
                        const int n = n0 + cc + i;
                        if (vec) {
                            if (n < n_store) {
                                if (OUT_F32) {
                                    *reinterpret_cast<float4*>(dst_f + cc + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                                } else {
                                    *reinterpret_cast<uint4*>(dst_h + cc + i) =
                                        make_uint4(cvt_bf16x2(v[i], v[i + 1]), cvt_bf16x2(v[i + 2], v[i + 3]),
                                                   cvt_bf16x2(v[i + 4], v[i + 5]), cvt_bf16x2(v[i + 6], v[i + 7]));
                                }
                            }
                        } else {
#pragma unroll
                            for (int u = 0; u < kVecElems; ++u) {
                                if (n + u < N) {
                                    if (OUT_F32) dst_f[cc + i + u] = v[i + u];
                                    else dst_h[cc + i + u] = __float2bfloat16_rn(v[i + u]);
                                }
                            }
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem, 2 * BN);
    }
}

// partial fp32 tiles -> bf16 or fp32 output (+ bias, ReLU)
__global__ void __launch_bounds__(256) splitk_reduce_out_kernel(const float* __restrict__ part, int splits, size_t split_stride, int M,
                                                               int N, int ldp, const float* __restrict__ bias, int relu, void* out_,
                                                               int ldc, int out_f32) {
    const size_t total = (size_t)M * N;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int r = (int)(i / N), n = (int)(i - (size_t)r * N);
        float acc = bias != nullptr ? __ldg(bias + n) : 0.0f;
        for (int z = 0; z < splits; ++z) acc += part[(size_t)z * split_stride + (size_t)r * ldp + n];
        if (relu) acc = fmaxf(acc, 0.0f);
        if (out_f32) static_cast<float*>(out_)[(size_t)r * ldc + n] = acc;
        else static_cast<__nv_bfloat16*>(out_)[(size_t)r * ldc + n] = __float2bfloat16_rn(acc);
    }
}

// Column sums (bias gradients).  CTA = kCols columns x one chunk of the rows: 32 row groups of 32 lanes (1024 threads)
// read 128 contiguous bytes per row - 32 fp32 columns, or 64 bf16 columns as pairs when the rows are 4-byte aligned -
// eight rows in flight per thread; the 32 partial sums of a column are added in a fixed order.  The row chunks of one
// column block are a thread-block cluster along y (up to 8 CTAs: a [15030 x 512] gradient has only 16 column blocks,
// which left 130 SMs idle): each leaves its column sums in shared memory and CTA 0 of the cluster adds them in chunk
// order through distributed shared memory - one launch, no scratch buffer, deterministic.
template <typename T, int V>      // V = columns per lane: 1, or 2 (bf16 pairs)
__global__ void __launch_bounds__(1024) colsum_kernel(const T* __restrict__ x, int M, int N, int ld, int rows_per_cta, float* __restrict__ out) {
    constexpr int kCols = 32 * V;
    __shared__ float part[32][kCols + 1];
    __shared__ __align__(16) float csum[kCols];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int n = blockIdx.x * kCols + V * lane;
    const int m_lo = blockIdx.y * rows_per_cta, m_hi = min(M, m_lo + rows_per_cta);
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.0f;
    auto load = [&](int m, float (&dst)[V]) {
        if (V == 2) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + (size_t)m * ld + n));
            dst[0] = f.x;
            dst[V - 1] = f.y;
        } else {
            dst[0] = (float)x[(size_t)m * ld + n];
        }
    };
    if (n < N) {         // V == 2: N is even, so a pair is inside or outside as a whole
        int m = m_lo + grp;
        for (; m + 7 * 32 < m_hi; m += 8 * 32) {
            float a[8][V];
#pragma unroll
            for (int u = 0; u < 8; ++u) load(m + 32 * u, a[u]);
#pragma unroll
            for (int v = 0; v < V; ++v)
                acc[v] += ((a[0][v] + a[1][v]) + (a[2][v] + a[3][v])) + ((a[4][v] + a[5][v]) + (a[6][v] + a[7][v]));
        }
        for (; m < m_hi; m += 32) {
            float a[V];
            load(m, a);
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] += a[v];
        }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) part[grp][V * lane + v] = acc[v];
    __syncthreads();
    if (threadIdx.x < kCols) {
        float t = 0.0f;
#pragma unroll
        for (int g = 0; g < 32; ++g) t += part[g][threadIdx.x];
        csum[threadIdx.x] = t;
    }
    const bool alone = cluster_nctarank() == 1;
    cluster_sync_all(alone);
    if (cluster_ctarank() == 0 && threadIdx.x < kCols && blockIdx.x * kCols + threadIdx.x < N) {
        const uint32_t addr = smem_u32(&csum[threadIdx.x]);
        const int chunks = (int)cluster_nctarank();
        float t = 0.0f;
        for (int r = 0; r < chunks; ++r) {
            uint32_t remote;
            float v;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(r));
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
            t += v;
        }
        out[blockIdx.x * kCols + threadIdx.x] = t;
    }
    cluster_sync_all(alone, true);    // the chunks' shared memory stays until CTA 0 has read it
}

}  // namespace asr

using namespace asr;

extern "C" int asr_colsum(const void* x, int is_bf16, int M, int N, int ld, float* out, void* stream) {
    ASR_REQUIRE(x && out && M > 0 && N > 0 && ld >= N, "asr_colsum: bad arguments");
    if (asr_device_ok() != 0) return 3;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool pairs = is_bf16 && (N % 2) == 0 && (ld % 2) == 0 && (reinterpret_cast<uintptr_t>(x) & 3u) == 0;
    const int cols = pairs ? 64 : 32;
    const int col_blocks = (N + cols - 1) / cols;
    // row chunks: enough CTAs for two per SM's worth of loads in flight, at least 1024 rows each (below that the cluster
    // launch costs more than the idle SMs: 3.4 -> 4.4 us at M = 1344), at most one cluster (8)
    int chunks = std::min(kG2MaxSplits, std::max(1, (2 * num_sms() + col_blocks - 1) / col_blocks));
    chunks = std::max(1, std::min(chunks, M / 1024));
    const int rows_per_cta = (M + chunks - 1) / chunks;
    chunks = (M + rows_per_cta - 1) / rows_per_cta;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)col_blocks, (unsigned)chunks, 1);
    cfg.blockDim = dim3(1024);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = (unsigned)chunks;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = chunks > 1 ? 1 : 0;
    if (pairs)
        ASR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, colsum_kernel<__nv_bfloat16, 2>, static_cast<const __nv_bfloat16*>(x), M, N, ld, rows_per_cta, out));
    else if (is_bf16)
        ASR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, colsum_kernel<__nv_bfloat16, 1>, static_cast<const __nv_bfloat16*>(x), M, N, ld, rows_per_cta, out));
    else
        ASR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, colsum_kernel<float, 1>, static_cast<const float*>(x), M, N, ld, rows_per_cta, out));
    ASR_LAUNCH_CHECK();
    return 0;
}

// ---- host side --------------------------------------------------------------------------------------------
namespace {

struct Plan {
    int bn, splits, stages_per_split, nk_total;
    dim3 grid;
};

// Tile width and split-K are chosen by a small cost model, fitted to CUDA-graph replays of the model's shapes on a B200
// (tools/gemm_split_bench.py --sweep; microseconds): a CTA costs  steps x t_step(BN) + t_tile(BN)  (+ the cluster
// reduction when it shares its tile), and the launch takes as many such rounds as the tiles need at the number of
// CTAs - or whole clusters - the GPU holds at once.  That last number is NOT num_sms / splits: a cluster lives inside
// one GPC, so e.g. clusters of 5 one-CTA-per-SM kernels leave SMs of every GPC idle (28 tiles x 5 took two rounds);
// cudaOccupancyMaxActiveClusters knows, and is asked once per (kernel, BN, cluster size).
struct G2Cost {
    float t_step[2];          // one K step (128 bytes of K) of a 128 x BN tile, BN = 128 / 256
    float t_tile[2];          // accumulator -> global memory, barriers, TMEM allocation
    float t_cluster;          // a tile shared by a cluster: two cluster barriers + the reduction, less what the row-contiguous stores save
    int ctas_per_sm;
    bool cluster_256;         // may a 256-wide tile be split inside a cluster (its partial tile must fit in the operand ring)
};
constexpr G2Cost kCostF32 = {{0.87f, 1.28f}, {2.4f, 5.0f}, 0.3f, 1, true};
constexpr G2Cost kCostB16 = {{0.225f, 0.43f}, {2.0f, 4.5f}, 2.3f, 2, false};

template <typename KernelT>
int query_max_clusters(KernelT kernel, int smem, int cluster_z) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1, 1, (unsigned)cluster_z);
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = (size_t)smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = (unsigned)cluster_z;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
        cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// clusters of `splits` CTAs (one per output tile) the GPU holds at once; 0 = such a cluster cannot be launched
int max_clusters(bool f32, int bn, int splits);      // defined below the kernels' launchers (needs their instances)

Plan make_plan(bool f32, int M, int N, int K, int bk, int force_bn, int force_splits, bool allow_split, bool in_cluster) {
    const G2Cost& cm = f32 ? kCostF32 : kCostB16;
    const long long mt = (M + kG2M - 1) / kG2M;
    const int nk_total = (K + bk - 1) / bk;
    Plan best = {};
    float best_cost = -1.0f;
    for (int bi = 0; bi < 2; ++bi) {
        const int bn = bi ? 256 : 128;
        if ((force_bn == 128 || force_bn == 256) && bn != force_bn) continue;
        if (bn == 256 && N <= 128 && force_bn != 256) continue;
        const long long tiles = mt * ((N + bn - 1) / bn);
        for (int want = 1; want <= kG2MaxSplits; ++want) {
            if (force_splits > 0 && want != std::max(1, std::min(force_splits, std::min(nk_total, kG2MaxSplits)))) continue;
            if (want > 1 && !allow_split) break;
            const int steps = (nk_total + want - 1) / want;
            const int splits = (nk_total + steps - 1) / steps;
            if (splits != want && force_splits <= 0) continue;            // the same plan as a smaller `want`
            if (splits > 1 && steps < 4 && force_splits <= 0) continue;   // too little work per CTA to pay for the hand-over
            const bool clustered = splits > 1 && in_cluster && (bn == 128 || cm.cluster_256);
            long long room;                                               // tiles in flight
            if (clustered) {
                room = max_clusters(f32, bn, splits);
                if (room <= 0) continue;
            } else {
                room = std::max<long long>(1, (long long)num_sms() * cm.ctas_per_sm / splits);
            }
            const long long rounds = (tiles + room - 1) / room;
            float cost = (float)rounds * ((float)steps * cm.t_step[bi] + cm.t_tile[bi] + (splits > 1 ? cm.t_cluster + 0.1f * splits : 0.0f));
            if (splits > 1 && !clustered) cost += 4.5f;                   // partial tiles through the workspace + the reduce kernel
            if (best_cost < 0.0f || cost < best_cost) {
                best_cost = cost;
                best.bn = bn;
                best.splits = splits;
                best.stages_per_split = steps;
                best.nk_total = nk_total;
                best.grid = dim3((unsigned)((N + bn - 1) / bn), (unsigned)mt, (unsigned)splits);
            }
        }
    }
    if (best_cost < 0.0f) {      // nothing admissible (cannot happen with splits = 1 allowed): one CTA per 128 x 128 tile
        best.bn = 128;
        best.splits = 1;
        best.stages_per_split = nk_total;
        best.nk_total = nk_total;
        best.grid = dim3((unsigned)((N + 127) / 128), (unsigned)mt, 1u);
    }
    return best;
}

// tensor map of one operand: K-major = [mn rows x k cols] (box ROWS x chunk), MN-major = [k rows x mn cols] (box chunk-k x chunk)
int make_operand_map(CUtensorMap* map, CUtensorMapDataType dt, int elem, const void* base, bool mn_major, int mn, int k, int ld,
                     int tile_rows) {
    const int chunk = kG2RowBytes / elem;
    if (!mn_major)
        return make_tmap_2d(map, dt, elem, base, (uint64_t)mn, (uint64_t)k, (uint64_t)ld * elem, (uint32_t)std::min(tile_rows, 256),
                            (uint32_t)chunk, CU_TENSOR_MAP_SWIZZLE_128B);
    return make_tmap_2d(map, dt, elem, base, (uint64_t)k, (uint64_t)mn, (uint64_t)ld * elem, (uint32_t)chunk, (uint32_t)chunk,
                        elem == 4 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
}

template <typename KernelT>
int set_smem_once(KernelT kernel, int bytes, bool& done) {
    if (!done) {
        ASR_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        done = true;
    }
    return 0;
}


// <<<grid, 192, smem, st>>> with the gridDim.z CTAs of an output tile as one cluster when cluster_z > 1
template <typename KernelT, typename... Args>
cudaError_t launch_g2(KernelT kernel, dim3 grid, int smem, cudaStream_t st, int cluster_z, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = (unsigned)cluster_z;
    cfg.attrs = attr;
    cfg.numAttrs = cluster_z > 1 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}


int max_clusters(bool f32, int bn, int splits) {
    static int cache[2][2][kG2MaxSplits + 1];      // 0 = not asked yet, -1 = cannot be launched
    if (splits < 2 || splits > kG2MaxSplits) return 0;
    int& slot = cache[f32 ? 1 : 0][bn == 256 ? 1 : 0][splits];
    if (slot == 0) {
        int n;
        if (f32)
            n = bn == 256 ? query_max_clusters(gemm_f32x3_kernel<256, false, false>, G2F32Cfg<256>::kSmem, splits)
                          : query_max_clusters(gemm_f32x3_kernel<128, false, false>, G2F32Cfg<128>::kSmem, splits);
        else
            n = bn == 256 ? query_max_clusters(gemm_bf16_kernel<256, false, false, true, false>, G2B16Cfg<256>::kSmem, splits)
                          : query_max_clusters(gemm_bf16_kernel<128, false, false, true, false>, G2B16Cfg<128>::kSmem, splits);
        slot = n > 0 ? n : -1;
    }
    return slot > 0 ? slot : 0;
}

}  // namespace

extern "C" size_t asr_gemm_workspace_bytes(int M, int N, int K) {
    (void)K;
    if (M <= 0 || N <= 0) return 0;
    return (size_t)8 * M * N * sizeof(float) + 256;      // up to 8 split-K partial tiles
}

extern "C" int asr_gemm_f32(const float* a, int a_mn_major, int lda, const float* b, int b_mn_major, int ldb, const float* bias,
                            int M, int N, int K, float* c, int ldc, void* ws, size_t ws_bytes, void* stream) {
    return asr_gemm_f32_ragged(a, a_mn_major, lda, b, b_mn_major, ldb, bias, M, N, K, c, ldc, nullptr, 0, 0, ws, ws_bytes, stream);
}

extern "C" int asr_gemm_f32_ragged(const float* a, int a_mn_major, int lda, const float* b, int b_mn_major, int ldb,
                                   const float* bias, int M, int N, int K, float* c, int ldc, const int* row_len, int group_rows,
                                   int skip_dead_output, void* ws, size_t ws_bytes, void* stream) {
    ASR_REQUIRE(a && b && c, "asr_gemm_f32: null pointer");
    int dead_mode = 0;
    if (row_len != nullptr) {
        ASR_REQUIRE(group_rows > 0, "asr_gemm_f32_ragged: group_rows must be positive");
        ASR_REQUIRE(!a_mn_major || b_mn_major, "asr_gemm_f32_ragged: with the rows as the contraction, both operands must be MN-major");
        ASR_REQUIRE((a_mn_major ? K : M) % group_rows == 0, "asr_gemm_f32_ragged: the row count must be a multiple of group_rows");
        dead_mode = a_mn_major ? 3 : (skip_dead_output ? 1 : 2);
    }
    ASR_REQUIRE(M > 0 && N > 0 && K > 0, "asr_gemm_f32: bad shape M=%d N=%d K=%d", M, N, K);
    ASR_REQUIRE(lda % 4 == 0 && ldb % 4 == 0, "asr_gemm_f32: operand row strides (%d, %d) must be multiples of 4 floats (16-byte rows for TMA)", lda, ldb);
    ASR_REQUIRE(lda >= (a_mn_major ? M : K) && ldb >= (b_mn_major ? N : K) && ldc >= N, "asr_gemm_f32: row stride smaller than the row");
    ASR_REQUIRE(aligned16(a) && aligned16(b), "asr_gemm_f32: operands must be 16-byte aligned");
    ASR_REQUIRE((long long)(M + kG2M - 1) / kG2M <= 65535, "asr_gemm_f32: M=%d exceeds the grid limit", M);
    if (asr_device_ok() != 0) return 3;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool can_split = ws != nullptr && ws_bytes >= asr_gemm_workspace_bytes(M, N, K);
    const bool cluster_mode = get_opt("gemm_split_mode") != 1;
    const Plan p = make_plan(true, M, N, K, 32, get_opt("gemm_f32_bn"), get_opt("gemm_split_k"), can_split, cluster_mode);
    CUtensorMap ta, tb;
    if (make_operand_map(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a, a_mn_major != 0, M, K, lda, kG2M) ||
        make_operand_map(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, b, b_mn_major != 0, N, K, ldb, p.bn))
        return 4;
    // split-K: inside a cluster (default), or through the workspace + a reduce kernel (option gemm_split_mode = 1)
    const bool in_cluster = p.splits > 1 && cluster_mode;
    const bool via_ws = p.splits > 1 && !in_cluster;
    // rows the epilogue cannot store 16 bytes at a time (an odd row stride: the 4233-wide vocabulary) leave through shared
    // memory, where a warp writes one contiguous row segment instead of 32 strided scalars (measured: 51 -> 39 us at
    // M = 896, N = 4233); option gemm_stage_out: 1 = never, 2 = always
    const bool direct_vec = (ldc & 3) == 0 && aligned16(c);
    const int stage_out = get_opt("gemm_stage_out") == 2 || (get_opt("gemm_stage_out") == 0 && !direct_vec) ? 1 : 0;
    float* part = nullptr;
    if (via_ws) part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    float* dst = via_ws ? part : c;
    const int ldd = via_ws ? N : ldc;
    const size_t split_stride = via_ws ? (size_t)M * N : 0;
#define ASR_G2F(BNV, AM, BM)                                                                                        \
    do {                                                                                                            \
        static bool done = false;                                                                                   \
        if (set_smem_once(gemm_f32x3_kernel<BNV, AM, BM>, G2F32Cfg<BNV>::kSmem, done)) return 1;                     \
        ASR_CHECK_CUDA(launch_g2(gemm_f32x3_kernel<BNV, AM, BM>, p.grid, G2F32Cfg<BNV>::kSmem, st,                   \
                                 in_cluster ? p.splits : 1, ta, tb, bias, dst, M, N, K, ldd, p.stages_per_split,    \
                                 split_stride, row_len, group_rows, dead_mode, stage_out));                        \
    } while (0)
    const int sel = (p.bn == 256 ? 4 : 0) | (a_mn_major ? 2 : 0) | (b_mn_major ? 1 : 0);
    switch (sel) {
        case 0: ASR_G2F(128, false, false); break;
        case 1: ASR_G2F(128, false, true); break;
        case 2: ASR_G2F(128, true, false); break;
        case 3: ASR_G2F(128, true, true); break;
        case 4: ASR_G2F(256, false, false); break;
        case 5: ASR_G2F(256, false, true); break;
        case 6: ASR_G2F(256, true, false); break;
        default: ASR_G2F(256, true, true); break;
    }
#undef ASR_G2F
    ASR_LAUNCH_CHECK();
    if (via_ws) {
        const size_t total = (size_t)M * N;
        const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)num_sms() * 8);
        splitk_reduce_f32_kernel<<<blocks, 256, 0, st>>>(part, p.splits, split_stride, M, N, N, bias, c, ldc);
        ASR_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int asr_gemm_bf16(const void* a, int a_mn_major, int lda, const void* b, int b_mn_major, int ldb, const float* bias,
                             int relu, int M, int N, int K, void* c, int ldc, int out_f32, void* ws, size_t ws_bytes, void* stream) {
    ASR_REQUIRE(a && b && c, "asr_gemm_bf16: null pointer");
    ASR_REQUIRE(M > 0 && N > 0 && K > 0, "asr_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
    ASR_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "asr_gemm_bf16: operand row strides (%d, %d) must be multiples of 8 elements (16-byte rows for TMA)", lda, ldb);
    ASR_REQUIRE(lda >= (a_mn_major ? M : K) && ldb >= (b_mn_major ? N : K) && ldc >= N, "asr_gemm_bf16: row stride smaller than the row");
    ASR_REQUIRE(aligned16(a) && aligned16(b), "asr_gemm_bf16: operands must be 16-byte aligned");
    ASR_REQUIRE((long long)(M + kG2M - 1) / kG2M <= 65535, "asr_gemm_bf16: M=%d exceeds the grid limit", M);
    ASR_REQUIRE(!(relu && out_f32), "asr_gemm_bf16: ReLU is fused for bf16 outputs only");
    if (asr_device_ok() != 0) return 3;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool can_split = ws != nullptr && ws_bytes >= asr_gemm_workspace_bytes(M, N, K);
    const bool cluster_mode = get_opt("gemm_split_mode") != 1;
    const int force_bn = get_opt("gemm_variant") == 1 ? 128 : (get_opt("gemm_variant") == 2 ? 256 : 0);
    Plan p = make_plan(false, M, N, K, 64, force_bn, get_opt("gemm_split_k"), can_split, cluster_mode);
    // Unsplit products with more 128-wide tiles than two CTAs per SM hold (the encoder's M = 15030 rows) take the persistent
    // kernel; its tile width by measurement (CUDA-graph replays, us, 128 / 256 wide): N = K = 512: 12.2 / 12.4, N = 2048:
    // 35.9 / 33.7, K = 2048: 30.3 / 27.6 - 256 wide (one CTA per SM, four stages) once N or K reaches 1024 and there is a tile
    // for every SM.
    bool persistent = false;
    {
        const int pers_opt = get_opt("gemm_persistent");
        const long long mt = (M + kG2M - 1) / kG2M;
        const long long tiles128 = mt * ((N + 127) / 128), tiles256 = mt * ((N + 255) / 256);
        if (p.splits == 1 && pers_opt != 1) {
            if (pers_opt == 2) {
                persistent = true;
            } else if (tiles128 > 2LL * num_sms()) {
                persistent = true;
                if (force_bn == 0) p.bn = ((N >= 1024 || K >= 1024) && tiles256 >= num_sms()) ? 256 : 128;
            }
            if (persistent) p.grid = dim3((unsigned)((N + p.bn - 1) / p.bn), (unsigned)mt, 1u);
        }
    }
    CUtensorMap ta, tb;
    if (make_operand_map(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, a_mn_major != 0, M, K, lda, kG2M) ||
        make_operand_map(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, b, b_mn_major != 0, N, K, ldb, p.bn))
        return 4;
    // 256-wide tiles: the partial tile does not fit in the bf16 kernel's operand ring - those go through the workspace
    const bool in_cluster = p.splits > 1 && p.bn == 128 && cluster_mode;
    // as in asr_gemm_f32: outputs whose rows cannot take 16-byte stores go out through shared memory (bf16, M = 896,
    // N = 4233: 50 -> 16 us)
    const bool direct_vec = (ldc % (out_f32 ? 4 : 8)) == 0 && aligned16(c);
    const int stage_out = get_opt("gemm_stage_out") == 2 || (get_opt("gemm_stage_out") == 0 && !direct_vec) ? 1 : 0;
    const bool split = p.splits > 1 && !in_cluster;
    float* part = split ? reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255) : nullptr;
    void* dst = split ? static_cast<void*>(part) : c;
    const int ldd = split ? N : ldc;
    const size_t split_stride = (size_t)M * N;
    const bool f32 = split || out_f32 != 0;
    const bool rl = relu != 0 && !split;
    // the persistent kernel (decided above; option gemm_persistent: 1 = never, 2 = whenever unsplit)
    const long long tiles = (long long)p.grid.x * p.grid.y;
    const int resident = (p.bn == 256 ? 1 : 2) * num_sms();
    if (persistent && stage_out == 0 && tiles < (1LL << 30)) {
        const int tiles_n = (int)p.grid.x, n_tiles = (int)tiles;
        const dim3 grid((unsigned)std::min<long long>(tiles, resident));
#define ASR_G2P(BNV, AM, BM, F32, RL)                                                                                       \
    do {                                                                                                                    \
        static bool done = false;                                                                                           \
        if (set_smem_once(gemm_bf16_persistent_kernel<BNV, AM, BM, F32, RL>, G2B16PCfg<BNV>::kSmem, done)) return 1;          \
        gemm_bf16_persistent_kernel<BNV, AM, BM, F32, RL><<<grid, 192, G2B16PCfg<BNV>::kSmem, st>>>(ta, tb, bias, c, M, N, K, \
                                                                                                   ldc, tiles_n, n_tiles, get_opt("gemm_debug"));  \
    } while (0)
#define ASR_G2P_LAYOUT(BNV, F32, RL)                                  \
    do {                                                              \
        if (!a_mn_major && !b_mn_major) ASR_G2P(BNV, false, false, F32, RL); \
        else if (!a_mn_major) ASR_G2P(BNV, false, true, F32, RL);     \
        else if (!b_mn_major) ASR_G2P(BNV, true, false, F32, RL);     \
        else ASR_G2P(BNV, true, true, F32, RL);                       \
    } while (0)
        if (p.bn == 256) {
            if (out_f32) ASR_G2P_LAYOUT(256, true, false);
            else if (relu) ASR_G2P_LAYOUT(256, false, true);
            else ASR_G2P_LAYOUT(256, false, false);
        } else {
            if (out_f32) ASR_G2P_LAYOUT(128, true, false);
            else if (relu) ASR_G2P_LAYOUT(128, false, true);
            else ASR_G2P_LAYOUT(128, false, false);
        }
#undef ASR_G2P_LAYOUT
#undef ASR_G2P
        ASR_LAUNCH_CHECK();
        return 0;
    }
#define ASR_G2H(BNV, AM, BM, F32, RL)                                                                                     \
    do {                                                                                                                  \
        static bool done = false;                                                                                         \
        if (set_smem_once(gemm_bf16_kernel<BNV, AM, BM, F32, RL>, G2B16Cfg<BNV>::kSmem, done)) return 1;                    \
        ASR_CHECK_CUDA(launch_g2(gemm_bf16_kernel<BNV, AM, BM, F32, RL>, p.grid, G2B16Cfg<BNV>::kSmem, st,                 \
                                 in_cluster ? p.splits : 1, ta, tb, bias, dst, M, N, K, ldd, p.stages_per_split,          \
                                 split_stride, stage_out));                                                              \
    } while (0)
#define ASR_G2H_LAYOUT(BNV, F32, RL)                                    \
    do {                                                                \
        if (!a_mn_major && !b_mn_major) ASR_G2H(BNV, false, false, F32, RL); \
        else if (!a_mn_major) ASR_G2H(BNV, false, true, F32, RL);        \
        else if (!b_mn_major) ASR_G2H(BNV, true, false, F32, RL);        \
        else ASR_G2H(BNV, true, true, F32, RL);                          \
    } while (0)
    if (p.bn == 256) {
        if (f32) ASR_G2H_LAYOUT(256, true, false);
        else if (rl) ASR_G2H_LAYOUT(256, false, true);
        else ASR_G2H_LAYOUT(256, false, false);
    } else {
        if (f32) ASR_G2H_LAYOUT(128, true, false);
        else if (rl) ASR_G2H_LAYOUT(128, false, true);
        else ASR_G2H_LAYOUT(128, false, false);
    }
#undef ASR_G2H_LAYOUT
#undef ASR_G2H
    ASR_LAUNCH_CHECK();
    if (split) {
        const size_t total = (size_t)M * N;
        const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)num_sms() * 8);
        splitk_reduce_out_kernel<<<blocks, 256, 0, st>>>(part, p.splits, split_stride, M, N, N, bias, relu, c, ldc, out_f32);
        ASR_LAUNCH_CHECK();
    }
    return 0;
}

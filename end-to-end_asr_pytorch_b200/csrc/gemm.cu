// gemm.cu - the linear layers around the attention core as tcgen05 GEMMs with fused epilogues.
//
//   asr_linear_act_bf16                : y = act(x W^T + b)                       act = identity | ReLU
//       w_1 of PositionwiseFeedForward (/root/reference/src/transformer/module.py:35-53: relu(w_1(x))) and the
//       q / k / v projections of MultiheadAttention (attention.py:40-45).  Since round 2 this entry point (and
//       asr_linear_f32) runs the general GEMMs of gemm2.cu; only the LayerNorm kernel lives here.
//   asr_linear_residual_layernorm_bf16 : y = LayerNorm(x W^T + b + residual) * gamma + beta,  N = d_model = 512
//       w_2 + residual + layer_norm of the feed-forward block (module.py:50-52) and fc + residual + layer_norm of
//       the attention block (attention.py:59-60), dropout off (evaluation, or p = 0).
//
// x [M,K] bf16 row-major, W [N,K] bf16 row-major (torch's Linear.weight: K-major, exactly what the tensor core
// wants for B), fp32 accumulation in tensor memory, fp32 epilogue, bf16 output.  Both operands arrive as
// [rows x 64] TMA tiles with the 128-byte swizzle (a 3- or 2-deep ring), one elected thread issues
// tcgen05.mma (M128, N = 128 or 256, K16 per instruction), four epilogue warps read the accumulator with
// tcgen05.ld: thread = output row.  The LayerNorm kernel keeps the WHOLE 512-wide row of a 128-row tile in
// tensor memory (all 512 columns), so mean and variance need no exchange between threads and the pre-norm
// activations never leave the SM.
#include "common.cuh"
#include "tcgen05.cuh"

#include <cuda_bf16.h>

namespace asr {

constexpr int kGM = 128;                 // rows of x per CTA
constexpr int kGK = 64;                  // K per pipeline stage (128-byte rows)
constexpr int kGATile = kGM * kGK * 2;   // 16 KB

struct __align__(8) GemmBarriers {
    uint64_t full[4];
    uint64_t empty[4];
    uint64_t acc_full;
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// ---- producer and issuer, shared by both kernels: BN rows of W per stage, nk stages of 64 in K -------------
template <int BN, int STAGES>
__device__ __forceinline__ void gemm_produce(const CUtensorMap* tm_x, const CUtensorMap* tm_w, unsigned char* sA, unsigned char* sB,
                                             GemmBarriers* bars, int m0, int n0, int nk) {
    constexpr int kBTile = BN * kGK * 2;
    tma_prefetch_desc(tm_x);
    tma_prefetch_desc(tm_w);
    for (int k = 0; k < nk; ++k) {
        const int s = k % STAGES;
        if (k >= STAGES) mbar_wait(&bars->empty[s], ((k / STAGES) - 1) & 1);
        mbar_arrive_expect_tx(&bars->full[s], kGATile + kBTile);
        tma_load_2d(sA + s * kGATile, tm_x, k * kGK, m0, &bars->full[s]);
#pragma unroll
        for (int part = 0; part < BN / 256 + (BN % 256 ? 1 : 0); ++part)      // a TMA box has at most 256 rows
            tma_load_2d(sB + s * kBTile + part * 256 * kGK * 2, tm_w, k * kGK, n0 + part * 256, &bars->full[s]);
    }
}

template <int BN, int STAGES>
__device__ __forceinline__ void gemm_issue(uint32_t tmem, unsigned char* sA, unsigned char* sB, GemmBarriers* bars, int nk) {
    constexpr int kBTile = BN * kGK * 2;
    constexpr int kNI = BN > 256 ? 256 : BN;                  // N of one instruction
    constexpr uint32_t idesc = make_idesc(kGM, kNI, 0, 0);    // A and B K-major
    for (int k = 0; k < nk; ++k) {
        const int s = k % STAGES;
        mbar_wait(&bars->full[s], (k / STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + s * kGATile);
        const uint32_t b_addr = smem_u32(sB + s * kBTile);
#pragma unroll
        for (int kk = 0; kk < kGK / 16; ++kk) {
#pragma unroll
            for (int nn = 0; nn < BN / kNI; ++nn)
                umma_bf16(tmem + nn * kNI, smem_desc_sw128(a_addr + kk * 32, 16, 1024),
                          smem_desc_sw128(b_addr + nn * kNI * kGK * 2 + kk * 32, 16, 1024), idesc, (k > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(&bars->empty[s]);       // the stage is free once these products have read it
    }
    tc_commit(&bars->acc_full);
}

// ---- y = LayerNorm(x W^T + b + residual) * gamma + beta, N = 512: the whole row lives in tensor memory -----
constexpr int kLnN = 512;
constexpr int kLnStages = 2;
constexpr int kLnSmem = kLnStages * (kGATile + kLnN * kGK * 2) + 256 + 3 * kLnN * 4 /*bias, gamma, beta*/;

__global__ void __launch_bounds__(192, 1)
linear_res_ln_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const float* __restrict__ bias,
                     const __nv_bfloat16* __restrict__ residual, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float eps, __nv_bfloat16* __restrict__ y, int M, int K) {
    extern __shared__ __align__(1024) unsigned char smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    constexpr int kBTile = kLnN * kGK * 2;
    unsigned char* sA = smem;
    unsigned char* sB = sA + kLnStages * kGATile;
    GemmBarriers* bars = reinterpret_cast<GemmBarriers*>(sB + kLnStages * kBTile);
    float* sPar = reinterpret_cast<float*>(sB + kLnStages * kBTile + 256);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kGM;
    const int nk = K / kGK;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kLnStages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        mbar_init(&bars->acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 5) {
        tmem_alloc(&bars->tmem_base, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    if (warp == 4) {
        if (elect_one_sync()) gemm_produce<kLnN, kLnStages>(&tm_x, &tm_w, sA, sB, bars, m0, 0, nk);
    } else if (warp == 5) {
        if (elect_one_sync()) gemm_issue<kLnN, kLnStages>(tmem, sA, sB, bars, nk);
    } else {
        const int row = m0 + warp * 32 + lane;
        const int rowc = min(row, M - 1);
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        const uint4* res = reinterpret_cast<const uint4*>(residual + (size_t)rowc * kLnN);      // 64 x 16 bytes per row
        // while the products run: bias / gamma / beta into shared memory (read back as broadcast float4), and the
        // first residual chunks into registers
        for (int i = threadIdx.x; i < kLnN; i += 128) {
            sPar[i] = bias ? __ldg(bias + i) : 0.0f;
            sPar[kLnN + i] = __ldg(gamma + i);
            sPar[2 * kLnN + i] = __ldg(beta + i);
        }
        bar_sync_named(1, 128);
        const float4* sBias4 = reinterpret_cast<const float4*>(sPar);
        const float4* sGamma4 = reinterpret_cast<const float4*>(sPar + kLnN);
        const float4* sBeta4 = reinterpret_cast<const float4*>(sPar + 2 * kLnN);
        constexpr int kPre = 2;                        // residual chunks (32 columns = 4 x 16 bytes) in flight
        uint4 rbuf[kPre][4];
#pragma unroll
        for (int pch = 0; pch < kPre; ++pch)
#pragma unroll
            for (int q = 0; q < 4; ++q) rbuf[pch][q] = __ldg(res + pch * 4 + q);
        mbar_wait(&bars->acc_full, 0);
        tc_fence_after();
        // pass 1: z = acc + bias + residual, written back to tensor memory; row sum
        float sum = 0.0f;
#pragma unroll
        for (int c = 0; c < kLnN; c += 32) {
            const int slot = (c / 32) % kPre;
            uint4 rcur[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) rcur[q] = rbuf[slot][q];
            if (c + 32 * kPre < kLnN) {
#pragma unroll
                for (int q = 0; q < 4; ++q) rbuf[slot][q] = __ldg(res + (c / 32 + kPre) * 4 + q);
            }
            float v[32];
            tmem_ld32(taddr + c, v);
            uint32_t z[32];
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                const uint32_t rw[4] = {rcur[i >> 3].x, rcur[i >> 3].y, rcur[i >> 3].z, rcur[i >> 3].w};
                const float4 ba = sBias4[(c + i) >> 2], bb = sBias4[((c + i) >> 2) + 1];
                const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float r0 = __uint_as_float(rw[u] << 16), r1 = __uint_as_float(rw[u] & 0xffff0000u);
                    const float z0 = v[i + 2 * u] + bv[2 * u] + r0, z1 = v[i + 2 * u + 1] + bv[2 * u + 1] + r1;
                    sum += z0 + z1;
                    z[i + 2 * u] = __float_as_uint(z0);
                    z[i + 2 * u + 1] = __float_as_uint(z1);
                }
            }
            tmem_st32(taddr + c, z);
        }
        tmem_st_wait();
        const float mean = sum * (1.0f / kLnN);
        // pass 2: variance around the mean (as torch's LayerNorm: biased, two-pass accuracy)
        float sq = 0.0f;
#pragma unroll 2
        for (int c = 0; c < kLnN; c += 32) {
            float v[32];
            tmem_ld32(taddr + c, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float d = v[i] - mean;
                sq = fmaf(d, d, sq);
            }
        }
        const float rstd = rsqrtf(sq * (1.0f / kLnN) + eps);
        // pass 3: normalise, scale, shift, store
        __nv_bfloat16* dst = y + (size_t)rowc * kLnN;
#pragma unroll 2
        for (int c = 0; c < kLnN; c += 32) {
            float v[32];
            tmem_ld32(taddr + c, v);
            if (row < M) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    const float4 ga = sGamma4[(c + i) >> 2], gb = sGamma4[((c + i) >> 2) + 1];
                    const float4 ea = sBeta4[(c + i) >> 2], eb = sBeta4[((c + i) >> 2) + 1];
                    const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
                    const float ev[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
                    uint32_t w[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float a0 = fmaf((v[i + 2 * u] - mean) * rstd, gv[2 * u], ev[2 * u]);
                        const float a1 = fmaf((v[i + 2 * u + 1] - mean) * rstd, gv[2 * u + 1], ev[2 * u + 1]);
                        w[u] = pack_bf16x2(a0, a1);
                    }
                    *reinterpret_cast<uint4*>(dst + c + i) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}


}  // namespace asr

using namespace asr;

extern "C" int asr_gemm_f32(const float* a, int a_mn_major, int lda, const float* b, int b_mn_major, int ldb, const float* bias, int M,
                            int N, int K, float* c, int ldc, void* ws, size_t ws_bytes, void* stream);

// y = x W^T + b in fp32 (three TF32 products per K step): the general GEMM of gemm2.cu with both operands K-major.  (Round 1
// had its own kernel for this entry point; it lacked the row-contiguous epilogue and the split-K plan and is gone.)
extern "C" int asr_linear_f32(const float* x, const float* w, const float* bias, int M, int N, int K, float* y, void* stream) {
    ASR_REQUIRE(x && w && y, "asr_linear_f32: null pointer");
    ASR_REQUIRE(M > 0 && N > 0 && K > 0, "asr_linear_f32: bad shape M=%d N=%d K=%d", M, N, K);
    ASR_REQUIRE(K % 4 == 0, "asr_linear_f32: K=%d must be a multiple of 4 (16-byte rows for TMA)", K);
    ASR_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y), "asr_linear_f32: pointers must be 16-byte aligned");
    return asr_gemm_f32(x, 0, K, w, 0, K, bias, M, N, K, y, N, nullptr, 0, stream);
}

static int make_rowmajor_bf16_map(CUtensorMap* map, const void* base, int rows, int cols, int box_rows) {
    return make_tmap_2d(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, (uint64_t)rows, (uint64_t)cols, (uint64_t)cols * 2,
                        (uint32_t)box_rows, (uint32_t)kGK, CU_TENSOR_MAP_SWIZZLE_128B);
}

extern "C" int asr_gemm_bf16(const void* a, int a_mn_major, int lda, const void* b, int b_mn_major, int ldb, const float* bias,
                             int relu, int M, int N, int K, void* c, int ldc, int out_f32, void* ws, size_t ws_bytes, void* stream);

// y = act(x W^T + b) in bf16: the general GEMM of gemm2.cu with both operands K-major - at the feed-forward shapes its
// persistent kernel (accumulator double-buffered in tensor memory, rows leaving through shared memory).  Round 1's own
// kernel for this entry point (one CTA per tile, a thread storing its row 16 bytes at a time) ran at 705 TFLOP/s where the
// persistent one reaches 1000 (M = 102400, N = 2048, K = 512), with the same output bits; it is gone.
extern "C" int asr_linear_act_bf16(const void* x, const void* w, const float* bias, int M, int N, int K, int relu, void* y,
                                   void* stream) {
    ASR_REQUIRE(x && w && y, "asr_linear_act_bf16: null pointer");
    ASR_REQUIRE(M > 0 && N > 0 && K > 0, "asr_linear_act_bf16: bad shape M=%d N=%d K=%d", M, N, K);
    ASR_REQUIRE(N % 128 == 0 && K % kGK == 0, "asr_linear_act_bf16: N=%d must be a multiple of 128 and K=%d of 64", N, K);
    ASR_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y) && aligned16(bias), "asr_linear_act_bf16: pointers must be 16-byte aligned");
    return asr_gemm_bf16(x, 0, K, w, 0, K, bias, relu, M, N, K, y, N, 0, nullptr, 0, stream);
}

extern "C" int asr_linear_residual_layernorm_bf16(const void* x, const void* w, const float* bias, const void* residual,
                                                  const float* gamma, const float* beta, float eps, int M, int N, int K, void* y,
                                                  void* stream) {
    ASR_REQUIRE(x && w && residual && gamma && beta && y, "asr_linear_residual_layernorm_bf16: null pointer");
    ASR_REQUIRE(M > 0 && K > 0, "asr_linear_residual_layernorm_bf16: bad shape M=%d K=%d", M, K);
    ASR_REQUIRE(N == kLnN, "asr_linear_residual_layernorm_bf16: N=%d, only d_model = %d is supported (the row must fill tensor memory)", N, kLnN);
    ASR_REQUIRE(K % kGK == 0, "asr_linear_residual_layernorm_bf16: K=%d must be a multiple of 64", K);
    ASR_REQUIRE(aligned16(x) && aligned16(w) && aligned16(residual) && aligned16(y),
                "asr_linear_residual_layernorm_bf16: pointers must be 16-byte aligned");
    if (asr_device_ok() != 0) return 3;
    CUtensorMap tx, tw;
    if (make_rowmajor_bf16_map(&tx, x, M, K, kGM) || make_rowmajor_bf16_map(&tw, w, kLnN, K, 256)) return 4;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ASR_CHECK_CUDA(cudaFuncSetAttribute(linear_res_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLnSmem));
    linear_res_ln_kernel<<<(M + kGM - 1) / kGM, 192, kLnSmem, st>>>(tx, tw, bias, static_cast<const __nv_bfloat16*>(residual), gamma, beta,
                                                                   eps, static_cast<__nv_bfloat16*>(y), M, K);
    ASR_LAUNCH_CHECK();
    return 0;
}

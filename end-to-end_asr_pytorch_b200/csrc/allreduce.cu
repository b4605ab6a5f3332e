// allreduce.cu - gradient all-reduce (mean) over NVLink / NVSwitch peer memory, one kernel per rank (sm_100a).
//
// Replaces the NCCL all-reduce of the data-parallel training step (the reference trains on one GPU:
// /root/reference/src/transformer/solver.py:141-161 is the loop that is wrapped; SURVEY.md 8(e): one process per GPU,
// gradients only).  Every rank holds its fp32 gradient bucket in SYMMETRIC memory: the same allocation is mapped into
// every rank's address space (peer pointers) and, where the box has an NVSwitch with multicast objects, behind one
// multicast address.  The host side (dp.PeerAllReduce) only allocates and exchanges the handles
// (torch.distributed._symmetric_memory: plumbing); the collective itself is this kernel, launched by every rank on its
// own stream:
//
//   multicast flavour (NVLS):  rank r owns elements [r * n / W, (r + 1) * n / W).  multimem.ld_reduce.add.v4.f32 on the
//       multicast address makes the SWITCH read the 16 bytes from all W GPUs and return their sum; the rank scales it by
//       1 / W and multimem.st broadcasts the result into all W buffers.  Per GPU and direction ~n * 4 bytes cross the
//       links once (reduce) + once (broadcast); no GPU ever holds another GPU's raw gradients.
//   peer flavour (no multicast): the rank reads its slice from every peer buffer with ordinary loads through the peer
//       mappings, adds in rank order (deterministic), scales and stores the result into every peer's buffer.
//
// Ranks synchronise through flag words in the symmetric signal pad: CTA b of rank r raises flag (b, r) in every peer's
// pad and waits for its own W flags (release / acquire at system scope); once before the first load (all buckets are
// final: they were written by earlier kernels of each rank's stream) and once after the last store (nobody leaves - and
// lets its stream overwrite the bucket - before every rank has written its slice everywhere).
// Few CTAs on purpose: the kernel runs next to HBM-bound kernels of the training step and is bound by the links, not by
// the SMs; `ctas` x 512 threads x 8 (multicast) or 16 (peer flavour) x 16 bytes are in flight.
#include "common.cuh"

namespace asr {

constexpr int kArThreads = 512;

__device__ __forceinline__ void flag_raise(uint32_t* addr) {
    // 0 -> 1, waiting for the previous round's flag to be consumed; release: this rank's earlier writes are visible first
    uint32_t old;
    do {
        asm volatile("atom.release.sys.global.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory");
    } while (old != 0u);
}
__device__ __forceinline__ void flag_consume(uint32_t* addr) {
    // 1 -> 0; acquire: the peer's writes before its raise are visible after this
    uint32_t old;
    do {
        asm volatile("atom.acquire.sys.global.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(addr) : "memory");
    } while (old != 1u);
}
// All W ranks' CTA `blockIdx.x` meet.  pads[p] = rank p's signal pad as mapped here; slot (b, r) of a pad belongs to
// CTA b of rank r.  `round` picks one of two slot sets so that the entry and the exit barrier of one launch (and the next
// launch's entry barrier) never share a word that a slow rank has not consumed yet.
__device__ __forceinline__ void ranks_barrier(uint32_t* const* pads, int rank, int world, int round) {
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const size_t slot = ((size_t)round * gridDim.x + blockIdx.x) * (size_t)world;
        flag_raise(pads[threadIdx.x] + slot + rank);
        flag_consume(pads[rank] + slot + threadIdx.x);
    }
    __syncthreads();
}

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc_addr) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc_addr)
                 : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float* mc_addr, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

struct AllReduceArgs {
    float* const* peers;        // device array [world]: every rank's bucket as mapped in this process
    float* mc;                  // multicast address of the bucket (null: peer flavour)
    uint32_t* const* pads;      // device array [world]: every rank's signal pad
    int rank, world;
    size_t off4;                // first float4 of the range inside the buckets
    size_t n4;                  // float4 elements
    float scale;                // 1 / world
};

// kUnroll 16-byte elements per thread and pass; the peer flavour reads each of them from kWorld buffers (kWorld = 0: any
// number of ranks, one element per pass), so kUnroll * kWorld loads are in flight per thread - the links want megabytes in
// flight (measured on two B200s, 209 MB: 16 CTAs x 512 threads x 8 loads 2.4 ms, 128 CTAs 0.40 ms).
template <bool kMulticast, int kWorld, int kUnroll>
__global__ void __launch_bounds__(kArThreads) allreduce_mean_kernel(const AllReduceArgs a) {
    ranks_barrier(a.pads, a.rank, a.world, 0);
    const int world = kWorld > 0 ? kWorld : a.world;
    // this rank's slice, in float4 units
    const size_t per = (a.n4 + world - 1) / world;
    const size_t lo = a.off4 + min(a.n4, per * (size_t)a.rank);
    const size_t hi = a.off4 + min(a.n4, per * (size_t)a.rank + per);
    const size_t stride = (size_t)gridDim.x * kArThreads;
    for (size_t i0 = lo + (size_t)blockIdx.x * kArThreads + threadIdx.x; i0 < hi; i0 += stride * kUnroll) {
        float4 v[kUnroll];
        if (kMulticast) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const size_t i = i0 + u * stride;
                if (i < hi) v[u] = multimem_ld_reduce_add(a.mc + 4 * i);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const size_t i = i0 + u * stride;
                if (i < hi) {
                    v[u].x *= a.scale; v[u].y *= a.scale; v[u].z *= a.scale; v[u].w *= a.scale;
                    multimem_st(a.mc + 4 * i, v[u]);
                }
            }
        } else if (kWorld > 0) {
            float4 w[kUnroll][kWorld > 0 ? kWorld : 1];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const size_t i = i0 + u * stride;
#pragma unroll
                for (int p = 0; p < kWorld; ++p)
                    if (i < hi) w[u][p] = __ldcg(reinterpret_cast<const float4*>(a.peers[p]) + i);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const size_t i = i0 + u * stride;
                if (i < hi) {
                    float4 t = w[u][0];
#pragma unroll
                    for (int p = 1; p < kWorld; ++p) {         // rank order: the same sum on every rank and run
                        t.x += w[u][p].x; t.y += w[u][p].y; t.z += w[u][p].z; t.w += w[u][p].w;
                    }
                    t.x *= a.scale; t.y *= a.scale; t.z *= a.scale; t.w *= a.scale;
#pragma unroll
                    for (int p = 0; p < kWorld; ++p) __stcg(reinterpret_cast<float4*>(a.peers[p]) + i, t);
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const size_t i = i0 + u * stride;
                if (i < hi) {
                    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int p = 0; p < world; ++p) {
                        const float4 w = __ldcg(reinterpret_cast<const float4*>(a.peers[p]) + i);
                        t.x += w.x; t.y += w.y; t.z += w.z; t.w += w.w;
                    }
                    t.x *= a.scale; t.y *= a.scale; t.z *= a.scale; t.w *= a.scale;
                    for (int p = 0; p < world; ++p) __stcg(reinterpret_cast<float4*>(a.peers[p]) + i, t);
                }
            }
        }
    }
    __threadfence_system();
    ranks_barrier(a.pads, a.rank, a.world, 1);
}

}  // namespace asr

using namespace asr;

extern "C" size_t asr_allreduce_signal_bytes(int world, int ctas) {
    return (size_t)2 * (size_t)(ctas > 0 ? ctas : 0) * (size_t)(world > 0 ? world : 0) * sizeof(uint32_t);
}

extern "C" int asr_allreduce_mean_f32(const void* peer_ptrs_dev, void* multicast_ptr, const void* signal_pads_dev, int rank, int world,
                                      size_t offset, size_t n, int ctas, void* stream) {
    ASR_REQUIRE(peer_ptrs_dev && signal_pads_dev, "asr_allreduce_mean_f32: null pointer table");
    ASR_REQUIRE(world >= 1 && world <= 32 && rank >= 0 && rank < world, "asr_allreduce_mean_f32: bad rank %d / world %d", rank, world);
    ASR_REQUIRE(n % 4 == 0 && offset % 4 == 0, "asr_allreduce_mean_f32: offset %zu and length %zu must be multiples of 4 floats", offset, n);
    ASR_REQUIRE(ctas >= 1 && ctas <= 128, "asr_allreduce_mean_f32: ctas %d outside [1, 128]", ctas);
    ASR_REQUIRE(multicast_ptr == nullptr || aligned16(multicast_ptr), "asr_allreduce_mean_f32: multicast pointer must be 16-byte aligned");
    if (asr_device_ok() != 0) return 3;
    if (n == 0) return 0;
    AllReduceArgs a;
    a.peers = static_cast<float* const*>(peer_ptrs_dev);
    a.mc = static_cast<float*>(multicast_ptr);
    a.pads = static_cast<uint32_t* const*>(signal_pads_dev);
    a.rank = rank;
    a.world = world;
    a.off4 = offset / 4;
    a.n4 = n / 4;
    a.scale = 1.0f / (float)world;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (multicast_ptr != nullptr) allreduce_mean_kernel<true, 0, 8><<<ctas, kArThreads, 0, st>>>(a);
    else if (world == 2) allreduce_mean_kernel<false, 2, 8><<<ctas, kArThreads, 0, st>>>(a);
    else if (world == 4) allreduce_mean_kernel<false, 4, 4><<<ctas, kArThreads, 0, st>>>(a);
    else if (world == 8) allreduce_mean_kernel<false, 8, 2><<<ctas, kArThreads, 0, st>>>(a);
    else allreduce_mean_kernel<false, 0, 2><<<ctas, kArThreads, 0, st>>>(a);
    ASR_LAUNCH_CHECK();
    return 0;
}

// common.cu - error plumbing, options, TMA descriptor encoding.
#include "common.cuh"

#include <stdarg.h>
#include <string.h>


namespace asr {

static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};
// Tuning options: a fixed table of atomics, read on every call without a lock (the names are compile-time constants
// of the callers, a dozen strcmp's are cheaper than a mutex'd std::map<std::string> lookup).
struct Opt {
    const char* name;
    std::atomic<int> value;
};
static Opt g_opts[] = {
    {"cif_fwd_variant", {0}}, {"cif_fwd_width", {0}}, {"cif_fwd_stages", {0}}, {"cif_fwd_rows", {0}},
    {"mha_variant", {0}}, {"ctc_fuse_apply", {0}}, {"ctc_lattice_variant", {0}}, {"ctc_chunks", {0}},
    {"ctc_finish_per_slice", {0}}, {"mha_bwd_groups", {0}}, {"gemm_variant", {0}}, {"gemm_f32_bn", {0}},
    {"gemm_split_k", {0}}, {"gemm_split_mode", {0}}, {"gemm_stage_out", {0}}, {"gemm_persistent", {0}}, {"gemm_debug", {0}}, {"ctc_lattice_split", {0}},
};
static Opt* find_opt(const char* key) {
    for (Opt& o : g_opts)
        if (strcmp(o.name, key) == 0) return &o;
    return nullptr;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int get_opt(const char* key) {
    const Opt* o = find_opt(key);
    return o ? o->value.load(std::memory_order_relaxed) : 0;
}

void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

int num_sms() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;
    }
    return cached;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_tmap_nd(CUtensorMap* map, CUtensorMapDataType dtype, size_t elem_bytes, const void* base, int rank,
                 const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                 CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = encode_fn();
    ASR_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
    // cuTensorMapEncodeTiled is a DRIVER call: it needs a context current on the calling thread.  A thread that has only
    // ever selected a device (torch's autograd worker threads do cudaSetDevice and nothing else before calling into this
    // library) may not have the primary context bound yet -> CUDA_ERROR_INVALID_CONTEXT (201).  One runtime call per
    // thread binds it.
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) {
        cudaFree(nullptr);
        ctx_bound = true;
    }
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bx[5];
    cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    (void)elem_bytes;
    CUresult r = fn(map, dtype, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ASR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu, box %u x %u)",
                (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
                rank > 1 ? box[1] : 0);
    return 0;
}

int make_tmap_2d(CUtensorMap* map, CUtensorMapDataType dtype, size_t elem_bytes, const void* base, uint64_t rows,
                 uint64_t cols, uint64_t row_stride_bytes, uint32_t box_rows, uint32_t box_cols,
                 CUtensorMapSwizzle swizzle) {
    uint64_t dims[2] = {cols, rows};
    uint64_t strides[1] = {row_stride_bytes};
    uint32_t box[2] = {box_cols, box_rows};
    return make_tmap_nd(map, dtype, elem_bytes, base, 2, dims, strides, box, swizzle);
}

}  // namespace asr

extern "C" {

int asr_abi_version(void) { return ASR_SM100_ABI_VERSION; }

const char* asr_last_error(void) { return asr::g_err; }

int asr_device_ok(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        asr::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
        return 1;
    }
    static std::atomic<uint64_t> ok_mask{0};      // devices already checked (this runs on every compute call)
    if (dev < 64 && (ok_mask.load(std::memory_order_relaxed) >> dev) & 1) return 0;
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) {
        asr::set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
        return 1;
    }
    if (major != 10) {
        asr::set_error("device %d has compute capability %d.x; libasr_sm100 is sm_100a only", dev, major);
        return 2;
    }
    if (dev < 64) ok_mask.fetch_or(1ull << dev, std::memory_order_relaxed);
    return 0;
}

int asr_set_option(const char* key, int value) {
    asr::Opt* o = key ? asr::find_opt(key) : nullptr;
    if (o == nullptr) {
        asr::set_error("unknown option '%s'", key ? key : "(null)");
        return 2;
    }
    o->value.store(value, std::memory_order_relaxed);
    return 0;
}

int asr_get_option(const char* key, int* value) {
    const asr::Opt* o = key ? asr::find_opt(key) : nullptr;
    if (o == nullptr || value == nullptr) {
        asr::set_error("unknown option '%s'", key ? key : "(null)");
        return 2;
    }
    *value = o->value.load(std::memory_order_relaxed);
    return 0;
}

uint64_t asr_launch_count(void) { return asr::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"

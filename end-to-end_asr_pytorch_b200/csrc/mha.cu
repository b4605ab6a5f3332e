// mha.cu - multi-head attention core (placeholder until the tcgen05 kernels land).
#include "common.cuh"

extern "C" int asr_mha_fwd_bf16(const void*, const void*, const void*, const int*, const uint8_t*, int, int, int, int,
                                int, int, float, void*, float*, void*) {
    asr::set_error("asr_mha_fwd_bf16: not built yet");
    return 9;
}
extern "C" size_t asr_mha_bwd_workspace_bytes(int, int, int, int, int) { return 0; }
extern "C" int asr_mha_bwd_bf16(const void*, const void*, const void*, const void*, const void*, const float*,
                                const int*, const uint8_t*, int, int, int, int, int, int, float, void*, void*, void*,
                                void*, size_t, void*) {
    asr::set_error("asr_mha_bwd_bf16: not built yet");
    return 9;
}
extern "C" int asr_mha_probs_f32(const void*, const void*, const int*, const uint8_t*, int, int, int, int, int, int,
                                 float, float*, void*) {
    asr::set_error("asr_mha_probs_f32: not built yet");
    return 9;
}

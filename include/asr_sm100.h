/*
 * asr_sm100.h - C ABI of libasr_sm100.so: the B200 (sm_100a) acoustic-model
 * training hot path of eastonYi/end-to-end_asr_pytorch.
 *
 * The reference has no FFI / operator registry: the drop-in
 * boundary is the Python API of three modules.  Every entry point below is what
 * the Python-side `torch.autograd.Function` wrappers bind through ctypes, and
 * each one names the reference code it replaces.
 *
 * Conventions
 *   - plain C symbols, plain pointers and sizes, no torch / C++ types;
 *   - return 0 on success, non-zero on failure (never throws, never exits);
 *     `asr_last_error()` returns a thread-local description of the last failure;
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - no allocation inside the library: outputs, saved-for-backward buffers and
 *     workspaces are owned by the caller (PyTorch caching allocator);
 *   - tensors are contiguous, row-major; base pointers 16-byte aligned;
 *   - `stream` is a `cudaStream_t` passed as void*; all calls are asynchronous
 *     and ordered on that stream: they start after the work already queued on it
 *     and the work queued after them sees their results.  asr_ctc_fwd_bwd_f32 may
 *     fork onto library-owned streams in between (event fork/join, also legal
 *     under CUDA-graph capture); calls are stateless, hence re-entrant per stream;
 *   - there is NO CPU fallback: on a machine without an sm_100 device every
 *     compute entry point returns an error.
 */
#ifndef ASR_SM100_H
#define ASR_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASR_SM100_ABI_VERSION 1

/* ---- housekeeping ------------------------------------------------------- */
int         asr_abi_version(void);
const char* asr_last_error(void);
/* 0 when the current CUDA device is compute capability 10.x, else non-zero. */
int         asr_device_ok(void);
/* Tuning knobs (kernel variants; all variants are sm_100a CUDA).  Unknown keys
 * return non-zero.  Keys: "cif_fwd_variant" (0 = auto, 1 = plain loads,
 * 2 = one-warp TMA pipeline, 3 = warp-specialised TMA pipeline, 4 = schedule kernel +
 * segment-parallel rows), "cif_fwd_width"
 * (0 = auto, 32/64/128 floats per warp), "cif_fwd_stages" (0 = auto),
 * "cif_fwd_rows" (variant 3: data warps per CTA, 0 = auto), "ctc_fuse_apply"
 * (0 = separate K3 pass applies the sparse gradient update (default, faster),
 * 1 = the lattice kernel applies it itself with RED.ADD), "ctc_lattice_variant"
 * (0 = bidirectional four-warp lattice (default), 1 = one warp per utterance),
 * "ctc_chunks" (asr_ctc_fwd_bwd_f32 / begin / finish slice the batch into this
 * many pieces and run each slice's lattice on a library-owned stream so that it
 * overlaps the HBM-bound row kernels of the next slice; 0 = auto, 1 = no slicing,
 * max 8; results are bit-identical for every value), "ctc_finish_per_slice"
 * (asr_ctc_finish_f32: 0 = one apply launch after all lattices, 1 = slice by
 * slice), "mha_variant" (0 = auto = 21 without dropout, 3 with,
 * 1 = one tile per CTA with four softmax warps, 2 = two-tile ping-pong, 3 = eight
 * softmax warps per tile, O accumulated in TMEM with lazy rescale, 4 = as 3 with P
 * kept in TMEM as the A operand of P V, 8 = two tiles per CTA sharing K/V, one thread
 * per query row, scores read from TMEM once, 10 = 8 with packed f32x2 arithmetic,
 * 21 = 10 with the tiles taking turns on the XU pipe, P in TMEM and one MMA issuer
 * per tile), "gemm_variant" (asr_linear_act_bf16 tiles: 0 = auto = 2,
 * 1 = 128 x 128 with a 3-deep ring, 2 = 128 x 256 with a 2-deep ring, both two CTAs per SM,
 * 3 = 128 x 256 with a 4-deep ring, one CTA per SM), "mha_bwd_groups" (softmax-backward
 * warps per CTA = 4 * groups; 0 = default (4 groups), 2). */
int         asr_set_option(const char* key, int value);
int         asr_get_option(const char* key, int* value);
/* Number of kernels launched by this library since load (all streams). */
uint64_t    asr_launch_count(void);

/* ---- CIF: continuous integrate-and-fire --------------------------------- */
/*
 * Replaces CIF_Model.cif, /root/reference/src/transformer/cif_model.py:57-106
 * (forward) and the autograd graph it builds (backward).
 *
 *   hidden  [B,T,H] f32   encoder frames
 *   alphas  [B,T]   f32   per-frame weights (already scaled, cif_model.py:48)
 *   L               rows of the output; the caller sizes it the reference's
 *                   way: max_b int(round(sum_t alphas[b,t]))  (cif_model.py:95-96)
 * outputs
 *   out       [B,L,H] f32  fired frames, zero rows beyond n_fired[b] (written
 *                          entirely by the kernel, no pre-zeroing needed)
 *   fire_t    [B,L]   i32  frame index of the k-th fire, -1 beyond n_fired[b]
 *   n_fired   [B]     i32  number of fires (may exceed L: rows >= L are dropped
 *                          and the Python wrapper raises, like the reference's
 *                          torch.zeros(negative) at cif_model.py:100)
 *   cur, rem  [B,T]   f32  saved for backward (cif_model.py:78-81)
 *   sched     [B,T]   i32  saved for backward: (#fires before t) << 1 | fire_t
 *   alpha_sum [B]     f32  sum_t alphas[b,t] (sequential fp32 order); with
 *   target_num[B]     f32  (may be NULL) the kernel also emits
 *   qua_term  [B]     f32  (alpha_sum - target_num)^2, the per-utterance term of
 *                          the quantity loss (transformer/loss.py:55)
 */
int asr_cif_fwd_f32(const float* hidden, const float* alphas, float threshold,
                    int B, int T, int H, int L,
                    float* out, int* fire_t, int* n_fired,
                    float* cur, float* rem, int* sched,
                    float* alpha_sum, const float* target_num, float* qua_term,
                    void* stream);
/*
 * Same call with a per-call kernel choice: kernel_hint = 0 lets the library choose (like asr_cif_fwd_f32), 1 = plain
 * one-warp kernel, 2 = one-warp TMA pipeline, 3 = warp-specialised TMA kernel (the one that disturbs co-running
 * latency-bound kernels least), 4 = schedule + segment-parallel rows.  Every choice produces identical bits; the hint is
 * tuning state that belongs to the call site (e.g. "queued between asr_ctc_begin_f32 and asr_ctc_finish_f32"), not to
 * the process-wide option table.
 */
int asr_cif_fwd_hint_f32(const float* hidden, const float* alphas, float threshold,
                         int B, int T, int H, int L,
                         float* out, int* fire_t, int* n_fired,
                         float* cur, float* rem, int* sched,
                         float* alpha_sum, const float* target_num, float* qua_term,
                         int kernel_hint, void* stream);

/*
 * Analytic backward of the above (SURVEY.md 8a row a2').
 *   g_out [B,L,H] -> g_hidden [B,T,H], g_alphas [B,T]
 * ws: asr_cif_bwd_workspace_bytes(B,T) bytes of device scratch.
 */
size_t asr_cif_bwd_workspace_bytes(int B, int T);
int asr_cif_bwd_f32(const float* hidden, const float* g_out,
                    const int* n_fired, const float* cur, const float* rem, const int* sched,
                    int B, int T, int H, int L,
                    float* g_hidden, float* g_alphas,
                    void* ws, size_t ws_bytes, void* stream);

/* ---- CIF weight producer --------------------------------- */
/*
 * Replaces the tail of Attention_Assigner.forward,
 * /root/reference/src/transformer/attentionAssigner.py:36-40
 *     alphas = sigmoid(linear(x).squeeze(-1)) * sequence_mask(input_lengths)
 * and the scaling glue of CIF_Model.forward, src/transformer/cif_model.py:43-48
 *     _num = alpha.sum(-1);  alpha *= (num_noise / _num)[:, None]
 * in one pass over x (padded frames are not read).
 *
 *   x [B,T,D] f32  assigner activations after its dropout; w [D], bias [1]: `linear`
 *   len [B] i32    valid frames; num_noise [B] f32 = #labels + U[0,1) - 0.5, drawn
 *                  by the caller (cif_model.py:46-47), or NULL: no scaling (decoding)
 * outputs
 *   alpha   [B,T]  the weights cif() consumes (0 at padded frames)
 *   a_raw   [B,T]  unscaled weights, saved for backward
 *   num_raw [B]    _num = sum_t a_raw, the quantity-loss input (loss.py:55)
 * backward: g_alpha [B,T], g_num [B] or NULL (gradient of the quantity loss w.r.t.
 * _num) -> g_x [B,T,D], g_w [D], g_bias [1].  All reductions have a fixed order.
 */
int asr_cif_alpha_fwd_f32(const float* x, const float* w, const float* bias,
                          const int* len, const float* num_noise,
                          int B, int T, int D,
                          float* alpha, float* a_raw, float* num_raw, void* stream);
size_t asr_cif_alpha_bwd_workspace_bytes(int B, int T, int D);
int asr_cif_alpha_bwd_f32(const float* x, const float* w, const int* len,
                          const float* num_noise, const float* a_raw, const float* num_raw,
                          const float* g_alpha, const float* g_num,
                          int B, int T, int D,
                          float* g_x, float* g_w, float* g_bias,
                          void* ws, size_t ws_bytes, void* stream);

/* ---- input side: low-frame-rate stacking ----------------- */
/*
 * Replaces build_LFR_features, /root/reference/src/utils/data.py:191-218 (numpy, per
 * utterance in the data loader) for a padded batch already on the device:
 *   out[b, i, j*D + d] = in[b, min(i*n + j, len[b]-1), d]   i < ceil(len[b]/n), else 0
 *   in [B,T,D] f32, len [B] i32 -> out [B, ceil(T/n), m*D] f32, out_len [B] = ceil(len/n).
 * A pure copy: bit-exact with the reference.
 */
int asr_lfr_f32(const float* in, const int* len, int B, int T, int D, int m, int n,
                float* out, int* out_len, void* stream);

/* ---- input side: SpecAugment ------------------------------ */
/*
 * Replaces the masking loops of spec_aug, /root/reference/src/utils/utils.py:168-194, for a
 * padded batch on the device, in place.  The caller draws the bands / spans (the reference
 * draws them with torch.rand, utils.py:178-181 and 186-189; the host mirror makes the same
 * calls in the same order) and passes them as [R,B] i32 arrays:
 *   freq_mean[b,t] = mean_v feats[b,t,:]            time_mean[b,v] = sum_t feats[b,:,v] / lens[b]
 *   feats[b,t,v]   = time_mean[b,v]  if t0[r,b] <= t < t0[r,b]+tw[r,b] for some r
 *                    freq_mean[b,t]  else if f0[r,b] <= v < f0[r,b]+fw[r,b] for some r
 * (both means of the batch as it was on entry; identical to the reference's mask-by-mask
 * order because every mask of a family writes the same value and the time spans are
 * applied last).  Bands / spans are clipped to [0,V) / [0,T).  V <= 1024.  The workspace
 * holds the two means and the per-64-frame column sums.
 */
size_t asr_spec_aug_workspace_bytes(int B, int T, int V);
int asr_spec_aug_f32(float* feats, const int* lens, const int* f0, const int* fw, const int* t0, const int* tw,
                     int R, int B, int T, int V, void* ws, size_t ws_bytes, void* stream);

/* ---- linear layers with fused epilogues ------------------- */
/*
 * y = act(x W^T + bias): x [M,K] bf16, W [N,K] bf16 (torch Linear.weight as it is), bias [N]
 * f32 or NULL, y [M,N] bf16, fp32 accumulation; relu != 0 applies max(.,0).  Replaces
 * relu(w_1(x)) of PositionwiseFeedForward, /root/reference/src/transformer/module.py:50,
 * and the w_qs / w_ks / w_vs projections of attention.py:40-45.  N % 128 == 0, K % 64 == 0.
 */
int asr_linear_act_bf16(const void* x, const void* w, const float* bias, int M, int N, int K, int relu,
                        void* y, void* stream);
/*
 * y = LayerNorm(x W^T + bias + residual) * gamma + beta over the last dimension, N = 512:
 * w_2 + residual + layer_norm of module.py:50-52 and fc + residual + layer_norm of
 * attention.py:59-60 with the dropout between them off (evaluation or p = 0).  residual
 * [M,512] bf16, gamma / beta [512] f32, eps as nn.LayerNorm (1e-5 by default).  K % 64 == 0.
 */
int asr_linear_residual_layernorm_bf16(const void* x, const void* w, const float* bias, const void* residual,
                                       const float* gamma, const float* beta, float eps, int M, int N, int K,
                                       void* y, void* stream);

/*
 * y = x W^T + bias in fp32 ON THE TENSOR CORES with fp32-level accuracy: every operand is split
 * into a TF32 head and an fp32 remainder and three tcgen05.mma kind::tf32 products are accumulated
 * (hi*hi + lo*hi + hi*lo; relative error ~1e-6 instead of TF32's 1e-3).  x [M,K], W [N,K] (torch
 * Linear.weight), bias [N] or NULL, y [M,N], all f32.  For the fp32 Linear layers of the model
 * shell (cuBLAS runs them as SIMT sgemm) and the ctc_fc / tgt_word_prj vocabulary projections
 * (/root/reference/src/transformer/cif_model.py:38), whose logits feed the CTC kernels at the
 * 1e-5 bar.  K % 4 == 0; any M, N (N = 4233 included).
 */
int asr_linear_f32(const float* x, const float* w, const float* bias, int M, int N, int K, float* y,
                   void* stream);

/*
 * General GEMMs for the TRAINING half of the linear layers and for the vocabulary projection
 * fused with the CTC loss (8(f1)):  C[M,N] = A B (+ bias[N]),  either operand K-major or MN-major, so that all
 * three products of y = x W^T run on the tensors as torch stores them, without transposed copies
 * (what autograd computes for /root/reference/src/transformer/module.py:46-53, attention.py:40-45,59-60,
 * cif_model.py:38):
 *     forward  y  = x  W^T :  A = x  (a_mn_major 0, [M,K], lda)   B = W  (b_mn_major 0, [N,K], ldb)
 *     dX       gx = gy W   :  A = gy (a_mn_major 0, [M,N])        B = W  (b_mn_major 1: stored [K_contract = N, N_out = K])
 *     dW       gW = gy^T x :  A = gy (a_mn_major 1: stored [K_contract = rows, M_out = N])   B = x (b_mn_major 1)
 * a_mn_major = 0: a is [M,K] with row stride lda;  1: a is [K,M] with row stride lda (same for b with N).
 * Row strides in elements, multiples of 16 bytes; sizes themselves are free (TMA zero-fills the edges, e.g. the
 * 4233-wide vocabulary inside rows padded to 4240).  c [M,N] with row stride ldc.
 * asr_gemm_f32 : fp32 in / out on the tensor cores at fp32-level accuracy (three TF32 products per K step).
 * asr_gemm_bf16: bf16 operands, fp32 accumulation, bf16 output (out_f32 = 0; relu = 1 fuses max(.,0)) or fp32 output
 *                (out_f32 = 1: weight gradients for fp32 master weights).
 * ws / ws_bytes: optional workspace of asr_gemm_workspace_bytes(M, N, K) bytes; passing it allows products with few
 * output tiles (dW of a small layer; at the model's shapes nearly every product) to be split along K.  The CTAs of one
 * output tile are a thread-block cluster and add their partial tiles through distributed shared memory in a fixed order
 * (deterministic; the workspace itself is only used for 256-wide bf16 tiles and with option gemm_split_mode = 1).
 */
size_t asr_gemm_workspace_bytes(int M, int N, int K);
/* out[n] = sum_m x[m, n] in fp32 (x fp32 or bf16 [M,N], row stride ld elements): the bias gradient of a linear layer
 * (the column sums of gy), one pass, fixed order (deterministic). */
int asr_colsum(const void* x, int is_bf16, int M, int N, int ld, float* out, void* stream);
/* LayerNorm(dropout(y) + residual) of a training step, and its backward - replaces what torch runs for
 * /root/reference/src/transformer/module.py:50-52, attention.py:59-60 and encoder.py:49 (nn.Dropout, +, nn.LayerNorm and
 * their autograd nodes).  y [M, D] fp32 or bf16 (y_bf16), residual [M, D] fp32 or NULL, gamma / beta [D], D = 256 / 512 /
 * 1024; z [M, D] = dropout(y) + residual is written for the backward (NULL: not written - only when it would equal y);
 * out [M, D], mean / rstd [M] fp32.  row_scale [M] or NULL: out is multiplied by row_scale[row] (the non-pad mask of
 * encoder.py:76-80 / decoder.py:628-634; constant, no gradient).  Dropout keeps an element when its Philox byte >= round(256 p_drop), scaled by
 * 1 / asr_ln_dropout_keep_prob(p_drop); the mask is regenerated in the backward from the same seed (seed_dev != NULL: the
 * seed is *seed_dev + seed, read on the device - CUDA-graph replays).
 * asr_ln_bwd: g_out, z, mean, rstd, gamma as saved -> g_z [M, D] fp32 (the residual's gradient; NULL: not wanted), g_y [M, D]
 * fp32 / bf16 (= g_z * keep / p_keep; NULL: not wanted), g_gamma_beta [2, D].  ws: asr_ln_bwd_workspace_bytes(M, D) bytes.
 * out_bf16 (asr_ln_fwd, optional): a bf16 copy of out for the next layer's bf16 GEMM (saves autocast's conversion kernel);
 * g_out_bf16 (asr_ln_bwd, optional): the gradient that came back through that copy, added to g_out (either may be NULL).
 * Fixed summation orders throughout (deterministic). */
float asr_ln_dropout_keep_prob(float p_drop);
int asr_ln_fwd(const void* y, int y_bf16, const float* residual, const float* gamma, const float* beta,
               const float* row_scale, int M, int D, float eps, float p_drop, uint64_t seed, const uint64_t* seed_dev, float* z, float* out,
               void* out_bf16, float* mean, float* rstd, void* stream);
size_t asr_ln_bwd_workspace_bytes(int M, int D);
int asr_ln_bwd(const float* g_out, const void* g_out_bf16, const float* z, const float* mean, const float* rstd, const float* gamma,
               const float* row_scale, int M, int D, float p_drop, uint64_t seed, const uint64_t* seed_dev, float* g_z, void* g_y, int y_bf16,
               float* g_gamma_beta, void* ws, size_t ws_bytes, void* stream);
/* Evaluation flavour of the above: out = LayerNorm(y + residual) * gamma + beta, y / residual (or NULL) / out bf16 [M, D],
 * nothing saved.  With asr_gemm_bf16 in front it is the faster route for module.py:50-52 / attention.py:59-60 in evaluation
 * (0.22 ms against 0.31 ms of the one-kernel asr_linear_residual_layernorm_bf16 at M = 102400, K = 2048). */
int asr_ln_eval_bf16(const void* y, const void* residual, const float* gamma, const float* beta, int M, int D, float eps,
                     void* out, void* stream);
/* out[i] = y[i] > 0 ? gy[i] : 0 (bf16, n elements, 16-byte aligned): the ReLU backward of module.py:50 from the saved
 * output, what torch runs as compare + cast + multiply. */
int asr_relu_bwd_bf16(const void* gy, const void* y, void* out, size_t n, void* stream);
/* keep [M, D] u8 = 1 where that dropout keeps the element (tests, inspection) */
int asr_ln_dropout_keep(uint8_t* keep, int M, int D, float p_drop, uint64_t seed, void* stream);
int asr_gemm_f32(const float* a, int a_mn_major, int lda, const float* b, int b_mn_major, int ldb,
                 const float* bias, int M, int N, int K, float* c, int ldc,
                 void* ws, size_t ws_bytes, void* stream);
/* asr_gemm_f32 on RAGGED rows (a padded batch of utterances): the row dimension - M when a is K-major, the contraction
 * when a is MN-major (then b must be MN-major too) - consists of groups of `group_rows` rows of which only the first
 * row_len[g] are valid; rows beyond are known to be zero (gradient rows of padded frames) or never read (their logits).
 * Row tiles / K steps that lie entirely in padding are skipped: for a K-major a, dead output tiles are left unwritten
 * (skip_dead_output = 1) or written as zeros (0).  Results on the valid rows are identical to asr_gemm_f32. */
int asr_gemm_f32_ragged(const float* a, int a_mn_major, int lda, const float* b, int b_mn_major, int ldb,
                        const float* bias, int M, int N, int K, float* c, int ldc,
                        const int* row_len, int group_rows, int skip_dead_output,
                        void* ws, size_t ws_bytes, void* stream);
int asr_gemm_bf16(const void* a, int a_mn_major, int lda, const void* b, int b_mn_major, int ldb,
                  const float* bias, int relu, int M, int N, int K, void* c, int ldc, int out_f32,
                  void* ws, size_t ws_bytes, void* stream);

/* ---- CTC loss (fused log-softmax, alpha-beta, gradient) ------------------ */
/*
 * Replaces log_softmax + torch.nn.functional.ctc_loss as called at
 * /root/reference/src/transformer/loss.py:39-43, src/ctcModel/loss.py:7-11 and
 * src/mask_lm/loss.py:38-41 (blank = V-1, targets 0-padded [B,S] int64).
 *
 *   logits   [B,T,V] f32
 *   targets  [B,S]   i64   0 = padding; tgt_len[b] = number of leading labels
 *   in_len   [B]     i32   valid frames per utterance (<= T)
 *   tgt_len  [B]     i32   labels per utterance (<= S)
 * outputs
 *   nll      [B]     f32   -log p(targets|logits); +inf when no alignment exists
 *   g_logits [B,T,V] f32   (may be NULL: forward only) d(mean loss)/d logits,
 *                          mean loss = mean_b(nll_b / max(tgt_len_b,1));
 *                          exactly 0 for t >= in_len[b]; NaN rows for an
 *                          infeasible utterance (zero_infinity=False semantics)
 * ws: asr_ctc_workspace_bytes(B,T,V,S) bytes of device scratch.  It holds the
 *     gathered log-probabilities [B,T,S+3] (overwritten in place by the per-frame
 *     label occupancies) and the repeated-label links [B,S]; the alpha/beta
 *     lattice itself never leaves shared memory.
 */
size_t asr_ctc_workspace_bytes(int B, int T, int V, int S);
int asr_ctc_fwd_bwd_f32(const float* logits, const int64_t* targets,
                        const int* in_len, const int* tgt_len,
                        int B, int T, int V, int S, int blank,
                        float* nll, float* g_logits,
                        void* ws, size_t ws_bytes, void* stream);
/* The same call on logits / gradient rows with a row stride ld >= V (in floats; ld % 4 == 0 gives 16-byte-aligned
 * rows).  g_logits may alias logits: every row is read completely before its gradient is written, so the gradient
 * can replace the logits in place - the form the fused vocabulary projection uses (asr_gemm_f32 writes the logits
 * into rows padded to a multiple of 4, the CTC kernels turn them into the gradient, two more GEMMs consume it). */
int asr_ctc_fwd_bwd_ld_f32(const float* logits, const int64_t* targets,
                           const int* in_len, const int* tgt_len,
                           int B, int T, int V, int ld, int S, int blank,
                           float* nll, float* g_logits,
                           void* ws, size_t ws_bytes, void* stream);
/* The same call in two phases.  begin queues the row kernels on `stream` and the
 * lattices on library-owned streams and returns a ticket; finish (same arguments,
 * same stream, the ticket) waits for the lattices and applies the sparse gradient
 * update.  Whatever the caller queues on `stream` between the two runs next to
 * the last slice's lattice, which is a latency-bound chain that leaves the GPU
 * almost idle - bench.py puts the CIF forward/backward pair there.  finish then
 * applies the whole batch in one launch (option "ctc_finish_per_slice" = 1: slice
 * by slice as each lattice completes, which is what asr_ctc_fwd_bwd_f32 does).  nll and
 * g_logits are complete (in stream order) after finish.  The library owns the
 * lattice streams and 4 ticket slots per device: a 5th begin before any finish
 * returns an error (it does not alias an open ticket), and finish on a ticket that
 * is not open returns an error.  asr_ctc_fwd_bwd_f32 == begin immediately followed
 * by finish.  A target label outside [0,V) or equal to `blank` makes that
 * utterance's nll and gradient rows NaN (all CTC entry points). */
int asr_ctc_begin_f32(const float* logits, const int64_t* targets,
                      const int* in_len, const int* tgt_len,
                      int B, int T, int V, int S, int blank,
                      float* nll, float* g_logits,
                      void* ws, size_t ws_bytes, void* stream, int* ticket);
int asr_ctc_finish_f32(const float* logits, const int64_t* targets,
                       const int* in_len, const int* tgt_len,
                       int B, int T, int V, int S, int blank,
                       float* nll, float* g_logits,
                       void* ws, size_t ws_bytes, void* stream, int ticket);
/* Same call split into its three kernels, for per-kernel timing with CUDA events
 * (bench.py): stages is a bit mask, 1 = K1 row pass (log-sum-exp, gather, dense
 * gradient), 2 = K2 lattice (alpha/beta, nll, occupancies), 4 = K3 sparse gradient
 * update.  Stages must be issued in order on one stream; no slicing, one stream. */
int asr_ctc_stages_f32(const float* logits, const int64_t* targets,
                       const int* in_len, const int* tgt_len,
                       int B, int T, int V, int S, int blank,
                       float* nll, float* g_logits,
                       void* ws, size_t ws_bytes, int stages, void* stream);
/* In-place g *= *scale_dev, skipped on the device when *scale_dev == 1.0f
 * (autograd's incoming gradient for the loss; it is 1 in the reference's
 * solvers, transformer/solver.py:153). */
int asr_scale_inplace_f32(float* g, size_t n, const float* scale_dev, void* stream);

/* ---- multi-head attention core ------------------------------------------ */
/*
 * Replaces ScaledDotProductAttention.forward,
 * /root/reference/src/transformer/attention.py:74-86 (and the head split /
 * merge copies at :47-49,:56-57), bf16 in, fp32 accumulate.
 *
 *   q [B,Lq,Hh,D], k,v [B,Lk,Hh,D] bf16 (the natural layout of the w_qs/w_ks/w_vs
 *   outputs, attention.py:43-45; D = 64), out [B,Lq,Hh,D] bf16 (what `fc` eats).
 *   kv_len [B] i32 or NULL: keys >= kv_len[b] are masked (get_attn_pad_mask /
 *   get_attn_key_pad_mask, utils/utils.py:147-165); causal != 0 adds the
 *   subsequent mask (utils.py:136-144); dense_mask [B,Lq,Lk] u8 or NULL is the
 *   general form (non-zero = masked).  lse [B,Hh,Lq] f32 is saved for backward.
 *   A fully masked row yields NaN like the reference's softmax over all -inf.
 */
int asr_mha_fwd_bf16(const void* q, const void* k, const void* v,
                     const int* kv_len, const uint8_t* dense_mask, int causal,
                     int B, int Hh, int Lq, int Lk, int D, float scale,
                     void* out, float* lse, void* stream);
size_t asr_mha_bwd_workspace_bytes(int B, int Hh, int Lq, int Lk, int D);
int asr_mha_bwd_bf16(const void* q, const void* k, const void* v, const void* out,
                     const void* g_out, const float* lse,
                     const int* kv_len, const uint8_t* dense_mask, int causal,
                     int B, int Hh, int Lq, int Lk, int D, float scale,
                     void* g_q, void* g_k, void* g_v,
                     void* ws, size_t ws_bytes, void* stream);
/* The same pair with the reference's dropout on the probabilities
 * (`attn = self.dropout(attn)`, attention.py:83, training mode): element
 * (b,h,q,k) of softmax(S) is zeroed with probability p and the kept ones are
 * scaled by 1/(1-p) before the product with V; backward regenerates the same
 * mask from (seed, coordinates) with a counter-based generator (Philox4x32-7),
 * nothing is stored.  p is quantised to 1/256: the drop probability actually used
 * is round(256 p)/256 and the scale its exact inverse keep probability
 * (asr_mha_dropout_keep_prob), so the expectation is unbiased.  p_drop = 0 is the
 * plain call.  asr_mha_dropout_keep_u8 writes the keep mask [B,Hh,Lq,Lk] u8 for
 * the same (p_drop, seed): tests and debugging. */
int asr_mha_fwd_dropout_bf16(const void* q, const void* k, const void* v,
                             const int* kv_len, const uint8_t* dense_mask, int causal,
                             int B, int Hh, int Lq, int Lk, int D, float scale,
                             float p_drop, uint64_t seed,
                             void* out, float* lse, void* stream);
int asr_mha_bwd_dropout_bf16(const void* q, const void* k, const void* v, const void* out,
                             const void* g_out, const float* lse,
                             const int* kv_len, const uint8_t* dense_mask, int causal,
                             int B, int Hh, int Lq, int Lk, int D, float scale,
                             float p_drop, uint64_t seed,
                             void* g_q, void* g_k, void* g_v,
                             void* ws, size_t ws_bytes, void* stream);
/* The same pair with the seed read ON THE DEVICE: the mask is that of seed = *seed_dev + seed_add.  For training
 * steps captured in a CUDA graph: the host-side arguments are frozen at capture time, so every attention call of the
 * step gets its own constant `seed_add` and the step bumps the one device word `*seed_dev` (a captured device-side
 * add) to draw fresh masks on every replay.  Forward and backward of one call must use the same pair. */
int asr_mha_fwd_dropout_dev_bf16(const void* q, const void* k, const void* v,
                                 const int* kv_len, const uint8_t* dense_mask, int causal,
                                 int B, int Hh, int Lq, int Lk, int D, float scale,
                                 float p_drop, const uint64_t* seed_dev, uint64_t seed_add,
                                 void* out, float* lse, void* stream);
int asr_mha_bwd_dropout_dev_bf16(const void* q, const void* k, const void* v, const void* out,
                                 const void* g_out, const float* lse,
                                 const int* kv_len, const uint8_t* dense_mask, int causal,
                                 int B, int Hh, int Lq, int Lk, int D, float scale,
                                 float p_drop, const uint64_t* seed_dev, uint64_t seed_add,
                                 void* g_q, void* g_k, void* g_v,
                                 void* ws, size_t ws_bytes, void* stream);
float asr_mha_dropout_keep_prob(float p_drop);
int asr_mha_dropout_keep_u8(int B, int Hh, int Lq, int Lk, float p_drop, uint64_t seed,
                            uint8_t* keep, void* stream);
/* attention probabilities [Hh*B, Lq, Lk] f32 in the reference's head-major row
 * order (attention.py:47,62) - only computed when a caller asks for `attn`. */
int asr_mha_probs_f32(const void* q, const void* k,
                      const int* kv_len, const uint8_t* dense_mask, int causal,
                      int B, int Hh, int Lq, int Lk, int D, float scale,
                      float* attn, void* stream);

/* ---- gradient all-reduce over NVLink / NVSwitch peer memory ----------------------------------------
 * The data-parallel wrapper of the training loop (/root/reference/src/transformer/solver.py:141-161 is the single-GPU loop
 * it wraps) averages the fp32 gradient buckets over the ranks.  Every rank calls this on its own stream with the same
 * (offset, n, ctas); the bucket lives in symmetric memory:
 *   peer_ptrs_dev    device array [world] of the bucket's base address in every rank, as mapped in THIS process
 *   multicast_ptr    the bucket's multicast address (NVSwitch multicast object), or NULL: without it the kernel reads
 *                    and writes the peer mappings directly
 *   signal_ptrs_dev  device array [world] of every rank's flag area (uint32, zero-initialised,
 *                    asr_allreduce_signal_bytes(world, ctas) bytes), as mapped in this process
 * In place: on return (stream order) elements [offset, offset + n) of EVERY rank's bucket hold the mean over the ranks.
 * offset and n in floats, multiples of 4.  ctas: thread blocks (1..128) - the kernel is bound by the links, a few dozen
 * suffice and leave the SMs to the kernels it overlaps with. */
size_t asr_allreduce_signal_bytes(int world, int ctas);
int asr_allreduce_mean_f32(const void* peer_ptrs_dev, void* multicast_ptr, const void* signal_ptrs_dev,
                           int rank, int world, size_t offset, size_t n, int ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ASR_SM100_H */

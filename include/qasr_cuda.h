/*
 * qasr_cuda.h - C ABI of libqasr_cuda.so: the B200 (sm_100a) replacement for the hot path of
 * the reference Qwen3-ASR engine (mel -> audio encoder -> decoder prefill -> greedy decode).
 *
 * Plain C, opaque handle, host pointers and sizes only.  Every function returns 0 on success
 * and a negative code on failure (qasr_cuda_last_error() gives the text); nothing here falls
 * back to the CPU: if no CUDA device / no sm_100 kernel image is available the call fails.
 *
 * Each entry point names the reference interface it replaces (paths relative to the
 * reference root).  INTEGRATION.md shows the host-side C shim that implements the
 * reference's seven symbols on top of this header so qwen_asr.c links unchanged.
 *
 * State contract (reference qwen_asr.c:763,1823): the caller owns the KV length.  Every
 * decoder call takes `kv_len` = the number of valid cached positions BEFORE the call
 * (the reference's ctx->kv_cache_len, which callers reset to 0 per segment or roll back
 * for streaming prefix reuse).  Truncation moves no data.
 *
 * Threading and process-wide state (reference: single caller thread, global thread pool and scratch): one
 * inference process per GPU is the supported mode.  Calls on ONE context must come from one thread at a time; contexts on
 * different devices of one process are independent.  Process-wide, read once: the QASR_* environment switches,
 * the split-K scratch per device, the cuTensorMapEncodeTiled entry point; per thread: the error text.
 */
#ifndef QASR_CUDA_H
#define QASR_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qasr_ctx qasr_ctx_t;

#define QASR_OK 0
#define QASR_ERR_CUDA (-1)     /* CUDA runtime/driver failure, or no usable GPU */
#define QASR_ERR_ARG (-2)      /* bad argument */
#define QASR_ERR_MODEL (-3)    /* checkpoint missing / malformed / unsupported dtype */
#define QASR_ERR_NOMEM (-4)
#define QASR_ERR_STATE (-5)    /* call order violated (e.g. encode before load) */

#define QASR_TOKEN_ENDOFTEXT 151643 /* reference qwen_asr.h:32 */
#define QASR_TOKEN_IM_END 151645    /* reference qwen_asr.h:31 */

/* ---------------------------------------------------------------- lifecycle */

/* Number of CUDA devices visible (0 if none / driver missing). */
int qasr_cuda_device_count(void);
/* Text of the last error on the calling thread. */
const char *qasr_cuda_last_error(void);

/* Create a context on `device` (one context = one sequence = one KV cache, like
 * qwen_ctx_t, reference qwen_asr.h:195-273). Returns NULL on failure. */
qasr_ctx_t *qasr_cuda_init(int device);
void qasr_cuda_free(qasr_ctx_t *ctx);

/* Upload a safetensors checkpoint directory to HBM once (bf16 verbatim; norms/biases f32).
 * Replaces qwen_encoder_load / qwen_decoder_load (reference qwen_asr_encoder.c:67,
 * qwen_asr_decoder.c:50) and the variant probe detect_config (qwen_asr.c:135-215). */
int qasr_cuda_load_dir(qasr_ctx_t *ctx, const char *model_dir);

/* The same one-time upload from tensors the caller already holds in host memory (its own mmap of the checkpoint):
 * exactly what the reference's loaders are handed - qwen_encoder_load / qwen_decoder_load receive a
 * multi_safetensors_t (reference qwen_asr.c:125-128,243,251; qwen_asr_safetensors.h:24-49) - so the host shim can
 * forward {name, data, dtype, shape} per tensor and qwen_asr.c needs no edit.  Tensor names are the checkpoint's
 * ("thinker.audio_tower...", "thinker.model..."); tensors the path does not use are ignored.  The data is copied to
 * HBM before the call returns. */
#define QASR_DTYPE_F32 0  /* same numbering as the reference's safetensor_dtype_t, qwen_asr_safetensors.h:15-23 */
#define QASR_DTYPE_F16 1
#define QASR_DTYPE_BF16 2
typedef struct {
    const char *name;
    const void *data;
    int dtype;            /* QASR_DTYPE_* */
    int ndim;
    const int64_t *shape; /* [ndim] */
} qasr_tensor_t;
int qasr_cuda_upload_tensors(qasr_ctx_t *ctx, const qasr_tensor_t *tensors, int count);

/* out[12] = enc_d_model, enc_layers, enc_heads, enc_ffn_dim, enc_output_dim, dec_hidden,
 * dec_layers, dec_heads, dec_kv_heads, dec_head_dim, dec_intermediate, vocab_size
 * (reference qwen_config_t, qwen_asr.h:46-76). */
int qasr_cuda_config(const qasr_ctx_t *ctx, int *out12);

/* GEMM operand precision for encoder/prefill tensor-core GEMMs: 2 (default) splits every f32
 * activation into bf16 hi+lo (two MMAs, ~2^-17 relative error: the f32-activation reference
 * is reproduced to greedy-id parity); 1 uses a single bf16 activation (north_star tolerance
 * 1e-2).  Weights are exact bf16 in both. */
int qasr_cuda_set_gemm_split(qasr_ctx_t *ctx, int nsplit);

/* ------------------------------------------------- level 1: entry points */

/* Frames produced for n_samples: floor(n/160) (reference qwen_asr_audio.c:311-312). */
int qasr_cuda_mel_frames(int n_samples);

/* Log-mel front end.  Replaces qwen_mel_spectrogram (reference qwen_asr_audio.c:293-394).
 * samples: host f32 mono 16 kHz.  mel_out: host [128, frames] f32 or NULL (result always
 * stays resident on the device for qasr_cuda_encode(mel=NULL)). */
int qasr_cuda_mel(qasr_ctx_t *ctx, const float *samples, int n_samples, float *mel_out, int *out_frames);

/* Audio encoder.  Replaces qwen_encoder_forward (reference qwen_asr_encoder.c:171-372).
 * mel: host [128, mel_frames] f32, or NULL to consume the device-resident mel of the last
 * qasr_cuda_mel call.  enc_out: host [T, enc_output_dim] f32 or NULL (result stays on the
 * device as the audio rows for qasr_cuda_prefill_prompt). */
int qasr_cuda_encode(qasr_ctx_t *ctx, const float *mel, int mel_frames, float *enc_out, int *out_tokens);
/* T for a given frame count: 13 per full 100-frame chunk + conv^3 of the tail. */
int qasr_cuda_encoder_tokens(int mel_frames);

/* Decoder prefill from host embeddings [seq_len, dec_hidden] f32 appended at kv_len.
 * Replaces qwen_decoder_prefill (reference qwen_asr_decoder.c:457-563). */
int qasr_cuda_prefill_embeds(qasr_ctx_t *ctx, const float *input_embeds, int seq_len, int kv_len);

/* Decoder prefill with on-device prompt assembly: rows =
 * embed(pre_ids[0..n_pre)) | the n_audio device-resident encoder rows of the last
 * qasr_cuda_encode | embed(suf_ids[0..n_suf)); ALL rows are prefilled except the last, which
 * is kept as the pending step input (reference qwen_asr.c:685-769).  Follow with
 * qasr_cuda_step_pending(). */
int qasr_cuda_prefill_prompt(qasr_ctx_t *ctx, const int *pre_ids, int n_pre, int n_audio,
                             const int *suf_ids, int n_suf, int kv_len);

/* One decode step from a host embedding; returns the greedy token in *out_token.
 * Replaces qwen_decoder_forward (reference qwen_asr_decoder.c:592-685). */
int qasr_cuda_step_embed(qasr_ctx_t *ctx, const float *input_embed, int kv_len, int *out_token);
/* Same, but the input row is gathered on the device from the tied embedding table, so only
 * a token id crosses PCIe (reference qwen_asr.c:816-817 does the gather on the host). */
int qasr_cuda_step_token(qasr_ctx_t *ctx, int token_id, int kv_len, int *out_token);
/* Step on the pending row left by qasr_cuda_prefill_prompt. */
int qasr_cuda_step_pending(qasr_ctx_t *ctx, int kv_len, int *out_token);
/* One step that also returns full logits [vocab] f32.
 * Replaces qwen_decoder_forward_logits (reference qwen_asr_decoder.c:691-783). */
int qasr_cuda_step_logits(qasr_ctx_t *ctx, const float *input_embed, int kv_len, float *logits);

/* Greedy loop on the device (reference qwen_asr.c:788-818): starting from `first_token`
 * (already produced by a step), repeatedly feed the last token back, stop after an EOS token
 * (151643/151645) or max_new ids.  out_ids[0] = first_token.  Returns the id count in
 * *out_n; *out_kv_len receives the new KV length.  Only ids cross PCIe. */
int qasr_cuda_generate(qasr_ctx_t *ctx, int first_token, int kv_len, int max_new, int *out_ids,
                       int *out_n, int *out_kv_len);

/* Whole offline segment (reference transcribe_segment, qwen_asr.c:649-842, default prompt:
 * no system text, no forced language, no past text): samples -> greedy ids.
 * timings_ms (nullable) = {mel, encoder, prefill+first step, decode} device-timed. */
int qasr_cuda_transcribe_ids(qasr_ctx_t *ctx, const float *samples, int n_samples, int max_new,
                             int *out_ids, int *out_n, double *timings_ms, int *out_enc_tokens);

/* Prompt used by qasr_cuda_transcribe_ids / _staged / _batch around the audio rows (default: no system text, no forced
 * language).  Replaces what qwen_set_prompt / qwen_set_force_language / past-text conditioning add to the token
 * sequence (reference qwen_asr.c:388-399,685-759): pre = [151644, 8948, 198] + system-prompt tokens +
 * [151645, 198, 151644, 872, 198, 151669]; suf = [151670, 151645, 198, 151644, 77091, 198] (+ "language X" tokens +
 * 151704) (+ past-text tokens + 151704).  n_suf >= 1. */
int qasr_cuda_set_prompt(qasr_ctx_t *ctx, const int *pre_ids, int n_pre, const int *suf_ids, int n_suf);

/* Independent units (the segments of -S mode, reference qwen_asr.c:941-1103 with --past-text no, or separate utterances)
 * transcribed together.  Up to qasr_cuda_max_batch() units (4 for 0.6B, 2 for 1.7B) share the persistent decode kernel
 * (one pass over the weights per step serves all of them; front end, encoder and prefill per unit).  More units take the
 * batched throughput path: encoder, prefill and every decode step run as GEMMs over the rows of a whole group (up to 128
 * units, QASR_BATCH_MAX), attention per unit over a pooled KV cache.  samples[i] / n_samples[i] / max_new[i] describe unit
 * i; its ids go to out_ids + i*ids_stride (max_new[i] <= ids_stride), its count to out_n[i].  Ids are those of
 * qasr_cuda_transcribe_ids on the same unit.  timings_ms (nullable) = {mel, encoder, prefill, decode} totals. */
int qasr_cuda_max_batch(const qasr_ctx_t *ctx);
/* How a call with `count` units would be split: *out_groups groups, the first (largest) of *out_group_size units = the
 * sequences that share one pass over the weights in every decode step (benchmark reporting). */
int qasr_cuda_batch_plan(const qasr_ctx_t *ctx, int count, int *out_groups, int *out_group_size);
int qasr_cuda_transcribe_batch(qasr_ctx_t *ctx, const float *const *samples, const int *n_samples, int count,
                               const int *max_new, int ids_stride, int *out_ids, int *out_n, double *timings_ms);

/* Streaming session with every buffer in HBM: the device-visible part of the reference's stream_impl
 * (qwen_asr.c:1273-1900).  stream_begin resets the session (window_sec = 8, max_windows = 4 in the reference).
 * stream_feed takes ALL audio received so far: windows completed since the last call are encoded once and cached on the
 * device, the partial tail is re-encoded, the prompt (qasr_cuda_set_prompt) is assembled on the device, the rows shared
 * with the previous chunk are reused from the KV cache (*out_reused), the rest is prefilled and up to max_new greedy ids
 * are produced.  The token bookkeeping that follows in the reference (:1906-2146) stays on the host. */
int qasr_cuda_stream_begin(qasr_ctx_t *ctx, float window_sec, int max_windows);
int qasr_cuda_stream_feed(qasr_ctx_t *ctx, const float *samples, int n_samples, int max_new, int *out_ids, int *out_n,
                          int *out_reused, int *out_rows);

/* Benchmark plumbing: keep a segment's samples resident in HBM and transcribe from there (no
 * per-call host->device copy), a CUDA-event stopwatch on the library's own stream (the stream
 * every kernel here is launched on), and the accumulated device time / step count of the greedy
 * decode-step launches since the last reset (for the decode roofline). */
int qasr_cuda_stage_audio(qasr_ctx_t *ctx, const float *samples, int n_samples);
/* Interleaved 16-bit PCM (the payload of a WAV "data" chunk) at any sample rate -> f32 mono 16 kHz on the device: channel
 * average, 1/32768 scaling and the reference's windowed-sinc resampler (qwen_parse_wav_buffer, qwen_asr_audio.c:81-164;
 * the RIFF chunk walk :40-79 stays host code).  The result is staged for qasr_cuda_transcribe_staged; `out` (nullable,
 * out_cap samples) receives a host copy.  *out_n = floor(n_frames * 16000 / sample_rate). */
int qasr_cuda_decode_pcm16(qasr_ctx_t *ctx, const int16_t *pcm, int n_frames, int channels, int sample_rate, float *out,
                           int out_cap, int *out_n);
int qasr_cuda_transcribe_staged(qasr_ctx_t *ctx, int max_new, int *out_ids, int *out_n, double *timings_ms,
                                int *out_enc_tokens);
int qasr_cuda_timer_start(qasr_ctx_t *ctx);
int qasr_cuda_timer_stop(qasr_ctx_t *ctx, double *out_ms);
int qasr_cuda_decode_stats(qasr_ctx_t *ctx, long long *steps, double *ms, int reset);

/* Test hooks: copy KV rows [0,len) of one layer to the host ([len, kv_heads*head_dim] f32),
 * and the bf16 embedding row of a token upcast to f32 (reference qwen_asr.c:412-419). */
int qasr_cuda_read_kv(qasr_ctx_t *ctx, int layer, int len, float *k_out, float *v_out);
int qasr_cuda_embed_token(qasr_ctx_t *ctx, int token_id, float *out);

/* Device-time of the kernels launched by the last decode step / generate call, in ms, and
 * the number of kernel launches this context has issued (bench.py "gpu_launches"). */
double qasr_cuda_last_decode_ms(const qasr_ctx_t *ctx);
long long qasr_cuda_launch_count(const qasr_ctx_t *ctx);

/* ------------------------------------- level 2: operator surface (test seam)
 * Host-pointer twins of the reference's qwen_asr_kernels.h ops: H2D -> sm_100a kernel -> D2H.
 * Not a production path (one PCIe round trip per op); used for op-level parity tests.
 * Argument meaning is identical to the reference function named in each comment. */

/* qwen_linear / qwen_linear_nobias (b may be NULL): y[seq,out] = x[seq,in] W[out,in]^T + b.
 * reference qwen_asr_kernels.h:31-35 */
int qasr_op_linear(qasr_ctx_t *ctx, float *y, const float *x, const float *W, const float *b,
                   int seq_len, int in_dim, int out_dim);
/* qwen_matmul_t: C[M,N] = A[M,K] B[N,K]^T. reference qwen_asr_kernels.h:28 */
int qasr_op_matmul_t(qasr_ctx_t *ctx, float *C, const float *A, const float *B, int M, int K, int N);
/* qwen_linear_bf16 / qwen_linear_nobias_bf16 (b may be NULL). seq_len==1 runs the decode GEMV,
 * seq_len>1 the tcgen05 GEMM. reference qwen_asr_kernels.h:38-42 */
int qasr_op_linear_bf16(qasr_ctx_t *ctx, float *y, const float *x, const uint16_t *W_bf16, const float *b,
                        int seq_len, int in_dim, int out_dim);
/* qwen_matmul_t_bf16. reference qwen_asr_kernels.h:52 */
int qasr_op_matmul_t_bf16(qasr_ctx_t *ctx, float *C, const float *A, const uint16_t *B_bf16, int M, int K, int N);
/* qwen_linear_nobias_bf16_qkv. reference qwen_asr_kernels.h:45-50 */
int qasr_op_linear_nobias_bf16_qkv(qasr_ctx_t *ctx, float *q, float *k, float *v, const float *x,
                                   const uint16_t *Wq, const uint16_t *Wk, const uint16_t *Wv,
                                   int in_dim, int q_dim, int kv_dim);
/* qwen_argmax_matvec_bf16 (returns the index in *out_index). reference qwen_asr_kernels.h:166 */
int qasr_op_argmax_matvec_bf16(qasr_ctx_t *ctx, const float *x, const uint16_t *W_bf16, int in_dim,
                               int out_dim, int *out_index);
/* qwen_conv2d. reference qwen_asr_kernels.h:68-70 */
int qasr_op_conv2d(qasr_ctx_t *ctx, float *out, const float *in, const float *weight, const float *bias,
                   int c_in, int c_out, int h_in, int w_in, int kh, int kw, int stride, int padding);
/* qwen_layer_norm. reference qwen_asr_kernels.h:92 */
int qasr_op_layer_norm(qasr_ctx_t *ctx, float *out, const float *x, const float *weight, const float *bias,
                       int seq_len, int hidden, float eps);
/* qwen_rms_norm. reference qwen_asr_kernels.h:96 */
int qasr_op_rms_norm(qasr_ctx_t *ctx, float *out, const float *x, const float *weight, int seq_len,
                     int hidden, float eps);
/* qwen_rms_norm_per_head (in place). reference qwen_asr_kernels.h:102 */
int qasr_op_rms_norm_per_head(qasr_ctx_t *ctx, float *x, const float *weight, int seq_len, int n_heads,
                              int head_dim, float eps);
/* qwen_gelu / qwen_silu / qwen_softmax (in place). reference qwen_asr_kernels.h:109-111 */
int qasr_op_gelu(qasr_ctx_t *ctx, float *x, int n);
int qasr_op_silu(qasr_ctx_t *ctx, float *x, int n);
int qasr_op_softmax(qasr_ctx_t *ctx, float *x, int rows, int cols);
/* qwen_swiglu_multiply (out may alias gate_up). reference qwen_asr_kernels.h:113 */
int qasr_op_swiglu_multiply(qasr_ctx_t *ctx, float *out, const float *gate_up, int seq_len, int intermediate);
/* qwen_bidirectional_attention. reference qwen_asr_kernels.h:128-131 */
int qasr_op_bidirectional_attention(qasr_ctx_t *ctx, float *out, const float *Q, const float *K, const float *V,
                                    int seq, int n_heads, int head_dim, float scale, const int *window_starts,
                                    int n_windows);
/* qwen_causal_attention. reference qwen_asr_kernels.h:140-142 */
int qasr_op_causal_attention(qasr_ctx_t *ctx, float *out, const float *Q, const float *K, const float *V,
                             int seq_q, int seq_k, int n_heads, int n_kv_heads, int head_dim, float scale,
                             int q_offset);
/* qwen_sinusoidal_pe. reference qwen_asr_kernels.h:153 */
int qasr_op_sinusoidal_pe(qasr_ctx_t *ctx, float *pe, int n_pos, int d_model);
/* qwen_compute_rope_neox / qwen_apply_rope_neox. reference qwen_asr_kernels.h:160-169 */
int qasr_op_compute_rope_neox(qasr_ctx_t *ctx, float *cos_out, float *sin_out, const int *positions, int seq,
                              int head_dim, float theta);
int qasr_op_apply_rope_neox(qasr_ctx_t *ctx, float *x, const float *cos_vals, const float *sin_vals, int seq,
                            int n_heads, int head_dim);
/* qwen_add_inplace / qwen_mul_inplace / qwen_scale / qwen_copy. reference qwen_asr_kernels.h:18-21 */
int qasr_op_add_inplace(qasr_ctx_t *ctx, float *a, const float *b, int n);
int qasr_op_mul_inplace(qasr_ctx_t *ctx, float *a, const float *b, int n);
int qasr_op_scale(qasr_ctx_t *ctx, float *x, float s, int n);
int qasr_op_copy(qasr_ctx_t *ctx, float *dst, const float *src, int n);

/* qwen_set_threads / qwen_get_num_cpus (reference qwen_asr_kernels.h:176-179): kept for ABI
 * compatibility; grid scheduling replaces the pthread pool, so these do nothing. */
void qasr_set_threads(int n);
int qasr_get_num_cpus(void);

#ifdef __cplusplus
}
#endif
#endif /* QASR_CUDA_H */

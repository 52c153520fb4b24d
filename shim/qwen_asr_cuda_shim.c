/*
 * qwen_asr_cuda_shim.c - the reference's seven hot-path symbols on top of libqasr_cuda.so.
 *
 * HOST C, written against the reference's own headers (qwen_asr.h, qwen_asr_audio.h,
 * qwen_asr_safetensors.h) and include/qasr_cuda.h.  It REPLACES two translation units of the
 * reference at link time - qwen_asr_encoder.c and qwen_asr_decoder.c - and the one function
 * qwen_mel_spectrogram of qwen_asr_audio.c (that file is compiled unmodified with
 * -Dqwen_mel_spectrogram=qwen_mel_spectrogram_cpu so its WAV / resampler code stays).  qwen_asr.c,
 * main.c, the tokenizer, the safetensors reader and the kernels stay untouched and call into this
 * file exactly where they called the CPU code:
 *
 *   qwen_encoder_load / qwen_decoder_load   qwen_asr.c:243,251 (externs :125-128)
 *   qwen_mel_spectrogram                    qwen_asr.c:661,1122 (qwen_asr_audio.h:32)
 *   qwen_encoder_forward                    qwen_asr.c:671,1126 (qwen_asr.h:352)
 *   qwen_decoder_prefill                    qwen_asr.c:765,1826 (qwen_asr.h:356)
 *   qwen_decoder_forward                    qwen_asr.c:769,817,1834,1887 (qwen_asr.h:359)
 *   qwen_decoder_forward_logits             thinker sampler, qwen_asr.c:2516-2589 (qwen_asr.h:362)
 *
 * Built by oracle/Makefile (target `shim`) together with the unmodified reference sources into
 * oracle/_ref/libqasr_ref_cuda.so; tests/test_gpu_shim.py drives the reference's own
 * qwen_transcribe_audio / qwen_transcribe_stream through it.
 *
 * Conventions kept: returned buffers are malloc'd and freed by the caller; the caller
 * owns ctx->kv_cache_len (reset per segment, rolled back for streaming prefix reuse) and it is passed
 * down on every call; failures are NULL / -1 / silent / EOS-as-error / zero logits like the CPU code;
 * one inference per process (file-scope device handle, like the reference's own globals).  There is no
 * CPU fallback: without a usable B200 the loaders fail and qwen_load returns NULL.
 */
#include "qwen_asr.h"
#include "qwen_asr_audio.h"
#include "qwen_asr_safetensors.h"

#include "qasr_cuda.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static qasr_ctx_t *g_dev; /* one device context per process */

static void shim_log(const char *what) {
    if (qwen_verbose >= 1) fprintf(stderr, "qasr shim: %s: %s\n", what, qasr_cuda_last_error());
}

static void shim_release(void) {
    if (g_dev) { qasr_cuda_free(g_dev); g_dev = NULL; }
}

/* qwen_asr_encoder.c:67 - called first by qwen_load (qwen_asr.c:243).  The whole checkpoint (encoder AND decoder
 * tensors) is handed to the device here, straight from the mmap the reference already holds: every entry of the
 * multi_safetensors_t becomes one {name, data, dtype, shape} row.  The qwen_encoder_t fields stay NULL (qwen_free
 * frees NULLs, qwen_asr.c:282-330); nothing on the host reads encoder weights any more. */
int qwen_encoder_load(qwen_encoder_t *enc, multi_safetensors_t *ms, const qwen_config_t *cfg) {
    (void)enc; (void)cfg;
    static int registered;
    if (!ms) return -1;
    int count = 0;
    for (int s = 0; s < ms->num_shards; s++) count += ms->shards[s]->num_tensors;
    if (count <= 0) return -1;
    qasr_tensor_t *tab = (qasr_tensor_t *)calloc((size_t)count, sizeof(qasr_tensor_t));
    if (!tab) return -1;
    int n = 0;
    for (int s = 0; s < ms->num_shards; s++) {
        safetensors_file_t *sf = ms->shards[s];
        for (int i = 0; i < sf->num_tensors; i++) {
            const safetensor_t *t = &sf->tensors[i];
            if (t->dtype != DTYPE_F32 && t->dtype != DTYPE_F16 && t->dtype != DTYPE_BF16) continue;
            tab[n].name = t->name;
            tab[n].data = safetensors_data(sf, t);
            tab[n].dtype = (int)t->dtype;       /* QASR_DTYPE_* uses the reference's numbering */
            tab[n].ndim = t->ndim;
            tab[n].shape = t->shape;
            n++;
        }
    }
    shim_release();                             /* a second qwen_load in the same process replaces the model */
    const char *dv = getenv("QASR_DEVICE");
    g_dev = qasr_cuda_init(dv ? atoi(dv) : 0);
    int rc = g_dev ? qasr_cuda_upload_tensors(g_dev, tab, n) : -1;
    free(tab);
    if (rc != 0) {
        fprintf(stderr, "qasr shim: GPU upload failed (no CPU fallback): %s\n", qasr_cuda_last_error());
        shim_release();
        return -1;
    }
    if (!registered) { atexit(shim_release); registered = 1; }
    return 0;
}

/* qwen_asr_decoder.c:50 - the only decoder field host code still reads is the bf16 embedding table
 * (tok_embed_bf16_to_f32, qwen_asr.c:412-419,702-757,816); it keeps pointing into the mmap. */
int qwen_decoder_load(qwen_decoder_t *dec, multi_safetensors_t *ms, const qwen_config_t *cfg) {
    (void)cfg;
    safetensors_file_t *sf = NULL;
    const safetensor_t *t = multi_safetensors_find(ms, "thinker.model.embed_tokens.weight", &sf);
    if (!t || !sf || !g_dev) return -1;
    dec->tok_embeddings_bf16 = safetensors_get_bf16_direct(sf, t);
    return dec->tok_embeddings_bf16 ? 0 : -1;
}

/* qwen_asr_audio.c:293 - malloc'd [128, frames]; NULL when the audio is too short (:313-317) */
float *qwen_mel_spectrogram(const float *samples, int n_samples, int *out_frames) {
    if (!g_dev || !samples) return NULL;
    const int frames = qasr_cuda_mel_frames(n_samples);
    if (frames <= 0) return NULL;
    float *mel = (float *)malloc((size_t)128 * frames * sizeof(float));
    int got = 0;
    if (!mel || qasr_cuda_mel(g_dev, samples, n_samples, mel, &got) != 0) { shim_log("mel"); free(mel); return NULL; }
    if (out_frames) *out_frames = got;
    return mel;
}

/* qwen_asr_encoder.c:171 - malloc'd [T, enc_output_dim]; the caller frees it (qwen_asr.c:672,729) */
float *qwen_encoder_forward(qwen_ctx_t *ctx, const float *mel, int mel_frames, int *out_seq_len) {
    if (!g_dev || !ctx || !mel || mel_frames <= 0) return NULL;
    const int T = qasr_cuda_encoder_tokens(mel_frames);
    float *out = (float *)malloc((size_t)T * ctx->config.enc_output_dim * sizeof(float));
    int got = 0;
    if (!out || qasr_cuda_encode(g_dev, mel, mel_frames, out, &got) != 0) { shim_log("encoder"); free(out); return NULL; }
    if (out_seq_len) *out_seq_len = got;
    return out;
}

/* qwen_asr_decoder.c:457 - appends at ctx->kv_cache_len; void and silent on failure (:471-477) */
void qwen_decoder_prefill(qwen_ctx_t *ctx, const float *input_embeds, int seq_len) {
    if (!g_dev || !ctx || !input_embeds || seq_len <= 0) return;
    if (qasr_cuda_prefill_embeds(g_dev, input_embeds, seq_len, ctx->kv_cache_len) == 0) ctx->kv_cache_len += seq_len;
    else shim_log("prefill");
}

/* qwen_asr_decoder.c:592 - returns the greedy token; QWEN_TOKEN_IM_END on internal failure (:621,625) */
int qwen_decoder_forward(qwen_ctx_t *ctx, const float *input_embed) {
    int tok = QWEN_TOKEN_IM_END;
    if (!g_dev || !ctx || !input_embed) return QWEN_TOKEN_IM_END;
    if (qasr_cuda_step_embed(g_dev, input_embed, ctx->kv_cache_len, &tok) != 0) { shim_log("decode step"); return QWEN_TOKEN_IM_END; }
    ctx->kv_cache_len += 1;
    return tok;
}

/* qwen_asr_decoder.c:691 - full logits [vocab]; zero-filled on failure (:721,727) */
void qwen_decoder_forward_logits(qwen_ctx_t *ctx, const float *input_embed, float *logits) {
    if (!ctx || !logits) return;
    if (!g_dev || !input_embed || qasr_cuda_step_logits(g_dev, input_embed, ctx->kv_cache_len, logits) != 0) {
        shim_log("logits step");
        memset(logits, 0, (size_t)ctx->config.vocab_size * sizeof(float));
        return;
    }
    ctx->kv_cache_len += 1;
}

/* qwen_asr_decoder.c:321 - page-in of the 30B-MoE expert weights (qwen_asr.c:368-372, `--moe-preload`).  MoE checkpoints
 * are outside this path (the loaders above reject them), so there is nothing to preload. */
void qwen_decoder_moe_preload(qwen_decoder_t *dec, const qwen_config_t *cfg) { (void)dec; (void)cfg; }

/*
 * ref_harness.c - thin driver around the UNMODIFIED reference sources.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is compiled together with the
 * reference's own .c files *where they lie* under /root/reference (see
 * oracle/Makefile) into oracle/_ref/libqasr_ref_*.so.  It adds no arithmetic:
 * every function forwards to a reference entry point so tests and the
 * bench's CPU arm can call them through ctypes:
 *
 *   qwen_load / qwen_free                 qwen_asr.c:221,282
 *   qwen_mel_spectrogram                  qwen_asr_audio.c:293
 *   qwen_encoder_forward                  qwen_asr_encoder.c:171
 *   qwen_decoder_prefill                  qwen_asr_decoder.c:457
 *   qwen_decoder_forward                  qwen_asr_decoder.c:592
 *   qwen_decoder_forward_logits           qwen_asr_decoder.c:691
 *   qwen_transcribe_audio / _stream       qwen_asr.c:900,2148 (ref_transcribe_text: sets the public
 *                                         ctx fields the CLI sets, main.c:262-300, then forwards)
 *
 * The same file is linked into oracle/_ref/libqasr_ref_cuda.so, where the reference's encoder /
 * decoder / mel translation units are replaced by shim/qwen_asr_cuda_shim.c (the drop-in test).
 *
 * ref_transcribe_ids() restates only the *driver* of transcribe_segment
 * (qwen_asr.c:649-818: prompt layout, prefill of total_seq-1 rows, greedy
 * loop, EOS stop) with an explicit max_new_tokens cap, because the CLI
 * hard-codes 2048 and random-init weights never emit EOS.
 */
#include "qwen_asr.h"
#include "qwen_asr_audio.h"
#include "qwen_asr_kernels.h"

#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#ifdef USE_OPENBLAS
extern void openblas_set_num_threads(int n);
#endif

static double now_ms(void) {
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec * 1000.0 + tv.tv_usec / 1000.0;
}

void *ref_load(const char *model_dir, int threads) {
    qwen_verbose = 0;
    if (threads <= 0) threads = qwen_get_num_cpus();
    if (threads > 16) threads = 16; /* QWEN_MAX_THREADS, qwen_asr_kernels.c:38 */
    qwen_set_threads(threads);
#ifdef USE_OPENBLAS
    openblas_set_num_threads(threads);
#endif
    return qwen_load(model_dir);
}

void ref_free(void *c) { qwen_free((qwen_ctx_t *)c); }

int ref_threads_used(int threads) {
    if (threads <= 0) threads = qwen_get_num_cpus();
    return threads > 16 ? 16 : threads;
}

void ref_config(void *c, int *out) {
    const qwen_config_t *cfg = &((qwen_ctx_t *)c)->config;
    out[0] = cfg->enc_d_model;  out[1] = cfg->enc_layers;   out[2] = cfg->enc_heads;
    out[3] = cfg->enc_ffn_dim;  out[4] = cfg->enc_output_dim;
    out[5] = cfg->dec_hidden;   out[6] = cfg->dec_layers;   out[7] = cfg->dec_heads;
    out[8] = cfg->dec_kv_heads; out[9] = cfg->dec_head_dim; out[10] = cfg->dec_intermediate;
    out[11] = cfg->vocab_size;
}

float *ref_mel(const float *samples, int n, int *frames) { return qwen_mel_spectrogram(samples, n, frames); }
float *ref_encode(void *c, const float *mel, int frames, int *T) {
    return qwen_encoder_forward((qwen_ctx_t *)c, mel, frames, T);
}
void ref_free_buf(void *p) { free(p); }
float *ref_parse_wav(const unsigned char *data, size_t size, int *n) { return qwen_parse_wav_buffer(data, size, n); }

void ref_set_kv_len(void *c, int n) { ((qwen_ctx_t *)c)->kv_cache_len = n; }
int ref_get_kv_len(void *c) { return ((qwen_ctx_t *)c)->kv_cache_len; }

void ref_prefill(void *c, const float *embeds, int seq) { qwen_decoder_prefill((qwen_ctx_t *)c, embeds, seq); }
int ref_step(void *c, const float *embed) { return qwen_decoder_forward((qwen_ctx_t *)c, embed); }
void ref_step_logits(void *c, const float *embed, float *logits) {
    qwen_decoder_forward_logits((qwen_ctx_t *)c, embed, logits);
}

void ref_embed_token(void *c, int tok, float *dst) {
    qwen_ctx_t *ctx = (qwen_ctx_t *)c;
    int dim = ctx->config.dec_hidden;
    const uint16_t *src = ctx->decoder.tok_embeddings_bf16 + (size_t)tok * dim;
    for (int i = 0; i < dim; i++) {
        uint32_t b = ((uint32_t)src[i]) << 16;
        memcpy(&dst[i], &b, 4);
    }
}

/* Copy KV rows [0,len) of one layer (for cache-level parity checks). */
void ref_read_kv(void *c, int layer, int len, float *k_out, float *v_out) {
    qwen_ctx_t *ctx = (qwen_ctx_t *)c;
    int kv_dim = ctx->config.dec_kv_heads * ctx->config.dec_head_dim;
    size_t base = (size_t)layer * ctx->kv_cache_max * kv_dim;
    memcpy(k_out, ctx->kv_cache_k + base, (size_t)len * kv_dim * sizeof(float));
    memcpy(v_out, ctx->kv_cache_v + base, (size_t)len * kv_dim * sizeof(float));
}

/*
 * Offline single-segment greedy transcription to token ids.
 * prompt layout: qwen_asr.c:388-399,685-759 (no system prompt, no forced
 * language, no past text).  timings_ms = {mel, encoder, prefill(+first tok), decode}.
 * Returns number of generated ids (including a terminating EOS if hit).
 */
int ref_transcribe_ids(void *c, const float *samples, int n_samples, int max_new,
                       int *out_ids, double *timings_ms, int *out_enc_tokens) {
    static const int PRE[] = {151644, 8948, 198, 151645, 198, 151644, 872, 198, 151669};
    static const int SUF[] = {151670, 151645, 198, 151644, 77091, 198};
    qwen_ctx_t *ctx = (qwen_ctx_t *)c;
    int dim = ctx->config.dec_hidden;

    double t0 = now_ms();
    int frames = 0, T = 0;
    float *mel = qwen_mel_spectrogram(samples, n_samples, &frames);
    if (!mel) return -1;
    double t1 = now_ms();
    float *enc = qwen_encoder_forward(ctx, mel, frames, &T);
    free(mel);
    if (!enc) return -1;
    double t2 = now_ms();

    int total = 9 + T + 6;
    float *emb = (float *)malloc((size_t)total * dim * sizeof(float));
    float *tmp = (float *)malloc((size_t)dim * sizeof(float));
    for (int i = 0; i < 9; i++) ref_embed_token(c, PRE[i], emb + (size_t)i * dim);
    memcpy(emb + (size_t)9 * dim, enc, (size_t)T * dim * sizeof(float));
    for (int i = 0; i < 6; i++) ref_embed_token(c, SUF[i], emb + (size_t)(9 + T + i) * dim);
    free(enc);

    ctx->kv_cache_len = 0;
    qwen_decoder_prefill(ctx, emb, total - 1);
    int tok = qwen_decoder_forward(ctx, emb + (size_t)(total - 1) * dim);
    free(emb);
    double t3 = now_ms();

    int n = 0;
    while (n < max_new) {
        out_ids[n++] = tok;
        if (tok == QWEN_TOKEN_ENDOFTEXT || tok == QWEN_TOKEN_IM_END) break;
        if (n >= max_new) break;
        ref_embed_token(c, tok, tmp);
        tok = qwen_decoder_forward(ctx, tmp);
    }
    double t4 = now_ms();
    free(tmp);
    if (timings_ms) {
        timings_ms[0] = t1 - t0; timings_ms[1] = t2 - t1;
        timings_ms[2] = t3 - t2; timings_ms[3] = t4 - t3;
    }
    if (out_enc_tokens) *out_enc_tokens = T;
    return n;
}

/*
 * The reference's own top-level entry points, unmodified: offline (-S segment_sec, -W search_sec) or
 * --stream, optional forced language (its tokens + <asr_text> join the prompt, qwen_asr.c:581-603, so
 * text is emitted from the first token even with random-init weights).  Returns the malloc'd text.
 */
char *ref_transcribe_text(void *c, const float *samples, int n_samples, float segment_sec, float search_sec,
                          int stream, const char *language) {
    qwen_ctx_t *ctx = (qwen_ctx_t *)c;
    ctx->segment_sec = segment_sec;
    ctx->search_sec = search_sec;
    if (qwen_set_force_language(ctx, language) != 0) return NULL;
    return stream ? qwen_transcribe_stream(ctx, samples, n_samples) : qwen_transcribe_audio(ctx, samples, n_samples);
}

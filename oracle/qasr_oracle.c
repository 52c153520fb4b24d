/*
 * qasr_oracle.c - plain-C CPU restatement of the reference's Qwen3-ASR hot path
 * (mel -> encoder -> prefill -> greedy decode) and of its operator surface.
 *
 * TEST INFRASTRUCTURE ONLY (see qasr_oracle.h).  Every function cites the reference
 * file:line whose arithmetic it follows (paths relative to /root/reference).  The code
 * is written from the numerics contract, not transcribed: plain
 * f32 loops + OpenMP, no BLAS, no intrinsics, no -ffast-math.  Summation order therefore
 * differs from the reference build (OpenBLAS / AVX lanes), so float outputs agree to
 * rounding (tests state the tolerance) while greedy ids agree exactly.
 */
#include "qasr_oracle.h"

#include "../smol-vision_b200/csrc/qasr_safetensors.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static inline float bf16_f32(uint16_t b) {
    uint32_t u = ((uint32_t)b) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

static inline float dotf(const float *a, const float *b, int n) {
    float s = 0.0f;
#pragma omp simd reduction(+ : s)
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

/* ======================================================================
 * Level 2: operator surface
 * ====================================================================== */

/* y = x W^T + b.  reference qwen_asr_kernels.c:196-224 (cblas_sgemm + bias loop). */
void qo_linear(float *y, const float *x, const float *W, const float *b, int seq, int in_dim, int out_dim) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int s = 0; s < seq; s++)
        for (int o = 0; o < out_dim; o++)
            y[(size_t)s * out_dim + o] = dotf(x + (size_t)s * in_dim, W + (size_t)o * in_dim, in_dim) + (b ? b[o] : 0.0f);
}

/* bf16 weights are upcast exactly (bits<<16) then used as f32.
 * reference qwen_asr_kernels.c:232-236,462-484 and qwen_asr_kernels_generic.c:9-22. */
void qo_linear_bf16(float *y, const float *x, const uint16_t *W, const float *b, int seq, int in_dim, int out_dim) {
#pragma omp parallel
    {
        float *row = (float *)malloc((size_t)in_dim * sizeof(float));
#pragma omp for schedule(static)
        for (int o = 0; o < out_dim; o++) {
            const uint16_t *w = W + (size_t)o * in_dim;
            for (int k = 0; k < in_dim; k++) row[k] = bf16_f32(w[k]);
            for (int s = 0; s < seq; s++)
                y[(size_t)s * out_dim + o] = dotf(x + (size_t)s * in_dim, row, in_dim) + (b ? b[o] : 0.0f);
        }
        free(row);
    }
}

/* 2-D convolution, zero padding. reference qwen_asr_kernels.c:566-590 (im2col), 643-685. */
void qo_conv2d(float *out, const float *in, const float *w, const float *bias, int c_in, int c_out,
               int h_in, int w_in, int kh, int kw, int stride, int padding) {
    int h_out = (h_in + 2 * padding - kh) / stride + 1;
    int w_out = (w_in + 2 * padding - kw) / stride + 1;
    int K = c_in * kh * kw, S = h_out * w_out;
    /* patches[pos][K]: one contiguous receptive field per output position */
    float *patches = (float *)malloc((size_t)S * K * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int p = 0; p < S; p++) {
        int oh = p / w_out, ow = p % w_out;
        float *dst = patches + (size_t)p * K;
        for (int ic = 0; ic < c_in; ic++)
            for (int ki = 0; ki < kh; ki++)
                for (int kj = 0; kj < kw; kj++) {
                    int ih = oh * stride - padding + ki, iw = ow * stride - padding + kj;
                    float v = 0.0f;
                    if (ih >= 0 && ih < h_in && iw >= 0 && iw < w_in) v = in[((size_t)ic * h_in + ih) * w_in + iw];
                    *dst++ = v;
                }
    }
#pragma omp parallel for collapse(2) schedule(static)
    for (int oc = 0; oc < c_out; oc++)
        for (int p = 0; p < S; p++)
            out[(size_t)oc * S + p] = dotf(w + (size_t)oc * K, patches + (size_t)p * K, K) + (bias ? bias[oc] : 0.0f);
    free(patches);
}

/* LayerNorm, biased variance. reference qwen_asr_kernels.c:691-799. */
void qo_layer_norm(float *out, const float *x, const float *w, const float *b, int seq, int hidden, float eps) {
    for (int s = 0; s < seq; s++) {
        const float *xr = x + (size_t)s * hidden;
        float *orow = out + (size_t)s * hidden;
        float mean = 0.0f;
        for (int i = 0; i < hidden; i++) mean += xr[i];
        mean /= hidden;
        float var = 0.0f;
        for (int i = 0; i < hidden; i++) { float d = xr[i] - mean; var += d * d; }
        var /= hidden;
        float inv = 1.0f / sqrtf(var + eps);
        for (int i = 0; i < hidden; i++) orow[i] = (xr[i] - mean) * inv * w[i] + b[i];
    }
}

/* RMSNorm. reference qwen_asr_kernels.c:801-860. */
void qo_rms_norm(float *out, const float *x, const float *w, int seq, int hidden, float eps) {
    for (int s = 0; s < seq; s++) {
        const float *xr = x + (size_t)s * hidden;
        float *orow = out + (size_t)s * hidden;
        float ss = 0.0f;
        for (int i = 0; i < hidden; i++) ss += xr[i] * xr[i];
        float inv = 1.0f / sqrtf(ss / hidden + eps);
        for (int i = 0; i < hidden; i++) orow[i] = xr[i] * inv * w[i];
    }
}

/* Per-head RMSNorm, shared [head_dim] weight. reference qwen_asr_kernels.c:862-924. */
void qo_rms_norm_per_head(float *x, const float *w, int seq, int n_heads, int head_dim, float eps) {
    for (int s = 0; s < seq; s++)
        for (int h = 0; h < n_heads; h++) {
            float *v = x + ((size_t)s * n_heads + h) * head_dim;
            float ss = 0.0f;
            for (int d = 0; d < head_dim; d++) ss += v[d] * v[d];
            float inv = 1.0f / sqrtf(ss / head_dim + eps);
            for (int d = 0; d < head_dim; d++) v[d] = v[d] * inv * w[d];
        }
}

/* tanh-approximation GELU. reference qwen_asr_kernels.c:937-944. */
void qo_gelu(float *x, int n) {
    for (int i = 0; i < n; i++) {
        float v = x[i];
        x[i] = 0.5f * v * (1.0f + tanhf(0.7978845608028654f * (v + 0.044715f * v * v * v)));
    }
}

/* reference qwen_asr_kernels.c:930-935 */
void qo_silu(float *x, int n) {
    for (int i = 0; i < n; i++) x[i] = x[i] / (1.0f + expf(-x[i]));
}

/* reference qwen_asr_kernels.c:1012-1029 */
void qo_softmax(float *x, int rows, int cols) {
    for (int r = 0; r < rows; r++) {
        float *row = x + (size_t)r * cols;
        float mx = row[0];
        for (int c = 1; c < cols; c++) if (row[c] > mx) mx = row[c];
        float sum = 0.0f;
        for (int c = 0; c < cols; c++) { row[c] = expf(row[c] - mx); sum += row[c]; }
        float inv = 1.0f / sum;
        for (int c = 0; c < cols; c++) row[c] *= inv;
    }
}

/* silu(g)*u over interleaved [g0,u0,g1,u1,...]; in-place allowed. reference qwen_asr_kernels.c:946-1010. */
void qo_swiglu_multiply(float *out, const float *gate_up, int seq, int inter) {
    for (int s = 0; s < seq; s++) {
        const float *gu = gate_up + (size_t)s * 2 * inter;
        float *o = out + (size_t)s * inter;
        for (int j = 0; j < inter; j++) {
            float g = gu[2 * j], u = gu[2 * j + 1];
            o[j] = (g / (1.0f + expf(-g))) * u;
        }
    }
}

/* One query row of online-softmax attention over keys [k0,k1).
 * reference qwen_asr_kernels.c:1069-1095 / 1120-1145 (max starts at -1e30). */
static void attend_row(float *o, const float *q, const float *K, const float *V, int k0, int k1, int kv_stride,
                       int head_dim, float scale) {
    float mx = -1e30f, sum = 0.0f;
    for (int d = 0; d < head_dim; d++) o[d] = 0.0f;
    for (int j = k0; j < k1; j++) {
        const float *kr = K + (size_t)j * kv_stride, *vr = V + (size_t)j * kv_stride;
        float sc = dotf(q, kr, head_dim) * scale;
        if (sc > mx) {
            float c = expf(mx - sc);
            sum = sum * c + 1.0f;
            for (int d = 0; d < head_dim; d++) o[d] = o[d] * c + vr[d];
            mx = sc;
        } else {
            float wt = expf(sc - mx);
            sum += wt;
            for (int d = 0; d < head_dim; d++) o[d] += wt * vr[d];
        }
    }
    if (sum > 0.0f) {
        float inv = 1.0f / sum;
        for (int d = 0; d < head_dim; d++) o[d] *= inv;
    }
}

/* Block-diagonal bidirectional MHA. reference qwen_asr_kernels.c:1054-1099. */
void qo_bidirectional_attention(float *out, const float *Q, const float *K, const float *V, int seq,
                                int n_heads, int head_dim, float scale, const int *window_starts, int n_windows) {
    (void)seq;
    int hidden = n_heads * head_dim;
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int h = 0; h < n_heads; h++)
        for (int w = 0; w < n_windows; w++) {
            int ws = window_starts[w], we = window_starts[w + 1];
            for (int i = ws; i < we; i++)
                attend_row(out + (size_t)i * hidden + h * head_dim, Q + (size_t)i * hidden + h * head_dim,
                           K + h * head_dim, V + h * head_dim, ws, we, hidden, head_dim, scale);
        }
}

/* Causal GQA attention. reference qwen_asr_kernels.c:1101-1148. */
void qo_causal_attention(float *out, const float *Q, const float *K, const float *V, int seq_q, int seq_k,
                         int n_heads, int n_kv_heads, int head_dim, float scale, int q_offset) {
    int per = n_heads / n_kv_heads, qh = n_heads * head_dim, kvh = n_kv_heads * head_dim;
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int h = 0; h < n_heads; h++)
        for (int i = 0; i < seq_q; i++) {
            int kv = h / per, k_end = q_offset + i + 1;
            if (k_end > seq_k) k_end = seq_k;
            attend_row(out + (size_t)i * qh + h * head_dim, Q + (size_t)i * qh + h * head_dim, K + kv * head_dim,
                       V + kv * head_dim, 0, k_end, kvh, head_dim, scale);
        }
}

/* reference qwen_asr_kernels.c:1198-1211 */
void qo_sinusoidal_pe(float *pe, int n_pos, int d_model) {
    int half = d_model / 2;
    float lt = logf(10000.0f) / (float)(half - 1);
    for (int p = 0; p < n_pos; p++)
        for (int d = 0; d < half; d++) {
            float ang = (float)p * expf(-(float)d * lt);
            pe[(size_t)p * d_model + d] = sinf(ang);
            pe[(size_t)p * d_model + half + d] = cosf(ang);
        }
}

/* reference qwen_asr_kernels.c:1213-1231 (and decoder cache qwen_asr_decoder.c:253-302) */
void qo_compute_rope_neox(float *cos_out, float *sin_out, const int *positions, int seq, int head_dim, float theta) {
    int half = head_dim / 2;
    for (int s = 0; s < seq; s++)
        for (int d = 0; d < half; d++) {
            float freq = 1.0f / powf(theta, (float)(2 * d) / (float)head_dim);
            float ang = (float)positions[s] * freq;
            float c = cosf(ang), sn = sinf(ang);
            cos_out[(size_t)s * head_dim + d] = cos_out[(size_t)s * head_dim + half + d] = c;
            sin_out[(size_t)s * head_dim + d] = sin_out[(size_t)s * head_dim + half + d] = sn;
        }
}

/* reference qwen_asr_kernels.c:1233-1298 */
void qo_apply_rope_neox(float *x, const float *cos_v, const float *sin_v, int seq, int n_heads, int head_dim) {
    int half = head_dim / 2;
    for (int s = 0; s < seq; s++)
        for (int h = 0; h < n_heads; h++) {
            float *v = x + ((size_t)s * n_heads + h) * head_dim;
            const float *c = cos_v + (size_t)s * head_dim, *sn = sin_v + (size_t)s * head_dim;
            for (int d = 0; d < half; d++) {
                float x1 = v[d], x2 = v[half + d];
                v[d] = x1 * c[d] - x2 * sn[d];
                v[half + d] = x2 * c[half + d] + x1 * sn[half + d];
            }
        }
}

/* Greedy head without logits; strict '>' from -1e30 => lowest index wins ties.
 * reference qwen_asr_kernels.c:486-543, qwen_asr_kernels_generic.c:24-47. */
int qo_argmax_matvec_bf16(const float *x, const uint16_t *W, int in_dim, int out_dim) {
    int best = 0;
    float best_v = -1e30f;
#pragma omp parallel
    {
        int lb = -1;
        float lv = -1e30f;
#pragma omp for schedule(static) nowait
        for (int o = 0; o < out_dim; o++) {
            const uint16_t *w = W + (size_t)o * in_dim;
            float s = 0.0f;
#pragma omp simd reduction(+ : s)
            for (int k = 0; k < in_dim; k++) s += bf16_f32(w[k]) * x[k];
            if (s > lv) { lv = s; lb = o; }
        }
#pragma omp critical
        {
            if (lb >= 0 && (lv > best_v || (lv == best_v && lb < best))) { best_v = lv; best = lb; }
        }
    }
    return best;
}

/* ======================================================================
 * Level 1: mel front end.  reference qwen_asr_audio.c:236-394
 * ====================================================================== */
#define N_FFT 400
#define N_FREQ 201
#define N_MEL 128
#define HOP 160

static float hz_to_mel(float f) { /* qwen_asr_audio.c:236-243 */
    if (f >= 1000.0f) return 15.0f + logf(f / 1000.0f) * (27.0f / logf(6.4f));
    return 3.0f * f / 200.0f;
}
static float mel_to_hz(float m) { /* qwen_asr_audio.c:245-252 */
    if (m >= 15.0f) return 1000.0f * expf((logf(6.4f) / 27.0f) * (m - 15.0f));
    return 200.0f * m / 3.0f;
}

static void mel_filterbank(float *fb /* [128][201] */) { /* qwen_asr_audio.c:254-287 */
    float pts[N_MEL + 2];
    float mmax = hz_to_mel(8000.0f), mmin = hz_to_mel(0.0f);
    for (int i = 0; i < N_MEL + 2; i++) pts[i] = mel_to_hz(mmin + (mmax - mmin) * (float)i / (float)(N_MEL + 1));
    for (int m = 0; m < N_MEL; m++) {
        float dl = pts[m + 1] - pts[m], dr = pts[m + 2] - pts[m + 1];
        if (dl == 0.0f) dl = 1e-6f;
        if (dr == 0.0f) dr = 1e-6f;
        float en = 2.0f / (pts[m + 2] - pts[m]);
        for (int f = 0; f < N_FREQ; f++) {
            float hz = (float)f * 8000.0f / (float)(N_FREQ - 1);
            float v = fminf((hz - pts[m]) / dl, (pts[m + 2] - hz) / dr);
            fb[m * N_FREQ + f] = (v < 0.0f ? 0.0f : v) * en;
        }
    }
}

float *qo_mel_spectrogram(const float *samples, int n, int *out_frames) {
    int pad = N_FFT / 2, plen = n + 2 * pad;
    int frames = (plen - N_FFT) / HOP + 1 - 1; /* last frame dropped, :311-312 */
    if (frames <= 0) return NULL;
    float *x = (float *)malloc((size_t)plen * sizeof(float));
    for (int i = 0; i < pad; i++) { /* reflect pad, :301-309 */
        int l = pad - i, r = n - 2 - i;
        x[i] = l < n ? samples[l] : 0.0f;
        x[pad + n + i] = r >= 0 ? samples[r] : 0.0f;
    }
    memcpy(x + pad, samples, (size_t)n * sizeof(float));

    float *fb = (float *)malloc(sizeof(float) * N_MEL * N_FREQ);
    mel_filterbank(fb);
    float win[N_FFT];
    for (int i = 0; i < N_FFT; i++) win[i] = 0.5f * (1.0f - cosf(2.0f * (float)M_PI * (float)i / (float)N_FFT));
    float *ct = (float *)malloc(sizeof(float) * N_FREQ * N_FFT), *st = (float *)malloc(sizeof(float) * N_FREQ * N_FFT);
    for (int k = 0; k < N_FREQ; k++)
        for (int j = 0; j < N_FFT; j++) { /* f32 angle, :330-335 */
            float ang = 2.0f * (float)M_PI * (float)k * (float)j / (float)N_FFT;
            ct[k * N_FFT + j] = cosf(ang);
            st[k * N_FFT + j] = sinf(ang);
        }
    float *lm = (float *)malloc((size_t)frames * N_MEL * sizeof(float));
    float gmax = -1e30f;
#pragma omp parallel
    {
        float fr[N_FFT], pw[N_FREQ], lmax = -1e30f;
#pragma omp for schedule(static)
        for (int t = 0; t < frames; t++) {
            for (int i = 0; i < N_FFT; i++) fr[i] = x[t * HOP + i] * win[i];
            for (int k = 0; k < N_FREQ; k++) {
                float re = dotf(fr, ct + k * N_FFT, N_FFT), im = dotf(fr, st + k * N_FFT, N_FFT);
                pw[k] = re * re + im * im;
            }
            for (int m = 0; m < N_MEL; m++) {
                float s = dotf(fb + m * N_FREQ, pw, N_FREQ);
                if (s < 1e-10f) s = 1e-10f;
                float v = log10f(s);
                lm[(size_t)t * N_MEL + m] = v;
                if (v > lmax) lmax = v;
            }
        }
#pragma omp critical
        if (lmax > gmax) gmax = lmax;
    }
    float *mel = (float *)malloc((size_t)N_MEL * frames * sizeof(float));
    float lo = gmax - 8.0f; /* dynamic-max clamp, :375-383 */
    for (int t = 0; t < frames; t++)
        for (int m = 0; m < N_MEL; m++) {
            float v = lm[(size_t)t * N_MEL + m];
            if (v < lo) v = lo;
            mel[(size_t)m * frames + t] = (v + 4.0f) / 4.0f;
        }
    free(x); free(fb); free(ct); free(st); free(lm);
    *out_frames = frames;
    return mel;
}

/* ======================================================================
 * Model container + loaders.  reference qwen_asr.c:135-215 (variant probe, hard-coded
 * dims), qwen_asr_encoder.c:67-165, qwen_asr_decoder.c:50-162.
 * ====================================================================== */
#define MAX_ENC 32
#define MAX_DEC 48

typedef struct {
    float *wq, *bq, *wk, *bk, *wv, *bv, *wo, *bo, *ln1w, *ln1b, *fc1w, *fc1b, *fc2w, *fc2b, *ln2w, *ln2b;
} enc_layer_t;
typedef struct {
    const uint16_t *wq, *wk, *wv, *wo, *down;
    uint16_t *gate_up; /* interleaved rows, qwen_asr_decoder.c:140-152 */
    float *qn, *kn, *in_norm, *post_norm;
} dec_layer_t;

struct qo_model {
    qst_dir_t *st;
    int d, enc_layers, enc_heads, F, H, dec_layers, heads, kv_heads, hd, I, V;
    float *c1w, *c1b, *c2w, *c2b, *c3w, *c3b, *conv_out, *lnpw, *lnpb, *p1w, *p1b, *p2w, *p2b;
    enc_layer_t enc[MAX_ENC];
    const uint16_t *emb;
    dec_layer_t dec[MAX_DEC];
    float *final_norm;
    float *kv_k, *kv_v;
    int kv_len, kv_max;
};

static float *get_f32(qo_model_t *m, const char *name) {
    const qst_tensor_t *t = qst_find(m->st, name);
    if (!t) { fprintf(stderr, "oracle: missing tensor %s\n", name); return NULL; }
    float *o = (float *)malloc(t->numel * sizeof(float));
    if (t->dtype == QST_BF16) {
        const uint16_t *s = (const uint16_t *)t->data;
        for (size_t i = 0; i < t->numel; i++) o[i] = bf16_f32(s[i]);
    } else if (t->dtype == QST_F32) {
        memcpy(o, t->data, t->numel * sizeof(float));
    } else { free(o); return NULL; }
    return o;
}
static const uint16_t *get_bf16(qo_model_t *m, const char *name) {
    const qst_tensor_t *t = qst_find(m->st, name);
    if (!t || t->dtype != QST_BF16) { fprintf(stderr, "oracle: missing bf16 tensor %s\n", name); return NULL; }
    return (const uint16_t *)t->data;
}

qo_model_t *qo_load(const char *dir) {
    qo_model_t *m = (qo_model_t *)calloc(1, sizeof(*m));
    m->st = qst_open_dir(dir);
    if (!m->st) { free(m); return NULL; }
    int big = qst_find(m->st, "thinker.audio_tower.layers.18.self_attn.q_proj.weight") != NULL;
    m->d = big ? 1024 : 896; m->enc_layers = big ? 24 : 18; m->enc_heads = big ? 16 : 14; m->F = big ? 4096 : 3584;
    m->H = big ? 2048 : 1024; m->dec_layers = 28; m->heads = 16; m->kv_heads = 8; m->hd = 128;
    m->I = big ? 6144 : 3072; m->V = 151936;
    char n[256];
#define E "thinker.audio_tower."
    m->c1w = get_f32(m, E "conv2d1.weight"); m->c1b = get_f32(m, E "conv2d1.bias");
    m->c2w = get_f32(m, E "conv2d2.weight"); m->c2b = get_f32(m, E "conv2d2.bias");
    m->c3w = get_f32(m, E "conv2d3.weight"); m->c3b = get_f32(m, E "conv2d3.bias");
    m->conv_out = get_f32(m, E "conv_out.weight");
    for (int i = 0; i < m->enc_layers; i++) {
        enc_layer_t *l = &m->enc[i];
#define G(field, suffix) snprintf(n, sizeof n, E "layers.%d." suffix, i); l->field = get_f32(m, n)
        G(wq, "self_attn.q_proj.weight"); G(bq, "self_attn.q_proj.bias");
        G(wk, "self_attn.k_proj.weight"); G(bk, "self_attn.k_proj.bias");
        G(wv, "self_attn.v_proj.weight"); G(bv, "self_attn.v_proj.bias");
        G(wo, "self_attn.out_proj.weight"); G(bo, "self_attn.out_proj.bias");
        G(ln1w, "self_attn_layer_norm.weight"); G(ln1b, "self_attn_layer_norm.bias");
        G(fc1w, "fc1.weight"); G(fc1b, "fc1.bias"); G(fc2w, "fc2.weight"); G(fc2b, "fc2.bias");
        G(ln2w, "final_layer_norm.weight"); G(ln2b, "final_layer_norm.bias");
#undef G
    }
    m->lnpw = get_f32(m, E "ln_post.weight"); m->lnpb = get_f32(m, E "ln_post.bias");
    m->p1w = get_f32(m, E "proj1.weight"); m->p1b = get_f32(m, E "proj1.bias");
    m->p2w = get_f32(m, E "proj2.weight"); m->p2b = get_f32(m, E "proj2.bias");
#undef E
    m->emb = get_bf16(m, "thinker.model.embed_tokens.weight");
    for (int i = 0; i < m->dec_layers; i++) {
        dec_layer_t *l = &m->dec[i];
#define P "thinker.model.layers.%d."
        snprintf(n, sizeof n, P "self_attn.q_proj.weight", i); l->wq = get_bf16(m, n);
        snprintf(n, sizeof n, P "self_attn.k_proj.weight", i); l->wk = get_bf16(m, n);
        snprintf(n, sizeof n, P "self_attn.v_proj.weight", i); l->wv = get_bf16(m, n);
        snprintf(n, sizeof n, P "self_attn.o_proj.weight", i); l->wo = get_bf16(m, n);
        snprintf(n, sizeof n, P "self_attn.q_norm.weight", i); l->qn = get_f32(m, n);
        snprintf(n, sizeof n, P "self_attn.k_norm.weight", i); l->kn = get_f32(m, n);
        snprintf(n, sizeof n, P "input_layernorm.weight", i); l->in_norm = get_f32(m, n);
        snprintf(n, sizeof n, P "post_attention_layernorm.weight", i); l->post_norm = get_f32(m, n);
        snprintf(n, sizeof n, P "mlp.down_proj.weight", i); l->down = get_bf16(m, n);
        snprintf(n, sizeof n, P "mlp.gate_proj.weight", i); const uint16_t *g = get_bf16(m, n);
        snprintf(n, sizeof n, P "mlp.up_proj.weight", i); const uint16_t *u = get_bf16(m, n);
#undef P
        if (!l->wq || !l->wk || !l->wv || !l->wo || !l->down || !g || !u) { qo_free(m); return NULL; }
        l->gate_up = (uint16_t *)malloc((size_t)2 * m->I * m->H * 2);
        for (int r = 0; r < m->I; r++) {
            memcpy(l->gate_up + (size_t)(2 * r) * m->H, g + (size_t)r * m->H, (size_t)m->H * 2);
            memcpy(l->gate_up + (size_t)(2 * r + 1) * m->H, u + (size_t)r * m->H, (size_t)m->H * 2);
        }
    }
    m->final_norm = get_f32(m, "thinker.model.norm.weight");
    if (!m->c1w || !m->conv_out || !m->emb || !m->final_norm || !m->p2w) { qo_free(m); return NULL; }
    return m;
}

void qo_free(qo_model_t *m) {
    if (!m) return;
    free(m->c1w); free(m->c1b); free(m->c2w); free(m->c2b); free(m->c3w); free(m->c3b); free(m->conv_out);
    free(m->lnpw); free(m->lnpb); free(m->p1w); free(m->p1b); free(m->p2w); free(m->p2b);
    for (int i = 0; i < MAX_ENC; i++) {
        enc_layer_t *l = &m->enc[i];
        free(l->wq); free(l->bq); free(l->wk); free(l->bk); free(l->wv); free(l->bv); free(l->wo); free(l->bo);
        free(l->ln1w); free(l->ln1b); free(l->fc1w); free(l->fc1b); free(l->fc2w); free(l->fc2b); free(l->ln2w); free(l->ln2b);
    }
    for (int i = 0; i < MAX_DEC; i++) {
        dec_layer_t *l = &m->dec[i];
        free(l->gate_up); free(l->qn); free(l->kn); free(l->in_norm); free(l->post_norm);
    }
    free(m->final_norm); free(m->kv_k); free(m->kv_v);
    if (m->st) qst_close(m->st);
    free(m);
}

void qo_config(const qo_model_t *m, int *o) {
    o[0] = m->d; o[1] = m->enc_layers; o[2] = m->enc_heads; o[3] = m->F; o[4] = m->H; o[5] = m->H;
    o[6] = m->dec_layers; o[7] = m->heads; o[8] = m->kv_heads; o[9] = m->hd; o[10] = m->I; o[11] = m->V;
}
void qo_free_buf(void *p) { free(p); }

/* ======================================================================
 * Level 1: encoder.  reference qwen_asr_encoder.c:171-372
 * ====================================================================== */
static int conv_out_len(int w) { return (w + 2 - 3) / 2 + 1; }

float *qo_encoder_forward(qo_model_t *m, const float *mel, int frames, int *out_T) {
    int d = m->d, nh = m->enc_heads, hd = 64, F = m->F, H = m->H, CH = 480, chunk = 100;
    int n_chunks = (frames + chunk - 1) / chunk, T = 0;
    for (int c = 0; c < n_chunks; c++) {
        int w = frames - c * chunk; if (w > chunk) w = chunk;
        T += conv_out_len(conv_out_len(conv_out_len(w)));
    }
    float *x = (float *)calloc((size_t)T * d, sizeof(float));
    float *pe = (float *)malloc((size_t)13 * d * sizeof(float));
    qo_sinusoidal_pe(pe, 13, d); /* positions restart at 0 in every chunk, :280-284 */
    int tok = 0;
    for (int c = 0; c < n_chunks; c++) { /* per-chunk conv stem, :221-287 */
        int s0 = c * chunk, w = frames - s0; if (w > chunk) w = chunk;
        int w1 = conv_out_len(w), w2 = conv_out_len(w1), w3 = conv_out_len(w2);
        float *cm = (float *)malloc((size_t)128 * w * sizeof(float));
        for (int r = 0; r < 128; r++) memcpy(cm + (size_t)r * w, mel + (size_t)r * frames + s0, (size_t)w * sizeof(float));
        float *a1 = (float *)malloc((size_t)CH * 64 * w1 * sizeof(float));
        qo_conv2d(a1, cm, m->c1w, m->c1b, 1, CH, 128, w, 3, 3, 2, 1); qo_gelu(a1, CH * 64 * w1); free(cm);
        float *a2 = (float *)malloc((size_t)CH * 32 * w2 * sizeof(float));
        qo_conv2d(a2, a1, m->c2w, m->c2b, CH, CH, 64, w1, 3, 3, 2, 1); qo_gelu(a2, CH * 32 * w2); free(a1);
        float *a3 = (float *)malloc((size_t)CH * 16 * w3 * sizeof(float));
        qo_conv2d(a3, a2, m->c3w, m->c3b, CH, CH, 32, w2, 3, 3, 2, 1); qo_gelu(a3, CH * 16 * w3); free(a2);
        float *flat = (float *)malloc((size_t)w3 * 7680 * sizeof(float)); /* [t][ch*16+f], :262-271 */
        for (int t = 0; t < w3; t++)
            for (int ch = 0; ch < CH; ch++)
                for (int f = 0; f < 16; f++) flat[(size_t)t * 7680 + ch * 16 + f] = a3[((size_t)ch * 16 + f) * w3 + t];
        free(a3);
        qo_linear(x + (size_t)tok * d, flat, m->conv_out, NULL, w3, 7680, d);
        free(flat);
        for (int i = 0; i < w3 * d; i++) x[(size_t)tok * d + i] += pe[i];
        tok += w3;
    }
    free(pe);
    int win = 13 * (800 / chunk); /* 104-token windows, :291-297 */
    int n_win = (T + win - 1) / win;
    int *ws = (int *)malloc((size_t)(n_win + 1) * sizeof(int));
    for (int i = 0; i < n_win; i++) ws[i] = i * win;
    ws[n_win] = T;

    size_t Td = (size_t)T * d;
    float *xn = (float *)malloc(Td * 4), *q = (float *)malloc(Td * 4), *k = (float *)malloc(Td * 4), *v = (float *)malloc(Td * 4);
    float *ao = (float *)malloc(Td * 4), *po = (float *)malloc(Td * 4), *mid = (float *)malloc((size_t)T * F * 4);
    float scale = 1.0f / sqrtf((float)hd);
    for (int L = 0; L < m->enc_layers; L++) { /* pre-LN block, :312-347 */
        enc_layer_t *l = &m->enc[L];
        qo_layer_norm(xn, x, l->ln1w, l->ln1b, T, d, 1e-5f);
        qo_linear(q, xn, l->wq, l->bq, T, d, d);
        qo_linear(k, xn, l->wk, l->bk, T, d, d);
        qo_linear(v, xn, l->wv, l->bv, T, d, d);
        qo_bidirectional_attention(ao, q, k, v, T, nh, hd, scale, ws, n_win);
        qo_linear(po, ao, l->wo, l->bo, T, d, d);
        for (size_t i = 0; i < Td; i++) x[i] += po[i];
        qo_layer_norm(xn, x, l->ln2w, l->ln2b, T, d, 1e-5f);
        qo_linear(mid, xn, l->fc1w, l->fc1b, T, d, F);
        qo_gelu(mid, T * F);
        qo_linear(po, mid, l->fc2w, l->fc2b, T, F, d);
        for (size_t i = 0; i < Td; i++) x[i] += po[i];
    }
    qo_layer_norm(x, x, m->lnpw, m->lnpb, T, d, 1e-5f); /* tail, :350-361 */
    qo_linear(xn, x, m->p1w, m->p1b, T, d, d);
    qo_gelu(xn, T * d);
    float *out = (float *)malloc((size_t)T * H * sizeof(float));
    qo_linear(out, xn, m->p2w, m->p2b, T, d, H);
    free(x); free(xn); free(q); free(k); free(v); free(ao); free(po); free(mid); free(ws);
    *out_T = T;
    return out;
}

/* ======================================================================
 * Level 1: decoder.  reference qwen_asr_decoder.c:168-216 (KV cache), 457-563 (prefill),
 * 592-685 (step), 691-783 (step + logits)
 * ====================================================================== */
void qo_set_kv_len(qo_model_t *m, int n) { m->kv_len = n; }
int qo_get_kv_len(const qo_model_t *m) { return m->kv_len; }

static int kv_reserve(qo_model_t *m, int need) {
    if (need <= m->kv_max) return 0;
    int kvd = m->kv_heads * m->hd, nm = m->kv_max ? m->kv_max : 1024;
    while (nm < need) nm *= 2;
    float *nk = (float *)calloc((size_t)m->dec_layers * nm * kvd, 4), *nv = (float *)calloc((size_t)m->dec_layers * nm * kvd, 4);
    if (!nk || !nv) { free(nk); free(nv); return -1; }
    for (int l = 0; l < m->dec_layers && m->kv_k; l++) {
        memcpy(nk + (size_t)l * nm * kvd, m->kv_k + (size_t)l * m->kv_max * kvd, (size_t)m->kv_len * kvd * 4);
        memcpy(nv + (size_t)l * nm * kvd, m->kv_v + (size_t)l * m->kv_max * kvd, (size_t)m->kv_len * kvd * 4);
    }
    free(m->kv_k); free(m->kv_v);
    m->kv_k = nk; m->kv_v = nv; m->kv_max = nm;
    return 0;
}

/* Shared block stack over `seq` rows starting at position kv_len; leaves the post-stack
 * residual in x (caller frees).  seq==1 is the decode step. */
static float *decoder_rows(qo_model_t *m, const float *embeds, int seq) {
    int H = m->H, qd = m->heads * m->hd, kvd = m->kv_heads * m->hd, I = m->I, hd = m->hd;
    int start = m->kv_len;
    if (kv_reserve(m, start + seq + (m->kv_max ? 0 : 1024)) != 0) return NULL;
    float *x = (float *)malloc((size_t)seq * H * 4), *xn = (float *)malloc((size_t)seq * H * 4);
    float *q = (float *)malloc((size_t)seq * qd * 4), *k = (float *)malloc((size_t)seq * kvd * 4), *v = (float *)malloc((size_t)seq * kvd * 4);
    float *ao = (float *)malloc((size_t)seq * qd * 4), *po = (float *)malloc((size_t)seq * H * 4);
    float *gu = (float *)malloc((size_t)seq * 2 * I * 4), *act = (float *)malloc((size_t)seq * I * 4);
    float *rc = (float *)malloc((size_t)seq * hd * 4), *rs = (float *)malloc((size_t)seq * hd * 4);
    int *pos = (int *)malloc((size_t)seq * sizeof(int));
    for (int i = 0; i < seq; i++) pos[i] = start + i;
    qo_compute_rope_neox(rc, rs, pos, seq, hd, 1e6f);
    memcpy(x, embeds, (size_t)seq * H * 4);
    float scale = 1.0f / sqrtf((float)hd);
    for (int L = 0; L < m->dec_layers; L++) {
        dec_layer_t *l = &m->dec[L];
        qo_rms_norm(xn, x, l->in_norm, seq, H, 1e-6f);
        qo_linear_bf16(q, xn, l->wq, NULL, seq, H, qd);
        qo_linear_bf16(k, xn, l->wk, NULL, seq, H, kvd);
        qo_linear_bf16(v, xn, l->wv, NULL, seq, H, kvd);
        qo_rms_norm_per_head(q, l->qn, seq, m->heads, hd, 1e-6f);
        qo_rms_norm_per_head(k, l->kn, seq, m->kv_heads, hd, 1e-6f);
        qo_apply_rope_neox(q, rc, rs, seq, m->heads, hd);
        qo_apply_rope_neox(k, rc, rs, seq, m->kv_heads, hd);
        float *Kc = m->kv_k + (size_t)L * m->kv_max * kvd, *Vc = m->kv_v + (size_t)L * m->kv_max * kvd;
        memcpy(Kc + (size_t)start * kvd, k, (size_t)seq * kvd * 4);
        memcpy(Vc + (size_t)start * kvd, v, (size_t)seq * kvd * 4);
        qo_causal_attention(ao, q, Kc, Vc, seq, start + seq, m->heads, m->kv_heads, hd, scale, start);
        qo_linear_bf16(po, ao, l->wo, NULL, seq, qd, H);
        for (size_t i = 0; i < (size_t)seq * H; i++) x[i] += po[i];
        qo_rms_norm(xn, x, l->post_norm, seq, H, 1e-6f);
        qo_linear_bf16(gu, xn, l->gate_up, NULL, seq, H, 2 * I);
        qo_swiglu_multiply(act, gu, seq, I);
        qo_linear_bf16(po, act, l->down, NULL, seq, I, H);
        for (size_t i = 0; i < (size_t)seq * H; i++) x[i] += po[i];
    }
    m->kv_len = start + seq;
    free(xn); free(q); free(k); free(v); free(ao); free(po); free(gu); free(act); free(rc); free(rs); free(pos);
    return x;
}

void qo_decoder_prefill(qo_model_t *m, const float *embeds, int seq) {
    float *x = decoder_rows(m, embeds, seq);
    free(x);
}

int qo_decoder_forward(qo_model_t *m, const float *embed) {
    float *x = decoder_rows(m, embed, 1);
    if (!x) return 151645; /* error-as-EOS, qwen_asr_decoder.c:621,625 */
    qo_rms_norm(x, x, m->final_norm, 1, m->H, 1e-6f);
    int t = qo_argmax_matvec_bf16(x, m->emb, m->H, m->V);
    free(x);
    return t;
}

void qo_decoder_forward_logits(qo_model_t *m, const float *embed, float *logits) {
    float *x = decoder_rows(m, embed, 1);
    if (!x) { memset(logits, 0, (size_t)m->V * 4); return; }
    qo_rms_norm(x, x, m->final_norm, 1, m->H, 1e-6f);
    qo_linear_bf16(logits, x, m->emb, NULL, 1, m->H, m->V);
    free(x);
}

void qo_embed_token(const qo_model_t *m, int tok, float *dst) { /* qwen_asr.c:412-419 */
    for (int i = 0; i < m->H; i++) dst[i] = bf16_f32(m->emb[(size_t)tok * m->H + i]);
}

void qo_read_kv(const qo_model_t *m, int layer, int len, float *k_out, float *v_out) {
    int kvd = m->kv_heads * m->hd;
    memcpy(k_out, m->kv_k + (size_t)layer * m->kv_max * kvd, (size_t)len * kvd * 4);
    memcpy(v_out, m->kv_v + (size_t)layer * m->kv_max * kvd, (size_t)len * kvd * 4);
}

static double now_ms(void) {
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec * 1000.0 + tv.tv_usec / 1000.0;
}

/* Driver of one offline segment: prompt layout qwen_asr.c:388-399,685-759; prefill of
 * total_seq-1 rows then the single-token step (:764-769); greedy loop with EOS stop
 * (:788-818) under an explicit cap. */
int qo_transcribe_ids(qo_model_t *m, const float *samples, int n_samples, int max_new, int *out_ids,
                      double *tm, int *out_T) {
    static const int PRE[] = {151644, 8948, 198, 151645, 198, 151644, 872, 198, 151669};
    static const int SUF[] = {151670, 151645, 198, 151644, 77091, 198};
    int H = m->H, frames = 0, T = 0;
    double t0 = now_ms();
    float *mel = qo_mel_spectrogram(samples, n_samples, &frames);
    if (!mel) return -1;
    double t1 = now_ms();
    float *enc = qo_encoder_forward(m, mel, frames, &T);
    free(mel);
    double t2 = now_ms();
    int total = 9 + T + 6;
    float *emb = (float *)malloc((size_t)total * H * 4), *tmp = (float *)malloc((size_t)H * 4);
    for (int i = 0; i < 9; i++) qo_embed_token(m, PRE[i], emb + (size_t)i * H);
    memcpy(emb + (size_t)9 * H, enc, (size_t)T * H * 4);
    for (int i = 0; i < 6; i++) qo_embed_token(m, SUF[i], emb + (size_t)(9 + T + i) * H);
    free(enc);
    m->kv_len = 0;
    qo_decoder_prefill(m, emb, total - 1);
    int tok = qo_decoder_forward(m, emb + (size_t)(total - 1) * H);
    free(emb);
    double t3 = now_ms();
    int n = 0;
    while (n < max_new) {
        out_ids[n++] = tok;
        if (tok == 151643 || tok == 151645) break;
        if (n >= max_new) break;
        qo_embed_token(m, tok, tmp);
        tok = qo_decoder_forward(m, tmp);
    }
    double t4 = now_ms();
    free(tmp);
    if (tm) { tm[0] = t1 - t0; tm[1] = t2 - t1; tm[2] = t3 - t2; tm[3] = t4 - t3; }
    if (out_T) *out_T = T;
    return n;
}


/* ---------------------------------------------------------------- WAV bytes -> f32 mono 16 kHz
 * Restates qwen_parse_wav_buffer (reference qwen_asr_audio.c:40-168): RIFF chunk walk (:52-69), 16-bit PCM only
 * (:71-75), channels averaged and scaled by 1/32768 (:81-94), then - if the file is not at 16 kHz - a windowed-sinc
 * resampler (:96-164): 32 taps around floor(i / ratio), sinc cut at min(ratio, 1), Kaiser window beta = 6 with I0 as a
 * 20-term power series, every output divided by the sum of the coefficients used.  All of it in double, like the
 * reference. */
static double qo_bessel_i0(double x) { /* :109-116 */
    double sum = 1.0, term = 1.0, xx = x * x;
    for (int k = 1; k <= 20; k++) {
        term *= xx / (4.0 * (double)k * (double)k);
        sum += term;
    }
    return sum;
}

float *qo_resample_to_16k(const float *in, int n, int rate, int *out_n) {
    const double pi = 3.14159265358979323846;
    int new_n = (int)((long long)n * 16000 / rate); /* :99 */
    float *out = (float *)malloc((size_t)(new_n > 0 ? new_n : 1) * sizeof(float));
    double ratio = 16000.0 / (double)rate, cutoff = ratio < 1.0 ? ratio : 1.0, inv_i0 = 1.0 / qo_bessel_i0(6.0);
    for (int i = 0; i < new_n; i++) {
        double pos = (double)i / ratio, acc = 0.0, wsum = 0.0;
        int c = (int)pos;
        for (int j = c - 15; j <= c + 16; j++) { /* SINC_HALF = 16: j in [c-16+1, c+16], :126-127 */
            double d = (double)j - pos, x = d * cutoff;
            double sv = fabs(x) < 1e-9 ? 1.0 : sin(pi * x) / (pi * x);
            double np_ = d / 16.0;
            double w = (np_ <= -1.0 || np_ >= 1.0) ? 0.0 : qo_bessel_i0(6.0 * sqrt(1.0 - np_ * np_)) * inv_i0;
            double coeff = sv * w * cutoff;
            if (j >= 0 && j < n) acc += (double)in[j] * coeff;
            wsum += coeff;
        }
        out[i] = wsum > 1e-9 ? (float)(acc / wsum) : 0.0f;
    }
    *out_n = new_n;
    return out;
}

static unsigned qo_rd16(const uint8_t *p) { return (unsigned)p[0] | ((unsigned)p[1] << 8); }
static unsigned qo_rd32(const uint8_t *p) { return (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24); }

float *qo_parse_wav_buffer(const uint8_t *data, size_t size, int *out_n) {
    if (size < 44 || memcmp(data, "RIFF", 4) || memcmp(data + 8, "WAVE", 4)) return NULL;
    int fmt = 0, ch = 0, rate = 0, bits = 0, pcm_bytes = 0;
    const uint8_t *pcm = NULL, *p = data + 12, *end = data + size;
    while (p + 8 <= end) {
        unsigned len = qo_rd32(p + 4);
        if (p + 8 + len > end) break;
        if (!memcmp(p, "fmt ", 4) && len >= 16) { fmt = qo_rd16(p + 8); ch = qo_rd16(p + 10); rate = (int)qo_rd32(p + 12); bits = qo_rd16(p + 22); }
        else if (!memcmp(p, "data", 4)) { pcm = p + 8; pcm_bytes = (int)len; }
        p += 8 + len + (len & 1);
    }
    if (fmt != 1 || bits != 16 || !pcm || ch < 1) return NULL;
    int n = pcm_bytes / (ch * 2);
    float *mono = (float *)malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
    for (int i = 0; i < n; i++) {
        if (ch == 1) {
            int16_t v; memcpy(&v, pcm + (size_t)i * 2, 2);
            mono[i] = v / 32768.0f;
        } else { /* float sum of the channels, then / channels / 32768, :86-92 */
            float sum = 0;
            for (int c = 0; c < ch; c++) { int16_t v; memcpy(&v, pcm + ((size_t)i * ch + c) * 2, 2); sum += v; }
            mono[i] = (sum / ch) / 32768.0f;
        }
    }
    if (rate != 16000) {
        int nn = 0;
        float *r = qo_resample_to_16k(mono, n, rate, &nn);
        free(mono);
        mono = r; n = nn;
    }
    *out_n = n;
    return mono;
}

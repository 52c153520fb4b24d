// qasr_api.cu - C ABI of libqasr_cuda.so (see include/qasr_cuda.h): context, one-time HBM
// upload of the safetensors checkpoint, the level-1 entry points (mel / encoder / prefill /
// decode step / greedy loop) and their CUDA-graph plumbing.  No CPU fallback anywhere: every
// path either launches sm_100a kernels or fails with an error code.
#include "qasr_ctx.h"
#include "qasr_safetensors.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <string>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
int set_err(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
const char *qasr_cuda_last_error(void) { return g_err; }

// ------------------------------------------------------------------ lifecycle
int qasr_cuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

qasr_ctx_t *qasr_cuda_init(int device) {
    int n = qasr_cuda_device_count();
    if (n <= 0) { set_err(QASR_ERR_CUDA, "no CUDA device visible: libqasr_cuda has no CPU fallback"); return nullptr; }
    if (device < 0 || device >= n) { set_err(QASR_ERR_ARG, "device %d out of range (%d visible)", device, n); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { set_err(QASR_ERR_CUDA, "cudaSetDevice(%d) failed", device); return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        set_err(QASR_ERR_CUDA, "device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
        return nullptr;
    }
    qasr_ctx_t *c = new qasr_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_err(QASR_ERR_CUDA, "cudaStreamCreate failed");
        delete c;
        return nullptr;
    }
    for (int i = 0; i < 5; i++) cudaEventCreate(&c->ev[i]);
    const char *ng = getenv("QASR_NO_GRAPH");
    c->use_graph = !(ng && ng[0] == '1');
    const char *dm = getenv("QASR_DECODE");
    c->use_stream = !(dm && strcmp(dm, "graph") == 0);
    if (c->use_stream && stream_init() != 0) {
        set_err(QASR_ERR_CUDA, "%s", stream_error());
        cudaStreamDestroy(c->stream);
        delete c;
        return nullptr;
    }
    if (gemm_tc_prepare() != 0) {
        set_err(QASR_ERR_CUDA, "%s", gemm_tc_error());
        cudaStreamDestroy(c->stream);
        delete c;
        return nullptr;
    }
    return c;
}

void qasr_cuda_free(qasr_ctx_t *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->graph) cudaGraphDestroy(c->graph);
    for (auto &ge : c->graph_cache) cudaGraphExecDestroy(ge.exec);
    for (void *p : c->owned) cudaFree(p);
    for (int q = 0; q < QASR_STREAM_MAX_SEQS; q++) { cudaFree(c->kv_ks[q]); cudaFree(c->kv_vs[q]); }
    cudaFree(c->rope_cos); cudaFree(c->rope_sin);
    if (c->h_tokens) cudaFreeHost(c->h_tokens);
    c->ws_samples.release(); c->ws_meltmp.release(); c->ws_mel.release(); c->ws_enc.release();
    c->ws_encout.release(); c->ws_pre.release(); c->ws_ids.release(); c->ws_geom.release();
    for (auto &w : c->st_win) w.rows.release();
    c->ws_pcm.release(); c->ws_mono.release();
    batch_release(c);
    for (int i = 0; i < 5; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 2; i++) if (c->tev[i]) cudaEventDestroy(c->tev[i]);
    cudaStreamDestroy(c->stream);
    delete c;
}

int qasr_cuda_set_gemm_split(qasr_ctx_t *c, int nsplit) {
    if (!c || (nsplit != 1 && nsplit != 2)) return set_err(QASR_ERR_ARG, "nsplit must be 1 or 2");
    c->nsplit = nsplit;
    return 0;
}

int qasr_cuda_config(const qasr_ctx_t *c, int *o) {
    if (!c || !o) return set_err(QASR_ERR_ARG, "null argument");
    if (!c->loaded) return set_err(QASR_ERR_STATE, "no model loaded");
    o[0] = c->d; o[1] = c->enc_layers; o[2] = c->enc_heads; o[3] = c->F; o[4] = c->H; o[5] = c->H;
    o[6] = c->dec_layers; o[7] = c->heads; o[8] = c->kv_heads; o[9] = c->hd; o[10] = c->I; o[11] = c->V;
    return 0;
}

double qasr_cuda_last_decode_ms(const qasr_ctx_t *c) { return c ? c->last_decode_ms : 0.0; }
long long qasr_cuda_launch_count(const qasr_ctx_t *c) { return c ? c->launches : 0; }
void qasr_set_threads(int n) { (void)n; }
int qasr_get_num_cpus(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }

// ------------------------------------------------------------------ weight upload
static inline float bf16_to_f32(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }

static void *dev_alloc(qasr_ctx_t *c, size_t bytes) {
    void *p = nullptr;
    if (cudaMalloc(&p, align_up(bytes, 256)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    c->owned.push_back(p);
    c->weight_bytes += bytes;
    return p;
}

static const qst_tensor_t *need(qst_dir_t *st, const char *name, int *rc) {
    const qst_tensor_t *t = qst_find(st, name);
    if (!t) *rc = set_err(QASR_ERR_MODEL, "tensor not found: %s", name);
    return t;
}

// f32-class tensor (norm weights, biases, conv1): BF16 upcast exactly, or F32 verbatim
// (reference safetensors_get_f32, qwen_asr_safetensors.c:255-278)
static float *up_f32(qasr_ctx_t *c, qst_dir_t *st, const char *name, size_t expect_numel, int *rc) {
    const qst_tensor_t *t = need(st, name, rc);
    if (!t) return nullptr;
    if (t->numel != expect_numel) { *rc = set_err(QASR_ERR_MODEL, "%s has %zu elements, expected %zu", name, t->numel, expect_numel); return nullptr; }
    std::vector<float> h(t->numel);
    if (t->dtype == QST_BF16) {
        const uint16_t *s = (const uint16_t *)t->data;
        for (size_t i = 0; i < t->numel; i++) h[i] = bf16_to_f32(s[i]);
    } else if (t->dtype == QST_F32) {
        memcpy(h.data(), t->data, t->numel * 4);
    } else { *rc = set_err(QASR_ERR_MODEL, "unsupported dtype for %s", name); return nullptr; }
    float *d = (float *)dev_alloc(c, t->numel * 4);
    if (!d || cudaMemcpy(d, h.data(), t->numel * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        *rc = set_err(QASR_ERR_CUDA, "upload failed: %s", name);
        return nullptr;
    }
    return d;
}

static const uint16_t *host_bf16(qst_dir_t *st, const char *name, size_t expect_numel, int *rc) {
    const qst_tensor_t *t = need(st, name, rc);
    if (!t) return nullptr;
    if (t->dtype != QST_BF16) { *rc = set_err(QASR_ERR_MODEL, "%s must be BF16 (matrix weights are uploaded verbatim)", name); return nullptr; }
    if (expect_numel && t->numel != expect_numel) { *rc = set_err(QASR_ERR_MODEL, "%s has %zu elements, expected %zu", name, t->numel, expect_numel); return nullptr; }
    return (const uint16_t *)t->data;
}

// Checkpoint -> HBM: two pinned staging buffers; the host copies mmap pages into one while the DMA engine
// drains the other (a pageable cudaMemcpy per tensor serialises page faults, staging and DMA: 1.9 s for the 4.1 GB 1.7B
// checkpoint).  Row interleaving (gate/up) happens while staging, so no second host copy of those matrices exists.
struct Uploader {
    static constexpr size_t CHUNK = (size_t)32 << 20;
    uint8_t *pin[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    cudaStream_t s = nullptr;
    int cur = 0;
    bool ok = true;
    bool init(cudaStream_t stream) {
        s = stream;
        for (int i = 0; i < 2; i++)
            if (cudaHostAlloc((void **)&pin[i], CHUNK, cudaHostAllocDefault) != cudaSuccess || cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                ok = false;
            }
        return ok;
    }
    uint8_t *acquire() { cudaEventSynchronize(done[cur]); return pin[cur]; }
    void submit(void *dst, size_t bytes) {
        if (cudaMemcpyAsync(dst, pin[cur], bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) ok = false;
        cudaEventRecord(done[cur], s);
        cur ^= 1;
    }
    void put(void *dst, const void *src, size_t bytes) {
        for (size_t off = 0; off < bytes && ok; off += CHUNK) {
            const size_t n = bytes - off < CHUNK ? bytes - off : CHUNK;
            memcpy(acquire(), (const uint8_t *)src + off, n);
            submit((uint8_t *)dst + off, n);
        }
    }
    // dst rows (2r, 2r+1) = (a row r, b row r): the gate/up interleave of reference qwen_asr_decoder.c:140-152
    void put_interleaved(void *dst, const void *a, const void *b, size_t rows, size_t row_bytes) {
        const size_t per = CHUNK / (2 * row_bytes);
        for (size_t r0 = 0; r0 < rows && ok; r0 += per) {
            const size_t n = rows - r0 < per ? rows - r0 : per;
            uint8_t *p = acquire();
            for (size_t r = 0; r < n; r++) {
                memcpy(p + (2 * r) * row_bytes, (const uint8_t *)a + (r0 + r) * row_bytes, row_bytes);
                memcpy(p + (2 * r + 1) * row_bytes, (const uint8_t *)b + (r0 + r) * row_bytes, row_bytes);
            }
            submit((uint8_t *)dst + 2 * r0 * row_bytes, 2 * n * row_bytes);
        }
    }
    ~Uploader() { finish(); }
    bool finish() {
        if (s && cudaStreamSynchronize(s) != cudaSuccess) ok = false;
        for (int i = 0; i < 2; i++) { if (pin[i]) cudaFreeHost(pin[i]); if (done[i]) cudaEventDestroy(done[i]); pin[i] = nullptr; done[i] = nullptr; }
        return ok;
    }
};
static thread_local Uploader *g_up = nullptr; // active during qasr_cuda_load_dir

// bf16 matrix uploaded verbatim from the mmap (north_star item 5)
static bf16_t *up_bf16(qasr_ctx_t *c, qst_dir_t *st, const char *name, size_t numel, int *rc) {
    const uint16_t *h = host_bf16(st, name, numel, rc);
    if (!h) return nullptr;
    bf16_t *d = (bf16_t *)dev_alloc(c, numel * 2);
    if (!d) { *rc = set_err(QASR_ERR_NOMEM, "cudaMalloc failed: %s", name); return nullptr; }
    if (g_up) g_up->put(d, h, numel * 2);
    else if (cudaMemcpy(d, h, numel * 2, cudaMemcpyHostToDevice) != cudaSuccess) { *rc = set_err(QASR_ERR_CUDA, "upload failed: %s", name); return nullptr; }
    return d;
}

// several bf16 matrices stacked along rows into one device matrix (fused QKV)
static bf16_t *up_bf16_cat(qasr_ctx_t *c, qst_dir_t *st, const char *const *names, const size_t *numels, int n, int *rc) {
    size_t total = 0;
    for (int i = 0; i < n; i++) total += numels[i];
    bf16_t *d = (bf16_t *)dev_alloc(c, total * 2);
    if (!d) { *rc = set_err(QASR_ERR_NOMEM, "cudaMalloc failed"); return nullptr; }
    size_t off = 0;
    for (int i = 0; i < n; i++) {
        const uint16_t *h = host_bf16(st, names[i], numels[i], rc);
        if (!h) return nullptr;
        if (g_up) g_up->put(d + off, h, numels[i] * 2);
        else if (cudaMemcpy(d + off, h, numels[i] * 2, cudaMemcpyHostToDevice) != cudaSuccess) { *rc = set_err(QASR_ERR_CUDA, "upload failed: %s", names[i]); return nullptr; }
        off += numels[i];
    }
    return d;
}

static float *up_f32_cat(qasr_ctx_t *c, qst_dir_t *st, const char *const *names, int n, int each, int *rc) {
    std::vector<float> h((size_t)n * each);
    for (int i = 0; i < n; i++) {
        const qst_tensor_t *t = need(st, names[i], rc);
        if (!t) return nullptr;
        if ((int)t->numel != each) { *rc = set_err(QASR_ERR_MODEL, "%s: bad size", names[i]); return nullptr; }
        if (t->dtype == QST_BF16) for (int k = 0; k < each; k++) h[(size_t)i * each + k] = bf16_to_f32(((const uint16_t *)t->data)[k]);
        else if (t->dtype == QST_F32) memcpy(&h[(size_t)i * each], t->data, (size_t)each * 4);
        else { *rc = set_err(QASR_ERR_MODEL, "unsupported dtype for %s", names[i]); return nullptr; }
    }
    float *d = (float *)dev_alloc(c, h.size() * 4);
    if (!d || cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) { *rc = set_err(QASR_ERR_CUDA, "upload failed"); return nullptr; }
    return d;
}

static bf16_t *up_host_vec(qasr_ctx_t *c, const std::vector<uint16_t> &h, int *rc) {
    bf16_t *d = (bf16_t *)dev_alloc(c, h.size() * 2);
    if (!d || cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) { *rc = set_err(QASR_ERR_CUDA, "upload failed"); return nullptr; }
    return d;
}
static float *up_host_f32(qasr_ctx_t *c, const std::vector<float> &h, int *rc) {
    float *d = (float *)dev_alloc(c, h.size() * 4);
    if (!d || cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) { *rc = set_err(QASR_ERR_CUDA, "upload failed"); return nullptr; }
    return d;
}

// Slaney mel filterbank, reference qwen_asr_audio.c:236-287.  Stored [201][128] (bin-major).
static float hz_to_mel(float f) { return f >= 1000.0f ? 15.0f + logf(f / 1000.0f) * (27.0f / logf(6.4f)) : 3.0f * f / 200.0f; }
static float mel_to_hz(float m) { return m >= 15.0f ? 1000.0f * expf((logf(6.4f) / 27.0f) * (m - 15.0f)) : 200.0f * m / 3.0f; }

static int build_tables(qasr_ctx_t *c) {
    int rc = 0;
    std::vector<float> ct(400 * 208, 0.0f), st(400 * 208, 0.0f), win(400), fb(201 * 128, 0.0f);
    for (int k = 0; k < 201; k++)
        for (int j = 0; j < 400; j++) { // f32 angle exactly as the reference forms it (:330-335)
            float ang = 2.0f * (float)M_PI * (float)k * (float)j / (float)400;
            ct[j * 208 + k] = cosf(ang);
            st[j * 208 + k] = sinf(ang);
        }
    for (int i = 0; i < 400; i++) win[i] = 0.5f * (1.0f - cosf(2.0f * (float)M_PI * (float)i / (float)400));
    float pts[130];
    const float mmin = hz_to_mel(0.0f), mmax = hz_to_mel(8000.0f);
    for (int i = 0; i < 130; i++) pts[i] = mel_to_hz(mmin + (mmax - mmin) * (float)i / 129.0f);
    for (int m = 0; m < 128; m++) {
        float dl = pts[m + 1] - pts[m], dr = pts[m + 2] - pts[m + 1];
        if (dl == 0.0f) dl = 1e-6f;
        if (dr == 0.0f) dr = 1e-6f;
        const float en = 2.0f / (pts[m + 2] - pts[m]);
        for (int f = 0; f < 201; f++) {
            const float hz = (float)f * 8000.0f / 200.0f;
            float v = fminf((hz - pts[m]) / dl, (pts[m + 2] - hz) / dr);
            fb[f * 128 + m] = (v < 0.0f ? 0.0f : v) * en;
        }
    }
    c->mel_cos = up_host_f32(c, ct, &rc); c->mel_sin = up_host_f32(c, st, &rc);
    c->mel_win = up_host_f32(c, win, &rc); c->mel_fb = up_host_f32(c, fb, &rc);
    // per-chunk sinusoidal PE table [13][d], reference qwen_asr_kernels.c:1198-1211
    std::vector<float> pe((size_t)13 * c->d);
    const int half = c->d / 2;
    const float lt = logf(10000.0f) / (float)(half - 1);
    for (int p = 0; p < 13; p++)
        for (int i = 0; i < half; i++) {
            const float ang = (float)p * expf(-(float)i * lt);
            pe[(size_t)p * c->d + i] = sinf(ang);
            pe[(size_t)p * c->d + half + i] = cosf(ang);
        }
    c->pe = up_host_f32(c, pe, &rc);
    return rc;
}

// RoPE cos/sin [pos][64] with the reference's f32 formula (qwen_asr_decoder.c:253-302)
int ensure_rope(qasr_ctx_t *c, int need_pos) {
    if (need_pos <= c->rope_cap) return 0;
    int cap = c->rope_cap ? c->rope_cap : 4096;
    while (cap < need_pos) cap *= 2;
    std::vector<float> hc((size_t)cap * 64), hs((size_t)cap * 64);
    float inv[64];
    for (int d = 0; d < 64; d++) inv[d] = 1.0f / powf(1e6f, (float)(2 * d) / 128.0f);
    for (int p = 0; p < cap; p++)
        for (int d = 0; d < 64; d++) {
            const float ang = (float)p * inv[d];
            hc[(size_t)p * 64 + d] = cosf(ang);
            hs[(size_t)p * 64 + d] = sinf(ang);
        }
    CK(cudaStreamSynchronize(c->stream));
    float *nc = nullptr, *ns = nullptr;
    CK(cudaMalloc(&nc, hc.size() * 4));
    CK(cudaMalloc(&ns, hs.size() * 4));
    CK(cudaMemcpy(nc, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ns, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice));
    cudaFree(c->rope_cos); cudaFree(c->rope_sin);
    c->rope_cos = nc; c->rope_sin = ns; c->rope_cap = cap;
    c->ws_gen++;
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; } // pointers changed
    return 0;
}

// KV cache growth keeps rows [0,keep) of every layer (reference kv_cache_grow, qwen_asr_decoder.c:179-206).
// All sequence caches share one capacity; a growth re-strides every allocated one (rows [0, kv_fill[q]) of the
// others, rows [0, keep) of the current sequence).
static void select_seq(qasr_ctx_t *c, int q) { c->seq = q; c->kv_k = c->kv_ks[q]; c->kv_v = c->kv_vs[q]; }

static int ensure_kv(qasr_ctx_t *c, int need_pos, int keep) {
    const size_t kvd = (size_t)c->kv_heads * c->hd;
    if (need_pos <= c->kv_max && c->kv_ks[c->seq]) return 0;
    int cap = c->kv_max;
    if (!cap) { const char *e = getenv("QASR_KV_INIT_ROWS"); cap = e && atoi(e) > 0 ? atoi(e) : 2048; } // reference: seq + 1024 (qwen_asr_decoder.c:168-177)
    while (cap < need_pos) cap *= 2;
    const size_t bytes = (size_t)c->dec_layers * cap * kvd * 4;
    CK(cudaStreamSynchronize(c->stream));
    // allocate and fill every new cache first; the pointers and kv_max are committed together only when all succeeded
    float *nk[QASR_STREAM_MAX_SEQS] = {}, *nv[QASR_STREAM_MAX_SEQS] = {};
    bool grow[QASR_STREAM_MAX_SEQS] = {};
    int rc = 0;
    for (int q = 0; q < QASR_STREAM_MAX_SEQS && rc == 0; q++) {
        if (!c->kv_ks[q] && q != c->seq) continue;
        if (c->kv_ks[q] && cap == c->kv_max) continue;
        grow[q] = true;
        if (cudaMalloc(&nk[q], bytes) != cudaSuccess || cudaMalloc(&nv[q], bytes) != cudaSuccess) {
            cudaGetLastError();
            rc = set_err(QASR_ERR_NOMEM, "KV cache allocation of %zu bytes failed", 2 * bytes);
            break;
        }
        const int rows = q == c->seq ? keep : c->kv_fill[q];
        if (c->kv_ks[q] && rows > 0)
            for (int l = 0; l < c->dec_layers && rc == 0; l++)
                if (cudaMemcpy(nk[q] + (size_t)l * cap * kvd, c->kv_ks[q] + (size_t)l * c->kv_max * kvd, (size_t)rows * kvd * 4, cudaMemcpyDeviceToDevice) != cudaSuccess ||
                    cudaMemcpy(nv[q] + (size_t)l * cap * kvd, c->kv_vs[q] + (size_t)l * c->kv_max * kvd, (size_t)rows * kvd * 4, cudaMemcpyDeviceToDevice) != cudaSuccess)
                    rc = set_err(QASR_ERR_CUDA, "KV cache re-stride copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (rc != 0) { // nothing was committed: the context keeps running on the old caches
        for (int q = 0; q < QASR_STREAM_MAX_SEQS; q++) { cudaFree(nk[q]); cudaFree(nv[q]); }
        return rc;
    }
    for (int q = 0; q < QASR_STREAM_MAX_SEQS; q++)
        if (grow[q]) { cudaFree(c->kv_ks[q]); cudaFree(c->kv_vs[q]); c->kv_ks[q] = nk[q]; c->kv_vs[q] = nv[q]; }
    c->kv_max = cap;
    select_seq(c, c->seq);
    c->ws_gen++;
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    return 0;
}

// Upload every tensor of the checkpoint view `st` (closed here) and build the device-side state.
static int load_from(qasr_ctx_t *c, qst_dir_t *st) {
    int rc = 0;
    Uploader up;
    if (!up.init(c->stream)) { qst_close(st); up.finish(); return set_err(QASR_ERR_NOMEM, "pinned staging buffers for the checkpoint upload"); }
    g_up = &up;
    struct UpGuard { ~UpGuard() { g_up = nullptr; } } up_guard;
    // variant probe, reference qwen_asr.c:146-203 (dims are hard-coded there; config.json is not read)
    if (qst_find(st, "thinker.audio_tower.layers.31.self_attn.q_proj.weight")) {
        qst_close(st);
        return set_err(QASR_ERR_MODEL, "Qwen3-Omni-30B-MoE checkpoints are out of scope for this path");
    }
    const bool big = qst_find(st, "thinker.audio_tower.layers.18.self_attn.q_proj.weight") != nullptr;
    c->d = big ? 1024 : 896; c->enc_layers = big ? 24 : 18; c->enc_heads = big ? 16 : 14; c->F = big ? 4096 : 3584;
    c->H = big ? 2048 : 1024; c->dec_layers = 28; c->heads = 16; c->kv_heads = 8; c->hd = 128; c->I = big ? 6144 : 3072;
    c->V = 151936;
    const int d = c->d, F = c->F, H = c->H, I = c->I;
    char n0[256], n1[256], n2[256];
#define E "thinker.audio_tower."
    c->c1w = up_f32(c, st, E "conv2d1.weight", 480 * 9, &rc); c->c1b = up_f32(c, st, E "conv2d1.bias", 480, &rc);
    c->c2b = up_f32(c, st, E "conv2d2.bias", 480, &rc); c->c3b = up_f32(c, st, E "conv2d3.bias", 480, &rc);
    for (int cv = 2; cv <= 3 && rc == 0; cv++) { // [oc][ic][ki][kj] -> [oc][(ki*3+kj)*480 + ic]
        snprintf(n0, sizeof n0, E "conv2d%d.weight", cv);
        const uint16_t *h = host_bf16(st, n0, (size_t)480 * 480 * 9, &rc);
        if (!h) break;
        std::vector<uint16_t> perm((size_t)480 * 4320);
        for (int oc = 0; oc < 480; oc++)
            for (int ic = 0; ic < 480; ic++)
                for (int t = 0; t < 9; t++) perm[(size_t)oc * 4320 + t * 480 + ic] = h[((size_t)oc * 480 + ic) * 9 + t];
        (cv == 2 ? c->c2w : c->c3w) = up_host_vec(c, perm, &rc);
    }
    if (rc == 0) { // conv_out [d][ch*16+f] -> [d][f*480+ch]
        const uint16_t *h = host_bf16(st, E "conv_out.weight", (size_t)d * 7680, &rc);
        if (h) {
            std::vector<uint16_t> perm((size_t)d * 7680);
            for (int n = 0; n < d; n++)
                for (int ch = 0; ch < 480; ch++)
                    for (int f = 0; f < 16; f++) perm[(size_t)n * 7680 + f * 480 + ch] = h[(size_t)n * 7680 + ch * 16 + f];
            c->conv_out = up_host_vec(c, perm, &rc);
        }
    }
    for (int l = 0; l < c->enc_layers && rc == 0; l++) {
        EncLayerW &L = c->enc[l];
        snprintf(n0, sizeof n0, E "layers.%d.self_attn.q_proj.weight", l);
        snprintf(n1, sizeof n1, E "layers.%d.self_attn.k_proj.weight", l);
        snprintf(n2, sizeof n2, E "layers.%d.self_attn.v_proj.weight", l);
        const char *wn[3] = {n0, n1, n2};
        const size_t ne[3] = {(size_t)d * d, (size_t)d * d, (size_t)d * d};
        L.wqkv = up_bf16_cat(c, st, wn, ne, 3, &rc);
        snprintf(n0, sizeof n0, E "layers.%d.self_attn.q_proj.bias", l);
        snprintf(n1, sizeof n1, E "layers.%d.self_attn.k_proj.bias", l);
        snprintf(n2, sizeof n2, E "layers.%d.self_attn.v_proj.bias", l);
        L.bqkv = up_f32_cat(c, st, wn, 3, d, &rc);
#define LN(field, suffix, n) snprintf(n0, sizeof n0, E "layers.%d." suffix, l); L.field = up_f32(c, st, n0, (size_t)(n), &rc)
        snprintf(n0, sizeof n0, E "layers.%d.self_attn.out_proj.weight", l); L.wo = up_bf16(c, st, n0, (size_t)d * d, &rc);
        LN(bo, "self_attn.out_proj.bias", d);
        LN(ln1w, "self_attn_layer_norm.weight", d); LN(ln1b, "self_attn_layer_norm.bias", d);
        snprintf(n0, sizeof n0, E "layers.%d.fc1.weight", l); L.fc1 = up_bf16(c, st, n0, (size_t)F * d, &rc);
        LN(fc1b, "fc1.bias", F);
        snprintf(n0, sizeof n0, E "layers.%d.fc2.weight", l); L.fc2 = up_bf16(c, st, n0, (size_t)d * F, &rc);
        LN(fc2b, "fc2.bias", d);
        LN(ln2w, "final_layer_norm.weight", d); LN(ln2b, "final_layer_norm.bias", d);
#undef LN
    }
    if (rc == 0) {
        c->lnpw = up_f32(c, st, E "ln_post.weight", d, &rc); c->lnpb = up_f32(c, st, E "ln_post.bias", d, &rc);
        c->p1w = up_bf16(c, st, E "proj1.weight", (size_t)d * d, &rc); c->p1b = up_f32(c, st, E "proj1.bias", d, &rc);
        c->p2w = up_bf16(c, st, E "proj2.weight", (size_t)H * d, &rc); c->p2b = up_f32(c, st, E "proj2.bias", H, &rc);
    }
#undef E
    if (rc == 0) c->emb = up_bf16(c, st, "thinker.model.embed_tokens.weight", (size_t)c->V * H, &rc);
    for (int l = 0; l < c->dec_layers && rc == 0; l++) {
        DecLayerW &L = c->dec[l];
#define P "thinker.model.layers.%d."
        snprintf(n0, sizeof n0, P "self_attn.q_proj.weight", l);
        snprintf(n1, sizeof n1, P "self_attn.k_proj.weight", l);
        snprintf(n2, sizeof n2, P "self_attn.v_proj.weight", l);
        const char *wn[3] = {n0, n1, n2};
        const size_t ne[3] = {(size_t)2048 * H, (size_t)1024 * H, (size_t)1024 * H};
        L.wqkv = up_bf16_cat(c, st, wn, ne, 3, &rc);
        snprintf(n0, sizeof n0, P "self_attn.o_proj.weight", l); L.wo = up_bf16(c, st, n0, (size_t)H * 2048, &rc);
        snprintf(n0, sizeof n0, P "mlp.down_proj.weight", l); L.wdown = up_bf16(c, st, n0, (size_t)H * I, &rc);
        snprintf(n0, sizeof n0, P "self_attn.q_norm.weight", l); L.qn = up_f32(c, st, n0, 128, &rc);
        snprintf(n0, sizeof n0, P "self_attn.k_norm.weight", l); L.kn = up_f32(c, st, n0, 128, &rc);
        snprintf(n0, sizeof n0, P "input_layernorm.weight", l); L.in_norm = up_f32(c, st, n0, H, &rc);
        snprintf(n0, sizeof n0, P "post_attention_layernorm.weight", l); L.post_norm = up_f32(c, st, n0, H, &rc);
        // gate/up rows interleaved [g0,u0,g1,u1,...] (reference qwen_asr_decoder.c:140-152) so SwiGLU
        // pairs are adjacent GEMV rows / GEMM columns and fuse into the epilogue
        snprintf(n0, sizeof n0, P "mlp.gate_proj.weight", l);
        snprintf(n1, sizeof n1, P "mlp.up_proj.weight", l);
#undef P
        if (rc) break;
        const uint16_t *g = host_bf16(st, n0, (size_t)I * H, &rc), *u = host_bf16(st, n1, (size_t)I * H, &rc);
        if (!g || !u) break;
        L.wgu = (bf16_t *)dev_alloc(c, (size_t)2 * I * H * 2);
        if (!L.wgu) { rc = set_err(QASR_ERR_NOMEM, "cudaMalloc failed: gate/up"); break; }
        g_up->put_interleaved(L.wgu, g, u, (size_t)I, (size_t)H * 2);
    }
    if (rc == 0) c->final_norm = up_f32(c, st, "thinker.model.norm.weight", H, &rc);
    const bool up_ok = up.finish(); // every staged copy has landed before the mmap goes away
    g_up = nullptr;
    qst_close(st);
    if (rc) return rc;
    if (!up_ok) return set_err(QASR_ERR_CUDA, "checkpoint upload failed");
    CKR(build_tables(c));

    // decode-step state
    auto dalloc = [&](size_t bytes) -> void * { void *p = dev_alloc(c, bytes); if (p) cudaMemset(p, 0, bytes); return p; };
    c->n_parts = argmax_num_parts(c->V);
    c->x = (float *)dalloc((size_t)QASR_STREAM_MAX_SEQS * H * 4); c->pending = (float *)dalloc((size_t)H * 4);
    c->hidden_buf = (float *)dalloc((size_t)QASR_STREAM_MAX_SEQS * H * 4);
    c->qkv = (float *)dalloc(4096 * 4); c->attn = (float *)dalloc(2048 * 4); c->act = (float *)dalloc((size_t)I * 4);
    c->attn_part = (float *)dalloc((size_t)8 * QASR_ATTN_SPLITS * 2 * QASR_ATTN_PART_STRIDE * 4);
    c->logits = (float *)dalloc((size_t)c->V * 4);
    c->part_val = (float *)dalloc((size_t)c->n_parts * 4); c->part_idx = (int *)dalloc((size_t)c->n_parts * 4);
    c->counters = (unsigned *)dalloc(8 * 4);
    c->d_pos = (int *)dalloc(4 * QASR_STREAM_MAX_SEQS); c->d_done = (int *)dalloc(4); c->d_step = (int *)dalloc(4);
    c->d_tokens = (int *)dalloc((size_t)c->max_steps * QASR_STREAM_MAX_SEQS * 4);
    c->d_gmax = (int *)dalloc(4 * 512); // one slot per unit of a batched front end
    if (c->use_stream) { // decode weight image: every decoder matrix + the tied lm_head re-tiled into per-warp streams
        const int G = stream_grid();
        std::vector<unsigned long long> off((size_t)G + 1);
        const size_t image_bytes = stream_image_layout(c->dec_layers, H, I, c->V, off.data());
        if (!image_bytes) return set_err(QASR_ERR_CUDA, "%s", stream_error());
        c->sk_image = (uint8_t *)dev_alloc(c, image_bytes);
        c->sk_cta_off = (unsigned long long *)dalloc(off.size() * 8);
        if (!c->sk_image || !c->sk_cta_off) return set_err(QASR_ERR_NOMEM, "decode weight image (%zu bytes)", image_bytes);
        CK(cudaMemcpy(c->sk_cta_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice));
        std::vector<const bf16_t *> mats;
        for (int l = 0; l < c->dec_layers; l++) {
            const DecLayerW &L = c->dec[l];
            mats.push_back(L.wqkv); mats.push_back(L.wo); mats.push_back(L.wgu); mats.push_back(L.wdown);
        }
        if (stream_build_image(c->stream, c->dec_layers, H, I, c->V, mats.data(), c->emb, c->sk_cta_off, c->sk_image, 0) != 0)
            return set_err(QASR_ERR_CUDA, "%s", stream_error());
        if (stream_use_rounds(H)) { // second copy, round-major, for the single-sequence producer / consumer kernel (qasr_stream_r.cu)
            c->sk_image_r = (uint8_t *)dev_alloc(c, image_bytes);
            if (!c->sk_image_r) return set_err(QASR_ERR_NOMEM, "decode weight image, round-major (%zu bytes)", image_bytes);
            if (stream_build_image(c->stream, c->dec_layers, H, I, c->V, mats.data(), c->emb, c->sk_cta_off, c->sk_image_r, 1) != 0)
                return set_err(QASR_ERR_CUDA, "%s", stream_error());
        }
        CK(cudaStreamSynchronize(c->stream));
        const size_t NS = QASR_STREAM_MAX_SEQS;
        c->ll_qkv = (unsigned long long *)dalloc(NS * 4096 * 8); c->ll_att = (unsigned long long *)dalloc(NS * QASR_STREAM_ATT_WORDS * 8);
        c->ll_xwo = (unsigned long long *)dalloc(NS * H * 8); c->ll_act = (unsigned long long *)dalloc(NS * I * 8);
        c->ll_xdn = (unsigned long long *)dalloc(NS * H * 8); c->ll_head = (unsigned long long *)dalloc(NS * 2048 * 8);
        if (!c->ll_head) return set_err(QASR_ERR_NOMEM, "exchange buffers");
    }
    if (getenv("QASR_MEGA_PROF")) c->mega_prof = (long long *)dalloc(3 * 4096 * 8);
    if (!c->x || !c->logits || !c->d_gmax) return set_err(QASR_ERR_NOMEM, "state allocation failed");
    CK(cudaHostAlloc((void **)&c->h_tokens, (size_t)c->max_steps * QASR_STREAM_MAX_SEQS * 4, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void **)&c->dh_tokens, c->h_tokens, 0));
    CKR(ensure_kv(c, 2048, 0));
    CKR(ensure_rope(c, 4096));
    c->loaded = true;
    return 0;
}

int qasr_cuda_load_dir(qasr_ctx_t *c, const char *model_dir) {
    if (!c || !model_dir) return set_err(QASR_ERR_ARG, "null argument");
    if (c->loaded) return set_err(QASR_ERR_STATE, "context already holds a model");
    CK(cudaSetDevice(c->device));
    qst_dir_t *st = qst_open_dir(model_dir);
    if (!st) return set_err(QASR_ERR_MODEL, "cannot open safetensors in %s", model_dir);
    return load_from(c, st);
}

// The same upload from tensors the caller already holds in host memory - what the reference's loaders receive
// (multi_safetensors_t: name, dtype, shape, pointer into its own mmap; qwen_asr_safetensors.h:24-49), so
// qwen_encoder_load / qwen_decoder_load can forward them without knowing the checkpoint directory.
int qasr_cuda_upload_tensors(qasr_ctx_t *c, const qasr_tensor_t *tensors, int count) {
    if (!c || !tensors || count <= 0) return set_err(QASR_ERR_ARG, "null argument");
    if (c->loaded) return set_err(QASR_ERR_STATE, "context already holds a model");
    CK(cudaSetDevice(c->device));
    std::vector<qst_tensor_t> tab((size_t)count);
    for (int i = 0; i < count; i++) {
        const qasr_tensor_t &t = tensors[i];
        qst_tensor_t &o = tab[i];
        memset(&o, 0, sizeof o);
        if (!t.name || !t.data || t.ndim < 0 || t.ndim > 8 || (t.ndim > 0 && !t.shape)) return set_err(QASR_ERR_ARG, "tensor %d: null name / data / shape", i);
        snprintf(o.name, sizeof o.name, "%s", t.name);
        o.dtype = t.dtype == QASR_DTYPE_F32 ? QST_F32 : t.dtype == QASR_DTYPE_F16 ? QST_F16 : t.dtype == QASR_DTYPE_BF16 ? QST_BF16 : QST_OTHER;
        o.ndim = t.ndim;
        o.numel = 1;
        for (int k = 0; k < t.ndim; k++) { if (t.shape[k] < 0) return set_err(QASR_ERR_ARG, "tensor %s: negative extent", t.name); o.shape[k] = t.shape[k]; o.numel *= (size_t)t.shape[k]; }
        o.data = t.data;
        o.nbytes = o.numel * qst_elem_size(o.dtype);
    }
    qst_dir_t *st = qst_from_table(tab.data(), count);
    if (!st) return set_err(QASR_ERR_NOMEM, "tensor table");
    return load_from(c, st);
}

// ------------------------------------------------------------------ mel
int qasr_cuda_mel_frames(int n_samples) { return n_samples / 160; }

int qasr_cuda_encoder_tokens(int frames) {
    int T = 0;
    for (int s = 0; s < frames; s += 100) {
        int w = frames - s < 100 ? frames - s : 100;
        w = (w - 1) / 2 + 1; w = (w - 1) / 2 + 1; w = (w - 1) / 2 + 1;
        T += w;
    }
    return T;
}

// samples == NULL: use the n device-resident samples staged by qasr_cuda_stage_audio
static int mel_device(qasr_ctx_t *c, const float *samples, int n, int *frames_out) {
    const int frames = n / 160; // (n + 400 - 400)/160 + 1 - 1, reference :311-312
    if (frames <= 0) return set_err(QASR_ERR_ARG, "audio too short (%d samples)", n); // reference returns NULL (:313-317)
    if ((samples && c->ws_samples.reserve((size_t)n * 4)) || c->ws_meltmp.reserve((size_t)frames * 128 * 4) || c->ws_mel.reserve((size_t)frames * 128 * 4))
        return set_err(QASR_ERR_NOMEM, "mel workspace allocation failed");
    if (samples) CK(cudaMemcpyAsync(c->ws_samples.p, samples, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    launch_mel(c->stream, c->ws_samples.as<float>(), n, frames, c->mel_cos, c->mel_sin, c->mel_win, c->mel_fb,
               c->ws_meltmp.as<float>(), c->d_gmax, c->ws_mel.as<float>(), frames, 0);
    c->launches += 3;
    CK(cudaGetLastError());
    c->mel_frames = frames;
    *frames_out = frames;
    return 0;
}

int qasr_cuda_mel(qasr_ctx_t *c, const float *samples, int n_samples, float *mel_out, int *out_frames) {
    if (!c || !samples || !out_frames) return set_err(QASR_ERR_ARG, "null argument");
    if (!c->loaded) return set_err(QASR_ERR_STATE, "no model loaded");
    CK(cudaSetDevice(c->device));
    int frames = 0;
    CKR(mel_device(c, samples, n_samples, &frames));
    if (mel_out) CK(cudaMemcpyAsync(mel_out, c->ws_mel.p, (size_t)frames * 128 * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *out_frames = frames;
    return 0;
}

// ------------------------------------------------------------------ encoder
int gemm(qasr_ctx_t *c, const bf16_t *a_hi, const bf16_t *a_lo, int M, int K, const bf16_t *W, int N, int mode,
                float *of, bf16_t *ohi, bf16_t *olo, const float *bias, int ldo, const GemmEpilogue *norm) {
    GemmEpilogue e;
    if (norm) { e = *norm; if (c->nsplit != 2) e.nx_lo = nullptr; }
    e.mode = mode; e.out_f32 = of; e.out_hi = ohi; e.out_lo = (c->nsplit == 2) ? olo : nullptr; e.bias = bias; e.ldo = ldo;
    if (launch_gemm_tc(c->stream, a_hi, c->nsplit == 2 ? a_lo : nullptr, M, K, W, N, e) != 0)
        return set_err(QASR_ERR_CUDA, "%s", gemm_tc_error());
    c->launches += 1;
    return 0;
}

// mel (device, [128, frames_total]) -> encoder output rows ([T, H] f32).  The mel may hold several independent units side
// by side along the frame axis (unit_frames[0..n_units)): chunking, per-chunk conv padding, positional rows and the
// attention windows restart at every unit boundary (reference qwen_asr_encoder.c:188-297 applied per unit), while every
// GEMM runs over the rows of all units at once.  mel_stride = row stride of d_mel (0: the units' frames fill it exactly).
// out = NULL writes c->ws_encout.  One unit: the launch chain is captured
// per frame count in a CUDA graph.
int encode_units_device(qasr_ctx_t *c, const float *d_mel, int mel_stride, const int *unit_frames, int n_units, float *out, int *T_out) {
    const int d = c->d, F = c->F, H = c->H;
    int frames = 0, nc = 0;
    for (int u = 0; u < n_units; u++) { frames += unit_frames[u]; nc += (unit_frames[u] + 99) / 100; }
    std::vector<int> geom((size_t)nc * 2 + 3 * (nc + 1));
    int *w0 = geom.data(), *m0 = w0 + nc, *o1 = m0 + nc, *o2 = o1 + nc + 1, *o3 = o2 + nc + 1;
    o1[0] = o2[0] = o3[0] = 0;
    {
        int i = 0, f0 = 0;
        for (int u = 0; u < n_units; u++) {
            for (int s0 = 0; s0 < unit_frames[u]; s0 += 100, i++) {
                const int w = unit_frames[u] - s0 < 100 ? unit_frames[u] - s0 : 100;
                const int w1 = (w - 1) / 2 + 1, w2 = (w1 - 1) / 2 + 1, w3 = (w2 - 1) / 2 + 1;
                w0[i] = w; m0[i] = f0 + s0;
                o1[i + 1] = o1[i] + w1 * 64; o2[i + 1] = o2[i] + w2 * 32; o3[i + 1] = o3[i] + w3 * 16;
            }
            f0 += unit_frames[u];
        }
    }
    const int tot1 = o1[nc], tot2 = o2[nc], tot3 = o3[nc], T = tot3 / 16;
    // per-token PE row (position restarts in every chunk) and window starts: 13 * (800/100) = 104 tokens per window, counted
    // from the first token of each unit (reference qwen_asr_encoder.c:291-297)
    std::vector<int> aux;
    aux.reserve((size_t)T + T / 104 + n_units + 2);
    for (int i = 0; i < nc; i++) { const int w3 = (o3[i + 1] - o3[i]) / 16; for (int k = 0; k < w3; k++) aux.push_back(k); }
    int nwin = 0;
    {
        int t0 = 0;
        for (int u = 0; u < n_units; u++) {
            const int Tu = qasr_cuda_encoder_tokens(unit_frames[u]);
            for (int w = 0; w * 104 < Tu; w++) { aux.push_back(t0 + w * 104); nwin++; }
            t0 += Tu;
        }
        aux.push_back(T);
    }
    if (c->ws_geom.reserve((geom.size() + aux.size()) * 4)) return set_err(QASR_ERR_NOMEM, "geom alloc");
    int *dg = c->ws_geom.as<int>();
    const int geom_key = n_units == 1 ? frames : -1; // the tables of a single unit depend on its frame count only: same length, nothing to upload and no host sync
    if (geom_key < 0 || c->geom_frames != geom_key || c->ws_geom.grew) {
        c->geom_frames = 0;
        CK(cudaMemcpyAsync(dg, geom.data(), geom.size() * 4, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(dg + geom.size(), aux.data(), aux.size() * 4, cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream)); // host vectors go out of scope below
        c->geom_frames = geom_key < 0 ? 0 : geom_key;
    }
    ConvGeom g;
    g.n_chunks = nc; g.d_w0 = dg; g.d_mel0 = dg + nc; g.d_off1 = dg + 2 * nc; g.d_off2 = g.d_off1 + nc + 1; g.d_off3 = g.d_off2 + nc + 1;
    g.total1 = tot1; g.total2 = tot2; g.total3 = tot3;
    const int *d_rowpos = dg + geom.size(), *d_win = d_rowpos + T;

    // workspace carve (all offsets 256-byte aligned)
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    static int conv_im2col = -1; // QASR_CONV_IM2COL=1: round 1's materialised patch matrices (A/B runs)
    if (conv_im2col < 0) { const char *e = getenv("QASR_CONV_IM2COL"); conv_im2col = e && e[0] == '1'; }
    const size_t o_act1 = carve((size_t)tot1 * 480 * 2 * 2), o_col2 = carve(conv_im2col ? (size_t)tot2 * 4320 * 2 * 2 : 0),
                 o_act2 = carve((size_t)tot2 * 480 * 2 * 2), o_col3 = carve(conv_im2col ? (size_t)tot3 * 4320 * 2 * 2 : 0),
                 o_act3 = carve((size_t)tot3 * 480 * 2 * 2), o_x = carve((size_t)T * d * 4),
                 o_xn = carve((size_t)T * d * 2 * 2), o_qkv = carve((size_t)T * 3 * d * 4),
                 o_att = carve((size_t)T * d * 2 * 2), o_mid = carve((size_t)T * F * 2 * 2);
    if (c->ws_enc.reserve(off) || (!out && c->ws_encout.reserve((size_t)T * H * 4))) return set_err(QASR_ERR_NOMEM, "encoder workspace (%zu bytes)", off);
    note_growth(c, c->ws_enc); note_growth(c, c->ws_encout); note_growth(c, c->ws_geom); note_growth(c, c->ws_mel);
    if (!out) out = c->ws_encout.as<float>();
    uint8_t *B = c->ws_enc.as<uint8_t>();
    auto enqueue = [&]() -> int {
#define HI(o) reinterpret_cast<bf16_t *>(B + (o))
#define LO(o, n) (reinterpret_cast<bf16_t *>(B + (o)) + (size_t)(n))
    cudaStream_t s = c->stream;
    const bool two = c->nsplit == 2;
    // conv stem, reference qwen_asr_encoder.c:221-276
    launch_conv1(s, d_mel, mel_stride > 0 ? mel_stride : frames, c->c1w, c->c1b, g, HI(o_act1), two ? LO(o_act1, (size_t)tot1 * 480) : nullptr);
    if (!conv_im2col) { // implicit GEMM: the kernel gathers the 3 x 3 patches itself
        GemmEpilogue e2;
        e2.mode = QASR_GEMM_GELU_SPLIT; e2.out_f32 = nullptr; e2.out_hi = HI(o_act2); e2.out_lo = two ? LO(o_act2, (size_t)tot2 * 480) : nullptr; e2.bias = c->c2b; e2.ldo = 480;
        if (launch_conv_gemm_tc(s, HI(o_act1), two ? LO(o_act1, (size_t)tot1 * 480) : nullptr, g, 2, c->c2w, e2) != 0) return set_err(QASR_ERR_CUDA, "%s", gemm_tc_error());
        GemmEpilogue e3 = e2;
        e3.out_hi = HI(o_act3); e3.out_lo = two ? LO(o_act3, (size_t)tot3 * 480) : nullptr; e3.bias = c->c3b;
        if (launch_conv_gemm_tc(s, HI(o_act2), two ? LO(o_act2, (size_t)tot2 * 480) : nullptr, g, 3, c->c3w, e3) != 0) return set_err(QASR_ERR_CUDA, "%s", gemm_tc_error());
        c->launches += 2;
    } else {
    launch_im2col_stage(s, HI(o_act1), HI(o_col2), g, 2);
    if (two) launch_im2col_stage(s, LO(o_act1, (size_t)tot1 * 480), LO(o_col2, (size_t)tot2 * 4320), g, 2);
    c->launches += two ? 2 : 1;
    CKR(gemm(c, HI(o_col2), LO(o_col2, (size_t)tot2 * 4320), tot2, 4320, c->c2w, 480, QASR_GEMM_GELU_SPLIT, nullptr,
             HI(o_act2), LO(o_act2, (size_t)tot2 * 480), c->c2b, 480));
    launch_im2col_stage(s, HI(o_act2), HI(o_col3), g, 3);
    if (two) launch_im2col_stage(s, LO(o_act2, (size_t)tot2 * 480), LO(o_col3, (size_t)tot3 * 4320), g, 3);
    c->launches += two ? 2 : 1;
    CKR(gemm(c, HI(o_col3), LO(o_col3, (size_t)tot3 * 4320), tot3, 4320, c->c3w, 480, QASR_GEMM_GELU_SPLIT, nullptr,
             HI(o_act3), LO(o_act3, (size_t)tot3 * 480), c->c3b, 480));
    }
    float *x = reinterpret_cast<float *>(B + o_x);
    CKR(gemm(c, HI(o_act3), LO(o_act3, (size_t)tot3 * 480), T, 7680, c->conv_out, d, QASR_GEMM_F32, x, nullptr, nullptr, nullptr, d));
    launch_add_rows(s, x, c->pe, d_rowpos, T, d);
    c->launches += 1;
    // transformer, reference qwen_asr_encoder.c:312-347
    float *qkv = reinterpret_cast<float *>(B + o_qkv);
    bf16_t *xn_hi = HI(o_xn), *xn_lo = LO(o_xn, (size_t)T * d), *at_hi = HI(o_att), *at_lo = LO(o_att, (size_t)T * d),
           *mid_hi = HI(o_mid), *mid_lo = LO(o_mid, (size_t)T * F);
    for (int l = 0; l < c->enc_layers; l++) {
        const EncLayerW &L = c->enc[l];
        launch_layernorm(s, x, L.ln1w, L.ln1b, 1e-5f, T, d, nullptr, xn_hi, two ? xn_lo : nullptr);
        CKR(gemm(c, xn_hi, xn_lo, T, d, L.wqkv, 3 * d, QASR_GEMM_F32, qkv, nullptr, nullptr, L.bqkv, 3 * d));
        launch_attn_windowed(s, qkv, qkv + d, qkv + 2 * d, 3 * d, c->enc_heads, d_win, nwin, 104, 0.125f, d, nullptr, at_hi, two ? at_lo : nullptr);
        CKR(gemm(c, at_hi, at_lo, T, d, L.wo, d, QASR_GEMM_RESIDUAL, x, nullptr, nullptr, L.bo, d));
        launch_layernorm(s, x, L.ln2w, L.ln2b, 1e-5f, T, d, nullptr, xn_hi, two ? xn_lo : nullptr);
        CKR(gemm(c, xn_hi, xn_lo, T, d, L.fc1, F, QASR_GEMM_GELU_SPLIT, nullptr, mid_hi, mid_lo, L.fc1b, F));
        CKR(gemm(c, mid_hi, mid_lo, T, F, L.fc2, d, QASR_GEMM_RESIDUAL, x, nullptr, nullptr, L.fc2b, d));
        c->launches += 3;
    }
    // tail, reference qwen_asr_encoder.c:350-361
    launch_layernorm(s, x, c->lnpw, c->lnpb, 1e-5f, T, d, nullptr, xn_hi, two ? xn_lo : nullptr);
    CKR(gemm(c, xn_hi, xn_lo, T, d, c->p1w, d, QASR_GEMM_GELU_SPLIT, nullptr, at_hi, at_lo, c->p1b, d));
    CKR(gemm(c, at_hi, at_lo, T, d, c->p2w, H, QASR_GEMM_F32, out, nullptr, nullptr, c->p2b, H));
    c->launches += 1;
    return 0;
    };
    {
        PdlScope pdl(T <= 256);
        if (n_units == 1 && out == c->ws_encout.as<float>()) CKR(run_cached_graph(c, 1, frames, c->nsplit, enqueue));
        else CKR(enqueue());
    }
#undef HI
#undef LO
    CK(cudaGetLastError());
    c->enc_T = T;
    *T_out = T;
    return 0;
}
static int encode_device(qasr_ctx_t *c, const float *d_mel, int frames, int *T_out) { return encode_units_device(c, d_mel, 0, &frames, 1, nullptr, T_out); }

int qasr_cuda_encode(qasr_ctx_t *c, const float *mel, int mel_frames, float *enc_out, int *out_tokens) {
    if (!c || !out_tokens) return set_err(QASR_ERR_ARG, "null argument");
    if (!c->loaded) return set_err(QASR_ERR_STATE, "no model loaded");
    CK(cudaSetDevice(c->device));
    const float *d_mel;
    if (mel) {
        if (mel_frames <= 0) return set_err(QASR_ERR_ARG, "mel_frames must be positive");
        if (c->ws_mel.reserve((size_t)mel_frames * 128 * 4)) return set_err(QASR_ERR_NOMEM, "mel alloc");
        CK(cudaMemcpyAsync(c->ws_mel.p, mel, (size_t)mel_frames * 128 * 4, cudaMemcpyHostToDevice, c->stream));
        c->mel_frames = mel_frames;
    } else {
        if (c->mel_frames <= 0) return set_err(QASR_ERR_STATE, "no device-resident mel: call qasr_cuda_mel first");
        mel_frames = c->mel_frames;
    }
    d_mel = c->ws_mel.as<float>();
    int T = 0;
    CKR(encode_device(c, d_mel, mel_frames, &T));
    if (enc_out) CK(cudaMemcpyAsync(enc_out, c->ws_encout.p, (size_t)T * c->H * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *out_tokens = T;
    return 0;
}

// ------------------------------------------------------------------ decoder prefill
// x: device [P, H] f32 rows inside ws_pre (offset 0).  reference qwen_asr_decoder.c:457-563
static int prefill_device(qasr_ctx_t *c, int P, int kv_len) {
    const int H = c->H, I = c->I;
    CKR(ensure_kv(c, kv_len + P + 1, kv_len));
    CKR(ensure_rope(c, kv_len + P + 1));
    uint8_t *B = c->ws_pre.as<uint8_t>();
    size_t off = align_up((size_t)P * H * 4, 256);
    auto carve = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    const size_t o_xn = carve((size_t)P * H * 4), o_qkv = carve((size_t)P * 4096 * 4), o_q = carve((size_t)P * 2048 * 4),
                 o_att = carve((size_t)P * 2048 * 4), o_act = carve((size_t)P * I * 4), o_ssq = carve((size_t)P * (H >> 7) * 4);
    if (off > c->ws_pre.cap) return set_err(QASR_ERR_STATE, "prefill workspace not reserved");
    float *x = reinterpret_cast<float *>(B);
    float *ssq = reinterpret_cast<float *>(B + o_ssq); // [P][H / 128] sums of squares of the fused RMSNorms
    const bool fuse = gemm_tc_can_fuse_norm(P, 2048, H) && gemm_tc_can_fuse_norm(P, I, H);
    const bool fuse_qk = gemm_tc_can_fuse_qk(P, H); // q/k-norm + RoPE + KV store in the QKV epilogue (GemmEpilogue::qk_*)
    bf16_t *xn_hi = reinterpret_cast<bf16_t *>(B + o_xn), *xn_lo = xn_hi + (size_t)P * H;
    float *qkv = reinterpret_cast<float *>(B + o_qkv), *q = reinterpret_cast<float *>(B + o_q);
    bf16_t *at_hi = reinterpret_cast<bf16_t *>(B + o_att), *at_lo = at_hi + (size_t)P * 2048;
    bf16_t *ac_hi = reinterpret_cast<bf16_t *>(B + o_act), *ac_lo = ac_hi + (size_t)P * I;
    cudaStream_t s = c->stream;
    const bool two = c->nsplit == 2;
    const size_t kvd = (size_t)c->kv_heads * c->hd;
    const float scale = 1.0f / sqrtf((float)c->hd);
    note_growth(c, c->ws_pre);
    // QASR_PREFILL_ABLATE (timing experiments only, results are wrong): bit mask of launches to leave out of a layer -
    // 1 RMSNorms, 2 q/k-norm+RoPE+KV store, 4 attention, 8 QKV, 16 WO, 32 gate/up, 64 down
    static int ablate = -1;
    if (ablate < 0) { const char *e = getenv("QASR_PREFILL_ABLATE"); ablate = e ? atoi(e) : 0; }
    auto enqueue = [&]() -> int {
    for (int l = 0; l < c->dec_layers; l++) {
        const DecLayerW &L = c->dec[l];
        float *kc = c->kv_k + (size_t)l * c->kv_max * kvd, *vc = c->kv_v + (size_t)l * c->kv_max * kvd;
        // RMSNorm: a stand-alone launch (layer 0, long prompts, ablation runs), or fused across the GEMMs on both sides of it: the
        // producer of x writes the planes of x * gamma and per-tile sums of squares, the consumer scales its rows (GemmEpilogue::nx_*)
        GemmEpilogue from_x, to_post, to_next; // consumer side of both norms; producer side of the post-attention / next input norm
        from_x.in_ssq = ssq; from_x.in_tiles = H >> 7; from_x.in_eps = 1e-6f;
        to_post.nx_gamma = L.post_norm; to_post.nx_hi = xn_hi; to_post.nx_lo = xn_lo; to_post.nx_ssq = ssq;
        to_next = to_post;
        to_next.nx_gamma = l + 1 < c->dec_layers ? c->dec[l + 1].in_norm : nullptr;
        const bool next_fused = fuse && to_next.nx_gamma;
        if (!(ablate & 1) && !(fuse && l > 0)) launch_rmsnorm(s, x, L.in_norm, 1e-6f, P, H, nullptr, xn_hi, two ? xn_lo : nullptr);
        GemmEpilogue qkv_epi; // consumer side of the input norm (layers > 0) + q/k-norm, RoPE and the KV store of this layer
        if (fuse && l > 0) qkv_epi = from_x;
        if (fuse_qk) {
            qkv_epi.qk_q = q; qkv_epi.qk_kc = kc; qkv_epi.qk_vc = vc; qkv_epi.qk_qn = L.qn; qkv_epi.qk_kn = L.kn;
            qkv_epi.qk_cos = c->rope_cos; qkv_epi.qk_sin = c->rope_sin; qkv_epi.qk_pos0 = kv_len; qkv_epi.qk_eps = 1e-6f;
        }
        if (!(ablate & 8)) CKR(gemm(c, xn_hi, xn_lo, P, H, L.wqkv, 4096, QASR_GEMM_F32, qkv, nullptr, nullptr, nullptr, 4096, &qkv_epi));
        if (!(ablate & 2) && !fuse_qk) launch_qk_norm_rope_store(s, qkv, L.qn, L.kn, c->rope_cos, c->rope_sin, kv_len, P, 1e-6f, q, kc, vc);
        if (!(ablate & 4)) launch_attn_prefill(s, q, kc, vc, kv_len, P, kv_len + P, c->heads, c->kv_heads, scale, nullptr, at_hi, two ? at_lo : nullptr);
        if (!(ablate & 16)) CKR(gemm(c, at_hi, at_lo, P, 2048, L.wo, H, QASR_GEMM_RESIDUAL, x, nullptr, nullptr, nullptr, H, fuse ? &to_post : nullptr));
        if (!(ablate & 1) && !fuse) launch_rmsnorm(s, x, L.post_norm, 1e-6f, P, H, nullptr, xn_hi, two ? xn_lo : nullptr);
        if (!(ablate & 32)) CKR(gemm(c, xn_hi, xn_lo, P, H, L.wgu, 2 * I, QASR_GEMM_SWIGLU_SPLIT, nullptr, ac_hi, ac_lo, nullptr, I, fuse ? &from_x : nullptr));
        if (!(ablate & 64)) CKR(gemm(c, ac_hi, ac_lo, P, I, L.wdown, H, QASR_GEMM_RESIDUAL, x, nullptr, nullptr, nullptr, H, next_fused ? &to_next : nullptr));
        c->launches += (fuse ? (l > 0 ? 2 : 3) : 4) - (fuse_qk ? 1 : 0);
    }
    return 0;
    };
    {
        PdlScope pdl(P <= 256);
        CKR(run_cached_graph(c, 2, P, (long long)kv_len * 8 + c->nsplit + ((long long)c->seq << 48), enqueue));
    }
    c->kv_fill[c->seq] = kv_len + P;
    CK(cudaGetLastError());
    return 0;
}

static int reserve_prefill(qasr_ctx_t *c, int P) {
    const size_t H = c->H, I = c->I, p = P;
    size_t bytes = align_up(p * H * 4, 256) * 2 + align_up(p * 4096 * 4, 256) + align_up(p * 2048 * 4, 256) * 2 + align_up(p * I * 4, 256) +
                   align_up(p * (H >> 7) * 4, 256);
    if (c->ws_pre.reserve(bytes)) return set_err(QASR_ERR_NOMEM, "prefill workspace (%zu bytes)", bytes);
    return 0;
}

int qasr_cuda_prefill_embeds(qasr_ctx_t *c, const float *embeds, int seq_len, int kv_len) {
    if (c) c->kv_epoch++;
    if (!c || !embeds) return set_err(QASR_ERR_ARG, "null argument");
    if (!c->loaded) return set_err(QASR_ERR_STATE, "no model loaded");
    if (seq_len <= 0) return 0;
    if (kv_len < 0) return set_err(QASR_ERR_ARG, "negative kv_len");
    CK(cudaSetDevice(c->device));
    CKR(reserve_prefill(c, seq_len));
    CK(cudaMemcpyAsync(c->ws_pre.p, embeds, (size_t)seq_len * c->H * 4, cudaMemcpyHostToDevice, c->stream));
    CKR(prefill_device(c, seq_len, kv_len));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// rows = embed(pre) | encoder rows | embed(suf); prefill all but the last row, keep it pending.
static int prefill_prompt_device(qasr_ctx_t *c, const int *pre, int n_pre, int n_audio, const int *suf, int n_suf, int kv_len) {
    if (n_audio > c->enc_T) return set_err(QASR_ERR_STATE, "n_audio=%d exceeds the %d encoder rows on the device", n_audio, c->enc_T);
    const int total = n_pre + n_audio + n_suf, H = c->H;
    if (total < 1) return set_err(QASR_ERR_ARG, "empty prompt");
    CKR(reserve_prefill(c, total));
    if (c->ws_ids.reserve((size_t)(n_pre + n_suf + 1) * 4)) return set_err(QASR_ERR_NOMEM, "ids alloc");
    int *d_ids = c->ws_ids.as<int>();
    float *x = c->ws_pre.as<float>();
    if (n_pre) CK(cudaMemcpyAsync(d_ids, pre, (size_t)n_pre * 4, cudaMemcpyHostToDevice, c->stream));
    if (n_suf) CK(cudaMemcpyAsync(d_ids + n_pre, suf, (size_t)n_suf * 4, cudaMemcpyHostToDevice, c->stream));
    launch_embed_gather(c->stream, c->emb, d_ids, n_pre, H, x);
    if (n_audio) CK(cudaMemcpyAsync(x + (size_t)n_pre * H, c->ws_encout.p, (size_t)n_audio * H * 4, cudaMemcpyDeviceToDevice, c->stream));
    launch_embed_gather(c->stream, c->emb, d_ids + n_pre, n_suf, H, x + (size_t)(n_pre + n_audio) * H);
    c->launches += 2;
    CK(cudaMemcpyAsync(c->pending, x + (size_t)(total - 1) * H, (size_t)H * 4, cudaMemcpyDeviceToDevice, c->stream));
    c->has_pending = true;
    if (total > 1) CKR(prefill_device(c, total - 1, kv_len));
    return 0;
}

int qasr_cuda_prefill_prompt(qasr_ctx_t *c, const int *pre_ids, int n_pre, int n_audio, const int *suf_ids, int n_suf, int kv_len) {
    if (c) c->kv_epoch++;
    if (!c || (n_pre > 0 && !pre_ids) || (n_suf > 0 && !suf_ids)) return set_err(QASR_ERR_ARG, "null argument");
    if (!c->loaded) return set_err(QASR_ERR_STATE, "no model loaded");
    if (n_pre < 0 || n_audio < 0 || n_suf < 0 || kv_len < 0) return set_err(QASR_ERR_ARG, "negative count");
    for (int i = 0; i < n_pre; i++) if (pre_ids[i] < 0 || pre_ids[i] >= c->V) return set_err(QASR_ERR_ARG, "token id out of range");
    for (int i = 0; i < n_suf; i++) if (suf_ids[i] < 0 || suf_ids[i] >= c->V) return set_err(QASR_ERR_ARG, "token id out of range");
    CK(cudaSetDevice(c->device));
    CKR(prefill_prompt_device(c, pre_ids, n_pre, n_audio, suf_ids, n_suf, kv_len));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------ decode step
// One token through the 28 blocks, input row in c->x, position in *d_pos.
// reference qwen_asr_decoder.c:632-678.  5 launches per layer.
static void enqueue_layers(qasr_ctx_t *c, cudaStream_t s) {
    const int H = c->H, I = c->I;
    const size_t kvd = (size_t)c->kv_heads * c->hd;
    for (int l = 0; l < c->dec_layers; l++) {
        const DecLayerW &L = c->dec[l];
        float *kc = c->kv_k + (size_t)l * c->kv_max * kvd, *vc = c->kv_v + (size_t)l * c->kv_max * kvd;
        launch_gemv_bf16(s, L.wqkv, c->x, L.in_norm, 1e-6f, c->qkv, nullptr, nullptr, 4096, H, QASR_EPI_STORE, c->d_done);
        launch_attn_decode(s, c->qkv, L.qn, L.kn, c->rope_cos, c->rope_sin, kc, vc, c->d_pos, c->attn_part, c->counters, c->attn, 1e-6f);
        launch_gemv_bf16(s, L.wo, c->attn, nullptr, 0.f, c->x, c->x, nullptr, H, 2048, QASR_EPI_RESIDUAL, c->d_done);
        launch_gemv_bf16(s, L.wgu, c->x, L.post_norm, 1e-6f, c->act, nullptr, nullptr, 2 * I, H, QASR_EPI_SWIGLU, c->d_done);
        launch_gemv_bf16(s, L.wdown, c->act, nullptr, 0.f, c->x, c->x, nullptr, H, I, QASR_EPI_RESIDUAL, c->d_done);
    }
}
static void enqueue_greedy_head(qasr_ctx_t *c, cudaStream_t s) {
    launch_argmax_gemv(s, c->emb, c->x, c->final_norm, 1e-6f, c->V, c->H, c->part_val, c->part_idx, c->d_done);
    launch_argmax_finalize(s, c->part_val, c->part_idx, c->n_parts, c->emb, c->H, c->x, c->d_tokens, c->d_step, c->d_pos,
                           c->d_done, c->dh_tokens, c->max_steps);
}
static const int kStepKernels = 28 * 5 + 2;

static int ensure_graph(qasr_ctx_t *c) {
    if (!c->use_graph || c->graph_exec) return 0;
    if (c->graph) { cudaGraphDestroy(c->graph); c->graph = nullptr; }
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    enqueue_layers(c, c->stream);
    enqueue_greedy_head(c, c->stream);
    CK(cudaStreamEndCapture(c->stream, &c->graph));
    CK(cudaGraphInstantiate(&c->graph_exec, c->graph, 0));
    return 0;
}

// Enqueue n greedy steps (each consumes c->x, leaves the next embedding in c->x).
static int enqueue_steps(qasr_ctx_t *c, int n, int nseq = 1) {
    if (nseq != 1 && !c->use_stream) return set_err(QASR_ERR_STATE, "batched decode needs the stream kernel (unset QASR_DECODE)");
    if (c->use_stream) { // one persistent cooperative launch runs all n steps (stops itself after an EOS token)
        StreamParams p = {};
        p.image = c->sk_image; p.image_r = c->sk_image_r; p.cta_off = c->sk_cta_off;
        p.n_layers = c->dec_layers; p.H = c->H; p.I = c->I; p.V = c->V; p.n_steps = n; p.eps = 1e-6f;
        p.emb = c->emb; p.final_norm = c->final_norm;
        for (int l = 0; l < c->dec_layers; l++) {
            const DecLayerW &L = c->dec[l];
            p.in_norm[l] = L.in_norm; p.post_norm[l] = L.post_norm; p.qn[l] = L.qn; p.kn[l] = L.kn;
        }
        p.x_io = c->x;
        p.nseq = nseq;
        for (int q = 0; q < QASR_STREAM_MAX_SEQS; q++) { p.kv_k[q] = nseq == 1 ? c->kv_k : c->kv_ks[q]; p.kv_v[q] = nseq == 1 ? c->kv_v : c->kv_vs[q]; }
        p.kv_layer_stride = (size_t)c->kv_max * c->kv_heads * c->hd;
        p.rope_cos = c->rope_cos; p.rope_sin = c->rope_sin;
        p.ll_qkv = c->ll_qkv; p.ll_att = c->ll_att; p.ll_xwo = c->ll_xwo; p.ll_act = c->ll_act; p.ll_xdn = c->ll_xdn; p.ll_head = c->ll_head;
        const unsigned span = (unsigned)n * (unsigned)(c->dec_layers + 1) + 1;
        if (c->sk_tag > 0xF0000000u) { // tag space exhausted: clear the exchange buffers and start over
            const size_t NS = QASR_STREAM_MAX_SEQS;
            CK(cudaMemsetAsync(c->ll_qkv, 0, NS * 4096 * 8, c->stream)); CK(cudaMemsetAsync(c->ll_att, 0, NS * QASR_STREAM_ATT_WORDS * 8, c->stream));
            CK(cudaMemsetAsync(c->ll_xwo, 0, NS * c->H * 8, c->stream)); CK(cudaMemsetAsync(c->ll_act, 0, NS * c->I * 8, c->stream));
            CK(cudaMemsetAsync(c->ll_xdn, 0, NS * c->H * 8, c->stream)); CK(cudaMemsetAsync(c->ll_head, 0, NS * 2048 * 8, c->stream));
            c->sk_tag = 1;
        }
        p.tag_base = c->sk_tag;
        c->sk_tag += span;
        p.d_pos = c->d_pos; p.d_step = c->d_step; p.d_tokens = c->d_tokens; p.h_tokens = c->dh_tokens;
        p.prof = c->mega_prof; p.prof_cap = 4096;
        { const char *dbg = getenv("QASR_MEGA_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
        { const char *tc = getenv("QASR_MEGA_TRACE_CTA"); p.trace_cta = tc ? atoi(tc) : 0; }
        { const char *e = getenv("QASR_SK_L2AHEAD"); p.l2_ahead_units = e ? atoi(e) : 8; }
        { const char *e = getenv("QASR_SK_L2ISSUE"); p.l2_issue = e ? atoi(e) : 2; }
        p.dbg_logits = c->dbg_logits; p.dbg_hidden = c->dbg_hidden; // test hooks, NULL outside qasr_cuda_step_logits / qasr_debug_stream_step
        if (launch_decode_stream(c->stream, p) != 0) return set_err(QASR_ERR_CUDA, "%s", stream_error());
        c->launches += 1;
        return 0;
    }
    if (c->use_graph) {
        CKR(ensure_graph(c));
        for (int i = 0; i < n; i++) CK(cudaGraphLaunch(c->graph_exec, c->stream));
    } else {
        for (int i = 0; i < n; i++) { enqueue_layers(c, c->stream); enqueue_greedy_head(c, c->stream); }
        CK(cudaGetLastError());
    }
    c->launches += (long long)n * kStepKernels;
    return 0;
}

static int step_common(qasr_ctx_t *c, int kv_len, int *out_token) {
    if (c) c->kv_epoch++;
    CKR(ensure_kv(c, kv_len + 2, kv_len));
    CKR(ensure_rope(c, kv_len + 2));
    launch_set_state(c->stream, c->d_pos, kv_len, c->d_done, 0, c->d_step, 0);
    c->launches += 1;
    CK(cudaEventRecord(c->ev[0], c->stream));
    CKR(enqueue_steps(c, 1));
    CK(cudaEventRecord(c->ev[1], c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    c->last_decode_ms = ms;
    const int tok = c->h_tokens[0];
    c->x_token = tok;
    if (out_token) *out_token = tok;
    return 0;
}

int qasr_cuda_step_embed(qasr_ctx_t *c, const float *embed, int kv_len, int *out_token) {
    if (!c || !embed || !out_token) return set_err(QASR_ERR_ARG, "null argument");
    if (!c->loaded) return set_err(QASR_ERR_STATE, "no model loaded");
    if (kv_len < 0) return set_err(QASR_ERR_ARG, "negative kv_len");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->x, embed, (size_t)c->H * 4, cudaMemcpyHostToDevice, c->stream));
    return step_common(c, kv_len, out_token);
}

static int load_token_embedding(qasr_ctx_t *c, int token_id) {
    if (token_id < 0 || token_id >= c->V) return set_err(QASR_ERR_ARG, "token id %d out of range", token_id);
    if (c->x_token == token_id) return 0; // already gathered by the previous step's finalize
    if (c->ws_ids.reserve(64)) return set_err(QASR_ERR_NOMEM, "ids alloc");
    CK(cudaMemcpyAsync(c->ws_ids.p, &token_id, 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream)); // token_id is a stack variable
    launch_embed_gather(c->stream, c->emb, c->ws_ids.as<int>(), 1, c->H, c->x);
    c->launches += 1;
    c->x_token = token_id;
    return 0;
}

int qasr_cuda_step_token(qasr_ctx_t *c, int token_id, int kv_len, int *out_token) {
    if (!c || !out_token) return set_err(QASR_ERR_ARG, "null argument");
    if (!c->loaded) return set_err(QASR_ERR_STATE, "no model loaded");
    if (kv_len < 0) return set_err(QASR_ERR_ARG, "negative kv_len");
    CK(cudaSetDevice(c->device));
    CKR(load_token_embedding(c, token_id));
    return step_common(c, kv_len, out_token);
}

int qasr_cuda_step_pending(qasr_ctx_t *c, int kv_len, int *out_token) {
    if (!c || !out_token) return set_err(QASR_ERR_ARG, "null argument");
    if (!c->loaded || !c->has_pending) return set_err(QASR_ERR_STATE, "no pending row: call qasr_cuda_prefill_prompt first");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->x, c->pending, (size_t)c->H * 4, cudaMemcpyDeviceToDevice, c->stream));
    c->has_pending = false;
    return step_common(c, kv_len, out_token);
}

// Full logits of one step.  The default route is the SAME kernel the greedy path runs (decode_stream_kernel, one step,
// with its HEAD phase also storing y * rsqrt(mean x^2 + eps) per vocab row), so every logits-level parity test pins the
// production kernel; QASR_DECODE=graph takes the per-phase kernels + the lm_head GEMV instead.
int qasr_cuda_step_logits(qasr_ctx_t *c, const float *embed, int kv_len, float *logits) {
    if (c) c->kv_epoch++;
    if (!c || !embed || !logits) return set_err(QASR_ERR_ARG, "null argument");
    if (!c->loaded) return set_err(QASR_ERR_STATE, "no model loaded");
    if (kv_len < 0) return set_err(QASR_ERR_ARG, "negative kv_len");
    CK(cudaSetDevice(c->device));
    CKR(ensure_kv(c, kv_len + 2, kv_len));
    CKR(ensure_rope(c, kv_len + 2));
    CK(cudaMemcpyAsync(c->x, embed, (size_t)c->H * 4, cudaMemcpyHostToDevice, c->stream));
    launch_set_state(c->stream, c->d_pos, kv_len, c->d_done, 0, c->d_step, 0);
    if (c->use_stream) {
        c->dbg_logits = c->logits;
        const int rc = enqueue_steps(c, 1);
        c->dbg_logits = nullptr;
        CKR(rc);
        c->launches += 1;
    } else {
        enqueue_layers(c, c->stream);
        // final RMSNorm fused into the lm_head GEMV (reference qwen_asr_decoder.c:781-782)
        launch_gemv_bf16(c->stream, c->emb, c->x, c->final_norm, 1e-6f, c->logits, nullptr, nullptr, c->V, c->H, QASR_EPI_STORE, nullptr);
        c->launches += 28 * 5 + 2;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(logits, c->logits, (size_t)c->V * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->kv_fill[c->seq] = kv_len + 1;
    c->x_token = -1;
    return 0;
}

// ---- test hooks of the batched decode kernel (not part of the public header; bound by tests only)
// qasr_debug_select_seq: route the single-sequence entry points (prefill_embeds, read_kv, ...) to KV cache `q`, so a test
// can fill several sequence caches with different prompts.
extern "C" int qasr_debug_select_seq(qasr_ctx_t *c, int q) {
    if (!c || !c->loaded || q < 0 || q >= QASR_STREAM_MAX_SEQS) return set_err(QASR_ERR_ARG, "bad sequence index");
    CK(cudaSetDevice(c->device));
    select_seq(c, q);
    CKR(ensure_kv(c, c->kv_max > 0 ? c->kv_max : 2048, c->kv_fill[q]));
    return 0;
}
// qasr_debug_stream_step: ONE step of decode_stream_kernel<nseq> (nseq = 1, 2, 4) on the caches of sequences 0..nseq-1:
// embeds [nseq][H] in, greedy tokens [nseq], full logits [nseq][V] and the post-final-norm hidden states [nseq][H] out
// (reference twin: qwen_decoder_forward_logits, qwen_asr_decoder.c:691-783, called once per sequence).
extern "C" int qasr_debug_stream_step(qasr_ctx_t *c, int nseq, const float *embeds, const int *kv_lens, int *tokens, float *logits, float *hidden) {
    if (c) c->kv_epoch++;
    if (!c || !embeds || !kv_lens) return set_err(QASR_ERR_ARG, "null argument");
    if (!c->loaded || !c->use_stream) return set_err(QASR_ERR_STATE, "needs a loaded model and the stream decode kernel");
    if (nseq != 1 && nseq != 2 && nseq != 4) return set_err(QASR_ERR_ARG, "nseq must be 1, 2 or 4");
    if (nseq > qasr_cuda_max_batch(c)) return set_err(QASR_ERR_ARG, "nseq exceeds qasr_cuda_max_batch");
    CK(cudaSetDevice(c->device));
    int cap = 0;
    for (int q = 0; q < nseq; q++) if (kv_lens[q] + 2 > cap) cap = kv_lens[q] + 2;
    for (int q = 0; q < nseq; q++) { select_seq(c, q); CKR(ensure_kv(c, cap, kv_lens[q])); }
    select_seq(c, 0);
    CKR(ensure_rope(c, cap));
    float *d_logits = nullptr;
    if (logits) CK(cudaMalloc(&d_logits, (size_t)nseq * c->V * 4));
    CK(cudaMemcpyAsync(c->x, embeds, (size_t)nseq * c->H * 4, cudaMemcpyHostToDevice, c->stream));
    launch_set_state(c->stream, c->d_pos, kv_lens[0], c->d_done, 0, c->d_step, 0);
    CK(cudaMemcpyAsync(c->d_pos, kv_lens, sizeof(int) * nseq, cudaMemcpyHostToDevice, c->stream));
    c->dbg_logits = d_logits; c->dbg_hidden = hidden ? c->hidden_buf : nullptr;
    const int rc = enqueue_steps(c, 1, nseq);
    c->dbg_logits = nullptr; c->dbg_hidden = nullptr;
    if (rc == 0 && logits) cudaMemcpyAsync(logits, d_logits, (size_t)nseq * c->V * 4, cudaMemcpyDeviceToHost, c->stream);
    if (rc == 0 && hidden) cudaMemcpyAsync(hidden, c->hidden_buf, (size_t)nseq * c->H * 4, cudaMemcpyDeviceToHost, c->stream);
    const cudaError_t se = cudaStreamSynchronize(c->stream);
    if (d_logits) cudaFree(d_logits);
    if (rc != 0) return rc;
    if (se != cudaSuccess) return set_err(QASR_ERR_CUDA, "stream step failed: %s", cudaGetErrorString(se));
    for (int q = 0; q < nseq; q++) { if (tokens) tokens[q] = c->h_tokens[q]; c->kv_fill[q] = kv_lens[q] + 1; }
    c->x_token = -1;
    return 0;
}

// Greedy loop on the device; only ids cross PCIe.  reference qwen_asr.c:788-818
static int generate_device(qasr_ctx_t *c, int first_token, int kv_len, int max_new, int *out_ids, int *out_n, int *out_kv) {
    int n = 0;
    if (max_new <= 0) { *out_n = 0; if (out_kv) *out_kv = kv_len; return 0; }
    out_ids[n++] = first_token;
    int tok = first_token;
    int pos = kv_len;
    if (tok != QASR_TOKEN_ENDOFTEXT && tok != QASR_TOKEN_IM_END && n < max_new) {
        CKR(ensure_kv(c, kv_len + max_new + 1, kv_len));
        CKR(ensure_rope(c, kv_len + max_new + 1));
        CKR(load_token_embedding(c, tok));
        CK(cudaEventRecord(c->ev[0], c->stream));
        bool done = false;
        while (!done && n < max_new) {
            int chunk = max_new - n;
            if (chunk > 16) chunk = 16;
            if (chunk > c->max_steps) chunk = c->max_steps;
            launch_set_state(c->stream, c->d_pos, pos, c->d_done, 0, c->d_step, 0);
            c->launches += 1;
            CKR(enqueue_steps(c, chunk));
            CK(cudaStreamSynchronize(c->stream));
            for (int i = 0; i < chunk; i++) {
                tok = c->h_tokens[i];
                out_ids[n++] = tok;
                pos++;
                if (tok == QASR_TOKEN_ENDOFTEXT || tok == QASR_TOKEN_IM_END) { done = true; break; }
            }
        }
        CK(cudaEventRecord(c->ev[1], c->stream));
        CK(cudaStreamSynchronize(c->stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
        c->last_decode_ms = ms;
        c->decode_ms_total += ms;
        c->decode_steps_total += n - 1;
        c->x_token = tok;
    }
    *out_n = n;
    if (out_kv) *out_kv = pos;
    return 0;
}

int qasr_cuda_generate(qasr_ctx_t *c, int first_token, int kv_len, int max_new, int *out_ids, int *out_n, int *out_kv_len) {
    if (c) c->kv_epoch++;
    if (!c || !out_ids || !out_n) return set_err(QASR_ERR_ARG, "null argument");
    if (!c->loaded) return set_err(QASR_ERR_STATE, "no model loaded");
    if (kv_len < 0) return set_err(QASR_ERR_ARG, "negative kv_len");
    CK(cudaSetDevice(c->device));
    return generate_device(c, first_token, kv_len, max_new, out_ids, out_n, out_kv_len);
}

// Whole offline segment. reference transcribe_segment, qwen_asr.c:649-842
static int transcribe_impl(qasr_ctx_t *c, const float *samples, int n_samples, int max_new, int *out_ids, int *out_n,
                           double *timings_ms, int *out_enc_tokens) {
    if (c) c->kv_epoch++;
    const int *PRE = c ? c->pre_ids.data() : nullptr, *SUF = c ? c->suf_ids.data() : nullptr; // qwen_asr.c:388-396 (+ prompt / language tokens)
    const int n_pre = c ? (int)c->pre_ids.size() : 0, n_suf = c ? (int)c->suf_ids.size() : 0;
    if (!c || !out_ids || !out_n) return set_err(QASR_ERR_ARG, "null argument");
    if (!c->loaded) return set_err(QASR_ERR_STATE, "no model loaded");
    CK(cudaSetDevice(c->device));
    int frames = 0, T = 0, first = 0;
    CK(cudaEventRecord(c->ev[0], c->stream));
    CKR(mel_device(c, samples, n_samples, &frames));
    CK(cudaEventRecord(c->ev[2], c->stream));
    CKR(encode_device(c, c->ws_mel.as<float>(), frames, &T));
    CK(cudaEventRecord(c->ev[3], c->stream));
    CKR(prefill_prompt_device(c, PRE, n_pre, T, SUF, n_suf, 0));
    CK(cudaMemcpyAsync(c->x, c->pending, (size_t)c->H * 4, cudaMemcpyDeviceToDevice, c->stream));
    c->has_pending = false;
    const int kv0 = n_pre + T + n_suf - 1;
    CKR(ensure_kv(c, kv0 + max_new + 2, kv0));
    CKR(ensure_rope(c, kv0 + max_new + 2));
    launch_set_state(c->stream, c->d_pos, kv0, c->d_done, 0, c->d_step, 0);
    c->launches += 1;
    CKR(enqueue_steps(c, 1));
    CK(cudaEventRecord(c->ev[4], c->stream));
    CK(cudaStreamSynchronize(c->stream));
    first = c->h_tokens[0];
    c->x_token = first;
    float t_mel = 0, t_enc = 0, t_pre = 0;
    cudaEventElapsedTime(&t_mel, c->ev[0], c->ev[2]);
    cudaEventElapsedTime(&t_enc, c->ev[2], c->ev[3]);
    cudaEventElapsedTime(&t_pre, c->ev[3], c->ev[4]);
    c->last_decode_ms = 0.0;
    int kv_out = 0;
    CKR(generate_device(c, first, kv0 + 1, max_new, out_ids, out_n, &kv_out));
    if (timings_ms) { timings_ms[0] = t_mel; timings_ms[1] = t_enc; timings_ms[2] = t_pre; timings_ms[3] = c->last_decode_ms; }
    if (out_enc_tokens) *out_enc_tokens = T;
    return 0;
}

int qasr_cuda_transcribe_ids(qasr_ctx_t *c, const float *samples, int n_samples, int max_new, int *out_ids, int *out_n,
                             double *timings_ms, int *out_enc_tokens) {
    if (!samples) return set_err(QASR_ERR_ARG, "null samples");
    return transcribe_impl(c, samples, n_samples, max_new, out_ids, out_n, timings_ms, out_enc_tokens);
}

// ------------------------------------------------------------------ batched segments / utterances
// Independent units (reference: the segments of -S mode, qwen_asr.c:941-1103, or separate files) are decoded
// B at a time by one decode_stream_kernel launch: sequence s rides in MMA columns 2s, 2s+1, so B greedy tokens
// cost the weight traffic of one.  Front end, encoder and prefill still run per unit into that unit's KV cache.
int qasr_cuda_max_batch(const qasr_ctx_t *c) {
    if (!c || !c->loaded) return 0;
    return c->use_stream ? stream_max_seqs(c->H, c->I) : 1;
}

// How qasr_cuda_transcribe_batch would split `count` units: number of groups and the size of the first (largest) one
// = sequences that share one pass over the weights in every decode step.
int qasr_cuda_batch_plan(const qasr_ctx_t *c, int count, int *out_groups, int *out_group_size) {
    if (!c || !c->loaded || count < 0) return set_err(QASR_ERR_ARG, "bad argument");
    const int maxb = qasr_cuda_max_batch(c);
    const char *e = getenv("QASR_BATCH");
    const bool gemm_path = e && !strcmp(e, "gemm") ? count > 1 : (e && !strcmp(e, "stream") ? false : count > maxb);
    if (gemm_path) return batch_plan(count, out_groups, out_group_size);
    const int B = count >= 4 && maxb >= 4 ? 4 : (count >= 2 && maxb >= 2 ? 2 : (count > 0 ? 1 : 0));
    if (out_group_size) *out_group_size = B;
    if (out_groups) { int g = 0; for (int i = 0; i < count; g++) i += count - i >= 4 && maxb >= 4 ? 4 : (count - i >= 2 && maxb >= 2 ? 2 : 1); *out_groups = g; }
    return 0;
}

static int transcribe_group(qasr_ctx_t *c, const float *const *samples, const int *n_samples, int B, const int *max_new, int ids_stride,
                            int *out_ids, int *out_n, double *tm) {
    const int *PRE = c ? c->pre_ids.data() : nullptr, *SUF = c ? c->suf_ids.data() : nullptr; // qwen_asr.c:388-396 (+ prompt / language tokens)
    const int n_pre = c ? (int)c->pre_ids.size() : 0, n_suf = c ? (int)c->suf_ids.size() : 0;
    int kv0[QASR_STREAM_MAX_SEQS] = {}, n[QASR_STREAM_MAX_SEQS] = {};
    bool done[QASR_STREAM_MAX_SEQS] = {};
    int cap_new = 0;
    for (int q = 0; q < B; q++) {
        select_seq(c, q);
        c->kv_fill[q] = 0;
        int frames = 0, T = 0;
        CK(cudaEventRecord(c->ev[0], c->stream));
        CKR(mel_device(c, samples[q], n_samples[q], &frames));
        CK(cudaEventRecord(c->ev[2], c->stream));
        CKR(encode_device(c, c->ws_mel.as<float>(), frames, &T));
        CK(cudaEventRecord(c->ev[3], c->stream));
        CKR(prefill_prompt_device(c, PRE, n_pre, T, SUF, n_suf, 0));
        CK(cudaMemcpyAsync(c->x + (size_t)q * c->H, c->pending, (size_t)c->H * 4, cudaMemcpyDeviceToDevice, c->stream));
        CK(cudaEventRecord(c->ev[4], c->stream));
        c->has_pending = false;
        if (tm) { // per-stage times were asked for: the only reason to wait here (everything is stream-ordered, the workspaces are reused in order)
            CK(cudaEventSynchronize(c->ev[4]));
            float a = 0, b = 0, d = 0;
            cudaEventElapsedTime(&a, c->ev[0], c->ev[2]); cudaEventElapsedTime(&b, c->ev[2], c->ev[3]); cudaEventElapsedTime(&d, c->ev[3], c->ev[4]);
            tm[0] += a; tm[1] += b; tm[2] += d;
        }
        kv0[q] = n_pre + T + n_suf - 1;
        if (max_new[q] > cap_new) cap_new = max_new[q];
    }
    int cap = 0;
    for (int q = 0; q < B; q++) if (kv0[q] > cap) cap = kv0[q];
    cap += cap_new + 34; // finished sequences keep stepping until the whole group is done
    for (int q = 0; q < B; q++) { select_seq(c, q); CKR(ensure_kv(c, cap, kv0[q])); }
    CKR(ensure_rope(c, cap));
    select_seq(c, 0);
    launch_set_state(c->stream, c->d_pos, kv0[0], c->d_done, 0, c->d_step, 0);
    CK(cudaMemcpyAsync(c->d_pos, kv0, sizeof(int) * B, cudaMemcpyHostToDevice, c->stream));
    c->launches += 1;
    CK(cudaEventRecord(c->ev[0], c->stream));
    long long ksteps = 0;
    bool first = true;
    for (;;) {
        int chunk = 0;
        for (int q = 0; q < B; q++) if (!done[q] && max_new[q] - n[q] > chunk) chunk = max_new[q] - n[q];
        if (chunk <= 0) break;
        if (first) chunk = 1; // the step on the pending prompt row (reference qwen_asr.c:769)
        if (chunk > 16) chunk = 16;
        if (!first) { launch_set_state(c->stream, c->d_step, 0, c->d_done, 0, c->d_step, 0); c->launches += 1; }
        CKR(enqueue_steps(c, chunk, B));
        CK(cudaStreamSynchronize(c->stream));
        ksteps += chunk;
        for (int i = 0; i < chunk; i++)
            for (int q = 0; q < B; q++) {
                if (done[q]) continue;
                const int tok = c->h_tokens[i * B + q];
                out_ids[(size_t)q * ids_stride + n[q]++] = tok;
                if (tok == QASR_TOKEN_ENDOFTEXT || tok == QASR_TOKEN_IM_END || n[q] >= max_new[q]) done[q] = true;
            }
        first = false;
    }
    CK(cudaEventRecord(c->ev[1], c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    c->last_decode_ms = ms;
    c->decode_ms_total += ms;
    c->decode_steps_total += ksteps;
    if (tm) tm[3] += ms;
    for (int q = 0; q < B; q++) { out_n[q] = n[q]; c->kv_fill[q] = 0; }
    c->x_token = -1;
    return 0;
}

int qasr_cuda_transcribe_batch(qasr_ctx_t *c, const float *const *samples, const int *n_samples, int count, const int *max_new,
                               int ids_stride, int *out_ids, int *out_n, double *timings_ms) {
    if (c) c->kv_epoch++;
    if (!c || !samples || !n_samples || !max_new || !out_ids || !out_n || count < 0) return set_err(QASR_ERR_ARG, "null argument");
    if (!c->loaded) return set_err(QASR_ERR_STATE, "no model loaded");
    for (int i = 0; i < count; i++)
        if (!samples[i] || max_new[i] < 1 || max_new[i] > ids_stride) return set_err(QASR_ERR_ARG, "unit %d: null samples or max_new outside [1, ids_stride]", i);
    CK(cudaSetDevice(c->device));
    if (timings_ms) timings_ms[0] = timings_ms[1] = timings_ms[2] = timings_ms[3] = 0.0;
    const int maxb = qasr_cuda_max_batch(c);
    // more units than the persistent decode kernel carries per launch (4 / 2): the GEMM-batched throughput path (qasr_batch.cu)
    static int batch_mode = -1; // QASR_BATCH=stream keeps every group on the persistent kernel (A/B runs)
    if (batch_mode < 0) { const char *e = getenv("QASR_BATCH"); batch_mode = e && !strcmp(e, "stream") ? 0 : (e && !strcmp(e, "gemm") ? 2 : 1); }
    if ((batch_mode == 1 && count > maxb) || (batch_mode == 2 && count > 1)) {
        const int rc = batch_transcribe(c, samples, n_samples, count, max_new, ids_stride, out_ids, out_n, timings_ms);
        select_seq(c, 0);
        c->x_token = -1;
        return rc;
    }
    int i = 0;
    while (i < count) {
        int B = count - i >= 4 && maxb >= 4 ? 4 : (count - i >= 2 && maxb >= 2 ? 2 : 1);
        const int rc = transcribe_group(c, samples + i, n_samples + i, B, max_new + i, ids_stride, out_ids + (size_t)i * ids_stride, out_n + i, timings_ms);
        if (rc != 0) { select_seq(c, 0); return rc; }
        i += B;
    }
    select_seq(c, 0);
    return 0;
}

// Prompt of the whole-segment entry points: pre = tokens before the audio rows, suf = tokens after them.  The reference
// builds them in qwen_set_prompt / qwen_set_force_language / transcribe_segment (qwen_asr.c:388-399,685-759):
// pre = [151644, 8948, 198] + system-prompt tokens + [151645, 198, 151644, 872, 198, 151669],
// suf = [151670, 151645, 198, 151644, 77091, 198] (+ "language X" tokens + 151704) (+ past-text tokens + 151704).
int qasr_cuda_set_prompt(qasr_ctx_t *c, const int *pre_ids, int n_pre, const int *suf_ids, int n_suf) {
    if (!c || n_pre < 0 || n_suf < 1 || (n_pre > 0 && !pre_ids) || !suf_ids) return set_err(QASR_ERR_ARG, "bad prompt (the suffix needs at least one token)");
    if (n_pre > 4096 || n_suf > 4096) return set_err(QASR_ERR_ARG, "prompt too long");
    for (int i = 0; i < n_pre; i++) if (pre_ids[i] < 0 || pre_ids[i] >= c->V) return set_err(QASR_ERR_ARG, "token id out of range");
    for (int i = 0; i < n_suf; i++) if (suf_ids[i] < 0 || suf_ids[i] >= c->V) return set_err(QASR_ERR_ARG, "token id out of range");
    c->pre_ids.assign(pre_ids, pre_ids + n_pre);
    c->suf_ids.assign(suf_ids, suf_ids + n_suf);
    return 0;
}

// ------------------------------------------------------------------ streaming session (device-resident)
// The device-visible part of the reference's stream_impl (qwen_asr.c:1273-1900) with every buffer in
// HBM: encoder rows of completed windows are computed once and cached on the device (:1601-1641), the partial tail is
// re-encoded (:1643-1659), at most `max_windows` windows are kept (:1670-1683), the prompt rows are assembled on the
// device, the reusable prefix is found from the window identities instead of a memcmp of float rows (:1811-1823: rows
// are equal exactly when they come from the same cached window at the same position), the delta is prefilled
// (:1825-1829) and up to max_new greedy tokens are decoded (:1880-1888).  Only samples go in and ids come out.
int qasr_cuda_stream_begin(qasr_ctx_t *c, float window_sec, int max_windows) {
    if (!c) return set_err(QASR_ERR_ARG, "null context");
    if (!c->loaded) return set_err(QASR_ERR_STATE, "no model loaded");
    if (!(window_sec > 0.05f) || max_windows < 1 || max_windows > 8) return set_err(QASR_ERR_ARG, "window_sec > 0.05 and 1 <= max_windows <= 8");
    c->st_window = (int)lroundf(window_sec * 16000.0f);
    c->st_max_windows = max_windows;
    for (auto &w : c->st_win) { w.index = -1; w.T = 0; }
    c->st_prev.clear();
    c->st_active = true;
    c->st_fed = false;
    return 0;
}

int qasr_cuda_stream_feed(qasr_ctx_t *c, const float *samples, int n_samples, int max_new, int *out_ids, int *out_n,
                          int *out_reused, int *out_rows) {
    if (!c || !samples || !out_ids || !out_n || n_samples <= 0 || max_new < 1) return set_err(QASR_ERR_ARG, "bad argument");
    if (!c->loaded || !c->st_active) return set_err(QASR_ERR_STATE, "call qasr_cuda_stream_begin first");
    CK(cudaSetDevice(c->device));
    select_seq(c, 0);
    const int H = c->H, W = c->st_window, MW = c->st_max_windows;
    const long long n_full = n_samples / W, first = n_full > MW ? n_full - MW : 0;
    // (1) newly completed windows: mel over the window's own span, encoder, rows into the cache slot of that window
    for (long long w = first; w < n_full; w++) {
        qasr_ctx::StreamWin &slot = c->st_win[w % MW];
        if (slot.index == w) continue;
        int frames = 0, T = 0;
        CKR(mel_device(c, samples + (size_t)w * W, W, &frames));
        CKR(encode_device(c, c->ws_mel.as<float>(), frames, &T));
        if (slot.rows.reserve((size_t)T * H * 4)) return set_err(QASR_ERR_NOMEM, "window cache");
        CK(cudaMemcpyAsync(slot.rows.p, c->ws_encout.p, (size_t)T * H * 4, cudaMemcpyDeviceToDevice, c->stream));
        slot.index = w; slot.T = T;
    }
    // (2) partial tail (re-encoded on every chunk); stays in ws_encout
    int T_tail = 0;
    const int tail_n = n_samples - (int)(n_full * W);
    if (tail_n > 0 && tail_n < 160) { // the reference's mel returns NULL below one frame and the chunk is skipped (qwen_asr.c:1643-1665)
        *out_n = 0;
        if (out_reused) *out_reused = 0;
        if (out_rows) *out_rows = 0;
        return 0;
    }
    if (tail_n >= 160) {
        int frames = 0;
        CKR(mel_device(c, samples + (size_t)n_full * W, tail_n, &frames));
        CKR(encode_device(c, c->ws_mel.as<float>(), frames, &T_tail));
    }
    // (3) prompt layout and reusable prefix
    const int n_pre = (int)c->pre_ids.size(), n_suf = (int)c->suf_ids.size();
    std::vector<long long> cur;
    int total = n_pre + T_tail + n_suf, reused = n_pre;
    for (long long w = first; w < n_full; w++) { cur.push_back(w); total += c->st_win[w % MW].T; }
    if (total - n_pre - n_suf <= 0) { // no encoder rows at all: skipped like the reference (qwen_asr.c:1686-1691)
        *out_n = 0;
        if (out_reused) *out_reused = 0;
        if (out_rows) *out_rows = 0;
        return 0;
    }
    // The prefix is reusable only if the cache still holds the previous chunk's rows: same prompt tokens, and no other entry
    // point wrote sequence 0's KV cache in between (the reference compares the embedding rows themselves, qwen_asr.c:1811-1823).
    if (!c->st_fed || c->kv_epoch != c->st_epoch || c->st_pre != c->pre_ids || c->st_suf != c->suf_ids) reused = 0;
    else
        for (size_t i = 0; i < cur.size() && i < c->st_prev.size() && cur[i] == c->st_prev[i]; i++) reused += c->st_win[cur[i] % MW].T;
    if (reused > total - 1) reused = total - 1;
    // (4) rows [reused, total) assembled in the prefill workspace (the reused prefix is already in the KV cache)
    const int P = total - reused; // includes the last row, which goes through the single-token step
    CKR(reserve_prefill(c, P));
    if (c->ws_ids.reserve((size_t)(n_pre + n_suf + 1) * 4)) return set_err(QASR_ERR_NOMEM, "ids alloc");
    int *d_ids = c->ws_ids.as<int>();
    float *x = c->ws_pre.as<float>();
    CK(cudaMemcpyAsync(d_ids, c->pre_ids.data(), (size_t)n_pre * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_ids + n_pre, c->suf_ids.data(), (size_t)n_suf * 4, cudaMemcpyHostToDevice, c->stream));
    int row = 0; // absolute row of the next segment
    auto place = [&](int seg_rows, auto &&emit) { // emit(dst_row_in_x, first_row_of_segment, count)
        const int lo = reused > row ? reused - row : 0;
        if (lo < seg_rows) emit(row + lo - reused, lo, seg_rows - lo);
        row += seg_rows;
    };
    int rc = 0;
    place(n_pre, [&](int dst, int off, int cnt) { launch_embed_gather(c->stream, c->emb, d_ids + off, cnt, H, x + (size_t)dst * H); c->launches += 1; });
    for (long long w : cur) {
        qasr_ctx::StreamWin &slot = c->st_win[w % MW];
        place(slot.T, [&](int dst, int off, int cnt) {
            if (cudaMemcpyAsync(x + (size_t)dst * H, slot.rows.as<float>() + (size_t)off * H, (size_t)cnt * H * 4, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess) rc = -1;
        });
    }
    place(T_tail, [&](int dst, int off, int cnt) {
        if (cudaMemcpyAsync(x + (size_t)dst * H, c->ws_encout.as<float>() + (size_t)off * H, (size_t)cnt * H * 4, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess) rc = -1;
    });
    place(n_suf, [&](int dst, int off, int cnt) { launch_embed_gather(c->stream, c->emb, d_ids + n_pre + off, cnt, H, x + (size_t)dst * H); c->launches += 1; });
    if (rc != 0) return set_err(QASR_ERR_CUDA, "assembling the prompt rows failed");
    CK(cudaMemcpyAsync(c->x, x + (size_t)(P - 1) * H, (size_t)H * 4, cudaMemcpyDeviceToDevice, c->stream));
    // (5) delta prefill at kv_len = reused (rollback moves no data, qwen_asr.c:1823), first step, greedy loop
    if (P > 1) CKR(prefill_device(c, P - 1, reused));
    const int kv0 = total - 1;
    CKR(ensure_kv(c, kv0 + max_new + 2, kv0));
    CKR(ensure_rope(c, kv0 + max_new + 2));
    launch_set_state(c->stream, c->d_pos, kv0, c->d_done, 0, c->d_step, 0);
    c->launches += 1;
    CKR(enqueue_steps(c, 1));
    CK(cudaStreamSynchronize(c->stream));
    const int first_tok = c->h_tokens[0];
    c->x_token = first_tok;
    c->has_pending = false;
    int kv_out = 0;
    CKR(generate_device(c, first_tok, kv0 + 1, max_new, out_ids, out_n, &kv_out));
    c->st_prev = cur;
    c->st_fed = true;
    c->st_pre = c->pre_ids; c->st_suf = c->suf_ids;
    c->st_epoch = ++c->kv_epoch;
    if (out_reused) *out_reused = reused;
    if (out_rows) *out_rows = total;
    return 0;
}

// Interleaved 16-bit PCM at any sample rate -> f32 mono 16 kHz on the device (reference qwen_parse_wav_buffer,
// qwen_asr_audio.c:81-164).  The result is left staged like qasr_cuda_stage_audio (so qasr_cuda_transcribe_staged can
// consume it without another copy) and optionally copied to the host.
int qasr_cuda_decode_pcm16(qasr_ctx_t *c, const int16_t *pcm, int n_frames, int channels, int sample_rate, float *out, int out_cap, int *out_n) {
    if (!c || !pcm || !out_n || n_frames <= 0 || channels < 1 || sample_rate < 1000) return set_err(QASR_ERR_ARG, "bad argument");
    CK(cudaSetDevice(c->device));
    const int new_n = sample_rate == 16000 ? n_frames : (int)((long long)n_frames * 16000 / sample_rate);
    if (new_n <= 0) return set_err(QASR_ERR_ARG, "audio too short");
    if (out && out_cap < new_n) return set_err(QASR_ERR_ARG, "output buffer holds %d samples, %d needed", out_cap, new_n);
    if (c->ws_pcm.reserve((size_t)n_frames * channels * 2) || c->ws_mono.reserve((size_t)n_frames * 4) || c->ws_samples.reserve((size_t)new_n * 4))
        return set_err(QASR_ERR_NOMEM, "audio buffers");
    CK(cudaMemcpyAsync(c->ws_pcm.p, pcm, (size_t)n_frames * channels * 2, cudaMemcpyHostToDevice, c->stream));
    float *mono = sample_rate == 16000 ? c->ws_samples.as<float>() : c->ws_mono.as<float>();
    launch_pcm16_to_mono(c->stream, c->ws_pcm.as<int16_t>(), n_frames, channels, mono);
    if (sample_rate != 16000) launch_resample_sinc(c->stream, mono, n_frames, sample_rate, c->ws_samples.as<float>(), new_n);
    c->launches += sample_rate != 16000 ? 2 : 1;
    CK(cudaGetLastError());
    if (out) CK(cudaMemcpyAsync(out, c->ws_samples.p, (size_t)new_n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->staged_samples = new_n;
    *out_n = new_n;
    return 0;
}

int qasr_cuda_stage_audio(qasr_ctx_t *c, const float *samples, int n_samples) {
    if (!c || !samples || n_samples <= 0) return set_err(QASR_ERR_ARG, "bad argument");
    CK(cudaSetDevice(c->device));
    if (c->ws_samples.reserve((size_t)n_samples * 4)) return set_err(QASR_ERR_NOMEM, "sample buffer");
    CK(cudaMemcpyAsync(c->ws_samples.p, samples, (size_t)n_samples * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->staged_samples = n_samples;
    return 0;
}

int qasr_cuda_transcribe_staged(qasr_ctx_t *c, int max_new, int *out_ids, int *out_n, double *timings_ms, int *out_enc_tokens) {
    if (!c || c->staged_samples <= 0) return set_err(QASR_ERR_STATE, "no staged audio: call qasr_cuda_stage_audio first");
    return transcribe_impl(c, nullptr, c->staged_samples, max_new, out_ids, out_n, timings_ms, out_enc_tokens);
}

// CUDA-event stopwatch on the context's stream (the stream every kernel of this library runs on)
int qasr_cuda_timer_start(qasr_ctx_t *c) {
    if (!c) return set_err(QASR_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    if (!c->tev[0]) { CK(cudaEventCreate(&c->tev[0])); CK(cudaEventCreate(&c->tev[1])); }
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventRecord(c->tev[0], c->stream));
    return 0;
}
int qasr_cuda_timer_stop(qasr_ctx_t *c, double *out_ms) {
    if (!c || !out_ms || !c->tev[0]) return set_err(QASR_ERR_ARG, "timer not started");
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->tev[1], c->stream));
    CK(cudaEventSynchronize(c->tev[1]));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->tev[0], c->tev[1]));
    *out_ms = ms;
    return 0;
}
int qasr_cuda_decode_stats(qasr_ctx_t *c, long long *steps, double *ms, int reset) {
    if (!c) return set_err(QASR_ERR_ARG, "null context");
    if (steps) *steps = c->decode_steps_total;
    if (ms) *ms = c->decode_ms_total;
    if (reset) { c->decode_steps_total = 0; c->decode_ms_total = 0.0; }
    return 0;
}

// ------------------------------------------------------------------ test hooks
int qasr_cuda_read_kv(qasr_ctx_t *c, int layer, int len, float *k_out, float *v_out) {
    if (!c || !k_out || !v_out) return set_err(QASR_ERR_ARG, "null argument");
    if (!c->loaded || layer < 0 || layer >= c->dec_layers || len < 0 || len > c->kv_max) return set_err(QASR_ERR_ARG, "bad layer/len");
    CK(cudaSetDevice(c->device));
    const size_t kvd = (size_t)c->kv_heads * c->hd;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(k_out, c->kv_k + (size_t)layer * c->kv_max * kvd, (size_t)len * kvd * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(v_out, c->kv_v + (size_t)layer * c->kv_max * kvd, (size_t)len * kvd * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int qasr_cuda_embed_token(qasr_ctx_t *c, int token_id, float *out) {
    if (!c || !out) return set_err(QASR_ERR_ARG, "null argument");
    if (!c->loaded || token_id < 0 || token_id >= c->V) return set_err(QASR_ERR_ARG, "bad token id");
    CK(cudaSetDevice(c->device));
    std::vector<uint16_t> h(c->H);
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(h.data(), c->emb + (size_t)token_id * c->H, (size_t)c->H * 2, cudaMemcpyDeviceToHost));
    for (int i = 0; i < c->H; i++) out[i] = bf16_to_f32(h[i]);
    return 0;
}

// debug: copy the megakernel phase stamps (2 x 4096 clock64 values) to the host
extern "C" int qasr_debug_mega_prof(qasr_ctx_t *c, long long *out) {
    if (!c || !c->mega_prof) return -1;
    cudaStreamSynchronize(c->stream);
    return cudaMemcpy(out, c->mega_prof, 3 * 4096 * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}

// ------------------------------------------------------------------ accessors for qasr_ops.cu
cudaStream_t qasr_internal_stream(qasr_ctx_t *c) { return c->stream; }
int qasr_internal_device(qasr_ctx_t *c) { return c->device; }
int qasr_internal_nsplit(qasr_ctx_t *c) { return c->nsplit; }
void qasr_internal_count(qasr_ctx_t *c, int n) { c->launches += n; }
int qasr_internal_err(int code, const char *msg) { return set_err(code, "%s", msg); }

// qasr_ctx.h - the context behind the opaque qasr_ctx_t handle and the helpers shared by the translation units that
// implement the C ABI (qasr_api.cu: lifecycle, single-sequence entry points; qasr_batch.cu: the batched throughput path).
#pragma once
#include "../../include/qasr_cuda.h"
#include "qasr_internal.h"

#include <string.h>

#include <vector>

// ------------------------------------------------------------------ errors
int set_err(int code, const char *fmt, ...); // records the thread's error text (qasr_cuda_last_error) and returns `code`

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return set_err(QASR_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define CKR(expr)              \
    do {                       \
        int r__ = (expr);      \
        if (r__ != 0) return r__; \
    } while (0)

// ------------------------------------------------------------------ context
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool grew = false;
    int reserve(size_t bytes) { // allocate before freeing: a failed grow leaves the old buffer (and the graphs that point into it) intact
        if (bytes <= cap) return 0;
        const size_t want = bytes + bytes / 4;
        void *np = nullptr;
        if (cudaMalloc(&np, want) != cudaSuccess) { cudaGetLastError(); return -1; }
        if (p) cudaFree(p);
        p = np;
        cap = want;
        grew = true;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct EncLayerW {
    bf16_t *wqkv, *wo, *fc1, *fc2;
    float *bqkv, *bo, *fc1b, *fc2b, *ln1w, *ln1b, *ln2w, *ln2b;
};
struct DecLayerW {
    bf16_t *wqkv, *wo, *wgu, *wdown;
    float *qn, *kn, *in_norm, *post_norm;
};

struct qasr_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool loaded = false;
    int nsplit = 2;
    // config (reference qwen_config_t)
    int d = 0, enc_layers = 0, enc_heads = 0, F = 0, H = 0, dec_layers = 0, heads = 16, kv_heads = 8, hd = 128, I = 0,
        V = 151936;
    // weights
    std::vector<void *> owned; // every cudaMalloc'd weight block
    float *c1w = nullptr, *c1b = nullptr, *c2b = nullptr, *c3b = nullptr;
    bf16_t *c2w = nullptr, *c3w = nullptr, *conv_out = nullptr, *p1w = nullptr, *p2w = nullptr;
    float *lnpw = nullptr, *lnpb = nullptr, *p1b = nullptr, *p2b = nullptr;
    EncLayerW enc[32];
    DecLayerW dec[48];
    bf16_t *emb = nullptr;
    float *final_norm = nullptr;
    size_t weight_bytes = 0;
    // constant tables
    float *mel_cos = nullptr, *mel_sin = nullptr, *mel_win = nullptr, *mel_fb = nullptr, *pe = nullptr;
    float *rope_cos = nullptr, *rope_sin = nullptr;
    int rope_cap = 0;
    // KV cache f32 [layers][kv_max][kv_heads*hd]
    float *kv_k = nullptr, *kv_v = nullptr; // cache of the CURRENT sequence (kv_ks[seq]); seq 0 = the single-sequence API
    float *kv_ks[QASR_STREAM_MAX_SEQS] = {}, *kv_vs[QASR_STREAM_MAX_SEQS] = {}; // batched decode: one cache per sequence, same capacity
    int kv_fill[QASR_STREAM_MAX_SEQS] = {};  // valid rows per sequence (what a growth has to preserve)
    int seq = 0;
    int kv_max = 0;
    // decode-step state
    float *x = nullptr, *qkv = nullptr, *attn = nullptr, *act = nullptr, *attn_part = nullptr, *logits = nullptr,
          *pending = nullptr, *part_val = nullptr;
    int *part_idx = nullptr, *d_pos = nullptr, *d_done = nullptr, *d_step = nullptr, *d_tokens = nullptr;
    unsigned *counters = nullptr;
    int *h_tokens = nullptr, *dh_tokens = nullptr; // mapped pinned ring
    int max_steps = 64;
    int n_parts = 0;
    bool has_pending = false;
    int x_token = -1; // token whose embedding currently sits in x (or -1)
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraph_t graph = nullptr;
    int graph_nodes = 0;
    bool use_graph = true;
    bool use_stream = true; // persistent cooperative decode kernel of qasr_stream.cu (default); QASR_DECODE=graph selects the per-phase kernels
    uint8_t *sk_image = nullptr;           // decode weight image (pre-tiled, per-warp streams)
    uint8_t *sk_image_r = nullptr;         // the same units round-major (single-sequence producer / consumer kernel)
    unsigned long long *sk_cta_off = nullptr;
    unsigned long long *ll_qkv = nullptr, *ll_att = nullptr, *ll_xwo = nullptr, *ll_act = nullptr, *ll_xdn = nullptr, *ll_head = nullptr;
    unsigned sk_tag = 1;                   // next free exchange tag
    float *dbg_logits = nullptr, *dbg_hidden = nullptr; // set around one launch by the logits entry points
    float *hidden_buf = nullptr;           // [QASR_STREAM_MAX_SEQS][H] landing buffer of dbg_hidden
    // prompt around the audio rows used by the whole-segment entry points (reference qwen_asr.c:388-399,685-759): default = no system text, no forced language
    std::vector<int> pre_ids = {151644, 8948, 198, 151645, 198, 151644, 872, 198, 151669};
    std::vector<int> suf_ids = {151670, 151645, 198, 151644, 77091, 198};
    // streaming session (qasr_cuda_stream_*): encoder rows of the completed windows stay in HBM
    struct StreamWin { long long index = -1; int T = 0; DevBuf rows; };
    StreamWin st_win[8];
    int st_window = 0, st_max_windows = 0;      // samples per window, windows kept
    std::vector<long long> st_prev;             // window indices of the previous chunk's prompt, in order
    bool st_active = false, st_fed = false;     // session open / at least one chunk fed (the KV cache holds its prompt)
    std::vector<int> st_pre, st_suf;            // prompt tokens the previous chunk was prefilled with
    long long kv_epoch = 0, st_epoch = -1;      // bumped by every entry point that writes sequence 0's KV cache; value after the previous chunk
    long long *mega_prof = nullptr;
    struct GraphEntry { long long key[4]; cudaGraphExec_t exec; long long n_launch; };
    std::vector<GraphEntry> graph_cache; // captured encoder / prefill launch sequences, keyed by shape
    long long ws_gen = 0;                // bumped whenever a workspace the graphs point into is reallocated
    // scratch
    DevBuf ws_samples, ws_meltmp, ws_mel, ws_enc, ws_encout, ws_pre, ws_ids, ws_geom, ws_pcm, ws_mono;
    int *d_gmax = nullptr;
    int mel_frames = 0, enc_T = 0;
    int geom_frames = 0; // frame count whose chunk / window tables sit in ws_geom
    cudaEvent_t ev[5] = {};
    cudaEvent_t tev[2] = {};
    double last_decode_ms = 0.0;
    double decode_ms_total = 0.0;
    long long decode_steps_total = 0;
    int staged_samples = 0;
    long long launches = 0;
    struct BatchState *batch = nullptr;  // buffers of the batched throughput path (qasr_batch.cu), created on first use
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline void note_growth(qasr_ctx_t *c, DevBuf &b) { if (b.grew) { c->ws_gen++; b.grew = false; } }


// Launch sequences of the encoder / prefill are captured once per shape into a CUDA graph and replayed:
// at these sizes the ~230 + ~170 launches of one utterance are CPU-launch-bound otherwise (each
// tensor-core GEMM launch also encodes two TMA descriptors on the host).
template <class F>
static inline int run_cached_graph(qasr_ctx_t *c, long long k0, long long k1, long long k2, F &&enqueue) {
    if (!c->use_graph) return enqueue();
    const long long key[4] = {k0, k1, k2, c->ws_gen};
    for (auto &ge : c->graph_cache)
        if (!memcmp(ge.key, key, sizeof key)) { CK(cudaGraphLaunch(ge.exec, c->stream)); c->launches += ge.n_launch; return 0; }
    const long long launches_before = c->launches;
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue();
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
    if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return set_err(QASR_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
    cudaGraphExec_t exec = nullptr;
    CK(cudaGraphInstantiate(&exec, graph, 0));
    cudaGraphDestroy(graph);
    if (c->graph_cache.size() >= 16) { cudaGraphExecDestroy(c->graph_cache.front().exec); c->graph_cache.erase(c->graph_cache.begin()); }
    qasr_ctx::GraphEntry ge;
    memcpy(ge.key, key, sizeof key);
    ge.exec = exec;
    ge.n_launch = c->launches - launches_before;
    c->graph_cache.push_back(ge);
    CK(cudaGraphLaunch(exec, c->stream));
    return 0;
}


// ---- helpers defined in qasr_api.cu, shared with qasr_batch.cu
int encode_units_device(qasr_ctx_t *c, const float *d_mel, int mel_stride, const int *unit_frames, int n_units, float *out, int *T_out);
int ensure_rope(qasr_ctx_t *c, int need_pos);
int gemm(qasr_ctx_t *c, const bf16_t *a_hi, const bf16_t *a_lo, int M, int K, const bf16_t *W, int N, int mode, float *of, bf16_t *ohi,
         bf16_t *olo, const float *bias, int ldo, const GemmEpilogue *norm = nullptr); // norm: nx_* / in_* fields of a fused RMSNorm (qasr_internal.h)
// batched path (qasr_batch.cu): independent units through batched front end / encoder / prefill / decode
int batch_transcribe(qasr_ctx_t *c, const float *const *samples, const int *n_samples, int count, const int *max_new, int ids_stride,
                     int *out_ids, int *out_n, double *timings_ms);
void batch_release(qasr_ctx_t *c);
int batch_plan(int count, int *out_groups, int *out_group_size);

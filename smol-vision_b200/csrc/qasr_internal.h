// qasr_internal.h - internal launcher prototypes shared by the .cu files of libqasr_cuda.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

typedef uint16_t bf16_t; // raw bf16 bits on the host side / in signatures

#ifdef __CUDACC__
#include <stdlib.h>
#include <utility>
// Programmatic dependent launch for the encoder / prefill launch chains (~440 small dependent kernels per utterance):
// every kernel of a chain starts with pdl_trigger() - the NEXT grid may be scheduled as soon as all CTAs of this one
// are resident, so its launch, CTA placement and (for the GEMMs) barrier / TMEM set-up and first weight tiles overlap
// the grids before it - and pdl_wait(): nothing it reads or writes is touched before the previous grid has completed
// and flushed.  Every kernel of a chain must wait, or completion would stop being transitive.  No deadlock: a grid
// triggers only once all its CTAs are resident, so at any time only the newest launched grid can have CTAs waiting for
// an SM, and the oldest unfinished grid always runs to completion.  QASR_PDL=0 = ordinary stream order.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("QASR_PDL"); on = !(e && e[0] == '0'); }
    return on != 0;
}
// cluster_y > 1: thread-block clusters of (1, cluster_y, 1) CTAs (gridDim.y must be a multiple)
// Chains are captured per shape; the ones whose GEMMs take the large-tile kernel (more than 256 rows) measured slower
// with early scheduling (30 s utterance: encoder 3.96 -> 4.23 ms) while the skinny-GEMM chains gain (3.6 s utterance:
// encoder + prefill 4.67 -> 4.30 ms), so the caller scopes it by row count.
inline int &pdl_scope() { static thread_local int on = 1; return on; }
struct PdlScope {
    int prev;
    explicit PdlScope(bool on) : prev(pdl_scope()) { pdl_scope() = on; }
    ~PdlScope() { pdl_scope() = prev; }
};
template <class... KA, class... A>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_y, A &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[2];
    unsigned n = 0;
    if (pdl_enabled() && pdl_scope()) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        n++;
    }
    if (cluster_y > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = 1; at[n].val.clusterDim.y = (unsigned)cluster_y; at[n].val.clusterDim.z = 1;
        n++;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...);
}
template <class... KA, class... A>
inline cudaError_t launch_pdl(void (*kernel)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A &&...args) {
    return launch_pdl_cluster(kernel, grid, block, smem, s, 1, std::forward<A>(args)...);
}
#endif

enum { QASR_EPI_STORE = 0, QASR_EPI_RESIDUAL = 1, QASR_EPI_SWIGLU = 2 };
// tensor-core GEMM epilogues
enum { QASR_GEMM_F32 = 0, QASR_GEMM_RESIDUAL = 1, QASR_GEMM_GELU_SPLIT = 2, QASR_GEMM_SWIGLU_SPLIT = 3 };

#define QASR_ATTN_SPLITS 16
#define QASR_ATTN_PART_STRIDE 132 /* 128 acc + m + l + pad */
#define QASR_ARGMAX_ROWS_PER_CTA 32

struct GemmEpilogue {
    int mode;          // QASR_GEMM_*
    float *out_f32;    // F32 / RESIDUAL target [M, ldo]
    bf16_t *out_hi;    // *_SPLIT targets [M, ldo]
    bf16_t *out_lo;    // may be NULL (nsplit == 1)
    const float *bias; // [N] or NULL
    int ldo;           // leading dimension of the output (N, or N/2 for SWIGLU)
    // RMSNorm fused across two GEMMs (skinny kernel only, gemm_tc_can_fuse_norm): the scalar 1/sqrt(mean x^2 + eps) of a row commutes
    // with the product, so the PRODUCER of the residual stream (mode RESIDUAL) also writes the next GEMM's operand planes hi/lo of
    // x * gamma and the sum of x^2 over its 128 columns of every row; the CONSUMER multiplies row m of its accumulator by
    // rsqrt(sum_t in_ssq[m][t] / K + eps) before anything else.  Reference: qwen_rms_norm, qwen_asr_kernels.c:801-860.
    const float *nx_gamma = nullptr; // producer: [N] weight of the norm that follows
    bf16_t *nx_hi = nullptr, *nx_lo = nullptr; // producer: planes [M, N] (lo may be NULL)
    float *nx_ssq = nullptr;         // producer: [M][N / 128] partial sums of squares
    const float *in_ssq = nullptr;   // consumer: [M][in_tiles] partial sums written by the producer of its input
    int in_tiles = 0;
    float in_eps = 0.0f;
    // q/k per-head RMSNorm + NeoX RoPE + KV-cache store fused into the QKV GEMM (mode F32, N = 4096 = 16 q | 8 k | 8 v heads of 128,
    // skinny split-K path: a warp of the reduction holds one head of one row).  Reference qwen_asr_decoder.c:510-524.
    float *qk_q = nullptr;           // [M][2048] roped queries (replaces out_f32, which is not written)
    float *qk_kc = nullptr, *qk_vc = nullptr; // this layer's cache rows [pos][1024]
    const float *qk_qn = nullptr, *qk_kn = nullptr, *qk_cos = nullptr, *qk_sin = nullptr; // norm weights [128], RoPE tables [pos][64]
    int qk_pos0 = 0;                 // row m sits at position qk_pos0 + m
    float qk_eps = 0.0f;
};

// ---- decode-step kernels (qasr_decode.cu)
void launch_gemv_bf16(cudaStream_t s, const bf16_t *W, const float *x, const float *gamma, float eps, float *out,
                      const float *res, const float *bias, int N, int K, int epi, const int *d_done);
void launch_argmax_gemv(cudaStream_t s, const bf16_t *E, const float *x, const float *gamma, float eps, int V, int K,
                        float *part_val, int *part_idx, const int *d_done);
int argmax_num_parts(int V);
void launch_argmax_finalize(cudaStream_t s, const float *part_val, const int *part_idx, int n_parts, const bf16_t *E,
                            int H, float *x_next, int *d_tokens, int *d_step, int *d_pos, int *d_done,
                            volatile int *h_tokens_mapped, int max_steps);
void launch_attn_decode(cudaStream_t s, const float *qkv, const float *qn, const float *kn, const float *rope_cos,
                        const float *rope_sin, float *kc, float *vc, const int *d_pos, float *part,
                        unsigned *counters, float *out, float eps);
void launch_set_state(cudaStream_t s, int *d_pos, int pos, int *d_done, int done, int *d_step, int step);
void launch_embed_gather(cudaStream_t s, const bf16_t *E, const int *d_ids, int n, int H, float *out);

// ---- streaming decode kernel (qasr_stream.cu): pre-tiled weight image + flag-in-data exchanges
struct StreamParams {
    const uint8_t *image;                  // decode weight image (units of 16x64 bf16 in A-fragment order, per-warp streams)
    const uint8_t *image_r;                // the same units round-major (qasr_stream_r.cu), or NULL
    const unsigned long long *cta_off;     // [grid+1] byte offsets of each CTA's 16 streams
    int n_layers, H, I, V, n_steps;
    float eps;
    const bf16_t *emb;                     // row-major tied embedding (next-input gather)
    const float *in_norm[28], *post_norm[28], *qn[28], *kn[28];
    const float *final_norm;
    int nseq;                              // sequences decoded together (1, 2 or 4): sequence s = MMA columns 2s, 2s+1
    float *x_io;                           // [nseq][H] input rows in / last gathered rows out
    float *kv_k[4], *kv_v[4];              // per-sequence KV caches
    size_t kv_layer_stride;
    const float *rope_cos, *rope_sin;      // [pos][64]
    unsigned long long *ll_qkv, *ll_att, *ll_xwo, *ll_act, *ll_xdn, *ll_head; // {f32, tag} exchange buffers, [nseq][...]
    unsigned tag_base;                     // tags of this launch: tag_base + step*(L+1) + layer + 1
    int *d_pos, *d_step, *d_tokens;        // d_pos[nseq]; d_tokens[step][nseq]
    volatile int *h_tokens;                // mapped pinned ring [step][nseq] (may be NULL)
    long long *prof;                       // optional clock64 stamps [2][prof_cap] (CTA 0, last CTA) or NULL
    int prof_cap;
    int debug;
    int trace_cta;                         // CTA whose warp 0 writes the per-unit trace (debug bit 6)
    int l2_issue;                          // units prefetched per poll iteration
    int l2_ahead_units;                    // second-level prefetch distance into L2, in 2 KB units per warp (0 = off)
    int sr_chunk;                          // producer of qasr_stream_r.cu: bytes per bulk copy
    int sr_chunk_head;                     // ... for the rounds of the lm_head phase (pure streaming: larger copies, fewer issues)
    float *dbg_logits;                     // test hook (NULL in production): [nseq][V] logits of the LAST step of the launch
    float *dbg_hidden;                     // test hook (NULL in production): [nseq][H] post-final-norm hidden state
};
int stream_init(void);
int stream_grid(void);
int stream_max_seqs(int H, int I);
size_t stream_image_layout(int L, int H, int I, int V, unsigned long long *cta_off_host /* [grid+1] */);
int stream_build_image(cudaStream_t s, int L, int H, int I, int V, const bf16_t *const *layer_mats /* [L*4] */, const bf16_t *emb,
                       const unsigned long long *d_cta_off, uint8_t *image, int round_major);
int launch_decode_stream(cudaStream_t s, const StreamParams &p);
bool stream_use_rounds(int H);
int launch_decode_rounds(cudaStream_t s, const StreamParams &p, int grid, char *err, size_t errlen); // qasr_stream_r.cu
const char *stream_error(void);
#define QASR_STREAM_MAX_SEQS 4
#define QASR_STREAM_ATT_WORDS (16 * 4 * 132) /* per sequence: 16 heads x SK_ATT_MAXS splits x (128 acc + m + l + 2 pad) */

// ---- row-wise / prefill / encoder kernels (qasr_rows.cu)
void launch_rmsnorm(cudaStream_t s, const float *x, const float *gamma, float eps, int M, int H, float *out_f32,
                    bf16_t *out_hi, bf16_t *out_lo);
void launch_layernorm(cudaStream_t s, const float *x, const float *w, const float *b, float eps, int M, int H,
                      float *out_f32, bf16_t *out_hi, bf16_t *out_lo);
void launch_rmsnorm_per_head(cudaStream_t s, float *x, const float *w, int seq, int n_heads, int head_dim, float eps);
void launch_qk_norm_rope_store(cudaStream_t s, const float *qkv, const float *qn, const float *kn, const float *rope_cos,
                               const float *rope_sin, int start_pos, int P, float eps, float *q_out, float *kc, float *vc);
void launch_attn_prefill(cudaStream_t s, const float *q, const float *kc, const float *vc, int q_offset, int P, int seq_k,
                         int n_heads, int n_kv_heads, float scale, float *out_f32, bf16_t *out_hi, bf16_t *out_lo);
void launch_attn_windowed(cudaStream_t s, const float *q, const float *k, const float *v, int ld, int n_heads,
                          const int *d_window_starts, int n_windows, int max_window, float scale, int out_ld,
                          float *out_f32, bf16_t *out_hi, bf16_t *out_lo);
// batched path (qasr_batch.cu): per-unit KV caches inside one pool, unit u's block of a layer at pool + u * unit_stride,
// head-major inside ([kv head][cap][128], head_stride = cap * 128)
void launch_qk_norm_rope_store_rows(cudaStream_t s, const float *qkv, const float *qn, const float *kn, const float *rope_cos,
                                    const float *rope_sin, const int *d_row_unit, const int *d_row_pos, int R, float eps, float *q_out,
                                    float *kpool, float *vpool, size_t unit_stride, size_t head_stride);
void launch_attn_prefill_batch(cudaStream_t s, const float *q, const float *kpool, const float *vpool, size_t unit_stride, size_t head_stride,
                               const int *d_row0, const int *d_P, int n_units, int max_P, int n_heads, int n_kv_heads, float scale, bf16_t *out_hi, bf16_t *out_lo);
void launch_attn_decode_batch(cudaStream_t s, const float *qkv, const float *qn, const float *kn, const float *rope_cos, const float *rope_sin,
                              float *kpool, float *vpool, size_t unit_stride, size_t head_stride, const int *d_pos, int B, float eps, float scale, bf16_t *ohi,
                              bf16_t *olo);
void launch_argmax_next(cudaStream_t s, const float *logits, int V, const bf16_t *E, int H, float *x_next, int *d_pos, int *d_step,
                        int *d_tokens, volatile int *h_tokens, int B, int max_steps);
void launch_assemble_prompts(cudaStream_t s, const bf16_t *E, int H, const int *d_pre, int n_pre, const int *d_suf, int n_suf, const float *enc,
                             const int *d_enc0, const int *d_T, const int *d_row0, int n_units, int max_total, float *X, float *x_first);
void launch_split_f32(cudaStream_t s, const float *x, size_t n, bf16_t *hi, bf16_t *lo);
void launch_add_rows(cudaStream_t s, float *x, const float *table, const int *d_row_idx, int M, int d);
void launch_eltwise(cudaStream_t s, int op, float *a, const float *b, float scalar, size_t n);
void launch_swiglu(cudaStream_t s, float *out, const float *gate_up, int seq, int inter);
void launch_softmax(cudaStream_t s, float *x, int rows, int cols);
void launch_rope_apply(cudaStream_t s, float *x, const float *c, const float *sn, int seq, int n_heads, int head_dim);
void launch_gemm_f32(cudaStream_t s, const float *A, const float *W, const float *bias, float *C, int M, int N, int K);
void launch_im2col_f32(cudaStream_t s, const float *in, float *cols, int c_in, int h_in, int w_in, int kh, int kw,
                       int stride, int padding, int h_out, int w_out);
void launch_transpose_bias(cudaStream_t s, const float *in, const float *bias, float *out, int S, int C);

// ---- conv stem (qasr_conv.cu)
struct ConvGeom { // per-chunk geometry tables live on the device
    int n_chunks;
    const int *d_w0;   // [n_chunks] mel frames in the chunk
    const int *d_mel0; // [n_chunks] first mel frame of the chunk
    const int *d_off1; // [n_chunks+1] position prefix for stage-1 output (w1*64 each)
    const int *d_off2; // stage-2 output (w2*32)
    const int *d_off3; // stage-3 output (w3*16)
    int total1, total2, total3;
};
void launch_conv1(cudaStream_t s, const float *mel, int frames, const float *w, const float *b, const ConvGeom &g,
                  bf16_t *out_hi, bf16_t *out_lo);
void launch_im2col_stage(cudaStream_t s, const bf16_t *src, bf16_t *dst, const ConvGeom &g, int stage /*2 or 3*/);

// ---- mel (qasr_mel.cu)
// mel_out is [128][out_stride]; columns [out_off, out_off + frames) are written (single unit: out_stride = frames, out_off = 0)
void launch_mel(cudaStream_t s, const float *samples, int n, int frames, const float *d_cos, const float *d_sin,
                const float *d_win, const float *d_fb, float *mel_tmp, int *d_gmax, float *mel_out, int out_stride, int out_off);

void launch_pcm16_to_mono(cudaStream_t s, const int16_t *pcm, int n, int channels, float *out);
void launch_resample_sinc(cudaStream_t s, const float *in, int n, int rate, float *out, int new_n);

// ---- tcgen05 GEMM (qasr_gemm_tc.cu)
// C[M,N] = A[M,K] * W[N,K]^T with A given as bf16 hi (+ optional lo) planes, W bf16, f32 accumulate in TMEM.
int gemm_tc_init(void); // resolves cuTensorMapEncodeTiled, sets smem attributes; 0 on success
int gemm_tc_prepare(void); // per-device split-K scratch (call once per device, outside stream capture)
int launch_gemm_tc(cudaStream_t s, const bf16_t *A_hi, const bf16_t *A_lo, int M, int K, const bf16_t *W, int N,
                   const GemmEpilogue &epi);
bool gemm_tc_can_fuse_norm(int M, int K, int N); // a RESIDUAL GEMM of this shape can carry nx_* (skinny kernel, split-K reduction path)
bool gemm_tc_can_scale_rows(int M);              // a GEMM with M rows can carry in_ssq (skinny kernel)
bool gemm_tc_can_fuse_qk(int M, int K);          // the QKV GEMM (N = 4096) of this shape can carry qk_*
const char *gemm_tc_error(void);
// conv stem stage 2 / 3 as an implicit GEMM: the 3 x 3 patches are gathered into the operand stage by the kernel itself
int launch_conv_gemm_tc(cudaStream_t s, const bf16_t *src_hi, const bf16_t *src_lo, const ConvGeom &g, int stage, const bf16_t *W,
                        const GemmEpilogue &epi);
// encode a 2-D bf16 [rows, K] row-major tensor map with the given box and 128B swizzle into *out (sizeof(CUtensorMap) = 128 bytes, host)
int tc_encode_map(void *out_map64, const bf16_t *ptr, int rows, int K, int box_cols, int box_rows);

// qasr_rows.cu - row-wise kernels for the encoder and the decoder prefill: norms that emit the
// bf16 hi/lo operand planes of the tensor-core GEMMs, q/k-norm + RoPE + KV store, causal and
// windowed attention, and the small f32 operators behind the level-2 test seam.
#include "qasr_common.cuh"
#include "qasr_internal.h"

// Store one value into whichever outputs are requested.
__device__ __forceinline__ void store_out(float v, size_t idx, float *of, bf16_t *ohi, bf16_t *olo) {
    if (of) of[idx] = v;
    if (ohi) {
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        ohi[idx] = __bfloat16_as_ushort(hi);
        if (olo) olo[idx] = __bfloat16_as_ushort(lo);
    }
}

// Row norms: one 256-thread CTA per row, the row lives in registers (<= 16 values per thread, H <= 4096),
// one block reduction per statistic.  Emits f32 and/or the bf16 hi/lo operand planes.
#define NORM_THREADS 256
#define NORM_MAX_PER 16
__device__ __forceinline__ float norm_block_sum(float v, float *red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < NORM_THREADS / 32; w++) t += red[w];
    return t;
}

// ------------------------------------------------------------------ RMSNorm (rows)
// out = x * rsqrt(mean(x^2)+eps) * w, reference qwen_asr_kernels.c:801-860.
__global__ void __launch_bounds__(NORM_THREADS)
rmsnorm_rows_kernel(const float *__restrict__ x, const float *__restrict__ gamma, float eps, int H, float *of, bf16_t *ohi,
                    bf16_t *olo) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[NORM_THREADS / 32];
    const size_t base = (size_t)blockIdx.x * H;
    float v[NORM_MAX_PER];
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_PER; i++) {
        const int e = threadIdx.x + i * NORM_THREADS;
        v[i] = e < H ? x[base + e] : 0.0f;
        ss = fmaf(v[i], v[i], ss);
    }
    const float inv = 1.0f / sqrtf(norm_block_sum(ss, red) / (float)H + eps);
#pragma unroll
    for (int i = 0; i < NORM_MAX_PER; i++) {
        const int e = threadIdx.x + i * NORM_THREADS;
        if (e < H) store_out(v[i] * inv * gamma[e], base + e, of, ohi, olo);
    }
}
void launch_rmsnorm(cudaStream_t s, const float *x, const float *gamma, float eps, int M, int H, float *out_f32,
                    bf16_t *out_hi, bf16_t *out_lo) {
    if (M > 0) launch_pdl(rmsnorm_rows_kernel, M, NORM_THREADS, 0, s, x, gamma, eps, H, out_f32, out_hi, out_lo);
}

// ------------------------------------------------------------------ LayerNorm (rows)
// (x-mean)*rsqrt(var+eps)*w+b, biased variance, reference qwen_asr_kernels.c:691-799.
__global__ void __launch_bounds__(NORM_THREADS)
layernorm_rows_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b, float eps, int H,
                      float *of, bf16_t *ohi, bf16_t *olo) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[NORM_THREADS / 32];
    const size_t base = (size_t)blockIdx.x * H;
    float v[NORM_MAX_PER];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_PER; i++) {
        const int e = threadIdx.x + i * NORM_THREADS;
        v[i] = e < H ? x[base + e] : 0.0f;
        sum += v[i];
    }
    const float mean = norm_block_sum(sum, red) / (float)H;
    float var = 0.0f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_PER; i++) {
        const int e = threadIdx.x + i * NORM_THREADS;
        const float d = e < H ? v[i] - mean : 0.0f;
        var = fmaf(d, d, var);
    }
    const float inv = 1.0f / sqrtf(norm_block_sum(var, red) / (float)H + eps);
#pragma unroll
    for (int i = 0; i < NORM_MAX_PER; i++) {
        const int e = threadIdx.x + i * NORM_THREADS;
        if (e < H) store_out((v[i] - mean) * inv * w[e] + b[e], base + e, of, ohi, olo);
    }
}
void launch_layernorm(cudaStream_t s, const float *x, const float *w, const float *b, float eps, int M, int H,
                      float *out_f32, bf16_t *out_hi, bf16_t *out_lo) {
    if (M > 0) launch_pdl(layernorm_rows_kernel, M, NORM_THREADS, 0, s, x, w, b, eps, H, out_f32, out_hi, out_lo);
}

// ------------------------------------------------------------------ per-head RMSNorm (in place)
// reference qwen_asr_kernels.c:862-924. One warp per (row, head).
__global__ void __launch_bounds__(256)
rmsnorm_per_head_kernel(float *x, const float *__restrict__ w, int n_vec, int head_dim, float eps) {
    const int v = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (v >= n_vec) return;
    float *p = x + (size_t)v * head_dim;
    float ss = 0.0f;
    for (int i = lane; i < head_dim; i += 32) ss = fmaf(p[i], p[i], ss);
    ss = warp_sum(ss);
    const float inv = 1.0f / sqrtf(ss / (float)head_dim + eps);
    for (int i = lane; i < head_dim; i += 32) p[i] = p[i] * inv * w[i];
}
void launch_rmsnorm_per_head(cudaStream_t s, float *x, const float *w, int seq, int n_heads, int head_dim, float eps) {
    const int n = seq * n_heads;
    if (n > 0) rmsnorm_per_head_kernel<<<(n + 7) / 8, 256, 0, s>>>(x, w, n, head_dim, eps);
}

// ------------------------------------------------------------------ prefill: q/k norm + RoPE + KV store
// qkv [P, 4096] = q(16x128) | k(8x128) | v(8x128).  grid (P, 32 head slots), 128 threads.
// reference qwen_asr_decoder.c:510-524.
__global__ void __launch_bounds__(128)
qk_norm_rope_store_kernel(const float *__restrict__ qkv, const float *__restrict__ qn, const float *__restrict__ kn,
                          const float *__restrict__ rope_cos, const float *__restrict__ rope_sin, int start_pos,
                          float eps, float *__restrict__ q_out, float *__restrict__ kc, float *__restrict__ vc) {
    pdl_trigger();
    pdl_wait();
    __shared__ float tmp[128];
    __shared__ float red[4];
    const int p = blockIdx.x, slot = blockIdx.y, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int pos = start_pos + p;
    const float v = qkv[(size_t)p * 4096 + slot * 128 + t];
    if (slot >= 24) { // V: plain copy into the cache
        vc[(size_t)pos * 1024 + (slot - 24) * 128 + t] = v;
        return;
    }
    float ss = warp_sum(v * v);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    ss = red[0] + red[1] + red[2] + red[3];
    const float *w = slot < 16 ? qn : kn;
    tmp[t] = v * (1.0f / sqrtf(ss / 128.0f + eps)) * w[t];
    __syncthreads();
    const int d = t & 63;
    const float c = rope_cos[(size_t)pos * 64 + d], sn = rope_sin[(size_t)pos * 64 + d];
    const float r = t < 64 ? tmp[t] * c - tmp[t + 64] * sn : tmp[t] * c + tmp[t - 64] * sn;
    if (slot < 16) q_out[(size_t)p * 2048 + slot * 128 + t] = r;
    else kc[(size_t)pos * 1024 + (slot - 16) * 128 + t] = r;
}
void launch_qk_norm_rope_store(cudaStream_t s, const float *qkv, const float *qn, const float *kn, const float *rope_cos,
                               const float *rope_sin, int start_pos, int P, float eps, float *q_out, float *kc, float *vc) {
    if (P > 0) launch_pdl(qk_norm_rope_store_kernel, dim3(P, 32), 128, 0, s, qkv, qn, kn, rope_cos, rope_sin, start_pos, eps, q_out, kc, vc);
}

// Batched variant (qasr_batch.cu): row r belongs to unit row_unit[r] at position row_pos[r]; K/V rows go to that unit's
// cache kpool + unit * unit_stride (this layer's head-major [kv head][cap][128] block): a head's keys are contiguous, so the
// decode attention streams them as one sequential range instead of 512-byte pieces 4 KB apart.
__global__ void __launch_bounds__(128)
qk_norm_rope_store_rows_kernel(const float *__restrict__ qkv, const float *__restrict__ qn, const float *__restrict__ kn,
                               const float *__restrict__ rope_cos, const float *__restrict__ rope_sin, const int *__restrict__ row_unit,
                               const int *__restrict__ row_pos, float eps, float *__restrict__ q_out, float *__restrict__ kpool,
                               float *__restrict__ vpool, size_t unit_stride, size_t head_stride) {
    __shared__ float tmp[128];
    __shared__ float red[4];
    const int p = blockIdx.x, slot = blockIdx.y, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int pos = row_pos[p];
    const size_t ub = (size_t)row_unit[p] * unit_stride;
    const float v = qkv[(size_t)p * 4096 + slot * 128 + t];
    if (slot >= 24) {
        vpool[ub + (size_t)(slot - 24) * head_stride + (size_t)pos * 128 + t] = v;
        return;
    }
    float ss = warp_sum(v * v);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    ss = red[0] + red[1] + red[2] + red[3];
    const float *w = slot < 16 ? qn : kn;
    tmp[t] = v * (1.0f / sqrtf(ss / 128.0f + eps)) * w[t];
    __syncthreads();
    const int d = t & 63;
    const float c = rope_cos[(size_t)pos * 64 + d], sn = rope_sin[(size_t)pos * 64 + d];
    const float r = t < 64 ? tmp[t] * c - tmp[t + 64] * sn : tmp[t] * c + tmp[t - 64] * sn;
    if (slot < 16) q_out[(size_t)p * 2048 + slot * 128 + t] = r;
    else kpool[ub + (size_t)(slot - 16) * head_stride + (size_t)pos * 128 + t] = r;
}
void launch_qk_norm_rope_store_rows(cudaStream_t s, const float *qkv, const float *qn, const float *kn, const float *rope_cos,
                                    const float *rope_sin, const int *d_row_unit, const int *d_row_pos, int R, float eps, float *q_out,
                                    float *kpool, float *vpool, size_t unit_stride, size_t head_stride) {
    if (R > 0) qk_norm_rope_store_rows_kernel<<<dim3(R, 32), 128, 0, s>>>(qkv, qn, kn, rope_cos, rope_sin, d_row_unit, d_row_pos, eps, q_out, kpool, vpool, unit_stride, head_stride);
}

// ------------------------------------------------------------------ online-softmax helpers
template <int NV>
__device__ __forceinline__ void soft_update(float sc, float &m, float &l, float (&acc)[NV], const float (&v)[NV]) {
    if (sc > m) {
        const float c = expf(m - sc);
        l = l * c + 1.0f;
#pragma unroll
        for (int i = 0; i < NV; i++) acc[i] = acc[i] * c + v[i];
        m = sc;
    } else {
        const float w = expf(sc - m);
        l += w;
#pragma unroll
        for (int i = 0; i < NV; i++) acc[i] = fmaf(w, v[i], acc[i]);
    }
}

// ------------------------------------------------------------------ tiled attention core (prefill + encoder)
// Lane-per-key scheme: for a tile of 32 keys every lane owns ONE key and computes its full head_dim dot product
// against 4 queries of the warp (K tile transposed in shared memory so the lane's float4 reads are conflict free,
// q rows read as broadcasts) - no per-key warp reduction.  The softmax of the tile needs two warp reductions per
// query per 32 keys (max, sum); the probabilities go through a per-warp scratch so that in the P.V pass every lane
// owns HD/32 output dims.  Online softmax across tiles with the reference's initial max of -1e30
// (qwen_asr_kernels.c:1054-1148); same arithmetic as the per-key recurrence, different rounding order.
#define ATT_KT 32
template <int HD>
struct AttSmem {
    float4 kt[HD / 4][ATT_KT + 1];   // K tile transposed: [dim/4][key] (+1 float4 of padding: conflict-free stores)
    float vs[ATT_KT][HD];            // V tile
    float qs[32][HD];                // 32 queries of the CTA
    float4 ps[8][ATT_KT];            // per-warp probabilities [key][4 queries]
};

// one 32-key tile for the 4 queries of a warp.  hi[j]: query j may attend absolute keys < hi[j] (and >= t0 - always true here)
template <int HD>
__device__ __forceinline__ void att_tile(AttSmem<HD> &sm, int warp, int lane, int t0, int nk, const int (&hi)[4], float scale,
                                         float (&m)[4], float (&l)[4], float (&acc)[4][HD / 32]) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int d4 = 0; d4 < HD / 4; d4++) {
        const float4 k4 = sm.kt[d4][lane];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float4 q4 = reinterpret_cast<const float4 *>(sm.qs[warp * 4 + j])[d4];
            s[j] = fmaf(q4.x, k4.x, fmaf(q4.y, k4.y, fmaf(q4.z, k4.z, fmaf(q4.w, k4.w, s[j]))));
        }
    }
    float pj[4];
    int kmax = 0; // keys of this tile any of the 4 queries attends (warp-uniform)
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool valid = lane < nk && t0 + lane < hi[j];
        const float sj = valid ? s[j] * scale : -INFINITY;
        const float mnew = fmaxf(m[j], warp_max(sj));
        const float pv = valid ? expf(sj - mnew) : 0.0f;
        const float c = expf(m[j] - mnew);
        l[j] = l[j] * c + warp_sum(pv);
#pragma unroll
        for (int i = 0; i < HD / 32; i++) acc[j][i] *= c;
        m[j] = mnew;
        pj[j] = pv;
        kmax = max(kmax, min(nk, hi[j] - t0));
    }
    sm.ps[warp][lane] = make_float4(pj[0], pj[1], pj[2], pj[3]);
    __syncwarp();
    for (int kk = 0; kk < kmax; kk++) {
        const float4 p4 = sm.ps[warp][kk];
        const float pq[4] = {p4.x, p4.y, p4.z, p4.w};
        float vv[HD / 32];
        if (HD == 128) {
            const float4 v4 = reinterpret_cast<const float4 *>(sm.vs[kk])[lane];
            vv[0] = v4.x; vv[1] = v4.y; vv[HD / 32 - 2] = v4.z; vv[HD / 32 - 1] = v4.w;
        } else {
            const float2 v2 = reinterpret_cast<const float2 *>(sm.vs[kk])[lane];
            vv[0] = v2.x; vv[1] = v2.y;
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < HD / 32; i++) acc[j][i] = fmaf(pq[j], vv[i], acc[j][i]);
    }
    __syncwarp();
}

// K/V tile = rows [t0, t0+nk) of k / v (row stride ld, head column offset col0).  The 256 threads of the CTA fetch it
// into registers (att_fetch_tile) and store it to shared memory, K transposed (att_store_tile), as two steps, so the
// global loads of tile t+1 are in flight while tile t is being consumed: one L2 round trip per tile, off the critical path.
template <int HD>
struct AttRegs {
    static constexpr int N = ATT_KT * (HD / 4) / 256; // float4 per thread per matrix
    float4 k[N], v[N];
};
template <int HD>
__device__ __forceinline__ void att_fetch_tile(AttRegs<HD> &r, const float *k, const float *v, size_t ld, int col0, int t0, int nk) {
    constexpr int C4 = HD / 4;
#pragma unroll
    for (int i = 0; i < AttRegs<HD>::N; i++) {
        const int e = threadIdx.x + i * 256;
        if (e < nk * C4) {
            const int kk = e / C4, c4 = e - kk * C4;
            r.k[i] = *reinterpret_cast<const float4 *>(k + (size_t)(t0 + kk) * ld + col0 + c4 * 4);
            r.v[i] = *reinterpret_cast<const float4 *>(v + (size_t)(t0 + kk) * ld + col0 + c4 * 4);
        }
    }
}
template <int HD>
__device__ __forceinline__ void att_store_tile(AttSmem<HD> &sm, const AttRegs<HD> &r, int nk) {
    constexpr int C4 = HD / 4;
#pragma unroll
    for (int i = 0; i < AttRegs<HD>::N; i++) {
        const int e = threadIdx.x + i * 256;
        if (e < nk * C4) {
            const int kk = e / C4, c4 = e - kk * C4;
            sm.kt[c4][kk] = r.k[i];
            reinterpret_cast<float4 *>(sm.vs[kk])[c4] = r.v[i];
        }
    }
}

// ------------------------------------------------------------------ causal GQA attention (prefill)
// reference qwen_asr_kernels.c:1101-1148.  head_dim = 128.  CTA = (kv head, 16 query positions) = 32 queries (both
// query heads of the kv head share every K/V tile); warp w owns queries 4w..4w+3 = positions p0+2w, p0+2w+1 x 2 heads.
// Query position i attends keys [0, q_offset + i].
__device__ __forceinline__ void attn_prefill_body(const float *__restrict__ q, const float *__restrict__ kc, const float *__restrict__ vc,
                                                  int q_offset, int P, int seq_k, int n_heads, int n_kv_heads, float scale, float *of, bf16_t *ohi,
                                                  bf16_t *olo, size_t out_row0, int kvh, int ib, int kld_override = 0, int kcol_override = -1) {
    extern __shared__ __align__(16) uint8_t att_raw[];
    AttSmem<128> &sm = *reinterpret_cast<AttSmem<128> *>(att_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = n_heads / n_kv_heads; // 2
    // K/V rows: [key][n_kv_heads * 128] with this head at column kvh * 128 (the reference's cache rows), or - batched pool -
    // a head-major block [key][128] the caller already points at
    const size_t qld = (size_t)n_heads * 128, kld = kld_override ? (size_t)kld_override : (size_t)n_kv_heads * 128;
    const int kcol = kcol_override >= 0 ? kcol_override : kvh * 128;
    const int kmax_cta = min(q_offset + min(ib + 16, P), seq_k); // keys needed by the last position of the CTA
    AttRegs<128> regs;
    att_fetch_tile<128>(regs, kc, vc, kld, kcol, 0, min(ATT_KT, kmax_cta));
    for (int e = threadIdx.x; e < 32 * 32; e += 256) { // 32 queries x 32 float4
        const int qi = e >> 5, c4 = e & 31, pos = ib + (qi >> 1), hh = kvh * per + (qi & 1);
        reinterpret_cast<float4 *>(sm.qs[qi])[c4] = pos < P ? *reinterpret_cast<const float4 *>(q + (size_t)pos * qld + hh * 128 + c4 * 4)
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    int hi[4];
    float m[4], l[4], acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int pos = ib + ((warp * 4 + j) >> 1);
        hi[j] = pos < P ? min(q_offset + pos + 1, seq_k) : 0;
        m[j] = -1e30f; l[j] = 0.0f;
        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
    }
    for (int t0 = 0; t0 < kmax_cta; t0 += ATT_KT) {
        const int nk = min(ATT_KT, kmax_cta - t0);
        __syncthreads(); // the previous tile has been consumed
        att_store_tile<128>(sm, regs, nk);
        __syncthreads();
        if (t0 + ATT_KT < kmax_cta) att_fetch_tile<128>(regs, kc, vc, kld, kcol, t0 + ATT_KT, min(ATT_KT, kmax_cta - t0 - ATT_KT));
        if (t0 < max(max(hi[0], hi[1]), max(hi[2], hi[3]))) att_tile<128>(sm, warp, lane, t0, nk, hi, scale, m, l, acc);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int qi = warp * 4 + j, pos = ib + (qi >> 1), hh = kvh * per + (qi & 1);
        if (pos >= P) continue;
        const float inv = l[j] > 0.0f ? 1.0f / l[j] : 0.0f;
        const size_t base = (out_row0 + (size_t)pos) * qld + hh * 128 + lane * 4;
#pragma unroll
        for (int c = 0; c < 4; c++) store_out(acc[j][c] * inv, base + c, of, ohi, olo);
    }
}

__global__ void __launch_bounds__(256)
attn_prefill_kernel(const float *__restrict__ q, const float *__restrict__ kc, const float *__restrict__ vc,
                    int q_offset, int P, int seq_k, int n_heads, int n_kv_heads, float scale, float *of, bf16_t *ohi,
                    bf16_t *olo) {
    pdl_trigger();
    pdl_wait();
    attn_prefill_body(q, kc, vc, q_offset, P, seq_k, n_heads, n_kv_heads, scale, of, ohi, olo, 0, blockIdx.x, blockIdx.y * 16);
}

// Batched variant (qasr_batch.cu): blockIdx.z = unit.  Unit u owns rows [row0[u], row0[u] + P[u]) of the concatenated
// q / output matrices and its own KV cache kpool + u * unit_stride (this layer's head-major [kv head][cap][128] block, head_stride = cap * 128);
// every unit starts at position 0 (a fresh segment / utterance, reference qwen_asr.c:763).
__global__ void __launch_bounds__(256)
attn_prefill_batch_kernel(const float *__restrict__ q, const float *__restrict__ kpool, const float *__restrict__ vpool, size_t unit_stride,
                          size_t head_stride, const int *__restrict__ row0, const int *__restrict__ Ps, int n_heads, int n_kv_heads, float scale,
                          bf16_t *ohi, bf16_t *olo) {
    const int u = blockIdx.z, P = Ps[u], ib = blockIdx.y * 16;
    if (ib >= P) return;
    const size_t r0 = (size_t)row0[u];
    attn_prefill_body(q + r0 * n_heads * 128, kpool + u * unit_stride + blockIdx.x * head_stride, vpool + u * unit_stride + blockIdx.x * head_stride, 0, P, P,
                      n_heads, n_kv_heads, scale, nullptr, ohi, olo, r0, blockIdx.x, ib, 128, 0);
}
// ------------------------------------------------------------------ causal GQA attention on tensor cores (prefill)
// The same attention as attn_prefill_body (reference qwen_asr_kernels.c:1101-1148), with Q.K^T and P.V on mma.sync
// (m16n8k16, bf16 in / f32 out).  The reference works in f32, so every operand is split a = hi + lo (two bf16) and each
// product is the three-term sum hi.hi + hi.lo + lo.hi (the dropped lo.lo term is ~2^-18 relative): scores and outputs carry
// ~16 mantissa bits, like the GEMMs.  CTA = (kv head, 32 query positions) = 64 query rows (both query heads of the
// kv head share every K/V tile); warp w owns 16 rows (head w >> 1, positions 16 (w & 1) ...).  K/V tiles of 64 keys are read
// from the f32 cache, split and stored as bf16 hi / lo planes in shared memory (16-byte chunks XOR-swizzled by the key, so
// ldmatrix is conflict free); B fragments come from ldmatrix (.trans for V).  Online softmax per row in f32 (initial max
// -1e30 like the reference).  kld / kcol: row pitch and first column of this kv head inside a K/V row.
#define ATC_KT 64
struct AttTcSmem {
    uint16_t k[2][ATC_KT * 128]; // [plane hi/lo][key][dim], 256 B per key, chunk c stored at c ^ (key & 7)
    uint16_t v[2][ATC_KT * 128];
};
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void split_pair(float a, float b, uint32_t &hi, uint32_t &lo) { // (a, b) -> bf16x2 hi and bf16x2 lo (a in the low half)
    const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
    const __nv_bfloat16 la = __float2bfloat16_rn(a - __bfloat162float(ha)), lb = __float2bfloat16_rn(b - __bfloat162float(hb));
    hi = (uint32_t)__bfloat16_as_ushort(ha) | ((uint32_t)__bfloat16_as_ushort(hb) << 16);
    lo = (uint32_t)__bfloat16_as_ushort(la) | ((uint32_t)__bfloat16_as_ushort(lb) << 16);
}

__device__ __forceinline__ void attn_prefill_tc_body(const float *__restrict__ q, const float *__restrict__ kc, const float *__restrict__ vc,
                                                     int q_offset, int P, int seq_k, int n_heads, float scale, float *of, bf16_t *ohi, bf16_t *olo,
                                                     size_t out_row0, int kvh, int ib, size_t kld, int kcol) {
    extern __shared__ __align__(16) uint8_t att_raw[];
    AttTcSmem &sm = *reinterpret_cast<AttTcSmem *>(att_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const size_t qld = (size_t)n_heads * 128;
    const int head = kvh * 2 + (warp >> 1), p0 = ib + (warp & 1) * 16; // this warp: 16 positions of one query head
    const int r0 = p0 + g, r1 = p0 + g + 8;                            // the two rows (positions) this thread holds
    // Q fragments (A operand), hi / lo: k-step kk covers dims 16 kk .. 16 kk + 15
    uint32_t qh[8][4], ql[8][4];
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
        const int d0 = 16 * kk + 2 * t;
        const float2 z = make_float2(0.f, 0.f);
        const float2 a0 = r0 < P ? *reinterpret_cast<const float2 *>(q + (size_t)r0 * qld + head * 128 + d0) : z;
        const float2 a1 = r1 < P ? *reinterpret_cast<const float2 *>(q + (size_t)r1 * qld + head * 128 + d0) : z;
        const float2 a2 = r0 < P ? *reinterpret_cast<const float2 *>(q + (size_t)r0 * qld + head * 128 + d0 + 8) : z;
        const float2 a3 = r1 < P ? *reinterpret_cast<const float2 *>(q + (size_t)r1 * qld + head * 128 + d0 + 8) : z;
        split_pair(a0.x, a0.y, qh[kk][0], ql[kk][0]);
        split_pair(a1.x, a1.y, qh[kk][1], ql[kk][1]);
        split_pair(a2.x, a2.y, qh[kk][2], ql[kk][2]);
        split_pair(a3.x, a3.y, qh[kk][3], ql[kk][3]);
    }
    const int hi0 = r0 < P ? min(q_offset + r0 + 1, seq_k) : 0, hi1 = r1 < P ? min(q_offset + r1 + 1, seq_k) : 0; // keys [0, hi) are attended
    const int warp_hi = min(q_offset + min(p0 + 16, P), seq_k);         // keys any row of this warp attends
    const int kmax_cta = min(q_offset + min(ib + 32, P), seq_k);
    float m0 = -1e30f, m1 = -1e30f, l0 = 0.0f, l1 = 0.0f;
    float o[16][4];
#pragma unroll
    for (int j = 0; j < 16; j++) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.0f;
    const uint32_t sk_hi = (uint32_t)__cvta_generic_to_shared(sm.k[0]), sk_lo = (uint32_t)__cvta_generic_to_shared(sm.k[1]);
    const uint32_t sv_hi = (uint32_t)__cvta_generic_to_shared(sm.v[0]), sv_lo = (uint32_t)__cvta_generic_to_shared(sm.v[1]);
    const int lm = lane >> 3, lr = lane & 7; // ldmatrix: this lane addresses row lr of matrix lm

    for (int t0 = 0; t0 < kmax_cta; t0 += ATC_KT) {
        __syncthreads(); // the previous tile has been consumed
        // K / V tile: 64 keys x 128 dims, f32 -> bf16 hi / lo planes (rows past the last needed key are zero)
        for (int e = tid; e < ATC_KT * 32; e += 128) {
            const int key = e >> 5, c4 = e & 31;
            float4 kv4 = make_float4(0.f, 0.f, 0.f, 0.f), vv4 = kv4;
            if (t0 + key < kmax_cta) {
                kv4 = *reinterpret_cast<const float4 *>(kc + (size_t)(t0 + key) * kld + kcol + c4 * 4);
                vv4 = *reinterpret_cast<const float4 *>(vc + (size_t)(t0 + key) * kld + kcol + c4 * 4);
            }
            const int off = key * 128 + ((((c4 >> 1) ^ (key & 7)) << 3) | ((c4 & 1) << 2)); // element offset of 4 consecutive dims
            uint32_t h0, h1, l0w, l1w;
            split_pair(kv4.x, kv4.y, h0, l0w); split_pair(kv4.z, kv4.w, h1, l1w);
            *reinterpret_cast<uint2 *>(&sm.k[0][off]) = make_uint2(h0, h1);
            *reinterpret_cast<uint2 *>(&sm.k[1][off]) = make_uint2(l0w, l1w);
            split_pair(vv4.x, vv4.y, h0, l0w); split_pair(vv4.z, vv4.w, h1, l1w);
            *reinterpret_cast<uint2 *>(&sm.v[0][off]) = make_uint2(h0, h1);
            *reinterpret_cast<uint2 *>(&sm.v[1][off]) = make_uint2(l0w, l1w);
        }
        __syncthreads();
        if (t0 >= warp_hi) continue; // causal: nothing in this tile for the rows of this warp (the barriers above stay CTA-uniform)
        // ---- S = Q K^T over the 64 keys of the tile: 8 n-tiles of 8 keys
        float sacc[8][4];
#pragma unroll
        for (int j = 0; j < 8; j++) sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int key = 8 * j + lr;
#pragma unroll
            for (int kk2 = 0; kk2 < 4; kk2++) { // dims 32 kk2 .. 32 kk2 + 31: matrices lm = 0..3 are the four 8-dim chunks
                const uint32_t offb = (uint32_t)(key * 256 + (((4 * kk2 + lm) ^ (key & 7)) << 4));
                uint32_t bh[4], bl[4];
                ldsm_x4(bh, sk_hi + offb);
                ldsm_x4(bl, sk_lo + offb);
                mma_bf16_16816(sacc[j], qh[2 * kk2], bh[0], bh[1]);
                mma_bf16_16816(sacc[j], qh[2 * kk2], bl[0], bl[1]);
                mma_bf16_16816(sacc[j], ql[2 * kk2], bh[0], bh[1]);
                mma_bf16_16816(sacc[j], qh[2 * kk2 + 1], bh[2], bh[3]);
                mma_bf16_16816(sacc[j], qh[2 * kk2 + 1], bl[2], bl[3]);
                mma_bf16_16816(sacc[j], ql[2 * kk2 + 1], bh[2], bh[3]);
            }
        }
        // ---- online softmax: this thread holds keys t0 + 8 j + 2 t, + 1 of rows r0 (c0, c1) and r1 (c2, c3)
        float mx0 = -1e30f, mx1 = -1e30f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int key = t0 + 8 * j + 2 * t;
            sacc[j][0] = key < hi0 ? sacc[j][0] * scale : -INFINITY;
            sacc[j][1] = key + 1 < hi0 ? sacc[j][1] * scale : -INFINITY;
            sacc[j][2] = key < hi1 ? sacc[j][2] * scale : -INFINITY;
            sacc[j][3] = key + 1 < hi1 ? sacc[j][3] * scale : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(sacc[j][0], sacc[j][1]));
            mx1 = fmaxf(mx1, fmaxf(sacc[j][2], sacc[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float c0 = expf(m0 - mn0), c1 = expf(m1 - mn1);
        m0 = mn0; m1 = mn1;
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            sacc[j][0] = expf(sacc[j][0] - mn0); sacc[j][1] = expf(sacc[j][1] - mn0);
            sacc[j][2] = expf(sacc[j][2] - mn1); sacc[j][3] = expf(sacc[j][3] - mn1);
            s0 += sacc[j][0] + sacc[j][1];
            s1 += sacc[j][2] + sacc[j][3];
        }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        l0 = l0 * c0 + s0; l1 = l1 * c1 + s1;
#pragma unroll
        for (int j = 0; j < 16; j++) { o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1; }
        // ---- O += P V: k-step kk covers keys 16 kk .. 16 kk + 15 = S n-tiles 2 kk and 2 kk + 1
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            uint32_t ph[4], pl[4];
            split_pair(sacc[2 * kk][0], sacc[2 * kk][1], ph[0], pl[0]);
            split_pair(sacc[2 * kk][2], sacc[2 * kk][3], ph[1], pl[1]);
            split_pair(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1], ph[2], pl[2]);
            split_pair(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3], ph[3], pl[3]);
            const int key = 16 * kk + (lm & 1) * 8 + lr;
#pragma unroll
            for (int j2 = 0; j2 < 8; j2++) { // dim tiles 2 j2 and 2 j2 + 1
                const uint32_t offb = (uint32_t)(key * 256 + (((2 * j2 + (lm >> 1)) ^ (key & 7)) << 4));
                uint32_t vh[4], vl[4];
                ldsm_x4_t(vh, sv_hi + offb);
                ldsm_x4_t(vl, sv_lo + offb);
                mma_bf16_16816(o[2 * j2], ph, vh[0], vh[1]);
                mma_bf16_16816(o[2 * j2], ph, vl[0], vl[1]);
                mma_bf16_16816(o[2 * j2], pl, vh[0], vh[1]);
                mma_bf16_16816(o[2 * j2 + 1], ph, vh[2], vh[3]);
                mma_bf16_16816(o[2 * j2 + 1], ph, vl[2], vl[3]);
                mma_bf16_16816(o[2 * j2 + 1], pl, vh[2], vh[3]);
            }
        }
    }
    const float inv0 = l0 > 0.0f ? 1.0f / l0 : 0.0f, inv1 = l1 > 0.0f ? 1.0f / l1 : 0.0f;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const int col = head * 128 + 8 * j + 2 * t;
        if (r0 < P) {
            const size_t base = (out_row0 + (size_t)r0) * qld + col;
            store_out(o[j][0] * inv0, base, of, ohi, olo);
            store_out(o[j][1] * inv0, base + 1, of, ohi, olo);
        }
        if (r1 < P) {
            const size_t base = (out_row0 + (size_t)r1) * qld + col;
            store_out(o[j][2] * inv1, base, of, ohi, olo);
            store_out(o[j][3] * inv1, base + 1, of, ohi, olo);
        }
    }
}

__global__ void __launch_bounds__(128)
attn_prefill_tc_kernel(const float *__restrict__ q, const float *__restrict__ kc, const float *__restrict__ vc, int q_offset, int P, int seq_k,
                       int n_heads, int n_kv_heads, float scale, float *of, bf16_t *ohi, bf16_t *olo) {
    pdl_trigger();
    pdl_wait();
    attn_prefill_tc_body(q, kc, vc, q_offset, P, seq_k, n_heads, scale, of, ohi, olo, 0, blockIdx.x, blockIdx.y * 32, (size_t)n_kv_heads * 128, blockIdx.x * 128);
}
__global__ void __launch_bounds__(128)
attn_prefill_tc_batch_kernel(const float *__restrict__ q, const float *__restrict__ kpool, const float *__restrict__ vpool, size_t unit_stride,
                             size_t head_stride, const int *__restrict__ row0, const int *__restrict__ Ps, int n_heads, float scale, bf16_t *ohi,
                             bf16_t *olo) {
    const int u = blockIdx.z, P = Ps[u], ib = blockIdx.y * 32;
    if (ib >= P) return;
    const size_t r0 = (size_t)row0[u];
    attn_prefill_tc_body(q + r0 * n_heads * 128, kpool + u * unit_stride + blockIdx.x * head_stride, vpool + u * unit_stride + blockIdx.x * head_stride, 0, P, P,
                         n_heads, scale, nullptr, ohi, olo, r0, blockIdx.x, ib, 128, 0);
}
static bool attn_use_tc() { // QASR_ATTN_FFMA=1: the f32 FFMA kernels of round 1 (A/B runs)
    static int v = -1;
    if (v < 0) { const char *e = getenv("QASR_ATTN_FFMA"); v = !(e && e[0] == '1'); }
    return v != 0;
}

static void attn_prefill_opt_in() { // per-device bit: the attribute belongs to the (function, device) pair
    static unsigned attr_set = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr_set >> (dev & 31) & 1u)) {
        cudaFuncSetAttribute(attn_prefill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AttSmem<128>));
        cudaFuncSetAttribute(attn_prefill_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AttSmem<128>));
        cudaFuncSetAttribute(attn_prefill_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AttTcSmem));
        cudaFuncSetAttribute(attn_prefill_tc_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AttTcSmem));
        attr_set |= 1u << (dev & 31);
    }
}
void launch_attn_prefill(cudaStream_t s, const float *q, const float *kc, const float *vc, int q_offset, int P, int seq_k,
                         int n_heads, int n_kv_heads, float scale, float *out_f32, bf16_t *out_hi, bf16_t *out_lo) {
    if (P <= 0) return;
    attn_prefill_opt_in();
    // short prompts keep the FFMA kernel: its 16-position CTAs give twice the CTAs, and at P = 61 the kernel is pure latency
    if (attn_use_tc() && n_heads == 2 * n_kv_heads && P >= 128) {
        launch_pdl(attn_prefill_tc_kernel, dim3(n_kv_heads, (P + 31) / 32), 128, sizeof(AttTcSmem), s, q, kc, vc, q_offset, P, seq_k, n_heads, n_kv_heads, scale, out_f32, out_hi, out_lo);
        return;
    }
    dim3 grid(n_kv_heads, (P + 15) / 16);
    launch_pdl(attn_prefill_kernel, grid, 256, sizeof(AttSmem<128>), s, q, kc, vc, q_offset, P, seq_k, n_heads, n_kv_heads, scale, out_f32, out_hi, out_lo);
}
void launch_attn_prefill_batch(cudaStream_t s, const float *q, const float *kpool, const float *vpool, size_t unit_stride, size_t head_stride,
                               const int *d_row0, const int *d_P, int n_units, int max_P, int n_heads, int n_kv_heads, float scale, bf16_t *out_hi, bf16_t *out_lo) {
    if (n_units <= 0 || max_P <= 0) return;
    attn_prefill_opt_in();
    if (attn_use_tc() && n_heads == 2 * n_kv_heads && max_P >= 128) {
        attn_prefill_tc_batch_kernel<<<dim3(n_kv_heads, (max_P + 31) / 32, n_units), 128, sizeof(AttTcSmem), s>>>(q, kpool, vpool, unit_stride, head_stride, d_row0, d_P, n_heads, scale, out_hi, out_lo);
        return;
    }
    dim3 grid(n_kv_heads, (max_P + 15) / 16, n_units);
    attn_prefill_batch_kernel<<<grid, 256, sizeof(AttSmem<128>), s>>>(q, kpool, vpool, unit_stride, head_stride, d_row0, d_P, n_heads, n_kv_heads, scale, out_hi, out_lo);
}

// ------------------------------------------------------------------ windowed bidirectional attention (encoder)
// reference qwen_asr_kernels.c:1054-1099.  head_dim = 64; CTA = (head, window, 32 queries), warp w owns queries
// 4w..4w+3; every key of the window is attended.  q/k/v may be column slices of one [T, ld] buffer (fused QKV output).
__global__ void __launch_bounds__(256)
attn_windowed_kernel(const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v, int ld,
                     const int *__restrict__ window_starts, float scale, int out_ld, float *of, bf16_t *ohi,
                     bf16_t *olo) {
    pdl_trigger();
    pdl_wait();
    __shared__ AttSmem<64> sm;
    const int h = blockIdx.x, w = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ws = window_starts[w], we = window_starts[w + 1];
    const int ib = ws + blockIdx.z * 32;
    if (ib >= we) return; // whole CTA
    AttRegs<64> regs;
    att_fetch_tile<64>(regs, k, v, (size_t)ld, h * 64, ws, min(ATT_KT, we - ws));
    for (int e = threadIdx.x; e < 32 * 16; e += 256) { // 32 queries x 16 float4
        const int qi = e >> 4, c4 = e & 15;
        reinterpret_cast<float4 *>(sm.qs[qi])[c4] = ib + qi < we ? *reinterpret_cast<const float4 *>(q + (size_t)(ib + qi) * ld + h * 64 + c4 * 4)
                                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    int hi[4];
    float m[4], l[4], acc[4][2];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        hi[j] = ib + warp * 4 + j < we ? we : 0;
        m[j] = -1e30f; l[j] = 0.0f;
        acc[j][0] = acc[j][1] = 0.0f;
    }
    for (int t0 = ws; t0 < we; t0 += ATT_KT) {
        const int nk = min(ATT_KT, we - t0);
        __syncthreads(); // the previous tile has been consumed
        att_store_tile<64>(sm, regs, nk);
        __syncthreads();
        if (t0 + ATT_KT < we) att_fetch_tile<64>(regs, k, v, (size_t)ld, h * 64, t0 + ATT_KT, min(ATT_KT, we - t0 - ATT_KT));
        if (hi[0] > 0) att_tile<64>(sm, warp, lane, t0, nk, hi, scale, m, l, acc);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int qi = ib + warp * 4 + j;
        if (qi >= we) break;
        const float inv = l[j] > 0.0f ? 1.0f / l[j] : 0.0f;
        const size_t base = (size_t)qi * out_ld + h * 64 + lane * 2;
        store_out(acc[j][0] * inv, base, of, ohi, olo);
        store_out(acc[j][1] * inv, base + 1, of, ohi, olo);
    }
}
void launch_attn_windowed(cudaStream_t s, const float *q, const float *k, const float *v, int ld, int n_heads,
                          const int *d_window_starts, int n_windows, int max_window, float scale, int out_ld,
                          float *out_f32, bf16_t *out_hi, bf16_t *out_lo) {
    if (n_windows <= 0) return;
    dim3 grid(n_heads, n_windows, (max_window + 31) / 32);
    launch_pdl(attn_windowed_kernel, grid, 256, 0, s, q, k, v, ld, d_window_starts, scale, out_ld, out_f32, out_hi, out_lo);
}

// ------------------------------------------------------------------ small element-wise kernels
__global__ void split_f32_kernel(const float *__restrict__ x, size_t n, bf16_t *hi, bf16_t *lo) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        store_out(x[i], i, nullptr, hi, lo);
}
void launch_split_f32(cudaStream_t s, const float *x, size_t n, bf16_t *hi, bf16_t *lo) {
    if (n == 0) return;
    const int blocks = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    split_f32_kernel<<<blocks, 256, 0, s>>>(x, n, hi, lo);
}

// x[m, :] += table[row_idx[m], :]   (per-chunk sinusoidal PE, reference qwen_asr_encoder.c:280-284)
__global__ void add_rows_kernel(float *x, const float *__restrict__ table, const int *__restrict__ row_idx, int d) {
    pdl_trigger();
    pdl_wait();
    const int m = blockIdx.x;
    const float *t = table + (size_t)row_idx[m] * d;
    for (int i = threadIdx.x; i < d; i += blockDim.x) x[(size_t)m * d + i] += t[i];
}
void launch_add_rows(cudaStream_t s, float *x, const float *table, const int *d_row_idx, int M, int d) {
    if (M > 0) launch_pdl(add_rows_kernel, M, 256, 0, s, x, table, d_row_idx, d);
}

// op: 0 add (a+=b) 1 mul (a*=b) 2 scale (a*=s) 3 gelu 4 silu
__global__ void eltwise_kernel(int op, float *a, const float *__restrict__ b, float sc, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v = a[i];
        switch (op) {
            case 0: v += b[i]; break;
            case 1: v *= b[i]; break;
            case 2: v *= sc; break;
            case 3: v = gelu_tanh(v); break;
            default: v = silu(v); break;
        }
        a[i] = v;
    }
}
void launch_eltwise(cudaStream_t s, int op, float *a, const float *b, float scalar, size_t n) {
    if (n == 0) return;
    const int blocks = (int)((n + 255) / 256 < 2368 ? (n + 255) / 256 : 2368);
    eltwise_kernel<<<blocks, 256, 0, s>>>(op, a, b, scalar, n);
}

// reference qwen_asr_kernels.c:946-1010 (out may alias gate_up: each thread reads its pair first,
// and writes index j <= 2j, so a row-serial in-place update is only safe per row; we stage through
// registers per row chunk and sync).
__global__ void swiglu_kernel(float *out, const float *gate_up, int inter) {
    const int srow = blockIdx.x;
    const float *gu = gate_up + (size_t)srow * 2 * inter;
    float *o = out + (size_t)srow * inter;
    for (int j0 = 0; j0 < inter; j0 += blockDim.x) {
        const int j = j0 + threadIdx.x;
        float r = 0.0f;
        if (j < inter) r = silu(gu[2 * j]) * gu[2 * j + 1];
        __syncthreads(); // all reads of this chunk (indices < 2*(j0+blockDim)) done before writes < j0+blockDim
        if (j < inter) o[j] = r;
        __syncthreads();
    }
}
void launch_swiglu(cudaStream_t s, float *out, const float *gate_up, int seq, int inter) {
    if (seq > 0) swiglu_kernel<<<seq, 256, 0, s>>>(out, gate_up, inter);
}

// reference qwen_asr_kernels.c:1012-1029. One CTA per row.
__global__ void softmax_kernel(float *x, int cols) {
    __shared__ float red[32];
    float *row = x + (size_t)blockIdx.x * cols;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) mx = fmaxf(mx, row[i]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < nw; w++) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.0f;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) { const float e = expf(row[i] - mx); row[i] = e; sum += e; }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.0f;
    for (int w = 0; w < nw; w++) sum += red[w];
    const float inv = 1.0f / sum;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) row[i] *= inv;
}
void launch_softmax(cudaStream_t s, float *x, int rows, int cols) {
    if (rows > 0) softmax_kernel<<<rows, 256, 0, s>>>(x, cols);
}

// reference qwen_asr_kernels.c:1233-1298; cos/sin given as [seq, head_dim] (duplicated halves)
__global__ void rope_apply_kernel(float *x, const float *__restrict__ c, const float *__restrict__ sn, int n_heads,
                                  int head_dim) {
    const int s = blockIdx.x, half = head_dim / 2;
    for (int idx = threadIdx.x; idx < n_heads * half; idx += blockDim.x) {
        const int h = idx / half, d = idx % half;
        float *v = x + ((size_t)s * n_heads + h) * head_dim;
        const float x1 = v[d], x2 = v[half + d];
        v[d] = x1 * c[(size_t)s * head_dim + d] - x2 * sn[(size_t)s * head_dim + d];
        v[half + d] = x2 * c[(size_t)s * head_dim + half + d] + x1 * sn[(size_t)s * head_dim + half + d];
    }
}
void launch_rope_apply(cudaStream_t s, float *x, const float *c, const float *sn, int seq, int n_heads, int head_dim) {
    if (seq > 0) rope_apply_kernel<<<seq, 256, 0, s>>>(x, c, sn, n_heads, head_dim);
}

// ------------------------------------------------------------------ exact f32 SIMT GEMM
// C[M,N] = A[M,K] W[N,K]^T + bias.  Serves the f32-weight operators of the level-2 seam
// (qwen_linear / qwen_matmul_t / qwen_conv2d take arbitrary f32 weights, which the bf16
// tensor-core path cannot represent exactly).  64x64 tile, 4x4 per thread, BK = 16.
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float *__restrict__ A, const float *__restrict__ W, const float *__restrict__ bias,
                float *__restrict__ C, int M, int N, int K) {
    __shared__ float As[16][65], Ws[16][65];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int e = threadIdx.x; e < 64 * 16; e += 256) {
            const int r = e >> 4, kk = e & 15;
            As[kk][r] = (m0 + r < M && k0 + kk < K) ? A[(size_t)(m0 + r) * K + k0 + kk] : 0.0f;
            Ws[kk][r] = (n0 + r < N && k0 + kk < K) ? W[(size_t)(n0 + r) * K + k0 + kk] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { a[i] = As[kk][ty * 4 + i]; b[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) C[(size_t)m * N + n] = acc[i][j] + (bias ? bias[n] : 0.0f);
        }
}
void launch_gemm_f32(cudaStream_t s, const float *A, const float *W, const float *bias, float *C, int M, int N, int K) {
    if (M <= 0 || N <= 0) return;
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    gemm_f32_kernel<<<grid, 256, 0, s>>>(A, W, bias, C, M, N, K);
}

// patches[pos][K] with K ordered (ic, ki, kj) as in the reference weight layout [C_out, C_in, kH, kW]
// (reference im2col, qwen_asr_kernels.c:566-590, transposed so the GEMM reads K-contiguous rows).
__global__ void im2col_f32_kernel(const float *__restrict__ in, float *__restrict__ cols, int c_in, int h_in, int w_in,
                                  int kh, int kw, int stride, int padding, int h_out, int w_out) {
    const int K = c_in * kh * kw;
    const size_t total = (size_t)h_out * w_out * K;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int p = (int)(e / K), kidx = (int)(e % K);
        const int ic = kidx / (kh * kw), ki = (kidx / kw) % kh, kj = kidx % kw;
        const int oh = p / w_out, ow = p % w_out;
        const int ih = oh * stride - padding + ki, iw = ow * stride - padding + kj;
        cols[e] = (ih >= 0 && ih < h_in && iw >= 0 && iw < w_in) ? in[((size_t)ic * h_in + ih) * w_in + iw] : 0.0f;
    }
}
void launch_im2col_f32(cudaStream_t s, const float *in, float *cols, int c_in, int h_in, int w_in, int kh, int kw,
                       int stride, int padding, int h_out, int w_out) {
    const size_t total = (size_t)h_out * w_out * c_in * kh * kw;
    if (total == 0) return;
    const int blocks = (int)((total + 255) / 256 < 4736 ? (total + 255) / 256 : 4736);
    im2col_f32_kernel<<<blocks, 256, 0, s>>>(in, cols, c_in, h_in, w_in, kh, kw, stride, padding, h_out, w_out);
}

// out[c][s] = in[s][c] + bias[c]   ([S,C] GEMM result -> reference's [C_out, H_out*W_out] layout)
__global__ void transpose_bias_kernel(const float *__restrict__ in, const float *__restrict__ bias, float *__restrict__ out,
                                      int S, int C) {
    __shared__ float tile[32][33];
    const int s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int s = s0 + r, c = c0 + threadIdx.x;
        tile[r][threadIdx.x] = (s < S && c < C) ? in[(size_t)s * C + c] : 0.0f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, s = s0 + threadIdx.x;
        if (c < C && s < S) out[(size_t)c * S + s] = tile[threadIdx.x][r] + (bias ? bias[c] : 0.0f);
    }
}
void launch_transpose_bias(cudaStream_t s, const float *in, const float *bias, float *out, int S, int C) {
    if (S <= 0 || C <= 0) return;
    dim3 grid((S + 31) / 32, (C + 31) / 32), block(32, 8);
    transpose_bias_kernel<<<grid, block, 0, s>>>(in, bias, out, S, C);
}

// ------------------------------------------------------------------ batched decode step kernels (qasr_batch.cu)
// One new token per sequence.  CTA = (kv head, sequence): per-head RMSNorm + split-half RoPE of the two query heads and of
// the new key (reference qwen_asr_decoder.c:632-646), append of the new K/V row to this sequence's cache at pos[u], then
// GQA attention of both query heads over keys [0, pos[u]] with the reference's online softmax (qwen_asr_kernels.c:1101-1148):
// warp w owns keys w, w+8, ...; a lane holds 4 dims of q (both heads) and of the key / value row; the 8 per-warp states are
// merged in fixed order.  Positions come from device memory so one captured CUDA graph serves every step.
#define ATTD_KEYS 4 /* keys per warp in flight */
__global__ void __launch_bounds__(256)
attn_decode_batch_kernel(const float *__restrict__ qkv /*[B][4096]*/, const float *__restrict__ qn, const float *__restrict__ kn,
                         const float *__restrict__ rope_cos, const float *__restrict__ rope_sin, float *__restrict__ kpool,
                         float *__restrict__ vpool, size_t unit_stride, size_t head_stride, const int *__restrict__ d_pos, float eps, float scale,
                         bf16_t *__restrict__ ohi, bf16_t *__restrict__ olo /*[B][2048] planes*/) {
    __shared__ float4 s_new[3][32];        // roped q0, q1, k_new (4 dims per lane)
    __shared__ float4 s_vnew[32];
    __shared__ float s_acc[8][2][128];
    __shared__ float s_ml[8][2][2];
    pdl_trigger();
    pdl_wait();
    const int kvh = blockIdx.x, u = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pos = d_pos[u];
    const float *row = qkv + (size_t)u * 4096;
    float *kc = kpool + (size_t)u * unit_stride + (size_t)kvh * head_stride, *vc = vpool + (size_t)u * unit_stride + (size_t)kvh * head_stride; // [key][128]
    const size_t hoff = (size_t)lane * 4;
    if (warp < 3) { // warps 0, 1: the two query heads of this kv head; warp 2: the new key
        const float *src = warp < 2 ? row + (2 * kvh + warp) * 128 : row + 2048 + kvh * 128;
        float4 v = *reinterpret_cast<const float4 *>(src + lane * 4);
        const float4 nw = *reinterpret_cast<const float4 *>((warp < 2 ? qn : kn) + lane * 4);
        const float s2 = warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
        const float inv = 1.0f / sqrtf(s2 / 128.0f + eps);
        v.x = v.x * inv * nw.x; v.y = v.y * inv * nw.y; v.z = v.z * inv * nw.z; v.w = v.w * inv * nw.w;
        float4 o; // dims d and d +- 64 live in lanes l and l ^ 16
        o.x = __shfl_xor_sync(0xffffffffu, v.x, 16); o.y = __shfl_xor_sync(0xffffffffu, v.y, 16);
        o.z = __shfl_xor_sync(0xffffffffu, v.z, 16); o.w = __shfl_xor_sync(0xffffffffu, v.w, 16);
        const float4 rc = *reinterpret_cast<const float4 *>(rope_cos + (size_t)pos * 64 + (lane & 15) * 4);
        const float4 rs = *reinterpret_cast<const float4 *>(rope_sin + (size_t)pos * 64 + (lane & 15) * 4);
        const float sgn = lane < 16 ? -1.0f : 1.0f;
        const float4 r = make_float4(v.x * rc.x + sgn * o.x * rs.x, v.y * rc.y + sgn * o.y * rs.y, v.z * rc.z + sgn * o.z * rs.z, v.w * rc.w + sgn * o.w * rs.w);
        s_new[warp][lane] = r;
        if (warp == 2) *reinterpret_cast<float4 *>(kc + (size_t)pos * 128 + hoff) = r;
    } else if (warp == 3) {
        const float4 v = *reinterpret_cast<const float4 *>(row + 3072 + kvh * 128 + lane * 4);
        s_vnew[lane] = v;
        *reinterpret_cast<float4 *>(vc + (size_t)pos * 128 + hoff) = v;
    }
    __syncthreads();
    const float4 q0 = s_new[0][lane], q1 = s_new[1][lane];
    float m0 = -1e30f, l0 = 0.0f, m1 = -1e30f, l1 = 0.0f;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    const int n_keys = pos + 1;
    for (int base = warp; base < n_keys; base += 8 * ATTD_KEYS) {
        float4 kr[ATTD_KEYS], vr[ATTD_KEYS];
#pragma unroll
        for (int i = 0; i < ATTD_KEYS; i++) {
            const int j = base + 8 * i;
            if (j < pos) {
                kr[i] = __ldcg(reinterpret_cast<const float4 *>(kc + (size_t)j * 128 + hoff));
                vr[i] = __ldcg(reinterpret_cast<const float4 *>(vc + (size_t)j * 128 + hoff));
            } else if (j == pos) { kr[i] = s_new[2][lane]; vr[i] = s_vnew[lane]; } // never read the row being appended from the cache
        }
        float d0[ATTD_KEYS], d1[ATTD_KEYS];
#pragma unroll
        for (int i = 0; i < ATTD_KEYS; i++) {
            const bool ok = base + 8 * i < n_keys;
            d0[i] = ok ? q0.x * kr[i].x + q0.y * kr[i].y + q0.z * kr[i].z + q0.w * kr[i].w : 0.0f;
            d1[i] = ok ? q1.x * kr[i].x + q1.y * kr[i].y + q1.z * kr[i].z + q1.w * kr[i].w : 0.0f;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int i = 0; i < ATTD_KEYS; i++) {
                d0[i] += __shfl_xor_sync(0xffffffffu, d0[i], o);
                d1[i] += __shfl_xor_sync(0xffffffffu, d1[i], o);
            }
#pragma unroll
        for (int i = 0; i < ATTD_KEYS; i++) {
            if (base + 8 * i >= n_keys) continue;
            float acc0[4] = {a0.x, a0.y, a0.z, a0.w}, acc1[4] = {a1.x, a1.y, a1.z, a1.w};
            const float vv[4] = {vr[i].x, vr[i].y, vr[i].z, vr[i].w};
            soft_update<4>(d0[i] * scale, m0, l0, acc0, vv);
            soft_update<4>(d1[i] * scale, m1, l1, acc1, vv);
            a0 = make_float4(acc0[0], acc0[1], acc0[2], acc0[3]);
            a1 = make_float4(acc1[0], acc1[1], acc1[2], acc1[3]);
        }
    }
    *reinterpret_cast<float4 *>(&s_acc[warp][0][lane * 4]) = a0;
    *reinterpret_cast<float4 *>(&s_acc[warp][1][lane * 4]) = a1;
    if (lane == 0) { s_ml[warp][0][0] = m0; s_ml[warp][0][1] = l0; s_ml[warp][1][0] = m1; s_ml[warp][1][1] = l1; }
    __syncthreads();
    { // 256 threads = 2 heads x 128 dims: merge the 8 warps in fixed order
        const int h = tid >> 7, dim = tid & 127;
        float M = -1e30f;
#pragma unroll
        for (int w = 0; w < 8; w++) M = fmaxf(M, s_ml[w][h][0]);
        float L = 0.0f, A = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const float e = expf(s_ml[w][h][0] - M);
            L += s_ml[w][h][1] * e;
            A += s_acc[w][h][dim] * e;
        }
        store_out(L > 0.0f ? A / L : 0.0f, (size_t)u * 2048 + (2 * kvh + h) * 128 + dim, nullptr, ohi, olo);
    }
}
void launch_attn_decode_batch(cudaStream_t s, const float *qkv, const float *qn, const float *kn, const float *rope_cos, const float *rope_sin,
                              float *kpool, float *vpool, size_t unit_stride, size_t head_stride, const int *d_pos, int B, float eps, float scale, bf16_t *ohi,
                              bf16_t *olo) {
    if (B > 0) launch_pdl(attn_decode_batch_kernel, dim3(8, B), 256, 0, s, qkv, qn, kn, rope_cos, rope_sin, kpool, vpool, unit_stride, head_stride, d_pos, eps, scale, ohi, olo);
}

// Greedy head of the batched step: argmax over the f32 logits row of sequence u (strict >, ties -> lowest index,
// reference qwen_asr_kernels.c:536-541), token bookkeeping, and the gather of the next input row (exact bf16 -> f32
// upcast, qwen_asr.c:412-419,816).  One CTA per sequence.
__global__ void __launch_bounds__(1024)
argmax_next_kernel(const float *__restrict__ logits, int V, const bf16_t *__restrict__ E, int H, float *__restrict__ x_next,
                   int *__restrict__ d_pos, const int *__restrict__ d_step, int *__restrict__ d_tokens, volatile int *h_tokens, int B, int max_steps) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sv[32];
    __shared__ int si[32];
    __shared__ int s_tok;
    const int u = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *row = logits + (size_t)u * V;
    float bv = -1e30f;
    int bi = 0x7fffffff;
    for (int i = tid; i < V; i += 1024) { // ascending per thread: strict > keeps the lowest index
        const float v = row[i];
        if (v > bv) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        bv = sv[lane]; bi = si[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
            if (bi < 0 || bi >= V) bi = 0; // a row of NaNs never compares greater: keep the index in range
            s_tok = bi;
            const int st = *d_step;
            if (st < max_steps) {
                d_tokens[st * B + u] = bi;
                if (h_tokens) h_tokens[st * B + u] = bi;
            }
            d_pos[u] += 1;
        }
    }
    __syncthreads();
    const int tok = s_tok;
    for (int e = tid; e < H; e += 1024) x_next[(size_t)u * H + e] = __uint_as_float(((uint32_t)E[(size_t)tok * H + e]) << 16);
}
__global__ void advance_step_kernel(int *d_step) {
    pdl_trigger();
    pdl_wait();
    *d_step += 1;
}
void launch_argmax_next(cudaStream_t s, const float *logits, int V, const bf16_t *E, int H, float *x_next, int *d_pos, int *d_step,
                        int *d_tokens, volatile int *h_tokens, int B, int max_steps) {
    if (B <= 0) return;
    launch_pdl(argmax_next_kernel, B, 1024, 0, s, logits, V, E, H, x_next, d_pos, (const int *)d_step, d_tokens, h_tokens, B, max_steps);
    launch_pdl(advance_step_kernel, 1, 1, 0, s, d_step);
}

// Prompt rows of a group of units (reference qwen_asr.c:685-759): unit u's sequence is embed(pre) | its T[u] encoder rows
// (enc rows [enc0[u], enc0[u] + T[u])) | embed(suf).  Rows 0 .. total-2 go to the prefill matrix at row0[u] + i, the last
// row becomes the unit's first decode input x_first[u] (the reference prefills total_seq - 1 rows and steps on the last, :764-769).
__global__ void __launch_bounds__(256)
assemble_prompts_kernel(const bf16_t *__restrict__ E, int H, const int *__restrict__ pre, int n_pre, const int *__restrict__ suf, int n_suf,
                        const float *__restrict__ enc, const int *__restrict__ enc0, const int *__restrict__ Ts, const int *__restrict__ row0,
                        float *__restrict__ X, float *__restrict__ x_first) {
    const int u = blockIdx.y, i = blockIdx.x, T = Ts[u], total = n_pre + T + n_suf;
    if (i >= total) return;
    float *dst = i < total - 1 ? X + ((size_t)row0[u] + i) * H : x_first + (size_t)u * H;
    if (i >= n_pre && i < n_pre + T) {
        const float *src = enc + ((size_t)enc0[u] + (i - n_pre)) * H;
        for (int e = threadIdx.x; e < H; e += 256) dst[e] = src[e];
    } else {
        const int tok = i < n_pre ? pre[i] : suf[i - n_pre - T];
        const bf16_t *src = E + (size_t)tok * H;
        for (int e = threadIdx.x; e < H; e += 256) dst[e] = __uint_as_float(((uint32_t)src[e]) << 16);
    }
}
void launch_assemble_prompts(cudaStream_t s, const bf16_t *E, int H, const int *d_pre, int n_pre, const int *d_suf, int n_suf, const float *enc,
                             const int *d_enc0, const int *d_T, const int *d_row0, int n_units, int max_total, float *X, float *x_first) {
    if (n_units > 0 && max_total > 0)
        assemble_prompts_kernel<<<dim3(max_total, n_units), 256, 0, s>>>(E, H, d_pre, n_pre, d_suf, n_suf, enc, d_enc0, d_T, d_row0, X, x_first);
}

// qasr_decode.cu - token-by-token decoder kernels (HBM-bound path).
//
// Replaces, for seq_len == 1, the reference's threaded bf16 matvec family
// (qwen_asr_kernels.c:336-373,389-460 -> qwen_asr_kernels_avx.c:25-155), the streaming
// argmax head (qwen_asr_kernels.c:486-543) and single-query causal GQA attention
// (qwen_asr_kernels.c:1101-1148), with RMSNorm (:801-860), per-head q/k RMSNorm (:862-924),
// NeoX RoPE (:1233-1298), SwiGLU (:946-1010) and the residual adds fused in.
//
// Layout: weights bf16 row-major [N,K] exactly as in the checkpoint; activations f32;
// KV cache f32 [layer][kv_max][kv_heads*128]; every kernel reads the current position from
// device memory (*d_pos) so one captured CUDA graph replays for every token.
#include "qasr_common.cuh"
#include "qasr_internal.h"

// --------------------------------------------------------------------------------------
// GEMV: one warp per output row, 128-bit streaming weight loads, x staged in shared memory
// (optionally RMS-normalised on the way in), warp-shuffle reduction, fused epilogue.
// Algorithmic bytes per launch = 2*N*K (weights) + 4*K + 4*N.
// --------------------------------------------------------------------------------------
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
gemv_bf16_kernel(const bf16_t *__restrict__ W, const float *__restrict__ x, const float *__restrict__ gamma,
                 float eps, float *out, const float *res /* may alias out */,
                 const float *__restrict__ bias, int N, int K, int epi, const int *__restrict__ d_done) {
    extern __shared__ __align__(16) float xs[];
    float *red = xs + K;
    if (d_done && *d_done) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (gamma) {
        float ss = 0.0f;
        for (int i = tid; i < K; i += WARPS * 32) {
            float v = x[i];
            xs[i] = v;
            ss = fmaf(v, v, ss);
        }
        ss = warp_sum(ss);
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        float tot = 0.0f;
#pragma unroll
        for (int w = 0; w < WARPS; w++) tot += red[w];
        const float inv = 1.0f / sqrtf(tot / (float)K + eps);
        for (int i = tid; i < K; i += WARPS * 32) xs[i] = xs[i] * inv * gamma[i];
    } else {
        for (int i = tid; i < K; i += WARPS * 32) xs[i] = x[i];
    }
    __syncthreads();

    const int row = blockIdx.x * WARPS + warp;
    float acc = 0.0f;
    if (row < N) {
        const uint4 *wr = reinterpret_cast<const uint4 *>(W + (size_t)row * K);
        const int nvec = K >> 3;
        for (int i0 = lane; i0 < nvec; i0 += 128) {
            uint4 w[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int i = i0 + j * 32;
                w[j] = (i < nvec) ? ld_stream_u4(wr + i) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int i = i0 + j * 32;
                if (i < nvec) {
                    const float4 a = *reinterpret_cast<const float4 *>(xs + i * 8);
                    const float4 b = *reinterpret_cast<const float4 *>(xs + i * 8 + 4);
                    acc = dot8(w[j], a, b, acc);
                }
            }
        }
        acc = warp_sum(acc);
    }

    if (epi == QASR_EPI_SWIGLU) {
        __syncthreads();
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (tid < WARPS / 2) {
            const int pair = (blockIdx.x * WARPS) / 2 + tid;
            if (2 * pair + 1 < N) out[pair] = silu(red[2 * tid]) * red[2 * tid + 1];
        }
    } else if (lane == 0 && row < N) {
        float v = acc;
        if (bias) v += bias[row];
        if (epi == QASR_EPI_RESIDUAL) v += res[row];
        out[row] = v;
    }
}

void launch_gemv_bf16(cudaStream_t s, const bf16_t *W, const float *x, const float *gamma, float eps, float *out,
                      const float *res, const float *bias, int N, int K, int epi, const int *d_done) {
    const size_t smem = (size_t)(K + 32) * sizeof(float);
    if (N <= 2048) {
        gemv_bf16_kernel<4><<<(N + 3) / 4, 128, smem, s>>>(W, x, gamma, eps, out, res, bias, N, K, epi, d_done);
    } else {
        gemv_bf16_kernel<8><<<(N + 7) / 8, 256, smem, s>>>(W, x, gamma, eps, out, res, bias, N, K, epi, d_done);
    }
}

// --------------------------------------------------------------------------------------
// Greedy head: argmax_o E[o,:] . rmsnorm(x) without materialising logits.
// Strict '>' starting from -1e30 with rows visited in ascending order => lowest index wins
// ties, as in the reference (qwen_asr_kernels.c:518-543, _generic.c:24-47).
// --------------------------------------------------------------------------------------
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__global__ void __launch_bounds__(256)
argmax_gemv_kernel(const bf16_t *__restrict__ E, const float *__restrict__ x, const float *__restrict__ gamma,
                   float eps, int V, int K, float *__restrict__ part_val, int *__restrict__ part_idx,
                   const int *__restrict__ d_done) {
    extern __shared__ __align__(16) float xs[];
    float *red = xs + K;
    int *redi = reinterpret_cast<int *>(red + 8);
    if (d_done && *d_done) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        float ss = 0.0f;
        for (int i = tid; i < K; i += 256) {
            float v = x[i];
            xs[i] = v;
            ss = fmaf(v, v, ss);
        }
        ss = warp_sum(ss);
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        float tot = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; w++) tot += red[w];
        const float inv = 1.0f / sqrtf(tot / (float)K + eps);
        for (int i = tid; i < K; i += 256) xs[i] = xs[i] * inv * gamma[i];
        __syncthreads();
    }
    const int row0 = blockIdx.x * QASR_ARGMAX_ROWS_PER_CTA + warp * 4;
    const int nvec = K >> 3;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = lane; i < nvec; i += 32) {
        uint4 w[4];
#pragma unroll
        for (int r = 0; r < 4; r++)
            w[r] = (row0 + r < V) ? ld_stream_u4(reinterpret_cast<const uint4 *>(E + (size_t)(row0 + r) * K) + i)
                                  : make_uint4(0, 0, 0, 0);
        const float4 a = *reinterpret_cast<const float4 *>(xs + i * 8);
        const float4 b = *reinterpret_cast<const float4 *>(xs + i * 8 + 4);
#pragma unroll
        for (int r = 0; r < 4; r++) acc[r] = dot8(w[r], a, b, acc[r]);
    }
    float bv = -1e30f;
    int bi = row0 < V ? row0 : V - 1;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const float v = warp_sum(acc[r]);
        if (row0 + r < V && v > bv) { bv = v; bi = row0 + r; }
    }
    __syncthreads();
    if (lane == 0) { red[warp] = bv; redi[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        float v = red[0];
        int i = redi[0];
        for (int w = 1; w < 8; w++)
            if (red[w] > v) { v = red[w]; i = redi[w]; }
        part_val[blockIdx.x] = v;
        part_idx[blockIdx.x] = i;
    }
}

int argmax_num_parts(int V) { return (V + QASR_ARGMAX_ROWS_PER_CTA - 1) / QASR_ARGMAX_ROWS_PER_CTA; }

void launch_argmax_gemv(cudaStream_t s, const bf16_t *E, const float *x, const float *gamma, float eps, int V, int K,
                        float *part_val, int *part_idx, const int *d_done) {
    const size_t smem = (size_t)(K + 32) * sizeof(float);
    argmax_gemv_kernel<<<argmax_num_parts(V), 256, smem, s>>>(E, x, gamma, eps, V, K, part_val, part_idx, d_done);
}

// Reduce the per-CTA winners, publish the token, advance the position, and gather the next
// input row from the tied embedding table so that only the id leaves the device
// (reference does the gather on the host: qwen_asr.c:412-419,816).
__global__ void __launch_bounds__(256)
argmax_finalize_kernel(const float *__restrict__ part_val, const int *__restrict__ part_idx, int n_parts,
                       const bf16_t *__restrict__ E, int H, float *__restrict__ x_next, int *d_tokens, int *d_step,
                       int *d_pos, int *d_done, volatile int *h_tokens, int max_steps) {
    __shared__ float sv[256];
    __shared__ int si[256];
    if (d_done && *d_done) return;
    const int tid = threadIdx.x;
    float bv = -1e30f;
    int bi = 0x7fffffff;
    for (int i = tid; i < n_parts; i += 256) {
        const float v = part_val[i];
        const int idx = part_idx[i];
        if (better(v, idx, bv, bi)) { bv = v; bi = idx; }
    }
    sv[tid] = bv;
    si[tid] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o && better(sv[tid + o], si[tid + o], sv[tid], si[tid])) { sv[tid] = sv[tid + o]; si[tid] = si[tid + o]; }
        __syncthreads();
    }
    const int tok = si[0];
    for (int i = tid; i < H; i += 256) x_next[i] = __uint_as_float(((uint32_t)E[(size_t)tok * H + i]) << 16);
    if (tid == 0) {
        const int step = *d_step;
        if (step < max_steps) {
            d_tokens[step] = tok;
            if (h_tokens) h_tokens[step] = tok;
        }
        *d_step = step + 1;
        *d_pos = *d_pos + 1;
        if (tok == 151643 || tok == 151645) *d_done = 1; // reference qwen_asr.c:792
    }
}

void launch_argmax_finalize(cudaStream_t s, const float *part_val, const int *part_idx, int n_parts, const bf16_t *E,
                            int H, float *x_next, int *d_tokens, int *d_step, int *d_pos, int *d_done,
                            volatile int *h_tokens_mapped, int max_steps) {
    argmax_finalize_kernel<<<1, 256, 0, s>>>(part_val, part_idx, n_parts, E, H, x_next, d_tokens, d_step, d_pos,
                                            d_done, h_tokens_mapped, max_steps);
}

__global__ void set_state_kernel(int *d_pos, int pos, int *d_done, int done, int *d_step, int step) {
    if (threadIdx.x == 0) {
        if (d_pos) *d_pos = pos;
        if (d_done) *d_done = done;
        if (d_step) *d_step = step;
    }
}
void launch_set_state(cudaStream_t s, int *d_pos, int pos, int *d_done, int done, int *d_step, int step) {
    set_state_kernel<<<1, 32, 0, s>>>(d_pos, pos, d_done, done, d_step, step);
}

// Embedding rows: bf16 -> f32 by bits<<16 (exact), reference qwen_asr.c:412-419.
__global__ void embed_gather_kernel(const bf16_t *__restrict__ E, const int *__restrict__ ids, int H,
                                    float *__restrict__ out) {
    const int r = blockIdx.x;
    const bf16_t *src = E + (size_t)ids[r] * H;
    for (int i = threadIdx.x; i < H; i += blockDim.x) out[(size_t)r * H + i] = __uint_as_float(((uint32_t)src[i]) << 16);
}
void launch_embed_gather(cudaStream_t s, const bf16_t *E, const int *d_ids, int n, int H, float *out) {
    if (n > 0) embed_gather_kernel<<<n, 256, 0, s>>>(E, d_ids, H, out);
}

// --------------------------------------------------------------------------------------
// Single-query causal GQA attention over the device-resident KV cache (flash-decoding).
// grid = (kv_heads, QASR_ATTN_SPLITS); the two query heads of a kv head share every K/V load.
// Prologue fuses per-head q/k RMSNorm + NeoX RoPE and the KV append for position *d_pos.
// The last CTA to finish a kv head merges the split partials in fixed order (deterministic).
// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_decode_kernel(const float *__restrict__ qkv, const float *__restrict__ qn, const float *__restrict__ kn,
                   const float *__restrict__ rope_cos, const float *__restrict__ rope_sin, float *__restrict__ kc,
                   float *__restrict__ vc, const int *__restrict__ d_pos, float *__restrict__ part,
                   unsigned *__restrict__ counters, float *__restrict__ out, float eps) {
    __shared__ __align__(16) float qs[2][128];
    __shared__ float tmp[3][128];
    __shared__ float red[3][4];
    __shared__ float wm[4][2], wl[4][2];
    __shared__ __align__(16) float wacc[4][2][128];
    __shared__ int s_last;

    const int h = blockIdx.x, split = blockIdx.y, S = gridDim.y;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int pos = *d_pos, n_keys = pos + 1;
    const int per = (n_keys + S - 1) / S;
    const int k0 = split * per;
    const int k1 = min(n_keys, k0 + per);
    const bool owner = (k0 < k1) && (k1 == n_keys); // this split covers the new position

    // --- prologue: q (2 heads) and, for the owner, k: RMSNorm over 128 then RoPE
    const float q0 = qkv[(2 * h) * 128 + t], q1 = qkv[(2 * h + 1) * 128 + t];
    const float kk = owner ? qkv[2048 + h * 128 + t] : 0.0f;
    float s0 = warp_sum(q0 * q0), s1 = warp_sum(q1 * q1), s2 = warp_sum(kk * kk);
    if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; red[2][warp] = s2; }
    __syncthreads();
    s0 = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    s1 = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    s2 = red[2][0] + red[2][1] + red[2][2] + red[2][3];
    tmp[0][t] = q0 * (1.0f / sqrtf(s0 / 128.0f + eps)) * qn[t];
    tmp[1][t] = q1 * (1.0f / sqrtf(s1 / 128.0f + eps)) * qn[t];
    tmp[2][t] = kk * (1.0f / sqrtf(s2 / 128.0f + eps)) * kn[t];
    __syncthreads();
    {
        const int d = t & 63;
        const float c = rope_cos[(size_t)pos * 64 + d], sn = rope_sin[(size_t)pos * 64 + d];
        const int partner = t < 64 ? t + 64 : t - 64;
        const float sgn = t < 64 ? -1.0f : 1.0f;
        qs[0][t] = tmp[0][t] * c + sgn * tmp[0][partner] * sn;
        qs[1][t] = tmp[1][t] * c + sgn * tmp[1][partner] * sn;
        if (owner) {
            kc[(size_t)pos * 1024 + h * 128 + t] = tmp[2][t] * c + sgn * tmp[2][partner] * sn;
            vc[(size_t)pos * 1024 + h * 128 + t] = qkv[3072 + h * 128 + t];
        }
    }
    __syncthreads();

    // --- main loop: warp w takes keys k0+w, k0+w+4, ...; lane owns dims 4l..4l+3
    const float scale = 0.08838834764831845f; // 1/sqrtf(128)
    const float4 qa = *reinterpret_cast<const float4 *>(&qs[0][lane * 4]);
    const float4 qb = *reinterpret_cast<const float4 *>(&qs[1][lane * 4]);
    float m0 = -1e30f, l0 = 0.0f, m1 = -1e30f, l1 = 0.0f;
    float4 a0 = make_float4(0, 0, 0, 0), a1 = make_float4(0, 0, 0, 0);
    for (int j = k0 + warp; j < k1; j += 4) {
        const float4 kr = *reinterpret_cast<const float4 *>(kc + (size_t)j * 1024 + h * 128 + lane * 4);
        const float4 vr = *reinterpret_cast<const float4 *>(vc + (size_t)j * 1024 + h * 128 + lane * 4);
        float p0 = qa.x * kr.x + qa.y * kr.y + qa.z * kr.z + qa.w * kr.w;
        float p1 = qb.x * kr.x + qb.y * kr.y + qb.z * kr.z + qb.w * kr.w;
        p0 = warp_sum(p0) * scale;
        p1 = warp_sum(p1) * scale;
        if (p0 > m0) {
            const float c = expf(m0 - p0);
            l0 = l0 * c + 1.0f;
            a0.x = a0.x * c + vr.x; a0.y = a0.y * c + vr.y; a0.z = a0.z * c + vr.z; a0.w = a0.w * c + vr.w;
            m0 = p0;
        } else {
            const float w = expf(p0 - m0);
            l0 += w;
            a0.x += w * vr.x; a0.y += w * vr.y; a0.z += w * vr.z; a0.w += w * vr.w;
        }
        if (p1 > m1) {
            const float c = expf(m1 - p1);
            l1 = l1 * c + 1.0f;
            a1.x = a1.x * c + vr.x; a1.y = a1.y * c + vr.y; a1.z = a1.z * c + vr.z; a1.w = a1.w * c + vr.w;
            m1 = p1;
        } else {
            const float w = expf(p1 - m1);
            l1 += w;
            a1.x += w * vr.x; a1.y += w * vr.y; a1.z += w * vr.z; a1.w += w * vr.w;
        }
    }
    if (lane == 0) { wm[warp][0] = m0; wl[warp][0] = l0; wm[warp][1] = m1; wl[warp][1] = l1; }
    *reinterpret_cast<float4 *>(&wacc[warp][0][lane * 4]) = a0;
    *reinterpret_cast<float4 *>(&wacc[warp][1][lane * 4]) = a1;
    __syncthreads();

    // --- merge the 4 warps, publish this split's partial
    float *pbase = part + ((size_t)(h * S + split) * 2) * QASR_ATTN_PART_STRIDE;
#pragma unroll
    for (int hd = 0; hd < 2; hd++) {
        float M = fmaxf(fmaxf(wm[0][hd], wm[1][hd]), fmaxf(wm[2][hd], wm[3][hd]));
        float L = 0.0f, A = 0.0f;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const float e = expf(wm[w][hd] - M);
            L += wl[w][hd] * e;
            A += wacc[w][hd][t] * e;
        }
        pbase[hd * QASR_ATTN_PART_STRIDE + t] = A;
        if (t == 0) { pbase[hd * QASR_ATTN_PART_STRIDE + 128] = M; pbase[hd * QASR_ATTN_PART_STRIDE + 129] = L; }
    }
    __threadfence();
    __syncthreads();
    if (t == 0) {
        const unsigned old = atomicAdd(&counters[h], 1u);
        s_last = (old == (unsigned)(S - 1));
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int hd = 0; hd < 2; hd++) {
        float M = -1e30f;
        for (int sp = 0; sp < S; sp++)
            M = fmaxf(M, __ldcg(part + ((size_t)(h * S + sp) * 2 + hd) * QASR_ATTN_PART_STRIDE + 128));
        float L = 0.0f, A = 0.0f;
        for (int sp = 0; sp < S; sp++) {
            const float *pb = part + ((size_t)(h * S + sp) * 2 + hd) * QASR_ATTN_PART_STRIDE;
            const float e = expf(__ldcg(pb + 128) - M);
            L += __ldcg(pb + 129) * e;
            A += __ldcg(pb + t) * e;
        }
        out[(2 * h + hd) * 128 + t] = L > 0.0f ? A / L : 0.0f;
    }
    if (t == 0) counters[h] = 0;
}

void launch_attn_decode(cudaStream_t s, const float *qkv, const float *qn, const float *kn, const float *rope_cos,
                        const float *rope_sin, float *kc, float *vc, const int *d_pos, float *part,
                        unsigned *counters, float *out, float eps) {
    dim3 grid(8, QASR_ATTN_SPLITS);
    attn_decode_kernel<<<grid, 128, 0, s>>>(qkv, qn, kn, rope_cos, rope_sin, kc, vc, d_pos, part, counters, out, eps);
}

// qasr_conv.cu - audio-encoder conv stem staging kernels.
//
// The reference runs, per 100-frame mel chunk, three 3x3/stride-2/pad-1 convolutions as
// im2col + sgemm (qwen_asr_encoder.c:221-258, qwen_asr_kernels.c:566-590,643-685).  Here all
// chunks of a segment are batched: activations are kept position-major / channel-minor
//     act[chunk][w][h][480]        (w = time, h = frequency)
// so that (a) a GEMM row is one output position and its epilogue writes the next layer's
// input directly, and (b) after stage 3 the 16x480 block of one token is contiguous, which is
// exactly the [T, 7680] operand of conv_out once its columns are permuted at upload
// (reference flatten order ch*16+f, qwen_asr_encoder.c:262-271).
// Stage 1 (C_in = 1, K = 9) is computed directly in f32; stages 2/3 gather K = 9*480 patches
// (tap-major, channel-minor - weights are permuted to match at upload) for the tcgen05 GEMM.
#include "qasr_common.cuh"
#include "qasr_internal.h"

// One CTA per (chunk, ow): 64 oh x 480 oc outputs, + bias, GELU, bf16 hi/lo split.
__global__ void __launch_bounds__(256)
conv1_kernel(const float *__restrict__ mel, int frames, const float *__restrict__ w /*[480][9]*/,
             const float *__restrict__ b, const int *__restrict__ w0s, const int *__restrict__ mel0s,
             const int *__restrict__ off1, bf16_t *__restrict__ ohi, bf16_t *__restrict__ olo) {
    pdl_trigger();
    pdl_wait();
    __shared__ float in_s[3][130]; // [kj][ih+1], zero padded
    __shared__ float w_s[480 * 9];
    __shared__ float b_s[480];
    const int c = blockIdx.x, ow = blockIdx.y;
    const int w0 = w0s[c], w1 = (w0 - 1) / 2 + 1;
    if (ow >= w1) return;
    const int mel0 = mel0s[c];
    for (int e = threadIdx.x; e < 3 * 130; e += 256) {
        const int kj = e / 130, r = e % 130; // r = ih + 1
        const int iw = 2 * ow - 1 + kj, ih = r - 1;
        float v = 0.0f;
        if (iw >= 0 && iw < w0 && ih >= 0 && ih < 128) v = mel[(size_t)ih * frames + mel0 + iw];
        in_s[kj][r] = v;
    }
    for (int e = threadIdx.x; e < 480 * 9; e += 256) w_s[e] = w[e];
    for (int e = threadIdx.x; e < 480; e += 256) b_s[e] = b[e];
    __syncthreads();
    const size_t base = ((size_t)off1[c] + (size_t)ow * 64) * 480;
    for (int e = threadIdx.x; e < 64 * 480; e += 256) {
        const int oh = e / 480, oc = e % 480;
        float acc = 0.0f; // same tap order as the reference's K index (ki, kj)
#pragma unroll
        for (int ki = 0; ki < 3; ki++)
#pragma unroll
            for (int kj = 0; kj < 3; kj++) acc = fmaf(w_s[oc * 9 + ki * 3 + kj], in_s[kj][2 * oh + ki], acc);
        const float v = gelu_tanh(acc + b_s[oc]);
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        ohi[base + e] = __bfloat16_as_ushort(hi);
        if (olo) olo[base + e] = __bfloat16_as_ushort(lo);
    }
}

void launch_conv1(cudaStream_t s, const float *mel, int frames, const float *w, const float *b, const ConvGeom &g,
                  bf16_t *out_hi, bf16_t *out_lo) {
    if (g.n_chunks <= 0) return;
    dim3 grid(g.n_chunks, 50);
    launch_pdl(conv1_kernel, grid, 256, 0, s, mel, frames, w, b, g.d_w0, g.d_mel0, g.d_off1, out_hi, out_lo);
}

// Patch gather for stage 2 (64 x w1 -> 32 x w2) or 3 (32 x w2 -> 16 x w3).
// dst[pos][tap*480 + ic], pos = off_out[c] + ow*Hout + oh; src[(off_in[c] + iw*Hin + ih)*480 + ic].
// One thread moves 16 bytes (8 channels); zero rows implement the padding at chunk edges.
__global__ void __launch_bounds__(256)
im2col_stage_kernel(const bf16_t *__restrict__ src, bf16_t *__restrict__ dst, const int *__restrict__ w0s,
                    const int *__restrict__ off_in, const int *__restrict__ off_out, int n_chunks, int stage,
                    int total_out) {
    pdl_trigger();
    pdl_wait();
    const int Hin = stage == 2 ? 64 : 32, Hout = Hin / 2;
    const long long n_items = (long long)total_out * 9 * 60;
    for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < n_items;
         it += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(it % 60);
        const int tap = (int)((it / 60) % 9);
        const int pos = (int)(it / 540);
        int lo = 0, hi = n_chunks - 1; // chunk containing pos
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (off_out[mid] <= pos) lo = mid; else hi = mid - 1;
        }
        const int c = lo;
        const int w0 = w0s[c], w1 = (w0 - 1) / 2 + 1, w2 = (w1 - 1) / 2 + 1;
        const int Win = stage == 2 ? w1 : w2;
        const int local = pos - off_out[c];
        const int ow = local / Hout, oh = local % Hout;
        const int ki = tap / 3, kj = tap % 3;
        const int ih = 2 * oh - 1 + ki, iw = 2 * ow - 1 + kj;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (ih >= 0 && ih < Hin && iw >= 0 && iw < Win)
            val = *reinterpret_cast<const uint4 *>(src + ((size_t)off_in[c] + (size_t)iw * Hin + ih) * 480 + v * 8);
        *reinterpret_cast<uint4 *>(dst + (size_t)pos * 4320 + tap * 480 + v * 8) = val;
    }
}

void launch_im2col_stage(cudaStream_t s, const bf16_t *src, bf16_t *dst, const ConvGeom &g, int stage) {
    const int total_out = stage == 2 ? g.total2 : g.total3;
    if (total_out <= 0) return;
    const long long n_items = (long long)total_out * 540;
    const int blocks = (int)((n_items + 255) / 256 < 148 * 16 ? (n_items + 255) / 256 : 148 * 16);
    launch_pdl(im2col_stage_kernel, blocks, 256, 0, s, src, dst, g.d_w0, stage == 2 ? g.d_off1 : g.d_off2,
                                              stage == 2 ? g.d_off2 : g.d_off3, g.n_chunks, stage, total_out);
}

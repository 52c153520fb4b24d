// qasr_mel.cu - log-mel front end (reference qwen_mel_spectrogram, qwen_asr_audio.c:293-394).
//
// Pass 1 (one CTA per 8 frames): reflect-pad + periodic Hann window into shared memory,
// direct 201-bin DFT of the 400-sample frame against the f32 cos/sin tables (built on the host
// with the reference's own f32 angle formula, :328-336, and kept resident), power, 128-bin
// Slaney filterbank, log10(max(.,1e-10)), running global max (atomicMax on an order-preserving
// int).  Pass 2: clamp at gmax-8, (v+4)/4, transpose to the reference's [128, frames] layout.
// FP32 FFMA throughout: the front end is <0.1 % of the flops of a segment.
#include "qasr_common.cuh"
#include "qasr_internal.h"

#define MEL_FPB 8     // frames per CTA
#define MEL_NFFT 400
#define MEL_NFREQ 201
#define MEL_TSTRIDE 208 // padded row stride of the [n][k] DFT tables and the [k][m]-major filterbank rows

__device__ __forceinline__ int float_to_ordered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void __launch_bounds__(256)
mel_pass1_kernel(const float *__restrict__ samples, int n, int frames, const float *__restrict__ ct,
                 const float *__restrict__ st, const float *__restrict__ win, const float *__restrict__ fb /*[201][128]*/,
                 float *__restrict__ mel_tmp /*[frames][128]*/, int *__restrict__ gmax) {
    __shared__ float fr[MEL_FPB][MEL_NFFT];
    __shared__ float pw[MEL_FPB][MEL_NFREQ + 3];
    __shared__ float red[8];
    const int t0 = blockIdx.x * MEL_FPB, tid = threadIdx.x;
    // The DFT tables (2 x 333 KB) and the filterbank (103 KB) have usually been evicted by the decode weight stream of the previous
    // utterance; without this every CTA walked them row by row at DRAM latency (~90 us for a 3.6 s utterance).  Pull them into L2 up front.
    for (int i = tid; i < MEL_NFFT * MEL_TSTRIDE / 32; i += 256) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ct + (size_t)i * 32) : "memory");
        asm volatile("prefetch.global.L2 [%0];" ::"l"(st + (size_t)i * 32) : "memory");
    }
    for (int i = tid; i < MEL_NFREQ * 128 / 32; i += 256) asm volatile("prefetch.global.L2 [%0];" ::"l"(fb + (size_t)i * 32) : "memory");
    // windowed frames; padded[i]: i<200 -> x[200-i]; i<200+n -> x[i-200]; else x[n-2-(i-200-n)]  (:301-309)
    for (int e = tid; e < MEL_FPB * MEL_NFFT; e += 256) {
        const int f = e / MEL_NFFT, i = e % MEL_NFFT, t = t0 + f;
        float v = 0.0f;
        if (t < frames) {
            const int p = t * 160 + i;
            int src;
            if (p < 200) src = 200 - p;
            else if (p < 200 + n) src = p - 200;
            else src = n - 2 - (p - 200 - n);
            v = (src >= 0 && src < n) ? samples[src] * win[i] : 0.0f;
        }
        fr[f][i] = v;
    }
    __syncthreads();
    if (tid < MEL_NFREQ) {
        float re[MEL_FPB], im[MEL_FPB];
#pragma unroll
        for (int f = 0; f < MEL_FPB; f++) re[f] = im[f] = 0.0f;
#pragma unroll 8 // the 16 table loads of 8 iterations are issued together: the loop was bound by one L2 round trip per iteration
        for (int j = 0; j < MEL_NFFT; j++) {
            const float c = ct[j * MEL_TSTRIDE + tid], s = st[j * MEL_TSTRIDE + tid];
#pragma unroll
            for (int f = 0; f < MEL_FPB; f++) {
                const float x = fr[f][j];
                re[f] = fmaf(x, c, re[f]);
                im[f] = fmaf(x, s, im[f]);
            }
        }
#pragma unroll
        for (int f = 0; f < MEL_FPB; f++) pw[f][tid] = re[f] * re[f] + im[f] * im[f];
    }
    __syncthreads();
    const int m = tid & 127, fg = tid >> 7; // 2 groups x 4 frames
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int k = 0; k < MEL_NFREQ; k++) {
        const float w = fb[k * 128 + m];
#pragma unroll
        for (int f = 0; f < 4; f++) acc[f] = fmaf(w, pw[fg * 4 + f][k], acc[f]);
    }
    float lmax = -1e30f;
#pragma unroll
    for (int f = 0; f < 4; f++) {
        const int t = t0 + fg * 4 + f;
        if (t < frames) {
            const float v = log10f(fmaxf(acc[f], 1e-10f));
            mel_tmp[(size_t)t * 128 + m] = v;
            lmax = fmaxf(lmax, v);
        }
    }
    lmax = warp_max(lmax);
    if ((tid & 31) == 0) red[tid >> 5] = lmax;
    __syncthreads();
    if (tid == 0) {
        float v = red[0];
        for (int w = 1; w < 8; w++) v = fmaxf(v, red[w]);
        atomicMax(gmax, float_to_ordered(v));
    }
}

// mel is [128][out_stride]; this call fills columns [out_off, out_off + frames) (out_stride == frames, out_off == 0 for a
// single unit; the batched path lays the units of a group side by side along the frame axis)
__global__ void mel_pass2_kernel(const float *__restrict__ mel_tmp, const int *__restrict__ gmax, int frames,
                                 float *__restrict__ mel, int out_stride, int out_off) {
    const float lo = ordered_to_float(*gmax) - 8.0f;
    const size_t total = (size_t)128 * frames;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(e / frames), t = (int)(e % frames);
        float v = mel_tmp[(size_t)t * 128 + m];
        if (v < lo) v = lo;
        mel[(size_t)m * out_stride + out_off + t] = (v + 4.0f) / 4.0f;
    }
}

__global__ void mel_init_kernel(int *gmax) { *gmax = float_to_ordered(-1e30f); }

void launch_mel(cudaStream_t s, const float *samples, int n, int frames, const float *d_cos, const float *d_sin,
                const float *d_win, const float *d_fb, float *mel_tmp, int *d_gmax, float *mel_out, int out_stride, int out_off) {
    mel_init_kernel<<<1, 1, 0, s>>>(d_gmax);
    mel_pass1_kernel<<<(frames + MEL_FPB - 1) / MEL_FPB, 256, 0, s>>>(samples, n, frames, d_cos, d_sin, d_win, d_fb,
                                                                     mel_tmp, d_gmax);
    const size_t total = (size_t)128 * frames;
    const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    mel_pass2_kernel<<<blocks, 256, 0, s>>>(mel_tmp, d_gmax, frames, mel_out, out_stride, out_off);
}


// ------------------------------------------------------------------ PCM16 -> f32 mono 16 kHz
// reference qwen_parse_wav_buffer, qwen_asr_audio.c:81-164: channels averaged in f32 and scaled by 1/32768, then - for
// files that are not at 16 kHz - a 32-tap windowed-sinc resampler evaluated in double (sinc cut at min(ratio, 1), Kaiser
// window beta = 6 with I0 as a 20-term power series, output divided by the sum of the coefficients).  One thread per sample.
__global__ void pcm16_to_mono_kernel(const int16_t *__restrict__ pcm, int n, int channels, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (channels == 1) { out[i] = (float)pcm[i] / 32768.0f; return; }
    float sum = 0.0f;
    for (int c = 0; c < channels; c++) sum += (float)pcm[(size_t)i * channels + c];
    out[i] = (sum / (float)channels) / 32768.0f;
}
__device__ __forceinline__ double bessel_i0_series(double x) {
    double sum = 1.0, term = 1.0;
    const double xx = x * x;
#pragma unroll 1
    for (int k = 1; k <= 20; k++) {
        term *= xx / (4.0 * (double)k * (double)k);
        sum += term;
    }
    return sum;
}
__global__ void resample_sinc_kernel(const float *__restrict__ in, int n, int rate, float *__restrict__ out, int new_n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= new_n) return;
    const double pi = 3.14159265358979323846;
    const double ratio = 16000.0 / (double)rate, cutoff = ratio < 1.0 ? ratio : 1.0, inv_i0 = 1.0 / bessel_i0_series(6.0);
    const double pos = (double)i / ratio;
    const int c = (int)pos;
    double acc = 0.0, wsum = 0.0;
#pragma unroll 1
    for (int j = c - 15; j <= c + 16; j++) {
        const double d = (double)j - pos, x = d * cutoff;
        const double sv = fabs(x) < 1e-9 ? 1.0 : sin(pi * x) / (pi * x);
        const double np_ = d / 16.0;
        const double w = (np_ <= -1.0 || np_ >= 1.0) ? 0.0 : bessel_i0_series(6.0 * sqrt(1.0 - np_ * np_)) * inv_i0;
        const double coeff = sv * w * cutoff;
        if (j >= 0 && j < n) acc += (double)in[j] * coeff;
        wsum += coeff;
    }
    out[i] = wsum > 1e-9 ? (float)(acc / wsum) : 0.0f;
}
void launch_pcm16_to_mono(cudaStream_t s, const int16_t *pcm, int n, int channels, float *out) {
    if (n > 0) pcm16_to_mono_kernel<<<(n + 255) / 256, 256, 0, s>>>(pcm, n, channels, out);
}
void launch_resample_sinc(cudaStream_t s, const float *in, int n, int rate, float *out, int new_n) {
    if (new_n > 0) resample_sinc_kernel<<<(new_n + 127) / 128, 128, 0, s>>>(in, n, rate, out, new_n);
}

// Microbenchmark: latency of one all-to-all exchange step among 148 co-resident CTAs (idle memory system):
//   A. LL all-gather: every CTA stores its share of N {f32, tag} words, every CTA polls all N words
//   B. flat grid barrier (red.release + ld.acquire polling) followed by a plain L2 read of N floats
//   C. LL all-gather with sentinel pre-poll (one word per producer) before the batch load
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ll_exchange ll_exchange.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
typedef unsigned long long u64;
__device__ __forceinline__ void ll_load2(const u64 *p, u64 &a, u64 &b) { asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); }
__device__ __forceinline__ void ll_store(u64 *p, float v, unsigned tag) {
    const u64 w = ((u64)tag << 32) | (u64)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
template <int NP>
__global__ void __launch_bounds__(512, 1) k(u64 *ll_base, float *plain, unsigned *gbar, long long *out, int N, int iters, int mode) {
    const int tid = threadIdx.x, b = blockIdx.x, G = gridDim.x;
    const int r0 = (int)((long long)N * b / G), r1 = (int)((long long)N * (b + 1) / G);
    float acc = 0.f;
    unsigned target = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 1; it <= iters; it++) {
        const unsigned tag = (unsigned)it;
        u64 *ll = ll_base + (size_t)(it & 1) * 8192;   // double-buffered: a fast CTA must not overwrite words a slow one still polls
        if (mode == 0) {
            if (r0 + tid < r1) ll_store(ll + r0 + tid, (float)(it + tid), tag);
            u64 w[NP][2];
#pragma unroll
            for (int i = 0; i < NP; i++) { const int p = tid + i * 512; if (2 * p < N) ll_load2(ll + 2 * p, w[i][0], w[i][1]); else w[i][0] = w[i][1] = (u64)tag << 32; }
            for (;;) {
                bool ok = true;
#pragma unroll
                for (int i = 0; i < NP; i++) ok = ok && (unsigned)(w[i][0] >> 32) == tag && (unsigned)(w[i][1] >> 32) == tag;
                if (__all_sync(0xffffffffu, ok)) break;
#pragma unroll
                for (int i = 0; i < NP; i++) if ((unsigned)(w[i][0] >> 32) != tag || (unsigned)(w[i][1] >> 32) != tag) ll_load2(ll + 2 * (tid + i * 512), w[i][0], w[i][1]);
            }
#pragma unroll
            for (int i = 0; i < NP; i++) acc += __uint_as_float((unsigned)w[i][0]);
            __syncthreads();
        } else if (mode == 1) {
            if (r0 + tid < r1) plain[r0 + tid] = (float)(it + tid);
            target += G;
            __syncthreads();
            if (tid == 0) {
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gbar) : "memory");
                unsigned c;
                do { asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(c) : "l"(gbar) : "memory"); } while ((int)(c - target) < 0);
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < NP; i++) { const int p = tid + i * 512; if (2 * p < N) { float2 v = __ldcg(reinterpret_cast<const float2 *>(plain) + p); acc += v.x; } }
            __syncthreads();
        } else {
            if (r0 + tid < r1) ll_store(ll + r0 + tid, (float)(it + tid), tag);
            if (tid < G) { // sentinel: last word of each producer
                const int s1 = (int)((long long)N * (tid + 1) / G), s0 = (int)((long long)N * tid / G);
                if (s1 > s0) { u64 w; do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(ll + s1 - 1) : "memory"); } while ((unsigned)(w >> 32) != tag); }
            }
            __syncthreads();
            u64 w[NP][2];
#pragma unroll
            for (int i = 0; i < NP; i++) { const int p = tid + i * 512; if (2 * p < N) ll_load2(ll + 2 * p, w[i][0], w[i][1]); else w[i][0] = w[i][1] = (u64)tag << 32; }
            for (;;) {
                bool ok = true;
#pragma unroll
                for (int i = 0; i < NP; i++) ok = ok && (unsigned)(w[i][0] >> 32) == tag && (unsigned)(w[i][1] >> 32) == tag;
                if (__all_sync(0xffffffffu, ok)) break;
#pragma unroll
                for (int i = 0; i < NP; i++) if ((unsigned)(w[i][0] >> 32) != tag || (unsigned)(w[i][1] >> 32) != tag) ll_load2(ll + 2 * (tid + i * 512), w[i][0], w[i][1]);
            }
#pragma unroll
            for (int i = 0; i < NP; i++) acc += __uint_as_float((unsigned)w[i][0]);
            __syncthreads();
        }
    }
    long long t1 = clock64();
    if (tid == 0) out[b] = t1 - t0;
    if (acc == 1234.5f) out[200] = 1;
}
static int g_grid = 148;
template <int NP> void run(int N, int mode, const char *name) {
    u64 *ll; float *plain; unsigned *gbar; long long *out;
    CK(cudaMalloc(&ll, 2 * 8192 * 8)); CK(cudaMemset(ll, 0, 2 * 8192 * 8));
    CK(cudaMalloc(&plain, 8192 * 4)); CK(cudaMalloc(&gbar, 4)); CK(cudaMemset(gbar, 0, 4)); CK(cudaMalloc(&out, 256 * 8));
    int iters = 2000;
    void *args[] = {&ll, &plain, &gbar, &out, &N, &iters, &mode};
    CK(cudaLaunchCooperativeKernel((const void *)k<NP>, dim3(g_grid), dim3(512), args, 0, 0));
    CK(cudaDeviceSynchronize());
    long long h[148]; CK(cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost));
    printf("grid %3d  %-44s N=%4d: %.3f us per exchange\n", g_grid, name, N, (double)h[0] / iters / 1965.0);
    cudaFree(ll); cudaFree(plain); cudaFree(gbar); cudaFree(out);
}
int main(int argc, char **argv) {
    for (int gi = 1; gi < argc; gi++) { g_grid = atoi(argv[gi]); run<2>(1024, 0, "A. LL all-gather"); run<2>(2048, 0, "A. LL all-gather"); run<6>(6144, 0, "A. LL all-gather"); }
    if (argc > 1) return 0;
    run<2>(2048, 0, "A. LL all-gather");
    run<6>(6144, 0, "A. LL all-gather");
    run<2>(2048, 1, "B. grid barrier + plain L2 read");
    run<6>(6144, 1, "B. grid barrier + plain L2 read");
    run<2>(2048, 2, "C. LL all-gather with sentinel pre-poll");
    run<6>(6144, 2, "C. LL all-gather with sentinel pre-poll");
    return 0;
}

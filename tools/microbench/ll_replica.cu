// Microbenchmark: LL all-gather among 148 CTAs with R replicas of the exchange buffer (CTA b polls replica b % R,
// producers store every word R times): does lower read contention on the polled lines shorten the exchange?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ll_replica ll_replica.cu
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
#define NMAX 8192
template <int NP>
__global__ void __launch_bounds__(512, 1) k(u64 *ll_base, long long *out, int N, int R, int iters) {
    const int tid = threadIdx.x, b = blockIdx.x, G = gridDim.x;
    const int r0 = (int)((long long)N * b / G), r1 = (int)((long long)N * (b + 1) / G), nr = r1 - r0;
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 1; it <= iters; it++) {
        const unsigned tag = (unsigned)it;
        u64 *half = ll_base + (size_t)(it & 1) * R * NMAX;
        for (int i = tid; i < nr * R; i += 512) { const int rep = i / nr, row = r0 + i % nr; ll_store(half + (size_t)rep * NMAX + row, (float)(it + row), tag); }
        const u64 *ll = half + (size_t)(b % R) * NMAX;
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
    long long t1 = clock64();
    if (tid == 0) out[b] = t1 - t0;
    if (acc == 1234.5f) out[200] = 1;
}
template <int NP> void run(int N, int R) {
    u64 *ll; long long *out;
    CK(cudaMalloc(&ll, (size_t)2 * R * NMAX * 8)); CK(cudaMemset(ll, 0, (size_t)2 * R * NMAX * 8));
    CK(cudaMalloc(&out, 256 * 8));
    int iters = 2000;
    void *args[] = {&ll, &out, &N, &R, &iters};
    CK(cudaLaunchCooperativeKernel((const void *)k<NP>, dim3(148), dim3(512), args, 0, 0));
    CK(cudaDeviceSynchronize());
    long long h[148]; CK(cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost));
    printf("LL all-gather N=%4d replicas=%3d: %.3f us per exchange\n", N, R, (double)h[0] / iters / 1965.0);
    cudaFree(ll); cudaFree(out);
}
int main() {
    for (int R : {1, 2, 4, 8, 16, 37, 74, 148}) run<2>(2048, R);
    for (int R : {1, 4, 8, 16, 37, 148}) run<6>(6144, R);
    for (int R : {1, 8, 37}) run<1>(512, R);
    return 0;
}

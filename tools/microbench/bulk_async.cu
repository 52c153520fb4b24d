// Microbenchmark: do cp.async.bulk copies make progress while the issuing warps sleep?
// Each of 12 warps per CTA (148 CTAs) issues NS bulk loads of 4 KB, sleeps `sleep_us`, then waits.
// Reports (per CTA 0 warp 0) cycles spent waiting after the sleep, and total time.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t su32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int NS>
__global__ void __launch_bounds__(384, 1) k(const uint8_t *src, size_t stride_cta, int sleep_ns, long long *out, int rounds) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 12 * NS * 4096);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) for (int s = 0; s < NS; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar[warp * NS + s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const uint8_t *base = src + (size_t)blockIdx.x * stride_cta + (size_t)warp * NS * 4096 * rounds;
    long long t_wait = 0, t0 = clock64();
    float acc = 0.f;
    for (int r = 0; r < rounds; r++) {
        if (lane == 0)
            for (int s = 0; s < NS; s++) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar[warp * NS + s])), "r"(4096) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(sm + (warp * NS + s) * 4096)),
                             "l"(base + ((size_t)r * NS + s) * 4096), "r"(4096), "r"(su32(&bar[warp * NS + s])) : "memory");
            }
        if (sleep_ns > 0) { long long ts = clock64(); while (clock64() - ts < (long long)sleep_ns * 2) {} } // ~ns at 1.965 GHz
        long long tw = clock64();
        for (int s = 0; s < NS; s++) {
            uint32_t done;
            do {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(done) : "r"(su32(&bar[warp * NS + s])), "r"(r & 1) : "memory");
            } while (!done);
            acc += reinterpret_cast<float *>(sm + (warp * NS + s) * 4096)[lane];
        }
        t_wait += clock64() - tw;
        __syncwarp();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t_wait; out[1] = clock64() - t0; }
    if (acc == 12345.678f) out[2] = 1;
}
template <int NS> void run(const uint8_t *d, size_t stride, long long *dout, int sleep_ns, int rounds, int grid = 148) {
    size_t smem = 12 * NS * 4096 + 1024;
    cudaFuncSetAttribute(k<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NS><<<grid, 384, smem>>>(d, stride, sleep_ns, dout, rounds); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<NS><<<grid, 384, smem>>>(d, stride, sleep_ns, dout, rounds);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[3]; cudaMemcpy(h, dout, sizeof h, cudaMemcpyDeviceToHost);
    double bytes = (double)grid * 12 * NS * 4096 * rounds;
    printf("grid=%3d NS=%d sleep=%5dns rounds=%d: total %.1f us (%.0f GB/s) | per round: wait %.2f us, total %.2f us | %s\n", grid, NS, sleep_ns, rounds, ms * 1e3, bytes / ms / 1e6,
           h[0] / 1965.0 / rounds, h[1] / 1965.0 / rounds, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    size_t total = (size_t)3 << 30; uint8_t *d; cudaMalloc(&d, total); cudaMemset(d, 1, total);
    long long *dout; cudaMalloc(&dout, 64);
    size_t stride = total / 148 / 4096 * 4096;
    for (int grid : {1, 8, 32, 74, 148}) {
        run<1>(d, stride, dout, 0, 64, grid);
        run<2>(d, stride, dout, 0, 64, grid);
        run<4>(d, stride, dout, 0, 64, grid);
    }
    return 0;
}

// Microbenchmark 3: the two exchange primitives of the cluster decode kernel (qasr_stream2.cu)
//   A. DSMEM all-gather inside a 4-CTA cluster: every CTA pushes W {f32, tag} words into all 4 shared memories
//      (st.shared::cluster.u64) and polls its own copy - no L2 round trip
//   B. all-reduce through L2 atomics: every cluster adds its partial vector (fixed-point i64 + arrival count in the top
//      byte, red.global.add.u64) into N accumulators; every CTA polls the N words until the count field has advanced by
//      the number of clusters.  Integer addition is associative => bitwise deterministic sums.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_exchange cluster_exchange.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
typedef unsigned long long u64;
#define THREADS 512
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, int rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int CS>
__global__ void __launch_bounds__(THREADS, 1) k_dsmem(long long *out, int W, int iters) {
    extern __shared__ __align__(16) uint8_t sm[];
    u64 *buf = reinterpret_cast<u64 *>(sm); // [2][CS * W]
    const int tid = threadIdx.x;
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = tid; i < 2 * CS * W; i += THREADS) buf[i] = 0;
    __syncthreads();
    cluster_sync();
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 1; it <= iters; it++) {
        u64 *b = buf + (it & 1) * CS * W;
        if (tid < W) {
            const u64 w = ((u64)(unsigned)it << 32) | (u64)__float_as_uint((float)(it + tid));
            const uint32_t a = smem_u32(b + rank * W + tid);
#pragma unroll
            for (int r = 0; r < CS; r++) asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(mapa(a, r)), "l"(w) : "memory");
        }
        for (int i = tid; i < CS * W; i += THREADS) {
            u64 v;
            do { asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(smem_u32(b + i)) : "memory"); } while ((unsigned)(v >> 32) != (unsigned)it);
            acc += __uint_as_float((unsigned)v);
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 1234.5f) out[255] = 1;
    cluster_sync();
}

// B: N accumulators, sector stride `stride` (32 = contiguous).  CTA rank r of each cluster contributes rows [r N/CS, (r+1) N/CS).
#define BIAS (1ull << 49)
template <int CS, int NP>
__global__ void __launch_bounds__(THREADS, 1) k_allreduce(uint8_t *base, long long *out, int N, int stride, int iters, size_t bufbytes) {
    const int tid = threadIdx.x, ncl = gridDim.x / CS;
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int per = N / CS;
    u64 prev[2][NP];
#pragma unroll
    for (int i = 0; i < NP; i++) prev[0][i] = prev[1][i] = 0;
    float acc = 0.f;
    cluster_sync();
    long long t0 = clock64();
#pragma unroll 1
    for (int it2 = 0; it2 < iters; it2 += 2) {
#pragma unroll
        for (int ph = 0; ph < 2; ph++) { // two alternating accumulator sets (as xwo / xdn in the kernel)
            uint8_t *buf = base + (size_t)ph * bufbytes;
            for (int i = tid; i < per; i += THREADS) {
                const int row = rank * per + i;
                const long long fx = __float2ll_rn((float)(row & 7) * 0.37f * 4294967296.0f);
                const u64 w = (u64)(fx + (long long)BIAS) + (1ull << 56);
                asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(buf + (size_t)(row >> 2) * stride + (row & 3) * 8), "l"(w) : "memory");
            }
            u64 w[NP];
            const uint8_t *pp[NP];
#pragma unroll
            for (int i = 0; i < NP; i++) {
                const int row = tid + i * THREADS;
                pp[i] = buf + (size_t)(row >> 2) * stride + (row & 3) * 8;
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w[i]) : "l"(pp[i]) : "memory");
            }
            for (;;) {
                bool ok = true;
#pragma unroll
                for (int i = 0; i < NP; i++) ok = ok && (unsigned)((w[i] - prev[ph][i]) >> 56) == (unsigned)ncl;
                if (__all_sync(0xffffffffu, ok)) break;
#pragma unroll
                for (int i = 0; i < NP; i++)
                    if ((unsigned)((w[i] - prev[ph][i]) >> 56) != (unsigned)ncl) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w[i]) : "l"(pp[i]) : "memory");
            }
#pragma unroll
            for (int i = 0; i < NP; i++) {
                const u64 d = w[i] - prev[ph][i];
                prev[ph][i] = w[i];
                const long long s = (long long)(d & ((1ull << 56) - 1)) - (long long)ncl * (long long)BIAS;
                acc += (float)s * (1.0f / 4294967296.0f);
            }
            __syncthreads();
        }
    }
    long long t1 = clock64();
    if (tid == 0) { out[blockIdx.x] = t1 - t0; reinterpret_cast<float *>(out + 200)[blockIdx.x & 1] = acc; }
}

template <class K, class... A>
static int launch_cluster(K kern, int CS, size_t smem, bool coop, int *G_out, A... args) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(CS);
    int ncl = 0;
    CK(cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg));
    const int G = ncl * CS;
    *G_out = G;
    cfg.gridDim = dim3(G);
    cfg.numAttrs = coop ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
    if (e != cudaSuccess) { printf("launch (CS=%d grid=%d coop=%d): %s\n", CS, G, (int)coop, cudaGetErrorString(e)); cudaGetLastError(); return -1; }
    CK(cudaDeviceSynchronize());
    return 0;
}

int main() {
    long long *out; CK(cudaMalloc(&out, 256 * 8));
    long long h[256];
    const size_t smem = 200 * 1024; // like the decode kernel: one CTA per SM
    for (int coop = 1; coop >= 0; coop--) {
        int G = 0, iters = 4000;
        for (int W : {96, 130, 384}) {
            if (launch_cluster(k_dsmem<4>, 4, smem, coop, &G, out, W, iters) == 0) {
                CK(cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost));
                printf("A. DSMEM all-gather CS=4 W=%3d coop=%d grid=%d: %.3f us per exchange\n", W, coop, G, (double)h[0] / iters / 1965.0);
            }
            if (launch_cluster(k_dsmem<8>, 8, smem, coop, &G, out, W, iters) == 0) {
                CK(cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost));
                printf("A. DSMEM all-gather CS=8 W=%3d coop=%d grid=%d: %.3f us per exchange\n", W, coop, G, (double)h[0] / iters / 1965.0);
            }
        }
        for (int N : {1024, 2048})
            for (int stride : {32, 256, 1024}) {
                uint8_t *base;
                const size_t bufbytes = (size_t)(N / 4 + 1) * stride + 4096;
                CK(cudaMalloc(&base, 2 * bufbytes)); CK(cudaMemset(base, 0, 2 * bufbytes));
                int r = N == 1024 ? launch_cluster(k_allreduce<4, 2>, 4, smem, coop, &G, base, out, N, stride, iters, bufbytes)
                                  : launch_cluster(k_allreduce<4, 4>, 4, smem, coop, &G, base, out, N, stride, iters, bufbytes);
                if (r == 0) {
                    CK(cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost));
                    printf("B. atomic all-reduce CS=4 N=%4d stride=%4d coop=%d grid=%d: %.3f us per exchange (check %.3f)\n", N, stride, coop, G, (double)h[0] / iters / 1965.0,
                           reinterpret_cast<float *>(h + 200)[0]);
                }
                CK(cudaMemset(base, 0, 2 * bufbytes));
                r = N == 1024 ? launch_cluster(k_allreduce<8, 2>, 8, smem, coop, &G, base, out, N, stride, iters, bufbytes)
                              : launch_cluster(k_allreduce<8, 4>, 8, smem, coop, &G, base, out, N, stride, iters, bufbytes);
                if (r == 0) {
                    CK(cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost));
                    printf("B. atomic all-reduce CS=8 N=%4d stride=%4d coop=%d grid=%d: %.3f us per exchange\n", N, stride, coop, G, (double)h[0] / iters / 1965.0);
                }
                cudaFree(base);
            }
    }
    return 0;
}

// Microbenchmark 2: what bounds one all-to-all exchange among the co-resident CTAs of the decode kernel?
//   layouts : {f32, tag} 8-byte words (T8) or 4-byte words whose mantissa LSB is a phase bit (P4);
//             32-byte sectors contiguous or spread with a stride (one sector per 256 B / 1 KB granule => more L2 slices)
//   readers : every CTA polls every word (flat), or the CTAs of a cluster poll 1/CS of the words each and forward them
//             into all CS shared memories over DSMEM, where the consumers poll locally (CL)
// Steady-state time per exchange = (clock64 after iters) / iters; every iteration depends on the previous one.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ll_exchange2 ll_exchange2.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
typedef unsigned long long u64;
#define THREADS 512
#ifndef ALIGNED
#define ALIGNED 1
#endif
#define NMAXB (64 * 1024 * 1024)

__device__ __forceinline__ void ld16(const uint8_t *p, uint32_t (&w)[4]) {
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void st8(uint8_t *p, uint32_t lo, uint32_t hi) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ void st4(uint8_t *p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// WB = bytes per word (8: value + tag, 4: value with phase bit).  A thread owns NP 16-byte halves of sectors.
template <int WB>
__device__ __forceinline__ bool half_ok(const uint32_t (&w)[4], unsigned tag) {
    if (WB == 8) return w[1] == tag && w[3] == tag;
    const unsigned ph = tag & 1u;
    return ((w[0] & 1u) == ph) && ((w[1] & 1u) == ph) && ((w[2] & 1u) == ph) && ((w[3] & 1u) == ph);
}
__device__ __forceinline__ size_t half_off(int h, int stride) { return (size_t)(h >> 1) * stride + (h & 1) * 16; }
template <int WB>
__device__ __forceinline__ size_t word_off(int i, int stride) { constexpr int WPS = 32 / WB; return (size_t)(i / WPS) * stride + (i % WPS) * WB; }

template <int WB, int NP>
__global__ void __launch_bounds__(THREADS, 1) k_flat(uint8_t *base, long long *out, int N, int stride, int iters, size_t bufbytes) {
    const int tid = threadIdx.x, b = blockIdx.x, G = gridDim.x;
    const int r0 = ALIGNED ? 16 * (int)((long long)(N >> 4) * b / G) : (int)((long long)N * b / G);
    const int r1 = ALIGNED ? 16 * (int)((long long)(N >> 4) * (b + 1) / G) : (int)((long long)N * (b + 1) / G);
    const int halves = N * WB / 16;
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 2; it < iters + 2; it++) {
        // two alternating buffers; P4: the phase of a buffer flips every second iteration
        const unsigned tag = WB == 8 ? (unsigned)it : (unsigned)(it >> 1);
        uint8_t *buf = base + (size_t)(it & 1) * bufbytes;
        if (r0 + tid < r1) {
            const float v = (float)(it + tid);
            if (WB == 8) st8(buf + word_off<WB>(r0 + tid, stride), __float_as_uint(v), tag);
            else st4(buf + word_off<WB>(r0 + tid, stride), (__float_as_uint(v) & ~1u) | (tag & 1u));
        }
        uint32_t w[NP][4];
#pragma unroll
        for (int i = 0; i < NP; i++) {
            const int h = tid + i * THREADS;
            if (h < halves) ld16(buf + half_off(h, stride), w[i]);
            else { w[i][0] = w[i][2] = tag & 1u; w[i][1] = w[i][3] = WB == 8 ? tag : (tag & 1u); }
        }
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int i = 0; i < NP; i++) ok = ok && half_ok<WB>(w[i], tag);
            if (__all_sync(0xffffffffu, ok)) break;
#pragma unroll
            for (int i = 0; i < NP; i++)
                if (!half_ok<WB>(w[i], tag)) ld16(buf + half_off(tid + i * THREADS, stride), w[i]);
        }
#pragma unroll
        for (int i = 0; i < NP; i++) acc += __uint_as_float(w[i][0]);
        __syncthreads();
    }
    long long t1 = clock64();
    if (tid == 0) out[b] = t1 - t0;
    if (acc == 1234.5f) out[200] = 1;
}

// Cluster variant: rank r of a CS-cluster polls halves h with h % CS == r (interleaved: all ranks finish together) and
// forwards them into every member's shared buffer; consumers poll their own shared memory.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, int rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
__device__ __forceinline__ void st_cluster16(uint32_t a, const uint32_t (&w)[4]) {
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}
template <int WB, int NPL /* halves polled from L2 per thread */, int NPS /* halves polled locally per thread */>
__global__ void __launch_bounds__(THREADS, 1) k_cluster(uint8_t *base, long long *out, int N, int stride, int iters, size_t bufbytes, int CS) {
    extern __shared__ __align__(16) uint8_t sm[]; // 2 buffers x N*WB bytes
    const int tid = threadIdx.x, b = blockIdx.x, G = gridDim.x;
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int r0 = ALIGNED ? 16 * (int)((long long)(N >> 4) * b / G) : (int)((long long)N * b / G);
    const int r1 = ALIGNED ? 16 * (int)((long long)(N >> 4) * (b + 1) / G) : (int)((long long)N * (b + 1) / G);
    const int halves = N * WB / 16;
    for (int i = tid; i < 2 * halves * 4; i += THREADS) reinterpret_cast<uint32_t *>(sm)[i] = 0;
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 2; it < iters + 2; it++) {
        const unsigned tag = WB == 8 ? (unsigned)it : (unsigned)(it >> 1);
        uint8_t *buf = base + (size_t)(it & 1) * bufbytes;
        uint8_t *sbuf = sm + (size_t)(it & 1) * halves * 16;
        if (r0 + tid < r1) {
            const float v = (float)(it + tid);
            if (WB == 8) st8(buf + word_off<WB>(r0 + tid, stride), __float_as_uint(v), tag);
            else st4(buf + word_off<WB>(r0 + tid, stride), (__float_as_uint(v) & ~1u) | (tag & 1u));
        }
        // (1) poll my share from L2
        uint32_t w[NPL][4];
#pragma unroll
        for (int i = 0; i < NPL; i++) {
            const int h = (tid + i * THREADS) * CS + (int)rank;
            if (h < halves) ld16(buf + half_off(h, stride), w[i]);
            else { w[i][0] = w[i][2] = tag & 1u; w[i][1] = w[i][3] = WB == 8 ? tag : (tag & 1u); }
        }
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int i = 0; i < NPL; i++) ok = ok && half_ok<WB>(w[i], tag);
            if (__all_sync(0xffffffffu, ok)) break;
#pragma unroll
            for (int i = 0; i < NPL; i++)
                if (!half_ok<WB>(w[i], tag)) ld16(buf + half_off((tid + i * THREADS) * CS + (int)rank, stride), w[i]);
        }
        // (2) forward to every member (tags / phase bits travel with the data)
#pragma unroll
        for (int i = 0; i < NPL; i++) {
            const int h = (tid + i * THREADS) * CS + (int)rank;
            if (h < halves) {
                const uint32_t a = smem_u32(sbuf + (size_t)h * 16);
                for (int r = 0; r < CS; r++) st_cluster16(mapa(a, r), w[i]);
            }
        }
        // (3) poll locally
#pragma unroll
        for (int i = 0; i < NPS; i++) {
            const int h = tid + i * THREADS;
            if (h < halves) {
                uint32_t v[4];
                for (;;) {
                    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(smem_u32(sbuf + (size_t)h * 16)) : "memory");
                    if (half_ok<WB>(v, tag)) break;
                }
                acc += __uint_as_float(v[0]);
            }
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (tid == 0) out[b] = t1 - t0;
    if (acc == 1234.5f) out[200] = 1;
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

static double report(long long *out, int G, int iters, const char *name, int N, int stride) {
    long long h[160];
    CK(cudaMemcpy(h, out, sizeof(long long) * G, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < G; i++) mx = h[i] > mx ? h[i] : mx;
    printf("%-34s N=%5d stride=%5d grid=%3d: %.3f us per exchange (cta0 %.3f)\n", name, N, stride, G, (double)mx / iters / 1965.0, (double)h[0] / iters / 1965.0);
    return 0;
}
template <int WB, int NP> void run_flat(int N, int stride, int G, const char *name) {
    uint8_t *base; long long *out;
    const size_t bufbytes = (size_t)(N * WB / 32 + 1) * stride + 4096;
    CK(cudaMalloc(&base, 2 * bufbytes)); CK(cudaMemset(base, 0, 2 * bufbytes)); CK(cudaMalloc(&out, 256 * 8));
    int iters = 4000;
    void *args[] = {&base, &out, &N, &stride, &iters, (void *)&bufbytes};
    CK(cudaLaunchCooperativeKernel((const void *)k_flat<WB, NP>, dim3(G), dim3(THREADS), args, 0, 0));
    CK(cudaDeviceSynchronize());
    report(out, G, iters, name, N, stride);
    cudaFree(base); cudaFree(out);
}
template <int WB, int NPL, int NPS> void run_cluster(int N, int stride, int CS, const char *name) {
    uint8_t *base; long long *out;
    const size_t bufbytes = (size_t)(N * WB / 32 + 1) * stride + 4096;
    CK(cudaMalloc(&base, 2 * bufbytes)); CK(cudaMemset(base, 0, 2 * bufbytes)); CK(cudaMalloc(&out, 256 * 8));
    int iters = 4000;
    const size_t smem = 140 * 1024; // one CTA per SM
    CK(cudaFuncSetAttribute(k_cluster<WB, NPL, NPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int ncl = 0;
    cfg.gridDim = dim3(CS);
    CK(cudaOccupancyMaxActiveClusters(&ncl, k_cluster<WB, NPL, NPS>, &cfg));
    const int G = (ncl * CS > 148 ? 148 / CS : ncl) * CS;
    cfg.gridDim = dim3(G);
    CK(cudaLaunchKernelEx(&cfg, k_cluster<WB, NPL, NPS>, base, out, N, stride, iters, bufbytes, CS));
    CK(cudaDeviceSynchronize());
    char nm[96];
    snprintf(nm, sizeof nm, "%s CS=%d", name, CS);
    report(out, G, iters, nm, N, stride);
    cudaFree(base); cudaFree(out);
}

// one-way latency: CTA 0 and CTA 1 bounce one {value, tag} word (thread 0 only)
__global__ void k_pingpong(u64 *w, long long *out, int iters) {
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    long long t0 = clock64();
    for (int it = 1; it <= iters; it++) {
        if ((it & 1) == me) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(w), "l"((u64)it) : "memory");
        else { u64 v; do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(w) : "memory"); } while (v != (u64)it); }
    }
    out[me] = clock64() - t0;
}
int main() {
    {
        u64 *w; long long *out; CK(cudaMalloc(&w, 4096)); CK(cudaMemset(w, 0, 4096)); CK(cudaMalloc(&out, 64));
        int iters = 20000;
        void *args[] = {&w, &out, &iters};
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaMemset(w, 0, 4096));
            CK(cudaLaunchCooperativeKernel((const void *)k_pingpong, dim3(2), dim3(32), args, 0, 0));
            CK(cudaDeviceSynchronize());
        }
        long long h[2]; CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
        printf("ping-pong one-way (store -> visible to a polling thread of another SM): %.3f us\n", (double)h[0] / iters / 1965.0);
    }
    printf("ALIGNED=%d\n", ALIGNED);
    for (int G : {148, 144, 136, 132, 128, 64}) run_flat<8, 1>(1024, 32, G, "flat T8");
    for (int G : {148, 144, 128}) run_flat<8, 1>(1024, 512, G, "flat T8");
    for (int G : {148, 128}) run_flat<4, 1>(1024, 1024, G, "flat P4");
    for (int G : {148, 128}) run_flat<8, 2>(2048, 32, G, "flat T8");
    for (int G : {148, 128}) run_flat<8, 2>(2048, 512, G, "flat T8");
    for (int G : {148, 128}) run_flat<8, 6>(6144, 32, G, "flat T8");
    for (int G : {148, 128}) run_flat<8, 6>(6144, 256, G, "flat T8");
    for (int G : {148, 128}) run_flat<4, 3>(6144, 256, G, "flat P4");
    for (int CS : {2, 4}) {
        run_cluster<8, 1, 1>(1024, 32, CS, "cluster T8");
        run_cluster<8, 1, 1>(1024, 512, CS, "cluster T8");
        run_cluster<4, 1, 1>(1024, 1024, CS, "cluster P4");
        run_cluster<8, 1, 2>(2048, 32, CS, "cluster T8");
        run_cluster<8, 1, 2>(2048, 512, CS, "cluster T8");
    }
    return 0;
}

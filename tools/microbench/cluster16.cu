// Feasibility microbenchmark for the cluster-resident decode kernel (qasr_mega3.cu):
//   1. can 8 clusters of 16 CTAs (512 threads, ~220 KB smem each) be co-resident on a B200?
//   2. latency of: cluster barrier | DSMEM broadcast + cluster barrier | cross-cluster "LL" exchange
//      (8-byte {value, tag} stores polled by the same-rank CTAs of the other clusters) | flat grid barrier
//   3. HBM streaming rate of 128 CTAs x 16 warps x NS-slot rings of 2 KB TMA boxes (16 rows x 64 cols,
//      128B swizzle) versus 148 CTAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster16 cluster16.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t su32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t smid() { uint32_t r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

struct ProbeOut { long long t[8]; unsigned sm[128]; };

__global__ void __launch_bounds__(512, 1) probe(unsigned long long *ll, unsigned *gbar, ProbeOut *out, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float *rx = reinterpret_cast<float *>(smem);          // [16][512] broadcast landing zone
    const int tid = threadIdx.x;
    const uint32_t rank = cluster_ctarank(), cid = cluster_id();
    const int nclusters = gridDim.x / 16;
    if (tid == 0) out->sm[blockIdx.x] = smid();
    cluster_sync();
    // 1. bare cluster barrier
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) cluster_sync();
    long long t1 = clock64();
    // 2. DSMEM broadcast: every thread stores one float into all 16 CTAs, then cluster barrier
    float acc = 0.f;
    for (int i = 0; i < iters; i++) {
        const uint32_t local = su32(rx + rank * 512 + tid);
#pragma unroll
        for (int d = 0; d < 16; d++) st_cluster_f32(mapa(local, d), (float)(i + tid));
        cluster_sync();
        acc += rx[((tid + i) & 15) * 512 + tid];
    }
    long long t2 = clock64();
    // 3. LL exchange among the same-rank CTAs of all clusters: 128 threads write {value, tag}, poll the others
    for (int i = 1; i <= iters; i++) {
        if (tid < 128) {
            const unsigned long long v = ((unsigned long long)(unsigned)i << 32) | (unsigned)(tid + i);
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(ll + ((size_t)cid * 16 + rank) * 128 + tid), "l"(v) : "memory");
            float s = 0.f;
            for (int c = 0; c < nclusters; c++) {
                unsigned long long w;
                do {
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(ll + ((size_t)c * 16 + rank) * 128 + tid) : "memory");
                } while ((unsigned)(w >> 32) != (unsigned)i);
                s += __uint_as_float((unsigned)w);
            }
            acc += s;
        }
        __syncthreads();
    }
    long long t3 = clock64();
    // 4. LL exchange + DSMEM broadcast of the reduced 128 values + cluster barrier (= the full cross-cluster step)
    for (int i = iters + 1; i <= 2 * iters; i++) {
        if (tid < 128) {
            const unsigned long long v = ((unsigned long long)(unsigned)i << 32) | (unsigned)(tid + i);
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(ll + ((size_t)cid * 16 + rank) * 128 + tid), "l"(v) : "memory");
            float s = 0.f;
            for (int c = 0; c < nclusters; c++) {
                unsigned long long w;
                do {
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(ll + ((size_t)c * 16 + rank) * 128 + tid) : "memory");
                } while ((unsigned)(w >> 32) != (unsigned)i);
                s += __uint_as_float((unsigned)w);
            }
            const uint32_t local = su32(rx + rank * 128 + tid);
#pragma unroll
            for (int d = 0; d < 16; d++) st_cluster_f32(mapa(local, d), s);
        }
        cluster_sync();
        acc += rx[tid];
    }
    long long t4 = clock64();
    // 5. flat grid barrier (monotonic counter)
    unsigned target = 0;
    for (int i = 0; i < iters; i++) {
        target += gridDim.x;
        __syncthreads();
        if (tid == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gbar) : "memory");
            unsigned c;
            do { asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(c) : "l"(gbar) : "memory"); } while ((int)(c - target) < 0);
        }
        __syncthreads();
    }
    long long t5 = clock64();
    if (blockIdx.x == 0 && tid == 0) {
        out->t[0] = t1 - t0; out->t[1] = t2 - t1; out->t[2] = t3 - t2; out->t[3] = t4 - t3; out->t[4] = t5 - t4;
    }
    if (acc == 1234.5f) out->t[7] = 1;
    cluster_sync();
}

// ---- streaming test: every warp streams its share of [rows, K] bf16 through an NS-slot ring of 2 KB boxes
template <int NS>
__global__ void __launch_bounds__(512, 1) stream_k(const __grid_constant__ CUtensorMap map, int K, int rows_per_cta, float *sink) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 16 * NS * 2048);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) for (int s = 0; s < NS; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar[warp * NS + s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int n_slices = K / 64, groups = rows_per_cta / 16;
    const int n_units = groups * n_slices;              // unit u = group * n_slices + slice; warp takes u % 16 == warp
    const int my_units = (n_units - warp + 15) / 16;
    const int row0 = blockIdx.x * rows_per_cta;
    auto issue = [&](int i) {
        const int u = warp + i * 16, g = u / n_slices, s = u % n_slices, slot = i % NS;
        const uint32_t b = su32(&bar[warp * NS + slot]), dst = su32(smem + (warp * NS + slot) * 2048);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(2048) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(dst), "l"(&map), "r"(b), "r"(s * 64), "r"(row0 + g * 16) : "memory");
    };
    if (lane == 0) for (int i = 0; i < NS && i < my_units; i++) issue(i);
    float acc = 0.f;
    for (int i = 0; i < my_units; i++) {
        const int slot = i % NS;
        uint32_t done;
        do {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }"
                         : "=r"(done) : "r"(su32(&bar[warp * NS + slot])), "r"((i / NS) & 1) : "memory");
        } while (!done);
        acc += reinterpret_cast<float *>(smem + (warp * NS + slot) * 2048)[lane];
        __syncwarp();
        if (lane == 0 && i + NS < my_units) issue(i + NS);
    }
    if (acc == 12345.678f) sink[0] = acc;
}

template <int NS>
static void run_stream(const CUtensorMap &map, int K, size_t rows, int grid, int cluster, float *sink) {
    const int rows_per_cta = (int)(rows / grid / 16 * 16);
    const size_t smem = 16 * NS * 2048 + 16 * NS * 8 + 1024;
    CK(cudaFuncSetAttribute(stream_k<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (cluster > 8) CK(cudaFuncSetAttribute(stream_k<NS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        CK(cudaLaunchKernelEx(&cfg, stream_k<NS>, map, K, rows_per_cta, sink));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double bytes = (double)grid * rows_per_cta * K * 2;
    printf("stream grid=%3d cluster=%2d NS=%d (%3d KB in flight/SM): %.1f us  %.0f GB/s\n", grid, cluster, NS, 16 * NS * 2, best * 1e3, bytes / best / 1e6);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("%s, %d SMs, smem/block optin %zu\n", prop.name, prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
    const size_t smem = 220 * 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (int cs : {8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 8); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe, &cfg);
        printf("cluster size %2d: max active clusters = %d (%s)\n", cs, n, cudaGetErrorString(e));
    }
    unsigned long long *ll; unsigned *gbar; ProbeOut *out;
    CK(cudaMalloc(&ll, 8 * 16 * 128 * 8)); CK(cudaMemset(ll, 0, 8 * 16 * 128 * 8));
    CK(cudaMalloc(&gbar, 4)); CK(cudaMemset(gbar, 0, 4));
    CK(cudaMalloc(&out, sizeof(ProbeOut))); CK(cudaMemset(out, 0, sizeof(ProbeOut)));
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(128); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = 2;
        const int iters = 2000;
        cudaError_t e = cudaLaunchKernelEx(&cfg, probe, ll, gbar, out, iters);
        printf("cooperative + cluster16 launch: %s\n", cudaGetErrorString(e));
        e = cudaDeviceSynchronize();
        printf("sync: %s\n", cudaGetErrorString(e));
        if (e == cudaSuccess) {
            ProbeOut h; CK(cudaMemcpy(&h, out, sizeof h, cudaMemcpyDeviceToHost));
            const char *nm[5] = {"cluster barrier", "DSMEM bcast(512 thr x16) + cluster barrier", "LL exchange (8 same-rank CTAs) + bar.sync",
                                 "LL exchange + DSMEM bcast + cluster barrier", "flat grid barrier (128 CTAs)"};
            for (int i = 0; i < 5; i++) printf("  %-48s %8.0f cycles  %.3f us\n", nm[i], (double)h.t[i] / iters, (double)h.t[i] / iters / 1965.0);
            printf("  smid per cluster:");
            for (int c = 0; c < 8; c++) { printf("\n   c%d:", c); for (int r = 0; r < 16; r++) printf(" %3u", h.sm[c * 16 + r]); }
            printf("\n");
        }
    }
    // streaming
    const int K = 2048; const size_t rows = (size_t)3 << 30 >> 12; // 3 GiB of bf16 rows
    uint16_t *w; CK(cudaMalloc(&w, rows * K * 2)); CK(cudaMemset(w, 1, rows * K * 2));
    float *sink; CK(cudaMalloc(&sink, 4));
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, 16}, estr[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); return 1; }
    run_stream<5>(map, K, rows, 148, 1, sink);
    run_stream<5>(map, K, rows, 128, 16, sink);
    run_stream<5>(map, K, rows, 128, 1, sink);
    run_stream<3>(map, K, rows, 128, 16, sink);
    run_stream<2>(map, K, rows, 128, 16, sink);
    run_stream<6>(map, K, rows, 128, 16, sink);
    run_stream<5>(map, K, rows, 112, 16, sink);
    return 0;
}

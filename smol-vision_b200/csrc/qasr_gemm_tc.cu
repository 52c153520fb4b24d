// qasr_gemm_tc.cu - tcgen05 / TMEM / TMA bf16 GEMM for the encoder and the decoder prefill.
//
//   C[M,N] (+epilogue) = A[M,K] * W[N,K]^T         A, W bf16 K-major (row-major), f32 accumulate
//
// replaces cblas_sgemm behind qwen_linear (reference qwen_asr_kernels.c:196-224) and the
// "convert the whole bf16 matrix to f32 scratch, then sgemm" path of qwen_linear_nobias_bf16
// for seq_len > 1 (:462-472).  The reference keeps activations in f32; to reproduce it on bf16
// tensor cores A is supplied as two bf16 planes, hi = RN(a) and lo = RN(a - hi), and each
// k-block issues MMA(A_hi, W) and MMA(A_lo, W) into the same TMEM accumulator (weights are
// exact bf16, so the product carries ~16 mantissa bits of A).  A_lo == NULL gives the plain
// single-bf16 GEMM.
//
// Structure (one CTA per 128 x BN output tile, 192 threads):
//   warp 0   : TMA producer  - cp.async.bulk.tensor.2d, 128B-swizzled [128 x 64] A tiles and
//              [BN x 64] W tiles into a 4-stage shared-memory ring, mbarrier expect_tx
//   warp 1   : TMEM allocator + MMA issuer - one elected lane issues tcgen05.mma
//              (cta_group::1, kind::f16, M=128, N=BN, K=16), tcgen05.commit frees ring slots
//   warps 2-5: epilogue - tcgen05.ld 32x32b from TMEM (warp%4 selects the 32-lane quarter),
//              bias / residual / GELU / SwiGLU, f32 or bf16 hi/lo stores
// Ragged M, N and K are handled by TMA zero fill on load and masking on store.
#include "qasr_common.cuh"
#include "qasr_internal.h"

#include <cuda.h>
#include <stdio.h>
#include <string.h>

#define TC_BM 128
#define TC_BK 64
#define TC_STAGES 4
#define TC_THREADS 192

static char g_tc_err[256] = "";
const char *gemm_tc_error(void) { return g_tc_err; }

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread i of the warp gets TMEM lane (base_lane + i), columns c..c+31
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile: rows are 128 bytes (64 bf16), 8-row groups are 1024 bytes
// apart.  Field layout per cute::UMMA::SmemDescriptor (mma_sm100_desc.hpp): start>>4 [0,14),
// LBO>>4 [16,30) (=1, unused for swizzled K-major), SBO>>4 [32,46) (=1024>>4), version=1 [46,48),
// layout_type [61,64) = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

struct TcParams {
    int M, N, K;
    int nsplit; // 1 or 2 A planes
    GemmEpilogue epi;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stage][A_hi 16K | A_lo 16K | B BN*128] then barriers
    constexpr uint32_t A_BYTES = TC_BM * TC_BK * 2;
    constexpr uint32_t B_BYTES = BN * TC_BK * 2;
    constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + B_BYTES;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + TC_STAGES * STAGE_BYTES);
    uint64_t *empty_bar = full_bar + TC_STAGES;
    uint64_t *tmem_full_bar = empty_bar + TC_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
    const int num_kb = (p.K + TC_BK - 1) / TC_BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        if (p.nsplit == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) { // TMEM allocation: BN f32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const uint32_t tx = (p.nsplit == 2 ? 2 * A_BYTES : A_BYTES) + B_BYTES;
            for (int kb = 0; kb < num_kb; kb++) {
                const int s = kb % TC_STAGES;
                const uint32_t ph = (kb / TC_STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t *st = smem + s * STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], tx);
                tma_load_2d(st, &tmA_hi, &full_bar[s], kb * TC_BK, m0);
                if (p.nsplit == 2) tma_load_2d(st + A_BYTES, &tmA_lo, &full_bar[s], kb * TC_BK, m0);
                tma_load_2d(st + 2 * A_BYTES, &tmB, &full_bar[s], kb * TC_BK, n0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (single elected lane) =====
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1,
            // B=bf16 [10,13)=1, A/B K-major [15],[16]=0, N>>3 [17,23), M>>4 [24,29)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(TC_BM >> 4) << 24);
            for (int kb = 0; kb < num_kb; kb++) {
                const int s = kb % TC_STAGES;
                const uint32_t ph = (kb / TC_STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES);
                const uint64_t dah = make_sw128_desc(a_hi);
                const uint64_t dal = make_sw128_desc(a_hi + A_BYTES);
                const uint64_t db = make_sw128_desc(a_hi + 2 * A_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / 16; k++) // +32 bytes (16 bf16) along K inside the swizzle atom
                    tc_mma_bf16(tmem_base, dah + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                if (p.nsplit == 2) {
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; k++)
                        tc_mma_bf16(tmem_base, dal + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
                }
                tc_commit(&empty_bar[s]); // frees the ring slot once the MMAs above have read it
            }
            tc_commit(tmem_full_bar); // accumulator complete
        }
    } else {
        // ===== epilogue warps 2..5 =====
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int q = warp & 3; // TMEM lane quarter this warp may access
        const int row = m0 + q * 32 + lane;
        const GemmEpilogue &e = p.epi;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            if (row < p.M) {
                const int nb = n0 + c0;
                if (e.mode == QASR_GEMM_F32 || e.mode == QASR_GEMM_RESIDUAL) {
                    float *orow = e.out_f32 + (size_t)row * e.ldo;
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const int n = nb + j;
                        if (n < p.N) {
                            float v = __uint_as_float(r[j]);
                            if (e.bias) v += e.bias[n];
                            if (e.mode == QASR_GEMM_RESIDUAL) v += orow[n];
                            orow[n] = v;
                        }
                    }
                } else if (e.mode == QASR_GEMM_GELU_SPLIT) {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const int n = nb + j;
                        if (n < p.N) {
                            float v = __uint_as_float(r[j]);
                            if (e.bias) v += e.bias[n];
                            v = gelu_tanh(v);
                            __nv_bfloat16 hi, lo;
                            split_bf16(v, hi, lo);
                            e.out_hi[(size_t)row * e.ldo + n] = __bfloat16_as_ushort(hi);
                            if (e.out_lo) e.out_lo[(size_t)row * e.ldo + n] = __bfloat16_as_ushort(lo);
                        }
                    }
                } else { // SWIGLU: columns (2j, 2j+1) = (gate_j, up_j), reference qwen_asr_decoder.c:140-152
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const int n = nb + j;
                        if (n + 1 < p.N) {
                            const float g = __uint_as_float(r[j]), u = __uint_as_float(r[j + 1]);
                            const float v = silu(g) * u;
                            __nv_bfloat16 hi, lo;
                            split_bf16(v, hi, lo);
                            e.out_hi[(size_t)row * e.ldo + (n >> 1)] = __bfloat16_as_ushort(hi);
                            if (e.out_lo) e.out_lo[(size_t)row * e.ldo + (n >> 1)] = __bfloat16_as_ushort(lo);
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

template <int BN>
static constexpr size_t tc_smem_bytes() {
    return (size_t)TC_STAGES * (2 * TC_BM * TC_BK * 2 + BN * TC_BK * 2) + 256 + 1024;
}

int gemm_tc_init(void) {
    if (g_encode) return 0;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) {
        snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled unavailable: %s", cudaGetErrorString(e));
        return -1;
    }
    g_encode = (PFN_encodeTiled)fn;
    e = cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes<128>());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes<64>());
    if (e != cudaSuccess) {
        snprintf(g_tc_err, sizeof g_tc_err, "cudaFuncSetAttribute(gemm_tc): %s", cudaGetErrorString(e));
        g_encode = nullptr;
        return -1;
    }
    return 0;
}

// 2-D bf16 tensor [rows, K] row-major; box = [box_rows x 64] with 128-byte swizzle; OOB -> zeros
static int make_map(CUtensorMap *m, const bf16_t *ptr, int rows, int K, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)ptr, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ptr=%p", (int)r, rows, K, (const void *)ptr);
        return -1;
    }
    return 0;
}

static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
int tc_encode_map(void *out_map64, const bf16_t *ptr, int rows, int K, int box_cols, int box_rows) {
    if (gemm_tc_init() != 0) return -1;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode((CUtensorMap *)out_map64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)ptr, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled failed (%d) rows=%d K=%d", (int)r, rows, K);
        return -1;
    }
    return 0;
}

int launch_gemm_tc(cudaStream_t s, const bf16_t *A_hi, const bf16_t *A_lo, int M, int K, const bf16_t *W, int N,
                   const GemmEpilogue &epi) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    if (gemm_tc_init() != 0) return -1;
    if ((K & 7) || ((uintptr_t)A_hi & 15) || ((uintptr_t)W & 15) || (A_lo && ((uintptr_t)A_lo & 15))) {
        snprintf(g_tc_err, sizeof g_tc_err, "gemm_tc: K must be a multiple of 8 and operands 16-byte aligned (K=%d)", K);
        return -1;
    }
    // Narrow tiles when the 128-wide grid would leave most of the 148 SMs idle.
    const int tiles128 = ((M + TC_BM - 1) / TC_BM) * ((N + 127) / 128);
    const bool bn64 = tiles128 < 120;
    TcParams p;
    p.M = M; p.N = N; p.K = K;
    p.nsplit = A_lo ? 2 : 1;
    p.epi = epi;
    CUtensorMap ma, ml, mb;
    if (make_map(&ma, A_hi, M, K, TC_BM) != 0) return -1;
    if (make_map(&ml, A_lo ? A_lo : A_hi, M, K, TC_BM) != 0) return -1;
    if (make_map(&mb, W, N, K, bn64 ? 64 : 128) != 0) return -1;
    if (bn64) {
        dim3 grid((N + 63) / 64, (M + TC_BM - 1) / TC_BM);
        gemm_tc_kernel<64><<<grid, TC_THREADS, tc_smem_bytes<64>(), s>>>(ma, ml, mb, p);
    } else {
        dim3 grid((N + 127) / 128, (M + TC_BM - 1) / TC_BM);
        gemm_tc_kernel<128><<<grid, TC_THREADS, tc_smem_bytes<128>(), s>>>(ma, ml, mb, p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_tc_err, sizeof g_tc_err, "gemm_tc launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

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
#include <stdlib.h>
#include <string.h>

#define TC_BM 128
#define TC_BK 64
#define TC_STAGES 4
#define TC_EPI_WARPS 8 /* at most; the launch picks 4 or 8 epilogue warps (warp e drains TMEM lane quarter e % 4, column part e / 4 of the tile).
                           Measured (profiles/r02_batched_path.txt): 8 warps help where the epilogue is exposed - short mainloops (K <= 1024: 0.6B
                           prefill gate/up M = 16440 327 -> 296 us) and few tiles per CTA (encoder fc1 / fc2 at M = 3120: 76 -> 62, 87 -> 71 us) - and cost
                           4-7 % on the long steady-state GEMMs (1.7B prefill at M = 25856), where the epilogue hides behind the next mainloop anyway */
#define TC_THREADS ((2 + TC_EPI_WARPS) * 32)

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
    int tiles_m, tiles_n; // persistent tile loop: tile t -> (m = t / tiles_n, n = t % tiles_n), n fastest so the CTAs running
                          // at the same time share A tiles (and the few W tiles of their column range) in L2
    GemmEpilogue epi;
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void st_v4(void *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_hi2(float a, float b, float &ra, float &rb) { // bf16 pair of (a, b) and the residuals a - hi, b - hi
    const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
    ra = a - __bfloat162float(ha); rb = b - __bfloat162float(hb);
    return (uint32_t)__bfloat16_as_ushort(ha) | ((uint32_t)__bfloat16_as_ushort(hb) << 16);
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
}

// Epilogue of one 32-row x 32-column block held as "thread = row, registers = 32 consecutive columns" (the layout
// tcgen05.ld delivers).  Every thread writes its own row with 16-byte vector stores: 8 (f32), 4 + 4 (bf16 hi / lo planes) or
// 2 + 2 (SwiGLU halves the width) store instructions per block instead of 32 / 64 scalar ones through a shared-memory
// transpose.  `vec` = the output row pitch and base allow aligned 16-byte accesses and the block lies inside [0, N).
__device__ __forceinline__ void tc_epilogue_block(const GemmEpilogue &e, const float (&v)[32], int row, int n, int N, bool row_ok, bool vec) {
    if (!row_ok) return;
    if (e.mode == QASR_GEMM_F32 || e.mode == QASR_GEMM_RESIDUAL) {
        float *o = e.out_f32 + (size_t)row * e.ldo + n;
        if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 r = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                if (e.bias) { const float4 bb = __ldg(reinterpret_cast<const float4 *>(e.bias + n + j)); r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w; }
                if (e.mode == QASR_GEMM_RESIDUAL) { const float4 old = *reinterpret_cast<const float4 *>(o + j); r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w; }
                *reinterpret_cast<float4 *>(o + j) = r;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (n + j < N) {
                    float r = v[j] + (e.bias ? e.bias[n + j] : 0.0f);
                    if (e.mode == QASR_GEMM_RESIDUAL) r += o[j];
                    o[j] = r;
                }
        }
    } else if (e.mode == QASR_GEMM_GELU_SPLIT) {
        bf16_t *oh = e.out_hi + (size_t)row * e.ldo + n, *ol = e.out_lo ? e.out_lo + (size_t)row * e.ldo + n : nullptr;
        if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint32_t h[4], l[4];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const float a = gelu_tanh(v[j + 2 * t] + (e.bias ? __ldg(e.bias + n + j + 2 * t) : 0.0f));
                    const float b = gelu_tanh(v[j + 2 * t + 1] + (e.bias ? __ldg(e.bias + n + j + 2 * t + 1) : 0.0f));
                    float ra, rb;
                    h[t] = pack_hi2(a, b, ra, rb);
                    l[t] = pack_bf2(ra, rb);
                }
                st_v4(oh + j, h[0], h[1], h[2], h[3]);
                if (ol) st_v4(ol + j, l[0], l[1], l[2], l[3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (n + j < N) {
                    const float a = gelu_tanh(v[j] + (e.bias ? e.bias[n + j] : 0.0f));
                    __nv_bfloat16 hi, lo;
                    split_bf16(a, hi, lo);
                    oh[j] = __bfloat16_as_ushort(hi);
                    if (ol) ol[j] = __bfloat16_as_ushort(lo);
                }
        }
    } else { // SWIGLU: columns (2j, 2j+1) = (gate_j, up_j), reference qwen_asr_decoder.c:140-152; output column (n >> 1) + j
        bf16_t *oh = e.out_hi + (size_t)row * e.ldo + (n >> 1), *ol = e.out_lo ? e.out_lo + (size_t)row * e.ldo + (n >> 1) : nullptr;
        if (vec) {
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
                uint32_t h[4], l[4];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const float a = silu(v[2 * (j + 2 * t)]) * v[2 * (j + 2 * t) + 1];
                    const float b = silu(v[2 * (j + 2 * t + 1)]) * v[2 * (j + 2 * t + 1) + 1];
                    float ra, rb;
                    h[t] = pack_hi2(a, b, ra, rb);
                    l[t] = pack_bf2(ra, rb);
                }
                st_v4(oh + j, h[0], h[1], h[2], h[3]);
                if (ol) st_v4(ol + j, l[0], l[1], l[2], l[3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (n + 2 * j + 1 < N) {
                    const float a = silu(v[2 * j]) * v[2 * j + 1];
                    __nv_bfloat16 hi, lo;
                    split_bf16(a, hi, lo);
                    oh[j] = __bfloat16_as_ushort(hi);
                    if (ol) ol[j] = __bfloat16_as_ushort(lo);
                }
        }
    }
}

// Persistent kernel: one CTA per SM walks the tile list; the accumulator is double-buffered in TMEM (2 x BN columns), so
// the epilogue of tile i (warps 2-5) overlaps the mainloop of tile i+1 (warps 0-1) and the operand ring never drains
// between tiles.
//
// GATHER = true is the implicit-GEMM form of the conv stem (stages 2 and 3, reference qwen_conv2d + im2col,
// qwen_asr_kernels.c:566-590,643-685): the A operand is not a matrix in memory.  Row `pos` of the GEMM is an output
// position (chunk, ow, oh), its K = 9 x 480 columns are the 3 x 3 input patch (tap-major, channel-minor) of the
// position-major activation act_in[(off_in[c] + iw * Hin + ih)][480]; four extra warps (thread = A row) gather every
// 64-column k-block straight into the 128B-swizzled operand stage with 16-byte cp.async (zero-fill for the padding at
// chunk edges and past K), wait for their own copies, fence them towards the async proxy and arrive on the stage's
// full barrier; the weights still arrive by TMA.  Nothing is materialised (round 1 wrote and re-read the 9 x expanded
// patch matrix: 2.5 GB per launch at 8 x 30 s).
struct ConvGather {
    const bf16_t *src_hi, *src_lo;   // activation planes of the previous stage, [positions][480]
    const int *w0s, *off_in, *off_out;
    int n_chunks, Hin;               // Hin = 64 (stage 2) or 32 (stage 3); Hout = Hin / 2
};
#define TC_GATHER_THREADS 128

template <int BN, bool GATHER = false, int STAGES = (BN > 128 ? 3 : TC_STAGES)>
__global__ void __launch_bounds__(TC_THREADS + (GATHER ? TC_GATHER_THREADS : 0), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB, const TcParams p, const ConvGather cg) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stage][A_hi 16K | A_lo 16K | B BN*128] then barriers
    constexpr uint32_t A_BYTES = TC_BM * TC_BK * 2;
    constexpr uint32_t B_BYTES = BN * TC_BK * 2;
    constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + B_BYTES;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full_bar = empty_bar + STAGES;   // [2] accumulator stage complete (MMA -> epilogue)
    uint64_t *tmem_empty_bar = tmem_full_bar + 2;   // [2] accumulator stage drained (4 epilogue warps -> MMA)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (p.K + TC_BK - 1) / TC_BK;
    const int total_tiles = p.tiles_m * p.tiles_n;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        if (p.nsplit == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], GATHER ? 1 + TC_GATHER_THREADS : 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], (blockDim.x >> 5) - 2 - (GATHER ? TC_GATHER_THREADS / 32 : 0)); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) { // TMEM allocation: two accumulator stages of BN f32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BN))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();    // barriers, tensor-map prefetch and the TMEM allocation above overlapped the previous grid

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const uint32_t tx = GATHER ? B_BYTES : (p.nsplit == 2 ? 2 * A_BYTES : A_BYTES) + B_BYTES;
            uint32_t it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int m0 = (t / p.tiles_n) * TC_BM, n0 = (t % p.tiles_n) * BN;
                for (int kb = 0; kb < num_kb; kb++, it++) {
                    const int s = it % STAGES;
                    mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                    uint8_t *st = smem + s * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], tx);
                    if (!GATHER) {
                        tma_load_2d(st, &tmA_hi, &full_bar[s], kb * TC_BK, m0);
                        if (p.nsplit == 2) tma_load_2d(st + A_BYTES, &tmA_lo, &full_bar[s], kb * TC_BK, m0);
                    }
                    tma_load_2d(st + 2 * A_BYTES, &tmB, &full_bar[s], kb * TC_BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (single elected lane) =====
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1,
            // B=bf16 [10,13)=1, A/B K-major [15],[16]=0, N>>3 [17,23), M>>4 [24,29)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(TC_BM >> 4) << 24);
            uint32_t it = 0, lt = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, lt++) {
                const uint32_t as = lt & 1;
                mbar_wait(&tmem_empty_bar[as], ((lt >> 1) & 1) ^ 1); // the epilogue has drained this accumulator stage
                tc_fence_after();
                const uint32_t tacc = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; kb++, it++) {
                    const int s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES);
                    const uint64_t dah = make_sw128_desc(a_hi);
                    const uint64_t dal = make_sw128_desc(a_hi + A_BYTES);
                    const uint64_t db = make_sw128_desc(a_hi + 2 * A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; k++) // +32 bytes (16 bf16) along K inside the swizzle atom
                        tc_mma_bf16(tacc, dah + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    if (p.nsplit == 2) {
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; k++)
                            tc_mma_bf16(tacc, dal + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
                    }
                    tc_commit(&empty_bar[s]); // frees the ring slot once the MMAs above have read it
                }
                tc_commit(&tmem_full_bar[as]); // accumulator of this tile complete
            }
        }
    } else if (GATHER && warp >= (int)(blockDim.x >> 5) - TC_GATHER_THREADS / 32) {
        // ===== A-operand gather warps (after the epilogue warps; conv stem): thread = row r of the tile =====
        const int r = threadIdx.x - ((int)blockDim.x - TC_GATHER_THREADS);
        const int Hin = cg.Hin, Hout = Hin >> 1;
        const uint32_t row_sm = (uint32_t)r * 128u, sw = (uint32_t)(r & 7);
        // A gather thread publishes k-block i as soon as its copies have landed and never blocks on a ring slot with more than
        // DEPTH k-blocks unpublished: with S stages the MMA warp consumes block i while i+1 is ready and i+2 .. i+S-1 are in flight.
        constexpr int DEPTH = STAGES - 2;
        uint32_t it = 0, arrived = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int pos = (t / p.tiles_n) * TC_BM + r;
            int oh = 0, ow = 0, in0 = 0, Win = 0; // Win = 0 (row past M): every tap is padding
            if (pos < p.M) {
                int lo = 0, hi = cg.n_chunks - 1; // chunk containing pos
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (cg.off_out[mid] <= pos) lo = mid; else hi = mid - 1;
                }
                const int w0 = cg.w0s[lo], w1 = (w0 - 1) / 2 + 1, w2 = (w1 - 1) / 2 + 1;
                Win = Hin == 64 ? w1 : w2;
                const int local = pos - cg.off_out[lo];
                ow = local / Hout; oh = local - ow * Hout; in0 = cg.off_in[lo];
            }
            // element offset of the 480-channel vector of a tap, -1 = zero padding (chunk edge, frequency edge, tap >= 9)
            auto tap_base = [&](int tap) {
                const int ki = tap / 3, ih = 2 * oh - 1 + ki, iw = 2 * ow - 1 + (tap - 3 * ki);
                return (tap < 9 && ih >= 0 && ih < Hin && iw >= 0 && iw < Win) ? (in0 + iw * Hin + ih) * 480 : -1;
            };
            for (int kb = 0; kb < num_kb; kb++, it++) {
                const int s = it % STAGES;
                mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                const uint32_t st = smem_u32(smem + s * STAGE_BYTES) + row_sm;
                const int k0 = kb * TC_BK, tapA = k0 / 480, cA = k0 - tapA * 480; // a k-block spans at most two taps (480 = 7.5 x 64)
                const int bA = tap_base(tapA), bB = tap_base(tapA + 1);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int ch = cA + 8 * j;                      // channel inside tap A, or 480 + channel inside tap B
                    const int b = ch < 480 ? bA : bB, c = ch < 480 ? ch : ch - 480;
                    const bool ok = b >= 0 && k0 + 8 * j < p.K;
                    const size_t off = ok ? (size_t)b + c : 0;
                    const uint32_t dst = st + ((uint32_t)(j ^ sw) << 4);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(cg.src_hi + off), "r"(ok ? 16 : 0) : "memory");
                    if (p.nsplit == 2)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + A_BYTES), "l"(cg.src_lo + off), "r"(ok ? 16 : 0) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (it >= DEPTH) { // the copies of k-block it - DEPTH have landed: publish them to the async proxy (tcgen05 reads)
                    asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH) : "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(&full_bar[arrived % STAGES]);
                    arrived++;
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (; arrived < it; arrived++) mbar_arrive(&full_bar[arrived % STAGES]);
    } else {
        // ===== epilogue warps: TMEM lane quarter q = rows m0 + 32 q .. + 31 (thread = row), column part (warp - 2) / 4 of the tile =====
        const int q = warp & 3, cpart = (warp - 2) >> 2;
        const int cwidth = BN / ((((int)(blockDim.x >> 5) - 2 - (GATHER ? TC_GATHER_THREADS / 32 : 0))) >> 2);
        const GemmEpilogue &e = p.epi;
        const int width = e.mode == QASR_GEMM_SWIGLU_SPLIT ? 2 : 1;
        const bool aligned = (e.ldo % 8 == 0) && ((e.mode == QASR_GEMM_F32 || e.mode == QASR_GEMM_RESIDUAL)
                                 ? (reinterpret_cast<uintptr_t>(e.out_f32) & 15) == 0 && (!e.bias || (reinterpret_cast<uintptr_t>(e.bias) & 15) == 0)
                                 : (reinterpret_cast<uintptr_t>(e.out_hi) & 15) == 0 && (!e.out_lo || (reinterpret_cast<uintptr_t>(e.out_lo) & 15) == 0));
        uint32_t lt = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, lt++) {
            const int m0 = (t / p.tiles_n) * TC_BM, n0 = (t % p.tiles_n) * BN;
            const uint32_t as = lt & 1;
            mbar_wait(&tmem_full_bar[as], (lt >> 1) & 1);
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
            const bool row_ok = row < p.M;
#pragma unroll 1
            for (int c0 = cpart * cwidth; c0 < (cpart + 1) * cwidth; c0 += 32) {
                const int n = n0 + c0;
                if (n >= p.N) break;
                uint32_t r[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + (uint32_t)c0, r);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
                tc_epilogue_block(e, v, row, n, p.N, row_ok, aligned && n + 32 <= p.N && (width == 1 || (p.N & 1) == 0));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)) : "memory");
    }
}


// ------------------------------------------------------------------ 2-CTA variant of the persistent kernel (cta_group::2)
// A CTA pair (cluster (2,1,1) = the two SMs of a TPC) owns a 256 x 256 output tile: CTA r loads ITS 128 rows of A (hi and lo
// planes) and ITS half of the 256 W rows (128 x 64 per k-block), and the leader issues ONE tcgen05.mma.cta_group::2 (M = 256,
// N = 256, K = 16) per k-step and plane; each SM's tensor core reads its own A rows and both halves of W.  Per CTA and
// k-block that is 48 KB through L2 -> shared memory instead of 64 KB (W is fetched once per pair) and half the W operand
// reads from shared memory, so the ring holds 4 stages instead of 3.  Barriers: TMA completions of BOTH CTAs land on the
// leader's full barrier (the follower passes the leader's barrier address); tcgen05.commit multicasts "slot free" /
// "accumulator complete" to the same barrier in both CTAs; the follower's epilogue warps signal "accumulator drained" on
// the leader's barrier through its cluster address.
#define TC2_STAGES 4
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void *dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint64_t *bar) { // arrives on `bar` (same offset) in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int BN = 256;                        // pair tile: 256 (M) x 256 (N)
    constexpr uint32_t A_BYTES = TC_BM * TC_BK * 2; // this CTA's 128 rows of one A plane
    constexpr uint32_t B_BYTES = 128 * TC_BK * 2;   // this CTA's half of the W rows
    constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + B_BYTES;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + TC2_STAGES * STAGE_BYTES);
    uint64_t *empty_bar = full_bar + TC2_STAGES;
    uint64_t *tmem_full_bar = empty_bar + TC2_STAGES;
    uint64_t *tmem_empty_bar = tmem_full_bar + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int num_kb = (p.K + TC_BK - 1) / TC_BK;
    const int total_tiles = p.tiles_m * p.tiles_n; // tiles_m counts 256-row pair tiles
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        if (p.nsplit == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
        for (int s = 0; s < TC2_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 2 * ((blockDim.x >> 5) - 2)); } // epilogue warps of both CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) { // both CTAs of the pair take part in the allocation (same columns in both tensor memories)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all(); // barriers of both CTAs are initialised before anything can signal them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own A rows, own half of the W rows; completions go to the leader's full barrier =====
        if (lane == 0) {
            const uint32_t tx = 2u * ((p.nsplit == 2 ? 2 * A_BYTES : A_BYTES) + B_BYTES); // bytes of BOTH CTAs per stage
            uint32_t it = 0;
            for (int t = pair; t < total_tiles; t += n_pairs) {
                const int m0 = (t / p.tiles_n) * 256 + (int)rank * 128, n0 = (t % p.tiles_n) * BN + (int)rank * 128;
                for (int kb = 0; kb < num_kb; kb++, it++) {
                    const int s = it % TC2_STAGES;
                    mbar_wait(&empty_bar[s], ((it / TC2_STAGES) & 1) ^ 1);
                    uint8_t *st = smem + s * STAGE_BYTES;
                    if (leader) mbar_expect_tx(&full_bar[s], tx);
                    const uint32_t lbar = mapa_u32(smem_u32(&full_bar[s]), 0);
                    tma_load_2d_2sm(st, &tmA_hi, lbar, kb * TC_BK, m0);
                    if (p.nsplit == 2) tma_load_2d_2sm(st + A_BYTES, &tmA_lo, lbar, kb * TC_BK, m0);
                    tma_load_2d_2sm(st + 2 * A_BYTES, &tmB, lbar, kb * TC_BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one elected lane of the LEADER CTA drives both tensor cores =====
        if (leader && lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            uint32_t it = 0, lt = 0;
            for (int t = pair; t < total_tiles; t += n_pairs, lt++) {
                const uint32_t as = lt & 1;
                mbar_wait(&tmem_empty_bar[as], ((lt >> 1) & 1) ^ 1); // both CTAs' epilogues have drained this accumulator stage
                tc_fence_after();
                const uint32_t tacc = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; kb++, it++) {
                    const int s = it % TC2_STAGES;
                    mbar_wait(&full_bar[s], (it / TC2_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES);
                    const uint64_t dah = make_sw128_desc(a_hi);
                    const uint64_t dal = make_sw128_desc(a_hi + A_BYTES);
                    const uint64_t db = make_sw128_desc(a_hi + 2 * A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; k++)
                        tc_mma_bf16_2sm(tacc, dah + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    if (p.nsplit == 2) {
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; k++)
                            tc_mma_bf16_2sm(tacc, dal + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
                    }
                    tc_commit_2sm(&empty_bar[s]); // the slot is free in both CTAs once these MMAs have read it
                }
                tc_commit_2sm(&tmem_full_bar[as]);
            }
        }
    } else {
        // ===== epilogue warps of both CTAs: this CTA's 128 rows of the pair tile, column part (warp - 2) / 4 =====
        const int q = warp & 3, cpart = (warp - 2) >> 2;
        const int cwidth = BN / (((int)(blockDim.x >> 5) - 2) >> 2);
        const GemmEpilogue &e = p.epi;
        const int width = e.mode == QASR_GEMM_SWIGLU_SPLIT ? 2 : 1;
        const bool aligned = (e.ldo % 8 == 0) && ((e.mode == QASR_GEMM_F32 || e.mode == QASR_GEMM_RESIDUAL)
                                 ? (reinterpret_cast<uintptr_t>(e.out_f32) & 15) == 0 && (!e.bias || (reinterpret_cast<uintptr_t>(e.bias) & 15) == 0)
                                 : (reinterpret_cast<uintptr_t>(e.out_hi) & 15) == 0 && (!e.out_lo || (reinterpret_cast<uintptr_t>(e.out_lo) & 15) == 0));
        uint32_t lt = 0;
        for (int t = pair; t < total_tiles; t += n_pairs, lt++) {
            const int m0 = (t / p.tiles_n) * 256 + (int)rank * 128, n0 = (t % p.tiles_n) * BN;
            const uint32_t as = lt & 1;
            mbar_wait(&tmem_full_bar[as], (lt >> 1) & 1);
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
            const bool row_ok = row < p.M;
#pragma unroll 1
            for (int c0 = cpart * cwidth; c0 < (cpart + 1) * cwidth; c0 += 32) {
                const int n = n0 + c0;
                if (n >= p.N) break;
                uint32_t r[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + (uint32_t)c0, r);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
                tc_epilogue_block(e, v, row, n, p.N, row_ok, aligned && n + 32 <= p.N && (width == 1 || (p.N & 1) == 0));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { // "drained" goes to the leader's barrier (the MMA issuer waits there for the epilogue warps of both CTAs)
                const uint32_t lbar = mapa_u32(smem_u32(&tmem_empty_bar[as]), 0);
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(lbar) : "memory");
            }
        }
    }
    tc_fence_before();
    cluster_sync_all(); // nobody leaves (or frees tensor memory) while the peer may still signal its barriers or read its operands
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------ skinny-M variant (M <= 256)
// At the single-utterance shapes of the encoder / prefill (M = 26...157 rows) a GEMM is a weight
// stream: bytes = 2*N*K, flops are irrelevant.  Roles are swapped so that the 128-row MMA operand is
// the WEIGHT tile (each weight byte enters shared memory exactly once) and the activations are the
// N-side operand (MP = M rounded up to 64/128/256 columns of TMEM); split-K spreads one weight
// matrix over ~148 CTAs.  D[n (TMEM lane), m (column)] = sum_k W[n,k] * A[m,k].
// Split-K: the S CTAs of a tile are launched as one thread-block cluster (1, S, 1).  Each leaves its partial tile in its
// own shared memory (the drained pipeline stages), the cluster synchronises, and CTA r adds rows [r*M/S, (r+1)*M/S) of
// all S partials over distributed shared memory in fixed split order and runs the epilogue => deterministic, no
// global scratch, no atomics, and the reduction itself is spread over the S CTAs.  (QASR_GEMM_SK_CLUSTER=0 keeps the
// earlier scheme: partials in a global workspace, the last CTA of a tile - ticket counter - reduces them; same sums.)
// Stores are coalesced along n.
struct SkParams {
    int M, N, K, nsplit, S, kb_per; // S = split-K factor, kb_per = k-blocks per split
    float *ws;                      // [n_tiles][S][MP][128] f32 partials (S > 1)
    unsigned *tickets;              // [n_tiles], zero between launches
    int fuse_planes;                // 1 (MP <= 128, nsplit == 2): ONE N = 2*MP MMA per k-step over the [hi | lo] planes, two accumulators
    int cluster;                    // 1: the S splits of a tile form one thread-block cluster (1, S, 1) and reduce through DSMEM
    GemmEpilogue epi;
};

template <int MP>
__device__ __forceinline__ void sk_epilogue_store(const SkParams &p, int n, int m, float v, int lane) {
    const GemmEpilogue &e = p.epi;
    if (e.mode == QASR_GEMM_SWIGLU_SPLIT) { // rows (2j, 2j+1) of W = (gate_j, up_j): neighbouring lanes
        const float up = __shfl_down_sync(0xffffffffu, v, 1);
        if (!(lane & 1) && n + 1 < p.N && m < p.M) {
            const float r = silu_fast(v) * up;
            __nv_bfloat16 hi, lo;
            split_bf16(r, hi, lo);
            e.out_hi[(size_t)m * e.ldo + (n >> 1)] = __bfloat16_as_ushort(hi);
            if (e.out_lo) e.out_lo[(size_t)m * e.ldo + (n >> 1)] = __bfloat16_as_ushort(lo);
        }
        return;
    }
    if (n >= p.N || m >= p.M) return;
    if (e.bias) v += e.bias[n];
    if (e.mode == QASR_GEMM_F32) {
        e.out_f32[(size_t)m * e.ldo + n] = v;
    } else if (e.mode == QASR_GEMM_RESIDUAL) {
        e.out_f32[(size_t)m * e.ldo + n] += v;
    } else { // GELU_SPLIT
        v = gelu_tanh(v);
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        e.out_hi[(size_t)m * e.ldo + n] = __bfloat16_as_ushort(hi);
        if (e.out_lo) e.out_lo[(size_t)m * e.ldo + n] = __bfloat16_as_ushort(lo);
    }
}

// epilogue of 4 consecutive reduced outputs (row m, columns nb..nb+3) of a split-K tile
template <int MP>
__device__ __forceinline__ void sk_reduced_store(const SkParams &p, int nb, int m, float4 v, const float *s_scale) {
    if (p.epi.in_ssq) { const float sc = s_scale[m]; v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; } // RMSNorm scalar of the input row
    if (p.epi.nx_hi) {
        // residual stream + the operand planes and sum of squares of the RMSNorm that follows (a warp = one row m, its 32 lanes =
        // the 128 columns of this tile; N % 128 == 0 and ldo % 4 == 0 are checked on the host)
        const GemmEpilogue &e = p.epi;
        float4 x = *reinterpret_cast<const float4 *>(e.out_f32 + (size_t)m * e.ldo + nb);
        if (e.bias) { const float4 b4 = *reinterpret_cast<const float4 *>(e.bias + nb); v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w; }
        x.x += v.x; x.y += v.y; x.z += v.z; x.w += v.w;
        *reinterpret_cast<float4 *>(e.out_f32 + (size_t)m * e.ldo + nb) = x;
        const float4 g = __ldg(reinterpret_cast<const float4 *>(e.nx_gamma + nb));
        float r0, r1, r2, r3;
        uint2 hi, lo;
        hi.x = pack_hi2(x.x * g.x, x.y * g.y, r0, r1);
        hi.y = pack_hi2(x.z * g.z, x.w * g.w, r2, r3);
        *reinterpret_cast<uint2 *>(e.nx_hi + (size_t)m * p.N + nb) = hi;
        if (e.nx_lo) {
            lo.x = pack_bf2(r0, r1); lo.y = pack_bf2(r2, r3);
            *reinterpret_cast<uint2 *>(e.nx_lo + (size_t)m * p.N + nb) = lo;
        }
        const float ss = warp_sum(fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, x.w * x.w))));
        if ((threadIdx.x & 31) == 0) e.nx_ssq[(size_t)m * (p.N >> 7) + (nb >> 7)] = ss;
        return;
    }
    if (p.epi.qk_q) {
        // QKV row m, head slot nb / 128 (16 q | 8 k | 8 v): this warp holds the head's 128 values, lane l dims 4l .. 4l+3; dims d and
        // d +- 64 of the split-half RoPE sit in lanes l and l ^ 16
        const GemmEpilogue &e = p.epi;
        const int slot = nb >> 7, lane = threadIdx.x & 31, pos = e.qk_pos0 + m;
        if (slot >= 24) { *reinterpret_cast<float4 *>(e.qk_vc + (size_t)pos * 1024 + (nb - 3072)) = v; return; }
        const float ss = warp_sum(fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w))));
        const float inv = 1.0f / sqrtf(ss / 128.0f + e.qk_eps);
        const float4 w = __ldg(reinterpret_cast<const float4 *>(slot < 16 ? e.qk_qn : e.qk_kn) + lane);
        v.x = v.x * inv * w.x; v.y = v.y * inv * w.y; v.z = v.z * inv * w.z; v.w = v.w * inv * w.w;
        float4 o;
        o.x = __shfl_xor_sync(0xffffffffu, v.x, 16); o.y = __shfl_xor_sync(0xffffffffu, v.y, 16);
        o.z = __shfl_xor_sync(0xffffffffu, v.z, 16); o.w = __shfl_xor_sync(0xffffffffu, v.w, 16);
        const float4 c = __ldg(reinterpret_cast<const float4 *>(e.qk_cos + (size_t)pos * 64) + (lane & 15));
        const float4 sn = __ldg(reinterpret_cast<const float4 *>(e.qk_sin + (size_t)pos * 64) + (lane & 15));
        float4 r;
        if (lane < 16) { r.x = v.x * c.x - o.x * sn.x; r.y = v.y * c.y - o.y * sn.y; r.z = v.z * c.z - o.z * sn.z; r.w = v.w * c.w - o.w * sn.w; }
        else { r.x = v.x * c.x + o.x * sn.x; r.y = v.y * c.y + o.y * sn.y; r.z = v.z * c.z + o.z * sn.z; r.w = v.w * c.w + o.w * sn.w; }
        if (slot < 16) *reinterpret_cast<float4 *>(e.qk_q + (size_t)m * 2048 + nb) = r;
        else *reinterpret_cast<float4 *>(e.qk_kc + (size_t)pos * 1024 + (nb - 2048)) = r;
        return;
    }
    if (p.epi.mode == QASR_GEMM_SWIGLU_SPLIT) { // (gate, up) pairs sit inside the float4
        const GemmEpilogue &e = p.epi;
        const float r[2] = {silu_fast(v.x) * v.y, silu_fast(v.z) * v.w};
#pragma unroll
        for (int j = 0; j < 2; j++)
            if (nb + 2 * j + 1 < p.N) {
                __nv_bfloat16 hi, lo;
                split_bf16(r[j], hi, lo);
                e.out_hi[(size_t)m * e.ldo + ((nb >> 1) + j)] = __bfloat16_as_ushort(hi);
                if (e.out_lo) e.out_lo[(size_t)m * e.ldo + ((nb >> 1) + j)] = __bfloat16_as_ushort(lo);
            }
    } else {
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) sk_epilogue_store<MP>(p, nb + j, m, vv[j], 0);
    }
}

// warps: 0 W producer, 1 MMA, 2 .. 2 + EPW - 1 epilogue, 2 + EPW activation producer.  One epilogue warp per TMEM lane quarter (4 warps,
// one per scheduler) left a CTA 10-20 us in its epilogue - 128 dependent (ld, SiLU, split, store) sequences per thread with nothing to hide
// their latency, and the next CTA cannot start before this one leaves (profiles/r02_batched_path.txt: pre.gu 22.5 us with, 12.3 us without the
// epilogue; contiguous bulk copies of pre-tiled weights, a rotated k order per CTA and a single activation plane all changed nothing).  EPW = 8 (MP = 64) or 16 warps: warp e works on lane quarter e % 4 and the column block e / 4 of the accumulator.
template <int MP> struct SkWarps { static constexpr int EPW = MP <= 32 ? 4 : (MP <= 64 ? 8 : 16), CW = MP / (EPW / 4), THREADS = (3 + EPW) * 32; };
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// tmA: 3-D map {K, M, planes} with box {64, MP, nsplit}: ONE TMA op brings the hi and lo planes of a
// k-block (a TMA op costs ~0.1 us of issue time on the issuing thread, so ops are kept few and the
// weight and activation streams are issued by different warps).
template <int MP, int STAGES>
__global__ void __launch_bounds__(SkWarps<MP>::THREADS, 1)
gemm_tc_skinny_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA, const SkParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr uint32_t W_BYTES = 128 * TC_BK * 2;  // 16 KB
    constexpr uint32_t A_BYTES = MP * TC_BK * 2;   // MP x 128 B
    constexpr uint32_t STAGE_BYTES = W_BYTES + 2 * A_BYTES;
    constexpr int TMEM_COLS = MP <= 128 ? 2 * MP : MP; // room for separate hi / lo accumulators (fuse_planes)
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full_bar = empty_bar + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);
    int *s_last = reinterpret_cast<int *>(tmem_slot + 1);
    float *s_scale = reinterpret_cast<float *>(s_last + 1); // [256] RMSNorm scalars of the input rows (epi.in_ssq)

    pdl_trigger(); // the next grid of the chain may be scheduled right away (it waits for this one where it must)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, split = blockIdx.y, n0 = tile * 128;
    const int total_kb = (p.K + TC_BK - 1) / TC_BK;
    const int kb0 = split * p.kb_per, kb1 = min(total_kb, kb0 + p.kb_per), num_kb = kb1 - kb0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // The weight producer (warp 0) does not wait for the previous grid: weights are constants, so its first STAGES
    // tiles stream from HBM while the previous kernel drains.  Every other warp (activation loads, MMAs that consume
    // them, output stores, split-K scratch) waits, so the grid as a whole still completes after its predecessor.
    if (warp != 0) pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tx = W_BYTES + (p.nsplit == 2 ? 2 * A_BYTES : A_BYTES);
            for (int i = 0; i < num_kb; i++) {
                const int s = i % STAGES;
                mbar_wait(&empty_bar[s], ((i / STAGES) & 1) ^ 1);
                uint8_t *st = smem + s * STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], tx); // covers both producers' bytes; the phase cannot complete before this arrive
                tma_load_2d(st, &tmW, &full_bar[s], (kb0 + i) * TC_BK, n0);
            }
        }
    } else if (warp == 2 + SkWarps<MP>::EPW) {
        if (lane == 0) { // activation producer
            for (int i = 0; i < num_kb; i++) {
                const int s = i % STAGES;
                mbar_wait(&empty_bar[s], ((i / STAGES) & 1) ^ 1);
                tma_load_3d(smem + s * STAGE_BYTES + W_BYTES, &tmA, &full_bar[s], (kb0 + i) * TC_BK, 0, 0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // M (instruction) = 128 weight rows, N (instruction) = MP activation rows
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * MP) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int i = 0; i < num_kb; i++) {
                const int s = i % STAGES;
                mbar_wait(&full_bar[s], (i / STAGES) & 1);
                tc_fence_after();
                const uint32_t w_addr = smem_u32(smem + s * STAGE_BYTES);
                const uint64_t dw = make_sw128_desc(w_addr);
                const uint64_t dah = make_sw128_desc(w_addr + W_BYTES);
                const uint64_t dal = make_sw128_desc(w_addr + W_BYTES + A_BYTES);
                if (MP <= 128 && p.fuse_planes) {
                    // the lo plane follows the hi plane in the stage (MP x 128 B each, same swizzle atoms): together they
                    // are one K-major operand of 2*MP rows, so one MMA reads the weight tile once for both planes;
                    // columns [0, MP) accumulate W.hi, columns [MP, 2 MP) W.lo, added in the epilogue
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; k++)
                        tc_mma_bf16(tmem_base, dw + (uint64_t)(k * 2), dah + (uint64_t)(k * 2), idesc2, (i | k) != 0);
                } else {
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; k++)
                        tc_mma_bf16(tmem_base, dw + (uint64_t)(k * 2), dah + (uint64_t)(k * 2), idesc, (i | k) != 0);
                    if (p.nsplit == 2) {
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; k++)
                            tc_mma_bf16(tmem_base, dw + (uint64_t)(k * 2), dal + (uint64_t)(k * 2), idesc, 1u);
                    }
                }
                tc_commit(&empty_bar[s]);
            }
            tc_commit(tmem_full_bar);
        }
    } else if (warp < 2 + SkWarps<MP>::EPW) {
        // ===== epilogue warps: thread <-> weight row n (TMEM lane quarter warp % 4), activation rows m of this warp's column block
        constexpr int CW = SkWarps<MP>::CW;
        if (p.epi.in_ssq) { // row scalars of the fused RMSNorm, computed while the mainloop runs (reference qwen_asr_kernels.c:801-860)
            for (int m = (int)threadIdx.x - 64; m < MP; m += SkWarps<MP>::EPW * 32) {
                float t = 0.0f;
                if (m < p.M)
                    for (int i = 0; i < p.epi.in_tiles; i++) t += p.epi.in_ssq[(size_t)m * p.epi.in_tiles + i];
                s_scale[m] = m < p.M ? 1.0f / sqrtf(t / (float)p.K + p.epi.in_eps) : 0.0f;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(SkWarps<MP>::EPW * 32) : "memory");
        }
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int q = warp & 3, nl = q * 32 + lane, n = n0 + nl, cb = (warp - 2) >> 2;
        float *wsp = p.cluster ? reinterpret_cast<float *>(smem) + nl // every stage has been consumed: the MMAs that read them are complete
                               : p.ws + ((size_t)(tile * p.S + split) * MP) * 128 + nl;
#pragma unroll 1
        for (int c0 = cb * CW; c0 < (cb + 1) * CW; c0 += 32) {
            if (c0 >= p.M) break; // columns beyond M hold products with zero-filled rows
            uint32_t r[32];
            tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            if (MP <= 128 && p.fuse_planes) {
                uint32_t r2[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(MP + c0), r2);
#pragma unroll
                for (int j = 0; j < 32; j++) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
            }
            if (p.S == 1) {
                if (p.epi.in_ssq) {
#pragma unroll
                    for (int j = 0; j < 32; j++) r[j] = __float_as_uint(__uint_as_float(r[j]) * s_scale[c0 + j]);
                }
#pragma unroll
                for (int j = 0; j < 32; j++) sk_epilogue_store<MP>(p, n, c0 + j, __uint_as_float(r[j]), lane);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++)
                    if (c0 + j < p.M) wsp[(size_t)(c0 + j) * 128] = __uint_as_float(r[j]);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
    if (p.S == 1) return;
    if (p.cluster) {
        // ---- split-K over the cluster: partial tiles sit in the S shared memories; this CTA owns a band of rows
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        const int mper = (p.M + p.S - 1) / p.S, mlo = split * mper, mhi = min(p.M, mlo + mper);
        uint32_t peer[8];
#pragma unroll
        for (int sp = 0; sp < 8; sp++)
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer[sp]) : "r"(smem_u32(smem)), "r"(sp < p.S ? sp : 0));
        for (int idx = mlo * 32 + threadIdx.x; idx < mhi * 32; idx += SkWarps<MP>::THREADS) {
            const int m = idx >> 5, c4 = idx & 31;
            const uint32_t off = (uint32_t)(m * 128 + c4 * 4) * 4u;
            float4 t[8];
#pragma unroll
            for (int sp = 0; sp < 8; sp++)
                if (sp < p.S)
                    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(t[sp].x), "=f"(t[sp].y), "=f"(t[sp].z), "=f"(t[sp].w) : "r"(peer[sp] + off) : "memory");
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int sp = 0; sp < 8; sp++)
                if (sp < p.S) { v.x += t[sp].x; v.y += t[sp].y; v.z += t[sp].z; v.w += t[sp].w; }
            sk_reduced_store<MP>(p, n0 + c4 * 4, m, v, s_scale);
        }
        // nobody leaves while a peer may still read its shared memory
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        return;
    }
    // ---- split-K through global scratch: the last CTA to finish this tile reduces the partials in fixed split order
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned old = atomicAdd(&p.tickets[tile], 1u);
        *s_last = (old == (unsigned)(p.S - 1));
        if (*s_last) p.tickets[tile] = 0; // ready for the next launch
    }
    __syncthreads();
    if (!*s_last) return;
    __threadfence();
    const float *wst = p.ws + ((size_t)tile * p.S * MP) * 128;
    // 4 output float4s per thread per pass, all S x 4 loads issued before the first add (one L2 round trip per pass)
    for (int base = 0; base < p.M * 32; base += SkWarps<MP>::THREADS * 4) {
        float4 acc4[4];
#pragma unroll
        for (int u = 0; u < 4; u++) acc4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int sp = 0; sp < p.S; sp++) {
            float4 t[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int idx = base + u * SkWarps<MP>::THREADS + threadIdx.x;
                t[u] = idx < p.M * 32 ? __ldcg(reinterpret_cast<const float4 *>(wst + ((size_t)sp * MP + (idx >> 5)) * 128) + (idx & 31))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) { acc4[u].x += t[u].x; acc4[u].y += t[u].y; acc4[u].z += t[u].z; acc4[u].w += t[u].w; }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
        const int idx = base + u * SkWarps<MP>::THREADS + threadIdx.x;
        if (idx >= p.M * 32) continue;
        sk_reduced_store<MP>(p, n0 + (idx & 31) * 4, idx >> 5, acc4[u], s_scale);
        }
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

template <int BN>
static constexpr size_t tc_smem_bytes() {
    return (size_t)(BN > 128 ? 3 : TC_STAGES) * (2 * TC_BM * TC_BK * 2 + BN * TC_BK * 2) + 256 + 1024;
}

static constexpr size_t tc2_smem_bytes() { return (size_t)TC2_STAGES * (2 * TC_BM * TC_BK * 2 + 128 * TC_BK * 2) + 256 + 1024; }

template <int MP, int STAGES>
static constexpr size_t sk_smem_bytes() {
    return (size_t)STAGES * (128 * TC_BK * 2 + 2 * MP * TC_BK * 2) + 256 + 1024 + 1024; // stages | barriers | row scalars | alignment
}

struct SkScratch { float *ws = nullptr; size_t ws_bytes = 0; unsigned *tickets = nullptr; };
static SkScratch g_sk[16]; // per device

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the (function, device) pair: the opt-in is repeated for every
// device a context is created on (bit per device), the driver entry point is resolved once per process.
static unsigned g_tc_attr_dev = 0;
int gemm_tc_init(void) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (g_encode && (g_tc_attr_dev >> (dev & 31) & 1u)) return 0;
    if (!g_encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) {
            snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled unavailable: %s", cudaGetErrorString(e));
            return -1;
        }
        g_encode = (PFN_encodeTiled)fn;
    }
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes<128>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes<256>());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes<64>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc2_smem_bytes());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes<128>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes<256>());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_tc_skinny_kernel<32, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sk_smem_bytes<32, 8>());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_tc_skinny_kernel<64, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sk_smem_bytes<64, 6>());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_tc_skinny_kernel<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sk_smem_bytes<128, 4>());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_tc_skinny_kernel<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sk_smem_bytes<256, 2>());
    if (e != cudaSuccess) {
        snprintf(g_tc_err, sizeof g_tc_err, "cudaFuncSetAttribute(gemm_tc): %s", cudaGetErrorString(e));
        return -1;
    }
    g_tc_attr_dev |= 1u << (dev & 31);
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
// Allocate the per-device split-K scratch up front so launches never allocate or synchronise
// (they may run inside a CUDA-graph stream capture).
int gemm_tc_prepare(void) {
    if (gemm_tc_init() != 0) return -1;
    int dev = 0;
    cudaGetDevice(&dev);
    SkScratch &sc = g_sk[dev & 15];
    if (sc.tickets) return 0;
    const size_t ws_bytes = (size_t)32 << 20;
    if (cudaMalloc(&sc.tickets, 4096 * sizeof(unsigned)) != cudaSuccess || cudaMemset(sc.tickets, 0, 4096 * sizeof(unsigned)) != cudaSuccess ||
        cudaMalloc(&sc.ws, ws_bytes) != cudaSuccess) {
        snprintf(g_tc_err, sizeof g_tc_err, "gemm_tc: scratch allocation failed");
        cudaGetLastError();
        return -1;
    }
    sc.ws_bytes = ws_bytes;
    return 0;
}

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

// Skinny path (M <= 256): activation columns MP, weight tiles of 128 rows, split factor S (<= 8 = the portable cluster
// size, also without clusters so that QASR_GEMM_SK_CLUSTER only changes where the partials travel) and k-blocks per split.
struct SkPlan { int MP, n_tiles, total_kb, S, kb_per; bool cluster_mode; };

static SkPlan sk_plan(int M, int K, int N) {
    static int target_ctas = 0, min_kb = 0, sk_cluster = -1;
    if (!target_ctas) {
        const char *e = getenv("QASR_GEMM_TARGET_CTAS"), *m = getenv("QASR_GEMM_MIN_KB"), *ev = getenv("QASR_GEMM_SK_CLUSTER");
        target_ctas = e && atoi(e) > 0 ? atoi(e) : 74;
        min_kb = m && atoi(m) > 0 ? atoi(m) : 2;
        sk_cluster = !(ev && ev[0] == '0');
    }
    SkPlan pl;
    static int mp32 = -1; // QASR_GEMM_MP32=0: rows <= 32 take the 64-column instantiation as before (A/B runs)
    if (mp32 < 0) { const char *e = getenv("QASR_GEMM_MP32"); mp32 = !(e && e[0] == '0'); }
    // activation columns of the MMA: per k-block a CTA takes in 16 KB of weights + 2 planes x MP x 128 B of activations, and that intake
    // (~84 GB/s per SM) is what paces the mainloop - up to 32 rows (batched decode on an 8-GPU shard, short prompts, stream deltas) use MP = 32
    pl.MP = (M <= 32 && mp32) ? 32 : (M <= 64 ? 64 : (M <= 128 ? 128 : 256));
    pl.n_tiles = (N + 127) / 128;
    pl.total_kb = (K + TC_BK - 1) / TC_BK;
    int S = (target_ctas + pl.n_tiles - 1) / pl.n_tiles; // split K until ~target_ctas CTAs stream weights
    if (S > pl.total_kb / min_kb) S = pl.total_kb / min_kb;
    if (S > 8) S = 8;
    if (S < 1) S = 1;
    pl.kb_per = (pl.total_kb + S - 1) / S;
    pl.S = (pl.total_kb + pl.kb_per - 1) / pl.kb_per; // no empty split
    pl.cluster_mode = sk_cluster != 0;
    return pl;
}
// host-only debug hook (tests/test_host_logic.py): out = {path (0 skinny, 1 large-tile), MP, n_tiles, total_kb, S, kb_per}
extern "C" int qasr_debug_gemm_plan(int M, int K, int N, int *out) {
    if (!out || M <= 0 || K <= 0 || N <= 0) return -1;
    if (M > 256) { out[0] = 1; out[1] = out[2] = out[3] = out[4] = out[5] = 0; return 0; }
    const SkPlan pl = sk_plan(M, K, N);
    out[0] = 0; out[1] = pl.MP; out[2] = pl.n_tiles; out[3] = pl.total_kb; out[4] = pl.S; out[5] = pl.kb_per;
    return 0;
}

static int skinny_max_m_cfg() { // QASR_GEMM_SKINNY_MAX_M: largest M that takes the skinny kernel (experiments; default 256)
    static int v = -1;
    if (v < 0) { const char *ev = getenv("QASR_GEMM_SKINNY_MAX_M"); v = ev ? atoi(ev) : 256; if (v > 256) v = 256; }
    return v;
}
// Fused RMSNorm (GemmEpilogue::nx_* / in_ssq) lives in the skinny kernel: the producer needs the split-K reduction path (a warp = one
// row of the tile) and whole 128-column tiles, the consumer only the skinny kernel.  QASR_GEMM_FUSE_NORM=0 switches it off (A/B runs).
static bool fuse_norm_enabled() {
    static int v = -1;
    if (v < 0) { const char *ev = getenv("QASR_GEMM_FUSE_NORM"); v = !(ev && ev[0] == '0'); }
    return v != 0;
}
bool gemm_tc_can_scale_rows(int M) { return fuse_norm_enabled() && M > 0 && M <= skinny_max_m_cfg(); }
bool gemm_tc_can_fuse_norm(int M, int K, int N) {
    if (!gemm_tc_can_scale_rows(M) || K <= 0 || N <= 0 || (N & 127)) return false;
    return sk_plan(M, K, N).S > 1;
}

bool gemm_tc_can_fuse_qk(int M, int K) { return gemm_tc_can_fuse_norm(M, K, 4096); }

// host-only debug hook (tests/test_host_logic.py): bit 0 the consumer side (row scalars), bit 1 the producer side of a fused RMSNorm
// (RESIDUAL GEMM of this shape), bit 2 q/k-norm + RoPE + KV store in a QKV GEMM with these M, K
extern "C" int qasr_debug_gemm_fusion(int M, int K, int N) {
    return (gemm_tc_can_scale_rows(M) ? 1 : 0) | (gemm_tc_can_fuse_norm(M, K, N) ? 2 : 0) | (gemm_tc_can_fuse_qk(M, K) ? 4 : 0);
}

int launch_gemm_tc(cudaStream_t s, const bf16_t *A_hi, const bf16_t *A_lo, int M, int K, const bf16_t *W, int N,
                   const GemmEpilogue &epi) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    if (gemm_tc_init() != 0) return -1;
    if (epi.qk_q && (epi.mode != QASR_GEMM_F32 || N != 4096 || epi.bias || !epi.qk_kc || !epi.qk_vc || !epi.qk_qn || !epi.qk_kn || !epi.qk_cos ||
                     !epi.qk_sin || ((uintptr_t)epi.qk_q & 15) || ((uintptr_t)epi.qk_kc & 15) || ((uintptr_t)epi.qk_vc & 15) || !gemm_tc_can_fuse_qk(M, K))) {
        snprintf(g_tc_err, sizeof g_tc_err, "gemm_tc: fused q/k norm + RoPE needs the F32 QKV GEMM (N = 4096) on the skinny split-K path (M=%d K=%d N=%d)", M, K, N);
        return -1;
    }
    if (epi.nx_hi && (epi.mode != QASR_GEMM_RESIDUAL || !epi.nx_gamma || !epi.nx_ssq || (epi.ldo & 3) || ((uintptr_t)epi.out_f32 & 15) ||
                      ((uintptr_t)epi.nx_hi & 7) || ((uintptr_t)epi.nx_lo & 7) || !gemm_tc_can_fuse_norm(M, K, N))) {
        snprintf(g_tc_err, sizeof g_tc_err, "gemm_tc: fused norm output needs a RESIDUAL skinny split-K GEMM with N %% 128 == 0 (M=%d K=%d N=%d)", M, K, N);
        return -1;
    }
    if (epi.in_ssq && (!gemm_tc_can_scale_rows(M) || epi.in_tiles <= 0)) {
        snprintf(g_tc_err, sizeof g_tc_err, "gemm_tc: row scalars need the skinny kernel (M=%d)", M);
        return -1;
    }
    if ((K & 7) || ((uintptr_t)A_hi & 15) || ((uintptr_t)W & 15) || (A_lo && ((uintptr_t)A_lo & 15))) {
        snprintf(g_tc_err, sizeof g_tc_err, "gemm_tc: K must be a multiple of 8 and operands 16-byte aligned (K=%d)", K);
        return -1;
    }
    const int skinny_max_m = skinny_max_m_cfg();
    if (M <= skinny_max_m) { // weight-streaming regime: skinny-M kernel with split-K over ~148 CTAs
        int dev = 0;
        cudaGetDevice(&dev);
        SkScratch &sc = g_sk[dev & 15];
        SkPlan pl = sk_plan(M, K, N);
        const int MP = pl.MP, n_tiles = pl.n_tiles, S = pl.S, kb_per = pl.kb_per;
        const bool sk_cluster = pl.cluster_mode;
        const size_t need = (size_t)n_tiles * S * MP * 128 * sizeof(float);
        if (!sk_cluster && (!sc.tickets || (S > 1 && need > sc.ws_bytes))) {
            snprintf(g_tc_err, sizeof g_tc_err, "gemm_tc: split-K scratch missing or too small (%zu B needed): call gemm_tc_prepare()", need);
            return -1;
        }
        if (n_tiles > 4096) { snprintf(g_tc_err, sizeof g_tc_err, "gemm_tc: N too large for the skinny path"); return -1; }
        SkParams sp;
        sp.M = M; sp.N = N; sp.K = K; sp.nsplit = A_lo ? 2 : 1; sp.S = S; sp.kb_per = kb_per;
        sp.ws = sc.ws; sp.tickets = sc.tickets; sp.epi = epi;
        sp.cluster = sk_cluster && S > 1;
        static int sk_fuse = -1;
        if (sk_fuse < 0) { const char *ev = getenv("QASR_GEMM_SK_FUSE"); sk_fuse = !(ev && ev[0] == '0'); }
        sp.fuse_planes = sk_fuse && A_lo && MP <= 128;
        if (A_lo && A_lo != A_hi + (size_t)M * K) {
            snprintf(g_tc_err, sizeof g_tc_err, "gemm_tc: lo plane must follow the hi plane (lo = hi + M*K)");
            return -1;
        }
        CUtensorMap mw, ma;
        if (make_map(&mw, W, N, K, 128) != 0) return -1;
        { // 3-D {K, M, planes}: box {64, MP, planes}
            cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)M, (cuuint64_t)(A_lo ? 2 : 1)};
            cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)M * K * 2};
            cuuint32_t box[3] = {TC_BK, (cuuint32_t)MP, (cuuint32_t)(A_lo ? 2 : 1)};
            cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = g_encode(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void *)A_hi, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled(3d) failed (%d) M=%d K=%d", (int)r, M, K); return -1; }
        }
        dim3 grid(n_tiles, S);
        if (MP == 32) launch_pdl_cluster(gemm_tc_skinny_kernel<32, 8>, grid, SkWarps<32>::THREADS, sk_smem_bytes<32, 8>(), s, sp.cluster ? S : 1, mw, ma, sp);
        else if (MP == 64) launch_pdl_cluster(gemm_tc_skinny_kernel<64, 6>, grid, SkWarps<64>::THREADS, sk_smem_bytes<64, 6>(), s, sp.cluster ? S : 1, mw, ma, sp);
        else if (MP == 128) launch_pdl_cluster(gemm_tc_skinny_kernel<128, 4>, grid, SkWarps<128>::THREADS, sk_smem_bytes<128, 4>(), s, sp.cluster ? S : 1, mw, ma, sp);
        else launch_pdl_cluster(gemm_tc_skinny_kernel<256, 2>, grid, SkWarps<256>::THREADS, sk_smem_bytes<256, 2>(), s, sp.cluster ? S : 1, mw, ma, sp);
        cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) { snprintf(g_tc_err, sizeof g_tc_err, "gemm_tc skinny launch: %s", cudaGetErrorString(le)); return -1; }
        return 0;
    }
    // Narrow tiles when the 128-wide grid would leave most of the 148 SMs idle.
    const int tiles128 = ((M + TC_BM - 1) / TC_BM) * ((N + 127) / 128);
    static int force_bn = -1;
    if (force_bn < 0) { const char *ev = getenv("QASR_GEMM_FORCE_BN"); force_bn = ev ? atoi(ev) : 0; }
    const bool bn64 = force_bn ? force_bn == 64 : tiles128 < 80; // measured crossover (tools/gemm_bench.py medium): 84-96 tiles are faster at 128 wide
    // 128 x 256 tiles (43 -> 65 MACs per byte of L2 -> shared traffic: the 128 x 128 kernel is L2-feed bound with hi/lo
    // operands) once there are enough tiles to fill the SMs for several waves
    static int bn256_min_tiles = -1;
    if (bn256_min_tiles < 0) { const char *ev = getenv("QASR_GEMM_BN256_MIN_TILES"); bn256_min_tiles = ev ? atoi(ev) : 400; }
    const bool bn256 = force_bn ? force_bn == 256 : (!bn64 && ((M + TC_BM - 1) / TC_BM) * ((N + 255) / 256) >= bn256_min_tiles);
    TcParams p;
    p.M = M; p.N = N; p.K = K;
    p.nsplit = A_lo ? 2 : 1;
    p.epi = epi;
    // epilogue warps: 8 where the epilogue is exposed (short mainloop, or so few tiles per CTA that its tail shows), else 4 (see TC_EPI_WARPS)
    static int force_epw = -1;
    if (force_epw < 0) { const char *ev = getenv("QASR_GEMM_EPI_WARPS"); force_epw = ev ? atoi(ev) : 0; }
    const int bn_sel = bn64 ? 64 : (bn256 ? 256 : 128);
    const long long tiles_sel = (long long)((M + TC_BM - 1) / TC_BM) * ((N + bn_sel - 1) / bn_sel);
    const bool wide_epi = force_epw ? force_epw == 8 : (K <= 1024 || tiles_sel <= 3 * 148);
    const int epi_threads = (2 + (wide_epi ? 8 : 4)) * 32, bn64_threads = (2 + 4) * 32; // 64-wide tiles: one 32-column chunk per quarter, nothing to split
    static int use_2cta = -1; // QASR_GEMM_2CTA=0: keep the one-CTA 128 x 256 tiles (A/B runs)
    if (use_2cta < 0) { const char *ev = getenv("QASR_GEMM_2CTA"); use_2cta = !(ev && ev[0] == '0'); }
    if (bn256 && use_2cta && M > 128) { // CTA pairs, 256 x 256 tiles
        p.tiles_m = (M + 255) / 256;
        p.tiles_n = (N + 255) / 256;
        CUtensorMap ma, ml, mb;
        if (make_map(&ma, A_hi, M, K, TC_BM) != 0) return -1;
        if (make_map(&ml, A_lo ? A_lo : A_hi, M, K, TC_BM) != 0) return -1;
        if (make_map(&mb, W, N, K, 128) != 0) return -1;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const long long total = (long long)p.tiles_m * p.tiles_n;
        const long long pairs = total < sms / 2 ? total : sms / 2;
        launch_pdl(gemm_tc2_kernel, dim3((unsigned)(2 * pairs)), epi_threads, tc2_smem_bytes(), s, ma, ml, mb, p);
        cudaError_t e2 = cudaGetLastError();
        if (e2 != cudaSuccess) { snprintf(g_tc_err, sizeof g_tc_err, "gemm_tc 2-CTA launch: %s", cudaGetErrorString(e2)); return -1; }
        return 0;
    }
    const int bn = bn64 ? 64 : (bn256 ? 256 : 128);
    p.tiles_m = (M + TC_BM - 1) / TC_BM;
    p.tiles_n = (N + bn - 1) / bn;
    CUtensorMap ma, ml, mb;
    if (make_map(&ma, A_hi, M, K, TC_BM) != 0) return -1;
    if (make_map(&ml, A_lo ? A_lo : A_hi, M, K, TC_BM) != 0) return -1;
    if (make_map(&mb, W, N, K, bn) != 0) return -1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    { static int sm_count[32] = {}; if (!sm_count[dev & 31]) cudaDeviceGetAttribute(&sm_count[dev & 31], cudaDevAttrMultiProcessorCount, dev); sms = sm_count[dev & 31] > 0 ? sm_count[dev & 31] : 148; }
    const long long total = (long long)p.tiles_m * p.tiles_n;
    dim3 grid((unsigned)(total < sms ? total : sms)); // persistent: one CTA per SM walks the tile list
    const ConvGather none = {};
    if (bn256) launch_pdl(gemm_tc_kernel<256>, grid, epi_threads, tc_smem_bytes<256>(), s, ma, ml, mb, p, none);
    else if (bn64) launch_pdl(gemm_tc_kernel<64>, grid, bn64_threads, tc_smem_bytes<64>(), s, ma, ml, mb, p, none);
    else launch_pdl(gemm_tc_kernel<128>, grid, epi_threads, tc_smem_bytes<128>(), s, ma, ml, mb, p, none);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_tc_err, sizeof g_tc_err, "gemm_tc launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

// Conv stem stage 2 / 3 as an implicit GEMM (see ConvGather): out[pos][oc] = epilogue(sum_k patch(pos)[k] * W[oc][k]),
// M = output positions of all chunks, K = 9 * 480, N = 480.
int launch_conv_gemm_tc(cudaStream_t s, const bf16_t *src_hi, const bf16_t *src_lo, const ConvGeom &g, int stage, const bf16_t *W,
                        const GemmEpilogue &epi) {
    const int M = stage == 2 ? g.total2 : g.total3, K = 4320, N = 480;
    if (M <= 0) return 0;
    if (gemm_tc_init() != 0) return -1;
    if (((uintptr_t)src_hi & 15) || (src_lo && ((uintptr_t)src_lo & 15)) || ((uintptr_t)W & 15)) {
        snprintf(g_tc_err, sizeof g_tc_err, "conv gemm: operands must be 16-byte aligned");
        return -1;
    }
    TcParams p;
    p.M = M; p.N = N; p.K = K;
    p.nsplit = src_lo ? 2 : 1;
    p.epi = epi;
    p.tiles_m = (M + TC_BM - 1) / TC_BM;
    const bool bn256 = (long long)p.tiles_m * 2 >= 296; // two 256-wide tiles cover N = 480; narrower tiles while they would leave SMs idle
    const int bn = bn256 ? 256 : 128;
    p.tiles_n = (N + bn - 1) / bn;
    ConvGather cg;
    cg.src_hi = src_hi; cg.src_lo = src_lo ? src_lo : src_hi;
    cg.w0s = g.d_w0; cg.off_in = stage == 2 ? g.d_off1 : g.d_off2; cg.off_out = stage == 2 ? g.d_off2 : g.d_off3;
    cg.n_chunks = g.n_chunks; cg.Hin = stage == 2 ? 64 : 32;
    CUtensorMap mb;
    if (make_map(&mb, W, N, K, bn) != 0) return -1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long total = (long long)p.tiles_m * p.tiles_n;
    dim3 grid((unsigned)(total < sms ? total : sms));
    if (bn256) launch_pdl(gemm_tc_kernel<256, true>, grid, 192 + TC_GATHER_THREADS, tc_smem_bytes<256>(), s, mb, mb, mb, p, cg);
    else launch_pdl(gemm_tc_kernel<128, true>, grid, 192 + TC_GATHER_THREADS, tc_smem_bytes<128>(), s, mb, mb, mb, p, cg);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(g_tc_err, sizeof g_tc_err, "conv gemm launch: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}

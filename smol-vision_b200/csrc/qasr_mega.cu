// qasr_mega.cu - persistent cooperative decode kernel: the whole greedy loop for a chunk of
// tokens in ONE launch (reference hot loop: qwen_asr.c:788-818 -> qwen_decoder_forward,
// qwen_asr_decoder.c:592-685; kernels qwen_asr_kernels.c:336-373,486-543,801-924,946-1010,
// 1101-1148,1233-1298).
//
// Why: a decode step is 3.44 GB (1.7B) / 1.19 GB (0.6B) of bf16 weights read exactly once, i.e.
// purely HBM-bound, but split over 142 dependent phases.  As separate kernels each phase pays
// launch + first-byte latency with an empty memory pipe (measured 35 % of HBM peak).  Here one CTA
// per SM (16 consumer warps) stays resident and every warp streams its static list of weight
// "units" through a private 3-slot shared-memory ring with TMA tensor copies (mbarrier
// complete_tx).  Weight addresses do not depend on activations, so the rings keep
// filling ACROSS the grid barriers that separate the dependent phases.
//
// Unit = 16 weight rows x 128 columns, fetched as two [16 x 64] cp.async.bulk.tensor.2d boxes with
// the 128-byte swizzle (ldmatrix is bank-conflict free; one elected lane issues 2 TMA ops per 4 KB -
// per-row cp.async.bulk copies cost ~0.1 us of issue time EACH and paced the stream, see
// profiles/).  The dot products run on tensor cores:
//     mma.sync.m16n8k16 (bf16 x bf16 -> f32):  A = the 16x16 weight tile straight from the ring,
//     B = the phase input x as two columns, x_hi = RN(x) and x_lo = RN(x - x_hi),
// so D[:,0] + D[:,1] = W . x to ~2^-17 relative (weights are exact bf16; same hi/lo device as the
// prefill GEMMs).  2 instructions (ldmatrix.x4 + mma) consume 256 weights, versus 17 on the FFMA
// path of round 1 whose consumer, not HBM, paced the stream (profiles/r01_megakernel_*).
// Warp w owns column slices {w, w+16, ...} of every 16-row group of the CTA; the <=16 per-slice
// partial sums of a row are added in fixed order => bitwise reproducible run to run.
//
// Phases per layer (grid barrier after each):
//   QKV  : qkv = Wqkv . rmsnorm(x)                         rows 4096
//   ATTN : per kv head x key split: q/k RMSNorm + RoPE + KV append + online-softmax partials
//   WO   : x += Wo . merge(partials)                       rows H
//   GU   : act = silu(g) * u,  [g;u] = Wgu . rmsnorm(x)    rows 2I (interleaved gate/up)
//   DOWN : x += Wdown . act                                rows H
// then HEAD: per-CTA argmax over its vocab rows of E . rmsnorm(x); barrier; every CTA reduces the
// per-CTA winners (lowest index wins ties, reference qwen_asr_kernels.c:536-541); CTA 0 publishes
// the token and gathers the next input row; barrier.
#include "qasr_common.cuh"
#include "qasr_internal.h"

#include <cuda.h>
#include <stdio.h>
#include <type_traits>

#define MG_WARPS 16
#define MG_THREADS (MG_WARPS * 32)
#define MG_SLOTS 3
#define MG_KS 128                       /* columns per unit */
#define MG_UNIT_BYTES 4096              /* two 128B-swizzled [16 rows x 64 cols] TMA boxes */
#define MG_MAX_K 6144
#define MG_CHUNK_GROUPS 6               /* 16-row groups reduced together (96 rows) */
#define MG_PSTRIDE 17

struct MegaSmem {
    uint8_t ring[MG_WARPS][MG_SLOTS][MG_UNIT_BYTES]; // 196608 B, 1024-byte aligned tiles
    float xs[MG_MAX_K];                              // phase input (f32, natural order)
    float partial[MG_CHUNK_GROUPS * 16][MG_PSTRIDE]; // [row in chunk][slice owner]
    uint64_t bar[MG_WARPS][MG_SLOTS];
    float red[MG_WARPS * 3];
    int redi[MG_WARPS];
    int s_tok;
};

__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, %0;" ::"n"(MG_THREADS) : "memory"); }
__device__ __forceinline__ uint32_t mg_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mg_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mg_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mg_mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(mg_smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// Grid barrier over a monotonically increasing arrival counter (zeroed by the host before every
// launch; all CTAs are co-resident: cooperative launch).  Arrival is a fire-and-forget release
// reduction (cumulative over the CTA's writes ordered by bar.sync); the k-th barrier completes
// when the counter reaches k * nblocks.
__device__ __forceinline__ void grid_barrier(unsigned *count, unsigned &target, unsigned nblocks, int debug) {
    if (debug & 1) { csync(); return; } // timing experiment only: no inter-CTA ordering (wrong results)
    target += nblocks;
    csync();
    if (threadIdx.x == 0) {
        if (debug & 16) __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(count) : "memory");
        unsigned c;
        do {
            asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(c) : "l"(count) : "memory");
        } while ((int)(c - target) < 0);
    }
    csync();
}

// ---- static weight schedule -----------------------------------------------------------------
struct PhaseGeom {
    const bf16_t *W;
    int K;          // columns
    int n_slices;   // K / 128
    int g0, g1;     // 16-row groups [g0, g1) of this CTA
    int wps;        // warps per slice when n_slices < 16 (else 1)
};

__device__ __forceinline__ PhaseGeom phase_geom(const MegaParams &p, int wp, int b, int G) {
    PhaseGeom g;
    unsigned N;
    if (wp < p.n_layers * 4) {
        const MegaLayer &L = p.layers[wp >> 2];
        switch (wp & 3) {
            case 0: g.W = L.wqkv; N = 4096; g.K = p.H; break;
            case 1: g.W = L.wo; N = p.H; g.K = 2048; break;
            case 2: g.W = L.wgu; N = 2 * p.I; g.K = p.H; break;
            default: g.W = L.wdown; N = p.H; g.K = p.I; break;
        }
    } else {
        g.W = p.emb; N = p.V; g.K = p.H;
    }
    g.n_slices = g.K / MG_KS;
    g.wps = g.n_slices < MG_WARPS ? MG_WARPS / g.n_slices : 1;
    const unsigned groups = N >> 4;
    g.g0 = (int)(groups * (unsigned)b / (unsigned)G);
    g.g1 = (int)(groups * (unsigned)(b + 1) / (unsigned)G);
    return g;
}
// slices owned by warp w: n_slices >= 16: {w, w+16, w+32}; else the single slice w % n_slices
__device__ __forceinline__ int warp_nsl(const PhaseGeom &g, int w) {
    return g.n_slices >= MG_WARPS ? (g.n_slices - w + MG_WARPS - 1) / MG_WARPS : 1;
}
__device__ __forceinline__ int warp_slice(const PhaseGeom &g, int w, int si) {
    return g.n_slices >= MG_WARPS ? w + si * MG_WARPS : w % g.n_slices;
}
__device__ __forceinline__ int warp_sub(const PhaseGeom &g, int w) { return g.n_slices >= MG_WARPS ? 0 : w / g.n_slices; }

struct UnitStream { // per-warp issue cursor: (step, weighted phase, row group, slice index)
    int step, wp, grp, si, nsl;
    PhaseGeom g;
};

__device__ __forceinline__ bool mg_better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__device__ __forceinline__ float block_sum(float v, float *red, int tid) {
    v = warp_sum(v);
    csync();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    csync();
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < MG_WARPS; w++) t += red[w];
    return t;
}

// x (global, written by other CTAs in earlier phases -> L2 loads) -> xs, optionally RMS-normalised.
// One pass: every thread keeps its <= 12 elements (and the norm weights) in registers across the
// block reduction, so the critical path is a single L2 round trip.
__device__ __forceinline__ void stage_x(MegaSmem &sm, const float *x, const float *gamma, int K, float eps, int tid) {
    constexpr int PER = MG_MAX_K / MG_THREADS;
    float v[PER], gm[PER];
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int e = tid + i * MG_THREADS;
        v[i] = e < K ? __ldcg(x + e) : 0.0f;
        gm[i] = (gamma && e < K) ? gamma[e] : 1.0f;
    }
    float inv = 1.0f;
    if (gamma) {
        float ss = 0.0f;
#pragma unroll
        for (int i = 0; i < PER; i++) ss = fmaf(v[i], v[i], ss);
        const float tot = block_sum(ss, sm.red, tid);
        inv = 1.0f / sqrtf(tot / (float)K + eps);
    }
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int e = tid + i * MG_THREADS;
        if (e < K) sm.xs[e] = gamma ? v[i] * inv * gm[i] : v[i];
    }
    csync();
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b); // .x = a (low half), .y = b
    return *reinterpret_cast<const uint32_t *>(&v);
}

__global__ void __launch_bounds__(MG_THREADS, 1) decode_mega_kernel(const MegaParams p) {
    extern __shared__ __align__(1024) uint8_t mg_raw[];
    MegaSmem &sm = *reinterpret_cast<MegaSmem *>((reinterpret_cast<uintptr_t>(mg_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3; // mma fragment coordinates
    const int b = blockIdx.x, G = gridDim.x;
    const int n_wp = p.n_layers * 4 + 1;

    if (lane == 0)
        for (int s = 0; s < MG_SLOTS; s++) mg_mbar_init(&sm.bar[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    // ---- per-warp weight stream: the issue cursor runs MG_SLOTS units ahead of consumption
    UnitStream is;
    is.step = 0; is.wp = 0; is.si = 0;
    is.g = phase_geom(p, 0, b, G);
    is.nsl = warp_nsl(is.g, warp);
    is.grp = is.g.g0 + warp_sub(is.g, warp);
    auto stream_seek = [&]() { // move to the next phase that holds a unit for this warp
        while (is.step < p.n_steps && is.grp >= is.g.g1) {
            is.wp++;
            if (is.wp == n_wp) { is.wp = 0; is.step++; }
            if (is.step < p.n_steps) {
                // the TMA unit has to fetch a descriptor before its first copy: warm the NEXT phase's map now
                if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(p.maps + (is.wp + 1 == n_wp ? 0 : is.wp + 1)) : "memory");
                is.g = phase_geom(p, is.wp, b, G);
                is.nsl = warp_nsl(is.g, warp);
                is.grp = is.g.g0 + warp_sub(is.g, warp);
                is.si = 0;
            }
        }
    };
    if (lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(p.maps) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(p.maps + 1) : "memory");
    }
    stream_seek();
    unsigned issued = 0, consumed = 0;
    auto top_up = [&]() {
        while (issued - consumed < MG_SLOTS && is.step < p.n_steps) {
            const int slot = issued % MG_SLOTS;
            if (lane == 0) { // one elected lane: two [16 rows x 64 cols] boxes (128B swizzle) = one 4 KB unit
                const uint32_t bar = mg_smem_u32(&sm.bar[warp][slot]), dst = mg_smem_u32(sm.ring[warp][slot]);
                const CUtensorMap *map = p.maps + is.wp;
                const int c0 = warp_slice(is.g, warp, is.si) * MG_KS, r0 = is.grp * 16;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(MG_UNIT_BYTES) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(r0) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(dst + 2048), "l"(map), "r"(bar), "r"(c0 + 64), "r"(r0) : "memory");
            }
            issued++;
            if (++is.si == is.nsl) { is.si = 0; is.grp += is.g.wps; }
            stream_seek();
        }
    };
    top_up();

    // ---- one weighted phase: y[row] = W[row,:] . xs for the CTA's rows, handed to `epi(row, y)`
    // in chunks of <= 96 rows.  NSL = column slices owned by this warp (1, 2 or 3).
    auto run_phase_n = [&](const PhaseGeom &g, auto nsl_c, auto &&epi, const float *resid) {
        constexpr int NSL = decltype(nsl_c)::value;
        // residual phases: the old x of this thread's row of the FIRST chunk is fetched now, off the critical path
        float res0 = 0.0f;
        if (resid && tid < (min(g.g1, g.g0 + MG_CHUNK_GROUPS) - g.g0) * 16) res0 = __ldcg(resid + g.g0 * 16 + tid);
        // B fragments of x for this warp's slices: lanes with gid 0 carry x_hi, gid 1 carry x_lo
        uint32_t bf[NSL][8][2];
#pragma unroll
        for (int si = 0; si < NSL; si++) {
            const int k0 = warp_slice(g, warp, si) * MG_KS + 2 * tig;
#pragma unroll
            for (int kb = 0; kb < 8; kb++) {
                float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
                if (gid < 2) {
                    const float2 a = *reinterpret_cast<const float2 *>(sm.xs + k0 + kb * 16);
                    const float2 c = *reinterpret_cast<const float2 *>(sm.xs + k0 + kb * 16 + 8);
                    v0 = a.x; v1 = a.y; v2 = c.x; v3 = c.y;
                    if (gid == 1) { // residual after rounding to bf16
                        v0 -= __bfloat162float(__float2bfloat16_rn(v0)); v1 -= __bfloat162float(__float2bfloat16_rn(v1));
                        v2 -= __bfloat162float(__float2bfloat16_rn(v2)); v3 -= __bfloat162float(__float2bfloat16_rn(v3));
                    }
                }
                bf[si][kb][0] = pack_bf16(v0, v1);
                bf[si][kb][1] = pack_bf16(v2, v3);
            }
        }
        const int sub = warp_sub(g, warp);
        const int pidx = g.n_slices >= MG_WARPS ? warp : warp % g.n_slices; // column of partial[][] this warp fills
        const int np = g.n_slices >= MG_WARPS ? MG_WARPS : g.n_slices;     // partial sums per row
        // ldmatrix.x4 lane address inside a tile: matrices (rows 0-7,k 0-7) (rows 8-15,k 0-7) (rows 0-7,k 8-15) (rows 8-15,k 8-15).
        // A box row is 128 bytes = 8 chunks of 16 B; TMA's 128B swizzle stores chunk c of row r at chunk c ^ (r & 7).
        const int lm_row = ((lane >> 3) & 1) * 8 + (lane & 7), lm_hi = lane >> 4;
        const uint32_t lm_rowoff = (uint32_t)(lm_row * 128);
        for (int cg0 = g.g0; cg0 < g.g1; cg0 += MG_CHUNK_GROUPS) {
            const int cg1 = min(cg0 + MG_CHUNK_GROUPS, g.g1);
            // this warp's groups in the chunk: (grp - g.g0) % wps == sub
            int first = cg0;
            { const int r = (first - g.g0) % g.wps; first += (sub - r + g.wps) % g.wps; }
            for (int grp = first; grp < cg1; grp += g.wps) {
                float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                for (int si = 0; si < NSL; si++) {
                    const int slot = consumed % MG_SLOTS;
                    const bool tr = (p.debug & 64) && p.prof && b == p.trace_cta && tid == 0 && consumed < 1300;
                    if (tr) p.prof[2 * p.prof_cap + 3 * consumed] = clock64();
                    mg_mbar_wait(&sm.bar[warp][slot], (consumed / MG_SLOTS) & 1);
                    if (tr) p.prof[2 * p.prof_cap + 3 * consumed + 1] = clock64();
                    const uint32_t tile = mg_smem_u32(sm.ring[warp][slot]) + lm_rowoff;
#pragma unroll
                    for (int kb = 0; kb < 8; kb++) {
                        uint32_t a0, a1, a2, a3;
                        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                                     : "r"(tile + (kb >> 2) * 2048 + (((((kb & 3) << 1) | lm_hi) ^ (lm_row & 7)) << 4)));
                        if (kb & 1)
                            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                         : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3)
                                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bf[si][kb][0]), "r"(bf[si][kb][1]));
                        else
                            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                         : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3)
                                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bf[si][kb][0]), "r"(bf[si][kb][1]));
                    }
                    if (tr) p.prof[2 * p.prof_cap + 3 * consumed + 2] = clock64();
                    __syncwarp();
                    consumed++;
                    top_up(); // the freed slot immediately takes the next unit (possibly of a later phase/token)
                }
                if (tig == 0) { // column 0 = x_hi sums, column 1 = x_lo sums; rows gid and gid+8
                    const int r = (grp - cg0) * 16 + gid;
                    sm.partial[r][pidx] = (c0 + d0) + (c1 + d1);
                    sm.partial[r + 8][pidx] = (c2 + d2) + (c3 + d3);
                }
            }
            csync();
            const int rows = (cg1 - cg0) * 16;
            if (tid < rows) {
                auto rowsum = [&](int r) { // fixed-order sum of the per-slice partials of chunk row r
                    float y = 0.0f;
                    for (int i = 0; i < np; i++) y += sm.partial[r][i];
                    return y;
                };
                epi(cg0 * 16 + tid, tid, rowsum, (resid && cg0 != g.g0) ? __ldcg(resid + cg0 * 16 + tid) : res0);
            }
            csync();
        }
    };
    auto run_phase = [&](const PhaseGeom &g, auto &&epi, const float *resid = nullptr) {
        switch (warp_nsl(g, warp)) {
            case 3: run_phase_n(g, std::integral_constant<int, 3>{}, epi, resid); break;
            case 2: run_phase_n(g, std::integral_constant<int, 2>{}, epi, resid); break;
            default: run_phase_n(g, std::integral_constant<int, 1>{}, epi, resid); break;
        }
    };

    // optional phase profiling (QASR_MEGA_PROF): thread 0 of CTA 0 and of the last CTA stamp clock64
    long long *prof = (p.prof && (b == 0 || b == G - 1) && tid == 0) ? p.prof + (b == 0 ? 0 : p.prof_cap) : nullptr;
    int prof_n = 0;
    auto mark = [&]() { if (prof && prof_n < p.prof_cap) prof[prof_n++] = clock64(); };
    const size_t kvd = 1024;
    const float scale = 0.08838834764831845f; // 1/sqrtf(128)
    int pos = *p.d_pos;
    int step = 0;
    unsigned bar_target = 0;
    bool stop = false;

    for (; step < p.n_steps && !stop; step++) {
        for (int l = 0; l < p.n_layers; l++) {
            const MegaLayer &L = p.layers[l];
            float *kc = p.kv_k + (size_t)l * p.kv_layer_stride, *vc = p.kv_v + (size_t)l * p.kv_layer_stride;
            // ---------------- QKV
            {
                const PhaseGeom g = phase_geom(p, l * 4 + 0, b, G);
                mark();
                stage_x(sm, p.x, L.in_norm, p.H, p.eps, tid);
                mark();
                run_phase(g, [&](int row, int r, auto &&rowsum, float) { p.qkv[row] = rowsum(r); });
                mark();
                mark();
            }
            grid_barrier(p.bar_count, bar_target, G, p.debug);
            mark();
            // ---------------- ATTN (flash-decoding partials); S splits of the pos+1 keys
            const int n_keys = pos + 1;
            int S = (n_keys + 95) / 96;
            S = S < 1 ? 1 : (S > QASR_ATTN_SPLITS ? QASR_ATTN_SPLITS : S);
            if (b < 8 * S) {
                const int h = b / S, split = b % S;
                const int per = (n_keys + S - 1) / S;
                const int k0 = split * per, k1 = min(n_keys, k0 + per);
                const bool owner = (k0 < k1) && (k1 == n_keys);
                float *qs = sm.xs;          // [2][128] roped queries
                float *tmp = sm.xs + 256;   // [3][128]
                float *wacc = sm.xs + 1024; // [MG_WARPS][2][128]
                float *wml = sm.xs + 1024 + MG_WARPS * 256; // [MG_WARPS][4] = m0,l0,m1,l1
                float q0 = 0.f, q1 = 0.f, kk = 0.f;
                if (tid < 128) {
                    q0 = __ldcg(p.qkv + (2 * h) * 128 + tid);
                    q1 = __ldcg(p.qkv + (2 * h + 1) * 128 + tid);
                    kk = owner ? __ldcg(p.qkv + 2048 + h * 128 + tid) : 0.0f;
                }
                { // three sums of squares in one block reduction (only warps 0-3 hold data)
                    const float a = warp_sum(q0 * q0), c2 = warp_sum(q1 * q1), d2 = warp_sum(kk * kk);
                    if (lane == 0 && warp < 4) { sm.red[warp] = a; sm.red[4 + warp] = c2; sm.red[8 + warp] = d2; }
                }
                csync();
                if (tid < 128) {
                    const float s0 = sm.red[0] + sm.red[1] + sm.red[2] + sm.red[3];
                    const float s1 = sm.red[4] + sm.red[5] + sm.red[6] + sm.red[7];
                    const float s2 = sm.red[8] + sm.red[9] + sm.red[10] + sm.red[11];
                    tmp[tid] = q0 * (1.0f / sqrtf(s0 / 128.0f + p.eps)) * L.qn[tid];
                    tmp[128 + tid] = q1 * (1.0f / sqrtf(s1 / 128.0f + p.eps)) * L.qn[tid];
                    tmp[256 + tid] = kk * (1.0f / sqrtf(s2 / 128.0f + p.eps)) * L.kn[tid];
                }
                csync();
                if (tid < 128) {
                    const int d = tid & 63;
                    const float c = p.rope_cos[(size_t)pos * 64 + d], sn = p.rope_sin[(size_t)pos * 64 + d];
                    const int partner = tid < 64 ? tid + 64 : tid - 64;
                    const float sgn = tid < 64 ? -1.0f : 1.0f;
                    qs[tid] = tmp[tid] * c + sgn * tmp[partner] * sn;
                    qs[128 + tid] = tmp[128 + tid] * c + sgn * tmp[128 + partner] * sn;
                    if (owner) {
                        kc[(size_t)pos * kvd + h * 128 + tid] = tmp[256 + tid] * c + sgn * tmp[256 + partner] * sn;
                        vc[(size_t)pos * kvd + h * 128 + tid] = __ldcg(p.qkv + 3072 + h * 128 + tid);
                    }
                }
                csync();
                const float4 qa = *reinterpret_cast<const float4 *>(qs + lane * 4);
                const float4 qb = *reinterpret_cast<const float4 *>(qs + 128 + lane * 4);
                float m0 = -1e30f, l0 = 0.0f, m1 = -1e30f, l1 = 0.0f;
                float4 a0 = make_float4(0, 0, 0, 0), a1 = make_float4(0, 0, 0, 0);
                auto attend = [&](const float4 kr, const float4 vr) {
                    float p0 = qa.x * kr.x + qa.y * kr.y + qa.z * kr.z + qa.w * kr.w;
                    float p1 = qb.x * kr.x + qb.y * kr.y + qb.z * kr.z + qb.w * kr.w;
                    p0 = warp_sum(p0) * scale;
                    p1 = warp_sum(p1) * scale;
                    if (p0 > m0) {
                        const float c = expf(m0 - p0);
                        l0 = l0 * c + 1.0f;
                        a0.x = a0.x * c + vr.x; a0.y = a0.y * c + vr.y; a0.z = a0.z * c + vr.z; a0.w = a0.w * c + vr.w;
                        m0 = p0;
                    } else {
                        const float w = expf(p0 - m0);
                        l0 += w;
                        a0.x += w * vr.x; a0.y += w * vr.y; a0.z += w * vr.z; a0.w += w * vr.w;
                    }
                    if (p1 > m1) {
                        const float c = expf(m1 - p1);
                        l1 = l1 * c + 1.0f;
                        a1.x = a1.x * c + vr.x; a1.y = a1.y * c + vr.y; a1.z = a1.z * c + vr.z; a1.w = a1.w * c + vr.w;
                        m1 = p1;
                    } else {
                        const float w = expf(p1 - m1);
                        l1 += w;
                        a1.x += w * vr.x; a1.y += w * vr.y; a1.z += w * vr.z; a1.w += w * vr.w;
                    }
                };
                for (int j = k0 + warp; j < k1; j += MG_WARPS)
                    attend(*reinterpret_cast<const float4 *>(kc + (size_t)j * kvd + h * 128 + lane * 4),
                           *reinterpret_cast<const float4 *>(vc + (size_t)j * kvd + h * 128 + lane * 4));
                if (lane == 0) { wml[warp * 4 + 0] = m0; wml[warp * 4 + 1] = l0; wml[warp * 4 + 2] = m1; wml[warp * 4 + 3] = l1; }
                *reinterpret_cast<float4 *>(wacc + (warp * 2 + 0) * 128 + lane * 4) = a0;
                *reinterpret_cast<float4 *>(wacc + (warp * 2 + 1) * 128 + lane * 4) = a1;
                csync();
                if (tid < 256) { // merge the warps in fixed order; thread = (head, dim)
                    const int hd = tid >> 7, d = tid & 127;
                    float M = -1e30f;
                    for (int w = 0; w < MG_WARPS; w++) M = fmaxf(M, wml[w * 4 + hd * 2]);
                    float Lsum = 0.0f, A = 0.0f;
                    for (int w = 0; w < MG_WARPS; w++) {
                        const float e = expf(wml[w * 4 + hd * 2] - M);
                        Lsum += wml[w * 4 + hd * 2 + 1] * e;
                        A += wacc[(w * 2 + hd) * 128 + d] * e;
                    }
                    float *pb = p.attn_part + ((size_t)(h * QASR_ATTN_SPLITS + split) * 2 + hd) * QASR_ATTN_PART_STRIDE;
                    pb[d] = A;
                    if (d == 0) { pb[128] = M; pb[129] = Lsum; }
                }
            }
            mark();
            grid_barrier(p.bar_count, bar_target, G, p.debug);
            mark();
            // ---------------- WO: input = merged attention output (every CTA merges the S partials itself)
            {
                const PhaseGeom g = phase_geom(p, l * 4 + 1, b, G);
                // softmax factors per (head, split) first, then one pass of independent coalesced loads
                float *fac = &sm.partial[0][0]; // [16 heads][16 splits] normalised weights (partial[] is free here)
                if (tid < 16) {
                    const int hh = tid >> 1, hd = tid & 1;
                    float m[QASR_ATTN_SPLITS], lv[QASR_ATTN_SPLITS], M = -1e30f;
                    for (int sp = 0; sp < S; sp++) {
                        const float *pb = p.attn_part + ((size_t)(hh * QASR_ATTN_SPLITS + sp) * 2 + hd) * QASR_ATTN_PART_STRIDE;
                        m[sp] = __ldcg(pb + 128);
                        lv[sp] = __ldcg(pb + 129);
                    }
                    for (int sp = 0; sp < S; sp++) M = fmaxf(M, m[sp]);
                    float Lsum = 0.0f;
                    for (int sp = 0; sp < S; sp++) { m[sp] = expf(m[sp] - M); Lsum += lv[sp] * m[sp]; }
                    const float invL = Lsum > 0.0f ? 1.0f / Lsum : 0.0f;
                    for (int sp = 0; sp < S; sp++) fac[tid * QASR_ATTN_SPLITS + sp] = m[sp] * invL;
                }
                csync();
                for (int e = tid; e < 2048; e += MG_THREADS) {
                    const int hq = e >> 7, d = e & 127, hh = hq >> 1, hd = hq & 1;
                    float A = 0.0f;
                    for (int sp = 0; sp < S; sp++)
                        A += __ldcg(p.attn_part + ((size_t)(hh * QASR_ATTN_SPLITS + sp) * 2 + hd) * QASR_ATTN_PART_STRIDE + d) *
                             fac[hq * QASR_ATTN_SPLITS + sp];
                    sm.xs[e] = A;
                }
                csync();
                mark();
                run_phase(g, [&](int row, int r, auto &&rowsum, float xold) { p.x[row] = xold + rowsum(r); }, p.x);
                mark();
                mark();
            }
            grid_barrier(p.bar_count, bar_target, G, p.debug);
            mark();
            // ---------------- GU + SwiGLU: rows (2j, 2j+1) = (gate_j, up_j) are neighbours in a chunk
            {
                const PhaseGeom g = phase_geom(p, l * 4 + 2, b, G);
                stage_x(sm, p.x, L.post_norm, p.H, p.eps, tid);
                mark();
                run_phase(g, [&](int row, int r, auto &&rowsum, float) {
                    if (!(row & 1)) p.act[row >> 1] = silu(rowsum(r)) * rowsum(r + 1);
                });
                mark();
                mark();
            }
            grid_barrier(p.bar_count, bar_target, G, p.debug);
            mark();
            // ---------------- DOWN
            {
                const PhaseGeom g = phase_geom(p, l * 4 + 3, b, G);
                stage_x(sm, p.act, nullptr, p.I, p.eps, tid);
                mark();
                run_phase(g, [&](int row, int r, auto &&rowsum, float xold) { p.x[row] = xold + rowsum(r); }, p.x);
                mark();
                mark();
            }
            grid_barrier(p.bar_count, bar_target, G, p.debug);
            mark();
        }
        // ---------------- HEAD: greedy argmax over this CTA's vocab rows
        {
            const PhaseGeom g = phase_geom(p, p.n_layers * 4, b, G);
            stage_x(sm, p.x, p.final_norm, p.H, p.eps, tid);
            float bv = -1e30f;
            int bi = 0x7fffffff;
            run_phase(g, [&](int row, int r, auto &&rowsum, float) { const float y = rowsum(r); if (mg_better(y, row, bv, bi)) { bv = y; bi = row; } });
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(QASR_FULL, bv, o);
                const int oi = __shfl_xor_sync(QASR_FULL, bi, o);
                if (mg_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { sm.red[warp] = bv; sm.redi[warp] = bi; }
            csync();
            if (tid == 0) {
                for (int w = 1; w < MG_WARPS; w++)
                    if (mg_better(sm.red[w], sm.redi[w], bv, bi)) { bv = sm.red[w]; bi = sm.redi[w]; }
                p.head_val[b] = bv;
                p.head_idx[b] = bi;
            }
        }
        grid_barrier(p.bar_count, bar_target, G, p.debug);
        {
            if (warp == 0) { // every CTA reduces the G winners identically
                float bv = -1e30f;
                int bi = 0x7fffffff;
                for (int i = lane; i < G; i += 32) {
                    const float v = __ldcg(p.head_val + i);
                    const int idx = __ldcg(p.head_idx + i);
                    if (mg_better(v, idx, bv, bi)) { bv = v; bi = idx; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(QASR_FULL, bv, o);
                    const int oi = __shfl_xor_sync(QASR_FULL, bi, o);
                    if (mg_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
                }
                if (lane == 0) sm.s_tok = bi;
            }
            csync();
            const int tok = sm.s_tok;
            pos++;
            if (b == 0) { // publish + gather the next input row (reference qwen_asr.c:412-419,816)
                for (int e = tid; e < p.H; e += MG_THREADS) p.x[e] = __uint_as_float(((uint32_t)p.emb[(size_t)tok * p.H + e]) << 16);
                if (tid == 0) {
                    p.d_tokens[step] = tok;
                    if (p.h_tokens) p.h_tokens[step] = tok;
                }
            }
            stop = (tok == 151643 || tok == 151645); // reference qwen_asr.c:792
        }
        grid_barrier(p.bar_count, bar_target, G, p.debug);
    }
    if (b == 0 && tid == 0) { *p.d_pos = pos; *p.d_step = step; }
    // drain bulk copies that were prefetched past an early stop before the CTA's shared memory goes away
    while (consumed < issued) {
        mg_mbar_wait(&sm.bar[warp][consumed % MG_SLOTS], (consumed / MG_SLOTS) & 1);
        consumed++;
    }
}

static char g_mega_err[256] = "";
const char *mega_error(void) { return g_mega_err; }
static int g_mega_grid = 0;

int mega_init(void) {
    if (g_mega_grid) return 0;
    int dev = 0, sms = 0, coop = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaError_t e = cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MegaSmem) + 1024);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_mega_kernel, MG_THREADS, sizeof(MegaSmem) + 1024);
    if (e != cudaSuccess || !coop || per_sm < 1 || sms < 1) {
        snprintf(g_mega_err, sizeof g_mega_err, "decode megakernel unavailable: %s (coop=%d, blocks/SM=%d, smem=%zu)",
                 cudaGetErrorString(e), coop, per_sm, sizeof(MegaSmem));
        cudaGetLastError();
        return -1;
    }
    g_mega_grid = sms;
    return 0;
}

int launch_decode_mega(cudaStream_t s, const MegaParams &p) {
    if (mega_init() != 0) return -1;
    auto k_ok = [](int K) { return K % MG_KS == 0 && K <= MG_MAX_K && (K / MG_KS <= MG_WARPS ? MG_WARPS % (K / MG_KS) == 0 : K / MG_KS <= 3 * MG_WARPS); };
    if (!k_ok(p.H) || !k_ok(p.I) || p.n_steps > 64 || p.n_layers > 28 || (p.V & 15) || (p.H & 15) || (p.I & 7)) {
        snprintf(g_mega_err, sizeof g_mega_err, "decode megakernel: unsupported dims H=%d I=%d V=%d", p.H, p.I, p.V);
        return -1;
    }
    void *args[] = {(void *)&p};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)decode_mega_kernel, dim3(g_mega_grid), dim3(MG_THREADS), args,
                                                sizeof(MegaSmem) + 1024, s);
    if (e != cudaSuccess) {
        snprintf(g_mega_err, sizeof g_mega_err, "decode megakernel launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

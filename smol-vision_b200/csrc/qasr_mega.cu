// qasr_mega.cu - persistent cooperative decode kernel: the whole greedy loop for a chunk of
// tokens in ONE launch (reference hot loop: qwen_asr.c:788-818 -> qwen_decoder_forward,
// qwen_asr_decoder.c:592-685; kernels qwen_asr_kernels.c:336-373,486-543,801-924,946-1010,
// 1101-1148,1233-1298).
//
// Why: a decode step is 3.44 GB (1.7B) / 1.19 GB (0.6B) of bf16 weights read exactly once, i.e.
// purely HBM-bound, but split over 142 dependent phases.  As separate kernels each phase pays
// launch + first-byte latency with an empty memory pipe (measured 35 % of HBM peak).  Here one CTA
// per SM (12 warps) stays resident; every warp owns a private 4-slot shared-memory ring fed by
// cp.async.bulk (TMA 1-D bulk copy, mbarrier complete_tx) over its static list of weight
// "units" (row x <=2048-column piece, <= 4 KB, contiguous in the checkpoint layout).  Weight
// addresses do not depend on activations, so the rings keep streaming ACROSS the grid barriers
// that separate the dependent phases: up to 192 KB per SM is in flight while a CTA waits.
//
// Phases per layer (grid barrier after each):
//   QKV  : qkv = Wqkv . rmsnorm(x)                         rows 4096
//   ATTN : per kv head x key split: q/k RMSNorm + RoPE + KV append + online-softmax partials
//   WO   : x += Wo . merge(partials)                       rows H
//   GU   : act = silu(g) * u,  [g;u] = Wgu . rmsnorm(x)    rows 2I (interleaved gate/up)
//   DOWN : x += Wdown . act                                rows H
// then HEAD: per-CTA argmax over its vocab rows of E . rmsnorm(x); barrier; every CTA reduces the
// 148 winners (lowest index wins ties, reference qwen_asr_kernels.c:536-541); CTA 0 publishes the
// token and gathers the next input row; barrier.  Every reduction order is fixed => bitwise
// reproducible run to run.
#include "qasr_common.cuh"
#include "qasr_internal.h"

#include <stdio.h>
#include <type_traits>

#define MG_THREADS 384 /* consumer threads */
#define MG_WARPS 12    /* consumer warps */
#define MG_ALL_THREADS (MG_THREADS + 32) /* + 1 producer warp */
#define MG_SLOTS 4
#define MG_SLOT_BYTES 4096
#define MG_MAX_K 6144
#define MG_MAX_UNITS 1100
#ifndef MG_PF_AHEAD
#define MG_PF_AHEAD 0 /* L2 prefetch distance in units per consumer warp; measured slower than none (TMA queue contention) */
#endif

struct MegaSmem {
    uint8_t ring[MG_WARPS][MG_SLOTS][MG_SLOT_BYTES]; // 196608 B, 16-byte aligned pieces of weight rows
    float xs[MG_MAX_K];                              // phase input, lane-interleaved (see stage_x)
    float partial[MG_MAX_UNITS];                     // one dot product per unit
    uint64_t bar[MG_WARPS][MG_SLOTS];   // full: bulk copy landed
    uint64_t empty[MG_WARPS][MG_SLOTS]; // empty: consumer warp released the slot
    volatile int issued[MG_WARPS];
    volatile int limit_step;           // producer issues units of steps < limit_step only
    float red[MG_WARPS * 2];
    int redi[MG_WARPS];
    int s_tok;
};

// barrier over the 12 consumer warps only (the producer warp never joins it)
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
// one elected lane: arm the barrier with the byte count, then start the bulk copy global -> smem
__device__ __forceinline__ void mg_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    // WAR on the slot (generic-proxy reads by the warp, then this async-proxy write) is ordered by the
    // __syncwarp() before the elected lane gets here, as in TMA producer/consumer pipelines
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mg_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(mg_smem_u32(dst)), "l"(src), "r"(bytes), "r"(mg_smem_u32(bar))
                 : "memory");
}

// Grid barrier over a monotonically increasing arrival counter (zeroed by the host before every
// launch; all CTAs are co-resident: cooperative launch).  Arrival is a fire-and-forget release
// reduction; the k-th barrier completes when the counter reaches k * nblocks.
__device__ __forceinline__ void grid_barrier(unsigned *count, unsigned &target, unsigned nblocks, int debug = 0) {
    if (debug & 1) { csync(); return; } // timing experiment only: no inter-CTA ordering (wrong results)
    target += nblocks;
    csync();
    if (threadIdx.x == 0) {
        if (debug & 16) __threadfence(); // not needed: red.release is cumulative over writes ordered by bar.sync
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
    int K, PC, KP;   // columns, pieces per row, columns per piece
    int row0, rows;  // first row / row count of this CTA
    int wpp;         // warps per piece: warp w owns piece w % PC and rows w / PC + j * wpp
};

__device__ __forceinline__ PhaseGeom phase_geom(const MegaParams &p, int wp, int b, int G) {
    PhaseGeom g;
    unsigned N;
    const int n_w = p.n_layers * 4;
    if (wp < n_w) {
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
    g.PC = (g.K + 2047) >> 11;
    g.KP = g.K / g.PC;
    g.wpp = MG_WARPS / g.PC;
    const unsigned groups = N >> 1; // rows are dealt in pairs so SwiGLU gate/up stay together (< 2^17, b < 2^8)
    const unsigned g0 = groups * (unsigned)b / (unsigned)G, g1 = groups * (unsigned)(b + 1) / (unsigned)G;
    g.row0 = 2 * (int)g0;
    g.rows = 2 * (int)(g1 - g0);
    return g;
}

struct UnitStream { // per-warp cursor over (step, weighted phase, j); unit = (row w/PC + j*wpp, piece w%PC)
    int step, wp, j;
    PhaseGeom g;
};

__device__ __forceinline__ bool warp_has_unit(const PhaseGeom &g, int warp, int j) {
    return warp < g.wpp * g.PC && warp / g.PC + j * g.wpp < g.rows;
}

__device__ __forceinline__ void stream_seek(UnitStream &s, const MegaParams &p, int b, int G, int warp, int n_wp) {
    // move to the next (wp, j) that holds a unit for this warp (or step == n_steps)
    while (s.step < p.n_steps && !warp_has_unit(s.g, warp, s.j)) {
        s.j = 0;
        s.wp++;
        if (s.wp == n_wp) { s.wp = 0; s.step++; }
        if (s.step < p.n_steps) s.g = phase_geom(p, s.wp, b, G);
    }
}

// Stage the K-vector of a phase into shared memory in the lane-interleaved order the dot loop
// reads it: element e = 8*(32*i + lane) + 4*half + off  ->  xs[((2*i + half)*32 + lane)*4 + off]
// so each warp-wide float4 read is one conflict-free 512-byte wavefront.
__device__ __forceinline__ int xs_index(int e) {
    const int i = e >> 8, lane = (e >> 3) & 31, half = (e >> 2) & 1, off = e & 3;
    return (((i << 1) + half) * 32 + lane) * 4 + off;
}

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
    float v[MG_MAX_K / MG_THREADS], gm[MG_MAX_K / MG_THREADS];
#pragma unroll
    for (int i = 0; i < MG_MAX_K / MG_THREADS; i++) {
        const int e = tid + i * MG_THREADS;
        v[i] = e < K ? __ldcg(x + e) : 0.0f;
        gm[i] = (gamma && e < K) ? gamma[e] : 1.0f;
    }
    float inv = 1.0f;
    if (gamma) {
        float ss = 0.0f;
#pragma unroll
        for (int i = 0; i < MG_MAX_K / MG_THREADS; i++) ss = fmaf(v[i], v[i], ss);
        const float tot = block_sum(ss, sm.red, tid);
        inv = 1.0f / sqrtf(tot / (float)K + eps);
    }
#pragma unroll
    for (int i = 0; i < MG_MAX_K / MG_THREADS; i++) {
        const int e = tid + i * MG_THREADS;
        if (e < K) sm.xs[xs_index(e)] = gamma ? v[i] * inv * gm[i] : v[i];
    }
    csync();
}

__device__ __forceinline__ bool mg_better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__device__ __forceinline__ bool mg_mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(mg_smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mg_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mg_smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(MG_ALL_THREADS, 1) decode_mega_kernel(const MegaParams p) {
    extern __shared__ __align__(128) uint8_t mg_raw[];
    MegaSmem &sm = *reinterpret_cast<MegaSmem *>(mg_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, G = gridDim.x;
    const int n_wp = p.n_layers * 4 + 1;

    if (warp < MG_WARPS && lane == 0)
        for (int s = 0; s < MG_SLOTS; s++) { mg_mbar_init(&sm.bar[warp][s], 1); mg_mbar_init(&sm.empty[warp][s], 1); }
    if (tid == 0) sm.limit_step = p.n_steps;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    if (warp == MG_WARPS) {
        // ===== producer warp: lane i streams consumer warp i's static unit list into its ring.
        // Issue is decoupled from consumption, so bulk copies for LATER phases keep entering free
        // slots while the consumers sit in a grid barrier or stage activations.
        if (lane < MG_WARPS && !(p.debug & 32)) {
            sm.issued[lane] = 0; // default (coupled) mode: every consumer warp issues its own bulk copies
        } else if (lane < MG_WARPS) {
            UnitStream is;
            is.step = 0; is.wp = 0; is.j = 0;
            is.g = phase_geom(p, 0, b, G);
            stream_seek(is, p, b, G, lane, n_wp);
            // second cursor: L2 prefetch runs MG_PF_AHEAD units ahead of the shared-memory ring, so HBM
            // keeps streaming into the 126 MB L2 while the (192 KB) ring is full during a stall and the
            // ring refills at L2 latency/bandwidth afterwards
            UnitStream pf = is;
            unsigned issued = 0, prefetched = 0;
            while (is.step < sm.limit_step) {
                if (prefetched < issued + MG_PF_AHEAD && pf.step < sm.limit_step) {
                    const int row = pf.g.row0 + lane / pf.g.PC + pf.j * pf.g.wpp, piece = lane % pf.g.PC;
                    const bf16_t *src = pf.g.W + (size_t)row * pf.g.K + (size_t)piece * pf.g.KP;
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)pf.g.KP * 2) : "memory");
                    prefetched++;
                    pf.j++;
                    stream_seek(pf, p, b, G, lane, n_wp);
                }
                const int slot = issued % MG_SLOTS;
                if (!mg_mbar_test(&sm.empty[lane][slot], ((issued / MG_SLOTS) & 1) ^ 1)) { __nanosleep(64); continue; } // ring full
                const int row = is.g.row0 + lane / is.g.PC + is.j * is.g.wpp, piece = lane % is.g.PC;
                const bf16_t *src = is.g.W + (size_t)row * is.g.K + (size_t)piece * is.g.KP;
                mg_bulk_load(sm.ring[lane][slot], src, (uint32_t)is.g.KP * 2, &sm.bar[lane][slot]);
                issued++;
                is.j++;
                stream_seek(is, p, b, G, lane, n_wp);
            }
            sm.issued[lane] = (int)issued;
        }
        __syncthreads(); // joins the consumers' final barrier
        return;
    }

    // ===== consumer warps
    unsigned consumed = 0, issued_c = 0;
    const bool coupled = (p.debug & 32) == 0; // bit 5 selects the dedicated producer warp instead (measured slower)
    UnitStream cs;
    cs.step = 0; cs.wp = 0; cs.j = 0;
    cs.g = phase_geom(p, 0, b, G);
    if (coupled) stream_seek(cs, p, b, G, warp, n_wp);
    auto top_up = [&]() { // coupled mode only: refill the slots this warp has freed
        while (issued_c - consumed < MG_SLOTS && cs.step < p.n_steps) {
            if (lane == 0) {
                const int row = cs.g.row0 + warp / cs.g.PC + cs.j * cs.g.wpp, piece = warp % cs.g.PC;
                const bf16_t *src = cs.g.W + (size_t)row * cs.g.K + (size_t)piece * cs.g.KP;
                const int slot = issued_c % MG_SLOTS;
                mg_bulk_load(sm.ring[warp][slot], src, (uint32_t)cs.g.KP * 2, &sm.bar[warp][slot]);
            }
            issued_c++;
            cs.j++;
            stream_seek(cs, p, b, G, warp, n_wp);
        }
    };
    if (coupled) top_up();

    // Dot products of this warp's units of one weighted phase -> sm.partial[row_local*PC + piece].
    // The warp's slice of the phase input (<= 2048 columns of its piece) is pulled from sm.xs into
    // 64 registers ONCE, so the per-unit shared-memory traffic is just the 4 KB of weights and the
    // consumer runs several times faster than HBM can refill the rings.
    auto run_units_n = [&](const PhaseGeom &g, auto nit_c) {
        constexpr int NIT = decltype(nit_c)::value; // 256-column iterations per unit (KP / 256)
        const int piece = warp % g.PC;
        float4 xa[NIT], xb[NIT];
        {
            const float4 *xv = reinterpret_cast<const float4 *>(sm.xs) + (size_t)piece * (g.KP >> 2);
#pragma unroll
            for (int it = 0; it < NIT; it++) { xa[it] = xv[it * 64 + lane]; xb[it] = xv[it * 64 + 32 + lane]; }
        }
        for (int rl = warp / g.PC; rl < g.rows; rl += g.wpp) {
            const int slot = consumed % MG_SLOTS;
            const bool tr = (p.debug & 64) && p.prof && b == 0 && tid == 0 && consumed < 1300;
            if (tr) p.prof[2 * p.prof_cap + 3 * consumed] = clock64();
            mg_mbar_wait(&sm.bar[warp][slot], (consumed / MG_SLOTS) & 1);
            if (tr) p.prof[2 * p.prof_cap + 3 * consumed + 1] = clock64();
            const uint4 *wv = reinterpret_cast<const uint4 *>(sm.ring[warp][slot]);
            float acc4[4] = {0.0f, 0.0f, 0.0f, 0.0f}; // 4 independent FMA chains (latency, not throughput, bounds a unit)
#pragma unroll
            for (int it = 0; it < NIT; it++) acc4[it & 3] = dot8(wv[it * 32 + lane], xa[it], xb[it], acc4[it & 3]);
            float acc = warp_sum((acc4[0] + acc4[1]) + (acc4[2] + acc4[3]));
            if (lane == 0) sm.partial[rl * g.PC + piece] = acc;
            if (tr) p.prof[2 * p.prof_cap + 3 * consumed + 2] = clock64();
            __syncwarp();
            if (lane == 0 && !coupled) mg_mbar_arrive(&sm.empty[warp][slot]); // slot free: the producer may refill it
            consumed++;
            if (coupled) top_up();
        }
    };
    auto run_units = [&](const PhaseGeom &g) {
        if (warp >= g.wpp * g.PC) return;
        switch (g.KP >> 8) {
            case 8: run_units_n(g, std::integral_constant<int, 8>{}); break;
            case 6: run_units_n(g, std::integral_constant<int, 6>{}); break;
            default: run_units_n(g, std::integral_constant<int, 4>{}); break;
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
                run_units(g);
                csync();
                mark();
                for (int r = tid; r < g.rows; r += MG_THREADS) p.qkv[g.row0 + r] = sm.partial[r];
                mark();
            }
            grid_barrier(p.bar_count, bar_target, G, p.debug);
            mark();
            // ---------------- ATTN (flash-decoding partials); S splits of the pos+1 keys
            const int n_keys = pos + 1;
            int S = (n_keys + 63) / 64;
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
                const float s0 = block_sum(q0 * q0, sm.red, tid);
                const float s1 = block_sum(q1 * q1, sm.red, tid);
                const float s2 = block_sum(kk * kk, sm.red, tid);
                if (tid < 128) {
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
                // merge the S split partials: softmax factors per (head, split) first, then one
                // pass of independent coalesced loads per output element
                float *fac = sm.partial; // [16 heads][S] normalised weights (partial[] is free here)
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
                    sm.xs[xs_index(e)] = A;
                }
                csync();
                mark();
                run_units(g);
                csync();
                mark();
                for (int r = tid; r < g.rows; r += MG_THREADS) p.x[g.row0 + r] = __ldcg(p.x + g.row0 + r) + sm.partial[r];
                mark();
            }
            grid_barrier(p.bar_count, bar_target, G, p.debug);
            mark();
            // ---------------- GU + SwiGLU
            {
                const PhaseGeom g = phase_geom(p, l * 4 + 2, b, G);
                stage_x(sm, p.x, L.post_norm, p.H, p.eps, tid);
                mark();
                run_units(g);
                csync();
                mark();
                for (int r = tid; r < g.rows / 2; r += MG_THREADS)
                    p.act[g.row0 / 2 + r] = silu(sm.partial[2 * r]) * sm.partial[2 * r + 1];
                mark();
            }
            grid_barrier(p.bar_count, bar_target, G, p.debug);
            mark();
            // ---------------- DOWN
            {
                const PhaseGeom g = phase_geom(p, l * 4 + 3, b, G);
                stage_x(sm, p.act, nullptr, p.I, p.eps, tid);
                mark();
                run_units(g);
                csync();
                mark();
                const int rows = g.rows;
                for (int r = tid; r < rows; r += MG_THREADS) {
                    float v = 0.0f;
                    for (int pc = 0; pc < g.PC; pc++) v += sm.partial[r * g.PC + pc];
                    p.x[g.row0 + r] = __ldcg(p.x + g.row0 + r) + v;
                }
                mark();
            }
            grid_barrier(p.bar_count, bar_target, G, p.debug);
            mark();
        }
        // ---------------- HEAD: greedy argmax over this CTA's vocab rows
        {
            const PhaseGeom g = phase_geom(p, p.n_layers * 4, b, G);
            stage_x(sm, p.x, p.final_norm, p.H, p.eps, tid);
            run_units(g);
            csync();
            float bv = -1e30f;
            int bi = 0x7fffffff;
            for (int r = tid; r < g.rows; r += MG_THREADS)
                if (mg_better(sm.partial[r], g.row0 + r, bv, bi)) { bv = sm.partial[r]; bi = g.row0 + r; }
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
            if (stop && tid == 0) sm.limit_step = step + 1;
        }
        grid_barrier(p.bar_count, bar_target, G, p.debug);
    }
    if (b == 0 && tid == 0) { *p.d_pos = pos; *p.d_step = step; }
    // join the producer, then drain bulk copies that were prefetched past an early stop before the
    // CTA's shared memory goes away
    __syncthreads();
    const unsigned issued = coupled ? issued_c : (unsigned)sm.issued[warp];
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
    cudaError_t e = cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MegaSmem));
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_mega_kernel, MG_ALL_THREADS, sizeof(MegaSmem));
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
    auto kp_ok = [](int K) { const int kp = K / ((K + 2047) / 2048); return kp == 1024 || kp == 1536 || kp == 2048; };
    if (p.H > MG_MAX_K || p.I > MG_MAX_K || !kp_ok(p.H) || !kp_ok(p.I) || p.n_steps > 64) {
        snprintf(g_mega_err, sizeof g_mega_err, "decode megakernel: unsupported dims H=%d I=%d", p.H, p.I);
        return -1;
    }
    void *args[] = {(void *)&p};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)decode_mega_kernel, dim3(g_mega_grid), dim3(MG_ALL_THREADS), args,
                                                sizeof(MegaSmem), s);
    if (e != cudaSuccess) {
        snprintf(g_mega_err, sizeof g_mega_err, "decode megakernel launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

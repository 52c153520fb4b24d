// qasr_stream_common.cuh - shared by the two decode kernels (qasr_stream.cu: shared-memory weight ring, up to 4 sequences;
// qasr_stream_r.cu: register-resident weight window, one sequence): the static schedule of the decode weight image,
// the flag-in-data exchange primitives and the B-fragment packing of a phase input.
#pragma once
#include "qasr_common.cuh"
#include "qasr_internal.h"

#include <stdio.h>
#include <stdlib.h>
#include <type_traits>

#define SK_WARPS 16
#define SK_THREADS (SK_WARPS * 32)
#define SK_UNIT 2048          /* 16 rows x 64 cols bf16, A-fragment order: [kb 0..3][lane][a0..a3] */
#define SK_CHUNK_GROUPS 8     /* 16-row groups reduced together (128 rows) */
#define SK_PSTRIDE 17
#define SK_MAX_K 6144
#define SK_MAX_H 2048
#ifndef SK_ATT_MAXS
#define SK_ATT_MAXS 4         /* key splits per (sequence, head); measured: 4 > 8 (the merge at the WO stage grows with it) */
#endif
#ifndef SK_ATT_BATCH
#define SK_ATT_BATCH 2        /* cached keys per warp whose K/V rows are loaded ahead of the q words; a split holds 32 keys per batch */
#endif
#define SK_ATT_STRIDE 132     /* 128 acc + m + l, padded to whole 32-byte sectors: writers of neighbouring blocks never share a sector (2.4 x slower exchange when they do, tools/microbench/ll_exchange2.cu) */
// p.debug bits: 4 = stall-driven L2 prefetch without the evict_last hint, 64 = per-unit trace of warp 0 of CTA p.trace_cta

typedef unsigned long long u64;

// ---- static schedule, shared by the re-tiling kernel and the decode kernel ---------------------
struct SkDims { int L, H, I, V, G; };
__host__ __device__ __forceinline__ void sk_phase_shape(const SkDims &d, int wp, int &N, int &K) {
    if (wp < 4 * d.L) {
        switch (wp & 3) {
            case 0: N = 4096; K = d.H; break;
            case 1: N = d.H; K = 2048; break;
            case 2: N = 2 * d.I; K = d.H; break;
            default: N = d.H; K = d.I; break;
        }
    } else { N = d.V; K = d.H; }
}
__host__ __device__ __forceinline__ int sk_g0(int NG, int b, int G) { return (int)((unsigned)NG * (unsigned)b / (unsigned)G); }
// units per warp of CTA b in phase wp
__host__ __device__ __forceinline__ int sk_phase_units(const SkDims &d, int wp, int b) {
    int N, K;
    sk_phase_shape(d, wp, N, K);
    return (sk_g0(N >> 4, b + 1, d.G) - sk_g0(N >> 4, b, d.G)) * (K >> 10);
}

// ---- device helpers ----------------------------------------------------------------------------
__device__ __forceinline__ void sk_csync() { asm volatile("bar.sync 1, %0;" ::"n"(SK_THREADS) : "memory"); }
__device__ __forceinline__ uint32_t sk_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ll_store(u64 *p, float v, unsigned tag) {
    const u64 w = ((u64)tag << 32) | (u64)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ void ll_store_u32(u64 *p, unsigned v, unsigned tag) {
    const u64 w = ((u64)tag << 32) | (u64)v;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ void ll_load2(const u64 *p, u64 &a, u64 &b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ u64 ll_load1(const u64 *p) {
    u64 a;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    return a;
}
// All poll loops are warp-uniform (the whole warp stays in the loop until every lane has its words), so the
// service functor - the weight-stream top-up, whose cursor state must stay identical across lanes - runs converged.
template <class Svc>
__device__ __forceinline__ float ll_wait1(const u64 *p, unsigned tag, bool active, Svc &&svc) {
    u64 w = active ? ll_load1(p) : (u64)tag << 32;
    for (;;) {
        const bool ok = (unsigned)(w >> 32) == tag;
        if (__all_sync(QASR_FULL, ok)) break;
        svc();
        if (!ok) w = ll_load1(p);
    }
    return __uint_as_float((unsigned)w);
}
// Poll NP pairs of consecutive words per thread (pair p = tid + i*512, valid while p < npairs).
template <int NP, int NT = SK_THREADS, class Svc>
__device__ __forceinline__ void ll_gather_pairs(const u64 *buf, int npairs, unsigned tag, int tid, float (&v)[NP][2], Svc &&svc) {
    u64 w[NP][2];
#pragma unroll
    for (int i = 0; i < NP; i++) {
        const int p = tid + i * NT;
        if (p < npairs) ll_load2(buf + 2 * p, w[i][0], w[i][1]);
        else w[i][0] = w[i][1] = (u64)tag << 32;
    }
    for (;;) {
        bool ok = true;
#pragma unroll
        for (int i = 0; i < NP; i++) ok = ok && (unsigned)(w[i][0] >> 32) == tag && (unsigned)(w[i][1] >> 32) == tag;
        if (__all_sync(QASR_FULL, ok)) break;
        svc();
#pragma unroll
        for (int i = 0; i < NP; i++)
            if ((unsigned)(w[i][0] >> 32) != tag || (unsigned)(w[i][1] >> 32) != tag) ll_load2(buf + 2 * (tid + i * NT), w[i][0], w[i][1]);
    }
#pragma unroll
    for (int i = 0; i < NP; i++) { v[i][0] = __uint_as_float((unsigned)w[i][0]); v[i][1] = __uint_as_float((unsigned)w[i][1]); }
}

__device__ __forceinline__ uint32_t sk_pack_bf16(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b); // .x = a (low half)
    return *reinterpret_cast<const uint32_t *>(&v);
}
__device__ __forceinline__ bool sk_better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

// elements (2p, 2p+1) of sequence s -> hi / lo words of the B-fragment image ([kb][lane = column*4 + tig][2 regs])
template <int NSEQ>
__device__ __forceinline__ void sk_put_pair(uint32_t *xf, int s, int p, float v0, float v1) {
    const float h0 = __bfloat162float(__float2bfloat16_rn(v0)), h1 = __bfloat162float(__float2bfloat16_rn(v1));
    const int kb = p >> 3, jj = p & 7, tig = jj & 3, reg = jj >> 2;
    uint32_t *q = xf + kb * (16 * NSEQ) + 16 * s + tig * 2 + reg;
    q[0] = sk_pack_bf16(v0, v1);            // column 2s   : x_hi
    q[8] = sk_pack_bf16(v0 - h0, v1 - h1);  // column 2s+1 : x_lo
}


// qasr_stream.cu - the decode kernel (v3): whole greedy loop for a chunk of tokens in ONE cooperative
// launch, weights streamed from a pre-tiled HBM image, phases chained by flag-in-data exchanges
// instead of grid barriers.  Reference hot loop: qwen_asr.c:788-818 -> qwen_decoder_forward,
// qwen_asr_decoder.c:592-685; kernels qwen_asr_kernels.c:336-373,486-543,801-924,946-1010,
// 1101-1148,1233-1298.
//
// What bounds a decode step: 3.44 GB (1.7B) / 1.19 GB (0.6B) of bf16 weights read exactly once ->
// HBM.  What actually limited rounds 1-2 (profiles/r01_*, r02_mega2_*): 142 dependent phases per
// token, each paying a grid barrier (~2.6 us with skew), a re-staging of the input vector from L2
// and, for attention, serial HBM-latency loads: 40 us per layer against 15.4 us of HBM time; and
// 2-D TMA boxes with 128-byte rows fetch at only 3.6 TB/s (tools/microbench/cluster16.cu).
//
// Design
//  * Weight image.  At load time every decoder matrix (and the tied lm_head) is re-tiled on the
//    device into "units" of 16 rows x 64 columns (2 KB) stored in mma.m16n8k16 A-fragment order, and
//    the units are laid out in HBM in exactly the order each (CTA, warp) consumes them.  A warp's
//    whole per-token weight stream is ONE contiguous byte range that it walks cyclically with 2 KB
//    1-D bulk copies (cp.async.bulk + mbarrier complete_tx) into a private 5-slot shared-memory
//    ring: no tensor maps, no address arithmetic, fully sequential DRAM pages, and the ring keeps
//    filling across phase (and token) boundaries because weight addresses never depend on data.
//  * Dot products on tensor cores: A = weight unit (one conflict-free LDS.128 per lane per 16x16
//    tile), B = the phase input as two columns x_hi = RN(x), x_lo = RN(x - x_hi) held in shared
//    memory in B-fragment order, so D[:,0] + D[:,1] = W.x to ~2^-17 relative (weights are exact bf16).
//    Warp w owns column slices {w, w+16, ...} of every 16-row group of its CTA; the 16 per-warp partial
//    sums of a row are added in fixed order => bitwise reproducible.
//  * Exchanges.  Every phase output is written to global memory as 8-byte {f32 value, u32 tag}
//    words (one atomic 64-bit store; tag = launch base + step*(L+1) + layer + 1) and every consumer
//    polls the words it needs until the tag matches: one L2 round trip after the data lands, no
//    separate barrier, no fence, no re-read.  The residual stream x lives in every CTA's shared
//    memory (replicated, updated identically), so nothing but the exchange buffers is shared.
//  * Attention: 16 q heads x S key splits (S = ceil(keys/128) <= 4) CTAs; K/V rows of the f32 cache
//    are prefetched to L2 at the top of the layer and then loaded 8 rows at a time per warp.
//
// Phases per layer:  QKV | ATTN | WO(+residual) | GU(+SwiGLU) | DOWN(+residual); then HEAD (per-CTA
// argmax, exchange of the 148 winners, lowest index wins ties, reference qwen_asr_kernels.c:536-541)
// and the embedding gather of the next input row by every CTA.
#include "qasr_common.cuh"
#include "qasr_internal.h"

#include <stdio.h>

#define SK_WARPS 16
#define SK_THREADS (SK_WARPS * 32)
#define SK_SLOTS 5
#define SK_UNIT 2048          /* 16 rows x 64 cols bf16, A-fragment order: [kb 0..3][lane][a0..a3] */
#define SK_CHUNK_GROUPS 8     /* 16-row groups reduced together (128 rows) */
#define SK_PSTRIDE 17
#define SK_MAX_K 6144
#define SK_MAX_H 2048
#define SK_ATT_MAXS 4
#define SK_ATT_STRIDE 130     /* 128 acc + m + l */

typedef unsigned long long u64;

struct SkSmem {
    uint8_t ring[SK_WARPS][SK_SLOTS][SK_UNIT];            // 163840 B
    uint32_t xf[SK_MAX_K];                                 // phase input, B-fragment order: [kb][8 lanes][2] (24576 B); attention scratch
    float x[SK_MAX_H];                                     // residual stream (replicated in every CTA)
    float partial[2][SK_CHUNK_GROUPS * 16][SK_PSTRIDE];    // [buffer][row in chunk][warp]
    uint64_t bar[SK_WARPS][SK_SLOTS];
    float red[64];
    int redi[SK_WARPS];
};

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

// ---- re-tiling: row-major [N, K] bf16 -> the image (one CTA per (cta b, phase wp)) -------------
struct RetileArgs {
    const bf16_t *src[28 * 4 + 1];
};
__global__ void __launch_bounds__(SK_THREADS) sk_retile_kernel(const RetileArgs a, SkDims d, const u64 *cta_off, uint8_t *image) {
    const int b = blockIdx.x, wp = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
    __shared__ int s_off;
    if (threadIdx.x == 0) {
        int o = 0;
        for (int q = 0; q < wp; q++) o += sk_phase_units(d, q, b);
        s_off = o;
    }
    __syncthreads();
    int N, K;
    sk_phase_shape(d, wp, N, K);
    const int g0 = sk_g0(N >> 4, b, d.G), g1 = sk_g0(N >> 4, b + 1, d.G), nj = K >> 10;
    const u64 slen = (cta_off[b + 1] - cta_off[b]) / SK_WARPS;
    uint8_t *dst = image + cta_off[b] + (u64)warp * slen + (u64)s_off * SK_UNIT;
    const uint32_t *W = reinterpret_cast<const uint32_t *>(a.src[wp]); // pairs of bf16
    const size_t ldw = (size_t)K >> 1;
    for (int g = g0; g < g1; g++)
        for (int j = 0; j < nj; j++) {
            const int slice = warp + 16 * j;
            uint4 *u = reinterpret_cast<uint4 *>(dst + (size_t)((g - g0) * nj + j) * SK_UNIT);
#pragma unroll
            for (int kb = 0; kb < 4; kb++) {
                const size_t c = (size_t)(slice * 64 + kb * 16 + 2 * tig) >> 1;
                const size_t r0 = (size_t)(g * 16 + gid) * ldw, r1 = (size_t)(g * 16 + gid + 8) * ldw;
                uint4 v;
                v.x = W[r0 + c]; v.y = W[r1 + c]; v.z = W[r0 + c + 4]; v.w = W[r1 + c + 4];
                u[kb * 32 + lane] = v;
            }
        }
}

// ---- device helpers ----------------------------------------------------------------------------
__device__ __forceinline__ void sk_csync() { asm volatile("bar.sync 1, %0;" ::"n"(SK_THREADS) : "memory"); }
__device__ __forceinline__ uint32_t sk_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sk_mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(sk_smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
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
__device__ __forceinline__ float ll_wait1(const u64 *p, unsigned tag) {
    u64 w = ll_load1(p);
    while ((unsigned)(w >> 32) != tag) w = ll_load1(p);
    return __uint_as_float((unsigned)w);
}
// Poll NP pairs of consecutive words per thread (pair p = tid + i*512, valid while p < npairs).
template <int NP>
__device__ __forceinline__ void ll_gather_pairs(const u64 *buf, int npairs, unsigned tag, int tid, float (&v)[NP][2]) {
    u64 w[NP][2];
#pragma unroll
    for (int i = 0; i < NP; i++) {
        const int p = tid + i * SK_THREADS;
        if (p < npairs) ll_load2(buf + 2 * p, w[i][0], w[i][1]);
        else w[i][0] = w[i][1] = (u64)tag << 32;
    }
    bool ok;
    do {
        ok = true;
#pragma unroll
        for (int i = 0; i < NP; i++)
            if ((unsigned)(w[i][0] >> 32) != tag || (unsigned)(w[i][1] >> 32) != tag) {
                ll_load2(buf + 2 * (tid + i * SK_THREADS), w[i][0], w[i][1]);
                ok = false;
            }
    } while (!ok);
#pragma unroll
    for (int i = 0; i < NP; i++) { v[i][0] = __uint_as_float((unsigned)w[i][0]); v[i][1] = __uint_as_float((unsigned)w[i][1]); }
}

__device__ __forceinline__ uint32_t sk_pack_bf16(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b); // .x = a (low half)
    return *reinterpret_cast<const uint32_t *>(&v);
}
// elements (2p, 2p+1) of the phase input -> hi / lo words of the B-fragment image
__device__ __forceinline__ void sk_put_pair(uint32_t *xf, int p, float v0, float v1) {
    const float h0 = __bfloat162float(__float2bfloat16_rn(v0)), h1 = __bfloat162float(__float2bfloat16_rn(v1));
    const int kb = p >> 3, jj = p & 7, tig = jj & 3, reg = jj >> 2;
    xf[kb * 16 + tig * 2 + reg] = sk_pack_bf16(v0, v1);               // lanes 0-3  (column 0: x_hi)
    xf[kb * 16 + 8 + tig * 2 + reg] = sk_pack_bf16(v0 - h0, v1 - h1); // lanes 4-7  (column 1: x_lo)
}
__device__ __forceinline__ bool sk_better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__device__ __forceinline__ float sk_block_sum(float v, float *red, int tid) {
    v = warp_sum(v);
    sk_csync();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    sk_csync();
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < SK_WARPS; w++) t += red[w];
    return t;
}

__global__ void __launch_bounds__(SK_THREADS, 1) decode_stream_kernel(const StreamParams p) {
    extern __shared__ __align__(1024) uint8_t sk_raw[];
    SkSmem &sm = *reinterpret_cast<SkSmem *>((reinterpret_cast<uintptr_t>(sk_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int b = blockIdx.x, G = gridDim.x;
    const int L = p.n_layers, H = p.H, I = p.I;

    if (lane == 0)
        for (int s = 0; s < SK_SLOTS; s++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sk_smem_u32(&sm.bar[warp][s])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int e = tid; e < H; e += SK_THREADS) sm.x[e] = p.x_io[e];
    __syncthreads();

    // ---- per-warp weight stream (contiguous, cyclic): the fetch cursor runs SK_SLOTS units ahead
    const u64 coff = p.cta_off[b];
    const uint32_t slen = (uint32_t)((p.cta_off[b + 1] - coff) / SK_WARPS);
    const uint8_t *sbase = p.image + coff + (u64)warp * slen;
    uint32_t foff = 0;
    int fsteps = 0;
    unsigned issued = 0, consumed = 0;
    auto top_up = [&]() {
        while (issued - consumed < SK_SLOTS && fsteps < p.n_steps) {
            if (lane == 0) {
                const unsigned slot = issued % SK_SLOTS;
                const uint32_t bar = sk_smem_u32(&sm.bar[warp][slot]), dst = sk_smem_u32(sm.ring[warp][slot]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "n"(SK_UNIT) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(sbase + foff), "n"(SK_UNIT), "r"(bar) : "memory");
            }
            issued++;
            foff += SK_UNIT;
            if (foff == slen) { foff = 0; fsteps++; }
        }
    };
    top_up();

    // ---- one weighted phase: y[row] = W[row,:] . x for the CTA's rows, handed to epi(row, r, rowsum)
    // in chunks of <= 128 rows; `rowsum(r)` adds the 16 per-warp partials of chunk row r in fixed order.
    int pbuf = 0;
    auto run_phase = [&](int N, int K, auto &&epi) {
        const int g0 = sk_g0(N >> 4, b, G), g1 = sk_g0(N >> 4, b + 1, G), nj = K >> 10;
        for (int cg0 = g0; cg0 < g1; cg0 += SK_CHUNK_GROUPS) {
            const int cg1 = min(cg0 + SK_CHUNK_GROUPS, g1);
            float(*part)[SK_PSTRIDE] = sm.partial[pbuf];
            for (int grp = cg0; grp < cg1; grp++) {
                float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
                for (int j = 0; j < nj; j++) {
                    const unsigned slot = consumed % SK_SLOTS;
                    sk_mbar_wait(&sm.bar[warp][slot], (consumed / SK_SLOTS) & 1);
                    const uint4 *tile = reinterpret_cast<const uint4 *>(sm.ring[warp][slot]) + lane;
                    const uint2 *xb = reinterpret_cast<const uint2 *>(sm.xf) + (size_t)(warp + 16 * j) * 32 + (lane & 7);
                    uint4 a[4];
                    uint2 bb[4];
#pragma unroll
                    for (int kb = 0; kb < 4; kb++) {
                        a[kb] = tile[kb * 32];
                        bb[kb] = xb[kb * 8];
                        if (lane >= 8) bb[kb] = make_uint2(0u, 0u);
                    }
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3) : "r"(a[0].x), "r"(a[0].y), "r"(a[0].z), "r"(a[0].w), "r"(bb[0].x), "r"(bb[0].y));
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3) : "r"(a[1].x), "r"(a[1].y), "r"(a[1].z), "r"(a[1].w), "r"(bb[1].x), "r"(bb[1].y));
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3) : "r"(a[2].x), "r"(a[2].y), "r"(a[2].z), "r"(a[2].w), "r"(bb[2].x), "r"(bb[2].y));
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3) : "r"(a[3].x), "r"(a[3].y), "r"(a[3].z), "r"(a[3].w), "r"(bb[3].x), "r"(bb[3].y));
                    __syncwarp();
                    consumed++;
                    top_up(); // the freed slot immediately takes the next unit (possibly of a later phase / token)
                }
                if (tig == 0) { // column 0 = x_hi sums, column 1 = x_lo sums; rows gid and gid+8
                    const int r = (grp - cg0) * 16 + gid;
                    part[r][warp] = (c0 + d0) + (c1 + d1);
                    part[r + 8][warp] = (c2 + d2) + (c3 + d3);
                }
            }
            sk_csync();
            const int rows = (cg1 - cg0) * 16;
            if (tid < rows) {
                auto rowsum = [&](int r) {
                    float y = 0.0f;
#pragma unroll
                    for (int i = 0; i < SK_WARPS; i++) y += part[r][i];
                    return y;
                };
                epi(cg0 * 16 + tid, tid, rowsum);
            }
            pbuf ^= 1; // the next chunk / phase writes the other buffer: one bar.sync per chunk is enough
        }
    };

    // x (shared, or gathered from an exchange buffer first) -> RMSNorm -> B-fragment image
    auto stage_norm = [&](const u64 *src, unsigned tag, const float *gamma) {
        float v[2][2];
        const int npairs = H >> 1;
        if (src) {
            ll_gather_pairs<2>(src, npairs, tag, tid, v);
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const int pr = tid + i * SK_THREADS;
                if (pr < npairs) *reinterpret_cast<float2 *>(sm.x + 2 * pr) = make_float2(v[i][0], v[i][1]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const int pr = tid + i * SK_THREADS;
                const float2 t = pr < npairs ? *reinterpret_cast<const float2 *>(sm.x + 2 * pr) : make_float2(0.f, 0.f);
                v[i][0] = t.x; v[i][1] = t.y;
            }
        }
        float ss = 0.0f;
#pragma unroll
        for (int i = 0; i < 2; i++) ss = fmaf(v[i][0], v[i][0], fmaf(v[i][1], v[i][1], ss));
        const float tot = sk_block_sum(ss, sm.red, tid);
        const float inv = 1.0f / sqrtf(tot / (float)H + p.eps);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int pr = tid + i * SK_THREADS;
            if (pr < npairs) {
                const float2 gm = *reinterpret_cast<const float2 *>(gamma + 2 * pr);
                sk_put_pair(sm.xf, pr, v[i][0] * inv * gm.x, v[i][1] * inv * gm.y);
            }
        }
        sk_csync();
    };

    long long *prof = (p.prof && (b == 0 || b == G - 1) && tid == 0) ? p.prof + (b == 0 ? 0 : p.prof_cap) : nullptr;
    int prof_n = 0;
    auto mark = [&]() { if (prof && prof_n < p.prof_cap) prof[prof_n++] = clock64(); };

    const size_t kvd = 1024;
    const float scale = 0.08838834764831845f; // 1/sqrtf(128)
    int pos = *p.d_pos;
    int step = 0;
    bool stop = false;

    for (; step < p.n_steps && !stop; step++) {
        for (int l = 0; l < L; l++) {
            const unsigned tag = p.tag_base + (unsigned)(step * (L + 1) + l + 1);
            float *kc = p.kv_k + (size_t)l * p.kv_layer_stride, *vc = p.kv_v + (size_t)l * p.kv_layer_stride;
            // attention role of this CTA: q head hq, key split sp of S
            const int n_keys = pos + 1;
            int S = (n_keys + 127) >> 7;
            S = S > SK_ATT_MAXS ? SK_ATT_MAXS : S;
            const bool att = b < 16 * S;
            const int hq = b / S, sp = b - hq * S, hkv = hq >> 1;
            const int per = (n_keys + S - 1) / S;
            const int k0 = sp * per, k1 = min(n_keys, k0 + per);
            if (att) // K/V rows of this split -> L2 while the QKV phase runs
                for (int j = k0 + tid; j < k1 && j < pos; j += SK_THREADS) {
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], 512;" ::"l"(kc + (size_t)j * kvd + hkv * 128) : "memory");
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], 512;" ::"l"(vc + (size_t)j * kvd + hkv * 128) : "memory");
                }
            mark();
            // ---------------- QKV
            stage_norm(l > 0 ? p.ll_xdn : nullptr, tag - 1, p.in_norm[l]);
            mark();
            run_phase(4096, H, [&](int row, int r, auto &&rowsum) { ll_store(p.ll_qkv + row, rowsum(r), tag); });
            mark();
            // ---------------- ATTN
            if (att) {
                float *sc = reinterpret_cast<float *>(sm.xf); // the QKV input image is dead: attention scratch
                float *qs = sc, *knew = sc + 128, *vnew = sc + 256, *tmp = sc + 384, *wacc = sc + 640, *wml = sc + 640 + SK_WARPS * 128;
                const bool has_new = (k1 == n_keys) && (k0 < k1);
                const bool writer = has_new && !(hq & 1);
                float val = 0.0f;
                const int d = tid & 127;
                if (tid < 384) {
                    const int src = tid < 128 ? hq * 128 + d : (tid < 256 ? 2048 + hkv * 128 + d : 3072 + hkv * 128 + d);
                    val = ll_wait1(p.ll_qkv + src, tag);
                }
                { // sums of squares of q (warps 0-3) and k (warps 4-7)
                    const float s2 = warp_sum(val * val);
                    if (lane == 0 && warp < 8) sm.red[warp] = s2;
                }
                sk_csync();
                if (tid < 256) {
                    const int q4 = (tid >> 7) * 4;
                    const float s = sm.red[q4] + sm.red[q4 + 1] + sm.red[q4 + 2] + sm.red[q4 + 3];
                    tmp[tid] = val * (1.0f / sqrtf(s / 128.0f + p.eps)) * (tid < 128 ? p.qn[l][d] : p.kn[l][d]);
                } else if (tid < 384) {
                    vnew[d] = val;
                    if (writer) vc[(size_t)pos * kvd + hkv * 128 + d] = val;
                }
                sk_csync();
                if (tid < 256) {
                    const int dd = d & 63, base = tid & 128;
                    const float c = p.rope_cos[(size_t)pos * 64 + dd], sn = p.rope_sin[(size_t)pos * 64 + dd];
                    const int partner = d < 64 ? d + 64 : d - 64;
                    const float sgn = d < 64 ? -1.0f : 1.0f;
                    const float r = tmp[tid] * c + sgn * tmp[base + partner] * sn;
                    if (tid < 128) qs[d] = r;
                    else {
                        knew[d] = r;
                        if (writer) kc[(size_t)pos * kvd + hkv * 128 + d] = r;
                    }
                }
                sk_csync();
                const float4 q4v = *reinterpret_cast<const float4 *>(qs + lane * 4);
                float m = -1e30f, lsum = 0.0f;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int base = k0; base < k1; base += 8 * SK_WARPS) {
                    float4 kr[8], vr[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int j = base + warp + SK_WARPS * i;
                        if (j < k1) {
                            if (j == pos) {
                                kr[i] = *reinterpret_cast<const float4 *>(knew + lane * 4);
                                vr[i] = *reinterpret_cast<const float4 *>(vnew + lane * 4);
                            } else {
                                kr[i] = __ldcg(reinterpret_cast<const float4 *>(kc + (size_t)j * kvd + hkv * 128 + lane * 4));
                                vr[i] = __ldcg(reinterpret_cast<const float4 *>(vc + (size_t)j * kvd + hkv * 128 + lane * 4));
                            }
                        }
                    }
                    float sc8[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int j = base + warp + SK_WARPS * i;
                        float d4 = 0.0f;
                        if (j < k1) d4 = q4v.x * kr[i].x + q4v.y * kr[i].y + q4v.z * kr[i].z + q4v.w * kr[i].w;
                        sc8[i] = warp_sum(d4) * scale;
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int j = base + warp + SK_WARPS * i;
                        if (j < k1) {
                            const float s = sc8[i];
                            if (s > m) {
                                const float c = expf(m - s);
                                lsum = lsum * c + 1.0f;
                                acc.x = acc.x * c + vr[i].x; acc.y = acc.y * c + vr[i].y; acc.z = acc.z * c + vr[i].z; acc.w = acc.w * c + vr[i].w;
                                m = s;
                            } else {
                                const float w = expf(s - m);
                                lsum += w;
                                acc.x += w * vr[i].x; acc.y += w * vr[i].y; acc.z += w * vr[i].z; acc.w += w * vr[i].w;
                            }
                        }
                    }
                }
                if (lane == 0) { wml[warp * 2] = m; wml[warp * 2 + 1] = lsum; }
                *reinterpret_cast<float4 *>(wacc + warp * 128 + lane * 4) = acc;
                sk_csync();
                if (tid < 128) { // merge the warps in fixed order
                    float M = -1e30f;
#pragma unroll
                    for (int w = 0; w < SK_WARPS; w++) M = fmaxf(M, wml[w * 2]);
                    float Ls = 0.0f, A = 0.0f;
#pragma unroll
                    for (int w = 0; w < SK_WARPS; w++) {
                        const float e = expf(wml[w * 2] - M);
                        Ls += wml[w * 2 + 1] * e;
                        A += wacc[w * 128 + tid] * e;
                    }
                    u64 *pb = p.ll_att + (size_t)(hq * SK_ATT_MAXS + sp) * SK_ATT_STRIDE;
                    ll_store(pb + tid, A, tag);
                    if (tid == 0) { ll_store(pb + 128, M, tag); ll_store(pb + 129, Ls, tag); }
                }
                sk_csync(); // scratch (xf) is rewritten by the WO staging below
            }
            mark();
            // ---------------- WO: input = attention output merged over the S key splits
            {
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const int pr = tid + i * SK_THREADS; // elements 2pr, 2pr+1 of the 2048-wide head-major vector
                    const int hh = pr >> 6, dd = (pr & 63) * 2;
                    const u64 *pb = p.ll_att + (size_t)(hh * SK_ATT_MAXS) * SK_ATT_STRIDE;
                    float o0 = 0.f, o1 = 0.f, Ls = 0.f, M = -1e30f;
                    float a0[SK_ATT_MAXS], a1[SK_ATT_MAXS], ms[SK_ATT_MAXS], ls[SK_ATT_MAXS];
#pragma unroll
                    for (int s = 0; s < SK_ATT_MAXS; s++) {
                        if (s < S) {
                            u64 w0, w1, w2, w3;
                            ll_load2(pb + s * SK_ATT_STRIDE + dd, w0, w1);
                            ll_load2(pb + s * SK_ATT_STRIDE + 128, w2, w3);
                            while ((unsigned)(w0 >> 32) != tag || (unsigned)(w1 >> 32) != tag) ll_load2(pb + s * SK_ATT_STRIDE + dd, w0, w1);
                            while ((unsigned)(w2 >> 32) != tag || (unsigned)(w3 >> 32) != tag) ll_load2(pb + s * SK_ATT_STRIDE + 128, w2, w3);
                            a0[s] = __uint_as_float((unsigned)w0); a1[s] = __uint_as_float((unsigned)w1);
                            ms[s] = __uint_as_float((unsigned)w2); ls[s] = __uint_as_float((unsigned)w3);
                            M = fmaxf(M, ms[s]);
                        }
                    }
#pragma unroll
                    for (int s = 0; s < SK_ATT_MAXS; s++)
                        if (s < S) {
                            const float e = expf(ms[s] - M);
                            Ls += ls[s] * e; o0 += a0[s] * e; o1 += a1[s] * e;
                        }
                    const float invL = Ls > 0.0f ? 1.0f / Ls : 0.0f;
                    sk_put_pair(sm.xf, pr, o0 * invL, o1 * invL);
                }
                sk_csync();
                mark();
                run_phase(H, 2048, [&](int row, int r, auto &&rowsum) { ll_store(p.ll_xwo + row, sm.x[row] + rowsum(r), tag); });
            }
            mark();
            // ---------------- GU + SwiGLU: rows (2j, 2j+1) = (gate_j, up_j) are neighbours in a chunk
            stage_norm(p.ll_xwo, tag, p.post_norm[l]);
            mark();
            run_phase(2 * I, H, [&](int row, int r, auto &&rowsum) {
                if (!(row & 1)) ll_store(p.ll_act + (row >> 1), silu(rowsum(r)) * rowsum(r + 1), tag);
            });
            mark();
            // ---------------- DOWN
            {
                float v[6][2];
                ll_gather_pairs<6>(p.ll_act, I >> 1, tag, tid, v);
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    const int pr = tid + i * SK_THREADS;
                    if (pr < (I >> 1)) sk_put_pair(sm.xf, pr, v[i][0], v[i][1]);
                }
                sk_csync();
                mark();
                run_phase(H, I, [&](int row, int r, auto &&rowsum) { ll_store(p.ll_xdn + row, sm.x[row] + rowsum(r), tag); });
            }
            mark();
        }
        // ---------------- HEAD: greedy argmax over this CTA's vocab rows of the tied embedding
        const unsigned htag = p.tag_base + (unsigned)(step * (L + 1) + L + 1);
        stage_norm(p.ll_xdn, htag - 1, p.final_norm);
        float bv = -1e30f;
        int bi = 0x7fffffff;
        run_phase(p.V, H, [&](int row, int r, auto &&rowsum) { const float y = rowsum(r); if (sk_better(y, row, bv, bi)) { bv = y; bi = row; } });
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(QASR_FULL, bv, o);
            const int oi = __shfl_xor_sync(QASR_FULL, bi, o);
            if (sk_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        sk_csync();
        if (lane == 0) { sm.red[warp] = bv; sm.redi[warp] = bi; }
        sk_csync();
        __threadfence(); // KV rows appended this step become visible device-wide before the token exchange
        if (tid == 0) {
#pragma unroll
            for (int w = 1; w < SK_WARPS; w++)
                if (sk_better(sm.red[w], sm.redi[w], bv, bi)) { bv = sm.red[w]; bi = sm.redi[w]; }
            ll_store(p.ll_head + 2 * b, bv, htag);
            ll_store_u32(p.ll_head + 2 * b + 1, (unsigned)bi, htag);
        }
        { // every CTA reduces the G winners identically
            float wv = -1e30f;
            int wi = 0x7fffffff;
            if (tid < G) {
                u64 w0, w1;
                ll_load2(p.ll_head + 2 * tid, w0, w1);
                while ((unsigned)(w0 >> 32) != htag || (unsigned)(w1 >> 32) != htag) ll_load2(p.ll_head + 2 * tid, w0, w1);
                wv = __uint_as_float((unsigned)w0);
                wi = (int)(unsigned)w1;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(QASR_FULL, wv, o);
                const int oi = __shfl_xor_sync(QASR_FULL, wi, o);
                if (sk_better(ov, oi, wv, wi)) { wv = ov; wi = oi; }
            }
            sk_csync();
            if (lane == 0) { sm.red[warp] = wv; sm.redi[warp] = wi; }
            sk_csync();
            wv = sm.red[0]; wi = sm.redi[0];
#pragma unroll
            for (int w = 1; w < SK_WARPS; w++)
                if (sk_better(sm.red[w], sm.redi[w], wv, wi)) { wv = sm.red[w]; wi = sm.redi[w]; }
            const int tok = wi;
            __threadfence();
            pos++;
            // next input row: exact bf16 -> f32 upcast of the embedding (reference qwen_asr.c:412-419,816)
            for (int e = tid; e < H; e += SK_THREADS) sm.x[e] = __uint_as_float(((uint32_t)p.emb[(size_t)tok * H + e]) << 16);
            if (b == 0 && tid == 0) {
                p.d_tokens[step] = tok;
                if (p.h_tokens) p.h_tokens[step] = tok;
            }
            stop = (tok == 151643 || tok == 151645); // reference qwen_asr.c:792
            sk_csync();
        }
        mark();
    }
    if (b == 0) {
        for (int e = tid; e < H; e += SK_THREADS) p.x_io[e] = sm.x[e];
        if (tid == 0) { *p.d_pos = pos; *p.d_step = step; }
    }
    // drain bulk copies that were prefetched past an early stop before the CTA's shared memory goes away
    while (consumed < issued) {
        sk_mbar_wait(&sm.bar[warp][consumed % SK_SLOTS], (consumed / SK_SLOTS) & 1);
        consumed++;
    }
}

// ---- host side ---------------------------------------------------------------------------------
static char g_sk_err[256] = "";
const char *stream_error(void) { return g_sk_err; }
static int g_sk_grid = 0;

int stream_init(void) {
    if (g_sk_grid) return 0;
    int dev = 0, sms = 0, coop = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaError_t e = cudaFuncSetAttribute(decode_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SkSmem) + 1024);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_stream_kernel, SK_THREADS, sizeof(SkSmem) + 1024);
    if (e != cudaSuccess || !coop || per_sm < 1 || sms < 1) {
        snprintf(g_sk_err, sizeof g_sk_err, "decode stream kernel unavailable: %s (coop=%d, blocks/SM=%d, smem=%zu)",
                 cudaGetErrorString(e), coop, per_sm, sizeof(SkSmem));
        cudaGetLastError();
        return -1;
    }
    g_sk_grid = sms;
    return 0;
}
int stream_grid(void) { return stream_init() == 0 ? g_sk_grid : 0; }

static bool sk_dims_ok(int L, int H, int I, int V) {
    return L >= 1 && L <= 28 && (H == 1024 || H == 2048) && I % 1024 == 0 && I <= SK_MAX_K && V % 16 == 0;
}

// Size of the image and the per-CTA byte offsets (host array of grid+1 entries).
size_t stream_image_layout(int L, int H, int I, int V, unsigned long long *cta_off_host) {
    if (stream_init() != 0 || !sk_dims_ok(L, H, I, V)) return 0;
    SkDims d{L, H, I, V, g_sk_grid};
    u64 off = 0;
    for (int b = 0; b < d.G; b++) {
        cta_off_host[b] = off;
        u64 units = 0;
        for (int wp = 0; wp <= 4 * L; wp++) units += (u64)sk_phase_units(d, wp, b);
        off += units * SK_UNIT * SK_WARPS;
    }
    cta_off_host[d.G] = off;
    return (size_t)off;
}

int stream_build_image(cudaStream_t s, int L, int H, int I, int V, const bf16_t *const *layer_mats /* [L*4] */, const bf16_t *emb,
                       const unsigned long long *d_cta_off, uint8_t *image) {
    if (stream_init() != 0) return -1;
    if (!sk_dims_ok(L, H, I, V)) { snprintf(g_sk_err, sizeof g_sk_err, "decode stream kernel: unsupported dims L=%d H=%d I=%d V=%d", L, H, I, V); return -1; }
    RetileArgs a;
    for (int i = 0; i < 4 * L; i++) a.src[i] = layer_mats[i];
    a.src[4 * L] = emb;
    SkDims d{L, H, I, V, g_sk_grid};
    sk_retile_kernel<<<dim3(d.G, 4 * L + 1), SK_THREADS, 0, s>>>(a, d, d_cta_off, image);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(g_sk_err, sizeof g_sk_err, "re-tile launch: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}

int launch_decode_stream(cudaStream_t s, const StreamParams &p) {
    if (stream_init() != 0) return -1;
    if (!sk_dims_ok(p.n_layers, p.H, p.I, p.V) || p.n_steps > 64 || p.n_steps < 1) {
        snprintf(g_sk_err, sizeof g_sk_err, "decode stream kernel: unsupported dims H=%d I=%d V=%d steps=%d", p.H, p.I, p.V, p.n_steps);
        return -1;
    }
    void *args[] = {(void *)&p};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)decode_stream_kernel, dim3(g_sk_grid), dim3(SK_THREADS), args,
                                                sizeof(SkSmem) + 1024, s);
    if (e != cudaSuccess) {
        snprintf(g_sk_err, sizeof g_sk_err, "decode stream kernel launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

// qasr_stream.cu - the decode kernel: whole greedy loop for a chunk of tokens (and up to 4 independent sequences)
// in ONE cooperative launch, weights streamed from a pre-tiled HBM image, phases chained by flag-in-data
// exchanges instead of grid barriers.  Reference hot loop: qwen_asr.c:788-818 -> qwen_decoder_forward,
// qwen_asr_decoder.c:592-685; kernels qwen_asr_kernels.c:336-373,486-543,801-924,946-1010,1101-1148,1233-1298.
//
// What bounds a decode step: 3.44 GB (1.7B) / 1.19 GB (0.6B) of bf16 weights read exactly once -> HBM.
// What actually limited the earlier kernels of this round (profiles/r01_megakernel_*, r01_mega2_*): 142 dependent
// phases per token, each paying a grid barrier (~2.6 us with skew), a re-staging of the input vector from L2 and,
// for attention, serial HBM-latency loads: 40 us per layer against 15.4 us of HBM time; 2-D TMA boxes with
// 128-byte rows fetch at only 3.6 TB/s (tools/microbench/cluster16.cu); and every bulk copy costs ~0.1 us of
// issue time on the issuing warp.
//
// Design
//  * Weight image.  At load time every decoder matrix (and the tied lm_head) is re-tiled on the device into
//    "units" of 16 rows x 64 columns (2 KB) stored in mma.m16n8k16 A-fragment order, and the units are laid out in
//    HBM in exactly the order each (CTA, warp) consumes them.  A warp's whole per-token weight stream is ONE
//    contiguous byte range that it walks cyclically: every lane copies its own 4 x 16 B of each unit with cp.async
//    (LDGSTS, 128-bit, L1 bypass) into its private bytes of the warp's ring (5 slots, 4 in the batched variants;
//    160 / 128 KB per SM) and later reads back exactly those bytes as its A fragments - a per-lane FIFO tracked by
//    the cp.async group counter alone: no mbarrier, no tensor map, no address arithmetic, sequential DRAM pages.
//    The ring keeps filling across phase (and token) boundaries because weight addresses never depend on data;
//    while a warp polls an exchange it also pulls the next units beyond its ring into L2.
//  * Dot products on tensor cores: A = weight unit (one conflict-free LDS.128 per lane per 16x16 tile), B = the
//    phase input as two columns x_hi = RN(x), x_lo = RN(x - x_hi) held in shared memory in B-fragment order, so
//    D[:,0] + D[:,1] = W.x to ~2^-17 relative (weights are exact bf16).  Sequence s of a batched launch uses
//    columns 2s, 2s+1.  Warp w owns column slices {w, w+16, ...} of every 16-row group of its CTA; the 16 per-warp
//    partial sums of a row are added in fixed order => bitwise reproducible.
//  * Exchanges.  Every phase output is written to global memory as 8-byte {f32 value, u32 tag} words (one atomic
//    64-bit store; tag = launch base + step*(L+1) + layer + 1) and every consumer polls the words it needs until the
//    tag matches: data and flag travel together, no separate barrier, no fence, no re-read.  The residual stream x
//    lives in every CTA's shared memory (replicated, updated identically), so nothing but the exchange buffers is
//    shared.  A writer can never lap a reader: every buffer is rewritten only after a later all-to-all exchange.
//  * Attention: (sequence, q head, key split) CTAs with S = ceil(keys/32) <= 4 splits; everything up to the final
//    merge is warp-local; K/V rows of the f32 cache are prefetched to L2 at the top of the layer and loaded into
//    registers before the q words are polled.
//
// Phases per layer:  QKV | ATTN | WO(+residual) | GU(+SwiGLU) | DOWN(+residual); then HEAD (per-CTA argmax,
// exchange of the 148 winners, lowest index wins ties, reference qwen_asr_kernels.c:536-541) and the embedding
// gather of the next input row by every CTA.
#include "qasr_stream_common.cuh"
#include <string.h>

// ---- re-tiling: row-major [N, K] bf16 -> the image (one CTA per (cta b, phase wp)) -------------
struct RetileArgs {
    const bf16_t *src[28 * 4 + 1];
};
// round_major: the 16 units of a (group, j) round contiguous, in the order the 8 consumer warps of qasr_stream_r.cu read
// them (warp w: old streams w, w + 8); otherwise one contiguous stream per warp.
__global__ void __launch_bounds__(SK_THREADS) sk_retile_kernel(const RetileArgs a, SkDims d, const u64 *cta_off, uint8_t *image, int round_major) {
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
    uint8_t *dst = round_major ? image + cta_off[b] + (u64)s_off * (SK_WARPS * SK_UNIT) + (u64)((warp & 7) * 2 + (warp >> 3)) * SK_UNIT
                               : image + cta_off[b] + (u64)warp * slen + (u64)s_off * SK_UNIT;
    const size_t ustride = round_major ? (size_t)SK_WARPS * SK_UNIT : (size_t)SK_UNIT;
    const uint32_t *W = reinterpret_cast<const uint32_t *>(a.src[wp]); // pairs of bf16
    const size_t ldw = (size_t)K >> 1;
    for (int g = g0; g < g1; g++)
        for (int j = 0; j < nj; j++) {
            const int slice = warp + 16 * j;
            uint4 *u = reinterpret_cast<uint4 *>(dst + (size_t)((g - g0) * nj + j) * ustride);
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

// Shared-memory carve-up (dynamic): ring | xf | x | partial | small.  NSEQ sequences share one launch: sequence s
// occupies MMA columns 2s (x_hi) and 2s+1 (x_lo) of the m16n8k16 B operand, so up to 4 independent segments /
// utterances are decoded for the weight traffic of one (the 6 columns a single sequence leaves idle are free).
template <int NSEQ, int SLOTS>
struct SkLayout {
    static constexpr int CH_GROUPS = SK_CHUNK_GROUPS / NSEQ;   // 16-row groups reduced together
    static constexpr int CH_ROWS = CH_GROUPS * 16;             // x NSEQ sequences = 128 epilogue threads
    static constexpr size_t ring_bytes = (size_t)SK_WARPS * SLOTS * SK_UNIT;
    static __host__ __device__ size_t xf_bytes(int kmax) { return (size_t)kmax * 4 * NSEQ; }            // [kb][8*NSEQ lanes][2] u32
    static __host__ __device__ size_t x_bytes(int H) { return (size_t)NSEQ * H * 4; }
    static constexpr size_t partial_bytes = (size_t)2 * 128 * SK_PSTRIDE * 4;
    static constexpr size_t small_bytes = (64 + NSEQ * SK_WARPS + SK_WARPS + NSEQ * 16 * SK_ATT_MAXS) * 4;
    static __host__ __device__ size_t total(int kmax, int H) { return ring_bytes + xf_bytes(kmax) + x_bytes(H) + partial_bytes + small_bytes + 1024; }
};

template <int NSEQ, int SLOTS>
__global__ void __launch_bounds__(SK_THREADS, 1) decode_stream_kernel(const StreamParams p) {
    using LY = SkLayout<NSEQ, SLOTS>;
    constexpr int CH_GROUPS = LY::CH_GROUPS, CH_ROWS = LY::CH_ROWS;
    extern __shared__ __align__(1024) uint8_t sk_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int b = blockIdx.x, G = gridDim.x;
    const int L = p.n_layers, H = p.H, I = p.I;
    const int kmax = max(max(H, I), 2048);
    uint8_t *sm0 = sk_raw; // used as is (16-byte alignment suffices): keeps the shared address space visible to the compiler (LDS/STS instead of generic LD/ST)
    uint8_t *sm_ring = sm0;
    uint32_t *sm_xf = reinterpret_cast<uint32_t *>(sm0 + LY::ring_bytes);
    float *sm_x = reinterpret_cast<float *>(sm0 + LY::ring_bytes + LY::xf_bytes(kmax));                  // [NSEQ][H] residual streams
    float(*sm_partial)[128][SK_PSTRIDE] = reinterpret_cast<float(*)[128][SK_PSTRIDE]>(sm0 + LY::ring_bytes + LY::xf_bytes(kmax) + LY::x_bytes(H));
    float *sm_red = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(sm_partial) + LY::partial_bytes); // [64]
    float *sm_ssq = sm_red + 64;                                                                            // [NSEQ][16]
    int *sm_redi = reinterpret_cast<int *>(sm_ssq + NSEQ * SK_WARPS);                                       // [16]
    float *sm_fac = reinterpret_cast<float *>(sm_redi + SK_WARPS);                                         // [NSEQ][16 heads][SK_ATT_MAXS] split weights

    for (int e = tid; e < NSEQ * H; e += SK_THREADS) sm_x[e] = p.x_io[e];
    __syncthreads();

    // ---- per-warp weight stream (contiguous, cyclic).  Every lane copies its own 4 x 16 B of each unit with
    // cp.async (LDGSTS, L1 bypass) into its private bytes of the warp's SLOTS-deep ring and
    // later reads back exactly those bytes as its A fragments: a per-lane FIFO, so completion is tracked by the
    // per-thread cp.async group counter alone (one group per unit, always SLOTS groups outstanding).
    const u64 coff = p.cta_off[b];
    const uint32_t slen = (uint32_t)((p.cta_off[b + 1] - coff) / SK_WARPS);
    const uint8_t *wbase = p.image + coff + (u64)warp * slen;
    const uint8_t *sbase = wbase + lane * 16;
    // Second-level prefetch, driven by the exchange waits: while a warp polls, it pulls the units that follow its
    // ring (up to `l2_window` of them) into L2, so HBM keeps streaming during the stalls the ring cannot
    // cover; the later cp.async fetches then hit L2.  Nothing is issued while the consumer is HBM-bound.
    const unsigned l2_window = (unsigned)p.l2_ahead_units;
    unsigned fpos = 0, lpos = 0;
    uint32_t loff = 0;
    uint32_t foff = 0;
    int fsteps = 0;
    unsigned consumed = 0;
    // L2 evict-first for the weight stream (read once per token) so KV rows, exchange words and L2-prefetched units stay resident
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const uint32_t ring0 = sk_smem_u32(sm_ring) + warp * (SLOTS * SK_UNIT) + lane * 16;
    auto fetch_into = [&](unsigned slot) { // next unit of the stream -> ring slot; always commits one group
        if (fsteps < p.n_steps) {
            const uint8_t *src = sbase + foff;
            uint32_t dst = ring0 + slot * SK_UNIT;
            // keep the shared address in ONE general register: when ptxas 12.9 splits it into a uniform base plus an offset
            // ([R+UR]) the cache-hinted LDGSTS is mis-assembled with undefined descriptor registers (CUDA_EXCEPTION_4,
            // "Warp Illegal Instruction", located with cuda-gdb); tests/test_abi.py checks the SASS for that form
            asm volatile("mov.u32 %0, %0;" : "+r"(dst));
#pragma unroll
            for (int kb = 0; kb < 4; kb++) {
                if constexpr (NSEQ == 1) // hinted copies only where the generated SASS is clean (see above); the batched variants use the default policy
                    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst + kb * 512), "l"(src + kb * 512), "l"(pol) : "memory");
                else
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + kb * 512), "l"(src + kb * 512) : "memory");
            }
            foff += SK_UNIT;
            if (foff == slen) { foff = 0; fsteps++; }
            fpos++;
            if (lpos < fpos) { lpos = fpos; loff = foff; }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int i = 0; i < SLOTS; i++) fetch_into(i);
    auto top_up = [&]() { // service hook of every poll loop (runs converged: the loops are warp-uniform)
#pragma unroll 1
        for (int rep = 0; rep < p.l2_issue; rep++) // units per poll iteration
        if (lpos - fpos < l2_window) {
            if (lane < 16) {
                if (p.debug & 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(wbase + loff + lane * 128) : "memory");
                else asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(wbase + loff + lane * 128) : "memory");
            }
            lpos++;
            loff += SK_UNIT;
            if (loff == slen) loff = 0;
        }
    };

    // ---- one weighted phase: y[s][row] = W[row,:] . x_s for the CTA's rows, handed to epi(s, row, r, rowsum) in
    // chunks of <= CH_ROWS rows; `rowsum(r)` adds the 16 per-warp partials of (sequence s, chunk row r) in fixed order.
    int pbuf = 0;
    auto run_phase = [&](int N, int K, auto &&epi) {
        const int g0 = sk_g0(N >> 4, b, G), g1 = sk_g0(N >> 4, b + 1, G), nj = K >> 10;
        for (int cg0 = g0; cg0 < g1; cg0 += CH_GROUPS) {
            const int cg1 = min(cg0 + CH_GROUPS, g1);
            float(*part)[SK_PSTRIDE] = sm_partial[pbuf];
            for (int grp = cg0; grp < cg1; grp++) {
                float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
                for (int j = 0; j < nj; j++) {
                    const bool tr = (p.debug & 64) && p.prof && b == p.trace_cta && tid == 0 && consumed < 1300;
                    if (tr) p.prof[2 * p.prof_cap + 3 * consumed] = clock64();
                    const uint2 *xb = reinterpret_cast<const uint2 *>(sm_xf) + (size_t)(warp + 16 * j) * (32 * NSEQ) + (lane & (8 * NSEQ - 1));
                    uint2 bb[4];
#pragma unroll
                    for (int kb = 0; kb < 4; kb++) {
                        bb[kb] = xb[kb * (8 * NSEQ)];
                        if (NSEQ < 4 && lane >= 8 * NSEQ) bb[kb] = make_uint2(0u, 0u);
                    }
                    const unsigned slot = consumed % SLOTS;
                    asm volatile("cp.async.wait_group %0;" ::"n"(SLOTS - 1) : "memory"); // this lane's bytes of the oldest unit have landed
                    if (tr) p.prof[2 * p.prof_cap + 3 * consumed + 1] = clock64();
                    const uint4 *tile = reinterpret_cast<const uint4 *>(sm_ring + (size_t)(warp * SLOTS + slot) * SK_UNIT) + lane;
                    uint4 a[4];
#pragma unroll
                    for (int kb = 0; kb < 4; kb++) a[kb] = tile[kb * 32];
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3) : "r"(a[0].x), "r"(a[0].y), "r"(a[0].z), "r"(a[0].w), "r"(bb[0].x), "r"(bb[0].y));
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3) : "r"(a[1].x), "r"(a[1].y), "r"(a[1].z), "r"(a[1].w), "r"(bb[1].x), "r"(bb[1].y));
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3) : "r"(a[2].x), "r"(a[2].y), "r"(a[2].z), "r"(a[2].w), "r"(bb[2].x), "r"(bb[2].y));
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3) : "r"(a[3].x), "r"(a[3].y), "r"(a[3].z), "r"(a[3].w), "r"(bb[3].x), "r"(bb[3].y));
                    if (tr) p.prof[2 * p.prof_cap + 3 * consumed + 2] = clock64();
                    consumed++;
                    fetch_into(slot); // the freed slot immediately takes the next unit (possibly of a later phase / token)
                }
                if (tig < NSEQ) { // D columns (2 tig, 2 tig + 1) = x_hi / x_lo sums of sequence tig; rows gid and gid+8
                    const int r = tig * CH_ROWS + (grp - cg0) * 16 + gid;
                    part[r][warp] = (c0 + d0) + (c1 + d1);
                    part[r + 8][warp] = (c2 + d2) + (c3 + d3);
                }
            }
            sk_csync();
            const int rows = (cg1 - cg0) * 16;
            const int es = tid / CH_ROWS, er = tid - es * CH_ROWS; // epilogue thread <-> (sequence, chunk row)
            if (tid < NSEQ * CH_ROWS && er < rows) {
                auto rowsum = [&](int r) {
                    float y = 0.0f;
#pragma unroll
                    for (int i = 0; i < SK_WARPS; i++) y += part[es * CH_ROWS + r][i];
                    return y;
                };
                epi(es, cg0 * 16 + er, er, rowsum);
            }
            pbuf ^= 1; // the next chunk / phase writes the other buffer: one bar.sync per chunk is enough
        }
    };

    // x (shared, or gathered from an exchange buffer first) -> x * gamma -> B-fragment image.  The RMSNorm scale
    // 1/sqrt(mean(x^2)+eps) is a scalar, so it is applied to the phase OUTPUT (norm_scale() in the epilogue):
    // the block reduction of the squares leaves the critical path between the gather and the first MMA.
    // All NSEQ vectors are gathered as one long vector (pair q = s * H/2 + pr), every load in flight before the first check.
    constexpr int NPX = NSEQ == 4 ? 4 : 2 * NSEQ; // 512 threads x NPX pairs cover NSEQ x H/2 (NSEQ = 4 only with H = 1024)
    auto stage_norm = [&](const u64 *src, unsigned tag, const float2 (&gm2)[2]) {
        float v[NPX][2];
        const int hp = H >> 1, npairs = NSEQ * hp;
        if (src) {
            ll_gather_pairs<NPX>(src, npairs, tag, tid, v, top_up);
#pragma unroll
            for (int i = 0; i < NPX; i++) {
                const int q = tid + i * SK_THREADS;
                if (q < npairs) *reinterpret_cast<float2 *>(sm_x + 2 * q) = make_float2(v[i][0], v[i][1]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < NPX; i++) {
                const int q = tid + i * SK_THREADS;
                const float2 t = q < npairs ? *reinterpret_cast<const float2 *>(sm_x + 2 * q) : make_float2(0.f, 0.f);
                v[i][0] = t.x; v[i][1] = t.y;
            }
        }
        // pair q belongs to sequence q / hp; with hp a multiple of 512 (H = 1024, 2048) each i maps to ONE sequence per thread
        float ss[NSEQ];
#pragma unroll
        for (int s = 0; s < NSEQ; s++) ss[s] = 0.0f;
#pragma unroll
        for (int i = 0; i < NPX; i++) {
            const int q = tid + i * SK_THREADS;
            if (q < npairs) {
                const int s = NSEQ == 1 ? 0 : q / hp, pr = q - s * hp;
                const float2 gm = (hp > SK_THREADS && (i & 1)) ? gm2[1] : gm2[0]; // pr = tid + 512 * (i mod hp/512)
                const float sq = fmaf(v[i][0], v[i][0], v[i][1] * v[i][1]);
#pragma unroll
                for (int t = 0; t < NSEQ; t++) if (t == s) ss[t] += sq;
                sk_put_pair<NSEQ>(sm_xf, s, pr, v[i][0] * gm.x, v[i][1] * gm.y);
            }
        }
#pragma unroll
        for (int s = 0; s < NSEQ; s++) {
            const float t = warp_sum(ss[s]);
            if (lane == 0) sm_ssq[s * SK_WARPS + warp] = t;
        }
        sk_csync();
    };
    auto norm_scale = [&](int s) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < SK_WARPS; w++) t += sm_ssq[s * SK_WARPS + w];
        return 1.0f / sqrtf(t / (float)H + p.eps);
    };
    auto load_gamma = [&](const float *g, float2 (&gm2)[2]) { // this thread's norm weights (pairs tid, tid + 512), loaded before the exchange waits
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int pr = tid + j * SK_THREADS;
            gm2[j] = pr < (H >> 1) ? __ldg(reinterpret_cast<const float2 *>(g) + pr) : make_float2(0.f, 0.f);
        }
    };

    long long *prof = (p.prof && (b == 0 || b == G - 1) && tid == 0) ? p.prof + (b == 0 ? 0 : p.prof_cap) : nullptr;
    int prof_n = 0;
    auto mark = [&]() { if (prof && prof_n < p.prof_cap) prof[prof_n++] = clock64(); };

    const size_t kvd = 1024;
    const float scale = 0.08838834764831845f; // 1/sqrtf(128)
    int pos[NSEQ];
#pragma unroll
    for (int s = 0; s < NSEQ; s++) pos[s] = p.d_pos[s];
    unsigned done_mask = 0; // sequences that have produced an EOS token (they keep running; the host cuts their ids)
    int step = 0;
    bool stop = false;
    const size_t att_words = (size_t)16 * SK_ATT_MAXS * SK_ATT_STRIDE;

    for (; step < p.n_steps && !stop; step++) {
        // attention role of this CTA for the whole step: (sequence sa, q head hq, key split sp of S)
        int maxkeys = 0;
#pragma unroll
        for (int s = 0; s < NSEQ; s++) maxkeys = max(maxkeys, pos[s] + 1);
        int S = (maxkeys + SK_ATT_BATCH * SK_WARPS - 1) / (SK_ATT_BATCH * SK_WARPS);
        S = min(S, min(SK_ATT_MAXS, G / (16 * NSEQ)));
        const bool att = b < NSEQ * 16 * S;
        const int sa = att ? b / (16 * S) : 0, hq = (b - sa * 16 * S) / S, sp = b - sa * 16 * S - hq * S, hkv = hq >> 1;
        int apos = pos[0];
#pragma unroll
        for (int s = 1; s < NSEQ; s++) if (s == sa) apos = pos[s];
        const int n_keys = apos + 1;
        const int per = (n_keys + S - 1) / S;
        const int k0 = sp * per, k1 = min(n_keys, k0 + per);
        const float4 rope_c = __ldg(reinterpret_cast<const float4 *>(p.rope_cos + (size_t)apos * 64) + (lane & 15));
        const float4 rope_s = __ldg(reinterpret_cast<const float4 *>(p.rope_sin + (size_t)apos * 64) + (lane & 15));
        for (int l = 0; l < L; l++) {
            const unsigned tag = p.tag_base + (unsigned)(step * (L + 1) + l + 1);
            float *kc = p.kv_k[sa] + (size_t)l * p.kv_layer_stride, *vc = p.kv_v[sa] + (size_t)l * p.kv_layer_stride;
            if (att) // K/V rows of this split -> L2 while the QKV phase runs
                for (int j = k0 * 8 + tid; j < k1 * 8 && j < apos * 8; j += SK_THREADS) { // 8 lines of 128 B per key (K row + V row)
                    const float *row = ((j & 4) ? vc : kc) + (size_t)(j >> 3) * kvd + hkv * 128 + (j & 3) * 32;
                    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(row) : "memory");
                }
            // small per-layer parameters: issue the loads now, use them after the exchange waits
            float2 g_in[2], g_post[2];
            load_gamma(p.in_norm[l], g_in);
            load_gamma(p.post_norm[l], g_post);
            const float4 qn4 = __ldg(reinterpret_cast<const float4 *>(p.qn[l]) + lane), kn4 = __ldg(reinterpret_cast<const float4 *>(p.kn[l]) + lane);
            mark();
            // ---------------- QKV
            stage_norm(l > 0 ? p.ll_xdn : nullptr, tag - 1, g_in);
            mark();
            run_phase(4096, H, [&](int s, int row, int r, auto &&rowsum) { ll_store(p.ll_qkv + s * 4096 + row, rowsum(r) * norm_scale(s), tag); });
            mark();
            // ---------------- ATTN (attention CTAs only): everything up to the final merge is warp-local.
            // Every warp gathers q of the head itself (4 dims per lane), RMS-normalises and ropes it with shuffles;
            // the warp that owns the new key does the same for k and appends k/v to the cache.  The K/V rows of the
            // cached keys do not depend on this layer's output: their loads are issued BEFORE the q words are polled.
            if (att) {
                float *sc = reinterpret_cast<float *>(sm_xf); // the QKV input image is dead: attention scratch
                float *wacc = sc, *wml = sc + SK_WARPS * 128;
                const bool has_new = (k1 == n_keys) && (k0 < k1);
                const int w_new = has_new ? ((apos - k0) & (SK_WARPS - 1)) : -1; // warp whose key list contains `apos`
                const size_t hoff = (size_t)hkv * 128 + lane * 4;
                float4 kr[SK_ATT_BATCH], vr[SK_ATT_BATCH];
#pragma unroll
                for (int i = 0; i < SK_ATT_BATCH; i++) {
                    const int j = k0 + warp + SK_WARPS * i;
                    if (j < k1 && j != apos) {
                        kr[i] = __ldcg(reinterpret_cast<const float4 *>(kc + (size_t)j * kvd + hoff));
                        vr[i] = __ldcg(reinterpret_cast<const float4 *>(vc + (size_t)j * kvd + hoff));
                    }
                }
                // q (all warps), k and v (owning warp): 4 consecutive words per lane
                u64 wq[4], wk[4], wv[4];
                const u64 *qkv = p.ll_qkv + sa * 4096;
                const u64 *pq = qkv + hq * 128 + lane * 4, *pk = qkv + 2048 + hkv * 128 + lane * 4, *pv = pk + 1024;
                const bool own = warp == w_new;
                ll_load2(pq, wq[0], wq[1]); ll_load2(pq + 2, wq[2], wq[3]);
                if (own) { ll_load2(pk, wk[0], wk[1]); ll_load2(pk + 2, wk[2], wk[3]); ll_load2(pv, wv[0], wv[1]); ll_load2(pv + 2, wv[2], wv[3]); }
                else {
#pragma unroll
                    for (int i = 0; i < 4; i++) wk[i] = wv[i] = (u64)tag << 32;
                }
                for (;;) {
                    bool ok = true;
#pragma unroll
                    for (int i = 0; i < 4; i++) ok = ok && (unsigned)(wq[i] >> 32) == tag && (unsigned)(wk[i] >> 32) == tag && (unsigned)(wv[i] >> 32) == tag;
                    if (__all_sync(QASR_FULL, ok)) break;
                    top_up();
                    if ((unsigned)(wq[0] >> 32) != tag || (unsigned)(wq[1] >> 32) != tag) ll_load2(pq, wq[0], wq[1]);
                    if ((unsigned)(wq[2] >> 32) != tag || (unsigned)(wq[3] >> 32) != tag) ll_load2(pq + 2, wq[2], wq[3]);
                    if (own) {
                        if ((unsigned)(wk[0] >> 32) != tag || (unsigned)(wk[1] >> 32) != tag) ll_load2(pk, wk[0], wk[1]);
                        if ((unsigned)(wk[2] >> 32) != tag || (unsigned)(wk[3] >> 32) != tag) ll_load2(pk + 2, wk[2], wk[3]);
                        if ((unsigned)(wv[0] >> 32) != tag || (unsigned)(wv[1] >> 32) != tag) ll_load2(pv, wv[0], wv[1]);
                        if ((unsigned)(wv[2] >> 32) != tag || (unsigned)(wv[3] >> 32) != tag) ll_load2(pv + 2, wv[2], wv[3]);
                    }
                }
                mark();
                // RMSNorm over the 128-vector (warp_sum), split-half RoPE: dims d and d+-64 live in lanes l and l^16
                auto norm_rope = [&](const u64 (&w)[4], const float4 nw) {
                    float4 v = make_float4(__uint_as_float((unsigned)w[0]), __uint_as_float((unsigned)w[1]), __uint_as_float((unsigned)w[2]), __uint_as_float((unsigned)w[3]));
                    const float s2 = warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
                    const float inv = 1.0f / sqrtf(s2 / 128.0f + p.eps);
                    v.x = v.x * inv * nw.x; v.y = v.y * inv * nw.y; v.z = v.z * inv * nw.z; v.w = v.w * inv * nw.w;
                    float4 o;
                    o.x = __shfl_xor_sync(QASR_FULL, v.x, 16); o.y = __shfl_xor_sync(QASR_FULL, v.y, 16);
                    o.z = __shfl_xor_sync(QASR_FULL, v.z, 16); o.w = __shfl_xor_sync(QASR_FULL, v.w, 16);
                    const float sgn = lane < 16 ? -1.0f : 1.0f;
                    return make_float4(v.x * rope_c.x + sgn * o.x * rope_s.x, v.y * rope_c.y + sgn * o.y * rope_s.y,
                                       v.z * rope_c.z + sgn * o.z * rope_s.z, v.w * rope_c.w + sgn * o.w * rope_s.w);
                };
                const float4 q4v = norm_rope(wq, qn4);
                float4 k_new = make_float4(0.f, 0.f, 0.f, 0.f), v_new = k_new;
                if (w_new >= 0) { // CTA-uniform branch; only the owning warp holds real k/v words
                    k_new = norm_rope(wk, kn4);
                    v_new = make_float4(__uint_as_float((unsigned)wv[0]), __uint_as_float((unsigned)wv[1]), __uint_as_float((unsigned)wv[2]), __uint_as_float((unsigned)wv[3]));
                    if (own) {
#pragma unroll
                        for (int i = 0; i < SK_ATT_BATCH; i++) if (k0 + warp + SK_WARPS * i == apos) { kr[i] = k_new; vr[i] = v_new; }
                        if (!(hq & 1)) { // one writer per kv head appends the new row (reference qwen_asr_decoder.c:640-646)
                            *reinterpret_cast<float4 *>(kc + (size_t)apos * kvd + hoff) = k_new;
                            *reinterpret_cast<float4 *>(vc + (size_t)apos * kvd + hoff) = v_new;
                        }
                    }
                }
                mark();
                float m = -1e30f, lsum = 0.0f;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int base = k0; base < k1; base += SK_ATT_BATCH * SK_WARPS) {
                    if (base != k0) { // later batches (only when a split holds more than SK_ATT_BATCH * 16 keys)
#pragma unroll
                        for (int i = 0; i < SK_ATT_BATCH; i++) {
                            const int j = base + warp + SK_WARPS * i;
                            if (j < k1) {
                                if (j == apos) { kr[i] = k_new; vr[i] = v_new; } // never read the row being appended from the cache
                                else {
                                    kr[i] = __ldcg(reinterpret_cast<const float4 *>(kc + (size_t)j * kvd + hoff));
                                    vr[i] = __ldcg(reinterpret_cast<const float4 *>(vc + (size_t)j * kvd + hoff));
                                }
                            }
                        }
                    }
                    float sc8[SK_ATT_BATCH];
#pragma unroll
                    for (int i = 0; i < SK_ATT_BATCH; i++) {
                        const int j = base + warp + SK_WARPS * i;
                        float d4 = 0.0f;
                        if (j < k1) d4 = q4v.x * kr[i].x + q4v.y * kr[i].y + q4v.z * kr[i].z + q4v.w * kr[i].w;
                        sc8[i] = warp_sum(d4) * scale;
                    }
#pragma unroll
                    for (int i = 0; i < SK_ATT_BATCH; i++) {
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
                mark();
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
                    u64 *pb = p.ll_att + sa * att_words + (size_t)(hq * SK_ATT_MAXS + sp) * SK_ATT_STRIDE;
                    ll_store(pb + tid, A, tag);
                    if (tid == 0) { ll_store(pb + 128, M, tag); ll_store(pb + 129, Ls, tag); }
                }
                sk_csync(); // scratch (xf) is rewritten by the WO staging below
            }
            mark();
            // ---------------- WO: input = attention output merged over the S key splits
            {
                if constexpr (NSEQ == 4) { // two-phase merge: fewer words per thread, one extra bar.sync (pays off only with 4 sequences)
                // (1) softmax weights of the key splits, once per (sequence, head): thread t < 16 NSEQ polls the S (m, l) pairs
                {
                    const bool act_a = tid < NSEQ * 16;
                    const u64 *pm = p.ll_att + (tid >> 4) * att_words + (size_t)((tid & 15) * SK_ATT_MAXS) * SK_ATT_STRIDE + 128;
                    u64 wm[SK_ATT_MAXS][2];
#pragma unroll
                    for (int t = 0; t < SK_ATT_MAXS; t++) {
                        if (act_a && t < S) ll_load2(pm + t * SK_ATT_STRIDE, wm[t][0], wm[t][1]);
                        else wm[t][0] = wm[t][1] = (u64)tag << 32;
                    }
                    for (;;) {
                        bool ok = true;
#pragma unroll
                        for (int t = 0; t < SK_ATT_MAXS; t++) ok = ok && (unsigned)(wm[t][0] >> 32) == tag && (unsigned)(wm[t][1] >> 32) == tag;
                        if (__all_sync(QASR_FULL, ok)) break;
                        top_up();
#pragma unroll
                        for (int t = 0; t < SK_ATT_MAXS; t++)
                            if ((unsigned)(wm[t][0] >> 32) != tag || (unsigned)(wm[t][1] >> 32) != tag) ll_load2(pm + t * SK_ATT_STRIDE, wm[t][0], wm[t][1]);
                    }
                    if (act_a) {
                        float M = -1e30f, Ls = 0.f, e[SK_ATT_MAXS];
#pragma unroll
                        for (int t = 0; t < SK_ATT_MAXS; t++)
                            if (t < S) M = fmaxf(M, __uint_as_float((unsigned)wm[t][0]));
#pragma unroll
                        for (int t = 0; t < SK_ATT_MAXS; t++) {
                            e[t] = t < S ? expf(__uint_as_float((unsigned)wm[t][0]) - M) : 0.0f;
                            Ls += t < S ? __uint_as_float((unsigned)wm[t][1]) * e[t] : 0.0f;
                        }
                        const float invL = Ls > 0.0f ? 1.0f / Ls : 0.0f;
#pragma unroll
                        for (int t = 0; t < SK_ATT_MAXS; t++) sm_fac[tid * SK_ATT_MAXS + t] = e[t] * invL;
                    }
                    sk_csync();
                }
                // (2) weighted sum of the split accumulators: PG pairs per thread in flight together
                auto merge_splits = [&](auto ns_c, auto pg_c) { // NS = compile-time bound on S
                    constexpr int NS = decltype(ns_c)::value, PG = decltype(pg_c)::value;
#pragma unroll 1
                    for (int i0 = 0; i0 < 2 * NSEQ; i0 += PG) {
                        u64 w[PG][NS][2];
                        const u64 *pb[PG];
#pragma unroll
                        for (int g = 0; g < PG; g++) {
                            const int q = tid + (i0 + g) * SK_THREADS;   // pair q = s * 1024 + pr: elements 2pr, 2pr+1 of sequence s' head-major vector
                            const int pr = q & 1023;
                            pb[g] = p.ll_att + (q >> 10) * att_words + (size_t)((pr >> 6) * SK_ATT_MAXS) * SK_ATT_STRIDE + (pr & 63) * 2;
#pragma unroll
                            for (int t = 0; t < NS; t++) {
                                if (t < S) ll_load2(pb[g] + t * SK_ATT_STRIDE, w[g][t][0], w[g][t][1]);
                                else w[g][t][0] = w[g][t][1] = (u64)tag << 32;
                            }
                        }
                        for (;;) {
                            bool ok = true;
#pragma unroll
                            for (int g = 0; g < PG; g++)
#pragma unroll
                                for (int t = 0; t < NS; t++) ok = ok && (unsigned)(w[g][t][0] >> 32) == tag && (unsigned)(w[g][t][1] >> 32) == tag;
                            if (__all_sync(QASR_FULL, ok)) break;
                            top_up();
#pragma unroll
                            for (int g = 0; g < PG; g++)
#pragma unroll
                                for (int t = 0; t < NS; t++)
                                    if ((unsigned)(w[g][t][0] >> 32) != tag || (unsigned)(w[g][t][1] >> 32) != tag) ll_load2(pb[g] + t * SK_ATT_STRIDE, w[g][t][0], w[g][t][1]);
                        }
#pragma unroll
                        for (int g = 0; g < PG; g++) {
                            const int q = tid + (i0 + g) * SK_THREADS, pr = q & 1023;
                            const float *f = sm_fac + ((q >> 10) * 16 + (pr >> 6)) * SK_ATT_MAXS;
                            float o0 = 0.f, o1 = 0.f;
#pragma unroll
                            for (int t = 0; t < NS; t++)
                                if (t < S) {
                                    o0 = fmaf(__uint_as_float((unsigned)w[g][t][0]), f[t], o0);
                                    o1 = fmaf(__uint_as_float((unsigned)w[g][t][1]), f[t], o1);
                                }
                            sk_put_pair<NSEQ>(sm_xf, q >> 10, pr, o0, o1);
                        }
                    }
                };
                if (S == 1) merge_splits(std::integral_constant<int, 1>{}, std::integral_constant<int, 2 * NSEQ>{});
                else if (S == 2) merge_splits(std::integral_constant<int, 2>{}, std::integral_constant<int, (NSEQ > 2 ? 4 : 2 * NSEQ)>{});
                else merge_splits(std::integral_constant<int, SK_ATT_MAXS>{}, std::integral_constant<int, 2>{});
                } else {
                auto merge_splits = [&](auto ns_c, auto pg_c) { // NS = compile-time bound on S; PG = pairs per thread whose loads fly together
                    constexpr int NS = decltype(ns_c)::value, PG = decltype(pg_c)::value;
#pragma unroll 1
                    for (int i0 = 0; i0 < 2 * NSEQ; i0 += PG) {
                        u64 w[PG][NS][4];
                        const u64 *pb[PG];
                        int dd[PG];
#pragma unroll
                        for (int g = 0; g < PG; g++) {
                            const int q = tid + (i0 + g) * SK_THREADS;   // pair q = s * 1024 + pr: elements 2pr, 2pr+1 of sequence s' head-major vector
                            const int s = q >> 10, pr = q & 1023;
                            dd[g] = (pr & 63) * 2;
                            pb[g] = p.ll_att + s * att_words + (size_t)((pr >> 6) * SK_ATT_MAXS) * SK_ATT_STRIDE;
#pragma unroll
                            for (int t = 0; t < NS; t++) {
                                if (t < S) {
                                    ll_load2(pb[g] + t * SK_ATT_STRIDE + dd[g], w[g][t][0], w[g][t][1]);
                                    ll_load2(pb[g] + t * SK_ATT_STRIDE + 128, w[g][t][2], w[g][t][3]);
                                } else w[g][t][0] = w[g][t][1] = w[g][t][2] = w[g][t][3] = (u64)tag << 32;
                            }
                        }
                        for (;;) {
                            bool ok = true;
#pragma unroll
                            for (int g = 0; g < PG; g++)
#pragma unroll
                                for (int t = 0; t < NS; t++)
                                    ok = ok && (unsigned)(w[g][t][0] >> 32) == tag && (unsigned)(w[g][t][1] >> 32) == tag && (unsigned)(w[g][t][2] >> 32) == tag && (unsigned)(w[g][t][3] >> 32) == tag;
                            if (__all_sync(QASR_FULL, ok)) break;
                            top_up();
#pragma unroll
                            for (int g = 0; g < PG; g++)
#pragma unroll
                                for (int t = 0; t < NS; t++) {
                                    if ((unsigned)(w[g][t][0] >> 32) != tag || (unsigned)(w[g][t][1] >> 32) != tag) ll_load2(pb[g] + t * SK_ATT_STRIDE + dd[g], w[g][t][0], w[g][t][1]);
                                    if ((unsigned)(w[g][t][2] >> 32) != tag || (unsigned)(w[g][t][3] >> 32) != tag) ll_load2(pb[g] + t * SK_ATT_STRIDE + 128, w[g][t][2], w[g][t][3]);
                                }
                        }
#pragma unroll
                        for (int g = 0; g < PG; g++) {
                            const int q = tid + (i0 + g) * SK_THREADS;
                            float M = -1e30f, Ls = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
                            for (int t = 0; t < NS; t++)
                                if (t < S) M = fmaxf(M, __uint_as_float((unsigned)w[g][t][2]));
#pragma unroll
                            for (int t = 0; t < NS; t++)
                                if (t < S) {
                                    const float e = expf(__uint_as_float((unsigned)w[g][t][2]) - M);
                                    Ls += __uint_as_float((unsigned)w[g][t][3]) * e;
                                    o0 += __uint_as_float((unsigned)w[g][t][0]) * e;
                                    o1 += __uint_as_float((unsigned)w[g][t][1]) * e;
                                }
                            const float invL = Ls > 0.0f ? 1.0f / Ls : 0.0f;
                            sk_put_pair<NSEQ>(sm_xf, q >> 10, q & 1023, o0 * invL, o1 * invL);
                        }
                    }
                };
                if (S == 1) merge_splits(std::integral_constant<int, 1>{}, std::integral_constant<int, 2 * NSEQ>{});
                else merge_splits(std::integral_constant<int, SK_ATT_MAXS>{}, std::integral_constant<int, 2>{});
                }
                sk_csync();
                mark();
                run_phase(H, 2048, [&](int s, int row, int r, auto &&rowsum) { ll_store(p.ll_xwo + s * H + row, sm_x[s * H + row] + rowsum(r), tag); });
            }
            mark();
            // ---------------- GU + SwiGLU: rows (2j, 2j+1) = (gate_j, up_j) are neighbours in a chunk
            stage_norm(p.ll_xwo, tag, g_post);
            mark();
            run_phase(2 * I, H, [&](int s, int row, int r, auto &&rowsum) {
                if (!(row & 1)) {
                    const float inv = norm_scale(s);
                    ll_store(p.ll_act + s * I + (row >> 1), silu(rowsum(r) * inv) * (rowsum(r + 1) * inv), tag);
                }
            });
            mark();
            // ---------------- DOWN
            {
                const int ip = I >> 1, total = NSEQ * ip; // the NSEQ activation vectors as one long vector
                constexpr int NPD = NSEQ == 4 ? 12 : 6;   // pairs per thread in flight per pass
#pragma unroll 1
                for (int base = 0; base < total; base += NPD * SK_THREADS) {
                    float v[NPD][2];
                    ll_gather_pairs<NPD>(p.ll_act + 2 * (size_t)base, min(total - base, NPD * SK_THREADS), tag, tid, v, top_up);
#pragma unroll
                    for (int i = 0; i < NPD; i++) {
                        const int q = base + tid + i * SK_THREADS;
                        if (q < total) {
                            const int s = NSEQ == 1 ? 0 : q / ip;
                            sk_put_pair<NSEQ>(sm_xf, s, q - s * ip, v[i][0], v[i][1]);
                        }
                    }
                }
                sk_csync();
                mark();
                run_phase(H, I, [&](int s, int row, int r, auto &&rowsum) { ll_store(p.ll_xdn + s * H + row, sm_x[s * H + row] + rowsum(r), tag); });
            }
            mark();
        }
        // ---------------- HEAD: greedy argmax over this CTA's vocab rows of the tied embedding, per sequence
        const unsigned htag = p.tag_base + (unsigned)(step * (L + 1) + L + 1);
        float2 g_fin[2];
        load_gamma(p.final_norm, g_fin);
        stage_norm(p.ll_xdn, htag - 1, g_fin); // argmax is invariant under the positive RMSNorm scale: not applied
        float bv = -1e30f;
        int bi = 0x7fffffff;
        if (p.dbg_hidden && b == 0) // test hook: the post-final-norm hidden state of this step (reference qwen_asr_decoder.c:683,781)
            for (int e = tid; e < NSEQ * H; e += SK_THREADS) p.dbg_hidden[e] = sm_x[e] * norm_scale(e / H) * __ldg(p.final_norm + (e % H));
        run_phase(p.V, H, [&](int s, int row, int r, auto &&rowsum) {
            const float y = rowsum(r);
            if (p.dbg_logits) p.dbg_logits[(size_t)s * p.V + row] = y * norm_scale(s); // test hook: full logits of the kernel the product runs
            if (sk_better(y, row, bv, bi)) { bv = y; bi = row; }
        });
        // epilogue thread t holds the winner of sequence t / CH_ROWS over rows == t (mod CH_ROWS): reduce per sequence
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(QASR_FULL, bv, o);
            const int oi = __shfl_xor_sync(QASR_FULL, bi, o);
            if (sk_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        sk_csync();
        if (lane == 0) { sm_red[warp] = bv; sm_redi[warp] = bi; }
        sk_csync();
        __threadfence(); // KV rows appended this step become visible device-wide before the token exchange
        if (tid < NSEQ) {
            constexpr int WPS = CH_ROWS / 32; // epilogue warps per sequence (4, 2, 1)
            float v = sm_red[tid * WPS];
            int ix = sm_redi[tid * WPS];
#pragma unroll
            for (int w = 1; w < WPS; w++)
                if (sk_better(sm_red[tid * WPS + w], sm_redi[tid * WPS + w], v, ix)) { v = sm_red[tid * WPS + w]; ix = sm_redi[tid * WPS + w]; }
            ll_store(p.ll_head + (size_t)tid * 2048 + 4 * b, v, htag);
            ll_store_u32(p.ll_head + (size_t)tid * 2048 + 4 * b + 1, (unsigned)ix, htag);
        }
        int toks[NSEQ];
#pragma unroll 1
        for (int s = 0; s < NSEQ; s++) { // every CTA reduces the G winners of every sequence identically
            float wv = -1e30f;
            int wi = 0x7fffffff;
            {
                const bool active = tid < G;
                const u64 *hp2 = p.ll_head + (size_t)s * 2048 + 4 * tid; // one 32-byte sector per CTA: two writers never share a sector
                u64 w0 = (u64)htag << 32, w1 = (u64)htag << 32;
                if (active) ll_load2(hp2, w0, w1);
                for (;;) {
                    const bool ok = (unsigned)(w0 >> 32) == htag && (unsigned)(w1 >> 32) == htag;
                    if (__all_sync(QASR_FULL, ok)) break;
                    top_up();
                    if (!ok) ll_load2(hp2, w0, w1);
                }
                if (active) { wv = __uint_as_float((unsigned)w0); wi = (int)(unsigned)w1; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(QASR_FULL, wv, o);
                const int oi = __shfl_xor_sync(QASR_FULL, wi, o);
                if (sk_better(ov, oi, wv, wi)) { wv = ov; wi = oi; }
            }
            sk_csync();
            if (lane == 0) { sm_red[warp] = wv; sm_redi[warp] = wi; }
            sk_csync();
            wv = sm_red[0]; wi = sm_redi[0];
#pragma unroll
            for (int w = 1; w < SK_WARPS; w++)
                if (sk_better(sm_red[w], sm_redi[w], wv, wi)) { wv = sm_red[w]; wi = sm_redi[w]; }
#pragma unroll
            for (int t = 0; t < NSEQ; t++) if (t == s) toks[t] = wi;
        }
        __threadfence();
#pragma unroll
        for (int s = 0; s < NSEQ; s++) {
            const int tok = toks[s];
            pos[s]++;
            // next input row: exact bf16 -> f32 upcast of the embedding (reference qwen_asr.c:412-419,816)
            for (int e = tid; e < H; e += SK_THREADS) sm_x[s * H + e] = __uint_as_float(((uint32_t)p.emb[(size_t)tok * H + e]) << 16);
            if (b == 0 && tid == 0) {
                p.d_tokens[step * NSEQ + s] = tok;
                if (p.h_tokens) p.h_tokens[step * NSEQ + s] = tok;
            }
            if (tok == 151643 || tok == 151645) done_mask |= 1u << s; // reference qwen_asr.c:792
        }
        stop = done_mask == (1u << NSEQ) - 1u;
        sk_csync();
        mark();
    }
    if (b == 0) {
        for (int e = tid; e < NSEQ * H; e += SK_THREADS) p.x_io[e] = sm_x[e];
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < NSEQ; s++) p.d_pos[s] = pos[s];
            *p.d_step = step;
        }
    }
    // drain copies that were prefetched past an early stop before the CTA's shared memory goes away
    asm volatile("cp.async.wait_all;" ::: "memory");
}

// ---- host side ---------------------------------------------------------------------------------
static char g_sk_err[256] = "";
const char *stream_error(void) { return g_sk_err; }
static int g_sk_grid_dev[32] = {}; // CTAs of the decode kernel per device (0 = not initialised there): the shared-memory opt-in belongs to the (function, device) pair

template <int NSEQ, int SLOTS>
static cudaError_t sk_prepare_variant(int *per_sm, int kmax, int hmax) {
    const size_t smem = SkLayout<NSEQ, SLOTS>::total(kmax, hmax);
    cudaError_t e = cudaFuncSetAttribute(decode_stream_kernel<NSEQ, SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, decode_stream_kernel<NSEQ, SLOTS>, SK_THREADS, smem);
    return e;
}

int stream_init(void) {
    int dev = 0, sms = 0, coop = 0, per_sm = 0;
    cudaGetDevice(&dev);
    if (g_sk_grid_dev[dev & 31]) return 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaError_t e = sk_prepare_variant<1, 5>(&per_sm, SK_MAX_K, SK_MAX_H);
    int p2 = 0;
    if (e == cudaSuccess && per_sm >= 1) e = sk_prepare_variant<2, 4>(&p2, SK_MAX_K, SK_MAX_H);
    if (e == cudaSuccess && per_sm >= 1) e = sk_prepare_variant<4, 4>(&p2, 3072, 1024); // 4 sequences only fit for the 0.6B dims
    if (e != cudaSuccess || !coop || per_sm < 1 || p2 < 1 || sms < 1) {
        snprintf(g_sk_err, sizeof g_sk_err, "decode stream kernel unavailable: %s (coop=%d, blocks/SM=%d/%d)", cudaGetErrorString(e), coop, per_sm, p2);
        cudaGetLastError();
        return -1;
    }
    // Fewer CTAs = cheaper exchanges and fewer idle pollers, less streaming headroom.  From 128 CTAs up the busiest CTA of every phase has
    // the same number of rounds for both models (256 QKV groups, 64 / 128 output groups, 384 / 768 gate-up groups: ceil(groups / G) does
    // not change between 128 and 148), so the extra CTAs only add exchange traffic: measured on B200 (profiles/r02_decode_latency.txt)
    // 1.7B 638.7 us per token at 148 CTAs, 616-618 at 130-134, 623 at 128, 713 at 126 (a seventh gate-up group); 0.6B 385 -> 383.
    // At least 16 CTAs per sequence: the attention roles are (sequence, head, split) CTAs and S = min(S, G / (16 * NSEQ)) must stay >= 1.
    int grid = sms >= 132 ? 132 : sms;
    { const char *e = getenv("QASR_SK_GRID"); if (e && atoi(e) >= 16 * QASR_STREAM_MAX_SEQS && atoi(e) <= sms) grid = atoi(e); }
    if (grid < 16 * QASR_STREAM_MAX_SEQS) {
        snprintf(g_sk_err, sizeof g_sk_err, "decode stream kernel needs at least %d SMs (device has %d)", 16 * QASR_STREAM_MAX_SEQS, sms);
        return -1;
    }
    g_sk_grid_dev[dev & 31] = grid;
    return 0;
}
int stream_grid(void) {
    if (stream_init() != 0) return 0;
    int dev = 0;
    cudaGetDevice(&dev);
    return g_sk_grid_dev[dev & 31];
}

static bool sk_dims_ok(int L, int H, int I, int V) {
    return L >= 1 && L <= 28 && (H == 1024 || H == 2048) && I % 1024 == 0 && I <= SK_MAX_K && V % 16 == 0;
}

// Size of the image and the per-CTA byte offsets (host array of grid+1 entries).
size_t stream_image_layout(int L, int H, int I, int V, unsigned long long *cta_off_host) {
    if (stream_init() != 0 || !sk_dims_ok(L, H, I, V)) return 0;
    SkDims d{L, H, I, V, stream_grid()};
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
                       const unsigned long long *d_cta_off, uint8_t *image, int round_major) {
    if (stream_init() != 0) return -1;
    if (!sk_dims_ok(L, H, I, V)) { snprintf(g_sk_err, sizeof g_sk_err, "decode stream kernel: unsupported dims L=%d H=%d I=%d V=%d", L, H, I, V); return -1; }
    RetileArgs a;
    for (int i = 0; i < 4 * L; i++) a.src[i] = layer_mats[i];
    a.src[4 * L] = emb;
    SkDims d{L, H, I, V, stream_grid()};
    sk_retile_kernel<<<dim3(d.G, 4 * L + 1), SK_THREADS, 0, s>>>(a, d, d_cta_off, image, round_major);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(g_sk_err, sizeof g_sk_err, "re-tile launch: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}

// Which single-sequence kernel: the producer / consumer kernel over round-major weights (qasr_stream_r.cu) or the per-lane
// cp.async ring kernel (which still serves 2 and 4 sequences per launch).  QASR_DECODE_KERNEL=ring | rounds; default rounds.
bool stream_use_rounds(int H) {
    static int mode = -1; // 0 = default, 1 = ring, 2 = rounds
    if (mode < 0) { const char *e = getenv("QASR_DECODE_KERNEL"); mode = e && !strcmp(e, "ring") ? 1 : (e && !strcmp(e, "rounds") ? 2 : 0); }
    (void)H;
    return mode != 1; // measured: 0.6B 455 -> 387 us per token, 1.7B 709 -> 641 (profiles/r02_decode_latency.txt)
}

// Largest number of sequences one launch can carry for these dims (shared memory: the DOWN input image is K x 4 B per sequence)
int stream_max_seqs(int H, int I) { return (H <= 1024 && I <= 3072) ? 4 : 2; }

int launch_decode_stream(cudaStream_t s, const StreamParams &p) {
    if (stream_init() != 0) return -1;
    if (!sk_dims_ok(p.n_layers, p.H, p.I, p.V) || p.n_steps > 64 || p.n_steps < 1 || (p.nseq != 1 && p.nseq != 2 && p.nseq != 4) ||
        p.nseq > stream_max_seqs(p.H, p.I)) {
        snprintf(g_sk_err, sizeof g_sk_err, "decode stream kernel: unsupported dims H=%d I=%d V=%d steps=%d seqs=%d", p.H, p.I, p.V, p.n_steps, p.nseq);
        return -1;
    }
    void *args[] = {(void *)&p};
    const int grid = stream_grid();
    const int kmax = p.I > 2048 ? p.I : 2048;
    if (p.nseq == 1 && p.image_r && stream_use_rounds(p.H)) return launch_decode_rounds(s, p, grid, g_sk_err, sizeof g_sk_err);
    cudaError_t e;
    if (p.nseq == 1)
        e = cudaLaunchCooperativeKernel((const void *)decode_stream_kernel<1, 5>, dim3(grid), dim3(SK_THREADS), args, SkLayout<1, 5>::total(kmax, p.H), s);
    else if (p.nseq == 2)
        e = cudaLaunchCooperativeKernel((const void *)decode_stream_kernel<2, 4>, dim3(grid), dim3(SK_THREADS), args, SkLayout<2, 4>::total(kmax, p.H), s);
    else
        e = cudaLaunchCooperativeKernel((const void *)decode_stream_kernel<4, 4>, dim3(grid), dim3(SK_THREADS), args, SkLayout<4, 4>::total(kmax, p.H), s);
    if (e != cudaSuccess) {
        snprintf(g_sk_err, sizeof g_sk_err, "decode stream kernel launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

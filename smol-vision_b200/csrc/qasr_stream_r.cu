// qasr_stream_r.cu - single-sequence decode kernel, producer / consumer form: one warp streams the weights with 1-D TMA
// bulk copies into a shared-memory ring of whole "rounds", eight warps consume them.
// Same job, same schedule, same exchanges and the same arithmetic per product as decode_stream_kernel<1, 5>
// (qasr_stream.cu; reference hot loop qwen_asr.c:788-818 -> qwen_decoder_forward, qwen_asr_decoder.c:592-685).
//
// What was measured on the per-lane cp.async ring (profiles/r02_decode_latency.txt): a phase of the 0.6B model gives a
// warp only 2-3 units, and between two exchanges a unit costs ~0.3 us, not because the data is late but because the
// consumers themselves issue the loads: an LDGSTS / LDG burst is accepted by the SM's memory pipeline at ~57 B/clk, the
// issuing warp is blocked for 0.13 us per unit, the LDS of the next unit queues behind it, and (register window variant)
// scoreboards shared between old and fresh loads make the next MMA wait for the refill.  28 x 10 such rounds per token.
// Here no consumer ever issues a weight load:
//  * the image is laid out round-major: a round = the 16 units (16 rows x 1024 columns, 32 KB) that the consumer warps
//    need together, contiguous in HBM in consumption order of the CTA;
//  * the producer warp (one elected lane) keeps a short ring of rounds in flight (3 slots for the 0.6B dims, 5 for 1.7B): mbarrier
//    expect_tx + the round as four or eight cp.async.bulk copies (1-D TMA, L2 evict-first), and a cp.async.bulk.prefetch.L2 a
//    few rounds further ahead; it waits on the `empty` barrier of a slot and on nothing else;
//  * consumer warp w reads its 4 KB of a round (column slices w + 16 j and w + 8 + 16 j of the same 16 rows) with eight
//    conflict-free LDS.128 after an mbarrier wait that has normally long completed, runs 8 MMAs on two independent
//    accumulator chains and releases the slot with one arrive per warp.
// 9 - 12 warps put three warps on one scheduler: the kernel is held to 168 registers and must not spill (with spills, a
// larger ring - i.e. a smaller L1 - was measurably slower; tests/test_abi.py checks registers and stack).
#include "qasr_stream_common.cuh"

#define SR_WARPS 8            /* consumer warps */
#define SR_THREADS (SR_WARPS * 32)
#define SR_MAX_PRODUCERS 4
#ifndef SR_DEFAULT_PRODUCERS
#define SR_DEFAULT_PRODUCERS 1
#endif
#define SR_ALL_THREADS (SR_THREADS + 32 * SR_MAX_PRODUCERS) /* launch bound: consumers + up to 4 producer warps (12 warps = 3 per scheduler => 168 registers) */
#define SR_ROUND (16 * SK_UNIT) /* 32 KB: 16 rows x 1024 columns */
#define SR_MAX_SLOTS 6
#define SR_PSTRIDE 9
#define SR_ATT_BATCH 4        /* cached keys per warp and batch: 8 warps x 4 = 32 keys per split, as in the ring kernel */
#define SR_ATT_STRIDE SK_ATT_STRIDE
#define SR_HEAD_STRIDE 4      /* words per CTA in the head exchange (one sector) */
#define SR_CHAINS 2           /* independent MMA accumulator chains per round (4 chains, all 16 operand loads of a round in flight, __expf in the softmax: all within noise) */

__device__ __forceinline__ void sr_csync() { asm volatile("bar.sync 1, %0;" ::"n"(SR_THREADS) : "memory"); }
__device__ __forceinline__ void sr_mma(float (&c)[4], const uint4 a, const uint2 b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ void sr_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void sr_mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void sr_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool sr_mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void sr_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, u64 pol) { // weights are read once per token: L2 evict-first
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void sr_bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

struct SrLayout { // dynamic shared memory: ring | xf | x | partial | small | barriers
    static __host__ __device__ size_t ring_bytes(int nslot) { return (size_t)nslot * SR_ROUND; }
    static __host__ __device__ size_t xf_bytes(int kmax) { return (size_t)kmax * 4; } // [kb][8 lanes][2] u32
    static __host__ __device__ size_t x_bytes(int H) { return (size_t)H * 4; }
    static constexpr size_t partial_bytes = (size_t)2 * 128 * SR_PSTRIDE * 4;
    static constexpr size_t small_bytes = (64 + 16 + 16 + 4) * 4 + 2 * SR_MAX_SLOTS * 8;
    static __host__ __device__ size_t rest(int kmax, int H) { return xf_bytes(kmax) + x_bytes(H) + partial_bytes + small_bytes + 256; }
    static __host__ __device__ size_t total(int nslot, int kmax, int H) { return ring_bytes(nslot) + rest(kmax, H); }
};

template <int NPX> // pairs of the hidden vector per consumer thread: H = 512 NPX
__global__ void __launch_bounds__(SR_ALL_THREADS, 1) decode_rounds_kernel(const StreamParams p, const int nslot) {
    extern __shared__ __align__(128) uint8_t sr_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int b = blockIdx.x, G = gridDim.x;
    const int L = p.n_layers, H = p.H, I = p.I;
    const int kmax = max(max(H, I), 2048);
    uint8_t *sm_ring = sr_raw;
    uint8_t *sm_rest = sr_raw + SrLayout::ring_bytes(nslot);
    uint32_t *sm_xf = reinterpret_cast<uint32_t *>(sm_rest);
    float *sm_x = reinterpret_cast<float *>(sm_rest + SrLayout::xf_bytes(kmax));
    float(*sm_partial)[128][SR_PSTRIDE] = reinterpret_cast<float(*)[128][SR_PSTRIDE]>(sm_rest + SrLayout::xf_bytes(kmax) + SrLayout::x_bytes(H));
    float *sm_red = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(sm_partial) + SrLayout::partial_bytes); // [64]
    float *sm_ssq = sm_red + 64;                                                                                // [16]
    int *sm_redi = reinterpret_cast<int *>(sm_ssq + 16);                                                        // [16]
    volatile unsigned *sm_ctl = reinterpret_cast<volatile unsigned *>(sm_redi + 16);                            // [0] stop flag, [1] rounds consumed at the stop
    const uint32_t bar_full = sk_smem_u32(const_cast<unsigned *>(sm_ctl) + 4), bar_empty = bar_full + SR_MAX_SLOTS * 8;

    if (tid == 0) {
        for (int i = 0; i < nslot; i++) { sr_mbar_init(bar_full + 8 * i, 1); sr_mbar_init(bar_empty + 8 * i, SR_WARPS); }
        sm_ctl[0] = 0; sm_ctl[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = tid; e < H; e += blockDim.x) sm_x[e] = p.x_io[e];
    __syncthreads();

    // ---- producer warps: rounds of this CTA in consumption order, cyclic over the steps of the launch.  A round is issued as
    // SR_ROUND / chunk bulk copies (measured: 8 KB chunks for the 1.7B dims, 4 KB for 0.6B - one 32 KB copy per round is slower,
    // 2 KB copies are issue-bound); with n_prod producer warps, producer q issues chunks q, q + n_prod, ... of every round, so the
    // ~0.1 us of issue time per copy is shared.  Producer 0 arms the round's `full` barrier (expect_tx of the whole round; the
    // transaction count may go negative until it does) and runs the L2 prefetch cursor.
    const u64 coff = p.cta_off[b];
    const uint32_t rounds_per_step = (uint32_t)((p.cta_off[b + 1] - coff) / SR_ROUND);
    if (warp >= SR_WARPS) {
        if (lane == 0) {
            const uint32_t q = (uint32_t)(warp - SR_WARPS), n_prod = (uint32_t)(blockDim.x >> 5) - SR_WARPS;
            const uint8_t *src = p.image_r + coff;
            uint32_t total = rounds_per_step * (uint32_t)p.n_steps;
            u64 pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            const uint32_t ahead = q == 0 ? (uint32_t)min(p.l2_ahead_units, 64) : 0u;
            for (uint32_t r = 0; r < ahead && r < rounds_per_step; r++) sr_bulk_prefetch_l2(src + (size_t)r * SR_ROUND, SR_ROUND);
            uint32_t slot = 0, pass = 0, rr = 0, pr = ahead % rounds_per_step; // ring slot, ring pass, round within the step, prefetch cursor
            // the rounds of the lm_head phase close every step: no exchange waits there, so a copy costs its ~0.1 us of issue time and
            // nothing else - with the 4 KB copies that suit the layer phases of the 0.6B dims the head was issue-bound (8 copies per round)
            const uint32_t head_from = rounds_per_step - (uint32_t)((sk_g0(p.V >> 4, b + 1, G) - sk_g0(p.V >> 4, b, G)) * (p.H >> 10));
            bool stopping = false;
            uint32_t drain_from = 0;
            for (uint32_t issued = 0; issued < total; issued++) {
                if (pass > 0) // the slot must have been drained by all consumer warps
                    while (!sr_mbar_try(bar_empty + 8 * slot, (pass - 1) & 1)) {
                        if (!stopping && sm_ctl[0]) {
                            // Early stop (EOS): the consumers drained `sm_ctl[1]` rounds and take no more.  EVERY producer still issues
                            // its share of the rounds whose slots are free (< drained + nslot: their `empty` phases are complete, so
                            // nobody blocks), which keeps the barriers of those rounds consistent among the producers; then it waits
                            // for them to land - no bulk copy may be in flight when the CTA's shared memory goes away.
                            stopping = true;
                            drain_from = sm_ctl[1];
                            total = min(total, drain_from + (uint32_t)nslot);
                        }
                        if (issued >= total) break;
                    }
                if (issued >= total) break;
                if (q == 0) sr_mbar_expect_tx(bar_full + 8 * slot, SR_ROUND);
                const uint32_t dst = sk_smem_u32(sm_ring + (size_t)slot * SR_ROUND);
                const uint8_t *g = src + (size_t)rr * SR_ROUND;
                const uint32_t chunk = (uint32_t)(rr >= head_from ? p.sr_chunk_head : p.sr_chunk), chunks_per_round = SR_ROUND / chunk;
                for (uint32_t c = q; c < chunks_per_round; c += n_prod) sr_bulk_load(dst + c * chunk, g + (size_t)c * chunk, chunk, bar_full + 8 * slot, pol);
                if (ahead) { sr_bulk_prefetch_l2(src + (size_t)pr * SR_ROUND, SR_ROUND); if (++pr == rounds_per_step) pr = 0; }
                if (++rr == rounds_per_step) rr = 0;
                if (++slot == (uint32_t)nslot) { slot = 0; pass++; }
            }
            if (stopping)
                for (uint32_t k = drain_from; k < total; k++)
                    while (!sr_mbar_try(bar_full + 8 * (k % (uint32_t)nslot), (k / (uint32_t)nslot) & 1)) {}
        }
        return;
    }
    uint32_t cslot = 0, cpar = 0, consumed = 0; // consumer cursor: ring slot, parity of its full barrier, rounds consumed
    auto no_svc = []() {}; // stall-driven L2 prefetch from the polling warps (the ring kernel's hook) was measured here too: no gain, dropped

    long long *prof = (p.prof && (b == 0 || b == G - 1) && tid == 0) ? p.prof + (b == 0 ? 0 : p.prof_cap) : nullptr;
    int prof_n = 0;
    auto mark = [&]() { if (prof && prof_n < p.prof_cap) prof[prof_n++] = clock64(); };
#ifdef SR_FINE_PROF // -DSR_FINE_PROF: finer stamps for tools/mega_prof_fine.py (costs registers: not in the product build)
    const bool fine = (p.debug & 128) != 0; // extra stamps inside every layer phase: MMAs done | after bar.sync | after the epilogue
    bool finer = (p.debug & 256) != 0;      // + per round: A fragments requested | MMAs issued
#else
    constexpr bool fine = false;
    bool finer = false;
#endif
    // one round for this warp: 16 rows x (64 + 64) columns against the phase input, two independent accumulator chains
    auto do_round = [&](int j, float(&c)[SR_CHAINS][4]) {
        const uint2 *xb = reinterpret_cast<const uint2 *>(sm_xf) + (size_t)(warp + 16 * j) * 32 + (lane & 7);
        uint2 b0[4];
#pragma unroll
        for (int kb = 0; kb < 4; kb++) b0[kb] = xb[kb * 8]; // lanes >= 8 feed D columns nobody reads
        while (!sr_mbar_try(bar_full + 8 * cslot, cpar)) {}
        const uint4 *ring = reinterpret_cast<const uint4 *>(sm_ring + (size_t)cslot * SR_ROUND + (size_t)warp * (2 * SK_UNIT)) + lane;
        {
            uint4 a0[4];
#pragma unroll
            for (int kb = 0; kb < 4; kb++) a0[kb] = ring[kb * 32];
            if (finer) mark();
#pragma unroll
            for (int kb = 0; kb < 4; kb++) sr_mma(c[kb & (SR_CHAINS - 1)], a0[kb], b0[kb]);
        }
        { // second half (column slice warp + 8 + 16 j): the operand registers of the first half are free again
            uint4 a1[4];
            uint2 b1[4];
#pragma unroll
            for (int kb = 0; kb < 4; kb++) { b1[kb] = xb[SR_WARPS * 32 + kb * 8]; a1[kb] = ring[128 + kb * 32]; }
#pragma unroll
            for (int kb = 0; kb < 4; kb++) sr_mma(c[kb & (SR_CHAINS - 1)], a1[kb], b1[kb]);
        }
        __syncwarp();
        if (lane == 0) sr_mbar_arrive(bar_empty + 8 * cslot); // operands were read at issue: the slot may be refilled
        if (finer) mark();
        consumed++;
        if (++cslot == (uint32_t)nslot) { cslot = 0; cpar ^= 1; }
    };

    // ---- one weighted phase: y[row] = W[row,:] . x for the CTA's rows [16 g0, 16 g1), handed to epi(row, r, y, active) in
    // chunks of <= 128 rows (warps 0-3, thread = row; every lane of those warps calls epi so that it may shuffle)
    float pend_ss = 0.0f;   // this thread's share of sum(x^2) of the staged input: reduced while the MMAs are in flight
    bool have_ss = false;
    auto flush_ss = [&]() {
        if (have_ss) {
            const float t = warp_sum(pend_ss);
            if (lane == 0) sm_ssq[warp] = t;
            have_ss = false;
        }
    };
    int pbuf = 0;
    auto run_phase = [&](int g0, int g1, int nj, bool layer_phase, auto &&epi) {
        for (int cg0 = g0; cg0 < g1; cg0 += SK_CHUNK_GROUPS) {
            const int cg1 = min(cg0 + SK_CHUNK_GROUPS, g1);
            float(*part)[SR_PSTRIDE] = sm_partial[pbuf];
            for (int grp = cg0; grp < cg1; grp++) {
                float c[SR_CHAINS][4];
#pragma unroll
                for (int i = 0; i < SR_CHAINS; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0f;
                for (int j = 0; j < nj; j++) do_round(j, c);
                if (grp + 1 == cg1) flush_ss(); // shuffles overlap the latency of the last MMAs
                if (tig == 0) { // D columns 0 / 1 = x_hi / x_lo sums; rows gid and gid + 8
                    const int r = (grp - cg0) * 16 + gid;
                    part[r][warp] = (c[0][0] + c[1][0]) + (c[0][1] + c[1][1]);
                    part[r + 8][warp] = (c[0][2] + c[1][2]) + (c[0][3] + c[1][3]);
                }
            }
            if (fine && layer_phase) mark();
            sr_csync();
            if (fine && layer_phase) mark();
            if (tid < 128) {
                float y = 0.0f;
#pragma unroll
                for (int i = 0; i < SR_WARPS; i++) y += part[tid][i];
                epi(cg0 * 16 + tid, tid, y, tid < (cg1 - cg0) * 16);
            }
            if (fine && layer_phase) mark();
            pbuf ^= 1;
        }
        flush_ss();       // CTAs without rows in this phase
    };

    // x (shared, or gathered from an exchange buffer first) -> x * gamma -> B-fragment image; the RMSNorm scalar is
    // applied to the phase output (norm_scale()), so the reduction of squares leaves the critical path
    auto stage_norm = [&](const u64 *src, unsigned tag, const float2(&gm2)[NPX]) {
        float v[NPX][2];
        const int hp = H >> 1;
        if (src) {
            ll_gather_pairs<NPX, SR_THREADS>(src, hp, tag, tid, v, no_svc);
#pragma unroll
            for (int i = 0; i < NPX; i++) {
                const int q = tid + i * SR_THREADS;
                if (q < hp) *reinterpret_cast<float2 *>(sm_x + 2 * q) = make_float2(v[i][0], v[i][1]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < NPX; i++) {
                const int q = tid + i * SR_THREADS;
                const float2 t = q < hp ? *reinterpret_cast<const float2 *>(sm_x + 2 * q) : make_float2(0.f, 0.f);
                v[i][0] = t.x; v[i][1] = t.y;
            }
        }
        float ss = 0.0f;
#pragma unroll
        for (int i = 0; i < NPX; i++) {
            const int q = tid + i * SR_THREADS;
            if (q < hp) {
                ss += fmaf(v[i][0], v[i][0], v[i][1] * v[i][1]);
                sk_put_pair<1>(sm_xf, 0, q, v[i][0] * gm2[i].x, v[i][1] * gm2[i].y);
            }
        }
        pend_ss = ss;
        have_ss = true;
        sr_csync();
    };
    const float inv_H = 1.0f / (float)H;
    auto norm_scale = [&]() { // valid after the bar.sync of the phase that follows stage_norm
        const float t = ((sm_ssq[0] + sm_ssq[1]) + (sm_ssq[2] + sm_ssq[3])) + ((sm_ssq[4] + sm_ssq[5]) + (sm_ssq[6] + sm_ssq[7]));
        return rsqrtf(fmaf(t, inv_H, p.eps));
    };
    auto load_gamma = [&](const float *g, float2(&gm2)[NPX]) {
#pragma unroll
        for (int j = 0; j < NPX; j++) {
            const int pr = tid + j * SR_THREADS;
            gm2[j] = pr < (H >> 1) ? __ldg(reinterpret_cast<const float2 *>(g) + pr) : make_float2(0.f, 0.f);
        }
    };

    // rows of this CTA per phase (16-row groups), fixed for the whole launch
    const int gq0 = sk_g0(4096 >> 4, b, G), gq1 = sk_g0(4096 >> 4, b + 1, G), gh0 = sk_g0(H >> 4, b, G), gh1 = sk_g0(H >> 4, b + 1, G);
    const int gg0 = sk_g0((2 * I) >> 4, b, G), gg1 = sk_g0((2 * I) >> 4, b + 1, G), gv0 = sk_g0(p.V >> 4, b, G), gv1 = sk_g0(p.V >> 4, b + 1, G);
    const int njH = H >> 10, njI = I >> 10;
    const size_t kvd = 1024;
    const float scale = 0.08838834764831845f; // 1/sqrtf(128)
    int pos = p.d_pos[0];
    int step = 0;
    bool stop = false;

    for (; step < p.n_steps && !stop; step++) {
        // attention role of this CTA for the whole step: (q head hq, key split sp of S)
        int S = (pos + 1 + SR_ATT_BATCH * SR_WARPS - 1) / (SR_ATT_BATCH * SR_WARPS);
        S = min(S, min(SK_ATT_MAXS, G / 16));
        const bool att = b < 16 * S;
        const int hq = b / S, sp = b - hq * S, hkv = hq >> 1;
        const int apos = pos, n_keys = apos + 1;
        const int per = (n_keys + S - 1) / S;
        const int k0 = sp * per, k1 = min(n_keys, k0 + per);
        const float4 rope_c = __ldg(reinterpret_cast<const float4 *>(p.rope_cos + (size_t)apos * 64) + (lane & 15));
        const float4 rope_s = __ldg(reinterpret_cast<const float4 *>(p.rope_sin + (size_t)apos * 64) + (lane & 15));
        for (int l = 0; l < L; l++) {
            const unsigned tag = p.tag_base + (unsigned)(step * (L + 1) + l + 1);
            float *kc = p.kv_k[0] + (size_t)l * p.kv_layer_stride, *vc = p.kv_v[0] + (size_t)l * p.kv_layer_stride;
            if (att) // K/V rows of this split -> L2 while the QKV phase runs
                for (int j = k0 * 8 + tid; j < k1 * 8 && j < apos * 8; j += SR_THREADS) { // 8 lines of 128 B per key (K row + V row)
                    const float *row = ((j & 4) ? vc : kc) + (size_t)(j >> 3) * kvd + hkv * 128 + (j & 3) * 32;
                    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(row) : "memory");
                }
            float2 g_in[NPX], g_post[NPX];
            load_gamma(p.in_norm[l], g_in);
            mark();
            // ---------------- QKV
            stage_norm(l > 0 ? p.ll_xdn : nullptr, tag - 1, g_in);
            mark();
            run_phase(gq0, gq1, njH, true, [&](int row, int r, float y, bool active) { if (active) ll_store(p.ll_qkv + row, y * norm_scale(), tag); });
            mark();
            // ---------------- ATTN (attention CTAs only); warp-local up to the final merge, as in the ring kernel
            if (att) {
                float *sc = reinterpret_cast<float *>(sm_xf); // the QKV input image is dead: attention scratch
                float *wacc = sc, *wml = sc + SR_WARPS * 128;
                const float4 qn4 = __ldg(reinterpret_cast<const float4 *>(p.qn[l]) + lane), kn4 = __ldg(reinterpret_cast<const float4 *>(p.kn[l]) + lane);
                const bool has_new = (k1 == n_keys) && (k0 < k1);
                const int w_new = has_new ? ((apos - k0) & (SR_WARPS - 1)) : -1; // warp whose key list contains `apos`
                const size_t hoff = (size_t)hkv * 128 + lane * 4;
                float4 kr[SR_ATT_BATCH], vr[SR_ATT_BATCH];
#pragma unroll
                for (int i = 0; i < SR_ATT_BATCH; i++) {
                    const int j = k0 + warp + SR_WARPS * i;
                    if (j < k1 && j != apos) {
                        kr[i] = __ldcg(reinterpret_cast<const float4 *>(kc + (size_t)j * kvd + hoff));
                        vr[i] = __ldcg(reinterpret_cast<const float4 *>(vc + (size_t)j * kvd + hoff));
                    }
                }
                u64 wq[4], wk[4], wv[4];
                const u64 *pq = p.ll_qkv + hq * 128 + lane * 4, *pk = p.ll_qkv + 2048 + hkv * 128 + lane * 4, *pv = pk + 1024;
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
                auto norm_rope = [&](const u64(&w)[4], const float4 nw) {
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
                        for (int i = 0; i < SR_ATT_BATCH; i++) if (k0 + warp + SR_WARPS * i == apos) { kr[i] = k_new; vr[i] = v_new; }
                        if (!(hq & 1)) { // one writer per kv head appends the new row (reference qwen_asr_decoder.c:640-646)
                            *reinterpret_cast<float4 *>(kc + (size_t)apos * kvd + hoff) = k_new;
                            *reinterpret_cast<float4 *>(vc + (size_t)apos * kvd + hoff) = v_new;
                        }
                    }
                }
                mark();
                float m = -1e30f, lsum = 0.0f;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int base = k0; base < k1; base += SR_ATT_BATCH * SR_WARPS) {
                    if (base != k0) { // later batches (only when a split holds more than 32 keys)
#pragma unroll
                        for (int i = 0; i < SR_ATT_BATCH; i++) {
                            const int j = base + warp + SR_WARPS * i;
                            if (j < k1) {
                                if (j == apos) { kr[i] = k_new; vr[i] = v_new; } // never read the row being appended from the cache
                                else {
                                    kr[i] = __ldcg(reinterpret_cast<const float4 *>(kc + (size_t)j * kvd + hoff));
                                    vr[i] = __ldcg(reinterpret_cast<const float4 *>(vc + (size_t)j * kvd + hoff));
                                }
                            }
                        }
                    }
                    float sc8[SR_ATT_BATCH];
#pragma unroll
                    for (int i = 0; i < SR_ATT_BATCH; i++) {
                        const int j = base + warp + SR_WARPS * i;
                        float d4 = 0.0f;
                        if (j < k1) d4 = q4v.x * kr[i].x + q4v.y * kr[i].y + q4v.z * kr[i].z + q4v.w * kr[i].w;
                        sc8[i] = warp_sum(d4) * scale;
                    }
#pragma unroll
                    for (int i = 0; i < SR_ATT_BATCH; i++) {
                        const int j = base + warp + SR_WARPS * i;
                        if (j < k1) {
                            const float s = sc8[i];
                            if (s > m) {
                                const float cc = expf(m - s);
                                lsum = lsum * cc + 1.0f;
                                acc.x = acc.x * cc + vr[i].x; acc.y = acc.y * cc + vr[i].y; acc.z = acc.z * cc + vr[i].z; acc.w = acc.w * cc + vr[i].w;
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
                sr_csync();
                if (tid < 128) { // merge the warps in fixed order
                    float M = -1e30f;
#pragma unroll
                    for (int w = 0; w < SR_WARPS; w++) M = fmaxf(M, wml[w * 2]);
                    float Ls = 0.0f, Aa = 0.0f;
#pragma unroll
                    for (int w = 0; w < SR_WARPS; w++) {
                        const float e = expf(wml[w * 2] - M);
                        Ls += wml[w * 2 + 1] * e;
                        Aa += wacc[w * 128 + tid] * e;
                    }
                    u64 *pb = p.ll_att + (size_t)(hq * SK_ATT_MAXS + sp) * SR_ATT_STRIDE;
                    ll_store(pb + tid, Aa, tag);
                    if (tid == 0) { ll_store(pb + 128, M, tag); ll_store(pb + 129, Ls, tag); }
                }
                sr_csync(); // scratch (xf) is rewritten by the WO staging below
            }
            mark();
            // ---------------- WO: input = attention output merged over the S key splits (1024 pairs, 4 per thread)
            load_gamma(p.post_norm[l], g_post);
            {
                // thread t: head t / 16, pairs (t % 16) + 16 i of that head: the S (m, l) words are loaded once per thread and
                // every exchange word of the thread is in flight before the first tag is checked (one L2 round trip)
                auto merge_splits = [&](auto ns_c) { // NS = compile-time bound on S
                    constexpr int NS = decltype(ns_c)::value;
                    const int hd = tid >> 4, p0 = tid & 15;
                    const u64 *hb = p.ll_att + (size_t)(hd * SK_ATT_MAXS) * SR_ATT_STRIDE;
                    u64 ml[NS][2], w[NS][4][2];
#pragma unroll
                    for (int t = 0; t < NS; t++) {
                        if (t < S) {
                            ll_load2(hb + t * SR_ATT_STRIDE + 128, ml[t][0], ml[t][1]);
#pragma unroll
                            for (int i = 0; i < 4; i++) ll_load2(hb + t * SR_ATT_STRIDE + 2 * (p0 + 16 * i), w[t][i][0], w[t][i][1]);
                        } else {
                            ml[t][0] = ml[t][1] = (u64)tag << 32;
#pragma unroll
                            for (int i = 0; i < 4; i++) w[t][i][0] = w[t][i][1] = (u64)tag << 32;
                        }
                    }
                    for (;;) {
                        bool ok = true;
#pragma unroll
                        for (int t = 0; t < NS; t++) {
                            ok = ok && (unsigned)(ml[t][0] >> 32) == tag && (unsigned)(ml[t][1] >> 32) == tag;
#pragma unroll
                            for (int i = 0; i < 4; i++) ok = ok && (unsigned)(w[t][i][0] >> 32) == tag && (unsigned)(w[t][i][1] >> 32) == tag;
                        }
                        if (__all_sync(QASR_FULL, ok)) break;
#pragma unroll
                        for (int t = 0; t < NS; t++) {
                            if ((unsigned)(ml[t][0] >> 32) != tag || (unsigned)(ml[t][1] >> 32) != tag) ll_load2(hb + t * SR_ATT_STRIDE + 128, ml[t][0], ml[t][1]);
#pragma unroll
                            for (int i = 0; i < 4; i++)
                                if ((unsigned)(w[t][i][0] >> 32) != tag || (unsigned)(w[t][i][1] >> 32) != tag) ll_load2(hb + t * SR_ATT_STRIDE + 2 * (p0 + 16 * i), w[t][i][0], w[t][i][1]);
                        }
                    }
                    float M = -1e30f, Ls = 0.f, e[NS];
#pragma unroll
                    for (int t = 0; t < NS; t++)
                        if (t < S) M = fmaxf(M, __uint_as_float((unsigned)ml[t][0]));
#pragma unroll
                    for (int t = 0; t < NS; t++) {
                        e[t] = t < S ? __expf(__uint_as_float((unsigned)ml[t][0]) - M) : 0.0f;
                        Ls += t < S ? __uint_as_float((unsigned)ml[t][1]) * e[t] : 0.0f;
                    }
                    const float invL = Ls > 0.0f ? __fdividef(1.0f, Ls) : 0.0f;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        float o0 = 0.f, o1 = 0.f;
#pragma unroll
                        for (int t = 0; t < NS; t++)
                            if (t < S) {
                                o0 = fmaf(__uint_as_float((unsigned)w[t][i][0]), e[t], o0);
                                o1 = fmaf(__uint_as_float((unsigned)w[t][i][1]), e[t], o1);
                            }
                        sk_put_pair<1>(sm_xf, 0, hd * 64 + p0 + 16 * i, o0 * invL, o1 * invL);
                    }
                };
                if (S == 1) merge_splits(std::integral_constant<int, 1>{});
                else if (S == 2) merge_splits(std::integral_constant<int, 2>{});
                else merge_splits(std::integral_constant<int, SK_ATT_MAXS>{});
                sr_csync();
                mark();
                run_phase(gh0, gh1, 2, true, [&](int row, int r, float y, bool active) { if (active) ll_store(p.ll_xwo + row, sm_x[row] + y, tag); });
            }
            mark();
            // ---------------- GU + SwiGLU: rows (2j, 2j+1) = (gate_j, up_j) are neighbours in a chunk
            stage_norm(p.ll_xwo, tag, g_post);
            mark();
            run_phase(gg0, gg1, njH, true, [&](int row, int r, float y, bool active) {
                const float u = __shfl_down_sync(QASR_FULL, y, 1); // up_j sits in the next row of the chunk
                if (active && !(row & 1)) {
                    const float inv = norm_scale(), g = y * inv;
                    ll_store(p.ll_act + (row >> 1), __fdividef(g, 1.0f + __expf(-g)) * (u * inv), tag);
                }
            });
            mark();
            // ---------------- DOWN
            {
                const int ip = I >> 1;
                constexpr int NPD = 6;
#pragma unroll 1
                for (int base = 0; base < ip; base += NPD * SR_THREADS) {
                    float v[NPD][2];
                    ll_gather_pairs<NPD, SR_THREADS>(p.ll_act + 2 * (size_t)base, min(ip - base, NPD * SR_THREADS), tag, tid, v, no_svc);
#pragma unroll
                    for (int i = 0; i < NPD; i++) {
                        const int q = base + tid + i * SR_THREADS;
                        if (q < ip) sk_put_pair<1>(sm_xf, 0, q, v[i][0], v[i][1]);
                    }
                }
                sr_csync();
                mark();
                run_phase(gh0, gh1, njI, true, [&](int row, int r, float y, bool active) { if (active) ll_store(p.ll_xdn + row, sm_x[row] + y, tag); });
            }
            mark();
        }
        // ---------------- HEAD: greedy argmax over this CTA's vocab rows of the tied embedding
        const unsigned htag = p.tag_base + (unsigned)(step * (L + 1) + L + 1);
        float2 g_fin[NPX];
        load_gamma(p.final_norm, g_fin);
        stage_norm(p.ll_xdn, htag - 1, g_fin); // argmax is invariant under the positive RMSNorm scale: not applied
        float bv = -1e30f;
        int bi = 0x7fffffff;
        const bool finer_saved = finer;
        finer = false;
        if (p.dbg_hidden) { // test hook: the post-final-norm hidden state of this step (reference qwen_asr_decoder.c:683,781)
            flush_ss();
            sr_csync();
            if (b == 0)
                for (int e = tid; e < H; e += SR_THREADS) p.dbg_hidden[e] = sm_x[e] * norm_scale() * __ldg(p.final_norm + e);
        }
        run_phase(gv0, gv1, njH, false, [&](int row, int r, float y, bool active) {
            if (active) {
                if (p.dbg_logits) p.dbg_logits[row] = y * norm_scale(); // test hook: full logits of the kernel the product runs
                if (sk_better(y, row, bv, bi)) { bv = y; bi = row; }
            }
        });
        finer = finer_saved;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(QASR_FULL, bv, o);
            const int oi = __shfl_xor_sync(QASR_FULL, bi, o);
            if (sk_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        sr_csync();
        if (lane == 0) { sm_red[warp] = bv; sm_redi[warp] = bi; }
        sr_csync();
        __threadfence(); // KV rows appended this step become visible device-wide before the token exchange
        if (tid == 0) {
            float v = sm_red[0];
            int ix = sm_redi[0];
#pragma unroll
            for (int w = 1; w < 4; w++) // epilogue threads are warps 0-3
                if (sk_better(sm_red[w], sm_redi[w], v, ix)) { v = sm_red[w]; ix = sm_redi[w]; }
            ll_store(p.ll_head + SR_HEAD_STRIDE * b, v, htag);
            ll_store_u32(p.ll_head + SR_HEAD_STRIDE * b + 1, (unsigned)ix, htag);
        }
        float wv = -1e30f;
        int wi = 0x7fffffff;
        {
            const bool active = tid < G;
            const u64 *hp2 = p.ll_head + SR_HEAD_STRIDE * tid;
            u64 x0 = (u64)htag << 32, x1 = (u64)htag << 32;
            if (active) ll_load2(hp2, x0, x1);
            for (;;) {
                const bool ok = (unsigned)(x0 >> 32) == htag && (unsigned)(x1 >> 32) == htag;
                if (__all_sync(QASR_FULL, ok)) break;
                if (!ok) ll_load2(hp2, x0, x1);
            }
            if (active) { wv = __uint_as_float((unsigned)x0); wi = (int)(unsigned)x1; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(QASR_FULL, wv, o);
            const int oi = __shfl_xor_sync(QASR_FULL, wi, o);
            if (sk_better(ov, oi, wv, wi)) { wv = ov; wi = oi; }
        }
        sr_csync();
        if (lane == 0) { sm_red[warp] = wv; sm_redi[warp] = wi; }
        sr_csync();
        wv = sm_red[0]; wi = sm_redi[0];
#pragma unroll
        for (int w = 1; w < SR_WARPS; w++)
            if (sk_better(sm_red[w], sm_redi[w], wv, wi)) { wv = sm_red[w]; wi = sm_redi[w]; }
        __threadfence();
        const int tok = wi;
        pos++;
        // next input row: exact bf16 -> f32 upcast of the embedding (reference qwen_asr.c:412-419,816)
        for (int e = tid; e < H; e += SR_THREADS) sm_x[e] = __uint_as_float(((uint32_t)p.emb[(size_t)tok * H + e]) << 16);
        if (b == 0 && tid == 0) {
            p.d_tokens[step] = tok;
            if (p.h_tokens) p.h_tokens[step] = tok;
        }
        stop = tok == 151643 || tok == 151645; // reference qwen_asr.c:792
        sr_csync();
        mark();
    }
    if (step < p.n_steps) { // early stop: tell the producer how far the ring was drained (it waits for what it has in flight)
        sr_csync();
        if (tid == 0) { sm_ctl[1] = consumed; __threadfence_block(); sm_ctl[0] = 1; }
    }
    if (prof && prof_n < p.prof_cap) prof[prof_n] = 0; // terminator: stamps of an earlier, longer launch may follow
    if (b == 0) {
        for (int e = tid; e < H; e += SR_THREADS) p.x_io[e] = sm_x[e];
        if (tid == 0) { p.d_pos[0] = pos; *p.d_step = step; }
    }
}

// ---- host side ---------------------------------------------------------------------------------
static int g_sr_slots_dev[32] = {}; // ring slots per device (0 = not initialised): the shared-memory opt-in belongs to the (function, device) pair

int launch_decode_rounds(cudaStream_t s, const StreamParams &p, int grid, char *err, size_t errlen) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (!g_sr_slots_dev[dev & 31]) {
        int optin = 0, per_sm = 0;
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaError_t e = cudaFuncSetAttribute(decode_rounds_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(decode_rounds_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_rounds_kernel<4>, SR_ALL_THREADS, (size_t)optin);
        if (e != cudaSuccess || per_sm < 1 || optin < (int)SrLayout::total(3, SK_MAX_K, SK_MAX_H)) {
            snprintf(err, errlen, "decode rounds kernel unavailable: %s (blocks/SM=%d, shared memory opt-in %d)", cudaGetErrorString(e), per_sm, optin);
            cudaGetLastError();
            return -1;
        }
        g_sr_slots_dev[dev & 31] = optin;
    }
    // Measured defaults (profiles/r02_decode_latency.txt): 0.6B dims - 3 slots, 3 rounds of L2 prefetch, 4 KB copies; 1.7B dims - all the
    // slots that fit (5), 6 rounds, 8 KB copies.
    const bool small = p.H <= 1024;
    const int kmax = p.I > 2048 ? p.I : 2048;
    int nslot = (int)((g_sr_slots_dev[dev & 31] - SrLayout::rest(kmax, p.H)) / SR_ROUND);
    if (nslot > SR_MAX_SLOTS) nslot = SR_MAX_SLOTS;
    static int e_slots = -2, e_chunk = -2, e_ahead = -2, e_prod = -2, e_pad = -2, e_chunk_head = -2;
    if (e_slots == -2) {
        const char *e;
        e = getenv("QASR_SR_SLOTS"); e_slots = e ? atoi(e) : -1;
        e = getenv("QASR_SR_CHUNK"); e_chunk = e ? atoi(e) : -1;
        e = getenv("QASR_SR_CHUNK_HEAD"); e_chunk_head = e ? atoi(e) : -1;
        e = getenv("QASR_SR_L2AHEAD"); e_ahead = e ? atoi(e) : -1;
        e = getenv("QASR_SR_PRODUCERS"); e_prod = e ? atoi(e) : -1;
        e = getenv("QASR_SR_PAD_KB"); e_pad = e ? atoi(e) : 0;
    }
    const int want_slots = e_slots >= 2 ? e_slots : (small ? 3 : SR_MAX_SLOTS);
    if (want_slots < nslot) nslot = want_slots;
    size_t smem = SrLayout::total(nslot, kmax, p.H);
    if (e_pad > 0 && smem + (size_t)e_pad * 1024 <= (size_t)g_sr_slots_dev[dev & 31]) smem += (size_t)e_pad * 1024; // experiment: unused shared memory (shrinks L1)
    StreamParams q = p;
    q.sr_chunk = (e_chunk >= 1024 && SR_ROUND % e_chunk == 0) ? e_chunk : (small ? 4096 : 8192);
    q.sr_chunk_head = (e_chunk_head >= 1024 && SR_ROUND % e_chunk_head == 0) ? e_chunk_head : 8192;
    q.l2_ahead_units = e_ahead >= 0 ? e_ahead : (small ? 3 : 6);
    const int n_prod = (e_prod >= 1 && e_prod <= SR_MAX_PRODUCERS) ? e_prod : SR_DEFAULT_PRODUCERS;
    void *args[] = {(void *)&q, (void *)&nslot};
    const void *kern = p.H <= 1024 ? (const void *)decode_rounds_kernel<2> : (const void *)decode_rounds_kernel<4>;
    cudaError_t e = cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(SR_THREADS + 32 * n_prod), args, smem, s);
    if (e != cudaSuccess) {
        snprintf(err, errlen, "decode rounds kernel launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

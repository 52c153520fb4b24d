// qasr_batch.cu - the throughput path behind qasr_cuda_transcribe_batch: many independent units (the segments of -S
// mode, reference qwen_asr.c:941-1103 with --past-text no, or separate utterances) go through the front end, the
// encoder, the prefill and the greedy decode TOGETHER, so that every weight matrix is read once per group instead of
// once per unit (the reference's loop is serial: one transcribe_segment per segment, qwen_asr.c:987).
//
//   front end   per unit (its own dynamic max, qwen_asr_audio.c:361-383) into one mel [128][frames of all units]
//   encoder     rows of all units of a sub-group concatenated (encode_units_device): conv / transformer GEMMs with
//               M = sum of the units' positions / tokens; chunk padding, positional rows and attention windows per unit
//   prefill     prompt rows of all units concatenated: M = sum (n_pre + T_u + n_suf - 1); q/k-norm + RoPE + KV store per
//               row with that row's (unit, position); causal attention per unit over its own KV cache
//   decode      one step = one token for EVERY live unit: the four projections of a layer are skinny tcgen05 GEMMs with the
//               B current rows as the N operand (weights streamed once per step for all B sequences), attention per
//               (unit, kv head) over that unit's cache, lm_head GEMM + per-row argmax (ties -> lowest index) + on-device
//               embedding gather.  A step is one CUDA graph (positions / step counter live in device memory) replayed
//               from the host, ids are read back every few steps to stop at EOS / the caps.
//
// KV pool: f32 [unit][layer][kv head][cap][128] for K and for V: the reference's f32 values (qwen_asr.h:205-206), head-major
// so that the decode attention of one (unit, kv head) streams one contiguous range.
// Groups of at most qasr_cuda_max_batch() units fall through to the persistent single-launch decode kernel instead
// (qasr_stream.cu, 2 / 4 sequences per weight pass): its per-step latency is lower than a chain of GEMM launches.
#include "qasr_ctx.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

struct BatchState {
    DevBuf samples, mel, meltmp, enc, x, ws, kv_k, kv_v, logits, meta, xb;
    int *d_pos = nullptr, *d_step = nullptr, *d_tokens = nullptr; // [QASR_BATCH_CAP], [1], [steps][B]
    int *h_tokens = nullptr, *dh_tokens = nullptr;                  // mapped pinned mirror of d_tokens
    int kv_cap = 0, kv_units = 0;
    static constexpr int STEP_CHUNK = 8;  // decode steps between host checks
    static constexpr int MAX_UNITS = 256;
};

void batch_release(qasr_ctx_t *c) {
    BatchState *b = c->batch;
    if (!b) return;
    b->samples.release(); b->mel.release(); b->meltmp.release(); b->enc.release(); b->x.release(); b->ws.release();
    b->kv_k.release(); b->kv_v.release(); b->logits.release(); b->meta.release(); b->xb.release();
    cudaFree(b->d_pos); cudaFree(b->d_step); cudaFree(b->d_tokens);
    if (b->h_tokens) cudaFreeHost(b->h_tokens);
    delete b;
    c->batch = nullptr;
}

static int batch_state(qasr_ctx_t *c, BatchState **out) {
    if (!c->batch) {
        BatchState *b = new BatchState();
        const size_t tok_bytes = (size_t)BatchState::STEP_CHUNK * BatchState::MAX_UNITS * 4;
        if (cudaMalloc(&b->d_pos, BatchState::MAX_UNITS * 4) != cudaSuccess || cudaMalloc(&b->d_step, 4) != cudaSuccess ||
            cudaMalloc(&b->d_tokens, tok_bytes) != cudaSuccess || cudaHostAlloc((void **)&b->h_tokens, tok_bytes, cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer((void **)&b->dh_tokens, b->h_tokens, 0) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(b->d_pos); cudaFree(b->d_step); cudaFree(b->d_tokens);
            if (b->h_tokens) cudaFreeHost(b->h_tokens);
            delete b;
            return set_err(QASR_ERR_NOMEM, "batch state allocation failed");
        }
        c->batch = b;
    }
    *out = c->batch;
    return 0;
}

// Largest group the batched path takes at once.  The per-step weight traffic is shared by the whole group while the KV
// read grows with it (229 376 B per cached position per sequence, f32 like the reference), so past a few dozen sequences
// the step time is dominated by attention: measured on one B200 (1.7B, 30 s units) 12.2k / 17.5k / 24.2k tokens/s at 32 / 64 / 128
// sequences per step (0.6B, 20 s units: 25.5k at 60, 35.8k at 120).
static int batch_group_max() {
    static int v = 0;
    if (!v) { const char *e = getenv("QASR_BATCH_MAX"); v = e && atoi(e) >= 1 ? atoi(e) : 128; if (v > BatchState::MAX_UNITS) v = BatchState::MAX_UNITS; }
    return v;
}
// equal-sized groups: 140 units with a group limit of 128 run as 70 + 70, not 128 + 12
static int batch_group_size(int remaining) {
    const int gmax = batch_group_max(), groups = (remaining + gmax - 1) / gmax;
    return (remaining + groups - 1) / groups;
}
int batch_plan(int count, int *out_groups, int *out_group_size) {
    int groups = 0, first = 0;
    for (int i = 0; i < count;) { const int B = batch_group_size(count - i); if (!groups) first = B; groups++; i += B; }
    if (out_groups) *out_groups = groups;
    if (out_group_size) *out_group_size = first;
    return 0;
}

struct GroupPlan {
    int B = 0;
    std::vector<int> frames, T, row0, P, enc0, kv0;
    int R = 0, T_total = 0, F_total = 0, max_P = 0, max_total = 0;
};

// ---- one decode step for B sequences (enqueued on the stream; captured into a graph by the caller)
static int enqueue_batch_step(qasr_ctx_t *c, BatchState *b, int B, float *xb, uint8_t *W, size_t unit_stride, size_t layer_stride) {
    const size_t head_stride = layer_stride / 8; // [kv head][cap][128] inside a layer block
    const int H = c->H, I = c->I;
    cudaStream_t s = c->stream;
    const bool two = c->nsplit == 2;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    const size_t o_xn = carve((size_t)B * H * 4), o_qkv = carve((size_t)B * 4096 * 4), o_att = carve((size_t)B * 2048 * 4), o_act = carve((size_t)B * I * 4),
                 o_ssq = carve((size_t)B * (H >> 7) * 4);
    bf16_t *xn_hi = reinterpret_cast<bf16_t *>(W + o_xn), *xn_lo = xn_hi + (size_t)B * H;
    // RMSNorms fused across the GEMMs around them (GemmEpilogue::nx_* / in_ssq, qasr_internal.h): only the first norm of a step is a launch
    float *ssq = reinterpret_cast<float *>(W + o_ssq);
    const bool fuse = gemm_tc_can_fuse_norm(B, 2048, H) && gemm_tc_can_fuse_norm(B, I, H);
    GemmEpilogue from_x, to_x;
    from_x.in_ssq = ssq; from_x.in_tiles = H >> 7; from_x.in_eps = 1e-6f;
    to_x.nx_hi = xn_hi; to_x.nx_lo = xn_lo; to_x.nx_ssq = ssq;
    float *qkv = reinterpret_cast<float *>(W + o_qkv);
    bf16_t *at_hi = reinterpret_cast<bf16_t *>(W + o_att), *at_lo = at_hi + (size_t)B * 2048;
    bf16_t *ac_hi = reinterpret_cast<bf16_t *>(W + o_act), *ac_lo = ac_hi + (size_t)B * I;
    const float scale = 1.0f / sqrtf((float)c->hd);
    for (int l = 0; l < c->dec_layers; l++) { // reference qwen_asr_decoder.c:632-678, B rows at a time
        const DecLayerW &L = c->dec[l];
        float *kl = b->kv_k.as<float>() + (size_t)l * layer_stride, *vl = b->kv_v.as<float>() + (size_t)l * layer_stride;
        const bool in_fused = fuse && l > 0;
        if (!in_fused) launch_rmsnorm(s, xb, L.in_norm, 1e-6f, B, H, nullptr, xn_hi, two ? xn_lo : nullptr);
        CKR(gemm(c, xn_hi, xn_lo, B, H, L.wqkv, 4096, QASR_GEMM_F32, qkv, nullptr, nullptr, nullptr, 4096, in_fused ? &from_x : nullptr));
        launch_attn_decode_batch(s, qkv, L.qn, L.kn, c->rope_cos, c->rope_sin, kl, vl, unit_stride, head_stride, b->d_pos, B, 1e-6f, scale, at_hi, two ? at_lo : nullptr);
        to_x.nx_gamma = L.post_norm;
        CKR(gemm(c, at_hi, at_lo, B, 2048, L.wo, H, QASR_GEMM_RESIDUAL, xb, nullptr, nullptr, nullptr, H, fuse ? &to_x : nullptr));
        if (!fuse) launch_rmsnorm(s, xb, L.post_norm, 1e-6f, B, H, nullptr, xn_hi, two ? xn_lo : nullptr);
        CKR(gemm(c, xn_hi, xn_lo, B, H, L.wgu, 2 * I, QASR_GEMM_SWIGLU_SPLIT, nullptr, ac_hi, ac_lo, nullptr, I, fuse ? &from_x : nullptr));
        to_x.nx_gamma = l + 1 < c->dec_layers ? c->dec[l + 1].in_norm : c->final_norm; // the norm that reads x next
        CKR(gemm(c, ac_hi, ac_lo, B, I, L.wdown, H, QASR_GEMM_RESIDUAL, xb, nullptr, nullptr, nullptr, H, fuse ? &to_x : nullptr));
        c->launches += 1 + (in_fused ? 0 : 1) + (fuse ? 0 : 1);
    }
    // head: final RMSNorm -> tied lm_head (reference qwen_asr_decoder.c:683-684) -> argmax -> next input row
    if (!fuse) launch_rmsnorm(s, xb, c->final_norm, 1e-6f, B, H, nullptr, xn_hi, two ? xn_lo : nullptr);
    CKR(gemm(c, xn_hi, xn_lo, B, H, c->emb, c->V, QASR_GEMM_F32, b->logits.as<float>(), nullptr, nullptr, nullptr, c->V, fuse ? &from_x : nullptr));
    launch_argmax_next(s, b->logits.as<float>(), c->V, c->emb, H, xb, b->d_pos, b->d_step, b->d_tokens, b->dh_tokens, B, BatchState::STEP_CHUNK);
    c->launches += fuse ? 2 : 3;
    return 0;
}

static size_t step_ws_bytes(const qasr_ctx_t *c, int B) {
    const size_t H = c->H, I = c->I, b = B;
    return align_up(b * H * 4, 256) + align_up(b * 4096 * 4, 256) + align_up(b * 2048 * 4, 256) + align_up(b * I * 4, 256) + align_up(b * (H >> 7) * 4, 256);
}

static int run_group(qasr_ctx_t *c, BatchState *b, const float *const *samples, const int *n_samples, int B, const int *max_new, int ids_stride,
                     int *out_ids, int *out_n, double *tm) {
    const int H = c->H, I = c->I, L = c->dec_layers;
    const int n_pre = (int)c->pre_ids.size(), n_suf = (int)c->suf_ids.size();
    cudaStream_t s = c->stream;
    const bool two = c->nsplit == 2;
    GroupPlan g;
    g.B = B;
    g.frames.resize(B); g.T.resize(B); g.row0.resize(B); g.P.resize(B); g.enc0.resize(B); g.kv0.resize(B);
    std::vector<size_t> soff(B + 1, 0);
    int cap_new = 0;
    for (int u = 0; u < B; u++) {
        g.frames[u] = n_samples[u] / 160;
        if (g.frames[u] <= 0) return set_err(QASR_ERR_ARG, "unit too short (%d samples)", n_samples[u]); // reference: mel returns NULL (qwen_asr_audio.c:313-317)
        g.T[u] = qasr_cuda_encoder_tokens(g.frames[u]);
        g.enc0[u] = g.T_total;
        g.T_total += g.T[u];
        g.row0[u] = g.R;
        g.P[u] = n_pre + g.T[u] + n_suf - 1;
        g.kv0[u] = g.P[u];
        g.R += g.P[u];
        g.F_total += g.frames[u];
        g.max_P = std::max(g.max_P, g.P[u]);
        g.max_total = std::max(g.max_total, g.P[u] + 1);
        soff[u + 1] = soff[u] + align_up((size_t)n_samples[u] * 4, 256);
        cap_new = std::max(cap_new, max_new[u]);
    }
    cudaEvent_t *ev = c->ev;
    CK(cudaEventRecord(ev[0], s));
    // ---- front end: one mel [128][F_total], every unit normalised by its own maximum
    if (b->samples.reserve(soff[B]) || b->mel.reserve((size_t)128 * g.F_total * 4)) return set_err(QASR_ERR_NOMEM, "batch front-end buffers");
    {
        int fmax = 0;
        for (int u = 0; u < B; u++) fmax = std::max(fmax, g.frames[u]);
        if (b->meltmp.reserve((size_t)B * fmax * 128 * 4)) return set_err(QASR_ERR_NOMEM, "batch front-end buffers");
        int f0 = 0;
        for (int u = 0; u < B; u++) {
            float *d_s = reinterpret_cast<float *>(b->samples.as<uint8_t>() + soff[u]);
            CK(cudaMemcpyAsync(d_s, samples[u], (size_t)n_samples[u] * 4, cudaMemcpyHostToDevice, s));
            launch_mel(s, d_s, n_samples[u], g.frames[u], c->mel_cos, c->mel_sin, c->mel_win, c->mel_fb, b->meltmp.as<float>() + (size_t)u * fmax * 128,
                       c->d_gmax + u, b->mel.as<float>(), g.F_total, f0);
            f0 += g.frames[u];
        }
        c->launches += 3 * B;
    }
    CK(cudaEventRecord(ev[2], s));
    // ---- encoder, in sub-groups of ~16 x 30 s of audio per pass (bounds the activation workspace: the stage-1 conv output is 184 MB per 30 s)
    if (b->enc.reserve((size_t)g.T_total * H * 4)) return set_err(QASR_ERR_NOMEM, "batch encoder output");
    {
        static int enc_frames_max = 0;
        if (!enc_frames_max) { const char *e = getenv("QASR_BATCH_ENC_FRAMES"); enc_frames_max = e && atoi(e) >= 100 ? atoi(e) : 48000; }
        // the units of a pass must be contiguous along the mel frame axis: conv1 reads mel[ih * F_total + frame]
        int u0 = 0, f0 = 0;
        while (u0 < B) {
            int u1 = u0, fsum = 0;
            while (u1 < B && (u1 == u0 || fsum + g.frames[u1] <= enc_frames_max)) fsum += g.frames[u1++];
            int Tsub = 0;
            // sub-group view of the mel: same row stride F_total, first frame f0 -> shift the base pointer, keep chunk offsets relative
            CKR(encode_units_device(c, b->mel.as<float>() + f0, g.F_total, &g.frames[u0], u1 - u0, b->enc.as<float>() + (size_t)g.enc0[u0] * H, &Tsub));
            f0 += fsum;
            u0 = u1;
        }
    }
    CK(cudaEventRecord(ev[3], s));
    // ---- per-group device tables: [row0 | P | enc0 | T | kv0] per unit, [unit | pos] per prefill row, prompt ids
    const size_t n_meta = (size_t)5 * B + (size_t)2 * g.R + n_pre + n_suf;
    if (b->meta.reserve(n_meta * 4)) return set_err(QASR_ERR_NOMEM, "batch tables");
    std::vector<int> meta(n_meta);
    int *m_row0 = meta.data(), *m_P = m_row0 + B, *m_enc0 = m_P + B, *m_T = m_enc0 + B, *m_kv0 = m_T + B, *m_runit = m_kv0 + B, *m_rpos = m_runit + g.R,
        *m_pre = m_rpos + g.R, *m_suf = m_pre + n_pre;
    for (int u = 0; u < B; u++) {
        m_row0[u] = g.row0[u]; m_P[u] = g.P[u]; m_enc0[u] = g.enc0[u]; m_T[u] = g.T[u]; m_kv0[u] = g.kv0[u];
        for (int i = 0; i < g.P[u]; i++) { m_runit[g.row0[u] + i] = u; m_rpos[g.row0[u] + i] = i; }
    }
    std::copy(c->pre_ids.begin(), c->pre_ids.end(), m_pre);
    std::copy(c->suf_ids.begin(), c->suf_ids.end(), m_suf);
    int *d_meta = b->meta.as<int>();
    CK(cudaMemcpyAsync(d_meta, meta.data(), n_meta * 4, cudaMemcpyHostToDevice, s));
    const int *d_row0 = d_meta, *d_P = d_meta + B, *d_enc0 = d_P + B, *d_T = d_enc0 + B, *d_kv0 = d_T + B, *d_runit = d_kv0 + B, *d_rpos = d_runit + g.R,
              *d_pre = d_rpos + g.R, *d_suf = d_pre + n_pre;
    // ---- KV pool and workspaces
    int need_cap = 0;
    for (int u = 0; u < B; u++) need_cap = std::max(need_cap, g.kv0[u]);
    need_cap += cap_new + BatchState::STEP_CHUNK + 2; // finished sequences keep stepping until the whole group is done
    if (need_cap > b->kv_cap || B > b->kv_units) {
        const int cap = std::max(b->kv_cap, (need_cap + 63) / 64 * 64), units = std::max(b->kv_units, B);
        const size_t bytes = (size_t)units * L * cap * 1024 * 4;
        CK(cudaStreamSynchronize(s));
        if (b->kv_k.reserve(bytes) || b->kv_v.reserve(bytes)) return set_err(QASR_ERR_NOMEM, "batched KV pool (%zu bytes x 2)", bytes);
        b->kv_cap = cap; b->kv_units = units;
        c->ws_gen++;
    }
    const size_t layer_stride = (size_t)b->kv_cap * 1024, unit_stride = (size_t)L * layer_stride;
    CKR(ensure_rope(c, need_cap));
    const size_t R = g.R;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    const size_t o_xn = carve(R * H * 4), o_qkv = carve(R * 4096 * 4), o_q = carve(R * 2048 * 4), o_att = carve(R * 2048 * 4), o_act = carve(R * I * 4);
    const size_t ws_need = std::max(off, step_ws_bytes(c, B));
    if (b->x.reserve(R * H * 4) || b->ws.reserve(ws_need) || b->xb.reserve((size_t)B * H * 4) || b->logits.reserve((size_t)B * c->V * 4))
        return set_err(QASR_ERR_NOMEM, "batched prefill workspace (%zu bytes)", ws_need);
    if (b->ws.grew || b->xb.grew || b->logits.grew) { c->ws_gen++; b->ws.grew = b->xb.grew = b->logits.grew = false; }
    uint8_t *W = b->ws.as<uint8_t>();
    float *x = b->x.as<float>(), *xb = b->xb.as<float>();
    // ---- prompt rows (on-device assembly) and the batched prefill, reference qwen_asr.c:685-769, qwen_asr_decoder.c:457-563
    launch_assemble_prompts(s, c->emb, H, d_pre, n_pre, d_suf, n_suf, b->enc.as<float>(), d_enc0, d_T, d_row0, B, g.max_total, x, xb);
    c->launches += 1;
    {
        PdlScope pdl(false);
        bf16_t *xn_hi = reinterpret_cast<bf16_t *>(W + o_xn), *xn_lo = xn_hi + R * H;
        float *qkv = reinterpret_cast<float *>(W + o_qkv), *q = reinterpret_cast<float *>(W + o_q);
        bf16_t *at_hi = reinterpret_cast<bf16_t *>(W + o_att), *at_lo = at_hi + R * 2048;
        bf16_t *ac_hi = reinterpret_cast<bf16_t *>(W + o_act), *ac_lo = ac_hi + R * I;
        const float scale = 1.0f / sqrtf((float)c->hd);
        for (int l = 0; l < L; l++) {
            const DecLayerW &Lw = c->dec[l];
            float *kl = b->kv_k.as<float>() + (size_t)l * layer_stride, *vl = b->kv_v.as<float>() + (size_t)l * layer_stride;
            launch_rmsnorm(s, x, Lw.in_norm, 1e-6f, g.R, H, nullptr, xn_hi, two ? xn_lo : nullptr);
            CKR(gemm(c, xn_hi, xn_lo, g.R, H, Lw.wqkv, 4096, QASR_GEMM_F32, qkv, nullptr, nullptr, nullptr, 4096));
            launch_qk_norm_rope_store_rows(s, qkv, Lw.qn, Lw.kn, c->rope_cos, c->rope_sin, d_runit, d_rpos, g.R, 1e-6f, q, kl, vl, unit_stride, layer_stride / 8);
            launch_attn_prefill_batch(s, q, kl, vl, unit_stride, layer_stride / 8, d_row0, d_P, B, g.max_P, c->heads, c->kv_heads, scale, at_hi, two ? at_lo : nullptr);
            CKR(gemm(c, at_hi, at_lo, g.R, 2048, Lw.wo, H, QASR_GEMM_RESIDUAL, x, nullptr, nullptr, nullptr, H));
            launch_rmsnorm(s, x, Lw.post_norm, 1e-6f, g.R, H, nullptr, xn_hi, two ? xn_lo : nullptr);
            CKR(gemm(c, xn_hi, xn_lo, g.R, H, Lw.wgu, 2 * I, QASR_GEMM_SWIGLU_SPLIT, nullptr, ac_hi, ac_lo, nullptr, I));
            CKR(gemm(c, ac_hi, ac_lo, g.R, I, Lw.wdown, H, QASR_GEMM_RESIDUAL, x, nullptr, nullptr, nullptr, H));
            c->launches += 4;
        }
    }
    CK(cudaMemcpyAsync(b->d_pos, d_kv0, (size_t)B * 4, cudaMemcpyDeviceToDevice, s));
    CK(cudaEventRecord(ev[4], s));
    CK(cudaGetLastError());
    // ---- greedy decode: STEP_CHUNK graph replays, then the host reads the ids back (stop at EOS / caps)
    std::vector<int> n(B, 0);
    std::vector<char> done(B, 0);
    long long ksteps = 0;
    for (;;) {
        int remaining = 0;
        for (int u = 0; u < B; u++) if (!done[u]) remaining = std::max(remaining, max_new[u] - n[u]);
        if (remaining <= 0) break;
        const int chunk = std::min(remaining, (int)BatchState::STEP_CHUNK);
        CK(cudaMemsetAsync(b->d_step, 0, 4, s));
        for (int i = 0; i < chunk; i++) {
            PdlScope pdl(true);
            CKR(run_cached_graph(c, 3, B, c->nsplit, [&]() -> int { return enqueue_batch_step(c, b, B, xb, W, unit_stride, layer_stride); }));
        }
        CK(cudaStreamSynchronize(s));
        ksteps += chunk;
        for (int i = 0; i < chunk; i++)
            for (int u = 0; u < B; u++) {
                if (done[u]) continue;
                const int tok = b->h_tokens[i * B + u];
                out_ids[(size_t)u * ids_stride + n[u]++] = tok;
                if (tok == QASR_TOKEN_ENDOFTEXT || tok == QASR_TOKEN_IM_END || n[u] >= max_new[u]) done[u] = 1; // reference qwen_asr.c:792
            }
    }
    CK(cudaEventRecord(ev[1], s));
    CK(cudaStreamSynchronize(s));
    for (int u = 0; u < B; u++) out_n[u] = n[u];
    float t_mel = 0.f, t_enc = 0.f, t_pre = 0.f, t_dec = 0.f;
    cudaEventElapsedTime(&t_mel, ev[0], ev[2]);
    cudaEventElapsedTime(&t_enc, ev[2], ev[3]);
    cudaEventElapsedTime(&t_pre, ev[3], ev[4]);
    cudaEventElapsedTime(&t_dec, ev[4], ev[1]);
    c->last_decode_ms = t_dec;
    c->decode_ms_total += t_dec;
    c->decode_steps_total += ksteps;
    if (tm) { tm[0] += t_mel; tm[1] += t_enc; tm[2] += t_pre; tm[3] += t_dec; }
    return 0;
}

int batch_transcribe(qasr_ctx_t *c, const float *const *samples, const int *n_samples, int count, const int *max_new, int ids_stride,
                     int *out_ids, int *out_n, double *timings_ms) {
    BatchState *b = nullptr;
    CKR(batch_state(c, &b));
    for (int i = 0; i < count;) {
        const int B = batch_group_size(count - i);
        CKR(run_group(c, b, samples + i, n_samples + i, B, max_new + i, ids_stride, out_ids + (size_t)i * ids_stride, out_n + i, timings_ms));
        i += B;
    }
    return 0;
}

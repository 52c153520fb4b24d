#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libqasr_ref_*.so).

The reference's own tests hold no tensor-level vectors for this path, so the pins
are outputs of the reference itself run in the build container on the deterministic synthetic
checkpoint (tools/synth_weights.c, seed 1234) and deterministic synthetic audio
(smol-vision_b200/synth.py).  Run:  python tools/make_golden.py
The reference does not exist on the GPU box; the fixtures written here are what travels.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle.bindings import RefLib  # noqa: E402

pkg = ge.load_package()
OUT = os.path.join(ROOT, "tests", "golden")
PRE = [151644, 8948, 198, 151645, 198, 151644, 872, 198, 151669]
SUF = [151670, 151645, 198, 151644, 77091, 198]
f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def rnd(shape, seed, scale=1.0):
    n = int(np.prod(shape))
    u = pkg.synth._uniform(n, seed, 17) + pkg.synth._uniform(n, seed, 18) + pkg.synth._uniform(n, seed, 19)
    return ((u - 1.5) * 2.0 * scale).astype(np.float32).reshape(shape)


def to_bf16(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return u.astype(np.uint16)


def segment_golden(ref, variant, seconds, seed, n_ids, n_logit_steps):
    audio = pkg.synth_audio(seconds, seed)
    mel = ref.mel(audio)
    enc = ref.encode(mel)
    T = enc.shape[0]
    ids, _ = ref.transcribe_ids(audio, n_ids)
    # teacher-forced logits: rebuild the prompt, prefill, then step with full logits
    H = ref.cfg["dec_hidden"]
    emb = np.zeros((len(PRE) + T + len(SUF), H), np.float32)
    for i, t in enumerate(PRE):
        emb[i] = ref.embed(t)
    emb[len(PRE):len(PRE) + T] = enc
    for i, t in enumerate(SUF):
        emb[len(PRE) + T + i] = ref.embed(t)
    ref.kv_len = 0
    ref.prefill(emb[:-1])
    P = emb.shape[0] - 1
    k0, v0 = ref.read_kv(0, P)
    k27, v27 = ref.read_kv(27, P)
    rows = [0, 1, P // 2, P - 1]
    top_idx, top_val, head = [], [], []
    x = emb[-1]
    for s in range(n_logit_steps):
        lg = ref.step_logits(x)
        order = np.argsort(-lg, kind="stable")[:16]
        top_idx.append(order.astype(np.int32))
        top_val.append(lg[order])
        head.append(lg[:2048].copy())
        assert int(order[0]) == int(ids[s]), "teacher-forced argmax must equal the greedy id"
        x = ref.embed(int(ids[s]))
    np.savez_compressed(
        os.path.join(OUT, f"segment_{variant}.npz"),
        seconds=np.float64(seconds), seed=np.int64(seed), audio_i16=np.round(audio * 32768).astype(np.int16),
        mel=mel, enc=enc, ids=ids.astype(np.int32), prefill_len=np.int32(P), kv_rows=np.array(rows, np.int32),
        k0=k0[rows], v0=v0[rows], k27=k27[rows], v27=v27[rows],
        logits_top_idx=np.stack(top_idx), logits_top_val=np.stack(top_val), logits_head=np.stack(head))
    print(f"segment_{variant}: frames={mel.shape[1]} T={T} ids={ids.tolist()}")


def ops_golden(ref):
    L = ref.lib
    g = {}
    # qwen_gelu / qwen_silu / qwen_softmax
    x = rnd((4, 640), 1, 3.0)
    y = x.copy(); L.qwen_gelu.argtypes = [f32p, C.c_int]; L.qwen_gelu(y, y.size); g["gelu_x"], g["gelu_y"] = x, y
    y = x.copy(); L.qwen_silu.argtypes = [f32p, C.c_int]; L.qwen_silu(y, y.size); g["silu_y"] = y
    y = x.copy(); L.qwen_softmax.argtypes = [f32p, C.c_int, C.c_int]; L.qwen_softmax(y, 4, 640); g["softmax_y"] = y
    # norms
    w, b = rnd((896,), 2, 1.0) + 1.0, rnd((896,), 3, 0.1)
    xn = rnd((5, 896), 4, 2.0)
    y = np.empty_like(xn)
    L.qwen_layer_norm.argtypes = [f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_float]
    L.qwen_layer_norm(y, xn, w, b, 5, 896, 1e-5)
    g["ln_x"], g["ln_w"], g["ln_b"], g["ln_y"] = xn, w, b, y
    y = np.empty_like(xn)
    L.qwen_rms_norm.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, C.c_float]
    L.qwen_rms_norm(y, xn, w, 5, 896, 1e-6)
    g["rms_y"] = y
    xh = rnd((3, 16 * 128), 5, 2.0)
    wh = rnd((128,), 6, 0.2) + 1.0
    y = xh.copy()
    L.qwen_rms_norm_per_head.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_float]
    L.qwen_rms_norm_per_head(y, wh, 3, 16, 128, 1e-6)
    g["rmsh_x"], g["rmsh_w"], g["rmsh_y"] = xh, wh, y
    # swiglu
    gu = rnd((3, 2 * 192), 7, 2.0)
    y = np.empty((3, 192), np.float32)
    L.qwen_swiglu_multiply.argtypes = [f32p, f32p, C.c_int, C.c_int]
    L.qwen_swiglu_multiply(y, gu, 3, 192)
    g["swiglu_x"], g["swiglu_y"] = gu, y
    # rope
    pos = np.array([0, 1, 7, 300, 2047], np.int32)
    c = np.empty((5, 128), np.float32); s = np.empty((5, 128), np.float32)
    L.qwen_compute_rope_neox.argtypes = [f32p, f32p, i32p, C.c_int, C.c_int, C.c_float]
    L.qwen_compute_rope_neox(c, s, pos, 5, 128, 1e6)
    xr = rnd((5, 8 * 128), 8, 1.0)
    y = xr.copy()
    L.qwen_apply_rope_neox.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, C.c_int]
    L.qwen_apply_rope_neox(y, c, s, 5, 8, 128)
    g["rope_pos"], g["rope_cos"], g["rope_sin"], g["rope_x"], g["rope_y"] = pos, c, s, xr, y
    pe = np.empty((13, 896), np.float32)
    L.qwen_sinusoidal_pe.argtypes = [f32p, C.c_int, C.c_int]
    L.qwen_sinusoidal_pe(pe, 13, 896)
    g["pe"] = pe
    # linear f32 / bf16 (seq 1 = matvec path, seq 5 = sgemm path) / qkv / argmax
    xl = rnd((5, 256), 9, 1.0)
    W = rnd((96, 256), 10, 0.2)
    bl = rnd((96,), 11, 0.1)
    y = np.empty((5, 96), np.float32)
    L.qwen_linear.argtypes = [f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int]
    L.qwen_linear(y, xl, W, bl, 5, 256, 96)
    g["lin_x"], g["lin_w"], g["lin_b"], g["lin_y"] = xl, W, bl, y
    Wb = to_bf16(rnd((200, 256), 12, 0.2))
    L.qwen_linear_nobias_bf16.argtypes = [f32p, f32p, u16p, C.c_int, C.c_int, C.c_int]
    y5 = np.empty((5, 200), np.float32); L.qwen_linear_nobias_bf16(y5, xl, Wb, 5, 256, 200)
    y1 = np.empty((1, 200), np.float32); L.qwen_linear_nobias_bf16(y1, np.ascontiguousarray(xl[:1]), Wb, 1, 256, 200)
    g["linbf_w"], g["linbf_y5"], g["linbf_y1"] = Wb, y5, y1
    L.qwen_argmax_matvec_bf16.restype = C.c_int
    L.qwen_argmax_matvec_bf16.argtypes = [f32p, u16p, C.c_int, C.c_int]
    g["argmax_idx"] = np.int32(L.qwen_argmax_matvec_bf16(np.ascontiguousarray(xl[0]), Wb, 256, 200))
    # conv2d (stem shape class: 3x3 stride 2 pad 1)
    xc = rnd((3, 16, 11), 13, 1.0)
    wc = rnd((4, 3, 3, 3), 14, 0.3)
    bc = rnd((4,), 15, 0.1)
    yc = np.empty((4, 8, 6), np.float32)
    L.qwen_conv2d.argtypes = [f32p, f32p, f32p, f32p] + [C.c_int] * 8
    L.qwen_conv2d(yc, xc, wc, bc, 3, 4, 16, 11, 3, 3, 2, 1)
    g["conv_x"], g["conv_w"], g["conv_b"], g["conv_y"] = xc, wc, bc, yc
    # attention
    Q = rnd((30, 2 * 64), 16, 1.0); K = rnd((30, 2 * 64), 17, 1.0); V = rnd((30, 2 * 64), 18, 1.0)
    ws = np.array([0, 13, 26, 30], np.int32)
    y = np.zeros_like(Q)
    L.qwen_bidirectional_attention.argtypes = [f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_float, i32p, C.c_int]
    L.qwen_bidirectional_attention(y, Q, K, V, 30, 2, 64, 0.125, ws, 3)
    g["battn_q"], g["battn_k"], g["battn_v"], g["battn_ws"], g["battn_y"] = Q, K, V, ws, y
    Qc = rnd((5, 4 * 128), 19, 1.0); Kc = rnd((12, 2 * 128), 20, 1.0); Vc = rnd((12, 2 * 128), 21, 1.0)
    y = np.zeros_like(Qc)
    L.qwen_causal_attention.argtypes = [f32p, f32p, f32p, f32p] + [C.c_int] * 5 + [C.c_float, C.c_int]
    L.qwen_causal_attention(y, Qc, Kc, Vc, 5, 12, 4, 2, 128, 1.0 / np.sqrt(128.0), 7)
    g["cattn_q"], g["cattn_k"], g["cattn_v"], g["cattn_y"] = Qc, Kc, Vc, y
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **g)
    print("ops:", sorted(g))


def wav_bytes(pcm, channels, rate):
    """Minimal RIFF/WAVE file around interleaved int16 PCM."""
    import struct
    data = np.ascontiguousarray(pcm, np.int16).tobytes()
    return (b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, channels, rate, rate * channels * 2, channels * 2, 16)
            + b"data" + struct.pack("<I", len(data)) + data)


def wav_golden(ref):
    """qwen_parse_wav_buffer (reference qwen_asr_audio.c:40-168) on small synthetic WAV files: mono/stereo, 16 kHz
    pass-through, down- and up-sampling."""
    g = {}
    rng = np.random.default_rng(2024)
    for i, (ch, rate, n) in enumerate([(1, 16000, 1500), (2, 44100, 3000), (1, 8000, 1200), (2, 48000, 2400), (1, 22050, 2345)]):
        t = np.arange(n) / rate
        base = 0.3 * np.sin(2 * np.pi * 440.0 * t) + 0.1 * np.sin(2 * np.pi * 3100.0 * t)
        pcm = np.stack([base * (1.0 - 0.2 * c) for c in range(ch)], axis=1) + 0.02 * rng.standard_normal((n, ch))
        pcm = np.clip(np.round(pcm * 32768.0), -32768, 32767).astype(np.int16)
        g[f"pcm{i}"], g[f"meta{i}"] = pcm.reshape(-1), np.array([ch, rate, n], np.int32)
        g[f"out{i}"] = ref.parse_wav(wav_bytes(pcm, ch, rate))
    g["count"] = np.array(5, np.int32)
    np.savez_compressed(os.path.join(OUT, "wav.npz"), **g)
    print("wav:", [len(g[f"out{i}"]) for i in range(5)])


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = RefLib()
    if "--wav-only" in sys.argv:
        wav_golden(ref)
        return
    ops_golden(ref)
    wav_golden(ref)
    ref.load(pkg.ensure_model_dir("0.6b"))
    segment_golden(ref, "0p6b", seconds=2.5, seed=7, n_ids=16, n_logit_steps=3)
    ref.close()


if __name__ == "__main__":
    main()

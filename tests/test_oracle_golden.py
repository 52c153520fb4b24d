"""CPU: the plain-C oracle (oracle/qasr_oracle.c) against the committed golden vectors, which
are outputs of the UNMODIFIED reference (tools/make_golden.py).  This pins the oracle."""
import ctypes as C

import numpy as np
import pytest

from conftest import prompt_embeds, rel_err

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
TOL = 2e-5  # f32 summation-order noise between the -ffast-math/BLAS reference and the plain loops


@pytest.fixture(scope="module")
def L(oracle_lib):
    return oracle_lib().lib


def test_eltwise(L, golden_ops):
    g = golden_ops
    for name, fn in (("gelu_y", L.qo_gelu), ("silu_y", L.qo_silu)):
        y = g["gelu_x"].copy()
        fn.argtypes = [f32p, C.c_int]
        fn(y, y.size)
        assert rel_err(y, g[name]) < TOL
    y = g["gelu_x"].copy()
    L.qo_softmax.argtypes = [f32p, C.c_int, C.c_int]
    L.qo_softmax(y, 4, 640)
    assert rel_err(y, g["softmax_y"]) < TOL


def test_norms(L, golden_ops):
    g = golden_ops
    y = np.empty_like(g["ln_x"])
    L.qo_layer_norm.argtypes = [f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_float]
    L.qo_layer_norm(y, g["ln_x"], g["ln_w"], g["ln_b"], 5, 896, 1e-5)
    assert rel_err(y, g["ln_y"]) < TOL
    L.qo_rms_norm.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, C.c_float]
    L.qo_rms_norm(y, g["ln_x"], g["ln_w"], 5, 896, 1e-6)
    assert rel_err(y, g["rms_y"]) < TOL
    y = g["rmsh_x"].copy()
    L.qo_rms_norm_per_head.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_float]
    L.qo_rms_norm_per_head(y, g["rmsh_w"], 3, 16, 128, 1e-6)
    assert rel_err(y, g["rmsh_y"]) < TOL


def test_swiglu_rope_pe(L, golden_ops):
    g = golden_ops
    y = np.empty((3, 192), np.float32)
    L.qo_swiglu_multiply.argtypes = [f32p, f32p, C.c_int, C.c_int]
    L.qo_swiglu_multiply(y, g["swiglu_x"], 3, 192)
    assert rel_err(y, g["swiglu_y"]) < TOL
    c = np.empty((5, 128), np.float32)
    s = np.empty((5, 128), np.float32)
    L.qo_compute_rope_neox.argtypes = [f32p, f32p, i32p, C.c_int, C.c_int, C.c_float]
    L.qo_compute_rope_neox(c, s, g["rope_pos"], 5, 128, 1e6)
    # f32 angle = pos*inv_freq has ulp ~1e-4 rad at pos 2047; -ffast-math reassociates the reference's powf
    assert np.abs(c - g["rope_cos"]).max() < 2e-4 and np.abs(s - g["rope_sin"]).max() < 2e-4
    y = g["rope_x"].copy()
    L.qo_apply_rope_neox.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, C.c_int]
    L.qo_apply_rope_neox(y, g["rope_cos"], g["rope_sin"], 5, 8, 128)
    assert rel_err(y, g["rope_y"]) < TOL
    pe = np.empty((13, 896), np.float32)
    L.qo_sinusoidal_pe.argtypes = [f32p, C.c_int, C.c_int]
    L.qo_sinusoidal_pe(pe, 13, 896)
    assert np.abs(pe - g["pe"]).max() < 1e-5


def test_linear_family(L, golden_ops):
    g = golden_ops
    y = np.empty((5, 96), np.float32)
    L.qo_linear.argtypes = [f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int]
    L.qo_linear(y, g["lin_x"], g["lin_w"], g["lin_b"], 5, 256, 96)
    assert rel_err(y, g["lin_y"]) < TOL
    L.qo_linear_bf16.argtypes = [f32p, f32p, u16p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    y5 = np.empty((5, 200), np.float32)
    L.qo_linear_bf16(y5, g["lin_x"], g["linbf_w"], None, 5, 256, 200)
    assert rel_err(y5, g["linbf_y5"]) < TOL
    y1 = np.empty((1, 200), np.float32)
    L.qo_linear_bf16(y1, np.ascontiguousarray(g["lin_x"][:1]), g["linbf_w"], None, 1, 256, 200)
    assert rel_err(y1, g["linbf_y1"]) < TOL
    L.qo_argmax_matvec_bf16.restype = C.c_int
    L.qo_argmax_matvec_bf16.argtypes = [f32p, u16p, C.c_int, C.c_int]
    assert L.qo_argmax_matvec_bf16(np.ascontiguousarray(g["lin_x"][0]), g["linbf_w"], 256, 200) == int(g["argmax_idx"])


def test_argmax_ties_lowest_index(L):
    """Ties resolve to the lowest row (reference strict '>' from -1e30, qwen_asr_kernels.c:536-541)."""
    W = np.zeros((64, 16), np.uint16)
    W[[9, 40, 41]] = 0x3F80  # 1.0 in bf16
    x = np.ones(16, np.float32)
    L.qo_argmax_matvec_bf16.restype = C.c_int
    L.qo_argmax_matvec_bf16.argtypes = [f32p, u16p, C.c_int, C.c_int]
    assert L.qo_argmax_matvec_bf16(x, W, 16, 64) == 9


def test_conv_attention(L, golden_ops):
    g = golden_ops
    y = np.empty((4, 8, 6), np.float32)
    L.qo_conv2d.argtypes = [f32p, f32p, f32p, f32p] + [C.c_int] * 8
    L.qo_conv2d(y, g["conv_x"], g["conv_w"], g["conv_b"], 3, 4, 16, 11, 3, 3, 2, 1)
    assert rel_err(y, g["conv_y"]) < TOL
    y = np.zeros_like(g["battn_q"])
    L.qo_bidirectional_attention.argtypes = [f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_float, i32p, C.c_int]
    L.qo_bidirectional_attention(y, g["battn_q"], g["battn_k"], g["battn_v"], 30, 2, 64, 0.125, g["battn_ws"], 3)
    assert rel_err(y, g["battn_y"]) < TOL
    y = np.zeros_like(g["cattn_q"])
    L.qo_causal_attention.argtypes = [f32p, f32p, f32p, f32p] + [C.c_int] * 5 + [C.c_float, C.c_int]
    L.qo_causal_attention(y, g["cattn_q"], g["cattn_k"], g["cattn_v"], 5, 12, 4, 2, 128, 1.0 / np.sqrt(128.0), 7)
    assert rel_err(y, g["cattn_y"]) < TOL


def test_segment_against_reference_golden(oracle06, golden_seg, pkg):
    g = golden_seg
    audio = pkg.synth_audio(float(g["seconds"]), int(g["seed"]))
    assert np.array_equal(np.round(audio * 32768).astype(np.int16), g["audio_i16"]), "synthetic audio drifted"
    mel = oracle06.mel(audio)
    d = np.abs(mel - g["mel"])
    # log of near-zero power amplifies f32 summation-order noise near the clamp floor
    assert d.max() < 5e-3 and d.mean() < 2e-5
    enc = oracle06.encode(g["mel"])
    assert rel_err(enc, g["enc"]) < 1e-4
    emb = prompt_embeds(oracle06, g["enc"])
    oracle06.kv_len = 0
    oracle06.prefill(emb[:-1])
    P = int(g["prefill_len"])
    assert oracle06.kv_len == P
    rows = g["kv_rows"]
    for layer, (gk, gv) in ((0, (g["k0"], g["v0"])), (27, (g["k27"], g["v27"]))):
        k, v = oracle06.read_kv(layer, P)
        assert rel_err(k[rows], gk) < 1e-4 and rel_err(v[rows], gv) < 1e-4
    x = emb[-1]
    for s in range(g["logits_top_idx"].shape[0]):
        lg = oracle06.step_logits(x)
        assert np.abs(lg[:2048] - g["logits_head"][s]).max() < 2e-4
        assert int(np.argmax(lg)) == int(g["logits_top_idx"][s][0])
        assert np.abs(lg[g["logits_top_idx"][s]] - g["logits_top_val"][s]).max() < 2e-4
        x = oracle06.embed(int(g["ids"][s]))
    ids, info = oracle06.transcribe_ids(audio, len(g["ids"]))
    assert info["enc_tokens"] == g["enc"].shape[0]
    assert ids.tolist() == g["ids"].tolist()


def test_wav_parse_and_resample_against_reference_golden(oracle_lib):
    """qo_parse_wav_buffer (restatement of qwen_parse_wav_buffer, reference qwen_asr_audio.c:40-168: RIFF walk, stereo
    average, 1/32768, windowed-sinc resampler in double) against outputs of the compiled reference (tests/golden/wav.npz,
    tools/make_golden.py --wav-only)."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from make_golden import wav_bytes
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wav.npz"))
    o = oracle_lib()
    for i in range(int(g["count"])):
        ch, rate, n = (int(v) for v in g[f"meta{i}"])
        out = o.parse_wav(wav_bytes(g[f"pcm{i}"].reshape(n, ch), ch, rate))
        assert out.shape == g[f"out{i}"].shape == ((n,) if rate == 16000 else (n * 16000 // rate,))
        assert np.abs(out - g[f"out{i}"]).max() <= 1e-7
    assert o.parse_wav(b"RIFFxxxxWAVEjunk" + b"\0" * 40) is None       # no fmt / data chunk: rejected like the reference

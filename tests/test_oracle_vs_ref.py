"""CPU: the oracle port against the compiled reference (oracle/_ref) run live, on edge cases the
golden files do not carry: ragged last chunk, sub-chunk audio, too-short audio, KV rollback.
Skipped when oracle/_ref was not built (it only builds where /root/reference exists)."""
import numpy as np
import pytest

from conftest import prompt_embeds, rel_err


@pytest.fixture(scope="module")
def ref06(ref_lib, model06):
    if ref_lib is None:
        pytest.skip("oracle/_ref not available")
    r = ref_lib().load(model06)
    yield r
    r.close()


@pytest.mark.parametrize("seconds", [0.35, 1.0, 2.07])
def test_mel_and_encoder_edge_lengths(ref06, oracle06, pkg, seconds):
    audio = pkg.synth_audio(seconds, seed=11)
    mr, mo = ref06.mel(audio), oracle06.mel(audio)
    assert mr.shape == mo.shape == (128, len(audio) // 160)
    assert np.abs(mr - mo).max() < 5e-3
    er, eo = ref06.encode(mr), oracle06.encode(mr)
    assert er.shape == eo.shape
    assert rel_err(eo, er) < 1e-4


def test_too_short_audio_returns_null(ref06, oracle06):
    x = np.zeros(100, np.float32)  # < 160 samples -> 0 frames (reference qwen_asr_audio.c:313-317)
    assert ref06.mel(x) is None and oracle06.mel(x) is None


def test_kv_rollback_and_delta_prefill(ref06, oracle06, pkg):
    """Streaming prefix reuse: the caller rewinds kv_len and prefills only the new rows
    (reference qwen_asr.c:1811-1829, qwen_asr_decoder.c:527-532)."""
    audio = pkg.synth_audio(1.0, seed=2)
    enc = ref06.encode(ref06.mel(audio))
    emb = prompt_embeds(ref06, enc)
    outs = []
    for eng in (ref06, oracle06):
        eng.kv_len = 0
        eng.prefill(emb[:-1])
        t_full = eng.step(emb[-1])
        eng.kv_len = 10                      # roll back to a 10-row prefix
        eng.prefill(emb[10:-1])              # delta prefill
        t_delta = eng.step(emb[-1])
        outs.append((t_full, t_delta, eng.kv_len))
    assert outs[0] == outs[1]
    assert outs[0][0] == outs[0][1]


def test_wav_parse_and_resample_live(ref_lib, oracle_lib):
    """Live: restated qwen_parse_wav_buffer vs the compiled reference on random PCM at several rates / channel counts."""
    if ref_lib is None:
        pytest.skip("oracle/_ref not available")
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from make_golden import wav_bytes
    r, o = ref_lib(), oracle_lib()
    rng = np.random.default_rng(9)
    for ch, rate, n in [(1, 16000, 4000), (2, 44100, 9000), (1, 8000, 3000), (3, 32000, 5000), (1, 11025, 2500)]:
        pcm = (rng.standard_normal((n, ch)) * 7000).astype(np.int16)
        w = wav_bytes(pcm, ch, rate)
        a, b = o.parse_wav(w), r.parse_wav(w)
        assert a.shape == b.shape and np.abs(a - b).max() <= 1e-7

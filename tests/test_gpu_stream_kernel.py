"""GPU: tensor-level parity of the kernel the product decodes with (decode_stream_kernel, qasr_stream.cu) and
path-level parity at the BASELINE.json shapes the smaller tests do not reach.

* logits of ONE step of the default decode kernel, for 1, 2 and 4 sequences per launch, against the reference's
  qwen_decoder_forward_logits (qwen_asr_decoder.c:691-783) on the same KV state: max-rel <= 1e-3 (north_star bar
  1e-2), top-1 equal wherever the reference's top-1/top-2 margin exceeds twice the observed error, and the token the
  kernel's own HEAD phase picked equals the argmax of the logits it stored;
* argmax ties across CTAs -> lowest index (qwen_asr_kernels.c:536-541), on a checkpoint with duplicated lm_head rows;
* greedy ids against the compiled reference at configs[4] (1.7B, 30 s, 128 tokens), a configs[2] segment (0.6B, 20 s,
  tokens_cap) and a configs[3] stream session with the reference's real parameters (8 s windows, 4 kept, 2 s chunks,
  32 tokens, long enough for window eviction).
The CPU side of these runs takes a few minutes in total.
"""
import numpy as np
import pytest

import variants
from conftest import prompt_embeds, rel_err

pytestmark = pytest.mark.gpu


def checker(ref_lib, oracle_lib, model_dir):
    """The compiled reference when it travelled, else the pinned C restatement."""
    return (ref_lib or oracle_lib)().load(model_dir)


def _fill_sequences(eng, cpu, pkg, nseq, seed0):
    """Different prompts in the KV caches of sequences 0..nseq-1; returns (last rows, kv_lens, reference logits)."""
    last, kv, want = [], [], []
    for q in range(nseq):
        audio = pkg.synth_audio(0.9 + 0.6 * q, seed=seed0 + q)
        enc = cpu.encode(cpu.mel(audio))
        emb = prompt_embeds(cpu, enc)
        eng.debug_select_seq(q)
        eng.kv_len = 0
        eng.prefill(emb[:-1])
        cpu.kv_len = 0
        cpu.prefill(emb[:-1])
        want.append(cpu.step_logits(emb[-1]))
        last.append(emb[-1])
        kv.append(len(emb) - 1)
    eng.debug_select_seq(0)
    return np.stack(last), kv, want


def _check_logits(toks, logits, hidden, want, eng):
    for q, ref in enumerate(want):
        err = float(np.abs(logits[q] - ref).max())
        assert err <= 1e-3 * float(np.abs(ref).max()), (q, err)
        top = np.argsort(ref)[-2:]
        margin = float(ref[top[1]] - ref[top[0]])
        assert int(toks[q]) == int(np.argmax(logits[q]))          # HEAD-phase winner == argmax of the stored logits
        if margin > 2 * err:
            assert int(toks[q]) == int(top[1])
        # the hidden-state hook and the logits hook describe the same step: logits[r] = E[r] . hidden
        rows = [int(top[1]), 0, 77777, eng.cfg["vocab_size"] - 1]
        for r in rows:
            e = eng.embed(r).astype(np.float64)
            assert abs(float(e @ hidden[q].astype(np.float64)) - float(logits[q][r])) <= 2e-3 * max(1.0, float(np.abs(ref).max()))


@pytest.mark.parametrize("nseq", [1, 2, 4])
def test_stream_kernel_logits_vs_reference_0p6b(gpu06, ref_lib, oracle_lib, model06, pkg, nseq):
    cpu = checker(ref_lib, oracle_lib, model06)
    try:
        last, kv, want = _fill_sequences(gpu06, cpu, pkg, nseq, seed0=300 + 10 * nseq)
        toks, logits, hidden = gpu06.debug_stream_step(last, kv)
        _check_logits(toks, logits, hidden, want, gpu06)
    finally:
        gpu06.debug_select_seq(0)
        cpu.close()


@pytest.mark.parametrize("nseq", [1, 2])
def test_stream_kernel_logits_vs_reference_1p7b(pkg, ref_lib, oracle_lib, model17, nseq):
    eng = pkg.QasrCuda(0).load(model17)
    cpu = checker(ref_lib, oracle_lib, model17)
    try:
        last, kv, want = _fill_sequences(eng, cpu, pkg, nseq, seed0=400 + 10 * nseq)
        toks, logits, hidden = eng.debug_stream_step(last, kv)
        _check_logits(toks, logits, hidden, want, eng)
    finally:
        eng.close()
        cpu.close()


def test_step_logits_entry_point_runs_the_stream_kernel(gpu06, ref_lib, oracle_lib, model06, pkg):
    """qasr_cuda_step_logits (what the shim's qwen_decoder_forward_logits calls) takes the default decode kernel:
    identical bits to the debug hook, and a launch count of one kernel + the state setter."""
    cpu = checker(ref_lib, oracle_lib, model06)
    try:
        last, kv, want = _fill_sequences(gpu06, cpu, pkg, 1, seed0=333)
        _, logits, _ = gpu06.debug_stream_step(last, kv)
        gpu06.kv_len = kv[0]
        n0 = gpu06.launch_count
        again = gpu06.step_logits(last[0])
        assert gpu06.launch_count - n0 <= 3
        assert np.array_equal(again, logits[0])
        assert rel_err(again, want[0]) < 1e-3
    finally:
        cpu.close()


def test_stream_kernel_argmax_ties_lowest_index(pkg, model06, gpu06, ref_lib, oracle_lib):
    """Duplicated lm_head rows give exactly equal logits in different CTAs of the decode kernel; the winner must be the
    lowest index, as in the reference (strict >, ascending scan).  Row t0 (first greedy token) is copied to a HIGH index
    owned by the last CTA -> t0 must still win; row t1 (second token) is copied to a LOW index owned by CTA 0 -> that
    index must win, and since its embedding equals t1's the continuation is unchanged."""
    audio = pkg.synth_audio(1.3, seed=3)
    base = gpu06.transcribe_ids(audio, 6)[0].tolist()
    t0, t1 = base[0], base[1]
    V = gpu06.cfg["vocab_size"]
    hi, lo = V - 3, 5
    if t0 >= V - 1024 or t1 < 1024 or t0 == t1:
        pytest.skip("synthetic ids fall inside the patched CTA ranges")

    def patch(E):
        E[hi] = E[t0]
        E[lo] = E[t1]
    vdir = variants.patched_model_dir(model06, f"ties_{t0}_{t1}", patch, full_vocab=False)
    want = [t0, lo] + base[2:]
    eng = pkg.QasrCuda(0).load(vdir)
    cpu = checker(ref_lib, oracle_lib, vdir)
    try:
        assert cpu.transcribe_ids(audio, 6)[0].tolist() == want           # the reference's own tie rule on this checkpoint
        assert eng.transcribe_ids(audio, 6)[0].tolist() == want           # decode_stream_kernel<1>
        others = [pkg.synth_audio(1.0 + 0.2 * i, seed=700 + i) for i in range(3)]
        got, _ = eng.transcribe_batch([audio] + others, 6)                # decode_stream_kernel<4>: sequence 0 must not change
        assert got[0].tolist() == want
        got2, _ = eng.transcribe_batch([others[0], audio], 6)             # decode_stream_kernel<2>, sequence 1
        assert got2[1].tolist() == want
    finally:
        eng.close()
        cpu.close()


def test_config5_unit_1p7b_30s_128_tokens_vs_reference(pkg, model17, ref_lib, oracle_lib):
    """BASELINE configs[4] unit: Qwen3-ASR-1.7B, one 30 s utterance (T = 390, prefill 404), 128 greedy tokens."""
    audio = pkg.synth_audio(30.0, seed=0)[:480000]
    eng = pkg.QasrCuda(0).load(model17)
    cpu = checker(ref_lib, oracle_lib, model17)
    try:
        ids, info = eng.transcribe_ids(audio, 128)
        want, _ = cpu.transcribe_ids(audio, 128)
        assert info["enc_tokens"] == 390
        assert ids.tolist() == want.tolist()
    finally:
        eng.close()
        cpu.close()


def test_config3_segment_0p6b_20s_vs_reference(pkg, model06, gpu06, ref_lib, oracle_lib):
    """BASELINE configs[2] unit: the first two -S 20 -W 3 segments of the synthetic recording (T ~ 260, 3 encoder
    windows), each with the bench's token cap ceil(4 s) + 8; single-sequence path and the batched path."""
    seg = pkg.segments
    rec = pkg.synth_audio(60.0, seed=0)
    ranges = seg.split_segments(rec, 20.0, 3.0)[:2]
    units = [seg.pad_short(np.ascontiguousarray(rec[a:b], np.float32)) for a, b in ranges]
    caps = [seg.tokens_cap(b - a) for a, b in ranges]
    cpu = checker(ref_lib, oracle_lib, model06)
    try:
        want = [cpu.transcribe_ids(u, c)[0].tolist() for u, c in zip(units, caps)]
    finally:
        cpu.close()
    got = [gpu06.transcribe_ids(u, c) for u, c in zip(units, caps)]
    assert [g[1]["enc_tokens"] for g in got] == [pkg.load_library().qasr_cuda_encoder_tokens(len(u) // 160) for u in units]
    assert [g[0].tolist() for g in got] == want
    batched, _ = gpu06.transcribe_batch(units, caps)
    assert [b.tolist() for b in batched] == want


def test_config4_stream_real_parameters_vs_reference(pkg, model06, gpu06, ref_lib, oracle_lib):
    """BASELINE configs[3]: --stream with the reference's parameters (2 s chunks, 8 s encoder windows, 4 windows kept,
    32 new tokens per chunk, qwen_asr.c:1273-1900) over 44 s = 22 chunks, so the fifth window evicts the first.  The
    device-resident session (qasr_cuda_stream_feed) must give, chunk by chunk, the ids / reused prefix / prompt length
    of the same session driven on the CPU reference."""
    audio = pkg.synth_audio(44.0, seed=12)
    cpu = checker(ref_lib, oracle_lib, model06)
    try:
        want = pkg.streaming.run_stream(cpu, audio, 2.0, window_sec=8.0, max_windows=4, max_new=32)
    finally:
        cpu.close()
    gpu06.stream_begin(8.0, 4)
    got = [gpu06.stream_feed(audio[:end], 32) for end in range(32000, len(audio) + 1, 32000)]
    assert len(got) == len(want) == 22
    for i, (g, w) in enumerate(zip(got, want)):
        assert (g["rows"], g["reused"]) == (w["rows"], w["reused"]), i
        assert g["ids"] == w["ids"], i
    assert max(w["rows"] for w in want) <= 9 + 4 * 104 + 78 + 6          # never more than 4 cached windows + a partial one
    n_pre = len(pkg.streaming.PROMPT_PRE)
    assert any(want[i]["reused"] == n_pre and want[i - 1]["reused"] > n_pre for i in range(1, 22))   # the eviction chunk reuses the prompt prefix only


_EOS_CHILD = r"""
import sys, json
sys.path.insert(0, {root!r})
import __graft_entry__ as ge
pkg = ge.load_package()
eng = pkg.QasrCuda(0).load({vdir!r})
out = [eng.transcribe_ids(pkg.synth_audio(1.0 + 0.37 * i, seed=500 + i), 40)[0].tolist() for i in range(4) for _ in range(2)]
print("IDS " + json.dumps(out))
"""


def test_single_sequence_eos_stop_matches_reference(pkg, model06, ref_lib, oracle_lib):
    """Early stop of the persistent single-sequence kernel (reference qwen_asr.c:788-793): on the EOS-capable checkpoint the launch
    ends before its step budget, the producer warp of decode_rounds_kernel drains what it has in flight, and the next
    utterances on the same context decode normally.  ids equal the reference's, terminating EOS included."""
    vdir = variants.eos_model_dir(model06, 2.5)
    units = [pkg.synth_audio(1.0 + 0.37 * i, seed=500 + i) for i in range(4)]
    cpu = checker(ref_lib, oracle_lib, vdir)
    try:
        want = [cpu.transcribe_ids(u, 40)[0].tolist() for u in units]
    finally:
        cpu.close()
    assert min(len(w) for w in want) < 40
    eng = pkg.QasrCuda(0).load(vdir)
    try:
        for rep in range(2):
            for u, w in zip(units, want):
                assert eng.transcribe_ids(u, 40)[0].tolist() == w
    finally:
        eng.close()
    # the same with several producer warps: every producer must issue its share of the last rounds and drain them (child process: the
    # knobs are read once per process)
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, QASR_SR_PRODUCERS="4", QASR_SR_CHUNK="2048")
    r = subprocess.run([sys.executable, "-c", _EOS_CHILD.format(root=root, vdir=vdir)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    got = json.loads([l for l in r.stdout.splitlines() if l.startswith("IDS ")][-1][4:])
    assert got == [w for w in want for _ in range(2)]


_OTHER_KERNEL = r"""
import sys, json
sys.path.insert(0, {root!r})
import __graft_entry__ as ge
pkg = ge.load_package()
out = {{}}
for variant in ("0.6b", "1.7b"):
    eng = pkg.QasrCuda(0).load(pkg.ensure_model_dir(variant))
    out[variant] = [eng.transcribe_ids(pkg.synth_audio(s, seed=900 + i), 10)[0].tolist() for i, s in enumerate((1.3, 3.64, 7.0))]
    eng.close()
print("IDS " + json.dumps(out))
"""


def test_ring_and_rounds_kernels_decode_the_same_ids(pkg, ref_lib, oracle_lib):
    """Both single-sequence decode kernels are kept (qasr_stream.cu: per-lane cp.async ring, default for the 1.7B dims;
    qasr_stream_r.cu: TMA producer warp + round-major image, default for the 0.6B dims).  Each is forced for BOTH models in
    a child process (the choice is read once per process) and must reproduce the reference's greedy ids."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    want = {}
    for variant in ("0.6b", "1.7b"):
        cpu = checker(ref_lib, oracle_lib, pkg.ensure_model_dir(variant))
        try:
            want[variant] = [cpu.transcribe_ids(pkg.synth_audio(s, seed=900 + i), 10)[0].tolist() for i, s in enumerate((1.3, 3.64, 7.0))]
        finally:
            cpu.close()
    # the third run exercises the optional knobs of the rounds kernel: three producer warps sharing the copies of a round, a two-slot ring
    for kernel, extra in (("ring", {}), ("rounds", {}), ("rounds", {"QASR_SR_PRODUCERS": "3", "QASR_SR_CHUNK": "4096", "QASR_SR_SLOTS": "2", "QASR_SR_L2AHEAD": "0"})):
        env = dict(os.environ, QASR_DECODE_KERNEL=kernel, **extra)
        r = subprocess.run([sys.executable, "-c", _OTHER_KERNEL.format(root=root)], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        got = json.loads([l for l in r.stdout.splitlines() if l.startswith("IDS ")][-1][4:])
        assert got == want, (kernel, extra)

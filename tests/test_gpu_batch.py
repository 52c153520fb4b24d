"""GPU: the batched throughput path (qasr_batch.cu behind qasr_cuda_transcribe_batch) against the CPU reference.

Groups larger than qasr_cuda_max_batch() run the front end per unit, then encoder, prefill and every decode step with the
rows of ALL units in one GEMM (weights read once per group), per-unit attention over a pooled KV cache, and a per-row
argmax.  The reference processes the same units one after the other (transcribe_segment per segment, qwen_asr.c:987);
greedy ids must be identical unit by unit."""
import numpy as np
import pytest

import variants

pytestmark = pytest.mark.gpu


def checker(ref_lib, oracle_lib, model_dir):
    return (ref_lib or oracle_lib)().load(model_dir)


def test_batched_group_ids_vs_reference_0p6b(gpu06, ref_lib, oracle_lib, model06, pkg):
    """11 units of ragged length (0.6-6.3 s: 1-2 encoder windows, tail chunks of every size) and ragged caps."""
    secs = [1.3, 2.6, 0.6, 3.4, 1.9, 6.3, 1.1, 4.05, 2.2, 0.9, 5.0]
    caps = [9, 12, 6, 17, 8, 21, 5, 14, 11, 7, 16]
    units = [pkg.synth_audio(s, seed=140 + i) for i, s in enumerate(secs)]
    assert len(units) > gpu06.max_batch
    cpu = checker(ref_lib, oracle_lib, model06)
    try:
        want = [cpu.transcribe_ids(u, c)[0].tolist() for u, c in zip(units, caps)]
    finally:
        cpu.close()
    got, tm = gpu06.transcribe_batch(units, caps)
    assert [g.tolist() for g in got] == want
    assert tm["decode_ms"] > 0 and tm["enc_ms"] > 0 and tm["prefill_ms"] > 0
    # the single-sequence entry points are unaffected by a batch (sequence 0's cache is the default again)
    assert gpu06.transcribe_ids(units[3], caps[3])[0].tolist() == want[3]
    # group splitting: more units than one group takes (QASR_BATCH_MAX) must give the same ids
    got2, _ = gpu06.transcribe_batch(units + units[:4], caps + caps[:4])
    assert [g.tolist() for g in got2] == want + want[:4]


def test_batched_group_stops_at_eos_like_reference(pkg, model06, ref_lib, oracle_lib):
    """EOS-capable checkpoint variant: sequences of a group end at different steps (or run into their cap); each unit's ids
    must equal the reference's, including the terminating EOS (qwen_asr.c:788-793), while the rest of the group goes on."""
    vdir = variants.eos_model_dir(model06, 2.5)
    units = [pkg.synth_audio(1.0 + 0.37 * i, seed=500 + i) for i in range(9)]
    cpu = checker(ref_lib, oracle_lib, vdir)
    try:
        want = [cpu.transcribe_ids(u, 40)[0].tolist() for u in units]
    finally:
        cpu.close()
    lens = sorted(len(w) for w in want)
    assert lens[0] < 40 and want[int(np.argmin([len(w) for w in want]))][-1] == variants.TOKEN_IM_END
    eng = pkg.QasrCuda(0).load(vdir)
    try:
        got, _ = eng.transcribe_batch(units, 40)
        assert [g.tolist() for g in got] == want
    finally:
        eng.close()


def test_batched_group_ids_vs_reference_1p7b(pkg, model17, ref_lib, oracle_lib):
    secs = [2.0, 3.1, 1.2, 4.4, 2.7]
    units = [pkg.synth_audio(s, seed=160 + i) for i, s in enumerate(secs)]
    cpu = checker(ref_lib, oracle_lib, model17)
    try:
        want = [cpu.transcribe_ids(u, 12)[0].tolist() for u in units]
    finally:
        cpu.close()
    eng = pkg.QasrCuda(0).load(model17)
    try:
        assert len(units) > eng.max_batch
        got, _ = eng.transcribe_batch(units, 12)
        assert [g.tolist() for g in got] == want
    finally:
        eng.close()


def test_batched_custom_prompt(gpu06, ref_lib, oracle_lib, model06, pkg):
    """qasr_cuda_set_prompt applies to every unit of a batched group (system text + forced language tokens)."""
    from conftest import PRE, SUF
    pre = [151644, 8948, 198, 2610, 525, 264, 1273, 13, 151645, 198, 151644, 872, 198, 151669]
    suf = [151670, 151645, 198, 151644, 77091, 198, 11528, 6364, 151704]
    units = [pkg.synth_audio(1.2 + 0.3 * i, seed=180 + i) for i in range(6)]
    cpu = checker(ref_lib, oracle_lib, model06)
    try:
        want = []
        for u in units:
            enc = cpu.encode(cpu.mel(u))
            rows = np.stack([cpu.embed(t) for t in pre] + list(enc) + [cpu.embed(t) for t in suf]).astype(np.float32)
            cpu.kv_len = 0
            cpu.prefill(rows[:-1])
            ids = [cpu.step(rows[-1])]
            for _ in range(5):
                ids.append(cpu.step(cpu.embed(ids[-1])))
            want.append(ids)
    finally:
        cpu.close()
    gpu06.set_prompt(pre, suf)
    try:
        got, _ = gpu06.transcribe_batch(units, 6)
        assert [g.tolist() for g in got] == want
    finally:
        gpu06.set_prompt(PRE, SUF)


def test_batched_config5_shape_vs_reference(pkg, model17, ref_lib, oracle_lib):
    """BASELINE configs[4] through the batched path at its real shape: three 30 s utterances (T = 390, prompt 405 rows each,
    4 encoder windows), Qwen3-ASR-1.7B, 64 greedy tokens - one group of 3 > qasr_cuda_max_batch() = 2."""
    units = [pkg.synth_audio(30.0, seed=i)[:480000] for i in range(3)]
    cpu = checker(ref_lib, oracle_lib, model17)
    try:
        want = [cpu.transcribe_ids(u, 64)[0].tolist() for u in units]
    finally:
        cpu.close()
    eng = pkg.QasrCuda(0).load(model17)
    try:
        assert eng.batch_plan(3) == (1, 3)
        got, _ = eng.transcribe_batch(units, 64)
        assert [g.tolist() for g in got] == want
    finally:
        eng.close()

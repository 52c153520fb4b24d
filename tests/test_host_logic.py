"""CPU: host-side logic - synthetic inputs, segment splitting (reference qwen_asr.c:617-643,
941-970) and the multi-GPU shard/gather plumbing (world_size 2 over gloo)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synth_audio_deterministic(pkg):
    a = pkg.synth_audio(3.0, seed=5)
    b = pkg.synth_audio(3.0, seed=5)
    c = pkg.synth_audio(3.0, seed=6)
    assert a.dtype == np.float32 and len(a) == 48000
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.abs(a).max() < 1.0 and np.abs(a).max() > 0.05
    assert np.array_equal(a, np.round(a * 32768) / 32768)  # s16-quantised like a WAV sample


def test_split_segments_matches_reference_rules(pkg):
    seg = pkg.segments
    audio = pkg.synth_audio(95.0, seed=1)
    ranges = seg.split_segments(audio, 20.0, 3.0)
    assert ranges[0][0] == 0 and ranges[-1][1] == len(audio)
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))  # contiguous, no overlap
    for (a, b) in ranges[:-1]:
        assert 17 * 16000 <= b - a <= 23 * 16000                 # target +- search window
    assert seg.split_segments(audio[:16000 * 10], 20.0, 3.0) == [(0, 160000)]  # fits one segment
    assert seg.split_segments(audio, 0.0, 3.0) == [(0, len(audio))]            # -S 0
    # search window is clamped to half the segment (qwen_asr.c:944-945)
    r2 = seg.split_segments(audio, 4.0, 3.0)
    assert all(b - a >= 2 * 16000 - 800 for a, b in r2[:-1])
    # the reference's 127-split cap (qwen_asr.c:968)
    long_audio = np.tile(audio, 3)
    capped = seg.split_segments(long_audio, 1.0, 0.2)
    assert len(capped) == 127
    assert len(seg.split_segments(long_audio, 1.0, 0.2, max_splits=None)) > 127


def test_find_split_point_picks_silence(pkg):
    seg = pkg.segments
    x = np.full(16000 * 10, 0.1, np.float32)
    x[16000 * 5 + 800:16000 * 5 + 2400] = 0.0
    assert abs(seg.find_split_point(x, 16000 * 5, 3.0) - (16000 * 5 + 1600)) <= 800
    assert seg.pad_short(np.ones(10, np.float32)).shape == (8000,)


def test_shard_ranges_cover_everything(pkg):
    seg = pkg.segments
    for n in (0, 1, 7, 8, 181):
        for world in (1, 2, 4, 8):
            parts = [seg.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import __graft_entry__ as ge
    seg = ge.load_package().segments
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_items = 7
    lo, hi = seg.shard_range(n_items, rank, world)
    local = [([100 * i, 100 * i + 1], {"seg": i, "rank": rank}) for i in range(lo, hi)]  # stand-in for engine output
    merged = seg.gather_in_order(local, n_items, rank, world, dist)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, [m[1]["seg"] for m in merged], [m[0][0] for m in merged]))


def test_two_rank_gloo_gather_in_segment_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, order, firsts in results:
        assert order == list(range(7)), f"rank {rank} saw {order}"
        assert firsts == [100 * i for i in range(7)]


# ---------------------------------------------------------------- streaming session (host logic)
class _FakeEngine:
    """Deterministic stand-in with the duck-typed engine interface (mel/encode/embed/prefill/step/kv_len):
    records what the session asks the device to do."""

    def __init__(self, H=8):
        self.H, self.kv_len, self.log = H, 0, []

    def embed(self, tok):
        return np.full(self.H, float(tok % 97), np.float32)

    def mel(self, samples):
        # like the real front end, every value depends on the whole span (dynamic max normalisation)
        return np.tile(np.float32(samples[: (len(samples) // 160) * 160: 160]) + np.float32(len(samples) * 1e-3), (128, 1))

    def encode(self, mel):
        T = max(1, mel.shape[1] // 8)
        return (mel[:1, : T * 8].reshape(T, 8).mean(axis=1, keepdims=True) * np.ones((1, self.H))).astype(np.float32)

    def prefill(self, embeds):
        self.log.append(("prefill", self.kv_len, len(embeds)))
        self.kv_len += len(embeds)

    def step(self, embed):
        self.log.append(("step", self.kv_len))
        self.kv_len += 1
        return 7


def test_stream_session_reuses_prefix_and_caches_windows(pkg):
    st = pkg.streaming
    rng = np.random.default_rng(0)
    audio = rng.standard_normal(16000 * 5).astype(np.float32)
    eng = _FakeEngine()
    sess = st.StreamSession(eng, window_sec=2.0, max_windows=2, max_new=3)
    r1 = sess.feed(audio[:16000])          # 1 s: only a tail
    assert r1["reused"] == 0 and r1["new_windows"] == 0 and r1["ids"] == [7, 7, 7]
    r2 = sess.feed(audio[:32000])          # 2 s: first full window completes, no tail
    assert r2["new_windows"] == 1 and r2["reused"] == len(st.PROMPT_PRE)   # prompt prefix rows reused
    r3 = sess.feed(audio[:48000])          # 3 s: window 0 cached (rows reused), new tail
    assert r3["new_windows"] == 0
    n_w0 = len(sess.win_rows[0])
    assert r3["reused"] == len(st.PROMPT_PRE) + n_w0
    # every chunk: kv_len rolled back to `reused`, the delta prefilled, last row stepped (qwen_asr.c:1823-1829)
    pre = [e for e in eng.log if e[0] == "prefill"][-1]
    assert pre[1] == r3["reused"] and pre[2] == r3["prefilled"] == r3["rows"] - 1 - r3["reused"]
    sess.feed(audio[:64000])               # 4 s: windows 0, 1
    r5 = sess.feed(audio[:80000])          # 5 s: still windows 0, 1 + tail
    assert sorted(sess.win_rows) == [0, 1] and r5["reused"] >= len(st.PROMPT_PRE) + n_w0
    sess.feed(np.concatenate([audio, audio[:16000]]))   # 6 s: window 2 completes -> window 0 evicted (max_windows=2)
    assert sorted(sess.win_rows) == [1, 2]


def test_common_prefix_rows_is_bitwise(pkg):
    st = pkg.streaming
    a = np.arange(12, dtype=np.float32).reshape(4, 3)
    b = a.copy()
    assert st.common_prefix_rows(a, b) == 4 and st.common_prefix_rows(a, None) == 0
    b[2, 1] = np.nextafter(b[2, 1], np.float32(100))
    assert st.common_prefix_rows(a, b) == 2
    assert st.common_prefix_rows(a[:1], b) == 1


def test_skinny_gemm_split_plan_invariants(pkg):
    """Host-side plan of the weight-streaming GEMM (M <= 256): the split factor fits one portable thread-block cluster,
    no split is empty, every k-block is covered exactly once, and the production shapes get the measured plans."""
    import ctypes as C
    lib = pkg.load_library()
    plan = lib.qasr_debug_gemm_plan
    plan.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    plan.restype = C.c_int
    out = (C.c_int * 6)()

    def get(M, K, N):
        assert plan(M, K, N, out) == 0
        return list(out)

    for M in (1, 13, 32, 33, 47, 61, 64, 65, 128, 129, 143, 256):
        for K in (8, 64, 72, 896, 1024, 2048, 3584, 4320, 6144, 7680):
            for N in (40, 128, 480, 896, 1024, 2048, 4096, 12288, 151936):
                path, MP, tiles, kb, S, per = get(M, K, N)
                assert path == 0 and MP == (32 if M <= 32 else 64 if M <= 64 else 128 if M <= 128 else 256)
                assert tiles == -(-N // 128) and kb == -(-K // 64)
                assert 1 <= S <= 8 and per >= 1
                assert (S - 1) * per < kb <= S * per          # every split owns at least one k-block, all are covered
                if kb >= 4:
                    assert S == 1 or per >= 2                   # at least QASR_GEMM_MIN_KB k-blocks per split
    assert get(257, 1024, 1024)[0] == 1 and get(6240, 1024, 4096)[0] == 1   # large-tile kernel
    assert plan(0, 8, 8, out) != 0 and plan(8, 8, 8, None) != 0
    # bench workload, Qwen3-ASR-1.7B: prefill QKV / WO / gate-up / down and encoder fc2 (tools/gemm_bench.py)
    assert [get(61, 2048, 4096)[4], get(61, 2048, 2048)[4], get(61, 2048, 12288)[4], get(61, 6144, 2048)[4], get(47, 4096, 1024)[4]] == [3, 5, 1, 5, 8]


def test_fused_epilogue_decisions(pkg):
    """Which GEMMs carry the fused RMSNorm / q-k-norm + RoPE epilogues (host logic of gemm_tc_can_fuse_*): the skinny split-K path only -
    at most 256 rows, whole 128-column tiles, a split factor above 1 - so the production chains fuse exactly where DESIGN 3.2 says."""
    import ctypes as C
    lib = pkg.load_library()
    f = lib.qasr_debug_gemm_fusion
    f.argtypes = [C.c_int, C.c_int, C.c_int]
    f.restype = C.c_int
    for H, I in ((1024, 3072), (2048, 6144)):                       # 0.6B / 1.7B decoder dims
        for M in (1, 23, 31, 61, 128, 157, 256):                    # batched decode groups, single-utterance prompts
            assert f(M, 2048, H) & 3 == 3                           # WO -> post-attention norm
            assert f(M, I, H) & 3 == 3                              # down -> next input norm
            assert f(M, H, 4096) & 4                                # QKV: q/k-norm + RoPE + KV store
            assert f(M, H, 2 * I) & 1 and f(M, H, 151936) & 1       # gate/up and lm_head scale their rows
        for M in (257, 274, 404, 25856):                            # long prompts, batched prefill: large-tile kernels, stand-alone norms
            assert f(M, 2048, H) == 0 and f(M, H, 4096) == 0
    assert f(61, 2048, 1000) & 2 == 0                               # ragged N: no whole tiles
    assert f(61, 64, 2048) & 2 == 0                                 # one k-block: no split-K reduction path


def test_find_split_point_equals_scalar_restatement(pkg):
    """The vectorised split search must pick exactly the window the reference's scalar loop picks (qwen_asr.c:617-643:
    100 ms windows every 50 ms inside +-search_sec, strict '<' so the first minimum wins, centre of the window)."""
    seg = pkg.segments

    def scalar(samples, target, search_sec):
        n = len(samples)
        half = int(search_sec * 16000)
        lo, hi = max(0, target - half), min(n, target + half)
        win = 1600
        best_e, best_c = np.float32(1e30), target
        pos = lo
        while pos + win <= hi:
            e = np.float32((samples[pos:pos + win].astype(np.float32) ** 2).sum(dtype=np.float32) / np.float32(win))
            if e < best_e:
                best_e, best_c = e, pos + win // 2
            pos += win // 2
        return best_c
    audio = pkg.synth_audio(70.0, seed=9)
    for target, search in ((16000 * 20, 3.0), (16000 * 33 + 123, 1.5), (16000 * 2, 3.0), (len(audio) - 8000, 3.0), (16000 * 50, 0.04)):
        assert seg.find_split_point(audio, target, search) == scalar(audio, target, search)
    flat = np.zeros(16000 * 8, np.float32)          # all windows tie: the first one wins
    assert seg.find_split_point(flat, 16000 * 4, 1.0) == scalar(flat, 16000 * 4, 1.0)


def test_checkpoint_variant_helpers(tmp_path):
    """tests/variants.py: the vocab in which every id decodes to a distinct string, and the embedding-row patcher."""
    import json
    import struct
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import variants
    b2u = variants.bytes_to_unicode()
    assert len(b2u) == 256 and len(set(b2u.values())) == 256 and b2u[ord("A")] == "A" and b2u[ord(" ")] == "Ġ"
    vp = tmp_path / "vocab.json"
    variants.write_full_vocab(str(vp), vocab_size=1000)
    vocab = json.load(open(vp, encoding="utf-8"))
    assert len(vocab) == 1000 and sorted(vocab.values()) == list(range(1000)) and vocab["t999."] == 999
    # a two-tensor safetensors file: the patcher rewrites rows of the embedding only
    base = tmp_path / "m"
    base.mkdir()
    emb = np.arange(8 * 4, dtype=np.uint16).reshape(8, 4)
    other = np.full(6, 7, np.uint16)
    header = {"x": {"dtype": "BF16", "shape": [6], "data_offsets": [0, 12]},
              variants.EMBED: {"dtype": "BF16", "shape": [8, 4], "data_offsets": [12, 12 + 64]}}
    hj = json.dumps(header).encode()
    with open(base / "model.safetensors", "wb") as f:
        f.write(struct.pack("<Q", len(hj)) + hj + other.tobytes() + emb.tobytes())
    (base / "vocab.json").write_text("{}")
    out = variants.patched_model_dir(str(base), "dup", lambda E: E.__setitem__(5, E[1]), full_vocab=False)
    off, nbytes, shape, dtype = variants.tensor_span(os.path.join(out, "model.safetensors"), variants.EMBED)
    raw = open(os.path.join(out, "model.safetensors"), "rb").read()
    got = np.frombuffer(raw[off:off + nbytes], np.uint16).reshape(shape)
    want = emb.copy()
    want[5] = want[1]
    assert np.array_equal(got, want) and raw[off - 12:off] == other.tobytes()
    assert np.array_equal(variants.f32_to_bf16(variants.bf16_to_f32(emb[2])), emb[2])


def test_batch_plan_is_host_only(pkg):
    """qasr_cuda_batch_plan needs a loaded context; without a GPU it must fail cleanly (no CPU fallback, no crash)."""
    import ctypes as C
    lib = pkg.load_library()
    g, b = C.c_int(0), C.c_int(0)
    assert lib.qasr_cuda_batch_plan(None, 10, C.byref(g), C.byref(b)) != 0

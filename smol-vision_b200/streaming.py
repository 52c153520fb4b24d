"""Host-side streaming session: the device-visible part of the reference's `stream_impl`
(qwen_asr.c:1273-1900) for `--stream` mode, written against a duck-typed engine interface
(mel / encode / embed / prefill / step / kv_len, optionally generate): `QasrCuda` implements it, and so do the
CPU checkers the tests drive through the very same session.

Per chunk of new audio (reference line numbers):
  * every newly completed `window_sec` (8 s) span is mel-normalised on its own span and encoded ONCE,
    its rows are cached (`stream_encode_span` :1122, :1601-1641); only the partial tail is re-encoded
    (:1643-1659); at most `max_windows` (4) full windows are kept (:1670-1683);
  * embeds = prompt prefix rows | cached window rows | tail rows | prompt suffix rows (:1756-1805);
  * rows equal to the previous chunk's rows are reused: kv_len = longest common prefix (:1811-1823),
    only the delta is prefilled (:1825-1829), the last row goes through the single-token step;
  * up to `max_new` (32) greedy tokens are decoded (:1880-1888).
The token bookkeeping that follows in the reference (rollback, LCP commit, overlap dedup, :1906-2146) is pure
host integer work on the returned ids and is not part of this path.
"""
import numpy as np

SAMPLE_RATE = 16000
PROMPT_PRE = (151644, 8948, 198, 151645, 198, 151644, 872, 198, 151669)   # qwen_asr.c:388-393
PROMPT_SUF = (151670, 151645, 198, 151644, 77091, 198)                    # qwen_asr.c:394-396
EOS = (151643, 151645)


def common_prefix_rows(a, b):
    """Number of leading rows of `a` that are bit-identical to the rows of `b` (the reference memcmp's
    float rows, qwen_asr.c:1811-1821)."""
    if a is None or b is None:
        return 0
    n = min(len(a), len(b))
    if n == 0:
        return 0
    same = (a[:n].view(np.uint32) == b[:n].view(np.uint32)).all(axis=1)
    bad = np.flatnonzero(~same)
    return int(bad[0]) if len(bad) else n


class StreamSession:
    def __init__(self, engine, window_sec=8.0, max_windows=4, max_new=32, pre_ids=PROMPT_PRE, suf_ids=PROMPT_SUF, enc_cache=True):
        self.eng = engine
        self.enc_cache = enc_cache    # False = re-encode every window on every chunk (reference QWEN_STREAM_NO_ENC_CACHE=1, qwen_asr.c:1352)
        self.window = int(round(window_sec * SAMPLE_RATE))
        self.max_windows = max_windows
        self.max_new = max_new
        self.pre = np.stack([engine.embed(t) for t in pre_ids]).astype(np.float32)
        self.suf = np.stack([engine.embed(t) for t in suf_ids]).astype(np.float32)
        self.win_rows = {}        # window index -> encoder rows [T, H]
        self.prev_embeds = None

    def _encode_span(self, span):
        """mel on the span alone (dynamic max of that span), then the encoder (stream_encode_span :1122-1126)."""
        if len(span) < 160:  # fewer than one frame: the reference's mel returns NULL (qwen_asr_audio.c:313-317)
            return None
        mel = self.eng.mel(np.ascontiguousarray(span, np.float32))
        if mel is None or mel.shape[1] == 0:
            return None
        return np.array(self.eng.encode(mel), np.float32)

    def feed(self, samples):
        """`samples` = ALL audio received so far. Returns dict(ids, reused, prefilled, rows, new_windows)."""
        n_full = len(samples) // self.window
        new_windows = 0
        for w in range(max(0, n_full - self.max_windows), n_full):
            if w not in self.win_rows or not self.enc_cache:
                self.win_rows[w] = self._encode_span(samples[w * self.window:(w + 1) * self.window])
                new_windows += 1
        for w in [k for k in self.win_rows if k < n_full - self.max_windows]:
            del self.win_rows[w]  # evicted beyond max_windows
        parts = [self.pre]
        for w in range(max(0, n_full - self.max_windows), n_full):
            parts.append(self.win_rows[w])
        tail_n = len(samples) - n_full * self.window
        if 0 < tail_n < 160:  # the reference skips the chunk when the partial span cannot be encoded (qwen_asr.c:1643-1665)
            return dict(ids=[], reused=0, prefilled=0, rows=0, new_windows=new_windows)
        tail = self._encode_span(samples[n_full * self.window:])
        if tail is not None:
            parts.append(tail)
        if len(parts) == 1:   # no encoder rows at all: skipped as well (qwen_asr.c:1686-1691)
            return dict(ids=[], reused=0, prefilled=0, rows=0, new_windows=new_windows)
        parts.append(self.suf)
        embeds = np.ascontiguousarray(np.concatenate(parts, axis=0), np.float32)
        total = len(embeds)
        reused = min(common_prefix_rows(embeds, self.prev_embeds), total - 1)
        self.eng.kv_len = reused                       # rollback moves no data (qwen_asr.c:1823)
        if total - 1 > reused:
            self.eng.prefill(embeds[reused:total - 1])
        tok = self.eng.step(embeds[total - 1])
        if hasattr(self.eng, "generate"):
            ids = list(self.eng.generate(tok, self.max_new))
        else:
            ids = [tok]
            while len(ids) < self.max_new and ids[-1] not in EOS:
                ids.append(self.eng.step(self.eng.embed(ids[-1])))
        self.prev_embeds = embeds
        return dict(ids=[int(t) for t in ids], reused=reused, prefilled=total - 1 - reused, rows=total,
                    new_windows=new_windows)


def run_stream(engine, samples, chunk_sec=2.0, **kw):
    """Feed `samples` in `chunk_sec` steps; returns the per-chunk results."""
    sess = StreamSession(engine, **kw)
    step = int(round(chunk_sec * SAMPLE_RATE))
    out = []
    for end in range(step, len(samples) + step, step):
        out.append(sess.feed(samples[:min(end, len(samples))]))
    return out

"""ctypes binding of the C ABI in include/qasr_cuda.h (libqasr_cuda.so).

`QasrCuda` is the Python mirror of the reference's hot-path entry points
(qwen_mel_spectrogram / qwen_encoder_forward / qwen_decoder_prefill / qwen_decoder_forward /
qwen_decoder_forward_logits, reference qwen_asr.h:351-362, qwen_asr_audio.h:32) with the same
argument meaning and the reference's caller-owned `kv_len` state contract, so parity tests read
like calls into the reference.  There is no CPU fallback: if the shared library is missing, has
not been built for sm_100a, or no B200 is visible, construction raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libqasr_cuda.so")

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
vp, ci, cf = C.c_void_p, C.c_int, C.c_float
ip = C.POINTER(C.c_int)

CFG_KEYS = ["enc_d_model", "enc_layers", "enc_heads", "enc_ffn_dim", "enc_output_dim", "dec_hidden",
            "dec_layers", "dec_heads", "dec_kv_heads", "dec_head_dim", "dec_intermediate", "vocab_size"]

# name -> (restype, argtypes); mirrors include/qasr_cuda.h one to one
SIGNATURES = {
    "qasr_cuda_device_count": (ci, []),
    "qasr_cuda_last_error": (C.c_char_p, []),
    "qasr_cuda_init": (vp, [ci]),
    "qasr_cuda_free": (None, [vp]),
    "qasr_cuda_load_dir": (ci, [vp, C.c_char_p]),
    "qasr_cuda_upload_tensors": (ci, [vp, vp, ci]),
    "qasr_cuda_config": (ci, [vp, i32p]),
    "qasr_cuda_set_gemm_split": (ci, [vp, ci]),
    "qasr_cuda_mel_frames": (ci, [ci]),
    "qasr_cuda_mel": (ci, [vp, f32p, ci, vp, ip]),
    "qasr_cuda_encode": (ci, [vp, vp, ci, vp, ip]),
    "qasr_cuda_encoder_tokens": (ci, [ci]),
    "qasr_cuda_prefill_embeds": (ci, [vp, f32p, ci, ci]),
    "qasr_cuda_prefill_prompt": (ci, [vp, i32p, ci, ci, i32p, ci, ci]),
    "qasr_cuda_step_embed": (ci, [vp, f32p, ci, ip]),
    "qasr_cuda_step_token": (ci, [vp, ci, ci, ip]),
    "qasr_cuda_step_pending": (ci, [vp, ci, ip]),
    "qasr_cuda_step_logits": (ci, [vp, f32p, ci, f32p]),
    "qasr_cuda_generate": (ci, [vp, ci, ci, ci, i32p, ip, ip]),
    "qasr_cuda_transcribe_ids": (ci, [vp, f32p, ci, ci, i32p, ip, vp, ip]),
    "qasr_cuda_set_prompt": (ci, [vp, i32p, ci, i32p, ci]),
    "qasr_cuda_max_batch": (ci, [vp]),
    "qasr_cuda_batch_plan": (ci, [vp, ci, ip, ip]),
    "qasr_cuda_transcribe_batch": (ci, [vp, vp, i32p, ci, i32p, ci, i32p, i32p, vp]),
    "qasr_cuda_stream_begin": (ci, [vp, cf, ci]),
    "qasr_cuda_stream_feed": (ci, [vp, f32p, ci, ci, i32p, ip, ip, ip]),
    "qasr_cuda_stage_audio": (ci, [vp, f32p, ci]),
    "qasr_cuda_decode_pcm16": (ci, [vp, vp, ci, ci, ci, vp, ci, ip]),
    "qasr_cuda_transcribe_staged": (ci, [vp, ci, i32p, ip, vp, ip]),
    "qasr_cuda_timer_start": (ci, [vp]),
    "qasr_cuda_timer_stop": (ci, [vp, C.POINTER(C.c_double)]),
    "qasr_cuda_decode_stats": (ci, [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_double), ci]),
    "qasr_cuda_read_kv": (ci, [vp, ci, ci, f32p, f32p]),
    "qasr_cuda_embed_token": (ci, [vp, ci, f32p]),
    "qasr_cuda_last_decode_ms": (C.c_double, [vp]),
    "qasr_cuda_launch_count": (C.c_longlong, [vp]),
    "qasr_op_linear": (ci, [vp, f32p, f32p, f32p, vp, ci, ci, ci]),
    "qasr_op_matmul_t": (ci, [vp, f32p, f32p, f32p, ci, ci, ci]),
    "qasr_op_linear_bf16": (ci, [vp, f32p, f32p, u16p, vp, ci, ci, ci]),
    "qasr_op_matmul_t_bf16": (ci, [vp, f32p, f32p, u16p, ci, ci, ci]),
    "qasr_op_linear_nobias_bf16_qkv": (ci, [vp, f32p, f32p, f32p, f32p, u16p, u16p, u16p, ci, ci, ci]),
    "qasr_op_argmax_matvec_bf16": (ci, [vp, f32p, u16p, ci, ci, ip]),
    "qasr_op_conv2d": (ci, [vp, f32p, f32p, f32p, vp, ci, ci, ci, ci, ci, ci, ci, ci]),
    "qasr_op_layer_norm": (ci, [vp, f32p, f32p, f32p, f32p, ci, ci, cf]),
    "qasr_op_rms_norm": (ci, [vp, f32p, f32p, f32p, ci, ci, cf]),
    "qasr_op_rms_norm_per_head": (ci, [vp, f32p, f32p, ci, ci, ci, cf]),
    "qasr_op_gelu": (ci, [vp, f32p, ci]),
    "qasr_op_silu": (ci, [vp, f32p, ci]),
    "qasr_op_softmax": (ci, [vp, f32p, ci, ci]),
    "qasr_op_swiglu_multiply": (ci, [vp, f32p, f32p, ci, ci]),
    "qasr_op_bidirectional_attention": (ci, [vp, f32p, f32p, f32p, f32p, ci, ci, ci, cf, i32p, ci]),
    "qasr_op_causal_attention": (ci, [vp, f32p, f32p, f32p, f32p, ci, ci, ci, ci, ci, cf, ci]),
    "qasr_op_sinusoidal_pe": (ci, [vp, f32p, ci, ci]),
    "qasr_op_compute_rope_neox": (ci, [vp, f32p, f32p, i32p, ci, ci, cf]),
    "qasr_op_apply_rope_neox": (ci, [vp, f32p, f32p, f32p, ci, ci, ci]),
    "qasr_op_add_inplace": (ci, [vp, f32p, f32p, ci]),
    "qasr_op_mul_inplace": (ci, [vp, f32p, f32p, ci]),
    "qasr_op_scale": (ci, [vp, f32p, cf, ci]),
    "qasr_op_copy": (ci, [vp, f32p, f32p, ci]),
    "qasr_set_threads": (None, [ci]),
    "qasr_get_num_cpus": (ci, []),
}


class QasrError(RuntimeError):
    pass


def load_library(path=None):
    """dlopen libqasr_cuda.so and attach the header's signatures. Raises if it is missing."""
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise QasrError(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback for this path)")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the header's symbol
        fn.restype = res
        fn.argtypes = args
    return lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _opt(a):
    """optional float32 array -> void pointer (None -> NULL)"""
    return None if a is None else a.ctypes.data_as(vp)


class QasrCuda:
    """One device context = one sequence = one KV cache (like the reference's qwen_ctx_t)."""

    def __init__(self, device=0, lib=None):
        self.lib = lib or load_library()
        if self.lib.qasr_cuda_device_count() <= 0:
            raise QasrError("no CUDA device visible; libqasr_cuda has no CPU fallback")
        self.ctx = self.lib.qasr_cuda_init(device)
        if not self.ctx:
            raise QasrError(self._err())
        self.device = device
        self.cfg = None
        self.kv_len = 0  # caller-owned position, reference ctx->kv_cache_len (qwen_asr.h:207)

    def _err(self):
        return (self.lib.qasr_cuda_last_error() or b"").decode(errors="replace")

    def _ck(self, rc):
        if rc != 0:
            raise QasrError(f"libqasr_cuda error {rc}: {self._err()}")

    def close(self):
        if self.ctx:
            self.lib.qasr_cuda_free(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- lifecycle
    def load(self, model_dir, threads=0):
        self._ck(self.lib.qasr_cuda_load_dir(self.ctx, model_dir.encode()))
        cfg = np.zeros(12, np.int32)
        self._ck(self.lib.qasr_cuda_config(self.ctx, cfg))
        self.cfg = dict(zip(CFG_KEYS, (int(v) for v in cfg)))
        return self

    def set_gemm_split(self, nsplit):
        self._ck(self.lib.qasr_cuda_set_gemm_split(self.ctx, nsplit))

    @property
    def launch_count(self):
        return int(self.lib.qasr_cuda_launch_count(self.ctx))

    @property
    def last_decode_ms(self):
        return float(self.lib.qasr_cuda_last_decode_ms(self.ctx))

    # ---- level 1 (the checkers used by tests expose the same method names)
    def mel(self, samples, want_host=True):
        samples = _f32(samples)
        frames = self.lib.qasr_cuda_mel_frames(len(samples))
        out = np.empty((128, max(frames, 1)), np.float32) if want_host else None
        fr = ci(0)
        self._ck(self.lib.qasr_cuda_mel(self.ctx, samples, len(samples), _opt(out), C.byref(fr)))
        return out[:, :fr.value] if want_host else fr.value

    def encode(self, mel=None, want_host=True, frames=None):
        """mel=None consumes the device-resident mel of the last mel() call."""
        if mel is not None:
            mel = _f32(mel)
            frames = mel.shape[1]
        elif frames is None:
            frames = 0
        T_pred = self.lib.qasr_cuda_encoder_tokens(frames) if frames else 4096
        out = np.empty((T_pred, self.cfg["enc_output_dim"]), np.float32) if want_host else None
        T = ci(0)
        self._ck(self.lib.qasr_cuda_encode(self.ctx, _opt(mel), frames, _opt(out), C.byref(T)))
        return out[:T.value] if want_host else T.value

    def prefill(self, embeds):
        embeds = _f32(embeds)
        self._ck(self.lib.qasr_cuda_prefill_embeds(self.ctx, embeds, embeds.shape[0], self.kv_len))
        self.kv_len += embeds.shape[0]

    def prefill_prompt(self, pre_ids, n_audio, suf_ids):
        pre = np.ascontiguousarray(pre_ids, np.int32)
        suf = np.ascontiguousarray(suf_ids, np.int32)
        self._ck(self.lib.qasr_cuda_prefill_prompt(self.ctx, pre, len(pre), n_audio, suf, len(suf), self.kv_len))
        self.kv_len += len(pre) + n_audio + len(suf) - 1

    def step(self, embed):
        tok = ci(0)
        self._ck(self.lib.qasr_cuda_step_embed(self.ctx, _f32(embed), self.kv_len, C.byref(tok)))
        self.kv_len += 1
        return tok.value

    def step_token(self, token_id):
        tok = ci(0)
        self._ck(self.lib.qasr_cuda_step_token(self.ctx, int(token_id), self.kv_len, C.byref(tok)))
        self.kv_len += 1
        return tok.value

    def step_pending(self):
        tok = ci(0)
        self._ck(self.lib.qasr_cuda_step_pending(self.ctx, self.kv_len, C.byref(tok)))
        self.kv_len += 1
        return tok.value

    def step_logits(self, embed):
        out = np.empty(self.cfg["vocab_size"], np.float32)
        self._ck(self.lib.qasr_cuda_step_logits(self.ctx, _f32(embed), self.kv_len, out))
        self.kv_len += 1
        return out

    def generate(self, first_token, max_new):
        ids = np.zeros(max(max_new, 1), np.int32)
        n, kv = ci(0), ci(0)
        self._ck(self.lib.qasr_cuda_generate(self.ctx, int(first_token), self.kv_len, max_new, ids, C.byref(n), C.byref(kv)))
        self.kv_len = kv.value
        return ids[:n.value].copy()

    def embed(self, tok):
        out = np.empty(self.cfg["dec_hidden"], np.float32)
        self._ck(self.lib.qasr_cuda_embed_token(self.ctx, int(tok), out))
        return out

    def read_kv(self, layer, length):
        kvd = self.cfg["dec_kv_heads"] * self.cfg["dec_head_dim"]
        k = np.empty((length, kvd), np.float32)
        v = np.empty((length, kvd), np.float32)
        self._ck(self.lib.qasr_cuda_read_kv(self.ctx, layer, length, k, v))
        return k, v

    def transcribe_ids(self, samples, max_new):
        samples = _f32(samples)
        ids = np.zeros(max(max_new, 1), np.int32)
        tm = np.zeros(4, np.float64)
        n, T = ci(0), ci(0)
        self._ck(self.lib.qasr_cuda_transcribe_ids(self.ctx, samples, len(samples), max_new, ids, C.byref(n),
                                                   tm.ctypes.data_as(vp), C.byref(T)))
        return ids[:n.value].copy(), dict(mel_ms=tm[0], enc_ms=tm[1], prefill_ms=tm[2], decode_ms=tm[3],
                                          enc_tokens=T.value)

    def set_prompt(self, pre_ids, suf_ids):
        """Prompt tokens around the audio rows for transcribe_ids / transcribe_batch (system text, forced language)."""
        pre = np.ascontiguousarray(pre_ids, np.int32)
        suf = np.ascontiguousarray(suf_ids, np.int32)
        self._ck(self.lib.qasr_cuda_set_prompt(self.ctx, pre, len(pre), suf, len(suf)))

    @property
    def max_batch(self):
        return int(self.lib.qasr_cuda_max_batch(self.ctx))

    def batch_plan(self, count):
        """(number of groups, size of the largest group) a transcribe_batch call with `count` units would use."""
        g, b = ci(0), ci(0)
        self._ck(self.lib.qasr_cuda_batch_plan(self.ctx, int(count), C.byref(g), C.byref(b)))
        return g.value, b.value

    def transcribe_batch(self, units, max_new):
        """Independent segments / utterances decoded together (up to `max_batch` per decode step).
        units: list of f32 sample arrays; max_new: int or per-unit list.  Returns ([ids_i], timings dict)."""
        units = [_f32(u) for u in units]
        n = len(units)
        caps = np.ascontiguousarray([max_new] * n if np.isscalar(max_new) else list(max_new), np.int32)
        stride = int(caps.max()) if n else 1
        ptrs = (C.c_void_p * max(n, 1))(*[u.ctypes.data for u in units])
        lens = np.ascontiguousarray([len(u) for u in units], np.int32)
        ids = np.zeros((max(n, 1), stride), np.int32)
        cnt = np.zeros(max(n, 1), np.int32)
        tm = np.zeros(4, np.float64)
        self._ck(self.lib.qasr_cuda_transcribe_batch(self.ctx, C.cast(ptrs, vp), lens, n, caps, stride, ids, cnt, tm.ctypes.data_as(vp)))
        return [ids[i, :cnt[i]].copy() for i in range(n)], dict(mel_ms=tm[0], enc_ms=tm[1], prefill_ms=tm[2], decode_ms=tm[3])

    def stream_begin(self, window_sec=8.0, max_windows=4):
        self._ck(self.lib.qasr_cuda_stream_begin(self.ctx, float(window_sec), int(max_windows)))

    def stream_feed(self, samples, max_new=32):
        """Device-resident streaming chunk: `samples` = all audio so far.  Returns dict(ids, reused, rows)."""
        samples = _f32(samples)
        ids = np.zeros(max(max_new, 1), np.int32)
        n, reused, rows = ci(0), ci(0), ci(0)
        self._ck(self.lib.qasr_cuda_stream_feed(self.ctx, samples, len(samples), max_new, ids, C.byref(n), C.byref(reused), C.byref(rows)))
        return dict(ids=[int(t) for t in ids[:n.value]], reused=reused.value, rows=rows.value)

    # ---- test hooks of the batched decode kernel (exports outside the public header)
    def debug_select_seq(self, q):
        """Route prefill / read_kv / step of this context to the KV cache of sequence q (0..max_batch-1)."""
        f = self.lib.qasr_debug_select_seq
        f.restype, f.argtypes = ci, [vp, ci]
        self._ck(f(self.ctx, int(q)))

    def debug_stream_step(self, embeds, kv_lens, want_logits=True, want_hidden=True):
        """ONE step of decode_stream_kernel<nseq> on sequences 0..nseq-1: (tokens, logits [nseq, V], hidden [nseq, H])."""
        f = self.lib.qasr_debug_stream_step
        f.restype, f.argtypes = ci, [vp, ci, f32p, i32p, i32p, vp, vp]
        embeds = _f32(embeds)
        nseq = embeds.shape[0]
        kv = np.ascontiguousarray(kv_lens, np.int32)
        toks = np.zeros(nseq, np.int32)
        logits = np.empty((nseq, self.cfg["vocab_size"]), np.float32) if want_logits else None
        hidden = np.empty((nseq, self.cfg["dec_hidden"]), np.float32) if want_hidden else None
        self._ck(f(self.ctx, nseq, embeds, kv, toks, _opt(logits), _opt(hidden)))
        return toks, logits, hidden

    # ---- benchmark plumbing
    def stage_audio(self, samples):
        samples = _f32(samples)
        self._ck(self.lib.qasr_cuda_stage_audio(self.ctx, samples, len(samples)))

    def decode_pcm16(self, pcm, channels, sample_rate, want_host=True):
        """Interleaved int16 PCM -> f32 mono 16 kHz (device resampler); the result is also left staged."""
        pcm = np.ascontiguousarray(pcm, np.int16).reshape(-1)
        n_frames = len(pcm) // channels
        cap = n_frames if sample_rate == 16000 else int(n_frames * 16000 // sample_rate)
        out = np.empty(max(cap, 1), np.float32) if want_host else None
        n = ci(0)
        self._ck(self.lib.qasr_cuda_decode_pcm16(self.ctx, pcm.ctypes.data_as(vp), n_frames, channels, sample_rate,
                                                 out.ctypes.data_as(vp) if want_host else None, cap, C.byref(n)))
        return out[:n.value] if want_host else n.value

    def transcribe_staged(self, max_new, ids_buf=None):
        ids = ids_buf if ids_buf is not None else np.zeros(max(max_new, 1), np.int32)
        tm = np.zeros(4, np.float64)
        n, T = ci(0), ci(0)
        self._ck(self.lib.qasr_cuda_transcribe_staged(self.ctx, max_new, ids, C.byref(n), tm.ctypes.data_as(vp), C.byref(T)))
        return ids[:n.value], dict(mel_ms=tm[0], enc_ms=tm[1], prefill_ms=tm[2], decode_ms=tm[3], enc_tokens=T.value)

    def timer_start(self):
        self._ck(self.lib.qasr_cuda_timer_start(self.ctx))

    def timer_stop(self):
        ms = C.c_double(0.0)
        self._ck(self.lib.qasr_cuda_timer_stop(self.ctx, C.byref(ms)))
        return ms.value

    def decode_stats(self, reset=False):
        steps, ms = C.c_longlong(0), C.c_double(0.0)
        self._ck(self.lib.qasr_cuda_decode_stats(self.ctx, C.byref(steps), C.byref(ms), 1 if reset else 0))
        return steps.value, ms.value

    # ---- level 2 operator surface (names follow the reference's qwen_* ops)
    def linear(self, x, W, b=None):
        x, W = _f32(x), _f32(W)
        y = np.empty((x.shape[0], W.shape[0]), np.float32)
        bb = None if b is None else _f32(b)
        self._ck(self.lib.qasr_op_linear(self.ctx, y, x, W, _opt(bb), x.shape[0], x.shape[1], W.shape[0]))
        return y

    def linear_bf16(self, x, W_bf16, b=None):
        x = _f32(x)
        W = np.ascontiguousarray(W_bf16, np.uint16)
        y = np.empty((x.shape[0], W.shape[0]), np.float32)
        bb = None if b is None else _f32(b)
        self._ck(self.lib.qasr_op_linear_bf16(self.ctx, y, x, W, _opt(bb), x.shape[0], x.shape[1], W.shape[0]))
        return y

    def linear_nobias_bf16_qkv(self, x, Wq, Wk, Wv):
        x = _f32(x)
        Wq, Wk, Wv = (np.ascontiguousarray(w, np.uint16) for w in (Wq, Wk, Wv))
        q = np.empty(Wq.shape[0], np.float32)
        k = np.empty(Wk.shape[0], np.float32)
        v = np.empty(Wv.shape[0], np.float32)
        self._ck(self.lib.qasr_op_linear_nobias_bf16_qkv(self.ctx, q, k, v, x, Wq, Wk, Wv, x.shape[-1], Wq.shape[0], Wk.shape[0]))
        return q, k, v

    def argmax_matvec_bf16(self, x, W_bf16):
        x = _f32(x)
        W = np.ascontiguousarray(W_bf16, np.uint16)
        idx = ci(-1)
        self._ck(self.lib.qasr_op_argmax_matvec_bf16(self.ctx, x, W, W.shape[1], W.shape[0], C.byref(idx)))
        return idx.value

    def conv2d(self, x, w, bias, stride, padding):
        x, w = _f32(x), _f32(w)
        c_in, h, wd = x.shape
        c_out, _, kh, kw = w.shape
        ho, wo = (h + 2 * padding - kh) // stride + 1, (wd + 2 * padding - kw) // stride + 1
        out = np.empty((c_out, ho, wo), np.float32)
        bb = None if bias is None else _f32(bias)
        self._ck(self.lib.qasr_op_conv2d(self.ctx, out, x, w, _opt(bb), c_in, c_out, h, wd, kh, kw, stride, padding))
        return out

    def layer_norm(self, x, w, b, eps):
        x = _f32(x)
        out = np.empty_like(x)
        self._ck(self.lib.qasr_op_layer_norm(self.ctx, out, x, _f32(w), _f32(b), x.shape[0], x.shape[1], eps))
        return out

    def rms_norm(self, x, w, eps):
        x = _f32(x)
        out = np.empty_like(x)
        self._ck(self.lib.qasr_op_rms_norm(self.ctx, out, x, _f32(w), x.shape[0], x.shape[1], eps))
        return out

    def rms_norm_per_head(self, x, w, n_heads, head_dim, eps):
        x = _f32(x).copy()
        self._ck(self.lib.qasr_op_rms_norm_per_head(self.ctx, x, _f32(w), x.shape[0], n_heads, head_dim, eps))
        return x

    def _inplace(self, fn, x, *extra):
        x = _f32(x).copy()
        self._ck(fn(self.ctx, x, *extra, x.size))
        return x

    def gelu(self, x):
        return self._inplace(self.lib.qasr_op_gelu, x)

    def silu(self, x):
        return self._inplace(self.lib.qasr_op_silu, x)

    def add(self, a, b):
        return self._inplace(self.lib.qasr_op_add_inplace, a, _f32(b))

    def mul(self, a, b):
        return self._inplace(self.lib.qasr_op_mul_inplace, a, _f32(b))

    def scale(self, x, s):
        return self._inplace(self.lib.qasr_op_scale, x, float(s))

    def copy(self, x):
        x = _f32(x)
        out = np.empty_like(x)
        self._ck(self.lib.qasr_op_copy(self.ctx, out, x, x.size))
        return out

    def softmax(self, x):
        x = _f32(x).copy()
        self._ck(self.lib.qasr_op_softmax(self.ctx, x, x.shape[0], x.shape[1]))
        return x

    def swiglu_multiply(self, gate_up):
        g = _f32(gate_up)
        out = np.empty((g.shape[0], g.shape[1] // 2), np.float32)
        self._ck(self.lib.qasr_op_swiglu_multiply(self.ctx, out, g, g.shape[0], g.shape[1] // 2))
        return out

    def bidirectional_attention(self, Q, K, V, n_heads, head_dim, scale, window_starts):
        Q, K, V = _f32(Q), _f32(K), _f32(V)
        ws = np.ascontiguousarray(window_starts, np.int32)
        out = np.empty_like(Q)
        self._ck(self.lib.qasr_op_bidirectional_attention(self.ctx, out, Q, K, V, Q.shape[0], n_heads, head_dim,
                                                          scale, ws, len(ws) - 1))
        return out

    def causal_attention(self, Q, K, V, n_heads, n_kv_heads, head_dim, scale, q_offset):
        Q, K, V = _f32(Q), _f32(K), _f32(V)
        out = np.empty_like(Q)
        self._ck(self.lib.qasr_op_causal_attention(self.ctx, out, Q, K, V, Q.shape[0], K.shape[0], n_heads,
                                                   n_kv_heads, head_dim, scale, q_offset))
        return out

    def sinusoidal_pe(self, n_pos, d_model):
        out = np.empty((n_pos, d_model), np.float32)
        self._ck(self.lib.qasr_op_sinusoidal_pe(self.ctx, out, n_pos, d_model))
        return out

    def compute_rope_neox(self, positions, head_dim, theta):
        pos = np.ascontiguousarray(positions, np.int32)
        c = np.empty((len(pos), head_dim), np.float32)
        s = np.empty((len(pos), head_dim), np.float32)
        self._ck(self.lib.qasr_op_compute_rope_neox(self.ctx, c, s, pos, len(pos), head_dim, theta))
        return c, s

    def apply_rope_neox(self, x, cos_vals, sin_vals, n_heads, head_dim):
        x = _f32(x).copy()
        self._ck(self.lib.qasr_op_apply_rope_neox(self.ctx, x, _f32(cos_vals), _f32(sin_vals), x.shape[0], n_heads, head_dim))
        return x

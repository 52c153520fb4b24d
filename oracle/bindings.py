"""ctypes bindings for the CHECKERS under oracle/ (test infrastructure only).

`RefLib`    -> oracle/_ref/libqasr_ref_{v3,v4}.so : the unmodified reference sources
               (compiled by oracle/Makefile from /root/reference) behind oracle/ref_harness.c.
`OracleLib` -> oracle/libqasr_oracle.so           : the plain-C restatement (oracle/qasr_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (smol-vision_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def _cpu_has_avx512():
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        return all(k in txt for k in (" avx512f", " avx512bw", " avx512vl", " avx512dq", " avx512cd"))
    except OSError:
        return False


def ref_lib_path():
    """Prebuilt reference library matching this host's ISA, or None."""
    order = ["v4", "v3"] if _cpu_has_avx512() else ["v3"]
    for v in order:
        p = os.path.join(HERE, "_ref", f"libqasr_ref_{v}.so")
        if os.path.exists(p):
            return p
    return None


def shim_lib_path():
    """oracle/_ref/libqasr_ref_cuda.so: the unmodified reference host code linked against
    shim/qwen_asr_cuda_shim.c + libqasr_cuda.so (the drop-in check), or None when it was not built."""
    p = os.path.join(HERE, "_ref", "libqasr_ref_cuda.so")
    return p if os.path.exists(p) else None


class RefLib:
    """The reference's own CPU implementation (kind = "reference").  With path=shim_lib_path() the same
    harness drives the reference's host code on top of the B200 path instead."""

    def __init__(self, path=None):
        path = path or ref_lib_path()
        if path is None:
            raise FileNotFoundError("oracle/_ref/libqasr_ref_*.so not built (run `make -C oracle ref`)")
        self.path = path
        L = self.lib = C.CDLL(path)
        L.ref_load.restype = C.c_void_p
        L.ref_load.argtypes = [C.c_char_p, C.c_int]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_threads_used.restype = C.c_int
        L.ref_threads_used.argtypes = [C.c_int]
        L.ref_config.argtypes = [C.c_void_p, i32p]
        L.ref_mel.restype = C.POINTER(C.c_float)
        L.ref_mel.argtypes = [f32p, C.c_int, C.POINTER(C.c_int)]
        L.ref_encode.restype = C.POINTER(C.c_float)
        L.ref_encode.argtypes = [C.c_void_p, f32p, C.c_int, C.POINTER(C.c_int)]
        L.ref_free_buf.argtypes = [C.c_void_p]
        L.ref_set_kv_len.argtypes = [C.c_void_p, C.c_int]
        L.ref_get_kv_len.restype = C.c_int
        L.ref_get_kv_len.argtypes = [C.c_void_p]
        L.ref_prefill.argtypes = [C.c_void_p, f32p, C.c_int]
        L.ref_step.restype = C.c_int
        L.ref_step.argtypes = [C.c_void_p, f32p]
        L.ref_step_logits.argtypes = [C.c_void_p, f32p, f32p]
        L.ref_embed_token.argtypes = [C.c_void_p, C.c_int, f32p]
        L.ref_read_kv.argtypes = [C.c_void_p, C.c_int, C.c_int, f32p, f32p]
        L.ref_transcribe_ids.restype = C.c_int
        L.ref_transcribe_ids.argtypes = [C.c_void_p, f32p, C.c_int, C.c_int, i32p, f64p, C.POINTER(C.c_int)]
        L.ref_transcribe_text.restype = C.c_void_p
        L.ref_transcribe_text.argtypes = [C.c_void_p, f32p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_char_p]
        self.ctx = None
        self.cfg = None

    def threads_used(self, threads=0):
        return self.lib.ref_threads_used(threads)

    def transcribe_text(self, samples, segment_sec=0.0, search_sec=3.0, stream=False, language=None):
        """qwen_transcribe_audio (-S segment_sec -W search_sec) or qwen_transcribe_stream, unmodified; returns the text."""
        samples = np.ascontiguousarray(samples, np.float32)
        p = self.lib.ref_transcribe_text(self.ctx, samples, len(samples), float(segment_sec), float(search_sec),
                                         1 if stream else 0, language.encode() if language else None)
        if not p:
            return None
        txt = C.string_at(p).decode("utf-8", errors="replace")
        self.lib.ref_free_buf(p)
        return txt

    def load(self, model_dir, threads=0):
        self.ctx = self.lib.ref_load(model_dir.encode(), threads)
        if not self.ctx:
            raise RuntimeError(f"reference failed to load {model_dir}")
        cfg = np.zeros(12, np.int32)
        self.lib.ref_config(self.ctx, cfg)
        keys = ["enc_d_model", "enc_layers", "enc_heads", "enc_ffn_dim", "enc_output_dim", "dec_hidden",
                "dec_layers", "dec_heads", "dec_kv_heads", "dec_head_dim", "dec_intermediate", "vocab_size"]
        self.cfg = dict(zip(keys, (int(v) for v in cfg)))
        return self

    def close(self):
        if self.ctx:
            self.lib.ref_free(self.ctx)
            self.ctx = None

    def _take(self, ptr, n):
        out = np.ctypeslib.as_array(ptr, shape=(n,)).copy()
        self.lib.ref_free_buf(ptr)
        return out

    def parse_wav(self, data):
        """WAV bytes -> f32 mono 16 kHz through the reference's qwen_parse_wav_buffer (None if unsupported)."""
        self.lib.ref_parse_wav.restype = C.POINTER(C.c_float)
        self.lib.ref_parse_wav.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int)]
        n = C.c_int(0)
        p = self.lib.ref_parse_wav(bytes(data), len(data), C.byref(n))
        return self._take(p, n.value) if p else None

    def mel(self, samples):
        samples = np.ascontiguousarray(samples, np.float32)
        fr = C.c_int(0)
        p = self.lib.ref_mel(samples, len(samples), C.byref(fr))
        if not p:
            return None
        return self._take(p, 128 * fr.value).reshape(128, fr.value)

    def encode(self, mel):
        mel = np.ascontiguousarray(mel, np.float32)
        T = C.c_int(0)
        p = self.lib.ref_encode(self.ctx, mel, mel.shape[1], C.byref(T))
        if not p:
            return None
        H = self.cfg["enc_output_dim"]
        return self._take(p, T.value * H).reshape(T.value, H)

    @property
    def kv_len(self):
        return self.lib.ref_get_kv_len(self.ctx)

    @kv_len.setter
    def kv_len(self, n):
        self.lib.ref_set_kv_len(self.ctx, int(n))

    def prefill(self, embeds):
        embeds = np.ascontiguousarray(embeds, np.float32)
        self.lib.ref_prefill(self.ctx, embeds, embeds.shape[0])

    def step(self, embed):
        return self.lib.ref_step(self.ctx, np.ascontiguousarray(embed, np.float32))

    def step_logits(self, embed):
        out = np.empty(self.cfg["vocab_size"], np.float32)
        self.lib.ref_step_logits(self.ctx, np.ascontiguousarray(embed, np.float32), out)
        return out

    def embed(self, tok):
        out = np.empty(self.cfg["dec_hidden"], np.float32)
        self.lib.ref_embed_token(self.ctx, int(tok), out)
        return out

    def read_kv(self, layer, length):
        kvd = self.cfg["dec_kv_heads"] * self.cfg["dec_head_dim"]
        k = np.empty((length, kvd), np.float32)
        v = np.empty((length, kvd), np.float32)
        self.lib.ref_read_kv(self.ctx, layer, length, k, v)
        return k, v

    def transcribe_ids(self, samples, max_new):
        samples = np.ascontiguousarray(samples, np.float32)
        ids = np.zeros(max_new, np.int32)
        tm = np.zeros(4, np.float64)
        T = C.c_int(0)
        n = self.lib.ref_transcribe_ids(self.ctx, samples, len(samples), max_new, ids, tm, C.byref(T))
        if n < 0:
            raise RuntimeError("reference transcribe failed")
        return ids[:n].copy(), dict(mel_ms=tm[0], enc_ms=tm[1], prefill_ms=tm[2], decode_ms=tm[3],
                                    enc_tokens=T.value)


class OracleLib:
    """The plain-C restatement oracle/qasr_oracle.c (kind = "port"). Same surface as RefLib."""

    def __init__(self, path=None):
        path = path or os.path.join(HERE, "libqasr_oracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError("oracle/libqasr_oracle.so not built (run `make -C oracle oracle`)")
        self.path = path
        L = self.lib = C.CDLL(path)
        L.qo_load.restype = C.c_void_p
        L.qo_load.argtypes = [C.c_char_p]
        L.qo_free.argtypes = [C.c_void_p]
        L.qo_config.argtypes = [C.c_void_p, i32p]
        L.qo_mel_spectrogram.restype = C.POINTER(C.c_float)
        L.qo_mel_spectrogram.argtypes = [f32p, C.c_int, C.POINTER(C.c_int)]
        L.qo_encoder_forward.restype = C.POINTER(C.c_float)
        L.qo_encoder_forward.argtypes = [C.c_void_p, f32p, C.c_int, C.POINTER(C.c_int)]
        L.qo_free_buf.argtypes = [C.c_void_p]
        L.qo_set_kv_len.argtypes = [C.c_void_p, C.c_int]
        L.qo_get_kv_len.restype = C.c_int
        L.qo_get_kv_len.argtypes = [C.c_void_p]
        L.qo_decoder_prefill.argtypes = [C.c_void_p, f32p, C.c_int]
        L.qo_decoder_forward.restype = C.c_int
        L.qo_decoder_forward.argtypes = [C.c_void_p, f32p]
        L.qo_decoder_forward_logits.argtypes = [C.c_void_p, f32p, f32p]
        L.qo_embed_token.argtypes = [C.c_void_p, C.c_int, f32p]
        L.qo_read_kv.argtypes = [C.c_void_p, C.c_int, C.c_int, f32p, f32p]
        L.qo_transcribe_ids.restype = C.c_int
        L.qo_transcribe_ids.argtypes = [C.c_void_p, f32p, C.c_int, C.c_int, i32p, f64p, C.POINTER(C.c_int)]
        self.ctx = None
        self.cfg = None

    def threads_used(self, threads=0):
        return os.cpu_count() or 1

    def load(self, model_dir, threads=0):
        self.ctx = self.lib.qo_load(model_dir.encode())
        if not self.ctx:
            raise RuntimeError(f"oracle failed to load {model_dir}")
        cfg = np.zeros(12, np.int32)
        self.lib.qo_config(self.ctx, cfg)
        keys = ["enc_d_model", "enc_layers", "enc_heads", "enc_ffn_dim", "enc_output_dim", "dec_hidden",
                "dec_layers", "dec_heads", "dec_kv_heads", "dec_head_dim", "dec_intermediate", "vocab_size"]
        self.cfg = dict(zip(keys, (int(v) for v in cfg)))
        return self

    def close(self):
        if self.ctx:
            self.lib.qo_free(self.ctx)
            self.ctx = None

    def _take(self, ptr, n):
        out = np.ctypeslib.as_array(ptr, shape=(n,)).copy()
        self.lib.qo_free_buf(ptr)
        return out

    def parse_wav(self, data):
        """WAV bytes -> f32 mono 16 kHz (restatement of qwen_parse_wav_buffer; None if unsupported)."""
        self.lib.qo_parse_wav_buffer.restype = C.POINTER(C.c_float)
        self.lib.qo_parse_wav_buffer.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int)]
        n = C.c_int(0)
        p = self.lib.qo_parse_wav_buffer(bytes(data), len(data), C.byref(n))
        return self._take(p, n.value) if p else None

    def mel(self, samples):
        samples = np.ascontiguousarray(samples, np.float32)
        fr = C.c_int(0)
        p = self.lib.qo_mel_spectrogram(samples, len(samples), C.byref(fr))
        if not p:
            return None
        return self._take(p, 128 * fr.value).reshape(128, fr.value)

    def encode(self, mel):
        mel = np.ascontiguousarray(mel, np.float32)
        T = C.c_int(0)
        p = self.lib.qo_encoder_forward(self.ctx, mel, mel.shape[1], C.byref(T))
        if not p:
            return None
        H = self.cfg["enc_output_dim"]
        return self._take(p, T.value * H).reshape(T.value, H)

    @property
    def kv_len(self):
        return self.lib.qo_get_kv_len(self.ctx)

    @kv_len.setter
    def kv_len(self, n):
        self.lib.qo_set_kv_len(self.ctx, int(n))

    def prefill(self, embeds):
        embeds = np.ascontiguousarray(embeds, np.float32)
        self.lib.qo_decoder_prefill(self.ctx, embeds, embeds.shape[0])

    def step(self, embed):
        return self.lib.qo_decoder_forward(self.ctx, np.ascontiguousarray(embed, np.float32))

    def step_logits(self, embed):
        out = np.empty(self.cfg["vocab_size"], np.float32)
        self.lib.qo_decoder_forward_logits(self.ctx, np.ascontiguousarray(embed, np.float32), out)
        return out

    def embed(self, tok):
        out = np.empty(self.cfg["dec_hidden"], np.float32)
        self.lib.qo_embed_token(self.ctx, int(tok), out)
        return out

    def read_kv(self, layer, length):
        kvd = self.cfg["dec_kv_heads"] * self.cfg["dec_head_dim"]
        k = np.empty((length, kvd), np.float32)
        v = np.empty((length, kvd), np.float32)
        self.lib.qo_read_kv(self.ctx, layer, length, k, v)
        return k, v

    def transcribe_ids(self, samples, max_new):
        samples = np.ascontiguousarray(samples, np.float32)
        ids = np.zeros(max_new, np.int32)
        tm = np.zeros(4, np.float64)
        T = C.c_int(0)
        n = self.lib.qo_transcribe_ids(self.ctx, samples, len(samples), max_new, ids, tm, C.byref(T))
        if n < 0:
            raise RuntimeError("oracle transcribe failed")
        return ids[:n].copy(), dict(mel_ms=tm[0], enc_ms=tm[1], prefill_ms=tm[2], decode_ms=tm[3],
                                    enc_tokens=T.value)

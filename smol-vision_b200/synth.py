"""Synthetic inputs for parity tests and benchmarks (no real checkpoints or datasets offline).

* `synth_audio`      - deterministic speech-like 16 kHz audio (band-limited bursts separated by
                       near-silence, s16-quantised), generated with an explicit integer hash so
                       it is bit-identical on every machine / numpy version.
* `ensure_model_dir` - runs tools/build/synth_weights to write a random-init safetensors
                       checkpoint of the named architecture (0.6b / 1.7b) under a cache dir.
"""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SAMPLE_RATE = 16000


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def _uniform(n, seed, stream):
    """n floats in [0,1) from a counter-based hash."""
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) + np.uint64((seed * 1000003 + stream) * 7919 + 1) * np.uint64(0x100000001B3)
        h = _splitmix64(idx)
    return (h >> np.uint64(40)).astype(np.float64) / float(1 << 24)


def synth_audio(seconds, seed=0):
    """Speech-like synthetic audio: float32 mono 16 kHz in [-1, 1), s16-quantised."""
    n = int(round(seconds * SAMPLE_RATE))
    if n <= 0:
        return np.zeros(0, np.float32)
    noise = _uniform(n, seed, 0) * 2.0 - 1.0
    # band-limit: difference of two one-pole smoothers implemented as cumulative moving averages
    def smooth(x, w):
        c = np.cumsum(np.concatenate([[0.0], x]))
        out = (c[w:] - c[:-w]) / w
        return np.concatenate([np.full(w - 1, out[0]), out])
    band = smooth(noise, 4) - smooth(noise, 64)
    t = np.arange(n) / SAMPLE_RATE
    f0 = 110.0 + 40.0 * np.sin(2 * np.pi * 0.37 * t + seed)
    voiced = 0.5 * np.sin(2 * np.pi * np.cumsum(f0) / SAMPLE_RATE) + 0.25 * np.sin(2 * np.pi * 2.0 * np.cumsum(f0) / SAMPLE_RATE)
    # burst envelope: 0.5-4 s bursts at amplitude 0.1 separated by 0.2-0.6 s near-silence (1e-3)
    env = np.full(n, 1e-3)
    u = _uniform(4096, seed, 1)
    pos, k = 0, 0
    while pos < n:
        burst = int((0.5 + 3.5 * u[k % 4096]) * SAMPLE_RATE)
        gap = int((0.2 + 0.4 * u[(k + 1) % 4096]) * SAMPLE_RATE)
        end = min(n, pos + burst)
        ramp = min(400, max(1, (end - pos) // 2))
        seg = np.full(end - pos, 0.1)
        seg[:ramp] *= np.linspace(0.01, 1.0, ramp)
        seg[-ramp:] *= np.linspace(1.0, 0.01, ramp)
        env[pos:end] = np.maximum(env[pos:end], seg)
        pos = end + gap
        k += 2
    x = env * (3.0 * band + 0.6 * voiced)
    q = np.clip(np.round(x * 32768.0), -32768, 32767)
    return (q / 32768.0).astype(np.float32)


def cache_root():
    return os.environ.get("QASR_CACHE_DIR", "/tmp/qasr_cache")


def synth_tool():
    return os.path.join(ROOT, "tools", "build", "synth_weights")


def ensure_model_dir(variant="0.6b", seed=1234):
    """Directory with model.safetensors + vocab.json for `variant`, generated on first use."""
    if variant not in ("0.6b", "1.7b"):
        raise ValueError(variant)
    d = os.path.join(cache_root(), f"synth_{variant}_{seed}")
    expect = {"0.6b": 1_500_000_000, "1.7b": 4_000_000_000}[variant]
    st = os.path.join(d, "model.safetensors")
    if os.path.exists(st) and os.path.getsize(st) > expect and os.path.exists(os.path.join(d, ".done")):
        return d
    tool = synth_tool()
    if not os.path.exists(tool):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tools")], check=True, capture_output=True)
    os.makedirs(d, exist_ok=True)
    subprocess.run([tool, variant, d, str(seed)], check=True, capture_output=True)
    with open(os.path.join(d, ".done"), "w") as f:
        f.write("ok\n")
    return d

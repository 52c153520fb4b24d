"""Debug: per-unit wait/compute trace of warp 0 of one CTA of the streaming decode kernel."""
import ctypes as C, os, sys
import numpy as np
os.environ["QASR_MEGA_PROF"] = "1"
os.environ["QASR_MEGA_DEBUG"] = str(64 | (int(sys.argv[3]) if len(sys.argv) > 3 else 0))
os.environ["QASR_MEGA_TRACE_CTA"] = sys.argv[1] if len(sys.argv) > 1 else "0"
variant = sys.argv[2] if len(sys.argv) > 2 else "1.7b"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
eng = pkg.QasrCuda(0).load(pkg.ensure_model_dir(variant))
audio = pkg.synth_audio(3.64, 100)
ids, info = eng.transcribe_ids(audio, 2)
eng.kv_len = info["enc_tokens"] + 15
out = eng.generate(int(ids[0]), 3)   # one launch, 2 steps
buf = np.zeros(3 * 4096, np.int64)
eng.lib.qasr_debug_mega_prof.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(dtype=np.int64)]
assert eng.lib.qasr_debug_mega_prof(eng.ctx, buf) == 0
t = buf[2 * 4096:2 * 4096 + 3 * 1300].reshape(-1, 3).astype(np.float64)
n = int((t[:, 0] != 0).sum())
t0 = t[0, 0]
print(f"CTA {os.environ['QASR_MEGA_TRACE_CTA']} warp 0, {n} units traced (us since first unit: t, wait for data, compute)")
for i in range(min(n, 90)):
    b, a, c = t[i]
    gap = (b - t[i - 1, 2]) / 1965 if i else 0.0
    print(f"{i:3d}  t={((b - t0) / 1965):8.2f}  gap_before={gap:6.2f}  wait={(a - b) / 1965:6.2f}  compute={(c - a) / 1965:5.2f}")

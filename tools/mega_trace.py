"""Debug: per-unit wait/compute trace of warp 0 / CTA 0 of the decode megakernel."""
import ctypes as C, os, sys
import numpy as np
os.environ["QASR_MEGA_PROF"] = "1"
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
os.environ["QASR_MEGA_DEBUG"] = str(64 | mode)
os.environ["QASR_MEGA_TRACE_CTA"] = sys.argv[2] if len(sys.argv) > 2 else "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
eng = pkg.QasrCuda(0).load(pkg.ensure_model_dir("1.7b"))
audio = pkg.synth_audio(3.64, 100)
ids, info = eng.transcribe_ids(audio, 2)
eng.kv_len = info["enc_tokens"] + 15
out = eng.generate(int(ids[0]), 3)   # one launch, 2 steps
buf = np.zeros(3 * 4096, np.int64)
eng.lib.qasr_debug_mega_prof.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(dtype=np.int64)]
assert eng.lib.qasr_debug_mega_prof(eng.ctx, buf) == 0
t = buf[2 * 4096:2 * 4096 + 4095].reshape(-1, 3)[:420].astype(np.float64)
t0 = t[0, 0]
# units per layer for warp 0: QKV 3, WO 2, GU 7, DOWN 4 = 16
names = ["?"] * 16
print("mode", mode, "first 3 layers (us since start: before-wait, wait, compute)")
for i in range(40):
    b, a, c = t[i]
    print(f"{i:3d} {names[i % 16]}  t={((b - t0) / 1965):8.2f}  wait={(a - b) / 1965:6.2f}  compute={(c - a) / 1965:5.2f}")

"""Debug: clock64 stamps of decode_regs_kernel with QASR_MEGA_DEBUG=128 (three extra stamps inside every layer phase that has
rows on the CTA: MMAs done | after bar.sync | after the epilogue).  Prints the mean interval between consecutive stamps of a layer."""
import ctypes as C, os, sys
import numpy as np
os.environ["QASR_MEGA_PROF"] = "1"
os.environ["QASR_MEGA_DEBUG"] = os.environ.get("QASR_MEGA_DEBUG", "128")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
if os.environ.get("QASR_LIB"):  # a -DSR_FINE_PROF build next to the default library
    sys.modules[pkg.__name__ + ".binding"].LIB_PATH = os.path.abspath(os.environ["QASR_LIB"])
variant = sys.argv[1] if len(sys.argv) > 1 else "0.6b"
eng = pkg.QasrCuda(0).load(pkg.ensure_model_dir(variant))
audio = pkg.synth_audio(3.64, 100)
ids, info = eng.transcribe_ids(audio, 4)
eng.kv_len = info["enc_tokens"] + 15 + 4 - 1
out = eng.generate(int(ids[-1]), 3)      # one launch, 2 steps
buf = np.zeros(3 * 4096, np.int64)
print("ids", out[:3])
eng.lib.qasr_debug_mega_prof.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(dtype=np.int64)]
assert eng.lib.qasr_debug_mega_prof(eng.ctx, buf) == 0
L = 28
for which, name in ((0, "CTA0"), (1, "CTAlast")):
    t = buf[which * 4096:(which + 1) * 4096]
    n = int(np.argmax(t == 0)) if (t == 0).any() else len(t)
    t = t[:n].astype(np.float64) / 1.965e3
    for M in range(10, 80):                  # find the per-layer stamp count: n = steps * (M * L + 1)
        if n % (M * L + 1) == 0:
            break
    else:
        print(name, "cannot factor", n); continue
    per_step = M * L + 1
    steps = n // per_step
    tt = t[(steps - 1) * per_step:]
    lay = tt[:M * L].reshape(L, M)
    d = np.diff(np.concatenate([lay, np.append(lay[1:, :1], [[tt[M * L]]], axis=0)], axis=1), axis=1)
    acc = d[1:L - 1].mean(axis=0)
    print(name, "stamps/layer", M, "steps", steps, " ".join(f"{v:.2f}" for v in acc), f"| layer {acc.sum():.2f} us")

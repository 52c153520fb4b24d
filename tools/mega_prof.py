"""Debug: per-phase clock64 breakdown of the decode megakernel (QASR_MEGA_PROF=1)."""
import ctypes as C, os, sys
import numpy as np
os.environ["QASR_MEGA_PROF"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
variant = sys.argv[1] if len(sys.argv) > 1 else "1.7b"
eng = pkg.QasrCuda(0).load(pkg.ensure_model_dir(variant))
audio = pkg.synth_audio(3.64, 100)
ids, info = eng.transcribe_ids(audio, 8)
eng.kv_len = info["enc_tokens"] + 15 + 8 - 1
tok = eng.step_token(int(ids[-1]))      # one single-step launch -> stamps of exactly one token
buf = np.zeros(3 * 4096, np.int64)
eng.lib.qasr_debug_mega_prof.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(dtype=np.int64)]
assert eng.lib.qasr_debug_mega_prof(eng.ctx, buf) == 0
for which, name in ((0, "CTA0"), (1, "CTAlast")):
    t = buf[which * 4096:(which + 1) * 4096]
    n = int((t != 0).sum())
    t = t[:n].astype(np.float64)
    # per layer: QKV 5 marks (start, staged, units, epi, barrier), ATTN 2 marks (pre-barrier, post), WO 4, GU 4 (no start mark), DOWN 4
    d = np.diff(t) / 1.965e3  # us at 1965 MHz
    per_layer = 5 + 2 + 4 + 4 + 4
    L = (n - 1) // per_layer
    names = ["qkv.stage", "qkv.units", "qkv.epi", "qkv.barrier", "attn.work", "attn.barrier", "wo.stage", "wo.units", "wo.epi", "wo.barrier",
             "gu.stage", "gu.units", "gu.epi", "gu.barrier", "down.stage", "down.units", "down.epi", "down.barrier", "->next"]
    acc = np.zeros(per_layer)
    for l in range(1, L - 1):
        acc += d[l * per_layer:(l + 1) * per_layer]
    acc /= max(L - 2, 1)
    print(name, "marks", n, "layers", L, "total us", (t[-1] - t[0]) / 1.965e3)
    for nm, v in zip(names, acc):
        print(f"   {nm:14s} {v:7.2f} us")
    print("   per-layer sum", acc.sum())

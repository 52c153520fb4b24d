"""Debug: per-phase clock64 breakdown of the decode kernel in batched mode (QASR_MEGA_PROF=1): python tools/mega_prof_batch.py [variant] [seconds]"""
import ctypes as C, os, sys
import numpy as np
os.environ["QASR_MEGA_PROF"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
variant = sys.argv[1] if len(sys.argv) > 1 else "0.6b"
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
eng = pkg.QasrCuda(0).load(pkg.ensure_model_dir(variant))
B = eng.max_batch
units = [pkg.synth_audio(secs, 200 + i) for i in range(B)]
ids, tm = eng.transcribe_batch(units, 8)     # launches: 1 step, then 7 steps -> stamps of the 7-step launch remain
buf = np.zeros(3 * 4096, np.int64)
eng.lib.qasr_debug_mega_prof.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(dtype=np.int64)]
assert eng.lib.qasr_debug_mega_prof(eng.ctx, buf) == 0
names10 = ["qkv.stage(wait xdn)", "qkv.units", "attn", "wo.stage(wait att)", "wo.units", "gu.stage(wait xwo)", "gu.units",
           "down.stage(wait act)", "down.units", "->next layer"]
names13 = names10[:2] + ["attn.wait qkv", "attn.norm+rope", "attn.keys", "attn.merge+store"] + names10[3:]
L = 28
print(f"{variant}: {B} sequences per step, {secs} s units, decode {tm['decode_ms'] / 8:.3f} ms per step")
for which, name in ((0, "CTA0"), (1, "CTAlast")):
    M = 13 if which == 0 else 10
    names = names13 if which == 0 else names10
    t = buf[which * 4096:(which + 1) * 4096]
    n = int((t != 0).sum())
    t = t[:n].astype(np.float64) / 1.965e3
    per_step = M * L + 1
    tt = t[per_step:2 * per_step + 1]          # second step of the launch
    lay = tt[:M * L].reshape(L, M)
    d = np.diff(np.concatenate([lay, np.append(lay[1:, :1], [[tt[M * L]]], axis=0)], axis=1), axis=1)
    acc = d[1:L - 1].mean(axis=0)
    print(name, f"step total {tt[per_step] - tt[0]:.1f} us; head {tt[M * L] - tt[M * L - 1]:.1f} us")
    for nm, v in zip(names, acc):
        print(f"   {nm:22s} {v:7.2f} us")
    print(f"   per-layer sum {acc.sum():.2f} us")

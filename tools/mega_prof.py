"""Debug: per-phase clock64 breakdown of the streaming decode kernel (QASR_MEGA_PROF=1).
Marks per layer (CTA 0 / last CTA, thread 0): layer top | QKV staged | QKV done | ATTN done | WO staged |
WO done | GU staged | GU done | DOWN staged | DOWN done; one more per step after the head."""
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
out = eng.generate(int(ids[-1]), 4)      # one launch, 3 steps -> stamps of 3 tokens
buf = np.zeros(3 * 4096, np.int64)
eng.lib.qasr_debug_mega_prof.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(dtype=np.int64)]
assert eng.lib.qasr_debug_mega_prof(eng.ctx, buf) == 0
names10 = ["qkv.stage(wait xdn)", "qkv.units", "attn", "wo.stage(wait att)", "wo.units", "gu.stage(wait xwo)", "gu.units",
           "down.stage(wait act)", "down.units", "->next layer"]
names13 = names10[:2] + ["attn.wait qkv", "attn.norm+rope", "attn.keys", "attn.merge+store"] + names10[3:]
L = 28
for which, name in ((0, "CTA0"), (1, "CTAlast")):
    t = buf[which * 4096:(which + 1) * 4096]
    n = int((t != 0).sum())
    t = t[:n].astype(np.float64) / 1.965e3  # us at 1965 MHz
    M = 13 if which == 0 else 10          # CTA 0 is always an attention CTA (4 sub-marks), the last CTA never
    names = names13 if which == 0 else names10
    per_step = M * L + 1
    steps = n // per_step
    s = 1 if steps > 1 else 0             # second token: steady state
    tt = t[s * per_step:(s + 1) * per_step + 1]
    lay = tt[:M * L].reshape(L, M)
    d = np.diff(np.concatenate([lay, np.append(lay[1:, :1], [[tt[M * L]]], axis=0)], axis=1), axis=1)  # [L,M]
    acc = d[1:L - 1].mean(axis=0)
    print(name, "marks", n, "steps", steps, f"step total {tt[per_step] - tt[0]:.1f} us; head {tt[M * L] - tt[M * L - 1]:.1f} us" if len(tt) > per_step else "")
    for nm, v in zip(names, acc):
        print(f"   {nm:22s} {v:7.2f} us")
    print(f"   per-layer sum {acc.sum():.2f} us")

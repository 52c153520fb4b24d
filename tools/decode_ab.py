"""A/B timing of the decode megakernel under QASR_MEGA_DEBUG variants (one model load)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
if os.environ.get("QASR_LIB"):  # A/B of library variants built next to the default one
    sys.modules[pkg.__name__ + ".binding"].LIB_PATH = os.path.abspath(os.environ["QASR_LIB"])
variant = sys.argv[1] if len(sys.argv) > 1 else "1.7b"
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 2, 4, 6, 16, 1]
eng = pkg.QasrCuda(0).load(pkg.ensure_model_dir(variant))
audio = pkg.synth_audio(3.64, 100)
ids, info = eng.transcribe_ids(audio, 2)
kv0 = info["enc_tokens"] + 15
for dbg in variants:
    os.environ["QASR_MEGA_DEBUG"] = str(dbg)
    best = 1e9
    out = None
    for rep in range(4):
        eng.kv_len = kv0
        eng.decode_stats(reset=True)
        out = eng.generate(int(ids[0]), 33)
        steps, ms = eng.decode_stats(reset=True)
        best = min(best, ms / max(steps, 1))
    print(f"debug={dbg:3d}  {best*1000:8.1f} us/token   ids[:4]={out[:4].tolist()}")

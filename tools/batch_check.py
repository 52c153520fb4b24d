"""Debug: batched-decode parity of an alternative build of the library (python tools/batch_check.py <lib.so> [variant])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
lib = pkg.load_library(os.path.abspath(sys.argv[1])) if len(sys.argv) > 1 and sys.argv[1] != "-" else None
variant = sys.argv[2] if len(sys.argv) > 2 else "0.6b"
eng = pkg.QasrCuda(0, lib=lib).load(pkg.ensure_model_dir(variant))
secs = [1.3, 2.6, 0.9, 3.4, 1.9, 2.2, 1.1]
caps = [9, 12, 6, 17, 8, 11, 5]
units = [pkg.synth_audio(s, seed=40 + i) for i, s in enumerate(secs)]
single = [eng.transcribe_ids(u, c)[0].tolist() for u, c in zip(units, caps)]
try:
    batched, tm = eng.transcribe_batch(units, caps)
    print(sys.argv[1:], "batch == single:", [b.tolist() for b in batched] == single)
except Exception as ex:
    print(sys.argv[1:], "FAILED:", str(ex)[:160])

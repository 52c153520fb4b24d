"""Time qasr_cuda_load_dir (checkpoint -> HBM, incl. the decode weight image): python tools/load_time.py [variant]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
variant = sys.argv[1] if len(sys.argv) > 1 else "1.7b"
d = pkg.ensure_model_dir(variant)
size = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d) if f.endswith(".safetensors"))
for rep in range(3):
    eng = pkg.QasrCuda(0)
    t0 = time.perf_counter()
    eng.load(d)
    dt = time.perf_counter() - t0
    print(f"{variant}: load {dt:.2f} s for {size / 1e9:.2f} GB of safetensors = {size / dt / 1e9:.1f} GB/s (page cache {'cold?' if rep == 0 else 'warm'})")
    eng.close()

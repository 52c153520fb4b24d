"""Small profiling target: N utterance passes of the bench workload (for ncu launch lists)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
variant = sys.argv[1] if len(sys.argv) > 1 else "1.7b"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 3.64175
max_new = int(sys.argv[4]) if len(sys.argv) > 4 else 32
eng = pkg.QasrCuda(0).load(pkg.ensure_model_dir(variant))
audio = pkg.synth_audio(secs, 100)[: int(secs * 16000)]
for i in range(n):
    ids, info = eng.transcribe_ids(audio, max_new)
    print(i, info, eng.launch_count)

"""Stage timings of the batched path: python tools/batch_profile.py [variant] [units] [seconds] [tokens] [repeat]
(run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "1.7b"
n_units = int(sys.argv[2]) if len(sys.argv) > 2 else 64
seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 30.0
tokens = int(sys.argv[4]) if len(sys.argv) > 4 else 128
repeat = int(sys.argv[5]) if len(sys.argv) > 5 else 2
pkg = ge.load_package()
eng = pkg.QasrCuda(0).load(pkg.ensure_model_dir(variant))
units = [pkg.synth_audio(seconds, seed=i)[: int(seconds * 16000)] for i in range(n_units)]
for r in range(repeat):
    t0 = time.perf_counter()
    ids, tm = eng.transcribe_batch(units, tokens)
    wall = (time.perf_counter() - t0) * 1e3
    steps, dec_ms = eng.decode_stats(reset=True)
    print(f"pass {r}: wall {wall:.1f} ms  " + "  ".join(f"{k} {v:.1f}" for k, v in tm.items()) +
          f"  | per unit {wall / n_units:.2f} ms  decode {dec_ms / max(steps, 1):.3f} ms/step x {steps} steps"
          f"  = {n_units * steps / max(dec_ms, 1e-9) * 1e3:.0f} tok/s  realtime x{seconds * n_units / wall * 1e3:.0f}", flush=True)
eng.close()

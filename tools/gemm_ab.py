"""Print a digest of the skinny split-K GEMM outputs at production shapes (debug hook).
Run once per setting of QASR_GEMM_SK_CLUSTER / QASR_PDL and diff the lines: the reduction order is fixed, so the
digests must be identical."""
import hashlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
eng = pkg.QasrCuda(0)
for M, K, N in [(61, 2048, 4096), (61, 6144, 2048), (47, 4096, 1024), (47, 1024, 1024), (143, 3584, 896), (208, 4320, 480), (13, 7680, 896), (130, 72, 40), (255, 2048, 2048)]:
    rng = np.random.default_rng(M * 31 + K + N)
    x = rng.standard_normal((M, K), dtype=np.float32)
    W = (rng.standard_normal((N, K), dtype=np.float32) * K ** -0.5).view(np.uint32) >> 16
    y = eng.linear_bf16(x, W.astype(np.uint16))
    print(M, K, N, hashlib.sha1(y.tobytes()).hexdigest()[:16], float(np.abs(y).mean()))

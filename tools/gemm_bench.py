"""Device-time the tensor-core GEMM at the encoder / prefill shapes (debug hook)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
eng = pkg.QasrCuda(0)
if os.environ.get("QASR_SPLIT"):  # 1 = plain bf16 activations (one plane): timing experiments
    eng.lib.qasr_cuda_set_gemm_split.argtypes = [C.c_void_p, C.c_int]
    print("set_gemm_split rc", eng.lib.qasr_cuda_set_gemm_split(eng.ctx, int(os.environ["QASR_SPLIT"])))
f = eng.lib.qasr_debug_gemm_bench
f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
shapes = [("pre.qkv", 61, 2048, 4096), ("pre.wo", 61, 2048, 2048), ("pre.gu", 61, 2048, 12288), ("pre.down", 61, 6144, 2048),
          ("enc.qkv", 47, 1024, 3072), ("enc.wo", 47, 1024, 1024), ("enc.fc1", 47, 1024, 4096), ("enc.fc2", 47, 4096, 1024),
          ("conv2", 1456, 4320, 480), ("conv3", 384, 4320, 480), ("convout", 47, 7680, 1024),
          ("jfk.qkv", 157, 1024, 4096), ("30s.gu", 404, 2048, 12288), ("30s.down", 404, 6144, 2048),
          ("b4.enc.fc1", 1560, 1024, 4096), ("b4.pre.gu", 1616, 2048, 12288), ("b16.enc.fc1", 6240, 1024, 4096), ("big", 8192, 4096, 8192)]
if len(sys.argv) > 1 and sys.argv[1] == "medium":
    shapes = [("06.pre.qkv", 274, 1024, 4096), ("06.pre.wo", 274, 2048, 1024), ("06.pre.gu", 274, 1024, 6144), ("06.pre.down", 274, 3072, 1024),
              ("06.enc.qkv", 260, 896, 2688), ("06.enc.wo", 260, 896, 896), ("06.enc.fc1", 260, 896, 3584), ("06.enc.fc2", 260, 3584, 896),
              ("17.pre.qkv", 404, 2048, 4096), ("17.pre.wo", 404, 2048, 2048), ("17.pre.gu", 404, 2048, 12288), ("17.pre.down", 404, 6144, 2048),
              ("17.enc.qkv", 390, 1024, 3072), ("17.enc.wo", 390, 1024, 1024), ("17.enc.fc1", 390, 1024, 4096), ("17.enc.fc2", 390, 4096, 1024),
              ("conv2.30s", 48000, 4320, 480), ("conv3.30s", 12000, 4320, 480)]
if len(sys.argv) > 1 and sys.argv[1] == "batch":   # the batched throughput path: 8 / 64 units of 30 s (1.7B), 60 segments of 20 s (0.6B)
    shapes = [("17.pre.qkv", 25856, 2048, 4096), ("17.pre.wo", 25856, 2048, 2048), ("17.pre.gu", 25856, 2048, 12288), ("17.pre.down", 25856, 6144, 2048),
              ("17.enc.qkv", 3120, 1024, 3072), ("17.enc.wo", 3120, 1024, 1024), ("17.enc.fc1", 3120, 1024, 4096), ("17.enc.fc2", 3120, 4096, 1024),
              ("conv2.8x30", 192000, 4320, 480), ("conv3.8x30", 48000, 4320, 480), ("convout", 3120, 7680, 1024),
              ("06.pre.gu", 16440, 1024, 6144), ("06.pre.down", 16440, 3072, 1024), ("06.enc.fc1", 3120, 896, 3584),
              ("dec.qkv.b64", 64, 2048, 4096), ("dec.wo.b64", 64, 2048, 2048), ("dec.gu.b64", 64, 2048, 12288), ("dec.down.b64", 64, 6144, 2048), ("dec.head.b64", 64, 2048, 151936),
              ("dec06.gu.b60", 60, 1024, 6144), ("dec06.head.b60", 60, 1024, 151936), ("big", 8192, 4096, 8192)]
def pipeline_mode(name):   # the epilogue each shape has in the pipeline: 0 f32 store, 1 residual add, 2 GELU -> hi/lo planes, 3 SwiGLU -> hi/lo planes
    if ".gu" in name: return 3
    if "fc1" in name or name.startswith("conv2") or name.startswith("conv3"): return 2
    if ".wo" in name or ".down" in name or "fc2" in name: return 1
    return 0
if len(sys.argv) > 1 and sys.argv[1] == "decode":   # one decode step of the batched path at 128 sequences (1.7B and 0.6B)
    shapes = [("dec.qkv", 128, 2048, 4096), ("dec.wo", 128, 2048, 2048), ("dec.gu", 128, 2048, 12288), ("dec.down", 128, 6144, 2048), ("dec.head", 128, 2048, 151936),
              ("dec06.qkv", 128, 1024, 4096), ("dec06.wo", 128, 2048, 1024), ("dec06.gu", 128, 1024, 6144), ("dec06.down", 128, 3072, 1024), ("dec06.head", 128, 1024, 151936)]
if len(sys.argv) > 1 and sys.argv[1] == "decode32":  # decode step of a 32-sequence group (an 8-GPU shard of configs[4]) and a 31-row prompt: QASR_GEMM_MP32=0|1
    shapes = [("dec.qkv", 32, 2048, 4096), ("dec.wo", 32, 2048, 2048), ("dec.gu", 32, 2048, 12288), ("dec.down", 32, 6144, 2048), ("dec.head", 32, 2048, 151936),
              ("dec06.qkv", 23, 1024, 4096), ("dec06.wo", 23, 2048, 1024), ("dec06.gu", 23, 1024, 6144), ("dec06.down", 23, 3072, 1024), ("dec06.head", 23, 1024, 151936)]
for name, M, K, N in shapes:
    us = C.c_double(0)
    rc = f(eng.ctx, M, K, N, 64, pipeline_mode(name), C.byref(us))
    mb = 2.0 * N * K / 1e6
    print(f"{name:14s} M={M:6d} K={K:5d} N={N:5d}  {us.value:8.1f} us   weights {mb:6.1f} MB -> {mb / us.value * 1e3 / 1e3:6.2f} TB/s   {2.0*M*N*K/us.value/1e6:7.1f} TFLOP/s algorithmic (x2 issued)" if rc == 0 else f"{name} failed {eng._err()}")

"""CPU: the C-ABI shared library loads and exports every symbol include/qasr_cuda.h declares;
without a GPU the product fails loudly instead of falling back to anything."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "qasr_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(qasr_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(pkg):
    syms = header_symbols()
    assert len(syms) >= 45
    out = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (qasr_[a-z0-9_]+)", out))
    missing = [s for s in syms if s not in exported]
    assert not missing, f"header declares symbols the library does not export: {missing}"


def test_binding_covers_header(pkg):
    assert sorted(pkg.SIGNATURES) == header_symbols()
    pkg.load_library()  # attaches every signature; AttributeError if a symbol is absent


def test_host_only_entry_points(pkg):
    lib = pkg.load_library()
    assert lib.qasr_cuda_mel_frames(176000) == 1100      # jfk.wav, SURVEY 8: 1100 frames
    assert lib.qasr_cuda_mel_frames(58268) == 364        # test_speech.wav
    assert lib.qasr_cuda_encoder_tokens(1100) == 143     # 11 chunks * 13
    assert lib.qasr_cuda_encoder_tokens(364) == 47       # 3*13 + conv^3(64) = 8
    assert lib.qasr_cuda_encoder_tokens(3000) == 390
    assert lib.qasr_cuda_encoder_tokens(1) == 1
    assert lib.qasr_get_num_cpus() >= 1
    lib.qasr_set_threads(4)  # compat no-op


def test_sass_is_blackwell_native(pkg):
    """The tensor-core GEMM must contain tcgen05 / TMA / TMEM instructions (B200_PROFILING.md)."""
    r = subprocess.run(["cuobjdump", "-sass", pkg.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in r.stdout, f"{mnemonic} missing from SASS"
    assert "HGMMA" not in r.stdout


def test_decode_stream_sass_has_no_undefined_descriptor(pkg):
    """Canary for a ptxas 12.9 mis-assembly seen while building qasr_stream.cu: cache-hinted LDGSTS (cp.async) whose
    shared address was split into [R+UR0] got descriptor registers UR0/UR1 that no instruction ever writes
    (CUDA_EXCEPTION_4 on the first copy).  Every uniform register an LDGSTS reads must be written somewhere."""
    r = subprocess.run(["cuobjdump", "-sass", pkg.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    funcs = re.split(r"\n\s*Function : ", r.stdout)
    checked = 0
    for f in funcs:
        if not f.startswith("_Z20decode_stream_kernel"):
            continue
        checked += 1
        used = set(re.findall(r"LDGSTS[^;]*?(UR\d+)", f)) | set(re.findall(r"LDGSTS[^;]*desc\[(UR\d+)\]", f))
        for ur in used:
            n = int(ur[2:])
            writers = [rf"\b{ur}\s*,", rf"\bUR{n - 1}\s*," if n % 2 else r"$^"]  # written directly, or as the high half of a 64-bit pair
            assert any(re.search(rf"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?(?!LDGSTS)[A-Z0-9_.]+\s+{w}", f, re.M) for w in writers), \
                f"{ur} is read by an LDGSTS but never written in {f[:60]}"
    assert checked >= 3


def test_no_cpu_fallback_without_gpu(pkg):
    lib = pkg.load_library()
    if lib.qasr_cuda_device_count() > 0:
        pytest.skip("a GPU is visible here")
    assert not lib.qasr_cuda_init(0)
    assert b"no CUDA device" in lib.qasr_cuda_last_error()
    with pytest.raises(pkg.QasrError):
        pkg.QasrCuda(0)


def test_product_never_touches_oracle():
    """Nothing under smol-vision_b200/ may import, link or execute anything under oracle/."""
    pkg_dir = os.path.join(ROOT, "smol-vision_b200")
    bad = re.compile(r"(import\s+oracle|from\s+oracle|oracle/|libqasr_oracle|qasr_oracle|libqasr_ref|ref_harness|OracleLib|RefLib)")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert not bad.search(txt), f"{f} references the oracle"


def test_hot_kernel_resource_budget(pkg):
    """Static launch budget of the hot kernels, read from the built library: the persistent decode kernel runs 512 threads
    per CTA, so more than 128 registers per thread cannot launch at all (65536 registers per SM) and spills would sit on
    the per-unit critical path; the tcgen05 GEMMs keep one elected thread per role and must not spill either."""
    r = subprocess.run(["cuobjdump", "--dump-resource-usage", pkg.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    usage = {}
    name = None
    for line in r.stdout.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and name:
            usage[name] = (int(m.group(1)), int(m.group(2)))
            name = None
    decode = {k: v for k, v in usage.items() if "decode_stream_kernel" in k}
    gemm = {k: v for k, v in usage.items() if "gemm_tc_kernel" in k or "gemm_tc_skinny_kernel" in k}
    assert len(decode) == 3 and len(gemm) == 6, sorted(usage)
    for k, (reg, stack) in decode.items():
        assert reg <= 128 and stack <= 16, (k, reg, stack)
    for k, (reg, stack) in gemm.items():
        assert reg <= 128 and stack == 0, (k, reg, stack)

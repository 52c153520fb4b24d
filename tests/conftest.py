import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as ge  # noqa: E402

PRE = [151644, 8948, 198, 151645, 198, 151644, 872, 198, 151669]  # reference qwen_asr.c:388-393
SUF = [151670, 151645, 198, 151644, 77091, 198]                    # reference qwen_asr.c:394-396


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _built(path):
    return os.path.exists(os.path.join(ROOT, path))


@pytest.fixture(scope="session")
def pkg():
    if not _built("smol-vision_b200/libqasr_cuda.so") or not _built("oracle/libqasr_oracle.so") \
            or not _built("tools/build/synth_weights"):
        ge.build()
    return ge.load_package()


@pytest.fixture(scope="session")
def model06(pkg):
    return pkg.ensure_model_dir("0.6b")


@pytest.fixture(scope="session")
def model17(pkg):
    return pkg.ensure_model_dir("1.7b")


@pytest.fixture(scope="session")
def oracle_lib(pkg):
    from oracle.bindings import OracleLib
    return OracleLib


@pytest.fixture(scope="session")
def oracle06(pkg, model06, oracle_lib):
    o = oracle_lib().load(model06)
    yield o
    o.close()


@pytest.fixture(scope="session")
def ref_lib(pkg):
    """The compiled reference (oracle/_ref), or None when it was not built / shipped."""
    from oracle.bindings import RefLib, ref_lib_path
    return RefLib if ref_lib_path() else None


@pytest.fixture(scope="session")
def golden_ops():
    return np.load(os.path.join(ROOT, "tests", "golden", "ops.npz"))


@pytest.fixture(scope="session")
def golden_seg():
    return np.load(os.path.join(ROOT, "tests", "golden", "segment_0p6b.npz"))


@pytest.fixture(scope="session")
def gpu06(pkg, model06):
    eng = pkg.QasrCuda(0).load(model06)
    yield eng
    eng.close()


def prompt_embeds(engine, enc):
    """[prefix | audio rows | suffix] embeddings exactly as transcribe_segment builds them
    (reference qwen_asr.c:685-759); `engine` is any of RefLib/OracleLib/QasrCuda."""
    H = enc.shape[1]
    rows = [engine.embed(t) for t in PRE] + [enc[i] for i in range(enc.shape[0])] + [engine.embed(t) for t in SUF]
    return np.stack(rows).astype(np.float32).reshape(-1, H)


def rel_err(a, b):
    """max |a-b| relative to max |b| (the tolerance form north_star states)."""
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))

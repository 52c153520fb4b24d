"""smol-vision_b200: B200-native (sm_100a) hot path for the Qwen3-ASR engine of
chitindotsh/smol-vision - log-mel front end, audio encoder, decoder prefill and greedy decode -
behind the C ABI of include/qasr_cuda.h (libqasr_cuda.so, built from csrc/).

The directory name contains a hyphen, so import it through `__graft_entry__.load_package()`
(registers it as module `smol_vision_b200`).
"""
from .binding import LIB_PATH, SIGNATURES, QasrCuda, QasrError, load_library  # noqa: F401
from .synth import SAMPLE_RATE, ensure_model_dir, synth_audio  # noqa: F401
from . import segments  # noqa: F401
from . import streaming  # noqa: F401

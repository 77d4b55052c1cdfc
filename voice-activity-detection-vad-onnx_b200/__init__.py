"""vadx -- B200-native batched voice-activity-detection engine (hot path only).

Raw 16 kHz audio in -> per-frame speech probability -> timestamps, for the model families of
DakeQQ/Voice-Activity-Detection-VAD-ONNX, computed by hand-written sm_100a CUDA kernels behind
the C ABI declared in include/vadx.h (libvadx.so).  There is no CPU fallback: every compute
entry point raises if the CUDA library or a CUDA device is missing.
"""
__version__ = "0.1.0"

from . import constants, lib, tables, weights, synth, distributed  # noqa: F401
from .session import InferenceSession, FireRedSession, FireRedStreamSession, FsmnSession, MarbleNetSession, SileroSession  # noqa: F401
from .postprocess import FramePostConfig, StreamVadPostprocessor, postprocess_frames  # noqa: F401
from .dfsmn_aec import DfsmnAecSession  # noqa: F401,E402

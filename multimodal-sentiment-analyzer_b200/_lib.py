"""ctypes binding of libmsa_b200.so (C ABI: include/msa_b200.h).  Fails loudly when the library is
missing — there is deliberately no fallback implementation."""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmsa_b200.so")

_lock = threading.Lock()
_lib = None

c_void_p, c_int, c_size_t, c_char_p = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_char_p

_SIGNATURES = {
    "msa_version": (c_int, []),
    "msa_strerror": (c_char_p, [c_int]),
    "msa_last_launch_count": (c_int, []),
    "msa_features_cluster_size": (c_int, [c_int]),
    "msa_features_smem_bytes": (c_int, [c_int, c_int]),
    "msa_features_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "msa_features_s16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "msa_features_workspace_bytes": (c_size_t, [c_int, c_int]),
    "msa_features_ws_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "msa_features_ws_s16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "msa_fusion_packed_bytes": (c_size_t, []),
    "msa_fusion_workspace_bytes": (c_size_t, [c_int]),
    "msa_fusion_num_tensors": (c_int, []),
    "msa_fusion_tensor_name": (c_char_p, [c_int]),
    "msa_fusion_tensor_numel": (c_size_t, [c_int]),
    "msa_fusion_pack": (c_int, [ctypes.POINTER(c_void_p), c_void_p, c_void_p]),
    "msa_fusion_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "msa_fusion_set_impl": (c_int, [c_int]),
    "msa_resample_out_len": (c_int, [c_int, c_int, c_int]),
    "msa_resample_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "msa_resample_s16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "msa_resample_kernel_host": (c_int, [c_int, c_int, c_void_p, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "msa_rows_layernorm": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, ctypes.c_float, c_void_p, c_int, c_int, c_void_p]),
    "msa_pitch_frames": (c_int, [c_int]),
    "msa_pitch_outputs": (c_int, [c_int]),
    "msa_voiced_frames": (c_int, [c_int]),
    "msa_pitch_track_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "msa_pitch_track_s16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "msa_spectral_frames": (c_int, [c_int]),
    "msa_spectral_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "msa_spectral_s16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "msa_softmax7": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "msa_pack_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "msa_nan_to_num": (c_int, [c_void_p, ctypes.c_longlong, c_void_p]),
    "msa_aggregate_speakers": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
}

FEAT_STRICT_NAN = 1
PART_WAVE, PART_MFCC, PART_PITCH, PART_ALL = 1, 2, 4, 7
DETAIL_STRIDE = 96


class MsaError(RuntimeError):
    pass


def lib():
    """The loaded library.  Raises if it has not been built (python __graft_entry__.py / build.py)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise MsaError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                   "(nvcc, sm_100a). msa_b200 has no CPU or PyTorch fallback.")
                l = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(l, name)          # AttributeError if the ABI and the header disagree
                    fn.restype = res
                    fn.argtypes = args
                _lib = l
    return _lib


def exported_symbols():
    return list(_SIGNATURES.keys())


def strerror(code: int) -> str:
    return lib().msa_strerror(code).decode()


def check(code: int, what: str):
    if code != 0:
        raise MsaError(f"{what} failed: [{code}] {strerror(code)}")


def ptr(t):
    """Device (or host) pointer of a torch tensor / None."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def current_stream_ptr(device=None):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def normalize_device(device):
    """torch.device with an explicit index: torch.device("cuda") != torch.device("cuda", 0), so every comparison and
    every ``with torch.cuda.device(...)`` in this package goes through the normalised form."""
    import torch
    dev = torch.device(device if device is not None else "cuda")
    if dev.type == "cuda" and dev.index is None and torch.cuda.is_available():
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def require_cuda(device):
    import torch
    if not torch.cuda.is_available():
        raise MsaError("msa_b200 needs a CUDA device (B200, sm_100a); there is no CPU path.")
    dev = normalize_device(device)
    if dev.type != "cuda":
        raise MsaError(f"msa_b200 needs a CUDA device, got {dev}; there is no CPU path.")
    return dev


def on_device(device):
    """Context manager: make `device` the current CUDA device for the C-ABI calls inside (the library launches on the
    current device and builds its constant tables per device)."""
    import torch
    return torch.cuda.device(device)

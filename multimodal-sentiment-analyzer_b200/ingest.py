"""The rows either side of the hot path, on the device.

``resample`` restates ``torchaudio.transforms.Resample(sr, 16000)(waveform)`` of
/root/reference/src/analyzers/audio_analyzer.py:74-77; the three normalizer classes mirror
/root/reference/src/utils/normalization.py:19-98 (same names, ``target_dim``, ``normalize``), and
``assemble_row`` is the explicit concatenation + ``nan_to_num`` of streaming_processor.py:230-300.
All arithmetic runs in hand-written kernels behind the C ABI (csrc/msa_ingest.cu); no torch / torchaudio
compute is used.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _lib


def resample(waveform: torch.Tensor, orig_freq: int, new_freq: int = 16000, device=None) -> torch.Tensor:
    """waveform [..., L] fp32 (or int16 PCM, scaled by 1/32768 like torchaudio.load) -> [..., ceil(new L / orig)] fp32."""
    dev = _lib.require_cuda(device if device is not None else (waveform.device if waveform.is_cuda else "cuda"))
    w = waveform.to(dev)
    if w.dtype != torch.int16:
        w = w.float()
    if int(orig_freq) == int(new_freq):
        return w.float() / 32768.0 if w.dtype == torch.int16 else w
    lead, L = w.shape[:-1], w.shape[-1]
    w2 = w.reshape(-1, L).contiguous()
    lib = _lib.lib()
    n = lib.msa_resample_out_len(L, int(orig_freq), int(new_freq))
    out = torch.empty(w2.shape[0], n, device=dev, dtype=torch.float32)
    fn = lib.msa_resample_s16 if w2.dtype == torch.int16 else lib.msa_resample_f32
    _lib.check(fn(_lib.ptr(w2), w2.shape[0], L, int(orig_freq), int(new_freq), _lib.ptr(out), n, _lib.current_stream_ptr(dev)),
               "msa_resample")
    return out.reshape(lead + (n,))


class FeatureNormalizer:
    """normalization.py:7-17: pad / truncate to ``target_dim`` and apply an (untrained) LayerNorm."""

    def __init__(self, target_dim: int, device="cuda"):
        self.target_dim = target_dim
        self.device = _lib.require_cuda(device)
        self.weight: Optional[torch.Tensor] = None      # LayerNorm gamma / beta: None = the reference's untrained 1 / 0
        self.bias: Optional[torch.Tensor] = None
        self.eps = 1e-5

    def normalize(self, tensor: torch.Tensor, nan_to_num: bool = False) -> torch.Tensor:
        if tensor.dim() == 1:
            tensor = tensor.unsqueeze(0)
        x = tensor.to(self.device).float()
        if x.stride(-1) != 1:
            x = x.contiguous()
        B, d_in = x.shape
        out = torch.empty(B, self.target_dim, device=self.device, dtype=torch.float32)
        rc = _lib.lib().msa_rows_layernorm(_lib.ptr(x), B, d_in, x.stride(0), self.target_dim, _lib.ptr(self.weight), _lib.ptr(self.bias),
                                           self.eps, _lib.ptr(out), self.target_dim, 1 if nan_to_num else 0,
                                           _lib.current_stream_ptr(self.device))
        _lib.check(rc, "msa_rows_layernorm")
        return out


class AudioFeatureNormalizer(FeatureNormalizer):
    """normalization.py:19-44 (8 emotions + pitch + intensity + 13 timbre + speech_rate + 3 rhythm + 4 quality = 31)."""

    def __init__(self, device="cuda"):
        super().__init__(8 + 1 + 1 + 13 + 1 + 3 + 4, device)


class FaceFeatureNormalizer(FeatureNormalizer):
    """normalization.py:46-71 (7 + 5 + 3 + 4 + 4 + 4 = 27)."""

    def __init__(self, device="cuda"):
        super().__init__(7 + 5 + 3 + 4 + 4 + 4, device)


class TextFeatureNormalizer(FeatureNormalizer):
    """normalization.py:73-98 (7 + 1 + 1 + 1 + 1 + 768 + 4 = 783)."""

    def __init__(self, device="cuda"):
        super().__init__(7 + 1 + 1 + 1 + 1 + 768 + 4, device)


def assemble_row(pieces: Sequence[torch.Tensor], device=None) -> torch.Tensor:
    """streaming_processor.py:230-300: every piece is made [B, d] (``ensure_batch``), cast to float, concatenated on
    dim 1 (a copy: data movement) and scrubbed with ``torch.nan_to_num(nan=0.0)`` (msa_nan_to_num, in place)."""
    dev = _lib.require_cuda(device if device is not None else (pieces[0].device if pieces[0].is_cuda else "cuda"))
    cols = [p.to(dev).float().reshape(1, -1) if p.dim() < 2 else p.to(dev).float() for p in pieces]
    row = torch.cat(cols, dim=1).contiguous()
    _lib.check(_lib.lib().msa_nan_to_num(_lib.ptr(row), row.numel(), _lib.current_stream_ptr(dev)), "msa_nan_to_num")
    return row

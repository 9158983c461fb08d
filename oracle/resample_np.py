"""numpy restatement of ``torchaudio.transforms.Resample(sr, 16000)`` as the reference uses it
(/root/reference/src/analyzers/audio_analyzer.py:74-77; torchaudio is not vendored there: pinned 2.5.1 in
requirements.txt:350, container 2.11.0).  TEST INFRASTRUCTURE (see oracle/__init__.py).

Published algorithm (torchaudio/functional/functional.py, ``_get_sinc_resample_kernel`` /
``_apply_sinc_resample_kernel``, defaults sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99):
a polyphase windowed-sinc FIR, one filter per output phase j in [0, new), applied with stride ``orig``
to the zero-padded input.  Pinned against torchaudio's own output in tests/golden/resample_golden.npz.
"""
from __future__ import annotations

import math

import numpy as np

LOWPASS_WIDTH = 6
ROLLOFF = 0.99


def reduced(orig_freq: int, new_freq: int):
    g = math.gcd(int(orig_freq), int(new_freq))
    return int(orig_freq) // g, int(new_freq) // g


def sinc_kernel(orig_freq: int, new_freq: int):
    """-> (kernel [new, K] fp32, width), K = 2 width + orig, exactly as torchaudio builds it (float64, then
    cast to fp32; the phase offsets -j/new go through a float32 division first, like ``arange(...)/new_freq``)."""
    orig, new = reduced(orig_freq, new_freq)
    base = min(orig, new) * ROLLOFF
    width = math.ceil(LOWPASS_WIDTH * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    phase = (np.arange(0, -new, -1).astype(np.float32) / np.float32(new)).astype(np.float64)[:, None]
    t = (phase + idx) * base
    t = np.clip(t, -LOWPASS_WIDTH, LOWPASS_WIDTH)
    window = np.cos(t * math.pi / LOWPASS_WIDTH / 2.0) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0.0, 1.0, np.sin(t) / t)
    k = k * (window * (base / orig))
    return k.astype(np.float32), width


def resample(x: np.ndarray, orig_freq: int, new_freq: int) -> np.ndarray:
    """x [..., L] -> [..., ceil(new L / orig)] (fp64 accumulation of the fp32 kernel and samples)."""
    if int(orig_freq) == int(new_freq):
        return np.asarray(x)
    orig, new = reduced(orig_freq, new_freq)
    k, width = sinc_kernel(orig_freq, new_freq)
    x = np.asarray(x, dtype=np.float32)
    lead, L = x.shape[:-1], x.shape[-1]
    x2 = x.reshape(-1, L).astype(np.float64)
    xp = np.pad(x2, ((0, 0), (width, width + orig)))
    K = k.shape[1]
    n_frames = (xp.shape[1] - K) // orig + 1
    frames = np.lib.stride_tricks.sliding_window_view(xp, K, axis=1)[:, ::orig][:, :n_frames]   # [n, frames, K]
    y = np.einsum("nfk,jk->nfj", frames, k.astype(np.float64)).reshape(x2.shape[0], -1)
    target = int(math.ceil(new * L / orig))
    return y[:, :target].reshape(lead + (target,))

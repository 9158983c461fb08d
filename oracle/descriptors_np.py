"""numpy restatement of the additive descriptors (TEST INFRASTRUCTURE).

The reference computes NO f0, voicing or class probabilities: these are additional outputs
named by the north star, so PARITY IS UNPINNED BY THE REFERENCE.  They are pinned instead against the same
third-party dependency the reference uses (torchaudio, pinned 2.5.1 in requirements.txt:350, container 2.11.0):

  pitch_lags / pitch_frequency   torchaudio.functional.detect_pitch_frequency(waveform, 16000) with its defaults
                                 (frame 10 ms, 85..3400 Hz, median window 30): NCCF -> best lag per frame (the
                                 first half of the lag range wins when within 1 %) -> lower-median smoothing
  voiced_frames                  the per-frame analogue of _analyze_speech_rate's ``energy > 0.1 * energy.mean()``
                                 (audio_analyzer.py:223-228) on the 400/160 rhythm frames of :239-249
  class_probs                    softmax over the 7 fused logits (fusion_model.py:94 leaves them as logits)
  spectral_descriptors           per frame of the MFCC's STFT grid (n_fft 400, hop 200, periodic Hann, centre + reflect):
                                 centroid = torchaudio.functional.spectral_centroid (pinned against it), roll-off = first
                                 bin whose cumulative MAGNITUDE reaches 85 % (librosa.feature.spectral_rolloff's rule),
                                 flux = L2 norm of the magnitude difference to the previous frame, onset = mean over the
                                 128 HTK mel bands of the positive part of the log-power difference (the core of
                                 librosa.onset.onset_strength); stated definitions, first frame 0 for the two differences

tests/golden/descriptors_golden.npz holds torchaudio's own outputs (oracle/make_golden_ingest.py).
"""
from __future__ import annotations

import math

import numpy as np

SR = 16000
FRAME = 160          # ceil(16000 * 0.01)
LAGS = 189           # ceil(16000 / 85)
LAG_MIN = 5          # ceil(16000 / 3400)
MEDIAN_WIN = 30
EPS = np.float32(1e-9)


def nccf(x: np.ndarray) -> np.ndarray:
    """[T] fp32 -> [n_frames, LAGS] fp32, column a <-> lag a + 1 (functional.py _compute_nccf)."""
    x = np.asarray(x, dtype=np.float32)
    T = x.shape[0]
    nf = int(math.ceil(T / FRAME))
    xp = np.concatenate([x, np.zeros(LAGS + nf * FRAME - T, np.float32)])
    s1 = xp[: nf * FRAME].reshape(nf, FRAME)
    n1 = (EPS + np.sqrt((s1 * s1).sum(-1, dtype=np.float32))) ** 2
    out = np.empty((nf, LAGS), np.float32)
    for lag in range(1, LAGS + 1):
        s2 = xp[lag: lag + nf * FRAME].reshape(nf, FRAME)
        n2 = (EPS + np.sqrt((s2 * s2).sum(-1, dtype=np.float32))) ** 2
        out[:, lag - 1] = (s1 * s2).sum(-1, dtype=np.float32) / n1 / n2
    return out


def pitch_lags(x: np.ndarray) -> np.ndarray:
    """Best lag (in samples) per 10 ms frame before smoothing (functional.py _find_max_per_frame)."""
    c = nccf(x)
    best_i = c[:, LAG_MIN:].argmax(-1)
    best_v = c[:, LAG_MIN:].max(-1)
    half = LAGS // 2
    half_i = c[:, LAG_MIN:half].argmax(-1)
    half_v = c[:, LAG_MIN:half].max(-1)
    mask = half_v > np.float32(0.99) * best_v
    return (np.where(mask, half_i, best_i) + LAG_MIN + 1).astype(np.int32)


def median_smooth(lags: np.ndarray) -> np.ndarray:
    """functional.py _median_smoothing: 14 copies of the first value in front, windows of 30, LOWER median."""
    pad = (MEDIAN_WIN - 1) // 2
    a = np.concatenate([np.full(pad, lags[0], lags.dtype), lags])
    if a.shape[0] < MEDIAN_WIN:
        return np.zeros(0, lags.dtype)
    w = np.lib.stride_tricks.sliding_window_view(a, MEDIAN_WIN)
    return np.sort(w, axis=-1)[:, (MEDIAN_WIN - 1) // 2]


def pitch_frequency(x: np.ndarray) -> np.ndarray:
    m = median_smooth(pitch_lags(x))
    # torch evaluates ``sample_rate / tensor`` as reciprocal(tensor) * sample_rate, in fp32
    return ((np.float32(1.0) / (EPS + m.astype(np.float32))) * np.float32(SR)).astype(np.float32)


def voiced_frames(x: np.ndarray) -> np.ndarray:
    """int32 [L], L = (T - 400) // 160 + 1: frame energy above a tenth of the mean frame energy."""
    x = np.asarray(x, dtype=np.float32)
    if x.shape[0] < 400:
        return np.zeros(0, np.int32)
    L = (x.shape[0] - 400) // 160 + 1
    fr = np.lib.stride_tricks.sliding_window_view(x, 400)[::160][:L].astype(np.float64)
    e = (fr * fr).sum(-1)
    return (e > 0.1 * e.mean()).astype(np.int32)


def class_probs(logits: np.ndarray) -> np.ndarray:
    z = np.asarray(logits, dtype=np.float64)
    z = z - z.max(-1, keepdims=True)
    p = np.exp(z)
    return p / p.sum(-1, keepdims=True)


def stft400_magnitude(x: np.ndarray) -> np.ndarray:
    """[T] -> [n_frames, 201] fp64 magnitude of the STFT inside torchaudio.transforms.MFCC (n_fft 400, hop 200, periodic
    Hann, center=True, reflect padding)."""
    x = np.asarray(x, dtype=np.float64)
    n = 400
    xp = np.pad(x, (n // 2, n // 2), mode="reflect")
    nf = x.shape[0] // 200 + 1
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)
    fr = np.lib.stride_tricks.sliding_window_view(xp, n)[::200][:nf]
    return np.abs(np.fft.rfft(fr * w, axis=-1))


def spectral_descriptors(x: np.ndarray) -> np.ndarray:
    """[T] -> [n_frames, 4] fp64: centroid [Hz], roll-off [Hz], flux, onset strength (definitions in the module header)."""
    from oracle import features_np as fx
    S = stft400_magnitude(x)
    f = 40.0 * np.arange(201)
    with np.errstate(invalid="ignore", divide="ignore"):
        cen = (S * f).sum(-1) / S.sum(-1)
    cum = np.cumsum(S, axis=-1)
    roll = 40.0 * (cum >= 0.85 * cum[:, -1:]).argmax(-1)
    flux = np.concatenate([[0.0], np.sqrt(((S[1:] - S[:-1]) ** 2).sum(-1))])
    L = 10.0 * np.log10(np.maximum((S * S) @ fx.mel_fbanks(), 1e-10))
    onset = np.concatenate([[0.0], np.maximum(L[1:] - L[:-1], 0.0).mean(-1)])
    return np.stack([cen, roll, flux, onset], axis=-1)

"""First-principles numpy restatement of the reference's audio feature path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference lines it restates; paths are relative to /root/reference.  The DSP
that the reference delegates to torchaudio (pinned 2.5.1 in
requirements.txt:350, container has 2.11.0; not vendored under
/root/reference) is restated from the published algorithm:
``torchaudio.transforms.MFCC`` (Spectrogram -> MelScale(htk) -> AmplitudeToDB
(power, top_db=80) -> DCT-II ortho) and ``torchaudio.transforms.PitchShift``
with ``n_steps=0`` (STFT 512/128 -> identity phase vocoder -> ISTFT).

Arithmetic is float64 unless ``dtype=np.float32`` is passed; the reference
computes in fp32, so agreement is to ~1e-5 relative (tests/golden pins it).
A waveform here is a 1-D array [T]; the reference's layout is [1, T] mono
(anything else falls to defaults).
"""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np

SR = 16000
N_FFT_MFCC, HOP_MFCC = 400, 200          # torchaudio MFCC defaults (melkwargs empty)
N_MELS, N_MFCC = 128, 13
N_FFT_PITCH, HOP_PITCH = 512, 128        # torchaudio PitchShift defaults (n_fft=512, hop=n_fft//4)
RHYTHM_WIN, RHYTHM_HOP = 400, 160        # audio_analyzer.py:52-53,239-240 at 16 kHz
BLOCK = 1600                             # audio_analyzer.py:317 (100 ms)
TOP_DB = 80.0
AMIN = 1e-10


# ------------------------------------------------------------------ tables
@lru_cache(maxsize=None)
def hann(n: int) -> np.ndarray:
    """Periodic Hann window (torch.hann_window default periodic=True)."""
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


@lru_cache(maxsize=None)
def mel_fbanks() -> np.ndarray:
    """[201, 128] HTK triangular filterbank, norm=None, f in [0, 8000]
    (torchaudio.functional.melscale_fbanks as called by MelScale defaults)."""
    n_freqs = N_FFT_MFCC // 2 + 1
    all_freqs = np.linspace(0.0, SR // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + 0.0 / 700.0)
    m_max = 2595.0 * math.log10(1.0 + (SR / 2.0) / 700.0)
    m_pts = np.linspace(m_min, m_max, N_MELS + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(0.0, np.minimum(down, up))


@lru_cache(maxsize=None)
def dct_matrix() -> np.ndarray:
    """[128, 13] DCT-II, norm='ortho' (torchaudio.functional.create_dct)."""
    n = np.arange(N_MELS, dtype=np.float64)
    k = np.arange(N_MFCC, dtype=np.float64)[:, None]
    dct = np.cos(math.pi / N_MELS * (n + 0.5) * k)
    dct[0] *= 1.0 / math.sqrt(2.0)
    dct *= math.sqrt(2.0 / N_MELS)
    return dct.T.copy()


# ------------------------------------------------------------------ helpers
def _std_unbiased(a: np.ndarray) -> float:
    """torch.std default (correction=1); one element -> NaN like torch."""
    n = a.size
    if n < 2:
        return float("nan")
    m = a.mean()
    return float(np.sqrt(((a - m) ** 2).sum() / (n - 1)))


def _frames_centered(x: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    """torch.stft(center=True, pad_mode='reflect') framing -> [n_frames, n_fft]."""
    pad = n_fft // 2
    if x.size <= pad:
        raise ValueError("reflect padding needs T > n_fft/2")
    xp = np.pad(x, (pad, pad), mode="reflect")
    n_frames = 1 + (xp.size - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    return xp[idx]


# ------------------------------------------------------------------ MFCC
def power_spectrogram(x: np.ndarray) -> np.ndarray:
    """[n_frames, 201] |STFT|^2, n_fft=400 hop=200 periodic Hann."""
    fr = _frames_centered(x, N_FFT_MFCC, HOP_MFCC) * hann(N_FFT_MFCC)[None, :]
    spec = np.fft.rfft(fr, axis=1)
    return spec.real ** 2 + spec.imag ** 2


def mel_db(x: np.ndarray) -> np.ndarray:
    """[n_frames, 128] dB mel spectrogram with the whole-segment top_db clamp
    (torchaudio.functional.amplitude_to_DB: 10*log10(clamp(x, 1e-10)), then
    max(x_db, x_db.max() - 80))."""
    mel = power_spectrogram(x) @ mel_fbanks()
    db = 10.0 * np.log10(np.maximum(mel, AMIN))
    return np.maximum(db, db.max() - TOP_DB)


def mfcc(x: np.ndarray) -> np.ndarray:
    """[13, n_frames] — torchaudio.transforms.MFCC(sample_rate=16000, n_mfcc=13)."""
    return (mel_db(x) @ dct_matrix()).T


def timbre(x: np.ndarray, dtype=np.float64) -> np.ndarray:
    """audio_analyzer.py:203-217 -> [13]."""
    x = np.asarray(x, dtype=dtype)
    try:
        c = mfcc(x)
        z = (c - c.mean()) / (_std_unbiased(c) + 1e-6)
        return z.mean(axis=1)
    except Exception:
        return np.zeros(N_MFCC)


def clarity(x: np.ndarray, dtype=np.float64) -> float:
    """audio_analyzer.py:295-311."""
    x = np.asarray(x, dtype=dtype)
    try:
        c = mfcc(x)
        hi = np.abs(c[6:]).mean()
        lo = np.abs(c[:6]).mean()
        v = float(hi / (lo + 1e-6))
        return min(max(v, 0), 1)
    except Exception:
        return 0.0


# ------------------------------------------------------------------ "pitch"
def pitch_roundtrip(x: np.ndarray) -> np.ndarray:
    """PitchShift(n_steps=0)(x): stft 512/128 -> phase_vocoder(rate=1.0)
    (returns its input) -> torch.istft(length=T).  [T]."""
    T = x.size
    w = hann(N_FFT_PITCH)
    fr = _frames_centered(x, N_FFT_PITCH, HOP_PITCH) * w[None, :]
    spec = np.fft.rfft(fr, axis=1)
    yfr = np.fft.irfft(spec, n=N_FFT_PITCH, axis=1) * w[None, :]
    n_frames = yfr.shape[0]
    total = N_FFT_PITCH + HOP_PITCH * (n_frames - 1)
    y = np.zeros(total)
    env = np.zeros(total)
    w2 = w * w
    for f in range(n_frames):
        y[f * HOP_PITCH:f * HOP_PITCH + N_FFT_PITCH] += yfr[f]
        env[f * HOP_PITCH:f * HOP_PITCH + N_FFT_PITCH] += w2
    start = N_FFT_PITCH // 2
    y = y[start:start + T]
    env = env[start:start + T]
    if np.abs(env).min() < 1e-11:
        raise ValueError("window overlap add min")
    y = y / env
    if y.size < T:
        y = np.pad(y, (0, T - y.size))
    return y


def pitch(x: np.ndarray, dtype=np.float64) -> float:
    """audio_analyzer.py:175-188 -> scalar (reference shape [1,1]).
    Mathematically zero: the time-mean of a globally z-scored array."""
    x = np.asarray(x, dtype=dtype)
    try:
        p = np.abs(x - pitch_roundtrip(x))
        p = (p - p.mean()) / (_std_unbiased(p) + 1e-6)
        return float(p.mean())
    except Exception:
        return 0.0


# ------------------------------------------------------------------ wave stats
def intensity(x: np.ndarray, dtype=np.float64) -> float:
    """audio_analyzer.py:190-201.  One channel -> std of one element -> NaN."""
    x = np.asarray(x, dtype=dtype)
    e = np.array([np.sum(x * x)])
    return float(((e - e.mean()) / (_std_unbiased(e) + 1e-6))[0])


def speech_rate(x: np.ndarray, dtype=np.float64) -> float:
    """audio_analyzer.py:219-233.  Mono: 1.0 iff energy > 0.1*energy."""
    x = np.asarray(x, dtype=dtype)
    e = x.dtype.type(np.sum(x * x))
    thr = e * x.dtype.type(0.1)
    return 1.0 if e > thr else 0.0


def rhythm(x: np.ndarray, dtype=np.float64) -> np.ndarray:
    """audio_analyzer.py:235-263 -> [mean, unbiased std, L/16000]; T<400 -> zeros."""
    x = np.asarray(x, dtype=dtype)
    if x.size < RHYTHM_WIN:
        return np.zeros(3)
    L = (x.size - RHYTHM_WIN) // RHYTHM_HOP + 1
    idx = np.arange(RHYTHM_WIN)[None, :] + RHYTHM_HOP * np.arange(L)[:, None]
    e = (x[idx] ** 2).sum(axis=1)
    return np.array([e.mean(), _std_unbiased(e), L / SR])


def snr(x: np.ndarray, dtype=np.float64) -> float:
    """audio_analyzer.py:278-293."""
    x = np.asarray(x, dtype=dtype)
    n = int(0.05 * x.size)
    if n == 0:
        # waveform[:, :0] is [1,0] and waveform[:, -0:] is [1,T]: torch.cat raises -> 0.0
        return 0.0
    noise = np.concatenate([x[:n], x[-n:]])
    noise_power = np.mean(noise ** 2)
    signal_power = np.mean(x ** 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        v = float(10.0 * np.log10(signal_power / (noise_power + 1e-6)))
    return min(max(v / 30, 0), 1)


def consistency(x: np.ndarray, dtype=np.float64) -> float:
    """audio_analyzer.py:313-329.  T<1600 -> unfold raises -> 0.0."""
    x = np.asarray(x, dtype=dtype)
    nb = x.size // BLOCK
    if nb < 1:
        return 0.0
    m = (x[:nb * BLOCK].reshape(nb, BLOCK) ** 2).mean(axis=1)
    cv = _std_unbiased(m) / (m.mean() + 1e-6)
    return 1.0 - min(float(cv), 1.0)


def audio_quality(x: np.ndarray, dtype=np.float64) -> float:
    """audio_analyzer.py:265-276."""
    return 0.4 * snr(x, dtype) + 0.3 * clarity(x, dtype) + 0.3 * consistency(x, dtype)


# ------------------------------------------------------------------ row assembly
def layer_norm(v: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """nn.LayerNorm with weight 1 / bias 0: biased variance over the last axis."""
    m = v.mean(axis=-1, keepdims=True)
    var = ((v - m) ** 2).mean(axis=-1, keepdims=True)
    return (v - m) / np.sqrt(var + eps)


def raw_features(x: np.ndarray, emo8=None, dtype=np.float64) -> np.ndarray:
    """The 27 un-normalised features in the order audio_analyzer.py:113-120
    concatenates them: emotion8, pitch, intensity, timbre13, speech_rate, rhythm3."""
    if emo8 is None:
        emo8 = np.full(8, 0.125)       # stubbed wav2vec2 -> uniform (audio_analyzer.py:171-173)
    return np.concatenate([np.asarray(emo8, dtype=np.float64), [pitch(x, dtype)], [intensity(x, dtype)],
                           timbre(x, dtype), [speech_rate(x, dtype)], rhythm(x, dtype)])


def quality4(x: np.ndarray, dtype=np.float64) -> np.ndarray:
    """[audio_quality, snr, clarity, consistency] (audio_analyzer.py:128-131)."""
    return np.array([audio_quality(x, dtype), snr(x, dtype), clarity(x, dtype), consistency(x, dtype)], dtype=np.float64)


def ln31(raw27: np.ndarray) -> np.ndarray:
    """AudioFeatureNormalizer.normalize (src/utils/normalization.py:26-44):
    zero-pad 27 -> 31, LayerNorm(31).  Returns all 31 values."""
    row = np.concatenate([raw27, np.zeros(4)])
    with np.errstate(invalid="ignore"):
        return layer_norm(row)


def audio_row31(x: np.ndarray, emo8=None, dtype=np.float64, finite_intensity: bool = False) -> np.ndarray:
    """The [31] row StreamingProcessor hands to the fusion model
    (streaming_processor.py:250-268, 295-298): LN31(raw27)[:27] ++ quality4,
    then torch.nan_to_num(nan=0.0).  ``finite_intensity`` replaces the mono NaN
    intensity by 0 so the LayerNorm arithmetic itself can be compared."""
    raw = raw_features(x, emo8, dtype)
    if finite_intensity:
        raw[9] = 0.0
    row = np.concatenate([ln31(raw)[:27], quality4(x, dtype)])
    fmax = np.finfo(np.float32).max
    return np.nan_to_num(row, nan=0.0, posinf=fmax, neginf=-fmax)

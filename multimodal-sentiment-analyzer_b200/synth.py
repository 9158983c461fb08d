"""Deterministic synthetic inputs for the hot path (bench.py and the tests).

numpy ``default_rng`` (PCG64) only, so the same seed gives the same bytes in
the build container and on the GPU box.  Shapes follow SURVEY.md section 8(d):
5 s / 16 kHz mono PCM segments, an 8-way emotion embedding, a 27-wide face
row and a 783-wide text row.

The waveform generator mirrors how audio reaches the reference: int16 PCM
(``pcm_s16le``, /root/reference/src/processors/offline_processor.py:87-91 and
streaming_processor.py:185-196) converted to fp32 by ``/32768``.
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
SEGMENT_SECONDS = 5.0
SEGMENT_SAMPLES = 80000

FACE_DIM = 27
AUDIO_DIM = 31
TEXT_DIM = 783
HIDDEN_DIM = 1024
OUT_DIM = 7


def segment_pcm(seed: int, n_samples: int = SEGMENT_SAMPLES) -> np.ndarray:
    """One voiced-speech-like segment as int16 PCM, shape [n_samples]."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / SAMPLE_RATE
    f0 = rng.uniform(80.0, 300.0)
    rate = rng.uniform(2.0, 6.0)
    phase = rng.uniform(0.0, 2.0 * np.pi, size=5)
    x = np.zeros(n_samples, dtype=np.float64)
    for k in range(1, 6):
        x += (0.3 / k) * np.sin(2.0 * np.pi * f0 * k * t + phase[k - 1])
    x *= 0.5 + 0.5 * np.sin(2.0 * np.pi * rate * t)
    x += 0.02 * rng.standard_normal(n_samples)
    x = np.clip(x, -1.0, 1.0)
    return np.round(x * 32767.0).astype(np.int16)


def segments_pcm(first_seed: int, count: int, n_samples: int = SEGMENT_SAMPLES) -> np.ndarray:
    """[count, n_samples] int16, segment i uses seed first_seed + i."""
    out = np.empty((count, n_samples), dtype=np.int16)
    for i in range(count):
        out[i] = segment_pcm(first_seed + i, n_samples)
    return out


def pcm_to_f32(pcm: np.ndarray) -> np.ndarray:
    """int16 -> fp32 the way torchaudio.load normalises PCM (x / 32768)."""
    return (pcm.astype(np.float32) / np.float32(32768.0)).astype(np.float32)


def fast_segments_pcm(seed: int, count: int, n_samples: int = SEGMENT_SAMPLES) -> np.ndarray:
    """Vectorised bulk generator for benchmark-sized batches (same signal
    family as ``segment_pcm`` but one RNG stream for the whole batch)."""
    rng = np.random.default_rng(seed)
    t = (np.arange(n_samples, dtype=np.float32) / np.float32(SAMPLE_RATE))[None, :]
    f0 = rng.uniform(80.0, 300.0, size=(count, 1)).astype(np.float32)
    rate = rng.uniform(2.0, 6.0, size=(count, 1)).astype(np.float32)
    x = np.zeros((count, n_samples), dtype=np.float32)
    for k in range(1, 6):
        ph = rng.uniform(0.0, 2.0 * np.pi, size=(count, 1)).astype(np.float32)
        x += np.float32(0.3 / k) * np.sin(np.float32(2.0 * np.pi) * f0 * k * t + ph)
    x *= 0.5 + 0.5 * np.sin(np.float32(2.0 * np.pi) * rate * t)
    x += 0.02 * rng.standard_normal((count, n_samples), dtype=np.float32)
    np.clip(x, -1.0, 1.0, out=x)
    return np.round(x * 32767.0).astype(np.int16)


# ---------------------------------------------------------------- adversarial
def adversarial_cases() -> dict:
    """Named fp32 waveforms [T] covering the edge cases SURVEY.md section 4 lists."""
    rng = np.random.default_rng(99)
    cases = {}
    cases["white_0p1"] = pcm_to_f32(np.round(np.clip(0.1 * rng.standard_normal(80000), -1, 1) * 32767).astype(np.int16))
    cases["zeros"] = np.zeros(80000, dtype=np.float32)
    cases["noise_1e-4"] = (1e-4 * rng.standard_normal(80000)).astype(np.float32)
    t = np.arange(80000) / SAMPLE_RATE
    cases["tone_220"] = (0.5 * np.sin(2 * np.pi * 220.0 * t) + 1e-3 * rng.standard_normal(80000)).astype(np.float32)
    cases["half_silence"] = np.concatenate(
        [pcm_to_f32(segment_pcm(7, 40000)), np.zeros(40000, dtype=np.float32)])
    cases["short_8000"] = pcm_to_f32(segment_pcm(11, 8000))
    cases["odd_12345"] = pcm_to_f32(segment_pcm(12, 12345))
    cases["short_1700"] = pcm_to_f32(segment_pcm(13, 1700))
    cases["short_500"] = pcm_to_f32(segment_pcm(14, 500))
    cases["short_300"] = pcm_to_f32(segment_pcm(15, 300))
    cases["long_10s"] = pcm_to_f32(segment_pcm(16, 160000))
    return cases


# ------------------------------------------------------------- other modalities
def emotion_probs(seed: int, count: int, uniform: bool = False) -> np.ndarray:
    """[count, 8] fp32.  ``uniform`` = the reference's wav2vec2-failure fallback
    (1/8 each, audio_analyzer.py:171-173); otherwise softmax(N(0,1)) over four
    classes duplicated to eight and renormalised (audio_analyzer.py:163-168)."""
    if uniform:
        return np.full((count, 8), 0.125, dtype=np.float32)
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((count, 4))
    p = np.exp(z - z.max(axis=1, keepdims=True))
    p /= p.sum(axis=1, keepdims=True)
    p8 = np.concatenate([p, p], axis=1)
    p8 /= p8.sum(axis=1, keepdims=True)
    return p8.astype(np.float32)


def face_rows(seed: int, count: int) -> np.ndarray:
    """[count, 27] fp32: 23 x N(0,1) then raw pixel box (x, y, w, h), the way
    streaming_processor.py:233-248 feeds un-normalised coordinates to face_norm."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((count, 23))
    box = np.stack([rng.integers(0, 640, count), rng.integers(0, 480, count),
                    rng.integers(50, 301, count), rng.integers(50, 301, count)], axis=1)
    return np.concatenate([a, box.astype(np.float64)], axis=1).astype(np.float32)


def text_rows(seed: int, count: int) -> np.ndarray:
    """[count, 783] fp32: 779 x N(0,1) then 4 x U(0,1) quality scalars."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((count, 779), dtype=np.float32)
    q = rng.random((count, 4), dtype=np.float32)
    return np.concatenate([a, q], axis=1)


def audio_rows(seed: int, count: int) -> np.ndarray:
    """[count, 31] fp32 stand-in for an audio row when fusion is tested alone:
    27 x N(0,1) then 4 x U(0,1)."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((count, 27), dtype=np.float32)
    q = rng.random((count, 4), dtype=np.float32)
    return np.concatenate([a, q], axis=1)


# ------------------------------------------------------------- fusion weights
_LINEARS = [
    ("face_proj", 1024, 27), ("audio_proj", 1024, 31), ("text_proj", 1024, 783),
    ("face_processor.3", 512, 1024), ("audio_processor.3", 512, 1024), ("text_processor.3", 512, 1024),
    ("fusion.0", 1024, 1536), ("fusion.4", 512, 1024), ("fusion.8", 7, 512),
    ("fusion2", 1024, 1024),
]
_NORMS = [
    ("face_norm", 27), ("audio_norm", 31), ("text_norm", 783),
    ("face_processor.0", 1024), ("audio_processor.0", 1024), ("text_processor.0", 1024),
    ("face_processor.4", 512), ("audio_processor.4", 512), ("text_processor.4", 512),
    ("fusion.1", 1024), ("fusion.5", 512),
]


def fusion_state(seed: int, trained_like: bool = False) -> dict:
    """A full 45-tensor state dict (names/shapes of SURVEY.md appendix A) as
    numpy fp32 arrays.  Xavier-uniform weights, zero bias, gamma=1/beta=0, the
    same distributions as fusion_model.py:114-120; ``trained_like`` perturbs
    biases and LayerNorm affine parameters so those code paths are exercised."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, fan_out, fan_in in _LINEARS:
        bound = np.sqrt(6.0 / (fan_in + fan_out))
        sd[name + ".weight"] = rng.uniform(-bound, bound, size=(fan_out, fan_in)).astype(np.float32)
        b = np.zeros(fan_out, dtype=np.float32)
        if trained_like:
            b = (0.05 * rng.standard_normal(fan_out)).astype(np.float32)
        sd[name + ".bias"] = b
    for name, dim in _NORMS:
        g = np.ones(dim, dtype=np.float32)
        b = np.zeros(dim, dtype=np.float32)
        if trained_like:
            g = (1.0 + 0.1 * rng.standard_normal(dim)).astype(np.float32)
            b = (0.05 * rng.standard_normal(dim)).astype(np.float32)
        sd[name + ".weight"] = g
        sd[name + ".bias"] = b
    sd["audio_weight"] = np.float32(0.3)
    sd["text_weight"] = np.float32(0.3)
    sd["face_weight"] = np.float32(0.4)
    return sd

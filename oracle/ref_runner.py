"""Run the UNMODIFIED reference (when its source tree is present) for bench.py's CPU arm.

TEST / BENCH INFRASTRUCTURE (see oracle/__init__.py).  BASELINE.md section 5 step 4: the reference is looked up at
``$MSA_REFERENCE_DIR`` and ``/root/reference``.  It is a pure-Python source tree that cannot travel to the GPU box,
so there ``available()`` is False and bench.py falls back to the pinned torch port (``cpu_baseline.kind = "port"``);
in the build container the CPU arm times the reference's own code (``kind = "reference"``).

The per-segment work timed here is the body of AudioAnalyzer.analyze (audio_analyzer.py:83-147) without the file read and
with the wav2vec2 classifier stubbed to the reference's own failure fallback, followed by the row assembly and
nan_to_num of streaming_processor.py:250-268, 295-298; the batched fusion forward is AdvancedFusionModel.forward in eval
mode (fusion_model.py:131-190).
"""
from __future__ import annotations

import os
import time
import warnings

import numpy as np


def reference_dir():
    for d in (os.environ.get("MSA_REFERENCE_DIR"), "/root/reference"):
        if d and os.path.isfile(os.path.join(d, "src", "analyzers", "audio_analyzer.py")):
            return d
    return None


def available() -> bool:
    return reference_dir() is not None


_cache = {}


def _modules():
    if "mods" not in _cache:
        os.environ["MSA_REFERENCE_DIR"] = reference_dir()
        from oracle import make_golden
        make_golden.REF = reference_dir()
        _cache["mods"] = make_golden.load_reference()
    return _cache["mods"]


def audio_row(ana, w):
    """[1, T] fp32 tensor -> [1, 31] with the reference's own methods."""
    import torch
    with warnings.catch_warnings(), torch.no_grad():
        warnings.simplefilter("ignore")
        feats = torch.cat([ana._analyze_emotion(w), ana._analyze_pitch(w), ana._analyze_intensity(w), ana._analyze_timbre(w),
                           ana._analyze_speech_rate(w), ana._analyze_rhythm(w)], dim=1)
        feats = ana.normalizer.normalize(feats)
        q = torch.tensor([ana._calculate_audio_quality(w), ana._calculate_signal_noise_ratio(w), ana._calculate_clarity(w),
                          ana._calculate_consistency(w)]).float()[None, :]
        return torch.nan_to_num(torch.cat([feats[:, :27].float(), q], dim=1), nan=0.0)


def worker(args):
    """Pool task with the signature of oracle.torch_port._worker: (rows [count, 31], seconds of compute)."""
    import logging
    import torch
    from oracle import synth
    seed, count = args
    torch.set_num_threads(1)
    logging.disable(logging.CRITICAL)            # the reference logs every stubbed wav2vec2 failure at ERROR level
    audio_mod, _ = _modules()
    if "ana" not in _cache:
        _cache["ana"] = audio_mod.AudioAnalyzer(device="cpu")
    ana = _cache["ana"]
    waves = [torch.from_numpy(synth.pcm_to_f32(synth.segment_pcm(seed + i)))[None, :] for i in range(count)]
    audio_row(ana, waves[0])
    t0 = time.perf_counter()
    rows = [audio_row(ana, w) for w in waves]
    dt = time.perf_counter() - t0
    return torch.cat(rows).numpy(), dt


def fusion_forward(sd: dict, face, audio, text=None):
    """AdvancedFusionModel.forward of the reference (eval mode) -> logits [n, 7]."""
    import torch
    _, fusion_mod = _modules()
    if "model" not in _cache:
        m = fusion_mod.AdvancedFusionModel(device="cpu")
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        m.eval()
        _cache["model"] = m
    with torch.no_grad():
        return _cache["model"](face, audio, text)["fused"]

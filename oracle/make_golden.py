"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Runs only in the build container (needs /root/reference, which does not exist
on the GPU box).  The reference modules are imported by file path with
``speechbrain.inference.foreign_class`` stubbed (speechbrain is not installed;
the stub raises inside ``classify_batch`` so ``_analyze_emotion`` takes its own
uniform-1/8 fallback, audio_analyzer.py:171-173), and ``torchaudio.load`` is
replaced by a ``wave``-module reader for ``analyze()`` (torchcodec is absent).

    python -m oracle.make_golden

Outputs (committed): tests/golden/features_golden.npz, fusion_golden.npz.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import types
import warnings
import wave

import numpy as np

REF = os.environ.get("MSA_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference():
    import torch  # noqa: F401
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    sb, inf = types.ModuleType("speechbrain"), types.ModuleType("speechbrain.inference")

    class _NoSER:
        def classify_batch(self, w):
            raise RuntimeError("wav2vec2 SER is out of scope (stub)")

    inf.foreign_class = lambda **kw: _NoSER()
    sys.modules["speechbrain"], sys.modules["speechbrain.inference"] = sb, inf

    def _load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m

    audio = _load("ref_audio_analyzer", "src/analyzers/audio_analyzer.py")
    fusion = _load("ref_fusion_model", "src/models/fusion_model.py")
    return audio, fusion


def _wave_reader(path):
    import torch
    with wave.open(path, "rb") as wf:
        sr = wf.getframerate()
        pcm = np.frombuffer(wf.readframes(wf.getnframes()), dtype=np.int16)
    return torch.from_numpy(pcm.astype(np.float32) / 32768.0)[None, :], sr


def _write_wav(path, pcm):
    with wave.open(path, "wb") as wf:
        wf.setnchannels(1)
        wf.setsampwidth(2)
        wf.setframerate(16000)
        wf.writeframes(pcm.astype(np.int16).tobytes())


def golden_features(audio_mod):
    import torch
    import torchaudio
    from oracle import synth

    ana = audio_mod.AudioAnalyzer(device="cpu")
    torchaudio.load = _wave_reader

    def per_method(x32):
        w = torch.from_numpy(np.ascontiguousarray(x32))[None, :]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = {
                "pitch": ana._analyze_pitch(w).detach().numpy().reshape(-1),
                "intensity": ana._analyze_intensity(w).detach().numpy().reshape(-1),
                "timbre": ana._analyze_timbre(w).detach().numpy().reshape(-1),
                "speech_rate": ana._analyze_speech_rate(w).detach().numpy().reshape(-1),
                "rhythm": ana._analyze_rhythm(w).detach().numpy().reshape(-1),
                "quality4": np.array([ana._calculate_audio_quality(w), ana._calculate_signal_noise_ratio(w),
                                      ana._calculate_clarity(w), ana._calculate_consistency(w)], dtype=np.float64),
            }
        return out

    data = {}
    # seeded 5 s segments
    seeds = list(range(1234, 1234 + 24))
    data["seeds"] = np.array(seeds)
    rows = [per_method(synth.pcm_to_f32(synth.segment_pcm(s))) for s in seeds]
    for k in rows[0]:
        data["seg_" + k] = np.stack([r[k] for r in rows])

    # full analyze() on the first 4 seeds through a real wav file (pins the LN31 NaN pattern
    # and the AudioAnalysis slices, audio_analyzer.py:113-147)
    an_rows = []
    with tempfile.TemporaryDirectory() as d:
        for s in seeds[:4]:
            p = os.path.join(d, "seg.wav")
            _write_wav(p, synth.segment_pcm(s))
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                a = ana.analyze(p, "spk0")
            an_rows.append(np.concatenate([
                a.emotion_probs.detach().numpy().reshape(-1), a.pitch.detach().numpy().reshape(-1),
                a.intensity.detach().numpy().reshape(-1), a.timbre.detach().numpy().reshape(-1),
                a.speech_rate.detach().numpy().reshape(-1), a.rhythm.detach().numpy().reshape(-1),
                [a.audio_quality, a.signal_noise_ratio, a.clarity, a.consistency]]))
    data["analyze_rows"] = np.stack(an_rows)           # [4, 31]: LN slices (NaN for mono) ++ quality

    # LayerNorm arithmetic with a finite row (intensity forced to 0): normalizer on raw27
    fin = []
    for s in seeds[:4]:
        r = rows[s - 1234]
        raw = np.concatenate([np.full(8, 0.125), r["pitch"], [0.0], r["timbre"], r["speech_rate"], r["rhythm"]])
        fin.append(ana.normalizer.normalize(torch.from_numpy(raw.astype(np.float32))[None, :]).detach().numpy()[0])
    data["ln31_finite"] = np.stack(fin)                # [4, 31]

    # adversarial / edge cases
    adv = synth.adversarial_cases()
    data["adv_names"] = np.array(list(adv.keys()))
    for name, x in adv.items():
        r = per_method(x)
        for k, v in r.items():
            data[f"adv_{name}_{k}"] = v
    return data


def golden_fusion(fusion_mod):
    import torch
    from oracle import synth

    data = {}
    for tag, trained in (("init", False), ("trained", True)):
        sd = synth.fusion_state(4321, trained_like=trained)
        model = fusion_mod.AdvancedFusionModel(device="cpu")
        missing = model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        model.eval()                                    # stated deviation: reference leaves dropout on
        n = 256
        face = torch.from_numpy(synth.face_rows(1, n))
        audio = torch.from_numpy(synth.audio_rows(2, n))
        text = torch.from_numpy(synth.text_rows(3, n))
        with torch.no_grad():
            r3 = model(face, audio, text)
            r2 = model(face, audio, None)
            rft = model(face, None, text)
            rat = model(None, audio, text)
            r1 = model(None, audio, None)
            rbad = model(face[:, :20], audio, text)
        data[f"{tag}_fused3"] = r3["fused"].numpy()
        data[f"{tag}_fused2"] = r2["fused"].numpy()
        data[f"{tag}_keys3"] = np.array(sorted(r3.keys()))
        data[f"{tag}_keys2"] = np.array(sorted(r2.keys()))
        data[f"{tag}_keys_face_text"] = np.array(sorted(rft.keys()))
        data[f"{tag}_keys_audio_text"] = np.array(sorted(rat.keys()))
        data[f"{tag}_keys_audio_only"] = np.array(sorted(r1.keys()))
        data[f"{tag}_keys_bad_dim"] = np.array(sorted(rbad.keys()))
        w = model.get_weights()
        data[f"{tag}_weights"] = np.array([w["audio"], w["text"], w["face"]])
    data["weight_seed"] = np.array(4321)
    data["row_seeds"] = np.array([1, 2, 3])
    return data


def golden_multichannel(audio_mod):
    """Per-method semantics for C >= 2 channels (audio_analyzer.py:190-217): intensity is a finite [1, C] z-score of
    the channel energies, timbre a [1, C, 13] z-score over all channels' MFCCs (analyze() itself never gets that far
    with multi-channel input: its torch.cat fails and it returns the default)."""
    import torch
    from oracle import synth
    ana = audio_mod.AudioAnalyzer(device="cpu")
    seeds = [4000, 4001, 4002]
    w = torch.from_numpy(np.stack([synth.pcm_to_f32(synth.segment_pcm(s)) for s in seeds]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return {"seeds": np.array(seeds), "intensity": ana._analyze_intensity(w).numpy(), "timbre": ana._analyze_timbre(w).numpy()}


def main():
    os.makedirs(OUT, exist_ok=True)
    audio_mod, fusion_mod = load_reference()
    np.savez_compressed(os.path.join(OUT, "multichannel_golden.npz"), **golden_multichannel(audio_mod))
    f = golden_features(audio_mod)
    np.savez_compressed(os.path.join(OUT, "features_golden.npz"), **f)
    g = golden_fusion(fusion_mod)
    np.savez_compressed(os.path.join(OUT, "fusion_golden.npz"), **g)
    for name in ("features_golden.npz", "fusion_golden.npz"):
        print(name, os.path.getsize(os.path.join(OUT, name)), "bytes")


if __name__ == "__main__":
    main()

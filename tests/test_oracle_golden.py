"""Pin the numpy oracle (oracle/features_np.py, oracle/fusion_np.py) against
golden vectors produced by the UNMODIFIED reference (oracle/make_golden.py).

Tolerances: timbre / rhythm[0:2] / quality floats rel 1e-3
with an absolute floor of 1e-6 — the fp64 oracle actually agrees with the fp32
reference to ~1e-5, which is what is asserted here; rhythm[2] and speech_rate
exact; pitch |v| <= 1e-6 absolute (the reference value is rounding noise);
intensity NaN for mono.
"""
import numpy as np
import pytest

from oracle import features_np as fx
from oracle import fusion_np as fu
from oracle import synth


def _close(a, b, rel=2e-5, floor=1e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b)), (a, b)      # NaN pattern must be identical
    ok = ~np.isnan(b)
    assert np.all(np.abs(a[ok] - b[ok]) <= rel * np.abs(b[ok]) + floor), (a, b)


def test_tables_match_torchaudio_shapes():
    fb = fx.mel_fbanks()
    assert fb.shape == (201, 128)
    assert int((fb > 1e-9).sum()) == 394                    # SURVEY appendix B (fp64 leaves one 1e-14 crumb)
    assert [int(i) for i in np.where(fb.max(axis=0) == 0)[0]] == [0, 3, 6, 13]
    d = fx.dct_matrix()
    assert d.shape == (128, 13)
    assert np.allclose(d.T @ d, np.eye(13), atol=1e-12)     # ortho


def test_seeded_segments_per_method(golden_features):
    g = golden_features
    for i, seed in enumerate(g["seeds"][:8]):
        x = synth.pcm_to_f32(synth.segment_pcm(int(seed)))
        _close(fx.timbre(x), g["seg_timbre"][i], rel=5e-5, floor=2e-5)
        r = fx.rhythm(x)
        _close(r[:2], g["seg_rhythm"][i][:2])
        assert np.float32(r[2]) == g["seg_rhythm"][i][2]
        assert fx.speech_rate(x) == g["seg_speech_rate"][i][0]
        assert abs(fx.pitch(x)) <= 1e-6 and abs(g["seg_pitch"][i][0]) <= 1e-6
        assert np.isnan(fx.intensity(x)) and np.isnan(g["seg_intensity"][i][0])
        _close(fx.quality4(x), g["seg_quality4"][i])


def test_pitch_roundtrip_is_identity():
    x = synth.pcm_to_f32(synth.segment_pcm(1234)).astype(np.float64)
    assert np.abs(fx.pitch_roundtrip(x) - x).max() < 1e-12
    x = synth.pcm_to_f32(synth.segment_pcm(5, 12345)).astype(np.float64)
    assert np.abs(fx.pitch_roundtrip(x) - x).max() < 1e-12


def test_analyze_rows_nan_pattern(golden_features):
    g = golden_features
    for i, seed in enumerate(g["seeds"][:4]):
        x = synth.pcm_to_f32(synth.segment_pcm(int(seed)))
        ref = g["analyze_rows"][i]
        assert np.all(np.isnan(ref[:27]))                   # mono -> intensity NaN -> LN row NaN
        raw = fx.raw_features(x)
        mine = np.concatenate([fx.ln31(raw)[:27], fx.quality4(x)])
        assert np.array_equal(np.isnan(mine), np.isnan(ref))
        _close(mine[27:], ref[27:])
        row = fx.audio_row31(x)
        assert np.all(row[:27] == 0.0)
        _close(row[27:], ref[27:])


def test_ln31_finite(golden_features):
    g = golden_features
    for i, seed in enumerate(g["seeds"][:4]):
        x = synth.pcm_to_f32(synth.segment_pcm(int(seed)))
        raw = fx.raw_features(x)
        raw[9] = 0.0
        _close(fx.ln31(raw), g["ln31_finite"][i], rel=5e-5, floor=2e-5)


@pytest.mark.parametrize("name", list(synth.adversarial_cases().keys()))
def test_adversarial(golden_features, name):
    g = golden_features
    x = synth.adversarial_cases()[name]
    ref = {k: g[f"adv_{name}_{k}"] for k in ("pitch", "intensity", "timbre", "speech_rate", "rhythm", "quality4")}
    # near-silent inputs put most mel bins at the 1e-10 floor or within fp32 noise of it
    rel, floor = (5e-5, 2e-5) if name not in ("noise_1e-4", "zeros") else (1e-3, 1e-4)
    _close(fx.timbre(x), ref["timbre"], rel=rel, floor=floor)
    r = fx.rhythm(x)
    _close(r[:2], ref["rhythm"][:2], rel=2e-5, floor=1e-9)
    assert np.float32(r[2]) == ref["rhythm"][2]
    assert fx.speech_rate(x) == ref["speech_rate"][0]
    assert abs(fx.pitch(x)) <= 1e-6 and abs(ref["pitch"][0]) <= 1e-6
    assert np.isnan(ref["intensity"][0]) and np.isnan(fx.intensity(x))
    q, qr = fx.quality4(x), ref["quality4"]
    assert np.array_equal(np.isnan(q), np.isnan(qr))
    ok = ~np.isnan(qr)
    _close(q[ok], qr[ok], rel=rel, floor=floor)


def test_fusion_against_reference(golden_fusion):
    g = golden_fusion
    n = g["init_fused3"].shape[0]
    face, audio, text = synth.face_rows(1, n), synth.audio_rows(2, n), synth.text_rows(3, n)
    for tag, trained in (("init", False), ("trained", True)):
        sd = synth.fusion_state(int(g["weight_seed"]), trained_like=trained)
        out3 = fu.forward(sd, face, audio, text)
        out2 = fu.forward(sd, face, audio, None)
        assert sorted(out3.keys()) == list(g[f"{tag}_keys3"])
        assert sorted(out2.keys()) == list(g[f"{tag}_keys2"])
        assert np.abs(out3["fused"] - g[f"{tag}_fused3"]).max() < 2e-5
        assert np.abs(out2["fused"] - g[f"{tag}_fused2"]).max() < 2e-5
        assert np.array_equal(out3["fused"].argmax(1), g[f"{tag}_fused3"].argmax(1))
        assert np.array_equal(out2["fused"].argmax(1), g[f"{tag}_fused2"].argmax(1))
        assert sorted(fu.forward(sd, face, None, text).keys()) == list(g[f"{tag}_keys_face_text"])
        assert sorted(fu.forward(sd, None, audio, text).keys()) == list(g[f"{tag}_keys_audio_text"])
        assert sorted(fu.forward(sd, None, audio, None).keys()) == list(g[f"{tag}_keys_audio_only"])
        assert sorted(fu.forward(sd, face[:, :20], audio, text).keys()) == list(g[f"{tag}_keys_bad_dim"])
        w = fu.get_weights(sd)
        assert np.allclose([w["audio"], w["text"], w["face"]], g[f"{tag}_weights"], rtol=1e-6)
    out3["face"] is face


def test_torch_port_matches_reference(golden_features, golden_fusion):
    """The timed CPU baseline (oracle/torch_port.py) reproduces the reference's numbers."""
    import torch
    from oracle import torch_port as tp
    g = golden_features
    ana = tp.PortedAnalyzer()
    for i, seed in enumerate(g["seeds"][:3]):
        w = torch.from_numpy(synth.pcm_to_f32(synth.segment_pcm(int(seed))))[None, :]
        row = ana.audio_row(w).numpy()[0]
        assert np.all(row[:27] == 0.0)
        _close(row[27:], g["seg_quality4"][i])
        _close(ana.timbre(w).numpy()[0], g["seg_timbre"][i], rel=1e-6, floor=1e-6)
        _close(ana.rhythm(w).numpy()[0], g["seg_rhythm"][i], rel=1e-6, floor=1e-7)
    gf = golden_fusion
    sd = tp.build_fusion(synth.fusion_state(int(gf["weight_seed"]), trained_like=True))
    n = gf["trained_fused3"].shape[0]
    f, a, t = (torch.from_numpy(v) for v in (synth.face_rows(1, n), synth.audio_rows(2, n), synth.text_rows(3, n)))
    assert np.abs(tp.fusion_forward(sd, f, a, t).numpy() - gf["trained_fused3"]).max() < 1e-5
    assert np.abs(tp.fusion_forward(sd, f, a, None).numpy() - gf["trained_fused2"]).max() < 1e-5

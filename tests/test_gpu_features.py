"""GPU parity of the fused feature kernel (through the C ABI) against the golden vectors of the
reference, the numpy oracle, and size-independent properties at the benchmark size.

Tolerances: timbre / rhythm[0:2] / quality floats rel 1e-3 (abs floor 1e-6; 2e-5 against the reference's fp32 LayerNorm golden rows);
rhythm[2] and speech_rate exact; "pitch" |v| <= 1e-6 absolute (the reference value is rounding
noise ~1e-9); intensity NaN for mono; NaN pattern of the LayerNorm row identical.
"""
import os
import wave

import numpy as np
import pytest
import torch

from oracle import features_np as fx
from oracle import synth
from tests.gpu_util import close, need_gpu, residual_is_fp16_noise

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ana():
    need_gpu()
    import msa_b200
    return msa_b200.AudioAnalyzer(device="cuda:0")


def _detail(ana, x, cluster=0, parts=7, emo=None, flags=None):
    from msa_b200 import _lib
    w = torch.from_numpy(np.ascontiguousarray(x)).to(ana.device)
    B, T = w.shape
    feat = torch.empty(B, 31, device=ana.device)
    det = torch.empty(B, 96, device=ana.device)
    mf = torch.empty(B, T // 200 + 1, 13, device=ana.device)
    e = None if emo is None else torch.from_numpy(emo).to(ana.device)
    fn = ana._lib.msa_features_s16 if w.dtype == torch.int16 else ana._lib.msa_features_f32
    rc = fn(_lib.ptr(w), B, T, _lib.ptr(e), _lib.ptr(feat), _lib.ptr(det), _lib.ptr(mf), ana._flags() if flags is None else flags,
            parts, cluster, _lib.current_stream_ptr(ana.device))
    assert rc == 0, _lib.strerror(rc)
    torch.cuda.synchronize()
    return feat.cpu().numpy(), det.cpu().numpy(), mf.cpu().numpy()


def _check_vs_reference_row(det, ref_timbre, ref_rhythm, ref_q, ref_rate, what):
    assert abs(det[8]) <= 1e-6, what
    assert np.isnan(det[9]), what
    close(det[10:23], ref_timbre, what=what + " timbre")
    close(det[24:26], ref_rhythm[:2], what=what + " rhythm")
    assert np.float32(det[26]) == np.float32(ref_rhythm[2]), what
    assert det[23] == ref_rate, what
    close(det[27:31], ref_q, what=what + " quality")


def test_golden_seeded_segments(ana, golden_features):
    g = golden_features
    seeds = [int(s) for s in g["seeds"]]
    x = np.stack([synth.pcm_to_f32(synth.segment_pcm(s)) for s in seeds])
    feat, det, _ = _detail(ana, x)
    for i in range(len(seeds)):
        _check_vs_reference_row(det[i], g["seg_timbre"][i], g["seg_rhythm"][i], g["seg_quality4"][i], g["seg_speech_rate"][i][0],
                                f"seed {seeds[i]}")
        assert np.all(feat[i, :27] == 0.0)                       # NaN row -> nan_to_num -> zeros (strict reference)
        close(feat[i, 27:], g["seg_quality4"][i])
    for i in range(4):                                          # analyze() rows: identical NaN pattern
        row = np.concatenate([det[i, 32:59], det[i, 27:31]])
        close(row, g["analyze_rows"][i], what="analyze row")


@pytest.mark.parametrize("cluster,T", [(1, 80000), (2, 80000), (4, 80000), (8, 80000), (16, 80000), (16, 8000), (1, 16000), (2, 16000), (2, 30001), (0, 80000), (0, 160000)])
def test_cluster_sizes_agree_with_oracle(ana, cluster, T):
    x = synth.pcm_to_f32(synth.segment_pcm(1234, T))[None]
    _, det, mf = _detail(ana, x, cluster=cluster)
    raw, q = fx.raw_features(x[0]), fx.quality4(x[0])
    close(det[0, 10:23], raw[10:23], what="timbre")
    close(det[0, 24:27], raw[24:27], what="rhythm")
    close(det[0, 27:31], q, what="quality")
    assert np.abs(mf[0] - fx.mfcc(x[0].astype(np.float64)).T).max() < 2e-3      # MFCC values reach ~170
    residual_is_fp16_noise(det[0], float(np.abs(x).max()))      # STFT -> ISTFT residual: std, max
    assert det[0, 72] == T                                       # every sample reconstructed exactly once
    from msa_b200 import _lib
    assert ana._lib.msa_features_f32(_lib.ptr(torch.zeros(1, 800000, device=ana.device)), 1, 800000, None,
                                     _lib.ptr(torch.zeros(1, 31, device=ana.device)), None, None, 1, 7, 1, None) == -2   # too long for 1 CTA


def test_int16_ingest_equals_f32(ana):
    pcm = synth.segments_pcm(2000, 3)
    f16, d16, _ = _detail(ana, pcm)
    f32, d32, _ = _detail(ana, synth.pcm_to_f32(pcm))
    assert np.array_equal(f16, f32)
    # every feature is bit-identical except the "pitch" slot: that value is the fp32 rounding residue of
    # an STFT -> ISTFT round trip (~1e-9, tolerance 1e-6 absolute), and the two template instantiations
    # of the kernel need not contract the same multiply-adds
    cols = [c for c in range(63) if c != 8]
    bad = [c for c in cols if not np.array_equal(d16[:, c], d32[:, c], equal_nan=True)]
    assert not bad, f"columns differ between int16 and fp32 ingest: {bad}"
    assert np.all(np.abs(d16[:, 8]) <= 1e-6) and np.all(np.abs(d32[:, 8]) <= 1e-6)


def test_finite_layernorm_and_emotion_embedding(ana, golden_features):
    g = golden_features
    seeds = [int(s) for s in g["seeds"][:4]]
    x = np.stack([synth.pcm_to_f32(synth.segment_pcm(s)) for s in seeds])
    feat, det, _ = _detail(ana, x, flags=0)                     # not strict: intensity 0 -> finite LayerNorm row
    for i in range(4):
        close(det[i, 32:63], g["ln31_finite"][i], rel=1e-3, floor=2e-5, what="ln31")
    emo = synth.emotion_probs(5, 4)
    feat, det, _ = _detail(ana, x, emo=emo, flags=0)
    for i in range(4):
        close(feat[i], fx.audio_row31(x[i], emo[i], finite_intensity=True), rel=1e-3, floor=2e-5, what="row with emotion")


@pytest.mark.parametrize("name", list(synth.adversarial_cases().keys()))
def test_adversarial_vs_golden(ana, golden_features, name):
    g = golden_features
    x = synth.adversarial_cases()[name]
    if x.size <= 256:
        pytest.skip("covered by the shim test")
    _, det, _ = _detail(ana, x[None])
    d = det[0]
    ref = {k: g[f"adv_{name}_{k}"] for k in ("pitch", "timbre", "speech_rate", "rhythm", "quality4")}
    rel, floor = (1e-3, 1e-6) if name not in ("noise_1e-4", "zeros") else (2e-3, 2e-4)
    assert abs(d[8]) <= 1e-6
    close(d[10:23], ref["timbre"], rel, floor, what="timbre")
    close(d[24:26], ref["rhythm"][:2], rel, 1e-9 + floor * 0, what="rhythm")
    assert np.float32(d[26]) == ref["rhythm"][2]
    assert d[23] == ref["speech_rate"][0]
    close(d[27:31], ref["quality4"], rel, floor, what="quality")


def test_reference_method_shims(ana, golden_features, tmp_path):
    g = golden_features
    x = synth.pcm_to_f32(synth.segment_pcm(1234))
    w = torch.from_numpy(x)[None].to(ana.device)
    assert ana._analyze_pitch(w).shape == (1, 1) and abs(ana._analyze_pitch(w).item()) <= 1e-6
    assert torch.isnan(ana._analyze_intensity(w)).all()
    close(ana._analyze_timbre(w).cpu().numpy()[0], g["seg_timbre"][0], what="timbre")
    assert ana._analyze_speech_rate(w).item() == 1.0
    r = ana._analyze_rhythm(w).cpu().numpy()[0]
    close(r[:2], g["seg_rhythm"][0][:2], what="rhythm")
    q = [ana._calculate_audio_quality(w), ana._calculate_signal_noise_ratio(w), ana._calculate_clarity(w), ana._calculate_consistency(w)]
    close(q, g["seg_quality4"][0], what="quality")
    assert torch.allclose(ana._analyze_emotion(w), torch.full((1, 8), 0.125, device=ana.device))
    # reference error convention: malformed input -> the method's default, never an exception
    assert torch.equal(ana._analyze_timbre(w[0]), torch.zeros(1, 13, device=ana.device))
    assert torch.equal(ana._analyze_rhythm(w[:, :300]), torch.zeros(1, 3, device=ana.device))
    assert ana._calculate_consistency(w[:, :1000]) == 0.0
    assert torch.equal(ana._analyze_pitch(w[:, :200]), torch.zeros(1, 1, device=ana.device))
    # analyze(path, speaker) through a real PCM wav file
    p = os.path.join(tmp_path, "seg.wav")
    with wave.open(p, "wb") as wf:
        wf.setnchannels(1); wf.setsampwidth(2); wf.setframerate(16000)
        wf.writeframes(synth.segment_pcm(1234).tobytes())
    a = ana.analyze(p, "spk7")
    assert a.speaker_id == "spk7" and a["timbre"].shape == (1, 13) and a.get("rhythm").shape == (1, 3)
    row = torch.cat([a.emotion_probs, a.pitch, a.intensity, a.timbre, a.speech_rate, a.rhythm], dim=1).cpu().numpy()[0]
    close(np.concatenate([row, [a.audio_quality, a.signal_noise_ratio, a.clarity, a.consistency]]), g["analyze_rows"][0], what="analyze")
    assert isinstance(a.to_dict(), dict)
    d = ana.analyze(os.path.join(tmp_path, "missing.wav"), "x")           # failure -> default analysis
    assert d.audio_quality == 0.0 and torch.equal(d.emotion_probs, torch.full((1, 8), 0.125, device=ana.device))


def test_full_size_batch_properties(ana):
    """BASELINE config 2 (1024 x 5 s): rows are independent and deterministic, flags exact."""
    pcm = torch.from_numpy(synth.fast_segments_pcm(7, 1024)).to(ana.device)
    f1, d1 = ana.analyze_batch(pcm, return_detail=True)
    f2, d2 = ana.analyze_batch(pcm, return_detail=True)
    torch.cuda.synchronize()
    assert torch.equal(f1, f2) and torch.equal(d1[:, :63].nan_to_num(7.0), d2[:, :63].nan_to_num(7.0))     # deterministic
    idx = [0, 511, 1023]
    fs, ds = ana.analyze_batch(pcm[idx].contiguous(), return_detail=True)
    assert torch.equal(fs, f1[idx])                                                  # batch == loop of [1,T] calls
    d = d1.cpu().numpy()
    assert np.all(np.abs(d[:, 8]) <= 1e-6) and np.all(np.isnan(d[:, 9])) and np.all(d[:, 23] == 1.0)
    assert np.all(d[:, 26] == np.float32(498 / 16000)) and np.all(d[:, 72] == 80000)
    for i in range(d.shape[0]):                                                      # reconstruction everywhere (fp16 round trip)
        residual_is_fp16_noise(d[i], 1.0)
    assert np.all((d[:, 27:31] >= 0) & (d[:, 27:31] <= 1))
    for i in idx:                                                                    # spot-check against the oracle
        x = synth.pcm_to_f32(pcm[i].cpu().numpy())
        close(d[i, 10:23], fx.timbre(x), what="timbre")
        close(d[i, 27:31], fx.quality4(x), what="quality")


@pytest.mark.parametrize("T", [1, 79, 80, 150, 200, 201, 256, 257, 300, 399, 400, 401, 513, 1599, 1600])
def test_tiny_lengths_match_oracle(ana, T):
    """Segments shorter than a transform's padding make the reference's method raise and return its default
    (zeros); the thresholds differ per feature (timbre needs T > 200, "pitch" T > 256, rhythm T >= 400, consistency
    a full 1600-sample block).  The oracle reproduces the reference on every one of these lengths."""
    x = synth.pcm_to_f32(synth.segment_pcm(50 + T, T))
    _, det, _ = _detail(ana, x[None])
    d = det[0]
    raw, q = fx.raw_features(x), fx.quality4(x)
    assert abs(d[8]) <= 1e-6 and np.isnan(d[9])
    close(d[10:23], raw[10:23], what="timbre")
    assert d[23] == raw[23]
    close(d[24:26], raw[24:26], what="rhythm")
    assert np.float32(d[26]) == np.float32(raw[26])
    close(d[27:31], q, what="quality")
    assert d[72] == (T if T > 256 else 0)                            # residual samples: all of them, or the part is off


def test_empty_batch_and_bad_arguments(ana):
    from msa_b200 import _lib
    lib = ana._lib
    feat = torch.zeros(1, 31, device=ana.device)
    w = torch.zeros(1, 80000, device=ana.device)
    assert lib.msa_features_f32(_lib.ptr(w), 0, 80000, None, _lib.ptr(feat), None, None, 1, 7, 0, None) == 0     # B = 0: nothing to do
    assert lib.msa_features_f32(None, 1, 80000, None, _lib.ptr(feat), None, None, 1, 7, 0, None) == -1
    assert lib.msa_features_f32(_lib.ptr(w), 1, 0, None, _lib.ptr(feat), None, None, 1, 7, 0, None) == -1
    assert lib.msa_features_f32(_lib.ptr(w), 1, 80000, None, _lib.ptr(feat), None, None, 1, 7, 3, None) == -1    # cluster size must be 1/2/4/8/16
    assert ana.analyze_batch(torch.zeros(0, 80000, device=ana.device)).shape == (0, 31)


def test_results_do_not_depend_on_cluster_size(ana):
    """One segment per CTA, or split over 2 / 4 / 8 / 16 CTAs (what small batches and streaming use): the same bits,
    whichever top_db path (patch list / clamped pass) each CTA takes.  Column 8 ("pitch") is the rounding residue of
    the STFT -> ISTFT round trip and depends on the summation tree (|v| <= 1e-6 either way)."""
    rng = np.random.default_rng(5)
    t = np.arange(80000) / 16000.0
    flip = 0.3 * np.sin(2 * np.pi * 180 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 3 * t)) + 0.004 * rng.standard_normal(80000)
    flip = np.round(np.clip(flip, -1, 1) * 32767).astype(np.int16).astype(np.float32) / np.float32(32768)
    adv = synth.adversarial_cases()
    x = np.stack([synth.pcm_to_f32(synth.segment_pcm(1234)), flip, adv["tone_220"], adv["half_silence"], adv["white_0p1"]])
    cols = [c for c in range(63) if c != 8]
    ref = None
    for c in (1, 2, 4, 8, 16):                             # 16: a non-portable cluster size, used where the device schedules it
        feat, det, mf = _detail(ana, x, cluster=c)
        if ref is None:
            ref = (feat, det, mf)
            continue
        assert np.array_equal(feat, ref[0]), c
        bad = [k for k in cols if not np.array_equal(det[:, k], ref[1][:, k], equal_nan=True)]
        assert not bad, (c, bad)
        assert np.array_equal(mf, ref[2]), c
        assert np.all(np.abs(det[:, 8]) <= 1e-6)


def test_batch_of_512_is_cluster_size_invariant(ana):
    """The sharded hour of BASELINE configs[2] runs 90 segments per GPU on 8 GPUs (2 CTAs per segment) and 720 on one
    (1 CTA per segment): every bit of the result table must agree, including the quality floats whose sums are
    partitioned differently."""
    pcm = synth.fast_segments_pcm(11, 512)
    f1, d1, _ = _detail(ana, pcm, cluster=1)
    f2, d2, _ = _detail(ana, pcm, cluster=2)
    assert np.array_equal(f1, f2)
    cols = [c for c in range(63) if c != 8]
    bad = [k for k in cols if not np.array_equal(d1[:, k], d2[:, k], equal_nan=True)]
    assert not bad, bad
    assert (f1[:, 28] > 0).any()                                    # some segments have a non-zero SNR term


def test_workspace_path_equals_recompute_path(ana):
    """Pause-heavy segments overflow the on-chip top_db candidate lists; with the scratch table the clamp is applied
    from saved dB values, without it the quads are recomputed: bit-identical results, and equal to the oracle."""
    from msa_b200 import _lib
    pcm = synth.fast_segments_pcm(21, 6)
    pcm[:, 8000:24000] = 0
    pcm[:, 48000:64000] = 0
    pcm[3, :] = synth.segment_pcm(1234)                               # one ordinary segment in the batch
    x = synth.pcm_to_f32(pcm)
    f0, d0, m0 = _detail(ana, x)                                        # plain entry point: no workspace
    w = torch.from_numpy(x).to(ana.device)
    B, T = w.shape
    ws = torch.full((ana._lib.msa_features_workspace_bytes(B, T) // 4,), float("nan"), device=ana.device)
    feat = torch.empty(B, 31, device=ana.device); det = torch.empty(B, 96, device=ana.device); mf = torch.empty(B, T // 200 + 1, 13, device=ana.device)
    for cluster in (1, 2, 8, 16):
        rc = ana._lib.msa_features_ws_f32(_lib.ptr(w), B, T, None, _lib.ptr(feat), _lib.ptr(det), _lib.ptr(mf), 1, 7, cluster,
                                          _lib.ptr(ws), ws.numel() * 4, _lib.current_stream_ptr(ana.device))
        assert rc == 0
        torch.cuda.synchronize()
        cols = [c for c in range(63) if c != 8]
        assert np.array_equal(feat.cpu().numpy(), f0)
        assert np.array_equal(det.cpu().numpy()[:, cols], d0[:, cols], equal_nan=True)
        assert np.array_equal(mf.cpu().numpy(), m0)
    assert d0[0, 75] == 1.0 and d0[3, 75] == 0.0                       # the paused segments took the overflow path
    for i in (0, 3):
        close(d0[i, 10:23], fx.timbre(x[i]), what="timbre")
        close(d0[i, 27:31], fx.quality4(x[i]), what="quality")


def test_multichannel_per_method_semantics(ana):
    """audio_analyzer.py:190-201 with a [C, T] waveform, C = 3: intensity is the finite [1, C] z-score of the channel
    energies (golden values from the unmodified reference, oracle/make_golden.py: tests/golden/multichannel_golden.npz).
    "pitch" and timbre keep their documented defaults for C >= 2 (their multi-channel values depend on the reference's
    fp32 round-trip noise and on a top_db maximum shared by all channels, see the method docstrings)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "multichannel_golden.npz"))
    w = torch.from_numpy(np.stack([synth.pcm_to_f32(synth.segment_pcm(int(s))) for s in g["seeds"]])).to(ana.device)
    it = ana._analyze_intensity(w)
    assert tuple(it.shape) == (1, 3)
    close(it.cpu().numpy(), g["intensity"], what="intensity [1, C]")
    assert tuple(ana._analyze_pitch(w).shape) == (1, 1) and float(ana._analyze_pitch(w).abs().max()) == 0.0
    assert tuple(ana._analyze_timbre(w).shape) == (1, 13) and float(ana._analyze_timbre(w).abs().max()) == 0.0

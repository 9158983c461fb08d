"""PCM ingest + resample and the feature-row normalizers.

CPU part: the numpy oracle and the library's host-built polyphase filter bank against golden vectors produced
by torchaudio / the reference (oracle/make_golden_ingest.py).  GPU part: the kernels through the C ABI."""
import ctypes
import os
import wave

import numpy as np
import pytest
import torch

from oracle import resample_np as rs
from oracle.make_golden_ingest import RESAMPLE_CASES, ingest_signal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "ingest_golden.npz"))


@pytest.mark.parametrize("sr,n", RESAMPLE_CASES)
def test_oracle_resample_matches_torchaudio_golden(golden, sr, n):
    x = ingest_signal(sr, sr, n).astype(np.float32) / np.float32(32768.0)
    y = rs.resample(x, sr, 16000)
    ref = golden[f"resample_{sr}"]
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() < 1e-6


@pytest.mark.parametrize("sr", [48000, 44100, 8000, 22050, 32000, 11025, 96000])
def test_host_filter_bank_matches_oracle(sr):
    import __graft_entry__ as g
    g.build()
    import msa_b200  # noqa: F401
    from msa_b200 import _lib
    l = _lib.lib()
    w, t, ph = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert l.msa_resample_kernel_host(sr, 16000, None, 0, ctypes.byref(w), ctypes.byref(t), ctypes.byref(ph)) == 0
    k = np.zeros((ph.value, t.value), np.float32)
    assert l.msa_resample_kernel_host(sr, 16000, k.ctypes.data_as(ctypes.c_void_p), k.size, ctypes.byref(w), ctypes.byref(t),
                                      ctypes.byref(ph)) == 0
    ko, wo = rs.sinc_kernel(sr, 16000)
    assert w.value == wo and k.shape == ko.shape
    assert np.abs(k - ko).max() <= 1e-7                        # same float64 formula, cast to fp32
    assert l.msa_resample_out_len(12001, sr, 16000) == int(np.ceil(16000 * 12001 / sr))


def test_oracle_layernorm_rows_match_reference_golden(golden):
    for name, dims in (("face", (27, 20, 30)), ("text", (783, 700, 800)), ("audio", (31, 27, 40))):
        D = dims[0]
        for d in dims:
            x = golden[f"norm_{name}_{d}_in"].astype(np.float64)
            xp = np.zeros((x.shape[0], D))
            xp[:, :min(d, D)] = x[:, :D]
            y = (xp - xp.mean(1, keepdims=True)) / np.sqrt(xp.var(1, keepdims=True) + 1e-5)
            assert np.abs(y - golden[f"norm_{name}_{d}_out"]).max() < 2e-5


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("sr,n", RESAMPLE_CASES)
def test_gpu_resample_matches_torchaudio_golden(golden, sr, n):
    from tests.gpu_util import need_gpu
    dev = need_gpu()
    import msa_b200
    pcm = ingest_signal(sr, sr, n)
    ref = golden[f"resample_{sr}"]
    y16 = msa_b200.resample(torch.from_numpy(pcm).to(dev), sr, 16000).cpu().numpy()               # int16 PCM ingest
    y32 = msa_b200.resample(torch.from_numpy(pcm.astype(np.float32) / np.float32(32768.0)).to(dev), sr, 16000).cpu().numpy()
    assert y16.shape == ref.shape and np.array_equal(y16, y32)
    assert np.abs(y16 - ref).max() < 2e-6                        # fp32 accumulation order differs from conv1d
    # batched, ragged against the tile size, against the fp64 oracle
    xb = np.stack([ingest_signal(7 + i, sr, n).astype(np.float32) / np.float32(32768.0) for i in range(3)])
    yb = msa_b200.resample(torch.from_numpy(xb).to(dev), sr, 16000).cpu().numpy()
    assert np.abs(yb - rs.resample(xb, sr, 16000)).max() < 2e-6


@pytest.mark.gpu
def test_gpu_resample_long_and_identity():
    from tests.gpu_util import need_gpu
    dev = need_gpu()
    import msa_b200
    x = ingest_signal(3, 48000, 5 * 48000).astype(np.float32) / np.float32(32768.0)                 # a whole 5 s segment at 48 kHz
    y = msa_b200.resample(torch.from_numpy(x).to(dev), 48000, 16000).cpu().numpy()
    assert y.shape == (80000,) and np.abs(y - rs.resample(x, 48000, 16000)).max() < 2e-6
    assert torch.equal(msa_b200.resample(torch.from_numpy(x[:100]).to(dev), 16000, 16000).cpu(), torch.from_numpy(x[:100]))


@pytest.mark.gpu
def test_gpu_normalizers_match_reference_golden(golden):
    from tests.gpu_util import need_gpu
    dev = need_gpu()
    import msa_b200
    for name, cls, dims in (("face", msa_b200.FaceFeatureNormalizer, (27, 20, 30)), ("text", msa_b200.TextFeatureNormalizer, (783, 700, 800)),
                            ("audio", msa_b200.AudioFeatureNormalizer, (31, 27, 40))):
        norm = cls(device=str(dev))
        assert norm.target_dim == dims[0]
        for d in dims:
            y = norm.normalize(torch.from_numpy(golden[f"norm_{name}_{d}_in"])).cpu().numpy()
            ref = golden[f"norm_{name}_{d}_out"]
            assert y.shape == ref.shape and np.abs(y - ref).max() < 2e-5
        y1 = norm.normalize(torch.from_numpy(golden[f"norm_{name}_{dims[0]}_in"][0])).cpu().numpy()    # 1-D input -> [1, D]
        assert np.abs(y1 - golden[f"norm_{name}_1d_out"]).max() < 2e-5
    bad = torch.tensor([[1.0, float("nan"), float("inf"), -float("inf")] + [0.5] * 23])
    row = msa_b200.assemble_row([bad[:, :7], bad[:, 7:]], device=str(dev)).cpu()
    assert torch.equal(row, torch.nan_to_num(bad, nan=0.0))
    assert torch.isnan(msa_b200.FaceFeatureNormalizer(str(dev)).normalize(bad)).all()                  # NaN poisons the LayerNorm row ...
    assert torch.equal(msa_b200.FaceFeatureNormalizer(str(dev)).normalize(bad, nan_to_num=True).cpu(), torch.zeros(1, 27))   # ... scrubbed to 0


@pytest.mark.gpu
@pytest.mark.parametrize("sr,secs", [(48000, 2.0), (44100, 1.5), (8000, 2.0)])
def test_gpu_analyze_resampled_wav_matches_reference(golden, tmp_path, sr, secs):
    """AudioAnalyzer.analyze on a non-16 kHz wav: load -> Resample -> features (audio_analyzer.py:71-131)."""
    from tests.gpu_util import close, need_gpu
    dev = need_gpu()
    import msa_b200
    pcm = ingest_signal(1000 + sr, sr, int(sr * secs))
    p = os.path.join(tmp_path, "a.wav")
    with wave.open(p, "wb") as wf:
        wf.setnchannels(1); wf.setsampwidth(2); wf.setframerate(sr)
        wf.writeframes(pcm.tobytes())
    ana = msa_b200.AudioAnalyzer(device=str(dev))
    a = ana.analyze(p, "spk")
    row = torch.cat([a.emotion_probs, a.pitch, a.intensity, a.timbre, a.speech_rate, a.rhythm], dim=1).cpu().numpy()[0]
    got = np.concatenate([row, [a.audio_quality, a.signal_noise_ratio, a.clarity, a.consistency]])
    ref = golden[f"analyze_{sr}"]
    assert np.array_equal(np.isnan(got), np.isnan(ref))                                             # mono: the LayerNorm row is NaN
    close(got[27:], ref[27:], what="quality")
    w = msa_b200.resample(torch.from_numpy(pcm)[None, :].to(dev), sr, 16000)
    close(ana._analyze_timbre(w).cpu().numpy()[0], golden[f"analyze_{sr}_timbre"], what="timbre")
    close(ana._analyze_rhythm(w).cpu().numpy()[0][:2], golden[f"analyze_{sr}_rhythm"][:2], what="rhythm")

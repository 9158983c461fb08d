"""CPU emulation of the fused CUDA feature kernel vs the oracle.

tests/emu/emu_features.cpp compiles the SAME kernel body the GPU runs
(csrc/msa_features_body.cuh) with g++ and executes one cluster per segment on OS
threads.  This checks the slice / halo / ownership / overlap-add / digit-reversal
logic without a GPU; the GPU parity tests proper are tests/test_gpu_*.py.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import features_np as fx
from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "emu_features.cpp")
LIB = os.path.join(ROOT, "tests", "emu", "libemu_features.so")
CSRC = os.path.join(ROOT, "multimodal-sentiment-analyzer_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("msa_features_body.cuh", "msa_fft.cuh", "msa_tables.hpp", "msa_hd.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-I", CSRC, SRC, "-o", LIB])
    return ctypes.CDLL(LIB)


def run(lib, x, nranks=4, nwarps=4, flags=1, parts=7, emo=None, scratch=False):
    x = np.ascontiguousarray(x)
    B, T = x.shape
    feat = np.zeros((B, 31), np.float32)
    det = np.zeros((B, 96), np.float32)
    dbg = np.zeros((B, T // 200 + 1, 13), np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    ws = np.full((B, (T // 200 + 1 + 3) // 4, 16, 32), np.nan, np.float32) if scratch else None      # poisoned dB scratch table
    rc = lib.emu_features_ws(p(x), int(x.dtype == np.int16), B, T, nranks, nwarps,
                             None if emo is None else p(np.ascontiguousarray(emo)), p(feat), p(det), p(dbg), flags, parts,
                             None if ws is None else p(ws))
    assert rc == 0
    return feat, det, dbg


def residual_is_fp16_noise(det, x):
    """The STFT -> ISTFT round trip runs in fp16 on the tensor cores (msa_pitch_tc.cuh): |x - x^| (det[65:68] = mean,
    std, max) must be fp16 rounding noise relative to the signal.  One wrong index anywhere in the four matrix stages,
    the twiddles, the transposes or the overlap-add ring gives a residual of the order of the signal itself."""
    amp = float(np.abs(x).max())
    assert det[67] <= 6e-3 * amp + 2e-6, (det[65:68], amp)
    assert det[66] <= 2e-3 * amp + 1e-6, (det[65:68], amp)


def check_against_oracle(det, feat, x, emo=None, rel=1e-3, floor=1e-6):
    """Tolerances of SURVEY.md section 8(d): rel 1e-3 (abs floor), pitch abs 1e-6, exact flags."""
    raw = fx.raw_features(x, emo)
    q = fx.quality4(x)
    got = det[:27].astype(np.float64)
    assert abs(got[8]) <= 1e-6 and abs(raw[8]) <= 1e-6                      # "pitch" is rounding noise
    assert np.isnan(got[9]) and np.isnan(raw[9])                            # mono intensity is NaN
    for sl in (slice(0, 8), slice(10, 23), slice(24, 26)):
        a, b = got[sl], raw[sl]
        assert np.array_equal(np.isnan(a), np.isnan(b))
        ok = ~np.isnan(b)
        assert np.all(np.abs(a[ok] - b[ok]) <= rel * np.abs(b[ok]) + floor), (a, b)
    assert got[23] == raw[23]                                               # speech_rate exact
    assert np.float32(got[26]) == np.float32(raw[26])                       # L/16000 exact
    gq = det[27:31].astype(np.float64)
    assert np.array_equal(np.isnan(gq), np.isnan(q))
    ok = ~np.isnan(q)
    assert np.all(np.abs(gq[ok] - q[ok]) <= rel * np.abs(q[ok]) + floor), (gq, q)
    row = fx.audio_row31(x, emo)
    assert np.all(np.abs(feat - row) <= rel * np.abs(row) + floor)


@pytest.mark.parametrize("nranks,nwarps", [(1, 8), (2, 4), (4, 8), (8, 2), (16, 8), (1, 16)])
def test_seeded_segment_all_cluster_sizes(emu, nranks, nwarps):
    x = synth.pcm_to_f32(synth.segment_pcm(1234))
    feat, det, dbg = run(emu, x[None], nranks, nwarps)
    check_against_oracle(det[0], feat[0], x)
    ref = fx.mfcc(x.astype(np.float64)).T
    assert np.abs(dbg[0] - ref).max() < 1e-3                                # MFCC matrix itself (values up to ~170)
    residual_is_fp16_noise(det[0], x)


def test_int16_ingest_matches_f32(emu):
    pcm = synth.segments_pcm(2000, 2)
    f16, d16, _ = run(emu, pcm, 4, 4)
    f32, d32, _ = run(emu, synth.pcm_to_f32(pcm), 4, 4)
    assert np.array_equal(f16, f32)
    assert np.array_equal(d16[:, :31], d32[:, :31], equal_nan=True)


def test_emotion_embedding_and_finite_ln(emu):
    x = synth.pcm_to_f32(synth.segment_pcm(1240))
    emo = synth.emotion_probs(5, 1)
    feat, det, _ = run(emu, x[None], 4, 4, flags=0, emo=emo)                # flags=0: intensity 0 instead of NaN
    raw = fx.raw_features(x, emo[0])
    raw[9] = 0.0
    ln = fx.ln31(raw)
    assert np.allclose(det[0, 32:63], ln, rtol=1e-3, atol=1e-5)
    row = fx.audio_row31(x, emo[0], finite_intensity=True)
    assert np.allclose(feat[0], row, rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("name", ["white_0p1", "zeros", "noise_1e-4", "tone_220", "half_silence",
                                  "short_8000", "odd_12345", "short_1700", "short_500", "long_10s"])
def test_adversarial(emu, name):
    x = synth.adversarial_cases()[name]
    T = x.size
    nranks = 8 if T > 100000 else (4 if T > 20000 else (2 if T > 4000 else 1))
    feat, det, _ = run(emu, x[None], nranks, 4)
    # near-silent input: mel bins sit at / next to the 1e-10 floor, where fp32 FFT noise decides the dB value
    rel, floor = (1e-3, 1e-6) if name not in ("noise_1e-4", "zeros") else (2e-3, 2e-4)
    check_against_oracle(det[0], feat[0], x, rel=rel, floor=floor)


def test_batch_rows_are_independent(emu):
    pcm = synth.segments_pcm(3000, 3)
    fb, db, _ = run(emu, pcm, 4, 2)
    for i in range(3):
        f1, d1, _ = run(emu, pcm[i:i + 1], 4, 2)
        assert np.array_equal(fb[i], f1[0])


@pytest.mark.parametrize("T,nranks", [(30001, 2), (30001, 1), (257, 1), (513, 1), (1025, 2), (80127, 4), (80129, 1)])
def test_every_sample_is_reconstructed_once(emu, T, nranks):
    """det[72] counts the samples of the STFT -> ISTFT residual: all T of them, whatever T mod 128 is."""
    x = synth.pcm_to_f32(synth.segment_pcm(77, T))
    feat, det, _ = run(emu, x[None], nranks, 4)
    assert det[0, 72] == T
    residual_is_fp16_noise(det[0], x)
    check_against_oracle(det[0], feat[0], x)


@pytest.mark.parametrize("name,fix,slow", [("seg1234", 1, 0), ("white_0p1", 0, 0), ("tone_220", 1, 1), ("half_silence", 1, 1)])
def test_top_db_clamp_paths(emu, name, fix, slow):
    """amplitude_to_DB(top_db=80): values 80 dB below the segment maximum are clamped.  The kernel patches
    the few affected frames from per-warp candidate lists (det[78]) and only redoes the pass when a list
    overflows (det[75]); both paths and the no-clamp path must give the reference MFCC."""
    x = synth.pcm_to_f32(synth.segment_pcm(1234)) if name == "seg1234" else synth.adversarial_cases()[name]
    for nranks, nwarps in ((1, 8), (2, 4)):
        _, det, dbg = run(emu, x[None], nranks, nwarps)
        assert (det[0, 78], det[0, 75]) == (fix, slow)
        assert det[0, 64] - 80.0 <= det[0, 77]                               # candidate level never below the real threshold
        ref = fx.mfcc(x.astype(np.float64)).T
        assert np.abs(dbg[0] - ref).max() < 1e-3


@pytest.mark.parametrize("name", ["seg1234", "tone_220", "half_silence", "white_0p1", "path_flip"])
def test_results_do_not_depend_on_the_partition(emu, name):
    """The same segment over different cluster sizes / warps per CTA takes different top_db paths (patch list vs
    clamped pass, decided per CTA by list overflow) and different reduction trees, and must still give the SAME
    BITS: sharding a batch differently (or streaming one segment over 8 CTAs) may not change a result."""
    if name == "path_flip":
        # a 180 Hz tone over a low noise floor: ~600 top_db candidates, so 8 warps patch them from their lists
        # while 2 warps overflow and redo the pass clamped (asserted below)
        rng = np.random.default_rng(5)
        t = np.arange(80000) / 16000.0
        x = 0.3 * np.sin(2 * np.pi * 180 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 3 * t)) + 0.004 * rng.standard_normal(80000)
        x = np.round(np.clip(x, -1, 1) * 32767).astype(np.int16).astype(np.float32) / np.float32(32768)
    else:
        x = synth.pcm_to_f32(synth.segment_pcm(1234)) if name == "seg1234" else synth.adversarial_cases()[name]
    ref = None
    paths = set()
    for nranks, nwarps, scratch in ((1, 8, False), (2, 8, False), (8, 8, True), (16, 8, False), (1, 2, False), (4, 3, True), (1, 2, True), (1, 8, True)):
        feat, det, dbg = run(emu, x[None], nranks, nwarps, scratch=scratch)
        paths.add((det[0, 78], det[0, 75]))
        cur = (feat.copy(), det[:, :63].copy(), dbg.copy())
        if ref is None:
            ref = cur
        else:
            cols = [c for c in range(63) if c != 8]                 # col 8 ("pitch") is rounding residue: order dependent
            assert np.array_equal(cur[0], ref[0])
            assert np.array_equal(cur[1][:, cols], ref[1][:, cols], equal_nan=True), (nranks, nwarps)
            assert np.array_equal(cur[2], ref[2]), (nranks, nwarps)
    if name == "path_flip":
        assert (1.0, 0.0) in paths and (1.0, 1.0) in paths           # both the patch-list and the clamped-pass path ran

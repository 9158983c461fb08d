"""Helpers shared by the -m gpu parity tests (they all go through the C ABI via msa_b200)."""
import numpy as np
import pytest
import torch


def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def close(a, b, rel=1e-3, floor=1e-6, what=""):
    """Tolerance of SURVEY.md section 8(d): |a-b| <= rel*|b| + floor, identical NaN pattern."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.array_equal(np.isnan(a), np.isnan(b)), (what, a, b)
    ok = ~np.isnan(b)
    bad = np.abs(a[ok] - b[ok]) > rel * np.abs(b[ok]) + floor
    assert not bad.any(), (what, a[ok][bad][:8], b[ok][bad][:8])


def residual_is_fp16_noise(det, amp):
    """The STFT -> ISTFT round trip of the "pitch" feature runs in fp16 on the tensor cores (csrc/msa_pitch_tc.cuh):
    detail[65:68] = mean, std, max of |x - x^| must be fp16 rounding noise relative to the signal amplitude `amp`.
    A wrong index anywhere in the four matrix stages, twiddles, transposes or the overlap-add ring gives a residual of
    the order of the signal.  (The feature itself is the mean of the z-scored residual: |v| <= 1e-6 either way.)"""
    det = np.asarray(det, dtype=np.float64)
    assert det[67] <= 6e-3 * amp + 2e-6, (det[65:68], amp)
    assert det[66] <= 2e-3 * amp + 1e-6, (det[65:68], amp)

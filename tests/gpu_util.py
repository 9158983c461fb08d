"""Helpers shared by the -m gpu parity tests (they all go through the C ABI via msa_b200)."""
import numpy as np
import pytest
import torch


def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def close(a, b, rel=1e-3, floor=1e-5, what=""):
    """Tolerance of SURVEY.md section 8(d): |a-b| <= rel*|b| + floor, identical NaN pattern."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.array_equal(np.isnan(a), np.isnan(b)), (what, a, b)
    ok = ~np.isnan(b)
    bad = np.abs(a[ok] - b[ok]) > rel * np.abs(b[ok]) + floor
    assert not bad.any(), (what, a[ok][bad][:8], b[ok][bad][:8])

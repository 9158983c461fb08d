"""GPU check of the opt-in FOLD variant of the feature kernel (MSA_FEAT_FOLD_WAVE) against the default kernel, through
the C ABI.  Kept in a file of its own that sorts after the parity tests proper: the variant is off by default."""
import numpy as np
import pytest

from oracle import features_np as fx
from oracle import synth
from tests.gpu_util import close
from tests.test_gpu_features import _detail, ana  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,cluster,dtype", [(80000, 0, np.float32), (80000, 2, np.int16), (80000, 8, np.float32), (12345, 1, np.float32),
                                             (80129, 1, np.int16), (30001, 2, np.float32), (513, 1, np.float32), (257, 1, np.float32)])
def test_fold_wave_statistics_bit_identical(ana, T, cluster, dtype):
    """MSA_FEAT_FOLD_WAVE: the STFT-512 quads form the wave statistics from the samples they hold (no separate pass
    over the segment).  Same entries, same atoms, same order: rows and raw features equal the default kernel's bit for
    bit (also checked without a GPU on the emulator, tests/test_emu_features.py, and by scripts/fold_check.cu)."""
    pcm = np.stack([synth.segment_pcm(4000 + i, T) for i in range(6)])
    pcm[2, T // 3: T // 2] = 0                                   # a stretch of digital silence: top_db patch / overflow paths
    x = pcm if dtype == np.int16 else synth.pcm_to_f32(pcm)
    f0, d0, m0 = _detail(ana, x, cluster=cluster, flags=1)
    f1, d1, m1 = _detail(ana, x, cluster=cluster, flags=1 | 4)
    assert np.all(d0[:, 79] == 0.0) and np.all(d1[:, 79] == 1.0)
    assert np.array_equal(f0, f1)
    assert np.array_equal(d0[:, :79], d1[:, :79], equal_nan=True)
    assert np.array_equal(m0, m1)
    raw, q = fx.raw_features(synth.pcm_to_f32(pcm[0])), fx.quality4(synth.pcm_to_f32(pcm[0]))
    close(d1[0, 24:27], raw[24:27], what="rhythm (fold)")
    close(d1[0, 27:31], q, what="quality (fold)")

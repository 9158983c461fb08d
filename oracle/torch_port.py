"""torch / torchaudio port of the reference's CPU path — the TIMED CPU BASELINE.

TEST INFRASTRUCTURE (see oracle/__init__.py).  /root/reference is pure Python and cannot travel
to the GPU box, so ``bench.py`` times this port there (``cpu_baseline.kind = "port"``).  It issues
the same library calls in the same order as /root/reference/src/analyzers/audio_analyzer.py and
src/models/fusion_model.py — including the reference's habit of constructing a new
``torchaudio.transforms.MFCC`` inside every timbre / clarity call (:207-210, :299-302) and of
computing clarity twice per segment (:130 and inside :270) — so its cost profile is the
reference's.  tests/test_oracle_golden.py pins it against the golden vectors.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch
import torchaudio

SR = 16000


class PortedAnalyzer:
    """Feature body of AudioAnalyzer.analyze (audio_analyzer.py:84-131) with the wav2vec2 call stubbed
    to the reference's own fallback (uniform 1/8)."""

    def __init__(self):
        self.pitch_model = torchaudio.transforms.PitchShift(sample_rate=SR, n_steps=0)      # :43-47
        self.layer_norm = torch.nn.LayerNorm(31)                                            # normalization.py:13

    def pitch(self, w):                                                                     # :175-188
        p = torch.abs(w - self.pitch_model(w))
        p = (p - p.mean()) / (p.std() + 1e-6)
        return p.mean(dim=1).unsqueeze(0)

    def intensity(self, w):                                                                 # :190-201
        e = torch.sum(w ** 2, dim=1)
        return ((e - e.mean()) / (e.std() + 1e-6)).unsqueeze(0)

    def _mfcc(self, w):
        return torchaudio.transforms.MFCC(sample_rate=SR, n_mfcc=13)(w)

    def timbre(self, w):                                                                    # :203-217
        m = self._mfcc(w)
        m = (m - m.mean()) / (m.std() + 1e-6)
        return m.mean(dim=2).squeeze().unsqueeze(0)

    def speech_rate(self, w):                                                               # :219-233
        e = torch.sum(w ** 2, dim=1)
        s = (e > e.mean() * 0.1).float()
        return (torch.sum(s) / len(s)).unsqueeze(0).unsqueeze(0)

    def rhythm(self, w):                                                                    # :235-263
        e = torch.nn.functional.unfold(w.unsqueeze(0).unsqueeze(0), kernel_size=(1, 400), stride=(1, 160))
        e = torch.sum(e ** 2, dim=1)
        return torch.cat([e.mean(dim=1), e.std(dim=1), torch.tensor([len(e[0]) / SR])]).unsqueeze(0)

    def snr(self, w):                                                                       # :278-293
        n = int(0.05 * w.shape[1])
        noise = torch.cat([w[:, :n], w[:, -n:]])
        v = 10 * torch.log10(torch.mean(w ** 2) / (torch.mean(noise ** 2) + 1e-6))
        return min(max(v.item() / 30, 0), 1)

    def clarity(self, w):                                                                   # :295-311
        m = self._mfcc(w)
        v = torch.mean(torch.abs(m[:, 6:])) / (torch.mean(torch.abs(m[:, :6])) + 1e-6)
        return min(max(v.item(), 0), 1)

    def consistency(self, w):                                                               # :313-329
        e = torch.mean(w.unfold(1, 1600, 1600) ** 2, dim=2)
        return 1.0 - min((torch.std(e) / (torch.mean(e) + 1e-6)).item(), 1.0)

    def audio_quality(self, w):                                                             # :265-276
        return 0.4 * self.snr(w) + 0.3 * self.clarity(w) + 0.3 * self.consistency(w)

    @torch.no_grad()
    def audio_row(self, w: torch.Tensor) -> torch.Tensor:
        """[1, T] -> [1, 31]: the per-segment work of analyze() plus the row assembly and nan_to_num
        of streaming_processor.py:250-268, 295-298."""
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            feats = torch.cat([torch.full((1, 8), 0.125), self.pitch(w), self.intensity(w), self.timbre(w),
                               self.speech_rate(w), self.rhythm(w)], dim=1)
            feats = self.layer_norm(torch.cat([feats, torch.zeros(1, 4)], dim=1))
            q = torch.tensor([[self.audio_quality(w), self.snr(w), self.clarity(w), self.consistency(w)]])
        return torch.nan_to_num(torch.cat([feats[:, :27], q], dim=1), nan=0.0)


def build_fusion(sd: dict) -> dict:
    return {k: torch.from_numpy(np.asarray(v)).float() for k, v in sd.items()}


@torch.no_grad()
def fusion_forward(sd: dict, face, audio, text=None) -> torch.Tensor:
    """_fuse_all / _fuse_face_audio in eval mode with torch ops (fusion_model.py:296-321, 386-408)."""
    F = torch.nn.functional

    def ln(x, name):
        return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)

    def lin(x, name):
        return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])

    def branch(x, m):
        h = lin(ln(x, m + "_norm"), m + "_proj")
        h = lin(torch.relu(ln(h, m + "_processor.0")), m + "_processor.3")
        return torch.relu(ln(h, m + "_processor.4"))

    parts = [branch(face, "face"), branch(audio, "audio")] + ([branch(text, "text")] if text is not None else [])
    h = lin(torch.cat(parts, dim=-1), "fusion.0" if text is not None else "fusion2")
    h = lin(torch.relu(ln(h, "fusion.1")), "fusion.4")
    return lin(torch.relu(ln(h, "fusion.5")), "fusion.8")


def _worker(args):
    """Pool task: synthesise `count` segments (untimed), then time the reference port over them.
    Returns (rows [count, 31], seconds of compute)."""
    import time
    seed, count = args
    from oracle import synth
    torch.set_num_threads(1)
    ana = PortedAnalyzer()
    waves = [torch.from_numpy(synth.pcm_to_f32(synth.segment_pcm(seed + i)))[None, :] for i in range(count)]
    ana.audio_row(waves[0])                              # first-call costs (window / filterbank caches) are not timed
    t0 = time.perf_counter()
    rows = [ana.audio_row(w) for w in waves]
    dt = time.perf_counter() - t0
    return torch.cat(rows).numpy(), dt

"""Streaming window: a device-resident PCM ring buffer feeding the same two kernels.

The reference's StreamingProcessor cuts TUMBLING windows of `duration` seconds
(/root/reference/src/processors/streaming_processor.py:411-443); BASELINE.json's streaming
configuration asks for a 5 s window advanced by a 0.5 s hop.  Reflect padding, the top_db clamp and
the z-scores are all window-global, so nothing can be carried over between hops: each hop uploads
only its 8000 new int16 samples (16 KB) and recomputes the whole 80000-sample window on device.
"""
from __future__ import annotations

from typing import Optional

import torch

from .audio_analyzer import AudioAnalyzer
from .fusion_model import AdvancedFusionModel


class StreamingWindow:
    """Mirror-ring: the ring holds 2 x window samples and every chunk is written at `pos` and at
    `pos + window`, so the most recent `window` samples are always one CONTIGUOUS slice
    ring[pos : pos + window] (pos = next write position) and no on-device slide is needed.  Per hop: one 16 KB
    host->device copy from a rotating pinned staging buffer (an event guards its reuse), one 16 KB
    device copy for the mirror, the feature kernel and the fusion chain."""

    N_STAGE = 4

    def __init__(self, analyzer: AudioAnalyzer, fusion: AdvancedFusionModel, window: int = 80000, hop: int = 8000):
        if window % hop:
            raise ValueError("window must be a multiple of hop")
        self.analyzer, self.fusion = analyzer, fusion
        self.window, self.hop = window, hop
        self.device = analyzer.device
        self.ring = torch.zeros(2 * window, dtype=torch.int16, device=self.device)
        self.stages = [torch.empty(hop, dtype=torch.int16).pin_memory() for _ in range(self.N_STAGE)]
        self.events = [None] * self.N_STAGE
        self.n_pushed = 0
        self.pos = 0            # where the next chunk goes, in [0, window)

    @property
    def filled(self) -> int:
        return min(self.window, self.n_pushed * self.hop)

    @torch.no_grad()
    def push(self, chunk_pcm: torch.Tensor, face: torch.Tensor, text: Optional[torch.Tensor] = None):
        """chunk_pcm: [hop] int16 on the host.  Returns None until the window is full, then the dict of
        streaming_processor.py:302-320: {"fused_emotion": logits [7], "argmax": int, "audio_row": [31]}."""
        k = self.n_pushed % self.N_STAGE
        if self.events[k] is not None:
            self.events[k].synchronize()            # the H2D copy that last read this staging buffer is done
        stage = self.stages[k]
        stage.copy_(chunk_pcm.reshape(-1))
        p = self.pos
        self.ring[p:p + self.hop].copy_(stage, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.events[k] = ev
        self.ring[p + self.window:p + self.window + self.hop].copy_(self.ring[p:p + self.hop])
        self.pos = (p + self.hop) % self.window
        self.n_pushed += 1
        if self.n_pushed * self.hop < self.window:
            return None
        win = self.ring[self.pos:self.pos + self.window]      # oldest sample of the current window sits at pos
        row = self.analyzer.analyze_batch(win[None, :])
        logits, amax = self.fusion.fused_with_argmax(face.reshape(1, -1), row, None if text is None else text.reshape(1, -1))
        return {"fused_emotion": logits[0], "argmax": amax[0], "audio_row": row[0]}

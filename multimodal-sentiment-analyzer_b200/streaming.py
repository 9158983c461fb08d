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
    def __init__(self, analyzer: AudioAnalyzer, fusion: AdvancedFusionModel, window: int = 80000, hop: int = 8000):
        if window % hop:
            raise ValueError("window must be a multiple of hop")
        self.analyzer, self.fusion = analyzer, fusion
        self.window, self.hop = window, hop
        self.device = analyzer.device
        self.ring = torch.zeros(window, dtype=torch.int16, device=self.device)       # logical order: oldest first
        self.stage = torch.empty(hop, dtype=torch.int16).pin_memory()
        self.filled = 0

    @torch.no_grad()
    def push(self, chunk_pcm: torch.Tensor, face: torch.Tensor, text: Optional[torch.Tensor] = None):
        """chunk_pcm: [hop] int16 on the host.  Returns None until the window is full, then the dict of
        streaming_processor.py:302-320: {"fused_emotion": logits [7], "argmax": int, "audio_row": [31]}."""
        self.stage.copy_(chunk_pcm.reshape(-1))
        # slide: drop the oldest hop, append the new one (a 144 KB on-device move + 16 KB H2D)
        self.ring[: self.window - self.hop] = self.ring[self.hop:].clone()
        self.ring[self.window - self.hop:].copy_(self.stage, non_blocking=True)
        self.filled = min(self.window, self.filled + self.hop)
        if self.filled < self.window:
            return None
        row = self.analyzer.analyze_batch(self.ring[None, :])
        logits, amax = self.fusion.fused_with_argmax(face.reshape(1, -1), row, None if text is None else text.reshape(1, -1))
        return {"fused_emotion": logits[0], "argmax": amax[0], "audio_row": row[0]}

"""Result containers of the hot path, mirroring /root/reference/src/structures/analysis.py:14-56
(``DictMixin`` and ``AudioAnalysis``: same field names, order and dict-style access)."""
from __future__ import annotations

from dataclasses import asdict, dataclass
from typing import Any, Dict

import torch


class DictMixin:
    """obj["field"], obj.get("field", default) and obj.to_dict() (analysis.py:14-24)."""

    def __getitem__(self, key: str) -> Any:
        return getattr(self, key)

    def get(self, key: str, default: Any = None) -> Any:
        return getattr(self, key, default)

    def to_dict(self) -> Dict:
        return asdict(self)


@dataclass
class AudioAnalysis(DictMixin):
    """analysis.py:42-56.  Tensors are [1, n] on the analyzer's device and DETACHED (the reference
    returns non-leaf tensors of a throw-away LayerNorm, on which its own to_dict() raises —
    SURVEY.md section 2.4; callers always .detach() first, so this is a benign divergence)."""
    speaker_id: str
    emotion_probs: torch.Tensor   # [1, 8]
    pitch: torch.Tensor           # [1, 1]
    intensity: torch.Tensor       # [1, 1]
    timbre: torch.Tensor          # [1, 13]
    speech_rate: torch.Tensor     # [1, 1]
    rhythm: torch.Tensor          # [1, 3]
    audio_quality: float
    signal_noise_ratio: float
    clarity: float
    consistency: float

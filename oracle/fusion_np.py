"""numpy restatement of AdvancedFusionModel.forward in eval mode.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
/root/reference/src/models/fusion_model.py:131-190 (dispatch and fallbacks),
:296-328 (_fuse_face_audio), :330-384 (the two always-failing pairs) and
:386-417 (_fuse_all).  Dropout is the identity (eval mode) — the reference never
calls .eval() on its inference path; training-mode dropout is
not reproducible, so eval mode is the stated oracle.

``sd`` is a state dict of numpy arrays with the reference's parameter names
(SURVEY.md appendix A); ``compute`` selects float64 (default) or float32 math.
"""
from __future__ import annotations

import numpy as np


def _ln(x, g, b, eps=1e-5):
    m = x.mean(axis=-1, keepdims=True)
    v = ((x - m) ** 2).mean(axis=-1, keepdims=True)
    return (x - m) / np.sqrt(v + eps) * g + b


def _lin(x, sd, name, dt):
    return x @ sd[name + ".weight"].astype(dt).T + sd[name + ".bias"].astype(dt)


def _norm(x, sd, name, dt):
    return _ln(x, sd[name + ".weight"].astype(dt), sd[name + ".bias"].astype(dt))


def _branch(x, sd, mod, dt):
    """<mod>_norm -> <mod>_proj -> <mod>_processor (LN, ReLU, Drop, Linear, LN, ReLU, Drop)."""
    h = _lin(_norm(x.astype(dt), sd, mod + "_norm", dt), sd, mod + "_proj", dt)
    h = np.maximum(_norm(h, sd, mod + "_processor.0", dt), 0)
    h = _lin(h, sd, mod + "_processor.3", dt)
    return np.maximum(_norm(h, sd, mod + "_processor.4", dt), 0)


def _fusion_tail(h, sd, dt):
    """fusion[1:]: LN, ReLU, Drop, Linear(1024->512), LN, ReLU, Drop, Linear(512->7)."""
    h = np.maximum(_norm(h, sd, "fusion.1", dt), 0)
    h = _lin(h, sd, "fusion.4", dt)
    h = np.maximum(_norm(h, sd, "fusion.5", dt), 0)
    return _lin(h, sd, "fusion.8", dt)


def fuse_all(sd, face, audio, text, compute=np.float64):
    """fusion_model.py:386-408 -> logits [B, 7]."""
    cat = np.concatenate([_branch(face, sd, "face", compute), _branch(audio, sd, "audio", compute),
                          _branch(text, sd, "text", compute)], axis=-1)
    return _fusion_tail(_lin(cat, sd, "fusion.0", compute), sd, compute)


def fuse_face_audio(sd, face, audio, compute=np.float64):
    """fusion_model.py:296-321 -> logits [B, 7] through fusion2 then fusion[1:]."""
    cat = np.concatenate([_branch(face, sd, "face", compute), _branch(audio, sd, "audio", compute)], axis=-1)
    return _fusion_tail(_lin(cat, sd, "fusion2", compute), sd, compute)


def forward(sd, face=None, audio=None, text=None, compute=np.float64) -> dict:
    """fusion_model.py:131-190: the dict the reference returns, including its
    fallbacks: one modality passes through; face+text and audio+text feed a
    1024-wide concat to Linear(1536, 1024), raise, and fall back to a one-key
    dict; a wrong feature width raises inside and falls back to the most
    reliable single modality (face, then audio, then text)."""
    present = [(n, t) for n, t in (("face", face), ("audio", audio), ("text", text)) if t is not None]
    if not present:
        raise ValueError("no modality")
    if len(present) == 1:
        return {present[0][0]: present[0][1]}

    def fallback():
        n, t = present[0]          # order above is already face, audio, text
        return {n: t}

    dims = {"face": sd["face_norm.weight"].shape[0], "audio": sd["audio_norm.weight"].shape[0],
            "text": sd["text_norm.weight"].shape[0]}
    if any(t.ndim < 1 or t.shape[-1] != dims[n] for n, t in present):
        return fallback()
    if len(present) == 2:
        names = {n for n, _ in present}
        if names == {"face", "audio"}:
            return {"face": face, "audio": audio, "fused": fuse_face_audio(sd, face, audio, compute)}
        return fallback()
    return {"face": face, "audio": audio, "text": text, "fused": fuse_all(sd, face, audio, text, compute)}


def get_weights(sd) -> dict:
    """fusion_model.py:192-203: softmax over (audio, text, face) scalars, fp32."""
    w = np.array([sd["audio_weight"], sd["text_weight"], sd["face_weight"]], dtype=np.float32)
    e = np.exp(w - w.max())
    p = e / e.sum()
    return {"audio": float(p[0]), "text": float(p[1]), "face": float(p[2])}

"""``AdvancedFusionModel`` with the reference's interface, forward on sm_100a tensor cores.

Mirrors /root/reference/src/models/fusion_model.py: constructor arguments and module names
(:17-112, so ``state_dict()`` has the reference's 45 keys and its checkpoints load unchanged),
Xavier-uniform / zero-bias initialisation (:114-120), ``forward`` with the reference's modality
dispatch and fallbacks (:131-190), ``get_weights`` (:192-203), ``save`` / ``load`` (:239-294) and
the ``FusionModel`` alias (:420).

The nn.Module only OWNS the parameters; the arithmetic of ``_fuse_all`` (:386-408) and
``_fuse_face_audio`` (:296-321) runs in ``msa_fusion_forward`` (csrc/msa_fusion_tc.cu, tcgen05
split-bf16 GEMMs with fused bias + LayerNorm + ReLU epilogues) through the C ABI.  Inference is
eval-mode: the reference never calls .eval() and so runs Dropout(0.3) at inference (SURVEY.md
section 2.4); that stochastic behaviour is deliberately not reproduced.
"""
from __future__ import annotations

import ctypes
import logging
from pathlib import Path
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib

logger = logging.getLogger(__name__)


def _processor(hidden_dim: int, dropout: float) -> nn.Sequential:
    # indices 0 / 3 / 4 carry parameters, exactly like the reference's Sequential (:54-62)
    return nn.Sequential(nn.LayerNorm(hidden_dim), nn.ReLU(), nn.Dropout(dropout),
                         nn.Linear(hidden_dim, hidden_dim // 2), nn.LayerNorm(hidden_dim // 2), nn.ReLU(),
                         nn.Dropout(dropout))


class AdvancedFusionModel(nn.Module):
    def __init__(self, face_dim: int = 27, audio_dim: int = 31, text_dim: int = 783, hidden_dim: int = 1024,
                 output_dim: int = 7, dropout: float = 0.3, device: Optional[str] = "cuda"):
        super().__init__()
        if (face_dim, audio_dim, text_dim, hidden_dim, output_dim) != (27, 31, 783, 1024, 7):
            raise ValueError("the sm_100a kernels are specialised for the reference's dimensions 27/31/783/1024/7")
        self.device = _lib.normalize_device(device)
        self.dropout = dropout
        self.audio_dim, self.text_dim, self.face_dim = audio_dim, text_dim, face_dim
        self.hidden_dim, self.output_dim = hidden_dim, output_dim

        self.audio_norm = nn.LayerNorm(audio_dim)
        self.text_norm = nn.LayerNorm(text_dim)
        self.face_norm = nn.LayerNorm(face_dim)
        self.audio_proj = nn.Linear(audio_dim, hidden_dim)
        self.text_proj = nn.Linear(text_dim, hidden_dim)
        self.face_proj = nn.Linear(face_dim, hidden_dim)
        self.audio_processor = _processor(hidden_dim, dropout)
        self.text_processor = _processor(hidden_dim, dropout)
        self.face_processor = _processor(hidden_dim, dropout)
        self.fusion = nn.Sequential(
            nn.Linear((hidden_dim // 2) * 3, hidden_dim), nn.LayerNorm(hidden_dim), nn.ReLU(), nn.Dropout(dropout),
            nn.Linear(hidden_dim, hidden_dim // 2), nn.LayerNorm(hidden_dim // 2), nn.ReLU(), nn.Dropout(dropout),
            nn.Linear(hidden_dim // 2, output_dim))
        self.fusion2 = nn.Linear((hidden_dim // 2) * 2, hidden_dim)
        self.audio_weight = nn.Parameter(torch.tensor(0.3))
        self.text_weight = nn.Parameter(torch.tensor(0.3))
        self.face_weight = nn.Parameter(torch.tensor(0.4))
        self.softmax = nn.Softmax(dim=0)
        for m in self.modules():                      # fusion_model.py:114-120
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)
        self.eval()
        self.to(self.device)
        self._packed: Optional[torch.Tensor] = None
        self._packed_key = None
        self._workspace: Optional[torch.Tensor] = None
        # bumped whenever _packed or _workspace is (re)allocated: captured CUDA graphs hold their raw addresses
        # (StreamingWindow drops its graphs when this changes)
        self.buffers_generation = 0
        # load_state_dict invalidates the packed copy at once (the streaming path does not re-derive the parameter
        # key on every hop; the batch path does)
        self.register_load_state_dict_post_hook(lambda module, incompatible: setattr(module, "_packed_key", None))

    # ------------------------------------------------------------------ device weights
    def _param_key(self):
        # (address, autograd version) of every parameter: catches load_state_dict, optimiser steps and in-place ops.
        # Writes through ``p.data`` do not bump the version: call repack() after those.
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def repack(self) -> None:
        """Repack the current parameters into the kernels' layout (fp32 copies + split-bf16 K-major
        GEMM operands).  Called lazily by forward whenever a parameter changed."""
        l = _lib.lib()
        dev = _lib.require_cuda(self.device)
        sd = self.state_dict()
        n = l.msa_fusion_num_tensors()
        host = []
        arr = (ctypes.c_void_p * n)()
        for i in range(n):
            name = l.msa_fusion_tensor_name(i).decode()
            t = sd[name].detach().to("cpu", torch.float32).contiguous()
            if t.numel() != l.msa_fusion_tensor_numel(i):
                raise _lib.MsaError(f"state_dict tensor {name} has {t.numel()} elements, expected {l.msa_fusion_tensor_numel(i)}")
            host.append(t)
            arr[i] = t.data_ptr()
        if self._packed is None or self._packed.device != dev:
            self._packed = torch.empty(l.msa_fusion_packed_bytes(), dtype=torch.uint8, device=dev)
            self.buffers_generation += 1
        # repacked IN PLACE: the blob keeps its address, so graphs captured earlier read the new weights
        with _lib.on_device(dev):
            _lib.check(l.msa_fusion_pack(arr, _lib.ptr(self._packed), _lib.current_stream_ptr(dev)), "msa_fusion_pack")
        self._packed_key = self._param_key()

    def _ensure_packed(self):
        if self._packed is None or self._packed_key != self._param_key():
            self.repack()

    def _ws(self, B: int, dev) -> torch.Tensor:
        need = _lib.lib().msa_fusion_workspace_bytes(B)
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != dev:
            self._workspace = None                                 # release the old block before taking the larger one
            self._workspace = torch.empty(need, dtype=torch.uint8, device=dev)
            self.buffers_generation += 1
        return self._workspace

    # ------------------------------------------------------------------ fused paths
    def _device_forward(self, face: torch.Tensor, audio: torch.Tensor, text: Optional[torch.Tensor], want_argmax=False):
        dev = _lib.require_cuda(self.device)
        if face.shape[-1] != self.face_dim or audio.shape[-1] != self.audio_dim or (text is not None and text.shape[-1] != self.text_dim):
            raise ValueError("feature width does not match the model (LayerNorm would raise in the reference)")
        lead = face.shape[:-1]
        prep = lambda t: t.detach().to(dev, torch.float32).reshape(-1, t.shape[-1]).contiguous()
        f, a = prep(face), prep(audio)
        t = prep(text) if text is not None else None
        B = f.shape[0]
        if a.shape[0] != B or (t is not None and t.shape[0] != B):
            raise ValueError("batch sizes differ")
        self._ensure_packed()
        logits = torch.empty(B, self.output_dim, device=dev, dtype=torch.float32)
        amax = torch.empty(B, device=dev, dtype=torch.int32) if want_argmax else None
        ws = self._ws(B, dev)
        with _lib.on_device(dev):
            rc = _lib.lib().msa_fusion_forward(_lib.ptr(f), _lib.ptr(a), _lib.ptr(t), B, _lib.ptr(self._packed), _lib.ptr(ws),
                                               ws.numel(), _lib.ptr(logits), _lib.ptr(amax), _lib.current_stream_ptr(dev))
        _lib.check(rc, "msa_fusion_forward")
        logits = logits.reshape(*lead, self.output_dim)
        return (logits, amax) if want_argmax else logits

    def forward_into(self, face: torch.Tensor, audio: torch.Tensor, text: Optional[torch.Tensor], logits_out: torch.Tensor,
                     argmax_out: Optional[torch.Tensor]) -> None:
        """Allocation-free forward for CUDA-graph capture and tight loops: contiguous fp32 device inputs
        [B, 27] / [B, 31] / [B, 783] or None, outputs written into caller-owned [B, 7] fp32 and [B] int32.
        Weights must already be packed and the workspace sized (``prepare(B)`` once outside the capture): inside a
        capture nothing may be allocated, so a missing or too small buffer is an error here."""
        dev = self.device
        B = face.shape[0]
        need = _lib.lib().msa_fusion_workspace_bytes(B)
        if self._packed is None or self._workspace is None or self._workspace.numel() < need:
            raise _lib.MsaError("forward_into: call prepare(B) first (weights packed, workspace sized)")
        ws = self._workspace
        rc = _lib.lib().msa_fusion_forward(_lib.ptr(face), _lib.ptr(audio), _lib.ptr(text), B, _lib.ptr(self._packed), _lib.ptr(ws),
                                           ws.numel(), _lib.ptr(logits_out), _lib.ptr(argmax_out), _lib.current_stream_ptr(dev))
        _lib.check(rc, "msa_fusion_forward")

    def prepare(self, B: int) -> None:
        """Pack the weights and size the workspace for batches of up to B rows (allocation happens here, never in
        ``forward_into``)."""
        dev = _lib.require_cuda(self.device)
        self._ensure_packed()
        self._ws(B, dev)

    def fused_with_argmax(self, face, audio, text=None):
        """Additive batched entry point: (logits [B,7], argmax [B] int32) in one call."""
        return self._device_forward(face, audio, text, want_argmax=True)

    def class_probs(self, logits: torch.Tensor) -> torch.Tensor:
        """Additive output: softmax over the 7 classes of the "fused" logits (the reference returns raw logits,
        fusion_model.py:94, and its consumers argmax them)."""
        x = logits.to(self.device).float().reshape(-1, 7).contiguous()
        out = torch.empty_like(x)
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().msa_softmax7(_lib.ptr(x), x.shape[0], _lib.ptr(out), _lib.current_stream_ptr(self.device)), "msa_softmax7")
        return out.reshape(logits.shape)

    # ------------------------------------------------------------------ reference API
    def forward(self, face_probs: Optional[torch.Tensor] = None, audio_probs: Optional[torch.Tensor] = None,
                text_probs: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """fusion_model.py:131-190.  Keys of the result are a subset of {face, audio, text, fused};
        the face/audio/text entries are the caller's own tensors (returned by reference, :408)."""
        try:
            present = [n for n, t in (("face", face_probs), ("audio", audio_probs), ("text", text_probs)) if t is not None]
            if not present:
                raise ValueError("Nenhuma modalidade disponível para fusão")
            if len(present) == 1:
                return {present[0]: {"face": face_probs, "audio": audio_probs, "text": text_probs}[present[0]]}
            if len(present) == 2:
                if "face" in present and "audio" in present:
                    return {"face": face_probs, "audio": audio_probs,
                            "fused": self._device_forward(face_probs, audio_probs, None)}       # :296-321
                # face+text and audio+text feed a 1024-wide concat to Linear(1536, 1024) in the reference,
                # which raises and falls back to one modality (:330-384)
                raise ValueError("fusion of this pair is not defined by the reference (dimension mismatch)")
            return {"face": face_probs, "audio": audio_probs, "text": text_probs,
                    "fused": self._device_forward(face_probs, audio_probs, text_probs)}          # :386-408
        except _lib.MsaError:
            raise                                     # a missing library / device is not a data error: fail loudly
        except Exception as e:  # noqa: BLE001 - reference convention: most reliable single modality (:180-190)
            logger.error("Erro no forward do FusionModel: %s", e)
            if face_probs is not None:
                return {"face": face_probs}
            if audio_probs is not None:
                return {"audio": audio_probs}
            if text_probs is not None:
                return {"text": text_probs}
            raise ValueError("Nenhuma modalidade disponível para retorno de fallback")

    def get_weights(self) -> Dict[str, float]:
        """fusion_model.py:192-203."""
        w = self.softmax(torch.stack([self.audio_weight, self.text_weight, self.face_weight]).detach().float().cpu())
        return {"audio": w[0].item(), "text": w[1].item(), "face": w[2].item()}

    def save(self, path: str):
        """fusion_model.py:239-257: same checkpoint dictionary."""
        Path(path).parent.mkdir(parents=True, exist_ok=True)
        torch.save({"model_state_dict": self.state_dict(), "weights": self.get_weights(), "audio_dim": self.audio_dim,
                    "text_dim": self.text_dim, "face_dim": self.face_dim, "hidden_dim": self.hidden_dim,
                    "output_dim": self.output_dim, "dropout": self.dropout}, path)

    @classmethod
    def load(cls, path: str, device: Optional[str] = None) -> "AdvancedFusionModel":
        """fusion_model.py:259-294, including its quirks: the three modality scalars are overwritten
        with the SOFTMAXED values stored in the checkpoint, and a missing file creates, saves and
        returns a freshly initialised model."""
        try:
            ck = torch.load(path, map_location="cpu")
            model = cls(audio_dim=ck["audio_dim"], text_dim=ck["text_dim"], face_dim=ck["face_dim"],
                        hidden_dim=ck["hidden_dim"], output_dim=ck["output_dim"], dropout=ck["dropout"], device=device)
            model.load_state_dict(ck["model_state_dict"])
            w = ck.get("weights", {"audio": 0.3, "text": 0.3, "face": 0.4})
            model.audio_weight.data.fill_(w["audio"])
            model.text_weight.data.fill_(w["text"])
            model.face_weight.data.fill_(w["face"])
            return model
        except FileNotFoundError:
            logger.warning("Checkpoint não encontrado em %s. Criando novo modelo...", path)
            Path(path).parent.mkdir(parents=True, exist_ok=True)
            model = cls(device=device)
            model.save(path)
            return model


FusionModel = AdvancedFusionModel

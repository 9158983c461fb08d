"""``AudioAnalyzer`` with the reference's interface, computed by the fused sm_100a feature kernel.

Mirrors /root/reference/src/analyzers/audio_analyzer.py: same constructor arguments and
attributes (:15-54), ``analyze(audio_path, speaker_id) -> AudioAnalysis`` (:56-150), the ten
``_analyze_*`` / ``_calculate_*`` feature methods (:152-329) and ``_get_default_analysis``
(:331-345), with the reference's error convention: nothing raises out of a feature call, a failure
is logged and the documented default is returned.

One kernel launch (``msa_features_*``, csrc/msa_features_body.cuh) computes every feature of a
segment; the per-method shims below run it with the matching part mask and slice its output.
``analyze_batch`` is the additive batched entry point: B independent ``[1, T]`` reference calls.

Out of scope: the wav2vec2 emotion classifier.  ``_analyze_emotion``
returns an injected embedding, or the reference's own failure fallback (uniform 1/8).
"""
from __future__ import annotations

import logging
import wave
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from .structures import AudioAnalysis

logger = logging.getLogger(__name__)

_RAW = slice(0, 27)


from .ingest import AudioFeatureNormalizer, resample  # noqa: E402  (normalization.py:19-44; audio_analyzer.py:74-77)


class AudioAnalyzer:
    def __init__(self, device: str = "cuda", sample_rate: int = 16000, strict_reference: bool = True):
        """device / sample_rate as audio_analyzer.py:15-19.  ``strict_reference`` keeps the reference's
        mono behaviour (intensity = NaN, hence an all-NaN normalised row)."""
        self.device = _lib.require_cuda(device)
        if sample_rate != 16000:
            raise ValueError("the CUDA feature kernel hard-wires the reference's 16 kHz constants (audio_analyzer.py:52-53)")
        self.sample_rate = sample_rate
        self.normalizer = AudioFeatureNormalizer(self.device)
        self.window_size = 0.025
        self.hop_length = 0.010
        self.strict_reference = strict_reference
        self.emotion_embedding: Optional[torch.Tensor] = None   # injected [1, 8] output of the out-of-scope SER model
        self._lib = _lib.lib()
        self._ws: Optional[torch.Tensor] = None
        self.buffers_generation = 0     # bumped when the scratch table is reallocated (captured graphs hold its address)
        logger.info("AudioAnalyzer (msa_b200, sm_100a) on %s, sample_rate %d", self.device, sample_rate)

    # ------------------------------------------------------------------ kernel entry
    def _flags(self) -> int:
        return _lib.FEAT_STRICT_NAN if self.strict_reference else 0

    def _run(self, waves: torch.Tensor, emo8: Optional[torch.Tensor], parts: int, want_mfcc: bool = False
             ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
        """waves: [B, T] fp32 or int16, contiguous, on self.device -> (feat31 [B,31], detail [B,96], mfcc)."""
        B, T = waves.shape
        feat = torch.empty(B, 31, device=self.device, dtype=torch.float32)
        detail = torch.empty(B, _lib.DETAIL_STRIDE, device=self.device, dtype=torch.float32)
        mfcc = torch.empty(B, T // 200 + 1, 13, device=self.device, dtype=torch.float32) if want_mfcc else None
        fn = self._lib.msa_features_ws_s16 if waves.dtype == torch.int16 else self._lib.msa_features_ws_f32
        ws = self._workspace(B, T)
        with _lib.on_device(self.device):
            rc = fn(_lib.ptr(waves), B, T, _lib.ptr(emo8), _lib.ptr(feat), _lib.ptr(detail), _lib.ptr(mfcc), self._flags(), parts, 0,
                    _lib.ptr(ws), ws.numel(), _lib.current_stream_ptr(self.device))
        _lib.check(rc, "msa_features")
        return feat, detail, mfcc

    def _workspace(self, B: int, T: int) -> torch.Tensor:
        """Grow-only scratch table for the top_db clamp of pause-heavy segments (msa_features_workspace_bytes)."""
        need = self._lib.msa_features_workspace_bytes(B, T)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(max(int(need), 1), dtype=torch.uint8, device=self.device)
            self.buffers_generation += 1
        return self._ws

    def _mono(self, waveform: torch.Tensor) -> torch.Tensor:
        """The reference's only working layout is [1, T]."""
        if not isinstance(waveform, torch.Tensor) or waveform.dim() != 2 or waveform.shape[0] != 1 or waveform.shape[1] < 1:
            raise ValueError(f"expected a mono waveform of shape [1, T], got {tuple(getattr(waveform, 'shape', ()))}")
        w = waveform.to(self.device)
        if w.dtype != torch.int16:
            w = w.float()
        return w.contiguous()

    def _emo(self, B: int, emotion_probs: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        e = emotion_probs if emotion_probs is not None else self.emotion_embedding
        if e is None:
            return None
        e = e.to(self.device).float().reshape(-1, 8)
        if e.shape[0] == 1 and B > 1:
            e = e.expand(B, 8)
        return e.contiguous()

    # ------------------------------------------------------------------ batched API (additive)
    def analyze_batch(self, waveforms: torch.Tensor, emotion_probs: Optional[torch.Tensor] = None,
                      return_detail: bool = False):
        """B independent segments: waveforms [B, T] fp32 in [-1, 1] or int16 PCM (device tensor).
        Returns the [B, 31] audio rows AdvancedFusionModel.forward takes (LayerNorm31[:27] ++ quality4,
        NaN -> 0, exactly what streaming_processor.py:250-268,295-298 builds per segment); with
        ``return_detail`` also the [B, 96] record (raw features, quality floats, LayerNorm row)."""
        w = waveforms.to(self.device)
        if w.dtype != torch.int16:
            w = w.float()
        w = w.contiguous()
        feat, detail, _ = self._run(w, self._emo(w.shape[0], emotion_probs), _lib.PART_ALL)
        return (feat, detail) if return_detail else feat

    def analyze_into(self, waveforms: torch.Tensor, out_rows: torch.Tensor, emotion_probs: Optional[torch.Tensor] = None,
                     workspace: Optional[torch.Tensor] = None) -> None:
        """``analyze_batch`` writing the [B, 31] rows into a caller-owned contiguous buffer (no detail record,
        no allocation): the chunked host-buffer pipeline fills one row table slice by slice.  ``workspace``: a
        caller-owned scratch table of ``msa_features_workspace_bytes(B, T)`` bytes (a CUDA-graph capture must not depend
        on this object's grow-only table, whose address changes when a larger batch comes along)."""
        B, T = waveforms.shape
        if (not waveforms.is_contiguous() or not out_rows.is_contiguous() or out_rows.shape != (B, 31)
                or out_rows.dtype != torch.float32 or waveforms.dtype not in (torch.int16, torch.float32)):
            raise ValueError("analyze_into needs contiguous [B, T] int16/fp32 waves and a contiguous fp32 [B, 31] output")
        fn = self._lib.msa_features_ws_s16 if waveforms.dtype == torch.int16 else self._lib.msa_features_ws_f32
        ws = workspace if workspace is not None else self._workspace(B, T)
        with _lib.on_device(self.device):
            rc = fn(_lib.ptr(waveforms), B, T, _lib.ptr(self._emo(B, emotion_probs)), _lib.ptr(out_rows), None, None, self._flags(),
                    _lib.PART_ALL, 0, _lib.ptr(ws), ws.numel(), _lib.current_stream_ptr(self.device))
        _lib.check(rc, "msa_features")

    def track_pitch(self, waveforms: torch.Tensor, with_voicing: bool = True):
        """Additive output (NOT computed by the reference): a real f0 track per segment,
        = torchaudio.functional.detect_pitch_frequency(waveform, 16000), the raw best lag per 10 ms frame and
        frame-level voicing flags.  waveforms [B, T] fp32 / int16 on the device ->
        {"f0": [B, n_frames - 15] Hz, "lags": [B, n_frames] int32, "voiced": [B, (T - 400) // 160 + 1] int32}."""
        w = waveforms.to(self.device)
        if w.dtype != torch.int16:
            w = w.float()
        w = w.contiguous()
        B, T = w.shape
        lib = self._lib
        lags = torch.empty(B, lib.msa_pitch_frames(T), device=self.device, dtype=torch.int32)
        f0 = torch.empty(B, lib.msa_pitch_outputs(T), device=self.device, dtype=torch.float32)
        voiced = torch.empty(B, lib.msa_voiced_frames(T), device=self.device, dtype=torch.int32) if with_voicing else None
        fn = lib.msa_pitch_track_s16 if w.dtype == torch.int16 else lib.msa_pitch_track_f32
        with _lib.on_device(self.device):
            _lib.check(fn(_lib.ptr(w), B, T, _lib.ptr(lags), _lib.ptr(f0), _lib.ptr(voiced), _lib.current_stream_ptr(self.device)),
                       "msa_pitch_track")
        return {"f0": f0, "lags": lags, "voiced": voiced}

    def spectral_descriptors(self, waveforms: torch.Tensor) -> dict:
        """Additive output (NOT computed by the reference): per frame of the MFCC's STFT grid
        (n_fft 400, hop 200) the spectral centroid [Hz] (= torchaudio.functional.spectral_centroid), the 85 % roll-off
        [Hz], the spectral flux and an onset-strength envelope.  waveforms [B, T] fp32 / int16 on the device ->
        {"centroid", "rolloff", "flux", "onset"}: [B, T // 200 + 1] fp32 each."""
        w = waveforms.to(self.device)
        if w.dtype != torch.int16:
            w = w.float()
        w = w.contiguous()
        B, T = w.shape
        out = torch.empty(B, self._lib.msa_spectral_frames(T), 4, device=self.device, dtype=torch.float32)
        fn = self._lib.msa_spectral_s16 if w.dtype == torch.int16 else self._lib.msa_spectral_f32
        with _lib.on_device(self.device):
            _lib.check(fn(_lib.ptr(w), B, T, _lib.ptr(out), _lib.current_stream_ptr(self.device)), "msa_spectral")
        return {"centroid": out[..., 0], "rolloff": out[..., 1], "flux": out[..., 2], "onset": out[..., 3]}

    # ------------------------------------------------------------------ reference API
    def analyze(self, audio_path: str, speaker_id: str) -> AudioAnalysis:
        """audio_analyzer.py:56-150: load, resample to 16 kHz, all features, LayerNorm(31), slices."""
        try:
            pcm, sr, channels = _read_wav(audio_path)
            if channels != 1:
                raise ValueError("multi-channel audio makes the reference's torch.cat fail (falls to the default analysis)")
            if sr != self.sample_rate:                                                # audio_analyzer.py:74-77
                w = resample(torch.from_numpy(pcm)[None, :].to(self.device), sr, self.sample_rate).contiguous()
            else:
                w = torch.from_numpy(pcm)[None, :].to(self.device).contiguous()      # int16 PCM ingest
            _, detail, _ = self._run(w, self._emo(1, None), _lib.PART_ALL)
            d = detail[0]
            ln = d[32:63]
            q = d[27:31].tolist()                                                     # the .item() syncs of :128-131
            return AudioAnalysis(
                speaker_id=speaker_id,
                emotion_probs=ln[0:8].reshape(1, 8).clone(), pitch=ln[8:9].reshape(1, 1).clone(),
                intensity=ln[9:10].reshape(1, 1).clone(), timbre=ln[10:23].reshape(1, 13).clone(),
                speech_rate=ln[23:24].reshape(1, 1).clone(), rhythm=ln[24:27].reshape(1, 3).clone(),
                audio_quality=q[0], signal_noise_ratio=q[1], clarity=q[2], consistency=q[3])
        except _lib.MsaError:
            raise               # a missing library, a wrong device or a failed launch is not a data error: fail loudly
        except Exception as e:                                                        # noqa: BLE001 - reference convention
            logger.error("audio analysis failed: %s", e, exc_info=True)
            return self._get_default_analysis(speaker_id)

    def _analyze_emotion(self, waveform: torch.Tensor) -> torch.Tensor:
        """audio_analyzer.py:152-173.  The wav2vec2 classifier is out of scope: returns the injected
        embedding or the reference's own fallback, a uniform 1/8 distribution."""
        if self.emotion_embedding is not None:
            return self.emotion_embedding.to(self.device).float().reshape(1, 8)
        return torch.ones(1, 8, device=self.device).float() / 8

    def _feature(self, waveform, parts, sl, shape, min_len):
        w = self._mono(waveform)
        if w.shape[1] < min_len:
            raise ValueError(f"segment of {w.shape[1]} samples is too short for this feature (reference raises too)")
        _, detail, _ = self._run(w, None, parts)
        return detail[0, sl].reshape(shape).clone()

    def _channels(self, waveform: torch.Tensor) -> torch.Tensor:
        """[C, T] with C >= 2 (direct callers of the per-method API; analyze() itself only sees mono files)."""
        if not isinstance(waveform, torch.Tensor) or waveform.dim() != 2 or waveform.shape[0] < 2 or waveform.shape[1] < 1:
            raise ValueError("not a multi-channel waveform")
        w = waveform.to(self.device)
        if w.dtype != torch.int16:
            w = w.float()
        return w.contiguous()

    def _analyze_pitch(self, waveform: torch.Tensor) -> torch.Tensor:
        """audio_analyzer.py:175-188 -> [1, 1].  (For C >= 2 channels the reference returns a finite [1, C] that depends
        on the relative size of its fp32 round-trip noise per channel; that noise is not reproduced here - the round trip
        runs in fp16 - so multi-channel input takes the method's documented default.)"""
        try:
            return self._feature(waveform, _lib.PART_PITCH, slice(8, 9), (1, 1), 257)
        except _lib.MsaError:
            raise
        except Exception as e:  # noqa: BLE001
            print(f"Erro na análise de pitch: {e}")
            return torch.zeros(1, 1, device=self.device)

    def _analyze_intensity(self, waveform: torch.Tensor) -> torch.Tensor:
        """audio_analyzer.py:190-201 -> [1, 1] (NaN for mono in strict mode); [1, C] for C >= 2 channels: the z-score of
        the channel energies (every channel runs through the kernel as a segment; the C-element z-score is host glue)."""
        try:
            if isinstance(waveform, torch.Tensor) and waveform.dim() == 2 and waveform.shape[0] >= 2:
                w = self._channels(waveform)
                _, detail, _ = self._run(w, None, _lib.PART_WAVE)
                e = detail[:, 68]                                                # e_total per channel
                return ((e - e.mean()) / (e.std() + 1e-6)).unsqueeze(0)
            return self._feature(waveform, _lib.PART_WAVE, slice(9, 10), (1, 1), 1)
        except _lib.MsaError:
            raise
        except Exception as e:  # noqa: BLE001
            print(f"Erro na análise de intensidade: {e}")
            return torch.zeros(1, 1, device=self.device)

    def _analyze_timbre(self, waveform: torch.Tensor) -> torch.Tensor:
        """audio_analyzer.py:203-217 -> [1, 13].  (For C >= 2 channels the reference returns [1, C, 13] whose top_db clamp
        uses the maximum over ALL channels - torchaudio packs the channels of a 3-D input into one clamp group - while
        the kernel clamps every segment against its own maximum; multi-channel input takes the documented default.)"""
        try:
            return self._feature(waveform, _lib.PART_MFCC, slice(10, 23), (1, 13), 201)
        except _lib.MsaError:
            raise
        except Exception as e:  # noqa: BLE001
            print(f"Erro na análise de timbre: {e}")
            return torch.zeros(1, 13, device=self.device)

    def _analyze_speech_rate(self, waveform: torch.Tensor) -> torch.Tensor:
        """audio_analyzer.py:219-233 -> [1, 1]."""
        try:
            return self._feature(waveform, _lib.PART_WAVE, slice(23, 24), (1, 1), 1)
        except _lib.MsaError:
            raise
        except Exception as e:  # noqa: BLE001
            print(f"Erro na análise de velocidade: {e}")
            return torch.zeros(1, 1, device=self.device)

    def _analyze_rhythm(self, waveform: torch.Tensor) -> torch.Tensor:
        """audio_analyzer.py:235-263 -> [1, 3]; T < 400 makes the reference's unfold raise -> zeros."""
        try:
            return self._feature(waveform, _lib.PART_WAVE, slice(24, 27), (1, 3), 400)
        except _lib.MsaError:
            raise
        except Exception as e:  # noqa: BLE001
            print(f"Erro na análise de ritmo: {e}")
            return torch.zeros(1, 3, device=self.device)

    def _quality(self, waveform, parts, idx, min_len):
        w = self._mono(waveform)
        if w.shape[1] < min_len:
            raise ValueError("too short")
        _, detail, _ = self._run(w, None, parts)
        return detail[0, 27 + idx].item()

    def _calculate_audio_quality(self, waveform: torch.Tensor) -> float:
        """audio_analyzer.py:265-276: 0.4 snr + 0.3 clarity + 0.3 consistency; a term whose own
        computation raises in the reference (too-short segment) contributes its 0.0 default, which the
        kernel reproduces per term."""
        try:
            return self._quality(waveform, _lib.PART_WAVE | _lib.PART_MFCC, 0, 1)
        except _lib.MsaError:
            raise
        except Exception:  # noqa: BLE001
            return 0.0

    def _calculate_signal_noise_ratio(self, waveform: torch.Tensor) -> float:
        """audio_analyzer.py:278-293; int(0.05*T) == 0 makes torch.cat raise -> 0.0."""
        try:
            return self._quality(waveform, _lib.PART_WAVE, 1, 20)
        except _lib.MsaError:
            raise
        except Exception:  # noqa: BLE001
            return 0.0

    def _calculate_clarity(self, waveform: torch.Tensor) -> float:
        """audio_analyzer.py:295-311."""
        try:
            return self._quality(waveform, _lib.PART_MFCC, 2, 201)
        except _lib.MsaError:
            raise
        except Exception:  # noqa: BLE001
            return 0.0

    def _calculate_consistency(self, waveform: torch.Tensor) -> float:
        """audio_analyzer.py:313-329; T < 1600 makes unfold raise -> 0.0."""
        try:
            return self._quality(waveform, _lib.PART_WAVE, 3, 1600)
        except _lib.MsaError:
            raise
        except Exception:  # noqa: BLE001
            return 0.0

    def _get_default_analysis(self, speaker_id: str) -> AudioAnalysis:
        """audio_analyzer.py:331-345."""
        z = lambda n: torch.zeros(1, n, device=self.device)
        return AudioAnalysis(speaker_id=speaker_id, emotion_probs=torch.ones(1, 8, device=self.device) / 8, pitch=z(1),
                             intensity=z(1), timbre=z(13), speech_rate=z(1), rhythm=z(3), audio_quality=0.0,
                             signal_noise_ratio=0.0, clarity=0.0, consistency=0.0)


def _read_wav(path: str):
    """PCM s16le reader (the wire format both processors write: offline_processor.py:87-91,
    streaming_processor.py:190-196); replaces torchaudio.load, which needs torchcodec."""
    with wave.open(path, "rb") as wf:
        if wf.getsampwidth() != 2:
            raise ValueError("only 16-bit PCM wav files are supported")
        sr, ch = wf.getframerate(), wf.getnchannels()
        pcm = np.frombuffer(wf.readframes(wf.getnframes()), dtype=np.int16)
    if ch > 1:
        pcm = pcm.reshape(-1, ch).T.copy()
        return pcm, sr, ch
    return pcm.copy(), sr, ch

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
    ring[pos : pos + window] (pos = next write position) and no on-device slide is needed.

    Per hop the device work is: one host->device copy from a pinned staging buffer (16 KB of PCM and the face / text
    rows), one copy of the chunk into the ring and its mirror, the feature kernel (one segment spread over a cluster of
    16 CTAs), the fusion chain (5 launches) and a 32-byte read-back.  A hop is launch-latency bound, so once the window is full the whole
    sequence is captured into one CUDA graph per ring position (window / hop of them) and replayed: one
    graph launch per chunk instead of ~10 stream operations.  ``use_graph=False`` keeps the eager path."""

    N_STAGE = 4

    def __init__(self, analyzer: AudioAnalyzer, fusion: AdvancedFusionModel, window: int = 80000, hop: int = 8000,
                 use_graph: bool = True):
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
        self.use_graph = use_graph
        self._graphs = {}       # (ring position, has_text) -> torch.cuda.CUDAGraph
        self._g = None          # static buffers of the graph path

    @property
    def filled(self) -> int:
        return min(self.window, self.n_pushed * self.hop)

    # ------------------------------------------------------------------ eager path (also the warm-up of the graph path)
    @torch.no_grad()
    def _push_eager(self, chunk_pcm: torch.Tensor, face: torch.Tensor, text: Optional[torch.Tensor]):
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

    # ------------------------------------------------------------------ graph path
    def _static(self):
        if self._g is None:
            dev = self.device
            lib = self.analyzer._lib
            # ONE pinned host buffer and ONE device buffer with the same layout [hop int16 | 27 floats | 783 floats]: a
            # hop uploads with a single copy (a graph node costs ~2-3 us of dependent latency; there were three);
            # the fusion kernels read the face / text rows where they land
            nb_pcm, nb_face, nb_text = 2 * self.hop, 4 * 27, 4 * 783
            if nb_pcm % 16:
                raise ValueError("hop must be a multiple of 8 samples")
            hbuf = torch.zeros(nb_pcm + nb_face + nb_text, dtype=torch.uint8).pin_memory()
            dbuf = torch.zeros(nb_pcm + nb_face + nb_text, dtype=torch.uint8, device=dev)
            cut = lambda b, dt, a, n: b[a:a + n].view(dt)
            obuf = torch.zeros(8, dtype=torch.float32, device=dev)           # 7 logits, then the argmax (int32 bits)
            out = torch.zeros(8, dtype=torch.float32).pin_memory()
            self._g = {"hbuf": hbuf, "dbuf": dbuf,
                       "stage": cut(hbuf, torch.int16, 0, nb_pcm),
                       "face_h": cut(hbuf, torch.float32, nb_pcm, nb_face).view(1, 27),
                       "text_h": cut(hbuf, torch.float32, nb_pcm + nb_face, nb_text).view(1, 783),
                       "pcm": cut(dbuf, torch.int16, 0, nb_pcm),
                       "face": cut(dbuf, torch.float32, nb_pcm, nb_face).view(1, 27),
                       "text": cut(dbuf, torch.float32, nb_pcm + nb_face, nb_text).view(1, 783),
                       "row": torch.zeros(1, 31, device=dev), "obuf": obuf, "logits": obuf[:7].view(1, 7),
                       "amax": obuf.view(torch.int32)[7:8],
                       "out": out, "out_argmax": out.view(torch.int32)[7:8],
                       # the window's OWN scratch table for the top_db clamp: a captured graph must not point into the
                       # analyzer's grow-only table, which is replaced when a larger batch comes along
                       "ws": torch.empty(max(1, lib.msa_features_workspace_bytes(1, self.window)), dtype=torch.uint8, device=dev),
                       "stream": torch.cuda.Stream(dev), "done": torch.cuda.Event(), "busy": False, "gen": None}
        return self._g

    def _device_hop(self, p: int, has_text: bool):
        """The device work of one hop at ring position p, on the current stream, with static buffers only: ONE upload
        (16 KB of PCM, the face row, the text row), one copy of the chunk into the ring and its mirror, the feature
        kernel, the fusion kernels and ONE 32-byte read-back are all inside the captured graph, so a hop costs the host
        one graph launch and the device nine dependent nodes (thirteen before: three uploads, the mirror copy, a
        conversion kernel and two read-backs)."""
        g = self._g
        g["dbuf"].copy_(g["hbuf"], non_blocking=True)
        self.ring.view(2, self.window)[:, p:p + self.hop].copy_(g["pcm"].expand(2, self.hop))   # chunk and mirror at once
        nxt = (p + self.hop) % self.window
        self.analyzer.analyze_into(self.ring[nxt:nxt + self.window][None, :], g["row"], workspace=g["ws"])
        self.fusion.forward_into(g["face"], g["row"], g["text"] if has_text else None, g["logits"], g["amax"])
        g["out"].copy_(g["obuf"], non_blocking=True)

    def _graph_for(self, p: int, has_text: bool):
        g = self._static()
        f = self.fusion
        if f._packed is None or f._packed_key is None or f._workspace is None:
            f.prepare(1)                                           # first hop, or load_state_dict since the last one
        gen = f.buffers_generation
        if g["gen"] != gen:                                        # the fusion model's packed blob or workspace moved:
            self._graphs.clear()                                   # graphs captured before hold stale addresses
            g["gen"] = gen
        key = (p, has_text)
        gr = self._graphs.get(key)
        if gr is None:
            s = g["stream"]
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                self._device_hop(p, has_text)                      # warm-up outside the capture (tables, attributes)
                s.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=s):
                    self._device_hop(p, has_text)
            torch.cuda.current_stream(self.device).wait_stream(s)
            self._graphs[key] = gr
        return gr

    def input_buffers(self):
        """The pinned host buffers a hop's graph uploads from: {"pcm": [hop] int16, "face": [1, 27], "text": [1, 783]}.
        A producer that writes the next chunk straight into them (``pipe.readinto(memoryview(buf["pcm"].numpy()))``, the
        way streaming_processor.py:185-196 reads its ffmpeg pipe) and then calls ``push_staged`` pays no host copy at all.
        They may be rewritten as soon as the previous hop's ``done`` event has fired."""
        g = self._static()
        return {"pcm": g["stage"], "face": g["face_h"], "text": g["text_h"]}

    @torch.no_grad()
    def push_staged(self, has_text: bool = True):
        """``push`` for inputs already written into ``input_buffers()`` (window must be full: use ``push`` to fill it)."""
        if not self.use_graph or (self.n_pushed + 1) * self.hop < self.window:
            g = self._static()
            return self.push(g["stage"].clone(), g["face_h"].clone(), g["text_h"].clone() if has_text else None)
        g = self._static()
        p = self.pos
        self._graph_for(p, has_text).replay()
        g["done"].record(torch.cuda.current_stream(self.device))
        g["busy"] = True
        self.pos = (p + self.hop) % self.window
        self.n_pushed += 1
        return {"fused_emotion": g["logits"][0], "argmax": g["amax"][0], "audio_row": g["row"][0], "host": g["out"],
                "host_argmax": g["out_argmax"], "done": g["done"]}

    @torch.no_grad()
    def push(self, chunk_pcm: torch.Tensor, face: torch.Tensor, text: Optional[torch.Tensor] = None):
        """chunk_pcm: [hop] int16 on the host.  Returns None until the window is full, then the dict of
        streaming_processor.py:302-320: {"fused_emotion": logits [7], "argmax": int, "audio_row": [31]}.
        On the graph path the results live in static buffers that the next push overwrites, and
        ``result["host"]`` is a pinned [8] float tensor (7 logits; the last slot carries the argmax's int32 bits, read it
        through ``result["host_argmax"]``, an int32 view of that slot) valid after ``result["done"].synchronize()``
        (an event recorded behind the hop's read-back: waiting on it is cheaper than synchronising the stream)."""
        if not self.use_graph or (self.n_pushed + 1) * self.hop < self.window:
            return self._push_eager(chunk_pcm, face, text)
        g = self._static()
        if g["busy"]:
            g["done"].synchronize()                                # the previous replay has consumed the staging buffers
        g["stage"].copy_(chunk_pcm.reshape(-1))                    # host -> pinned (a pinned chunk costs a memcpy of 16 KB)
        g["face_h"].copy_(face.reshape(1, -1))
        if text is not None:
            g["text_h"].copy_(text.reshape(1, -1))
        p = self.pos
        self._graph_for(p, text is not None).replay()
        g["done"].record(torch.cuda.current_stream(self.device))
        g["busy"] = True
        self.pos = (p + self.hop) % self.window
        self.n_pushed += 1
        return {"fused_emotion": g["logits"][0], "argmax": g["amax"][0], "audio_row": g["row"][0], "host": g["out"],
                "host_argmax": g["out_argmax"], "done": g["done"]}

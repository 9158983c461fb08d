"""Batched segment pipeline: audio features -> fusion -> per-segment result rows, sharded over the
GPUs of one box, gathered, and aggregated per speaker.

This is the data-parallel loop of /root/reference/src/processors/offline_processor.py:255-257
(``for segment in segments: process_segment(...)``) and its speaker grouping (:259-298), with the
per-segment body of streaming_processor.py:250-320 (audio row -> nan_to_num -> fusion -> argmax).

Sharding: segments are independent, so rank r of N takes the contiguous range
``shard_range(S, N, r)``, runs the same two kernels on its slice, and ONE collective
(``all_gather`` of 40 x 32-bit words per segment over NCCL) assembles the result table; there is
no other exchange on the data path.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib
from .audio_analyzer import AudioAnalyzer
from .fusion_model import AdvancedFusionModel

ROW_WORDS = 40          # audio row 31 | logits 7 | argmax | segment id  (160 bytes / segment)


def shard_range(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of rank `rank`: sizes differ by at most one, earlier ranks get the extra."""
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def pack_rows(audio_row: torch.Tensor, logits: torch.Tensor, argmax: torch.Tensor, first_id: int) -> torch.Tensor:
    """[n, 40] fp32 result table on the device (msa_pack_rows); the two integer columns are stored bit-exactly
    (int32 viewed as fp32)."""
    if not audio_row.is_cuda:
        raise _lib.MsaError("pack_rows needs device tensors (msa_b200 has no CPU path)")
    n = audio_row.shape[0]
    rows = torch.empty(n, ROW_WORDS, device=audio_row.device, dtype=torch.float32)
    a, l, m = audio_row.float().contiguous(), logits.float().contiguous(), argmax.to(torch.int32).contiguous()
    rc = _lib.lib().msa_pack_rows(_lib.ptr(a), _lib.ptr(l), _lib.ptr(m), int(first_id), n, _lib.ptr(rows),
                                  _lib.current_stream_ptr(audio_row.device))
    _lib.check(rc, "msa_pack_rows")
    return rows


def unpack_rows(rows: torch.Tensor) -> Dict[str, torch.Tensor]:
    ints = rows.view(torch.int32)
    return {"audio_row": rows[:, 0:31], "logits": rows[:, 31:38], "argmax": ints[:, 38], "segment_id": ints[:, 39]}


def gather_rows(local_rows: torch.Tensor, n_total: int, world_size: int, rank: int, group=None) -> torch.Tensor:
    """The only collective of the path: all_gather of the per-rank result tables (ragged shards are
    padded to the largest shard and trimmed after the gather).  world_size 1 is a no-op."""
    if world_size == 1:
        return local_rows
    import torch.distributed as dist
    sizes = [shard_range(n_total, world_size, r)[1] - shard_range(n_total, world_size, r)[0] for r in range(world_size)]
    cap = max(sizes)
    buf = torch.zeros(cap, ROW_WORDS, device=local_rows.device, dtype=torch.float32)
    buf[: local_rows.shape[0]] = local_rows
    out = torch.empty(world_size * cap, ROW_WORDS, device=local_rows.device, dtype=torch.float32)
    dist.all_gather_into_tensor(out, buf, group=group)
    return torch.cat([out[r * cap: r * cap + sizes[r]] for r in range(world_size)], dim=0)


class PendingGather:
    """Result of ``gather_rows_async``: the all_gather runs on NCCL's own stream behind the kernels that produced the
    local rows, and the compute stream goes on with the next batch.  ``wait()`` makes the current stream wait for it and
    returns the gathered [n_total, 40] table."""

    def __init__(self, out, work, sizes, cap, keep):
        self._out, self._work, self._sizes, self._cap, self._keep = out, work, sizes, cap, keep

    def wait(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()                    # stream-level wait (the host does not block)
            self._work = None
        if self._sizes is None or all(n == self._cap for n in self._sizes):
            return self._out                     # equal shards: the gathered buffer IS the table (no copy)
        return torch.cat([self._out[r * self._cap: r * self._cap + n] for r, n in enumerate(self._sizes)], dim=0)


def gather_rows_async(local_rows: torch.Tensor, n_total: int, world_size: int, rank: int, group=None) -> PendingGather:
    """``gather_rows`` without stalling the compute stream: the collective is latency-bound (160 bytes per segment), so a
    step that waits for it pays ~45 us for nothing; issued asynchronously it overlaps the next batch's kernels
    (offline_processor.py only needs the table when all segments are done, :259)."""
    if world_size == 1:
        return PendingGather(local_rows, None, None, 0, None)
    import torch.distributed as dist
    sizes = [shard_range(n_total, world_size, r)[1] - shard_range(n_total, world_size, r)[0] for r in range(world_size)]
    cap = max(sizes)
    if local_rows.shape[0] == cap:
        buf = local_rows
    else:
        buf = torch.zeros(cap, ROW_WORDS, device=local_rows.device, dtype=torch.float32)
        buf[: local_rows.shape[0]] = local_rows
    out = torch.empty(world_size * cap, ROW_WORDS, device=local_rows.device, dtype=torch.float32)
    work = dist.all_gather_into_tensor(out, buf, group=group, async_op=True)
    return PendingGather(out, work, sizes, cap, (buf, local_rows))


def aggregate_speakers(argmax: torch.Tensor, speaker: torch.Tensor, n_speakers: int) -> Dict[str, torch.Tensor]:
    """offline_processor.py:259-298 on device: per-speaker label histogram, dominant emotion (mode)
    and the starts of three-in-a-row "patterns" within each speaker's own segment sequence."""
    dev = argmax.device
    a = argmax.to(torch.int32).contiguous()
    s = speaker.to(dev, torch.int32).contiguous()
    S = a.shape[0]
    hist = torch.empty(n_speakers, 7, device=dev, dtype=torch.int32)
    dom = torch.empty(n_speakers, device=dev, dtype=torch.int32)
    run3 = torch.zeros(max(S, 1), device=dev, dtype=torch.int32)
    rc = _lib.lib().msa_aggregate_speakers(_lib.ptr(a), _lib.ptr(s), S, n_speakers, _lib.ptr(hist), _lib.ptr(dom), _lib.ptr(run3),
                                           _lib.current_stream_ptr(dev))
    _lib.check(rc, "msa_aggregate_speakers")
    return {"hist": hist, "dominant": dom, "run3": run3[:S]}


class SegmentPipeline:
    """features -> fusion for a batch of segments on one GPU (one rank)."""

    def __init__(self, analyzer: AudioAnalyzer, fusion: AdvancedFusionModel):
        self.analyzer = analyzer
        self.fusion = fusion
        self.device = analyzer.device

    @torch.no_grad()
    def run(self, waves: torch.Tensor, face: torch.Tensor, text: Optional[torch.Tensor] = None,
            emotion_probs: Optional[torch.Tensor] = None, first_id: int = 0) -> torch.Tensor:
        """waves [n, T] (fp32 or int16 PCM), face [n, 27], text [n, 783] or None (the live streaming
        path has no text, streaming_processor.py:420-424) -> result table [n, 40]."""
        audio_row = self.analyzer.analyze_batch(waves, emotion_probs)
        logits, amax = self.fusion.fused_with_argmax(face, audio_row, text)
        return pack_rows(audio_row, logits, amax, first_id)

    @torch.no_grad()
    def run_host(self, pcm_host: torch.Tensor, face_host: torch.Tensor, text_host: Optional[torch.Tensor],
                 rows_host: torch.Tensor, first_id: int = 0, chunk: int = 128) -> torch.Tensor:
        """Same as ``run`` for HOST buffers (pinned: int16 PCM [n, T], face [n, 27], text [n, 783] or None).
        The PCM is cut into chunks of `chunk` segments: a copy stream uploads chunk i+1 into the next slot of
        a ring of three staging buffers while the compute stream runs the feature kernel on chunk i (the upload is the
        longer of the two: 160 KB per segment over PCIe).  The face / text rows go up once, the fusion chain
        runs once over all n rows when the last chunk's features are done, and the [n, 40] result table comes
        back with one device->host copy into `rows_host` (pinned).  The returned device table and `rows_host`
        are valid once the current stream has been synchronised.  Calls pipeline with each other: every staging
        buffer is guarded by its own event (the kernel that last read it), so the first upload of a call starts
        while the previous call's last kernels are still running and the copy engine never idles between calls."""
        n, T = pcm_host.shape
        dev = self.device
        if n == 0:                                                # an empty batch has nothing to upload or compute
            return torch.empty(0, ROW_WORDS, device=dev, dtype=torch.float32)
        cur = torch.cuda.current_stream(dev)
        st = self._host_state(chunk, T, text_host is not None, n)
        copy_s = st["copy_stream"]
        audio_rows = st["audio"][:n]
        # chunk boundaries; the last chunk is cut short so that little compute is left once the upload ends
        bounds = list(range(0, n, chunk)) + [n]
        tail = max(1, chunk // 4)
        if bounds[-1] - bounds[-2] > tail:
            bounds.insert(-1, n - tail)
        side = None
        for i in range(len(bounds) - 1):
            b, e = bounds[i], bounds[i + 1]
            k = st["next"]                                        # staging ring position, carried across calls
            st["next"] = (k + 1) % len(st["pcm"])
            with torch.cuda.stream(copy_s):
                if st["free"][k] is not None:
                    copy_s.wait_event(st["free"][k])              # the kernel that read this half is done
                st["pcm"][k][: e - b].copy_(pcm_host[b:e], non_blocking=True)
                up = torch.cuda.Event()
                up.record(copy_s)
                if i == 0:                                        # small side inputs ride behind the first chunk
                    if st["side_free"] is not None:
                        copy_s.wait_event(st["side_free"])        # the previous call's fusion has read them
                    st["face"][:n].copy_(face_host, non_blocking=True)
                    if text_host is not None:
                        st["text"][:n].copy_(text_host, non_blocking=True)
                    side = torch.cuda.Event()
                    side.record(copy_s)
            cur.wait_event(up)
            self.analyzer.analyze_into(st["pcm"][k][: e - b], audio_rows[b:e])
            done = torch.cuda.Event()
            done.record(cur)
            st["free"][k] = done
        cur.wait_event(side)
        text = st["text"][:n] if text_host is not None else None
        logits, amax = self.fusion.fused_with_argmax(st["face"][:n], audio_rows, text)
        st["side_free"] = torch.cuda.Event()
        st["side_free"].record(cur)
        rows = pack_rows(audio_rows, logits, amax, first_id)
        rows_host[:n].copy_(rows, non_blocking=True)
        return rows

    def _host_state(self, chunk: int, T: int, with_text: bool, n: int):
        key = (chunk, T, with_text)
        st = getattr(self, "_hs", None)
        if st is None or st["key"] != key or st["audio"].shape[0] < n:
            dev = self.device
            if st is not None:
                torch.cuda.current_stream(dev).synchronize()      # the old buffers may still be in use
                st["copy_stream"].synchronize()
            st = {"key": key, "copy_stream": torch.cuda.Stream(dev), "free": [None] * 3, "side_free": None, "next": 0,
                  "pcm": [torch.empty(chunk, T, dtype=torch.int16, device=dev) for _ in range(3)],
                  "face": torch.empty(n, 27, device=dev),
                  "text": torch.empty(n, 783, device=dev) if with_text else None,
                  "audio": torch.empty(n, 31, device=dev)}
            self._hs = st
        return st

    @torch.no_grad()
    def run_sharded(self, waves_local, face_local, text_local, n_total: int, world_size: int, rank: int,
                    emotion_probs=None, group=None, async_gather: bool = False):
        """This rank's shard through the pipeline, then the single result gather.  ``async_gather=True`` returns a
        ``PendingGather`` (the collective overlaps whatever the caller launches next; ``.wait()`` gives the table)."""
        begin, _ = shard_range(n_total, world_size, rank)
        rows = self.run(waves_local, face_local, text_local, emotion_probs, first_id=begin)
        if async_gather:
            return gather_rows_async(rows, n_total, world_size, rank, group)
        return gather_rows(rows, n_total, world_size, rank, group)

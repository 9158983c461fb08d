#!/usr/bin/env python
"""Benchmark of the audio-features -> fusion hot path (BASELINE.json metric: audio-sec/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A "step" is one pass of the hot path over one batch of synthetic input: 1024 five-second 16 kHz
mono segments per GPU (BASELINE.json configs[1]) through the fused feature kernel and the 3-modal
fusion forward, plus result-row packing and (N > 1) the one NCCL all-gather of the rows.

  value      whole-job audio-seconds per second, inputs already resident in HBM (fp32 waveforms)
  e2e        same metric through the public API with HOST buffers: pinned int16 PCM + face/text rows
             are copied host->device and the result rows device->host inside the timed region
  roofline   dominant kernel (features_kernel): 320,124 algorithmic bytes per segment
             x 1024 segments / its CUDA-event time, against the measured HBM copy bandwidth
  cpu_baseline  the torch/torchaudio port of the reference's CPU path (oracle/torch_port.py) timed on
             this box's host cores over a bounded sample (the reference itself is Python under
             /root/reference and does not exist on the GPU box)

--impl reference times that same CPU port with all host cores (one process per core, one torch
thread each: a single reference call does not scale with intra-op threads) and prints the same
JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEG_SAMPLES = 80000
SEG_SECONDS = 5.0
ALGO_BYTES_PER_SEGMENT = 80000 * 4 + 31 * 4          # waveform read once + 31-float row written
FP32_FLOP_PER_SEGMENT = 18.5e6                        # split-radix count of the path's FFTs
FP32_PEAK_TFLOPS = 74.0                               # 148 SMs x 128 lanes x 2 x 1.965 GHz
METRIC = "audio-sec/s (features+fusion)"
UNIT = "audio-s/s"


def measured_traffic(segments, lib_version):
    """DRAM bytes per launch of the feature kernel from the committed ncu --set full capture: only when that capture
    was taken at this batch size from THIS build of the library (profiles/features_traffic.json is stamped with
    msa_version(); a capture of another build gives null)."""
    p = os.path.join(ROOT, "profiles", "features_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        if int(d["segments"]) == int(segments) and int(d["msa_version"]) == int(lib_version):
            return int(d["dram_bytes_read"]) + int(d["dram_bytes_write"])
    except Exception:  # noqa: BLE001
        pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def measured_tensor_peak():
    """Sustained dense bf16 TFLOP/s (the fusion chain is timed inside a long step)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), "measured"
    return 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------ CPU arm
def cpu_kind():
    """"reference" when the reference's source tree is on this machine ($MSA_REFERENCE_DIR or /root/reference: the build
    container), else "port" (the GPU box: the pinned torch port of the same calls)."""
    from oracle import ref_runner
    return "reference" if ref_runner.available() else "port"


def cpu_port_throughput(budget_s: float, workers: int, segs_per_task: int = 2):
    """audio-s/s of the reference (or its pinned port) on host cores: (features + fusion) per segment."""
    import numpy as np
    import torch
    from oracle import ref_runner, synth, torch_port as tp

    use_ref = ref_runner.available()
    sd_np = synth.fusion_state(4321, trained_like=True)
    sd = tp.build_fusion(sd_np)
    work = ref_runner.worker if use_ref else tp._worker
    fuse = (lambda f, a, t: ref_runner.fusion_forward(sd_np, f, a, t)) if use_ref else (lambda f, a, t: tp.fusion_forward(sd, f, a, t))
    if workers <= 1 and use_ref:
        ref_runner.worker((1234, 1))                                     # imports, first-call caches
        t0 = time.perf_counter()
        ref_runner.worker((1234, 1))
        per = max((time.perf_counter() - t0) / 2.0, 1e-3)
        n = int(min(4096, max(8, budget_s / per)))
        rows, dt = ref_runner.worker((1234, n))
        torch.set_num_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        fuse(torch.from_numpy(synth.face_rows(1, n)), torch.from_numpy(rows), torch.from_numpy(synth.text_rows(3, n)))
        dt += time.perf_counter() - t0
        return n * SEG_SECONDS / dt, n, 1
    if workers <= 1:
        ana = tp.PortedAnalyzer()
        waves = [torch.from_numpy(synth.pcm_to_f32(synth.segment_pcm(1234 + i)))[None, :] for i in range(8)]
        ana.audio_row(waves[0])
        t0 = time.perf_counter()
        ana.audio_row(waves[1])
        per = max(time.perf_counter() - t0, 1e-3)
        n = int(min(4096, max(8, budget_s / per)))
        waves = [torch.from_numpy(synth.pcm_to_f32(synth.segment_pcm(1234 + i)))[None, :] for i in range(n)]
        t0 = time.perf_counter()
        a = torch.cat([ana.audio_row(w) for w in waves])
        tp.fusion_forward(sd, torch.from_numpy(synth.face_rows(1, n)), a, torch.from_numpy(synth.text_rows(3, n)))
        dt = time.perf_counter() - t0
        return n * SEG_SECONDS / dt, n, torch.get_num_threads()
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        probe = pool.map(work, [(1, 2)] * workers)                    # warm-up + per-segment cost probe
        per = max(max(dt for _, dt in probe) / 2.0, 1e-3)
        count = int(min(2048, max(4, budget_s / per)))
        out = pool.map(work, [(5000 + i * count, count) for i in range(workers)], chunksize=1)
    a = torch.from_numpy(np.concatenate([r for r, _ in out]))
    n = a.shape[0]
    torch.set_num_threads(workers)
    t0 = time.perf_counter()
    fuse(torch.from_numpy(synth.face_rows(1, n)), a, torch.from_numpy(synth.text_rows(3, n)))
    t_fus = time.perf_counter() - t0
    # the workers run concurrently: job time = slowest worker's compute + the batched fusion forward
    dt = max(d for _, d in out) + t_fus
    return n * SEG_SECONDS / dt, n, workers


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals, n_total = [], 0
    for _ in range(max(1, args.warmup > 0)):
        cpu_port_throughput(2.0, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, n, used = cpu_port_throughput(max(4.0, 40.0 / args.steps), cores)
        vals.append(v)
        n_total += n
    wall = time.perf_counter() - t0
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "5 s 16 kHz mono segments, feature body of AudioAnalyzer.analyze + 3-modal fusion forward (BASELINE configs[1] shape), CPU",
                   "segments_per_step": n_total // max(1, args.steps)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": cpu_kind(),
                         "sample": f"{n_total} segments over {args.steps} steps, {cores} processes x 1 torch thread"
                                   + (" (the reference's own source tree)" if cpu_kind() == "reference" else " (oracle/torch_port.py: the reference tree is not on this machine)")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------ GPU arm
def bind_to_gpu_numa(index: int):
    """Pin this process to the CPUs NVML reports as local to GPU `index` BEFORE the pinned host buffers are allocated
    (first touch puts them on that NUMA node): with 4-8 ranks uploading at once, buffers that sit on the other socket
    cross the inter-socket link and cap the box's aggregate host-to-device rate.  Returns a short description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return {"cpus_before": before, "cpus_after": len(os.sched_getaffinity(0)), "how": "nvmlDeviceSetCpuAffinity"}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:80]}


def run_native(args):
    # NCCL prints its "NCCL version ..." banner (any NCCL_DEBUG level from VERSION up) on stdout when the first
    # communicator is created; stdout must carry ONE JSON line.  Send NCCL's log to stderr, and, because a pod may
    # pin NCCL_DEBUG_FILE itself, also point file descriptor 1 at stderr until the first collective has run.
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_numa(local)
    if world > 1:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(8, device=dev)
            dist.all_reduce(warm)                                      # creates the communicator (and prints the banner)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    if not os.path.exists(os.path.join(ROOT, "multimodal-sentiment-analyzer_b200", "libmsa_b200.so")):
        if local == 0:
            entry.build()
        if world > 1:
            dist.barrier()

    import msa_b200
    from msa_b200 import _lib
    from msa_b200.pipeline import ROW_WORDS, gather_rows_async, pack_rows
    from msa_b200 import synth

    S = args.segments
    ana = msa_b200.AudioAnalyzer(device=str(dev))
    model = msa_b200.AdvancedFusionModel(device=str(dev))
    sd = synth.fusion_state(4321, trained_like=True)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})

    pcm_host = torch.from_numpy(synth.fast_segments_pcm(100 + rank, S)).pin_memory()
    face_host = torch.from_numpy(synth.face_rows(200 + rank, S)).pin_memory()
    text_host = torch.from_numpy(synth.text_rows(300 + rank, S)).pin_memory()
    rows_host = torch.empty(S, ROW_WORDS, dtype=torch.float32).pin_memory()
    pcm_dev = pcm_host.to(dev)
    wav_dev = (pcm_dev.float() / 32768.0).contiguous()                 # fp32 waveforms resident in HBM (328 MB > 126 MB L2)
    face_dev, text_dev = face_host.to(dev), text_host.to(dev)
    lib = _lib.lib()

    launches = {"n": 0}
    pending = {"g": None}

    def step_resident():
        row = ana.analyze_batch(wav_dev)
        launches["n"] += lib.msa_last_launch_count()
        logits, amax = model.fused_with_argmax(face_dev, row, text_dev)
        launches["n"] += lib.msa_last_launch_count()
        rows = pack_rows(row, logits, amax, rank * S)
        launches["n"] += lib.msa_last_launch_count()
        # the one collective of the path runs behind this step's kernels on NCCL's stream and overlaps the next step; the
        # gather of the step before is waited for here (stream-level), so every step's table is complete inside the timed region
        pend = gather_rows_async(rows, S * world, world, rank)
        prev, pending["g"] = pending["g"], pend
        return prev.wait() if prev is not None else None

    pipe = msa_b200.SegmentPipeline(ana, model)

    def step_e2e():
        # public host-buffer API: chunked upload on a copy stream overlapped with the kernels, one D2H of the rows
        return pipe.run_host(pcm_host, face_host, text_host, rows_host, first_id=rank * S, chunk=args.chunk)

    def timed(fn, steps, warmup, tail=None):
        for _ in range(warmup):
            fn()
        if tail is not None:
            tail()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if tail is not None:
            tail()                                                     # the last step's gather is inside the timed region
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches["n"] = 0

    def resident_loop_tail():
        if pending["g"] is not None:
            pending["g"].wait()
            pending["g"] = None
    ms_total = timed(step_resident, args.steps, args.warmup, tail=resident_loop_tail)
    launches_timed = launches["n"] * args.steps // max(1, args.steps + args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, args.steps, args.warmup)

    # the upload alone (same pinned buffers, one stream): the floor of the host-buffer path
    pcm_stage = torch.empty_like(pcm_dev)
    def h2d_only():
        pcm_stage.copy_(pcm_host, non_blocking=True)
        face_dev.copy_(face_host, non_blocking=True)
        text_dev.copy_(text_host, non_blocking=True)
    ms_h2d = timed(h2d_only, args.steps, args.warmup) / args.steps
    del pcm_stage

    # dominant kernel alone (the features kernel), same stream, CUDA events
    feat = torch.empty(S, 31, device=dev)
    def feat_only():
        rc = lib.msa_features_f32(_lib.ptr(wav_dev), S, SEG_SAMPLES, None, _lib.ptr(feat), None, None, ana._flags(), 7, 0,
                                  _lib.current_stream_ptr(dev))
        assert rc == 0
    ms_feat = timed(feat_only, args.steps, args.warmup) / args.steps
    def fus_only():
        model.fused_with_argmax(face_dev, feat, text_dev)
    ms_fus = timed(fus_only, args.steps, args.warmup) / args.steps

    # streaming mode (BASELINE configs[4]): 5 s window, 0.5 s hop, one chunk at a time; host-visible latency of
    # push() = 16 KB upload + feature kernel + fusion chain + read-back of the 7 logits
    stream = None
    if rank == 0 and world == 1 and not args.no_streaming:
        sw = msa_b200.StreamingWindow(ana, model, window=SEG_SAMPLES, hop=8000)
        chunks = pcm_host[:64].reshape(-1, 8000)
        face1, text1 = face_host[0], text_host[0]                     # host rows: they ride in the hop's graph like the PCM chunk
        lat = []
        for i in range(args.stream_chunks + 20):
            t0 = time.perf_counter()
            out = sw.push(chunks[i % chunks.shape[0]], face1, text1)
            if out is not None and "done" in out:
                out["done"].synchronize()                                # event behind the 32-byte read-back of this hop
                logits_host = out["host"][:7]                            # pinned
            else:
                torch.cuda.synchronize()
                if out is not None:
                    logits_host = out["fused_emotion"].cpu()
            if i >= 20:
                lat.append((time.perf_counter() - t0) * 1e3)
        # the same hop with the producer writing straight into the window's pinned input buffers (no host-side copies)
        bufs = sw.input_buffers()
        lat2 = []
        for i in range(args.stream_chunks // 2 + 20):
            out["done"].synchronize()
            t0 = time.perf_counter()
            bufs["pcm"].copy_(chunks[i % chunks.shape[0]])           # stands for pipe.readinto(): 16 KB into pinned memory
            out = sw.push_staged(has_text=True)
            out["done"].synchronize()
            logits_host = out["host"][:7]
            if i >= 20:
                lat2.append((time.perf_counter() - t0) * 1e3)
        lat2.sort()
        lat.sort()
        stream = {"p50_ms": lat[len(lat) // 2], "p99_ms": lat[min(len(lat) - 1, int(len(lat) * 0.99))], "chunks": len(lat),
                  "staged_p50_ms": lat2[len(lat2) // 2], "staged_p99_ms": lat2[min(len(lat2) - 1, int(len(lat2) * 0.99))],
                  "window_s": 5.0, "hop_s": 0.5, "timing": "host perf_counter around StreamingWindow.push (one CUDA graph per hop: PCM/face/text upload, feature kernel, fusion chain, logits read-back) + wait on the hop's event"}

    # the reference's own per-segment call (BASELINE configs[0] shape): AudioAnalyzer.analyze(path, speaker) on a 5 s wav
    # file followed by the fusion forward on its row, host wall-clock per call (file read, upload, kernels, read-back)
    api = None
    if rank == 0 and world == 1 and not args.no_streaming:
        import tempfile
        import wave as wave_mod
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "seg.wav")
            with wave_mod.open(path, "wb") as wf:
                wf.setnchannels(1); wf.setsampwidth(2); wf.setframerate(16000)
                wf.writeframes(pcm_host[0].numpy().tobytes())
            lat = []
            for i in range(60):
                t0 = time.perf_counter()
                a = ana.analyze(path, "spk")
                row = msa_b200.assemble_row([a.emotion_probs, a.pitch, a.intensity, a.timbre, a.speech_rate, a.rhythm,
                                             torch.tensor([a.audio_quality, a.signal_noise_ratio, a.clarity, a.consistency], device=dev)])
                out = model(face_dev[:1], row, text_dev[:1])
                out["fused"].argmax(dim=1).item()
                if i >= 10:
                    lat.append((time.perf_counter() - t0) * 1e3)
            lat.sort()
            api = {"p50_ms": lat[len(lat) // 2], "calls": len(lat),
                   "what": "AudioAnalyzer.analyze(wav path) + row assembly + AdvancedFusionModel.forward + argmax, one 5 s segment"}

    # ---- BASELINE configs[2]: one hour of audio = 720 tumbling 5 s segments, STRONG-scaled over the ranks (contiguous shards),
    # one all_gather of the result rows, speaker aggregation on the gathered table (offline_processor.py:255-298)
    hour = None
    if not args.no_extra:
        HS, NSPK = 720, 4
        from msa_b200.pipeline import shard_range, unpack_rows
        hb, he = shard_range(HS, world, rank)
        h_pcm = torch.from_numpy(synth.fast_segments_pcm(900, HS)[hb:he]).to(dev)
        h_face = torch.from_numpy(synth.face_rows(901, HS)[hb:he]).to(dev)
        h_text = torch.from_numpy(synth.text_rows(902, HS)[hb:he]).to(dev)
        spk = torch.from_numpy(np.random.default_rng(903).integers(0, NSPK, HS).astype(np.int32)).to(dev)
        def hour_step():
            table = pipe.run_sharded(h_pcm, h_face, h_text, HS, world, rank)
            return msa_b200.aggregate_speakers(unpack_rows(table)["argmax"], spk, NSPK)
        ms_hour = timed(hour_step, args.steps, args.warmup) / args.steps
        hour = {"segments": HS, "audio_s": HS * SEG_SECONDS, "ms": ms_hour, "value": HS * SEG_SECONDS / (ms_hour / 1000.0), "unit": UNIT,
                "scaling": "strong", "n_gpus": world, "segments_per_gpu": he - hb, "input": "int16 PCM resident in HBM",
                "what": "features + 3-modal fusion per shard, one all_gather of [S/N, 40] rows, speaker aggregation (histogram, dominant label, 3-in-a-row patterns)"}
        del h_pcm, h_face, h_text

    # ---- BASELINE configs[3]: fusion forward only, batch 65,536, synthetic face / audio / text rows (rank 0's GPU at N = 1)
    fus65 = None
    if world == 1 and not args.no_extra:
        FB = 65536
        f_face = torch.from_numpy(synth.face_rows(910, FB)).to(dev)
        f_audio = torch.from_numpy(synth.audio_rows(911, FB)).to(dev)
        f_text = torch.from_numpy(synth.text_rows(912, FB)).to(dev)
        def fus65_step():
            model.fused_with_argmax(f_face, f_audio, f_text)
        ms_f65 = timed(fus65_step, max(3, args.steps // 2), args.warmup) / max(3, args.steps // 2)
        tpeak, tsrc = measured_tensor_peak()
        algo = 9069568.0 * FB / (ms_f65 / 1000.0) / 1e12
        fus65 = {"rows": FB, "ms": ms_f65, "rows_per_s": FB / (ms_f65 / 1000.0), "algorithmic_tflops": algo, "issued_bf16_tflops": 3.0 * algo,
                 "peak": tpeak, "peak_source": tsrc + " (sustained dense bf16)", "frac": algo / tpeak, "frac_issued": 3.0 * algo / tpeak,
                 "note": "algorithmic = 9,069,568 FLOP per row counted once; every product is three bf16 MMAs (split-bf16, fp32 accumulate)"}
        del f_face, f_audio, f_text

    if rank == 0:
        audio_s = S * SEG_SECONDS * world
        value = audio_s * args.steps / (ms_total / 1000.0)
        e2e_v = audio_s * args.steps / (ms_e2e / 1000.0)
        peak, peak_src = measured_peaks()
        achieved = ALGO_BYTES_PER_SEGMENT * S / (ms_feat / 1000.0) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, n, threads = cpu_port_throughput(12.0, 1)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": cpu_kind(),
                   "sample": f"{n} segments, one process, {threads} torch intra-op threads ("
                             + ("the reference's own source tree)" if cpu_kind() == "reference" else "oracle/torch_port.py)")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{S} synthetic 5 s 16 kHz mono segments per GPU: fused feature kernel + 3-modal fusion forward (BASELINE configs[1])",
                       "segments_per_gpu": S, "segment_samples": SEG_SAMPLES, "fusion": "3-modal, split-bf16 tcgen05, fp32 accumulate",
                       "l2": "inputs larger than L2 (328 MB fp32 per GPU vs 126 MB)", "parallelism": f"segments sharded x{world}, one all_gather of result rows"},
            "roofline": {"bound": "hbm", "kernel": "features_kernel<float>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(S, lib.msa_version()), "traffic_unit": "bytes per launch (ncu dram read+write; null unless profiles/features_traffic.json was captured from this build)",
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SEGMENT * S, "peak_source": peak_src,
                         "ms_per_launch": ms_feat, "msa_version": lib.msa_version(),
                         "fp32": {"flop_per_segment": FP32_FLOP_PER_SEGMENT, "achieved_tflops": FP32_FLOP_PER_SEGMENT * S / (ms_feat / 1000.0) / 1e12,
                                  "peak_tflops": FP32_PEAK_TFLOPS, "frac": FP32_FLOP_PER_SEGMENT * S / (ms_feat / 1000.0) / 1e12 / FP32_PEAK_TFLOPS,
                                  "note": "split-radix FLOP count of the path's transforms (626 x 2 real 512-point + 401 real 400-point per segment) against the CUDA-core fp32 peak; the 512-point round trip now runs as fp16 matrix products on the tensor pipe, so this fraction only says how far the kernel is from an ideal fp32 FFT machine"},
                         "note": "not HBM-bound: ~58 FLOP per waveform byte vs a ridge of ~11; see DESIGN.md 4.1 for what bounds it (latency at 16 warps per SM; shared-memory pipe ~57 %, tensor pipe ~27 %, issue slots ~45 %)"},
            "kernels_ms": {"features": ms_feat, "fusion_chain": ms_fus},
            "fusion_tensor": (lambda pk: {"bound": "tensor", "achieved": 9069568.0 * S / (ms_fus / 1000.0) / 1e12, "peak": pk[0], "unit": "TFLOP/s",
                                          "frac": 9069568.0 * S / (ms_fus / 1000.0) / 1e12 / pk[0], "peak_source": pk[1],
                                          "note": "algorithmic FLOP (9,069,568 per row, counted once; three bf16 MMAs are issued per product). "
                                                  "At this batch the 5-launch chain is latency-bound; batch 65536 reaches 248 TFLOP/s "
                                                  "(profiles/r1_v5_fusion_only_times.jsonl)"})(measured_tensor_peak()),
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(pcm_host.numel() * 2 + face_host.numel() * 4 + text_host.numel() * 4) * world,
                    "d2h_bytes_per_step": int(rows_host.numel() * 4) * world, "ms_per_step": ms_e2e / args.steps, "input": "int16 PCM from pinned host memory", "h2d_only_ms": ms_h2d,
                    "pipeline": f"SegmentPipeline.run_host: {args.chunk}-segment chunks, upload overlapped with compute"},
            "hour_sharded": hour,
            "fusion_only_65536": fus65,
            "stream_latency": stream,
            "reference_api_call": api,
            "host_affinity": affinity,
            "gpu_launches": int(launches_timed),
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--segments", type=int, default=1024)
    ap.add_argument("--chunk", type=int, default=128, help="segments per upload chunk of the host-buffer (e2e) path")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-streaming", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the hour_sharded (configs[2]) and fusion_only_65536 (configs[3]) measurements")
    ap.add_argument("--stream-chunks", type=int, default=500)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()

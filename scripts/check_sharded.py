"""BASELINE configs[2] on N GPUs: one hour of synthetic 16 kHz audio cut into 720 tumbling 5 s segments, sharded
over the ranks, ONE NCCL all_gather of the [S/N, 40] result tables, speaker aggregation on every rank.
Run under torchrun; rank 0 also computes all 720 segments alone and checks the gathered table bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/check_sharded.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import msa_b200
from msa_b200 import synth
from msa_b200.pipeline import shard_range, unpack_rows

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

S, T, N_SPK = int(os.environ.get("MSA_CHECK_SEGMENTS", "720")), 80000, 4   # 720 x 5 s = one hour
hour = synth.fast_segments_pcm(11, S).reshape(-1)                    # 57.6 M samples, identical on every rank (seeded)
face, text = synth.face_rows(12, S), synth.text_rows(13, S)
speaker = np.random.default_rng(14).integers(0, N_SPK, S).astype(np.int32)
b, e = shard_range(S, world, rank)

ana = msa_b200.AudioAnalyzer(device=str(dev))
model = msa_b200.AdvancedFusionModel(device=str(dev))
model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.fusion_state(4321, trained_like=True).items()})
pipe = msa_b200.SegmentPipeline(ana, model)

seg = torch.from_numpy(hour[b * T:e * T].reshape(e - b, T)).to(dev)   # this rank holds only its slice of the hour
f, t = torch.from_numpy(face[b:e]).to(dev), torch.from_numpy(text[b:e]).to(dev)
for _ in range(2):                                                   # warm-up: lazy kernel loading, NCCL set-up
    table = pipe.run_sharded(seg, f, t, S, world, rank)
    msa_b200.aggregate_speakers(unpack_rows(table)["argmax"], torch.from_numpy(speaker), N_SPK)
torch.cuda.synchronize(); dist.barrier()
spk_dev = torch.from_numpy(speaker).to(dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
ev[0].record()
rows = pipe.run(seg, f, t, None, first_id=b)                       # this rank's shard: feature kernel + fusion + packing
ev[1].record()
from msa_b200.pipeline import gather_rows
table = gather_rows(rows, S, world, rank)                            # the one collective
ev[2].record()
u = unpack_rows(table)
agg = msa_b200.aggregate_speakers(u["argmax"], spk_dev, N_SPK)
ev[3].record(); torch.cuda.synchronize()
stage_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
ms = torch.tensor([ev[0].elapsed_time(ev[3])] + stage_ms, device=dev, dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
stage_ms = [float(v) for v in ms[1:].tolist()]
ms = ms[:1]

ok = True
if rank == 0:
    full = pipe.run(torch.from_numpy(hour.reshape(S, T)).to(dev), torch.from_numpy(face).to(dev), torch.from_numpy(text).to(dev))
    torch.cuda.synchronize()
    ok = bool(torch.equal(full.view(torch.int32), table.view(torch.int32)))
    hist = np.stack([np.bincount(u["argmax"].cpu().numpy()[speaker == s], minlength=7) for s in range(N_SPK)])
    ok = ok and np.array_equal(agg["hist"].cpu().numpy(), hist) and u["segment_id"].tolist() == list(range(S))
    print(json.dumps({"check": "sharded hour == single GPU (bit-exact), speaker histogram == numpy", "ok": ok, "n_gpus": world,
                      "segments": S, "ms": float(ms.item()), "stage_ms": {"shard": stage_ms[0], "gather": stage_ms[1], "aggregate": stage_ms[2]}, "audio_s_per_s": S * 5.0 / (float(ms.item()) / 1e3),
                      "dominant": agg["dominant"].tolist()}))
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)

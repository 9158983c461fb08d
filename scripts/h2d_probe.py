"""Host -> device upload rate with 1..N ranks uploading AT ONCE (the e2e path's limiter), for three kinds of host memory:
torch pinned (cudaHostAlloc default: what SegmentPipeline.run_host uses), cudaHostAlloc write-combined, and pinned memory
allocated after binding the process to the GPU's NUMA-local CPUs.  One cudaMemcpyAsync per chunk, CUDA events.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/h2d_probe.py
"""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
NBYTES = 160 * 1024 * 1024
dst = torch.empty(NBYTES, dtype=torch.uint8, device=dev)
rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")


def host_alloc(flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(NBYTES), ctypes.c_uint(flags))
    assert rc == 0, rc
    ctypes.memset(p, 1, NBYTES)
    return p


def timed_copy(src_ptr, reps=6):
    stream = torch.cuda.current_stream(dev).cuda_stream
    def go():
        rc = rt.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr()), src_ptr, ctypes.c_size_t(NBYTES), ctypes.c_int(1), ctypes.c_void_p(stream))
        assert rc == 0, rc
    go(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        go()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


out = {"n_ranks": world, "bytes_per_rank": NBYTES}
pinned = torch.empty(NBYTES, dtype=torch.uint8).pin_memory()
out["torch_pinned_ms"] = timed_copy(ctypes.c_void_p(pinned.data_ptr()))
wc = host_alloc(0x04)                                     # cudaHostAllocWriteCombined
out["write_combined_ms"] = timed_copy(wc)
try:
    import pynvml
    pynvml.nvmlInit()
    before = len(os.sched_getaffinity(0))
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    out["cpus_before_after"] = [before, len(os.sched_getaffinity(0))]
    local_pin = host_alloc(0x00)
    out["numa_bound_pinned_ms"] = timed_copy(local_pin)
except Exception as e:  # noqa: BLE001
    out["numa_error"] = str(e)[:100]
for k in list(out):
    if k.endswith("_ms"):
        out[k.replace("_ms", "_GBs_per_rank")] = NBYTES / out[k] / 1e6
        out[k.replace("_ms", "_GBs_aggregate")] = world * NBYTES / out[k] / 1e6
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()

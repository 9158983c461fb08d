import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msa_b200
from msa_b200 import _lib
dev = torch.device("cuda:0"); lib = _lib.lib()
for sr in (48000, 32000, 8000, 44100):
    for B in (1, 16, 256):
        L = 5 * sr
        pcm = torch.randint(-20000, 20000, (B, L), dtype=torch.int16, device=dev)
        n = lib.msa_resample_out_len(L, sr, 16000)
        out = torch.empty(B, n, device=dev)
        for _ in range(2):
            assert lib.msa_resample_s16(_lib.ptr(pcm), B, L, sr, 16000, _lib.ptr(out), n, None) == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(5): lib.msa_resample_s16(_lib.ptr(pcm), B, L, sr, 16000, _lib.ptr(out), n, None)
        e1.record(); torch.cuda.synchronize()
        print(sr, B, "gpu ms %.3f" % (e0.elapsed_time(e1) / 5), "host ms %.3f" % ((time.perf_counter() - t0) / 5 * 1e3))

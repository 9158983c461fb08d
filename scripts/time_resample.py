import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msa_b200
from msa_b200 import synth
dev = torch.device("cuda:0")
for sr in (48000, 44100):
    n = 5 * sr
    pcm = torch.randint(-20000, 20000, (256, n), dtype=torch.int16, device=dev)
    for _ in range(2): msa_b200.resample(pcm, sr, 16000)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): y = msa_b200.resample(pcm, sr, 16000)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"resample {sr} -> 16000, 256 x 5 s int16: {ms:.3f} ms ({256*5/ms*1e3/1e6:.2f} M audio-s/s, {(pcm.numel()*2 + y.numel()*4)/ms/1e6:.0f} GB/s)")

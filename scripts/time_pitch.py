import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msa_b200
from msa_b200 import synth
dev = torch.device("cuda:0")
ana = msa_b200.AudioAnalyzer(device="cuda:0")
wav = torch.from_numpy(synth.pcm_to_f32(synth.fast_segments_pcm(3, 1024))).to(dev)
for _ in range(2): ana.track_pitch(wav)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ana.track_pitch(wav)
e1.record(); torch.cuda.synchronize()
print("track_pitch 1024 x 5 s: %.3f ms" % (e0.elapsed_time(e1) / 5))

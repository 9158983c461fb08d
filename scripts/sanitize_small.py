"""Small invocation of every kernel of the library for compute-sanitizer (one tool per gpurun call):
   compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import msa_b200
from msa_b200 import synth
dev = torch.device("cuda:0")
ana = msa_b200.AudioAnalyzer(device="cuda:0")
m = msa_b200.AdvancedFusionModel(device="cuda:0")
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.fusion_state(4321, trained_like=True).items()})
pipe = msa_b200.SegmentPipeline(ana, m)
for n, T in ((3, 80000), (2, 12345), (1, 513), (2, 30001)):
    pcm = torch.from_numpy(synth.fast_segments_pcm(5, n, T)).to(dev)
    rows = pipe.run(pcm, torch.from_numpy(synth.face_rows(1, n)).to(dev), torch.from_numpy(synth.text_rows(2, n)).to(dev))
    f, d = ana.analyze_batch((pcm.float() / 32768.0).contiguous(), return_detail=True)
    ana.track_pitch(pcm); ana.spectral_descriptors(pcm)
n = 200
pcm = torch.from_numpy(synth.fast_segments_pcm(6, n, 16000)).to(dev)
rows = pipe.run(pcm, torch.from_numpy(synth.face_rows(1, n)).to(dev), torch.from_numpy(synth.text_rows(2, n)).to(dev))
rows2 = pipe.run(pcm, torch.from_numpy(synth.face_rows(1, n)).to(dev), None)
torch.cuda.synchronize()
print("sanitize_small ok", rows.shape, float(rows[:, 31:38].abs().max()))

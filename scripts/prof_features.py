"""Small driver for ncu: a few launches of the feature kernel (and optionally the fusion chain)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import msa_b200
from msa_b200 import _lib
from oracle import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
what = sys.argv[2] if len(sys.argv) > 2 else "features"
cluster = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda:0")
ana = msa_b200.AudioAnalyzer(device="cuda:0")
wav = torch.from_numpy(synth.pcm_to_f32(synth.fast_segments_pcm(3, B))).to(dev)
feat = torch.empty(B, 31, device=dev)
lib = _lib.lib()
if what in ("features", "both"):
    for _ in range(3):
        rc = lib.msa_features_f32(_lib.ptr(wav), B, 80000, None, _lib.ptr(feat), None, None, ana._flags(), 7, cluster, None)
        assert rc == 0
    torch.cuda.synchronize()
if what in ("fusion", "both"):
    m = msa_b200.AdvancedFusionModel(device="cuda:0")
    f = torch.from_numpy(synth.face_rows(1, B)).to(dev); a = torch.from_numpy(synth.audio_rows(2, B)).to(dev); t = torch.from_numpy(synth.text_rows(3, B)).to(dev)
    for _ in range(3):
        m.fused_with_argmax(f, a, t)
    torch.cuda.synchronize()
print("done", B, what)

"""BASELINE configs[3]: fusion forward only, batch 65536 (or argv[1]) with synthetic face/text embeddings.
Prints CUDA-event time per forward and algorithmic TFLOP/s (9,069,568 FLOP per row, counted once although the
split-bf16 path issues three MMAs per product)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import msa_b200
from msa_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
m = msa_b200.AdvancedFusionModel(device="cuda:0")
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.fusion_state(4321, trained_like=True).items()})
f, a, t = (torch.from_numpy(x).to(dev) for x in (synth.face_rows(1, B), synth.audio_rows(2, B), synth.text_rows(3, B)))
for _ in range(3):
    m.fused_with_argmax(f, a, t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    m.fused_with_argmax(f, a, t)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
flop = 9069568.0 * B
print(json.dumps({"config": f"fusion forward only, batch {B}, 3-modal", "ms": ms, "rows_per_s": B / ms * 1e3,
                  "algorithmic_tflops": flop / ms / 1e9, "issued_bf16_tflops": 3 * flop / ms / 1e9}))

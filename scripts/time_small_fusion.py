"""Fusion forward for a handful of rows (the streaming path): matrix-vector kernels (default for batch <= 8)
against the tensor-core chain forced onto the same rows.  Device time per forward: 20 forwards captured in one
CUDA graph (the streaming path replays a graph too), 20 replays between CUDA events."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import msa_b200
from msa_b200 import synth, _lib
dev = torch.device("cuda:0")
m = msa_b200.AdvancedFusionModel(device="cuda:0")
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.fusion_state(4321, trained_like=True).items()})
m._ensure_packed()
for B in (1, 2, 4, 8):
    f, a, t = (torch.from_numpy(x).to(dev) for x in (synth.face_rows(1, B), synth.audio_rows(2, B), synth.text_rows(3, B)))
    lo, am = torch.empty(B, 7, device=dev), torch.empty(B, dtype=torch.int32, device=dev)
    out = {"config": f"fusion forward only, batch {B}, 3-modal"}
    for impl, name in ((0, "matrix_vector_ms"), (2, "tcgen05_ms")):
        assert _lib.lib().msa_fusion_set_impl(impl) == 0
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for _ in range(5):
                m.forward_into(f, a, t, lo, am)
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for _ in range(20):
                    m.forward_into(f, a, t, lo, am)
            g.replay(); st.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(20):
                g.replay()
            e1.record(st); st.synchronize()
        out[name] = e0.elapsed_time(e1) / 400
    _lib.lib().msa_fusion_set_impl(0)
    print(json.dumps(out))

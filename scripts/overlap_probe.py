"""Do the 1024-row fusion chain and the feature kernel of the NEXT batch overlap when they are issued on two streams?
Times K feature launches alone, K fusion forwards alone, and both interleaved on two streams (fusion on a high-priority
stream), with CUDA events around the whole loop.  python scripts/overlap_probe.py"""
import json
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msa_b200  # noqa: E402
from msa_b200 import synth  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    S, T, K = 1024, 80000, 20
    ana = msa_b200.AudioAnalyzer(device="cuda:0")
    model = msa_b200.AdvancedFusionModel().to(dev).eval()
    wav = (torch.rand(S, T, device=dev) - 0.5) * 0.2
    face = torch.rand(S, 27, device=dev)
    text = torch.rand(S, 783, device=dev)
    rows = [torch.zeros(S, 31, device=dev) for _ in range(2)]
    logits = torch.zeros(S, 7, device=dev)
    amax = torch.zeros(S, dtype=torch.int32, device=dev)
    model.prepare(S)
    ana.analyze_into(wav, rows[0])
    model.forward_into(face, rows[0], text, logits, amax)
    torch.cuda.synchronize()
    sa = torch.cuda.Stream(dev)
    sb = torch.cuda.Stream(dev, priority=-1)

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K

    def feats_only():
        for i in range(K):
            ana.analyze_into(wav, rows[i & 1])

    def fusion_only():
        for i in range(K):
            model.forward_into(face, rows[i & 1], text, logits, amax)

    def serial():
        for i in range(K):
            ana.analyze_into(wav, rows[i & 1])
            model.forward_into(face, rows[i & 1], text, logits, amax)

    def pipelined():
        cur = torch.cuda.current_stream(dev)
        sa.wait_stream(cur)
        sb.wait_stream(cur)
        evs = []
        for i in range(K):
            with torch.cuda.stream(sa):
                if i >= 2:
                    sa.wait_event(evs[i - 2][1])                     # the fusion that read this row buffer two steps ago is done
                ana.analyze_into(wav, rows[i & 1])
                ef = torch.cuda.Event()
                ef.record(sa)
            with torch.cuda.stream(sb):
                sb.wait_event(ef)
                model.forward_into(face, rows[i & 1], text, logits, amax)
                eg = torch.cuda.Event()
                eg.record(sb)
            evs.append((ef, eg))
        cur.wait_stream(sa)
        cur.wait_stream(sb)

    out = {}
    for name, fn in (("features_only", feats_only), ("fusion_only", fusion_only), ("serial", serial), ("pipelined_two_streams", pipelined)):
        fn()
        out[name + "_ms_per_step"] = round(min(timed(fn) for _ in range(3)), 4)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

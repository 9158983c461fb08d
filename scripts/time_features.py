"""Time the feature kernel alone (CUDA events) for the current MSA_FEAT_* environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import msa_b200
from msa_b200 import _lib
from oracle import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dtype = sys.argv[2] if len(sys.argv) > 2 else "f32"
dev = torch.device("cuda:0")
ana = msa_b200.AudioAnalyzer(device="cuda:0")
pcm_np = synth.fast_segments_pcm(3, B)
if "pauses" in sys.argv[3:]:          # speech-like pauses: digital silence in 40 % of every segment
    pcm_np[:, 8000:24000] = 0
    pcm_np[:, 48000:64000] = 0
pcm = torch.from_numpy(pcm_np).to(dev)
wav = pcm if dtype == "s16" else (pcm.float() / 32768.0).contiguous()
feat = torch.empty(B, 31, device=dev)
lib = _lib.lib()
use_ws = "ws" in sys.argv[3:]
parts = 7
for a in sys.argv[3:]:
    if a.startswith("parts="):
        parts = int(a[6:])                                  # 1 wave statistics, 2 MFCC, 4 STFT-512 -> ISTFT residual
ws = torch.empty(max(1, lib.msa_features_workspace_bytes(B, 80000)), dtype=torch.uint8, device=dev) if use_ws else None
fn = (lib.msa_features_ws_s16 if dtype == "s16" else lib.msa_features_ws_f32) if use_ws else (lib.msa_features_s16 if dtype == "s16" else lib.msa_features_f32)
def run():
    if use_ws:
        rc = fn(_lib.ptr(wav), B, 80000, None, _lib.ptr(feat), None, None, ana._flags(), parts, 0, _lib.ptr(ws), ws.numel(), None)
    else:
        rc = fn(_lib.ptr(wav), B, 80000, None, _lib.ptr(feat), None, None, ana._flags(), parts, 0, None)
    assert rc == 0, rc
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"threads={os.environ.get('MSA_FEAT_THREADS','dflt')} cluster={lib.msa_features_cluster_size(80000)} "
      f"B={B} {dtype}{' pauses' if 'pauses' in sys.argv[3:] else ''}{' ws' if use_ws else ''}{' parts=%d' % parts if parts != 7 else ''}: {ms:.3f} ms  -> {B*5/ms*1e3/1e6:.2f} M audio-s/s, {320124*B/ms/1e6:.1f} GB/s algorithmic")

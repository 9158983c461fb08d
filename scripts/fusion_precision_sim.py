"""What operand precision does the fusion chain need?  fp64 oracle (oracle/fusion_np.py) against the same network with the
INPUT of every Linear rounded to one fp16 plane (two MMAs per product with split weights) or to bf16 hi + lo (what the
kernels do, three MMAs per product).  CPU only: python scripts/fusion_precision_sim.py
20,000 rows: one fp16 plane -> max |dlogit| 2.2e-3, rms 4.3e-4 (over the 1e-3 bar); bf16 hi + lo -> max 2.4e-5."""
import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fusion_np as fu, synth
def q(x, mode):
    if mode=='f16': return x.astype(np.float16).astype(np.float64)
    if mode=='bf16x2':  # hi+lo bf16 ~ 16 bits
        import struct
        x32=x.astype(np.float32); b=x32.view(np.uint32); hi=((b+0x7FFF+((b>>16)&1))&0xFFFF0000).view(np.float32)
        lo=(x32-hi); bl=lo.view(np.uint32); lo2=((bl+0x7FFF+((bl>>16)&1))&0xFFFF0000).view(np.float32)
        return (hi.astype(np.float64)+lo2.astype(np.float64))
    if mode=='f16x2':
        h=x.astype(np.float16).astype(np.float64); l=(x-h).astype(np.float16).astype(np.float64); return h+l
    return x
def lin(x, sd, name, mode):
    return q(x,mode) @ sd[name+'.weight'].astype(np.float64).T + sd[name+'.bias'].astype(np.float64)
def norm(x, sd, name): return fu._ln(x, sd[name+'.weight'].astype(np.float64), sd[name+'.bias'].astype(np.float64))
def branch(x, sd, mod, mode):
    h=lin(norm(x,sd,mod+'_norm'),sd,mod+'_proj',mode)
    h=np.maximum(norm(h,sd,mod+'_processor.0'),0.0)
    h=lin(h,sd,mod+'_processor.3',mode)
    return np.maximum(norm(h,sd,mod+'_processor.4'),0.0)
def fuse(sd,f,a,t,mode):
    h=np.concatenate([branch(f.astype(np.float64),sd,'face',mode),branch(a.astype(np.float64),sd,'audio',mode),branch(t.astype(np.float64),sd,'text',mode)],1)
    h=lin(h,sd,'fusion.0',mode); h=np.maximum(norm(h,sd,'fusion.1'),0.0)
    h=lin(h,sd,'fusion.4',mode); h=np.maximum(norm(h,sd,'fusion.5'),0.0)
    return h @ sd['fusion.8.weight'].astype(np.float64).T + sd['fusion.8.bias'].astype(np.float64)   # last layer fp32 on CUDA cores
n=20000
f,a,t=synth.face_rows(1,n),synth.audio_rows(2,n),synth.text_rows(3,n)
for trained in (False,True):
    sd=synth.fusion_state(4321,trained_like=trained)
    ref=fuse(sd,f,a,t,'exact')
    for mode in ('f16','bf16x2'):
        got=fuse(sd,f,a,t,mode)
        d=np.abs(got-ref)
        srt=np.sort(ref,1); margin=srt[:,-1]-srt[:,-2]
        print('trained' if trained else 'init', mode, 'max',d.max(), 'p99.9',np.quantile(d,0.999), 'rms',np.sqrt((d**2).mean()), 'argmax agree',(got.argmax(1)==ref.argmax(1)).mean(), 'logit scale',np.abs(ref).mean())

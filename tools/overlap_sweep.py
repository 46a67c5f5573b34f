import sys, json, subprocess
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from jxlatte_b200 import synth, default_frame_params, _lib
from jxlatte_b200.host import Reconstructor, qm_generate
W,H=7680,4320
p=default_frame_params(W,H,epf_iters=3); qw,qo=qm_generate()
st=synth.make_state(W,H,seed=0x4A584C02,params=p,qm_weights=qw,qm_offsets=qo)
dev=torch.device('cuda:0'); stream=torch.cuda.Stream(); torch.cuda.set_stream(stream)
r=Reconstructor(0); r.set_stream(stream.cuda_stream); r.setWeights(qw,qo)
d={k:torch.from_numpy(np.ascontiguousarray(st[k])).to(dev) for k in ("qcoeff","lf","dct_select","block_origin","hf_mul","sharpness","x_from_y","b_from_y")}
out=torch.empty((3,H,W),dtype=torch.float32,device=dev)
def run():
    r.reconstruct_dev(p,[d["qcoeff"][c].data_ptr() for c in range(3)],[d["lf"][c].data_ptr() for c in range(3)],d["dct_select"].data_ptr(),d["block_origin"].data_ptr(),d["hf_mul"].data_ptr(),d["x_from_y"].data_ptr(),d["b_from_y"].data_ptr(),d["sharpness"].data_ptr(),[out[c].data_ptr() for c in range(3)])
ref=None
for rows in (0,512,768,1024,1536,2048):
    r.set_option(_lib.OPT_OVERLAP_ROWS, rows)
    for _ in range(3): run()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(10): run()
    b.record(stream); torch.cuda.synchronize()
    o=out.clone()
    if ref is None: ref=o
    print(rows, '%.3f ms'%(a.elapsed_time(b)/10), 'identical' if torch.equal(o,ref) else 'DIFFERENT')

"""Per-slab event timeline of the pipelined host entry point (JXLB200_TIMELINE=1): where the milliseconds of one 8K call go.
    python tools/e2e_timeline.py 2> gpurun_out/timeline.txt"""
import os, sys, time
os.environ["JXLB200_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from jxlatte_b200.host import Reconstructor

W, H = 7680, 4320
p, st, qw, qo = bench.make_inputs(W, H, 0x4A584C00 + 2, 3)
rec = Reconstructor(0)
rec.setWeights(qw, qo)
keys = ("qcoeff", "lf", "dct_select", "block_origin", "hf_mul", "sharpness", "x_from_y", "b_from_y")
hst = {k: torch.from_numpy(np.ascontiguousarray(st[k])).pin_memory() for k in keys}
hnp = {k: v.numpy() for k, v in hst.items()}
hout = torch.empty((3, H, W), dtype=torch.float32).pin_memory().numpy()
h16 = dict(hnp)
q16 = torch.from_numpy(np.ascontiguousarray(st["qcoeff"]).astype(np.int16)).pin_memory()
h16["qcoeff"] = q16.numpy()
hpk = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory().numpy()
for name, fn in (("int32 in, float32 out", lambda: rec.reconstruct(p, hnp, out=hout)),
                 ("int16 in, float32 out", lambda: rec.reconstruct(p, h16, out=hout, narrow=True)),
                 ("int16 in, 8-bit packed out", lambda: rec.reconstruct_packed(p, h16, bits=8, out=hpk, narrow=True))):
    for i in range(3):
        sys.stderr.write("==== %s, call %d\n" % (name, i)); sys.stderr.flush()
        t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
        sys.stderr.write("==== wall %.3f ms\n" % (dt * 1e3)); sys.stderr.flush()
# plain copies for the bus rates
a = torch.empty(210_000_000, dtype=torch.uint8).pin_memory(); d = torch.empty(210_000_000, dtype=torch.uint8, device="cuda")
for nm, f in (("H2D 210 MB", lambda: d.copy_(a, non_blocking=True)), ("D2H 210 MB", lambda: a.copy_(d, non_blocking=True))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): f()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    sys.stderr.write("==== %s: %.3f ms = %.1f GB/s\n" % (nm, dt * 1e3, 0.21 / dt))

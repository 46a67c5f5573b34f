"""Slab height / stage-1 fan-out sweep of the packed host entry point (8K frame, int16 in, 8-bit samples out), and of the float32 one.
    python tools/packed_sweep.py > gpurun_out/packed_sweep.txt"""
import os, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from jxlatte_b200 import _lib
from jxlatte_b200.host import Reconstructor

W, H = 7680, 4320
p, st, qw, qo = bench.make_inputs(W, H, 0x4A584C00 + 2, 3)
rec = Reconstructor(0)
rec.setWeights(qw, qo)
keys = ("qcoeff", "lf", "dct_select", "block_origin", "hf_mul", "sharpness", "x_from_y", "b_from_y")
hst = {k: torch.from_numpy(np.ascontiguousarray(st[k])).pin_memory() for k in keys}
hnp = {k: v.numpy() for k, v in hst.items()}
h16 = dict(hnp)
q16 = torch.from_numpy(np.ascontiguousarray(st["qcoeff"]).astype(np.int16)).pin_memory()
h16["qcoeff"] = q16.numpy()
hpk = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory().numpy()
hout = torch.empty((3, H, W), dtype=torch.float32).pin_memory().numpy()

def wall(fn, n=8):
    for _ in range(2):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3

for rows in (512, 768, 1024, 1536, 2304, 4352):
    for fan in (0, 1):
        rec.set_option(_lib.OPT_PIPE_ROWS, rows)
        rec.set_option(_lib.OPT_PIPE_FANOUT, fan)
        a = wall(lambda: rec.reconstruct_packed(p, h16, bits=8, out=hpk, narrow=True))
        crc = zlib.crc32(hpk.tobytes())
        b = wall(lambda: rec.reconstruct(p, h16, out=hout, narrow=True))
        c = wall(lambda: rec.reconstruct(p, hnp, out=hout))
        print("rows %4d fanout %d: png8 %.3f ms (crc %08x)  int16->f32 %.3f ms  int32->f32 %.3f ms" % (rows, fan, a, crc, b, c), flush=True)

"""Device-resident step time of one 8K frame for the library named by JXLB200_LIB (kernel-variant experiments), plus a
checksum of the planes so that variants can be held to the default build's bits."""
import sys, os, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from jxlatte_b200.host import Reconstructor

W, H = 7680, 4320
p, st, qw, qo = bench.make_inputs(W, H, 0x4A584C00 + 2, 3)
dev = torch.device("cuda", 0)
rec = Reconstructor(0)
s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s); rec.set_stream(s.cuda_stream)
rec.setWeights(qw, qo)
if os.environ.get("JXLB200_STAGE2"):
    from jxlatte_b200 import _lib
    rec.set_option(_lib.OPT_STAGE2, int(os.environ["JXLB200_STAGE2"]))
d = {k: torch.from_numpy(np.ascontiguousarray(st[k])).to(dev) for k in ("qcoeff", "lf", "dct_select", "block_origin", "hf_mul", "sharpness", "x_from_y", "b_from_y")}
out = torch.empty((3, H, W), dtype=torch.float32, device=dev)
xyb = torch.empty((3, H, W), dtype=torch.float32, device=dev)
q = [d["qcoeff"][c].data_ptr() for c in range(3)]; lf = [d["lf"][c].data_ptr() for c in range(3)]
o = [out[c].data_ptr() for c in range(3)]; x = [xyb[c].data_ptr() for c in range(3)]
m = [d[k].data_ptr() for k in ("dct_select", "block_origin", "hf_mul", "x_from_y", "b_from_y", "sharpness")]

def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(n):
        fn()
    b.record(s)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n

step = timed(lambda: rec.reconstruct_dev(p, q, lf, *m, o))
rec.invert_dev(p, q, lf, m[0], m[1], m[2], m[3], m[4], x, W)
t2 = timed(lambda: rec.restore_dev(p, None, x, W, m[2], m[5], o))
rec.sync()
crc = zlib.crc32(out.cpu().numpy().tobytes())
# host-buffer entry point (pinned), int32 and int16 coefficients
if os.environ.get('SKIP_HOST'):
    print('%s step %.3f ms stage2 %.3f ms crc %08x' % (os.path.basename(os.environ.get('JXLB200_LIB', 'default')), step, t2, crc)); sys.exit(0)
import time
hst = {k: torch.from_numpy(np.ascontiguousarray(st[k])).pin_memory().numpy() for k in ("qcoeff", "lf", "dct_select", "block_origin", "hf_mul", "sharpness", "x_from_y", "b_from_y")}
hout = torch.empty((3, H, W), dtype=torch.float32).pin_memory().numpy()
rec.set_stream(None)
def host(narrow, n=8):
    for _ in range(2):
        rec.reconstruct(p, hst, out=hout, narrow=narrow)
    t0 = time.perf_counter()
    for _ in range(n):
        rec.reconstruct(p, hst, out=hout, narrow=narrow)
    return (time.perf_counter() - t0) / n * 1e3
e32 = host(False)
crc_h = zlib.crc32(hout.tobytes())
hst["qcoeff"] = torch.from_numpy(hst["qcoeff"].astype(np.int16)).pin_memory().numpy()
e16 = host(True)
crc_h16 = zlib.crc32(hout.tobytes())
print("%s step %.3f ms stage2 %.3f ms crc %08x | host i32 %.3f ms crc %08x, i16 %.3f ms crc %08x" % (os.path.basename(os.environ.get("JXLB200_LIB", "default")), step, t2, crc, e32, crc_h, e16, crc_h16))

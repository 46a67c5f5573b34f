"""Stage-2 time of the two bit-exact fused kernels for every (gab, epf_iters) and a few frame shapes: JXLB200_OPT_STAGE2 = 0 (the
library's choice), 5 (k2_stream forced), 6 (the tile kernel k2_exact forced); one line per case with the three times and whether the
planes agree bit for bit.  The dispatch rule in csrc/jxlb200.cu (stream_pays) is read off this table."""
import os, sys, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jxlatte_b200 import _lib, default_frame_params
from jxlatte_b200.host import Reconstructor

dev = torch.device("cuda", 0)
rec = Reconstructor(0)
s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s); rec.set_stream(s.cuda_stream)

def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(n): fn()
    b.record(s); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

g = torch.Generator(device=dev); g.manual_seed(5)
for (W, H) in ((7680, 4320), (2048, 2048), (1280, 720), (512, 512)):
    xyb = (torch.rand((3, H, W), device=dev, generator=g) * 0.2 + 0.4) * torch.tensor([0.05, 1.0, 1.0], device=dev)[:, None, None]
    hm = torch.randint(1, 5, (H // 8, W // 8), device=dev, dtype=torch.int32, generator=g)
    sh = torch.randint(0, 8, (H // 8, W // 8), device=dev, dtype=torch.int32, generator=g)
    out = torch.empty_like(xyb)
    x = [xyb[c].data_ptr() for c in range(3)]; o = [out[c].data_ptr() for c in range(3)]
    for gab in (1, 0):
        for it in (3, 2, 1):
            p = default_frame_params(W, H, epf_iters=it, gab=bool(gab))
            res = {}
            for opt in (_lib.STAGE2_AUTO, _lib.STAGE2_STREAM, _lib.STAGE2_TILE):
                rec.set_option(_lib.OPT_STAGE2, opt)
                t = timed(lambda: rec.restore_dev(p, None, x, W, hm.data_ptr(), sh.data_ptr(), o))
                rec.sync()
                res[opt] = (t, zlib.crc32(out.cpu().numpy().tobytes()))
            rec.set_option(_lib.OPT_STAGE2, _lib.STAGE2_AUTO)
            a, f, b = res[_lib.STAGE2_AUTO], res[_lib.STAGE2_STREAM], res[_lib.STAGE2_TILE]
            print("%5dx%-5d gab %d epf %d: default %.3f ms  stream %.3f ms  tile %.3f ms  default/tile %.2f  stream/tile %.2f  same bits %s"
                  % (W, H, gab, it, a[0], f[0], b[0], a[0] / b[0], f[0] / b[0], a[1] == b[1] and f[1] == b[1]), flush=True)

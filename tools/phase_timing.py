"""Where a k2_stream tick goes: cycles warp 1 of every CTA spends working in each stage and waiting at the barrier that ends it.
Needs a diagnostic build:  JXLB_SO=$PWD/jxlatte_b200/libjxlb200_clk.so JXLB_EXTRA_FLAGS=-DK2S_PHASE_TIMING python -m jxlatte_b200.build --force
then  JXLB200_LIB=$PWD/jxlatte_b200/libjxlb200_clk.so python tools/phase_timing.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jxlatte_b200 import _lib, default_frame_params
from jxlatte_b200.host import Reconstructor

W, H = 7680, 4320
dev = torch.device("cuda", 0)
rec = Reconstructor(0)
lib = C.CDLL(os.environ["JXLB200_LIB"])
g = torch.Generator(device=dev); g.manual_seed(5)
xyb = (torch.rand((3, H, W), device=dev, generator=g) * 0.2 + 0.4) * torch.tensor([0.05, 1.0, 1.0], device=dev)[:, None, None]
hm = torch.randint(1, 5, (H // 8, W // 8), device=dev, dtype=torch.int32, generator=g)
sh = torch.randint(0, 8, (H // 8, W // 8), device=dev, dtype=torch.int32, generator=g)
out = torch.empty_like(xyb)
x = [xyb[c].data_ptr() for c in range(3)]; o = [out[c].data_ptr() for c in range(3)]
p = default_frame_params(W, H, epf_iters=3, gab=True)
buf = (C.c_ulonglong * 16)()
for _ in range(2):
    rec.restore_dev(p, None, x, W, hm.data_ptr(), sh.data_ptr(), o); rec.sync()
lib.jxlb200_debug_phase_clocks(buf)
n = 5
for _ in range(n):
    rec.restore_dev(p, None, x, W, hm.data_ptr(), sh.data_ptr(), o)
rec.sync()
lib.jxlb200_debug_phase_clocks(buf)
v = np.array(list(buf), dtype=np.float64) / n / 148
names = ["G", "D0", "W0", "D1", "W1", "P2", "TMA wait"]
tot = v.sum()
print("cycles per CTA per launch: %.0f (%.3f ms at 1.9 GHz)" % (tot, tot / 1.9e6))
for i, nm in enumerate(names):
    print("%-9s work %5.1f%%  barrier wait %5.1f%%" % (nm, 100 * v[2 * i] / tot, 100 * v[2 * i + 1] / tot))

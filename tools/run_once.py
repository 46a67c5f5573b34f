"""One 8K frame through the device path a few times (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from jxlatte_b200.host import Reconstructor

W, H = (7680, 4320) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
iters = 3
p, st, qw, qo = bench.make_inputs(W, H, 0x4A584C00 + 2, iters)
dev = torch.device("cuda", 0)
rec = Reconstructor(0)
s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s); rec.set_stream(s.cuda_stream)
rec.setWeights(qw, qo)
d = {k: torch.from_numpy(np.ascontiguousarray(st[k])).to(dev) for k in ("qcoeff", "lf", "dct_select", "block_origin", "hf_mul", "sharpness", "x_from_y", "b_from_y")}
out = torch.empty((3, H, W), dtype=torch.float32, device=dev)
for _ in range(3):
    rec.reconstruct_dev(p, [d["qcoeff"][c].data_ptr() for c in range(3)], [d["lf"][c].data_ptr() for c in range(3)],
                        d["dct_select"].data_ptr(), d["block_origin"].data_ptr(), d["hf_mul"].data_ptr(),
                        d["x_from_y"].data_ptr(), d["b_from_y"].data_ptr(), d["sharpness"].data_ptr(), [out[c].data_ptr() for c in range(3)])
rec.sync()
print("ok", float(out.abs().max()))

"""The price of bit-exactness in stage 2, measured: for the 8K configuration, stage-2 time and the
error of the tolerance mode (JXLB200_OPT_STAGE2 = 2, k2_fused: re-associated / FMA-contracted EPF sums) against the bit-exact
default -- max abs error on the linear planes and the number of pixels whose sRGB-quantised sample moves by more than 1 LSB
at 8 and at 16 bits.  Writes one JSON object (gpurun_out/r2_exact_vs_fast.json); profiles/r2_exact_vs_fast.md is made from it.

    python tools/exact_vs_fast.py [W H]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from jxlatte_b200 import _lib
from jxlatte_b200.host import Reconstructor

W, H = (7680, 4320) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
dev = torch.device("cuda", 0)
rec = Reconstructor(0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
rec.set_stream(stream.cuda_stream)
res = {"frame": [W, H], "modes": {}}
for iters in (3, 1):
    p, st, qw, qo = bench.make_inputs(W, H, 0x4A584C00 + 2, iters)
    rec.setWeights(qw, qo)
    d = {k: torch.from_numpy(np.ascontiguousarray(st[k])).to(dev) for k in ("qcoeff", "lf", "dct_select", "block_origin", "hf_mul", "sharpness", "x_from_y", "b_from_y")}
    xyb = torch.empty((3, H, W), dtype=torch.float32, device=dev)
    outs = {}
    rec.invert_dev(p, [d["qcoeff"][c].data_ptr() for c in range(3)], [d["lf"][c].data_ptr() for c in range(3)], d["dct_select"].data_ptr(),
                   d["block_origin"].data_ptr(), d["hf_mul"].data_ptr(), d["x_from_y"].data_ptr(), d["b_from_y"].data_ptr(), [xyb[c].data_ptr() for c in range(3)], W)
    rec.sync()

    def srgb_q(lin, bits):
        a = torch.where(lin <= 0.0031308, lin * 12.92, 1.055 * torch.clamp(lin, min=0).pow(1.0 / 2.4) - 0.055)
        mx = (1 << bits) - 1
        return torch.clamp((a * mx + 0.5).to(torch.int64), 0, mx)

    for name, opt in (("exact", _lib.STAGE2_AUTO), ("staged", _lib.STAGE2_STAGED), ("tolerance", _lib.STAGE2_FUSED)):
        rec.set_option(_lib.OPT_STAGE2, opt)
        out = torch.empty((3, H, W), dtype=torch.float32, device=dev)

        def go():
            rec.restore_dev(p, None, [xyb[c].data_ptr() for c in range(3)], W, d["hf_mul"].data_ptr(), d["sharpness"].data_ptr(), [out[c].data_ptr() for c in range(3)])

        for _ in range(3):
            go()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(10):
            go()
        b.record(stream)
        torch.cuda.synchronize()
        outs[name] = (out, a.elapsed_time(b) / 10)
    rec.set_option(_lib.OPT_STAGE2, _lib.STAGE2_AUTO)
    ex = outs["exact"][0]
    for name in ("exact", "staged", "tolerance"):
        o, ms = outs[name]
        diff = (o - ex).abs()
        rec_ = {"stage2_ms": ms, "max_abs_err_linear": float(diff.max().item()), "bit_identical_to_exact": bool(torch.equal(o, ex))}
        for bits in (8, 16):
            dq = (srgb_q(o, bits) - srgb_q(ex, bits)).abs().amax(dim=0)
            rec_["px_moved_at_%d_bits" % bits] = int((dq > 0).sum().item())
            rec_["px_over_1_lsb_at_%d_bits" % bits] = int((dq > 1).sum().item())
            rec_["max_lsb_at_%d_bits" % bits] = int(dq.max().item())
        res["modes"]["epf%d_%s" % (iters, name)] = rec_
    del d, xyb, outs
    torch.cuda.empty_cache()
res["pixels"] = W * H
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r2_exact_vs_fast.json", "w"), indent=1)

"""Per-stage error of the CUDA path against the oracle (diagnostic; run on a GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from jxlatte_b200 import synth, default_frame_params
from jxlatte_b200.host import Reconstructor, qm_generate
from jxlatte_b200.params import TRANSFORM_NAMES
from oracle import oracle as O

def srgb_q(lin, bits):
    lin = lin.astype(np.float32)
    a = np.where(lin <= np.float32(0.0031308), lin * np.float32(12.92),
                 np.float32(1.055) * np.power(np.maximum(lin, 0), np.float32(1.0 / 2.4), dtype=np.float32) - np.float32(0.055))
    mx = (1 << bits) - 1
    return np.clip((a.astype(np.float32) * np.float32(mx) + np.float32(0.5)).astype(np.int64), 0, mx)

W, H, iters = 1280, 720, 1
p = default_frame_params(W, H, epf_iters=iters, gab=True)
qw, qo = qm_generate()
st = synth.make_state(W, H, seed=11 + W, params=p, qm_weights=qw, qm_offsets=qo)
r = Reconstructor(0); r.setWeights(qw, qo)
x_ref = O.vardct_invert(p, st, nthreads=8); x_got = r.invertVarDCT(p, st)
e = np.abs(x_got - x_ref)
print("stage1 max err per channel", e.reshape(3, -1).max(1))
# error by transform type
ds_px = np.kron(st["dct_select"], np.ones((8, 8), np.uint8))
for t in np.unique(ds_px):
    m = ds_px == t
    print("  type %-10s  maxerr X %.2e Y %.2e B %.2e" % (TRANSFORM_NAMES[t], e[0][m].max(), e[1][m].max(), e[2][m].max()))
g_ref = O.gab(p, x_ref); g_got = r.performGabConvolution(p, x_ref)
print("gab (same input) max err", np.abs(g_got - g_ref).max())
f_ref = O.epf(p, g_ref, st["hf_mul"], st["sharpness"], nthreads=8); f_got = r.performEdgePreservingFilter(p, g_ref, st["hf_mul"], st["sharpness"])
print("epf (same input) max err", np.abs(f_got - f_ref).max())
c_ref = O.color(p, f_ref); c_got = r.performColorTransforms(p, f_ref)
print("color (same input) max err", np.abs(c_got - c_ref).max())
full_ref = O.vardct_reconstruct(p, st, nthreads=8); full = r.reconstruct(p, st)
err = np.abs(full - full_ref)
print("full max err", err.max())
for bits in (8, 16):
    d = np.abs(srgb_q(full, bits) - srgb_q(full_ref, bits))
    print(bits, "bit: max LSB diff", d.max(), "count>1:", int((d > 1).sum()))
    if d.max() > 1:
        for (c, y, x) in np.argwhere(d > 1)[:8]:
            print("   c%d y%d x%d lin ref %.8g got %.8g  xyb ref %s got %s type %s" % (c, y, x, full_ref[c, y, x], full[c, y, x],
                  x_ref[:, y, x], x_got[:, y, x], TRANSFORM_NAMES[ds_px[y, x]]))

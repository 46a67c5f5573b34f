"""Fused vs staged stage 2 on the same stage-1 planes (diagnostic; run on a GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from jxlatte_b200 import synth, default_frame_params, _lib
from jxlatte_b200.host import Reconstructor, qm_generate

W, H = 1024, 768
qw, qo = qm_generate()
r = Reconstructor(0); r.setWeights(qw, qo)
for iters, gab in ((3, True), (1, True), (2, False)):
    p = default_frame_params(W, H, epf_iters=iters, gab=gab, color_mode=0)
    st = synth.make_state(W, H, seed=11 + W, params=p, qm_weights=qw, qm_offsets=qo)
    r.set_option(_lib.OPT_STAGE2, _lib.STAGE2_STAGED); a = r.reconstruct(p, st)
    r.set_option(_lib.OPT_STAGE2, _lib.STAGE2_FUSED); b = r.reconstruct(p, st)
    d = np.abs(a - b)
    print("iters", iters, "gab", gab, "XYB-domain max diff per channel", d.reshape(3, -1).max(1), "rel to range", d.reshape(3,-1).max(1) / np.abs(a).reshape(3,-1).max(1))
    for c in range(3):
        y, x = np.unravel_index(np.argmax(d[c]), d[c].shape)
        print("   c%d worst at y=%d x=%d (y%%32=%d x%%64=%d) staged %.9g fused %.9g" % (c, y, x, y % 32, x % 64, a[c, y, x], b[c, y, x]))
    # error histogram
    print("   frac > 1e-7:", (d > 1e-7).mean(), " > 1e-6:", (d > 1e-6).mean())
    p.color_mode = 1
    r.set_option(_lib.OPT_STAGE2, _lib.STAGE2_STAGED); a = r.reconstruct(p, st)
    r.set_option(_lib.OPT_STAGE2, _lib.STAGE2_FUSED); b = r.reconstruct(p, st)
    print("   linear max diff", np.abs(a - b).max())

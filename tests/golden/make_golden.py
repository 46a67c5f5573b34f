"""Regenerates tests/golden/*.npz: small inputs with the CPU oracle's outputs.

The reference (jxlatte) ships no golden vectors and cannot run here (no JVM), so these are vectors of the RESTATEMENT, not
of the Java: they pin the oracle against drift (compiler, libm, box) and let the GPU tests check the CUDA path against
committed bytes as well as against the oracle built on the box.   Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from jxlatte_b200 import synth, default_frame_params  # noqa: E402
from jxlatte_b200.host import qm_generate  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    qw, qo = qm_generate()
    # 1. one 8x8 block per 8x8-class type + one 16x8 / 32x32 block: dequantised coefficients -> pixels
    rng = np.random.default_rng(20261017)
    blocks = {}
    for t in (0, 1, 2, 3, 12, 13, 14, 15, 16, 17, 6, 5):
        info = O.tt_info(t)
        c = (rng.standard_normal((info["pixel_h"], info["pixel_w"])) * 0.1).astype(np.float32)
        blocks["in_%d" % t] = c
        blocks["out_%d" % t] = O.invert_varblock(c, t)
    np.savez_compressed(os.path.join(HERE, "varblocks.npz"), **blocks)
    # 2. a 128 x 64 frame, mixed small/medium partition, full path with gab + EPF 3 + XYB->linear
    W, H = 128, 64
    p = default_frame_params(W, H, epf_iters=3)
    st = synth.make_state(W, H, seed=synth.SEED_BASE + 9, mix="small", params=p, qm_weights=qw, qm_offsets=qo)
    xyb = O.vardct_invert(p, st)
    out = O.vardct_reconstruct(p, st)
    np.savez_compressed(os.path.join(HERE, "frame_128x64.npz"), qcoeff=st["qcoeff"], lf=st["lf"], dct_select=st["dct_select"],
                        block_origin=st["block_origin"], hf_mul=st["hf_mul"], sharpness=st["sharpness"],
                        x_from_y=st["x_from_y"], b_from_y=st["b_from_y"], xyb=xyb, out=out)
    # 3. QM tables: digest of the 394752 default weights + spot values
    import hashlib
    np.savez_compressed(os.path.join(HERE, "qm.npz"), sha256=np.frombuffer(hashlib.sha256(qw.tobytes()).digest(), np.uint8),
                        offsets=qo, head=qw[:64], tail=qw[-64:])
    # 4. modular
    x = rng.integers(-300, 300, size=(19, 23)).astype(np.int32)
    ah, rh = O.modular_forward_squeeze(x, True)
    av, rv = O.modular_forward_squeeze(x, False)
    ch = rng.integers(-1000, 1000, size=(3, 9, 11)).astype(np.int32)
    pal = rng.integers(0, 256, size=(3, 12)).astype(np.int32)
    idx = rng.integers(-40, 90, size=(13, 17)).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "modular.npz"), x=x, ah=ah, rh=rh, av=av, rv=rv, ch=ch,
                        rct=np.stack([O.modular_rct(ch, t) for t in range(42)]), pal=pal, idx=idx,
                        pal_out=O.modular_palette(idx, pal, 5, 5, 8))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()

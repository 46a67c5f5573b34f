"""Committed golden vectors (tests/golden/, made by make_golden.py from the CPU oracle).  CPU: the oracle built on this
box reproduces them bit for bit.  GPU: the CUDA path reproduces them bit for bit (staged stage 2)."""
import hashlib
import os

import numpy as np
import pytest

from jxlatte_b200 import default_frame_params
from jxlatte_b200.host import qm_generate

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _frame():
    z = np.load(os.path.join(G, "frame_128x64.npz"))
    st = {k: z[k] for k in ("qcoeff", "lf", "dct_select", "block_origin", "hf_mul", "sharpness", "x_from_y", "b_from_y")}
    st["qm_weights"], st["qm_offsets"] = qm_generate()
    return default_frame_params(128, 64, epf_iters=3), st, z["xyb"], z["out"]


def test_oracle_reproduces_golden_varblocks(orc):
    z = np.load(os.path.join(G, "varblocks.npz"))
    for t in (0, 1, 2, 3, 12, 13, 14, 15, 16, 17, 6, 5):
        assert np.array_equal(orc.invert_varblock(z["in_%d" % t], t), z["out_%d" % t]), t


def test_oracle_reproduces_golden_frame(orc):
    p, st, xyb, out = _frame()
    assert np.array_equal(orc.vardct_invert(p, st), xyb)
    assert np.array_equal(orc.vardct_reconstruct(p, st), out)


def test_qm_tables_match_golden_digest():
    z = np.load(os.path.join(G, "qm.npz"))
    w, off = qm_generate()
    assert np.array_equal(off, z["offsets"]) and np.array_equal(w[:64], z["head"]) and np.array_equal(w[-64:], z["tail"])
    assert hashlib.sha256(w.tobytes()).digest() == z["sha256"].tobytes()


def test_oracle_reproduces_golden_modular(orc):
    z = np.load(os.path.join(G, "modular.npz"))
    assert np.array_equal(orc.modular_squeeze(z["ah"], z["rh"], True), z["x"])
    assert np.array_equal(orc.modular_squeeze(z["av"], z["rv"], False), z["x"])
    for t in range(42):
        assert np.array_equal(orc.modular_rct(z["ch"], t), z["rct"][t])
    assert np.array_equal(orc.modular_palette(z["idx"], z["pal"], 5, 5, 8), z["pal_out"])


@pytest.mark.gpu
def test_cuda_reproduces_golden_frame(recon):
    from jxlatte_b200 import _lib
    p, st, xyb, out = _frame()
    assert np.array_equal(recon.invertVarDCT(p, st), xyb)
    assert np.array_equal(recon.reconstruct(p, st), out)          # default: fused exact kernel
    recon.set_option(_lib.OPT_STAGE2, _lib.STAGE2_STAGED)
    try:
        assert np.array_equal(recon.reconstruct(p, st), out)
    finally:
        recon.set_option(_lib.OPT_STAGE2, _lib.STAGE2_AUTO)


@pytest.mark.gpu
def test_cuda_reproduces_golden_modular(recon):
    z = np.load(os.path.join(G, "modular.npz"))
    assert np.array_equal(recon.inverseHorizontalSqueeze(z["ah"], z["rh"]), z["x"])
    assert np.array_equal(recon.inverseVerticalSqueeze(z["av"], z["rv"]), z["x"])
    for t in range(42):
        assert np.array_equal(np.stack(recon.inverseRCT(z["ch"], t)), z["rct"][t])
    assert np.array_equal(np.stack(recon.inversePalette(z["idx"], z["pal"], 5, 5, 8)), z["pal_out"])

"""The C-ABI library loads on a CPU-only box and exports every symbol include/jxlb200.h declares; the host-side pieces
(QM table build, argument checking that needs no device) behave like the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from jxlatte_b200 import _lib, default_frame_params
from jxlatte_b200.host import qm_default_params, qm_generate, InvalidBitstreamError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "jxlb200.h")).read()
    declared = set(re.findall(r"\b(jxlb200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 24
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), "libjxlb200.so does not export %s" % name
    assert declared == set(_lib.SYMBOLS), "ctypes prototypes and header drifted: %s" % (declared ^ set(_lib.SYMBOLS))


def test_frame_params_layout_matches_header():
    # 2+1+2+3+1+1+2+6+1+6+1+8+3+3+1+9+3+1 four-byte fields
    assert C.sizeof(default_frame_params(8, 8)) == 4 * 54
    from oracle.oracle import OrcFrameParams
    assert C.sizeof(OrcFrameParams) == 4 * 54


def test_qm_tables_bit_identical_to_oracle(orc):
    w, off = qm_generate()
    w2, off2 = orc.qm_default_weights()
    assert np.array_equal(off, off2)
    assert np.array_equal(w, w2)


def test_qm_custom_params_and_errors(orc):
    prm = qm_default_params()
    oprm = orc.qm_default_params()
    for q in (prm, oprm):
        q[0].dct_param[1][0] = 700.0
        q[0].dct_param[1][3] = -0.7
        q[5].dct_param[2][2] = 0.25
        q[10].param[0][6] = 0.5
    w, off = qm_generate(prm)
    rc, w2, off2 = orc.qm_generate(oprm)
    assert rc == 0 and np.array_equal(w, w2) and np.array_equal(off, off2)
    prm[2].param[0][0] = -1.0   # a non-positive weight: "Negative or infinite weight" (HFGlobal.java:425-426)
    with pytest.raises(InvalidBitstreamError):
        qm_generate(prm)


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from jxlatte_b200.host import Reconstructor
    with pytest.raises(IOError):
        Reconstructor(0)


def test_default_frame_params_reject_unpadded_sizes():
    with pytest.raises(ValueError):
        default_frame_params(100, 64)

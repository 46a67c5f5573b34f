"""The register-level transform templates the CUDA kernels use (jxlatte_b200/csrc/transforms.cuh), compiled for the
host, must reproduce the oracle BIT-EXACTLY: same products, same order of float additions."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hostlib():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "host")], stdout=subprocess.DEVNULL)
    return C.CDLL(os.path.join(HERE, "host", "libtransforms_host.so"))


FP = C.POINTER(C.c_float)


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32])
def test_ref_idct_bit_exact(hostlib, orc, n):
    rng = np.random.default_rng(n)
    for _ in range(100):
        x = (rng.standard_normal(n) * (rng.random(n) < 0.4)).astype(np.float32)
        out = np.zeros(n, np.float32)
        assert hostlib.jxlb_test_idct1d(x.ctypes.data_as(FP), out.ctypes.data_as(FP), n) == 0
        assert np.array_equal(out, orc.inverse_dct_1d(x))


@pytest.mark.parametrize("t", [0, 1, 2, 3, 12, 13, 14, 15, 16, 17])
def test_8x8_class_bit_exact(hostlib, orc, t):
    rng = np.random.default_rng(t)
    basis = orc.afv_basis().astype(np.float32)
    for _ in range(50):
        x = rng.standard_normal((8, 8)).astype(np.float32)
        out = np.zeros((8, 8), np.float32)
        assert hostlib.jxlb_test_block8(t, x.ctypes.data_as(FP), out.ctypes.data_as(FP), basis.ctypes.data_as(FP)) == 0
        assert np.array_equal(out, orc.invert_varblock(x, t))


def test_cosine_table_is_exactly_symmetric():
    """k1_big and RefIDCT form each product once for out[k] and out[N-1-k]: needs lut[n-1][N-1-k] == (-1)^n lut[n-1][k]."""
    for s in (2, 4, 8, 16, 32, 64, 128, 256):
        n = np.arange(1, s)[:, None].astype(np.float64)
        k = np.arange(s)[None, :].astype(np.float64)
        t = (np.sqrt(2.0) * np.cos(np.pi * n * (k + 0.5) / s)).astype(np.float32)
        sign = np.where((np.arange(1, s) % 2 == 0)[:, None], 1.0, -1.0).astype(np.float32)
        assert np.array_equal(t[:, ::-1] * sign, t)

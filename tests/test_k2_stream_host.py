"""The stage-2 stream kernel (jxlatte_b200/csrc/k2_stream.cuh) under the host emulator (tests/host/k2_stream_host.cpp: the same
source, one host thread per CUDA thread, shuffles / barriers / TMA boxes emulated) against the CPU oracle, BIT FOR BIT.

This is what lets the ring arithmetic, the stage lags, the frame-edge mirroring and the operation order be settled on a box
without a GPU; the `-m gpu` tests then hold the compiled sm_100a kernel to the same oracle through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from jxlatte_b200 import default_frame_params
from jxlatte_b200._lib import Slab
from jxlatte_b200.params import FrameParams

HERE = os.path.dirname(os.path.abspath(__file__))
HALO = 8


@pytest.fixture(scope="module", params=["libk2stream_host.so", "libk2stream_host8.so", "libk2stream_host18.so"], ids=["band16", "band8", "band18"])
def emu(request):
    subprocess.check_call(["make", "-C", os.path.join(HERE, "host")], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(HERE, "host", request.param))
    L.k2s_host_run.restype = C.c_int
    L.k2s_host_run.argtypes = [C.POINTER(FrameParams), C.POINTER(Slab), C.c_int, C.POINTER(C.c_void_p), C.c_longlong,
                               C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int]
    L.k2s_host_plan.restype = None
    L.k2s_host_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    return L


def _planes(rng, H, W):
    pl = rng.random((3, H, W), dtype=np.float32) * np.array([0.05, 1.0, 1.0], np.float32)[:, None, None]
    return np.ascontiguousarray(pl * 0.2 + 0.4)


def _maps(rng, H, W):
    hm = rng.integers(1, 5, size=(H // 8, W // 8)).astype(np.int32)
    sh = rng.integers(0, 8, size=(H // 8, W // 8)).astype(np.int32)      # 0 = pass-through blocks
    return hm, sh


def _oracle(orc, p, planes, hm, sh):
    x = planes
    if p.gab:
        x = orc.gab(p, x, nthreads=4)
    if p.epf_iters:
        x = orc.epf(p, x, hm, sh, nthreads=4)
    return orc.color(p, x, nthreads=4)


def _run(emu, p, planes_ext, row0, hm_ext, brow0, sh_ext, slab, n_frames, rows, grid):
    """planes_ext: [3, R, W] array whose row `row0` is the slab's first own row; hm_ext / sh_ext likewise in block rows."""
    W = p.width
    out = np.full((3, rows * n_frames, W), np.nan, np.float32)
    pitch = planes_ext.shape[2]
    ins = (C.c_void_p * 3)(*[planes_ext[c].ctypes.data + row0 * pitch * 4 for c in range(3)])
    outs = (C.c_void_p * 3)(*[out[c].ctypes.data for c in range(3)])
    rc = emu.k2s_host_run(C.byref(p), C.byref(slab) if slab is not None else None, n_frames, ins, pitch,
                          hm_ext.ctypes.data + brow0 * (W // 8) * 4, sh_ext.ctypes.data + brow0 * (W // 8) * 4, outs, grid)
    assert rc == 0
    return out


@pytest.mark.parametrize("cfg", [
    dict(W=128, H=64, iters=3, gab=1, grid=2),
    dict(W=72, H=40, iters=3, gab=1, grid=1),          # one strip, narrower than a strip: both column mirrors in one warp
    dict(W=8, H=8, iters=3, gab=1, grid=1),            # the smallest frame: every row and column mirrors
    dict(W=232, H=104, iters=2, gab=1, grid=3),        # three strips, the last one 8 columns wide
    dict(W=120, H=72, iters=1, gab=1, grid=2),
    dict(W=136, H=48, iters=3, gab=0, grid=2),
    dict(W=112, H=56, iters=1, gab=0, grid=1),
    dict(W=64, H=136, iters=2, gab=0, grid=1),
    dict(W=104, H=200, iters=3, gab=1, grid=3),        # several chunks per strip: a CTA's first item starts inside the frame
    dict(W=240, H=160, iters=2, gab=1, grid=5),
])
def test_whole_frame_bit_exact(emu, orc, cfg):
    W, H = cfg["W"], cfg["H"]
    p = default_frame_params(W, H, epf_iters=cfg["iters"], gab=bool(cfg["gab"]))
    rng = np.random.default_rng(W * 1000 + H + cfg["iters"])
    planes = _planes(rng, H, W)
    hm, sh = _maps(rng, H, W)
    want = _oracle(orc, p, planes, hm, sh)
    got = _run(emu, p, planes, 0, hm, 0, sh, None, 1, H, cfg["grid"])
    assert np.array_equal(got, want), "max abs err %g at %s" % (np.nanmax(np.abs(got - want)), np.argwhere(~(got == want))[:4])


@pytest.mark.parametrize("iters", [1, 2, 3])
def test_non_positive_hf_multipliers_bit_exact(emu, orc, iters):
    """HFMetadata takes the HF multiplier as 1 + a Modular-decoded value (HFMetadata.java:49), so a crafted stream can make it zero or
    negative: sigma and 1/sigma are then infinite or negative and epfWeight's 1 - x exceeds 1.  The row functions clamp with a
    saturating subtract only where every 1/sigma of the warp's row is non-negative; rows with such blocks take the literal form."""
    W, H = 136, 72
    p = default_frame_params(W, H, epf_iters=iters)
    rng = np.random.default_rng(4242 + iters)
    planes = _planes(rng, H, W)
    hm, sh = _maps(rng, H, W)
    hm[rng.random(hm.shape) < 0.2] = -3
    hm[rng.random(hm.shape) < 0.1] = 0
    sh[(hm < 0) & (sh == 0)] = 5        # sharpness 0 (sigma = -0, 1/sigma = -inf) would breed NaNs, which the reference's clamp passes on and a hardware max does not
    want = _oracle(orc, p, planes, hm, sh)
    got = _run(emu, p, planes, 0, hm, 0, sh, None, 1, H, 2)
    assert np.isfinite(want).all()
    assert np.array_equal(got, want), "max abs err %g" % np.abs(got - want).max()


def test_many_items_per_cta_bit_exact(emu, orc):
    """One emulated CTA streams every item of the frame back to back (several chunks per strip, several strips): the rolling
    state and the rings run across item boundaries."""
    W, H = 152, 136
    p = default_frame_params(W, H, epf_iters=3)
    rng = np.random.default_rng(77)
    planes = _planes(rng, H, W)
    hm, sh = _maps(rng, H, W)
    want = _oracle(orc, p, planes, hm, sh)
    plan = (C.c_int * 5)()
    emu.k2s_host_plan(W, H, 1, 1, plan)
    assert plan[4] == plan[2] * plan[3] and plan[2] == 2
    got = _run(emu, p, planes, 0, hm, 0, sh, None, 1, H, 1)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("iters,gab", [(3, 1), (1, 1), (2, 0)])
def test_slabs_with_halo_rows_bit_exact(emu, orc, iters, gab):
    """A frame cut into three slabs of group rows (the multi-GPU split and the pipelined host entry point): each slab sees the
    8 neighbouring rows and one neighbouring block row, mirrors only at the true frame top / bottom."""
    W, H = 136, 96
    p = default_frame_params(W, H, epf_iters=iters, gab=bool(gab))
    rng = np.random.default_rng(5 + iters)
    planes = _planes(rng, H, W)
    hm, sh = _maps(rng, H, W)
    want = _oracle(orc, p, planes, hm, sh)
    got = np.empty_like(want)
    for (y0, rows) in ((0, 32), (32, 40), (72, 24)):
        ps = default_frame_params(W, rows, epf_iters=iters, gab=bool(gab))
        slab = Slab(y0, rows, H, 1 if y0 > 0 else 0, 1 if y0 + rows < H else 0)
        got[:, y0:y0 + rows] = _run(emu, ps, planes, y0, hm, y0 // 8, sh, slab, 1, rows, 2)
    assert np.array_equal(got, want)


def test_stacked_frames_bit_exact(emu, orc):
    """jxlb200_vardct_reconstruct_batch_dev's layout: frames stacked vertically, each mirroring at its own edges."""
    W, H, n = 120, 64, 3
    p = default_frame_params(W, H, epf_iters=1)
    rng = np.random.default_rng(9)
    frames = [_planes(rng, H, W) for _ in range(n)]
    maps = [_maps(rng, H, W) for _ in range(n)]
    stack = np.ascontiguousarray(np.concatenate(frames, axis=1))
    hm = np.ascontiguousarray(np.concatenate([m[0] for m in maps], axis=0))
    sh = np.ascontiguousarray(np.concatenate([m[1] for m in maps], axis=0))
    got = _run(emu, p, stack, 0, hm, 0, sh, None, n, H, 2)
    for f in range(n):
        want = _oracle(orc, p, frames[f], maps[f][0], maps[f][1])
        assert np.array_equal(got[:, f * H:(f + 1) * H], want), "frame %d" % f


def test_plan_covers_the_frame(emu):
    for (W, H, n, cta) in ((7680, 4320, 1, 148), (2048, 2048, 16, 148), (16384, 2048, 1, 148), (8, 8, 1, 148), (520, 264, 1, 4)):
        plan = (C.c_int * 5)()
        emu.k2s_host_plan(W, H, n, cta, plan)
        ch, ir, cols, chunks, items = list(plan)
        assert ch % 8 == 0 and ir == ch + 16 and cols * 112 >= W and chunks * ch >= H and (chunks - 1) * ch < H
        assert items == n * cols * chunks

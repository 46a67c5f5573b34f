"""Known-answer checks that pin the CPU oracle as far as it can be pinned without a JVM (the reference ships no tests
or golden outputs: PARITY UNPINNED, see oracle/jxl_oracle.h).  Independent sources: scipy's DCTs, closed forms, inverse-of-
forward identities and hand-computed impulse responses."""
import numpy as np
import pytest
import scipy.fft as sf

from jxlatte_b200.params import TRANSFORM_TYPES


@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 32, 64, 128, 256])
def test_idct_1d_matches_scipy_and_inverts_forward(orc, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32)
    y = orc.inverse_dct_1d(x)
    # MathHelper.inverseDCTHorizontal == orthonormal DCT-III * sqrt(N)
    ref = sf.idct(x.astype(np.float64), type=2, norm="ortho") * np.sqrt(n)
    assert np.abs(y - ref).max() <= 4e-6 * np.sqrt(n)
    assert np.abs(orc.forward_dct_1d(y) - x).max() <= 4e-6


@pytest.mark.parametrize("shape", [(8, 8), (16, 8), (8, 32), (4, 8), (64, 32), (32, 64), (256, 128)])
def test_idct_2d_matches_scipy(orc, shape):
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    x = (rng.standard_normal((h, w)) / np.sqrt(h * w)).astype(np.float32)
    ref = sf.idctn(x.astype(np.float64), type=2, norm="ortho") * np.sqrt(h * w)
    assert np.abs(orc.inverse_dct_2d(x) - ref).max() <= 2e-5
    assert np.abs(orc.inverse_dct_2d(x, transposed=True) - ref.T).max() <= 2e-5   # W rows x H cols
    assert np.abs(orc.forward_dct_2d(orc.inverse_dct_2d(x)) - x).max() <= 1e-5


def test_llf_scale_closed_form(orc):
    i = np.arange(32)
    cf = 1.0 / (np.cos(np.pi * i / 512) * np.cos(np.pi * i / 256) * np.cos(np.pi * i / 128))
    # DCT256: 32 x 32 corner, index step 1
    for k in range(32):
        assert abs(orc.llf_scale(24, k, 0) - cf[k]) <= 2e-7 * cf[k]
        assert abs(orc.llf_scale(24, 0, k) - cf[k]) <= 2e-7 * cf[k]
    # DCT16: 2 x 2 corner, index step 16
    assert abs(orc.llf_scale(4, 1, 1) - cf[16] ** 2) <= 1e-6
    assert orc.llf_scale(0, 0, 0) == 1.0


def test_afv_basis_is_orthonormal(orc):
    a = orc.afv_basis()
    assert np.abs(a @ a.T - np.eye(16)).max() <= 5e-7


def test_mirror_coordinate(orc):
    assert [orc.mirror_coordinate(c, 10) for c in range(-4, 14)] == [3, 2, 1, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 8, 7, 6]
    assert orc.mirror_coordinate(-1, 1) == 0 and orc.mirror_coordinate(1, 1) == 0


def test_transform_type_table(orc):
    for t, (param, method, ph, pw) in enumerate(TRANSFORM_TYPES):
        info = orc.tt_info(t)
        assert (info["param_index"], info["method"], info["pixel_h"], info["pixel_w"]) == (param, method, ph, pw)
        assert info["flip"] == int(ph > pw or (method == 0 and ph == pw))


def test_qm_default_weights_known_answers(orc):
    w, off = orc.qm_default_weights()
    assert list(off[:4]) == [0, 64, 128, 192] and off[50] + 128 * 256 == w.size
    assert np.isfinite(w).all() and (w > 0).all()
    # DCT8 (param 0): weight (0,0) = 1/params[0] for X, Y, B
    for c, v in enumerate((3150.0, 560.0, 512.0)):
        assert w[off[c]] == np.float32(1.0) / np.float32(v)
    # Hornuss (param 1): w[0][0] = 1, w[0][1] = w[1][0] = 1/param[1], w[1][1] = 1/param[2], rest 1/param[0]
    o = off[3]
    assert w[o] == 1.0 and w[o + 1] == w[o + 8] == np.float32(1.0) / np.float32(3160.0)
    assert w[o + 9] == np.float32(1.0) / np.float32(3160.0) and w[o + 10] == np.float32(1.0) / np.float32(280.0)
    # DCT2 (param 2): bands by max(y, x): {1}, {2,3}, {4..7}
    o = off[6]
    m = w[o:o + 64].reshape(8, 8)
    assert m[0, 0] == 1.0 and m[0, 1] == m[1, 0] == np.float32(1) / np.float32(3840) and m[1, 1] == np.float32(1) / np.float32(2560)
    assert m[0, 3] == m[3, 1] == np.float32(1) / np.float32(1280) and m[2, 3] == np.float32(1) / np.float32(640)
    assert m[7, 0] == np.float32(1) / np.float32(480) and m[5, 6] == np.float32(1) / np.float32(300)
    # DCT quant weights are non-decreasing along the first row for monotone-negative params (Y channel of DCT16, param 4)
    y16 = w[off[4 * 3 + 1]:off[4 * 3 + 1] + 256].reshape(16, 16)
    assert (np.diff(y16[0]) >= 0).all()


def test_special_8x8_transforms_on_impulses(orc):
    z = np.zeros((8, 8), np.float32)
    # DC only: every 8x8-class transform reproduces a flat block (AFV and the 4x8 pair too)
    for t in (0, 1, 2, 3, 12, 13, 14, 15, 16, 17):
        c = z.copy()
        c[0, 0] = 0.5
        out = orc.invert_varblock(c, t)
        assert np.abs(out - 0.5).max() <= 1e-6, t
    # DCT2, by hand from PassGroup.auxDCT2 (:154-165): c01 = coeffs[iy][ix + num] enters r00, r01 with + and r10, r11 with -,
    # and r10 / r11 land on the odd ROWS -> stored coefficient (0, 1) is the top/bottom contrast, (1, 0) left/right
    c = z.copy(); c[0, 1] = 1.0
    out = orc.invert_varblock(c, 2)
    assert (out[:4] == 1).all() and (out[4:] == -1).all()
    c = z.copy(); c[1, 0] = 1.0
    out = orc.invert_varblock(c, 2)
    assert (out[:, :4] == 1).all() and (out[:, 4:] == -1).all()
    # Hornuss: a lone residual at slot (iy=1, ix=0) of quadrant (0,0) lifts that pixel and lowers the quadrant by 1/16
    c = z.copy(); c[2, 0] = 1.6
    out = orc.invert_varblock(c, 1)
    assert abs(out[1, 0] - (1.6 - 0.1)) <= 1e-6 and abs(out[3, 3] + 0.1) <= 1e-6 and np.abs(out[:, 4:]).max() == 0
    # DCT8_4 writes two 8-row x 4-col halves side by side, DCT4_8 two 4x8 halves stacked
    c = z.copy(); c[1, 0] = 1.0
    assert np.allclose(orc.invert_varblock(c, 13)[:, :4], 1) and np.allclose(orc.invert_varblock(c, 13)[:, 4:], -1)
    assert np.allclose(orc.invert_varblock(c, 12)[:4], 1) and np.allclose(orc.invert_varblock(c, 12)[4:], -1)


def test_every_type_preserves_dc_and_energy(orc):
    rng = np.random.default_rng(5)
    for t, (_, method, ph, pw) in enumerate(TRANSFORM_TYPES):
        c = np.zeros((ph, pw), np.float32)
        c[0, 0] = 0.25
        assert np.abs(orc.invert_varblock(c, t) - 0.25).max() <= 2e-6, t
        if method == 0:   # orthonormal up to the 1/N scale: mean of squares of pixels = sum of squares of coefficients
            c = (rng.standard_normal((ph, pw)) / np.sqrt(ph * pw)).astype(np.float32)
            out = orc.invert_varblock(c, t).astype(np.float64)
            assert abs((out ** 2).mean() - (c.astype(np.float64) ** 2).sum()) <= 1e-4


def test_gab_epf_invariants(orc):
    from jxlatte_b200 import default_frame_params
    p = default_frame_params(64, 48, epf_iters=3)
    flat = np.full((3, 48, 64), 0.3, np.float32)
    flat[0] = 0.01
    assert np.abs(orc.gab(p, flat) - flat).max() <= 1e-7              # weights sum to 1
    hm = np.ones((6, 8), np.int32)
    sh = np.full((6, 8), 4, np.int32)
    assert np.abs(orc.epf(p, flat, hm, sh) - flat).max() <= 1e-7      # constant plane is a fixed point
    rng = np.random.default_rng(1)
    noisy = rng.random((3, 48, 64), dtype=np.float32)
    sh0 = np.zeros((6, 8), np.int32)                                  # sharpness 0 -> sigma 0 -> pass-through
    assert np.array_equal(orc.epf(p, noisy, hm, sh0), noisy)
    with pytest.raises(RuntimeError):
        orc.epf(p, noisy, hm, np.full((6, 8), 8, np.int32))


def test_invert_xyb_of_known_colours(orc):
    """Forward XYB of sRGB-linear grey/white/primaries, written independently here, must invert to the colour."""
    from jxlatte_b200 import default_frame_params
    p = default_frame_params(8, 8)
    m = np.array(list(p.opsin_matrix), np.float64).reshape(3, 3)
    fwd = np.linalg.inv(m)
    bias = np.array(list(p.opsin_bias), np.float64)
    cols = np.array([[1, 1, 1], [0.5, 0.5, 0.5], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0.2, 0.6, 0.1], [0, 0, 0]], np.float64)
    planes = np.zeros((3, 8, 8), np.float32)
    for i, rgb in enumerate(cols):
        mix = fwd @ rgb
        g = np.cbrt(mix - bias) + np.cbrt(bias)
        planes[:, 0, i] = [(g[0] - g[1]) / 2, (g[0] + g[1]) / 2, g[2]]
    out = orc.color(p, planes)
    for i, rgb in enumerate(cols):
        assert np.abs(out[:, 0, i] - rgb).max() <= 2e-5, (rgb, out[:, 0, i])


def test_place_blocks_is_raster_first_fit(orc):
    # DCT16 then DCT8s: the 8x8s fill the gap to the right of / below the 16x16 in raster order
    rc, ds, bo, hm = orc.place_blocks(4, 4, [4, 0, 0, 0], [1, 2, 3, 4])
    assert rc == 4
    assert (ds[:2, :2] == 4).all() and bo[0, 0] == 1 and bo[0, 2] == 1 and bo[0, 3] == 1 and bo[1, 2] == 1
    assert hm[0, 2] == 2 and hm[0, 3] == 3 and hm[1, 2] == 4 and (hm[:2, :2] == 1).all()
    # a block that is too wide at the current column moves to the next row
    rc, ds, bo, hm = orc.place_blocks(4, 4, [0, 0, 0, 9], [0, 0, 0, 0])
    assert rc == 4 and bo[1, 0] == 1 and (ds[1] == 9).all()
    rc, *_ = orc.place_blocks(2, 2, [5], [0])
    assert rc == -2


def test_synthetic_partition_replays_through_placeblock(orc):
    """The generator's partitions are what HFMetadata.placeBlock produces from the type sequence in blockList order."""
    from jxlatte_b200 import synth
    rng = np.random.default_rng(9)
    for aligned in (True, False):
        ds, origin, oy, ox = synth.make_partition(64, 96, rng, aligned=aligned)
        ys, xs = np.nonzero(origin)
        types = ds[ys, xs].astype(np.int32)
        rc, ds2, bo2, _ = orc.place_blocks(64, 96, types, np.zeros_like(types))
        assert rc == len(types)
        assert np.array_equal(ds2, ds) and np.array_equal(bo2, origin)

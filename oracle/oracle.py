"""ctypes binding of the CPU oracle (oracle/liborc.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
module.  The product package (jxlatte_b200) never does.  PARITY UNPINNED: see oracle/jxl_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OrcFrameParams(C.Structure):
    # field order == oracle/jxl_oracle.h:orc_frame_params == include/jxlb200.h:jxlb200_frame_params
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("global_scale", C.c_int32),
        ("xqm_scale", C.c_int32), ("bqm_scale", C.c_int32),
        ("quant_bias", C.c_float * 3), ("quant_bias_numerator", C.c_float),
        ("color_factor", C.c_int32), ("base_corr_x", C.c_float), ("base_corr_b", C.c_float),
        ("shift_x", C.c_int32 * 3), ("shift_y", C.c_int32 * 3),
        ("gab", C.c_int32), ("gab_w1", C.c_float * 3), ("gab_w2", C.c_float * 3),
        ("epf_iters", C.c_int32), ("epf_sharp_lut", C.c_float * 8), ("epf_channel_scale", C.c_float * 3),
        ("epf_pass0_sigma_scale", C.c_float), ("epf_pass2_sigma_scale", C.c_float), ("epf_border_sad_mul", C.c_float),
        ("color_mode", C.c_int32),
        ("opsin_matrix", C.c_float * 9), ("opsin_bias", C.c_float * 3), ("intensity_target", C.c_float),
    ]


class OrcQmParams(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("n_dct", C.c_int32), ("n_param", C.c_int32), ("n_4x4", C.c_int32),
        ("denominator", C.c_float),
        ("dct_param", (C.c_float * 17) * 3), ("param", (C.c_float * 9) * 3), ("params4x4", (C.c_float * 17) * 3),
        ("raw", C.POINTER(C.c_float) * 3),
    ]


QM_TOTAL = 3 * 131584


def build(force=False):
    """Compile oracle/liborc.so with the committed Makefile (gcc, -ffp-contract=off)."""
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in ("vardct_oracle.c", "modular_oracle.c", "jxl_oracle.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_llf_scale.restype = C.c_float
        _LIB.orc_afv_basis.restype = C.POINTER(C.c_float)
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _planes(arrs, t):
    return (C.POINTER(t) * len(arrs))(*[_p(a, t) for a in arrs])


def as_params(p):
    """Accept an OrcFrameParams or any ctypes struct with the same layout (the product's FrameParams)."""
    if isinstance(p, OrcFrameParams):
        return p
    assert C.sizeof(p) == C.sizeof(OrcFrameParams), "frame-params layout drifted between product and oracle"
    return OrcFrameParams.from_buffer_copy(bytes(p))


def qm_default_weights():
    """HFGlobal default params -> (weights float32[3*131584], offsets int32[51])."""
    L = lib()
    prm = (OrcQmParams * 17)()
    L.orc_qm_default_params(prm)
    w = np.zeros(QM_TOTAL, np.float32)
    off = np.zeros(51, np.int32)
    rc = L.orc_qm_generate(prm, _p(w, C.c_float), _p(off, C.c_int32))
    assert rc == 0
    return w, off


def qm_default_params():
    L = lib()
    prm = (OrcQmParams * 17)()
    L.orc_qm_default_params(prm)
    return prm


def qm_generate(prm):
    L = lib()
    w = np.zeros(QM_TOTAL, np.float32)
    off = np.zeros(51, np.int32)
    rc = L.orc_qm_generate(prm, _p(w, C.c_float), _p(off, C.c_int32))
    return rc, w, off


def _c(a, dt):
    a = np.ascontiguousarray(a, dtype=dt)
    return a


def vardct_invert(p, st, nthreads=1, want_dequant=False):
    """PassGroup.invertVarDCT over the whole frame.  st: dict of frame-level arrays (see synth.py)."""
    L = lib()
    p = as_params(p)
    H, W = p.height, p.width
    q = [_c(st["qcoeff"][c], np.int32) for c in range(3)]
    lf = [_c(st["lf"][c], np.float32) for c in range(3)]
    dims = [(H >> p.shift_y[c], W >> p.shift_x[c]) for c in range(3)]      # chroma-subsampled channels are smaller
    out = [np.zeros(dims[c], np.float32) for c in range(3)]
    dq = [np.zeros(dims[c], np.float32) for c in range(3)] if want_dequant else None
    ds, bo = _c(st["dct_select"], np.uint8), _c(st["block_origin"], np.uint8)
    hm = _c(st["hf_mul"], np.int32)
    xf, bf = _c(st["x_from_y"], np.int32), _c(st["b_from_y"], np.int32)
    qw, qo = _c(st["qm_weights"], np.float32), _c(st["qm_offsets"], np.int32)
    rc = L.orc_vardct_invert(C.byref(p), _planes(q, C.c_int32), _planes(lf, C.c_float),
                             _p(ds, C.c_uint8), _p(bo, C.c_uint8), _p(hm, C.c_int32),
                             _p(xf, C.c_int32), _p(bf, C.c_int32), _p(qw, C.c_float), _p(qo, C.c_int32),
                             _planes(out, C.c_float), _planes(dq, C.c_float) if dq else None, int(nthreads))
    if rc:
        raise RuntimeError("orc_vardct_invert rc=%d" % rc)
    if len(set(dims)) > 1:
        return (out, dq) if want_dequant else out
    return (np.stack(out), np.stack(dq)) if want_dequant else np.stack(out)


def invert_subsampling(p, planes):
    """Frame.invertSubsampling: per-channel (H >> sy) x (W >> sx) planes -> three H x W planes."""
    L = lib()
    p = as_params(p)
    inp = [_c(planes[c], np.float32) for c in range(3)]
    out = [np.zeros((p.height, p.width), np.float32) for _ in range(3)]
    L.orc_invert_subsampling(C.byref(p), _planes(inp, C.c_float), _planes(out, C.c_float))
    return np.stack(out)


def gab(p, planes, nthreads=1):
    L = lib()
    p = as_params(p)
    inp = [_c(planes[c], np.float32) for c in range(3)]
    out = [np.zeros_like(inp[0]) for _ in range(3)]
    L.orc_gab(C.byref(p), _planes(inp, C.c_float), _planes(out, C.c_float), int(nthreads))
    return np.stack(out)


def epf(p, planes, hf_mul, sharpness, nthreads=1):
    L = lib()
    p = as_params(p)
    buf = [np.array(planes[c], dtype=np.float32, order="C", copy=True) for c in range(3)]
    hm, sh = _c(hf_mul, np.int32), _c(sharpness, np.int32)
    rc = L.orc_epf(C.byref(p), _planes(buf, C.c_float), _p(hm, C.c_int32), _p(sh, C.c_int32), int(nthreads))
    if rc:
        raise RuntimeError("orc_epf rc=%d" % rc)
    return np.stack(buf)


def lf_dequant(lf_quant, extra_precision, scaled_dequant, kx, kb, cfl=True, smooth=True):
    """LFCoefficients: quantised LF planes [3, hb, wb] (X, Y, B order) -> dequantised, LF chroma-from-luma, adaptive smoothing."""
    L = lib()
    q = [np.ascontiguousarray(lf_quant[c], np.int32) for c in range(3)]
    hb, wb = q[0].shape
    ep = np.ascontiguousarray(extra_precision, np.uint8).reshape(-1)
    assert ep.size == ((hb + 255) // 256) * ((wb + 255) // 256)
    sd = (C.c_float * 3)(*[float(v) for v in scaled_dequant])
    out = [np.zeros((hb, wb), np.float32) for _ in range(3)]
    L.orc_lf_dequant.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_lf_dequant.restype = None
    L.orc_lf_dequant(hb, wb, sd, float(kx), float(kb), int(bool(cfl)), int(bool(smooth)), _planes(q, C.c_int32), ep.ctypes.data, _planes(out, C.c_float))
    return np.stack(out)


def epf_uniform(p, planes, sigma_for_modular, nthreads=1):
    """performEdgePreservingFilter on a Modular-encoded frame: invModularSigma = 1f / epfSigmaForModular everywhere."""
    L = lib()
    p = as_params(p)
    buf = [np.array(planes[c], dtype=np.float32, order="C", copy=True) for c in range(3)]
    L.orc_epf_uniform.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_int32]
    L.orc_epf_uniform.restype = C.c_int32
    rc = L.orc_epf_uniform(C.byref(p), _planes(buf, C.c_float), float(sigma_for_modular), int(nthreads))
    if rc:
        raise RuntimeError("orc_epf_uniform rc=%d" % rc)
    return np.stack(buf)


def color(p, planes, nthreads=1):
    L = lib()
    p = as_params(p)
    buf = [np.array(planes[c], dtype=np.float32, order="C", copy=True) for c in range(3)]
    L.orc_color(C.byref(p), _planes(buf, C.c_float), int(nthreads))
    return np.stack(buf)


def vardct_reconstruct(p, st, nthreads=1):
    """Whole path: invertVarDCT -> Gaborish -> EPF -> performColorTransforms."""
    L = lib()
    p = as_params(p)
    H, W = p.height, p.width
    q = [_c(st["qcoeff"][c], np.int32) for c in range(3)]
    lf = [_c(st["lf"][c], np.float32) for c in range(3)]
    out = [np.zeros((H, W), np.float32) for _ in range(3)]
    ds, bo = _c(st["dct_select"], np.uint8), _c(st["block_origin"], np.uint8)
    hm, sh = _c(st["hf_mul"], np.int32), _c(st["sharpness"], np.int32)
    xf, bf = _c(st["x_from_y"], np.int32), _c(st["b_from_y"], np.int32)
    qw, qo = _c(st["qm_weights"], np.float32), _c(st["qm_offsets"], np.int32)
    rc = L.orc_vardct_reconstruct(C.byref(p), _planes(q, C.c_int32), _planes(lf, C.c_float),
                                  _p(ds, C.c_uint8), _p(bo, C.c_uint8), _p(hm, C.c_int32),
                                  _p(xf, C.c_int32), _p(bf, C.c_int32), _p(sh, C.c_int32),
                                  _p(qw, C.c_float), _p(qo, C.c_int32), _planes(out, C.c_float), int(nthreads))
    if rc:
        raise RuntimeError("orc_vardct_reconstruct rc=%d" % rc)
    return np.stack(out)


def inverse_dct_1d(x):
    x = _c(x, np.float32)
    out = np.zeros_like(x)
    lib().orc_inverse_dct_1d(_p(x, C.c_float), _p(out, C.c_float), x.shape[0])
    return out


def forward_dct_1d(x):
    x = _c(x, np.float32)
    out = np.zeros_like(x)
    lib().orc_forward_dct_1d(_p(x, C.c_float), _p(out, C.c_float), x.shape[0])
    return out


def inverse_dct_2d(x, transposed=False):
    x = _c(x, np.float32)
    h, w = x.shape
    out = np.zeros((w, h) if transposed else (h, w), np.float32)
    lib().orc_inverse_dct_2d(_p(x, C.c_float), _p(out, C.c_float), h, w, int(transposed))
    return out


def forward_dct_2d(x):
    x = _c(x, np.float32)
    h, w = x.shape
    out = np.zeros((h, w), np.float32)
    lib().orc_forward_dct_2d(_p(x, C.c_float), _p(out, C.c_float), h, w)
    return out


def invert_varblock(coeffs, t):
    """PassGroup.invertVarDCT for one varblock x channel: coeffs float32[pixelH, pixelW] -> pixels."""
    coeffs = _c(coeffs, np.float32)
    h, w = coeffs.shape
    out = np.zeros((h, w), np.float32)
    rc = lib().orc_invert_varblock(_p(coeffs, C.c_float), w, _p(out, C.c_float), w, int(t))
    assert rc == 0
    return out


def mirror_coordinate(c, size):
    return int(lib().orc_mirror_coordinate(int(c), int(size)))


def llf_scale(t, y, x):
    return float(lib().orc_llf_scale(int(t), int(y), int(x)))


def afv_basis():
    return np.ctypeslib.as_array(lib().orc_afv_basis(), shape=(16, 16)).copy()


def tt_info(t):
    v = [C.c_int32() for _ in range(5)]
    rc = lib().orc_tt_info(int(t), *[C.byref(x) for x in v])
    assert rc == 0
    return dict(param_index=v[0].value, method=v[1].value, pixel_h=v[2].value, pixel_w=v[3].value, flip=v[4].value)


def place_blocks(hb, wb, types, muls):
    """HFMetadata.placeBlock replay for one LF group of hb x wb blocks."""
    types, muls = _c(types, np.int32), _c(muls, np.int32)
    ds = np.full((hb, wb), 255, np.uint8)
    bo = np.zeros((hb, wb), np.uint8)
    hm = np.zeros((hb, wb), np.int32)
    rc = lib().orc_place_blocks(hb, wb, len(types), _p(types, C.c_int32), _p(muls, C.c_int32),
                                _p(ds, C.c_uint8), _p(bo, C.c_uint8), _p(hm, C.c_int32), wb)
    return rc, ds, bo, hm


def modular_rct(ch, rct_type):
    """ch: int32[3][h][w] -> channels after the inverse RCT + permutation."""
    v = [np.array(ch[c], dtype=np.int32, order="C", copy=True) for c in range(3)]
    out = [np.zeros_like(v[0]) for _ in range(3)]
    h, w = v[0].shape
    lib().orc_modular_rct(_planes(v, C.c_int32), h, w, int(rct_type), _planes(out, C.c_int32))
    return np.stack(out)


def modular_palette(idx, palette, nb_deltas, d_pred, bit_depth):
    idx = _c(idx, np.int32)
    palette = _c(palette, np.int32)
    num_c, nb_colors = palette.shape
    h, w = idx.shape
    out = [np.zeros((h, w), np.int32) for _ in range(num_c)]
    rc = lib().orc_modular_palette(_p(idx, C.c_int32), _p(palette, C.c_int32), h, w, num_c, nb_colors,
                                   int(nb_deltas), int(d_pred), int(bit_depth), _planes(out, C.c_int32))
    if rc:
        raise RuntimeError("orc_modular_palette rc=%d" % rc)
    return np.stack(out)


def modular_squeeze(avg, res, horizontal):
    avg, res = _c(avg, np.int32), _c(res, np.int32)
    ha, wa = avg.shape
    hr, wr = res.shape
    out = np.zeros((ha, wa + wr) if horizontal else (ha + hr, wa), np.int32)
    rc = lib().orc_modular_squeeze(_p(avg, C.c_int32), _p(res, C.c_int32), ha, wa, hr, wr, int(horizontal),
                                   _p(out, C.c_int32))
    if rc:
        raise RuntimeError("orc_modular_squeeze rc=%d" % rc)
    return out


def modular_forward_squeeze(x, horizontal):
    x = _c(x, np.int32)
    h, w = x.shape
    if horizontal:
        avg, res = np.zeros((h, (w + 1) // 2), np.int32), np.zeros((h, w // 2), np.int32)
    else:
        avg, res = np.zeros(((h + 1) // 2, w), np.int32), np.zeros((h // 2, w), np.int32)
    lib().orc_modular_forward_squeeze(_p(x, C.c_int32), h, w, int(horizontal), _p(avg, C.c_int32), _p(res, C.c_int32))
    return avg, res


def tendency(a, b, c):
    return int(lib().orc_tendency(int(a), int(b), int(c)))


def _rect(a):
    """(pointer to the first element, pitch in elements) of a 2-D view whose rows are contiguous."""
    if a is None:
        return None, 0
    assert a.ndim == 2 and (a.shape[1] <= 1 or a.strides[1] == a.itemsize) and a.strides[0] % a.itemsize == 0
    return C.c_void_p(a.ctypes.data), a.strides[0] // a.itemsize


def blend(op, canvas, a, b, fa=None, ra=None):
    """JXLCodestreamDecoder.blend* on one rectangle; op = dict(mode, is_int, is_alpha, has_extra, clamp, premult); the
    arguments are equally sized 2-D numpy views (canvas is written in place)."""
    L = lib()
    L.orc_blend.argtypes = [C.c_int32] * 8 + [C.c_void_p, C.c_int64] * 5
    L.orc_blend.restype = C.c_int32
    h, w = canvas.shape
    (cp, cpi), (ap, api), (bp, bpi), (fp, fpi), (rp, rpi) = _rect(canvas), _rect(a), _rect(b), _rect(fa), _rect(ra)
    rc = L.orc_blend(op["mode"], op["is_int"], op["is_alpha"], op["has_extra"], op["clamp"], op["premult"], h, w,
                     cp, cpi, ap, api, bp, bpi, fp, fpi, rp, rpi)
    if rc:
        raise RuntimeError("orc_blend rc=%d" % rc)


def upsample(plane, k, weights):
    """Frame.performUpsampling on one float32 channel; weights float32 [k][k][5][5]."""
    L = lib()
    a = _c(plane, np.float32)
    wt = _c(weights, np.float32)
    h, w = a.shape
    out = np.zeros((h * k, w * k), np.float32)
    L.orc_upsample.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.orc_upsample.restype = None
    L.orc_upsample(a.ctypes.data, h, w, int(k), wt.ctypes.data, out.ctypes.data)
    return out


def noise(planes, group_dim, seed0, lut, base_x, base_b):
    """Frame.initializeNoise + synthesizeNoise on X, Y, B planes [3, h, w]; returns the new planes."""
    L = lib()
    buf = [np.array(planes[c], dtype=np.float32, order="C", copy=True) for c in range(3)]
    h, w = buf[0].shape
    lt = _c(lut, np.float32)
    L.orc_noise.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_float, C.c_float]
    L.orc_noise.restype = None
    L.orc_noise(_planes(buf, C.c_float), h, w, int(group_dim), int(seed0), lt.ctypes.data, float(base_x), float(base_b))
    return np.stack(buf)


def splines(planes, splines_list, quant_adjust, base_x, base_b):
    """Frame.renderSplines on X, Y, B planes [3, h, w]; splines_list: [{"points": [x0, y0, x1, y1, ...], "coeff": 128 ints}]."""
    L = lib()
    buf = [np.array(planes[c], dtype=np.float32, order="C", copy=True) for c in range(3)]
    h, w = buf[0].shape
    npts = np.array([len(s["points"]) // 2 for s in splines_list], np.int32)
    pts = np.array([v for s in splines_list for v in s["points"]], np.int32)
    cf = np.array([v for s in splines_list for v in s["coeff"]], np.int32)
    L.orc_splines.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_float]
    L.orc_splines.restype = C.c_int32
    rc = L.orc_splines(_planes(buf, C.c_float), h, w, len(splines_list), npts.ctypes.data, pts.ctypes.data, cf.ctypes.data,
                       int(quant_adjust), float(base_x), float(base_b))
    if rc:
        raise RuntimeError("orc_splines rc=%d" % rc)
    return np.stack(buf)


def pack_samples(channels, depths, n_color, linear, bits):
    """PNGWriter's sample pipeline -> uint8 [h, w, C * bits/8] (big-endian for 16 bit)."""
    L = lib()
    ch = [np.ascontiguousarray(c) for c in channels]
    h, w = ch[0].shape
    n = len(ch)
    ptrs = (C.c_void_p * n)(*[c.ctypes.data for c in ch])
    is_int = np.array([c.dtype != np.float32 for c in ch], np.int32)
    dep = np.array(depths, np.int32)
    out = np.zeros((h, w, n * (2 if bits > 8 else 1)), np.uint8)
    L.orc_pack_samples.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.orc_pack_samples.restype = None
    L.orc_pack_samples(ptrs, is_int.ctypes.data, dep.ctypes.data, n, int(n_color), int(bool(linear)), h, w, int(bits), out.ctypes.data)
    return out

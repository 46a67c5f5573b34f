"""JXLDecoder / JXLImage / PNGWriter mirrors: .jxl bytes -> C++ front end (entropy decoding, headers; libjxlfront.so)
-> CUDA reconstruction (libjxlb200.so) -> image planes.

Mirrors the reference's only public API (J/JXLDecoder.java:17-46, J/JXLImage.java, J/io/PNGWriter.java:61-65,191-212,
J/io/PFMWriter.java:22-49) for the part of the format this build covers:
  * VarDCT frames, 4:4:4, XYB or YCbCr, any TransformType mix, custom or default quant tables, Gaborish, EPF;
  * Modular frames / extra channels: channels decoded on the host, RCT / Palette / Squeeze undone on the GPU;
  * frame blending mode REPLACE onto the canvas (crop offsets honoured).
Everything the frame loop of JXLCodestreamDecoder.decode (:588-626) does beyond that -- upsampling, noise, splines,
patches, the other blend modes, LF frames, chroma subsampling -- raises NotImplementedError.

`engine` is the reconstruction back end.  The default is the CUDA library (jxlatte_b200.host.Reconstructor); there is no
CPU fallback in this module -- the CPU tests inject the oracle explicitly (tests/test_frontend.py).
"""
import ctypes as C
import struct
import zlib

import numpy as np

from . import frontend
from .params import FrameParams

ENC_VARDCT, ENC_MODULAR = 0, 1
FLAG_NOISE, FLAG_PATCHES, FLAG_SPLINES, FLAG_USE_LF_FRAME = 1, 2, 16, 32
TF_SRGB, TF_LINEAR = (1 << 24) + 13, (1 << 24) + 8


class CudaEngine:
    """The product back end: every data-parallel stage on the GPU through the C ABI."""

    def __init__(self, device=0):
        from . import host
        self._host = host
        self.rec = host.Reconstructor(device)

    def qm_default(self):
        return self._host.qm_generate()

    def qm_generate(self, prm):
        return self._host.qm_generate(prm)

    def qm_params(self):
        return self._host.qm_default_params()

    def reconstruct(self, p, st):
        self.rec.setWeights(st["qm_weights"], st["qm_offsets"])
        return self.rec.reconstruct(p, st)

    def modular(self, channels, transforms, bit_depth):
        return self._host.ModularTransforms(self.rec, bit_depth).applyTransforms(channels, transforms)

    def close(self):
        self.rec.close()


def _f32(x):
    return np.float32(x)


def _conversion_matrix(color):
    """ColorManagement.getConversionMatrix(target = the image's tagged primaries / white point, current = sRGB / D65)
    (J/color/ColorManagement.java:150-161).  Identity for sRGB/D65-tagged images, which is every lossy sample here."""
    srgb = [0.639998686, 0.330010138, 0.300003784, 0.600003357, 0.150002046, 0.059997204]
    same_prim = all(abs(a - b) < 1e-6 for a, b in zip(color["prim_xy"], srgb))
    same_wp = abs(color["white_xy"][0] - 0.3127) < 1e-6 and abs(color["white_xy"][1] - 0.3290) < 1e-6
    if same_prim and same_wp:
        return np.eye(3, dtype=np.float32)

    def xyz(xy):
        inv = _f32(1.0) / _f32(xy[1])
        return np.array([_f32(xy[0]) * inv, _f32(1.0), (_f32(1.0) - _f32(xy[0]) - _f32(xy[1])) * inv], np.float32)

    def prim_to_xyz(prim, wp):
        m = np.stack([xyz(prim[0:2]), xyz(prim[2:4]), xyz(prim[4:6])]).T.astype(np.float32)
        s = np.linalg.inv(m.astype(np.float64)).astype(np.float32) @ xyz(wp)
        return (m @ np.diag(s)).astype(np.float32)

    bradford = np.array([[0.8951, 0.2664, -0.1614], [-0.7502, 1.7135, 0.0367], [0.0389, -0.0685, 1.0296]], np.float32)
    wp_t, wp_c = color["white_xy"], [0.3127, 0.3290]
    adapt = np.eye(3, dtype=np.float32)
    if not same_wp:
        lc, lt = bradford @ xyz(wp_c), bradford @ xyz(wp_t)
        adapt = (np.linalg.inv(bradford.astype(np.float64)).astype(np.float32) @ np.diag(lt / lc) @ bradford).astype(np.float32)
    forward = prim_to_xyz(srgb, wp_c)
    reverse = np.linalg.inv(prim_to_xyz(color["prim_xy"], wp_t).astype(np.float64)).astype(np.float32)
    return (reverse @ adapt @ forward).astype(np.float32)


class JXLImage:
    """Decoded image: `planes` float32 [C, H, W] (colour first, then extra channels), linear light in the tagged primaries
    for XYB images (what PFMWriter writes), or integer-valued samples scaled to [0, 1] for non-XYB images."""

    def __init__(self, planes, info, linear):
        self.planes = planes
        self.info = info
        self.linear = linear

    @property
    def width(self):
        return self.planes.shape[2]

    @property
    def height(self):
        return self.planes.shape[1]

    def to_int(self, bits=8):
        """PNGWriter's sample pipeline: TF_SRGB.fromLinearF for linear images, then (int)(v * max + 0.5f) clamped
        (J/color/TransferFunction.java:39-43, J/util/ImageBuffer.java:129-145)."""
        v = self.planes.astype(np.float32)
        ncol = self.info["color_channels"]
        if self.linear:
            c = v[:ncol]
            with np.errstate(invalid="ignore"):
                hi = np.float32(1.055) * np.power(c.astype(np.float64), 0.4166666666666667).astype(np.float32) + np.float32(-0.055)
            v = v.copy()
            v[:ncol] = np.where(c < np.float32(0.00313066844250063), c * np.float32(12.92), hi)
        maxv = (1 << bits) - 1
        q = (v * np.float32(maxv) + np.float32(0.5))
        q = np.where(np.isnan(q), 0, q)
        return np.clip(q.astype(np.int64), 0, maxv).astype(np.uint16 if bits > 8 else np.uint8)

    def write_png(self, path, bits=8):
        q = self.to_int(bits)
        c = q.shape[0]
        ctype = {1: 0, 2: 4, 3: 2, 4: 6}[c]
        rows = np.moveaxis(q, 0, 2)
        if bits > 8:
            rows = rows.astype(">u2")
        raw = b"".join(b"\x00" + rows[y].tobytes() for y in range(rows.shape[0]))

        def chunk(tag, data):
            return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
        with open(path, "wb") as f:
            f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", q.shape[2], q.shape[1], 16 if bits > 8 else 8, ctype, 0, 0, 0)) +
                    chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))

    def write_pfm(self, path):
        p = self.planes[:3] if self.planes.shape[0] >= 3 else self.planes[:1]
        with open(path, "wb") as f:
            f.write(("PF\n" if p.shape[0] == 3 else "Pf\n").encode() + ("%d %d\n1.0\n" % (self.width, self.height)).encode())
            f.write(np.moveaxis(p, 0, 2)[::-1].astype(">f4").tobytes())


class JXLDecoder:
    def __init__(self, source, engine=None):
        if isinstance(source, (bytes, bytearray, memoryview)):
            self.data = bytes(source)
        elif isinstance(source, str):
            with open(source, "rb") as f:
                self.data = f.read()
        else:
            self.data = source.read()
        self._own = engine is None
        self.engine = engine
        self.timings = {}

    def close(self):
        if self._own and self.engine is not None:
            self.engine.close()
            self.engine = None

    # ---- frame-level pieces ----
    def frame_params(self, info, f):
        p = FrameParams()
        p.width, p.height = f["padded_width"], f["padded_height"]
        p.global_scale = f["global_scale"]
        p.xqm_scale, p.bqm_scale = f["xqm_scale"], f["bqm_scale"]
        p.quant_bias[:] = info["quant_bias"]
        p.quant_bias_numerator = info["quant_bias_numerator"]
        p.color_factor = f["color_factor"]
        p.base_corr_x, p.base_corr_b = f["base_corr_x"], f["base_corr_b"]
        p.shift_x[:] = f["shift_x"]
        p.shift_y[:] = f["shift_y"]
        p.gab = 1 if f["gab"] else 0
        p.gab_w1[:] = f["gab_w1"]
        p.gab_w2[:] = f["gab_w2"]
        p.epf_iters = f["epf_iters"]
        p.epf_sharp_lut[:] = f["epf_sharp_lut"]
        p.epf_channel_scale[:] = f["epf_channel_scale"]
        p.epf_pass0_sigma_scale = f["epf_pass0_sigma_scale"]
        p.epf_pass2_sigma_scale = f["epf_pass2_sigma_scale"]
        p.epf_border_sad_mul = f["epf_border_sad_mul"]
        p.color_mode = 1 if info["xyb_encoded"] else (2 if f["do_ycbcr"] else 0)
        m = (_conversion_matrix(info["color"]) @ np.array(info["opsin_inverse"], np.float32).reshape(3, 3)).astype(np.float32)
        p.opsin_matrix[:] = [float(x) for x in m.reshape(-1)]
        p.opsin_bias[:] = info["opsin_bias"]
        p.intensity_target = info["intensity_target"]
        return p

    def quant_tables(self, parsed, k, f):
        if f["quant_all_default"]:
            return self.engine.qm_default()
        prm = self.engine.qm_params()          # library defaults, overwritten where the stream codes its own
        keep = []
        for i, q in enumerate(f["quant_params"]):
            if q["mode"] == 0:
                continue
            prm[i].mode, prm[i].n_dct, prm[i].n_param, prm[i].n_4x4 = q["mode"], q["n_dct"], q["n_param"], q["n_4x4"]
            prm[i].denominator = q["denominator"]
            for c in range(3):
                prm[i].dct_param[c][:] = q["dct_param"][17 * c:17 * c + 17]
                prm[i].param[c][:] = q["param"][9 * c:9 * c + 9]
                prm[i].params4x4[c][:] = q["params4x4"][17 * c:17 * c + 17]
                if q["mode"] == 7:
                    raw = np.ascontiguousarray(parsed.array(k, "qraw", 3 * i + c), np.float32)
                    keep.append(raw)
                    prm[i].raw[c] = raw.ctypes.data_as(C.POINTER(C.c_float))
        out = self.engine.qm_generate(prm)
        del keep
        return out

    def decode_frame(self, parsed, k):
        """-> float32 planes [C, h, w] of frame k (frame size, before blending), colour transform applied."""
        info, f = parsed.info, parsed.frames[k]
        if f["flags"] & (FLAG_NOISE | FLAG_PATCHES | FLAG_SPLINES | FLAG_USE_LF_FRAME):
            raise NotImplementedError("noise / patches / splines / LF frames (flags=%d): SURVEY.md 8f-3/4" % f["flags"])
        if f["upsampling"] != 1 or any(u != 1 for u in f["ec_upsampling"]):
            raise NotImplementedError("upsampling: SURVEY.md 8f-4")
        h, w = f["height"], f["width"]
        ncol = 3 if (info["xyb_encoded"] or f["encoding"] == ENC_VARDCT) else info["color_channels"]
        planes = []
        bits = info["bits_per_sample"]
        mod = None
        if f["modular"]["channels"]:
            tr = [dict(tr=t["tr"], begin_c=t["begin_c"], rct_type=t["rct_type"], num_c=t["num_c"], nb_colors=t["nb_colors"],
                       nb_deltas=t["nb_deltas"], d_pred=t["d_pred"], sp=[(bool(s[0]), bool(s[1]), s[2], s[3]) for s in t["sp"]])
                  for t in f["modular"]["transforms"]]
            mod = self.engine.modular(parsed.modular_channels(k), tr, bits)
        if f["encoding"] == ENC_VARDCT:
            st = parsed.vardct_state(k)
            st["qm_weights"], st["qm_offsets"] = self.quant_tables(parsed, k, f)
            p = self.frame_params(info, f)
            rec = self.engine.reconstruct(p, st)
            planes = [rec[c][:h, :w] for c in range(3)]
            linear = bool(info["xyb_encoded"])
            if not linear:
                planes = planes        # YCbCr -> RGB already in [0, 1]
        else:
            if info["xyb_encoded"]:
                raise NotImplementedError("XYB-encoded Modular frames")
            if info["exp_bits"]:
                raise NotImplementedError("float samples in Modular frames")
            scale = np.float32(1.0) / np.float32((1 << bits) - 1)
            planes = [mod[c][:h, :w].astype(np.float32) * scale for c in range(ncol)]
            linear = False
        if mod is not None:
            first = 0 if f["encoding"] == ENC_VARDCT else ncol
            for e, ec in enumerate(info["extra_channels"]):
                scale = np.float32(1.0) / np.float32((1 << ec["bits_per_sample"]) - 1)
                planes.append(mod[first + e][:h, :w].astype(np.float32) * scale)
        return np.stack(planes), linear

    def decode(self):
        import time
        t0 = time.perf_counter()
        parsed = frontend.parse(self.data)
        self.timings["front_end_s"] = time.perf_counter() - t0
        if self.engine is None:
            self.engine = CudaEngine()
        info = parsed.info
        canvas, linear = None, False
        t1 = time.perf_counter()
        for k, f in enumerate(parsed.frames):
            if f["type"] in (1, 2):
                raise NotImplementedError("LF frames / reference-only frames: SURVEY.md 8f-3")
            planes, linear = self.decode_frame(parsed, k)
            if canvas is None:
                canvas = np.zeros((planes.shape[0], info["height"], info["width"]), np.float32)
            if f["blend_mode"] != 0:
                raise NotImplementedError("blend mode %d: SURVEY.md 8f-3" % f["blend_mode"])
            y0, x0 = f["y0"], f["x0"]
            ys, xs = max(0, y0), max(0, x0)
            ye, xe = min(info["height"], y0 + planes.shape[1]), min(info["width"], x0 + planes.shape[2])
            if ye > ys and xe > xs:
                canvas[:, ys:ye, xs:xe] = planes[:, ys - y0:ye - y0, xs - x0:xe - x0]
        self.timings["reconstruct_s"] = time.perf_counter() - t1
        o = info["orientation"]
        if o != 1:                    # JXLCodestreamDecoder.transposeBuffer
            if o in (2, 3):
                canvas = canvas[:, :, ::-1] if o == 2 else canvas[:, ::-1, ::-1]
            elif o == 4:
                canvas = canvas[:, ::-1, :]
            elif o == 5:
                canvas = canvas.transpose(0, 2, 1)
            elif o == 6:
                canvas = canvas.transpose(0, 2, 1)[:, :, ::-1]
            elif o == 7:
                canvas = canvas[:, ::-1, ::-1].transpose(0, 2, 1)
            else:
                canvas = canvas.transpose(0, 2, 1)[:, ::-1, :]
            canvas = np.ascontiguousarray(canvas)
        parsed.close()
        return JXLImage(canvas, info, linear)

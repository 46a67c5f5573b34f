"""C++ front end (libjxlfront.so) + decoder glue on real .jxl files (tests/golden/samples, copied sample INPUTS of the
reference).  The reference ships no decoded outputs and no JVM exists here, so what pins the front end is (a) every ANS
stream of every file ending in its mandatory final state, (b) the decoded pictures (checked by eye when the fixtures were
made, pinned here by hash of the oracle's 8-bit output), (c) the GPU decode being bit-identical to the oracle's decode."""
import hashlib
import json
import os

import numpy as np
import pytest

from jxlatte_b200 import frontend
from jxlatte_b200.decoder import JXLDecoder, JXLImage

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
S = os.path.join(G, "samples")
PINS = json.load(open(os.path.join(G, "frontend_pins.json")))


def _parse(name, flags=0):
    return frontend.parse_file(os.path.join(S, name + ".jxl"), flags)


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_abi_symbols():
    L = frontend.lib()
    for sym in ("jxlf_decode", "jxlf_free", "jxlf_error", "jxlf_describe", "jxlf_array"):
        assert hasattr(L, sym)
    hdr = open(os.path.join(os.path.dirname(G), "..", "include", "jxlfront.h")).read()
    for sym in ("jxlf_decode", "jxlf_free", "jxlf_error", "jxlf_describe", "jxlf_array"):
        assert sym + "(" in hdr


@pytest.mark.parametrize("name", ["lenna", "bbb", "white", "bench", "quilt", "art", "patches-lossless", "blendmodes_5", "wb-rainbow"])
def test_headers_and_state(name):
    p = _parse(name)
    pin = PINS[name]
    i = p.info
    assert (i["width"], i["height"], i["xyb_encoded"], i["orientation"]) == tuple(pin["image"])
    k = len(p.frames) - 1
    f = p.frames[k]
    assert [f["encoding"], f["width"], f["height"], f["gab"], f["epf_iters"], f["num_groups"]] == pin["frame"]
    if f["encoding"] == 0:
        st = p.vardct_state(k)
        got = {n: _digest(st[n]) for n in ("qcoeff", "lf", "dct_select", "block_origin", "hf_mul", "sharpness", "x_from_y", "b_from_y")}
        assert got == pin["state"]
        # structural invariants of the varblock partition
        ds, bo = st["dct_select"], st["block_origin"]
        from jxlatte_b200.params import TRANSFORM_TYPES
        area = sum((TRANSFORM_TYPES[t][2] // 8) * (TRANSFORM_TYPES[t][3] // 8) for t in ds[bo == 1])
        assert area == ds.size
        assert st["hf_mul"].min() >= 1 and 0 <= st["sharpness"].min() and st["sharpness"].max() <= 7
    else:
        ch = p.modular_channels(k)
        assert [_digest(c) for c in ch] == pin["modular"]
    p.close()


def test_truncated_and_garbage_inputs():
    data = open(os.path.join(S, "lenna.jxl"), "rb").read()
    with pytest.raises(frontend.InvalidBitstreamError):
        frontend.parse(data[:20000])
    with pytest.raises(frontend.InvalidBitstreamError):
        frontend.parse(b"not a jxl file at all")
    bad = bytearray(data)
    bad[5000] ^= 0x55                      # a flipped bit inside an ANS stream must trip a final-state or range check
    p = frontend.parse(bytes(bad), strict=False)
    assert p.status in (-2, 0)
    if p.status == 0:
        assert p.vardct_state(0)["qcoeff"].shape == (3, 512, 512)


def test_host_transforms_flag_matches_glue():
    """The frame-level modular transforms undone by the front end itself (test flag) and by the decoder glue with the
    oracle's transforms agree: two independent implementations of ModularStream.applyTransforms on real data."""
    from oracle_engine import OracleEngine
    for name in ("quilt", "art"):
        a = _parse(name, frontend.FLAG_HOST_TRANSFORMS)
        b = _parse(name)
        f = b.frames[0]
        tr = [dict(tr=t["tr"], begin_c=t["begin_c"], rct_type=t["rct_type"], num_c=t["num_c"], nb_colors=t["nb_colors"],
                   nb_deltas=t["nb_deltas"], d_pred=t["d_pred"], sp=[(bool(s[0]), bool(s[1]), s[2], s[3]) for s in t["sp"]])
              for t in f["modular"]["transforms"]]
        want = OracleEngine().modular(b.modular_channels(0), tr, b.info["bits_per_sample"])
        got = a.modular_channels(0)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)


@pytest.mark.parametrize("name", ["lenna", "white", "quilt", "patches-lossless", "blendmodes_5", "wb-rainbow"])
def test_decode_with_oracle_engine(name):
    from oracle_engine import OracleEngine
    img = JXLDecoder(os.path.join(S, name + ".jxl"), engine=OracleEngine()).decode()
    assert isinstance(img, JXLImage)
    assert _digest(img.to_int(8)) == PINS[name]["png8"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lenna", "bbb", "white", "bench", "quilt", "art", "patches-lossless", "blendmodes_5", "wb-rainbow"])
def test_gpu_decode_matches_oracle_decode(name):
    """BASELINE configs[1] (+ the modular art files): the CUDA path and the oracle decode the same real file to the same
    bits -- planes equal, hence 8- and 16-bit PNG samples equal."""
    from oracle_engine import OracleEngine
    path = os.path.join(S, name + ".jxl")
    want = JXLDecoder(path, engine=OracleEngine()).decode()
    dec = JXLDecoder(path)
    got = dec.decode()
    dec.close()
    assert got.planes.shape == want.planes.shape and got.planes.dtype == want.planes.dtype
    if name == "wb-rainbow":
        # splines evaluate (float)Math.exp(double): the device's double exp and the host libm may differ in the last bit of
        # the double, which flips the float rounding once in ~1e8 evaluations; everything else in this file is exact
        assert float(np.abs(got.planes - want.planes).max()) <= 1e-6
        assert int(np.abs(got.to_int(16).astype(np.int64) - want.to_int(16).astype(np.int64)).max()) <= 1
        return
    assert np.array_equal(got.planes, want.planes, equal_nan=True)
    assert np.array_equal(got.to_int(16), want.to_int(16))
    if name in PINS and "png8" in PINS[name]:
        assert _digest(got.to_int(8)) == PINS[name]["png8"]




@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ants", "george-tiled", "sollevante-hdr"])
def test_gpu_decode_matches_oracle_decode_large(name):
    """BASELINE configs[0] (ants.jxl: JPEG-recompressed, chroma-subsampled YCbCr, raw quant tables), a 135-frame tiled
    image and a 4K HDR photo: CUDA decode == oracle decode, bit for bit."""
    from oracle_engine import OracleEngine
    path = os.path.join(S, name + ".jxl")       # committed since round 2: configs[0] must not skip on the GPU box
    want = JXLDecoder(path, engine=OracleEngine()).decode()
    dec = JXLDecoder(path)
    got = dec.decode()
    dec.close()
    assert np.array_equal(got.planes, want.planes)


@pytest.mark.gpu
def test_blend_kernel_matches_oracle(recon):
    """jxlb200_blend vs orc_blend: every mode x flag combination on strided rectangles, canvas aliasing the frame."""
    from oracle import oracle as orc
    rng = np.random.default_rng(11)
    H, W, h, w = 37, 53, 21, 30
    for mode in (1, 2, 3, 4):
        for is_alpha in (0, 1):
            for has_extra in (0, 1):
                for clamp in (0, 1):
                    for premult in (0, 1):
                        if mode == 3 and has_extra and is_alpha:
                            continue          # plain copy, done by the caller
                        big = [rng.uniform(-0.5, 1.5, (H, W)).astype(np.float32) for _ in range(4)]
                        a, b, fa, ra = [x[5:5 + h, 7:7 + w] for x in big]
                        op = dict(mode=mode, is_int=0, is_alpha=is_alpha, has_extra=has_extra, clamp=clamp, premult=premult)
                        want = a.copy()
                        orc.blend(op, want, a.copy(), b, fa, ra)
                        got_full = big[0].copy()
                        got = got_full[5:5 + h, 7:7 + w]
                        recon.blend(op, got, got, b, fa, ra)          # canvas aliases the frame operand, as for patches
                        assert np.array_equal(got, want, equal_nan=True), op
                        assert np.array_equal(got_full[:5], big[0][:5]) and np.array_equal(got_full[:, :7], big[0][:, :7])
    ia = rng.integers(-2 ** 31, 2 ** 31 - 1, (H, W), dtype=np.int64).astype(np.int32)
    ib = rng.integers(-2 ** 31, 2 ** 31 - 1, (H, W), dtype=np.int64).astype(np.int32)
    op = dict(mode=1, is_int=1, is_alpha=0, has_extra=1, clamp=0, premult=0)
    want = np.zeros((h, w), np.int32)
    orc.blend(op, want, ia[:h, :w], ib[:h, :w])
    got = np.zeros((h, w), np.int32)
    recon.blend(op, got, ia[:h, :w], ib[:h, :w])
    assert np.array_equal(got, want)
    with pytest.raises(ValueError):
        recon.blend(dict(op, mode=4), got, ia[:h, :w], ib[:h, :w])


@pytest.mark.gpu
def test_blend_batch_matches_sequential_oracle(recon):
    """jxlb200_blend_batch: 300 overlapping rectangles over six planes (two written, canvas aliasing the `frame` operand as patches
    do), blended on the device in order, == the oracle applying them one after the other on the host arrays."""
    from oracle import oracle as orc
    rng = np.random.default_rng(21)
    H, W = 72, 90
    planes = [rng.uniform(-0.2, 1.2, (H, W)).astype(np.float32) for _ in range(6)]
    want = [p.copy() for p in planes]
    got = [p.copy() for p in planes]
    items = []
    for k in range(300):
        h, w = int(rng.integers(1, 30)), int(rng.integers(1, 40))
        mode = int(rng.integers(1, 5))
        is_alpha = int(rng.integers(0, 2)) if mode != 3 else 0
        op = dict(mode=mode, is_int=0, is_alpha=is_alpha, has_extra=1, clamp=int(rng.integers(0, 2)), premult=int(rng.integers(0, 2)))
        cv = int(rng.integers(0, 2))                       # planes 0 and 1 are canvases
        pos = [(int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1))) for _ in range(5)]
        refs = [(cv, *pos[0]), (cv, *pos[0]) if k % 2 else (2, *pos[1]), (3, *pos[2]), (4, *pos[3]), (5, *pos[4])]
        items.append((op, (h, w), refs))
        v = [want[r[0]][r[1]:r[1] + h, r[2]:r[2] + w] for r in refs]
        orc.blend(op, v[0], v[1], v[2], v[3], v[4])
    recon.blend_batch(got, [True, True, False, False, False, False], items)
    for i in range(6):
        assert np.array_equal(got[i], want[i], equal_nan=True), "plane %d" % i
    with pytest.raises(Exception):
        recon.blend_batch(got, [False] * 6, items[:1])   # writes a plane not marked writable


def test_mutated_inputs_never_crash():
    """Bit flips, truncations and overwritten runs: the C++ front end must come back with a status, not a crash or a hang
    (run under -fsanitize=address,undefined when the fixtures were made: clean)."""
    rng = np.random.default_rng(2024)
    seen = set()
    for name in ("lenna", "white", "quilt", "art", "blendmodes_5", "patches-lossless"):
        data = open(os.path.join(S, name + ".jxl"), "rb").read()
        for i in range(25 if len(data) > 1000 else 60):
            m = bytearray(data)
            kind = int(rng.integers(0, 3))
            if kind == 0:
                for _ in range(int(rng.integers(1, 5))):
                    m[int(rng.integers(0, len(m)))] ^= 1 << int(rng.integers(0, 8))
            elif kind == 1:
                m = m[:int(rng.integers(0, len(m)))]
            else:
                at = int(rng.integers(0, len(m)))
                m[at:at + 8] = bytes(rng.integers(0, 256, min(8, len(m) - at), dtype=np.uint8).tolist())
            p = frontend.parse(bytes(m), strict=False)
            assert p.status in (0, -1, -2, -3)
            seen.add(p.status)
            p.close()
    assert -2 in seen


@pytest.mark.gpu
@pytest.mark.parametrize("k", [2, 4, 8])
def test_upsampling_kernel_exact(recon, k):
    from oracle import oracle as orc
    from jxlatte_b200.upsampling import up_weights
    rng = np.random.default_rng(k)
    for shape in ((37, 61), (2, 3), (1, 1), (64, 5)):
        a = rng.uniform(-1.0, 1.5, shape).astype(np.float32)
        wts = up_weights(k)
        assert np.array_equal(recon.performUpsampling(a, k, wts), orc.upsample(a, k, wts))
    neg = -rng.uniform(0.1, 1.0, (9, 9)).astype(np.float32)          # all-negative window: the Float.MIN_VALUE quirk caps at ~0
    assert np.array_equal(recon.performUpsampling(neg, k, up_weights(k)), orc.upsample(neg, k, up_weights(k)))


@pytest.mark.gpu
def test_noise_kernels_exact(recon):
    from oracle import oracle as orc
    rng = np.random.default_rng(3)
    for (h, w, gd) in ((300, 530, 256), (64, 40, 128), (257, 17, 256)):
        pl = rng.uniform(-0.1, 0.9, (3, h, w)).astype(np.float32)
        lut = rng.uniform(0, 0.6, 8).astype(np.float32)
        seed0 = (3 << 32) | 1
        assert np.array_equal(recon.synthesizeNoise(pl, gd, seed0, lut, 0.0, 1.0), orc.noise(pl, gd, seed0, lut, 0.0, 1.0))


@pytest.mark.gpu
def test_spline_rendering_matches_oracle(recon):
    from oracle import oracle as orc
    rng = np.random.default_rng(4)
    h, w = 200, 320
    pl = rng.uniform(0, 0.5, (3, h, w)).astype(np.float32)
    coeff = [0] * 128
    coeff[0], coeff[32], coeff[64], coeff[96], coeff[33], coeff[97] = 60, 900, 300, 12, -40, 3
    sp = [{"points": [20, 30, 150, 60, 90, 170, 300, 120], "coeff": coeff}, {"points": [250, 20], "coeff": [5] * 128},
          {"points": [10, 190, 310, 10], "coeff": [1] * 128}]
    got, want = recon.renderSplines(pl, sp, 2, 0.0, 1.0), orc.splines(pl, sp, 2, 0.0, 1.0)
    assert float(np.abs(want - pl).max()) > 0.05                       # something was drawn
    assert float(np.abs(got - want).max()) <= 1e-6
    assert float((got != want).mean()) < 1e-4


def test_png_sample_packing_oracle_vs_numpy_mirror():
    """orc_pack_samples (C restatement of TF_SRGB.fromLinearF + castToInt + interleave) == JXLImage.to_int (numpy mirror)."""
    from oracle_engine import OracleEngine
    eng = OracleEngine()
    for name in ("lenna", "patches-lossless", "blendmodes_5"):
        img = JXLDecoder(os.path.join(S, name + ".jxl"), engine=eng).decode()
        for bits in (8, 16):
            assert np.array_equal(img.packed(bits, engine=eng), img.packed(bits))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lenna", "patches-lossless", "blendmodes_5", "bench"])
def test_gpu_png_sample_packing(name):
    """jxlb200_pack_samples vs the oracle: the device's double pow and the host libm may round the last bit of a double
    differently, so a sample may differ by one step once in ~1e8; anything more is a bug."""
    from oracle_engine import OracleEngine
    from jxlatte_b200.decoder import CudaEngine
    eng, ref = CudaEngine(), OracleEngine()
    img = JXLDecoder(os.path.join(S, name + ".jxl"), engine=ref).decode()
    for bits in (8, 16):
        got, want = img.packed(bits, engine=eng), img.packed(bits, engine=ref)
        assert got.shape == want.shape
        if bits == 8:
            d = np.abs(got.astype(np.int32) - want.astype(np.int32))
        else:
            d = np.abs(got.view(">u2").astype(np.int32) - want.view(">u2").astype(np.int32))
        assert int(d.max()) <= 1 and float((d != 0).mean()) < 1e-6
    eng.close()


class _BitWriter:
    def __init__(self):
        self.bits = []

    def put(self, value, n):
        for i in range(n):
            self.bits.append((value >> i) & 1)

    def align(self):
        while len(self.bits) % 8:
            self.bits.append(0)

    def tobytes(self):
        self.align()
        out = bytearray(len(self.bits) // 8)
        for i, b in enumerate(self.bits):
            out[i >> 3] |= b << (i & 7)
        return bytes(out)


def _u32_toc_len(w, v):
    """U32(u(10), 1024 + u(14), 17408 + u(22), 4211712 + u(30)) as Frame.readTOC reads it."""
    for sel, (base, nbits) in enumerate(((0, 10), (1024, 14), (17408, 22), (4211712, 30))):
        if v - base < (1 << nbits) and v >= base:
            w.put(sel, 2)
            w.put(v - base, nbits)
            return
    raise ValueError(v)


def test_permuted_toc_crafted_from_lenna():
    """No sample uses a permuted TOC, so one is crafted: lenna's sections are stored with the first two swapped and a TOC
    permutation (a Lehmer code behind a one-symbol prefix code) that swaps them back.  The parsed state must be identical."""
    data = open(os.path.join(S, "lenna.jxl"), "rb").read()
    base = frontend.parse(data)
    f = base.frames[0]
    assert data[:2] == b"\xff\x0a" and not base.info["coverage"]["permuted_toc"]
    lengths, first, tbit = f["toc_lengths"], f["toc_first_section"], f["toc_bit_offset"]
    assert len(lengths) == 7
    secs, at = [], first
    for n in lengths:
        secs.append(data[at:at + n])
        at += n
    w = _BitWriter()
    for i in range(tbit):                       # everything before the TOC, bit for bit
        w.bits.append((data[i >> 3] >> (i & 7)) & 1)
    w.put(1, 1)                                 # permuted
    # EntropyStream(8 contexts): no LZ77; simple cluster map with 0 bits per entry; prefix codes; hybrid config 15-0-0
    w.put(0, 1)
    w.put(1, 1); w.put(0, 2)
    w.put(1, 1)
    w.put(15, 4)
    w.put(1, 1); w.put(0, 4); w.put(0, 0)       # alphabet size 1 + (1 << 0) + u(0) = 2
    w.put(1, 2); w.put(0, 2); w.put(1, 1)       # simple prefix code, one symbol, value 1 (1 bit wide)
    # end = 1, lehmer[0] = 1: both are the symbol "1", which costs no bits -> permutation [1, 0, 2, 3, 4, 5, 6]
    w.align()
    phys = [secs[1], secs[0]] + secs[2:]        # logical section i lives at physical position perm[i]
    for s_ in phys:
        _u32_toc_len(w, len(s_))
    w.align()
    crafted = w.tobytes() + b"".join(phys) + data[at:]
    got = frontend.parse(crafted)
    assert got.info["coverage"]["permuted_toc"] == 1
    a, b = base.vardct_state(0), got.vardct_state(0)
    for k in ("qcoeff", "lf", "dct_select", "block_origin", "hf_mul", "sharpness", "x_from_y", "b_from_y"):
        assert np.array_equal(a[k], b[k]), k


class _FakeParsed:
    """A ParsedImage stand-in assembled from real parsed data, to drive decoder paths no sample file reaches."""

    def __init__(self, info, states, modulars):
        self.info, self._states, self._mod = info, states, modulars

    @property
    def frames(self):
        return self.info["frames"]

    def vardct_state(self, k, copy=True):
        return dict(self._states[k])

    def modular_channels(self, k):
        return self._mod[k]

    def array(self, *a, **kw):
        raise KeyError(a)

    def close(self):
        pass


def test_lf_frame_and_lossy_modular_paths(monkeypatch):
    """LF frames and XYB-encoded Modular frames occur in no sample.  Build the pair from lenna: frame 0 = an LF frame
    (type 1, lf_level 1) coded as a lossy Modular frame whose integer Y, X, B-Y channels are lenna's own LF planes quantised
    finely; frame 1 = lenna with USE_LF_FRAME set and empty LF planes.  The decode must come out within the LF quantisation
    step of the ordinary decode."""
    import copy
    from oracle_engine import OracleEngine
    from jxlatte_b200 import decoder as dec_mod
    path = os.path.join(S, "lenna.jxl")
    real = frontend.parse_file(path)
    want = JXLDecoder(path, engine=OracleEngine()).decode().planes
    st = real.vardct_state(0)
    lf = st["lf"]                                              # X, Y, B float planes 64 x 64
    step = np.float32(1.0 / 65536.0)
    yq = np.rint(lf[1] / step).astype(np.int32)
    xq = np.rint(lf[0] / step).astype(np.int32)
    bq = np.rint(lf[2] / step).astype(np.int32) - yq
    info = copy.deepcopy(real.info)
    f1 = copy.deepcopy(info["frames"][0])
    f0 = copy.deepcopy(f1)
    f0.update(type=1, lf_level=1, encoding=1, width=64, height=64, padded_width=64, padded_height=64, gab=False, epf_iters=0,
              lf_dequant=[float(step)] * 3, flags=0, is_last=False, save_as_reference=0, num_groups=1,
              modular=dict(nb_meta=0, transformed=False, channels=[dict(h=64, w=64, hshift=0, vshift=0)] * 3, transforms=[]))
    f1.update(flags=f1["flags"] | 32, modular=dict(nb_meta=0, transformed=False, channels=[], transforms=[]))
    info["frames"] = [f0, f1]
    st1 = dict(st)
    st1["lf"] = np.zeros_like(lf)
    fake = _FakeParsed(info, {1: st1}, {0: [yq, xq, bq], 1: []})
    monkeypatch.setattr(dec_mod.frontend, "parse", lambda data, flags=0, strict=True: fake)
    got = JXLDecoder(b"unused", engine=OracleEngine()).decode().planes
    assert got.shape == want.shape
    assert float(np.abs(got - want).max()) < 2e-3              # LF quantised to 2^-16: far below this, far above a wrong path
    assert float(np.abs(got - want).max()) > 0.0


def test_decode_iterates_displayed_images():
    """JXLDecoder.decode() returns the next displayed image and None at the end (atEnd()), like the reference's API."""
    from oracle_engine import OracleEngine
    d = JXLDecoder(os.path.join(S, "white.jxl"), engine=OracleEngine())
    assert not d.atEnd()
    img = d.decode()
    assert img is not None and (img.width, img.height) == (320, 240)
    assert d.atEnd() and d.decode() is None


def test_container_box_size_overflow_is_rejected():
    """a 64-bit extended box size near 2^64 used to wrap `at + size`, pass the bounds check and walk backwards
    forever (36-byte file).  It must come back as an error, immediately."""
    import struct
    sig = bytes([0, 0, 0, 0x0C, 0x4A, 0x58, 0x4C, 0x20, 0x0D, 0x0A, 0x87, 0x0A])
    free8 = struct.pack(">I4s", 8, b"free")
    for tag in (b"free", b"jxlc", b"jxlp"):
        evil = struct.pack(">I4sQ", 1, tag, 0xFFFFFFFFFFFFFFF8)
        with pytest.raises(Exception):
            frontend.parse(sig + free8 + evil)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lenna", "bbb"])
def test_tolerance_mode_for_8_bit_png_output(name):
    """JXLOptions(outputFormat=PNG, outputDepth=8): the decoder may use the re-associated stage 2 (JXLB200_OPT_STAGE2 = 2); the
    north star's tolerance is what it is held to -- 1e-4 on the linear planes, 1 LSB on the 8-bit samples -- against the bit-exact
    decode.  Any other output keeps the bit-exact kernels."""
    from jxlatte_b200.decoder import JXLOptions
    path = os.path.join(S, name + ".jxl")
    from jxlatte_b200.decoder import CudaEngine
    exact = JXLDecoder(path).decode()
    eng = CudaEngine(allow_tolerance_mode=True)
    fast = JXLDecoder(path, engine=eng, options=JXLOptions(JXLOptions.OUTPUT_PNG, 8)).decode()
    eng.close()
    # not enabled on the engine (the default: the tolerance kernel is not faster, profiles/r2_exact_vs_fast.md) -> exact kernels
    assert np.array_equal(JXLDecoder(path, options=JXLOptions(JXLOptions.OUTPUT_PNG, 8)).decode().planes, exact.planes)
    assert float(np.abs(fast.planes - exact.planes).max()) <= 1e-4
    assert int(np.abs(fast.to_int(8).astype(np.int64) - exact.to_int(8).astype(np.int64)).max()) <= 1
    assert not np.array_equal(fast.planes, exact.planes)          # it really took the other kernel
    for opts in (JXLOptions(JXLOptions.OUTPUT_PNG, 16), JXLOptions(JXLOptions.OUTPUT_PFM), JXLOptions()):
        assert np.array_equal(JXLDecoder(path, options=opts).decode().planes, exact.planes)


def _fake_modular_frame(real_info, chans, w, h, gab, iters, sigma, xyb, ncolor):
    import copy
    info = copy.deepcopy(real_info)
    f = copy.deepcopy(info["frames"][0])
    f.update(type=0, lf_level=0, encoding=1, width=w, height=h, padded_width=w, padded_height=h, gab=gab, epf_iters=iters,
             epf_sigma_for_modular=sigma, lf_dequant=[1.0 / 4096] * 3, flags=0, is_last=True, save_as_reference=0, num_groups=1,
             upsampling=1, do_ycbcr=False,
             modular=dict(nb_meta=0, transformed=False, channels=[dict(h=h, w=w, hshift=0, vshift=0)] * len(chans), transforms=[]))
    info["frames"] = [f]
    info.update(width=w, height=h, xyb_encoded=xyb, color_channels=ncolor, bits_per_sample=8, orientation=1)
    return _FakeParsed(info, {}, {0: list(chans)})


@pytest.mark.parametrize("cfg", [dict(w=67, h=45, gab=True, iters=2, ncolor=3), dict(w=40, h=33, gab=True, iters=3, ncolor=1),
                                 dict(w=64, h=48, gab=False, iters=1, ncolor=3)])
def test_modular_frame_with_gaborish_and_epf_oracle(monkeypatch, orc, cfg):
    """Gaborish / EPF on a Modular-encoded frame occur in no sample: fabricate one -- integer RGB (or grey)
    channels, frame size not a multiple of 8 -- and hold the decoder to the stages written out by hand from Frame.java:505-679:
    cast to float, Gaborish, EPF with invModularSigma = 1 / epfSigmaForModular, channel 0 for every distance term when grey."""
    from oracle_engine import OracleEngine
    from jxlatte_b200 import decoder as dec_mod
    real = frontend.parse_file(os.path.join(S, "lenna.jxl"))
    rng = np.random.default_rng(cfg["w"])
    w, h, nc = cfg["w"], cfg["h"], cfg["ncolor"]
    base = rng.integers(90, 160, size=(nc, h, w)) + (rng.random((nc, h, w)) < 0.02) * 60
    chans = [np.ascontiguousarray(base[c], np.int32) for c in range(nc)]
    sigma = 1.75
    fake = _fake_modular_frame(real.info, chans, w, h, cfg["gab"], cfg["iters"], sigma, False, nc)
    monkeypatch.setattr(dec_mod.frontend, "parse", lambda data, flags=0, strict=True: fake)
    dec = JXLDecoder(b"unused", engine=OracleEngine())
    got = dec.decode().planes
    # by hand
    p = dec.frame_params(fake.info, dict(fake.info["frames"][0]))
    p.color_mode = 0
    fl = [(chans[c].astype(np.float32) * (np.float32(1.0) / np.float32(255))) for c in range(nc)]
    if nc == 1:
        for c in (1, 2):
            p.gab_w1[c], p.gab_w2[c] = p.gab_w1[0], p.gab_w2[0]
        fl = [fl[0]] * 3
    x = np.stack(fl)
    if cfg["gab"]:
        x = orc.gab(p, x)
    x = orc.epf_uniform(p, x, sigma)
    assert got.shape == (nc, h, w)
    assert np.array_equal(got, x[:nc])
    assert not np.array_equal(got, np.stack(fl)[:nc])


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [dict(w=67, h=45, gab=True, iters=2), dict(w=40, h=33, gab=True, iters=3), dict(w=264, h=136, gab=True, iters=3),
                                 dict(w=64, h=48, gab=False, iters=1), dict(w=9, h=5, gab=True, iters=0)])
def test_modular_frame_filters_kernel_matches_oracle(recon, orc, cfg):
    """jxlb200_restore_uniform (one sigma for the frame; sizes that are not multiples of 8 take the staged kernels, the others the
    fused tile kernel) == Gaborish + EPF of the oracle, bit for bit; sigma below 0.3 copies through like the reference."""
    from jxlatte_b200 import default_frame_params
    w, h = cfg["w"], cfg["h"]
    p = default_frame_params(((w + 7) // 8) * 8, ((h + 7) // 8) * 8, epf_iters=cfg["iters"], gab=cfg["gab"], color_mode=0)
    p.width, p.height = w, h
    rng = np.random.default_rng(w * 7 + h)
    planes = (rng.random((3, h, w), dtype=np.float32) * np.float32(0.3) + np.float32(0.3))
    for sigma in (1.5, 0.2):
        want = orc.gab(p, planes) if cfg["gab"] else planes
        if cfg["iters"]:
            want = orc.epf_uniform(p, want, sigma)
        got = recon.restoreModularFrame(p, planes, sigma)
        assert np.array_equal(got, want), "sigma %g: max abs err %g" % (sigma, np.abs(got - want).max())


@pytest.mark.parametrize("name", ["lenna", "bbb"])
def test_front_end_lf_equals_oracle_lf_dequant(orc, name):
    """The quantised LF planes the front end now hands out, through the oracle's restatement of LFCoefficients.java:61-103,
    113-179, must give the front end's own dequantised LF bit for bit (two independent restatements of the same Java)."""
    p = _parse(name)
    k = len(p.frames) - 1
    lfq = p.lf_quantised(k)
    assert lfq is not None
    q, ep, sd, kx, kb, smooth = lfq
    want = p.vardct_state(k)["lf"]
    got = orc.lf_dequant(q, ep, sd, kx, kb, cfl=True, smooth=smooth)
    assert np.array_equal(got, want)
    p.close()


def test_lf_dequant_oracle_properties(orc):
    """A constant field is a fixed point of the smoothing; without smoothing and CfL it is q * scaledDequant / 2^extraPrecision."""
    q = np.full((3, 20, 30), 7, np.int32)
    sd = [np.float32(0.01), np.float32(0.02), np.float32(0.03)]
    out = orc.lf_dequant(q, [1], sd, 0.0, 0.0, cfl=False, smooth=False)
    for c in range(3):
        assert np.array_equal(out[c], np.full((20, 30), np.float32(7) * (sd[c] / np.float32(2)), np.float32))
    sm = orc.lf_dequant(q, [1], sd, 0.0, 0.0, cfl=False, smooth=True)
    assert np.abs(sm - out).max() < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [dict(hb=300, wb=520, cfl=True, smooth=True), dict(hb=64, wb=64, cfl=True, smooth=False),
                                 dict(hb=257, wb=2, cfl=False, smooth=True), dict(hb=5, wb=700, cfl=True, smooth=True)])
def test_lf_dequant_kernel_matches_oracle(recon, orc, cfg):
    """jxlb200_lf_dequant == orc_lf_dequant: several LF groups with their own extraPrecision, groups too small to smooth."""
    rng = np.random.default_rng(cfg["hb"] + cfg["wb"])
    hb, wb = cfg["hb"], cfg["wb"]
    q = rng.integers(-900, 900, size=(3, hb, wb)).astype(np.int32)
    ng = ((hb + 255) // 256) * ((wb + 255) // 256)
    ep = rng.integers(0, 4, size=ng).astype(np.uint8)
    sd = [np.float32(0.0021), np.float32(0.00037), np.float32(0.0011)]
    kx, kb = np.float32(0.013), np.float32(1.0 + 3.0 / 84.0)
    want = orc.lf_dequant(q, ep, sd, kx, kb, cfl=cfg["cfl"], smooth=cfg["smooth"])
    got = recon.dequantLF(q, ep, sd, kx, kb, cfl=cfg["cfl"], smooth=cfg["smooth"])
    assert np.array_equal(got, want), "max abs err %g" % np.abs(got - want).max()

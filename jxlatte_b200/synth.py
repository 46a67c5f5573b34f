"""Deterministic synthetic post-entropy frame state.

This is what the reference holds right before Frame.decodePassGroups' invertVarDCT loop
(J/frame/Frame.java:361-374): quantised HF coefficients, dequantised LF, the varblock partition
(HFMetadata.dctSelect / blockList / hfMultiplier, J/frame/vardct/HFMetadata.java:23-54), the CfL tiles and the
EPF sharpness map -- stitched to frame level.  Both the CUDA path and the CPU oracle read the same arrays.

Partition modes
  aligned=True   every varblock is aligned to its own size (what libjxl encoders emit) and never crosses a
                 256x256 group or the frame edge.
  aligned=False  raster first-fit of randomly drawn types inside each group (HFMetadata.placeBlock order),
                 which produces varblocks straddling 64x64 CfL tiles -- exercises the reference's
                 chromaFromLuma cache quirk (J/frame/vardct/HFCoefficients.java:155-185).
"""
import numpy as np

from .params import TRANSFORM_TYPES

SEED_BASE = 0x4A584C00

CLASS8 = [0, 1, 2, 3, 12, 13, 14, 15, 16, 17]
MIXES = {
    # area fractions of the 8x8 / 16-32 / 64 / 128 / 256 classes
    "mixed": (0.30, 0.25, 0.20, 0.15, 0.10),
    "small": (0.55, 0.45, 0.0, 0.0, 0.0),
    "dct8": None,
}


def _dims(t):
    _, _, ph, pw = TRANSFORM_TYPES[t]
    return ph >> 3, pw >> 3


class _Layout:
    """Partition of a region of hb x wb 8x8-cells."""

    def __init__(self, hb, wb):
        self.ds = np.full((hb, wb), 255, np.uint8)
        self.dy = np.zeros((hb, wb), np.int16)  # offset of the cell from its varblock's top-left
        self.dx = np.zeros((hb, wb), np.int16)
        self.hb, self.wb = hb, wb

    def fits(self, y, x, t):
        bh, bw = _dims(t)
        return y + bh <= self.hb and x + bw <= self.wb and (self.ds[y:y + bh, x:x + bw] == 255).all()

    def put(self, y, x, t):
        bh, bw = _dims(t)
        self.ds[y:y + bh, x:x + bw] = t
        self.dy[y:y + bh, x:x + bw] = np.arange(bh, dtype=np.int16)[:, None]
        self.dx[y:y + bh, x:x + bw] = np.arange(bw, dtype=np.int16)[None, :]


def _fill_aligned(L, rng, y0, x0, size, probs, force=None):
    """Recursive aligned subdivision of the size x size-cell square at (y0, x0).  force = 0/1/2 picks the
    square / two tall / two wide arrangement at this level unconditionally."""
    if y0 >= L.hb or x0 >= L.wb:
        return
    p8, p32, p64, p128, p256 = probs
    if size >= 8:
        level = {32: 0, 16: 1, 8: 2}[size]
        rest = [p256 + p128 + p64 + p32 + p8, p128 + p64 + p32 + p8, p64 + p32 + p8][level]
        take = [p256, p128, p64][level]
        sq, tall, wide = [(24, 25, 26), (21, 22, 23), (18, 19, 20)][level]
        if force is not None or (rest > 0 and rng.random() < take / rest):
            k = rng.integers(0, 3) if force is None else force
            half = size // 2
            cand = [[(y0, x0, sq)], [(y0, x0, tall), (y0, x0 + half, tall)], [(y0, x0, wide), (y0 + half, x0, wide)]][k]
            if all(L.fits(y, x, t) for (y, x, t) in cand):
                for (y, x, t) in cand:
                    L.put(y, x, t)
                return
        h = size // 2
        for (yy, xx) in ((y0, x0), (y0, x0 + h), (y0 + h, x0), (y0 + h, x0 + h)):
            _fill_aligned(L, rng, yy, xx, h, probs)
        return
    # size == 4 cells (32 px): 16/32 class tilings, else 8x8 class
    rest = p32 + p8
    if rest > 0 and rng.random() < p32 / rest:
        k = rng.integers(0, 7)
        if k == 0:
            cand = [(y0, x0, 5)]
        elif k == 1:
            cand = [(y0, x0, 10), (y0, x0 + 2, 10)]                     # 32x16 side by side
        elif k == 2:
            cand = [(y0, x0, 11), (y0 + 2, x0, 11)]                     # 16x32 stacked
        elif k == 3:
            cand = [(y0, x0 + i, 8) for i in range(4)]                  # four 32x8
        elif k == 4:
            cand = [(y0 + i, x0, 9) for i in range(4)]                  # four 8x32
        elif k == 5:
            cand = [(y0, x0, 4), (y0, x0 + 2, 4), (y0 + 2, x0, 4), (y0 + 2, x0 + 2, 4)]
        else:
            cand = []
            for (yy, xx) in ((y0, x0), (y0, x0 + 2), (y0 + 2, x0), (y0 + 2, x0 + 2)):
                if rng.random() < 0.5:
                    cand += [(yy, xx, 6), (yy, xx + 1, 6)]              # two 16x8
                else:
                    cand += [(yy, xx, 7), (yy + 1, xx, 7)]              # two 8x16
        if all(y + _dims(t)[0] <= L.hb and x + _dims(t)[1] <= L.wb for (y, x, t) in cand):
            for (y, x, t) in cand:
                L.put(y, x, t)
            return
    for y in range(y0, min(y0 + 4, L.hb)):
        for x in range(x0, min(x0 + 4, L.wb)):
            L.put(y, x, CLASS8[rng.integers(0, len(CLASS8))])


def _fill_first_fit(L, rng, probs):
    """Raster first-fit of random types (no alignment): the order HFMetadata.placeBlock would produce."""
    p8, p32, p64, p128, p256 = probs
    classes = [CLASS8, [4, 5, 6, 7, 8, 9, 10, 11], [18, 19, 20], [21, 22, 23], [24, 25, 26]]
    # draw probability per free cell ~ area fraction / typical area (in cells)
    wts = np.array([p8 / 1.0, p32 / 6.0, p64 / 48.0, p128 / 190.0, p256 / 700.0])
    wts = wts / wts.sum()
    for y in range(L.hb):
        for x in range(L.wb):
            if L.ds[y, x] != 255:
                continue
            for _ in range(3):
                cls = classes[rng.choice(5, p=wts)]
                t = cls[rng.integers(0, len(cls))]
                if L.fits(y, x, t):
                    break
            else:
                t = CLASS8[rng.integers(0, len(CLASS8))]
            L.put(y, x, t)


def make_partition(hb, wb, rng, mix="mixed", aligned=True, bank=27):
    """-> dct_select u8[hb,wb], block_origin u8[hb,wb], (oy, ox) int32 maps of each cell's varblock origin."""
    if mix == "dct8":
        ds = np.zeros((hb, wb), np.uint8)
        oy, ox = np.meshgrid(np.arange(hb, dtype=np.int32), np.arange(wb, dtype=np.int32), indexing="ij")
        return ds, np.ones((hb, wb), np.uint8), oy.copy(), ox.copy()
    probs = MIXES[mix] if isinstance(mix, str) else tuple(mix)

    def gen(h, w, force=None):
        L = _Layout(h, w)
        if aligned:
            _fill_aligned(L, rng, 0, 0, 32, probs, force)
        else:
            _fill_first_fit(L, rng, probs)
        assert (L.ds != 255).all()
        return L

    gh, gw = (hb + 31) // 32, (wb + 31) // 32
    n_full = (hb // 32) * (wb // 32)
    layouts = [gen(32, 32) for _ in range(min(bank, n_full))] if n_full else []
    if aligned and probs[4] > 0 and n_full >= 24:
        # a frame this large must contain all three 256-class kinds (DCT256, DCT256_128, DCT128_256)
        layouts += [gen(32, 32, force=k) for k in range(3)]
    ds = np.zeros((hb, wb), np.uint8)
    dy = np.zeros((hb, wb), np.int32)
    dx = np.zeros((hb, wb), np.int32)
    for gy in range(gh):
        for gx in range(gw):
            h, w = min(32, hb - gy * 32), min(32, wb - gx * 32)
            L = layouts[rng.integers(0, len(layouts))] if (h == 32 and w == 32) else gen(h, w)
            sl = (slice(gy * 32, gy * 32 + h), slice(gx * 32, gx * 32 + w))
            ds[sl], dy[sl], dx[sl] = L.ds, L.dy, L.dx
    yy, xx = np.meshgrid(np.arange(hb, dtype=np.int32), np.arange(wb, dtype=np.int32), indexing="ij")
    oy, ox = yy - dy, xx - dx
    origin = ((dy == 0) & (dx == 0)).astype(np.uint8)
    return ds, origin, oy, ox


def _smooth_field(hb, wb, rng, coarse=16):
    ch, cw = hb // coarse + 2, wb // coarse + 2
    g = rng.random((ch, cw), dtype=np.float32)
    ys = (np.arange(hb, dtype=np.float32) + 0.5) / coarse
    xs = (np.arange(wb, dtype=np.float32) + 0.5) / coarse
    y0, x0 = ys.astype(np.int32), xs.astype(np.int32)
    fy, fx = (ys - y0)[:, None], (xs - x0)[None, :]
    a = g[y0][:, x0]
    b = g[y0][:, x0 + 1]
    c = g[y0 + 1][:, x0]
    d = g[y0 + 1][:, x0 + 1]
    return (a * (1 - fy) * (1 - fx) + b * (1 - fy) * fx + c * fy * (1 - fx) + d * fy * fx).astype(np.float32)


def make_state(width, height, seed=SEED_BASE, mix="mixed", aligned=True, density=(0.30, 0.45, 0.25),
               decay=(7.0, 4.0, 8.0), amplitude=(0.004, 0.03, 0.03), sharp_zero_fraction=0.0,
               params=None, qm_weights=None, qm_offsets=None, survey_spec=False):
    """Synthetic frame state.  width/height: padded size (multiples of 8).  Returns a dict of numpy arrays:
    qcoeff i32[3,H,W], lf f32[3,H/8,W/8], dct_select u8, block_origin u8, hf_mul i32, sharpness i32 (all [H/8,W/8]),
    x_from_y / b_from_y i32[ceil(H/64), ceil(W/64)].  Plane order is X, Y, B (Frame buffers, J/frame/Frame.java:42).

    If params + QM tables are given, coefficient sparsity is amplitude-aware: a coefficient whose dequantisation
    step (scaleFactor/hfMul * weight, HFCoefficients.java:299,314) exceeds amplitude[c] is mostly zero, the way an
    encoder quantises -- this keeps the XYB planes in an image-like range so the 1e-4 tolerance means something."""
    assert width % 8 == 0 and height % 8 == 0
    rng = np.random.default_rng(seed)
    H, W = height, width
    hb, wb = H // 8, W // 8
    ds, origin, oy, ox = make_partition(hb, wb, rng, mix=mix, aligned=aligned)

    # hf_mul: U{1..4}, constant per varblock (HFMetadata.placeBlock fills the whole block, :111-114)
    hm_cell = rng.integers(1, 5, size=(hb, wb), dtype=np.int32)
    hf_mul = hm_cell[oy, ox]
    sharp = rng.integers(1, 8, size=(hb, wb), dtype=np.int32)
    if sharp_zero_fraction > 0:
        sharp[rng.random((hb, wb)) < sharp_zero_fraction] = 0
    th, tw = (H + 63) // 64, (W + 63) // 64
    x_from_y = rng.integers(-4, 5, size=(th, tw), dtype=np.int32)     # kX = 0 + v/84
    b_from_y = rng.integers(-8, 9, size=(th, tw), dtype=np.int32)     # kB = 1 + v/84

    # LF: smooth luminance with a little block noise; X small; B follows Y
    s1, s2, s3 = (_smooth_field(hb, wb, rng) for _ in range(3))
    lf = np.empty((3, hb, wb), np.float32)
    lf[1] = 0.15 + 0.55 * s1 + 0.02 * (rng.random((hb, wb), dtype=np.float32) - 0.5)
    lf[0] = 0.02 * (s2 - 0.5) + 0.002 * (rng.random((hb, wb), dtype=np.float32) - 0.5)
    lf[2] = lf[1] + 0.1 * (s3 - 0.5)

    # HF: sparse, frequency-decaying.  P(nonzero) = density[c] * exp(-decay[c] * r), r = normalised radial frequency
    # of the coefficient inside its own varblock; magnitude 1 + Geometric(0.6), bounded by the amplitude budget.
    tt = np.array(TRANSFORM_TYPES, np.int32)  # (param, method, ph, pw)
    up = lambda a: np.repeat(np.repeat(a, 8, axis=0), 8, axis=1)
    ds_px = up(ds)
    ph, pw = tt[ds_px, 2], tt[ds_px, 3]
    ly = np.arange(H, dtype=np.int32)[:, None] - up(oy) * 8
    lx = np.arange(W, dtype=np.int32)[None, :] - up(ox) * 8
    fy = ly.astype(np.float32) / ph.astype(np.float32)
    fx = lx.astype(np.float32) / pw.astype(np.float32)
    r = np.sqrt(fy * fy + fx * fx)
    del fy, fx
    unit = None
    if params is not None and qm_weights is not None:
        flip = (ph > pw) | ((tt[ds_px, 1] == 0) & (ph == pw))           # TransformType.flip
        mw = np.maximum(ph, pw)
        wy = np.where(flip, lx, ly)
        wx = np.where(flip, ly, lx)
        widx = wy * mw + wx
        gs = np.float32(65536.0) / np.float32(params.global_scale)
        sf = [gs * np.float32(0.8) ** (params.xqm_scale - 2), gs, gs * np.float32(0.8) ** (params.bqm_scale - 2)]
        hm_px = up(hf_mul).astype(np.float32)
        par = tt[ds_px, 0]
        unit = [np.float32(sf[c]) / hm_px * qm_weights[qm_offsets[par * 3 + c] + widx] for c in range(3)]
        del wy, wx, widx, mw, flip, hm_px, par
    del ly, lx, ph, pw, ds_px
    qcoeff = np.empty((3, H, W), np.int32)
    for c in range(3):
        prob = np.float32(density[c]) * np.exp(np.float32(-decay[c]) * r)
        mag = rng.geometric(0.6, size=(H, W)).astype(np.int32)
        if unit is not None:
            room = np.float32(amplitude[c]) / unit[c]                   # how many steps fit in the budget
            prob = prob * np.minimum(np.float32(1.0), room * room)
            mag = np.minimum(mag, np.maximum(1, (2.0 * room).astype(np.int32)))
        mag = np.minimum(mag, 63)
        nz = rng.random((H, W), dtype=np.float32) < prob
        sign = rng.integers(0, 2, size=(H, W), dtype=np.int32) * 2 - 1
        qcoeff[c] = np.where(nz, mag * sign, 0)
    if survey_spec:
        # SURVEY.md 8(d) as written: 85 % zeros wherever the coefficient sits, else sign * (1 + Geometric(0.5)) clipped to +-63, and
        # chroma-from-luma factors U{-32..32}.  Not image-like (|B| reaches several units) -- kept as the stress case for the sparse
        # column walk of k1_big, reported beside the amplitude-aware default (profiles/r2_coefficient_density.md has real files).
        for c in range(3):
            mag = np.minimum(1 + rng.geometric(0.5, size=(H, W)).astype(np.int32), 63)
            nz = rng.random((H, W), dtype=np.float32) < np.float32(0.15)
            sign = rng.integers(0, 2, size=(H, W), dtype=np.int32) * 2 - 1
            qcoeff[c] = np.where(nz, mag * sign, 0)
        x_from_y = rng.integers(-32, 33, size=(th, tw), dtype=np.int32)
        b_from_y = rng.integers(-32, 33, size=(th, tw), dtype=np.int32)
    st = dict(qcoeff=qcoeff, lf=lf, dct_select=ds, block_origin=origin, hf_mul=hf_mul, sharpness=sharp,
              x_from_y=x_from_y, b_from_y=b_from_y, width=W, height=H)
    if qm_weights is not None:
        st["qm_weights"], st["qm_offsets"] = qm_weights, qm_offsets
    return st


def epf_active_fraction(p, st):
    """Share of 8x8 blocks EPF actually filters: invSigma <= 1/0.3 (J/frame/Frame.java:563-570, 608)."""
    gs = np.float32(65536.0) / np.float32(p.global_scale)
    lut = np.array(list(p.epf_sharp_lut), np.float32)
    with np.errstate(divide="ignore"):
        sigma = gs * lut[st["sharpness"]] / st["hf_mul"].astype(np.float32)
        inv = np.float32(1.0) / sigma
    return float((inv <= np.float32(1.0) / np.float32(0.3)).mean())

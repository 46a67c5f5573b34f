"""Pins against a JVM run of the reference, when fixtures exist (tests/golden/from_jvm/README.md; none are committed because no
JDK exists in this image -- the tests skip and say so).  For every dumped frame: the oracle, and on a GPU box the CUDA path through
the C ABI, must reproduce the reference's own planes after stage 1 (invertVarDCT), after Gaborish + EPF, and after the colour
transform."""
import glob
import json
import os

import numpy as np
import pytest

from jxlatte_b200 import default_frame_params

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "golden", "from_jvm")
DT = {"f32": np.float32, "i32": np.int32, "u8": np.uint8}


def _fixtures():
    out = []
    for d in sorted(glob.glob(os.path.join(ROOT, "*", "manifest.json"))):
        base = os.path.dirname(d)
        for pj in sorted(glob.glob(os.path.join(base, "f*_params.json"))):
            out.append((base, os.path.basename(pj).split("_")[0]))
    return out


FIX = _fixtures()


def _f(bits):
    return float(np.array([bits], np.int32).view(np.float32)[0])


def load_frame(base, tag):
    """-> (FrameParams, state dict, {stage: planes}) from DumpState's files."""
    man = {m["file"]: m for m in json.load(open(os.path.join(base, "manifest.json")))}

    def arr(name):
        m = man[name]
        return np.fromfile(os.path.join(base, name), dtype=DT[m["dtype"]]).reshape(m["shape"])

    pj = json.load(open(os.path.join(base, tag + "_params.json")))
    p = default_frame_params(pj["width"], pj["height"], epf_iters=pj["epf_iters"], gab=bool(pj["gab"]),
                             global_scale=pj["global_scale"])
    p.xqm_scale, p.bqm_scale = pj["xqm_scale"], pj["bqm_scale"]
    for i in range(3):
        p.quant_bias[i] = _f(pj["quant_bias"][i])
        p.shift_x[i], p.shift_y[i] = pj["shift_x"][i], pj["shift_y"][i]
        p.gab_w1[i], p.gab_w2[i] = _f(pj["gab_w1"][i]), _f(pj["gab_w2"][i])
        p.epf_channel_scale[i] = _f(pj["epf_channel_scale"][i])
    p.quant_bias_numerator = _f(pj["quant_bias_numerator"])
    p.color_factor = pj["color_factor"]
    p.base_corr_x, p.base_corr_b = _f(pj["base_corr_x"]), _f(pj["base_corr_b"])
    for i in range(8):
        p.epf_sharp_lut[i] = _f(pj["epf_sharp_lut"][i])
    p.epf_pass0_sigma_scale, p.epf_pass2_sigma_scale = _f(pj["epf_pass0_sigma_scale"]), _f(pj["epf_pass2_sigma_scale"])
    p.epf_border_sad_mul = _f(pj["epf_border_sad_mul"])
    p.color_mode = 2 if pj["do_ycbcr"] else 0
    cj = os.path.join(base, tag + "_color.json")
    if os.path.exists(cj):
        c = json.load(open(cj))
        for i in range(9):
            p.opsin_matrix[i] = _f(c["opsin_matrix"][i])
        for i in range(3):
            p.opsin_bias[i] = _f(c["opsin_bias"][i])
        p.intensity_target = _f(c["intensity_target"])
        p.color_mode |= 1
    st = {"qcoeff": [arr("%s_qcoeff_%d.i32" % (tag, c)) for c in range(3)], "lf": [arr("%s_lf_%d.f32" % (tag, c)) for c in range(3)]}
    for k, ext in (("dct_select", "u8"), ("block_origin", "u8"), ("hf_mul", "i32"), ("sharpness", "i32"), ("x_from_y", "i32"), ("b_from_y", "i32")):
        st[k] = arr("%s_%s.%s" % (tag, k, ext))
    if not any(pj["shift_x"]) and not any(pj["shift_y"]):
        st["qcoeff"], st["lf"] = np.stack(st["qcoeff"]), np.stack(st["lf"])
    st["qm_weights"] = arr(tag + "_qm_weights.f32").reshape(-1)
    from oracle import oracle
    _, st["qm_offsets"] = oracle.qm_default_weights()          # the flattened layout is fixed by the TransformType table
    st["height"], st["width"] = pj["height"], pj["width"]
    stages = {}
    for s in ("after_idct", "after_filters", "after_color"):
        names = ["%s_%s_%d.f32" % (tag, s, c) for c in range(3)]
        if all(n in man for n in names):
            stages[s] = np.stack([arr(n) for n in names])
    return p, st, stages


def _report(name, got, want):
    if np.array_equal(got, want):
        return None
    d = np.abs(got.astype(np.float64) - want.astype(np.float64))
    return "%s differs from the JVM's planes: %d of %d samples, max abs %g" % (name, int((got != want).sum()), got.size, d.max())


@pytest.mark.skipif(not FIX, reason="no fixtures from a JVM run of the reference (tests/golden/from_jvm/README.md): parity unpinned")
@pytest.mark.parametrize("base,tag", FIX or [("none", "f0")])
def test_oracle_against_jvm(orc, base, tag):
    p, st, stages = load_frame(base, tag)
    msgs = []
    if "after_idct" in stages and not any(p.shift_x) and not any(p.shift_y):
        msgs.append(_report("oracle stage 1", orc.vardct_invert(p, st, nthreads=8), stages["after_idct"]))
    if "after_color" in stages:
        msgs.append(_report("oracle whole path", orc.vardct_reconstruct(p, st, nthreads=8), stages["after_color"]))
    msgs = [m for m in msgs if m]
    assert not msgs, "; ".join(msgs)


@pytest.mark.gpu
@pytest.mark.skipif(not FIX, reason="no fixtures from a JVM run of the reference (tests/golden/from_jvm/README.md): parity unpinned")
@pytest.mark.parametrize("base,tag", FIX or [("none", "f0")])
def test_cuda_against_jvm(recon, base, tag):
    p, st, stages = load_frame(base, tag)
    recon.setWeights(st["qm_weights"], st["qm_offsets"])
    try:
        got = recon.reconstruct(p, st)
    finally:
        recon.generateWeights()
    msg = _report("CUDA whole path", got, stages["after_color"])
    assert not msg, msg


def test_loader_round_trip(tmp_path, orc):
    """The loader itself, on a fixture written here in DumpState's format from a synthetic frame (so the format cannot rot
    while no JVM fixture exists): oracle planes in, the same planes out, every scalar as a bit pattern."""
    from jxlatte_b200 import synth
    W, H = 64, 48
    p = default_frame_params(W, H, epf_iters=2)
    qw, qo = orc.qm_default_weights()
    st = synth.make_state(W, H, seed=5, params=p, qm_weights=qw, qm_offsets=qo, mix="small")
    want = orc.vardct_reconstruct(p, st, nthreads=2)
    man = []

    def put(name, a, dt):
        a = np.ascontiguousarray(a)
        a.tofile(os.path.join(tmp_path, name))
        man.append({"file": name, "dtype": dt, "shape": [int(a.shape[0]), int(a.shape[1])] if a.ndim == 2 else [1, int(a.size)]})

    def b(x):
        return int(np.array([x], np.float32).view(np.int32)[0])

    for c in range(3):
        put("f0_qcoeff_%d.i32" % c, st["qcoeff"][c], "i32")
        put("f0_lf_%d.f32" % c, st["lf"][c], "f32")
        put("f0_after_color_%d.f32" % c, want[c], "f32")
    for k, ext in (("dct_select", "u8"), ("block_origin", "u8"), ("hf_mul", "i32"), ("sharpness", "i32"), ("x_from_y", "i32"), ("b_from_y", "i32")):
        put("f0_%s.%s" % (k, ext), st[k], ext)
    put("f0_qm_weights.f32", qw, "f32")
    json.dump(man, open(os.path.join(tmp_path, "manifest.json"), "w"))
    pj = {"width": W, "height": H, "global_scale": p.global_scale, "xqm_scale": p.xqm_scale, "bqm_scale": p.bqm_scale,
          "quant_bias": [b(v) for v in p.quant_bias], "quant_bias_numerator": b(p.quant_bias_numerator),
          "color_factor": p.color_factor, "base_corr_x": b(p.base_corr_x), "base_corr_b": b(p.base_corr_b),
          "shift_x": [0, 0, 0], "shift_y": [0, 0, 0], "gab": p.gab, "gab_w1": [b(v) for v in p.gab_w1], "gab_w2": [b(v) for v in p.gab_w2],
          "epf_iters": p.epf_iters, "epf_sharp_lut": [b(v) for v in p.epf_sharp_lut], "epf_channel_scale": [b(v) for v in p.epf_channel_scale],
          "epf_pass0_sigma_scale": b(p.epf_pass0_sigma_scale), "epf_pass2_sigma_scale": b(p.epf_pass2_sigma_scale),
          "epf_border_sad_mul": b(p.epf_border_sad_mul), "do_ycbcr": 0}
    json.dump(pj, open(os.path.join(tmp_path, "f0_params.json"), "w"))
    json.dump({"opsin_matrix": [b(v) for v in p.opsin_matrix], "opsin_bias": [b(v) for v in p.opsin_bias],
               "intensity_target": b(p.intensity_target)}, open(os.path.join(tmp_path, "f0_color.json"), "w"))
    p2, st2, stages = load_frame(str(tmp_path), "f0")
    assert bytes(p2) == bytes(p)
    assert np.array_equal(orc.vardct_reconstruct(p2, st2, nthreads=2), stages["after_color"])

"""Regenerates tests/golden/frontend_pins.json from tests/golden/samples/*.jxl with the C++ front end and the oracle engine.
Run after checking the decoded pictures by eye (tests/tools/decode_with_oracle.py in.jxl out.png)."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from jxlatte_b200 import frontend  # noqa: E402
from jxlatte_b200.decoder import JXLDecoder  # noqa: E402
from oracle_engine import OracleEngine  # noqa: E402


def dg(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


pins = {}
for name in ("lenna", "bbb", "white", "bench", "quilt", "art", "patches-lossless", "blendmodes_5", "wb-rainbow"):
    path = os.path.join(HERE, "samples", name + ".jxl")
    p = frontend.parse_file(path)
    i, f = p.info, p.frames[-1]
    e = {"image": [i["width"], i["height"], i["xyb_encoded"], i["orientation"]],
         "frame": [f["encoding"], f["width"], f["height"], f["gab"], f["epf_iters"], f["num_groups"]]}
    if f["encoding"] == 0:
        st = p.vardct_state(len(p.frames) - 1)
        e["state"] = {k: dg(st[k]) for k in ("qcoeff", "lf", "dct_select", "block_origin", "hf_mul", "sharpness", "x_from_y", "b_from_y")}
    else:
        e["modular"] = [dg(c) for c in p.modular_channels(len(p.frames) - 1)]
    img = JXLDecoder(path, engine=OracleEngine()).decode()
    e["png8"] = dg(img.to_int(8))
    pins[name] = e
json.dump(pins, open(os.path.join(HERE, "frontend_pins.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(pins, indent=1)[:600])

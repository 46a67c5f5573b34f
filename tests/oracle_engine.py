"""TEST INFRASTRUCTURE: a reconstruction back end for jxlatte_b200.decoder.JXLDecoder that runs the CPU oracle instead of
the CUDA library, so the C++ front end and the decoder glue can be exercised on a machine without a GPU, and so the GPU
decode of a real .jxl file can be compared with the oracle's decode of the same file.  Never imported by the package."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

from jxlatte_b200 import host  # noqa: E402


class _OracleModularOps:
    def inverseRCT(self, channels, rct_type):
        return list(orc.modular_rct(channels, rct_type))

    def inversePalette(self, index_channel, palette, nb_deltas, d_pred, bit_depth):
        return list(orc.modular_palette(index_channel, palette, nb_deltas, d_pred, bit_depth))

    def inverseHorizontalSqueeze(self, a, r):
        return orc.modular_squeeze(a, r, True)

    def inverseVerticalSqueeze(self, a, r):
        return orc.modular_squeeze(a, r, False)


class OracleEngine:
    def __init__(self, nthreads=None):
        self.nthreads = nthreads or (os.cpu_count() or 1)

    def qm_default(self):
        return host.qm_generate()          # host-side table build of the product (no GPU involved)

    def qm_generate(self, prm):
        return host.qm_generate(prm)

    def qm_params(self):
        return host.qm_default_params()

    def reconstruct(self, p, st):
        return orc.vardct_reconstruct(p, st, nthreads=self.nthreads)

    def modular(self, channels, transforms, bit_depth):
        return host.ModularTransforms(_OracleModularOps(), bit_depth).applyTransforms(channels, transforms)

    def restore_modular(self, p, planes, sigma):
        x = np.stack([np.asarray(pl, np.float32) for pl in planes])
        if p.gab:
            x = orc.gab(p, x, nthreads=self.nthreads)
        if p.epf_iters:
            x = orc.epf_uniform(p, x, sigma, nthreads=self.nthreads)
        return x

    def color(self, p, planes):
        return orc.color(p, planes, nthreads=self.nthreads)

    def blend(self, op, canvas, a, b, fa, ra):
        orc.blend(op, canvas, a, b, fa, ra)

    def upsample(self, plane, k, weights):
        return orc.upsample(plane, k, weights)

    def pack(self, channels, depths, n_color, linear, bits):
        return orc.pack_samples(channels, depths, n_color, linear, bits)

    def noise(self, planes, group_dim, seed0, lut, base_x, base_b):
        return orc.noise(planes, group_dim, seed0, lut, base_x, base_b)

    def splines(self, planes, splines, quant_adjust, base_x, base_b):
        return orc.splines(planes, splines, quant_adjust, base_x, base_b)

    def close(self):
        pass

"""How sparse are real files' quantised HF coefficients?  (the synthetic generator uses ~2-4 % non-zero, SURVEY.md
8(d) sketched 15 %; k1_big's column walk skips zeros, so the headline depends on which is realistic.)  Counts, per sample file of
the reference, the share of non-zero coefficients outside the LLF corners, overall and for varblocks with a side >= 64.

    python tools/coefficient_density.py > profiles/r2_coefficient_density.md"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from jxlatte_b200 import frontend, synth, default_frame_params
from jxlatte_b200.params import TRANSFORM_TYPES

S = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "samples")


def stats(st):
    q = st["qcoeff"]
    ds, bo = st["dct_select"], st["block_origin"]
    if isinstance(q, (list, tuple)) and any(q[c].shape != q[1].shape for c in range(3)):
        q = [q[1]]                                        # chroma-subsampled: luma only
    H, W = q[0].shape
    llf = np.zeros((H, W), bool)
    big = np.zeros((H, W), bool)
    for by, bx in zip(*np.nonzero(bo)):
        _, _, ph, pw = TRANSFORM_TYPES[ds[by, bx]]
        llf[by * 8:by * 8 + ph // 8, bx * 8:bx * 8 + pw // 8] = True
        if max(ph, pw) >= 64:
            big[by * 8:by * 8 + ph, bx * 8:bx * 8 + pw] = True
    nz = np.zeros((H, W), np.float64)
    for c in range(len(q)):
        nz += (np.asarray(q[c]) != 0)
    nz /= len(q)
    hf = ~llf
    return float(nz[hf].mean()), (float(nz[hf & big].mean()) if (hf & big).any() else None), float(big.mean())


print("# Non-zero share of quantised HF coefficients (LLF corners excluded)\n")
print("Made by `tools/coefficient_density.py` from the reference's sample files through `libjxlfront.so`, and from the synthetic generator.\n")
print("| source | non-zero, all varblocks | non-zero, varblocks with a side >= 64 | area in such varblocks |")
print("|---|---|---|---|")
for name in ("lenna", "bbb", "sollevante-hdr", "ants", "bench"):
    p = frontend.parse_file(os.path.join(S, name + ".jxl"))
    k = max(i for i, f in enumerate(p.frames) if f["encoding"] == 0)
    a, b, area = stats(p.vardct_state(k))
    print("| `%s.jxl` | %.2f %% | %s | %.1f %% |" % (name, 100 * a, "%.2f %%" % (100 * b) if b is not None else "-", 100 * area))
    p.close()
from jxlatte_b200.host import qm_generate      # the product library's tables (bit-identical to the checker's: tests/test_abi.py)
qw, qo = qm_generate()
for label, kw in (("synthetic default (amplitude-aware, the bench headline)", {}), ("synthetic, SURVEY 8(d) as written (`survey_spec=True`)", {"survey_spec": True})):
    pr = default_frame_params(2048, 2048, epf_iters=3)
    st = synth.make_state(2048, 2048, seed=synth.SEED_BASE + 2, params=pr, qm_weights=qw, qm_offsets=qo, **kw)
    a, b, area = stats(st)
    print("| %s | %.2f %% | %.2f %% | %.1f %% |" % (label, 100 * a, 100 * b, 100 * area))
print("\nReading: libjxl's lossy VarDCT samples carry 3.5-8 % non-zero HF coefficients (and use no varblock with a side >= 64 at all);")
print("JPEG-recompressed files (`ants`, `bench`: 8x8 blocks only) 32-42 %.  The default synthetic generator (1.7 %) is SPARSER than the real")
print("lossy files, the 15 % sketch of SURVEY.md 8(d) denser; `bench.py` therefore reports stage 1 at both densities")
print("(`roofline.stage_ms.stage1_at_survey_spec_density_15pct`).  Only the column pass of varblocks with a side >= 64 depends on the density")
print("(it walks non-zero coefficients only); every other kernel of the path does the same work whatever the coefficients are.")

"""Parity at the sizes BASELINE.json names (the benchmarked shapes were never compared with the oracle).

  * configs[2]: one synthetic 7680x4320 frame, all 27 TransformTypes, Gaborish + EPF 3, through the HOST entry point
    (jxlb200_vardct_reconstruct: the call the e2e figure times) against the oracle on all host cores -- bit for bit;
  * configs[4] batch: a stack of 16 frames of 2048x2048 through jxlb200_vardct_reconstruct_batch_dev (what
    `bench.py --workload batch2048` times), every frame against the oracle;
  * configs[4] split: a 16384-wide strip cut into group-row slabs with explicit halo rows (the device path of the
    multi-GPU split; the halo transport is a device copy here) against the oracle on the whole strip;
  * the NCCL split itself (tools/verify_split.py) when the box shows two or more GPUs.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from jxlatte_b200 import synth, default_frame_params, _lib
from jxlatte_b200.host import qm_generate, Slab
from jxlatte_b200.multigpu import slab_rows, split_state

pytestmark = pytest.mark.gpu
HALO = _lib.HALO_ROWS
NTHREADS = os.cpu_count() or 8
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_8k_frame_host_entry_matches_oracle(recon, orc):
    W, H = 7680, 4320
    p = default_frame_params(W, H, epf_iters=3)
    qw, qo = qm_generate()
    st = synth.make_state(W, H, seed=synth.SEED_BASE + 2, params=p, qm_weights=qw, qm_offsets=qo)
    assert len(np.unique(st["dct_select"])) == 27
    got = recon.reconstruct(p, st)
    want = orc.vardct_reconstruct(p, st, nthreads=NTHREADS)
    assert np.array_equal(got, want), "max abs err %g" % np.abs(got - want).max()
    # the int16-coefficient entry point returns the same planes
    assert np.array_equal(recon.reconstruct(p, st, narrow=True), want)


def test_batch_of_16_frames_2048_matches_oracle(recon, orc):
    import torch
    W = H = 2048
    n = 16
    p = default_frame_params(W, H, epf_iters=1)
    qw, qo = qm_generate()
    sts = [synth.make_state(W, H, seed=synth.SEED_BASE + 4 + 16 * f, params=p, qm_weights=qw, qm_offsets=qo) for f in range(n)]
    dev = torch.device("cuda", 0)
    keys = ("qcoeff", "lf", "dct_select", "block_origin", "hf_mul", "sharpness", "x_from_y", "b_from_y")
    d = {k: torch.from_numpy(np.ascontiguousarray(np.concatenate([s[k] for s in sts], axis=-2))).to(dev) for k in keys}
    out = torch.full((3, H * n, W), np.nan, dtype=torch.float32, device=dev)
    recon.reconstruct_batch_dev(p, n, [d["qcoeff"][c].data_ptr() for c in range(3)], [d["lf"][c].data_ptr() for c in range(3)],
                                d["dct_select"].data_ptr(), d["block_origin"].data_ptr(), d["hf_mul"].data_ptr(),
                                d["x_from_y"].data_ptr(), d["b_from_y"].data_ptr(), d["sharpness"].data_ptr(),
                                [out[c].data_ptr() for c in range(3)])
    recon.sync()
    got = out.cpu().numpy()
    for f in range(n):
        want = orc.vardct_reconstruct(p, sts[f], nthreads=NTHREADS)
        assert np.array_equal(got[:, f * H:(f + 1) * H], want), "frame %d of the stack differs from the oracle" % f


def test_16384_wide_strip_as_slabs_matches_oracle(recon, orc):
    """Two slabs of one group row each, 16384 px wide (the width of BASELINE's split config): stage 1 per slab, halo rows
    copied between the slabs on the device, stage 2 per slab with has_top / has_bottom."""
    import torch
    W, H, parts = 16384, 512, 2
    p = default_frame_params(W, H, epf_iters=3)
    qw, qo = qm_generate()
    st = synth.make_state(W, H, seed=synth.SEED_BASE + 5, params=p, qm_weights=qw, qm_offsets=qo)
    want = orc.vardct_reconstruct(p, st, nthreads=NTHREADS)
    dev = torch.device("cuda", 0)
    wb = W // 8
    spans = [slab_rows(H, parts, r) for r in range(parts)]
    xybs, maps, ps = [], [], []
    keep = []
    for (y0, rows) in spans:
        s = split_state(st, y0, rows)
        ps_ = default_frame_params(W, rows, epf_iters=3)
        d = {k: torch.from_numpy(np.ascontiguousarray(s[k])).to(dev) for k in
             ("qcoeff", "lf", "dct_select", "block_origin", "hf_mul", "sharpness", "x_from_y", "b_from_y")}
        xyb = torch.zeros((3, rows + 2 * HALO, W), dtype=torch.float32, device=dev)
        base = [xyb[c].data_ptr() + HALO * W * 4 for c in range(3)]
        recon.invert_dev(ps_, [d["qcoeff"][c].data_ptr() for c in range(3)], [d["lf"][c].data_ptr() for c in range(3)],
                         d["dct_select"].data_ptr(), d["block_origin"].data_ptr(), d["hf_mul"].data_ptr(),
                         d["x_from_y"].data_ptr(), d["b_from_y"].data_ptr(), base, W)
        m = torch.ones((2, rows // 8 + 2, wb), dtype=torch.int32, device=dev)
        m[0, 1:-1] = d["hf_mul"]
        m[1, 1:-1] = d["sharpness"]
        xybs.append(xyb); maps.append(m); ps.append(ps_); keep.append(d)
    recon.sync()
    for i, (y0, rows) in enumerate(spans):
        if i > 0:
            pr = spans[i - 1][1]
            xybs[i][:, :HALO] = xybs[i - 1][:, pr:pr + HALO]
            maps[i][:, 0] = maps[i - 1][:, -2]
        if i < parts - 1:
            xybs[i][:, HALO + rows:] = xybs[i + 1][:, HALO:2 * HALO]
            maps[i][:, -1] = maps[i + 1][:, 1]
    torch.cuda.synchronize()
    got = np.empty((3, H, W), np.float32)
    for i, (y0, rows) in enumerate(spans):
        out = torch.empty((3, rows, W), dtype=torch.float32, device=dev)
        slab = Slab(y0, rows, H, 1 if i > 0 else 0, 1 if i < parts - 1 else 0)
        base = [xybs[i][c].data_ptr() + HALO * W * 4 for c in range(3)]
        recon.restore_dev(ps[i], slab, base, W, maps[i][0].data_ptr() + wb * 4, maps[i][1].data_ptr() + wb * 4,
                          [out[c].data_ptr() for c in range(3)])
        recon.sync()
        got[:, y0:y0 + rows] = out.cpu().numpy()
    assert np.array_equal(got, want), "max abs err %g" % np.abs(got - want).max()


def test_nccl_group_row_split_is_bit_identical():
    """tools/verify_split.py under torchrun on every visible GPU (2, 4 or 8): one frame split by group rows, halo rows over
    NCCL, compared on rank 0 with the whole frame.  Skips on a one-GPU box."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two or more GPUs")
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tools", "verify_split.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "verify_split: OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])

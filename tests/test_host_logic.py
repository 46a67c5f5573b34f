"""Host logic above the C ABI: ModularStream channel bookkeeping and the multi-GPU partitioning."""
import numpy as np

from jxlatte_b200.host import default_squeeze_params, forward_channel_layout
from jxlatte_b200.multigpu import slab_rows, frame_owner, split_state


def test_default_squeeze_params_follow_modularstream():
    # 3 channels of 64 x 48, no meta channels (ModularStream.java:110-131)
    sp = default_squeeze_params([(48, 64)] * 3, 0)
    assert sp[0] == (True, False, 1, 2) and sp[1] == (False, False, 1, 2)   # chroma first, not in place
    rest = sp[2:]
    assert rest[0] == (True, True, 0, 3)          # h < w: no leading vertical step
    hs = sum(1 for s in rest if s[0])
    vs = sum(1 for s in rest if not s[0])
    assert hs == 3 and vs == 3                     # 64 -> 8 and 48 -> 6
    # tall single channel: leading vertical step
    sp = default_squeeze_params([(100, 10)], 0)
    assert sp[0] == (False, True, 0, 1)


def test_forward_channel_layout_shapes():
    sizes = [(5, 9)]
    out = forward_channel_layout(sizes, [(True, True, 0, 1), (False, True, 0, 1)])
    # horizontal: (5, 5) + residual (5, 4); then vertical on channel 0: (3, 5) + residual (2, 5) inserted right after it
    assert out == [(3, 5), (2, 5), (5, 4)]


def test_slab_rows_partition_the_frame_by_group_rows():
    for H in (256, 4320, 16384, 8, 4096 + 8):
        for world in (1, 2, 3, 4, 8):
            spans = [slab_rows(H, world, r) for r in range(world)]
            assert spans[0][0] == 0
            assert sum(r for _, r in spans) == H
            for (y0, rows), (y1, _) in zip(spans, spans[1:]):
                assert y0 + rows == y1
            for y0, rows in spans:
                assert y0 % 256 == 0 or rows == 0
    assert slab_rows(16384, 8, 3) == (3 * 2048, 2048)


def test_frame_owner_round_robin():
    assert [frame_owner(i, 4) for i in range(6)] == [0, 1, 2, 3, 0, 1]


def test_split_state_cuts_every_map_consistently():
    from jxlatte_b200 import synth
    st = synth.make_state(64, 512, seed=3, mix="dct8")
    s = split_state(st, 256, 256)
    assert s["qcoeff"].shape == (3, 256, 64) and s["lf"].shape == (3, 32, 8) and s["x_from_y"].shape == (4, 1)
    assert np.array_equal(s["hf_mul"], st["hf_mul"][32:])


def test_host_entry_slab_schedule_partitions_the_frame():
    """jxlb200_host_slab_schedule (the pipelined host entry point's slab plan; needs no device): slabs start on group rows,
    cover the frame once, and tall frames get two one-group-row slabs at each end so the pipeline fills and drains fast."""
    import ctypes as C
    from jxlatte_b200 import _lib
    L = _lib.lib()
    for H in (8, 64, 256, 264, 512, 776, 1032, 2048, 2056, 4320, 16384, 65536):
        buf = (C.c_int32 * 512)()
        n = L.jxlb200_host_slab_schedule(H, buf, 512)
        assert n >= 1
        starts = list(buf[:n])
        assert starts[0] == 0 and all(s % 256 == 0 for s in starts) and starts == sorted(set(starts)) and starts[-1] < H
        rows = [b - a for a, b in zip(starts, starts[1:] + [H])]
        assert sum(rows) == H and all(r >= 8 and r % 8 == 0 for r in rows)
        assert max(rows) <= 512
        if H >= 6 * 256:
            assert rows[0] == rows[1] == 256 and rows[-2] == 256 and rows[-1] <= 256
    assert L.jxlb200_host_slab_schedule(4320, None, 0) == 11          # count only
    assert L.jxlb200_host_slab_schedule(12, None, 0) == _lib.E_ARG    # heights are padded to 8
    assert L.jxlb200_host_slab_schedule(0, None, 0) == _lib.E_ARG


def test_host_entry_stage2_ranges_follow_the_slabs_by_one_halo():
    """jxlb200_host_stage2_ranges (needs no device): after stage 1 of slab i the host entry point runs stage 2 on the slab
    shifted up by the 8 halo rows, so a range never reads stage-1 rows of a slab that has not been uploaded yet, and the ranges
    cover the frame exactly once."""
    import ctypes as C
    from jxlatte_b200 import _lib
    L = _lib.lib()
    for H in (8, 64, 256, 264, 512, 520, 776, 1032, 2048, 2056, 4320, 16384):
        starts = (C.c_int32 * 512)()
        n = L.jxlb200_host_slab_schedule(H, starts, 512)
        a, b = (C.c_int32 * 512)(), (C.c_int32 * 512)()
        assert L.jxlb200_host_stage2_ranges(H, a, b, 512) == n
        s = list(starts[:n]) + [H]
        assert a[0] == 0 and b[n - 1] == H
        for i in range(n):
            assert b[i] - a[i] >= 8 and a[i] % 8 == 0 and b[i] % 8 == 0
            if i:
                assert a[i] == b[i - 1] == s[i] - _lib.HALO_ROWS          # contiguous, one halo above the slab's first row
            # rows the range reads (itself plus HALO_ROWS above and below, clipped to the frame) are stage-1 output of slabs <= i
            assert min(b[i] + _lib.HALO_ROWS, H) <= s[i + 1]
    assert L.jxlb200_host_stage2_ranges(12, None, None, 0) == _lib.E_ARG


def test_narrow_coefficients_checks_the_range_and_keeps_int16_planes():
    """host.narrow_coefficients feeds jxlb200_vardct_reconstruct_i16: int16 planes go through without a copy, int32 planes
    are narrowed only when every coefficient fits."""
    import numpy as np
    import pytest
    from jxlatte_b200.host import narrow_coefficients
    rng = np.random.default_rng(5)
    q = [rng.integers(-32768, 32768, size=(16, 24), dtype=np.int32) for _ in range(3)]
    n = narrow_coefficients(q)
    assert all(a.dtype == np.int16 and a.flags.c_contiguous and np.array_equal(a, b) for a, b in zip(n, q))
    again = narrow_coefficients(n)
    assert all(a is b or np.shares_memory(a, b) for a, b in zip(again, n))
    stacked = np.stack(n)                                    # one (3, H, W) array works like a list of planes
    assert all(np.shares_memory(a, stacked) for a in narrow_coefficients(stacked))
    for bad in (32768, -32769):
        w = [a.copy() for a in q]
        w[2][3, 4] = bad
        with pytest.raises(ValueError):
            narrow_coefficients(w)
    with pytest.raises(ValueError):
        narrow_coefficients([a.astype(np.float32) for a in q])


def test_locate_view_finds_the_plane_of_a_blend_rectangle():
    """The batched compositing call names rectangles as (plane, y, x): recovered from numpy views, through [C, H, W] owners too."""
    import numpy as np
    from jxlatte_b200.decoder import locate_view
    a = np.arange(40 * 50, dtype=np.float32).reshape(40, 50)
    pl, y, x = locate_view(a[7:19, 11:30])
    assert pl.ctypes.data == a.ctypes.data and pl.shape == (40, 50) and (y, x) == (7, 11)
    b = np.zeros((3, 16, 24), np.int32)
    pl, y, x = locate_view(b[2][5:9, 3:20])
    assert pl.ctypes.data == b.ctypes.data and pl.shape == (48, 24) and (y, x) == (32 + 5, 3)      # [C, H, W] owner = one tall plane
    pl, y, x = locate_view(np.ascontiguousarray(b[1][:16, :24])[0:16, 0:24])     # already contiguous: still a view of b
    assert pl.shape == (48, 24) and (y, x) == (16, 0)
    pl, y, x = locate_view(a[::2, :])                     # every other row IS a window: of the owner seen 100 samples wide
    assert pl.shape == (20, 100) and (y, x) == (0, 0) and np.array_equal(pl[:, :50], a[::2, :])
    assert locate_view(a[:, ::2]) is None
    assert locate_view(a.astype(np.float64)[1:3, 1:3]) is None
    assert locate_view(a.T[1:3, 1:3]) is None

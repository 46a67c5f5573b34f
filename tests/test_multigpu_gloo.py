"""world_size-2 (and 3) run of the group-row split's halo exchange on CPU tensors over gloo: the same
exchange_halos() the NCCL path uses."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jxlatte_b200.multigpu import exchange_halos, slab_rows


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, halo, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(3 * H * W, dtype=torch.float32).reshape(3, H, W)   # every rank knows the whole frame
        y0, rows = slab_rows(H, world, rank)
        ext = torch.full((3, rows + 2 * halo, W), -1.0)
        ext[:, halo:halo + rows] = full[:, y0:y0 + rows]
        exchange_halos(ext, halo, rank, world)
        ok = True
        if rank > 0:
            ok &= bool(torch.equal(ext[:, :halo], full[:, y0 - halo:y0]))
        else:
            ok &= bool((ext[:, :halo] == -1).all())          # frame top: untouched, the kernel mirrors instead
        if rank < world - 1:
            ok &= bool(torch.equal(ext[:, halo + rows:], full[:, y0 + rows:y0 + rows + halo]))
        else:
            ok &= bool((ext[:, halo + rows:] == -1).all())
        ok &= bool(torch.equal(ext[:, halo:halo + rows], full[:, y0:y0 + rows]))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 1024, 40, 8, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(r, True) for r in range(world)]

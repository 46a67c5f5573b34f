"""Multi-GPU partitioning of the VarDCT reconstruction path.  One process per GPU.

Two ways the path shards, exactly as the north star states:
  * a batch of frames: frame i goes to rank i % world -- no communication at all (bench.py --workload 8k / batch2048);
  * one very large frame: contiguous GROUP ROWS (256 px) per rank.  Stage 1 needs nothing from the neighbours (varblocks
    never cross a group boundary); stage 2 needs JXLB200_HALO_ROWS rows of the neighbours' stage-1 output above and below
    (Gaborish 1 + EPF 3 + 2 + 1 = 7 used) plus one block row of hf_mul / sharpness.  On GPUs the exchange is inside the
    library (jxlb200_vardct_reconstruct_split_dev: ncclSend / ncclRecv over NVLink on a side stream, overlapped with the
    slab's own work); exchange_halos() below is the same hand-over on torch.distributed tensors, kept for the gloo/CPU
    tests of the row arithmetic.  The true frame top and bottom mirror inside the kernel instead.
"""
import numpy as np

from ._lib import HALO_ROWS, Slab

GROUP = 256


def slab_rows(height, world, rank):
    """Rows [y0, y0 + rows) of a frame of padded `height` owned by `rank`: contiguous group rows, the first
    (groups % world) ranks get one more.  Ranks beyond the number of group rows get rows == 0."""
    groups = (height + GROUP - 1) // GROUP
    base, rem = divmod(groups, world)
    g0 = rank * base + min(rank, rem)
    g1 = g0 + base + (1 if rank < rem else 0)
    y0 = min(height, g0 * GROUP)
    y1 = min(height, g1 * GROUP)
    return y0, y1 - y0


def frame_owner(frame_index, world):
    """Per-image sharding of a batch."""
    return frame_index % world


def exchange_halos(ext, halo, rank, world, group=None):
    """ext: tensor [..., halo + rows + halo, W] whose interior rows are filled.  Fills the top halo with the last
    `halo` interior rows of rank-1 and the bottom halo with the first `halo` interior rows of rank+1.
    Returns after the transfers are complete from the caller's stream's point of view."""
    import torch
    import torch.distributed as dist
    rows = ext.shape[-2] - 2 * halo
    ops = []
    keep = []
    if rank > 0:
        send_up = ext[..., halo:2 * halo, :].contiguous()          # my first rows -> bottom halo of rank-1
        recv_up = torch.empty_like(send_up)                       # last rows of rank-1 -> my top halo
        ops += [dist.P2POp(dist.isend, send_up, rank - 1, group), dist.P2POp(dist.irecv, recv_up, rank - 1, group)]
        keep.append(("top", recv_up))
    if rank < world - 1:
        send_dn = ext[..., rows:rows + halo, :].contiguous()       # my last rows -> top halo of rank+1
        recv_dn = torch.empty_like(send_dn)
        ops += [dist.P2POp(dist.isend, send_dn, rank + 1, group), dist.P2POp(dist.irecv, recv_dn, rank + 1, group)]
        keep.append(("bottom", recv_dn))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for side, t in keep:
        if side == "top":
            ext[..., 0:halo, :].copy_(t)
        else:
            ext[..., halo + rows:, :].copy_(t)
    return ext


class SplitFrame:
    """One rank's slab of a frame split by group rows; step() = stage 1, halo rows to / from the neighbours, stage 2.

    The whole step is ONE C-ABI call, jxlb200_vardct_reconstruct_split_dev: the library owns the communicator (NCCL, joined
    here with an id rank 0 makes and torch.distributed merely carries to the other ranks), sends the boundary rows on its own
    stream right after the boundary group rows' stage 1 and runs the rest of the slab meanwhile (csrc/split_nccl.cuh).  It
    enqueues on the Reconstructor's stream; bind that to the stream the caller times or synchronises (rec.set_stream)."""

    def __init__(self, rec, p_slab, dev_state, y0, rows, frame_height, rank, world, device):
        import torch
        import torch.distributed as dist
        self.rec, self.p, self.d = rec, p_slab, dev_state
        self.rank, self.world = rank, world
        self.slab = Slab(y0, rows, frame_height, 1 if rank > 0 else 0, 1 if rank < world - 1 else 0)
        self.out = dev_state["out"]
        if world > 1:
            idt = torch.zeros(128, dtype=torch.uint8, device=device)
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(rec.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            rec.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)

    def step(self):
        d = self.d
        self.rec.reconstruct_split_dev(self.p, self.slab, [d["qcoeff"][c].data_ptr() for c in range(3)],
                                       [d["lf"][c].data_ptr() for c in range(3)], d["dct_select"].data_ptr(), d["block_origin"].data_ptr(),
                                       d["hf_mul"].data_ptr(), d["x_from_y"].data_ptr(), d["b_from_y"].data_ptr(), d["sharpness"].data_ptr(),
                                       [self.out[c].data_ptr() for c in range(3)])


def split_state(st, y0, rows):
    """Cut a frame-level synthetic state (numpy dict) down to the slab [y0, y0 + rows)."""
    b0, b1 = y0 // 8, (y0 + rows) // 8
    t0, t1 = y0 // 64, (y0 + rows + 63) // 64
    out = dict(st)
    out["qcoeff"] = np.ascontiguousarray(st["qcoeff"][:, y0:y0 + rows])
    out["lf"] = np.ascontiguousarray(st["lf"][:, b0:b1])
    for k in ("dct_select", "block_origin", "hf_mul", "sharpness"):
        out[k] = np.ascontiguousarray(st[k][b0:b1])
    for k in ("x_from_y", "b_from_y"):
        out[k] = np.ascontiguousarray(st[k][t0:t1])
    out["height"] = rows
    return out

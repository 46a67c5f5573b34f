"""Group-row split across N GPUs with NCCL halo exchange, checked bit-for-bit against the whole frame on rank 0.
Run:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/verify_split.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from jxlatte_b200 import synth, default_frame_params
from jxlatte_b200.host import Reconstructor, qm_generate
from jxlatte_b200.multigpu import slab_rows, split_state, SplitFrame

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok_all = True
for (W, H, iters, gab) in ((520, 256 * world * 2 + 8, 3, True), (1024, 256 * world, 1, True), (264, 256 * world + 256, 2, False)):
    p = default_frame_params(W, H, epf_iters=iters, gab=gab)
    qw, qo = qm_generate()
    st = synth.make_state(W, H, seed=123 + W, params=p, qm_weights=qw, qm_offsets=qo)    # same frame on every rank
    y0, rows = slab_rows(H, world, rank)
    rec = Reconstructor(local)
    s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s); rec.set_stream(s.cuda_stream)
    rec.setWeights(qw, qo)
    mine = split_state(st, y0, rows)
    ps = default_frame_params(W, rows, epf_iters=iters, gab=gab)
    d = {k: torch.from_numpy(np.ascontiguousarray(mine[k])).to(dev) for k in
         ("qcoeff", "lf", "dct_select", "block_origin", "hf_mul", "sharpness", "x_from_y", "b_from_y")}
    d["out"] = torch.empty((3, rows, W), dtype=torch.float32, device=dev)
    sf = SplitFrame(rec, ps, d, y0, rows, H, rank, world, dev)
    sf.step()
    rec.sync()
    torch.cuda.synchronize()
    parts = [torch.empty((3, slab_rows(H, world, r)[1], W), dtype=torch.float32, device=dev) for r in range(world)]
    # all_gather needs equal shapes: pad to the largest slab
    mx = max(t.shape[1] for t in parts)
    padded = torch.zeros((3, mx, W), dtype=torch.float32, device=dev); padded[:, :rows] = d["out"]
    gathered = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded)
    if rank == 0:
        full = np.concatenate([gathered[r][:, :slab_rows(H, world, r)[1]].cpu().numpy() for r in range(world)], axis=1)
        rec.set_stream(None)
        whole = rec.reconstruct(p, st)
        same = np.array_equal(full, whole)
        ok_all &= same
        print("split %dx%d over %d GPUs, gab=%s EPF=%d: %s (max abs diff %g)" % (W, H, world, gab, iters, "bit-identical to the whole frame" if same else "MISMATCH", np.abs(full - whole).max()))
    rec.close()
dist.barrier()
if rank == 0:
    print("verify_split:", "OK" if ok_all else "FAILED")
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)

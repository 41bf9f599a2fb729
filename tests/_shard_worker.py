"""Worker for tests/test_sharding.py: one rank of a world_size-N gloo job that
emulates the GPU path's y-slab stepping on the CPU.

The sharding LOGIC under test is the product's: chemsim_lbm_slab_rows and
chemsim_lbm_halo_plan from libchemsim_lbm.so decide which rows a rank owns and
which (peer, population, row) messages it issues per step, in which order.  The
arithmetic is done by the oracle (the checker), the transport by gloo instead of
NCCL.  Rank 0 gathers the slabs and compares them bit for bit with the
unsharded oracle run.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from chemsim_b200 import _ffi, lbm, scenarios  # noqa: E402
from oracle import lbm_oracle as O  # noqa: E402


def exchange(planes, plan, h):
    """planes: (9, h+2, w) with ghost rows 0 and h+1.  Executes the plan in order."""
    row_index = {_ffi.ROW_FIRST: 1, _ffi.ROW_LAST: h, _ffi.ROW_GHOST_ABOVE: 0, _ffi.ROW_GHOST_BELOW: h + 1}
    ops, recvs = [], []
    for is_send, peer, q, row in plan:
        if is_send:
            t = torch.from_numpy(planes[q, row_index[row]].copy())
            ops.append(dist.P2POp(dist.isend, t, peer))
        else:
            t = torch.empty(planes.shape[2], dtype=torch.from_numpy(planes).dtype)
            ops.append(dist.P2POp(dist.irecv, t, peer))
            recvs.append((q, row_index[row], t))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for q, r, t in recvs:
        planes[q, r] = t.numpy()


def main():
    w, hg, steps, edge, dtype_name = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    dtype = np.float32 if dtype_name == "f32" else np.float64
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    r0, h = lbm.slab_rows(hg, rank, world)
    plan = lbm.halo_plan(rank, world, edge)
    rho, vx, vy, solid = scenarios.random_state(w, hg, dtype, seed=17)
    f_full = O.compute_equilibrium(rho, vx, vy)
    cur = np.zeros((9, h + 2, w), dtype=dtype)       # zero ghost rows = the zero-fill edge
    cur[:, 1:h + 1] = f_full[:, r0:r0 + h]
    nxt = np.zeros_like(cur)
    my_solid = np.ascontiguousarray(solid[r0:r0 + h])
    exchange(cur, plan, h)                           # ensure_ghosts()
    for _ in range(steps):
        O.step_fused_slab(cur, nxt, my_solid, edge, 0.8)
        exchange(nxt, plan, h)
        cur, nxt = nxt, cur
    mine = torch.from_numpy(np.ascontiguousarray(cur[:, 1:h + 1]))
    gathered = [None] * world
    dist.gather_object(mine.numpy(), gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        got = np.concatenate(gathered, axis=1)
        ref = O.step_fused(f_full, solid, steps, 0.8, edge)
        u = np.uint32 if dtype == np.float32 else np.uint64
        ok = bool((got.view(u) == ref.view(u)).all())
        print("SHARD_OK" if ok else "SHARD_MISMATCH", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

"""Worker for tests/test_sharding.py: one rank of a world_size-N gloo job that
emulates the GPU path's y-slab stepping on the CPU.

The sharding LOGIC under test is the product's: chemsim_lbm_slab_rows and
chemsim_lbm_halo_plan from libchemsim_lbm.so decide which rows a rank owns and
which (peer, population, row) messages it issues per exchange, in which order; the
emulation follows the host runtime's schedule (lattice.cu: step_impl): pairs of steps
as ONE pass over the slab extended by its two ghost rows, an odd step on its own, one
exchange per pass, the geometry's halo rows refreshed before every batch.
The arithmetic is done by the oracle (the checker), the transport by gloo instead of
NCCL.  Rank 0 gathers the slabs and compares them bit for bit with the unsharded
oracle run.
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

G = 2   # ghost rows per side (kernels.cuh: GHOST)


def row_index(h):
    return {_ffi.ROW_FIRST: G, _ffi.ROW_SECOND: G + 1, _ffi.ROW_LAST: G + h - 1, _ffi.ROW_SECOND_LAST: G + h - 2,
            _ffi.ROW_GHOST_ABOVE: G - 1, _ffi.ROW_GHOST_ABOVE2: G - 2, _ffi.ROW_GHOST_BELOW: G + h,
            _ffi.ROW_GHOST_BELOW2: G + h + 1}


def exchange(planes, plan, h):
    """planes: (9, h+2G, w).  Executes the plan in order."""
    idx = row_index(h)
    ops, recvs = [], []
    for is_send, peer, q, row in plan:
        if is_send:
            t = torch.from_numpy(planes[q, idx[row]].copy())
            ops.append(dist.P2POp(dist.isend, t, peer))
        else:
            t = torch.empty(planes.shape[2], dtype=torch.from_numpy(planes).dtype)
            ops.append(dist.P2POp(dist.irecv, t, peer))
            recvs.append((q, idx[row], t))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for q, r, t in recvs:
        planes[q, r] = t.numpy()


def exchange_mask(mask, rank, world, periodic, h):
    """mask: (h+2, w) with one halo row per side (lattice.cu: exchange_mask)."""
    up, down = (rank - 1) % world, (rank + 1) % world
    has_up, has_down = periodic or rank > 0, periodic or rank < world - 1
    ops, recvs = [], []
    if has_down:
        ops.append(dist.P2POp(dist.isend, torch.from_numpy(mask[h].copy()), down))
    if has_up:
        ops.append(dist.P2POp(dist.isend, torch.from_numpy(mask[1].copy()), up))
    if has_up:
        t = torch.empty(mask.shape[1], dtype=torch.uint8)
        ops.append(dist.P2POp(dist.irecv, t, up))
        recvs.append((0, t))
    if has_down:
        t = torch.empty(mask.shape[1], dtype=torch.uint8)
        ops.append(dist.P2POp(dist.irecv, t, down))
        recvs.append((h + 1, t))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for r, t in recvs:
        mask[r] = t.numpy()


def single_step(cur, nxt, mask, h, edge):
    """rows 0..h-1 from the inner ghost rows: the slab with ONE ghost row is a view of the array."""
    src = np.ascontiguousarray(cur[:, G - 1:G + h + 1])
    dst = np.zeros_like(src)
    O.step_fused_slab(src, dst, np.ascontiguousarray(mask[1:h + 1]), edge, 0.8)
    nxt[:, G:G + h] = dst[:, 1:h + 1]


def double_step(cur, nxt, mask, h, edge, r0, hg, periodic):
    """step2_impl.cuh on the CPU: step n+1 on rows -1..h (the slab plus one rim row per side, read
    from both ghost rows), step n+2 on rows 0..h-1 from that."""
    mid = np.zeros_like(cur)
    O.step_fused_slab(cur, mid, np.ascontiguousarray(mask), edge, 0.8)      # (9, (h+2)+2, w): height h+2, one ghost row
    if not periodic:                      # rim rows outside a zero-fill lattice hold 0 (never computed)
        if r0 == 0:
            mid[:, G - 1] = 0
        if r0 + h == hg:
            mid[:, G + h] = 0
    single_step(mid, nxt, mask, h, edge)


def main():
    w, hg, steps, edge, dtype_name = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    dtype = np.float32 if dtype_name == "f32" else np.float64
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    periodic = edge == _ffi.EDGE_PERIODIC
    r0, h = lbm.slab_rows(hg, rank, world)
    plan = lbm.halo_plan(hg, rank, world, edge)
    deep = hg // world >= 2
    rho, vx, vy, solid = scenarios.random_state(w, hg, dtype, seed=17)
    f_full = O.compute_equilibrium(rho, vx, vy)
    cur = np.zeros((9, h + 2 * G, w), dtype=dtype)   # zero ghost rows = the zero-fill edge
    cur[:, G:G + h] = f_full[:, r0:r0 + h]
    nxt = np.zeros_like(cur)
    mask = np.zeros((h + 2, w), np.uint8)
    mask[1:h + 1] = solid[r0:r0 + h]
    done = 0
    # batches like the callers': step(1), step(2), step(rest) — every batch refreshes the mask halo
    for batch in (1, 2, steps - 3):
        if deep:
            exchange_mask(mask, rank, world, periodic, h)
        if done == 0:
            exchange(cur, plan, h)                   # begin_sharded: ghosts of the uploaded state
        left = batch
        while left > 0:
            n = 2 if (left >= 2 and deep) else 1
            if n == 2:
                double_step(cur, nxt, mask, h, edge, r0, hg, periodic)
            else:
                single_step(cur, nxt, mask, h, edge)
            exchange(nxt, plan, h)
            cur, nxt = nxt, cur
            left -= n
            done += n
    assert done == steps
    mine = np.ascontiguousarray(cur[:, G:G + h])
    gathered = [None] * world
    dist.gather_object(mine, gathered if rank == 0 else None, dst=0)
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

"""Worker for the multi-GPU parity test: one rank per GPU (torchrun, NCCL).
Each rank steps its y-slab through the C ABI; rank 0 gathers all slabs and compares
them bit for bit with (a) the oracle and (b) an unsharded run on its own GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from chemsim_b200 import lbm, scenarios  # noqa: E402
from oracle import lbm_oracle as O  # noqa: E402


def main():
    w, hg, steps, edge, dtype_name = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    dtype = np.float32 if dtype_name == "f32" else np.float64
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ident = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        ident = torch.frombuffer(bytearray(lbm.nccl_unique_id()), dtype=torch.uint8).to(dev)
    dist.broadcast(ident, 0)
    state = lbm.State.create((w, hg), lbm.BGK(0.8), dtype=dtype, edge=edge, device=local, rank=rank, nranks=world,
                             nccl_id=ident.cpu().numpy().tobytes())
    halo = sys.argv[6] if len(sys.argv) > 6 else "nccl"
    if halo == "p2p":
        assert state.enable_p2p_halo(), "peer-memory halo could not be enabled on this box"
    assert state.halo_mode() == halo
    r0, h = state.row_offset, state.local_height
    assert (r0, h) == lbm.slab_rows(hg, rank, world)
    rho, vx, vy, solid = scenarios.random_state(w, hg, dtype, seed=23)
    state.init_equilibrium(rho[r0:r0 + h], vx[r0:r0 + h], vy[r0:r0 + h])
    state.geometry = solid[r0:r0 + h]
    m0 = state.total_mass(global_=True)
    # mix single steps and batches so that every dependency edge of the pipeline is used
    done = 0
    for n in (1, 2, steps - 3):
        state.step(n)
        done += n
    state.synchronize()          # also reports a halo time-out
    mine = state.populations_array()
    mass = state.total_mass(global_=True)
    image = state.render(state.RENDER_SPEED)          # collective: mean/stdev all-reduced over the slabs
    gathered = [None] * world
    dist.gather_object(mine, gathered if rank == 0 else None, dst=0)
    images = [None] * world
    dist.gather_object(image, images if rank == 0 else None, dst=0)
    # f-3 / f-4 on a sharded lattice: every rank paints the same GLOBAL rectangle (it straddles a
    # slab face: rows around hg/2 and around the lattice's first rows), checkpoints, steps, restores
    # and replays; rank 0 compares with the oracle run on the host-edited mask.
    ry, rh = max(0, hg // 2 - 3), min(7, hg - max(0, hg // 2 - 3))
    state.paint_rect(w // 4, ry, max(1, w // 2), rh, True)
    state.paint_rect(-3, -2, 9, 4, True)
    state.step(2)
    blob = state.checkpoint()
    state.step(3)
    first = state.populations_array()
    state.restore(blob)
    state.step(3)
    state.synchronize()
    again = state.populations_array()
    replay_ok = bool((first.view(np.uint8) == again.view(np.uint8)).all()) and abs(state.time - (steps + 5)) < 1e-6
    painted = [None] * world
    dist.gather_object((again, state.geometry, replay_ok), painted if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        got = np.concatenate(gathered, axis=1)
        f0 = O.compute_equilibrium(rho, vx, vy)
        ref = O.step_fused(f0, solid, steps, 0.8, edge)
        u = np.uint32 if dtype == np.float32 else np.uint64
        ok = bool((got.view(u) == ref.view(u)).all())
        single = lbm.State.create((w, hg), lbm.BGK(0.8), dtype=dtype, edge=edge, device=local)
        single.init_equilibrium(rho, vx, vy)
        single.geometry = solid
        single.step(steps)
        single.synchronize()
        ok = ok and bool((single.populations_array().view(u) == got.view(u)).all())
        whole = single.render(single.RENDER_SPEED).astype(np.int16)
        ok = ok and int(np.abs(np.concatenate(images, axis=0).astype(np.int16) - whole).max()) <= 1
        ok = ok and abs(mass - O.total_mass(ref)) <= 1e-12 * abs(mass)
        solid2 = solid.copy()
        solid2[ry:ry + rh, w // 4:w // 4 + max(1, w // 2)] = 1
        solid2[0:2, 0:6] = 1
        ref2 = O.step_fused(ref, solid2, 5, 0.8, edge)
        ok = ok and all(p[2] for p in painted)
        ok = ok and bool((np.concatenate([p[1] for p in painted], axis=0) == solid2.astype(bool)).all())
        ok = ok and bool((np.concatenate([p[0] for p in painted], axis=1).view(u) == ref2.view(u)).all())
        ok = ok and abs(m0 - O.total_mass(f0)) <= 1e-12 * abs(m0)
        print(("MULTIGPU_OK" if ok else "MULTIGPU_MISMATCH") + f" world={world} {w}x{hg} edge={edge} {dtype_name} halo={halo}", flush=True)
    state.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

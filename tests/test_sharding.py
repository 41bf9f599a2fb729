"""CPU tests of the multi-GPU host logic (world_size 2 and 3 over gloo): the slab
partition and the halo plan exported by the C-ABI library, executed by an emulation
whose result must equal the unsharded oracle bit for bit."""
import os
import subprocess
import sys

import pytest

from chemsim_b200 import _ffi, lbm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_rows_partition_the_lattice():
    for hg, n in ((32768, 8), (16384 * 8, 8), (10, 3), (7, 7), (4097, 4)):
        rows = [lbm.slab_rows(hg, r, n) for r in range(n)]
        assert rows[0][0] == 0
        for a, b in zip(rows, rows[1:]):
            assert a[0] + a[1] == b[0] and a[1] >= 1
        assert rows[-1][0] + rows[-1][1] == hg
    with pytest.raises(lbm.LbmError):
        lbm.slab_rows(3, 0, 4)


def test_halo_plan_pairs_up():
    """Every send has exactly one matching receive on the peer, in the same order per
    (sender, receiver) pair — also when both neighbours are the same rank."""
    pairs = {(_ffi.ROW_LAST, _ffi.ROW_GHOST_ABOVE), (_ffi.ROW_FIRST, _ffi.ROW_GHOST_BELOW),
             (_ffi.ROW_SECOND_LAST, _ffi.ROW_GHOST_ABOVE2), (_ffi.ROW_SECOND, _ffi.ROW_GHOST_BELOW2)}
    for n in (2, 3, 8):
        for hg, per_face in ((64, 12), (n, 3)):            # two-row halo / one row per rank: the one-row form
            for edge in (_ffi.EDGE_ZEROFILL, _ffi.EDGE_PERIODIC):
                plans = [lbm.halo_plan(hg, r, n, edge) for r in range(n)]
                assert all(len(p) <= _ffi.HALO_PLAN_MAX for p in plans)
                for a in range(n):
                    for b in range(n):
                        sends = [(q, row) for s, peer, q, row in plans[a] if s and peer == b]
                        recvs = [(q, row) for s, peer, q, row in plans[b] if not s and peer == a]
                        assert len(sends) == len(recvs)
                        for (qs, rs), (qr, rr) in zip(sends, recvs):
                            assert qs == qr and (rs, rr) in pairs
                count = sum(len(p) for p in plans)
                faces = n if edge == _ffi.EDGE_PERIODIC else n - 1
                assert count == faces * 2 * per_face * 2   # faces x directions x messages x (send+recv)
    # the outer row of a face carries all nine populations, the next one only the movers towards it
    plan = lbm.halo_plan(64, 1, 3, _ffi.EDGE_ZEROFILL)
    assert sorted(q for s, _, q, row in plan if s and row == _ffi.ROW_LAST) == list(range(9))
    assert sorted(q for s, _, q, row in plan if s and row == _ffi.ROW_SECOND_LAST) == [3, 6, 7]
    assert sorted(q for s, _, q, row in plan if s and row == _ffi.ROW_SECOND) == [1, 5, 8]
    assert lbm.halo_plan(64, 0, 1, _ffi.EDGE_PERIODIC) == []


@pytest.mark.parametrize("world,w,hg,edge,dtype", [
    (2, 24, 13, _ffi.EDGE_PERIODIC, "f32"),
    (2, 20, 8, _ffi.EDGE_ZEROFILL, "f64"),
    (3, 16, 10, _ffi.EDGE_PERIODIC, "f64"),
    (3, 12, 7, _ffi.EDGE_ZEROFILL, "f32"),      # two-row slabs: every row is a face row
    (3, 12, 3, _ffi.EDGE_ZEROFILL, "f32"),      # one row per rank: the one-row halo, single steps only
])
def test_gloo_emulation_matches_unsharded_oracle(world, w, hg, edge, dtype):
    port = 29500 + (os.getpid() + world * 7 + hg) % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "_shard_worker.py"), str(w), str(hg), "8", str(edge), dtype]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "SHARD_OK" in res.stdout


@pytest.mark.parametrize("world,edge", [(2, _ffi.EDGE_PERIODIC), (3, _ffi.EDGE_PERIODIC), (4, _ffi.EDGE_ZEROFILL)])
def test_peer_memory_flag_protocol_model(world, edge):
    """Executable model of the fused peer-memory halo: ranks are threads, "peer
    memory" is shared numpy storage, the arithmetic is the oracle's slab step.  A pass advances
    the slab by one step or by two (step2_impl.cuh); each rank waits until both neighbours have
    published step t, updates its slab from buffer A into B, stores its face rows into the
    NEIGHBOURS' ghost rows of B — outermost row: all nine populations, next row: the three that move
    towards the face — and then publishes t + (steps of the pass).  Random delays shake the
    interleavings; the invariant under test is that the A-B buffers plus one step counter per face
    are enough (a neighbour is never more than one PASS ahead), i.e. the result equals the unsharded
    run bit for bit and nothing deadlocks."""
    import threading
    import time

    import numpy as np

    from chemsim_b200 import scenarios
    from oracle import lbm_oracle as O

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _shard_worker as W

    G = W.G
    dtype, w, hg, steps = np.float32, 16, 4 * world + 1, 25
    rho, vx, vy, solid = scenarios.random_state(w, hg, dtype, seed=5)
    f0 = O.compute_equilibrium(rho, vx, vy)
    slabs = [lbm.slab_rows(hg, r, world) for r in range(world)]
    periodic = edge == _ffi.EDGE_PERIODIC
    up = [(r - 1) % world if (periodic or r > 0) else None for r in range(world)]
    down = [(r + 1) % world if (periodic or r < world - 1) else None for r in range(world)]
    bufs, masks = [], []                        # bufs[r][parity] : (9, h+2G, w) with ghost rows
    for r0, h in slabs:
        a = np.zeros((9, h + 2 * G, w), dtype)
        a[:, G:G + h] = f0[:, r0:r0 + h]
        bufs.append([a, np.zeros_like(a)])
        m = np.zeros((h + 2, w), np.uint8)
        m[1:h + 1] = solid[r0:r0 + h]
        masks.append(m)

    def deliver(r, b):
        """rank r's face rows of buffer b -> its neighbours' ghost rows of buffer b"""
        r0, h = slabs[r]
        src = bufs[r][b]
        if down[r] is not None:
            dst = bufs[down[r]][b]
            dst[:, G - 1] = src[:, G + h - 1]
            dst[[3, 6, 7], G - 2] = src[[3, 6, 7], G + h - 2]
        if up[r] is not None:
            dst, hu = bufs[up[r]][b], slabs[up[r]][1]
            dst[:, G + hu] = src[:, G]
            dst[[1, 5, 8], G + hu + 1] = src[[1, 5, 8], G + 1]

    for r in range(world):                      # the first exchange (NCCL in the product) + the mask halo
        deliver(r, 0)
        if down[r] is not None:
            masks[down[r]][0] = masks[r][slabs[r][1]]
        if up[r] is not None:
            masks[up[r]][slabs[up[r]][1] + 1] = masks[r][1]
    flags = np.zeros((world, 2), dtype=np.int64)      # [rank][0: from_up, 1: from_down], "step published"
    cond = threading.Condition()
    rng = np.random.default_rng(world)
    delays = rng.random((world, steps)) * 2e-3
    errors = []
    passes = [2] * (steps // 2) + [1] * (steps % 2)   # pairs first, the odd step on its own

    def rank(r):
        r0, h = slabs[r]
        try:
            t = 0
            for i, n in enumerate(passes):
                with cond:                      # wait_flag: both neighbours have published step t
                    ok = cond.wait_for(lambda: (up[r] is None or flags[r, 0] >= t) and
                                       (down[r] is None or flags[r, 1] >= t), timeout=20)
                    assert ok, f"rank {r} timed out at step {t}"
                time.sleep(delays[r, i])
                src, dst = bufs[r][i % 2], bufs[r][(i + 1) % 2]
                if n == 2:
                    W.double_step(src, dst, masks[r], h, edge, r0, hg, periodic)
                else:
                    W.single_step(src, dst, masks[r], h, edge)
                deliver(r, (i + 1) % 2)
                t += n
                with cond:                      # publish (after the stores)
                    if down[r] is not None:
                        flags[down[r], 0] = t
                    if up[r] is not None:
                        flags[up[r], 1] = t
                    cond.notify_all()
        except Exception as e:                  # pragma: no cover
            errors.append(e)
            with cond:
                cond.notify_all()

    threads = [threading.Thread(target=rank, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(60)
    assert not errors, errors
    got = np.concatenate([bufs[r][len(passes) % 2][:, G:G + slabs[r][1]] for r in range(world)], axis=1)
    ref = O.step_fused(f0, solid, steps, 0.8, edge)
    np.testing.assert_array_equal(got.view(np.uint32), ref.view(np.uint32))

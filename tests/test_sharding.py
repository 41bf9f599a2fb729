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
    for n in (2, 3, 8):
        for edge in (_ffi.EDGE_ZEROFILL, _ffi.EDGE_PERIODIC):
            plans = [lbm.halo_plan(r, n, edge) for r in range(n)]
            for a in range(n):
                for b in range(n):
                    sends = [(q, row) for s, peer, q, row in plans[a] if s and peer == b]
                    recvs = [(q, row) for s, peer, q, row in plans[b] if not s and peer == a]
                    assert len(sends) == len(recvs)
                    for (qs, rs), (qr, rr) in zip(sends, recvs):
                        assert qs == qr
                        assert (rs, rr) in ((_ffi.ROW_LAST, _ffi.ROW_GHOST_ABOVE), (_ffi.ROW_FIRST, _ffi.ROW_GHOST_BELOW))
            count = sum(len(p) for p in plans)
            faces = n if edge == _ffi.EDGE_PERIODIC else n - 1
            assert count == faces * 2 * 3 * 2       # faces x directions x 3 populations x (send+recv)
    assert lbm.halo_plan(0, 1, _ffi.EDGE_PERIODIC) == []


@pytest.mark.parametrize("world,w,hg,edge,dtype", [
    (2, 24, 13, _ffi.EDGE_PERIODIC, "f32"),
    (2, 20, 8, _ffi.EDGE_ZEROFILL, "f64"),
    (3, 16, 10, _ffi.EDGE_PERIODIC, "f64"),
    (3, 12, 3, _ffi.EDGE_ZEROFILL, "f32"),      # one row per rank: every row is a face row
])
def test_gloo_emulation_matches_unsharded_oracle(world, w, hg, edge, dtype):
    port = 29500 + (os.getpid() + world * 7 + hg) % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "_shard_worker.py"), str(w), str(hg), "6", str(edge), dtype]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "SHARD_OK" in res.stdout

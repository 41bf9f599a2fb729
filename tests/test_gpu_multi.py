"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): a y-slab
sharded run must be bit-identical to the unsharded run and to the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    import torch
    return torch.cuda.device_count()


def run_worker(w, hg, edge, dtype, halo):
    world = min(n_gpus(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    port = 29700 + (os.getpid() + hg + len(halo)) % 1000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "_multigpu_worker.py"), str(w), str(hg), "12", str(edge), dtype, halo]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MULTIGPU_OK" in res.stdout


@pytest.mark.parametrize("w,hg,edge,dtype", [(256, 96, 1, "f32"), (260, 50, 0, "f64"), (1024, 64, 1, "f64"),
                                             (37, 21, 1, "f32")])
def test_sharded_equals_unsharded(w, hg, edge, dtype):
    run_worker(w, hg, edge, dtype, "nccl")


@pytest.mark.parametrize("w,hg,edge,dtype", [(256, 96, 1, "f32"), (260, 50, 0, "f64"), (1024, 16, 1, "f64"),
                                             (4096, 40, 0, "f32")])
def test_sharded_with_fused_peer_memory_halo_equals_unsharded(w, hg, edge, dtype):
    """The face kernel stores the halo into the neighbours' ghost rows itself (cudaIpc + NVLink)."""
    run_worker(w, hg, edge, dtype, "p2p")

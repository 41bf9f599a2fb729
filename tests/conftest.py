import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ONE_GPU_REASON = ("this box has ONE GPU: the multi-GPU parity tests need >= 2 (run them with "
                  "`gpurun --gpus 2 -- python -m pytest tests -m gpu`); at N > 1 bench.py's `parity_sharded` "
                  "sub-record repeats the sharded-vs-unsharded-vs-oracle bit comparison in the driver's own run")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 CUDA devices on the box (always combined with gpu)")


def pytest_runtest_setup(item):
    """A `-m gpu` run on a box without a GPU fails loudly instead of skipping: the CUDA path is the
    product and there is no CPU fallback.  Multi-GPU tests never skip on a multi-GPU box; on a 1-GPU
    box they skip with the reason spelled out."""
    if item.get_closest_marker("gpu") is None:
        return
    import torch
    if not torch.cuda.is_available():
        pytest.fail("test is marked gpu but no CUDA device is visible: the D2Q9 path has no CPU fallback", pytrace=False)
    if item.get_closest_marker("multigpu") is not None and torch.cuda.device_count() < 2:
        pytest.skip(ONE_GPU_REASON)

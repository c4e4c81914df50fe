"""Multi-GPU parity (needs >= 2 GPUs on the box): launches tests/mgpu_worker.py with one process
per GPU and checks it reports MGPU_OK.  Skipped on 1-GPU boxes."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests._procs import run_torchrun  # noqa: E402


def _run(nproc, extra=()):
    # free rendezvous port, own process group: a timeout kills agent AND workers (no orphans
    # holding GPUs / ports)
    return run_torchrun(os.path.join(ROOT, "tests", "mgpu_worker.py"), nproc, dict(os.environ),
                        timeout=900, cwd=ROOT, python=sys.executable, args=extra)


@pytest.mark.gpu
@pytest.mark.multigpu
@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_reshard_kernels_match_oracle(nproc):
    if not torch.cuda.is_available() or torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    rc, out, err = _run(nproc)
    assert rc == 0 and f"MGPU_OK world={nproc}" in out, out[-3000:] + "\n" + err[-3000:]

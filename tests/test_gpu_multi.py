"""Multi-GPU parity (needs >= 2 GPUs on the box): launches tests/mgpu_worker.py with one process
per GPU and checks it reports MGPU_OK.  Skipped on 1-GPU boxes."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, port, extra=()):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mgpu_worker.py"), *extra]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)


@pytest.mark.gpu
@pytest.mark.multigpu
@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_reshard_kernels_match_oracle(nproc):
    if not torch.cuda.is_available() or torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    r = _run(nproc, 29530 + nproc)
    assert r.returncode == 0 and f"MGPU_OK world={nproc}" in r.stdout, \
        r.stdout[-3000:] + "\n" + r.stderr[-3000:]

"""Reference-in-the-loop parity (CPU container only: needs /root/reference).  The unmodified
reference traces, annotates and solves; its lowering (A) and the drop-in
easydist_b200.lowering.sharding_transform (B) are run on the same plan and inputs over gloo and
compared with each other and with vanilla PyTorch — tests/ref/auto_worker.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.refonly


@pytest.mark.parametrize("mesh,nproc,port,planner,model", [
    ("2", 2, 29791, "GREEDY", "foo"), ("2x2", 4, 29792, "GREEDY", "foo"),
    ("2x2", 4, 29793, "REPLICATE", "foo"), ("2x2", 4, 29794, "P2P", "foo"),
    ("2", 2, 29795, "GREEDY", "gpt"),   # the reference's GPT test model: views, expand, bmm
])
def test_dropin_lowering_equals_reference_lowering(mesh, nproc, port, planner, model):
    if not os.path.isdir("/root/reference/easydist"):
        pytest.skip("reference not present (GPU box)")
    env = dict(os.environ, EDB_TEST_MESH=mesh, OMP_NUM_THREADS="1", EDB_PLANNER=planner,
               EDB_MODEL=model)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "ref", "auto_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=280, cwd=ROOT, env=env)
    line = next((l for l in r.stdout.splitlines() if l.startswith("AUTO_PARITY")), "")
    assert r.returncode == 0 and "ok=True" in line, r.stdout[-2000:] + r.stderr[-3000:]
    # same communication structure as the reference's lowering
    ref_hist = line.split("hist_ref=")[1].split(" hist_b200=")[0]
    my_hist = line.split("hist_b200=")[1].rsplit(" [", 1)[0]
    assert ref_hist == my_hist, line

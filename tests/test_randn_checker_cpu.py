"""The randn reshard battery of tests/mgpu_worker.py (`randn`, seed 1234 +
rank; data movement bit-exact, fp32 reductions within 1e-5, bf16 within one ulp) run here against
the gloo stand-ins of the ten callables: checks the checker — shapes, oracle calls, tolerances — so
that a failure on GPUs means the kernels, not the test."""
import os

import torch
import torch.distributed as dist

from tests._procs import run_world


def _worker(rank, world, port, q):
    os.environ["OMP_NUM_THREADS"] = "1"
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                            world_size=world)
    import numpy as np
    from tests import gloo_ops, mgpu_worker as W
    gloo_ops.init_groups(np.arange(world).reshape((world,)))
    try:
        n = W.run_randn_cases(rank, world, list(range(world)), ops=gloo_ops, device="cpu")
        res = (True, "", n)
    except AssertionError as e:
        res = (False, str(e), 0)
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_randn_battery_logic_against_the_gloo_stand_ins():
    ok, msg, n = run_world(_worker, 2, lambda r, port, q: (r, 2, port, q), timeout=240)
    assert ok, msg
    assert n == 2 * (3 + 2 + 3 + 3)

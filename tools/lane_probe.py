"""EXPERIMENTAL probe (round-2 preparation): communication-lane collectives on N GPUs.
Launch: python -m torch.distributed.run --nproc-per-node 2 tools/lane_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from easydist_b200 import runtime
from tests import mgpu_worker as W

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rt = runtime.init(rank, world, local, heap_bytes=2 << 30)
rt.set_option("spin_timeout_ms", 3000)
n = W.run_lane(rank, world, list(range(world)))
torch.cuda.synchronize()
errs = rt.error_flags()
if rank == 0:
    print(f"LANE_PROBE world={world} checks={n} error_flags={errs}", flush=True)
dist.barrier()
dist.destroy_process_group()

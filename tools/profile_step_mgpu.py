"""Kernel-time breakdown of one eager compiled zero3 step on N GPUs (torch.profiler, rank 0 prints).
Launch: python -m torch.distributed.run --nproc-per-node N tools/profile_step_mgpu.py"""
import dataclasses, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from easydist_b200 import runtime
from easydist_b200.api import easydist_compile
from easydist_b200.device_mesh import set_device_mesh
from easydist_b200.workloads import GPT2, GPT2_CONFIGS, gpt2_train_step, synthetic_tokens

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rt = runtime.init(rank, world, local, heap_bytes=8 << 30)
set_device_mesh(list(range(world)), ["dp"], rank=rank)
cfg = dataclasses.replace(GPT2_CONFIGS["gpt2-medium"], attn="sdpa")
torch.manual_seed(0)
model = GPT2(cfg).to(device="cuda", dtype=torch.bfloat16)
opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, foreach=True)
tok, tgt = synthetic_tokens(cfg, 8, 512, rank, device="cuda")
step = easydist_compile(gpt2_train_step, parallel_mode="zero3", cuda_graph=True)
for _ in range(4):
    step(tok, tgt, model, opt)
torch.cuda.synchronize()
dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step(tok, tgt, model, opt)
    torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    ev = [e for e in prof.key_averages() if e.device_time_total > 0]
    tot = sum(e.self_device_time_total for e in ev)
    print(f"rank 0, world {world}: total device time {tot / 2e3:.2f} ms per step over "
          f"{sum(e.count for e in ev) // 2} launches")
    for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:28]:
        print(f"{e.self_device_time_total / 2e3:8.3f} ms {100 * e.self_device_time_total / tot:5.1f}% "
              f"n={e.count // 2:5d}  {e.key[:120]}")
dist.barrier()
dist.destroy_process_group()

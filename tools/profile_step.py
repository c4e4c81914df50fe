"""Kernel-time breakdown of one eager compiled step (torch.profiler), to rank what to optimise."""
import dataclasses, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from easydist_b200 import runtime
from easydist_b200.api import easydist_compile
from easydist_b200.device_mesh import set_device_mesh
from easydist_b200.workloads import GPT2, GPT2_CONFIGS, gpt2_train_step, synthetic_tokens

attn = sys.argv[1] if len(sys.argv) > 1 else "sdpa"
rt = runtime.init(0, 1, 0, heap_bytes=4 << 30)
set_device_mesh([0], ["dp"], rank=0)
cfg = dataclasses.replace(GPT2_CONFIGS["gpt2-medium"], attn=attn)
torch.manual_seed(0)
model = GPT2(cfg).to(device="cuda", dtype=torch.bfloat16)
opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, foreach=True)
tok, tgt = synthetic_tokens(cfg, 8, 512, 0, device="cuda")
step = easydist_compile(gpt2_train_step, parallel_mode="zero3", cuda_graph=False)
for _ in range(2):
    step(tok, tgt, model, opt)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(tok, tgt, model, opt)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.self_device_time_total for e in ev)
print(f"total device time {tot/1e3:.2f} ms over {sum(e.count for e in ev)} launches")
for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:40]:
    print(f"{e.self_device_time_total/1e3:8.3f} ms {100*e.self_device_time_total/tot:5.1f}% n={e.count:5d}  {e.key[:110]}")

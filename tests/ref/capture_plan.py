"""Capture-only recording (CPU container, needs /root/reference): the unmodified reference traces,
annotates and solves the benchmark GPT (EDB_GPT=depth,dim,heads,batch,seq) on a 1-D mesh of WORLD_SIZE
gloo ranks; rank 0 writes graph + plan as a bundle (EDB_RECORD=...json.gz); nothing is executed.
Run under `ulimit -s unlimited` (the front end recurses deeply at 24 layers), e.g.
  EDB_GPT=24,1024,16,8,128 EDB_RECORD=/tmp/m8.json.gz python -m torch.distributed.run \\
      --nproc-per-node 8 --master-addr 127.0.0.1 tests/ref/capture_plan.py
Used for tests/golden/auto_gpt2medium_s128_mesh8.json.gz (568 s)."""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, "/root/reference")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.set_num_threads(int(os.environ.get("EDB_THREADS_RANK0" if rank == 0 else "EDB_THREADS", "1")))
dist.init_process_group("gloo")
from oracle import refcompat
refcompat.install()
from easydist import easydist_setup
from easydist.torch.device_mesh import set_device_mesh
import easydist.torch.compile_auto as ref_auto
from torch.distributed.device_mesh import DeviceMesh
easydist_setup(backend="torch", device="cpu", allow_tf32=False)
set_device_mesh(DeviceMesh("cpu", torch.arange(world), mesh_dim_names=["spmd0"]))
from benchmark.torch.model.gpt import GPT
from easydist_b200 import graph_io
depth, dim, heads, gb, gs = (int(v) for v in os.environ.get("EDB_GPT", "24,1024,16,8,128").split(","))
torch.manual_seed(42)
with torch.device("meta" if os.environ.get("EDB_META") == "1" else "cpu"):
    model = GPT(depth=depth, dim=dim, num_heads=heads)
opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, foreach=True)
x = torch.randn(gb, gs, dim)
def train_step(input, model, opt):
    out = model(input); loss = out.mean(); loss.backward(); opt.step(); opt.zero_grad(); return out
class Done(Exception): pass
def capture(fx_module, opt_strategy, state_io_map):
    if rank == 0:
        text = graph_io.dump_bundle(fx_module, opt_strategy, [(a.name, b.name) for a, b in state_io_map.items()],
                                    extra={"mesh": [world], "model": f"GPT(depth={depth}, dim={dim}, num_heads={heads})",
                                           "seed": 42, "batch": [gb, gs, dim], "planner": "GREEDY"})
        import gzip
        with gzip.open(os.environ["EDB_RECORD"], "wt") as f:
            f.write(text)
        print(f"CAPTURED nodes={len(list(fx_module.graph.nodes))} plan={len(opt_strategy)} bytes={len(text)} t={time.time()-t0:.0f}s", flush=True)
    raise Done()
ref_auto.sharding_transform = capture
t0 = time.time()
try:
    ref_auto._compile_auto(train_step, "fake", None, "capture", (x, model, opt), {})
except Done:
    pass
dist.barrier()
dist.destroy_process_group()

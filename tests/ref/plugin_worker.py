"""Hook A with the reference in the loop (CPU container only): the
reference's `easydist_compile(parallel_mode="b200_<mode>")` decorator — its own
CompiledFuncWrapper, input-signature registry and run loop — drives this backend's compiled
object, registered through `easydist_b200.api.register()`.  World 2 over gloo; compared with
vanilla full-batch training (the reference's comparator, rtol 1e-4 / atol 1e-5)."""
import copy
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


class Foo(torch.nn.Module):
    def __init__(self, d=32):
        super().__init__()
        self.norm = torch.nn.LayerNorm(d)
        self.linear = torch.nn.Linear(d, d)

    def forward(self, x):
        return self.linear(self.norm(x)).relu()


def train_step(input, model, opt):
    out = model(input)
    loss = out.mean()
    loss.backward()
    opt.step()
    opt.zero_grad()
    return loss


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    mode = os.environ.get("EDB_PLUGIN_MODE", "b200_ddp")
    torch.set_num_threads(1)
    dist.init_process_group("gloo")
    from oracle import refcompat
    refcompat.install()
    from easydist import easydist_setup
    from easydist.torch.api import easydist_compile
    from easydist.torch.device_mesh import set_device_mesh
    from torch.distributed.device_mesh import DeviceMesh
    easydist_setup(backend="torch", device="cpu", allow_tf32=False)
    # the reference's set_device_mesh insists on an "spmd*" dim of size > 1 (device_mesh.py:36-41,
    # 131-139; its own examples/torch/simple_ddp.py mesh ["dp", "placeholder"] no longer passes
    # that check); this backend's DP modes take the only dim of a 1-D mesh
    set_device_mesh(DeviceMesh("cpu", torch.arange(world), mesh_dim_names=["spmd0"]))

    from easydist_b200 import api
    from tests import gloo_ops
    extra = {}
    if os.environ.get("EDB_TEST_AUTO_PF") == "1":
        # Hook C with the parameter-prefetch rewrite of the product path too (stand-in runtime)
        extra["fuse_rt"] = gloo_ops.FakeSymmRuntime()
    api.register(ops=gloo_ops, native=False, **extra)
    if mode in ("auto", "b200_auto"):
        # Hook B: register() also rebinds compile_auto.sharding_transform; the reference's auto
        # path (annotation, solver, executor) then runs on this backend's lowering
        import numpy as np
        gloo_ops.init_groups(np.arange(world))

    torch.manual_seed(42)
    g = torch.Generator().manual_seed(7)
    if os.environ.get("EDB_PLUGIN_MODEL", "foo") == "gpt":
        # the reference's own test model (TEST_GPT, tests/test_torch/test_utils.py:55-68): attention
        # with views / expand / bmm, GeLU MLPs, LayerNorms
        from benchmark.torch.model.gpt import GPT
        model0 = GPT(depth=2, dim=64, num_heads=4)
        batches = [torch.randn(world * 2, 32, 64, generator=g) for _ in range(3)]
    else:
        model0 = Foo()
        batches = [torch.randn(world * 4, 32, generator=g) for _ in range(3)]
    model = copy.deepcopy(model0)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, foreach=True)
    step = easydist_compile(train_step, mode, "fake", cuda_graph=False)   # the REFERENCE's decorator
    vmodel = copy.deepcopy(model0)
    vopt = torch.optim.SGD(vmodel.parameters(), lr=0.1, momentum=0.9, foreach=True)
    ok, msgs = True, []
    sl = slice(rank * 4, (rank + 1) * 4)
    for b in batches:
        if mode in ("auto", "b200_auto"):   # SPMD: every rank passes the global batch, gets the global loss
            loss = step(b, model, opt).detach().clone()
            want = train_step(b, vmodel, vopt).detach()
        else:
            loss = step(b[sl], model, opt).detach().clone()
            want = train_step(b, vmodel, vopt).detach()
            dist.all_reduce(loss)
            loss /= world
        if not torch.allclose(loss, want, rtol=1e-4, atol=1e-5):
            ok = False
            msgs.append(f"loss {loss} vs {want}")
    cf = step.compiled_func
    for name, p_ref in vmodel.named_parameters():
        p = cf.named_parameters()[name]
        p = p.to_local() if hasattr(p, "to_local") else p
        if mode in ("auto", "b200_auto") and p.shape != p_ref.shape:
            continue  # placement is the solver's choice; outputs / losses are the comparator here
        if p.shape != p_ref.shape:
            parts = [torch.empty_like(p) for _ in range(world)]
            dist.all_gather(parts, p.contiguous())
            p = torch.cat(parts).view(p_ref.shape)
        if not torch.allclose(p, p_ref.detach(), rtol=1e-4, atol=1e-5):
            ok = False
            msgs.append(f"param {name} differs by {(p - p_ref).abs().max()}")
    source = None
    if mode == "b200_auto" and os.environ.get("EDB_PLAN_CACHE_DIR"):
        # second compilation of the same step in this job: the plan must come from the cache (no
        # tracing / annotation / ILP by the reference) and train identically
        source = [api.LAST_AUTO_SOURCE[0]]
        import easydist.torch.compile_auto as ref_auto
        calls = {"n": 0}
        orig = ref_auto._compile_auto

        def counting(*a, **k):
            calls["n"] += 1
            return orig(*a, **k)

        ref_auto._compile_auto = counting
        model2 = copy.deepcopy(model0)
        opt2 = torch.optim.SGD(model2.parameters(), lr=0.1, momentum=0.9, foreach=True)
        step2 = easydist_compile(train_step, mode, "fake", cuda_graph=False)
        vmodel2 = copy.deepcopy(model0)
        vopt2 = torch.optim.SGD(vmodel2.parameters(), lr=0.1, momentum=0.9, foreach=True)
        for b in batches:
            loss = step2(b, model2, opt2).detach().clone()
            want = train_step(b, vmodel2, vopt2).detach()
            if not torch.allclose(loss, want, rtol=1e-4, atol=1e-5):
                ok = False
                msgs.append(f"cached plan: loss {loss} vs {want}")
        source.append(api.LAST_AUTO_SOURCE[0])
        if calls["n"] != 0 or source != ["solved", "cache"]:
            ok = False
            msgs.append(f"plan cache not used: reference compiles={calls['n']} sources={source}")
    if rank == 0:
        print(f"PLUGIN_PARITY ok={ok} mode={mode} wrapper={type(step).__module__}.{type(step).__name__} "
              f"compiled={type(cf).__module__}.{type(cf).__name__} plan_source={source} "
              f"fused={getattr(cf, 'info', {}).get('fused')} {msgs}",
              flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""BASELINE.json config 4 (Llama-2) at CPU scale: RMSNorm / rotary / SwiGLU / untied-head graph
through the DP modes on gloo, world 2, vs vanilla full-batch training (the reference's comparator:
rtol 1e-4 / atol 1e-5, tests/test_torch/test_spmd.py:67); also checks that the cross-entropy and
SGD rewrites of the native path match on this graph."""
import os

import pytest
import torch
import torch.distributed as dist

from tests._procs import run_world


def _worker(rank, world, port, mode, q):
    os.environ["OMP_NUM_THREADS"] = "1"
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from easydist_b200 import api, lowering, workloads
    from easydist_b200.device_mesh import set_device_mesh
    from tests import gloo_ops
    set_device_mesh(list(range(world)), ["dp"], rank=rank)
    cfg = workloads.LLAMA_CONFIGS["llama-tiny"]
    torch.manual_seed(0)
    model, ref = workloads.Llama(cfg), workloads.Llama(cfg)
    ref.load_state_dict(model.state_dict())
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9, foreach=True)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, foreach=True)
    g = torch.Generator().manual_seed(5)
    toks = [torch.randint(0, cfg.vocab_size, (world * 2, 33), generator=g) for _ in range(3)]
    sl = slice(rank * 2, (rank + 1) * 2)
    compiled = api._compile_dp(workloads.gpt2_train_step, mode, "fake",
                               (toks[0][sl, :-1].contiguous(), toks[0][sl, 1:].contiguous(), model, opt), {},
                               ops=gloo_ops,
                               native=False)
    n_ce = lowering.fuse_cross_entropy(compiled.graph)
    n_opt = lowering.fuse_optimizer_updates(compiled.graph)
    ok, msg = (n_ce == 1 and n_opt == 1), f"rewrites: ce {n_ce} opt {n_opt}"
    for t in toks:
        loss = compiled(t[sl, :-1].contiguous(), t[sl, 1:].contiguous(), model, opt)
        rloss = workloads.gpt2_train_step(t[:, :-1].contiguous(), t[:, 1:].contiguous(), ref, ropt)
        la = loss.detach().clone()
        dist.all_reduce(la)
        la /= world
        if not torch.allclose(la, rloss.detach(), rtol=1e-4, atol=1e-5):
            ok, msg = False, f"loss {la} vs {rloss}"
    params = compiled.named_parameters()
    for name, p_ref in ref.named_parameters():
        p = params[name]
        if p.shape != p_ref.shape:
            parts = [torch.empty_like(p) for _ in range(world)]
            dist.all_gather(parts, p.contiguous())
            p = torch.cat(parts).view(p_ref.shape)
        if not torch.allclose(p, p_ref.detach(), rtol=1e-4, atol=1e-5):
            ok, msg = False, f"param {name} differs by {(p - p_ref).abs().max()}"
    if rank == 0:
        q.put((ok, msg))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["ddp", "zero3"])
def test_tiny_llama_dp_matches_vanilla(mode):
    ok, msg = run_world(_worker, 2, lambda r, port, q: (r, 2, port, mode, q), timeout=300)
    assert ok, msg

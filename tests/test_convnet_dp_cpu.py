"""BASELINE.json config 3 (ResNet-50) at CPU scale: a residual conv + BatchNorm network through
the DP modes of the executor on gloo, world 2.  What this pins beyond the MLP / GPT cases: module
*buffers* (BatchNorm running statistics, num_batches_tracked) are traced as graph state and written
back every step, 4-D parameters are flattened / sharded by zero2 / zero3, and the comparison is
against per-rank eager training with averaged gradients (BatchNorm statistics are per rank in the
reference's DP modes too: compile_dp.py never synchronises them; SURVEY.md §7 BatchNorm caveat).
Tolerance: the reference's comparator, rtol 1e-4 / atol 1e-5 (tests/test_torch/test_spmd.py:67)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from tests._procs import run_world


class Block(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.c1 = nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.b1 = nn.BatchNorm2d(c)
        self.c2 = nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.b2 = nn.BatchNorm2d(c)

    def forward(self, x):
        return F.relu(x + self.b2(self.c2(F.relu(self.b1(self.c1(x))))))


class TinyResNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.stem = nn.Conv2d(3, 8, 3, padding=1, bias=False)
        self.bn = nn.BatchNorm2d(8)
        self.blk = Block(8)
        self.fc = nn.Linear(8, 10)

    def forward(self, x):
        x = self.blk(F.relu(self.bn(self.stem(x))))
        return self.fc(F.adaptive_avg_pool2d(x, 1).flatten(1))


def train_step(x, y, model, opt):
    loss = F.cross_entropy(model(x), y)
    loss.backward()
    opt.step()
    opt.zero_grad(True)
    return loss


def _worker(rank, world, port, mode, q, arch="tiny"):
    os.environ["OMP_NUM_THREADS"] = "1"
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from easydist_b200 import api
    from easydist_b200.device_mesh import set_device_mesh
    from tests import gloo_ops
    set_device_mesh(list(range(world)), ["dp"], rank=rank)
    torch.manual_seed(0)
    if arch == "resnet50":
        # the real thing (torchvision's ResNet-50: 53 convolutions, 53 BatchNorms, 161 parameters,
        # 159 buffers) at CPU-sized inputs
        import torchvision
        model, ref = torchvision.models.resnet50(num_classes=10), torchvision.models.resnet50(num_classes=10)
        res, lr, steps = 32, 0.01, 2
    else:
        model, ref = TinyResNet(), TinyResNet()
        res, lr, steps = 16, 0.1, 3
    ref.load_state_dict(model.state_dict())
    opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=0.9, foreach=True)
    ropt = torch.optim.SGD(ref.parameters(), lr=lr, momentum=0.9, foreach=True)
    g = torch.Generator().manual_seed(3)
    xs = [torch.randn(world * 4, 3, res, res, generator=g) for _ in range(steps)]
    ys = [torch.randint(0, 10, (world * 4,), generator=g) for _ in range(steps)]
    sl = slice(rank * 4, (rank + 1) * 4)
    compiled = api._compile_dp(train_step, mode, "fake", (xs[0][sl], ys[0][sl], model, opt), {},
                               ops=gloo_ops, native=False)
    ok, msg = True, ""
    for x, y in zip(xs, ys):
        loss = compiled(x[sl], y[sl], model, opt)
        # per-rank eager step with averaged gradients (= what DDP computes)
        rloss = F.cross_entropy(ref(x[sl]), y[sl])
        rloss.backward()
        for p in ref.parameters():
            dist.all_reduce(p.grad)
            p.grad /= world
        ropt.step()
        ropt.zero_grad(True)
        if not torch.allclose(loss, rloss.detach(), rtol=1e-4, atol=1e-5):
            ok, msg = False, f"loss {loss} vs {rloss}"
    params, bufs = compiled.named_parameters(), compiled.named_buffers()
    for name, p_ref in ref.named_parameters():
        p = params[name]
        if p.shape != p_ref.shape:  # flat 1/n shard (zero3)
            parts = [torch.empty_like(p) for _ in range(world)]
            dist.all_gather(parts, p.contiguous())
            p = torch.cat(parts).view(p_ref.shape)
        if not torch.allclose(p, p_ref.detach(), rtol=1e-4, atol=1e-5):
            ok, msg = False, f"param {name} differs by {(p - p_ref).abs().max()}"
    for name, b_ref in ref.named_buffers():
        if not torch.allclose(bufs[name].float(), b_ref.float(), rtol=1e-4, atol=1e-5):
            ok, msg = False, f"buffer {name} differs"
    if rank == 0:
        q.put((ok, msg, compiled.info["comm_nodes"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["ddp", "zero2", "zero3"])
def test_conv_batchnorm_network_matches_per_rank_eager(mode):
    ok, msg, hist = run_world(_worker, 2, lambda r, port, q: (r, 2, port, mode, q), timeout=240)
    assert ok, msg
    if mode == "ddp":
        # one all-reduce per parameter tensor: stem, bn (2), 2 x (conv + bn (2)), fc (2)
        assert hist.get("all_reduce_start", 0) == 11, hist
    else:
        assert hist.get("reduce_scatter_start", 0) + hist.get("all_reduce_start", 0) >= 1, hist


def test_torchvision_resnet50_ddp_matches_per_rank_eager():
    """BASELINE.json config 3's model itself (torchvision ResNet-50) through the ddp mode on gloo,
    world 2, 32x32 inputs: traced with its 159 BatchNorm buffers as graph state, 161 gradient
    all-reduces, losses / every parameter / every buffer equal per-rank eager training with averaged
    gradients (rtol 1e-4, atol 1e-5).  On GPUs its convolutions and BatchNorms stay ATen/cuDNN kernels
    (no native convolution path: DESIGN.md section 7); what runs natively there is the collectives."""
    pytest.importorskip("torchvision")
    ok, msg, hist = run_world(_worker, 2, lambda r, port, q: (r, 2, port, "ddp", q, "resnet50"),
                              timeout=500)
    assert ok, msg
    assert hist.get("all_reduce_start", 0) == 161, hist

"""easydist_compile wrapper behaviour (reference: easydist/torch/api.py:53-224): one compilation,
graphs keyed by input signature, the reference's "Input mismatch" error, enable_mono_graph
re-lowering for new input shapes over the same state, compile_only."""
import pytest
import torch

from easydist_b200 import api
from easydist_b200.device_mesh import set_device_mesh
from tests import gloo_ops


class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.norm = torch.nn.LayerNorm(16)
        self.fc = torch.nn.Linear(16, 4)

    def forward(self, x):
        return self.fc(self.norm(x))


def train_step(x, y, model, opt):
    loss = torch.nn.functional.cross_entropy(model(x), y)
    loss.backward()
    opt.step()
    opt.zero_grad(True)
    return loss


def _setup():
    set_device_mesh([0], ["dp"], rank=0)
    torch.manual_seed(0)
    model, ref = Net(), Net()
    ref.load_state_dict(model.state_dict())
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, foreach=True)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9, foreach=True)
    g = torch.Generator().manual_seed(1)
    data = [(torch.randn(b, 16, generator=g), torch.randint(0, 4, (b,), generator=g)) for b in (8, 8, 12, 8, 12)]
    return model, ref, opt, ropt, data


def test_input_signature_depends_on_shapes_dtypes_and_objects():
    m, _, opt, _, _ = _setup()
    a = api.input_signature((torch.zeros(2, 3), m, opt), {})
    assert a == api.input_signature((torch.ones(2, 3), m, opt), {})          # values do not matter
    assert a != api.input_signature((torch.zeros(3, 3), m, opt), {})         # shapes do
    assert a != api.input_signature((torch.zeros(2, 3, dtype=torch.bfloat16), m, opt), {})
    assert a != api.input_signature((torch.zeros(2, 3), Net(), opt), {})     # another module
    assert a != api.input_signature((torch.zeros(2, 3), m), {"opt": opt})    # call structure


def test_new_input_shape_raises_the_reference_error_without_mono_graph():
    model, ref, opt, ropt, data = _setup()
    step = api.easydist_compile(train_step, parallel_mode="ddp", cuda_graph=False,
                                ops=gloo_ops, native=False)
    for x, y in data[:2]:
        assert torch.allclose(step(x, y, model, opt), train_step(x, y, ref, ropt).detach(), rtol=1e-5)
    with pytest.raises(RuntimeError, match="Input mismatch"):
        step(*data[2], model, opt)
    assert len(step.all_input_signature) == 2 and len(step.graph_list) == 1


def test_mono_graph_relowers_for_new_shapes_over_the_same_state():
    model, ref, opt, ropt, data = _setup()
    step = api.easydist_compile(train_step, parallel_mode="ddp", cuda_graph=False,
                                enable_mono_graph=True, ops=gloo_ops, native=False)
    for x, y in data:                                      # batch 8, 8, 12, 8, 12
        got = step(x, y, model, opt)
        want = train_step(x, y, ref, ropt)
        assert torch.allclose(got, want.detach(), rtol=1e-5, atol=1e-6)
    assert len(step.graph_list) == 2                       # one lowered graph per signature
    params = step.compiled_func.named_parameters()
    for name, p in ref.named_parameters():
        assert torch.allclose(params[name], p.detach(), rtol=1e-4, atol=1e-6), name


def test_compile_only_returns_the_compiled_object():
    model, _, opt, _, data = _setup()
    step = api.easydist_compile(train_step, parallel_mode="ddp", cuda_graph=False, compile_only=True,
                                ops=gloo_ops, native=False)
    compiled = step(*data[0], model, opt)
    assert hasattr(compiled, "run_with_graph") and hasattr(compiled, "graph")
    assert set(compiled.named_parameters()) == {n for n, _ in model.named_parameters()}


def test_unknown_parallel_mode_is_rejected_like_the_reference():
    with pytest.raises(NotImplementedError):
        api.easydist_compile(train_step, parallel_mode="nope")


@pytest.mark.parametrize("make", [
    lambda ps: torch.optim.SGD(ps, lr=0.1, momentum=0.9, weight_decay=0.1, foreach=True),
    lambda ps: torch.optim.AdamW(ps, lr=1e-2, weight_decay=0.1, fused=True),
])
def test_compiling_does_not_change_parameters_or_seed_optimizer_state(make):
    """The warm-up step that materialises optimizer state must leave the model as it was, also with
    weight decay (parameters would shrink; SGD's momentum buffers would start at wd * p)."""
    from easydist_b200.compile import warm_up_optimizer
    set_device_mesh([0], ["dp"], rank=0)
    torch.manual_seed(0)
    model = Net()
    before = {k: v.detach().clone() for k, v in model.named_parameters()}
    opt = make(model.parameters())
    states = warm_up_optimizer(model, opt)
    for k, v in model.named_parameters():
        assert torch.equal(v, before[k]), k
    assert opt.param_groups[0]["weight_decay"] == 0.1
    for name, st in states.items():
        for key, t in st.items():
            if isinstance(t, torch.Tensor) and key != "step":
                assert float(t.abs().max()) == 0.0, (name, key)
            if key == "step":
                assert float(t) == 0.0

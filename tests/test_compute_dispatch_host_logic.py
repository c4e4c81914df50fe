"""Host side of the compute dispatch (norm.py, loss.py, optim.py) with the C-ABI mocked out: which
shapes / dtypes / layouts reach which entry point with which arguments, and what goes to ATen
(counted) instead — no GPU, no compute."""
import pytest
import torch

from easydist_b200 import _lib, loss, norm, optim


class _FakeLib:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def fn(*args):
            self.calls.append((name, args))
            if name.endswith("_workspace"):
                args[-1]._obj.value = 1024          # byref(c_size_t)
            return 0
        return fn


@pytest.fixture
def lib(monkeypatch):
    fake = _FakeLib()
    monkeypatch.setattr(_lib, "load", lambda *a, **k: fake)
    for mod in (norm, loss):
        monkeypatch.setattr(mod, "_stream", lambda t: None)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: type("S", (), {"cuda_stream": 0})())
    for m in (norm, loss, optim):
        m.reset_stats()
    return fake


def test_layer_norm_shapes_the_kernel_takes(lib):
    x = torch.zeros(6, 7, 1024, dtype=torch.bfloat16)
    w = torch.ones(1024, dtype=torch.bfloat16)
    y, mean, rstd = norm.native_layer_norm(x, [1024], w, w, 1e-5)
    (name, a), = lib.calls
    assert name == "edb_layer_norm_fwd" and a[6:8] == (42, 1024)       # rows, H
    assert a[9] == _lib.DTYPE_CODES["bfloat16"] and abs(a[8] - 1e-5) < 1e-12
    assert y.shape == x.shape and mean.shape == (6, 7, 1) and mean.dtype == torch.float32
    # H = 100 is outside the kernel's contract: ATen, counted
    norm.native_layer_norm(torch.zeros(4, 100), [100], torch.ones(100), None, 1e-5)
    assert norm.stats()["edb_ln_fwd"] == 1 and norm.stats()["aten_ln"] == 1


def test_column_sum_dispatch_conditions(lib):
    x = torch.zeros(4096, 1024, dtype=torch.bfloat16)
    out = norm.sum_dim_intlist(x, [0], True)
    name, a = lib.calls[-1]
    assert name == "edb_colsum" and a[3:6] == (4096, 1024, 1024) and out.shape == (1, 1024)
    lib.calls.clear()
    for bad in (torch.zeros(8, 1024, dtype=torch.bfloat16),          # too few rows
                torch.zeros(4096, 1030, dtype=torch.bfloat16),       # cols % 8
                torch.zeros(2, 64, 64)):                             # not 2-D
        norm.sum_dim_intlist(bad, [0], True)
    assert not lib.calls and norm.stats()["aten_sum"] == 3
    assert norm.sum_dim_intlist(torch.ones(4, 4), [1], False).tolist() == [4.0] * 4


def test_cross_entropy_passes_strides_and_builds_a_tma_legal_gradient(lib):
    rows, vocab = 8, 50257
    logits = torch.zeros(rows, 50264, dtype=torch.bfloat16)[:, :vocab]     # padded GEMM output
    target = torch.zeros(rows, dtype=torch.int64)
    l, tw, lse = loss.cross_entropy_fwd(logits, target, -100, 1)
    name, a = lib.calls[-1]
    assert name == "edb_cross_entropy_fwd"
    assert a[5] == 50264 and a[7:12] == (rows, vocab, -100, 1, _lib.DTYPE_CODES["bfloat16"])
    assert l.shape == () and tw.shape == () and lse.shape == (rows,) and lse.dtype == torch.float32
    dx = loss.cross_entropy_bwd(torch.ones(()), logits, target, lse, tw, -100, 1)
    name, a = lib.calls[-1]
    assert name == "edb_cross_entropy_bwd" and a[1] == 50264 and a[3] == 50264
    assert dx.shape == (rows, vocab) and dx.stride() == (50264, 1) and dx.dtype == torch.bfloat16
    # 3-D logits or int32 targets are not this kernel's case
    lib.calls.clear()
    loss.cross_entropy_fwd(torch.zeros(4, 10, dtype=torch.float64), torch.zeros(4, dtype=torch.int64), -100, 1)
    assert not lib.calls and loss.stats()["aten_ce"] == 1


def test_sgd_splits_the_lists_between_the_kernel_and_aten(lib):
    def trio(n, dtype=torch.float32):
        return torch.zeros(n, dtype=dtype), torch.ones(n, dtype=dtype), torch.zeros(n, dtype=dtype)

    a, b = trio(64), trio(32, torch.bfloat16)
    base = torch.zeros(40)
    c = (base[1:33], torch.ones(32), torch.zeros(32))                    # not 16-byte aligned
    params, grads, bufs = zip(a, b, c)
    optim.sgd_momentum_(list(params), list(grads), list(bufs), 0.9, 1, -0.1)
    kinds = sorted((name, args[0], args[8]) for name, args in lib.calls)
    assert kinds == sorted([("edb_sgd_momentum", 1, _lib.DTYPE_CODES["bfloat16"]),
                            ("edb_sgd_momentum", 1, _lib.DTYPE_CODES["float32"])])
    assert optim.stats() == {"edb_sgd": 2, "aten_sgd": 1}
    assert torch.allclose(c[0], torch.full((32,), -0.1))                 # the ATen group really ran
    with pytest.raises(ValueError):
        optim.sgd_momentum_([a[0]], [], [a[2]], 0.9, 1, -0.1)

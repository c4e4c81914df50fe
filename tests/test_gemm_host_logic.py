"""Host side of easydist_b200.gemm with the C-ABI mocked out: operand-layout classification
(K-major / MN-major for both operands), leading dimensions, padded staging for strides that break
TMA's 16-byte rule (vocab 50257), padded outputs — no GPU, no compute."""
import pytest
import torch

from easydist_b200 import _lib, gemm


class _FakeLib:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def fn(*args):
            self.calls.append((name, args))
            return 0
        return fn


@pytest.fixture
def lib(monkeypatch):
    fake = _FakeLib()
    monkeypatch.setattr(_lib, "load", lambda *a, **k: fake)
    monkeypatch.setattr(gemm, "_stream", lambda t: None)
    gemm.reset_stats()
    return fake


def _gemm_args(call):
    name, a = call
    assert name == "edb_gemm_bf16"
    keys = ["C", "A", "B", "bias", "M", "N", "K", "lda", "ldb", "ldc", "a_k", "b_k", "acc", "stream"]
    return dict(zip(keys, a))


@pytest.mark.parametrize("a_k,b_k", [(True, True), (True, False), (False, True), (False, False)])
def test_operand_layouts_map_to_kmajor_flags_without_copies(lib, a_k, b_k):
    M, N, K = 64, 48, 32
    A = torch.zeros(M, K, dtype=torch.bfloat16)
    B = torch.zeros(K, N, dtype=torch.bfloat16)
    a = A if a_k else A.t().contiguous().t()          # [M,K] with stride (1, M)
    b = B.t().contiguous().t() if b_k else B          # [K,N] with stride (1, K) = stored [N,K]
    out = gemm._launch(a, b, None)
    (call,) = lib.calls                               # no staging copy
    g = _gemm_args(call)
    assert (g["M"], g["N"], g["K"]) == (M, N, K)
    assert (g["a_k"], g["b_k"]) == (int(a_k), int(b_k))
    assert g["lda"] == (K if a_k else M) and g["ldb"] == (K if b_k else N) and g["ldc"] == N
    assert g["A"] == a.data_ptr() and g["B"] == b.data_ptr() and g["bias"] is None
    assert out.shape == (M, N) and out.is_contiguous()
    assert gemm.recorded_calls()[-1][:5] == (M, N, K, a_k, b_k)


def test_unaligned_leading_dims_are_staged_and_unaligned_n_gets_a_padded_output(lib):
    V, H, T = 50257, 64, 16
    x = torch.zeros(T, H, dtype=torch.bfloat16)
    W = torch.zeros(V, H, dtype=torch.bfloat16)
    logits = gemm._launch(x, W.t(), None)             # N = 50257: padded output, operands legal
    g = _gemm_args(lib.calls[-1])
    assert g["ldc"] == 50264 and logits.shape == (T, V) and logits.stride() == (50264, 1)
    assert [c[0] for c in lib.calls] == ["edb_gemm_bf16"]
    lib.calls.clear()
    dl = torch.zeros(T, V, dtype=torch.bfloat16)      # row stride 50257: not a multiple of 8
    gemm._launch(dl, W, None)                         # dgrad: A staged into a padded buffer
    names = [c[0] for c in lib.calls]
    assert names == ["edb_box_copy_local", "edb_gemm_bf16"]
    g = _gemm_args(lib.calls[-1])
    assert g["lda"] == 50264 and g["a_k"] == 1 and g["K"] == V
    assert gemm.stats()["padded_operands"] == 1
    lib.calls.clear()
    # the cross-entropy kernel hands over a gradient that already has a legal stride: no staging
    dl_padded = torch.zeros(T, 50264, dtype=torch.bfloat16)[:, :V]
    gemm._launch(dl_padded, W, None)
    gemm._launch(dl_padded.t(), x, None)              # wgrad: A = dl^T, MN-major view of the same buffer
    assert [c[0] for c in lib.calls] == ["edb_gemm_bf16", "edb_gemm_bf16"]
    g = _gemm_args(lib.calls[-1])
    assert (g["M"], g["K"], g["a_k"], g["lda"]) == (V, T, 0, 50264)


def test_shapes_the_kernel_does_not_take(lib):
    a = torch.zeros(8, 1, dtype=torch.bfloat16)
    assert gemm._launch(a, torch.zeros(1, 8, dtype=torch.bfloat16), None) is None   # degenerate strides
    s = torch.zeros(8, 16, dtype=torch.bfloat16)[:, ::2]                            # no unit stride
    assert gemm._launch(s, torch.zeros(8, 8, dtype=torch.bfloat16), None) is None
    # bias needs N % 8 == 0
    assert gemm._launch(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 12, dtype=torch.bfloat16),
                        torch.zeros(12, dtype=torch.bfloat16)) is None
    # CPU tensors never reach the kernel through the public entry point: ATen, counted
    out = gemm.mm(torch.ones(4, 4, dtype=torch.bfloat16), torch.ones(4, 4, dtype=torch.bfloat16))
    assert gemm.stats()["aten_mm"] == 1 and float(out[0, 0]) == 4.0

"""Sharded-op kernel dispatch for dense contractions.

The sharded FX graph the reference executes calls ATen for every compute node
(easydist/torch/compile_auto.py:752-756 runs the GraphModule op by op; after
passes/fix_bias.py the Linear layers are `aten.mm` + `aten.add`).  Here bf16 `aten.mm` nodes are
dispatched to the hand-written tcgen05 GEMM of libedb.so; shapes the kernel does not cover
(unaligned leading dimensions such as the 50257-wide LM head, non-bf16 dtypes) are routed to
ATen/cuBLAS and counted, so the share of native GEMMs is visible in `stats()`.
"""
import torch
from torch._subclasses.fake_tensor import FakeTensor

from . import _lib
from ._lib import check
from .runtime import get_runtime

_stats = {"edb_gemm": 0, "aten_mm": 0, "unsupported": {}}
_calls = []  # (M, N, K, a_kmajor, b_kmajor) of the native launches since reset_stats()


def stats():
    return {"edb_gemm": _stats["edb_gemm"], "aten_mm": _stats["aten_mm"],
            "unsupported": dict(_stats["unsupported"])}


def reset_stats():
    _stats["edb_gemm"] = 0
    _stats["aten_mm"] = 0
    _stats["unsupported"] = {}
    del _calls[:]


def recorded_calls():
    return list(_calls)


def _operand_layout(t, inner_is_dim1):
    """(kmajor, ld) of a 2-D operand view or None.  For A=[M,K]: K-major iff stride(1)==1.
    For B=[K,N]: 'K-major' means stored [N,K] row-major, i.e. stride(0)==1."""
    s0, s1 = t.stride()
    if t.shape[0] == 1 or t.shape[1] == 1:
        return None  # degenerate strides: leave to ATen
    if inner_is_dim1:
        if s1 == 1:
            return True, s0
        if s0 == 1:
            return False, s1
    else:
        if s0 == 1:
            return True, s1
        if s1 == 1:
            return False, s0
    return None


def gemm_supported(a, b):
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16 or a.dim() != 2 or b.dim() != 2:
        return None
    la = _operand_layout(a, True)
    lb = _operand_layout(b, False)
    if la is None or lb is None:
        return None
    M, K = a.shape
    N = b.shape[1]
    if (la[1] % 8) or (lb[1] % 8) or (N % 8):
        return None
    if (la[0] and K % 8) or (not la[0] and M % 8):
        return None
    if (a.data_ptr() % 16) or (b.data_ptr() % 16):
        return None
    return la, lb


def mm(a, b):
    """aten.mm.default(a, b) with bf16 operands on the tcgen05 kernel."""
    if isinstance(a, FakeTensor) or isinstance(b, FakeTensor) or a.is_meta:
        return torch.ops.aten.mm.default(a, b)
    lay = gemm_supported(a, b) if a.is_cuda else None
    if lay is None:
        key = (tuple(a.shape), tuple(a.stride()), tuple(b.shape), tuple(b.stride()), str(a.dtype))
        _stats["unsupported"][key] = _stats["unsupported"].get(key, 0) + 1
        _stats["aten_mm"] += 1
        return torch.ops.aten.mm.default(a, b)
    (a_k, lda), (b_k, ldb) = lay
    M, K = a.shape
    N = b.shape[1]
    out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    lib = _lib.load()
    stream = torch.cuda.current_stream(a.device).cuda_stream
    check(lib.edb_gemm_bf16(out.data_ptr(), a.data_ptr(), b.data_ptr(), M, N, K, lda, ldb, N,
                            1 if a_k else 0, 1 if b_k else 0, 0, stream))
    _stats["edb_gemm"] += 1
    if len(_calls) < 8192:
        _calls.append((M, N, K, bool(a_k), bool(b_k)))
    return out


def addmm(bias, a, b):
    """aten.addmm.default(bias, a, b) = bias + a @ b with the product on the tcgen05 kernel."""
    return torch.ops.aten.add.Tensor(mm(a, b), bias)

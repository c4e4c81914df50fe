"""Sharded-op kernel dispatch for LayerNorm: `aten.native_layer_norm` and
`aten.native_layer_norm_backward` nodes of the compiled graph run on the HBM-streaming kernels of
libedb.so (edb_norm.cu); shapes they do not cover go to ATen and are counted."""
from ctypes import byref, c_size_t

import torch
from torch._subclasses.fake_tensor import FakeTensor

from . import _lib
from ._lib import check

_stats = {"edb_ln_fwd": 0, "edb_ln_bwd": 0, "aten_ln": 0, "edb_colsum": 0, "aten_sum": 0}
_DT = {torch.bfloat16: _lib.DTYPE_CODES["bfloat16"], torch.float32: _lib.DTYPE_CODES["float32"]}
_workspaces = {}
aten = torch.ops.aten


def stats():
    return dict(_stats)


def reset_stats():
    for k in _stats:
        _stats[k] = 0


def _supported(x, normalized_shape, weight):
    if isinstance(x, FakeTensor) or not x.is_cuda or x.dtype not in _DT:
        return False
    if len(normalized_shape) != 1 or weight is None or weight.dtype != x.dtype:
        return False
    H = int(normalized_shape[0])
    per, hmax = (256, 2048) if x.dtype == torch.bfloat16 else (128, 1024)
    return H % per == 0 and H <= hmax and (H // per) in (1, 2, 3, 4, 6, 8) and x.numel() > 0


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def native_layer_norm(input, normalized_shape, weight, bias, eps):
    if not _supported(input, normalized_shape, weight) or (bias is not None and bias.dtype != input.dtype):
        if not isinstance(input, FakeTensor):
            _stats["aten_ln"] += 1
        return aten.native_layer_norm.default(input, normalized_shape, weight, bias, eps)
    x = input.contiguous()
    H = int(normalized_shape[0])
    rows = x.numel() // H
    y = torch.empty_like(x)
    stat_shape = list(x.shape[:-1]) + [1]
    mean = torch.empty(stat_shape, dtype=torch.float32, device=x.device)
    rstd = torch.empty(stat_shape, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    check(lib.edb_layer_norm_fwd(y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), x.data_ptr(),
                                 weight.data_ptr(), bias.data_ptr() if bias is not None else None,
                                 rows, H, float(eps), _DT[x.dtype], _stream(x)))
    _stats["edb_ln_fwd"] += 1
    return y, mean, rstd


def _workspace(H, device):
    key = (H, device)
    ws = _workspaces.get(key)
    if ws is None:
        nbytes = c_size_t()
        check(_lib.load().edb_layer_norm_bwd_workspace(H, byref(nbytes)))
        ws = torch.empty(max(16, nbytes.value), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def native_layer_norm_backward(grad_out, input, normalized_shape, mean, rstd, weight, bias,
                               output_mask, *, _add=None):
    """`_add`: a tensor of input's shape added to dx inside the kernel (the gradient accumulation
    `aten.add.Tensor(dx, residual_grad)` that follows in the traced backward pass)."""
    ok = (_supported(input, normalized_shape, weight) and grad_out.dtype == input.dtype
          and mean.dtype == torch.float32 and rstd.dtype == torch.float32 and output_mask[0])
    if ok and _add is not None:
        ok = _add.dtype == input.dtype and _add.shape == input.shape
    if not ok:
        if not isinstance(input, FakeTensor):
            _stats["aten_ln"] += 1
        res = aten.native_layer_norm_backward.default(grad_out, input, normalized_shape, mean, rstd,
                                                      weight, bias, output_mask)
        if _add is not None:
            res = (aten.add.Tensor(res[0], _add),) + tuple(res[1:])
        return res
    x = input.contiguous()
    dy = grad_out.contiguous()
    H = int(normalized_shape[0])
    rows = x.numel() // H
    dx = torch.empty_like(x)
    dw = torch.empty_like(weight) if output_mask[1] else None
    db = torch.empty_like(weight) if output_mask[2] else None
    ws = _workspace(H, x.device)
    lib = _lib.load()
    add = _add.contiguous() if _add is not None else None
    check(lib.edb_layer_norm_bwd_add(dx.data_ptr(), dw.data_ptr() if dw is not None else None,
                                     db.data_ptr() if db is not None else None, dy.data_ptr(),
                                     x.data_ptr(), mean.contiguous().data_ptr(),
                                     rstd.contiguous().data_ptr(), weight.data_ptr(),
                                     add.data_ptr() if add is not None else None, ws.data_ptr(), rows,
                                     H, _DT[x.dtype], _stream(x)))
    _stats["edb_ln_bwd"] += 1
    return dx, dw, db


_cs_workspaces = {}


def sum_dim_intlist(x, dim, keepdim=False, *, dtype=None):
    """aten.sum.dim_IntList; the column-sum case (2-D, dim == [0]) — every bias gradient of the
    train step — runs on edb_colsum, everything else on ATen."""
    ok = (not isinstance(x, FakeTensor) and x.is_cuda and x.dim() == 2 and list(dim) == [0]
          and dtype is None and x.dtype in _DT and x.stride(1) == 1 and x.numel() > 0)
    if ok:
        epv = 8 if x.dtype == torch.bfloat16 else 4
        ok = x.shape[1] % epv == 0 and x.stride(0) % epv == 0 and x.data_ptr() % 16 == 0 \
            and x.shape[0] >= 64
    if not ok:
        if not isinstance(x, FakeTensor):
            _stats["aten_sum"] += 1
        return aten.sum.dim_IntList(x, dim, keepdim, dtype=dtype)
    rows, cols = x.shape
    key = (cols, x.device)
    ws = _cs_workspaces.get(key)
    lib = _lib.load()
    if ws is None:
        nbytes = c_size_t()
        check(lib.edb_colsum_workspace(cols, byref(nbytes)))
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=x.device)
        _cs_workspaces[key] = ws
    out = torch.empty((1, cols) if keepdim else (cols,), dtype=x.dtype, device=x.device)
    check(lib.edb_colsum(out.data_ptr(), x.data_ptr(), ws.data_ptr(), rows, cols, x.stride(0),
                         _DT[x.dtype], _stream(x)))
    _stats["edb_colsum"] += 1
    return out

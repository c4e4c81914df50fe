"""Sharded-op kernel dispatch for the cross-entropy tail of a traced train step.

`lowering.fuse_cross_entropy` rewrites
    [_to_copy(fp32)] -> _log_softmax(dim=-1) -> nll_loss_forward        (forward)
    nll_loss_backward -> _log_softmax_backward_data -> [_to_copy(lp)]   (backward)
into `cross_entropy_fwd` / `cross_entropy_bwd` below, which run on edb_loss.cu through the C-ABI:
the logits are read once per direction in their storage dtype and the gradient is produced directly
in a TMA-legal layout for the LM-head GEMMs.  Semantics are those of the ATen ops with weight=None
(the graph the reference traces for F.cross_entropy: examples/torch/gpt_train.py:37-43).
"""
import torch
from torch._subclasses.fake_tensor import FakeTensor

from . import _lib
from ._lib import check

aten = torch.ops.aten
_stats = {"edb_ce_fwd": 0, "edb_ce_bwd": 0, "aten_ce": 0}
_DT = {torch.bfloat16: _lib.DTYPE_CODES["bfloat16"], torch.float32: _lib.DTYPE_CODES["float32"]}


def stats():
    return dict(_stats)


def reset_stats():
    for k in _stats:
        _stats[k] = 0


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _native_ok(logits, target):
    return (not isinstance(logits, FakeTensor) and logits.is_cuda and logits.dim() == 2
            and logits.dtype in _DT and logits.stride(1) == 1 and logits.numel() > 0
            and target.dtype == torch.int64 and target.dim() == 1 and target.is_contiguous())


def _aten_fwd(logits, target, ignore_index, reduction):
    ls = aten._log_softmax.default(logits.float(), 1, False)
    loss, tw = aten.nll_loss_forward.default(ls, target, None, reduction, ignore_index)
    # logsumexp per row, recovered from any column: lse = x - log_softmax(x)
    lse = logits[:, 0].float() - ls[:, 0]
    return loss, tw, lse


def cross_entropy_fwd(logits, target, ignore_index, reduction):
    """-> (loss fp32 scalar, total_weight fp32 scalar, lse fp32 [rows]) for logits [rows, vocab]."""
    if not _native_ok(logits, target):
        if not isinstance(logits, FakeTensor):
            _stats["aten_ce"] += 1
        return _aten_fwd(logits, target, ignore_index, reduction)
    rows, vocab = logits.shape
    dev = logits.device
    scal = torch.empty(2, dtype=torch.float32, device=dev)
    per_row = torch.empty((2, rows), dtype=torch.float32, device=dev)
    lib = _lib.load()
    check(lib.edb_cross_entropy_fwd(scal.data_ptr(), scal.data_ptr() + 4, per_row.data_ptr(),
                                    per_row.data_ptr() + 4 * rows, logits.data_ptr(), logits.stride(0),
                                    target.data_ptr(), rows, vocab, int(ignore_index), int(reduction),
                                    _DT[logits.dtype], _stream(logits)))
    _stats["edb_ce_fwd"] += 1
    return scal[0], scal[1], per_row[0]


def cross_entropy_bwd(grad_out, logits, target, lse, total_weight, ignore_index, reduction):
    """-> d loss / d logits, same dtype as `logits`, [rows, vocab] view of a buffer whose row stride
    is a multiple of 8 elements."""
    if not _native_ok(logits, target) or grad_out.numel() != 1:
        if not isinstance(logits, FakeTensor):
            _stats["aten_ce"] += 1
        ls = logits.float() - lse.unsqueeze(1)
        g = aten.nll_loss_backward.default(grad_out, ls, target, None, reduction, ignore_index,
                                           total_weight)
        return aten._log_softmax_backward_data.default(g, ls, 1, torch.float32).to(logits.dtype)
    rows, vocab = logits.shape
    ld_out = (vocab + 7) // 8 * 8
    out = torch.empty((rows, ld_out), dtype=logits.dtype, device=logits.device)
    g = grad_out if grad_out.dtype == torch.float32 else grad_out.float()
    lib = _lib.load()
    check(lib.edb_cross_entropy_bwd(out.data_ptr(), ld_out, logits.data_ptr(), logits.stride(0),
                                    target.data_ptr(), lse.data_ptr(), g.data_ptr(),
                                    total_weight.data_ptr(), rows, vocab, int(ignore_index),
                                    int(reduction), _DT[logits.dtype], _stream(logits)))
    _stats["edb_ce_bwd"] += 1
    return out if ld_out == vocab else out[:, :vocab]

"""CHECKER (test infrastructure, not product): end-to-end parity of a compiled train step against
vanilla PyTorch — the reference's own comparator (tests/test_torch/test_spmd.py:97-113 runs the
same train_step vanilla and compiled, then compares outputs, every parameter, every buffer and
every optimizer state with assert_close(rtol 1e-4, atol 1e-5)).

Used by bench.py's pre-timing `parity` leg (outside the timed region) and by the GPU tests.  The
reference side here is plain eager PyTorch in fp32 on the GPU ("highest" matmul precision, TF32
off): N micro-batches per step (one per data-parallel rank) with the gradients averaged, which is
the whole-global-batch step up to fp32 rounding.  bf16 runs cannot meet rtol 1e-4 (8 mantissa
bits), so their tolerance is calibrated instead of guessed: the same vanilla step in bf16 eager
PyTorch (cuBLAS) is measured against the fp32 reference too, and the compiled N-GPU run has to be
about as close to fp32 as that is.
"""
import math

import torch


def _sgd_step_fp32(model, opt, micro_batches, loss_fn):
    losses = []
    n = len(micro_batches)
    opt.zero_grad(set_to_none=True)
    for tok, tgt in micro_batches:
        loss = loss_fn(tok, tgt, model)
        (loss / n).backward()
        losses.append(float(loss.detach()))
    opt.step()
    opt.zero_grad(set_to_none=True)
    return losses


def vanilla_run(make_model, state_dict, steps, make_opt, dtype, device):
    """steps: list over optimisation steps of lists of (tokens, targets) micro-batches (all ranks'
    batches of that step, rank order).  Returns (per-step per-rank losses, params, optimizer
    states) of an eager run in `dtype`."""
    prev = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        model = make_model().to(device=device, dtype=dtype)
        model.load_state_dict({k: v.to(device=device, dtype=dtype) for k, v in state_dict.items()})
        opt = make_opt(model.parameters())
        losses = []
        for micro in steps:
            micro = [(t.to(device), y.to(device)) for t, y in micro]
            losses.append(_sgd_step_fp32(model, opt, micro, lambda t, y, m: m(t, y)))
        params = {k: v.detach().float() for k, v in model.named_parameters()}
        states = {}
        for name, p in model.named_parameters():
            st = opt.state.get(p, {})
            states[name] = {k: v.detach().float() for k, v in st.items()
                            if isinstance(v, torch.Tensor) and v.dim() > 0}
        return losses, params, states
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev


def gather_full(local, full_shape, group_size, process_group=None):
    """Parameter / optimizer-state shard of a zero2/zero3 run -> the full tensor (flat 1/n shards
    in rank order, compile_dp.py:330-343); tensors that are not sharded pass through."""
    numel = math.prod(full_shape)
    if local.numel() == numel:
        return local.reshape(full_shape)
    assert local.numel() * group_size == numel, (tuple(local.shape), tuple(full_shape), group_size)
    import torch.distributed as dist
    out = torch.empty(numel, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous().view(-1), group=process_group)
    return out.view(full_shape)


def rel_l2(got, want):
    want = want.float()
    den = float(want.norm())
    num = float((got.float() - want).norm())
    return num / den if den > 0 else (0.0 if num == 0 else float("inf"))


def bf16_ulps(got, want):
    """max |got - bf16(want)| in units of the bf16 spacing at max(|want|, rms(want)) (got: bf16
    values).  The rms floor keeps elements that are ~0 by cancellation (freshly updated zero-init
    biases) from turning the tensor's ordinary gradient noise into thousands of 'ulps'."""
    w = want.float().to(torch.bfloat16).float()
    g = got.float()
    rms = float(w.pow(2).mean().sqrt())
    mag = w.abs().clamp_min(max(rms, 1e-30))
    ulp = torch.exp2(torch.floor(torch.log2(mag)) - 7)
    return float(((g - w).abs() / ulp).max())


def compare(got_params, got_states, ref_params, ref_states, low_precision):
    """-> dict(checks, max_rel_err, param_max_ulp | param_rel, state_rel_l2, worst).
    fp32: elementwise relative error of every parameter and state (the reference's assert_close
    quantity, |a-b| / (atol/rtol + |b|) form reported as the largest violation ratio at rtol 1e-4 /
    atol 1e-5).  bf16: parameters in bf16 ulps of the fp32 reference, optimizer states as relative
    L2 error per tensor."""
    res = {"checks": 0, "worst": None}
    worst = 0.0
    if low_precision:
        ulps = 0.0
        for name, ref in ref_params.items():
            ulps = max(ulps, bf16_ulps(got_params[name], ref))
            res["checks"] += 1
        res["param_max_ulp"] = ulps
        srel = 0.0
        for name, st in ref_states.items():
            for key, ref in st.items():
                e = rel_l2(got_states[name][key], ref)
                res["checks"] += 1
                if e > srel:
                    srel, res["worst"] = e, f"{name}.{key}"
        res["state_rel_l2"] = srel
        res["max_rel_err"] = srel
        return res
    rtol, atol = 1e-4, 1e-5

    def violation(got, ref):
        return float(((got.float() - ref).abs() / (atol + rtol * ref.abs())).max())

    for name, ref in ref_params.items():
        v = violation(got_params[name], ref)
        res["checks"] += 1
        if v > worst:
            worst, res["worst"] = v, name
    for name, st in ref_states.items():
        for key, ref in st.items():
            v = violation(got_states[name][key], ref)
            res["checks"] += 1
            if v > worst:
                worst, res["worst"] = v, f"{name}.{key}"
    res["assert_close_violation"] = worst  # <= 1 passes the reference's assert_close
    res["max_rel_err"] = max(
        [rel_l2(got_params[n], r) for n, r in ref_params.items()] +
        [rel_l2(got_states[n][k], r) for n, st in ref_states.items() for k, r in st.items()] + [0.0])
    return res


def compiled_state(compiled, ref_params, ref_states, group_size, process_group=None):
    """Full parameters / optimizer states of an EDCompiledFunc (gathering zero2/zero3 shards)."""
    params, _, named_states = compiled.get_state()
    got_p = {n: gather_full(params[n].detach(), ref.shape, group_size, process_group)
             for n, ref in ref_params.items()}
    got_s = {}
    for n, st in ref_states.items():
        got_s[n] = {k: gather_full(named_states[n][k].detach(), ref.shape, group_size, process_group)
                    for k, ref in st.items()}
    return got_p, got_s

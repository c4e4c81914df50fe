"""ORACLE (test infrastructure): the reference's path for a train step restated on CPU.

In the reference the compiled train step is an FX graph of ATen ops executed eagerly op by op
(easydist/torch/compile_auto.py:752-756) with gloo collectives between ranks on CPU.  At world
size 1 the graph has no collective, so the CPU restatement is the same train step run eagerly by
PyTorch on the host in fp32 (the reference's CPU runs are fp32; its own comparator is exactly
"vanilla PyTorch vs compiled", tests/test_torch/test_spmd.py:97-113).  Used as
  * the parity checker of smoke() and the GPU train-step tests (loss trajectory), and
  * bench.py's `cpu_baseline` / `--impl reference` leg (timed on a bounded sample).
Never imported by easydist_b200/.
"""
import dataclasses
import time

import torch


def build(model_name, attn, seq, seed=0, dtype=torch.float32, device="cpu"):
    from easydist_b200.workloads import GPT2, GPT2_CONFIGS
    cfg = dataclasses.replace(GPT2_CONFIGS[model_name], attn=attn,
                              block_size=max(seq, GPT2_CONFIGS[model_name].block_size))
    torch.manual_seed(seed)
    model = GPT2(cfg).to(device=device, dtype=dtype)
    return cfg, model


def train_losses(model_name, attn, batch, seq, steps, lr=1e-3, momentum=0.9, seed=0,
                 state_dict=None):
    """Loss of each of `steps` SGD-momentum steps on CPU/fp32 (same synthetic batches as the GPU
    side: synthetic_tokens(cfg, batch, seq, seed=1000*b))."""
    from easydist_b200.workloads import gpt2_train_step, synthetic_tokens
    cfg, model = build(model_name, attn, seq, seed)
    if state_dict is not None:
        model.load_state_dict({k: v.float().cpu() for k, v in state_dict.items()})
    opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=momentum, foreach=True)
    losses = []
    for b in range(steps):
        tok, tgt = synthetic_tokens(cfg, batch, seq, seed=1000 * b)
        losses.append(float(gpt2_train_step(tok, tgt, model, opt).detach()))
    return losses, model


def time_cpu_train_step(model_name, attn, n_seqs, seq, steps=1, warmup=0):
    from easydist_b200.workloads import gpt2_train_step, synthetic_tokens
    cfg, model = build(model_name, attn, seq)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, foreach=True)
    tok, tgt = synthetic_tokens(cfg, n_seqs, seq, seed=0)
    loss = gpt2_train_step(tok[:1], tgt[:1], model, opt)  # allocations, thread pool
    for _ in range(warmup):
        loss = gpt2_train_step(tok, tgt, model, opt)
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = gpt2_train_step(tok, tgt, model, opt)
    return time.perf_counter() - t0, float(loss.detach())

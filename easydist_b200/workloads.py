"""Workload definitions for the measured configurations (BASELINE.json `configs`).

The reference ships no GPT-2 / Llama definition (benchmark/torch/model/__init__.py advertises
LLAMA but defines none; its GPT in benchmark/torch/model/gpt.py takes embeddings, not token ids —
SURVEY.md App. C-7), so the token-level GPT-2 used for config 2 is defined here from the published
architecture: learned token + position embeddings, pre-LN blocks, GELU MLP, tied LM head,
cross-entropy loss.  Synthetic tokens and random-init weights (no network for data/checkpoints).
"""
import math
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class GPT2Config:
    n_layer: int = 24
    n_head: int = 16
    n_embd: int = 1024
    vocab_size: int = 50257
    block_size: int = 512
    attn: str = "sdpa"  # "sdpa" (ATen fused attention) or "unfused" (matmul-softmax-matmul, as in
    #                      the reference's benchmark/torch/model/gpt.py:23-42)


GPT2_CONFIGS = {
    "gpt2-medium": GPT2Config(24, 16, 1024),
    "gpt2-small": GPT2Config(12, 12, 768),
    "gpt2-tiny": GPT2Config(2, 4, 128, vocab_size=512, block_size=64),
}


class Attention(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.c_attn = nn.Linear(cfg.n_embd, 3 * cfg.n_embd)
        self.c_proj = nn.Linear(cfg.n_embd, cfg.n_embd)
        self.n_head, self.n_embd, self.mode = cfg.n_head, cfg.n_embd, cfg.attn
        if cfg.attn == "unfused":
            mask = torch.tril(torch.ones(cfg.block_size, cfg.block_size, dtype=torch.bool))
            self.register_buffer("mask", mask.view(1, 1, cfg.block_size, cfg.block_size),
                                 persistent=False)

    def forward(self, x):
        B, T, C = x.shape
        q, k, v = self.c_attn(x).split(self.n_embd, dim=2)
        hd = C // self.n_head
        q = q.view(B, T, self.n_head, hd).transpose(1, 2)
        k = k.view(B, T, self.n_head, hd).transpose(1, 2)
        v = v.view(B, T, self.n_head, hd).transpose(1, 2)
        if self.mode == "sdpa":
            y = F.scaled_dot_product_attention(q, k, v, is_causal=True)
        else:
            att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(hd))
            att = att.masked_fill(~self.mask[:, :, :T, :T], float("-inf"))
            y = F.softmax(att, dim=-1) @ v
        return self.c_proj(y.transpose(1, 2).contiguous().view(B, T, C))


class Block(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.ln_1 = nn.LayerNorm(cfg.n_embd)
        self.attn = Attention(cfg)
        self.ln_2 = nn.LayerNorm(cfg.n_embd)
        self.c_fc = nn.Linear(cfg.n_embd, 4 * cfg.n_embd)
        self.c_proj = nn.Linear(4 * cfg.n_embd, cfg.n_embd)

    def forward(self, x):
        x = x + self.attn(self.ln_1(x))
        return x + self.c_proj(F.gelu(self.c_fc(self.ln_2(x)), approximate="tanh"))


class GPT2(nn.Module):
    def __init__(self, cfg: GPT2Config):
        super().__init__()
        self.cfg = cfg
        self.wte = nn.Embedding(cfg.vocab_size, cfg.n_embd)
        self.wpe = nn.Embedding(cfg.block_size, cfg.n_embd)
        self.h = nn.ModuleList(Block(cfg) for _ in range(cfg.n_layer))
        self.ln_f = nn.LayerNorm(cfg.n_embd)
        self.lm_head = nn.Linear(cfg.n_embd, cfg.vocab_size, bias=False)
        self.lm_head.weight = self.wte.weight  # tied, as published
        self.apply(self._init)

    @staticmethod
    def _init(m):
        if isinstance(m, (nn.Linear, nn.Embedding)):
            nn.init.normal_(m.weight, mean=0.0, std=0.02)
        if isinstance(m, nn.Linear) and m.bias is not None:
            nn.init.zeros_(m.bias)

    def forward(self, idx, targets):
        B, T = idx.shape
        pos = torch.arange(T, device=idx.device)
        x = self.wte(idx) + self.wpe(pos)
        for blk in self.h:
            x = blk(x)
        logits = self.lm_head(self.ln_f(x))
        return F.cross_entropy(logits.view(-1, logits.size(-1)).float(), targets.view(-1))

    def matmul_params(self):
        """Parameters that take part in a GEMM per token (for the 6*P*tokens closed form)."""
        return sum(p.numel() for n, p in self.named_parameters() if p.dim() == 2 and "wpe" not in n)


def gpt2_train_step(tokens, targets, model, opt):
    """One optimisation step; same shape as the reference's examples (gpt_train.py:37-43)."""
    loss = model(tokens, targets)
    loss.backward()
    opt.step()
    opt.zero_grad(True)
    return loss


def synthetic_tokens(cfg, batch, seq, seed, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    tok = torch.randint(0, cfg.vocab_size, (batch, seq + 1), generator=g)
    return tok[:, :-1].contiguous().to(device), tok[:, 1:].contiguous().to(device)


def train_flops_per_step(cfg, batch, seq):
    """Dense-contraction FLOPs of one fwd+bwd step: 6 * P_matmul * tokens + attention
    (12 * L * H * S per token), SURVEY.md §8(d)."""
    tokens = batch * seq
    p = cfg.n_layer * 12 * cfg.n_embd ** 2 + cfg.vocab_size * cfg.n_embd
    return 6 * p * tokens + 12 * cfg.n_layer * cfg.n_embd * seq * tokens

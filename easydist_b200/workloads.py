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
    pos_as_buffer: bool = False  # position ids from a registered buffer instead of torch.arange:
    #                              the reference's sharding discovery cannot annotate tensor-less
    #                              factory ops (used when a plan is recorded with its solver)


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
        if cfg.pos_as_buffer:
            self.register_buffer("pos_ids", torch.arange(cfg.block_size), persistent=False)
        self.apply(self._init)

    @staticmethod
    def _init(m):
        if isinstance(m, (nn.Linear, nn.Embedding)):
            nn.init.normal_(m.weight, mean=0.0, std=0.02)
        if isinstance(m, nn.Linear) and m.bias is not None:
            nn.init.zeros_(m.bias)

    def forward(self, idx, targets):
        B, T = idx.shape
        if self.cfg.pos_as_buffer:
            pos = self.pos_ids if T == self.pos_ids.shape[0] else self.pos_ids[:T]
        else:
            pos = torch.arange(T, device=idx.device)
        x = self.wte(idx) + self.wpe(pos)
        for blk in self.h:
            x = blk(x)
        logits = self.lm_head(self.ln_f(x))
        return F.cross_entropy(logits.view(-1, logits.size(-1)).float(), targets.view(-1))

    def matmul_params(self):
        """Parameters that take part in a GEMM per token (for the 6*P*tokens closed form)."""
        return sum(p.numel() for n, p in self.named_parameters() if p.dim() == 2 and "wpe" not in n)


@dataclass
class LlamaConfig:
    n_layer: int = 32
    n_head: int = 32
    n_embd: int = 4096
    ffn: int = 11008
    vocab_size: int = 32000
    block_size: int = 2048
    eps: float = 1e-5


LLAMA_CONFIGS = {
    "llama2-7b": LlamaConfig(),
    "llama-tiny": LlamaConfig(2, 4, 64, 176, vocab_size=256, block_size=64),
}


class RMSNorm(nn.Module):
    def __init__(self, d, eps):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.eps = eps

    def forward(self, x):
        v = x.float().pow(2).mean(-1, keepdim=True)
        return (x.float() * torch.rsqrt(v + self.eps)).to(x.dtype) * self.weight


def _rope(x, cos, sin):
    # x: [B, H, T, hd]; rotate pairs (first half, second half), the published Llama-2 layout
    h = x.shape[-1] // 2
    x1, x2 = x[..., :h], x[..., h:]
    return torch.cat((x1 * cos - x2 * sin, x2 * cos + x1 * sin), dim=-1)


class LlamaBlock(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        d = cfg.n_embd
        self.n_head = cfg.n_head
        self.ln1 = RMSNorm(d, cfg.eps)
        self.wq = nn.Linear(d, d, bias=False)
        self.wk = nn.Linear(d, d, bias=False)
        self.wv = nn.Linear(d, d, bias=False)
        self.wo = nn.Linear(d, d, bias=False)
        self.ln2 = RMSNorm(d, cfg.eps)
        self.w_gate = nn.Linear(d, cfg.ffn, bias=False)
        self.w_up = nn.Linear(d, cfg.ffn, bias=False)
        self.w_down = nn.Linear(cfg.ffn, d, bias=False)

    def forward(self, x, cos, sin):
        B, T, C = x.shape
        h = self.ln1(x)
        q, k, v = (w(h).view(B, T, self.n_head, C // self.n_head).transpose(1, 2)
                   for w in (self.wq, self.wk, self.wv))
        y = F.scaled_dot_product_attention(_rope(q, cos, sin), _rope(k, cos, sin), v, is_causal=True)
        x = x + self.wo(y.transpose(1, 2).contiguous().view(B, T, C))
        h = self.ln2(x)
        return x + self.w_down(F.silu(self.w_gate(h)) * self.w_up(h))


class Llama(nn.Module):
    """Llama-2 architecture (BASELINE.json config 4) from the published description: RMSNorm,
    rotary position embedding, SwiGLU MLP, no biases, untied LM head.  Random init; the reference
    ships no Llama definition either (benchmark/torch/model/__init__.py advertises one)."""

    def __init__(self, cfg: LlamaConfig):
        super().__init__()
        self.cfg = cfg
        self.tok = nn.Embedding(cfg.vocab_size, cfg.n_embd)
        self.h = nn.ModuleList(LlamaBlock(cfg) for _ in range(cfg.n_layer))
        self.norm = RMSNorm(cfg.n_embd, cfg.eps)
        self.lm_head = nn.Linear(cfg.n_embd, cfg.vocab_size, bias=False)
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Embedding)):
                nn.init.normal_(m.weight, mean=0.0, std=0.02)

    def forward(self, idx, targets):
        B, T = idx.shape
        hd = self.cfg.n_embd // self.cfg.n_head
        inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, device=idx.device).float() / hd))
        ang = torch.arange(T, device=idx.device).float()[:, None] * inv[None, :]
        x = self.tok(idx)
        cos, sin = ang.cos().to(x.dtype), ang.sin().to(x.dtype)
        for blk in self.h:
            x = blk(x, cos, sin)
        logits = self.lm_head(self.norm(x))
        return F.cross_entropy(logits.view(-1, logits.size(-1)).float(), targets.view(-1))


class _RefAttn(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.query, self.key, self.value, self.dense = (nn.Linear(dim, dim) for _ in range(4))

    def forward(self, x):
        B, T, C = x.shape
        q, k, v = (f(x).view(B, T, self.heads, C // self.heads).permute(0, 2, 1, 3)
                   for f in (self.query, self.key, self.value))
        att = (q @ k.transpose(-1, -2)) / math.sqrt(C // self.heads)
        keep = torch.ones(T, T, dtype=torch.uint8, device=x.device).tril().view(1, 1, T, T).bool()
        att = torch.where(keep, att, -1e4).softmax(-1)
        return self.dense((att @ v).transpose(1, 2).reshape(B, T, C))


class _RefMlp(nn.Module):
    def __init__(self, dim, ratio):
        super().__init__()
        self.dense_h_to_4h = nn.Linear(dim, dim * ratio)
        self.dense_4h_to_h = nn.Linear(dim * ratio, dim)

    def forward(self, x):
        return self.dense_4h_to_h(F.gelu(self.dense_h_to_4h(x)))


class _RefBlock(nn.Module):
    def __init__(self, dim, heads, ratio):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _RefAttn(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _RefMlp(dim, ratio)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class EmbeddingGPT(nn.Module):
    """The architecture of the reference's example / test model (benchmark/torch/model/gpt.py:
    pre-LN blocks with eps 1e-6, separate q/k/v/dense Linears, matmul-softmax attention with a
    -1e4 causal fill, GELU MLP; input = embeddings, no token layer) with the SAME parameter names,
    so that plan bundles recorded with the reference's solver on that model (SURVEY.md config 1:
    depth 4, dim 1024, 32 heads) can be lowered and run where the reference is not importable."""

    def __init__(self, depth, dim, num_heads, mlp_ratio=4):
        super().__init__()
        self.blocks = nn.ModuleList(_RefBlock(dim, num_heads, mlp_ratio) for _ in range(depth))

    def forward(self, x):
        for blk in self.blocks:
            x = blk(x)
        return x


def embedding_gpt_train_step(input, model, opt):
    """examples/torch/gpt_train.py:37-43 / tests/test_torch/test_utils.py train_step."""
    out = model(input)
    loss = out.mean()
    loss.backward()
    opt.step()
    opt.zero_grad()
    return out


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

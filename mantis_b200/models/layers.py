"""nn.Module building blocks whose forward() runs the hand-written CUDA kernels (mantis_b200.ops).

They subclass the torch modules they replace so that parameter names, state-dict layout, peft/LoRA leaf-name
matching (`q_proj`, `fc1`, ...) and HF weight init keep working unchanged.
"""
import math

import torch
from torch import nn

from .. import ops


class B200Linear(nn.Linear):
    """nn.Linear whose forward/backward are the tcgen05 GEMM (bf16) or the SIMT GEMM (fp32 / tiny shapes)."""

    def forward(self, x, act=None, residual=None):
        return ops.linear(x, self.weight, self.bias, act=act, residual=residual)


class B200LayerNorm(nn.LayerNorm):
    def forward(self, x):
        return ops.layer_norm(x, self.weight, self.bias, self.eps)


class B200RMSNorm(nn.Module):
    """LlamaRMSNorm / MistralRMSNorm / Idefics2RMSNorm (transformers llama/modeling_llama.py:53-68)"""

    def __init__(self, hidden_size, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps

    def forward(self, x):
        return ops.rms_norm(x, self.weight, self.variance_epsilon)

    def extra_repr(self):
        return f"{tuple(self.weight.shape)}, eps={self.variance_epsilon}"


class B200Embedding(nn.Embedding):
    def forward(self, ids):
        return ops.embedding(ids, self.weight)


def default_inv_freq(head_dim, theta, device=None):
    """HF default RoPE init: 1 / theta^(arange(0, dim, 2) / dim) computed in fp32 (bit-identical construction)."""
    return 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).to(device=device, dtype=torch.float) / head_dim))


def llama3_inv_freq(head_dim, theta, rope_scaling, device=None):
    """llama3 rope scaling (transformers modeling_rope_utils._compute_llama3_parameters)"""
    inv_freq = default_inv_freq(head_dim, theta, device)
    factor = rope_scaling["factor"]
    low = rope_scaling["low_freq_factor"]
    high = rope_scaling["high_freq_factor"]
    old_len = rope_scaling["original_max_position_embeddings"]
    low_wl, high_wl = old_len / low, old_len / high
    wavelen = 2 * math.pi / inv_freq
    inv_llama = torch.where(wavelen > low_wl, inv_freq / factor, inv_freq)
    smooth = (old_len / wavelen - low) / (high - low)
    smoothed = (1 - smooth) * inv_llama / factor + smooth * inv_llama
    is_medium = ~(wavelen < high_wl) * ~(wavelen > low_wl)
    return torch.where(is_medium, smoothed, inv_llama)

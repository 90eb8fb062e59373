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


# ---------------------------------------------------------------------------------------------- HF plumbing helpers
def _fresh(t):
    """transformers >= 5 flags tensors that were just loaded from a checkpoint; never re-initialise those"""
    return t is not None and not getattr(t, "_is_hf_initialized", False)


def init_module_weights(module, std):
    """normal(0, std) for Linear / Conv / Embedding weights, zeros for biases, ones/zeros for norms -- the rule the
    reference's `_init_weights` applies (mantis/models/mllava/modeling_llava.py:153-170) -- skipping loaded tensors."""
    if hasattr(module, "class_embedding") and _fresh(module.class_embedding):
        module.class_embedding.data.normal_(mean=0.0, std=std)
    if isinstance(module, (nn.Linear, nn.Conv2d)):
        if _fresh(module.weight):
            module.weight.data.normal_(mean=0.0, std=std)
        if _fresh(module.bias):
            module.bias.data.zero_()
    elif isinstance(module, nn.Embedding):
        if _fresh(module.weight):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.padding_idx is not None:
                module.weight.data[module.padding_idx].zero_()
    elif isinstance(module, nn.LayerNorm):
        if _fresh(module.weight):
            module.weight.data.fill_(1.0)
        if _fresh(module.bias):
            module.bias.data.zero_()
    elif isinstance(module, B200RMSNorm):
        if _fresh(module.weight):
            module.weight.data.fill_(1.0)


from contextlib import contextmanager


@contextmanager
def hf_key_remap_disabled(*model_types):
    """transformers >= 5 registers checkpoint *key renamings* per `model_type` (e.g. "llava": language_model.model.* ->
    language_model.*) for its own re-organised classes.  Our classes keep the reference's module tree / key layout, so
    those renamings must not be applied while loading or saving them.  Restores the registry afterwards."""
    try:
        from transformers import conversion_mapping as cm
    except Exception:       # older transformers: nothing to disable
        yield
        return
    cm.get_checkpoint_conversion_mapping("llava")          # make sure the registry is built
    cache = cm._checkpoint_conversion_mapping_cache
    saved = {mt: cache.pop(mt) for mt in model_types if mt in cache}
    try:
        yield
    finally:
        cache.update(saved)

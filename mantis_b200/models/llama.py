"""LLaMA / Mistral decoder stack on the mantis_b200 kernels.

Mirrors transformers' LlamaForCausalLM / LlamaModel / MistralModel (llama/modeling_llama.py:53-499,
mistral/modeling_mistral.py) -- same module tree and parameter names (`model.embed_tokens`, `model.layers.N.
self_attn.{q,k,v,o}_proj`, `mlp.{gate,up,down}_proj`, `input_layernorm`, `post_attention_layernorm`, `model.norm`,
`lm_head`) so reference checkpoints load unchanged -- but every op is one of our CUDA kernels:
RMSNorm -> tcgen05 GEMMs (q,k,v) -> RoPE with caller-supplied position_ids -> causal GQA attention with key
padding mask (+ KV cache) -> o_proj with fused residual -> RMSNorm -> gate/up GEMMs -> SwiGLU -> down_proj with
fused residual.
"""
import torch
from torch import nn
from transformers import PreTrainedModel
from transformers.modeling_outputs import BaseModelOutputWithPast, CausalLMOutputWithPast

from .. import ops
from .kv_cache import B200KVCache
from .layers import (B200Embedding, B200Linear, B200RMSNorm, default_inv_freq, init_module_weights,
                     llama3_inv_freq)

try:
    from transformers import LlamaConfig, MistralConfig
except Exception:  # pragma: no cover
    LlamaConfig = MistralConfig = None


def _plain(*mods):
    """True if every module is an unwrapped, bias-free B200Linear (so its weight can be used directly)"""
    return all(type(m) is B200Linear and m.bias is None for m in mods)


def _rope_params(config):
    theta = getattr(config, "rope_theta", None)
    rs = getattr(config, "rope_scaling", None)
    rp = getattr(config, "rope_parameters", None)
    if isinstance(rp, dict):
        theta = rp.get("rope_theta", theta)
        if rp.get("rope_type", "default") != "default":
            rs = rp
    if theta is None:
        theta = 10000.0
    return float(theta), rs


def window_key_mask(key_mask, total, window, batch, device):
    """decode step over a context longer than the sliding window: the new token (index total-1) sees key j iff
    (total-1) - j < window, so the keys before total - window are masked out of the (optional) padding mask"""
    keep = torch.arange(total, device=device) >= (total - window)
    if key_mask is None:
        return keep.to(torch.int64).unsqueeze(0).expand(batch, total).contiguous()
    return key_mask * keep.to(key_mask.dtype).unsqueeze(0)


class B200Attention(nn.Module):
    def __init__(self, config, layer_idx):
        super().__init__()
        self.layer_idx = layer_idx
        self.hidden_size = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.num_kv_heads = getattr(config, "num_key_value_heads", None) or self.num_heads
        self.head_dim = getattr(config, "head_dim", None) or self.hidden_size // self.num_heads
        bias = bool(getattr(config, "attention_bias", False))
        self.q_proj = B200Linear(self.hidden_size, self.num_heads * self.head_dim, bias=bias)
        self.k_proj = B200Linear(self.hidden_size, self.num_kv_heads * self.head_dim, bias=bias)
        self.v_proj = B200Linear(self.hidden_size, self.num_kv_heads * self.head_dim, bias=bias)
        self.o_proj = B200Linear(self.num_heads * self.head_dim, self.hidden_size, bias=bias)
        self.scaling = self.head_dim ** -0.5
        # Mistral (Idefics2's text model): key j is visible to query i iff i - j < sliding_window
        # (transformers mistral/modeling_mistral.py); LLaMA configs carry no such field
        self.sliding_window = getattr(config, "sliding_window", None)

    def forward(self, x, residual, position_ids, inv_freq, rope_scale, key_mask, cache, kbits=None, rope_tab=None,
                segs=None):
        B, S, _ = x.shape
        if _plain(self.q_proj, self.k_proj, self.v_proj):
            q, k, v = ops.multi_linear(x, self.q_proj.weight, self.k_proj.weight, self.v_proj.weight)
        else:                                  # biased or wrapped (peft LoRA) projections: go through the modules
            q, k, v = self.q_proj(x), self.k_proj(x), self.v_proj(x)
        q = q.view(B, S, self.num_heads, self.head_dim)
        k = k.view(B, S, self.num_kv_heads, self.head_dim)
        v = v.view(B, S, self.num_kv_heads, self.head_dim)
        q, k = ops.rope(q, k, position_ids, inv_freq, rope_scale, rope_tab)
        if cache is not None:
            if (S == 1 and self.head_dim == 128 and q.dtype == torch.bfloat16 and not torch.is_grad_enabled()
                    and not ops.FORCE_GENERIC):
                ctx = cache.write(k, v, self.layer_idx)                 # split-KV decode kernel reads the pages in place
                o = ops.decode_attention_paged(q, cache, self.layer_idx, ctx, key_mask, self.scaling, kbits=kbits)
                return self.o_proj(o.view(B, S, self.num_heads * self.head_dim), residual=residual)
            k, v = cache.append(k, v, self.layer_idx)
        if segs is not None:                   # packed sequences: block-diagonal causal attention, one launch per segment
            o = ops.attention_varlen(q, k, v, segs, kmask=key_mask, scale=self.scaling)
        else:
            o = ops.attention(q, k, v, causal=True, kmask=key_mask, scale=self.scaling, window=self.sliding_window)
        return self.o_proj(o.view(B, S, self.num_heads * self.head_dim), residual=residual)


class B200MLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        bias = bool(getattr(config, "mlp_bias", False))
        self.gate_proj = B200Linear(config.hidden_size, config.intermediate_size, bias=bias)
        self.up_proj = B200Linear(config.hidden_size, config.intermediate_size, bias=bias)
        self.down_proj = B200Linear(config.intermediate_size, config.hidden_size, bias=bias)

    def forward(self, x, residual=None):
        if (_plain(self.gate_proj, self.up_proj, self.down_proj)
                and ops.swiglu_mlp_ok(x, self.gate_proj.weight, self.up_proj.weight, self.down_proj.weight)):
            return ops.swiglu_mlp(x, self.gate_proj.weight, self.up_proj.weight, self.down_proj.weight, residual)
        if _plain(self.gate_proj, self.up_proj):
            g, u = ops.multi_linear(x, self.gate_proj.weight, self.up_proj.weight)
        else:
            g, u = self.gate_proj(x), self.up_proj(x)
        return self.down_proj(ops.swiglu(g, u), residual=residual)


class B200DecoderLayer(nn.Module):
    def __init__(self, config, layer_idx):
        super().__init__()
        self.self_attn = B200Attention(config, layer_idx)
        self.mlp = B200MLP(config)
        self.input_layernorm = B200RMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.post_attention_layernorm = B200RMSNorm(config.hidden_size, eps=config.rms_norm_eps)

    def forward(self, x, position_ids, inv_freq, rope_scale, key_mask, cache, kbits=None, rope_tab=None, segs=None):
        n1, n2 = self.input_layernorm, self.post_attention_layernorm
        if torch.is_grad_enabled() and x.requires_grad:
            # training: the residual stream leaves each norm together with the normed activations, so the two gradients that
            # meet at x (through the norm and around it) are summed inside the norm's backward kernel
            h, x = ops.rms_norm_res(x, n1.weight, n1.variance_epsilon)
            x = self.self_attn(h, x, position_ids, inv_freq, rope_scale, key_mask, cache, kbits, rope_tab, segs)
            h, x = ops.rms_norm_res(x, n2.weight, n2.variance_epsilon)
            return self.mlp(h, residual=x)
        x = self.self_attn(n1(x), x, position_ids, inv_freq, rope_scale, key_mask, cache, kbits, rope_tab, segs)
        x = self.mlp(n2(x), residual=x)
        return x


class B200DecoderPreTrainedModel(PreTrainedModel):
    base_model_prefix = "model"
    supports_gradient_checkpointing = True
    _no_split_modules = ["B200DecoderLayer"]
    _supports_sdpa = True
    _supports_flash_attn = True
    _supports_flash_attn_2 = True

    def _init_weights(self, module):
        init_module_weights(module, getattr(self.config, "initializer_range", 0.02))


class B200DecoderModel(B200DecoderPreTrainedModel):
    """== LlamaModel / MistralModel.  Mistral's sliding window is honoured once the context exceeds it: prefill through the
    window-aware attention kernel, single-token decode by masking the keys that have slid out (`window_key_mask`)."""

    def __init__(self, config):
        super().__init__(config)
        self.padding_idx = getattr(config, "pad_token_id", None)
        self.vocab_size = config.vocab_size
        self.embed_tokens = B200Embedding(config.vocab_size, config.hidden_size, self.padding_idx)
        self.layers = nn.ModuleList([B200DecoderLayer(config, i) for i in range(config.num_hidden_layers)])
        self.norm = B200RMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.gradient_checkpointing = False
        self._inv_freq = {}
        self.post_init()

    def get_input_embeddings(self):
        return self.embed_tokens

    def set_input_embeddings(self, value):
        self.embed_tokens = value

    def rope_tables(self, device):
        key = str(device)
        if key not in self._inv_freq:
            hd = self.layers[0].self_attn.head_dim
            theta, rs = _rope_params(self.config)
            scale = 1.0
            if rs and rs.get("rope_type", rs.get("type", "default")) == "llama3":
                inv = llama3_inv_freq(hd, theta, rs, device)
            else:
                inv = default_inv_freq(hd, theta, device)
            self._inv_freq[key] = (inv.contiguous(), scale)
        return self._inv_freq[key]

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                inputs_embeds=None, use_cache=None, output_attentions=None, output_hidden_states=None,
                return_dict=None, cache_position=None, **kwargs):
        if (input_ids is None) == (inputs_embeds is None):
            raise ValueError("You must specify exactly one of input_ids or inputs_embeds")
        if inputs_embeds is None:
            inputs_embeds = self.embed_tokens(input_ids)
        B, S, _ = inputs_embeds.shape
        use_cache = bool(use_cache) if use_cache is not None else False
        cache = None
        if use_cache or past_key_values is not None:
            if past_key_values is None or not isinstance(past_key_values, B200KVCache):
                if past_key_values is not None and hasattr(past_key_values, "get_seq_length") and past_key_values.get_seq_length() > 0:
                    raise ValueError("mantis_b200 needs its own B200KVCache for cached decoding")
                past_key_values = B200KVCache()
            cache = past_key_values
            if cache.n_layers is None:
                cache.n_layers = len(self.layers)
        past = cache.get_seq_length() if cache is not None else 0
        if position_ids is None:
            position_ids = torch.arange(past, past + S, device=inputs_embeds.device).unsqueeze(0).expand(B, S)
        elif position_ids.shape[-1] != S:
            raise ValueError(f"position_ids cover {position_ids.shape[-1]} positions, the input has {S} tokens")
        key_mask = None
        segs = kwargs.get("cu_segments")
        if attention_mask is not None and attention_mask.dim() == 4:
            # sequence packing (ref: mantis/train/data.py:1622-1645): [B, 1, S, S] block-diagonal 0/1 mask + per-sample
            # position ids.  The blocks are recovered from the position-id restarts, the key padding from the mask columns.
            if cache is not None or position_ids is None:
                raise ValueError("a packed (4-D) attention_mask needs position_ids and cannot be combined with a KV cache")
            if attention_mask.shape[-2:] != (S, S):
                raise ValueError(f"packed attention_mask must be [B, 1, {S}, {S}]")
            if position_ids.dim() == 1:
                position_ids = position_ids.unsqueeze(0)
            if segs is None:
                segs = ops.packed_segments(position_ids)
            attention_mask = (attention_mask != 0).any(dim=1).any(dim=1).to(torch.int64)       # [B, S] key validity
        elif segs is not None and cache is not None:
            raise ValueError("packed sequences cannot be combined with a KV cache")
        if attention_mask is not None:
            if attention_mask.dim() != 2:
                raise ValueError("mantis_b200 expects a 2-D [batch, kv_len] or a packed 4-D attention_mask")
            key_mask = attention_mask
            if key_mask.shape[1] != past + S:
                raise ValueError(f"attention_mask length {key_mask.shape[1]} != past({past}) + seq({S})")
        window = getattr(self.config, "sliding_window", None)
        if window and S == 1 and cache is not None and past + 1 > window:
            key_mask = window_key_mask(key_mask, past + 1, window, B, inputs_embeds.device)
        inv_freq, rope_scale = self.rope_tables(inputs_embeds.device)
        x = inputs_embeds
        kbits = None
        if (S == 1 and cache is not None and key_mask is not None and x.dtype == torch.bfloat16
                and not torch.is_grad_enabled()):
            kbits = ops.kmask_bits(key_mask)                 # one bitmask per decode step, shared by all layers
        rope_tab = None
        if x.dtype == torch.bfloat16 and x.is_cuda and S > 1:
            rope_tab = ops.rope_table(position_ids, inv_freq, self.layers[0].self_attn.head_dim, rope_scale, x.dtype)
        all_hidden = () if output_hidden_states else None
        for layer in self.layers:
            if output_hidden_states:
                all_hidden += (x,)
            if self.gradient_checkpointing and self.training and cache is None:
                x = torch.utils.checkpoint.checkpoint(layer, x, position_ids, inv_freq, rope_scale, key_mask, None, None,
                                                      rope_tab, segs, use_reentrant=False)
            else:
                x = layer(x, position_ids, inv_freq, rope_scale, key_mask, cache, kbits, rope_tab, segs)
        x = self.norm(x)
        if output_hidden_states:
            all_hidden += (x,)
        return BaseModelOutputWithPast(last_hidden_state=x, past_key_values=cache, hidden_states=all_hidden,
                                       attentions=None)


class B200CausalLM(B200DecoderPreTrainedModel):
    """== LlamaForCausalLM: `.model` + `.lm_head` (untied for LLaMA-3)."""
    _tied_weights_keys = {}

    def __init__(self, config):
        super().__init__(config)
        self.model = B200DecoderModel(config)
        self.vocab_size = config.vocab_size
        self.lm_head = B200Linear(config.hidden_size, config.vocab_size, bias=False)
        self.post_init()

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def set_input_embeddings(self, value):
        self.model.embed_tokens = value

    def get_output_embeddings(self):
        return self.lm_head

    def set_output_embeddings(self, new):
        self.lm_head = new

    def resize_token_embeddings(self, *args, **kwargs):
        out = super().resize_token_embeddings(*args, **kwargs)
        # HF builds plain nn.Embedding / nn.Linear replacements: put them back on the CUDA-kernel classes
        emb, head = self.get_input_embeddings(), self.get_output_embeddings()
        if type(emb) is nn.Embedding:
            emb.__class__ = B200Embedding
        if type(head) is nn.Linear:
            head.__class__ = B200Linear
        return self.get_input_embeddings() if out is not None else out

    def set_decoder(self, decoder):
        self.model = decoder

    def get_decoder(self):
        return self.model

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                inputs_embeds=None, labels=None, use_cache=None, output_attentions=None, output_hidden_states=None,
                return_dict=None, logits_to_keep=0, **kwargs):
        out = self.model(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                         past_key_values=past_key_values, inputs_embeds=inputs_embeds, use_cache=use_cache,
                         output_hidden_states=output_hidden_states, cu_segments=kwargs.get("cu_segments"))
        h = out.last_hidden_state
        if isinstance(logits_to_keep, int) and logits_to_keep > 0:
            h = h[:, -logits_to_keep:, :]
        logits = self.lm_head(h)
        loss = None
        if labels is not None:
            eff, count = ops.shift_labels(labels, None)
            loss = ops.cross_entropy(logits.reshape(-1, logits.shape[-1]), eff.reshape(-1), count)
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=out.past_key_values,
                                      hidden_states=out.hidden_states, attentions=None)

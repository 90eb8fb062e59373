"""Python handle of the native decode step (mantis_b200/csrc/decode_engine.cu): builds the weight / cache pointer tables
once and then issues ONE C call per generated token."""
import ctypes

import torch

from .. import _lib, ops
from .kv_cache import B200KVCache


native_steps = 0          # process-wide count of tokens decoded through the C++ engine (tests / bench read it)


class DecodeEngine:
    def __init__(self, decoder, lm_head_weight, cache: B200KVCache, reserve_tokens=1024):
        cfg = decoder.config
        layer0 = decoder.layers[0].self_attn
        self.L = len(decoder.layers)
        self.D = cfg.hidden_size
        self.H, self.Hkv, self.hd = layer0.num_heads, layer0.num_kv_heads, layer0.head_dim
        self.I = cfg.intermediate_size
        self.V = lm_head_weight.shape[0]
        self.eps = float(cfg.rms_norm_eps)
        self.decoder, self.lm_head_weight, self.cache = decoder, lm_head_weight, cache
        self.device = lm_head_weight.device
        self.inv_freq, self.rope_scale = decoder.rope_tables(self.device)
        self.B = cache.batch
        cache.reserve(cache.get_seq_length() + reserve_tokens)
        self._build_tables()
        self.ld_logits = (self.V + 7) // 8 * 8
        self.logits = torch.empty((self.B, self.ld_logits), dtype=torch.bfloat16, device=self.device)
        self.next_ids = torch.empty((self.B,), dtype=torch.int64, device=self.device)
        ws_bytes = _lib.lib().mb200_decode_ws_bytes(self.B, self.D, self.H, self.Hkv, self.hd, self.I, 0)
        self.ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=self.device)

    @staticmethod
    def eligible(decoder, cache, x_dtype):
        if not isinstance(cache, B200KVCache) or len(cache) != len(decoder.layers) or ops.FORCE_GENERIC:
            return False
        a = decoder.layers[0].self_attn
        if x_dtype != torch.bfloat16 or cache.dtype != torch.bfloat16 or a.head_dim != 128 or not 0 < cache.batch <= 16:
            return False
        for layer in decoder.layers:
            a, m = layer.self_attn, layer.mlp
            # every projection the engine reads raw weights of (_build_tables): a peft wrapper or a bias on ANY of them
            # (e.g. LoRA on k_proj/v_proj or down_proj only) sends the step down the Python path, which honours it
            for lin in (a.q_proj, a.k_proj, a.v_proj, a.o_proj, m.gate_proj, m.up_proj, m.down_proj):
                if type(lin).__name__ != "B200Linear" or lin.bias is not None:
                    return False
        return True

    def _build_tables(self):
        ptrs = []
        for i, layer in enumerate(self.decoder.layers):
            a, m = layer.self_attn, layer.mlp
            ptrs += [a.q_proj.weight, a.k_proj.weight, a.v_proj.weight, a.o_proj.weight, m.gate_proj.weight,
                     m.up_proj.weight, m.down_proj.weight, layer.input_layernorm.weight,
                     layer.post_attention_layernorm.weight, None, None]      # KV: paged, reached through the block table
        self._keep = ptrs
        self.layer_tab = (ctypes.c_void_p * len(ptrs))(*[t.data_ptr() if t is not None else None for t in ptrs])

    def step(self, ids, pos, kbits=None, next_out=None):
        """ids, pos: int64 [B] on device.  Appends one token per sequence to the cache, returns (logits [B,V], next_ids).
        The returned tensors are VIEWS of the engine's persistent buffers: the next step() overwrites them (callers that keep
        logits across steps must clone).  `next_out`: optional contiguous int64 [B] destination of the arg-max tokens (a row of
        a time-major token buffer, so a whole generation needs no copy kernels)."""
        ctx = self.cache.get_seq_length()
        self.cache.ensure(ctx + 1)                    # new pages only extend the block table; nothing is copied
        table = self.cache.device_table()
        dims = (ctypes.c_int * 12)(self.L, self.D, self.H, self.Hkv, self.hd, self.I, self.V, self.B, ctx,
                                   table.shape[1] * self.cache.PAGE, kbits.shape[1] if kbits is not None else 0,
                                   table.shape[1])
        fparm = (ctypes.c_float * 2)(self.eps, float(self.rope_scale))
        ids = ids.contiguous(); pos = pos.contiguous()
        misc = (ctypes.c_void_p * 10)(self.decoder.embed_tokens.weight.data_ptr(), self.decoder.norm.weight.data_ptr(),
                                     self.lm_head_weight.data_ptr(), self.inv_freq.data_ptr(), ids.data_ptr(),
                                     pos.data_ptr(), kbits.data_ptr() if kbits is not None else None,
                                     self.logits.data_ptr(),
                                     (next_out if next_out is not None else self.next_ids).data_ptr(), table.data_ptr())
        ops._call("mb200_llama_decode_step", dims, fparm, self.layer_tab, misc, ops._p(self.ws), self.ld_logits, ops._st())
        self.cache.advance(1)
        return self.logits[:, : self.V], (next_out if next_out is not None else self.next_ids)


def greedy_decode_loop(decoder, lm_head, cache, first_tokens, pos0, n_steps, kbits=None, eos_ids=None, check_every=32):
    """`n_steps` greedy decode steps after a prefill, ONE C call per token and nothing else: token t+1 is written by the engine's
    arg-max kernel straight into row t+1 of a time-major buffer, which is also the next step's input; positions are rows of a
    precomputed table.  Without EOS ids the host never synchronises; with them it looks every `check_every` steps.
    first_tokens int64 [B] (arg-max of the prefill), pos0 int64 [B] (position id of first_tokens).  Returns [B, <= n_steps + 1]
    tokens (first_tokens included; trailing steps after every sequence has finished are cut by the caller)."""
    global native_steps
    dev = first_tokens.device
    B = first_tokens.shape[0]
    eng = getattr(cache, "_engine", None)
    if eng is None or eng.decoder is not decoder or eng.B != cache.batch or eng.lm_head_weight is not lm_head.weight:
        eng = DecodeEngine(decoder, lm_head.weight, cache, reserve_tokens=n_steps + 8)
        cache._engine = eng
    else:
        cache.reserve(cache.get_seq_length() + n_steps + 8)
    toks = torch.empty((n_steps + 1, B), dtype=torch.int64, device=dev)
    toks[0].copy_(first_tokens)
    pos = (pos0.to(torch.int64).unsqueeze(0) + torch.arange(n_steps + 1, device=dev, dtype=torch.int64).unsqueeze(1)).contiguous()
    eos = torch.tensor(sorted(eos_ids), device=dev) if eos_ids else None
    done_at = n_steps
    for t in range(n_steps):
        eng.step(toks[t], pos[t], kbits, next_out=toks[t + 1])
        native_steps += 1
        if eos is not None and (t + 1) % check_every == 0 and bool(torch.isin(toks[: t + 2], eos).any(dim=0).all()):
            done_at = t + 1
            break
    return toks[: done_at + 1].t()


def native_decode_logits(decoder, lm_head, cache, input_ids, x_dtype, attention_mask, position_ids, labels=None,
                         output_hidden_states=False):
    """Single-token decode of any model whose text stack is a B200DecoderModel + a bias-free LM head (LLaVA, LLaVA-NeXT,
    Idefics2, Idefics3): one C call per token through the engine cached on the KV cache.  Returns logits [B, V] (bf16), or
    None when the step is not eligible (then the caller takes the Python path)."""
    if (input_ids is None or input_ids.shape[1] != 1 or labels is not None or output_hidden_states
            or torch.is_grad_enabled() or not isinstance(cache, B200KVCache) or cache.get_seq_length() == 0):
        return None
    if type(lm_head).__name__ != "B200Linear" or lm_head.bias is not None:
        return None
    if not DecodeEngine.eligible(decoder, cache, x_dtype):
        return None
    eng = getattr(cache, "_engine", None)
    if eng is None or eng.decoder is not decoder or eng.B != cache.batch or eng.lm_head_weight is not lm_head.weight:
        eng = DecodeEngine(decoder, lm_head.weight, cache)
        cache._engine = eng
    ctx = cache.get_seq_length()
    kbits = None
    if attention_mask is not None and (attention_mask.dim() != 2 or attention_mask.shape[1] != ctx + 1):
        return None
    window = getattr(decoder.config, "sliding_window", None)
    if window and ctx + 1 > window:                       # Mistral: keys that slid out of the window are masked
        from .llama import window_key_mask
        attention_mask = window_key_mask(attention_mask, ctx + 1, window, input_ids.shape[0], input_ids.device)
    if attention_mask is not None:
        kbits = ops.kmask_bits(attention_mask)
    if position_ids is None:
        position_ids = torch.full((input_ids.shape[0], 1), ctx, dtype=torch.int64, device=input_ids.device)
    logits, _ = eng.step(input_ids[:, 0], position_ids[:, -1].to(torch.int64), kbits)
    global native_steps
    native_steps += 1
    return logits

"""Idefics2ForConditionalGeneration (Mantis-8B-Idefics2) on the mantis_b200 CUDA kernels.

Drop-in for mantis.models.idefics2.modeling_idefics2 (reference file cited as `ref:`):
  Idefics2VisionEmbeddings (ref:155-210)   im2col + tcgen05 GEMM; NaViT fractional-bucket position ids computed once per
                                           distinct patch grid with the reference's exact fp32 arange/bucketize recipe
                                           (one host read of the tiny patch mask instead of one `.cpu()` per image)
  Idefics2VisionTransformer (ref:214-765)  shared pre-LN encoder blocks (bidirectional attention + key padding mask)
  Idefics2Connector (ref:1320-1334)        SwiGLU modality projection + perceiver resampler (ref:812-910,1187-1317)
  inputs_merger (ref:1545-1565)            row-copy kernel driven by an index map (same kernel as the LLaVA merge)
  Idefics2ForConditionalGeneration         Mistral decoder on our kernels, fp32 logits, CE with ignore_index=image_token_id
  (ref:1729-2000)                          (fused chunked LM-head+CE in training), image_hidden_states cached for generate()
State-dict keys are identical to the reference's (`model.vision_model.*`, `model.connector.*`, `model.text_model.*`,
`lm_head.weight`).
"""
from dataclasses import dataclass
from typing import List, Optional, Tuple, Union

import torch
from torch import nn
from transformers import PreTrainedModel
from transformers.generation import GenerationMixin
from transformers.modeling_outputs import BaseModelOutput, ModelOutput, SequenceClassifierOutputWithPast
from transformers.models.idefics2.configuration_idefics2 import Idefics2Config, Idefics2VisionConfig

from ... import ops
from ..kv_cache import B200KVCache
from ..layers import B200LayerNorm, B200Linear, B200RMSNorm, hf_key_remap_disabled, init_module_weights
from ..llama import B200DecoderModel
from ..vision import B200VisionEncoder, PatchEmbedGemm


@dataclass
class Idefics2BaseModelOutputWithPast(ModelOutput):
    last_hidden_state: torch.FloatTensor = None
    past_key_values: Optional[Tuple[Tuple[torch.FloatTensor]]] = None
    hidden_states: Optional[Tuple[torch.FloatTensor]] = None
    attentions: Optional[Tuple[torch.FloatTensor]] = None
    image_hidden_states: Optional[Tuple[torch.FloatTensor]] = None


@dataclass
class Idefics2CausalLMOutputWithPast(ModelOutput):
    loss: Optional[torch.FloatTensor] = None
    logits: torch.FloatTensor = None
    past_key_values: Optional[List[torch.FloatTensor]] = None
    hidden_states: Optional[Tuple[torch.FloatTensor]] = None
    attentions: Optional[Tuple[torch.FloatTensor]] = None
    image_hidden_states: Optional[Tuple[torch.FloatTensor]] = None


def navit_position_ids(patch_attention_mask_cpu: torch.Tensor, num_patches_per_side: int) -> torch.Tensor:
    """Exact restatement of ref:191-206 on the host (fp32 arange + bucketize(right=True)); the per-(h,w) result is
    cached, so a batch of equal-sized images costs one evaluation."""
    N, gh, gw = patch_attention_mask_cpu.shape
    boundaries = torch.arange(1 / num_patches_per_side, 1.0, 1 / num_patches_per_side)
    position_ids = torch.full(size=(N, gh * gw), fill_value=0)
    cache = {}
    for i in range(N):
        m = patch_attention_mask_cpu[i]
        nb_h = m[:, 0].sum(); nb_w = m[0].sum()
        key = (int(nb_h), int(nb_w))
        if key not in cache:
            fh = torch.arange(0, 1 - 1e-6, 1 / nb_h)
            fw = torch.arange(0, 1 - 1e-6, 1 / nb_w)
            bh = torch.bucketize(fh, boundaries, right=True)
            bw = torch.bucketize(fw, boundaries, right=True)
            cache[key] = (bh[:, None] * num_patches_per_side + bw).flatten()
        position_ids[i][m.view(-1)] = cache[key]
    return position_ids


class Idefics2VisionEmbeddings(nn.Module):
    def __init__(self, config: Idefics2VisionConfig):
        super().__init__()
        self.embed_dim = config.hidden_size
        self.image_size = config.image_size
        self.patch_size = config.patch_size
        self.patch_embedding = nn.Conv2d(config.num_channels, self.embed_dim, kernel_size=self.patch_size,
                                         stride=self.patch_size, padding="valid")
        self.num_patches_per_side = self.image_size // self.patch_size
        self.num_patches = self.num_patches_per_side ** 2
        self.num_positions = self.num_patches
        self.position_embedding = nn.Embedding(self.num_positions, self.embed_dim)
        self._pe = PatchEmbedGemm()
        self._full_ids = {}

    def forward(self, pixel_values, patch_attention_mask, mask_is_full):
        N, _, H, W = pixel_values.shape
        gh, gw = H // self.patch_size, W // self.patch_size
        x = self._pe(pixel_values, self.patch_embedding, self.patch_size)                   # [N*gh*gw, d]
        if mask_is_full:
            key = (gh, gw, str(pixel_values.device))
            if key not in self._full_ids:
                ids = navit_position_ids(torch.ones(1, gh, gw, dtype=torch.bool), self.num_patches_per_side)
                self._full_ids[key] = ids.to(pixel_values.device)
            pos = self._full_ids[key].expand(N, gh * gw)
        else:
            pos = navit_position_ids(patch_attention_mask.cpu(), self.num_patches_per_side).to(pixel_values.device)
        x = ops.add_rows(x, self.position_embedding.weight, idx=pos.reshape(-1))
        return x.view(N, gh * gw, self.embed_dim)


class Idefics2VisionTransformer(nn.Module):
    def __init__(self, config: Idefics2VisionConfig):
        super().__init__()
        self.config = config
        self.embeddings = Idefics2VisionEmbeddings(config)
        self.encoder = B200VisionEncoder(config)
        self.post_layernorm = B200LayerNorm(config.hidden_size, eps=config.layer_norm_eps)

    def get_input_embeddings(self):
        return self.embeddings

    def forward(self, pixel_values, patch_attention_mask=None, mask_is_full=None, **kw):
        N = pixel_values.size(0)
        if patch_attention_mask is None:
            mask_is_full = True
        elif mask_is_full is None:
            mask_is_full = not bool(torch.any(~patch_attention_mask))
        x = self.embeddings(pixel_values, patch_attention_mask, mask_is_full)
        key_mask = None if mask_is_full else patch_attention_mask.view(N, -1).to(torch.int64)
        x, _ = self.encoder(x, key_mask=key_mask)
        return BaseModelOutput(last_hidden_state=self.post_layernorm(x))


class Idefics2MLP(nn.Module):
    def __init__(self, hidden_size, intermediate_size, output_size, hidden_act):
        super().__init__()
        if hidden_act != "silu":
            raise ValueError("mantis_b200 Idefics2MLP implements the SwiGLU (silu) form used by Idefics2-8B")
        self.gate_proj = B200Linear(hidden_size, intermediate_size, bias=False)
        self.up_proj = B200Linear(hidden_size, intermediate_size, bias=False)
        self.down_proj = B200Linear(intermediate_size, output_size, bias=False)

    def forward(self, x, residual=None):
        return self.down_proj(ops.swiglu(self.gate_proj(x), self.up_proj(x)), residual=residual)


class Idefics2PerceiverAttention(nn.Module):
    def __init__(self, config, layer_idx=None):
        super().__init__()
        self.hidden_size = config.text_config.hidden_size
        self.num_heads = config.perceiver_config.resampler_n_heads
        self.head_dim = config.perceiver_config.resampler_head_dim
        self.num_key_value_heads = config.perceiver_config.num_key_value_heads
        self.q_proj = B200Linear(self.hidden_size, self.num_heads * self.head_dim, bias=False)
        self.k_proj = B200Linear(self.hidden_size, self.num_key_value_heads * self.head_dim, bias=False)
        self.v_proj = B200Linear(self.hidden_size, self.num_key_value_heads * self.head_dim, bias=False)
        self.o_proj = B200Linear(self.num_heads * self.head_dim, self.hidden_size, bias=False)

    def forward(self, latents, context, key_mask, residual):
        B, Lq, _ = latents.shape
        hs = torch.cat([context, latents], dim=-2)                                      # ref:857 (layout plumbing)
        Lk = hs.shape[1]
        q = self.q_proj(latents).view(B, Lq, self.num_heads, self.head_dim)
        k = self.k_proj(hs).view(B, Lk, self.num_key_value_heads, self.head_dim)
        v = self.v_proj(hs).view(B, Lk, self.num_key_value_heads, self.head_dim)
        o = ops.attention(q, k, v, causal=False, kmask=key_mask, scale=self.head_dim ** -0.5)
        return self.o_proj(o.view(B, Lq, self.num_heads * self.head_dim), residual=residual)


class Idefics2PerceiverLayer(nn.Module):
    def __init__(self, config, layer_idx):
        super().__init__()
        hidden = config.text_config.hidden_size
        eps = config.text_config.rms_norm_eps
        self.input_latents_norm = B200RMSNorm(hidden, eps=eps)
        self.input_context_norm = B200RMSNorm(hidden, eps=eps)
        self.self_attn = Idefics2PerceiverAttention(config, layer_idx)
        self.post_attention_layernorm = B200RMSNorm(hidden, eps=eps)
        self.mlp = Idefics2MLP(hidden, hidden * 4, hidden, config.perceiver_config.hidden_act)

    def forward(self, latents, context, key_mask):
        latents = self.self_attn(self.input_latents_norm(latents), self.input_context_norm(context), key_mask, latents)
        return self.mlp(self.post_attention_layernorm(latents), residual=latents)


class Idefics2PerceiverResampler(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.hidden_size = config.text_config.hidden_size
        self.n_latents = config.perceiver_config.resampler_n_latents
        self.depth = config.perceiver_config.resampler_depth
        self.latents = nn.Parameter(torch.ones(self.n_latents, self.hidden_size))
        self.layers = nn.ModuleList([Idefics2PerceiverLayer(config, i) for i in range(self.depth)])
        self.norm = B200RMSNorm(self.hidden_size, eps=config.text_config.rms_norm_eps)

    def forward(self, context, attention_mask):
        B = context.shape[0]
        latents = self.latents.unsqueeze(0).expand(B, *self.latents.shape).contiguous()
        key_mask = None
        if attention_mask is not None:
            ones = torch.ones((B, self.n_latents), dtype=torch.int64, device=context.device)
            key_mask = torch.cat([attention_mask.to(torch.int64), ones], dim=-1)
        x = latents
        for layer in self.layers:
            x = layer(x, context, key_mask)
        return self.norm(x)


class Idefics2Connector(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.modality_projection = Idefics2MLP(config.vision_config.hidden_size, config.text_config.intermediate_size,
                                               config.text_config.hidden_size, config.text_config.hidden_act)
        self.perceiver_resampler = Idefics2PerceiverResampler(config)

    def forward(self, image_hidden_states, attention_mask):
        return self.perceiver_resampler(self.modality_projection(image_hidden_states), attention_mask)


class Idefics2PreTrainedModel(PreTrainedModel):
    config_class = Idefics2Config
    base_model_prefix = "model"
    supports_gradient_checkpointing = True
    _no_split_modules = ["B200VisionEncoderLayer", "B200DecoderLayer", "Idefics2PerceiverLayer"]
    _skip_keys_device_placement = "past_key_values"
    _supports_flash_attn_2 = True
    _supports_flash_attn = True
    _supports_sdpa = True

    def _init_weights(self, module):
        std = getattr(self.config, "initializer_range", None) or getattr(self.config.text_config, "initializer_range", 0.02)
        init_module_weights(module, std)

    @classmethod
    def from_pretrained(cls, *args, **kwargs):
        with hf_key_remap_disabled("idefics2", "mistral"):
            return super().from_pretrained(*args, **kwargs)

    def save_pretrained(self, *args, **kwargs):
        with hf_key_remap_disabled("idefics2", "mistral"):
            return super().save_pretrained(*args, **kwargs)


class Idefics2Model(Idefics2PreTrainedModel):
    def __init__(self, config: Idefics2Config):
        super().__init__(config)
        self.padding_idx = self.config.text_config.pad_token_id
        self.vocab_size = self.config.text_config.vocab_size
        self.vision_model = Idefics2VisionTransformer(config.vision_config)
        self.connector = Idefics2Connector(config)
        self.text_model = B200DecoderModel(config.text_config)
        self.image_seq_len = config.perceiver_config.resampler_n_latents
        self.image_token_id = self.config.image_token_id
        self.post_init()

    def enable_input_require_grads(self):
        def make_inputs_require_grads(module, input, output):
            output.requires_grad_(True)
        self._text_require_grads_hook = self.get_input_embeddings().register_forward_hook(make_inputs_require_grads)
        self._vision_require_grads_hook = self.vision_model.embeddings.register_forward_hook(make_inputs_require_grads)

    def get_input_embeddings(self):
        return self.text_model.get_input_embeddings()

    def set_input_embeddings(self, value):
        self.text_model.set_input_embeddings(value)

    def inputs_merger(self, input_ids, inputs_embeds, image_hidden_states):
        """new_embeds[input_ids == image_token_id] = image rows in order (ref:1545-1565); sequence length unchanged."""
        B, T, D = inputs_embeds.shape
        mask = (input_ids == self.image_token_id).reshape(-1)
        rank = torch.cumsum(mask.to(torch.int64), 0) - 1
        t_idx = torch.arange(T, device=input_ids.device, dtype=torch.int64).repeat(B)
        srcmap = torch.where(mask, -(rank) - 2, t_idx).to(torch.int32).view(B, T)
        n_img_rows = image_hidden_states.shape[0] * image_hidden_states.shape[1]
        n_tok = int(mask.sum().item())
        if n_tok != n_img_rows:
            raise RuntimeError(f"shape mismatch: {n_tok} image tokens cannot take {n_img_rows} image hidden states")
        return ops._MergeRowsFn.apply(inputs_embeds, image_hidden_states, srcmap, T)

    def encode_images(self, pixel_values, pixel_attention_mask):
        batch_size, num_images, C, H, W = pixel_values.shape
        pixel_values = pixel_values.to(dtype=self.dtype).view(batch_size * num_images, C, H, W)
        # drop all-zero padding images (ref:1637-1639): row-wise zero scan on device, one tiny host read
        flags = ops.rows_all_zero(pixel_values.reshape(batch_size * num_images, -1))
        real = (flags == 0)
        all_real = bool(real.all().item())
        if not all_real:
            pixel_values = pixel_values[real].contiguous()
        mask_is_full = pixel_attention_mask is None
        if pixel_attention_mask is None:
            patch_attention_mask = None
        else:
            pam = pixel_attention_mask.view(batch_size * num_images, H, W)
            if not all_real:
                pam = pam[real].contiguous()
            p = self.config.vision_config.patch_size
            sub = pam.unfold(1, p, p).unfold(2, p, p)
            patch_attention_mask = (sub.sum(dim=(-1, -2)) > 0).bool()                        # ref:1655-1658
            mask_is_full = not bool(torch.any(~patch_attention_mask))
        x = self.vision_model(pixel_values, patch_attention_mask, mask_is_full).last_hidden_state
        am = None if mask_is_full else patch_attention_mask.view(pixel_values.size(0), -1)
        return self._connect(x, am)

    def _connect(self, x, patch_key_mask):
        return self.connector(x, patch_key_mask)

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                inputs_embeds=None, pixel_values=None, pixel_attention_mask=None, image_hidden_states=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None, **kw):
        if input_ids is None and inputs_embeds is None:
            raise ValueError("You have to specify either input_ids or inputs_embeds")
        past_seen = past_key_values.get_seq_length() if (past_key_values is not None and hasattr(past_key_values, "get_seq_length")) else 0
        if inputs_embeds is not None and input_ids is None and past_seen == 0:
            raise ValueError("When first calling the model, if input_embeds are passed, input_ids should not be None.")
        if inputs_embeds is None:
            inputs_embeds = self.text_model.get_input_embeddings()(input_ids)
        if pixel_values is not None and image_hidden_states is not None:
            raise ValueError("You cannot specify both pixel_values and image_hidden_states at the same time")
        elif pixel_values is not None:
            image_hidden_states = self.encode_images(pixel_values, pixel_attention_mask)
        elif image_hidden_states is not None:
            image_hidden_states = image_hidden_states.to(dtype=self.dtype, device=inputs_embeds.device)
        if past_seen == 0 and image_hidden_states is not None:
            inputs_embeds = self.inputs_merger(input_ids, inputs_embeds, image_hidden_states)
        out = self.text_model(inputs_embeds=inputs_embeds, attention_mask=attention_mask, position_ids=position_ids,
                              past_key_values=past_key_values, use_cache=use_cache,
                              output_hidden_states=output_hidden_states, cu_segments=kw.get("cu_segments"))
        return Idefics2BaseModelOutputWithPast(last_hidden_state=out.last_hidden_state, past_key_values=out.past_key_values,
                                               hidden_states=out.hidden_states, attentions=None,
                                               image_hidden_states=image_hidden_states)


class Idefics2ForConditionalGeneration(Idefics2PreTrainedModel, GenerationMixin):
    _tied_weights_keys = {}
    _backbone_cls = None          # set below (Idefics2Model); the Idefics3 shell swaps in its own backbone

    def _loss_ignore_index(self):
        return self.image_token_id                               # CrossEntropyLoss(ignore_index=image_token_id), ref:1894

    def __init__(self, config):
        super().__init__(config)
        self.model = (self._backbone_cls or Idefics2Model)(config)
        self.image_token_id = self.config.image_token_id
        self.lm_head = B200Linear(config.text_config.hidden_size, config.text_config.vocab_size, bias=False)
        self.vocab_size = config.text_config.vocab_size
        self.materialize_logits_in_training = False
        self.post_init()

    def enable_input_require_grads(self):
        self.model.enable_input_require_grads()

    def get_input_embeddings(self):
        return self.model.text_model.get_input_embeddings()

    def set_input_embeddings(self, value):
        self.model.text_model.set_input_embeddings(value)

    def get_output_embeddings(self):
        return self.lm_head

    def set_output_embeddings(self, new_embeddings):
        self.lm_head = new_embeddings

    def tie_weights(self, *args, **kwargs):
        if getattr(self.config, "tie_word_embeddings", False):
            self.lm_head.weight = self.get_input_embeddings().weight

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                inputs_embeds=None, pixel_values=None, pixel_attention_mask=None, image_hidden_states=None,
                labels=None, use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None,
                logits_to_keep=0, **kw):
        if (input_ids is not None and input_ids.shape[1] == 1 and inputs_embeds is None and labels is None
                and isinstance(past_key_values, B200KVCache) and not output_hidden_states):
            # decode step: the image states were merged at prefill (ref:1961-1990 caches them), the text stack runs through
            # the native engine -- one C call per generated token instead of ~420 Python launches
            from ..decode_engine import native_decode_logits
            lg = native_decode_logits(self.model.text_model, self.lm_head, past_key_values, input_ids,
                                      self.model.text_model.embed_tokens.weight.dtype, attention_mask, position_ids)
            if lg is not None:
                return Idefics2CausalLMOutputWithPast(loss=None, logits=lg.float().unsqueeze(1),
                                                      past_key_values=past_key_values, hidden_states=None, attentions=None,
                                                      image_hidden_states=image_hidden_states)
        outputs = self.model(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                             past_key_values=past_key_values, inputs_embeds=inputs_embeds, pixel_values=pixel_values,
                             pixel_attention_mask=pixel_attention_mask, image_hidden_states=image_hidden_states,
                             use_cache=use_cache, output_hidden_states=output_hidden_states,
                             cu_segments=kw.get("cu_segments"))
        hidden = outputs.last_hidden_state
        loss = None
        logits = None
        if attention_mask is not None and attention_mask.dim() == 4:       # packed batch: loss mask = key validity
            attention_mask = (attention_mask != 0).any(dim=1).any(dim=1).to(torch.int64)
        want_logits = (not (self.training and torch.is_grad_enabled())) or self.materialize_logits_in_training \
            or labels is None
        if labels is not None:
            eff, count = ops.shift_labels(labels, attention_mask, self._loss_ignore_index())         # ref:1887-1899
        if want_logits:
            h = hidden[:, -logits_to_keep:, :] if (isinstance(logits_to_keep, int) and logits_to_keep > 0) else hidden
            lg = self.lm_head(h)
            if labels is not None and not (isinstance(logits_to_keep, int) and logits_to_keep > 0):
                loss = ops.cross_entropy(lg.reshape(-1, lg.shape[-1]), eff.reshape(-1), count)
            logits = lg.float()                                                                      # ref:1884
        elif labels is not None:
            hint = kw.get("merge_hint")                                  # {"valid_rows": n} from the collator: no read-back
            loss = ops.lm_head_ce(hidden, self.lm_head.weight, eff, count,
                                  valid_rows_hint=hint.get("valid_rows") if hint else None)
        return Idefics2CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=outputs.past_key_values,
                                              hidden_states=outputs.hidden_states, attentions=None,
                                              image_hidden_states=outputs.image_hidden_states)

    def prepare_inputs_for_generation(self, input_ids, past_key_values=None, attention_mask=None, inputs_embeds=None,
                                      **kwargs):
        has_past = past_key_values is not None and past_key_values.get_seq_length() > 0
        if has_past:
            past_length = past_key_values.get_seq_length()
            if attention_mask is not None and attention_mask.shape[1] > input_ids.shape[1]:
                input_ids = input_ids[:, -(attention_mask.shape[1] - past_length):]
            elif past_length < input_ids.shape[1]:
                input_ids = input_ids[:, past_length:]
        elif past_key_values is None or not isinstance(past_key_values, B200KVCache):
            past_key_values = B200KVCache()
        position_ids = kwargs.get("position_ids", None)
        if position_ids is not None and position_ids.shape[-1] > input_ids.shape[1]:
            # transformers >= 5 may hand over position ids for the WHOLE sequence; the model wants those of the new tokens
            position_ids = position_ids[..., -input_ids.shape[1]:]
        if attention_mask is not None and position_ids is None:
            position_ids = attention_mask.long().cumsum(-1) - 1
            position_ids.masked_fill_(attention_mask == 0, 1)
            if has_past:
                position_ids = position_ids[:, -input_ids.shape[1]:]
        if inputs_embeds is not None and not has_past:
            model_inputs = {"inputs_embeds": inputs_embeds}
        else:
            model_inputs = {"input_ids": input_ids}
        image_hidden_states = kwargs.get("image_hidden_states", None)
        if image_hidden_states is not None:
            pixel_values = None; pixel_attention_mask = None
        else:
            pixel_values = kwargs.get("pixel_values", None)
            pixel_attention_mask = kwargs.get("pixel_attention_mask", None)
        model_inputs.update({"position_ids": position_ids, "past_key_values": past_key_values,
                             "use_cache": kwargs.get("use_cache", True), "attention_mask": attention_mask,
                             "pixel_values": pixel_values, "pixel_attention_mask": pixel_attention_mask,
                             "image_hidden_states": image_hidden_states, "logits_to_keep": 1})
        return model_inputs

    def _update_model_kwargs_for_generation(self, outputs, model_kwargs, is_encoder_decoder=False, **kwargs):
        model_kwargs = super()._update_model_kwargs_for_generation(outputs=outputs, model_kwargs=model_kwargs,
                                                                   is_encoder_decoder=is_encoder_decoder, **kwargs)
        model_kwargs["image_hidden_states"] = outputs.image_hidden_states
        return model_kwargs

    def _reorder_cache(self, past_key_values, beam_idx):
        past_key_values.reorder_cache(beam_idx)
        return past_key_values


class Idefics2ForSequenceClassification(Idefics2PreTrainedModel):
    """Classification head variant (ref:2017-2310): last-non-pad-token pooling over the same backbone.
    Not on the north-star path; kept importable with the same state-dict keys (`model.*`, `score.weight`)."""

    def __init__(self, config):
        super().__init__(config)
        self.num_labels = config.num_labels
        self.model = Idefics2Model(config)
        self.score = B200Linear(config.text_config.hidden_size, self.num_labels, bias=False)
        self.post_init()

    def get_input_embeddings(self):
        return self.model.text_model.get_input_embeddings()

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                inputs_embeds=None, pixel_values=None, pixel_attention_mask=None, image_hidden_states=None,
                labels=None, use_cache=None, **kw):
        out = self.model(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                         past_key_values=past_key_values, inputs_embeds=inputs_embeds, pixel_values=pixel_values,
                         pixel_attention_mask=pixel_attention_mask, image_hidden_states=image_hidden_states,
                         use_cache=use_cache)
        logits = self.score(out.last_hidden_state)
        B = logits.shape[0]
        pad = self.config.text_config.pad_token_id
        if pad is None or input_ids is None:
            last = torch.full((B,), logits.shape[1] - 1, device=logits.device)
        else:
            last = (torch.eq(input_ids, pad).int().argmax(-1) - 1) % input_ids.shape[-1]
        pooled = logits[torch.arange(B, device=logits.device), last]
        loss = None
        if labels is not None:
            count = torch.tensor([float(B)], device=logits.device)
            loss = ops.cross_entropy(pooled.contiguous(), labels.view(-1).to(torch.int64), count)
        return SequenceClassifierOutputWithPast(loss=loss, logits=pooled, past_key_values=out.past_key_values)

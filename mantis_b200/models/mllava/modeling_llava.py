"""LlavaForConditionalGeneration / MLlavaForConditionalGeneration on the mantis_b200 CUDA kernels.

Drop-in for mantis.models.mllava.modeling_llava (reference: mantis/models/mllava/modeling_llava.py): same class
names, constructor (`config, vision_tower=None, language_model=None`), forward() signature, output dataclass,
state-dict keys and error behaviour.  What changes is what runs underneath:

  reference                                              here
  ---------------------------------------------------    --------------------------------------------------------
  AutoModel vision tower, all layers + head (:456)       B200 vision tower, only the layers feeding hidden_states[-2]
  projector = 2 cuBLAS GEMMs + ATen gelu (:110-118)      tcgen05 GEMMs (+bias epilogue), CUDA gelu
  _merge_... ~20 ATen launches, 4 host syncs (:293-360)  plan + index + row-copy kernels, 1 host sync
  LlamaForCausalLM via HF/ATen/flash-attn (:510)         mantis_b200.models.llama (CUDA kernels, paged-style KV cache)
  full [B,S,V] logits + boolean gather + CE (:523-537)   fused chunked LM-head + CE (training); logits on demand
  decode-time K-cache scan + torch.where sync (:480-508) pad slots tracked on the host-free path (mask carried in cache)
"""
from dataclasses import dataclass
import os
from typing import List, Optional, Tuple, Union

import torch
from torch import nn
from transformers import PreTrainedModel
from transformers.generation import GenerationMixin
from transformers.modeling_outputs import ModelOutput

from ... import ops
from ..kv_cache import B200KVCache
from ..layers import B200Linear, hf_key_remap_disabled, init_module_weights
from ..llama import B200CausalLM
from ..vision import B200VisionEncoder, build_vision_tower
from .configuration_llava import LlavaConfig


@dataclass
class LlavaCausalLMOutputWithPast(ModelOutput):
    loss: Optional[torch.FloatTensor] = None
    logits: torch.FloatTensor = None
    past_key_values: Optional[List[torch.FloatTensor]] = None
    hidden_states: Optional[Tuple[torch.FloatTensor]] = None
    attentions: Optional[Tuple[torch.FloatTensor]] = None
    image_hidden_states: Optional[Tuple[torch.FloatTensor]] = None


class LlavaMultiModalProjector(nn.Module):
    """linear_2(act(linear_1(x)))   (reference :106-118)"""

    def __init__(self, config: LlavaConfig):
        super().__init__()
        self.linear_1 = B200Linear(config.vision_config.hidden_size, config.text_config.hidden_size, bias=True)
        self.act_name = config.projector_hidden_act
        self.linear_2 = B200Linear(config.text_config.hidden_size, config.text_config.hidden_size, bias=True)

    def forward(self, image_features):
        return self.linear_2(self.linear_1(image_features, act=self.act_name))


class LlavaPreTrainedModel(PreTrainedModel):
    config_class = LlavaConfig
    base_model_prefix = "model"
    supports_gradient_checkpointing = True
    _no_split_modules = ["B200VisionEncoderLayer", "B200DecoderLayer"]
    _skip_keys_device_placement = "past_key_values"
    _supports_flash_attn_2 = True
    _supports_flash_attn = True
    _supports_sdpa = True

    def _init_weights(self, module):
        # same rule as the reference (:153-170): normal(0, initializer_range) for Linear/Conv/Embedding
        std = getattr(self.config, "initializer_range", None)
        if std is None:
            std = getattr(self.config.text_config, "initializer_range", 0.02)
        init_module_weights(module, std)

    # transformers >= 5 would rename `language_model.model.*` keys for model_type "llava" (its own re-organised class);
    # ours keeps the reference layout, so checkpoints are read / written verbatim.
    @classmethod
    def from_pretrained(cls, *args, **kwargs):
        with hf_key_remap_disabled("llava"):
            return super().from_pretrained(*args, **kwargs)

    def save_pretrained(self, *args, **kwargs):
        with hf_key_remap_disabled("llava"):
            return super().save_pretrained(*args, **kwargs)


class LlavaForConditionalGeneration(LlavaPreTrainedModel, GenerationMixin):
    def __init__(self, config: LlavaConfig, vision_tower=None, language_model=None):
        super().__init__(config)
        self.vision_tower = build_vision_tower(config.vision_config) if vision_tower is None else vision_tower
        self.multi_modal_projector = LlavaMultiModalProjector(config)
        self.vocab_size = getattr(config, "vocab_size", None) or config.text_config.vocab_size
        self.language_model = B200CausalLM(config.text_config) if language_model is None else language_model
        self.pad_token_id = self.config.pad_token_id if self.config.pad_token_id is not None else -1
        # training: skip the [B,S,V] logits tensor unless the caller asks for it (set True for parity checks)
        self.materialize_logits_in_training = False
        self.post_init()

    # ---- embedding / decoder plumbing (reference :264-291) ----
    def get_input_embeddings(self):
        return self.language_model.get_input_embeddings()

    def set_input_embeddings(self, value):
        self.language_model.set_input_embeddings(value)

    def get_output_embeddings(self):
        return self.language_model.get_output_embeddings()

    def set_output_embeddings(self, new_embeddings):
        self.language_model.set_output_embeddings(new_embeddings)

    def set_decoder(self, decoder):
        self.language_model.set_decoder(decoder)

    def get_decoder(self):
        return self.language_model.get_decoder()

    def tie_weights(self, *args, **kwargs):
        return self.language_model.tie_weights(*args, **kwargs)

    def resize_token_embeddings(self, new_num_tokens: Optional[int] = None, pad_to_multiple_of=None, **kw) -> nn.Embedding:
        model_embeds = self.language_model.resize_token_embeddings(new_num_tokens, pad_to_multiple_of)
        self.config.text_config.vocab_size = model_embeds.num_embeddings
        self.config.vocab_size = model_embeds.num_embeddings
        self.vocab_size = model_embeds.num_embeddings
        return model_embeds

    # ---- the merge (reference :293-360) ----
    def _merge_input_ids_with_image_features(self, image_features, inputs_embeds, input_ids, attention_mask, labels,
                                             plan_hint=None):
        return ops.merge_input_ids_with_image_features(
            image_features, inputs_embeds, input_ids, attention_mask, labels,
            image_token_index=self.config.image_token_index, pad_token_id=self.pad_token_id,
            ignore_index=self.config.ignore_index, plan_hint=plan_hint)

    # ---- vision path ----
    def _select(self, feats, strategy):
        if strategy == "default":
            return feats[:, 1:]
        if strategy == "full":
            return feats
        raise ValueError(f"Unexpected select feature strategy: {self.config.vision_feature_select_strategy}")

    def _image_features(self, pixel_values, vision_feature_layer, vision_feature_select_strategy):
        tower_dtype = self.vision_tower.dtype
        if pixel_values.dtype != tower_dtype:
            pixel_values = pixel_values.type(tower_dtype)
        if hasattr(self.vision_tower, "features"):
            feats = self.vision_tower.features(pixel_values, vision_feature_layer)
        else:   # a user-supplied tower (3-arg constructor): generic HF protocol
            feats = self.vision_tower(pixel_values, output_hidden_states=True).hidden_states[vision_feature_layer]
        feats = self._select(feats, vision_feature_select_strategy)
        feats = self._post_select(feats)
        return self.multi_modal_projector(feats)

    def _post_select(self, feats):
        return feats

    def forward(
        self,
        input_ids: torch.LongTensor = None,
        pixel_values: torch.FloatTensor = None,
        attention_mask: Optional[torch.Tensor] = None,
        position_ids: Optional[torch.LongTensor] = None,
        past_key_values=None,
        inputs_embeds: Optional[torch.FloatTensor] = None,
        vision_feature_layer: Optional[int] = None,
        vision_feature_select_strategy: Optional[str] = None,
        labels: Optional[torch.LongTensor] = None,
        use_cache: Optional[bool] = None,
        output_attentions: Optional[bool] = None,
        output_hidden_states: Optional[bool] = None,
        return_dict: Optional[bool] = None,
        logits_to_keep: int = 0,
        **kwargs,
    ) -> Union[Tuple, LlavaCausalLMOutputWithPast]:
        output_hidden_states = (output_hidden_states if output_hidden_states is not None
                                else getattr(self.config, "output_hidden_states", False))
        return_dict = return_dict if return_dict is not None else getattr(self.config, "return_dict", True)
        vision_feature_layer = (vision_feature_layer if vision_feature_layer is not None
                                else self.config.vision_feature_layer)
        vision_feature_select_strategy = (vision_feature_select_strategy if vision_feature_select_strategy is not None
                                          else self.config.vision_feature_select_strategy)
        merged = False
        if inputs_embeds is None:
            inputs_embeds = self.get_input_embeddings()(input_ids)                                    # :427
            if pixel_values is not None and input_ids.shape[1] != 1:
                if isinstance(pixel_values, list):
                    pixel_values = torch.cat([x for x in pixel_values if x is not None], dim=0)       # :431-432
                image_features = self._image_features(pixel_values, vision_feature_layer,
                                                      vision_feature_select_strategy)
                inputs_embeds, attention_mask, labels, position_ids = self._merge_input_ids_with_image_features(
                    image_features, inputs_embeds, input_ids, attention_mask, labels,
                    plan_hint=kwargs.get("merge_hint"))               # optional: sync-free merge (train.data.Collator)
                merged = True
                if isinstance(past_key_values, B200KVCache):
                    past_key_values.prefill_mask = attention_mask
            elif (past_key_values is not None and pixel_values is not None and input_ids.shape[1] == 1
                  and attention_mask is not None):
                # decode step (reference :477-508): the reference rediscovers padded cache slots by scanning layer-0
                # keys for exact zeros (+ a host sync per token); the slots are exactly the zero-mask positions of
                # the merged prefill mask, which we carry with the cache.
                pm = getattr(past_key_values, "prefill_mask", None)
                ctx = past_key_values.get_seq_length()
                if pm is not None:
                    ext = torch.ones((attention_mask.shape[0], ctx + 1 - pm.shape[1]), dtype=pm.dtype, device=pm.device)
                    attention_mask = torch.cat((pm, ext), dim=1)
                else:
                    target = ctx + 1
                    attention_mask = torch.ones((attention_mask.shape[0], target), dtype=attention_mask.dtype,
                                                device=attention_mask.device)
                position_ids = torch.sum(attention_mask, dim=1).unsqueeze(-1) - 1                      # :508

        lm = self.language_model
        fast = self._native_decode(input_ids, inputs_embeds, attention_mask, position_ids, past_key_values, labels,
                                   output_hidden_states)
        if fast is not None:
            return fast
        out = lm.model(attention_mask=attention_mask, position_ids=position_ids, past_key_values=past_key_values,
                       inputs_embeds=inputs_embeds, use_cache=use_cache, output_hidden_states=output_hidden_states)
        hidden = out.last_hidden_state

        loss = None
        logits = None
        want_logits = (not (self.training and torch.is_grad_enabled())) or self.materialize_logits_in_training \
            or labels is None
        if labels is not None:
            eff, count = ops.shift_labels(labels, attention_mask, self.config.ignore_index)          # :526-531
        if want_logits:
            h = hidden[:, -logits_to_keep:, :] if (isinstance(logits_to_keep, int) and logits_to_keep > 0) else hidden
            logits = lm.lm_head(h)
            if labels is not None and not (isinstance(logits_to_keep, int) and logits_to_keep > 0):
                loss = ops.cross_entropy(logits.reshape(-1, logits.shape[-1]), eff.reshape(-1), count)
        elif labels is not None:
            hint = kwargs.get("merge_hint") if merged else None          # the collator's host-side count of supervised rows
            loss = ops.lm_head_ce(hidden, lm.lm_head.weight, eff, count,
                                  valid_rows_hint=hint.get("valid_rows") if hint else None)
        if labels is None and merged:
            # reference :475-476 substitutes all-ignore labels, so its loss is the mean over zero rows == NaN
            loss = torch.full((), float("nan"), dtype=torch.float32, device=hidden.device)

        if not return_dict:
            output = (logits,) + (out.past_key_values, out.hidden_states)
            return (loss,) + output if loss is not None else output
        return LlavaCausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=out.past_key_values,
                                           hidden_states=out.hidden_states, attentions=None)

    def _native_decode(self, input_ids, inputs_embeds, attention_mask, position_ids, cache, labels, output_hidden_states):
        """single-token decode through the C++ engine (one call per token); None if not applicable"""
        from ..decode_engine import native_decode_logits
        logits = native_decode_logits(self.language_model.model, self.language_model.lm_head, cache, input_ids,
                                      inputs_embeds.dtype, attention_mask, position_ids, labels, output_hidden_states)
        if logits is None:
            return None
        return LlavaCausalLMOutputWithPast(loss=None, logits=logits.unsqueeze(1), past_key_values=cache,
                                           hidden_states=None, attentions=None)

    # ---- generation glue (reference :551-605) ----
    def prepare_inputs_for_generation(self, input_ids, past_key_values=None, inputs_embeds=None, pixel_values=None,
                                      attention_mask=None, **kwargs):
        has_past = past_key_values is not None and past_key_values.get_seq_length() > 0
        if has_past:
            cache_length = past_length = past_key_values.get_seq_length()
            if attention_mask is not None and attention_mask.shape[1] > input_ids.shape[1]:
                input_ids = input_ids[:, -(attention_mask.shape[1] - past_length):]
            elif past_length < input_ids.shape[1]:
                input_ids = input_ids[:, past_length:]
            elif self.config.image_token_index in input_ids:
                input_ids = input_ids[:, input_ids.shape[1] - 1:]
            if cache_length < past_length and attention_mask is not None:
                attention_mask = attention_mask[:, -(cache_length + input_ids.shape[1]):]
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
        model_inputs.update({"position_ids": position_ids, "past_key_values": past_key_values,
                             "use_cache": kwargs.get("use_cache", True), "attention_mask": attention_mask,
                             "pixel_values": pixel_values, "logits_to_keep": 1})
        return model_inputs

    def _reorder_cache(self, past_key_values, beam_idx):
        past_key_values.reorder_cache(beam_idx)
        return past_key_values

    # ---- generate(): greedy decoding without per-token host round trips -------------------------------------------------
    _FAST_GENERATE_KEYS = {"input_ids", "pixel_values", "attention_mask", "max_new_tokens", "max_length", "min_new_tokens",
                           "do_sample", "num_beams", "eos_token_id", "pad_token_id", "use_cache", "num_return_sequences"}

    def _fast_greedy_plan(self, inputs, generation_config, kwargs):
        """-> dict of arguments for the sync-free greedy loop, or None when the call needs anything beyond plain greedy decoding
        (then transformers' GenerationMixin.generate runs as before)."""
        if os.environ.get("MB200_FAST_GENERATE", "1") != "1" or inputs is not None or generation_config is not None:
            return None
        if not set(kwargs) <= self._FAST_GENERATE_KEYS or kwargs.get("input_ids") is None or kwargs.get("attention_mask") is None:
            return None
        gc = self.generation_config
        ids = kwargs["input_ids"]
        if (kwargs.get("do_sample", gc.do_sample) or kwargs.get("num_beams", gc.num_beams) != 1
                or kwargs.get("num_return_sequences", 1) != 1 or kwargs.get("use_cache", True) is False
                or (gc.repetition_penalty or 1.0) != 1.0 or gc.no_repeat_ngram_size or gc.bad_words_ids
                or gc.forced_bos_token_id is not None or gc.forced_eos_token_id is not None or gc.suppress_tokens
                or gc.begin_suppress_tokens or gc.return_dict_in_generate or gc.output_scores or gc.output_logits
                or getattr(gc, "sequence_bias", None) or getattr(gc, "renormalize_logits", False)):
            return None
        if kwargs.get("max_new_tokens", gc.max_new_tokens) is not None:
            n_new = int(kwargs.get("max_new_tokens", gc.max_new_tokens))
        elif kwargs.get("max_length") is not None or gc.max_length is not None:
            n_new = int(kwargs.get("max_length") or gc.max_length) - ids.shape[1]
        else:
            return None
        eos = kwargs.get("eos_token_id", gc.eos_token_id)
        eos = [] if eos is None else ([int(e) for e in eos] if isinstance(eos, (list, tuple)) else [int(eos)])
        min_new = kwargs.get("min_new_tokens", gc.min_new_tokens) or 0
        if n_new < 1 or min_new > n_new:
            return None
        pad = kwargs.get("pad_token_id", gc.pad_token_id)
        if eos and pad is None:
            pad = eos[0]                                    # GenerationMixin: "Setting pad_token_id to eos_token_id"
        lm = self.language_model
        a = lm.model.layers[0].self_attn
        if (not ids.is_cuda or lm.lm_head.weight.dtype != torch.bfloat16 or a.head_dim != 128 or not 0 < ids.shape[0] <= 16
                or ops.FORCE_GENERIC):
            return None
        return {"n_new": n_new, "eos": eos, "pad": pad, "min_new": int(min_new)}

    @torch.no_grad()
    def generate(self, inputs=None, generation_config=None, **kwargs):
        """transformers' `generate()` (ref: the call mantis/models/mllava/utils.py:88 makes).  Plain greedy decoding -- what
        chat_mllava asks for by default -- takes a loop that issues exactly one C call per token and does not synchronise with
        the device until the end (or every 32 tokens when EOS ids are set): prefill through forward(), then
        decode_engine.greedy_decode_loop.  Output conventions are GenerationMixin's (prompt + new tokens, pad after a sequence's
        EOS, stop when every sequence has finished).  Every other configuration goes through GenerationMixin unchanged."""
        plan = self._fast_greedy_plan(inputs, generation_config, kwargs)
        if plan is None:
            return super().generate(inputs, generation_config=generation_config, **kwargs)
        from ..decode_engine import DecodeEngine, greedy_decode_loop
        ids, am = kwargs["input_ids"], kwargs["attention_mask"]
        cache = B200KVCache()
        out = self(input_ids=ids, pixel_values=kwargs.get("pixel_values"), attention_mask=am, past_key_values=cache,
                   use_cache=True, logits_to_keep=1)
        first = out.logits[:, -1, :].argmax(-1)
        lm = self.language_model
        if plan["n_new"] == 1 or not DecodeEngine.eligible(lm.model, cache, lm.lm_head.weight.dtype):
            if plan["n_new"] > 1:                             # e.g. a peft-wrapped projection: the per-step Python path
                return super().generate(inputs, generation_config=generation_config, **kwargs)
            new = first[:, None]
        else:
            pm = getattr(cache, "prefill_mask", None)         # merged prompt mask (zero = padded slot), ref :477-508
            S = cache.get_seq_length()
            n_steps = plan["n_new"] - 1
            kbits, pos0 = None, torch.full((ids.shape[0],), S, dtype=torch.int64, device=ids.device)
            if pm is not None and not bool((pm != 0).all()):
                full = torch.cat((pm, torch.ones((pm.shape[0], n_steps + 1), dtype=pm.dtype, device=pm.device)), dim=1)
                kbits = ops.kmask_bits(full)                  # one bitmask for the whole generation: step t reads its first S+t+1 bits
                pos0 = pm.sum(dim=1).to(torch.int64)          # position id of the first generated token (ref :508)
            new = greedy_decode_loop(lm.model, lm.lm_head, cache, first, pos0, n_steps, kbits, plan["eos"])
        if plan["eos"]:
            eos = torch.tensor(plan["eos"], device=new.device)
            hit = torch.isin(new, eos)
            if plan["min_new"] and bool(hit[:, : plan["min_new"]].any()):
                # GenerationMixin would have suppressed this EOS (MinNewTokensLengthLogitsProcessor) and continued differently
                return super().generate(inputs, generation_config=generation_config, **kwargs)
            after = (hit.cumsum(dim=1) - hit.long()) > 0       # strictly after a sequence's first EOS
            new = torch.where(after, torch.full_like(new, plan["pad"]), new)
            finished = hit.any(dim=1)
            if bool(finished.all()):
                last = int((hit.long().argmax(dim=1)).max())  # GenerationMixin stops right after the last sequence finishes
                new = new[:, : last + 1]
        return torch.cat((ids, new.to(ids.dtype)), dim=1)

    @torch.no_grad()
    def greedy_generate(self, input_ids, pixel_values=None, attention_mask=None, max_new_tokens=32, eos_token_id=None):
        """Minimal greedy loop over forward() + B200KVCache (what chat_mllava's generate(num_beams=1,
        do_sample=False) does), independent of GenerationMixin internals."""
        cache = B200KVCache()
        eos = set(eos_token_id if isinstance(eos_token_id, (list, tuple)) else ([eos_token_id] if eos_token_id is not None else []))
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        out = self(input_ids=input_ids, pixel_values=pixel_values, attention_mask=attention_mask,
                   past_key_values=cache, use_cache=True, logits_to_keep=1)
        tokens = [input_ids]
        finished = torch.zeros(input_ids.shape[0], dtype=torch.bool, device=input_ids.device)
        for step in range(max_new_tokens):
            nxt = out.logits[:, -1, :].argmax(-1)
            tokens.append(nxt[:, None])
            if eos:
                finished |= torch.isin(nxt, torch.tensor(sorted(eos), device=nxt.device))
                if bool(finished.all()):
                    break
            if step == max_new_tokens - 1:
                break
            attention_mask = torch.cat([attention_mask, torch.ones_like(nxt[:, None])], dim=1)
            out = self(input_ids=nxt[:, None], pixel_values=pixel_values, attention_mask=attention_mask,
                       past_key_values=cache, use_cache=True, logits_to_keep=1)
        return torch.cat(tokens, dim=1)


class MLlavaForConditionalGeneration(LlavaForConditionalGeneration):
    """Ablation variant (reference :615-792): adds image-index type embeddings and a CLIPEncoder stack
    (`vision_xatten_layers`) between feature selection and the projector."""

    def __init__(self, config: LlavaConfig):
        super().__init__(config)
        config.vision_config.type_vocab_size = 144
        self.image_type_embeddings = nn.Embedding(config.vision_config.type_vocab_size, config.vision_config.hidden_size)
        self.vision_xatten_layers = B200VisionEncoder(config.vision_config)
        self.post_init()

    def _post_select(self, feats):
        num_images, P, d = feats.shape
        idx = torch.arange(num_images, device=feats.device).repeat_interleave(P)
        feats = ops.add_rows(feats.reshape(num_images * P, d), self.image_type_embeddings.weight.to(feats.dtype), idx=idx)
        feats, _ = self.vision_xatten_layers(feats.view(num_images, P, d))
        return feats

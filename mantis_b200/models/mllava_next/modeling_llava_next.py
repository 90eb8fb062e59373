"""LlavaNextForConditionalGeneration as Mantis ships it on the mantis_b200 CUDA kernels.

Drop-in for mantis.models.mllava_next.modeling_llava_next (reference file cited as `ref:`).  The reference keeps the
LLaVA-NeXT interface (per-image stacks of any-resolution crops in `pixel_values`, `image_sizes`, the `image_newline`
parameter) but hard-disables the any-resolution branch (`if image_feature.shape[0] > 1 and False: # debug`, ref:563), so its
arithmetic is: tower + projector over every crop, keep ONLY the first (base) crop of each image, append the
`image_newline` row -> P + 1 rows per image, then the LLaVA merge with one extra step (rows that came from pad tokens are
zeroed, ref:455-461).  This shell therefore is the LLaVA path of this package with
  * the tower run on the base crops only -- the other crops' features are discarded by the reference, so skipping them is
    output-identical and removes (crops - 1) / crops of the vision FLOPs,
  * `image_newline` concatenated after the projector,
  * `zero_pad_rows=True` in the merge,
  * `image_sizes` accepted and carried through `prepare_inputs_for_generation` (ref:677-737) but unused, like the reference.
State-dict keys are the reference's: `vision_tower.*`, `multi_modal_projector.linear_{1,2}.*`, `image_newline`,
`language_model.model.*`, `language_model.lm_head.weight`.
"""
from typing import List, Optional, Union

import torch
from torch import nn
from transformers.models.llava_next.configuration_llava_next import LlavaNextConfig

from ... import ops
from ..layers import hf_key_remap_disabled
from ..mllava.modeling_llava import LlavaCausalLMOutputWithPast, LlavaForConditionalGeneration

LlavaNextCausalLMOutputWithPast = LlavaCausalLMOutputWithPast


class LlavaNextForConditionalGeneration(LlavaForConditionalGeneration):
    config_class = LlavaNextConfig

    def __init__(self, config: LlavaNextConfig, vision_tower=None, language_model=None):
        if getattr(config, "ignore_index", None) is None:
            config.ignore_index = -100
        if not hasattr(config, "pad_token_id"):
            config.pad_token_id = None
        super().__init__(config, vision_tower=vision_tower, language_model=language_model)
        std = getattr(config, "initializer_range", None) or getattr(config.text_config, "initializer_range", 0.02)
        self.image_newline = nn.Parameter(torch.randn(config.text_config.hidden_size) * std)     # ref:324 leaves it empty

    @classmethod
    def from_pretrained(cls, *args, **kwargs):
        with hf_key_remap_disabled("llava_next"):
            return super().from_pretrained(*args, **kwargs)

    def save_pretrained(self, *args, **kwargs):
        with hf_key_remap_disabled("llava_next"):
            return super().save_pretrained(*args, **kwargs)

    @staticmethod
    def _base_crops(pixel_values) -> Optional[torch.Tensor]:
        """list of [crops_i, C, H, W] (or one [N, crops, C, H, W] tensor) -> [N, C, H, W]: the first crop of every image."""
        if pixel_values is None:
            return None
        if isinstance(pixel_values, (list, tuple)):
            return torch.stack([p[0] for p in pixel_values], dim=0)
        if pixel_values.dim() == 5:
            return pixel_values[:, 0]
        return pixel_values

    def _image_features(self, pixel_values, vision_feature_layer, vision_feature_select_strategy):
        feats = super()._image_features(pixel_values, vision_feature_layer, vision_feature_select_strategy)    # [N, P, D]
        nl = self.image_newline.to(feats.dtype)[None, None, :].expand(feats.shape[0], 1, feats.shape[2])
        return torch.cat([feats, nl], dim=1)                                                       # ref:588-589

    def _merge_input_ids_with_image_features(self, image_features, inputs_embeds, input_ids, attention_mask, labels,
                                             plan_hint=None):
        return ops.merge_input_ids_with_image_features(
            image_features, inputs_embeds, input_ids, attention_mask, labels,
            image_token_index=self.config.image_token_index, pad_token_id=self.pad_token_id,
            ignore_index=self.config.ignore_index, plan_hint=plan_hint, zero_pad_rows=True)       # ref:370-468

    def forward(self, input_ids: torch.LongTensor = None, pixel_values: Union[torch.FloatTensor, List[torch.Tensor]] = None,
                image_sizes: Optional[torch.LongTensor] = None, attention_mask: Optional[torch.Tensor] = None, **kwargs):
        return super().forward(input_ids=input_ids, pixel_values=self._base_crops(pixel_values),
                               attention_mask=attention_mask, **kwargs)

    def prepare_inputs_for_generation(self, input_ids, past_key_values=None, inputs_embeds=None, pixel_values=None,
                                      image_sizes=None, attention_mask=None, **kwargs):
        model_inputs = super().prepare_inputs_for_generation(input_ids, past_key_values=past_key_values,
                                                             inputs_embeds=inputs_embeds,
                                                             pixel_values=self._base_crops(pixel_values),
                                                             attention_mask=attention_mask, **kwargs)
        model_inputs["image_sizes"] = image_sizes
        return model_inputs

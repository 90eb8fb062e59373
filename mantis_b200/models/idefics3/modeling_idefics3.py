"""Idefics3ForConditionalGeneration (Mantis-8B-Idefics3 family) on the mantis_b200 CUDA kernels.

Drop-in for mantis.models.idefics3.modeling_idefics3 (reference file cited as `ref:`).  Idefics3 is Idefics2 with the
perceiver resampler replaced by a pixel shuffle, a LLaMA-3 text model and the default CE ignore index, so the shell reuses
the Idefics2 modules of this package:
  Idefics3VisionEmbeddings / VisionTransformer (ref:129-186,536-609)  = the NaViT SigLIP tower of idefics2 (same weights layout)
  Idefics3Connector (ref:642-668)   pixel_shuffle (pure data movement: scale_factor^2 neighbouring patches are folded into
                                    the channel dim) + ONE bias-free projection on the tcgen05 GEMM (`modality_projection.proj`)
  inputs_merger (ref:869-893)       masked row scatter, sequence length unchanged -> merge_rows kernel
  Idefics3ForConditionalGeneration  fp32 logits, attention-mask-gathered shifted CE with ignore_index = -100 (ref:1166-1180;
  (ref:1024-1286)                   Idefics2 ignores image_token_id instead), cached image_hidden_states for generate()
State-dict keys are identical to the reference's (`model.vision_model.*`, `model.connector.modality_projection.proj.weight`,
`model.text_model.*`, `lm_head.weight`).
"""
from torch import nn

try:                                                     # transformers >= 4.46 ships the config; the reference vendors a copy
    from transformers.models.idefics3.configuration_idefics3 import Idefics3Config, Idefics3VisionConfig
except ImportError:                                      # pragma: no cover - older transformers
    from transformers import CONFIG_MAPPING, PretrainedConfig

    class Idefics3VisionConfig(PretrainedConfig):
        model_type = "idefics3_vision"

        def __init__(self, hidden_size=1152, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=16,
                     num_channels=3, image_size=224, patch_size=32, hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6,
                     attention_dropout=0.0, initializer_range=0.02, **kwargs):
            super().__init__(**kwargs)
            self.hidden_size, self.intermediate_size = hidden_size, intermediate_size
            self.num_hidden_layers, self.num_attention_heads = num_hidden_layers, num_attention_heads
            self.num_channels, self.image_size, self.patch_size = num_channels, image_size, patch_size
            self.hidden_act, self.layer_norm_eps = hidden_act, layer_norm_eps
            self.attention_dropout, self.initializer_range = attention_dropout, initializer_range

    class Idefics3Config(PretrainedConfig):
        model_type = "idefics3"

        def __init__(self, use_cache=True, image_token_id=128257, tie_word_embeddings=False, vision_config=None,
                     text_config=None, scale_factor=2, pad_token_id=128002, **kwargs):
            self.image_token_id, self.use_cache, self.scale_factor = image_token_id, use_cache, scale_factor
            self.vision_config = (Idefics3VisionConfig(**(vision_config or {})) if not isinstance(vision_config, PretrainedConfig)
                                  else vision_config)
            if isinstance(text_config, dict):
                text_config = CONFIG_MAPPING[text_config.get("model_type", "llama")](**text_config)
            self.text_config = text_config if text_config is not None else CONFIG_MAPPING["llama"](pad_token_id=pad_token_id)
            super().__init__(**kwargs, pad_token_id=pad_token_id, tie_word_embeddings=tie_word_embeddings)

from ..idefics2.modeling_idefics2 import (Idefics2BaseModelOutputWithPast, Idefics2CausalLMOutputWithPast,
                                          Idefics2ForConditionalGeneration, Idefics2Model, Idefics2PreTrainedModel,
                                          Idefics2VisionTransformer)
from ..layers import B200Linear, hf_key_remap_disabled
from ..llama import B200DecoderModel

Idefics3BaseModelOutputWithPast = Idefics2BaseModelOutputWithPast
Idefics3CausalLMOutputWithPast = Idefics2CausalLMOutputWithPast


class Idefics3SimpleMLP(nn.Module):
    def __init__(self, input_size, output_size):
        super().__init__()
        self.proj = B200Linear(input_size, output_size, bias=False)

    def forward(self, x):
        return self.proj(x)


class Idefics3Connector(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.scale_factor = config.scale_factor
        self.modality_projection = Idefics3SimpleMLP(config.vision_config.hidden_size * (self.scale_factor ** 2),
                                                     config.text_config.hidden_size)

    @staticmethod
    def pixel_shuffle(x, scale_factor=2):
        """[N, h*w, C] -> [N, h*w / s^2, C * s^2]: each s x s block of neighbouring patches becomes one token whose channels are
        the block's patches in (row-in-block, column-in-block) order -- the reference's two view/permute passes (ref:651-661)
        collapsed into one permute."""
        n, seq, c = x.shape
        side = int(seq ** 0.5)
        s = scale_factor
        if side * side != seq or side % s:
            raise ValueError(f"pixel_shuffle needs a square patch grid divisible by {s}, got {seq} patches")
        x = x.view(n, side // s, s, side // s, s, c)              # [n, H/s, i, W/s, j, c]   (h = H'*s + i, w = W'*s + j)
        x = x.permute(0, 1, 3, 2, 4, 5).contiguous()               # [n, H/s, W/s, i, j, c]
        return x.view(n, seq // (s * s), c * s * s)

    def forward(self, image_hidden_states):
        return self.modality_projection(self.pixel_shuffle(image_hidden_states, self.scale_factor))


class Idefics3PreTrainedModel(Idefics2PreTrainedModel):
    config_class = Idefics3Config
    _no_split_modules = ["B200VisionEncoderLayer", "B200DecoderLayer"]

    @classmethod
    def from_pretrained(cls, *args, **kwargs):
        with hf_key_remap_disabled("idefics3", "llama"):
            return super(Idefics2PreTrainedModel, cls).from_pretrained(*args, **kwargs)

    def save_pretrained(self, *args, **kwargs):
        with hf_key_remap_disabled("idefics3", "llama"):
            return super(Idefics2PreTrainedModel, self).save_pretrained(*args, **kwargs)


class Idefics3Model(Idefics3PreTrainedModel, Idefics2Model):
    """vision tower -> pixel shuffle + projection -> masked scatter into the text sequence -> LLaMA decoder (ref:816-1021)"""

    def __init__(self, config: Idefics3Config):
        Idefics2PreTrainedModel.__init__(self, config)
        self.padding_idx = self.config.text_config.pad_token_id
        self.vocab_size = self.config.text_config.vocab_size
        self.vision_model = Idefics2VisionTransformer(config.vision_config)
        self.connector = Idefics3Connector(config)
        self.text_model = B200DecoderModel(config.text_config)
        self.image_seq_len = int(((config.vision_config.image_size // config.vision_config.patch_size) ** 2)
                                 / (config.scale_factor ** 2))                                          # ref:826-828
        self.image_token_id = self.config.image_token_id
        self.post_init()

    def _connect(self, x, patch_key_mask):
        return self.connector(x)


class Idefics3ForConditionalGeneration(Idefics3PreTrainedModel, Idefics2ForConditionalGeneration):
    _tied_weights_keys = {}
    _backbone_cls = Idefics3Model

    def _loss_ignore_index(self):
        return -100                                               # plain CrossEntropyLoss() (ref:1178)

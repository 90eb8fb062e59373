"""LlavaConfig -- same fields / defaults as mantis.models.mllava.configuration_llava.LlavaConfig
(reference: mantis/models/mllava/configuration_llava.py:32-133)."""
from transformers import PretrainedConfig
from transformers.models.auto import CONFIG_MAPPING


class LlavaConfig(PretrainedConfig):
    model_type = "llava"
    is_composition = False
    sub_configs = {}

    def __init__(self, vision_config=None, text_config=None, ignore_index=-100, image_token_index=32000,
                 projector_hidden_act="gelu", vision_feature_select_strategy="default", vision_feature_layer=-2,
                 vocab_size=32000, **kwargs):
        self.ignore_index = ignore_index
        self.image_token_index = image_token_index
        self.projector_hidden_act = projector_hidden_act
        self.vision_feature_select_strategy = vision_feature_select_strategy
        self.vision_feature_layer = vision_feature_layer
        self.vocab_size = vocab_size

        if isinstance(vision_config, dict):
            vision_config = dict(vision_config)
            vision_config.setdefault("model_type", "clip_vision_model")
            vision_config = CONFIG_MAPPING[vision_config["model_type"]](**vision_config)
        elif vision_config is None:
            vision_config = CONFIG_MAPPING["clip_vision_model"](
                intermediate_size=4096, hidden_size=1024, patch_size=14, image_size=336, num_hidden_layers=24,
                num_attention_heads=16, vocab_size=32000, projection_dim=768)
        self.vision_config = vision_config

        if isinstance(text_config, dict):
            text_config = dict(text_config)
            text_config.setdefault("model_type", "llama")
            text_config = CONFIG_MAPPING[text_config["model_type"]](**text_config)
            self.vocab_size = text_config.vocab_size
        elif text_config is None:
            text_config = CONFIG_MAPPING["llama"]()
        self.text_config = text_config
        super().__init__(**kwargs)

    def to_dict(self):
        out = super().to_dict()
        for k in ("vision_config", "text_config"):
            v = getattr(self, k, None)
            if v is not None and hasattr(v, "to_dict"):
                out[k] = v.to_dict()
        return out


# Constants of the released Mantis-8B-SigLIP-LLaMA-3 (no config.json is available offline)
def mantis_8b_siglip_llama3_config(num_vision_layers=27, num_text_layers=32, **overrides):
    from transformers import LlamaConfig, SiglipVisionConfig
    vc = SiglipVisionConfig(hidden_size=1152, intermediate_size=4304, num_hidden_layers=num_vision_layers,
                            num_attention_heads=16, image_size=384, patch_size=14, layer_norm_eps=1e-6,
                            hidden_act="gelu_pytorch_tanh")
    tc = LlamaConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=num_text_layers,
                     num_attention_heads=32, num_key_value_heads=8, vocab_size=128258, rms_norm_eps=1e-5,
                     rope_theta=500000.0, max_position_embeddings=8192, tie_word_embeddings=False,
                     bos_token_id=128000, eos_token_id=128001)
    cfg = dict(vision_config=vc, text_config=tc, image_token_index=128256, pad_token_id=128257, vocab_size=128258,
               vision_feature_select_strategy="default", vision_feature_layer=-2, projector_hidden_act="gelu")
    cfg.update(overrides)
    return LlavaConfig(**cfg)

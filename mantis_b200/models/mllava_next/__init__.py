from .modeling_llava_next import LlavaNextConfig, LlavaNextForConditionalGeneration  # noqa: F401

__all__ = ["LlavaNextForConditionalGeneration", "LlavaNextConfig"]

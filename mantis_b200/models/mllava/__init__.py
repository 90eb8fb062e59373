from .configuration_llava import LlavaConfig, mantis_8b_siglip_llama3_config
from .modeling_llava import (LlavaCausalLMOutputWithPast, LlavaForConditionalGeneration,
                             MLlavaForConditionalGeneration)

__all__ = ["LlavaConfig", "LlavaForConditionalGeneration", "MLlavaForConditionalGeneration",
           "LlavaCausalLMOutputWithPast", "mantis_8b_siglip_llama3_config"]
from .processing_llava import MLlavaProcessor  # noqa: E402
from .utils import chat_mllava, chat_mllava_stream  # noqa: E402

__all__ += ["MLlavaProcessor", "chat_mllava", "chat_mllava_stream"]
from .image_processing import B200ImageProcessor  # noqa: F401,E402

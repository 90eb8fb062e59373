from mantis_b200.models.mllava import (LlavaConfig, LlavaForConditionalGeneration,  # noqa: F401
                                       MLlavaForConditionalGeneration)
from mantis_b200.models.mllava.processing_llava import MLlavaProcessor  # noqa: F401
from mantis_b200.models.mllava.utils import chat_mllava, chat_mllava_stream  # noqa: F401

from mantis_b200.models.mllava import (LlavaConfig, LlavaForConditionalGeneration,  # noqa: F401
                                       MLlavaForConditionalGeneration)
try:  # processor / chat helpers (callers of the hot path)
    from mantis_b200.models.mllava.processing_llava import MLlavaProcessor  # noqa: F401
    from mantis_b200.models.mllava.utils import chat_mllava  # noqa: F401
except Exception:  # pragma: no cover
    pass

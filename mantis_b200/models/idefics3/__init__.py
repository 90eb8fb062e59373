from .modeling_idefics3 import (Idefics3Config, Idefics3ForConditionalGeneration, Idefics3Model,  # noqa: F401
                                Idefics3VisionConfig)

__all__ = ["Idefics3ForConditionalGeneration", "Idefics3Model", "Idefics3Config", "Idefics3VisionConfig"]

from .modeling_idefics2 import (Idefics2ForConditionalGeneration, Idefics2ForSequenceClassification,
                                Idefics2Model)

__all__ = ["Idefics2ForConditionalGeneration", "Idefics2ForSequenceClassification", "Idefics2Model"]

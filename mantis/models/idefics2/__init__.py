from mantis_b200.models.idefics2 import (Idefics2ForConditionalGeneration,  # noqa: F401
                                         Idefics2ForSequenceClassification)

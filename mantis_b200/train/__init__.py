from .engine import B200Trainer, FlatState  # noqa: F401
from .data import (ChatDataset, Collator, PackingDataset, assistant_labels, llava_valid_rows, merged_length,  # noqa: F401
                   partition_balanced, plain_valid_rows)

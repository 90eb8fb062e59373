from .engine import B200Trainer, flat_grad_buffer  # noqa: F401
from .data import Collator, PackingDataset, merged_length, partition_balanced  # noqa: F401

from .engine import B200Trainer, flat_grad_buffer  # noqa: F401
from .data import Collator, PackingDataset  # noqa: F401

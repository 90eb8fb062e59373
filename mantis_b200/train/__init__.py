from .engine import B200Trainer, flat_grad_buffer  # noqa: F401

"""Collation for the hot path (mirror of mantis/train/data.py:1375-1544 `Collator`, caller-side).

The reference delegates to `MLlavaProcessor._right_pad_inputs_with_attention_mask`, which asserts batch == 1.  This
collator keeps that contract when given one sample and additionally right-pads real batches (input_ids with the pad id,
attention_mask with 0, labels with -100, position_ids with 0) so that the B > 1 merge path of the model can be used;
`pixel_values` stays a python list of per-sample tensors exactly like the reference (modeling_llava.py:431-432 cats it).
"""
from typing import Dict, List

import torch


class Collator:
    def __init__(self, processor=None, max_length=None, pad_token_id=None, label_pad=-100):
        self.processor = processor
        self.max_length = max_length
        self.pad_token_id = pad_token_id
        self.label_pad = label_pad

    def _pad_id(self):
        if self.pad_token_id is not None:
            return self.pad_token_id
        tok = getattr(self.processor, "tokenizer", None)
        pid = getattr(tok, "pad_token_id", None)
        if pid is None:
            raise ValueError("Collator needs a pad_token_id to pad batches of more than one sample")
        return pid

    def _pad_to(self, t: torch.Tensor, length: int, value) -> torch.Tensor:
        if t.shape[1] >= length:
            return t
        pad = torch.full((t.shape[0], length - t.shape[1]), value, dtype=t.dtype, device=t.device)
        return torch.cat([t, pad], dim=1)

    def __call__(self, batch: List[Dict]):
        if len(batch) == 1 and self.processor is not None and hasattr(self.processor, "_right_pad_inputs_with_attention_mask"):
            return self.processor._right_pad_inputs_with_attention_mask(model_inputs=batch)      # the reference's path
        out = {}
        for k in batch[0].keys():
            vals = [b[k] for b in batch]
            if "pixel_values" in k:
                out[k] = [v for v in vals]                       # list of per-sample tensors (None kept, model skips them)
            elif vals[0] is None:
                out[k] = None
            else:
                L = max(v.shape[1] for v in vals)
                if self.max_length is not None:
                    L = min(L, self.max_length)
                fill = self._pad_id() if k == "input_ids" else (self.label_pad if k == "labels" else 0)
                out[k] = torch.cat([self._pad_to(v[:, :L], L, fill) for v in vals], dim=0)
        return out

"""Sequence-packed fine-tuning of an Idefics2-style model on the B200 engine: the reference's `PackingDataset` layout
(mantis/train/data.py:1546-1671) + the launch script's cosine schedule (mantis/train/scripts/train_mllava.sh:162-165).

Several short samples share one row; attention stays block-diagonal (one tcgen05 launch per sample on strided views of the
packed q/k/v -- no S x S mask is read), RoPE restarts at every sample, the loss mask is the per-sample key mask.

    python examples/packed_finetune.py            # synthetic data, 2-layer model by default (LAYERS=32 for the 8B shape)
"""
import os

import torch

from bench import idefics2_8b_config
from mantis_b200.models.idefics2 import Idefics2ForConditionalGeneration
from mantis_b200.train import B200Trainer, Collator, PackingDataset


class Synthetic(torch.utils.data.Dataset):
    """text-only chat turns of ragged length (images are merged before packing in the Idefics2 family, so packing sees
    plain token rows; pixel_values ride along untouched)"""

    def __init__(self, n=64, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.rows = [torch.randint(3, 32000, (1, int(torch.randint(64, 700, (1,), generator=g))), generator=g) for _ in range(n)]

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, i):
        ids = self.rows[i]
        return {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": ids.clone(), "pixel_values": None}


def main():
    torch.cuda.set_device(0)
    layers = int(os.environ.get("LAYERS", "2"))
    torch.set_default_dtype(torch.bfloat16)
    with torch.device("cuda"):
        model = Idefics2ForConditionalGeneration(idefics2_8b_config(layers, 1))
    torch.set_default_dtype(torch.float32)
    packed = PackingDataset(Synthetic(), max_self_attn_len=2048, dense_mask=False)     # 2-D key mask + segment table
    collate = Collator(pad_token_id=0)
    steps = len(packed)
    trainer = B200Trainer(model.train(), lr=1e-5, grad_accum=1, lr_schedule="cosine", total_steps=steps, warmup_ratio=0.03)
    for step in range(steps):
        batch = collate([packed[step]])
        batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()
                 if v is not None and not (k == "pixel_values" and all(x is None for x in v))}      # text-only rows
        lr = trainer.current_lr()
        loss = trainer.train_step([batch])
        n_seg = len(batch["cu_segments"])
        print(f"step {step}: {n_seg} samples packed into {batch['input_ids'].shape[1]} tokens, lr {lr:.2e}, loss {loss.item():.4f}")


if __name__ == "__main__":
    main()

"""Minimal instruction-tuning loop on the B200 engine (what mantis/train/train_mllava.py + HF Trainer do for the hot path):
per-device batch 1, gradient accumulation, frozen vision tower, fused AdamW, one overlapped all-reduce per step.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_step.py
"""
import os

import torch
import torch.distributed as dist

from mantis_b200.models.mllava import LlavaForConditionalGeneration, mantis_8b_siglip_llama3_config
from mantis_b200.train import B200Trainer


def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.set_default_dtype(torch.bfloat16)
    with torch.device("cuda"):
        model = LlavaForConditionalGeneration(mantis_8b_siglip_llama3_config(num_text_layers=int(os.environ.get("LAYERS", "32"))))
    torch.set_default_dtype(torch.float32)
    trainer = B200Trainer(model.train(), lr=1e-5, grad_accum=4)
    g = torch.Generator().manual_seed(local)
    for step in range(3):
        batches = []
        for _ in range(4):
            ids = torch.randint(0, 128000, (1, 2048), generator=g)
            ids[0, 16::256] = 128256
            labels = ids.clone(); labels[ids == 128256] = -100
            batches.append(dict(input_ids=ids.cuda(), attention_mask=torch.ones_like(ids).cuda(), labels=labels.cuda(),
                                pixel_values=torch.randn(8, 3, 384, 384, generator=g).bfloat16().cuda()))
        loss = trainer.train_step(batches)
        if local == 0:
            print(f"step {step}: loss {loss.item():.4f}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Minimal data-parallel instruction-tuning engine for the hot path (what HF Trainer + accelerate/DeepSpeed do for
mantis/train/train_mllava.py:312-329 with per_device_train_batch_size=1 + gradient accumulation,
mantis/train/scripts/train_mllava.sh:137-168), reduced to what the step needs:

  * micro-batches of one sample: loss/accum -> backward (gradients accumulate in ONE flat bf16 buffer that every
    trainable parameter's .grad is a view of),
  * one NCCL all-reduce (sum, then 1/world) of that flat buffer per optimizer step -- the only exchange of the
    data-parallel path; the vision tower is frozen and excluded,
  * optional global-norm clipping + fused AdamW (bf16 params, fp32 moments) on our CUDA kernel.
"""
import math

import torch
import torch.distributed as dist

from .. import ops


def flat_grad_buffer(params):
    """Allocates one contiguous buffer and makes every param.grad a view into it. Returns (flat, views)."""
    params = [p for p in params if p.requires_grad]
    total = sum((p.numel() + 7) // 8 * 8 for p in params)      # keep every view 16-byte aligned
    dtype = params[0].dtype
    assert all(p.dtype == dtype for p in params), "trainable parameters must share a dtype"
    flat = torch.zeros(total, dtype=dtype, device=params[0].device)
    off = 0
    for p in params:
        n = p.numel()
        p.grad = flat[off:off + n].view_as(p)
        off += (n + 7) // 8 * 8
    return flat


class _GradReady(torch.autograd.Function):
    """identity whose backward fires a callback: placed at a decoder layer's input, it runs once that layer's backward
    (and therefore every gradient of the layers above it) has been produced"""

    @staticmethod
    def forward(ctx, x, cb):
        ctx.cb = cb
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        ctx.cb()
        return g, None


def lr_lambda(step, total_steps, warmup_steps, kind="cosine"):
    """Multiplier of the base LR for optimizer step `step` (0-based) -- the schedules HF Trainer builds for
    `--lr_scheduler_type {cosine,linear,constant}` with `--warmup_ratio` (mantis/train/scripts/train_mllava.sh:162-165;
    transformers/optimization.py get_cosine_schedule_with_warmup / get_linear_schedule_with_warmup): linear warm-up from 0,
    then half a cosine period (or a straight line) down to 0 at `total_steps`."""
    if kind == "constant" or not total_steps:
        return 1.0 if step >= warmup_steps or not warmup_steps else float(step) / float(max(1, warmup_steps))
    if step < warmup_steps:
        return float(step) / float(max(1, warmup_steps))
    if kind == "linear":
        return max(0.0, float(total_steps - step) / float(max(1, total_steps - warmup_steps)))
    if kind == "cosine":
        progress = float(step - warmup_steps) / float(max(1, total_steps - warmup_steps))
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * 2.0 * 0.5 * progress)))
    raise ValueError(f"unknown lr schedule {kind!r}")


class B200Trainer:
    def __init__(self, model, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=1.0,
                 grad_accum=1, freeze_vision=True, fused_wgrad_accum=True, overlap_allreduce=True,
                 lr_schedule="constant", total_steps=None, warmup_ratio=0.0, warmup_steps=None):
        self.model = model
        self.lr_schedule, self.total_steps = lr_schedule, total_steps
        # HF Trainer: warmup_steps wins over warmup_ratio; the ratio is rounded up (TrainingArguments.get_warmup_steps)
        self.warmup_steps = (warmup_steps if warmup_steps is not None
                             else int(math.ceil((total_steps or 0) * warmup_ratio)))
        if freeze_vision:                                   # mantis/train/train_mllava.py:239-242
            for n, p in model.named_parameters():
                if "vision_tower" in n or "vision_model" in n:
                    p.requires_grad_(False)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.flat_grad = flat_grad_buffer(self.params)
        if fused_wgrad_accum:
            # linear layers accumulate dW straight into the flat buffer from the wgrad GEMM epilogue
            from ..models.layers import B200Linear
            for mod in model.modules():
                if isinstance(mod, B200Linear) and mod.weight.requires_grad and mod.weight.grad is not None:
                    mod.weight._b200_fused_grad = True
        self.m = [torch.zeros(p.shape, dtype=torch.float32, device=p.device) for p in self.params]
        self.v = [torch.zeros(p.shape, dtype=torch.float32, device=p.device) for p in self.params]
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.max_grad_norm = max_grad_norm
        self.grad_accum = grad_accum
        self.step_count = 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self._norm = torch.zeros(1, dtype=torch.float32, device=self.flat_grad.device)
        self._overlap = False
        self._works = []
        self._reduced_from = None
        if overlap_allreduce and self.world > 1:
            self._install_overlap_hooks()

    # ---- overlapped gradient all-reduce (N > 1): during the LAST micro-batch's backward, the slice of the flat buffer
    # belonging to decoder layer i+1 (and everything after it) is reduced as soon as layer i's backward has run, so the
    # collective hides under the remaining backward; the head of the buffer is reduced after backward.
    def _install_overlap_hooks(self):
        layers = None
        for name, mod in self.model.named_modules():
            if name.endswith("language_model.model.layers") or name.endswith("text_model.layers"):
                layers = mod
        if layers is None:
            return
        base = self.flat_grad.data_ptr(); esz = self.flat_grad.element_size()
        starts = []
        for layer in layers:
            ps = [p for p in layer.parameters() if p.requires_grad]
            starts.append(min((p.grad.data_ptr() - base) // esz for p in ps) if ps else None)
        if any(s is None for s in starts) or starts != sorted(starts):
            return
        self._layer_starts = starts
        n = len(layers)

        def make_cb(i):
            def cb():
                if not self._overlap:
                    return
                lo = self._layer_starts[i + 1] if i + 1 < n else None
                if lo is None or self._reduced_from is None or lo >= self._reduced_from:
                    return
                self._works.append(dist.all_reduce(self.flat_grad[lo:self._reduced_from], op=dist.ReduceOp.SUM, async_op=True))
                self._reduced_from = lo
            return cb

        def make_hook(i):
            cb = make_cb(i)

            def hook(mod, args):
                if self._overlap and torch.is_grad_enabled() and args and args[0].requires_grad:
                    return (_GradReady.apply(args[0], cb),) + tuple(args[1:])
                return None
            return hook

        for i, layer in enumerate(layers):
            layer.register_forward_pre_hook(make_hook(i))
        self._has_hooks = True

    def current_lr(self):
        """learning rate of the NEXT optimizer step (scheduler value after `step_count` completed steps)"""
        return self.lr * lr_lambda(self.step_count, self.total_steps, self.warmup_steps, self.lr_schedule)

    # ---- checkpoint / resume of the optimizer side (the model itself goes through save_pretrained) ----
    def state_dict(self):
        return {"step": self.step_count, "lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.wd,
                "lr_schedule": self.lr_schedule, "total_steps": self.total_steps, "warmup_steps": self.warmup_steps,
                "exp_avg": [m.detach().cpu() for m in self.m], "exp_avg_sq": [v.detach().cpu() for v in self.v]}

    def load_state_dict(self, state):
        if len(state["exp_avg"]) != len(self.m):
            raise ValueError(f"optimizer state has {len(state['exp_avg'])} tensors, the model has {len(self.m)} trainable ones")
        for dst, src in zip(self.m, state["exp_avg"]):
            if dst.shape != src.shape:
                raise ValueError("optimizer state does not match the trainable parameters")
            dst.copy_(src)
        for dst, src in zip(self.v, state["exp_avg_sq"]):
            dst.copy_(src)
        self.step_count = int(state["step"])
        self.lr, self.betas, self.eps, self.wd = state["lr"], tuple(state["betas"]), state["eps"], state["weight_decay"]
        self.lr_schedule, self.total_steps = state["lr_schedule"], state["total_steps"]
        self.warmup_steps = state["warmup_steps"]

    def zero_grad(self):
        self.flat_grad.zero_()

    def micro_step(self, batch):
        """forward + backward of one micro-batch (gradients accumulate). Returns the detached loss."""
        out = self.model(**batch)
        (out.loss / self.grad_accum).backward()
        return out.loss.detach()

    def _restore_grad_views(self):
        # autograd accumulates in place into existing .grad tensors, so the views stay bound; assert cheaply
        p = self.params[0]
        assert p.grad.data_ptr() == self.flat_grad.data_ptr(), "param.grad was rebound away from the flat buffer"

    def reduce_gradients(self):
        """the one collective of the data-parallel path: sum the flat gradient buffer over ranks (the part not already
        reduced under the last backward).  Returns the scale (1/world) still to be applied (folded into AdamW)."""
        if self.world > 1:
            rf = getattr(self, "_reduced_from", None)
            hi = rf if rf is not None else self.flat_grad.numel()
            if hi > 0:
                dist.all_reduce(self.flat_grad[:hi], op=dist.ReduceOp.SUM)
            for w in getattr(self, "_works", []):
                w.wait()
            self._works = []
            self._reduced_from = None
        return 1.0 / self.world

    def optimizer_step(self):
        self._restore_grad_views()
        scale = self.reduce_gradients()
        if self.max_grad_norm is not None and self.max_grad_norm > 0:
            self._norm.zero_()
            ops.sumsq(self.flat_grad, self._norm)
            # clip factor computed on device; read back lazily (no sync needed: pass through a tiny host value
            # only when logging).  The fused kernel takes the scale as a host float, so one 4-byte readback here.
            total = math.sqrt(float(self._norm.item())) * scale
            if total > self.max_grad_norm:
                scale *= self.max_grad_norm / (total + 1e-6)
        ops.check_deferred()                       # errors of sync-free merges surface here, after the step's one readback
        lr = self.current_lr()
        self.step_count += 1
        for p, m, v in zip(self.params, self.m, self.v):
            ops.adamw_step(p.data, p.grad, m, v, lr, self.betas[0], self.betas[1], self.eps, self.wd,
                           self.step_count, grad_scale=scale)
        self.zero_grad()

    def train_step(self, micro_batches):
        losses = []
        for i, b in enumerate(micro_batches):
            last = i == len(micro_batches) - 1
            if last and self.world > 1 and getattr(self, "_has_hooks", False):
                self._overlap = True
                self._reduced_from = self.flat_grad.numel()
            try:
                losses.append(self.micro_step(b))
            finally:
                self._overlap = False
        self.optimizer_step()
        return torch.stack(losses).mean()

"""Minimal data-parallel instruction-tuning engine for the hot path (what HF Trainer + accelerate/DeepSpeed do for
mantis/train/train_mllava.py:312-329 with per_device_train_batch_size=1 + gradient accumulation,
mantis/train/scripts/train_mllava.sh:137-168, zero_configs/zero3.json "bf16": enabled), reduced to what the step needs:

  * FLAT training state, one allocation each, every tensor's slice starting on a 1024-element boundary:
      P   bf16  the model's weights (every trainable parameter is re-pointed at its slice)
      LO  int16 the 16 low bits of each weight's fp32 master copy -- (P, LO) together ARE the fp32 master weights DeepSpeed's
                bf16 optimizer keeps (an lr = 1e-5 AdamW step is far below half a bf16 ulp of a typical weight and would be
                rounded away if the bf16 weights were updated in place); 2 extra bytes per parameter instead of 4
      G   fp32  the main gradient: micro-batches accumulate here in fp32 (wgrad GEMMs write it from their epilogue; gradients
                autograd produces in bf16 are folded in by a post-accumulate hook), and it is the ONE buffer the data-parallel
                all-reduce exchanges; the vision tower is frozen and excluded
      M,V fp32  AdamW moments
  * one fused launch per optimizer step: global-norm clip factor computed on the device from sum(G^2) (no host read-back),
    AdamW on the fp32 master, bf16 weights re-rounded, G zeroed.
"""
import math

import torch
import torch.distributed as dist

from .. import ops

ALIGN = 1024        # elements; one param-group byte per block (ops.adamw_flat)


def _aligned(n):
    return (n + ALIGN - 1) // ALIGN * ALIGN


class FlatState:
    """The flat buffers above for an ordered list of parameters."""

    def __init__(self, params, grad_dtype=torch.float32):
        self.params = params
        dtype = params[0].dtype
        if not all(p.dtype == dtype for p in params):
            raise ValueError("trainable parameters must share a dtype")
        dev = params[0].device
        self.offsets, off = [], 0
        for p in params:
            self.offsets.append(off)
            off += _aligned(p.numel())
        self.total = off
        self.split = dtype == torch.bfloat16
        self.P = torch.zeros(self.total, dtype=dtype, device=dev)
        self.LO = torch.zeros(self.total, dtype=torch.int16, device=dev) if self.split else None
        self.G = torch.zeros(self.total, dtype=grad_dtype if self.split else torch.float32, device=dev)
        self.M = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.V = torch.zeros(self.total, dtype=torch.float32, device=dev)
        groups = torch.zeros(self.total // ALIGN, dtype=torch.uint8)
        for p, o in zip(params, self.offsets):
            n = p.numel()
            with torch.no_grad():
                self.P[o:o + n].copy_(p.detach().reshape(-1))
            p.data = self.P[o:o + n].view(p.shape)                     # the model now computes with the flat buffer
            p._b200_main_grad = self.G[o:o + n].view(p.shape)
            if p.dim() <= 1:                                           # biases / norm weights: no weight decay (HF Trainer's
                groups[o // ALIGN:(o + _aligned(n)) // ALIGN] = 1      # get_decay_parameter_names excludes exactly these)
        self.groups = groups.to(dev)

    def view(self, buf, i):
        p, o = self.params[i], self.offsets[i]
        return buf[o:o + p.numel()].view(p.shape)

    def master(self, i):
        """fp32 master copy of parameter i (reconstructed from the bf16 weight and its low half)"""
        p, o = self.params[i], self.offsets[i]
        n = p.numel()
        if not self.split:
            return self.P[o:o + n].view(p.shape).clone()
        return ops.master_join(self.P[o:o + n], self.LO[o:o + n]).view(p.shape)

    def set_master(self, i, value):
        p, o = self.params[i], self.offsets[i]
        n = p.numel()
        src = value.to(device=self.P.device, dtype=torch.float32).reshape(-1).contiguous()
        if not self.split:
            self.P[o:o + n].copy_(src)
        else:
            ops.master_split(src, self.P[o:o + n], self.LO[o:o + n])


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
                 lr_schedule="constant", total_steps=None, warmup_ratio=0.0, warmup_steps=None, grad_dtype=torch.float32):
        self.model = model
        self.lr_schedule, self.total_steps = lr_schedule, total_steps
        # HF Trainer: warmup_steps wins over warmup_ratio; the ratio is rounded up (TrainingArguments.get_warmup_steps)
        self.warmup_steps = (warmup_steps if warmup_steps is not None
                             else int(math.ceil((total_steps or 0) * warmup_ratio)))
        if freeze_vision:                                   # mantis/train/train_mllava.py:239-242
            for n, p in model.named_parameters():
                if "vision_tower" in n or "vision_model" in n:
                    p.requires_grad_(False)
        self._handles = []
        self.params, self._layer_ranges = self._ordered_params(model)
        self.state = FlatState(self.params, grad_dtype=grad_dtype)
        self.flat_grad = self.state.G
        from ..models.layers import B200Linear
        fused = set()
        if fused_wgrad_accum:
            # linear layers accumulate dW straight into the flat main gradient from the wgrad GEMM epilogue
            for mod in model.modules():
                if isinstance(mod, B200Linear) and mod.weight.requires_grad and mod.weight.dim() == 2:
                    fused.add(id(mod.weight))
        for p in self.params:
            # Gradients that reach a parameter through autograd (norm weights, biases, embeddings, the LM head's dW out of the
            # fused LM-head/CE, any weight used outside ops.linear) arrive in the parameter dtype: fold each into the fp32 main
            # gradient as soon as it is produced and drop the temporary.  Fused-wgrad weights normally never see one.
            p._b200_unfused_main_grad = p._b200_main_grad
            if id(p) not in fused:
                p._b200_main_grad = None                       # ops._LinearFn: not a fused-wgrad weight
            self._handles.append(p.register_post_accumulate_grad_hook(self._fold_grad))
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.max_grad_norm = max_grad_norm
        self.grad_accum = grad_accum
        self.step_count = 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self._norm = torch.zeros(1, dtype=torch.float32, device=self.flat_grad.device)
        self.last_grad_norm_sq = self._norm               # device scalar: sum(G^2) of the last step (before the 1/world scale)
        self._overlap = False
        self._works = []
        self._reduced_from = None
        if overlap_allreduce and self.world > 1:
            self._install_overlap_hooks()

    @staticmethod
    def _fold_grad(p):
        mg = p._b200_unfused_main_grad
        g = p.grad
        if g is None:
            return
        fresh = getattr(p, "_b200_grad_fresh", False)
        if g.is_cuda and mg.dtype == torch.float32:
            ops.accum_f32(mg.reshape(-1), g.contiguous().view(-1), accumulate=not fresh)
        elif fresh:                            # bf16 main gradient, or the CPU/gloo plumbing tests (no kernels there)
            mg.copy_(g)
        else:
            mg.add_(g.to(mg.dtype))
        p._b200_grad_fresh = False
        p.grad = None

    @staticmethod
    def _ordered_params(model):
        """Trainable parameters in flat-buffer order + the [start, end) parameter-index range of each decoder layer.
        Order: everything that is NOT part of the text decoder's layers / final norm / LM head first (projector, connector,
        MLlava's image_type_embeddings and vision_xatten_layers, token embeddings: their gradients are only final at the very
        end of backward), then decoder layer 0 .. L-1, then the final norm and the LM head (final first).  The overlapped
        all-reduce walks this buffer from the tail, so a slice is only ever reduced after all of its gradients exist."""
        layers = None
        for name, mod in model.named_modules():
            if name.endswith("language_model.model.layers") or name.endswith("text_model.layers"):
                layers = mod
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        if layers is None:
            return [p for _, p in named], None
        layer_ids = [{id(p) for p in layer.parameters()} for layer in layers]
        in_layer = set().union(*layer_ids) if layer_ids else set()
        tail_names = ("language_model.model.norm.", "language_model.lm_head.", "text_model.norm.", "lm_head.")
        head, tail = [], []
        for n, p in named:
            if id(p) in in_layer:
                continue
            (tail if any(t in n for t in tail_names) and "vision" not in n and "perceiver" not in n else head).append(p)
        order, ranges = list(head), []
        for layer in layers:
            ps = [p for p in layer.parameters() if p.requires_grad]
            ranges.append((len(order), len(order) + len(ps)))
            order += ps
        order += tail
        seen, uniq = set(), []
        for p in order:                                     # tied weights appear once
            if id(p) not in seen:
                seen.add(id(p)); uniq.append(p)
        if len(uniq) != len(order):
            return [p for _, p in named], None
        return order, ranges

    # ---- overlapped gradient all-reduce (N > 1): during the LAST micro-batch's backward, the slice of the flat buffer
    # belonging to decoder layer i+1 (and everything after it) is reduced as soon as layer i's backward has run, so the
    # collective hides under the remaining backward; the head of the buffer is reduced after backward.
    def _install_overlap_hooks(self):
        layers = None
        for name, mod in self.model.named_modules():
            if name.endswith("language_model.model.layers") or name.endswith("text_model.layers"):
                layers = mod
        if layers is None or self._layer_ranges is None or any(a == b for a, b in self._layer_ranges):
            return
        self._layer_starts = [self.state.offsets[a] for a, _ in self._layer_ranges]      # ascending by construction
        n = len(layers)

        def make_cb(i):
            def cb():
                if not self._overlap:
                    return
                lo = self._layer_starts[i + 1] if i + 1 < n else None
                if lo is None or self._reduced_from is None or lo >= self._reduced_from:
                    return
                if self.flat_grad.is_cuda:
                    ops.join_wgrad_stream(self.flat_grad.device)
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
            self._handles.append(layer.register_forward_pre_hook(make_hook(i)))
        self._has_hooks = True

    def close(self):
        """Detach the trainer from the model: hooks removed, gradient / moment / master-low buffers released (112 GB for
        Mantis-8B).  The weights stay where they are (the flat bf16 buffer), so the model keeps working, e.g. for generate()."""
        for h in self._handles:
            h.remove()
        self._handles = []
        for p in self.params:
            for name in ("_b200_main_grad", "_b200_unfused_main_grad"):
                if hasattr(p, name):
                    delattr(p, name)
        st = self.state
        st.G = st.M = st.V = st.LO = None
        self.flat_grad = None
        self._has_hooks = False

    def current_lr(self):
        """learning rate of the NEXT optimizer step (scheduler value after `step_count` completed steps)"""
        return self.lr * lr_lambda(self.step_count, self.total_steps, self.warmup_steps, self.lr_schedule)

    # ---- checkpoint / resume of the optimizer side (the model itself goes through save_pretrained) ----
    def state_dict(self):
        st = self.state
        n = len(self.params)
        return {"step": self.step_count, "lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.wd,
                "lr_schedule": self.lr_schedule, "total_steps": self.total_steps, "warmup_steps": self.warmup_steps,
                "exp_avg": [st.view(st.M, i).detach().cpu().clone() for i in range(n)],
                "exp_avg_sq": [st.view(st.V, i).detach().cpu().clone() for i in range(n)],
                # the fp32 master weights (DeepSpeed checkpoints them too): without them a resumed run restarts from the
                # bf16 rounding of every weight
                "master_params": [st.master(i).detach().cpu() for i in range(n)]}

    def load_state_dict(self, state):
        st = self.state
        if len(state["exp_avg"]) != len(self.params):
            raise ValueError(f"optimizer state has {len(state['exp_avg'])} tensors, the model has {len(self.params)} trainable ones")
        for i, src in enumerate(state["exp_avg"]):
            if st.view(st.M, i).shape != src.shape:
                raise ValueError("optimizer state does not match the trainable parameters")
            st.view(st.M, i).copy_(src)
        for i, src in enumerate(state["exp_avg_sq"]):
            st.view(st.V, i).copy_(src)
        for i, src in enumerate(state.get("master_params") or []):
            st.set_master(i, src)
        self.step_count = int(state["step"])
        self.lr, self.betas, self.eps, self.wd = state["lr"], tuple(state["betas"]), state["eps"], state["weight_decay"]
        self.lr_schedule, self.total_steps = state["lr_schedule"], state["total_steps"]
        self.warmup_steps = state["warmup_steps"]

    def zero_grad(self):
        self.flat_grad.zero_()
        for p in self.params:
            p._b200_grad_fresh = False

    def _mark_fresh(self):
        """after an optimizer step nothing is zero-filled: the next gradient written to a parameter's slice overwrites it"""
        for p in self.params:
            p._b200_grad_fresh = True

    def _zero_untouched(self):
        """a parameter that received no gradient since the last optimizer step still holds the previous step's: clear it"""
        for p in self.params:
            if getattr(p, "_b200_grad_fresh", False):
                p._b200_unfused_main_grad.zero_()
                p._b200_grad_fresh = False

    def micro_step(self, batch):
        """forward + backward of one micro-batch (gradients accumulate). Returns the detached loss."""
        out = self.model(**batch)
        (out.loss / self.grad_accum).backward()
        return out.loss.detach()

    def _check_views(self):
        p = self.params[0]
        if p.data_ptr() != self.state.P.data_ptr():
            raise RuntimeError("a trainable parameter was re-allocated after B200Trainer was built (model.to()/load with "
                               "assign=True?): its storage must stay the trainer's flat buffer")

    def reduce_gradients(self):
        """the one collective of the data-parallel path: sum the flat gradient buffer over ranks (the part not already
        reduced under the last backward).  Returns the scale (1/world) still to be applied (folded into AdamW).
        With `time_comm` set, the time the compute stream spends in / waiting on collectives after backward (the EXPOSED part of
        the exchange) is bracketed with CUDA events; read it with exposed_comm_ms() after a synchronize."""
        if self.world > 1:
            timing = getattr(self, "time_comm", False) and self.flat_grad.is_cuda
            if timing:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            rf = getattr(self, "_reduced_from", None)
            hi = rf if rf is not None else self.flat_grad.numel()
            if hi > 0:
                dist.all_reduce(self.flat_grad[:hi], op=dist.ReduceOp.SUM)
            for w in getattr(self, "_works", []):
                w.wait()
            self._works = []
            self._reduced_from = None
            if timing:
                e1.record()
                self.__dict__.setdefault("_comm_events", []).append((e0, e1))
        return 1.0 / self.world

    def exposed_comm_ms(self):
        """sum of the bracketed collective waits since the last call (call after torch.cuda.synchronize())"""
        evs = self.__dict__.pop("_comm_events", [])
        return float(sum(a.elapsed_time(b) for a, b in evs))

    def optimizer_step(self):
        self._check_views()
        if self.flat_grad.is_cuda:
            ops.join_wgrad_stream(self.flat_grad.device)      # weight-gradient GEMMs issued on the side stream (if enabled)
        self._zero_untouched()
        scale = self.reduce_gradients()
        st = self.state
        clip = self.max_grad_norm is not None and self.max_grad_norm > 0
        if clip:
            self._norm.zero_()
            ops.sumsq(self.flat_grad, self._norm)        # the clip factor is derived from it INSIDE the AdamW kernel: no read-back
        ops.check_deferred(block=False)                  # errors of sync-free merges surface here once their flag has arrived
        lr = self.current_lr()
        self.step_count += 1
        ops.adamw_flat(st.P, st.LO, st.G, st.M, st.V, st.groups, lr, self.betas[0], self.betas[1], self.eps, self.wd,
                       self.step_count, grad_scale=scale, norm_sq=self._norm if clip else None,
                       max_norm=self.max_grad_norm if clip else 0.0, zero_grad=False)
        self._mark_fresh()

    def grad_norm(self):
        """global gradient norm of the last optimizer step (after the 1/world scale, before clipping); one host read-back"""
        return math.sqrt(float(self._norm.item())) / self.world

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

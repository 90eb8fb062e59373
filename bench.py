#!/usr/bin/env python
"""bench.py -- the driver contract.

  python bench.py --gpus N --steps K --warmup W            our arm  (one process per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU forward/backward (rank 0 only)

Workload (BASELINE.json configs[1] / SURVEY.md section 8d config 2): Mantis-8B-SigLIP-LLaMA-3, random init, bf16, one
data-parallel rank = 4 samples x (8 images 384x384 + 2048 text tokens) per optimizer step (merged S = 7864 per sample,
31,456 merged tokens per step), processed as 4 micro-batches of one sample with gradient accumulation exactly like the
reference recipe (per_device_train_batch_size 1, mantis/train/scripts/train_mllava.sh:137), then one gradient
all-reduce (N > 1) and one fused AdamW step.  A "step" = those 4 forward+backward passes + all-reduce + AdamW.
metric = merged tokens / s over the whole job.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG_TOKEN, N_IMG, T_TEXT, IMG_RES, SAMPLES_PER_STEP = 128256, 8, 2048, 384, 4
FLOP_PER_STEP = 1.634e15          # 408.6 TFLOP/sample * 4 (ViT fwd only, rest x3)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--text-layers", type=int, default=32, help="debug only: fewer layers => INVALID as a bench number")
    ap.add_argument("--vision-layers", type=int, default=27)
    ap.add_argument("--samples", type=int, default=SAMPLES_PER_STEP)
    ap.add_argument("--micro-batch", type=int, default=0,
                    help="samples per forward/backward (0 = workload default: 1 for mllava, whose merged sequence is 7864 "
                         "tokens per sample; 4 for idefics2, whose sequences stay at 2048 tokens)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-gemm", action="store_true", help="only run the per-kernel roofline section")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="SURVEY 8d config 4: fix the GLOBAL batch (32) and split it over the ranks (strong scaling: 32/N "
                         "micro-batches of one sample per rank per optimizer step); default 0 = weak scaling, --samples per rank")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the sub-records of the N=1 line (generate configs[4], idefics2 configs[2], gpu_incumbent)")
    ap.add_argument("--no-hints", action="store_true",
                    help="do not attach the collator's host-side hints (merged length / supervised-row count): the merge and the "
                         "LM-head compaction then read their sizes back from the device (2 host syncs per micro-batch)")
    ap.add_argument("--new-tokens", type=int, default=512, help="generate(): new tokens per sequence (configs[4]: 512)")
    ap.add_argument("--workload", default="mllava", choices=["mllava", "idefics2"],
                    help="mllava = BASELINE configs[1] (the headline); idefics2 = configs[2] (perceiver-resampler path)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ synthetic data
def make_sample(i, torch):
    g = torch.Generator().manual_seed(1234 + i)
    ids = torch.randint(0, 128000, (1, T_TEXT), generator=g)
    for j in range(N_IMG):
        ids[0, j * 256 + 16] = IMG_TOKEN
    labels = ids.clone()
    labels[ids == IMG_TOKEN] = -100
    gp = torch.Generator().manual_seed(4321 + i)
    pv = torch.randn(N_IMG, 3, IMG_RES, IMG_RES, generator=gp).to(torch.bfloat16)
    return dict(input_ids=ids, attention_mask=torch.ones_like(ids), labels=labels, pixel_values=pv)


def make_sample_idefics2(i, torch):
    """configs[2]: 2048 tokens containing 8 runs of <fake><image>x64<fake>; pixel_values [1, 8, 3, 384, 384]"""
    g = torch.Generator().manual_seed(1234 + i)
    ids = torch.randint(3, 32000, (1, T_TEXT), generator=g)
    for j in range(N_IMG):
        s0 = j * 256 + 16
        ids[0, s0] = 32000; ids[0, s0 + 1:s0 + 65] = 32001; ids[0, s0 + 65] = 32000
    labels = ids.clone(); labels[ids == 32001] = 32001
    gp = torch.Generator().manual_seed(4321 + i)
    pv = torch.randn(1, N_IMG, 3, IMG_RES, IMG_RES, generator=gp).to(torch.bfloat16)
    return dict(input_ids=ids, attention_mask=torch.ones_like(ids), labels=labels, pixel_values=pv)


def idefics2_8b_config(text_layers=32, vision_layers=27):
    from transformers import Idefics2Config
    return Idefics2Config(
        vision_config=dict(hidden_size=1152, intermediate_size=4304, num_hidden_layers=vision_layers, num_attention_heads=16,
                           image_size=980, patch_size=14, hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6),
        perceiver_config=dict(resampler_n_latents=64, resampler_depth=3, resampler_n_heads=16, resampler_head_dim=96,
                              num_key_value_heads=4, hidden_act="silu", hidden_size=4096, rms_norm_eps=1e-5),
        text_config=dict(model_type="mistral", hidden_size=4096, intermediate_size=14336, num_hidden_layers=text_layers,
                         num_attention_heads=32, num_key_value_heads=8, vocab_size=32003, rms_norm_eps=1e-5,
                         rope_theta=10000.0, sliding_window=4096, max_position_embeddings=32768, pad_token_id=0),
        image_token_id=32001, tie_word_embeddings=False)


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True); self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = sorted(int(float(r[0])) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_baseline(n_timed=1, n_warm=0):
    """The reference's own forward+backward (mantis/models/mllava/modeling_llava.py:364-549, imported unmodified via
    oracle/ref_shim.py from baseline/_ref or /root/reference) on the host cores.  Bounded sample: full-width
    Mantis-8B-SigLIP at reduced depth (1+1 and 3+3 layers), one sample of 1 image + 256 text tokens (S = 983), fp32;
    per-layer cost from the depth difference (two layers apart and the fastest of the timed iterations, so that one-off
    costs of the first call do not swamp it), extrapolated linearly to 27 ViT + 32 LLaMA layers."""
    import torch
    from oracle.ref_shim import find_ref_root, ref_llava_classes
    if find_ref_root() is None:
        return None
    from transformers import LlamaConfig, SiglipVisionConfig
    LlavaConfig, RefLlava, _ = ref_llava_classes()
    # thread count: "all the host threads it can use" -- more threads than the BLAS scales to make the reference slower (128
    # threads were 3x slower than 8 on the layer GEMMs), so calibrate on the MLP GEMM of this workload and keep the fastest
    ncpu = os.cpu_count() or 1
    a = torch.randn(983, 4096); b = torch.randn(4096, 14336)
    best = (float("inf"), ncpu)
    for nt in sorted({ncpu, max(1, ncpu // 2), max(1, ncpu // 4), min(ncpu, 32), min(ncpu, 16), min(ncpu, 8)}, reverse=True):
        torch.set_num_threads(nt)
        a @ b
        t0 = time.perf_counter()
        for _ in range(3):
            a @ b
        dt = time.perf_counter() - t0
        if dt < best[0]:
            best = (dt, nt)
    cores = best[1]
    torch.set_num_threads(cores)
    del a, b

    def build(depth):
        vc = SiglipVisionConfig(hidden_size=1152, intermediate_size=4304, num_hidden_layers=depth + 1,
                                num_attention_heads=16, image_size=384, patch_size=14, layer_norm_eps=1e-6,
                                hidden_act="gelu_pytorch_tanh")   # +1: hidden_states[-2] drops the last layer
        tc = LlamaConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=depth, num_attention_heads=32,
                         num_key_value_heads=8, vocab_size=128258, rms_norm_eps=1e-5, rope_theta=500000.0)
        cfg = LlavaConfig(vision_config=vc, text_config=tc, image_token_index=IMG_TOKEN, pad_token_id=128257,
                          vocab_size=128258)
        torch.manual_seed(0)
        m = RefLlava(cfg)
        for n, p in m.named_parameters():
            if "vision_tower" in n:
                p.requires_grad_(False)
        return m.train()

    g = torch.Generator().manual_seed(7)
    ids = torch.randint(0, 128000, (1, 256), generator=g); ids[0, 16] = IMG_TOKEN
    labels = ids.clone(); labels[ids == IMG_TOKEN] = -100
    pv = torch.randn(1, 3, IMG_RES, IMG_RES, generator=g)
    times = {}
    for depth in (1, 3):
        m = build(depth)
        ts = []
        warm = n_warm if depth == 1 else 0          # process-wide one-offs (thread pool, oneDNN primitives) happen once
        for it in range(warm + n_timed):
            t0 = time.perf_counter()
            out = m(input_ids=ids, pixel_values=pv, attention_mask=torch.ones_like(ids), labels=labels)
            out.loss.backward()
            m.zero_grad(set_to_none=True)
            if it >= warm:
                ts.append(time.perf_counter() - t0)
        times[depth] = min(ts)
        del m
    per_layer = max((times[3] - times[1]) / 2.0, 1e-9)
    fixed = max(times[1] - per_layer, 0.0)
    full = fixed + 32 * per_layer
    S = 256 + 727
    return {"value": S / full, "unit": "tokens/s", "cores": cores, "kind": "reference",
            "sample": (f"reference fwd+bwd fp32 on {cores} host threads, full-width Mantis-8B-SigLIP at depth 1+1 "
                       f"({times[1]:.2f} s) and 3+3 ({times[3]:.2f} s), 1 image + 256 text tokens (S=983); linear "
                       f"extrapolation to 27+32 layers = {full:.1f} s/sample (short sequence favours the reference)"),
            "seconds_per_sample_extrapolated": full}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    try:
        res = cpu_reference_baseline(n_timed=max(1, min(args.steps, 3)), n_warm=min(args.warmup, 1))
    except Exception as e:  # noqa
        res = None
        err = f"{type(e).__name__}: {e}"
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": locals().get("err", "reference tree not found (baseline/_ref)")}))
        return
    line = {"impl": "reference", "metric": "training tokens/sec Mantis-8B-SigLIP 8-img/2048-tok", "value": res["value"],
            "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["seconds_per_sample_extrapolated"] * SAMPLES_PER_STEP * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Mantis-8B-SigLIP-LLaMA-3 fwd+bwd, reference CPU path, bounded sample (see cpu_baseline.sample)"},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def _latest_profile(pattern):
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    return files[-1] if files else None


def ncu_evidence(kernel, pattern, pick):
    """Per-launch DRAM traffic / tensor-pipe numbers of `kernel` from the newest `tools/ncu_to_json.py` file under profiles/.
    The file carries the hash of the kernel's sources; if the kernel in the tree has changed since the capture the numbers are
    REFUSED (null + a loud note) instead of being quoted for a kernel they no longer describe."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from kernel_hash import kernel_hash
    path = _latest_profile(pattern)
    if path is None:
        return {"traffic": None, "note": f"no profiles/{pattern}"}
    doc = json.load(open(path))
    if doc.get("kernel_hash") != kernel_hash(kernel):
        msg = (f"STALE ncu capture {os.path.relpath(path, ROOT)}: kernel hash {doc.get('kernel_hash')} != tree "
               f"{kernel_hash(kernel)} -- re-run tools/ncu_to_json.py; traffic / tensor-pipe numbers withheld")
        print("bench.py: " + msg, file=sys.stderr)
        return {"traffic": None, "note": msg}
    launch = next((l for l in doc["launches"] if pick in l["kernel"]), doc["launches"][0])
    return {"traffic": launch["traffic_bytes"], "tensor_pipe_active_pct_ncu": launch.get("tensor_pipe_active_pct"),
            "ncu_duration_us": launch.get("duration_us"), "ncu_sm_ghz": launch.get("sm_ghz"),
            "traffic_source": f"{os.path.relpath(path, ROOT)} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, "
                              f"launch {launch['kernel']}; kernel hash {doc['kernel_hash']})"}


def gemm_roofline(torch, ops, peaks):
    """Per-kernel roofline of the dominant kernel (the tcgen05 GEMM): CUDA events on the launching stream, operand sets
    cycled so every launch reads cold-in-L2 data.  `achieved` is the gate/up-projection forward GEMM (M=7864, N=14336,
    K=4096 -- the shape that carries most of the step's FLOPs and the one captured with `ncu --set full`);
    `all_linear_shapes` aggregates fwd/dgrad/wgrad (wgrad into the fp32 main gradient, as the step runs it) of all seven
    linears of a decoder layer."""
    dev = torch.device("cuda")
    M = 7864
    shapes = [(4096, 4096), (1024, 4096), (1024, 4096), (4096, 4096), (14336, 4096), (14336, 4096), (4096, 14336)]
    flops = 0.0; t_ms = 0.0; launches = 0
    nset = 3
    head = None
    for (N, K) in shapes:
        xs = [torch.randn(M, K, device=dev).bfloat16() for _ in range(nset)]
        ws = [(torch.randn(N, K, device=dev) * 0.02).bfloat16() for _ in range(nset)]
        gs = [torch.randn(M, N, device=dev).bfloat16() for _ in range(nset)]
        g32 = [torch.zeros(N, K, device=dev) for _ in range(nset)]
        for kind in ("fwd", "dgrad", "wgrad"):
            def run(i):
                if kind == "fwd":
                    ops.gemm(xs[i], ws[i])
                elif kind == "dgrad":
                    ops.gemm(gs[i], ws[i], trans_a=False, trans_b=False)
                else:
                    ops.gemm(gs[i], xs[i], trans_a=True, trans_b=False, addend=g32[i], out=g32[i])
            for i in range(nset):
                run(i)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            reps = 6
            e0.record()
            for r in range(reps):
                run(r % nset)
            e1.record(); torch.cuda.synchronize()
            dt = e0.elapsed_time(e1)
            t_ms += dt; flops += reps * 2.0 * M * N * K; launches += reps
            if head is None and (N, K) == (14336, 4096) and kind == "fwd":
                head = (2.0 * M * N * K / (dt / reps * 1e-3) / 1e12, dt / reps)
        del xs, ws, gs, g32
    ach_all = flops / (t_ms * 1e-3) / 1e12
    peak = peaks["bf16_tflops"]
    ach, ms = head
    out = {"bound": "tensor", "kernel": "gemm_sm100_2cta_kernel (tcgen05 cta_group::2, 256x256x64 per SM pair), gate/up fwd "
           "M=7864 N=14336 K=4096", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
           "avg_launch_ms": ms, "algorithmic_bytes": 2.0 * (M * 4096 + 14336 * 4096 + M * 14336),
           "all_linear_shapes": {"achieved": ach_all, "frac": ach_all / peak, "launches_timed": launches}}
    out.update(ncu_evidence("gemm_sm100_2cta_kernel", "ncu_gemm2cta_r*.json", "<0, 0>"))
    return out


def scatter_roofline(torch, ops, peaks, B=4):
    """HBM roofline of the image-token scatter (merge_rows_kernel) at BASELINE config 2's batch (B = 4 samples, S = 7864):
    algorithmic bytes = read every source row once + write every output row once = 2 * B * S * D * 2 = 516 MB."""
    dev = torch.device("cuda")
    T, P, D = T_TEXT, 728, 4096
    S = T + N_IMG * (P - 1)
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, 128000, (B, T), generator=g)
    for j in range(N_IMG):
        ids[:, j * 256 + 16] = IMG_TOKEN
    ids = ids.to(dev)
    nset = 3 if B > 1 else 8
    embs = [torch.randn(B, T, D, device=dev).bfloat16() for _ in range(nset)]
    feats = [torch.randn(B * N_IMG, P, D, device=dev).bfloat16() for _ in range(nset)]
    outs = [torch.empty((B, S, D), dtype=torch.bfloat16, device=dev) for _ in range(nset)]
    att = torch.ones_like(ids)
    ws, hdr = ops.merge_plan(ids, embs[0], P, IMG_TOKEN, 128257)
    srcmap = torch.empty((B, S), dtype=torch.int32, device=dev)
    om = torch.empty((B, S), dtype=torch.int64, device=dev); op_ = torch.empty_like(om); ol = torch.empty_like(om)
    ops._call("mb200_merge_index", ops._p(ids), ops._p(att), ops._p(ids), ops._p(ws), B, T, P, S, int(hdr[1]), IMG_TOKEN, -100,
              ops._p(srcmap), ops._p(om), ops._p(ol), ops._p(op_), ops._st())

    def run(i):
        f2 = feats[i % nset].reshape(-1, D)
        ops._call("mb200_merge_rows", ops._p(srcmap), ops._p(embs[i % nset]), ops._p(f2), ops._p(outs[i % nset]), B, S, T, D * 2,
                  f2.shape[0], ops._st())
    for i in range(nset):
        run(i)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 30
    e0.record()
    for i in range(reps):
        run(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bytes_ = 2.0 * B * S * D * 2
    ach = bytes_ / (ms * 1e-3) / 1e9
    out = {"bound": "hbm", "kernel": f"merge_rows_kernel (image-token scatter), B={B} S=7864 D=4096", "achieved": ach,
           "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "avg_launch_ms": ms,
           "algorithmic_bytes": bytes_, "l2": f"{nset} rotating operand sets of {bytes_ / 1e6:.0f} MB (> 126 MB L2)"}
    out.update(ncu_evidence("merge_rows_kernel", "ncu_merge_rows_r*.json", "merge_rows_kernel"))
    return out


def build_model(torch, dev, workload, text_layers, vision_layers):
    from mantis_b200.models.mllava import LlavaForConditionalGeneration, mantis_8b_siglip_llama3_config
    torch.manual_seed(0)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        if workload == "idefics2":
            from mantis_b200.models.idefics2 import Idefics2ForConditionalGeneration
            with torch.device(dev):
                model = Idefics2ForConditionalGeneration(idefics2_8b_config(text_layers, vision_layers))
        else:
            cfg = mantis_8b_siglip_llama3_config(num_vision_layers=vision_layers, num_text_layers=text_layers)
            with torch.device(dev):
                model = LlavaForConditionalGeneration(cfg)
    finally:
        torch.set_default_dtype(old)
    return model


def train_measure(torch, dist, ops, model, workload, args, world, rank, dev, steps, warmup, e2e_too, samples):
    """W warm-up + K timed optimizer steps (device-resident inputs), then optionally K more through the public API with
    host (pinned) inputs and a loss read-back per step.  Returns a dict of raw measurements."""
    from mantis_b200.train import B200Trainer
    model.train()
    mb = args.micro_batch or (4 if workload == "idefics2" else 1)
    mb = max(1, min(mb, samples))
    while samples % mb:
        mb -= 1
    trainer = B200Trainer(model, lr=1e-5, weight_decay=0.0, max_grad_norm=1.0, grad_accum=samples // mb)
    trainer.time_comm = world > 1
    n_train = sum(p.numel() for p in trainer.params)
    mk = make_sample_idefics2 if workload == "idefics2" else make_sample
    host = [mk(rank * samples + i, torch) for i in range(samples)]
    if mb > 1:                                   # same-shape synthetic samples: a micro-batch is a plain concatenation
        host = [{k: torch.cat([s[k] for s in host[i:i + mb]], dim=0) for k in host[0]} for i in range(0, len(host), mb)]
    host = [{k: v.pin_memory() for k, v in s.items()} for s in host]
    # what train.Collator attaches to a batch from the HOST copies of input_ids / labels (merged length, padding side, number of
    # supervised rows): with it neither the image-token merge nor the LM-head row compaction reads anything back from the device
    from mantis_b200.train import llava_valid_rows, plain_valid_rows
    if workload == "idefics2":
        hints = [{"valid_rows": plain_valid_rows(s["labels"], s["attention_mask"], 32001)} for s in host]
    else:
        hints = [{"max_image_tokens": N_IMG, "left_padding": True,
                  "valid_rows": llava_valid_rows(s["input_ids"], s["labels"], s["attention_mask"], IMG_TOKEN, True)} for s in host]
    if args.no_hints:
        hints = [None] * len(host)
    resident = [dict({k: v.to(dev, non_blocking=True) for k, v in s.items()}, **({"merge_hint": h} if h else {}))
                for s, h in zip(host, hints)]
    torch.cuda.synchronize()
    h2d = sum(v.numel() * v.element_size() for s in host for v in s.values())

    def step(e2e):
        if e2e:
            batches = [dict({k: v.to(dev, non_blocking=True) for k, v in s.items()}, **({"merge_hint": h} if h else {}))
                       for s, h in zip(host, hints)]
        else:
            batches = resident
        loss = trainer.train_step(batches)
        return float(loss.item()) if e2e else loss

    def timed(e2e, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        l0 = ops.launch_count
        trainer.exposed_comm_ms()                  # drop the brackets of earlier (warm-up) steps
        e0.record()
        last = None
        for _ in range(k):
            last = step(e2e)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1), trainer.exposed_comm_ms()], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms[0].item(), ops.launch_count - l0, last, ms[1].item() / max(k, 1)

    for _ in range(warmup):
        step(False)
    sampler = ClockSampler(dev.index or 0); sampler.start()
    ms, launches, last_loss, comm_ms = timed(False, steps)
    clocks = sampler.stop()
    res = {"ms": ms, "launches": launches, "loss": float(last_loss), "clocks": clocks, "comm_exposed_ms": comm_ms,
           "mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "n_train": n_train, "micro_batch": mb,
           "grad_accum": samples // mb, "h2d": h2d, "host": host, "e2e_ms": None}
    if e2e_too:
        step(True)
        res["e2e_ms"], _, _, _ = timed(True, steps)
    ops.check_deferred()
    trainer.close()                                # frees G / M / V / LO; the model keeps its (flat) weights
    del trainer, resident
    return res


def generate_record(torch, ops, model, peaks, bs, new_tokens=512):
    """BASELINE configs[4]: 8 images + 256-token prompt (merged prefill S = 6072) -> `new_tokens` greedy tokens THROUGH
    model.generate() (transformers' GenerationMixin, the call mantis/models/mllava/utils.py:88 makes).  Two calls with
    identical inputs: max_new_tokens = 1 (vision tower + merge + prefill + first token) and max_new_tokens = N; decode time is
    their difference over N - 1 steps.  HBM bound of a decode step = 15.01 GB of weights + 131,072 B x context per sequence."""
    from mantis_b200.models import decode_engine
    dev = next(model.parameters()).device
    g = torch.Generator().manual_seed(5 + bs)
    ids = torch.randint(0, 128000, (bs, 256), generator=g)
    for j in range(N_IMG):
        ids[:, j * 32 + 4] = IMG_TOKEN
    pv = torch.randn(bs * N_IMG, 3, IMG_RES, IMG_RES, generator=g).bfloat16().pin_memory()
    ids = ids.pin_memory()
    model.eval()
    S = 256 + N_IMG * 727

    def run(n):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ids_d = ids.to(dev, non_blocking=True)
        out = model.generate(input_ids=ids_d, pixel_values=pv.to(dev, non_blocking=True), attention_mask=torch.ones_like(ids_d),
                             max_new_tokens=n, min_new_tokens=n, do_sample=False, num_beams=1, pad_token_id=128257)
        out = out.cpu()                                   # the user-visible result
        e1.record(); torch.cuda.synchronize()
        assert out.shape == (bs, 256 + n), out.shape
        return e0.elapsed_time(e1) * 1e-3
    run(4)                                                # warm-up (allocations, kernel attributes, page slabs)
    n0 = decode_engine.native_steps
    l0 = ops.launch_count
    t1 = run(1)
    tn = run(new_tokens)
    native = decode_engine.native_steps - n0
    t_dec = (tn - t1) / (new_tokens - 1)
    ctx_avg = S + new_tokens / 2
    bytes_step = 15.01e9 + 131072.0 * ctx_avg * bs
    return {"bs": bs, "prefill_len": S, "new_tokens": new_tokens, "api": "model.generate (GenerationMixin, greedy)",
            "prefill_tok_s": bs * S / t1, "prefill_s": t1, "decode_tok_s": bs / t_dec, "decode_ms_per_step": t_dec * 1e3,
            "e2e_s": tn, "decode_hbm_gbs": bytes_step / t_dec / 1e9, "hbm_bound_bytes_per_step": bytes_step,
            "frac": bytes_step / t_dec / 1e9 / peaks["hbm_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s", "bound": "hbm",
            "native_engine_steps": native, "gpu_launches": ops.launch_count - l0,
            "includes": "H2D of prompt + 8 x 384^2 images per sequence, vision tower, merge, prefill, D2H of the tokens"}


def gpu_incumbent(torch, peaks):
    """Context line (not a target): what a Mantis user gets TODAY on this B200 -- the unmodified reference forward+backward
    (baseline/_ref through oracle/ref_shim.py) in bf16 through PyTorch (cuBLAS GEMMs + SDPA attention), full width at depth 1+1
    and 3+3 on one config-2 sample (8 images, S = 7864), extrapolated linearly to 27+32 layers like the CPU leg."""
    from oracle.ref_shim import find_ref_root, ref_llava_classes
    if find_ref_root() is None:
        return None
    from transformers import LlamaConfig, SiglipVisionConfig
    LlavaConfig, RefLlava, _ = ref_llava_classes()
    dev = torch.device("cuda")
    s = make_sample(0, torch)
    batch = {k: v.to(dev) for k, v in s.items()}
    times = {}
    for depth in (1, 3):
        vc = SiglipVisionConfig(hidden_size=1152, intermediate_size=4304, num_hidden_layers=depth + 1, num_attention_heads=16,
                                image_size=384, patch_size=14, layer_norm_eps=1e-6, hidden_act="gelu_pytorch_tanh")
        tc = LlamaConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=depth, num_attention_heads=32,
                         num_key_value_heads=8, vocab_size=128258, rms_norm_eps=1e-5, rope_theta=500000.0,
                         max_position_embeddings=8192)
        cfg = LlavaConfig(vision_config=vc, text_config=tc, image_token_index=IMG_TOKEN, pad_token_id=128257, vocab_size=128258)
        torch.manual_seed(0)
        m = RefLlava(cfg).to(dev).to(torch.bfloat16).train()
        for n, p in m.named_parameters():
            if "vision_tower" in n:
                p.requires_grad_(False)
        ts = []
        for it in range(4):
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            out = m(**batch)
            out.loss.backward()
            m.zero_grad(set_to_none=True)
            e1.record(); torch.cuda.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1) * 1e-3)
        times[depth] = min(ts)
        del m, out
        torch.cuda.empty_cache()
    per_layer = max((times[3] - times[1]) / 2.0, 1e-9)
    full = max(times[1] - per_layer, 0.0) + 32 * per_layer       # ViT depth scales with the same index (27 vs 32: upper bound)
    S = T_TEXT + N_IMG * 727
    return {"value": S / full, "unit": "tokens/s", "kind": "reference on GPU (torch eager: cuBLAS + SDPA), bf16, no optimizer step",
            "seconds_per_sample_extrapolated": full,
            "sample": f"full-width Mantis-8B-SigLIP reference fwd+bwd, depth 1+1 ({times[1]:.3f} s) and 3+3 ({times[3]:.3f} s), "
                      f"one config-2 sample (S = {S}); linear extrapolation to 27+32 layers"}


def run_ours(args):
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"        # keep NCCL's version banner off stdout: rank 0 prints ONE line, the JSON
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from mantis_b200 import _lib, ops
    assert _lib.lib().mb200_check_device() == 0, _lib.lib().mb200_last_error()
    peaks, peaks_src = measured_peaks()

    if args.profile_gemm:
        print(json.dumps(gemm_roofline(torch, ops, peaks)))
        return

    samples = args.samples
    scaling = "weak"
    if args.global_batch:                        # the GLOBAL batch is fixed, ranks split it
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} ranks")
        samples = args.global_batch // world
        scaling = "strong"
    model = build_model(torch, dev, args.workload, args.text_layers, args.vision_layers)
    r = train_measure(torch, dist, ops, model, args.workload, args, world, rank, dev, args.steps, args.warmup,
                      not args.no_e2e, samples)
    S_merged = T_TEXT if args.workload == "idefics2" else T_TEXT + N_IMG * 727
    tokens_per_step_rank = samples * S_merged
    ms = r["ms"]
    full = (args.text_layers == 32 and args.vision_layers == 27)
    extras = world == 1 and full and not args.no_extras
    gen = None
    if extras and args.workload == "mllava":
        # BASELINE configs[4] on the SAME weights: generate() prefill / decode tok/s (the trainer's state was released above)
        torch.cuda.empty_cache()
        try:
            gen = {f"bs{bs}": generate_record(torch, ops, model, peaks, bs, args.new_tokens) for bs in (1, 16)}
        except Exception as e:  # noqa
            gen = {"error": f"{type(e).__name__}: {e}"}
    # the per-kernel rooflines are measured on EVERY rank's own GPU (no collective; rank 0's goes into the line) so that
    # the N > 1 lines of the scaling run carry them too
    roof = gemm_roofline(torch, ops, peaks) if (full or world == 1) else None
    scat = scatter_roofline(torch, ops, peaks) if (full or world == 1) else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    e2e = None
    if r["e2e_ms"] is not None:
        e2e = {"value": world * tokens_per_step_rank * args.steps / (r["e2e_ms"] * 1e-3), "unit": "tokens/s",
               "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 4}
    value = world * tokens_per_step_rank * args.steps / (ms * 1e-3)
    per_sample_flop = (449e12 if args.workload == "idefics2" else FLOP_PER_STEP) / SAMPLES_PER_STEP
    # executed FLOPs: the fused LM-head/CE skips rows whose label is ignored (no loss, no gradient): 3 x 2*V*D per row
    ign = 32001 if args.workload == "idefics2" else -100
    valid_rows = sum(int((s_["labels"][:, 1:] != ign).sum()) for s_ in r["host"])
    V_ = 32003 if args.workload == "idefics2" else 128258
    skipped = (samples * S_merged - valid_rows) * 6.0 * V_ * 4096 if ops.LM_HEAD_SKIP_IGNORED else 0.0
    step_tflops = world * per_sample_flop * samples * args.steps / (ms * 1e-3) / 1e12
    line = {
        "metric": ("training tokens/sec Mantis-8B-SigLIP 8-img/2048-tok" if args.workload == "mllava"
                   else "training tokens/sec Mantis-8B-Idefics2 8-img/2048-tok"), "value": value, "unit": "tokens/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": ("Mantis-8B-SigLIP-LLaMA-3 instruction-tuning step, random init (configs[1])"
                                if args.workload == "mllava" else
                                "Mantis-8B-Idefics2 instruction-tuning step, random init (configs[2])"),
                   "samples_per_rank_per_step": samples, "global_batch": samples * world, "images_per_sample": N_IMG,
                   "text_tokens": T_TEXT, "merged_seq_len": S_merged, "micro_batch": r["micro_batch"],
                   "grad_accum": r["grad_accum"], "parallelism": f"dp{world}",
                   "optimizer": "ONE fused AdamW launch over flat buffers: fp32 master weights (bf16 weight + 16 low bits), "
                                "fp32 main-gradient accumulation, fp32 moments, device-side global-norm clip",
                   "trainable_params": r["n_train"], "vision_tower": "frozen (train_mllava.py:239-242)",
                   "l2": "working set >> L2 (35 GB of activations + 16 GB weights per micro-batch), no flush needed",
                   "text_layers": args.text_layers, "vision_layers": args.vision_layers,
                   "valid": full and (args.global_batch or samples == SAMPLES_PER_STEP)},
        "step_tflops": step_tflops, "step_frac_of_sustained_peak": step_tflops / world / peaks["bf16_tflops_sustained"],
        "step_tflops_executed": world * (per_sample_flop * samples - skipped) * args.steps / (ms * 1e-3) / 1e12,
        "flop_note": ("step_tflops uses the nominal 1.634 PFLOP/step of SURVEY 8d (full-sequence LM head, as the reference "
                      "computes it); step_tflops_executed subtracts the LM-head rows with ignored labels that the fused "
                      "LM-head/CE provably skips (identical loss and gradients)"),
        "peaks": peaks_src, "gpu_launches": r["launches"], "max_mem_gb": r["mem_gb"], "clocks": r["clocks"],
        "loss": r["loss"], "comm_exposed_ms": r["comm_exposed_ms"] if world > 1 else 0.0,
        "e2e": e2e, "roofline": roof, "roofline_scatter": scat, "generate": gen,
    }
    if extras and args.workload == "mllava":
        # BASELINE configs[2] (Idefics2: NaViT tower + perceiver resampler + Mistral-7B) as a sub-record of the same line
        del model
        torch.cuda.empty_cache()
        try:
            m2 = build_model(torch, dev, "idefics2", 32, 27)
            r2 = train_measure(torch, dist, ops, m2, "idefics2", args, 1, 0, dev, max(2, args.steps // 2), 2, False,
                               SAMPLES_PER_STEP)
            k2 = max(2, args.steps // 2)
            v2 = SAMPLES_PER_STEP * T_TEXT * k2 / (r2["ms"] * 1e-3)
            tf2 = 449e12 * k2 / (r2["ms"] * 1e-3) / 1e12
            line["idefics2"] = {"metric": "training tokens/sec Mantis-8B-Idefics2 8-img/2048-tok (configs[2])", "value": v2,
                                "unit": "tokens/s", "ms_per_step": r2["ms"] / k2, "steps": k2, "warmup": 2,
                                "step_tflops": tf2, "step_frac_of_sustained_peak": tf2 / peaks["bf16_tflops_sustained"],
                                "micro_batch": r2["micro_batch"], "gpu_launches": r2["launches"], "max_mem_gb": r2["mem_gb"],
                                "loss": r2["loss"], "trainable_params": r2["n_train"]}
            del m2, r2
        except Exception as e:  # noqa
            line["idefics2"] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
        try:
            line["gpu_incumbent"] = gpu_incumbent(torch, peaks)
        except Exception as e:  # noqa
            line["gpu_incumbent"] = {"error": f"{type(e).__name__}: {e}"}
    if not args.no_cpu_baseline and world == 1:
        try:
            cb = cpu_reference_baseline(1, 1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")} if cb else None
        except Exception as e:  # noqa
            line["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

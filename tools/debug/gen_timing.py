import sys, os, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mantis_b200 import ops
from mantis_b200.models.kv_cache import B200KVCache
from mantis_b200.models.decode_engine import greedy_decode_loop
from mantis_b200.models.mllava import LlavaForConditionalGeneration, mantis_8b_siglip_llama3_config
dev = torch.device("cuda")
cfg = mantis_8b_siglip_llama3_config()
torch.manual_seed(0); torch.set_default_dtype(torch.bfloat16)
with torch.device(dev):
    model = LlavaForConditionalGeneration(cfg)
torch.set_default_dtype(torch.float32); model.eval()
g = torch.Generator().manual_seed(5)
ids = torch.randint(0, 128000, (1, 256), generator=g)
for j in range(8):
    ids[:, j * 32 + 4] = 128256
pv = torch.randn(8, 3, 384, 384, generator=g).bfloat16().to(dev); ids = ids.to(dev); am = torch.ones_like(ids)
with torch.no_grad():
    for n in (32, 128, 512):
        for rep in range(2):
            cache = B200KVCache()
            out = model(input_ids=ids, pixel_values=pv, attention_mask=am, past_key_values=cache, use_cache=True, logits_to_keep=1)
            first = out.logits[:, -1].argmax(-1)
            S = cache.get_seq_length()
            pos0 = torch.full((1,), S, dtype=torch.int64, device=dev)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record()
            toks = greedy_decode_loop(model.language_model.model, model.language_model.lm_head, cache, first, pos0, n)
            e1.record(); t_issue = time.perf_counter() - t0
            torch.cuda.synchronize(); t_all = time.perf_counter() - t0
        print(f"loop n={n}: device {e0.elapsed_time(e1)/n:.3f} ms/step, host issue {t_issue/n*1e3:.3f} ms/step, wall {t_all/n*1e3:.3f}", flush=True)
    # phases of generate() itself
    import mantis_b200.models.decode_engine as de
    orig = de.greedy_decode_loop
    def timed_loop(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = orig(*a, **k)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"   inside generate: loop issue {(t1 - t0) * 1e3:.1f} ms, +drain {(t2 - t1) * 1e3:.1f} ms, n_steps {a[5]}", flush=True)
        return r
    de.greedy_decode_loop = timed_loop
    for n in (1, 128, 512):
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            o = model.generate(input_ids=ids, pixel_values=pv, attention_mask=am, max_new_tokens=n, min_new_tokens=n, do_sample=False, num_beams=1, pad_token_id=128257)
            o = o.cpu(); t = time.perf_counter() - t0
        print(f"generate n={n}: {t*1e3:.1f} ms", flush=True)

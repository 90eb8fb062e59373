import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_fixture
from transformers import Idefics2Config
from mantis_b200.models.idefics2 import Idefics2ForConditionalGeneration
from mantis_b200.models.kv_cache import B200KVCache
fx = load_fixture("idefics2_sliding.pt")
dev = torch.device("cuda")
model = Idefics2ForConditionalGeneration(Idefics2Config(**fx["cfg"]))
model.load_state_dict(fx["state_dict"]); model = model.to(dev).eval()
ids = fx["inputs"]["input_ids"].to(dev); pv = fx["inputs"]["pixel_values"].to(dev)
gen = fx["generated"].to(dev)
with torch.no_grad():
    # cache-free over 41, 42 tokens
    for n in (40, 41, 42):
        seq = gen[:, :n]
        lg = model(input_ids=seq, attention_mask=torch.ones_like(seq), pixel_values=pv).logits
        print("cache-free", n, "argmax", lg[0, -1].argmax().item(), "expect", gen[0, n].item())
    cache = B200KVCache()
    out = model(input_ids=ids, attention_mask=torch.ones_like(ids), pixel_values=pv, past_key_values=cache, use_cache=True)
    print("prefill argmax", out.logits[0, -1].argmax().item(), cache.get_seq_length())
    am = torch.ones_like(ids)
    for n in (40, 41):
        tok = gen[:, n:n + 1]
        am = torch.cat([am, torch.ones_like(tok)], 1)
        pos = torch.tensor([[n]], device=dev)
        out2 = model(input_ids=tok, attention_mask=am, position_ids=pos, past_key_values=cache, use_cache=True,
                     image_hidden_states=out.image_hidden_states)
        seq = gen[:, :n + 1]
        ref = model(input_ids=seq, attention_mask=torch.ones_like(seq), pixel_values=pv).logits[0, -1]
        print("cached step", n, "argmax", out2.logits[0, -1].argmax().item(), "expect", gen[0, n + 1].item(),
              "max|diff| vs cache-free", (out2.logits[0, -1] - ref).abs().max().item(), "scale", ref.abs().max().item())

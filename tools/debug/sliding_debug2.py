import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_fixture
from transformers import Idefics2Config
from mantis_b200.models.idefics2 import Idefics2ForConditionalGeneration
fx = load_fixture("idefics2_sliding.pt")
dev = torch.device("cuda")
model = Idefics2ForConditionalGeneration(Idefics2Config(**fx["cfg"]))
model.load_state_dict(fx["state_dict"]); model = model.to(dev).eval()
ids = fx["inputs"]["input_ids"].to(dev); pv = fx["inputs"]["pixel_values"].to(dev)
def pre(mod, args, k):
    am = k.get("attention_mask"); pid = k.get("position_ids"); pk = k.get("past_key_values")
    print("forward: ids", None if k.get("input_ids") is None else tuple(k["input_ids"].shape),
          "am", None if am is None else (tuple(am.shape), int(am.sum())), "pos", None if pid is None else pid.tolist()[0][-3:],
          "cache", type(pk).__name__, (pk.get_seq_length() if pk is not None else None),
          "other", {x: (v if isinstance(v, (int, bool, type(None))) else type(v).__name__) for x, v in k.items()
                    if x not in ("input_ids", "attention_mask", "position_ids", "past_key_values", "pixel_values", "image_hidden_states")})
def post(mod, args, k, out):
    print("   -> argmax", out.logits[0, -1].argmax().item())
model.register_forward_pre_hook(pre, with_kwargs=True)
model.register_forward_hook(post, with_kwargs=True)
def tpre(mod, args, k):
    am = k.get("attention_mask")
    print("   text_model: embeds", tuple(k["inputs_embeds"].shape) if k.get("inputs_embeds") is not None else None, "am",
          None if am is None else (tuple(am.shape), int(am.sum())), "pos", None if k.get("position_ids") is None else k["position_ids"].tolist()[0][-2:])
model.model.text_model.register_forward_pre_hook(tpre, with_kwargs=True)
gen = model.generate(input_ids=ids, attention_mask=torch.ones_like(ids), pixel_values=pv, max_new_tokens=4, do_sample=False, num_beams=1)
print(gen[0, 38:].tolist(), "expect", fx["generated"][0, 38:44].tolist())

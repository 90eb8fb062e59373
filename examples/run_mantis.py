"""Same script as the reference's examples/run_mantis.py, with the import line swapped (or left untouched when the repo
root is on PYTHONPATH, thanks to the `mantis/` alias package).  Needs the released weights on disk (no network here)."""
import sys

import torch
from PIL import Image

from mantis.models.mllava import LlavaForConditionalGeneration, MLlavaProcessor, chat_mllava   # == mantis_b200.models.mllava

path = sys.argv[1] if len(sys.argv) > 1 else "TIGER-Lab/Mantis-8B-siglip-llama3"
processor = MLlavaProcessor.from_pretrained(path)
model = LlavaForConditionalGeneration.from_pretrained(path, device_map="cuda", torch_dtype=torch.bfloat16)

generation_kwargs = {"max_new_tokens": 1024, "num_beams": 1, "do_sample": False}
images = [Image.open(p) for p in sys.argv[2:4]] if len(sys.argv) > 3 else []
text = "Describe the difference of <image> and <image> as much as you can."
response, history = chat_mllava(text, images, model, processor, **generation_kwargs)
print("USER:", text)
print("ASSISTANT:", response)

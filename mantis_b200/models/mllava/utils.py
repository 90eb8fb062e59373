"""chat_mllava: one-call multi-image chat on top of generate().

Behavioural mirror of mantis/models/mllava/utils.py:10-97 (template choice by language-model name, history bookkeeping,
image loading, prompt -> processor -> generate -> decode), written independently for this package.
"""
from typing import List, Optional, Tuple, Union

import torch

from ..conversation import conv_mllava_v1, conv_templates


def _pick_template(model, processor):
    """llama-3 checkpoints use the llama_3 template and stop on <|eot_id|> as well as EOS; everything else mllava_v1."""
    lm_name = (getattr(model.language_model, "name_or_path", "") or "").lower()
    if "llama-3" not in lm_name:
        return conv_mllava_v1.copy(), None
    tok = processor.tokenizer
    return conv_templates["llama_3"].copy(), [tok.eos_token_id, tok.convert_tokens_to_ids("<|eot_id|>")]


def _extend_dialogue(conv, text: str, history: Optional[List[dict]]):
    """Replays `history` into `conv`, appends the new user turn (if any) and an empty assistant turn; returns history."""
    user, assistant = conv.roles
    conv.messages = []
    if history is None:
        conv.append_message(user, text)
        conv.append_message(assistant, "")
        return [{"role": user, "text": text}, {"role": assistant, "text": ""}]
    for turn in history:
        if turn["role"] not in conv.roles:
            raise AssertionError(f"unknown role {turn['role']!r} in history")
        conv.append_message(turn["role"], turn["text"])
    last_role, last_text = conv.messages[-1]
    if text:
        if last_role != assistant:
            raise AssertionError("The last message in the history should be the assistant, if the given text is not empty")
        for role, msg in ((user, text), (assistant, "")):
            conv.append_message(role, msg)
            history.append({"role": role, "text": msg})
    elif last_role == assistant:
        if last_text != "":
            raise AssertionError("No user message should be provided")
    else:
        if last_role != user:
            raise AssertionError("The last message in the history should be the user, if the given text is empty")
        conv.append_message(user, "")
        history.append({"role": user, "text": ""})
    return history


def _to_device(inputs, device):
    for key in list(inputs.keys()):
        val = inputs[key]
        if val is None:
            continue
        if torch.is_tensor(val):
            inputs[key] = val.to(device)
        elif isinstance(val, list):
            inputs[key] = [t.to(device) for t in val]
        else:
            raise ValueError(f"Invalid input type: {type(val)}")
    return inputs


def chat_mllava(text: str, images: List[Union["PIL.Image.Image", str]], model, processor, max_input_length: int = None,
                history: List[dict] = None, **kwargs) -> Tuple[str, List[dict]]:
    """Returns (generated_text, history); `history` items are {"role": ..., "text": ...} like the reference's."""
    conv, terminators = _pick_template(model, processor)
    kwargs["eos_token_id"] = terminators
    history = _extend_dialogue(conv, text, history)
    tail_role, tail_text = conv.messages[-1]
    assert tail_role == conv.roles[1] and tail_text == "", "Format check"
    if images:
        import PIL.Image
        images[:] = [PIL.Image.open(im).convert("RGB") if isinstance(im, str) else im for im in images]
    inputs = processor(images=images, text=conv.get_prompt(), return_tensors="pt", truncation=True,
                       max_length=max_input_length)
    inputs = _to_device(inputs, model.device)
    prompt_len = inputs["input_ids"].shape[-1]
    new_ids = model.generate(**inputs, **kwargs)[0][prompt_len:]
    reply = processor.decode(new_ids, skip_special_tokens=True)
    history[-1]["text"] = reply
    return reply, history

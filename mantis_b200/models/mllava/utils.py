"""chat_mllava / chat_mllava_stream: one-call multi-image chat on top of generate().

Behavioural mirror of mantis/models/mllava/utils.py:10-97 and :100-186 (template choice by language-model name, history
bookkeeping, image loading, prompt -> processor -> generate -> decode; the streaming variant runs generate() on a worker
thread and yields the growing reply), written independently for this package.
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
    inputs, history = _prepare_chat(text, images, model, processor, max_input_length, history, kwargs, convert_rgb=True)
    prompt_len = inputs["input_ids"].shape[-1]
    new_ids = model.generate(**inputs, **kwargs)[0][prompt_len:]
    reply = processor.decode(new_ids, skip_special_tokens=True)
    history[-1]["text"] = reply
    return reply, history


def _prepare_chat(text, images, model, processor, max_input_length, history, kwargs, convert_rgb):
    conv, terminators = _pick_template(model, processor)
    kwargs["eos_token_id"] = terminators
    history = _extend_dialogue(conv, text, history)
    tail_role, tail_text = conv.messages[-1]
    assert tail_role == conv.roles[1] and tail_text == "", "Format check"
    if images:
        import PIL.Image
        if convert_rgb:
            images[:] = [PIL.Image.open(im).convert("RGB") if isinstance(im, str) else im for im in images]
        else:                                   # the streaming variant opens files without converting (ref :166-169)
            images[:] = [PIL.Image.open(im) if isinstance(im, str) else im for im in images]
    inputs = processor(images=images, text=conv.get_prompt(), return_tensors="pt", truncation=True,
                       max_length=max_input_length)
    return _to_device(inputs, model.device), history


def chat_mllava_stream(text: str, images: List[Union["PIL.Image.Image", str]], model, processor,
                       max_input_length: int = None, history: List[dict] = None, **kwargs):
    """Generator version of chat_mllava (ref: mantis/models/mllava/utils.py:100-186): `model.generate` runs on a worker
    thread feeding a transformers.TextIteratorStreamer; every decoded piece is appended to the last history entry and
    (reply_so_far, history) is yielded.

    The worker thread drives the CUDA kernels through the C ABI: the library keeps no per-thread state that changes results
    (one-time kernel attributes are set under C++ static-initialisation guards, the launch mode of the decode engine is scoped
    to the call), but a new thread starts on device 0 and on the default stream -- so the worker first binds the model's device
    and orders itself after the caller's stream."""
    from threading import Thread

    from transformers import TextIteratorStreamer
    inputs, history = _prepare_chat(text, images, model, processor, max_input_length, history, kwargs, convert_rgb=False)
    streamer = TextIteratorStreamer(processor, skip_prompt=True, skip_special_tokens=True)
    kwargs["streamer"] = streamer
    inputs.update(kwargs)
    device = model.device
    ready = torch.cuda.Event() if device.type == "cuda" else None
    if ready is not None:
        ready.record(torch.cuda.current_stream(device))
    failure = []

    def work():
        try:
            if ready is not None:
                torch.cuda.set_device(device)
                torch.cuda.current_stream(device).wait_event(ready)
            model.generate(**inputs)
        except BaseException as e:             # surface worker failures to the consumer instead of hanging the iterator
            failure.append(e)
            streamer.end()

    thread = Thread(target=work, daemon=True)
    thread.start()
    for piece in streamer:
        history[-1]["text"] += piece
        yield history[-1]["text"], history
    thread.join()
    if failure:
        raise failure[0]

"""CPU tests of the caller-side helpers: placeholder balancing / image tags of MLlavaProcessor and the conversation
templates (expected strings derived from mantis/models/mllava/processing_llava.py:84-155 and conversation.py:43-145)."""
import torch

from mantis_b200.models.conversation import conv_templates
from mantis_b200.models.mllava import MLlavaProcessor


class _Img:
    size = (4, 4)


class _Tok:
    model_input_names = ["input_ids", "attention_mask"]

    def convert_tokens_to_ids(self, t):
        return 7

    def __call__(self, text, return_tensors=None, padding=False, truncation=None, max_length=None):
        ids = [[7 if w == "<image>" else 1 for w in t.replace("<image>", " <image> ").split()] for t in text]
        if max_length:
            ids = [i[:max_length] for i in ids]
        L = max(len(i) for i in ids)
        ids = [i + [0] * (L - len(i)) for i in ids]
        t = torch.tensor(ids)
        return {"input_ids": t, "attention_mask": (t != 0).long()}


class _IP:
    model_input_names = ["pixel_values"]

    def __call__(self, images, return_tensors=None):
        return {"pixel_values": torch.zeros(len(images), 3, 4, 4)}


def test_placeholder_balancing_and_tags():
    p = MLlavaProcessor(_IP(), _Tok())
    texts, images = p.preprocess_interleaved_images_and_text("USER: compare <image> please", [_Img(), _Img()])
    assert texts == ["USER:(image 1: <Image><image></Image>) compare (image 2: <Image><image></Image>) please"]
    texts, _ = p.preprocess_interleaved_images_and_text("a <image> b <image> c <image>", [_Img()])
    assert texts == ["a (image 1: <Image><image></Image>) b  c "]
    texts, _ = p.preprocess_interleaved_images_and_text("no tag", [_Img()])
    assert texts == ["(image 1: <Image><image></Image>)no tag"]


def test_call_truncation_drops_images_and_collate_contract():
    p = MLlavaProcessor(_IP(), _Tok())
    out = p(text="x <image> y <image> z", images=[_Img(), _Img()], truncation=True, max_length=6)
    assert int((out["input_ids"] == 7).sum()) == out["pixel_values"].shape[0]
    col = p._right_pad_inputs_with_attention_mask([dict(out)])
    assert isinstance(col["pixel_values"], list) and col["input_ids"].shape[0] == 1


def test_conversation_templates():
    c = conv_templates["llama_3"].copy(); c.messages = []
    c.append_message("user", "hi <image>"); c.append_message("assistant", "")
    assert c.get_prompt().endswith("<|eot_id|><|start_header_id|>user<|end_header_id|>\n\nhi <image><|eot_id|>"
                                   "<|start_header_id|>assistant<|end_header_id|>\n\n")
    c = conv_templates["mllava_v1"].copy(); c.messages = []
    c.append_message("USER", "q"); c.append_message("ASSISTANT", "")
    assert c.get_prompt().endswith("</s>USER: q</s>ASSISTANT:")
    c = conv_templates["idefics_2"].copy(); c.messages = []
    c.append_message("User", "q"); c.append_message("Assistant", "")
    assert c.get_prompt() == "User:q<end_of_utterance>\nAssistant:"


def test_chat_history_bookkeeping():
    from mantis_b200.models.mllava.utils import _extend_dialogue
    c = conv_templates["mllava_v1"].copy()
    h = _extend_dialogue(c, "hello <image>", None)
    assert h == [{"role": "USER", "text": "hello <image>"}, {"role": "ASSISTANT", "text": ""}]
    assert c.messages[-1] == ["ASSISTANT", ""]
    h[-1]["text"] = "hi"
    c2 = conv_templates["mllava_v1"].copy()
    h2 = _extend_dialogue(c2, "and now?", h)
    assert [t["role"] for t in h2] == ["USER", "ASSISTANT", "USER", "ASSISTANT"] and h2[-1]["text"] == ""
    assert c2.get_prompt().endswith("USER: and now?</s>ASSISTANT:")
    import pytest
    with pytest.raises(AssertionError):
        _extend_dialogue(conv_templates["mllava_v1"].copy(), "x", [{"role": "USER", "text": "q"}])


def test_collator_single_and_batched():
    from mantis_b200.train.data import Collator
    p = MLlavaProcessor(_IP(), _Tok())
    one = dict(p(text="x <image> y", images=[_Img()]))
    one["labels"] = one["input_ids"].clone()
    col = Collator(p)([one])
    assert isinstance(col["pixel_values"], list) and col["input_ids"].shape[0] == 1      # reference batch==1 contract
    two = dict(p(text="a b c d <image> e", images=[_Img()]))
    two["labels"] = two["input_ids"].clone()
    out = Collator(p, pad_token_id=0)([one, two])
    L = max(one["input_ids"].shape[1], two["input_ids"].shape[1])
    assert out["input_ids"].shape == (2, L) and out["attention_mask"].shape == (2, L) and out["labels"].shape == (2, L)
    short = 0 if one["input_ids"].shape[1] < L else 1
    assert out["attention_mask"][short, -1] == 0 and out["labels"][short, -1] == -100 and out["input_ids"][short, -1] == 0
    assert len(out["pixel_values"]) == 2


def test_processor_batches_more_than_one_sample():
    """the reference asserts batch == 1 in _right_pad_inputs_with_attention_mask (processing_llava.py:279); ours
    right-pads real batches the way the reference's training collator does"""
    import torch
    from mantis_b200.models.mllava import MLlavaProcessor

    class Tok:
        pad_token_id = 7
    proc = MLlavaProcessor(image_processor=None, tokenizer=Tok())
    a = {"input_ids": torch.tensor([[1, 2, 3, 4, 5]]), "attention_mask": torch.ones(1, 5, dtype=torch.long),
         "labels": torch.tensor([[1, 2, 3, 4, 5]]), "pixel_values": torch.zeros(2, 3, 4, 4)}
    b = {"input_ids": torch.tensor([[9, 8]]), "attention_mask": torch.ones(1, 2, dtype=torch.long),
         "labels": torch.tensor([[9, 8]]), "pixel_values": None}
    one = proc._right_pad_inputs_with_attention_mask([a])
    assert torch.equal(one["input_ids"], a["input_ids"]) and isinstance(one["pixel_values"], list) and len(one["pixel_values"]) == 1
    two = proc._right_pad_inputs_with_attention_mask([a, b])
    assert two["input_ids"].tolist() == [[1, 2, 3, 4, 5], [9, 8, 7, 7, 7]]
    assert two["attention_mask"].tolist() == [[1, 1, 1, 1, 1], [1, 1, 0, 0, 0]]
    assert two["labels"].tolist() == [[1, 2, 3, 4, 5], [9, 8, -100, -100, -100]]
    assert isinstance(two["pixel_values"], list) and two["pixel_values"][1] is None


def test_processor_matches_the_live_reference():
    """Where the reference tree is present: MLlavaProcessor (placeholder balancing, "(image j: <Image><image></Image>)" tags,
    tokenisation, dropping images whose placeholders were truncated away) against the UNMODIFIED reference processor
    (mantis/models/mllava/processing_llava.py:66-252) on random interleaved prompts, with the same stub tokenizer / image
    processor on both sides."""
    import random
    import pytest
    from PIL import Image
    from oracle.ref_shim import find_ref_root
    if find_ref_root() is None:
        pytest.skip("reference tree not available here")
    from oracle.ref_shim import load_reference_processor
    RefProc = load_reference_processor().MLlavaProcessor

    class IP(_IP):
        def __call__(self, images, return_tensors=None):
            if not images:
                return {"pixel_values": torch.zeros(0, 3, 2, 2)}
            return {"pixel_values": torch.stack([torch.full((3, 2, 2), float(im.size[0])) for im in images])}

    def make(cls):
        p = object.__new__(cls)                      # ProcessorMixin.__init__ type-checks real tokenizers; not needed here
        p.image_processor, p.tokenizer, p.image_token_index = IP(), _Tok(), None
        return p
    ref, ours = make(RefProc), MLlavaProcessor(IP(), _Tok())
    rnd = random.Random(5)
    words = ["USER:", "Human:", "HUMAN:", "look", "at", "<image>", "and", "<image>", "tell", "me", "ASSISTANT:"]
    n = 0
    for _ in range(200):
        n_img = rnd.randint(1, 4)
        imgs = [Image.new("RGB", (3 + i, 5)) for i in range(n_img)]          # width identifies the image downstream
        text = " ".join(rnd.choice(words) for _ in range(rnd.randint(1, 12)))
        kw = dict(truncation=True, max_length=rnd.randint(3, 40)) if rnd.random() < 0.5 else {}
        for batched in (False, True):
            t_in = [text, text + " again <image>"] if batched else text
            i_in = [list(imgs), list(imgs)] if batched else list(imgs)
            a_t, a_i = ref.preprocess_interleaved_images_and_text(t_in if not batched else list(t_in), [list(x) for x in i_in] if batched else list(i_in))
            b_t, b_i = ours.preprocess_interleaved_images_and_text(t_in if not batched else list(t_in), [list(x) for x in i_in] if batched else list(i_in))
            assert a_t == b_t and [[im.size for im in g] for g in a_i] == [[im.size for im in g] for g in b_i]
            a = ref(text=t_in if not batched else list(t_in), images=[list(x) for x in i_in] if batched else list(i_in), **kw)
            b = ours(text=t_in if not batched else list(t_in), images=[list(x) for x in i_in] if batched else list(i_in), **kw)
            assert torch.equal(a["input_ids"], b["input_ids"]) and torch.equal(a["attention_mask"], b["attention_mask"])
            assert torch.equal(a["pixel_values"], b["pixel_values"])
            n += 1
    assert n == 400


def test_conversation_templates_match_the_live_reference():
    """every template this package ships renders exactly like mantis/models/conversation.py for random dialogues
    (incl. the empty trailing assistant turn that opens a generation and a non-empty system prompt)"""
    import os
    import random
    import pytest
    from oracle import ref_shim
    root = ref_shim.find_ref_root()
    if root is None:
        pytest.skip("reference tree not available here")
    ref = ref_shim._load("_mantis_ref_conversation", os.path.join(root, "mantis", "models", "conversation.py"))
    rnd = random.Random(3)
    vocab = ["hi", "<image>", "what is this?", "a cat", "", "compare (image 1: <Image><image></Image>) please", "ok\nfine"]
    for name, tmpl in conv_templates.items():
        assert name in ref.conv_templates
        for trial in range(50):
            a, b = tmpl.copy(), ref.conv_templates[name].copy()
            a.messages, b.messages = [], []
            if trial % 5 == 0:
                a.system = b.system = "SYSTEM PROMPT"
            for turn in range(rnd.randint(1, 6)):
                msg = rnd.choice(vocab)
                a.append_message(a.roles[turn % 2], msg); b.append_message(b.roles[turn % 2], msg)
            if rnd.random() < 0.5:
                a.append_message(a.roles[1], None); b.append_message(b.roles[1], None)
            assert a.get_prompt() == b.get_prompt(), (name, a.messages)


def test_chat_mllava_matches_the_live_reference():
    """chat_mllava (template choice by language-model name, history bookkeeping, prompt, terminators, decoding of the new
    tokens only) vs mantis/models/mllava/utils.py:10-97 with recording stubs for the model and the processor"""
    import copy
    import random
    import pytest
    from oracle.ref_shim import find_ref_root
    if find_ref_root() is None:
        pytest.skip("reference tree not available here")
    from oracle.ref_shim import load_reference_chat_utils
    from mantis_b200.models.mllava import chat_mllava
    ref_chat = load_reference_chat_utils().chat_mllava

    class Tok:
        eos_token_id = 2

        def convert_tokens_to_ids(self, t):
            return 128009

    class Proc:
        def __init__(self):
            self.tokenizer, self.calls = Tok(), []

        def __call__(self, images=None, text=None, **kw):
            self.calls.append((text, None if images is None else len(images), kw))
            return {"input_ids": torch.arange(len(text.split()))[None], "pixel_values": None if not images else [torch.zeros(1)]}

        def decode(self, ids, skip_special_tokens=True):
            return "reply:" + ",".join(str(int(i)) for i in ids)

    class LM:
        def __init__(self, name):
            self.name_or_path = name

    class Model:
        device = torch.device("cpu")

        def __init__(self, name):
            self.language_model, self.calls = LM(name), []

        def generate(self, **kw):
            self.calls.append({k: (v if not torch.is_tensor(v) else v.clone()) for k, v in kw.items()})
            return torch.cat([kw["input_ids"], torch.tensor([[41, 42, 43]])], dim=1)

    rnd = random.Random(8)
    for trial in range(60):
        name = rnd.choice(["meta-llama/Meta-Llama-3-8B-Instruct", "lmsys/vicuna-7b", ""])
        roles = ("user", "assistant") if "llama-3" in name.lower() else ("USER", "ASSISTANT")
        history = None
        if rnd.random() < 0.6:
            history = []
            for t in range(rnd.randint(1, 3)):
                history += [{"role": roles[0], "text": f"q{t} <image>"}, {"role": roles[1], "text": f"a{t}"}]
        text = rnd.choice(["what about <image> and <image>?", "again", "compare them"])
        images = [object()] * rnd.randint(0, 3)
        outs = []
        for fn in (ref_chat, chat_mllava):
            m, p = Model(name), Proc()
            h = copy.deepcopy(history)
            reply, h2 = fn(text, list(images), m, p, max_input_length=77, history=h, max_new_tokens=5, do_sample=False)
            gen = m.calls[0]
            outs.append((reply, h2, p.calls, gen["eos_token_id"], gen["max_new_tokens"], gen["do_sample"], gen["input_ids"].tolist()))
        assert outs[0] == outs[1], (name, history, text)


def test_chat_mllava_stream_matches_the_live_reference():
    """chat_mllava_stream (generate() on a worker thread feeding a TextIteratorStreamer, the reply growing in history[-1])
    vs mantis/models/mllava/utils.py:100-186 with recording stubs: same yielded (reply, history) sequence, same generate kwargs"""
    import copy
    import threading
    import pytest
    from oracle.ref_shim import find_ref_root
    if find_ref_root() is None:
        pytest.skip("reference tree not available here")
    from oracle.ref_shim import load_reference_chat_utils
    from mantis_b200.models.mllava import chat_mllava_stream
    ref_stream = load_reference_chat_utils().chat_mllava_stream

    class Tok:
        eos_token_id = 2

        def convert_tokens_to_ids(self, t):
            return 128009

    class Proc:                                         # doubles as the streamer's tokenizer, like MLlavaProcessor does
        def __init__(self):
            self.tokenizer = Tok()

        def __call__(self, images=None, text=None, **kw):
            return {"input_ids": torch.arange(len(text.split()))[None], "pixel_values": None}

        def decode(self, ids, skip_special_tokens=True, **kw):
            return "".join(f"w{int(i)} " for i in ids)

    class LM:
        name_or_path = "meta-llama/Meta-Llama-3-8B-Instruct"

    class Model:
        device = torch.device("cpu")
        language_model = LM()

        def __init__(self):
            self.calls, self.threads = [], []

        def generate(self, streamer=None, **kw):
            self.calls.append(kw)
            self.threads.append(threading.current_thread())
            streamer.put(kw["input_ids"])               # the prompt (skipped by the streamer)
            for t in (41, 42, 43, 44):
                streamer.put(torch.tensor([t]))
            streamer.end()

    outs = []
    for fn in (ref_stream, chat_mllava_stream):
        m = Model()
        hist = [{"role": "user", "text": "look <image>"}, {"role": "assistant", "text": "ok"}]
        seq = [(r, copy.deepcopy(h)) for r, h in fn("and now?", [], m, Proc(), max_input_length=50, history=hist,
                                                    max_new_tokens=4, do_sample=False)]
        assert m.threads[0] is not threading.main_thread()            # generate() really ran on a worker thread
        gen = m.calls[0]
        outs.append((seq, gen["eos_token_id"], gen["max_new_tokens"], gen["do_sample"], gen["input_ids"].tolist()))
    assert outs[0] == outs[1]
    assert outs[1][0][-1][0].split() == ["w41", "w42", "w43", "w44"]


def test_mantis_alias_package_exports_fail_loudly():
    """the `mantis.models.mllava` alias shim re-exports the processor and both chat helpers without swallowing import errors"""
    import inspect
    import mantis.models.mllava as alias
    for name in ("LlavaForConditionalGeneration", "MLlavaForConditionalGeneration", "LlavaConfig", "MLlavaProcessor",
                 "chat_mllava", "chat_mllava_stream"):
        assert hasattr(alias, name), name
    assert "except" not in inspect.getsource(alias)

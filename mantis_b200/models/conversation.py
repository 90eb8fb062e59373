"""Conversation templates of the hot-path families (prompt layout the models see).

Mirrors the pieces of mantis/models/conversation.py the mllava / idefics2 path uses: the `Conversation` container,
`SeparatorStyle.{SINGLE,TWO,LLAMA_3,IDEFICS_2,IDEFICS_3}` rendering (ref :43-158) and the templates `mllava_v1`,
`mllava_v1_mmtag`, `llama_3`, `idefics_2`, `idefics_3` (ref :452-500).  Other families' templates are out of scope.
"""
import dataclasses
from enum import Enum, auto
from typing import List, Optional


class SeparatorStyle(Enum):
    SINGLE = auto()
    TWO = auto()
    LLAMA_3 = auto()
    IDEFICS_2 = auto()
    IDEFICS_3 = auto()


@dataclasses.dataclass
class Conversation:
    system: str
    roles: List[str]
    messages: List[List[str]]
    offset: int
    sep_style: SeparatorStyle = SeparatorStyle.SINGLE
    sep: str = "###"
    sep2: Optional[str] = None
    version: str = "Unknown"
    name: Optional[str] = None

    def get_prompt(self) -> str:
        msgs = [(r, m[0] if isinstance(m, tuple) else m) for r, m in self.messages]
        if self.sep_style == SeparatorStyle.SINGLE:
            ret = self.system + self.sep
            for role, message in msgs:
                ret += (role + ": " + message + self.sep) if message else (role + ":")
        elif self.sep_style == SeparatorStyle.TWO:
            seps = [self.sep, self.sep2]
            ret = self.system + seps[0]
            for i, (role, message) in enumerate(msgs):
                ret += (role + ": " + message + seps[i % 2]) if message else (role + ":")
        elif self.sep_style == SeparatorStyle.LLAMA_3:
            ret = self.system + self.sep
            for role, message in msgs:
                head = f"<|start_header_id|>{role}<|end_header_id|>\n\n"
                ret += (head + message + self.sep) if message else head
        elif self.sep_style == SeparatorStyle.IDEFICS_2:
            ret = (self.system + self.sep) if self.system else ""
            for role, message in msgs:
                ret += (role + ":" + message + self.sep + "\n") if message else (role + ":")
        elif self.sep_style == SeparatorStyle.IDEFICS_3:          # IDEFICS_2 layout behind the LLaMA-3 BOS token (ref :146-158)
            ret = "<|begin_of_text|>" + ((self.system + self.sep) if self.system else "")
            for role, message in msgs:
                ret += (role + ":" + message + self.sep + "\n") if message else (role + ":")
        else:
            raise ValueError(f"Invalid style: {self.sep_style}")
        return ret

    def append_message(self, role, message):
        self.messages.append([role, message])

    def copy(self):
        return Conversation(system=self.system, roles=self.roles, messages=[[x, y] for x, y in self.messages],
                            offset=self.offset, sep_style=self.sep_style, sep=self.sep, sep2=self.sep2,
                            version=self.version, name=self.name)


conv_mllava_v1_mmtag = Conversation(
    system="A chat between a curious user and an artificial intelligence assistant. "
           "The assistant is able to understand the multiple visual contents that the user provides, and assist the user with a variety of tasks using natural language."
           "Each visual content will be provided with the following format: <Image>visual content</Image>.",
    roles=("USER", "ASSISTANT"), messages=(), offset=0, sep_style=SeparatorStyle.SINGLE, sep="</s>", version="v1_mmtag")

conv_mllava_v1 = Conversation(
    system="A chat between a curious human and an artificial intelligence assistant. "
           "The assistant gives helpful, detailed, and polite answers to the human's questions.",
    roles=("USER", "ASSISTANT"), version="v1", messages=(), offset=0, sep_style=SeparatorStyle.SINGLE, sep="</s>")

conv_llama_3 = Conversation(
    system="<|start_header_id|>system<|end_header_id|>\n\nYou are a pirate chatbot who always responds in pirate speak!",
    roles=("user", "assistant"), messages=(), offset=0, sep_style=SeparatorStyle.LLAMA_3, sep="<|eot_id|>")

conv_idefics_2 = Conversation(system="", roles=("User", "Assistant"), messages=(), offset=0,
                              sep_style=SeparatorStyle.IDEFICS_2, sep="<end_of_utterance>")

conv_idefics_3 = Conversation(system="", roles=("User", "Assistant"), messages=(), offset=0,
                              sep_style=SeparatorStyle.IDEFICS_3, sep="<end_of_utterance>")

default_conversation = conv_mllava_v1
conv_templates = {"mllava_v1": conv_mllava_v1, "mllava_v1_mmtag": conv_mllava_v1_mmtag, "llama_3": conv_llama_3,
                  "idefics_2": conv_idefics_2, "idefics_3": conv_idefics_3}

"""Collation for the hot path (mirror of mantis/train/data.py:1375-1544 `Collator`, caller-side).

The reference delegates to `MLlavaProcessor._right_pad_inputs_with_attention_mask`, which asserts batch == 1.  This
collator keeps that contract when given one sample and additionally right-pads real batches (input_ids with the pad id,
attention_mask with 0, labels with -100, position_ids with 0) so that the B > 1 merge path of the model can be used;
with an MLlavaProcessor `pixel_values` stays a python list of per-sample tensors (modeling_llava.py:431-432 cats it); with
any other processor (Idefics2/3) the reference's own `_right_pad_inputs_with_attention_mask` (data.py:1392-1532) is followed:
`pixel_values` [1, N, C, H, W] are concatenated on dim 0, 4-D masks (packed [1,1,q,kv] and `pixel_attention_mask`
[1,N,H,W]) are zero-padded on their last two dims.
"""
from typing import Dict, List

import torch


def llava_valid_rows(input_ids, labels, attention_mask, image_token_index, left_padding, ignore_index=-100):
    """Number of supervised rows of the SHIFTED loss after the image-token merge, from the host copies of the batch (no device
    work): the count the model would otherwise read back to size the compacted LM-head GEMMs.
    ref modeling_llava.py:293-360 + 523-531: a text token keeps its label and mask, image slots and alignment padding get
    ignore_index / mask 0, and the loss pairs logits[s] with labels[s+1] where mask[s+1] != 0 -- so every valid text token
    counts except one sitting on merged position 0, which is original token 0 of a row that is not shifted right (right padding,
    or a left-padded row carrying the batch's maximum number of images)."""
    is_img = input_ids == image_token_index
    valid = (labels != ignore_index) & ~is_img
    if attention_mask is not None and attention_mask.dim() == 2:
        valid &= attention_mask != 0
    n_img = is_img.sum(dim=-1)
    at_zero = (n_img == n_img.max()) if left_padding else torch.ones_like(n_img, dtype=torch.bool)
    return int(valid.sum()) - int((valid[:, 0] & at_zero).sum())


def plain_valid_rows(labels, attention_mask, ignore_index):
    """same for models whose sequence is not expanded by a merge (Idefics2/3: ref modeling_idefics2.py:1883-1899): positions
    1.. whose label is not `ignore_index` and whose mask is set"""
    valid = labels[:, 1:] != ignore_index
    if attention_mask is not None and attention_mask.dim() == 2:
        valid = valid & (attention_mask[:, 1:] != 0)
    return int(valid.sum())


class Collator:
    def __init__(self, processor=None, max_length=None, pad_token_id=None, label_pad=-100, image_token_index=None):
        self.processor = processor
        self.max_length = max_length
        self.pad_token_id = pad_token_id
        self.label_pad = label_pad
        # when set, every batch carries `merge_hint` = what the model's image-token merge would otherwise read back from
        # the device ("emitting merged lengths so the scatter needs no host sync"): the largest number of
        # <image> placeholders in a row and the padding side, both from the host copy of input_ids
        self.image_token_index = image_token_index

    def merge_hint(self, input_ids, pad_token_id, labels=None, attention_mask=None):
        n = int((input_ids == self.image_token_index).sum(dim=-1).max())
        left = not bool((input_ids[:, -1] == pad_token_id).any())          # ref: modeling_llava.py:296
        hint = {"max_image_tokens": n, "left_padding": left}
        if labels is not None:
            hint["valid_rows"] = llava_valid_rows(input_ids, labels, attention_mask, self.image_token_index, left,
                                                  self.label_pad)
        return hint

    def _pad_id(self):
        if self.pad_token_id is not None:
            return self.pad_token_id
        tok = getattr(self.processor, "tokenizer", None)
        pid = getattr(tok, "pad_token_id", None)
        if pid is None:
            raise ValueError("Collator needs a pad_token_id to pad batches of more than one sample")
        return pid

    def _pad_to(self, t: torch.Tensor, length: int, value) -> torch.Tensor:
        if t.shape[1] >= length:
            return t
        pad = torch.full((t.shape[0], length - t.shape[1]), value, dtype=t.dtype, device=t.device)
        return torch.cat([t, pad], dim=1)

    def __call__(self, batch: List[Dict]):
        # no processor, or an MLlavaProcessor: the LLaVA data contract (pixel_values = list of per-sample [n_img, C, H, W]);
        # any other processor (Idefics2/3): the reference Collator's own padding rules
        mllava = self.processor is None or hasattr(self.processor, "_right_pad_inputs_with_attention_mask")
        if len(batch) == 1 and "cu_segments" not in batch[0] and self.processor is not None and mllava:
            return self.processor._right_pad_inputs_with_attention_mask(model_inputs=batch)      # the reference's path
        packed = any("cu_segments" in b or (isinstance(b.get("attention_mask"), torch.Tensor) and b["attention_mask"].dim() == 4)
                     for b in batch)
        out = {}
        for k in batch[0].keys():
            vals = [b[k] for b in batch]
            if vals[0] is None:
                out[k] = None
            elif "pixel_values" in k and (mllava or isinstance(vals[0], list)):
                # MLlavaProcessor contract (processing_llava.py / modeling_llava.py:431-432): a python list of per-sample
                # tensors that the model concatenates itself (None kept, the model skips them)
                out[k] = [v for v in vals]
            elif k == "cu_segments":                             # packed rows (PackingDataset): renumber the batch row
                out[k] = [(i, s0, s1) for i, segs in enumerate(vals) for (_, s0, s1) in segs]
            elif "attention_mask" in k and isinstance(vals[0], torch.Tensor) and vals[0].dim() == 4:
                # ref data.py:1444-1479: zero-pad dim 2 to the batch max and dim 3 to the batch max, independently -- packed
                # block-diagonal masks [1, 1, q, kv] and Idefics2's pixel_attention_mask [1, N, H, W] both come through here
                d2 = max(v.shape[2] for v in vals); d3 = max(v.shape[3] for v in vals)
                padded = []
                for v in vals:
                    m = torch.zeros((v.shape[0], v.shape[1], d2, d3), dtype=v.dtype, device=v.device)
                    m[:, :, : v.shape[2], : v.shape[3]] = v
                    padded.append(m)
                out[k] = torch.cat(padded, dim=0)
            elif k == "input_ids" or k == "labels" or "attention_mask" in k or "position_ids" in k:
                L = max(v.shape[1] for v in vals)
                if self.max_length is not None and not packed:   # truncating a packed row would cut through its segment
                    L = min(L, self.max_length)                  # table and 4-D mask: packed rows are sized by PackingDataset
                fill = self._pad_id() if k == "input_ids" else (self.label_pad if k == "labels" else 0)
                out[k] = torch.cat([self._pad_to(v[:, :L], L, fill) for v in vals], dim=0)
            else:
                # everything else (Idefics2/3 pixel_values [1, N, C, H, W], ...): plain concatenation like the reference
                # (data.py:1529)
                out[k] = torch.cat(vals, dim=0)
        if self.image_token_index is not None and out.get("input_ids") is not None and "cu_segments" not in out:
            out["merge_hint"] = self.merge_hint(out["input_ids"], self._pad_id(), out.get("labels"), out.get("attention_mask"))
        return out


class PackingDataset(torch.utils.data.Dataset):
    """Sequence packing for the self-attention models (ref: mantis/train/data.py:1546-1671).

    Consecutive items of `dataset` (dicts with `input_ids [1, T]`, `attention_mask [1, T]`, `labels [1, T]` or `[T]`,
    `pixel_values`) are concatenated until the running length exceeds `max_self_attn_len` -- the item that crosses the
    limit is included, like the reference -- and returned as ONE row with
      * `attention_mask` [1, 1, S, S] int32 block-diagonal (block i = item i's key mask broadcast over its queries),
      * `position_ids` restarting at 0 for every item,
      * `cu_segments`: the (row, start, end) table the varlen attention launches from (no device sync needed).
    Differences from the reference, which cannot be instantiated as written (it reads `self.packing_same_mm_media` before
    assigning it, data.py:1555-1558): `position_ids` / `labels` keep a leading batch dim of 1 so the row can go straight into
    `model(**batch)`, and the packing interval is estimated lazily from the first `probe` packs.
    The block-diagonal mask is only materialised for callers that want the reference's tensors (`dense_mask=True`):
    it is S^2 int32 (268 MB at S = 8192) and the model needs just the segment table + a [1, S] key mask."""

    def __init__(self, dataset, max_self_attn_len, dense_mask=True, probe=20):
        super().__init__()
        self.dataset = dataset
        self.max_self_attn_len = max_self_attn_len
        self.dense_mask = dense_mask
        self.packing_same_mm_media = getattr(dataset, "packing_same_mm_media", False)
        assert not self.packing_same_mm_media, "Packing same mm media is not supported for self-attention based models"
        self.average_packing_interval = self._infer_interval(probe)
        self.num_last_packed_items = self.average_packing_interval

    def _take(self, start):
        items, total, i = [], 0, start
        while True:
            item = self.dataset[i % len(self.dataset)]
            total += item["input_ids"].shape[1]
            items.append(item)
            if (self.max_self_attn_len and total > self.max_self_attn_len) or len(items) >= len(self.dataset):
                return items
            i += 1

    def _infer_interval(self, probe):
        counts, i = [], 0
        for _ in range(min(probe, max(1, len(self.dataset)))):
            n = len(self._take(i))
            counts.append(n)
            i += n
        return max(1, -(-sum(counts) // len(counts)))

    def __len__(self):
        return max(1, len(self.dataset) // self.average_packing_interval)

    def __getitem__(self, idx):
        offset = self.num_last_packed_items - self.average_packing_interval
        items = self._take(idx * self.average_packing_interval + offset)
        self.num_last_packed_items = len(items)
        return self.pack_batch(items)

    def pack_batch(self, items):
        ids = torch.cat([x["input_ids"] for x in items], dim=1)
        S = ids.shape[1]
        pv = [x.get("pixel_values") for x in items]
        if any(isinstance(p, list) for p in pv):
            pixel_values = sum([p or [] for p in pv], []) or None
        elif all(p is None for p in pv):
            pixel_values = None
        else:
            pixel_values = torch.cat([torch.as_tensor(p) for p in pv if p is not None], dim=0)
        key_mask = torch.cat([x["attention_mask"].reshape(1, -1) for x in items], dim=1).to(torch.int32)
        segs, pos, acc = [], [], 0
        for x in items:
            n = x["input_ids"].shape[1]
            segs.append((0, acc, acc + n))
            pos.append(torch.arange(n, dtype=torch.long))
            acc += n
        out = {
            "input_ids": ids,
            "pixel_values": pixel_values,
            "position_ids": torch.cat(pos).unsqueeze(0),
            "labels": torch.cat([x["labels"].reshape(1, -1) for x in items], dim=1),
            "cu_segments": segs,
        }
        if self.dense_mask:
            mask = torch.zeros((1, 1, S, S), dtype=torch.int32)
            for (_, s0, s1) in segs:
                mask[0, 0, s0:s1, s0:s1] = key_mask[:, s0:s1].expand(s1 - s0, s1 - s0)
            out["attention_mask"] = mask
        else:
            out["attention_mask"] = key_mask.to(torch.long)
        for k in items[0].keys():
            if k in out or k in ("encoder_attention_mask", "encoder_position_ids"):
                continue
            v0 = items[0][k]
            if isinstance(v0, torch.Tensor):
                out[k] = torch.cat([x[k] for x in items], dim=0)
            elif isinstance(v0, list):
                out[k] = sum([x[k] for x in items], [])
            else:
                out[k] = [x[k] for x in items]
        return out


def merged_length(num_text_tokens: int, num_images: int, patches_per_image: int) -> int:
    """length of a sample after the image-token merge: every <image> placeholder grows to `patches_per_image` rows
    (ref: modeling_llava.py:305, max_embed_dim = num_special_image_tokens * (num_image_patches - 1) + sequence_length)"""
    return num_text_tokens + num_images * (patches_per_image - 1)


def partition_balanced(lengths, world_size: int, samples_per_rank=None):
    """Data-parallel partition of one global step ("ragged sequence lengths in real data => sort/bucket by
    merged length to balance ranks").  Every rank runs its samples one by one and the step ends with the gradient
    all-reduce, so the step time is the time of the rank with the most tokens (attention makes it superlinear, which the
    quadratic term of the cost accounts for).  Longest-processing-time greedy: samples sorted by cost, each handed to the
    currently lightest rank that still has a free slot.  Returns `world_size` lists of indices into `lengths`.
    Deterministic (ties by index), every rank gets the same number of samples when `samples_per_rank` is given."""
    n = len(lengths)
    if samples_per_rank is None:
        samples_per_rank = -(-n // world_size)
    if n > samples_per_rank * world_size:
        raise ValueError(f"{n} samples do not fit {world_size} ranks x {samples_per_rank}")
    cost = [float(L) + float(L) * float(L) / 65536.0 for L in lengths]     # linear layers + causal attention (S^2 / 2 x 32 heads ...)
    order = sorted(range(n), key=lambda i: (-cost[i], i))
    load = [0.0] * world_size
    parts = [[] for _ in range(world_size)]
    for i in order:
        r = min((r for r in range(world_size) if len(parts[r]) < samples_per_rank), key=lambda r: (load[r], r))
        parts[r].append(i)
        load[r] += cost[i]
    return parts


def assistant_labels(input_ids: torch.Tensor, sep_id: int, style: str, has_system: bool = True, sep_offset: int = 0,
                     ignore_index: int = -100) -> torch.Tensor:
    """Labels of one tokenised conversation: the assistant turns keep their token ids, everything else (system prompt, user
    turns, image placeholders' surroundings) is `ignore_index` -- the supervision mask of instruction tuning, derived from
    the positions of the turn separator exactly like the reference's ChatDataset (mantis/train/data.py:421-452).

    style "single" / "llama_3" (mllava_v1, llama_3 templates): the text after every ODD separator (0-based: 0 closes the
        system prompt, 1 the first user turn, ...) up to and including the next separator is supervised; after the last
        separator, if it is odd, everything to the end.
    style "idefics" (idefics_2 / idefics_3 templates, separator = the stripped `conv.sep`): same walk, but the parity flips
        when there is no system prompt (then separator 0 already closes a user turn) and the supervised span starts
        `1 + sep_offset` tokens after the separator (skipping the "\nAssistant:" role tokens is left to sep_offset)."""
    ids = input_ids.reshape(-1)
    labels = torch.full_like(ids, ignore_index)
    seps = torch.nonzero(ids == sep_id, as_tuple=True)[0].tolist()
    if style in ("single", "llama_3"):
        skip_parity, start = 0, 1
    elif style == "idefics":
        skip_parity, start = (0 if has_system else 1), 1 + sep_offset
    else:
        raise ValueError(f"unknown separator style {style!r}")
    for i, pos in enumerate(seps):
        if i % 2 == skip_parity:
            continue
        end = ids.numel() if i == len(seps) - 1 else seps[i + 1] + 1
        labels[pos + start:end] = ids[pos + start:end]
    return labels.view_as(input_ids)


class ChatDataset(torch.utils.data.Dataset):
    """Conversation-style fine-tuning items for the hot path (the part of mantis/train/data.py:94-505 `ChatDataset` that
    produces model inputs; dataset download / caching / video are out of scope): every record is
    {"conversation": [{"role"|"from": "human"/"user"/"gpt"/"assistant", "content"|"text"|"value": str}, ...], "images": [PIL...]}.
    Missing `<image>` placeholders are prepended to the first turn (ref:403-406), the prompt is rendered with a
    `mantis_b200.models.conversation` template, tokenised + image-processed by the processor (truncation to `max_seq_len`),
    and `labels` supervise the assistant turns only (`assistant_labels`)."""

    _STYLE = {"SINGLE": "single", "LLAMA_3": "llama_3", "IDEFICS_2": "idefics", "IDEFICS_3": "idefics"}

    def __init__(self, records, processor, conv, max_seq_len=None, image_key="images", ignore_index=-100):
        self.records, self.processor, self.conv = records, processor, conv
        self.max_seq_len, self.image_key, self.ignore_index = max_seq_len, image_key, ignore_index

    def __len__(self):
        return len(self.records)

    def messages(self, record):
        conv = self.conv
        roles = {"human": conv.roles[0], "user": conv.roles[0], "gpt": conv.roles[1], "assistant": conv.roles[1]}
        turns = record["conversation" if "conversation" in record else "conversations"]
        if roles[turns[0].get("from", turns[0].get("role"))] != conv.roles[0]:
            turns = turns[1:]                                                     # a leading non-user turn is dropped (ref:355-357)
        out = []
        for j, t in enumerate(turns):
            role = roles[t.get("from", t.get("role"))]
            if role != conv.roles[j % 2]:
                raise ValueError("conversation turns must alternate user / assistant")
            out.append([role, t.get("content", t.get("text", t.get("value", "")))])
        return out

    def __getitem__(self, idx):
        record = self.records[idx]
        msgs = self.messages(record)
        images = record.get(self.image_key)
        if images is not None and not isinstance(images, list):
            images = [images]
        n_tok = sum(m[1].count("<image>") for m in msgs)
        if isinstance(images, list) and n_tok < len(images):
            msgs[0][1] = "<image>" * (len(images) - n_tok) + msgs[0][1]
        conv = self.conv.copy()
        conv.messages = msgs
        enc = self.processor(conv.get_prompt(), images, return_tensors="pt", truncation=True, max_length=self.max_seq_len)
        style = self._STYLE[conv.sep_style.name]
        sep_tok = conv.sep.strip(" \n") if style == "idefics" else conv.sep
        sep_id = self.processor.tokenizer.convert_tokens_to_ids(sep_tok)
        enc["labels"] = assistant_labels(enc["input_ids"], sep_id, style, has_system=bool(conv.system),
                                         sep_offset=getattr(conv, "sep_offset", 0), ignore_index=self.ignore_index)
        return enc

"""PackingDataset / Collator host logic (ref: mantis/train/data.py:1546-1671 `PackingDataset.pack_batch`)."""
import torch

from mantis_b200.train import Collator, PackingDataset


class _DS(torch.utils.data.Dataset):
    lens = [5, 3, 7, 4, 6, 2, 9]

    def __len__(self):
        return len(self.lens)

    def __getitem__(self, i):
        n = self.lens[i]
        am = torch.ones(1, n, dtype=torch.long)
        if i == 2:
            am[0, -2:] = 0                                     # a sample that carries its own padding
        return {"input_ids": torch.arange(n).unsqueeze(0) + 100 * i, "attention_mask": am,
                "labels": torch.arange(n).unsqueeze(0) + 100 * i, "pixel_values": torch.full((1, 3, 2, 2), float(i))}


def test_pack_batch_layout_follows_the_reference():
    pd = PackingDataset(_DS(), max_self_attn_len=10)
    assert pd.average_packing_interval == 3 and len(pd) == 2
    row = pd[0]                                               # items 0, 1, 2: 5 + 3 = 8 <= 10, the 7-token item crosses it
    S = 15
    assert row["input_ids"].tolist() == [[0, 1, 2, 3, 4, 100, 101, 102, 200, 201, 202, 203, 204, 205, 206]]
    assert row["labels"].tolist() == row["input_ids"].tolist()
    assert row["position_ids"].tolist() == [[0, 1, 2, 3, 4, 0, 1, 2, 0, 1, 2, 3, 4, 5, 6]]
    assert row["cu_segments"] == [(0, 0, 5), (0, 5, 8), (0, 8, 15)]
    assert row["pixel_values"].shape == (3, 3, 2, 2) and row["pixel_values"][:, 0, 0, 0].tolist() == [0.0, 1.0, 2.0]
    m = row["attention_mask"]
    assert m.shape == (1, 1, S, S) and m.dtype == torch.int32
    expect = torch.zeros(S, S, dtype=torch.int32)
    expect[0:5, 0:5] = 1; expect[5:8, 5:8] = 1; expect[8:15, 8:13] = 1      # block i = item i's key mask, every query row
    assert torch.equal(m[0, 0], expect)
    # the sync-free variant carries a 2-D key mask instead of the S^2 tensor
    row2 = PackingDataset(_DS(), max_self_attn_len=10, dense_mask=False)[0]
    assert row2["attention_mask"].tolist() == [[1] * 13 + [0, 0]]
    assert torch.equal((m != 0).any(dim=1).any(dim=1).long(), row2["attention_mask"])


def test_collator_batches_packed_rows():
    pd = PackingDataset(_DS(), max_self_attn_len=10)
    a, b = pd[0], pd[1]
    Sa, Sb = a["input_ids"].shape[1], b["input_ids"].shape[1]
    batch = Collator(pad_token_id=7)([a, b])
    L = max(Sa, Sb)
    assert batch["input_ids"].shape == (2, L) and batch["attention_mask"].shape == (2, 1, L, L)
    assert batch["attention_mask"][1, 0, Sb:, :].sum() == 0 and batch["attention_mask"][1, 0, :, Sb:].sum() == 0
    assert batch["labels"][1, Sb:].tolist() == [-100] * (L - Sb) and batch["input_ids"][1, Sb:].tolist() == [7] * (L - Sb)
    assert batch["cu_segments"][: len(a["cu_segments"])] == a["cu_segments"]
    assert [s[0] for s in batch["cu_segments"][len(a["cu_segments"]):]] == [1] * len(b["cu_segments"])
    assert isinstance(batch["pixel_values"], list) and len(batch["pixel_values"]) == 2


def test_packed_segments_from_position_ids():
    from mantis_b200 import ops
    pos = torch.tensor([[0, 1, 2, 0, 1, 0, 1, 2, 3], [0, 1, 2, 3, 4, 5, 0, 1, 2]])
    assert ops.packed_segments(pos) == [(0, 0, 3), (0, 3, 5), (0, 5, 9), (1, 0, 6), (1, 6, 9)]
    assert ops.packed_segments(torch.tensor([5, 6, 0, 1])) == [(0, 0, 2), (0, 2, 4)]    # 1-D ids (reference layout)


def test_partition_balanced_by_merged_length():
    from mantis_b200.train import merged_length, partition_balanced
    assert merged_length(2048, 8, 728) == 7864                          # the bench sample
    import random
    rnd = random.Random(0)
    lens = [merged_length(rnd.randint(64, 2048), rnd.randint(1, 20), 728) for _ in range(32)]
    parts = partition_balanced(lens, world_size=8, samples_per_rank=4)
    assert sorted(i for p in parts for i in p) == list(range(32)) and all(len(p) == 4 for p in parts)
    tok = [sum(lens[i] for i in p) for p in parts]
    naive = [sum(lens[r * 4:(r + 1) * 4]) for r in range(8)]              # the order the sampler happened to deliver
    assert max(tok) <= max(naive) and max(tok) - min(tok) < max(naive) - min(naive)
    assert max(tok) <= 1.15 * (sum(tok) / 8)                             # within 15 % of the perfect split
    assert partition_balanced(lens, 8, 4) == parts                       # deterministic
    assert partition_balanced([5, 5, 5], 2) == [[0, 2], [1]]
    import pytest
    with pytest.raises(ValueError):
        partition_balanced(lens, 4, 4)


def _ref_data():
    import pytest
    from oracle.ref_shim import find_ref_root
    if find_ref_root() is None:
        pytest.skip("reference tree not available here")
    from oracle.ref_shim import load_reference_train_data
    return load_reference_train_data()


def test_pack_batch_matches_the_live_reference():
    """PackingDataset.pack_batch vs the unmodified mantis/train/data.py:1607-1671 (its __init__ cannot run -- it reads
    `self.packing_same_mm_media` before assigning it -- so the method is called on a bare instance).  The reference emits 1-D
    position_ids / labels (it concatenates per-item 1-D tensors); ours keeps a leading batch dim of 1, compared flattened."""
    ref = _ref_data()
    ds = _DS()
    items = [ds[i] for i in (0, 2, 3, 5)]
    r = object.__new__(ref.PackingDataset).pack_batch([dict(it, labels=it["labels"][0]) for it in items])
    o = PackingDataset(ds, max_self_attn_len=10).pack_batch(items)
    assert torch.equal(o["input_ids"], r["input_ids"])
    assert torch.equal(o["attention_mask"], r["attention_mask"]) and o["attention_mask"].dtype == r["attention_mask"].dtype
    assert torch.equal(o["position_ids"].reshape(-1), r["position_ids"]) and torch.equal(o["labels"].reshape(-1), r["labels"])
    assert torch.equal(o["pixel_values"], r["pixel_values"])


def test_collator_matches_the_live_reference():
    """right padding of real batches (ids with the pad id, masks with 0 -- 2-D and packed 4-D --, labels with -100, position ids
    with 0) vs the reference Collator's own `_right_pad_inputs_with_attention_mask` (data.py:1392-1527)"""
    ref = _ref_data()

    class Tok:
        pad_token_id = 7

    class Proc:                                    # no `_right_pad_inputs_with_attention_mask` -> the Collator's own is used
        tokenizer = Tok()
    rc = ref.Collator(Proc())
    rc.tokenizer = Tok()                           # the reference's method reads self.tokenizer, which its __init__ never sets
    oc = Collator(pad_token_id=7)
    ds = _DS()
    plain = [{k: v for k, v in ds[i].items() if k != "pixel_values"} for i in (0, 1, 6)]
    for it in plain:
        it["position_ids"] = torch.arange(it["input_ids"].shape[1])[None]
    r, o = rc([dict(x) for x in plain]), oc([dict(x) for x in plain])
    for k in ("input_ids", "attention_mask", "labels", "position_ids"):
        assert torch.equal(o[k], r[k]), k
    pd = PackingDataset(ds, max_self_attn_len=10)
    packed = [{k: v for k, v in pd[i].items() if k not in ("pixel_values", "cu_segments")} for i in (0, 1)]
    r, o = rc([dict(x) for x in packed]), oc([dict(x) for x in packed])
    for k in ("input_ids", "attention_mask", "labels", "position_ids"):
        assert torch.equal(o[k], r[k]), k


def test_collator_idefics2_pixel_tensors_match_the_live_reference():
    """Idefics2/3 processors have no `_right_pad_inputs_with_attention_mask`: the reference Collator then concatenates
    `pixel_values` [1, N, C, H, W] on dim 0 and zero-pads the 4-D `pixel_attention_mask` [1, N, H, W] on its last two dims
    independently (H != W, different sizes per sample) -- data.py:1444-1479,1529."""
    ref = _ref_data()

    class Tok:
        pad_token_id = 7

    class Proc:
        tokenizer = Tok()
    rc = ref.Collator(Proc()); rc.tokenizer = Tok()
    oc = Collator(processor=Proc(), pad_token_id=7)
    g = torch.Generator().manual_seed(3)
    items = []
    for T, (H, W) in ((5, (28, 42)), (9, (28, 42)), (7, (28, 42))):
        items.append({"input_ids": torch.randint(8, 50, (1, T), generator=g), "attention_mask": torch.ones(1, T, dtype=torch.long),
                      "labels": torch.randint(8, 50, (1, T), generator=g),
                      "pixel_values": torch.randn(1, 2, 3, H, W, generator=g),
                      "pixel_attention_mask": torch.ones(1, 2, H, W, dtype=torch.bool)})
    # ragged masks (the reference pads them; pixel_values of equal size are concatenated)
    items[1]["pixel_attention_mask"] = torch.ones(1, 2, 14, 56, dtype=torch.bool)
    r, o = rc([dict(x) for x in items]), oc([dict(x) for x in items])
    assert set(r.keys()) <= set(o.keys())
    for k in r:
        assert isinstance(o[k], torch.Tensor) and o[k].shape == r[k].shape and torch.equal(o[k], r[k]), k
    assert o["pixel_values"].shape == (3, 2, 3, 28, 42) and o["pixel_attention_mask"].shape == (3, 2, 28, 56)


def test_collator_max_length_leaves_packed_rows_alone():
    ds = _DS()
    pd = PackingDataset(ds, max_self_attn_len=10)
    rows = [pd[0], pd[1]]
    out = Collator(pad_token_id=7, max_length=4)([dict(x) for x in rows])
    L = max(r["input_ids"].shape[1] for r in rows)
    assert out["input_ids"].shape[1] == L and out["attention_mask"].shape[-2:] == (L, L)
    assert all(s1 <= L for (_, _, s1) in out["cu_segments"])
    plain = [{k: v for k, v in ds[i].items() if k != "pixel_values"} for i in (0, 1)]
    out = Collator(pad_token_id=7, max_length=4)([dict(x) for x in plain])
    assert out["input_ids"].shape[1] == 4 and out["labels"].shape[1] == 4


def test_llava_valid_rows_hint_matches_the_merge_oracle():
    """the collator's host-side count of supervised rows (what lets the LM-head row compaction run without a device read-back)
    == the count obtained by actually merging (numpy restatement of ref modeling_llava.py:293-360, pinned to the reference in
    tests/test_oracle.py) and applying the shifted loss mask of :523-531 -- left / right padding, uneven image counts, masked
    text tokens, ignored labels, image placeholder at position 0"""
    import numpy as np
    from oracle.merge_oracle import merge_oracle
    from mantis_b200.train import llava_valid_rows
    rng = np.random.default_rng(0)
    IMG, PAD = 90, 91
    for trial in range(300):
        B, T, P, D = int(rng.integers(1, 5)), int(rng.integers(4, 24)), int(rng.integers(2, 6)), 2
        left = bool(rng.integers(0, 2))
        ids = rng.integers(1, 80, size=(B, T))
        att = np.ones((B, T), dtype=np.int64)
        n_img = []
        for b in range(B):
            npad = int(rng.integers(0, T // 2)) if B > 1 else 0
            if b == 0 and not left:
                npad = max(npad, 1) if B > 1 else 0           # right padding is recognised by a pad token in the last column
            if left and b == 0:
                npad = 0                                      # left padding: no pad in the last column of any row
            body = T - npad
            k = int(rng.integers(0, min(3, body) + 1))
            pos = rng.choice(body, size=k, replace=False)
            row = ids[b, :body].copy(); row[pos] = IMG
            if left:
                ids[b, npad:] = row; ids[b, :npad] = PAD; att[b, :npad] = 0
            else:
                ids[b, :body] = row; ids[b, body:] = PAD; att[b, body:] = 0
            n_img.append(k)
        if not left and not (ids[:, -1] == PAD).any():
            continue                                          # would be classified as left padding by the reference
        labels = ids.copy()
        labels[rng.random((B, T)) < 0.3] = -100
        labels[ids == IMG] = -100
        att[(rng.random((B, T)) < 0.1) & (ids != IMG)] = 0     # a few masked text tokens
        feats = np.ones((sum(n_img), P, D), dtype=np.float32)
        emb = np.ones((B, T, D), dtype=np.float32) * 0.5
        emb[ids == PAD] = 0.5                                  # pad rows are ordinary (non-zero) text rows in this check
        try:
            _, fmask, flabels, _, _ = merge_oracle(feats, emb, ids, att, labels, IMG, PAD)
        except ValueError:
            continue
        expect = int(((flabels[:, 1:] != -100) & (fmask[:, 1:] != 0)).sum())
        is_left = not bool((ids[:, -1] == PAD).any())
        got = llava_valid_rows(torch.from_numpy(ids), torch.from_numpy(labels), torch.from_numpy(att), IMG, is_left)
        assert got == expect, (trial, left, ids.tolist(), labels.tolist(), att.tolist(), got, expect)

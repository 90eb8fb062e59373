"""Pre-allocated KV cache for generate(): replaces DynamicCache's torch.cat growth (K18 of SURVEY.md section 2.2).

Layout per layer: K and V as [B, capacity, Hkv, hd] (token-major, so appending a decode step is one contiguous
row write per sequence and the attention kernels read it with the same strides as activations).  Capacity grows
geometrically in pages of `page` tokens.  Implements the subset of the HF `Cache` protocol that
GenerationMixin and the reference's prepare_inputs_for_generation use (mantis/models/mllava/modeling_llava.py:
551-602): get_seq_length(), seen_tokens, legacy indexing cache[layer] -> (k, v) in [B, Hkv, S, hd] view.
"""
import torch


class B200KVCache:
    is_compileable = False

    def __init__(self, page: int = 256):
        self.page = page
        self.k = []
        self.v = []
        self.lengths = []

    # ---- HF Cache protocol (subset) ----
    def __len__(self):
        return len(self.k)

    def get_seq_length(self, layer_idx: int = 0) -> int:
        return self.lengths[layer_idx] if layer_idx < len(self.lengths) else 0

    @property
    def seen_tokens(self) -> int:
        return self.get_seq_length(0)

    def get_max_cache_shape(self):
        return None

    def get_mask_sizes(self, cache_position, layer_idx=0):
        return self.get_seq_length(layer_idx) + cache_position.shape[0], 0

    def __getitem__(self, layer_idx):
        n = self.lengths[layer_idx]
        return (self.k[layer_idx][:, :n].permute(0, 2, 1, 3), self.v[layer_idx][:, :n].permute(0, 2, 1, 3))

    def __iter__(self):
        for i in range(len(self.k)):
            yield self[i]

    def _ensure(self, layer_idx, B, need, Hkv, hd, dtype, device):
        while len(self.k) <= layer_idx:
            self.k.append(None); self.v.append(None); self.lengths.append(0)
        cur = self.k[layer_idx]
        if cur is None or cur.shape[1] < need or cur.shape[0] != B:
            cap = max(self.page, (need + self.page - 1) // self.page * self.page)
            if cur is not None and cur.shape[0] == B:
                cap = max(cap, 2 * cur.shape[1])
            nk = torch.empty((B, cap, Hkv, hd), dtype=dtype, device=device)
            nv = torch.empty((B, cap, Hkv, hd), dtype=dtype, device=device)
            n = self.lengths[layer_idx]
            if cur is not None and cur.shape[0] == B and n:
                nk[:, :n].copy_(cur[:, :n]); nv[:, :n].copy_(self.v[layer_idx][:, :n])
            self.k[layer_idx], self.v[layer_idx] = nk, nv

    def append(self, k_new, v_new, layer_idx):
        """k_new/v_new: [B, S_new, Hkv, hd]. Returns views [B, total, Hkv, hd] of the cache including the new tokens."""
        B, S_new, Hkv, hd = k_new.shape
        n = self.get_seq_length(layer_idx)
        self._ensure(layer_idx, B, n + S_new, Hkv, hd, k_new.dtype, k_new.device)
        self.k[layer_idx][:, n:n + S_new].copy_(k_new)
        self.v[layer_idx][:, n:n + S_new].copy_(v_new)
        self.lengths[layer_idx] = n + S_new
        return self.k[layer_idx][:, :n + S_new], self.v[layer_idx][:, :n + S_new]

    def reserve(self, capacity):
        """grow every layer to at least `capacity` tokens (one reallocation; afterwards pointers are stable, which the
        native decode engine relies on)"""
        for i in range(len(self.k)):
            cur = self.k[i]
            if cur is not None and cur.shape[1] < capacity:
                B, _, Hkv, hd = cur.shape
                self._ensure(i, B, capacity, Hkv, hd, cur.dtype, cur.device)

    def capacity(self):
        return min(k.shape[1] for k in self.k) if self.k else 0

    def advance(self, n=1):
        for i in range(len(self.lengths)):
            self.lengths[i] += n

    def reorder_cache(self, beam_idx):
        for i in range(len(self.k)):
            self.k[i] = self.k[i].index_select(0, beam_idx)
            self.v[i] = self.v[i].index_select(0, beam_idx)

    def crop(self, max_length):
        for i in range(len(self.lengths)):
            self.lengths[i] = min(self.lengths[i], max_length)

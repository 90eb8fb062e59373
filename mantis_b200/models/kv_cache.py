"""Paged KV cache for generate(): replaces DynamicCache's torch.cat growth (K18 of SURVEY.md section 2.2).

Memory layout (B200-first: sized for 180 GB of HBM, no re-layout on growth):

  page   = [L][2 (k, v)][128 tokens][Hkv][hd]      one page holds 128 tokens of ALL layers of one sequence
  slab   = [n_pages] pages                         separate device allocations, added when the free list runs dry
  table  = int64 [B, max_blocks]                   page base ADDRESSES, shared by every layer (layer l adds l * layer_stride)

The cache therefore grows by appending table entries -- old tokens are never copied -- and the decode kernels
(csrc/decode.cu: rope_append_kernel, decode_attn_kernel) resolve `token j -> table[b][j >> 7] + (j & 127) * row`
themselves.  128 tokens/page is the split-KV chunk of the decode attention kernel, so one split reads one page.

Implements the subset of the HF `Cache` protocol that GenerationMixin and the reference's
prepare_inputs_for_generation use (mantis/models/mllava/modeling_llava.py:551-602): get_seq_length(), seen_tokens,
legacy indexing cache[layer] -> (k, v) as [B, Hkv, S, hd] (gathered from the pages), crop, reorder_cache.
"""
import torch


def _ops():
    from .. import ops
    return ops


class B200KVCache:
    is_compileable = False
    PAGE = 128

    def __init__(self, n_layers=None, slab_tokens: int = 1024):
        self.n_layers = n_layers
        self.slab_tokens = slab_tokens           # per-sequence head-room of a new slab
        self.lengths = []
        self.batch = 0
        self.slabs = []                          # device tensors [n_pages, L, 2, PAGE, Hkv, hd]
        self.free = []                           # free page addresses
        self.blocks = []                         # per sequence: list of page addresses (host mirror of the table)
        self.table = None                        # int64 [B, max_blocks] on device
        self._table_dirty = False
        self.Hkv = self.hd = 0
        self.dtype = self.device = None

    # ---- HF Cache protocol (subset) ----
    def __len__(self):
        return len(self.lengths)

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
        k, v = self.gather(layer_idx)
        return k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3)

    def __iter__(self):
        for i in range(len(self.lengths)):
            yield self[i]

    # ---- geometry ----
    @property
    def row_elems(self):
        return self.Hkv * self.hd

    @property
    def v_off(self):                 # elements from a page's K rows to its V rows (same layer)
        return self.PAGE * self.row_elems

    @property
    def layer_stride(self):          # elements between consecutive layers inside a page
        return 2 * self.v_off

    @property
    def page_elems(self):
        return self.n_layers * self.layer_stride

    def capacity(self):
        return min((len(b) for b in self.blocks), default=0) * self.PAGE

    def table_stride(self):
        return self.table.shape[1] if self.table is not None else 0

    def pages_in_use(self):
        return sum(len(b) for b in self.blocks)

    # ---- page allocator ----
    def _configure(self, B, Hkv, hd, dtype, device):
        if self.n_layers is None:
            raise ValueError("B200KVCache needs n_layers (the decoder sets it before the first append)")
        if self.batch and (B != self.batch or Hkv != self.Hkv or hd != self.hd or dtype != self.dtype):
            self.reset()
        if not self.batch:
            if (Hkv * hd * torch.empty((), dtype=dtype).element_size()) % 16:
                raise ValueError("KV rows must be multiples of 16 bytes")
            self.batch, self.Hkv, self.hd, self.dtype, self.device = B, Hkv, hd, dtype, device
            self.blocks = [[] for _ in range(B)]
            self.lengths = [0] * self.n_layers

    def reset(self):
        self.slabs, self.free, self.blocks, self.table = [], [], [], None
        self.__dict__.pop("_engine", None)         # the native decode engine holds buffers sized for the old batch
        self.batch = 0
        self.lengths = []

    def _new_slab(self, min_pages):
        n = max(min_pages, self.batch * max(1, self.slab_tokens // self.PAGE))
        slab = torch.empty((n, self.n_layers, 2, self.PAGE, self.Hkv, self.hd), dtype=self.dtype, device=self.device)
        self.slabs.append(slab)
        step = self.page_elems * slab.element_size()
        base = slab.data_ptr()
        self.free.extend(base + i * step for i in range(n - 1, -1, -1))

    def ensure(self, tokens):
        """make sure every sequence owns pages for `tokens` tokens (no copies: new pages only extend the table)"""
        need = (tokens + self.PAGE - 1) // self.PAGE
        missing = sum(max(0, need - len(b)) for b in self.blocks)
        if not missing:
            return
        if missing > len(self.free):
            self._new_slab(missing - len(self.free))
        for b in self.blocks:
            while len(b) < need:
                b.append(self.free.pop())
        self._table_dirty = True

    def reserve(self, tokens):
        if self.batch:
            self.ensure(tokens)

    def device_table(self):
        """int64 [B, max_blocks] page addresses (uploaded lazily; the width grows in powers of two)"""
        width = max(len(b) for b in self.blocks)
        if self.table is None or self.table.shape[1] < width:
            w = 8
            while w < width:
                w *= 2
            self.table = torch.zeros((self.batch, w), dtype=torch.int64, device=self.device)
            self._table_dirty = True
        if self._table_dirty:
            host = torch.zeros((self.batch, self.table.shape[1]), dtype=torch.int64)
            for i, b in enumerate(self.blocks):
                if b:
                    host[i, : len(b)] = torch.tensor(b, dtype=torch.int64)
            self.table.copy_(host)
            self._table_dirty = False
        return self.table

    # ---- data movement (csrc/decode.cu: kv_page_copy_kernel) ----
    def _copy(self, k_lin, v_lin, layer_idx, start, to_pages):
        B, S = k_lin.shape[0], k_lin.shape[1]
        es = k_lin.element_size()
        assert k_lin.stride() == v_lin.stride() and k_lin.stride(3) == 1 and k_lin.stride(2) == self.hd
        tab = self.device_table()
        ops = _ops()
        ops._call("mb200_kv_page_copy", ops._p(k_lin), ops._p(v_lin), ops._p(tab), tab.shape[1],
                  layer_idx * self.layer_stride * es, self.v_off * es, B, S, start, self.row_elems * es,
                  k_lin.stride(0) * es, k_lin.stride(1) * es, int(to_pages), ops._st())

    def write(self, k_new, v_new, layer_idx):
        """store k_new/v_new [B, S_new, Hkv, hd] at the end of layer `layer_idx` (pages allocated on demand)"""
        B, S_new, Hkv, hd = k_new.shape
        self._configure(B, Hkv, hd, k_new.dtype, k_new.device)
        n = self.lengths[layer_idx]
        self.ensure(n + S_new)
        es = k_new.element_size()
        if (k_new.stride() != v_new.stride() or k_new.stride(3) != 1 or k_new.stride(2) != hd
                or (k_new.stride(0) * es) % 16 or (k_new.stride(1) * es) % 16
                or (k_new.data_ptr() | v_new.data_ptr()) % 16):
            k_new, v_new = k_new.contiguous(), v_new.contiguous()
        self._copy(k_new, v_new, layer_idx, n, True)
        self.lengths[layer_idx] = n + S_new
        return n + S_new

    def gather(self, layer_idx, n=None):
        """contiguous copies [B, n, Hkv, hd] of the first n cached tokens of a layer"""
        n = self.lengths[layer_idx] if n is None else n
        k = torch.empty((self.batch, n, self.Hkv, self.hd), dtype=self.dtype, device=self.device)
        v = torch.empty_like(k)
        if n:
            self._copy(k, v, layer_idx, 0, False)
        return k, v

    def append(self, k_new, v_new, layer_idx):
        """k_new/v_new: [B, S_new, Hkv, hd].  Stores them and returns contiguous [B, total, Hkv, hd] K/V including the new
        tokens: the inputs themselves for a prefill into an empty cache (no copy), a gather otherwise.  The bf16 decode
        step does not come through here (llama.py: write() + ops.decode_attention_paged read the pages directly)."""
        total = self.write(k_new, v_new, layer_idx)
        if total == k_new.shape[1]:
            return k_new, v_new
        return self.gather(layer_idx, total)

    def advance(self, n=1):
        for i in range(len(self.lengths)):
            self.lengths[i] += n

    def crop(self, max_length):
        for i in range(len(self.lengths)):
            self.lengths[i] = min(self.lengths[i], max_length)
        keep = (max(self.lengths, default=0) + self.PAGE - 1) // self.PAGE
        for b in self.blocks:
            while len(b) > keep:
                self.free.append(b.pop())
                self._table_dirty = True

    def reorder_cache(self, beam_idx):
        """beam search: sequence i continues from old sequence beam_idx[i].  Full pages are immutable, so beams could share
        them; the trailing partial page cannot be shared, so the simple, always-correct route is taken: re-materialise."""
        idx = beam_idx.to(self.device)
        lens = list(self.lengths)
        layers = [tuple(t.index_select(0, idx) for t in self.gather(i)) for i in range(len(lens))]
        nl, st = self.n_layers, self.slab_tokens
        self.reset()
        self.n_layers, self.slab_tokens = nl, st
        for i, (k, v) in enumerate(layers):
            self.write(k, v, i)
        self.lengths = lens

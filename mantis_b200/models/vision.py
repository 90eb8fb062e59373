"""SigLIP / CLIP vision towers on the mantis_b200 kernels.

Mirror transformers' SiglipVisionModel (siglip/modeling_siglip.py:116-710) and CLIPVisionModel
(clip/modeling_clip.py) -- same module tree / parameter names (`vision_model.embeddings.patch_embedding`,
`...position_embedding`, `encoder.layers.N.{layer_norm1,self_attn.{q,k,v,out}_proj,layer_norm2,mlp.fc{1,2}}`,
`post_layernorm`, SigLIP `head.*`, CLIP `class_embedding` / `pre_layrnorm`) so reference checkpoints load
unchanged.  Patch embedding is an im2col + tcgen05 GEMM (K padded to a TMA-legal multiple), attention is
bidirectional.  Only the encoder layers needed for `hidden_states[vision_feature_layer]` are evaluated: the
reference computes (and discards) the last layer, post_layernorm and the SigLIP MAP head
(mantis/models/mllava/modeling_llava.py:456-458).
"""
import torch
from torch import nn
from transformers import PreTrainedModel
from transformers.modeling_outputs import BaseModelOutputWithPooling

from .. import ops
from .layers import B200LayerNorm, B200Linear, init_module_weights


class B200VisionAttention(nn.Module):
    """Bidirectional MHA of SigLIP / CLIP (hf: siglip/modeling_siglip.py:252-312).

    Frozen bf16 towers take the tensor-core route: head_dim (72 for so400m) is zero-padded to 128 *inside cached fused
    QKV / out-proj weights*, so q/k/v come out of ONE tcgen05 GEMM already in the [B, L, H, 128] layout the tcgen05
    flash-attention kernel wants (padded lanes are exactly zero, so q.k and p.v are unchanged; the softmax scale stays
    head_dim**-0.5).  Trainable / fp32 / tiny towers use the per-projection kernels + SIMT attention."""

    PAD_HD = 128

    def __init__(self, config):
        super().__init__()
        self.embed_dim = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = self.embed_dim // self.num_heads
        self.scale = self.head_dim ** -0.5
        self.q_proj = B200Linear(self.embed_dim, self.embed_dim)
        self.k_proj = B200Linear(self.embed_dim, self.embed_dim)
        self.v_proj = B200Linear(self.embed_dim, self.embed_dim)
        self.out_proj = B200Linear(self.embed_dim, self.embed_dim)
        self._pad_cache = None
        self._pad_key = None

    def _padded_weights(self):
        ws = (self.q_proj.weight, self.k_proj.weight, self.v_proj.weight, self.out_proj.weight)
        key = tuple((w.data_ptr(), w._version) for w in ws) + (ws[0].dtype, ws[0].device)
        if self._pad_key != key:
            H, hd, P, d = self.num_heads, self.head_dim, self.PAD_HD, self.embed_dim
            with torch.no_grad():
                wqkv = torch.zeros((3, H, P, d), dtype=ws[0].dtype, device=ws[0].device)
                bqkv = torch.zeros((3, H, P), dtype=ws[0].dtype, device=ws[0].device)
                for i, lin in enumerate((self.q_proj, self.k_proj, self.v_proj)):
                    wqkv[i, :, :hd] = lin.weight.view(H, hd, d)
                    if lin.bias is not None:
                        bqkv[i, :, :hd] = lin.bias.view(H, hd)
                wo = torch.zeros((d, H, P), dtype=ws[0].dtype, device=ws[0].device)
                wo[:, :, :hd] = self.out_proj.weight.view(d, H, hd)
            self._pad_cache = (wqkv.view(3 * H * P, d), bqkv.view(3 * H * P), wo.view(d, H * P))
            self._pad_key = key
        return self._pad_cache

    def _can_pad(self, x, key_mask):
        frozen = not (torch.is_grad_enabled() and (self.q_proj.weight.requires_grad or x.requires_grad))
        return (frozen and x.dtype == torch.bfloat16 and self.head_dim < self.PAD_HD and x.shape[1] >= 64
                and not ops.FORCE_GENERIC)

    def forward(self, x, residual, key_mask=None):
        B, L, _ = x.shape
        if self._can_pad(x, key_mask):
            wqkv, bqkv, wo = self._padded_weights()
            H, P = self.num_heads, self.PAD_HD
            qkv = ops.gemm(x.reshape(B * L, self.embed_dim), wqkv, bias=bqkv).view(B, L, 3, H, P)
            o, _ = ops.attention_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], False, key_mask, self.scale)
            out = ops.gemm(o.view(B * L, H * P), wo, bias=self.out_proj.bias,
                           addend=residual.reshape(B * L, self.embed_dim) if residual is not None else None)
            return out.view(B, L, self.embed_dim)
        q = self.q_proj(x).view(B, L, self.num_heads, self.head_dim)
        k = self.k_proj(x).view(B, L, self.num_heads, self.head_dim)
        v = self.v_proj(x).view(B, L, self.num_heads, self.head_dim)
        o = ops.attention(q, k, v, causal=False, kmask=key_mask, scale=self.scale)
        return self.out_proj(o.view(B, L, self.embed_dim), residual=residual)


class B200VisionMLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.act = config.hidden_act
        self.fc1 = B200Linear(config.hidden_size, config.intermediate_size)
        self.fc2 = B200Linear(config.intermediate_size, config.hidden_size)

    def forward(self, x, residual=None):
        return self.fc2(self.fc1(x, act=self.act), residual=residual)


class B200VisionEncoderLayer(nn.Module):
    """pre-LN block shared by SigLIP, CLIP and the Idefics2 vision transformer"""

    def __init__(self, config):
        super().__init__()
        self.self_attn = B200VisionAttention(config)
        self.layer_norm1 = B200LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.mlp = B200VisionMLP(config)
        self.layer_norm2 = B200LayerNorm(config.hidden_size, eps=config.layer_norm_eps)

    def forward(self, x, key_mask=None):
        x = self.self_attn(self.layer_norm1(x), x, key_mask)
        x = self.mlp(self.layer_norm2(x), residual=x)
        return x


class B200VisionEncoder(nn.Module):
    """== CLIPEncoder / SiglipEncoder: `.layers`"""

    def __init__(self, config):
        super().__init__()
        self.layers = nn.ModuleList([B200VisionEncoderLayer(config) for _ in range(config.num_hidden_layers)])
        self.gradient_checkpointing = False

    def forward(self, x, num_layers=None, collect=False, key_mask=None):
        hs = [x] if collect else None
        n = len(self.layers) if num_layers is None else num_layers
        for layer in self.layers[:n]:
            if self.gradient_checkpointing and self.training and x.requires_grad:
                x = torch.utils.checkpoint.checkpoint(layer, x, key_mask, use_reentrant=False)
            else:
                x = layer(x, key_mask)
            if collect:
                hs.append(x)
        return x, hs


class _PatchEmbedFn(torch.autograd.Function):
    """conv2d(kernel = stride = patch) as im2col + GEMM; differentiable w.r.t. the conv weight / bias
    (pixels never need a gradient on this path)."""

    @staticmethod
    def forward(ctx, pixel_values, weight, bias, patch, w2d, k_pad):
        patches = ops.im2col(pixel_values, patch, k_pad, weight.dtype)
        y = ops.gemm(patches, w2d, bias=bias)
        ctx.save_for_backward(patches if weight.requires_grad else None)
        ctx.wshape = weight.shape
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        (patches,) = ctx.saved_tensors
        gw = gb = None
        gy = gy.contiguous()
        if ctx.needs_input_grad[1] and patches is not None:
            K = ctx.wshape[1] * ctx.wshape[2] * ctx.wshape[3]
            gw2 = ops.gemm(gy, patches, trans_a=True, trans_b=False)            # [out, k_pad]
            gw = gw2[:, :K].reshape(ctx.wshape)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = ops.colsum(gy)
        return None, gw, gb, None, None, None


class PatchEmbedGemm:
    """Keeps a K-padded copy of the flattened conv weight (TMA needs 16-byte multiples: 3*14*14 = 588 -> 640)."""

    def __init__(self):
        self._w = None
        self._key = None

    def weight2d(self, conv_weight, k_pad):
        key = (conv_weight.data_ptr(), conv_weight._version, conv_weight.dtype, k_pad, conv_weight.device)
        if self._key != key:
            out_ch = conv_weight.shape[0]
            w2 = conv_weight.detach().reshape(out_ch, -1)
            if w2.shape[1] != k_pad:
                w = torch.zeros((out_ch, k_pad), dtype=w2.dtype, device=w2.device)
                w[:, : w2.shape[1]] = w2
                w2 = w
            self._w, self._key = w2.contiguous(), key
        return self._w

    def __call__(self, pixel_values, conv, patch):
        K = conv.weight.shape[1] * patch * patch
        k_pad = (K + 63) // 64 * 64 if conv.weight.dtype == torch.bfloat16 else (K + 7) // 8 * 8
        return _PatchEmbedFn.apply(pixel_values, conv.weight, conv.bias, patch, self.weight2d(conv.weight, k_pad), k_pad)


class B200SiglipVisionEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.embed_dim = config.hidden_size
        self.patch_size = config.patch_size
        self.patch_embedding = nn.Conv2d(config.num_channels, self.embed_dim, kernel_size=self.patch_size,
                                         stride=self.patch_size, padding="valid")
        self.num_patches = (config.image_size // self.patch_size) ** 2
        self.num_positions = self.num_patches
        self.position_embedding = nn.Embedding(self.num_positions, self.embed_dim)
        self._pe = PatchEmbedGemm()

    def forward(self, pixel_values):
        N = pixel_values.shape[0]
        x = self._pe(pixel_values, self.patch_embedding, self.patch_size)            # [N*L, d]
        L = x.shape[0] // N
        if L != self.num_positions:
            raise ValueError(f"image gives {L} patches but position table has {self.num_positions}")
        x = ops.add_rows(x, self.position_embedding.weight, period=L)
        return x.view(N, L, self.embed_dim)


class B200SiglipHeadParams(nn.Module):
    """Parameter holder for SiglipMultiheadAttentionPoolingHead so checkpoints round-trip. The MAP head output is
    computed and discarded by the reference hot path (modeling_llava.py:456-458); it is not evaluated here."""

    def __init__(self, config):
        super().__init__()
        d = config.hidden_size
        self.probe = nn.Parameter(torch.randn(1, 1, d))
        self.attention = nn.MultiheadAttention(d, config.num_attention_heads, batch_first=True)
        self.layernorm = nn.LayerNorm(d, eps=config.layer_norm_eps)
        self.mlp = B200VisionMLP(config)


class B200SiglipVisionTransformer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embeddings = B200SiglipVisionEmbeddings(config)
        self.encoder = B200VisionEncoder(config)
        self.post_layernorm = B200LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.use_head = getattr(config, "vision_use_head", True)
        if self.use_head:
            self.head = B200SiglipHeadParams(config)


class B200CLIPVisionEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.embed_dim = config.hidden_size
        self.patch_size = config.patch_size
        self.class_embedding = nn.Parameter(torch.randn(self.embed_dim))
        self.patch_embedding = nn.Conv2d(config.num_channels, self.embed_dim, kernel_size=self.patch_size,
                                         stride=self.patch_size, bias=False)
        self.num_patches = (config.image_size // self.patch_size) ** 2
        self.num_positions = self.num_patches + 1
        self.position_embedding = nn.Embedding(self.num_positions, self.embed_dim)
        self._pe = PatchEmbedGemm()

    def forward(self, pixel_values):
        N = pixel_values.shape[0]
        pe = self._pe(pixel_values, self.patch_embedding, self.patch_size)
        L = pe.shape[0] // N
        x = torch.empty((N, L + 1, self.embed_dim), dtype=pe.dtype, device=pe.device)
        x[:, 0] = self.class_embedding.to(pe.dtype)                                 # layout plumbing (1 row / image)
        x[:, 1:] = pe.view(N, L, self.embed_dim)
        x = ops.add_rows(x.view(N * (L + 1), self.embed_dim), self.position_embedding.weight, period=L + 1)
        return x.view(N, L + 1, self.embed_dim)


class B200CLIPVisionTransformer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embeddings = B200CLIPVisionEmbeddings(config)
        self.pre_layrnorm = B200LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.encoder = B200VisionEncoder(config)
        self.post_layernorm = B200LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class B200VisionPreTrainedModel(PreTrainedModel):
    base_model_prefix = "vision_model"
    main_input_name = "pixel_values"
    supports_gradient_checkpointing = True
    _no_split_modules = ["B200VisionEncoderLayer"]
    _supports_sdpa = True
    _supports_flash_attn = True
    _supports_flash_attn_2 = True

    def _init_weights(self, module):
        init_module_weights(module, getattr(self.config, "initializer_range", 0.02))

    def get_input_embeddings(self):
        return self.vision_model.embeddings.patch_embedding

    def features(self, pixel_values, feature_layer):
        """hidden_states[feature_layer] of the HF model (tuple of N+1 entries), evaluating only what is needed."""
        vm = self.vision_model
        n_layers = len(vm.encoder.layers)
        idx = feature_layer if feature_layer >= 0 else n_layers + 1 + feature_layer
        if not 0 <= idx <= n_layers:
            raise ValueError(f"vision_feature_layer {feature_layer} out of range")
        x = vm.embeddings(pixel_values.to(vm.embeddings.patch_embedding.weight.dtype))
        if hasattr(vm, "pre_layrnorm"):
            x = vm.pre_layrnorm(x)
        x, _ = vm.encoder(x, num_layers=idx)
        return x

    def forward(self, pixel_values=None, output_attentions=None, output_hidden_states=None, return_dict=None, **kw):
        vm = self.vision_model
        x = vm.embeddings(pixel_values.to(vm.embeddings.patch_embedding.weight.dtype))
        if hasattr(vm, "pre_layrnorm"):
            x = vm.pre_layrnorm(x)
        x, hs = vm.encoder(x, collect=bool(output_hidden_states))
        last = vm.post_layernorm(x) if not hasattr(vm, "pre_layrnorm") else x
        pooled = None
        if hasattr(vm, "pre_layrnorm"):
            pooled = vm.post_layernorm(x[:, 0, :])
        return BaseModelOutputWithPooling(last_hidden_state=last, pooler_output=pooled,
                                          hidden_states=tuple(hs) if hs is not None else None, attentions=None)


class B200SiglipVisionModel(B200VisionPreTrainedModel):
    def __init__(self, config):
        super().__init__(config)
        self.vision_model = B200SiglipVisionTransformer(config)
        self.post_init()


class B200CLIPVisionModel(B200VisionPreTrainedModel):
    def __init__(self, config):
        super().__init__(config)
        self.vision_model = B200CLIPVisionTransformer(config)
        self.post_init()


def build_vision_tower(vision_config):
    mt = getattr(vision_config, "model_type", "")
    if mt in ("siglip_vision_model", "siglip"):
        return B200SiglipVisionModel(vision_config)
    if mt in ("clip_vision_model", "clip"):
        return B200CLIPVisionModel(vision_config)
    raise ValueError(f"mantis_b200 supports SigLIP and CLIP vision towers, got model_type={mt!r}")

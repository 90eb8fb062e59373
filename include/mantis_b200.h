/* mantis_b200 -- C ABI of the B200-native (sm_100a) kernels behind the Mantis interleaved multi-image hot path.
 *
 * The reference (TIGER-AI-Lab/Mantis) is pure Python over HuggingFace transformers: it has no FFI / plugin
 * layer, its boundary is the nn.Module API.  This header is therefore the boundary
 * *beneath* the module shell in mantis_b200/models: every entry point replaces the ATen / cuBLAS / flash-attn
 * call sequence of the reference code cited next to it.
 *
 * Conventions: raw device pointers + explicit sizes/strides; no allocation, no ownership transfer, no implicit
 * synchronisation; `stream` is a cudaStream_t passed as void*; `dtype` is MB200_DTYPE_{F32,BF16}; every function
 * returns 0 or a negative errno (-EINVAL bad argument, -ENOTSUP operands not eligible for the fast path, -EIO launch
 * failure; mb200_last_error() has the detail).  All entry points are re-entrant and thread-safe.
 */
#ifndef MANTIS_B200_H
#define MANTIS_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MB200_DTYPE_F32 0
#define MB200_DTYPE_BF16 1

/* ---- library ---------------------------------------------------------------------------------------------- */
int mb200_version(void);
const char* mb200_last_error(void);
int mb200_check_device(void);
int mb200_num_sms(void);

/* ---- image-token merge: LlavaForConditionalGeneration._merge_input_ids_with_image_features
 *      (mantis/models/mllava/modeling_llava.py:293-360), bit-exact integer outputs ------------------------- */
long long mb200_merge_ws_bytes(int B, int T_len);
int mb200_merge_plan(const int64_t* ids, const void* embeds, int dtype, int B, int T_len, int D, int P,
                     int64_t image_token, int64_t pad_token, void* ws, int64_t* header_dev, void* stream);
int mb200_merge_index(const int64_t* ids, const int64_t* attn, const int64_t* labels, const void* ws, int B,
                      int T_len, int P, int S, int left_padding, int64_t image_token, int64_t ignore_index,
                      int32_t* srcmap, int64_t* out_mask, int64_t* out_labels, int64_t* out_pos, void* stream);
int mb200_merge_rows(const int32_t* srcmap, const void* text, const void* img, void* out, int B, int S, int T_len,
                     long long row_bytes, long long n_img_rows, void* stream);
int mb200_merge_rows_bwd(const int32_t* srcmap, const void* gout, void* gtext, void* gimg, int B, int S, int T_len,
                         long long row_bytes, long long n_img_rows, void* stream);

/* ---- token embedding gather / scatter-add (modeling_llava.py:427 -> nn.Embedding) ------------------------- */
int mb200_embedding_fwd(const int64_t* ids, const void* table, void* out, long long n, int D, long long V, int dtype,
                        void* stream);
int mb200_embedding_bwd(const int64_t* ids, const void* gout, void* gtable, long long n, int D, long long V, int dtype,
                        void* stream);

/* ---- norms: LlamaRMSNorm (transformers llama/modeling_llama.py:53-68), Idefics2RMSNorm
 *      (mantis/models/idefics2/modeling_idefics2.py:795-809), nn.LayerNorm (siglip/modeling_siglip.py:334-336) */
int mb200_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, long long n, int D, float eps, int dtype,
                      void* stream);
int mb200_norm_bwd_parts(long long n);
int mb200_rmsnorm_bwd(const void* x, const void* w, const void* dy, const float* rstd, void* dx, float* dw_part,
                      void* dw, int accumulate_dw, int accumulate_dx, long long n, int D, int dtype, void* stream);
/* dx = rmsnorm backward of dy + dres (the residual-branch gradient that by-passes the norm, llama/modeling_llama.py:
 * hidden = residual + sublayer(norm(hidden))): one pass instead of the norm backward + autograd's elementwise sum */
int mb200_rmsnorm_bwd_res(const void* x, const void* w, const void* dy, const float* rstd, const void* dres, void* dx,
                          float* dw_part, void* dw, int accumulate_dw, long long n, int D, int dtype, void* stream);
int mb200_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, long long n,
                        int D, float eps, int dtype, void* stream);
int mb200_layernorm_bwd(const void* x, const void* w, const void* dy, const float* mean, const float* rstd, void* dx,
                        float* dw_part, float* db_part, void* dw, void* db, int accumulate, long long n, int D,
                        int dtype, void* stream);

/* ---- RoPE, rotate_half form (transformers llama/modeling_llama.py:146-168); y may alias x ------------------- */
int mb200_rope(const void* x, void* y, const int64_t* pos, const float* inv_freq, long long n_tok, int H, int hd,
               long long tok_stride, long long out_stride, float attn_scaling, int backward, int dtype, void* stream);

/* cos/sin table of one forward pass (shared by all layers) + q and k rotated in one vectorised launch from it */
int mb200_rope_table(const int64_t* pos, const float* inv_freq, void* tab, long long n_tok, int hd, float attn_scaling,
                     int dtype, void* stream);
int mb200_rope2_bf16(const void* q, const void* k, void* qo, void* ko, const void* tab, long long n_tok, int Hq, int Hk,
                     int hd, long long q_stride, long long k_stride, int backward, void* stream);

/* ---- SwiGLU (llama/modeling_llama.py:182-184; idefics2 :506-521) and GELU family
 *      (projector modeling_llava.py:110-118; SigLIP / CLIP MLP).  kind: 0 erf, 1 tanh, 2 quick ------------- */
int mb200_swiglu_fwd(const void* gate, const void* up, void* out, long long n, int dtype, void* stream);
int mb200_swiglu_bwd(const void* gate, const void* up, const void* dout, void* dgate, void* dup, long long n,
                     int dtype, void* stream);
int mb200_act_fwd(const void* x, void* y, long long n, int kind, int dtype, void* stream);
int mb200_act_bwd(const void* x, const void* dy, void* dx, long long n, int kind, int dtype, void* stream);

/* ---- residual add, position-embedding add, bias gradient, im2col (patch-embed conv as GEMM:
 *      siglip/modeling_siglip.py:124-130,175-186), dtype cast ----------------------------------------------- */
int mb200_add(const void* a, const void* b, void* y, long long n, int dtype, void* stream);
int mb200_add_rows(const void* x, const void* table, const int64_t* idx, void* y, long long n, int D, long long period,
                   int dtype, void* stream);
int mb200_colsum_parts(long long n);
int mb200_colsum(const void* x, float* part, void* out, int accumulate, long long n, int N, long long ld, int dtype,
                 void* stream);
int mb200_im2col(const void* px, int px_dtype, void* out, int out_dtype, int N, int C, int H, int W, int p, int Kpad,
                 void* stream);
int mb200_cast(const void* x, int in_dtype, void* y, int out_dtype, long long n, void* stream);
/* Image processor tail on the device (replaces the numpy rescale + normalize + transpose of the HF image
 * processor called from MLlavaProcessor.__call__, mantis/models/mllava/processing_llava.py:226-252): uint8 pixels
 * [N,H,W,C] (channels_last) or [N,C,H,W] -> out[n,c,h,w] = lut[c][pixel]; lut is fp32 [C][256] on the device. */
int mb200_image_normalize_u8(const void* px, const float* lut, void* out, int out_dtype, int N, int C, int H, int W,
                             int channels_last, void* stream);
/* One pass of Pillow's antialiased 8-bit resampler (what `PIL.Image.resize` does inside the reference's image processor):
 * [in_h, in_w, C] uint8 -> [in_h, out_len, C] (horizontal) or [out_len, in_w, C]; bounds int32 [out_len][2] = (first tap,
 * taps), coef int32 [out_len][ksize] in 22-bit fixed point, both computed on the host like Pillow's precompute_coeffs. */
int mb200_resize_u8_pass(const void* in, void* out, const int* bounds, const int* coef, int ksize, int in_h, int in_w,
                         int out_len, int C, int horizontal, void* stream);
/* flags[r] = 1 iff row r is all zeros: Idefics2 padding-image removal (mantis/models/idefics2/modeling_idefics2.py:1637-1639) */
int mb200_rows_all_zero(const void* x, long long n_rows, long long row_elems, int* flags, int dtype, void* stream);

/* ---- shifted masked cross-entropy (modeling_llava.py:523-537; modeling_idefics2.py:1883-1899) and AdamW ---- */
int mb200_shift_labels(const int64_t* labels, const int64_t* mask, int64_t* out, int B, int S, int64_t ignore_index,
                       float* count_out, void* stream);
int mb200_ce_fwd_bwd(const void* logits, const int64_t* labels, float* loss_rows, float* lse_rows, void* dlogits,
                     long long n, int V, long long ld, const float* gscale_ptr, float gscale_const, int dtype,
                     void* stream);
/* idx_out[j] = index of the j-th row with labels[row] >= 0 (ascending; at most cap entries written), *count_out = their number.
 * Replaces the boolean-mask gather of the reference's loss (modeling_llava.py:526-531 `shift_logits[shift_attention_mask != 0]`
 * + ignore_index) and the host read-back its data-dependent size costs: the caller knows the count from the host labels. */
int mb200_compact_valid_rows(const int64_t* labels, long long n, int64_t* idx_out, long long cap, int* count_out, void* stream);
int mb200_ce_reduce(const float* loss_rows, const int64_t* labels, long long n, int V, float* out2, int accumulate,
                    void* stream);
/* Fused AdamW over FLAT parameter / gradient / moment buffers (one launch per optimizer step) with fp32 master weights.
 * Replaces torch.optim.AdamW on DeepSpeed's fp32 master copy (mantis/train/zero_configs/zero3.json "bf16": enabled,
 * mantis/train/scripts/train_mllava.sh:148,162-165 `--bf16 True --learning_rate 1e-5 --weight_decay 0.`).
 *   p_dtype BF16: p = bf16 weights (what the model computes with), lo = their 16 low fp32 bits (uint16): the pair IS the
 *                 fp32 master, master_bits = (bf16_bits << 16) + (int16) lo; lo must not be NULL.
 *   p_dtype F32 : p = fp32 weights, lo ignored.
 *   g (g_dtype F32 or BF16): accumulated gradients, multiplied by grad_scale (1 / world size) and by the global-norm clip
 *                 factor min(1, max_norm / (sqrt(*norm_sq) * grad_scale + 1e-6)) computed ON THE DEVICE (norm_sq NULL or
 *                 max_norm <= 0: no clipping); zero_grad != 0 clears g on the way out.
 *   blk_group[n / 1024] (nullable): 1 = this 1024-element block takes no weight decay (biases, norm weights).
 * n must be a multiple of 8 and all buffers 16-byte aligned (every tensor's slice starts on a 1024-element boundary). */
int mb200_adamw_flat(void* p, void* lo, void* g, float* m, float* v, const unsigned char* blk_group, long long n, float lr,
                     float beta1, float beta2, float eps, float wd, int step, float grad_scale, const float* norm_sq,
                     float max_norm, int zero_grad, int p_dtype, int g_dtype, void* stream);
/* fp32 master <-> (bf16 weight, uint16 low half) over a flat range (trainer construction, optimizer checkpoints). lo may be
 * NULL in join (treated as zero). */
int mb200_master_split(const float* master, void* hi_bf16, void* lo_u16, long long n, void* stream);
int mb200_master_join(const void* hi_bf16, const void* lo_u16, float* master, long long n, void* stream);
/* dst (fp32) = (accumulate ? dst : 0) + scale * src (src_dtype): folds a gradient autograd produced in the parameter dtype
 * into the fp32 main gradient buffer (what DeepSpeed's fp32 gradient accumulation does); accumulate = 0 for the first
 * micro-batch of an optimizer step (the buffer is then never zero-filled). */
int mb200_accum_f32(float* dst, const void* src, long long n, float scale, int accumulate, int src_dtype, void* stream);
int mb200_sumsq(const void* g, long long n, float* out, int dtype, void* stream);

/* ---- GEMM: nn.Linear forward / dgrad / wgrad (llama/modeling_llama.py:171-184,238-249,487; siglip :270-273,
 *      :320-321; projector modeling_llava.py:110-118).  mb200_gemm_bf16 = tcgen05 + TMEM + TMA path;
 *      mb200_gemm_generic = shape-agnostic SIMT path (tiny configs, fp32 parity, cross-check) ---------------- */
int mb200_gemm_generic(const void* A, const void* B, void* C, const void* bias, int M, int N, int K, long long lda,
                       long long ldb, long long ldc, int transA, int transB, float alpha, float beta, int batch,
                       long long strideA, long long strideB, long long strideC, int dtype_ab, int dtype_c,
                       void* stream);
int mb200_gemm_bf16(const void* A, const void* B, void* C, const void* bias, const void* addend, int M, int N, int K,
                    long long lda, long long ldb, long long ldc, long long ld_add, int transA, int transB, int act,
                    void* stream);

/* C32[M,N] (fp32) = (accumulate ? C32 : 0) + op(A) op(B): bf16 operands, fp32 accumulation kept in fp32 all the way to
 * memory -- the weight-gradient product dW += dy^T x into the fp32 main-gradient buffer (same operand rules as above). */
int mb200_gemm_bf16_acc32(const void* A, const void* B, float* C32, int M, int N, int K, long long lda, long long ldb,
                          long long ldc, int transA, int transB, int accumulate, void* stream);
/* SwiGLU MLP with the activation fused into the projections around it (llama/modeling_llama.py:182-184
 * `down_proj(act_fn(gate_proj(x)) * up_proj(x))`; idefics2 :506-521), same tcgen05 kernels and operand rules as above:
 *   swiglu_fwd: U[M,N] = X[M,K] Wu[N,K]^T, Act[M,N] = silu(G) * U  (G = the gate projection computed just before; U is kept for
 *               backward).  Replaces the up-projection GEMM + the elementwise SwiGLU kernel.
 *   swiglu_bwd: d_act[M,N] = dY[M,K] Wd[K,N] is formed in TMEM only; dG = d_act * U * silu'(G), dU = d_act * silu(G) are
 *               written.  Replaces the down-projection dgrad GEMM + the elementwise SwiGLU backward.
 * G, U, Act, dG, dU share the leading dimension ld; arithmetic and bf16 rounding points equal mb200_swiglu_fwd / _bwd. */
int mb200_gemm_bf16_swiglu_fwd(const void* X, const void* Wu, const void* G, void* U, void* Act, int M, int N, int K,
                               long long lda, long long ldb, long long ld, void* stream);
int mb200_gemm_bf16_swiglu_bwd(const void* dY, const void* Wd, const void* G, const void* U, void* dG, void* dU, int M, int N,
                               int K, long long lda, long long ldb, long long ld, void* stream);
/* CTA-pair (cta_group::2, 256x256 tile per SM pair) variant of mb200_gemm_bf16; identical contract. */
int mb200_gemm_bf16_2cta(const void* A, const void* B, void* C, const void* bias, const void* addend, int M, int N, int K,
                         long long lda, long long ldb, long long ldc, long long ld_add, int transA, int transB, int act,
                         void* stream);

/* ---- attention: softmax(q k^T * scale + causal/padding mask) v, GQA (llama/modeling_llama.py:199-289;
 *      siglip/modeling_siglip.py:229-303; idefics2 perceiver :812-910).  q/o [B,Sq,H,hd], k/v [B,Sk,Hkv,hd];
 *      strides = {q_b,q_s,q_h, k_b,k_s,k_h, v_b,v_s,v_h, o_b,o_s,o_h} in elements.
 *      `causal` of the generic kernels: 0 = none, 1 = causal (key j visible to query i iff j <= i + Sk - Sq), W > 1 = causal
 *      with Mistral's sliding window of W keys (additionally (i + Sk - Sq) - j < W; mistral/modeling_mistral.py) ----------- */
int mb200_attn_generic_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Hkv,
                           int Sq, int Sk, int hd, const long long* strides, float scale, int causal,
                           const int64_t* kmask, long long kmask_sb, int dtype, void* stream);
int mb200_attn_generic_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout,
                           const float* lse, float* delta, void* dq, void* dk, void* dv, int B, int H, int Hkv, int Sq,
                           int Sk, int hd, const long long* strides, float scale, int causal, const int64_t* kmask,
                           long long kmask_sb, int dtype, void* stream);

/* tcgen05 flash attention (bf16, head_dim 128): same semantics as mb200_attn_generic_fwd; -ENOTSUP if ineligible.
 * kbits_ws: B * mb200_attn_kbits_words(Sk) uint32 of device scratch, needed only when kmask != NULL. */
long long mb200_attn_kbits_words(int Sk);
int mb200_kmask_bits(const int64_t* kmask, long long kmask_sb, void* bits, int B, int Sk, void* stream);
int mb200_attn_fwd_bf16(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Hkv, int Sq,
                        int Sk, int hd, const long long* strides, float scale, int causal, const int64_t* kmask,
                        long long kmask_sb, void* kbits_ws, void* stream);

/* two-query-tiles-per-CTA variant of mb200_attn_fwd_bf16 (tiles ping-pong on the tensor pipe and share K/V loads);
 * kbits = bitmask from mb200_kmask_bits (NULL = no key padding mask) */
int mb200_attn_fwd2_bf16(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Hkv, int Sq,
                         int Sk, int hd, const long long* strides, float scale, int causal, const void* kbits,
                         int kbits_stride, void* stream);

/* tcgen05 flash attention backward (dK/dV kernel + dQ kernel, deterministic, no atomics).  dq/dk/dv/dout contiguous;
 * kbits = the bitmask scratch filled by mb200_attn_fwd_bf16 for the same kmask (NULL iff kmask == NULL);
 * delta = fp32 scratch of 2 * B * H * mb200_attn_bwd_sq_pad(Sq) floats. */
long long mb200_attn_bwd_sq_pad(int Sq);
int mb200_attn_bwd_bf16(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                        float* delta, void* dq, void* dk, void* dv, int B, int H, int Hkv, int Sq, int Sk, int hd,
                        const long long* strides, float scale, int causal, const int64_t* kmask, long long kmask_sb,
                        const void* kbits, void* stream);
/* Single-pass backward for self-attention (Sq == Sk): the dK/dV kernel writes dS^T (bf16) once into ds_ws
 * (mb200_attn_bwd_ds_bytes(B, H, Sq) bytes) and dQ = scale * dS K is a tensor-core GEMM over it -- S and dP are computed and
 * read back from tensor memory once instead of once per kernel.  ds_ws == NULL or Sq != Sk: same as mb200_attn_bwd_bf16. */
long long mb200_attn_bwd_ds_bytes(int B, int H, int Sq);
int mb200_attn_bwd_bf16_sp(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                           float* delta, void* dq, void* dk, void* dv, int B, int H, int Hkv, int Sq, int Sk, int hd,
                           const long long* strides, float scale, int causal, const int64_t* kmask, long long kmask_sb,
                           const void* kbits, void* ds_ws, void* stream);

/* ---- decode-time kernels of generate() (q_len 1): skinny GEMM (M <= 16 rows, every weight byte read once), KV-cache
 *      append, split-KV attention + combine (hf: llama/modeling_llama.py:269-270; ref modeling_llava.py:477-519) ---- */
int mb200_skinny_gemm_bf16(const void* X, const void* W, void* C, const void* bias, const void* addend, int M, int N,
                           int K, long long ldx, long long ldw, long long ldc, long long ld_add, void* stream);
int mb200_skinny_gemm3_bf16(const void* X, const void* W0, const void* W1, const void* W2, void* C0, void* C1, void* C2,
                            int M, int N0, int N1, int N2, int K, long long ldx, long long ldw, void* stream);
int mb200_skinny_swiglu_bf16(const void* X, const void* Wg, const void* Wu, void* C, int M, int N, int K, long long ldx,
                             long long ldw, long long ldc, void* stream);
int mb200_rope_append_bf16(const void* q, const void* k, const void* v, void* q_out, void* k_cache, void* v_cache,
                           const int64_t* pos, const float* inv_freq, int B, int H, int Hkv, int hd, int ctx,
                           long long capacity, float rope_scale, void* stream);
int mb200_kv_append(const void* k_new, const void* v_new, void* k_cache, void* v_cache, const int* pos_dev, int pos_const,
                    int B, int row_elems, long long ld_new, long long capacity, void* stream);
int mb200_decode_attn_splits(int ctx);
int mb200_decode_attn_bf16(const void* q, const void* k, const void* v, void* o, float* part, int B, int H, int Hkv,
                           int ctx, int hd, long long q_sb, long long q_sh, long long kv_sb, long long kv_ss,
                           long long kv_sh, long long o_sb, long long o_sh, float scale, const void* kbits,
                           int kbits_stride, void* stream);

/* ---- native decode step: the whole single-token LLaMA/Mistral step (all layers, final norm, LM head, greedy argmax)
 *      enqueued by one call (hf: llama/modeling_llama.py:375-427 at q_len 1).  See decode_engine.cu for the tables. ---- */
/* RMSNorm fused into the decode projections (hf: llama/modeling_llama.py:53-68 + 171-184 / 225-262): the activations are
 * normalised inside the skinny GEMM when M <= 2, else into xn_scratch [M, K] first.  Same results as mb200_rmsnorm_fwd +
 * mb200_skinny_gemm3_bf16 / mb200_skinny_swiglu_bf16 up to the summation order of the row norm. */
int mb200_skinny_gemm3_norm_bf16(const void* X, const void* gamma, float eps, void* xn_scratch, const void* W0, const void* W1,
                                 const void* W2, void* C0, void* C1, void* C2, int M, int N0, int N1, int N2, int K,
                                 long long ldw, void* stream);
int mb200_skinny_swiglu_norm_bf16(const void* X, const void* gamma, float eps, void* xn_scratch, const void* Wg,
                                  const void* Wu, void* C, int M, int N, int K, long long ldw, long long ldc, void* stream);
/* Paged KV cache for generate() (replaces transformers' DynamicCache torch.cat growth used by the reference's
 * prepare_inputs_for_generation, mantis/models/mllava/modeling_llava.py:551-602).  A page holds
 * mb200_kv_page_tokens() tokens of all layers, [L][2 (k,v)][128][Hkv][hd]; the block table is int64 [B, table_stride]
 * of page base addresses.  *_off arguments are in elements for the bf16 kernels and in bytes for mb200_kv_page_copy. */
int mb200_kv_page_tokens(void);
int mb200_kv_page_copy(void* k_lin, void* v_lin, const int64_t* table, int table_stride, long long layer_off_bytes,
                       long long v_off_bytes, int B, int S, int start, int row_bytes, long long lin_sb, long long lin_ss,
                       int to_pages, void* stream);
int mb200_rope_append_paged_bf16(const void* q, const void* k, const void* v, void* q_out, const int64_t* table,
                                 int table_stride, long long layer_off, long long v_off, const int64_t* pos,
                                 const float* inv_freq, int B, int H, int Hkv, int hd, int ctx, float rope_scale,
                                 void* stream);
int mb200_decode_attn_paged_bf16(const void* q, const int64_t* table, int table_stride, long long layer_off, long long v_off,
                                 void* o, float* part, int B, int H, int Hkv, int ctx, int hd, long long q_sb,
                                 long long q_sh, long long o_sb, long long o_sh, float scale, const void* kbits,
                                 int kbits_stride, void* stream);
int mb200_argmax_bf16(const void* logits, long long ld, int B, int V, int64_t* out, void* stream);
long long mb200_decode_ws_bytes(int B, int hidden, int n_heads, int n_kv_heads, int head_dim, int inter, int ctx_max);
int mb200_llama_decode_step(const int* dims, const float* fparm, const void* const* layers, const void* const* misc,
                            void* ws, long long ld_logits, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MANTIS_B200_H */

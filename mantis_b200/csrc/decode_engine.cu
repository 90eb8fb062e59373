// Native decode step: one C call enqueues the whole LLaMA/Mistral single-token step (all layers + final norm + LM head +
// greedy argmax) on a stream -- ~15 launches per layer issued from C++ instead of ~420 Python/ctypes round trips per
// token, which is what bounds generate() once the kernels themselves are HBM-bound.
// Mirrors LlamaModel.forward at q_len == 1 (hf: llama/modeling_llama.py:375-427,225-333) with the KV cache of
// mantis_b200/models/kv_cache.py (token-major [B, capacity, Hkv, hd]).
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "../../include/mantis_b200.h"
#include <stdlib.h>

namespace {
__global__ void __launch_bounds__(1024)
argmax_kernel(const bf16* __restrict__ logits, long long ld, int V, int64_t* __restrict__ out) {
  __shared__ float bv[32]; __shared__ int bi[32];
  mb::pdl_trigger();
  mb::pdl_wait();
  const bf16* row = logits + (size_t)blockIdx.x * ld;
  float best = -INFINITY; int idx = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float v = __bfloat162float(row[i]);
    if (v > best || (v == best && i < idx)) { best = v; idx = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { bv[w] = best; bi[w] = idx; }
  __syncthreads();
  if (w == 0) {
    best = bv[lane]; idx = bi[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
    }
    if (lane == 0) out[blockIdx.x] = idx;
  }
}

// Cluster variant (the default): 8 CTAs of one cluster split a row, partial winners meet in CTA 0's shared memory over
// DSMEM -- 16 elements per thread instead of 125, no scratch buffer, no second launch.  Ties -> smallest index.
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(1024)
argmax_cluster_kernel(const bf16* __restrict__ logits, long long ld, int V, int64_t* __restrict__ out) {
  __shared__ float bv[32]; __shared__ int bi[32];
  __shared__ float cv[8]; __shared__ int ci[8];
  mb::pdl_trigger();
  mb::pdl_wait();
  const uint32_t rank = sm100::cluster_ctarank();
  const bf16* row = logits + (size_t)blockIdx.y * ld;
  const int nvec = (V + 7) / 8, per = (nvec + 7) / 8;
  const int v_end = min(nvec, (int)(rank + 1) * per);
  float best = -INFINITY; int idx = 0x7fffffff;
  for (int v = rank * per + threadIdx.x; v < v_end; v += blockDim.x) {
    const int4 raw = mb::ld_stream(reinterpret_cast<const int4*>(row) + v);       // ld % 8 == 0: the row tail is padding
    const bf16* e = reinterpret_cast<const bf16*>(&raw);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = v * 8 + j;
      const float x = __bfloat162float(e[j]);
      if (i < V && x > best) { best = x; idx = i; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { bv[w] = best; bi[w] = idx; }
  __syncthreads();
  if (w == 0) {
    best = bv[lane]; idx = bi[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
    }
    if (lane == 0) {          // deposit in CTA 0's cv/ci[rank]
      uint32_t rv, ri;
      asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(rv) : "r"(sm100::smem_u32(&cv[rank])));
      asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(ri) : "r"(sm100::smem_u32(&ci[rank])));
      asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(rv), "f"(best) : "memory");
      asm volatile("st.shared::cluster.s32 [%0], %1;" :: "r"(ri), "r"(idx) : "memory");
    }
  }
  sm100::cluster_sync_all();
  if (rank == 0 && threadIdx.x == 0) {
    best = cv[0]; idx = ci[0];
#pragma unroll
    for (int r = 1; r < 8; ++r)
      if (cv[r] > best || (cv[r] == best && ci[r] < idx)) { best = cv[r]; idx = ci[r]; }
    out[blockIdx.y] = idx;
  }
}
}  // namespace

#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

extern "C" {

int mb200_argmax_bf16(const void* logits, long long ld, int B, int V, int64_t* out, void* stream) {
  if (B <= 0) return MB200_OK;
  if ((ld & 7) == 0 && !(reinterpret_cast<uintptr_t>(logits) & 15))
    mb::launch_ex(argmax_cluster_kernel, dim3(8, B), dim3(1024), 0, (cudaStream_t)stream, mb::pdl_mode() != 0,
                  (const bf16*)logits, ld, V, out);
  else
    mb::launch_ex(argmax_kernel, dim3(B), dim3(1024), 0, (cudaStream_t)stream, mb::pdl_mode() != 0, (const bf16*)logits, ld, V, out);
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

// dims  : {n_layers, hidden, n_heads, n_kv_heads, head_dim, intermediate, vocab, B, ctx, capacity, kbits_stride,
//          table_stride}   table_stride > 0 selects the paged cache (misc[9] = int64 block table [B, table_stride] of page
//          addresses, page = [L][2][128][Hkv][hd]); layers[..][9..10] are then unused and capacity = table_stride * 128.
// fparm : {rms_eps, rope_scaling}
// layers: n_layers x 11 device pointers {wq, wk, wv, wo, w_gate, w_up, w_down, ln1, ln2, k_cache, v_cache}
// misc  : {embed, final_norm, lm_head, inv_freq, ids(int64[B]), pos(int64[B]), kbits (or null), logits_out [B, ld_logits] bf16,
//          next_ids (int64[B]) or null, block table (paged cache) or null}
// ws    : device scratch, mb200_decode_ws_bytes(...) bytes.   bf16 only; head_dim 128.
long long mb200_decode_ws_bytes(int B, int hidden, int n_heads, int n_kv_heads, int head_dim, int inter, int ctx_max) {
  const long long e = 2;
  long long b = 0;
  b += 2LL * B * hidden * e;                                   // x, xn
  b += (long long)B * (n_heads + 2 * n_kv_heads) * head_dim * e * 2;   // q,k,v + rotated q,k (over-allocated)
  b += (long long)B * n_heads * head_dim * e;                  // attn out
  b += 3LL * B * inter * e;                                    // gate, up, act
  b += (long long)B * n_heads * 64 * (head_dim + 2) * 4;       // split-KV partials (<= 64 splits)
  return b + 4096;
}

// Every kernel of the step after the embedding gather is launched with programmatic stream serialization: each one
// triggers its dependents at entry and calls griddepcontrol.wait before touching the previous kernel's outputs, so launch
// latency, block scheduling and the weight-stream ramp of kernel n+1 hide under the tail of kernel n (MB200_PDL=0 disables).
struct PdlScope {            // thread-local launch mode, restored on every exit path
  explicit PdlScope(int on) { mb::pdl_mode() = on; }
  ~PdlScope() { mb::pdl_mode() = 0; }
};
int mb200_llama_decode_step(const int* dims, const float* fparm, const void* const* layers, const void* const* misc,
                            void* ws, long long ld_logits, void* stream) {
  const int L = dims[0], D = dims[1], H = dims[2], Hkv = dims[3], hd = dims[4], I = dims[5], V = dims[6], B = dims[7];
  const int ctx = dims[8]; const long long cap = dims[9]; const int kbs = dims[10]; const int tstride = dims[11];
  if (B <= 0 || B > 16 || hd != 128) return -ENOTSUP;
  const int64_t* table = tstride > 0 ? (const int64_t*)misc[9] : nullptr;
  if (tstride > 0 && !table) return -EINVAL;
  const long long page_tok = mb200_kv_page_tokens();
  const long long v_off = page_tok * Hkv * hd, layer_stride = 2 * v_off;
  if (ctx + 1 > cap) return -EINVAL;
  const float eps = fparm[0], rope_scale = fparm[1];
  const void* embed = misc[0]; const void* fnorm = misc[1]; const void* lm_head = misc[2];
  const float* inv_freq = (const float*)misc[3];
  const int64_t* ids = (const int64_t*)misc[4]; const int64_t* pos = (const int64_t*)misc[5];
  const void* kbits = misc[6]; void* logits = const_cast<void*>(misc[7]); int64_t* next_ids = (int64_t*)const_cast<void*>(misc[8]);
  const int dt = MB200_DTYPE_BF16;
  char* p = (char*)ws;
  auto take = [&](long long bytes) { void* r = p; p += (bytes + 255) / 256 * 256; return r; };
  void* x = take(2LL * B * D); void* xn = take(2LL * B * D);
  void* q = take(2LL * B * H * hd); void* k = take(2LL * B * Hkv * hd); void* v = take(2LL * B * Hkv * hd);
  void* qr = take(2LL * B * H * hd); void* kr = take(2LL * B * Hkv * hd);
  void* ao = take(2LL * B * H * hd);
  void* g = take(2LL * B * I); void* u = take(2LL * B * I); void* act = take(2LL * B * I);
  float* part = (float*)take(4LL * B * H * 64 * (hd + 2));
  const float scale = 1.0f / sqrtf((float)hd);
  const int HD = H * hd, KD = Hkv * hd;

  // process-wide switches, read once (thread-safe static initialisation); the launch mode itself is scoped to this call on
  // the calling thread (PdlScope), so a decode step issued from a worker thread behaves exactly like one from the main thread
  static const int pdl = [] { const char* e = getenv("MB200_PDL"); return (e && e[0] == '0') ? 0 : 1; }();
  static const int fuse_norm = [] { const char* e = getenv("MB200_DECODE_FUSE_NORM"); return (e && e[0] == '0') ? 0 : 1; }();
  TRY(mb200_embedding_fwd(ids, embed, x, B, D, V, dt, stream));      // plain launch: its inputs come from torch kernels
  PdlScope scope(pdl);
  for (int l = 0; l < L; ++l) {
    const void* const* w = layers + (size_t)l * 11;
    if (fuse_norm) {
      TRY(mb200_skinny_gemm3_norm_bf16(x, w[7], eps, xn, w[0], w[1], w[2], q, k, v, B, HD, KD, KD, D, D, stream));
    } else {
      TRY(mb200_rmsnorm_fwd(x, w[7], xn, nullptr, B, D, eps, dt, stream));
      TRY(mb200_skinny_gemm3_bf16(xn, w[0], w[1], w[2], q, k, v, B, HD, KD, KD, D, D, D, stream));
    }
    if (table) {
      TRY(mb200_rope_append_paged_bf16(q, k, v, qr, table, tstride, l * layer_stride, v_off, pos, inv_freq, B, H, Hkv, hd,
                                       ctx, rope_scale, stream));
      TRY(mb200_decode_attn_paged_bf16(qr, table, tstride, l * layer_stride, v_off, ao, part, B, H, Hkv, ctx + 1, hd, HD, hd,
                                       HD, hd, scale, kbits, kbs, stream));
    } else {
      TRY(mb200_rope_append_bf16(q, k, v, qr, const_cast<void*>(w[9]), const_cast<void*>(w[10]), pos, inv_freq, B, H, Hkv,
                                 hd, ctx, cap, rope_scale, stream));
      TRY(mb200_decode_attn_bf16(qr, w[9], w[10], ao, part, B, H, Hkv, ctx + 1, hd, HD, hd, cap * KD, KD, hd, HD, hd, scale,
                                 kbits, kbs, stream));
    }
    TRY(mb200_skinny_gemm_bf16(ao, w[3], x, nullptr, x, B, D, HD, HD, HD, D, D, stream));          // x += o_proj(attn)
    if (fuse_norm) {
      TRY(mb200_skinny_swiglu_norm_bf16(x, w[8], eps, xn, w[4], w[5], act, B, I, D, D, I, stream));
    } else {
      TRY(mb200_rmsnorm_fwd(x, w[8], xn, nullptr, B, D, eps, dt, stream));
      TRY(mb200_skinny_swiglu_bf16(xn, w[4], w[5], act, B, I, D, D, D, I, stream));                 // silu(gate) * up
    }
    TRY(mb200_skinny_gemm_bf16(act, w[6], x, nullptr, x, B, D, I, I, I, D, D, stream));            // x += down(act)
  }
  TRY(mb200_rmsnorm_fwd(x, fnorm, xn, nullptr, B, D, eps, dt, stream));
  TRY(mb200_skinny_gemm_bf16(xn, lm_head, logits, nullptr, nullptr, B, V, D, D, D, ld_logits, 0, stream));
  if (next_ids) TRY(mb200_argmax_bf16(logits, ld_logits, B, V, next_ids, stream));
  return MB200_OK;
}

}  // extern "C"

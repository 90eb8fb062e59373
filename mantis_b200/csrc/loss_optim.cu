// Shifted, masked cross-entropy over a logits chunk (fused forward + dlogits), and fused AdamW.
//   CE: reference mantis/models/mllava/modeling_llava.py:523-537 (CrossEntropyLoss, ignore_index=-100, mean)
//       and mantis/models/idefics2/modeling_idefics2.py:1883-1899.
//   AdamW: HF Trainer default optimiser (mantis/train/scripts/train_mllava.sh:162-165) == torch.optim.AdamW.
#include "common.cuh"

namespace {
using mb::Cvt;

// One CTA per row.  logits [n, ld] (V valid columns).  labels[r] < 0 => ignored row (dlogits = 0).
// loss_rows[r] = lse - logit[target] (fp32).  If dlogits != nullptr: dlogits = (softmax - onehot) * gscale
// written in T (may alias logits).
template <typename T>
__global__ void __launch_bounds__(512)
ce_fwd_bwd_kernel(const T* logits, const int64_t* __restrict__ labels, float* __restrict__ loss_rows,
                  float* __restrict__ lse_rows, T* dlogits /* may alias logits */, long long n, int V, long long ld,
                  const float* __restrict__ gscale_ptr, float gscale_const) {
  __shared__ float red[33];
  const long long r = blockIdx.x;
  const T* lr = logits + (size_t)r * ld;
  const long long y = labels[r];
  const bool ignored = (y < 0 || y >= V);
  if (ignored && !lse_rows) {
    if (loss_rows && threadIdx.x == 0) loss_rows[r] = 0.f;
    if (dlogits) { T* dr = dlogits + (size_t)r * ld; for (int i = threadIdx.x; i < V; i += blockDim.x) dr[i] = Cvt<T>::from_f(0.f); }
    return;
  }
  // every thread reads the target logit BEFORE the block reductions: dlogits may alias logits, and the first warps to leave the
  // last reduction start overwriting the row while thread 0 is still composing the loss
  const float tgt = ignored ? 0.f : Cvt<T>::to_f(lr[y]);
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, Cvt<T>::to_f(lr[i]));
  mx = mb::block_max(mx, red);
  float se = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) se += __expf(Cvt<T>::to_f(lr[i]) - mx);
  se = mb::block_sum(se, red);
  const float lse = mx + logf(se);
  if (threadIdx.x == 0) {
    if (lse_rows) lse_rows[r] = lse;
    if (loss_rows) loss_rows[r] = ignored ? 0.f : (lse - tgt);
  }
  if (dlogits) {
    const float gs = ignored ? 0.f : (gscale_ptr ? *gscale_ptr : gscale_const);
    T* dr = dlogits + (size_t)r * ld;
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
      float p = __expf(Cvt<T>::to_f(lr[i]) - lse);
      if (i == y) p -= 1.f;
      dr[i] = Cvt<T>::from_f(p * gs);
    }
  }
}

// bf16 fast path: 16-byte vector loads, online (max, sum-exp) in ONE read pass, then one read + one write pass for dlogits
// (algorithmic minimum for forward+backward is 1 read + 1 write; the second read mostly hits L2: a row is 256 KB).
__global__ void __launch_bounds__(512)
ce_fwd_bwd_vec_kernel(const bf16* logits, const int64_t* __restrict__ labels, float* __restrict__ loss_rows,
                      float* __restrict__ lse_rows, bf16* dlogits /* may alias logits */, long long n, int V, long long ld,
                      const float* __restrict__ gscale_ptr, float gscale_const) {
  __shared__ float red[33];
  const long long r = blockIdx.x;
  const bf16* lr = logits + (size_t)r * ld;
  const long long y = labels[r];
  const bool ignored = (y < 0 || y >= V);
  const int nv = V >> 3;
  if (ignored && !lse_rows) {
    if (loss_rows && threadIdx.x == 0) loss_rows[r] = 0.f;
    if (dlogits) {
      bf16* dr = dlogits + (size_t)r * ld;
      const int4 z = make_int4(0, 0, 0, 0);
      for (int i = threadIdx.x; i < nv; i += blockDim.x) reinterpret_cast<int4*>(dr)[i] = z;
      for (int i = nv * 8 + threadIdx.x; i < V; i += blockDim.x) dr[i] = __float2bfloat16_rn(0.f);
    }
    return;
  }
  const float tgt = ignored ? 0.f : __bfloat162float(lr[y]);      // read before anything is overwritten (see above)
  float m = -INFINITY, sacc = 0.f;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    float f[8];
    mb::Vec8<bf16>::load(lr + (size_t)i * 8, f);
    float vm = f[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) vm = fmaxf(vm, f[j]);
    const float mn = fmaxf(m, vm);
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += __expf(f[j] - mn);
    sacc = sacc * __expf(m - mn) + t;      // exp(-inf) = 0 on the first vector
    m = mn;
  }
  for (int i = nv * 8 + threadIdx.x; i < V; i += blockDim.x) {
    const float v = __bfloat162float(lr[i]);
    const float mn = fmaxf(m, v);
    sacc = sacc * __expf(m - mn) + __expf(v - mn);
    m = mn;
  }
  const float M = mb::block_max(m, red);
  const float se = mb::block_sum((m == -INFINITY) ? 0.f : sacc * __expf(m - M), red);
  const float lse = M + logf(se);
  if (threadIdx.x == 0) {
    if (lse_rows) lse_rows[r] = lse;
    if (loss_rows) loss_rows[r] = ignored ? 0.f : (lse - tgt);
  }
  if (dlogits) {
    const float gs = ignored ? 0.f : (gscale_ptr ? *gscale_ptr : gscale_const);
    bf16* dr = dlogits + (size_t)r * ld;
    const int yv = ignored ? -1 : (int)(y >> 3);
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      float f[8];
      mb::Vec8<bf16>::load(lr + (size_t)i * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __expf(f[j] - lse);
      if (i == yv) f[y & 7] -= 1.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= gs;
      mb::Vec8<bf16>::store(dr + (size_t)i * 8, f);
    }
    for (int i = nv * 8 + threadIdx.x; i < V; i += blockDim.x) {
      float pv = __expf(__bfloat162float(lr[i]) - lse);
      if (i == y) pv -= 1.f;
      dr[i] = __float2bfloat16_rn(pv * gs);
    }
  }
}

// sum of loss_rows and count of valid labels -> out[0] = sum, out[1] = count  (single CTA, deterministic)
__global__ void __launch_bounds__(1024)
ce_reduce_kernel(const float* __restrict__ loss_rows, const int64_t* __restrict__ labels, long long n, int V,
                 float* __restrict__ out, int accumulate) {
  __shared__ float red[33];
  float s = 0.f, c = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const long long y = labels[i];
    if (y >= 0 && y < V) { s += loss_rows[i]; c += 1.f; }
  }
  s = mb::block_sum(s, red); c = mb::block_sum(c, red);
  if (threadIdx.x == 0) {
    if (accumulate) { out[0] += s; out[1] += c; } else { out[0] = s; out[1] = c; }
  }
}

// eff_labels[b, s] = (mask[b, s+1] != 0) ? labels[b, s+1] : -100   for s < S-1; last position -> -100
// (the reference drops rows with shift_attention_mask == 0 *before* CE; CE then ignores label -100)
__global__ void __launch_bounds__(256)
shift_labels_kernel(const int64_t* __restrict__ labels, const int64_t* __restrict__ mask /*nullable*/,
                    int64_t* __restrict__ out, int B, int S, int64_t ignore_index, float* __restrict__ count_out) {
  __shared__ float red[33];
  const long long total = (long long)B * S;
  float c = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(i % S);
    int64_t v = -100;
    if (s + 1 < S) {
      v = labels[i + 1];
      if (v == ignore_index) v = -100;
      if (mask && mask[i + 1] == 0) v = -100;
    }
    out[i] = v;
    c += (v >= 0) ? 1.f : 0.f;
  }
  if (count_out) {
    c = mb::block_sum(c, red);
    if (threadIdx.x == 0 && c != 0.f) atomicAdd(count_out, c);   // integer-valued fp32 adds: exact, order independent (< 2^24)
  }
}

// sum of squares (for grad-norm clipping), fp32 partial per CTA then atomic add
template <typename T>
__global__ void __launch_bounds__(256)
sumsq_kernel(const T* __restrict__ g, long long n, float* __restrict__ out) {
  __shared__ float red[33];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = Cvt<T>::to_f(g[i]); s += v * v;
  }
  s = mb::block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}


// ------------------------------------------------------------------ row compaction for the LM head
// idx_out[j] = i for the j-th row with labels[i] >= 0 (ascending order), *count_out = number of such rows.  The reference gathers
// the rows whose shifted label is not ignored AFTER computing every row's logits (modeling_llava.py:526-531); here they are
// compacted before the LM-head GEMMs.  One CTA (n is a sequence length: <= a few 10^4), ballot + warp-sum scan per 1024 rows;
// the output length is known on the host from the collator's label count, so nothing is read back.
__global__ void __launch_bounds__(1024)
compact_valid_rows_kernel(const int64_t* __restrict__ labels, long long n, int64_t* __restrict__ idx_out, long long cap,
                          int* __restrict__ count_out) {
  __shared__ int warp_sums[32];
  __shared__ int base_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (long long start = 0; start < n; start += 1024) {
    const long long i = start + tid;
    const bool flag = (i < n) && (labels[i] >= 0);
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int prefix = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_sums[wid] = __popc(bal);
    __syncthreads();
    int v = warp_sums[lane];                       // every warp scans the 32 warp totals itself (no second barrier needed)
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const int excl_w = __shfl_sync(0xffffffffu, incl - v, wid);
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const long long pos = (long long)base_s + excl_w + prefix;
    if (flag && pos < cap) idx_out[pos] = i;
    __syncthreads();
    if (tid == 0) base_s += total;
    __syncthreads();
  }
  if (tid == 0) *count_out = base_s;
}

// ------------------------------------------------------------------ flat AdamW with fp32 master weights
// The reference recipe (mantis/train/scripts/train_mllava.sh:148,162 `--bf16 True --learning_rate 1e-5`,
// zero_configs/zero3.json "bf16": enabled) keeps an fp32 master copy of every weight inside DeepSpeed's optimizer; the
// bf16 weights are its rounding.  Updating bf16 weights in place instead rounds an lr = 1e-5 step to nothing (ulp(0.02) in
// bf16 = 1.2e-4).  Here the master is stored SPLIT: the bf16 weight the model computes with + the 16 low bits of the fp32
// word, master_bits = (bf16_bits << 16) + (int16) lo, with the bf16 half rounded to nearest (ties away from zero) so that
// it is at the same time the model's weight -- an exact fp32 master for 2 extra bytes per parameter (16 GB, not 32 GB, for
// Mantis-8B).  One launch covers the whole flat parameter buffer; param groups (weight decay on / off) are resolved per
// 1024-element block (every tensor's slice of the flat buffers starts on a 1024 boundary); the gradient-norm clip factor is
// computed on the device from the squared norm, and the gradient buffer is zeroed on the way out.
__device__ __forceinline__ float clip_factor(const float* norm_sq, float max_norm, float grad_scale) {
  if (!norm_sq || max_norm <= 0.f) return 1.f;
  const float total = sqrtf(*norm_sq) * grad_scale;
  const float c = max_norm / (total + 1e-6f);              // torch.nn.utils.clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max = 1)
  return (c < 1.f || !(c == c)) ? c : 1.f;                 // NaN norm propagates like torch
}
__device__ __forceinline__ float master_join(uint16_t hi, uint16_t lo) {
  return __uint_as_float(((uint32_t)hi << 16) + (uint32_t)(int32_t)(int16_t)lo);
}
__device__ __forceinline__ void master_split(float x, uint16_t& hi, uint16_t& lo) {
  const uint32_t b = __float_as_uint(x);
  if ((b & 0x7f800000u) == 0x7f800000u) { hi = (uint16_t)(b >> 16) | ((b & 0xffffu) ? 0x40u : 0u); lo = 0; return; }   // inf / nan
  hi = (uint16_t)((b + 0x8000u) >> 16);
  lo = (uint16_t)(b - ((uint32_t)hi << 16));
}

struct AdamHyper { float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, grad_scale, max_norm; };

template <bool SPLIT, typename GT>
__global__ void __launch_bounds__(256)
adamw_flat_kernel(void* __restrict__ p_, uint16_t* __restrict__ lo_, GT* __restrict__ g_, float* __restrict__ m_,
                  float* __restrict__ v_, const unsigned char* __restrict__ blk_group, long long n8, AdamHyper h,
                  const float* __restrict__ norm_sq, int zero_grad) {
  const float gs = h.grad_scale * clip_factor(norm_sq, h.max_norm, h.grad_scale);
  const float step_size = h.lr / h.bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const long long e0 = i * 8;
    const float wd = (blk_group && blk_group[e0 >> 10]) ? 0.f : h.wd;
    float p[8], g[8], m[8], v[8];
    uint16_t hi[8], lo[8];
    if (SPLIT) {
      const int4 ph = mb::ld_stream(reinterpret_cast<const int4*>(reinterpret_cast<uint16_t*>(p_) + e0));
      const int4 pl = mb::ld_stream(reinterpret_cast<const int4*>(lo_ + e0));
      const uint16_t* phs = reinterpret_cast<const uint16_t*>(&ph); const uint16_t* pls = reinterpret_cast<const uint16_t*>(&pl);
#pragma unroll
      for (int j = 0; j < 8; ++j) p[j] = master_join(phs[j], pls[j]);
    } else {
      mb::Vec8<float>::load(reinterpret_cast<const float*>(p_) + e0, p);
    }
    mb::Vec8<GT>::load(g_ + e0, g);
    mb::Vec8<float>::load(m_ + e0, m);
    mb::Vec8<float>::load(v_ + e0, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gv = g[j] * gs;
      float pv = p[j] * (1.f - h.lr * wd);
      m[j] = h.beta1 * m[j] + (1.f - h.beta1) * gv;
      v[j] = h.beta2 * v[j] + (1.f - h.beta2) * gv * gv;
      const float denom = sqrtf(v[j]) / h.bc2_sqrt + h.eps;
      pv -= step_size * (m[j] / denom);
      p[j] = pv;
    }
    mb::Vec8<float>::store(m_ + e0, m);
    mb::Vec8<float>::store(v_ + e0, v);
    if (SPLIT) {
#pragma unroll
      for (int j = 0; j < 8; ++j) master_split(p[j], hi[j], lo[j]);
      mb::st_stream(reinterpret_cast<int4*>(reinterpret_cast<uint16_t*>(p_) + e0), *reinterpret_cast<const int4*>(hi));
      mb::st_stream(reinterpret_cast<int4*>(lo_ + e0), *reinterpret_cast<const int4*>(lo));
    } else {
      mb::Vec8<float>::store(reinterpret_cast<float*>(p_) + e0, p);
    }
    if (zero_grad) {
      float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      mb::Vec8<GT>::store(g_ + e0, z);
    }
  }
}

// master <-> (bf16, lo) conversion of a flat range (trainer construction / optimizer checkpoints)
__global__ void __launch_bounds__(256)
master_split_kernel(const float* __restrict__ master, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    master_split(master[i], hi[i], lo[i]);
}
__global__ void __launch_bounds__(256)
master_join_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, float* __restrict__ master, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    master[i] = master_join(hi[i], lo ? lo[i] : (uint16_t)0);
}

// dst (fp32) += src (bf16 / fp32): gradients that autograd produced in the parameter dtype folded into the fp32 main gradient
template <typename T>
__global__ void __launch_bounds__(256)
accum_f32_kernel(float* __restrict__ dst, const T* __restrict__ src, long long n, float scale, int accumulate) {
  const long long n8 = n >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float a[8], b[8];
    if (accumulate) mb::Vec8<float>::load(dst + i * 8, a);
    mb::Vec8<T>::load(src + i * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (accumulate ? a[j] : 0.f) + b[j] * scale;
    mb::Vec8<float>::store(dst + i * 8, a);
  }
  for (long long i = n8 * 8 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = (accumulate ? dst[i] : 0.f) + Cvt<T>::to_f(src[i]) * scale;
}

// vectorised sum of squares (fp32 / bf16), fp32 partial per CTA then one atomic
template <typename T>
__global__ void __launch_bounds__(256)
sumsq_vec_kernel(const T* __restrict__ g, long long n8, float* __restrict__ out) {
  __shared__ float red[33];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    mb::Vec8<T>::load(g + i * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j] * f[j];
  }
  s = mb::block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

}  // namespace

#define DISPATCH_T(dtype, ...)                                             \
  if ((dtype) == MB200_DTYPE_BF16) { typedef bf16 T; __VA_ARGS__; }        \
  else if ((dtype) == MB200_DTYPE_F32) { typedef float T; __VA_ARGS__; }   \
  else return -EINVAL;

extern "C" {

int mb200_ce_fwd_bwd(const void* logits, const int64_t* labels, float* loss_rows, float* lse_rows, void* dlogits,
                     long long n, int V, long long ld, const float* gscale_ptr, float gscale_const, int dtype,
                     void* stream) {
  if (n <= 0) return MB200_OK;
  if (n > 2147483647LL) return -EINVAL;
  if (dtype == MB200_DTYPE_BF16 && (ld & 7) == 0 && !(reinterpret_cast<uintptr_t>(logits) & 15) &&
      (!dlogits || !(reinterpret_cast<uintptr_t>(dlogits) & 15))) {
    ce_fwd_bwd_vec_kernel<<<(unsigned)n, 512, 0, (cudaStream_t)stream>>>((const bf16*)logits, labels, loss_rows, lse_rows,
                                                                        (bf16*)dlogits, n, V, ld, gscale_ptr, gscale_const);
    MB200_CHECK_LAUNCH(); return MB200_OK;
  }
  DISPATCH_T(dtype, (ce_fwd_bwd_kernel<T><<<(unsigned)n, 512, 0, (cudaStream_t)stream>>>(
                        (const T*)logits, labels, loss_rows, lse_rows, (T*)dlogits, n, V, ld, gscale_ptr, gscale_const)));
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_ce_reduce(const float* loss_rows, const int64_t* labels, long long n, int V, float* out2, int accumulate,
                    void* stream) {
  ce_reduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(loss_rows, labels, n, V, out2, accumulate);
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_shift_labels(const int64_t* labels, const int64_t* mask, int64_t* out, int B, int S, int64_t ignore_index,
                       float* count_out, void* stream) {
  if (B <= 0 || S <= 0) return MB200_OK;
  long long total = (long long)B * S;
  int grid = (int)((total + 255) / 256); if (grid > mb::num_sms() * 4) grid = mb::num_sms() * 4;
  shift_labels_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(labels, mask, out, B, S, ignore_index, count_out);
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_compact_valid_rows(const int64_t* labels, long long n, int64_t* idx_out, long long cap, int* count_out, void* stream) {
  if (n < 0 || cap < 0) return -EINVAL;
  compact_valid_rows_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(labels, n, idx_out, cap, count_out);
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_adamw_flat(void* p, void* lo, void* g, float* m, float* v, const unsigned char* blk_group, long long n, float lr,
                     float beta1, float beta2, float eps, float wd, int step, float grad_scale, const float* norm_sq,
                     float max_norm, int zero_grad, int p_dtype, int g_dtype, void* stream) {
  if (n <= 0) return MB200_OK;
  if ((n & 7) || (reinterpret_cast<uintptr_t>(p) & 15) || (reinterpret_cast<uintptr_t>(g) & 15) ||
      (reinterpret_cast<uintptr_t>(m) & 15) || (reinterpret_cast<uintptr_t>(v) & 15) || (reinterpret_cast<uintptr_t>(lo) & 15))
    return -EINVAL;
  if (p_dtype == MB200_DTYPE_BF16 && !lo) return -EINVAL;            // bf16 weights need their low halves (no silent RTN path)
  AdamHyper h;
  h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.wd = wd; h.grad_scale = grad_scale; h.max_norm = max_norm;
  h.bc1 = 1.f - powf(beta1, (float)step);
  h.bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  const long long n8 = n >> 3;
  long long g0 = (n8 + 255) / 256; const long long cap = (long long)mb::num_sms() * 8; if (g0 > cap) g0 = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (p_dtype == MB200_DTYPE_BF16 && g_dtype == MB200_DTYPE_F32)
    adamw_flat_kernel<true, float><<<(int)g0, 256, 0, st>>>(p, (uint16_t*)lo, (float*)g, m, v, blk_group, n8, h, norm_sq, zero_grad);
  else if (p_dtype == MB200_DTYPE_BF16 && g_dtype == MB200_DTYPE_BF16)
    adamw_flat_kernel<true, bf16><<<(int)g0, 256, 0, st>>>(p, (uint16_t*)lo, (bf16*)g, m, v, blk_group, n8, h, norm_sq, zero_grad);
  else if (p_dtype == MB200_DTYPE_F32 && g_dtype == MB200_DTYPE_F32)
    adamw_flat_kernel<false, float><<<(int)g0, 256, 0, st>>>(p, nullptr, (float*)g, m, v, blk_group, n8, h, norm_sq, zero_grad);
  else return -EINVAL;
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_master_split(const float* master, void* hi_bf16, void* lo_u16, long long n, void* stream) {
  if (n <= 0) return MB200_OK;
  long long g0 = (n + 255) / 256; const long long cap = (long long)mb::num_sms() * 16; if (g0 > cap) g0 = cap;
  master_split_kernel<<<(int)g0, 256, 0, (cudaStream_t)stream>>>(master, (uint16_t*)hi_bf16, (uint16_t*)lo_u16, n);
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_master_join(const void* hi_bf16, const void* lo_u16, float* master, long long n, void* stream) {
  if (n <= 0) return MB200_OK;
  long long g0 = (n + 255) / 256; const long long cap = (long long)mb::num_sms() * 16; if (g0 > cap) g0 = cap;
  master_join_kernel<<<(int)g0, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)hi_bf16, (const uint16_t*)lo_u16, master, n);
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_accum_f32(float* dst, const void* src, long long n, float scale, int accumulate, int src_dtype, void* stream) {
  if (n <= 0) return MB200_OK;
  if ((reinterpret_cast<uintptr_t>(dst) & 31) || (reinterpret_cast<uintptr_t>(src) & 15)) return -EINVAL;
  long long g0 = ((n >> 3) + 255) / 256; if (g0 < 1) g0 = 1;
  const long long cap = (long long)mb::num_sms() * 8; if (g0 > cap) g0 = cap;
  typedef float F;
  if (src_dtype == MB200_DTYPE_BF16) accum_f32_kernel<bf16><<<(int)g0, 256, 0, (cudaStream_t)stream>>>(dst, (const bf16*)src, n, scale, accumulate);
  else if (src_dtype == MB200_DTYPE_F32) accum_f32_kernel<F><<<(int)g0, 256, 0, (cudaStream_t)stream>>>(dst, (const F*)src, n, scale, accumulate);
  else return -EINVAL;
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_sumsq(const void* g, long long n, float* out, int dtype, void* stream) {
  if (n <= 0) return MB200_OK;
  if (!(n & 7) && !(reinterpret_cast<uintptr_t>(g) & 31)) {
    long long b0 = ((n >> 3) + 255) / 256; const long long bcap = (long long)mb::num_sms() * 8; if (b0 > bcap) b0 = bcap;
    DISPATCH_T(dtype, (sumsq_vec_kernel<T><<<(int)b0, 256, 0, (cudaStream_t)stream>>>((const T*)g, n >> 3, out)));
    MB200_CHECK_LAUNCH(); return MB200_OK;
  }
  long long g0 = (n + 255) / 256; long long cap = (long long)mb::num_sms() * 8; if (g0 > cap) g0 = cap;
  DISPATCH_T(dtype, (sumsq_kernel<T><<<(int)g0, 256, 0, (cudaStream_t)stream>>>((const T*)g, n, out)));
  MB200_CHECK_LAUNCH(); return MB200_OK;
}

}  // extern "C"

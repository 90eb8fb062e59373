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
ce_fwd_bwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels, float* __restrict__ loss_rows,
                  float* __restrict__ lse_rows, T* __restrict__ dlogits, long long n, int V, long long ld,
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
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, Cvt<T>::to_f(lr[i]));
  mx = mb::block_max(mx, red);
  float se = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) se += __expf(Cvt<T>::to_f(lr[i]) - mx);
  se = mb::block_sum(se, red);
  const float lse = mx + logf(se);
  if (threadIdx.x == 0) {
    if (lse_rows) lse_rows[r] = lse;
    if (loss_rows) loss_rows[r] = ignored ? 0.f : (lse - Cvt<T>::to_f(lr[y]));
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
ce_fwd_bwd_vec_kernel(const bf16* __restrict__ logits, const int64_t* __restrict__ labels, float* __restrict__ loss_rows,
                      float* __restrict__ lse_rows, bf16* __restrict__ dlogits, long long n, int V, long long ld,
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
    if (loss_rows) loss_rows[r] = ignored ? 0.f : (lse - __bfloat162float(lr[y]));
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

// ------------------------------------------------------------------ AdamW
template <typename T>
__global__ void __launch_bounds__(256)
adamw_kernel(T* __restrict__ p, const T* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             long long n, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt,
             float grad_scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float pv = Cvt<T>::to_f(p[i]);
    const float gv = Cvt<T>::to_f(g[i]) * grad_scale;
    pv *= (1.f - lr * wd);
    const float mv = beta1 * m[i] + (1.f - beta1) * gv;
    const float vv = beta2 * v[i] + (1.f - beta2) * gv * gv;
    m[i] = mv; v[i] = vv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv -= (lr / bc1) * (mv / denom);
    p[i] = Cvt<T>::from_f(pv);
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
int mb200_adamw(void* p, const void* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                float eps, float wd, int step, float grad_scale, int dtype, void* stream) {
  if (n <= 0) return MB200_OK;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  long long g0 = (n + 255) / 256; long long cap = (long long)mb::num_sms() * 16; if (g0 > cap) g0 = cap;
  DISPATCH_T(dtype, (adamw_kernel<T><<<(int)g0, 256, 0, (cudaStream_t)stream>>>(
                        (T*)p, (const T*)g, m, v, n, lr, beta1, beta2, eps, wd, bc1, bc2s, grad_scale)));
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_sumsq(const void* g, long long n, float* out, int dtype, void* stream) {
  if (n <= 0) return MB200_OK;
  long long g0 = (n + 255) / 256; long long cap = (long long)mb::num_sms() * 8; if (g0 > cap) g0 = cap;
  DISPATCH_T(dtype, (sumsq_kernel<T><<<(int)g0, 256, 0, (cudaStream_t)stream>>>((const T*)g, n, out)));
  MB200_CHECK_LAUNCH(); return MB200_OK;
}

}  // extern "C"

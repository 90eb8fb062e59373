// Shape-agnostic flash-style attention (SIMT, fp32 math): forward, dQ and dK/dV kernels.
// Any head_dim <= 256, GQA (H % Hkv == 0), optional causal mask with KV-cache offset, optional
// per-batch key padding mask.  Used for tiny / odd shapes (tiny parity configs have head_dim 16),
// fp32 parity, decode with a KV cache, and as the on-GPU cross-check of the tensor-core kernels.
// Semantics follow the reference stack's eager attention (transformers llama/modeling_llama.py:199-221,
// siglip/modeling_siglip.py:229-249): softmax(q k^T * scale + mask) in fp32, then @ v.
// Rows whose keys are all masked produce 0 (the reference produces an unspecified uniform average
// there; such rows are padding and never reach the loss).
//
// Layout: q/o/dq/do [B, Sq, H, hd], k/v/dk/dv [B, Sk, Hkv, hd] with explicit element strides for
// (batch, seq, head); hd contiguous.  lse/delta: [B, H, Sq] fp32.
#include "common.cuh"

namespace {
using mb::Cvt;
constexpr int KT = 32;        // keys (or queries) per shared-memory tile == warp width
constexpr int WARPS = 8;      // rows handled per CTA
constexpr int MAXR = 8;       // head_dim <= 32 * MAXR

struct AttnP {
  int B, H, Hkv, Sq, Sk, hd;
  long long q_sb, q_ss, q_sh, k_sb, k_ss, k_sh, v_sb, v_ss, v_sh, o_sb, o_ss, o_sh;
  float scale; int causal;    // causal: key j visible to query i iff j <= i + (Sk - Sq)
  int window;                 // > 0: additionally (i + (Sk - Sq)) - j < window  (Mistral sliding-window attention,
                              //      transformers mistral/modeling_mistral.py sliding_window_overlay: kv_idx > q_idx - window)
  const int64_t* kmask;       // [B, Sk] or null
  long long kmask_sb;
};

template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
attn_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ o,
                float* __restrict__ lse, AttnP p) {
  extern __shared__ float sm[];
  const int hd = p.hd, hdp = hd + 1;
  float* Ks = sm;                       // [KT][hdp]
  float* Vs = Ks + KT * hdp;            // [KT][hdp]
  float* Qs = Vs + KT * hdp;            // [WARPS][hd]
  __shared__ int kvalid[KT];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.z, h = blockIdx.y, hk = h / (p.H / p.Hkv);
  const int qi = blockIdx.x * WARPS + w;
  const bool row_ok = qi < p.Sq;
  const int off = p.Sk - p.Sq;
  if (row_ok) {
    const T* qp = q + (size_t)b * p.q_sb + (size_t)qi * p.q_ss + (size_t)h * p.q_sh;
    for (int d = lane; d < hd; d += 32) Qs[w * hd + d] = Cvt<T>::to_f(qp[d]) * p.scale;
  }
  float m = -INFINITY, l = 0.f, acc[MAXR];
#pragma unroll
  for (int i = 0; i < MAXR; ++i) acc[i] = 0.f;
  // last key any row of this CTA can see
  int k_end = p.Sk;
  if (p.causal) { int last_q = min(p.Sq - 1, blockIdx.x * WARPS + WARPS - 1); k_end = min(p.Sk, last_q + off + 1); }
  for (int k0 = 0; k0 < k_end; k0 += KT) {
    __syncthreads();
    for (int e = threadIdx.x; e < KT * hd; e += blockDim.x) {
      const int j = e / hd, d = e % hd, kj = k0 + j;
      float kv = 0.f, vv = 0.f;
      if (kj < p.Sk) {
        kv = Cvt<T>::to_f(k[(size_t)b * p.k_sb + (size_t)kj * p.k_ss + (size_t)hk * p.k_sh + d]);
        vv = Cvt<T>::to_f(v[(size_t)b * p.v_sb + (size_t)kj * p.v_ss + (size_t)hk * p.v_sh + d]);
      }
      Ks[j * hdp + d] = kv; Vs[j * hdp + d] = vv;
    }
    if (threadIdx.x < KT) {
      const int kj = k0 + threadIdx.x;
      kvalid[threadIdx.x] = (kj < p.Sk) && (!p.kmask || p.kmask[(size_t)b * p.kmask_sb + kj] != 0);
    }
    __syncthreads();
    if (!row_ok) continue;
    const int kj = k0 + lane;
    bool vis = kvalid[lane] && (!p.causal || (kj <= qi + off && (p.window <= 0 || qi + off - kj < p.window)));
    float s = -INFINITY;
    if (vis) {
      s = 0.f;
      const float* kr = Ks + lane * hdp; const float* qr = Qs + w * hd;
      for (int d = 0; d < hd; ++d) s = fmaf(qr[d], kr[d], s);
    }
    const float mt = mb::warp_max(s);
    if (mt == -INFINITY) continue;            // whole tile masked for this row
    const float m_new = fmaxf(m, mt);
    const float corr = (m == -INFINITY) ? 0.f : __expf(m - m_new);
    const float pj = vis ? __expf(s - m_new) : 0.f;
    l = l * corr + mb::warp_sum(pj);
    m = m_new;
#pragma unroll
    for (int i = 0; i < MAXR; ++i) acc[i] *= corr;
    for (int j = 0; j < KT; ++j) {
      const float pb = __shfl_sync(0xffffffffu, pj, j);
      if (pb != 0.f) {
#pragma unroll
        for (int i = 0; i < MAXR; ++i) { const int d = lane + 32 * i; if (d < hd) acc[i] = fmaf(pb, Vs[j * hdp + d], acc[i]); }
      }
    }
  }
  if (!row_ok) return;
  const float inv = (l > 0.f) ? 1.f / l : 0.f;
  T* op = o + (size_t)b * p.o_sb + (size_t)qi * p.o_ss + (size_t)h * p.o_sh;
#pragma unroll
  for (int i = 0; i < MAXR; ++i) { const int d = lane + 32 * i; if (d < hd) op[d] = Cvt<T>::from_f(acc[i] * inv); }
  if (lse && lane == 0) lse[((size_t)b * p.H + h) * p.Sq + qi] = (l > 0.f) ? (m + logf(l)) : -INFINITY;
}

// dQ (and delta = rowsum(dO * O)) ; one warp per query row
template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
attn_bwd_dq_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ o,
                   const T* __restrict__ dout, const float* __restrict__ lse, float* __restrict__ delta,
                   T* __restrict__ dq, AttnP p) {
  extern __shared__ float sm[];
  const int hd = p.hd, hdp = hd + 1;
  float* Ks = sm; float* Vs = Ks + KT * hdp; float* Qs = Vs + KT * hdp; float* Ds = Qs + WARPS * hd;
  __shared__ int kvalid[KT];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.z, h = blockIdx.y, hk = h / (p.H / p.Hkv);
  const int qi = blockIdx.x * WARPS + w;
  const bool row_ok = qi < p.Sq;
  const int off = p.Sk - p.Sq;
  float L = 0.f, dl = 0.f;
  if (row_ok) {
    const size_t qo = (size_t)b * p.q_sb + (size_t)qi * p.q_ss + (size_t)h * p.q_sh;
    const size_t oo = (size_t)b * p.o_sb + (size_t)qi * p.o_ss + (size_t)h * p.o_sh;
    float part = 0.f;
    for (int d = lane; d < hd; d += 32) {
      Qs[w * hd + d] = Cvt<T>::to_f(q[qo + d]) * p.scale;
      const float g = Cvt<T>::to_f(dout[oo + d]);
      Ds[w * hd + d] = g;
      part += g * Cvt<T>::to_f(o[oo + d]);
    }
    dl = mb::warp_sum(part);
    L = lse[((size_t)b * p.H + h) * p.Sq + qi];
    if (delta && lane == 0) delta[((size_t)b * p.H + h) * p.Sq + qi] = dl;
  }
  float acc[MAXR];
#pragma unroll
  for (int i = 0; i < MAXR; ++i) acc[i] = 0.f;
  int k_end = p.Sk;
  if (p.causal) { int last_q = min(p.Sq - 1, blockIdx.x * WARPS + WARPS - 1); k_end = min(p.Sk, last_q + off + 1); }
  for (int k0 = 0; k0 < k_end; k0 += KT) {
    __syncthreads();
    for (int e = threadIdx.x; e < KT * hd; e += blockDim.x) {
      const int j = e / hd, d = e % hd, kj = k0 + j;
      float kv = 0.f, vv = 0.f;
      if (kj < p.Sk) {
        kv = Cvt<T>::to_f(k[(size_t)b * p.k_sb + (size_t)kj * p.k_ss + (size_t)hk * p.k_sh + d]);
        vv = Cvt<T>::to_f(v[(size_t)b * p.v_sb + (size_t)kj * p.v_ss + (size_t)hk * p.v_sh + d]);
      }
      Ks[j * hdp + d] = kv; Vs[j * hdp + d] = vv;
    }
    if (threadIdx.x < KT) {
      const int kj = k0 + threadIdx.x;
      kvalid[threadIdx.x] = (kj < p.Sk) && (!p.kmask || p.kmask[(size_t)b * p.kmask_sb + kj] != 0);
    }
    __syncthreads();
    if (!row_ok || L == -INFINITY) continue;
    const int kj = k0 + lane;
    const bool vis = kvalid[lane] && (!p.causal || (kj <= qi + off && (p.window <= 0 || qi + off - kj < p.window)));
    float ds = 0.f;
    if (vis) {
      float s = 0.f, dp = 0.f;
      const float* kr = Ks + lane * hdp; const float* vr = Vs + lane * hdp;
      const float* qr = Qs + w * hd; const float* gr = Ds + w * hd;
      for (int d = 0; d < hd; ++d) { s = fmaf(qr[d], kr[d], s); dp = fmaf(gr[d], vr[d], dp); }
      ds = __expf(s - L) * (dp - dl) * p.scale;
    }
    for (int j = 0; j < KT; ++j) {
      const float db = __shfl_sync(0xffffffffu, ds, j);
      if (db != 0.f) {
#pragma unroll
        for (int i = 0; i < MAXR; ++i) { const int d = lane + 32 * i; if (d < hd) acc[i] = fmaf(db, Ks[j * hdp + d], acc[i]); }
      }
    }
  }
  if (!row_ok) return;
  T* dp_ = dq + (size_t)b * p.q_sb + (size_t)qi * p.q_ss + (size_t)h * p.q_sh;
#pragma unroll
  for (int i = 0; i < MAXR; ++i) { const int d = lane + 32 * i; if (d < hd) dp_[d] = Cvt<T>::from_f(acc[i]); }
}

// dK, dV ; one warp per key row, loops over the query heads of its group and over query tiles
template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
attn_bwd_dkv_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                    const T* __restrict__ dout, const float* __restrict__ lse, const float* __restrict__ delta,
                    T* __restrict__ dk, T* __restrict__ dv, AttnP p) {
  extern __shared__ float sm[];
  const int hd = p.hd, hdp = hd + 1;
  float* Qt = sm; float* Gt = Qt + KT * hdp; float* Kw = Gt + KT * hdp; float* Vw = Kw + WARPS * hd;
  __shared__ float Ls[KT], Dl[KT];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.z, hk = blockIdx.y, G = p.H / p.Hkv;
  const int kj = blockIdx.x * WARPS + w;
  const bool row_ok = kj < p.Sk;
  const int off = p.Sk - p.Sq;
  bool kvis = row_ok && (!p.kmask || p.kmask[(size_t)b * p.kmask_sb + kj] != 0);
  if (row_ok) {
    for (int d = lane; d < hd; d += 32) {
      Kw[w * hd + d] = Cvt<T>::to_f(k[(size_t)b * p.k_sb + (size_t)kj * p.k_ss + (size_t)hk * p.k_sh + d]);
      Vw[w * hd + d] = Cvt<T>::to_f(v[(size_t)b * p.v_sb + (size_t)kj * p.v_ss + (size_t)hk * p.v_sh + d]);
    }
  }
  float dka[MAXR], dva[MAXR];
#pragma unroll
  for (int i = 0; i < MAXR; ++i) { dka[i] = 0.f; dva[i] = 0.f; }
  // first query that can see any key of this CTA
  int q_begin = 0;
  if (p.causal) { q_begin = blockIdx.x * WARPS - off; if (q_begin < 0) q_begin = 0; q_begin = (q_begin / KT) * KT; }
  for (int g = 0; g < G; ++g) {
    const int h = hk * G + g;
    for (int q0 = q_begin; q0 < p.Sq; q0 += KT) {
      __syncthreads();
      for (int e = threadIdx.x; e < KT * hd; e += blockDim.x) {
        const int i = e / hd, d = e % hd, qi = q0 + i;
        float qv = 0.f, gv = 0.f;
        if (qi < p.Sq) {
          qv = Cvt<T>::to_f(q[(size_t)b * p.q_sb + (size_t)qi * p.q_ss + (size_t)h * p.q_sh + d]);
          gv = Cvt<T>::to_f(dout[(size_t)b * p.o_sb + (size_t)qi * p.o_ss + (size_t)h * p.o_sh + d]);
        }
        Qt[i * hdp + d] = qv; Gt[i * hdp + d] = gv;
      }
      if (threadIdx.x < KT) {
        const int qi = q0 + threadIdx.x;
        Ls[threadIdx.x] = (qi < p.Sq) ? lse[((size_t)b * p.H + h) * p.Sq + qi] : -INFINITY;
        Dl[threadIdx.x] = (qi < p.Sq) ? delta[((size_t)b * p.H + h) * p.Sq + qi] : 0.f;
      }
      __syncthreads();
      if (!kvis) continue;
      const int qi = q0 + lane;
      const bool vis = (qi < p.Sq) && (Ls[lane] != -INFINITY) && (!p.causal || (kj <= qi + off && (p.window <= 0 || qi + off - kj < p.window)));
      float pij = 0.f, ds = 0.f;
      if (vis) {
        float s = 0.f, dp = 0.f;
        const float* qr = Qt + lane * hdp; const float* gr = Gt + lane * hdp;
        const float* kr = Kw + w * hd; const float* vr = Vw + w * hd;
        for (int d = 0; d < hd; ++d) { s = fmaf(qr[d], kr[d], s); dp = fmaf(gr[d], vr[d], dp); }
        pij = __expf(s * p.scale - Ls[lane]);
        ds = pij * (dp - Dl[lane]) * p.scale;
      }
      for (int i2 = 0; i2 < KT; ++i2) {
        const float pb = __shfl_sync(0xffffffffu, pij, i2);
        const float db = __shfl_sync(0xffffffffu, ds, i2);
        if (pb != 0.f || db != 0.f) {
#pragma unroll
          for (int i = 0; i < MAXR; ++i) {
            const int d = lane + 32 * i;
            if (d < hd) { dva[i] = fmaf(pb, Gt[i2 * hdp + d], dva[i]); dka[i] = fmaf(db, Qt[i2 * hdp + d], dka[i]); }
          }
        }
      }
    }
  }
  if (!row_ok) return;
  T* dkp = dk + (size_t)b * p.k_sb + (size_t)kj * p.k_ss + (size_t)hk * p.k_sh;
  T* dvp = dv + (size_t)b * p.v_sb + (size_t)kj * p.v_ss + (size_t)hk * p.v_sh;
#pragma unroll
  for (int i = 0; i < MAXR; ++i) {
    const int d = lane + 32 * i;
    if (d < hd) { dkp[d] = Cvt<T>::from_f(dka[i]); dvp[d] = Cvt<T>::from_f(dva[i]); }
  }
}

inline AttnP make_params(int B, int H, int Hkv, int Sq, int Sk, int hd, const long long* st, float scale,
                         int causal, const int64_t* kmask, long long kmask_sb) {
  AttnP p;
  p.B = B; p.H = H; p.Hkv = Hkv; p.Sq = Sq; p.Sk = Sk; p.hd = hd;
  p.q_sb = st[0]; p.q_ss = st[1]; p.q_sh = st[2]; p.k_sb = st[3]; p.k_ss = st[4]; p.k_sh = st[5];
  p.v_sb = st[6]; p.v_ss = st[7]; p.v_sh = st[8]; p.o_sb = st[9]; p.o_ss = st[10]; p.o_sh = st[11];
  // `causal` argument: 0 = none, 1 = causal, W > 1 = causal with a sliding window of W keys
  p.scale = scale; p.causal = causal != 0; p.window = causal > 1 ? causal : 0; p.kmask = kmask; p.kmask_sb = kmask_sb;
  return p;
}
}  // namespace

#define DISPATCH_T(dtype, ...)                                             \
  if ((dtype) == MB200_DTYPE_BF16) { typedef bf16 T; __VA_ARGS__; }        \
  else if ((dtype) == MB200_DTYPE_F32) { typedef float T; __VA_ARGS__; }   \
  else return -EINVAL;

extern "C" {

// strides: 12 element strides {q_b,q_s,q_h, k_b,k_s,k_h, v_b,v_s,v_h, o_b,o_s,o_h}; dq uses q strides,
// dk/dv use k/v strides, dout uses o strides.
int mb200_attn_generic_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Hkv,
                           int Sq, int Sk, int hd, const long long* strides, float scale, int causal,
                           const int64_t* kmask, long long kmask_sb, int dtype, void* stream) {
  if (B <= 0 || Sq <= 0) return MB200_OK;
  if (hd <= 0 || hd > 32 * MAXR || H % Hkv != 0 || Sk <= 0) return -EINVAL;
  AttnP p = make_params(B, H, Hkv, Sq, Sk, hd, strides, scale, causal, kmask, kmask_sb);
  dim3 grid((Sq + WARPS - 1) / WARPS, H, B);
  const size_t smem = (size_t)(2 * KT * (hd + 1) + WARPS * hd) * sizeof(float);
  DISPATCH_T(dtype, {
    cudaFuncSetAttribute(attn_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn_fwd_kernel<T><<<grid, WARPS * 32, smem, (cudaStream_t)stream>>>((const T*)q, (const T*)k, (const T*)v, (T*)o, lse, p);
  });
  MB200_CHECK_LAUNCH(); return MB200_OK;
}

int mb200_attn_generic_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout,
                           const float* lse, float* delta, void* dq, void* dk, void* dv, int B, int H, int Hkv,
                           int Sq, int Sk, int hd, const long long* strides, float scale, int causal,
                           const int64_t* kmask, long long kmask_sb, int dtype, void* stream) {
  if (B <= 0 || Sq <= 0) return MB200_OK;
  if (hd <= 0 || hd > 32 * MAXR || H % Hkv != 0 || Sk <= 0) return -EINVAL;
  AttnP p = make_params(B, H, Hkv, Sq, Sk, hd, strides, scale, causal, kmask, kmask_sb);
  dim3 gq((Sq + WARPS - 1) / WARPS, H, B), gk((Sk + WARPS - 1) / WARPS, Hkv, B);
  const size_t smem = (size_t)(2 * KT * (hd + 1) + 2 * WARPS * hd) * sizeof(float);
  DISPATCH_T(dtype, {
    cudaFuncSetAttribute(attn_bwd_dq_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(attn_bwd_dkv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn_bwd_dq_kernel<T><<<gq, WARPS * 32, smem, (cudaStream_t)stream>>>((const T*)q, (const T*)k, (const T*)v, (const T*)o,
                                                                        (const T*)dout, lse, delta, (T*)dq, p);
    attn_bwd_dkv_kernel<T><<<gk, WARPS * 32, smem, (cudaStream_t)stream>>>((const T*)q, (const T*)k, (const T*)v,
                                                                         (const T*)dout, lse, delta, (T*)dk, (T*)dv, p);
  });
  MB200_CHECK_LAUNCH(); return MB200_OK;
}

}  // extern "C"

// HBM-bound row / elementwise kernels of the LLaVA / Idefics2 hot path (sm_100a).
//   embedding gather + scatter-add backward   (modeling_llava.py:427 -> nn.Embedding)
//   RMSNorm fwd/bwd                           (transformers llama/modeling_llama.py:53-68; idefics2 :795-809)
//   LayerNorm fwd/bwd                         (transformers siglip/modeling_siglip.py:334-336, clip)
//   RoPE (rotate_half form) fwd/bwd in place  (transformers llama/modeling_llama.py:146-168)
//   SwiGLU silu(gate)*up fwd/bwd              (transformers llama/modeling_llama.py:182-184)
//   GELU (erf / tanh / quick) fwd/bwd         (projector :110-118, SigLIP MLP, CLIP MLP)
//   bias add, column-sum (bias grad), add, row-broadcast add (position embeddings), im2col
// All reductions accumulate in fp32; in bf16 mode intermediate roundings follow the reference's
// op-by-op bf16 evaluation order so that results track the HF path as closely as possible.
#include "common.cuh"

namespace {

using mb::Cvt; using mb::Vec8;

// ------------------------------------------------------------------ embedding
template <typename T>
__global__ void __launch_bounds__(256)
embedding_fwd_kernel(const int64_t* __restrict__ ids, const T* __restrict__ table, T* __restrict__ out,
                     long long n, int D, long long V) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long row_bytes = (long long)D * sizeof(T);
  for (long long r = warp0; r < n; r += nwarps) {
    long long id = ids[r];
    T* dst = out + (size_t)r * D;
    if (id < 0 || id >= V) { for (int i = lane; i < D; i += 32) dst[i] = Cvt<T>::from_f(0.f); continue; }
    const T* src = table + (size_t)id * D;
    if ((row_bytes & 15) == 0) {
      const int4* s4 = reinterpret_cast<const int4*>(src); int4* d4 = reinterpret_cast<int4*>(dst);
      const int nv = (int)(row_bytes >> 4);
      for (int i = lane; i < nv; i += 32) d4[i] = __ldg(s4 + i);
    } else {
      for (int i = lane; i < D; i += 32) dst[i] = src[i];
    }
  }
}

__device__ __forceinline__ void atomic_add_T(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add_T(bf16* p, float v) { atomicAdd(p, __float2bfloat16_rn(v)); }

template <typename T>
__global__ void __launch_bounds__(256)
embedding_bwd_kernel(const int64_t* __restrict__ ids, const T* __restrict__ gout, T* __restrict__ gtable,
                     long long n, int D, long long V) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < n; r += nwarps) {
    long long id = ids[r];
    if (id < 0 || id >= V) continue;
    const T* src = gout + (size_t)r * D;
    T* dst = gtable + (size_t)id * D;
    if (sizeof(T) == 2 && (D & 1) == 0) {
      const bf162* s2 = reinterpret_cast<const bf162*>(src); bf162* d2 = reinterpret_cast<bf162*>(dst);
      for (int i = lane; i < D / 2; i += 32) atomicAdd(d2 + i, s2[i]);
    } else {
      for (int i = lane; i < D; i += 32) atomic_add_T(dst + i, Cvt<T>::to_f(src[i]));
    }
  }
}

// ------------------------------------------------------------------ RMSNorm
// y = w * T(x_f32 * rsqrt(mean(x^2) + eps))     (LlamaRMSNorm: cast to input dtype, then scale)
template <typename T>
__global__ void __launch_bounds__(256)
rmsnorm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, T* __restrict__ y,
                   float* __restrict__ rstd_out, long long n, int D, float eps) {
  __shared__ float red[33];
  for (long long r = blockIdx.x; r < n; r += gridDim.x) {
    const T* xr = x + (size_t)r * D; T* yr = y + (size_t)r * D;
    float ss = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) { float v = Cvt<T>::to_f(xr[i]); ss += v * v; }
    ss = mb::block_sum(ss, red);
    const float rstd = rsqrtf(ss / (float)D + eps);
    if (threadIdx.x == 0 && rstd_out) rstd_out[r] = rstd;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      float v = mb::rnd<T>(Cvt<T>::to_f(xr[i]) * rstd);
      yr[i] = Cvt<T>::from_f(Cvt<T>::to_f(w[i]) * v);
    }
  }
}

// dx = rstd * (g - xhat * mean(g*xhat)), g = dy*w ; dw partial sums per CTA into dw_part[gridDim.x, D]
template <typename T>
__global__ void __launch_bounds__(256)
rmsnorm_bwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ dy,
                   const float* __restrict__ rstd_in, T* __restrict__ dx, float* __restrict__ dw_part,
                   long long n, int D, int accumulate_dx, const T* __restrict__ dres) {
  extern __shared__ float dw_acc[];   // D floats
  __shared__ float red[33];
  for (int i = threadIdx.x; i < D; i += blockDim.x) dw_acc[i] = 0.f;
  __syncthreads();
  for (long long r = blockIdx.x; r < n; r += gridDim.x) {
    const T* xr = x + (size_t)r * D; const T* gr = dy + (size_t)r * D; T* dxr = dx + (size_t)r * D;
    const float rstd = rstd_in[r];
    float dot = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      float xh = Cvt<T>::to_f(xr[i]) * rstd, g = Cvt<T>::to_f(gr[i]);
      dot += g * Cvt<T>::to_f(w[i]) * xh;
      dw_acc[i] += g * xh;
    }
    dot = mb::block_sum(dot, red) / (float)D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      float xh = Cvt<T>::to_f(xr[i]) * rstd, g = Cvt<T>::to_f(gr[i]) * Cvt<T>::to_f(w[i]);
      float v = rstd * (g - xh * dot);
      if (accumulate_dx) v += Cvt<T>::to_f(dxr[i]);
      if (dres) v += Cvt<T>::to_f(dres[(size_t)r * D + i]);
      dxr[i] = Cvt<T>::from_f(v);
    }
  }
  __syncthreads();
  if (dw_part) for (int i = threadIdx.x; i < D; i += blockDim.x) dw_part[(size_t)blockIdx.x * D + i] = dw_acc[i];
}

// Few rows (the decode step: one row per sequence): one 128-thread CTA per row so the load -> reduce -> scale chain is
// 4 vectors deep per thread instead of 16; gamma is constant and fetched before griddepcontrol.wait.
template <int NV>   // 16-byte vectors per thread: D = NV * 1024
__global__ void __launch_bounds__(128)
rmsnorm_fwd_row_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ y,
                       float* __restrict__ rstd_out, float eps) {
  constexpr int D = NV * 1024;
  __shared__ float red[4];
  mb::pdl_trigger();
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long r = blockIdx.x;
  int4 wv[NV], xv[NV];
  const int4* wr = reinterpret_cast<const int4*>(w);
#pragma unroll
  for (int v = 0; v < NV; ++v) wv[v] = __ldg(wr + v * 128 + tid);
  mb::pdl_wait();
  const int4* xr = reinterpret_cast<const int4*>(x + (size_t)r * D);
#pragma unroll
  for (int v = 0; v < NV; ++v) xv[v] = mb::ld_stream(xr + v * 128 + tid);
  float ss = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const bf162* h = reinterpret_cast<const bf162*>(&xv[v]);
#pragma unroll
    for (int j = 0; j < 4; ++j) { float2 f = __bfloat1622float2(h[j]); ss += f.x * f.x + f.y * f.y; }
  }
  ss = mb::warp_sum(ss);
  if (lane == 0) red[wid] = ss;
  __syncthreads();
  ss = (red[0] + red[1]) + (red[2] + red[3]);
  const float rstd = rsqrtf(ss / (float)D + eps);
  if (tid == 0 && rstd_out) rstd_out[r] = rstd;
  int4* yr = reinterpret_cast<int4*>(y + (size_t)r * D);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const bf162* h = reinterpret_cast<const bf162*>(&xv[v]);
    const bf162* wh = reinterpret_cast<const bf162*>(&wv[v]);
    int4 o; bf162* oh = reinterpret_cast<bf162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h[j]), g = __bfloat1622float2(wh[j]);
      const bf162 nx = __floats2bfloat162_rn(f.x * rstd, f.y * rstd);        // .to(input_dtype) first, like the reference
      const float2 nf = __bfloat1622float2(nx);
      oh[j] = __floats2bfloat162_rn(g.x * nf.x, g.y * nf.y);
    }
    yr[v * 128 + tid] = o;
  }
}

// ---- vectorised RMSNorm (bf16, D % 256 == 0, D <= 8192): one warp per row, the row lives in registers,
//      16-byte loads/stores, no block barriers.  HBM traffic = 1 read + 1 write of the activation.
template <int NV>   // 16-byte vectors per lane: D = NV * 256
__global__ void __launch_bounds__(256)
rmsnorm_fwd_vec_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ y,
                       float* __restrict__ rstd_out, long long n, float eps) {
  constexpr int D = NV * 256;
  mb::pdl_trigger();          // no-ops unless launched with the programmatic-serialization attribute (decode engine)
  mb::pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < n; r += nwarps) {
    const int4* xr = reinterpret_cast<const int4*>(x + (size_t)r * D);
    int4 xv[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) xv[v] = mb::ld_stream(xr + v * 32 + lane);
    float ss = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const bf162* h = reinterpret_cast<const bf162*>(&xv[v]);
#pragma unroll
      for (int j = 0; j < 4; ++j) { float2 f = __bfloat1622float2(h[j]); ss += f.x * f.x + f.y * f.y; }
    }
    ss = mb::warp_sum(ss);
    const float rstd = rsqrtf(ss / (float)D + eps);
    if (lane == 0 && rstd_out) rstd_out[r] = rstd;
    int4* yr = reinterpret_cast<int4*>(y + (size_t)r * D);
    const int4* wr = reinterpret_cast<const int4*>(w);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int4 wv = __ldg(wr + v * 32 + lane);
      const bf162* h = reinterpret_cast<const bf162*>(&xv[v]);
      const bf162* wh = reinterpret_cast<const bf162*>(&wv);
      int4 o; bf162* oh = reinterpret_cast<bf162*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h[j]), g = __bfloat1622float2(wh[j]);
        const bf162 nx = __floats2bfloat162_rn(f.x * rstd, f.y * rstd);        // .to(input_dtype) first, like the reference
        const float2 nf = __bfloat1622float2(nx);
        oh[j] = __floats2bfloat162_rn(g.x * nf.x, g.y * nf.y);
      }
      mb::st_stream(yr + v * 32 + lane, o);
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(128)
rmsnorm_bwd_dx_vec_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ dy,
                          const float* __restrict__ rstd_in, bf16* __restrict__ dx, long long n,
                          const bf16* __restrict__ dres /* nullable: gradient of the residual branch, added to dx */) {
  constexpr int D = NV * 256;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int4* wr = reinterpret_cast<const int4*>(w);
  for (long long r = warp0; r < n; r += nwarps) {
    const int4* xr = reinterpret_cast<const int4*>(x + (size_t)r * D);
    const int4* gr = reinterpret_cast<const int4*>(dy + (size_t)r * D);
    int4 xv[NV], gv[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) { xv[v] = mb::ld_stream(xr + v * 32 + lane); gv[v] = mb::ld_stream(gr + v * 32 + lane); }
    const float rstd = rstd_in[r];
    float dot = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int4 wv = __ldg(wr + v * 32 + lane);
      const bf162* xh = reinterpret_cast<const bf162*>(&xv[v]);
      const bf162* gh = reinterpret_cast<const bf162*>(&gv[v]);
      const bf162* wh = reinterpret_cast<const bf162*>(&wv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 xf = __bfloat1622float2(xh[j]), gf = __bfloat1622float2(gh[j]), wf = __bfloat1622float2(wh[j]);
        dot += gf.x * wf.x * xf.x + gf.y * wf.y * xf.y;
      }
    }
    dot = mb::warp_sum(dot) * rstd / (float)D;       // mean(g*w*xhat)
    int4* dxr = reinterpret_cast<int4*>(dx + (size_t)r * D);
    const int4* rr = dres ? reinterpret_cast<const int4*>(dres + (size_t)r * D) : nullptr;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int4 wv = __ldg(wr + v * 32 + lane);
      int4 rv = make_int4(0, 0, 0, 0);                 // bf16 zeros
      if (rr) rv = mb::ld_stream(rr + v * 32 + lane);
      const bf162* xh = reinterpret_cast<const bf162*>(&xv[v]);
      const bf162* gh = reinterpret_cast<const bf162*>(&gv[v]);
      const bf162* wh = reinterpret_cast<const bf162*>(&wv);
      const bf162* rh = reinterpret_cast<const bf162*>(&rv);
      int4 o; bf162* oh = reinterpret_cast<bf162*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 xf = __bfloat1622float2(xh[j]), gf = __bfloat1622float2(gh[j]), wf = __bfloat1622float2(wh[j]);
        const float2 rf = __bfloat1622float2(rh[j]);
        oh[j] = __floats2bfloat162_rn(rstd * (gf.x * wf.x - xf.x * rstd * dot) + rf.x,
                                      rstd * (gf.y * wf.y - xf.y * rstd * dot) + rf.y);
      }
      mb::st_stream(dxr + v * 32 + lane, o);
    }
  }
}

// dx AND the dw partials in one pass over x / dy (the two-kernel version read both tensors twice: 69 + 38 + 21 us per norm
// at [7864, 4096], 2.4 % of the training step).  A thread owns 8 consecutive columns of every row its CTA visits: the
// weight-gradient partial sums stay in 8 registers, the row dot product is a block reduction shared by R rows per barrier.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS >= 512) ? 2 : 4)
rmsnorm_bwd_fused_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ dy,
                         const float* __restrict__ rstd_in, bf16* __restrict__ dx, const bf16* __restrict__ dres,
                         float* __restrict__ dw_part /* [gridDim.x, D] or null */, long long n, int prefetch) {
  constexpr int D = THREADS * 8, NW = THREADS / 32, R = 2;
  __shared__ float red[2][R][NW];                    // double-buffered by iteration parity: one barrier per R rows
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = threadIdx.x * 8;
  float wf[8], acc[8];
  {
    const int4 wv = __ldg(reinterpret_cast<const int4*>(w + c0));
    const bf162* wh = reinterpret_cast<const bf162*>(&wv);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(wh[j]); wf[2 * j] = t.x; wf[2 * j + 1] = t.y; acc[2 * j] = acc[2 * j + 1] = 0.f; }
  }
  const int4 z = make_int4(0, 0, 0, 0);
  int par = 0;
  for (long long r0 = (long long)blockIdx.x * R; r0 < n; r0 += (long long)gridDim.x * R, par ^= 1) {
    int4 xv[R], gv[R], rv[R]; float rs[R], dot[R];
    if (prefetch && (lane & 7) == 0) {               // the rows of the NEXT visit start their trip from DRAM to the L2 now
#pragma unroll
      for (int rr = 0; rr < R; ++rr) {
        const long long r = r0 + (long long)gridDim.x * R + rr;
        if (r < n) {
          asm volatile("prefetch.global.L2 [%0];" :: "l"(x + (size_t)r * D + c0));
          asm volatile("prefetch.global.L2 [%0];" :: "l"(dy + (size_t)r * D + c0));
          if (dres) asm volatile("prefetch.global.L2 [%0];" :: "l"(dres + (size_t)r * D + c0));
        }
      }
    }
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const long long r = r0 + rr; const bool ok = r < n;
      xv[rr] = ok ? mb::ld_stream(reinterpret_cast<const int4*>(x + (size_t)r * D + c0)) : z;
      gv[rr] = ok ? mb::ld_stream(reinterpret_cast<const int4*>(dy + (size_t)r * D + c0)) : z;
      rv[rr] = (ok && dres) ? mb::ld_stream(reinterpret_cast<const int4*>(dres + (size_t)r * D + c0)) : z;
      rs[rr] = ok ? rstd_in[r] : 0.f;
    }
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const bf162* xh = reinterpret_cast<const bf162*>(&xv[rr]);
      const bf162* gh = reinterpret_cast<const bf162*>(&gv[rr]);
      float d = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 xf = __bfloat1622float2(xh[j]), gf = __bfloat1622float2(gh[j]);
        d += gf.x * wf[2 * j] * xf.x + gf.y * wf[2 * j + 1] * xf.y;
        acc[2 * j] += gf.x * xf.x * rs[rr]; acc[2 * j + 1] += gf.y * xf.y * rs[rr];
      }
      dot[rr] = mb::warp_sum(d);
    }
    if (lane == 0) {
#pragma unroll
      for (int rr = 0; rr < R; ++rr) red[par][rr][warp] = dot[rr];
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const long long r = r0 + rr;
      if (r >= n) break;
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < NW; ++k) t += red[par][rr][k];
      const float rstd = rs[rr];
      const float dm = t * rstd / (float)D;          // mean(g*w*xhat)
      const bf162* xh = reinterpret_cast<const bf162*>(&xv[rr]);
      const bf162* gh = reinterpret_cast<const bf162*>(&gv[rr]);
      const bf162* rh = reinterpret_cast<const bf162*>(&rv[rr]);
      int4 o; bf162* oh = reinterpret_cast<bf162*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 xf = __bfloat1622float2(xh[j]), gf = __bfloat1622float2(gh[j]), rf = __bfloat1622float2(rh[j]);
        oh[j] = __floats2bfloat162_rn(rstd * (gf.x * wf[2 * j] - xf.x * rstd * dm) + rf.x,
                                      rstd * (gf.y * wf[2 * j + 1] - xf.y * rstd * dm) + rf.y);
      }
      mb::st_stream(reinterpret_cast<int4*>(dx + (size_t)r * D + c0), o);
    }
  }
  if (dw_part) {
    float4* dst = reinterpret_cast<float4*>(dw_part + (size_t)blockIdx.x * D + c0);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// dw partials: part[blockIdx.x, c] = sum over this CTA's row slab of dy[r,c] * x[r,c] * rstd[r]  (thread owns 8-wide column groups)
__global__ void __launch_bounds__(256)
rmsnorm_bwd_dw_vec_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, const float* __restrict__ rstd_in,
                          float* __restrict__ part, long long n, int D) {
  const long long rows_per = (n + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * rows_per, r1 = min(n, r0 + rows_per);
  for (int c0 = threadIdx.x * 8; c0 < D; c0 += 256 * 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (long long r = r0; r < r1; ++r) {
      const int4 xv = mb::ld_stream(reinterpret_cast<const int4*>(x + (size_t)r * D + c0));
      const int4 gv = mb::ld_stream(reinterpret_cast<const int4*>(dy + (size_t)r * D + c0));
      const float rs = rstd_in[r];
      const bf162* xh = reinterpret_cast<const bf162*>(&xv); const bf162* gh = reinterpret_cast<const bf162*>(&gv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 xf = __bfloat1622float2(xh[j]), gf = __bfloat1622float2(gh[j]);
        acc[2 * j] += gf.x * xf.x * rs; acc[2 * j + 1] += gf.y * xf.y * rs;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) part[(size_t)blockIdx.x * D + c0 + j] = acc[j];
  }
}

// out[i] (+)= sum_p part[p, i]   -> T.   64 columns x 4 interleaved part groups per CTA (launch with COLSUM_GRID(D) x 256):
// the one-thread-per-column version walked the ~300 part rows serially in 16 CTAs (21 us for D = 4096)
#define COLSUM_GRID(D) (((D) + 63) / 64)
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partials_kernel(const float* __restrict__ part, int nparts, int D, T* __restrict__ out, int accumulate) {
  __shared__ float red[4][64];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int i = blockIdx.x * 64 + tx;
  float s = 0.f;
  if (i < D)
    for (int p = ty; p < nparts; p += 4) s += part[(size_t)p * D + i];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && i < D) {
    s = (red[0][tx] + red[1][tx]) + (red[2][tx] + red[3][tx]);
    if (accumulate) s += Cvt<T>::to_f(out[i]);
    out[i] = Cvt<T>::from_f(s);
  }
}

// ------------------------------------------------------------------ LayerNorm
template <typename T>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ b,
                     T* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                     long long n, int D, float eps) {
  __shared__ float red[33];
  for (long long r = blockIdx.x; r < n; r += gridDim.x) {
    const T* xr = x + (size_t)r * D; T* yr = y + (size_t)r * D;
    float s = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) s += Cvt<T>::to_f(xr[i]);
    const float mean = mb::block_sum(s, red) / (float)D;
    float ss = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) { float d = Cvt<T>::to_f(xr[i]) - mean; ss += d * d; }
    const float rstd = rsqrtf(mb::block_sum(ss, red) / (float)D + eps);
    if (threadIdx.x == 0) { if (mean_out) mean_out[r] = mean; if (rstd_out) rstd_out[r] = rstd; }
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      float v = (Cvt<T>::to_f(xr[i]) - mean) * rstd;
      v = v * Cvt<T>::to_f(w[i]) + (b ? Cvt<T>::to_f(b[i]) : 0.f);
      yr[i] = Cvt<T>::from_f(v);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ dy,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                     T* __restrict__ dx, float* __restrict__ dw_part, float* __restrict__ db_part,
                     long long n, int D) {
  extern __shared__ float acc[];   // 2*D floats: dw, db
  __shared__ float red[33];
  float* dw_acc = acc; float* db_acc = acc + D;
  for (int i = threadIdx.x; i < D; i += blockDim.x) { dw_acc[i] = 0.f; db_acc[i] = 0.f; }
  __syncthreads();
  for (long long r = blockIdx.x; r < n; r += gridDim.x) {
    const T* xr = x + (size_t)r * D; const T* gr = dy + (size_t)r * D; T* dxr = dx + (size_t)r * D;
    const float mean = mean_in[r], rstd = rstd_in[r];
    float s1 = 0.f, s2 = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      float xh = (Cvt<T>::to_f(xr[i]) - mean) * rstd, g0 = Cvt<T>::to_f(gr[i]);
      float g = g0 * Cvt<T>::to_f(w[i]);
      s1 += g; s2 += g * xh;
      dw_acc[i] += g0 * xh; db_acc[i] += g0;
    }
    s1 = mb::block_sum(s1, red) / (float)D;
    s2 = mb::block_sum(s2, red) / (float)D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      float xh = (Cvt<T>::to_f(xr[i]) - mean) * rstd, g = Cvt<T>::to_f(gr[i]) * Cvt<T>::to_f(w[i]);
      dxr[i] = Cvt<T>::from_f(rstd * (g - s1 - xh * s2));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    if (dw_part) dw_part[(size_t)blockIdx.x * D + i] = dw_acc[i];
    if (db_part) db_part[(size_t)blockIdx.x * D + i] = db_acc[i];
  }
}

// ------------------------------------------------------------------ RoPE (in place)
// x: [n_tok, H, hd] (row stride `tok_stride` elements), pos: [n_tok] int64, inv_freq: [hd/2] fp32
// sign=+1 forward, -1 backward (rotation by -angle).
template <typename T>
__global__ void __launch_bounds__(256)
rope_kernel(const T* x, T* y, const int64_t* __restrict__ pos, const float* __restrict__ inv_freq,
            long long n_tok, int H, int hd, long long tok_stride, long long out_stride, float attn_scaling, float sign) {
  const int half = hd >> 1;
  const long long total = n_tok * H * half;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % half);
    const long long th = idx / half;
    const int h = (int)(th % H);
    const long long t = th / H;
    const float ang = (float)pos[t] * inv_freq[i];
    float sn, cs;
    sincosf(ang, &sn, &cs);
    cs = mb::rnd<T>(cs * attn_scaling); sn = mb::rnd<T>(sn * attn_scaling) * sign;
    const T* p = x + (size_t)t * tok_stride + (size_t)h * hd;
    T* q = y + (size_t)t * out_stride + (size_t)h * hd;
    const float x1 = Cvt<T>::to_f(p[i]), x2 = Cvt<T>::to_f(p[i + half]);
    // q_embed = (q * cos) + (rotate_half(q) * sin), each product rounded to T like the reference
    const float y1 = mb::rnd<T>(x1 * cs) + mb::rnd<T>(-x2 * sn);
    const float y2 = mb::rnd<T>(x2 * cs) + mb::rnd<T>(x1 * sn);
    q[i] = Cvt<T>::from_f(y1); q[i + half] = Cvt<T>::from_f(y2);
  }
}

// cos/sin table for one forward pass: tab[t, i] = {T-rounded cos, T-rounded sin}(pos[t] * inv_freq[i]) -- computed once per
// micro-batch and shared by every layer's q and k, forward and backward (the per-element sincosf dominated rope_kernel).
template <typename T>
__global__ void __launch_bounds__(256)
rope_table_kernel(const int64_t* __restrict__ pos, const float* __restrict__ inv_freq, float2* __restrict__ tab,
                  long long n_tok, int half, float attn_scaling) {
  const long long total = n_tok * half;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % half); const long long t = idx / half;
    float sn, cs;
    sincosf((float)pos[t] * inv_freq[i], &sn, &cs);
    tab[idx] = make_float2(mb::rnd<T>(cs * attn_scaling), mb::rnd<T>(sn * attn_scaling));
  }
}

// q and k rotated in ONE launch from the table, 8 elements (16 B) per thread and per half.
// x layout: [n_tok, H, hd] with token stride; heads [0, Hq) come from q, [Hq, Hq + Hk) from k.
__global__ void __launch_bounds__(256)
rope2_bf16_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, bf16* __restrict__ qo, bf16* __restrict__ ko,
                  const float2* __restrict__ tab, long long n_tok, int Hq, int Hk, int hd, long long q_stride,
                  long long k_stride, float sign) {
  const int half = hd >> 1, vec_per_head = half >> 3;           // 8-element vectors per half head
  const long long total = n_tok * (Hq + Hk) * vec_per_head;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(idx % vec_per_head);
    const long long th = idx / vec_per_head;
    const int h = (int)(th % (Hq + Hk));
    const long long t = th / (Hq + Hk);
    const bool isq = h < Hq;
    const bf16* src = isq ? q + (size_t)t * q_stride + (size_t)h * hd : k + (size_t)t * k_stride + (size_t)(h - Hq) * hd;
    bf16* dst = isq ? qo + ((size_t)t * Hq + h) * hd : ko + ((size_t)t * Hk + (h - Hq)) * hd;
    float x1[8], x2[8], y1[8], y2[8];
    mb::Vec8<bf16>::load(src + v * 8, x1);
    mb::Vec8<bf16>::load(src + half + v * 8, x2);
    const float2* tb = tab + (size_t)t * half + v * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 cs = tb[j];
      const float sn = cs.y * sign;
      // q_embed = (q * cos) + (rotate_half(q) * sin), each product rounded to bf16 like the reference
      y1[j] = mb::rnd<bf16>(x1[j] * cs.x) + mb::rnd<bf16>(-x2[j] * sn);
      y2[j] = mb::rnd<bf16>(x2[j] * cs.x) + mb::rnd<bf16>(x1[j] * sn);
    }
    mb::Vec8<bf16>::store(dst + v * 8, y1);
    mb::Vec8<bf16>::store(dst + half + v * 8, y2);
  }
}

// ------------------------------------------------------------------ SwiGLU
__device__ __forceinline__ float silu_f(float g) { return g / (1.f + __expf(-g)); }

template <typename T>
__global__ void __launch_bounds__(256)
swiglu_fwd_kernel(const T* __restrict__ gate, const T* __restrict__ up, T* __restrict__ out, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float g[8], u[8], o[8];
    Vec8<T>::load(gate + i * 8, g); Vec8<T>::load(up + i * 8, u);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = mb::rnd<T>(silu_f(g[j])) * u[j];
    Vec8<T>::store(out + i * 8, o);
  }
}
template <typename T>
__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(const T* __restrict__ gate, const T* __restrict__ up, const T* __restrict__ dout,
                  T* __restrict__ dgate, T* __restrict__ dup, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float g[8], u[8], d[8], dg[8], du[8];
    Vec8<T>::load(gate + i * 8, g); Vec8<T>::load(up + i * 8, u); Vec8<T>::load(dout + i * 8, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sg = 1.f / (1.f + __expf(-g[j]));
      const float s = g[j] * sg;
      du[j] = d[j] * s;
      dg[j] = d[j] * u[j] * (sg * (1.f + g[j] * (1.f - sg)));
    }
    Vec8<T>::store(dgate + i * 8, dg); Vec8<T>::store(dup + i * 8, du);
  }
}

// ------------------------------------------------------------------ GELU family (kind: 0 erf, 1 tanh, 2 quick)
__device__ __forceinline__ float act_fwd(float x, int kind) {
  if (kind == 0) return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
  if (kind == 1) { const float k = 0.79788456080286535588f; return 0.5f * x * (1.f + tanhf(k * (x + 0.044715f * x * x * x))); }
  return x / (1.f + __expf(-1.702f * x));
}
__device__ __forceinline__ float act_bwd(float x, int kind) {
  if (kind == 0) {
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
    return cdf + x * 0.39894228040143267794f * __expf(-0.5f * x * x);
  }
  if (kind == 1) {
    const float k = 0.79788456080286535588f, c = 0.044715f;
    const float u = k * (x + c * x * x * x), t = tanhf(u);
    return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * k * (1.f + 3.f * c * x * x);
  }
  const float s = 1.f / (1.f + __expf(-1.702f * x));
  return s + 1.702f * x * s * (1.f - s);
}
template <typename T>
__global__ void __launch_bounds__(256)
act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long n8, int kind) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float a[8], o[8];
    Vec8<T>::load(x + i * 8, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = act_fwd(a[j], kind);
    Vec8<T>::store(y + i * 8, o);
  }
}
template <typename T>
__global__ void __launch_bounds__(256)
act_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, long long n8, int kind) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float a[8], d[8], o[8];
    Vec8<T>::load(x + i * 8, a); Vec8<T>::load(dy + i * 8, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = d[j] * act_bwd(a[j], kind);
    Vec8<T>::store(dx + i * 8, o);
  }
}

// ------------------------------------------------------------------ add / bias / row-broadcast add
template <typename T>
__global__ void __launch_bounds__(256)
add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float p[8], q[8];
    Vec8<T>::load(a + i * 8, p); Vec8<T>::load(b + i * 8, q);
#pragma unroll
    for (int j = 0; j < 8; ++j) p[j] += q[j];
    Vec8<T>::store(y + i * 8, p);
  }
}
// y[r, :] = x[r, :] + table[idx ? idx[r] : (r % period), :]
template <typename T>
__global__ void __launch_bounds__(256)
add_rows_kernel(const T* __restrict__ x, const T* __restrict__ table, const int64_t* __restrict__ idx,
                T* __restrict__ y, long long n, int D, long long period) {
  const long long total = n * D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / D; const int c = (int)(e % D);
    const long long tr = idx ? idx[r] : (r % period);
    y[e] = Cvt<T>::from_f(Cvt<T>::to_f(x[e]) + Cvt<T>::to_f(table[(size_t)tr * D + c]));
  }
}
// column sums of a [n, N] matrix (bias gradient): two stage, fp32 partials
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, float* __restrict__ part, long long n, int N, long long ld) {
  // grid.x over column tiles of 256, grid.y over row slabs
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const long long rows_per = (n + gridDim.y - 1) / gridDim.y;
  const long long r0 = (long long)blockIdx.y * rows_per, r1 = min(n, r0 + rows_per);
  if (c >= N) return;
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s += Cvt<T>::to_f(x[(size_t)r * ld + c]);
  part[(size_t)blockIdx.y * N + c] = s;
}

// ------------------------------------------------------------------ im2col (patch embed as GEMM)
// pixels [N, C, H, W] (Tin) -> patches [N*gh*gw, Kpad] (Tout), K order (c, ky, kx) == conv weight.flatten(1)
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256)
im2col_kernel(const Tin* __restrict__ px, Tout* __restrict__ out, int N, int C, int H, int W, int p,
              int gh, int gw, int K, int Kpad) {
  const long long total = (long long)N * gh * gw * Kpad;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % Kpad);
    const long long row = e / Kpad;
    float v = 0.f;
    if (k < K) {
      const int kx = k % p, ky = (k / p) % p, c = k / (p * p);
      const int gx = (int)(row % gw), gy = (int)((row / gw) % gh);
      const long long n = row / ((long long)gw * gh);
      v = Cvt<Tin>::to_f(px[(((size_t)n * C + c) * H + (gy * p + ky)) * W + (gx * p + kx)]);
    }
    out[e] = Cvt<Tout>::from_f(v);
  }
}

// ------------------------------------------------------------------ cast
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256)
cast_kernel(const Tin* __restrict__ x, Tout* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = Cvt<Tout>::from_f(Cvt<Tin>::to_f(x[i]));
}

// flags[r] = 1 iff every element of row r is exactly zero (Idefics2 padding-image detection,
// mantis/models/idefics2/modeling_idefics2.py:1637-1639); one CTA per row, early exit per chunk
template <typename T>
__global__ void __launch_bounds__(256)
rows_all_zero_kernel(const T* __restrict__ x, long long row_elems, int* __restrict__ flags) {
  __shared__ int nz;
  if (threadIdx.x == 0) nz = 0;
  __syncthreads();
  const T* row = x + (size_t)blockIdx.x * row_elems;
  for (long long base = 0; base < row_elems; base += 256 * 16) {
    bool any = false;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      long long i = base + (long long)j * 256 + threadIdx.x;
      if (i < row_elems) any |= !(Cvt<T>::to_f(row[i]) == 0.0f);
    }
    if (any) nz = 1;
    __syncthreads();
    if (nz) break;
  }
  if (threadIdx.x == 0) flags[blockIdx.x] = nz ? 0 : 1;
}

inline int ew_grid(long long work_items, int threads = 256) {
  long long g = (work_items + threads - 1) / threads;
  long long cap = (long long)mb::num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

#define DISPATCH_T(dtype, ...)                                             \
  if ((dtype) == MB200_DTYPE_BF16) { typedef bf16 T; __VA_ARGS__; }        \
  else if ((dtype) == MB200_DTYPE_F32) { typedef float T; __VA_ARGS__; }   \
  else return -EINVAL;


// ---------------------------------------------------------------- antialiased resize of 8-bit images (caller side, 8f-2)
// One pass of Pillow's two-pass resampler (src/libImaging/Resample.c, ImagingResampleHorizontal_8bpc / Vertical_8bpc), which
// the reference's image processors call through PIL for every image: acc = 2^21 + sum_k pixel * coef (22-bit fixed point,
// int32), out = clip8(acc >> 22).  The per-output-index tap ranges and coefficients come from the host (double precision,
// same operation order as Pillow), so the device result is bit-identical to Image.resize.  HWC uint8 in and out.
namespace {
template <bool HORIZ>
__global__ void __launch_bounds__(256)
resize_pass_u8_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, const int* __restrict__ bounds,
                      const int* __restrict__ coef, int ksize, int in_h, int in_w, int out_len, int C) {
  const int oh = HORIZ ? in_h : out_len, ow = HORIZ ? out_len : in_w;
  const long long total = (long long)oh * ow * C;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const int x = (int)((t / C) % ow);
    const int y = (int)(t / ((long long)C * ow));
    const int o = HORIZ ? x : y;
    const int lo = bounds[2 * o], n = bounds[2 * o + 1];
    const int* k = coef + (size_t)o * ksize;
    int acc = 1 << 21;
    if (HORIZ) {
      const unsigned char* p = in + ((size_t)y * in_w + lo) * C + c;
      for (int i = 0; i < n; ++i) acc += (int)p[(size_t)i * C] * k[i];
    } else {
      const unsigned char* p = in + ((size_t)lo * in_w + x) * C + c;
      for (int i = 0; i < n; ++i) acc += (int)p[(size_t)i * in_w * C] * k[i];
    }
    acc >>= 22;
    out[t] = (unsigned char)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
  }
}
}  // namespace

// ---------------------------------------------------------------- image rescale + normalize + layout (caller side, 8f-2)
// uint8 pixels (HWC as decoded, or CHW as HF processors hand them over) -> normalized [N, C, H, W] fp32/bf16 through a
// 256-entry table per channel.  The table holds (float32(float64(v) * rescale) - mean) / std evaluated exactly like the
// reference's numpy image processor, so the result is bit-identical to it for every pixel value; the device only gathers.
// One thread = 8 consecutive pixels of one output row: 16-byte (bf16) / 2 x 16-byte (fp32) coalesced stores.
namespace {
template <typename OT>
__global__ void __launch_bounds__(256)
image_normalize_u8_kernel(const unsigned char* __restrict__ px, const float* __restrict__ lut, OT* __restrict__ out, int N, int C,
                          int H, int W, int channels_last) {
  __shared__ float tab[4 * 256];
  for (int i = threadIdx.x; i < C * 256; i += blockDim.x) tab[i] = lut[i];
  __syncthreads();
  const int wv = (W + 7) / 8;
  const long long total = (long long)N * C * H * wv;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int w0 = (int)(t % wv) * 8;
    long long r = t / wv;
    const int h = (int)(r % H); r /= H;
    const int c = (int)(r % C);
    const long long n = r / C;
    OT* o = out + (((size_t)n * C + c) * H + h) * W + w0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (w0 + j < W) {
        const size_t src = channels_last ? ((((size_t)n * H + h) * W + w0 + j) * C + c) : ((((size_t)n * C + c) * H + h) * W + w0 + j);
        mb::stf(o + j, tab[c * 256 + px[src]]);
      }
    }
  }
}
}  // namespace

extern "C" {

int mb200_embedding_fwd(const int64_t* ids, const void* table, void* out, long long n, int D, long long V,
                        int dtype, void* stream) {
  if (n <= 0) return MB200_OK;
  DISPATCH_T(dtype, (embedding_fwd_kernel<T><<<ew_grid(n * 32), 256, 0, (cudaStream_t)stream>>>(
                        ids, (const T*)table, (T*)out, n, D, V)));
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_embedding_bwd(const int64_t* ids, const void* gout, void* gtable, long long n, int D, long long V,
                        int dtype, void* stream) {
  if (n <= 0) return MB200_OK;
  DISPATCH_T(dtype, (embedding_bwd_kernel<T><<<ew_grid(n * 32), 256, 0, (cudaStream_t)stream>>>(
                        ids, (const T*)gout, (T*)gtable, n, D, V)));
  MB200_CHECK_LAUNCH(); return MB200_OK;
}

int mb200_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, long long n, int D, float eps,
                      int dtype, void* stream) {
  if (n <= 0) return MB200_OK;
  if (dtype == MB200_DTYPE_BF16 && (D == 4096 || D == 2048 || D == 1024) &&
      !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w)) & 15)) {
    long long g = (n + 7) / 8; const long long cap = (long long)mb::num_sms() * 8; if (g > cap) g = cap;
    cudaStream_t st = (cudaStream_t)stream;
    const bool pdl = mb::pdl_mode() != 0;
    if (n <= 32) {
      if (D == 4096) mb::launch_ex(rmsnorm_fwd_row_kernel<4>, dim3((int)n), dim3(128), 0, st, pdl, (const bf16*)x, (const bf16*)w, (bf16*)y, rstd, eps);
      else if (D == 2048) mb::launch_ex(rmsnorm_fwd_row_kernel<2>, dim3((int)n), dim3(128), 0, st, pdl, (const bf16*)x, (const bf16*)w, (bf16*)y, rstd, eps);
      else mb::launch_ex(rmsnorm_fwd_row_kernel<1>, dim3((int)n), dim3(128), 0, st, pdl, (const bf16*)x, (const bf16*)w, (bf16*)y, rstd, eps);
      MB200_CHECK_LAUNCH(); return MB200_OK;
    }
    if (D == 4096) mb::launch_ex(rmsnorm_fwd_vec_kernel<16>, dim3((int)g), dim3(256), 0, st, pdl, (const bf16*)x, (const bf16*)w, (bf16*)y, rstd, n, eps);
    else if (D == 2048) mb::launch_ex(rmsnorm_fwd_vec_kernel<8>, dim3((int)g), dim3(256), 0, st, pdl, (const bf16*)x, (const bf16*)w, (bf16*)y, rstd, n, eps);
    else mb::launch_ex(rmsnorm_fwd_vec_kernel<4>, dim3((int)g), dim3(256), 0, st, pdl, (const bf16*)x, (const bf16*)w, (bf16*)y, rstd, n, eps);
    MB200_CHECK_LAUNCH(); return MB200_OK;
  }
  int grid = (int)(n < (long long)mb::num_sms() * 8 ? n : (long long)mb::num_sms() * 8);
  DISPATCH_T(dtype, (rmsnorm_fwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
                        (const T*)x, (const T*)w, (T*)y, rstd, n, D, eps)));
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
// number of partial rows the bwd kernels write (= grid size); caller allocates [parts, D] fp32 per output
int mb200_norm_bwd_parts(long long n) {
  long long g = (long long)mb::num_sms() * 2; if (n < g) g = n; if (g < 1) g = 1; return (int)g;
}
static int rmsnorm_bwd_impl(const void* x, const void* w, const void* dy, const float* rstd, const void* dres, void* dx,
                            float* dw_part, void* dw, int accumulate_dw, int accumulate_dx, long long n, int D,
                            int dtype, void* stream) {
  if (n <= 0) return MB200_OK;
  if (dtype == MB200_DTYPE_BF16 && (D == 4096 || D == 2048 || D == 1024) && !accumulate_dx &&
      !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) |
         reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(dres)) & 15)) {
    cudaStream_t st = (cudaStream_t)stream;
    static const int fused_on = [] { const char* e = getenv("MB200_RMSNORM_BWD_FUSED"); return (e && e[0] == '0') ? 0 : 1; }();
    static const int pf = [] { const char* e = getenv("MB200_RMSNORM_BWD_PREFETCH"); return (e && e[0] == '0') ? 0 : 1; }();
    if (fused_on && dw && dw_part && !(reinterpret_cast<uintptr_t>(dw_part) & 15)) {
      // one pass: dx + per-CTA dw partials (grid <= the `parts` rows the caller allocated, one wave of resident CTAs)
      int grid = mb200_norm_bwd_parts(n);
      const long long need = (n + 1) / 2; if (grid > need) grid = (int)need;
      if (D == 4096) rmsnorm_bwd_fused_kernel<512><<<grid, 512, 0, st>>>((const bf16*)x, (const bf16*)w, (const bf16*)dy, rstd, (bf16*)dx, (const bf16*)dres, dw_part, n, pf);
      else if (D == 2048) rmsnorm_bwd_fused_kernel<256><<<grid, 256, 0, st>>>((const bf16*)x, (const bf16*)w, (const bf16*)dy, rstd, (bf16*)dx, (const bf16*)dres, dw_part, n, pf);
      else rmsnorm_bwd_fused_kernel<128><<<grid, 128, 0, st>>>((const bf16*)x, (const bf16*)w, (const bf16*)dy, rstd, (bf16*)dx, (const bf16*)dres, dw_part, n, pf);
      colsum_partials_kernel<bf16><<<COLSUM_GRID(D), 256, 0, st>>>(dw_part, grid, D, (bf16*)dw, accumulate_dw);
      MB200_CHECK_LAUNCH(); return MB200_OK;
    }
    long long g = (n + 3) / 4; const long long cap = (long long)mb::num_sms() * 12; if (g > cap) g = cap;
    if (D == 4096) rmsnorm_bwd_dx_vec_kernel<16><<<(int)g, 128, 0, st>>>((const bf16*)x, (const bf16*)w, (const bf16*)dy, rstd, (bf16*)dx, n, (const bf16*)dres);
    else if (D == 2048) rmsnorm_bwd_dx_vec_kernel<8><<<(int)g, 128, 0, st>>>((const bf16*)x, (const bf16*)w, (const bf16*)dy, rstd, (bf16*)dx, n, (const bf16*)dres);
    else rmsnorm_bwd_dx_vec_kernel<4><<<(int)g, 128, 0, st>>>((const bf16*)x, (const bf16*)w, (const bf16*)dy, rstd, (bf16*)dx, n, (const bf16*)dres);
    if (dw && dw_part) {
      const int parts = mb200_norm_bwd_parts(n);
      rmsnorm_bwd_dw_vec_kernel<<<parts, 256, 0, st>>>((const bf16*)x, (const bf16*)dy, rstd, dw_part, n, D);
      colsum_partials_kernel<bf16><<<COLSUM_GRID(D), 256, 0, st>>>(dw_part, parts, D, (bf16*)dw, accumulate_dw);
    }
    MB200_CHECK_LAUNCH(); return MB200_OK;
  }
  const int grid = mb200_norm_bwd_parts(n);
  const size_t smem = (size_t)D * sizeof(float);
  DISPATCH_T(dtype, {
    cudaFuncSetAttribute(rmsnorm_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rmsnorm_bwd_kernel<T><<<grid, 256, smem, (cudaStream_t)stream>>>((const T*)x, (const T*)w, (const T*)dy, rstd,
                                                                   (T*)dx, dw_part, n, D, accumulate_dx, (const T*)dres);
    if (dw && dw_part)
      colsum_partials_kernel<T><<<COLSUM_GRID(D), 256, 0, (cudaStream_t)stream>>>(dw_part, grid, D, (T*)dw, accumulate_dw);
  });
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_rmsnorm_bwd(const void* x, const void* w, const void* dy, const float* rstd, void* dx,
                      float* dw_part, void* dw, int accumulate_dw, int accumulate_dx, long long n, int D,
                      int dtype, void* stream) {
  return rmsnorm_bwd_impl(x, w, dy, rstd, nullptr, dx, dw_part, dw, accumulate_dw, accumulate_dx, n, D, dtype, stream);
}
// dx = rmsnorm_backward(dy) + dres: the gradient of the residual branch that by-passes the norm is summed in the same pass
// (autograd would otherwise add the two [n, D] tensors with a separate elementwise kernel).
int mb200_rmsnorm_bwd_res(const void* x, const void* w, const void* dy, const float* rstd, const void* dres, void* dx,
                          float* dw_part, void* dw, int accumulate_dw, long long n, int D, int dtype, void* stream) {
  return rmsnorm_bwd_impl(x, w, dy, rstd, dres, dx, dw_part, dw, accumulate_dw, 0, n, D, dtype, stream);
}

int mb200_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd,
                        long long n, int D, float eps, int dtype, void* stream) {
  if (n <= 0) return MB200_OK;
  int grid = (int)(n < (long long)mb::num_sms() * 8 ? n : (long long)mb::num_sms() * 8);
  DISPATCH_T(dtype, (layernorm_fwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
                        (const T*)x, (const T*)w, (const T*)b, (T*)y, mean, rstd, n, D, eps)));
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_layernorm_bwd(const void* x, const void* w, const void* dy, const float* mean, const float* rstd,
                        void* dx, float* dw_part, float* db_part, void* dw, void* db, int accumulate,
                        long long n, int D, int dtype, void* stream) {
  if (n <= 0) return MB200_OK;
  const int grid = mb200_norm_bwd_parts(n);
  const size_t smem = (size_t)2 * D * sizeof(float);
  DISPATCH_T(dtype, {
    cudaFuncSetAttribute(layernorm_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    layernorm_bwd_kernel<T><<<grid, 256, smem, (cudaStream_t)stream>>>((const T*)x, (const T*)w, (const T*)dy, mean, rstd,
                                                                     (T*)dx, dw_part, db_part, n, D);
    if (dw && dw_part)
      colsum_partials_kernel<T><<<COLSUM_GRID(D), 256, 0, (cudaStream_t)stream>>>(dw_part, grid, D, (T*)dw, accumulate);
    if (db && db_part)
      colsum_partials_kernel<T><<<COLSUM_GRID(D), 256, 0, (cudaStream_t)stream>>>(db_part, grid, D, (T*)db, accumulate);
  });
  MB200_CHECK_LAUNCH(); return MB200_OK;
}

int mb200_rope(const void* x, void* y, const int64_t* pos, const float* inv_freq, long long n_tok, int H, int hd,
               long long tok_stride, long long out_stride, float attn_scaling, int backward, int dtype, void* stream) {
  if (n_tok <= 0) return MB200_OK;
  if (hd & 1) return -EINVAL;
  const long long total = n_tok * H * (hd / 2);
  DISPATCH_T(dtype, (rope_kernel<T><<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)x, (T*)y, pos, inv_freq, n_tok, H, hd, tok_stride, out_stride, attn_scaling,
                        backward ? -1.f : 1.f)));
  MB200_CHECK_LAUNCH(); return MB200_OK;
}

int mb200_rope_table(const int64_t* pos, const float* inv_freq, void* tab, long long n_tok, int hd, float attn_scaling,
                     int dtype, void* stream) {
  if (n_tok <= 0) return MB200_OK;
  if (hd & 1) return -EINVAL;
  const long long total = n_tok * (hd / 2);
  DISPATCH_T(dtype, (rope_table_kernel<T><<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(pos, inv_freq, (float2*)tab, n_tok, hd / 2,
                                                                                           attn_scaling)));
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
// bf16, hd % 16 == 0, 16-byte aligned rows; outputs contiguous [n_tok, Hq, hd] / [n_tok, Hk, hd]
int mb200_rope2_bf16(const void* q, const void* k, void* qo, void* ko, const void* tab, long long n_tok, int Hq, int Hk,
                     int hd, long long q_stride, long long k_stride, int backward, void* stream) {
  if (n_tok <= 0) return MB200_OK;
  if ((hd & 15) || (q_stride & 7) || (k_stride & 7)) return -ENOTSUP;
  const long long total = n_tok * (Hq + Hk) * (hd / 16);
  rope2_bf16_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>((const bf16*)q, (const bf16*)k, (bf16*)qo, (bf16*)ko,
                                                                     (const float2*)tab, n_tok, Hq, Hk, hd, q_stride, k_stride,
                                                                     backward ? -1.f : 1.f);
  MB200_CHECK_LAUNCH(); return MB200_OK;
}

int mb200_swiglu_fwd(const void* gate, const void* up, void* out, long long n, int dtype, void* stream) {
  if (n <= 0) return MB200_OK;
  if (n & 7) return -EINVAL;
  DISPATCH_T(dtype, (swiglu_fwd_kernel<T><<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)gate, (const T*)up, (T*)out, n / 8)));
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_swiglu_bwd(const void* gate, const void* up, const void* dout, void* dgate, void* dup, long long n,
                     int dtype, void* stream) {
  if (n <= 0) return MB200_OK;
  if (n & 7) return -EINVAL;
  DISPATCH_T(dtype, (swiglu_bwd_kernel<T><<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)gate, (const T*)up, (const T*)dout, (T*)dgate, (T*)dup, n / 8)));
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_act_fwd(const void* x, void* y, long long n, int kind, int dtype, void* stream) {
  if (n <= 0) return MB200_OK;
  if ((n & 7) || kind < 0 || kind > 2) return -EINVAL;
  DISPATCH_T(dtype, (act_fwd_kernel<T><<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, n / 8, kind)));
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_act_bwd(const void* x, const void* dy, void* dx, long long n, int kind, int dtype, void* stream) {
  if (n <= 0) return MB200_OK;
  if ((n & 7) || kind < 0 || kind > 2) return -EINVAL;
  DISPATCH_T(dtype, (act_bwd_kernel<T><<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)x, (const T*)dy, (T*)dx, n / 8, kind)));
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_add(const void* a, const void* b, void* y, long long n, int dtype, void* stream) {
  if (n <= 0) return MB200_OK;
  if (n & 7) return -EINVAL;
  DISPATCH_T(dtype, (add_kernel<T><<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((const T*)a, (const T*)b, (T*)y, n / 8)));
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_add_rows(const void* x, const void* table, const int64_t* idx, void* y, long long n, int D,
                   long long period, int dtype, void* stream) {
  if (n <= 0) return MB200_OK;
  if (!idx && period <= 0) return -EINVAL;
  DISPATCH_T(dtype, (add_rows_kernel<T><<<ew_grid(n * D), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)x, (const T*)table, idx, (T*)y, n, D, period)));
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_colsum_parts(long long n) { long long g = (n + 255) / 256; if (g > 64) g = 64; if (g < 1) g = 1; return (int)g; }
int mb200_colsum(const void* x, float* part, void* out, int accumulate, long long n, int N, long long ld,
                 int dtype, void* stream) {
  if (n <= 0 || N <= 0) return MB200_OK;
  const int parts = mb200_colsum_parts(n);
  dim3 grid((N + 255) / 256, parts);
  DISPATCH_T(dtype, {
    colsum_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, part, n, N, ld);
    colsum_partials_kernel<T><<<COLSUM_GRID(N), 256, 0, (cudaStream_t)stream>>>(part, parts, N, (T*)out, accumulate);
  });
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_im2col(const void* px, int px_dtype, void* out, int out_dtype, int N, int C, int H, int W, int p,
                 int Kpad, void* stream) {
  if (N <= 0) return MB200_OK;
  const int gh = H / p, gw = W / p, K = C * p * p;
  if (Kpad < K) return -EINVAL;
  const long long total = (long long)N * gh * gw * Kpad;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = ew_grid(total);
  if (px_dtype == MB200_DTYPE_F32 && out_dtype == MB200_DTYPE_F32)
    im2col_kernel<float, float><<<g, 256, 0, st>>>((const float*)px, (float*)out, N, C, H, W, p, gh, gw, K, Kpad);
  else if (px_dtype == MB200_DTYPE_F32 && out_dtype == MB200_DTYPE_BF16)
    im2col_kernel<float, bf16><<<g, 256, 0, st>>>((const float*)px, (bf16*)out, N, C, H, W, p, gh, gw, K, Kpad);
  else if (px_dtype == MB200_DTYPE_BF16 && out_dtype == MB200_DTYPE_BF16)
    im2col_kernel<bf16, bf16><<<g, 256, 0, st>>>((const bf16*)px, (bf16*)out, N, C, H, W, p, gh, gw, K, Kpad);
  else if (px_dtype == MB200_DTYPE_BF16 && out_dtype == MB200_DTYPE_F32)
    im2col_kernel<bf16, float><<<g, 256, 0, st>>>((const bf16*)px, (float*)out, N, C, H, W, p, gh, gw, K, Kpad);
  else return -EINVAL;
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_rows_all_zero(const void* x, long long n_rows, long long row_elems, int* flags, int dtype, void* stream) {
  if (n_rows <= 0) return MB200_OK;
  DISPATCH_T(dtype, (rows_all_zero_kernel<T><<<(unsigned)n_rows, 256, 0, (cudaStream_t)stream>>>((const T*)x, row_elems, flags)));
  MB200_CHECK_LAUNCH(); return MB200_OK;
}
int mb200_cast(const void* x, int in_dtype, void* y, int out_dtype, long long n, void* stream) {
  if (n <= 0) return MB200_OK;
  cudaStream_t st = (cudaStream_t)stream; const int g = ew_grid(n);
  if (in_dtype == MB200_DTYPE_F32 && out_dtype == MB200_DTYPE_BF16) cast_kernel<float, bf16><<<g, 256, 0, st>>>((const float*)x, (bf16*)y, n);
  else if (in_dtype == MB200_DTYPE_BF16 && out_dtype == MB200_DTYPE_F32) cast_kernel<bf16, float><<<g, 256, 0, st>>>((const bf16*)x, (float*)y, n);
  else if (in_dtype == MB200_DTYPE_F32 && out_dtype == MB200_DTYPE_F32) cast_kernel<float, float><<<g, 256, 0, st>>>((const float*)x, (float*)y, n);
  else if (in_dtype == MB200_DTYPE_BF16 && out_dtype == MB200_DTYPE_BF16) cast_kernel<bf16, bf16><<<g, 256, 0, st>>>((const bf16*)x, (bf16*)y, n);
  else return -EINVAL;
  MB200_CHECK_LAUNCH(); return MB200_OK;
}


// px: uint8 [N,H,W,C] (channels_last = 1) or [N,C,H,W]; lut: fp32 [C][256] on device; out: [N,C,H,W] fp32 / bf16.  C <= 4.
int mb200_image_normalize_u8(const void* px, const float* lut, void* out, int out_dtype, int N, int C, int H, int W,
                             int channels_last, void* stream) {
  if (N <= 0) return MB200_OK;
  if (C < 1 || C > 4 || H <= 0 || W <= 0) return -EINVAL;
  const long long total = (long long)N * C * H * ((W + 7) / 8);
  const int g = ew_grid(total);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == MB200_DTYPE_F32)
    image_normalize_u8_kernel<float><<<g, 256, 0, st>>>((const unsigned char*)px, lut, (float*)out, N, C, H, W, channels_last);
  else if (out_dtype == MB200_DTYPE_BF16)
    image_normalize_u8_kernel<bf16><<<g, 256, 0, st>>>((const unsigned char*)px, lut, (bf16*)out, N, C, H, W, channels_last);
  else return -EINVAL;
  MB200_CHECK_LAUNCH(); return MB200_OK;
}


// One resampling pass over an [in_h, in_w, C] uint8 image: horizontal -> out [in_h, out_len, C], else -> [out_len, in_w, C].
// bounds: int32 [out_len][2] = (first tap, tap count); coef: int32 [out_len][ksize] 22-bit fixed point (device pointers).
int mb200_resize_u8_pass(const void* in, void* out, const int* bounds, const int* coef, int ksize, int in_h, int in_w,
                         int out_len, int C, int horizontal, void* stream) {
  if (in_h <= 0 || in_w <= 0 || out_len <= 0 || C <= 0 || ksize <= 0) return -EINVAL;
  const long long total = (long long)(horizontal ? in_h : out_len) * (horizontal ? out_len : in_w) * C;
  const int g = ew_grid(total);
  cudaStream_t st = (cudaStream_t)stream;
  if (horizontal)
    resize_pass_u8_kernel<true><<<g, 256, 0, st>>>((const unsigned char*)in, (unsigned char*)out, bounds, coef, ksize, in_h, in_w, out_len, C);
  else
    resize_pass_u8_kernel<false><<<g, 256, 0, st>>>((const unsigned char*)in, (unsigned char*)out, bounds, coef, ksize, in_h, in_w, out_len, C);
  MB200_CHECK_LAUNCH(); return MB200_OK;
}

}  // extern "C"

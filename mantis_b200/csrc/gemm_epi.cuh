// Epilogue shared by the tcgen05 GEMM kernels (gemm_sm100.cu, gemm_sm100_2cta.cu): one thread owns one output row and
// 32 consecutive accumulator columns read from TMEM.
//   bf16 mode : C = bf16( act(acc + bias) + addend )                       (forward / dgrad / bf16 wgrad)
//   fp32 mode : C32 = acc (+ C32 when accumulating)                          (wgrad into the fp32 main-gradient buffer:
//               what DeepSpeed's fp32 gradient accumulation does for the reference recipe,
//               mantis/train/zero_configs/zero3.json + scripts/train_mllava.sh:148 `--bf16 True`)
#pragma once
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace gemm_epi {

struct GemmEpi {
  void* C; long long ldc;               // bf16* or float* (c_f32); ldc in elements of that type
  const bf16* bias;
  const void* addend; long long ld_add; // same element type as C
  int act;                              // 0 none, 1 gelu(erf), 2 gelu(tanh), 3 quick_gelu
  int c_f32;                            // 1: C / addend are fp32
  // SwiGLU fused into the projections around it (llama/modeling_llama.py:182-184 `down(act(gate(x)) * up(x))`):
  //   mode 1 (up-projection forward):   acc = u.   C = u, C2 = silu(aux0 = gate) * u
  //   mode 2 (down-projection dgrad):   acc = d_act. C = d_gate, C2 = d_up from aux0 = gate, aux1 = up
  // -- the elementwise kernels (3 resp. 5 passes over [tokens, 14336]) disappear into epilogues that have time to spare.
  int mode;
  const bf16* aux0; const bf16* aux1; long long ld_aux;
  bf16* C2; long long ldc2;
  int tma_store;                        // output written (fp32 accumulate: reduce-added) by TMA from a swizzled smem tile
  unsigned long long pol_a, pol_b, pol_c;   // L2 eviction priorities of the A / B loads and the C stores (2-CTA kernel)
};

struct SwigluArgs { int mode; const bf16* aux0; const bf16* aux1; long long ld_aux; bf16* C2; long long ldc2; };

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// 8 consecutive bf16 <-> floats
__device__ __forceinline__ void ld8(const bf16* p, float* f) {
  const int4 v = *reinterpret_cast<const int4*>(p);
  const bf162* h = reinterpret_cast<const bf162*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}
__device__ __forceinline__ void st8(bf16* p, const float* f) {
  int4 v; bf162* h = reinterpret_cast<bf162*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
  *reinterpret_cast<int4*>(p) = v;
}

// The operands an epilogue chunk reads from global memory (fp32 / bf16 addend, or gate (+ up) of the SwiGLU modes), fetched one
// chunk AHEAD of their use: the epilogue warp issues chunk c+1's loads, then converts / stores chunk c, so the L2 round trip
// (microseconds while the TMA producer keeps L2 busy) is paid under useful work instead of once per chunk.
struct Prefetch {
  int4 a[8];          // fp32 addend: 8 x float4 | bf16 addend: a[0..3] | swiglu: gate a[0..3], up a[4..7]
  bool vec;           // the 32 columns are in range and every pointer / leading dimension allows 16-byte accesses
};

__device__ __forceinline__ bool chunk_vec(const GemmEpi& epi, int col0, int N) {
  if (col0 + 32 > N) return false;
  if (epi.mode)
    return !((epi.ldc | epi.ldc2 | epi.ld_aux) & 7) &&
           !((reinterpret_cast<uintptr_t>(epi.C) | reinterpret_cast<uintptr_t>(epi.C2) | reinterpret_cast<uintptr_t>(epi.aux0) |
              reinterpret_cast<uintptr_t>(epi.aux1)) & 15);
  if (epi.c_f32)
    return ((epi.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(epi.C) & 15) == 0) &&
           (!epi.addend || (((epi.ld_add & 3) == 0) && ((reinterpret_cast<uintptr_t>(epi.addend) & 15) == 0)));
  return ((epi.ldc & 7) == 0) && ((reinterpret_cast<uintptr_t>(epi.C) & 15) == 0) &&
         (!epi.addend || (((epi.ld_add & 7) == 0) && ((reinterpret_cast<uintptr_t>(epi.addend) & 15) == 0)));
}

__device__ __forceinline__ void prefetch32(const GemmEpi& epi, int row, int col0, int N, Prefetch& pf) {
  pf.vec = chunk_vec(epi, col0, N);
  if (!pf.vec) return;
  if (epi.mode) {
    const int4* g = reinterpret_cast<const int4*>(epi.aux0 + (size_t)row * epi.ld_aux + col0);
#pragma unroll
    for (int q = 0; q < 4; ++q) pf.a[q] = g[q];
    if (epi.aux1) {
      const int4* u = reinterpret_cast<const int4*>(epi.aux1 + (size_t)row * epi.ld_aux + col0);
#pragma unroll
      for (int q = 0; q < 4; ++q) pf.a[4 + q] = u[q];
    }
  } else if (epi.addend) {
    if (epi.c_f32) {
      const int4* a = reinterpret_cast<const int4*>(reinterpret_cast<const float*>(epi.addend) + (size_t)row * epi.ld_add + col0);
#pragma unroll
      for (int q = 0; q < 8; ++q) pf.a[q] = a[q];
    } else {
      const int4* a = reinterpret_cast<const int4*>(reinterpret_cast<const bf16*>(epi.addend) + (size_t)row * epi.ld_add + col0);
#pragma unroll
      for (int q = 0; q < 4; ++q) pf.a[q] = a[q];
    }
  }
}

__device__ __forceinline__ void unpack8(const int4& v, float* f) {
  const bf162* h = reinterpret_cast<const bf162*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}

// same arithmetic (and the same bf16 rounding points) as swiglu_fwd_kernel / swiglu_bwd_kernel in elementwise.cu
__device__ __forceinline__ void swiglu_one(int mode, float acc_raw, float g, float u, float& o1, float& o2) {
  const float acc = bf16_round(acc_raw);                 // the value the unfused path would have stored and re-read
  if (mode == 1) {                                       // acc = up:   o1 = up, o2 = silu(gate) * up
    o1 = acc;
    o2 = bf16_round(g / (1.f + __expf(-g))) * acc;
  } else {                                               // acc = d_act: o1 = d_gate, o2 = d_up
    const float sg = 1.f / (1.f + __expf(-g));
    o2 = acc * (g * sg);
    o1 = acc * u * (sg * (1.f + g * (1.f - sg)));
  }
}
__device__ __forceinline__ void swiglu_store32(const GemmEpi& epi, int row, int col0, int N, float (&v)[32], const Prefetch& pf) {
  bf16* c1 = reinterpret_cast<bf16*>(epi.C) + (size_t)row * epi.ldc + col0;
  bf16* c2 = epi.C2 + (size_t)row * epi.ldc2 + col0;
  const bf16* g0 = epi.aux0 + (size_t)row * epi.ld_aux + col0;
  const bf16* u0 = epi.aux1 ? epi.aux1 + (size_t)row * epi.ld_aux + col0 : nullptr;
  if (pf.vec) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {                        // 8 columns at a time keeps the live registers low
      float g[8], u[8], o1[8], o2[8];
      unpack8(pf.a[q], g);
      if (u0) unpack8(pf.a[4 + q], u);
#pragma unroll
      for (int j = 0; j < 8; ++j) swiglu_one(epi.mode, v[q * 8 + j], g[j], u0 ? u[j] : 0.f, o1[j], o2[j]);
      st8(c1 + q * 8, o1); st8(c2 + q * 8, o2);
    }
  } else {
    for (int j = 0; j < 32; ++j) {
      if (col0 + j < N) {
        float o1, o2;
        swiglu_one(epi.mode, v[j], __bfloat162float(g0[j]), u0 ? __bfloat162float(u0[j]) : 0.f, o1, o2);
        c1[j] = __float2bfloat16_rn(o1); c2[j] = __float2bfloat16_rn(o2);
      }
    }
  }
}

__device__ __forceinline__ float epi_act(float x, int kind) {
  if (kind == 1) return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
  if (kind == 2) { const float k = 0.79788456080286535588f; return 0.5f * x * (1.f + tanhf(k * (x + 0.044715f * x * x * x))); }
  if (kind == 3) return x / (1.f + __expf(-1.702f * x));
  return x;
}

// v[32] = raw accumulators of (row, col0 .. col0+31); pf = prefetch32() of the same chunk; caller guarantees row < M, col0 < N
__device__ __forceinline__ void store32(const GemmEpi& epi, int row, int col0, int N, float (&v)[32], const Prefetch& pf) {
  if (epi.mode) { swiglu_store32(epi, row, col0, N, v, pf); return; }
  if (epi.c_f32) {
    float* crow = reinterpret_cast<float*>(epi.C) + (size_t)row * epi.ldc;
    const float* arow = epi.addend ? reinterpret_cast<const float*>(epi.addend) + (size_t)row * epi.ld_add : nullptr;
    if (pf.vec) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float4 o = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        if (arow) { const float4 a = *reinterpret_cast<const float4*>(&pf.a[g]); o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
        *reinterpret_cast<float4*>(crow + col0 + 4 * g) = o;
      }
    } else {
      for (int j = 0; j < 32; ++j)
        if (col0 + j < N) crow[col0 + j] = v[j] + (arow ? arow[col0 + j] : 0.f);
    }
    return;
  }
  bf16* crow = reinterpret_cast<bf16*>(epi.C) + (size_t)row * epi.ldc;
  const bf16* arow = epi.addend ? reinterpret_cast<const bf16*>(epi.addend) + (size_t)row * epi.ld_add : nullptr;
  if (epi.bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j) if (col0 + j < N) v[j] += __bfloat162float(__ldg(epi.bias + col0 + j));
  }
  if (epi.act) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = epi_act(v[j], epi.act);
  }
  if (pf.vec) {
    if (arow) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float f[8];
        unpack8(pf.a[g], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[g * 8 + j] += f[j];
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) st8(crow + col0 + g * 8, v + g * 8);
  } else {
    for (int j = 0; j < 32; ++j) {
      if (col0 + j < N) {
        float x = v[j];
        if (arow) x += __bfloat162float(arow[col0 + j]);
        crow[col0 + j] = __float2bfloat16_rn(x);
      }
    }
  }
}

// Output through TMA: the warp converts its 32 rows, parks them in its private 4 KB shared-memory tile in the 128B-swizzled
// layout (16-byte chunk j of row r at r * 128 + ((j ^ (r & 7)) << 4): conflict-free 16-byte stores), and ONE thread hands the
// tile to the TMA unit, which writes full 128-byte lines and clips rows / columns beyond M / N by itself.
//   bf16 C (+ bias / activation / bf16 addend, which is prefetched per thread one chunk ahead): 64-column tiles, TMA store
//   fp32 C (wgrad into the main gradient): 32-column tiles; accumulation is a TMA REDUCE-ADD -- the previous value of the
//   gradient is added inside the L2 and never read by the SM
// Replaces 16-byte stores that each touch half a sector of 32 different lines (and, for fp32, the addend loads).
template <int NCHUNK>
__device__ __forceinline__ void epilogue_rows_tma(const GemmEpi& epi, const CUtensorMap* tmC, uint8_t* stage, uint32_t taddr,
                                                  int row, bool row_ok, int n0, int N, int lane) {
  const int row0 = row - lane;
  const bool f32 = epi.c_f32 != 0;
  const bool pre = !f32 && epi.addend != nullptr;            // bf16 addend: per-thread loads, one chunk ahead
  Prefetch pf[2];
  pf[0].vec = false; pf[1].vec = false;
  if (pre && row_ok && n0 < N) prefetch32(epi, row, n0, N, pf[0]);
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    const int col0 = n0 + c * 32;
    if (col0 < N) {                                          // uniform across the warp
      uint32_t r[32];
      sm100::tmem_ld_32x32b_x32(taddr + c * 32, r);
      if (pre && c + 1 < NCHUNK && row_ok && col0 + 32 < N) prefetch32(epi, row, col0 + 32, N, pf[(c + 1) & 1]);
      if (f32 || (c & 1) == 0) {
        if (lane == 0) sm100::tma_store_wait_read0();        // the previous bulk operation has finished READING the tile
        __syncwarp();
      }
      sm100::tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (f32) {
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<float4*>(stage + lane * 128 + ((g ^ (lane & 7)) << 4)) =
              make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
      } else {
        if (epi.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (col0 + j < N) v[j] += __bfloat162float(__ldg(epi.bias + col0 + j));
        }
        if (epi.act) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = epi_act(v[j], epi.act);
        }
        if (pre) {
          if (pf[c & 1].vec) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float f[8];
              unpack8(pf[c & 1].a[g], f);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[g * 8 + j] += f[j];
            }
          } else if (row_ok) {                               // column tail: scalar addend reads
            const bf16* arow = reinterpret_cast<const bf16*>(epi.addend) + (size_t)row * epi.ld_add;
            for (int j = 0; j < 32; ++j) if (col0 + j < N) v[j] += __bfloat162float(arow[col0 + j]);
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          int4 o; bf162* oh = reinterpret_cast<bf162*>(&o);
#pragma unroll
          for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
          const int chunk = (c & 1) * 4 + g;
          *reinterpret_cast<int4*>(stage + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = o;
        }
      }
      const bool flush = f32 || (c & 1) == 1 || c + 1 == NCHUNK || col0 + 32 >= N;
      if (flush) {
        sm100::fence_proxy_async();                          // generic-proxy writes -> visible to the async (TMA) proxy
        __syncwarp();
        if (lane == 0) {
          const int tc0 = f32 ? col0 : col0 - (c & 1) * 32;
          if (f32 && epi.addend) sm100::tma_reduce_add_2d_hint(tmC, stage, tc0, row0, epi.pol_c);
          else sm100::tma_store_2d_hint(tmC, stage, tc0, row0, epi.pol_c);
          sm100::tma_store_commit();
        }
      }
    }
  }
}

// One epilogue warp's share of an accumulator tile: NCHUNK chunks of 32 columns starting at TMEM column `tcol` / global column
// `n0`, for the warp's 32 rows (TMEM lanes 32 * (warp % 4) ...).  Software-pipelined: chunk c+1's global operands are in
// flight while chunk c is converted and stored.
template <int NCHUNK>
__device__ __forceinline__ void epilogue_rows(const GemmEpi& epi, uint32_t taddr, int row, bool row_ok, int n0, int N) {
  Prefetch pf[2];
  pf[0].vec = false; pf[1].vec = false;
  if (row_ok && n0 < N) prefetch32(epi, row, n0, N, pf[0]);
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    uint32_t r[32];
    sm100::tmem_ld_32x32b_x32(taddr + c * 32, r);
    const int col0 = n0 + c * 32;
    if (c + 1 < NCHUNK && row_ok && col0 + 32 < N) prefetch32(epi, row, col0 + 32, N, pf[(c + 1) & 1]);
    sm100::tmem_ld_wait();
    if (row_ok && col0 < N) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      store32(epi, row, col0, N, v, pf[c & 1]);
    }
  }
}

}  // namespace gemm_epi

// Epilogue shared by the tcgen05 GEMM kernels (gemm_sm100.cu, gemm_sm100_2cta.cu): one thread owns one output row and
// 32 consecutive accumulator columns read from TMEM.
//   bf16 mode : C = bf16( act(acc + bias) + addend )                       (forward / dgrad / bf16 wgrad)
//   fp32 mode : C32 = acc (+ C32 when accumulating)                          (wgrad into the fp32 main-gradient buffer:
//               what DeepSpeed's fp32 gradient accumulation does for the reference recipe,
//               mantis/train/zero_configs/zero3.json + scripts/train_mllava.sh:148 `--bf16 True`)
#pragma once
#include "common.cuh"

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

// same arithmetic (and the same bf16 rounding points) as swiglu_fwd_kernel / swiglu_bwd_kernel in elementwise.cu
__device__ __forceinline__ void swiglu_store32(const GemmEpi& epi, int row, int col0, int N, float (&v)[32]) {
  bf16* c1 = reinterpret_cast<bf16*>(epi.C) + (size_t)row * epi.ldc + col0;
  bf16* c2 = epi.C2 + (size_t)row * epi.ldc2 + col0;
  const bf16* g0 = epi.aux0 + (size_t)row * epi.ld_aux + col0;
  const bf16* u0 = epi.aux1 ? epi.aux1 + (size_t)row * epi.ld_aux + col0 : nullptr;
  const bool vec = (col0 + 32 <= N) && !((epi.ldc | epi.ldc2 | epi.ld_aux) & 7) &&
                   !((reinterpret_cast<uintptr_t>(epi.C) | reinterpret_cast<uintptr_t>(epi.C2) | reinterpret_cast<uintptr_t>(epi.aux0) |
                      reinterpret_cast<uintptr_t>(epi.aux1)) & 15);
  float g[32], u[32];
  if (vec) {                                            // every load of the chunk before the first store (see store32)
#pragma unroll
    for (int q = 0; q < 4; ++q) ld8(g0 + q * 8, g + q * 8);
    if (u0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) ld8(u0 + q * 8, u + q * 8);
    }
  } else {
    for (int j = 0; j < 32; ++j) {
      const bool ok = col0 + j < N;
      g[j] = ok ? __bfloat162float(g0[j]) : 0.f;
      u[j] = (ok && u0) ? __bfloat162float(u0[j]) : 0.f;
    }
  }
  float o1[32], o2[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float acc = bf16_round(v[j]);                  // the value the unfused path would have stored and re-read
    if (epi.mode == 1) {                                 // acc = up
      o1[j] = acc;
      o2[j] = bf16_round(g[j] / (1.f + __expf(-g[j]))) * acc;
    } else {                                             // acc = d_act
      const float sg = 1.f / (1.f + __expf(-g[j]));
      o2[j] = acc * (g[j] * sg);                                          // d_up
      o1[j] = acc * u[j] * (sg * (1.f + g[j] * (1.f - sg)));              // d_gate
    }
  }
  if (vec) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { st8(c1 + q * 8, o1 + q * 8); st8(c2 + q * 8, o2 + q * 8); }
  } else {
    for (int j = 0; j < 32; ++j)
      if (col0 + j < N) { c1[j] = __float2bfloat16_rn(o1[j]); c2[j] = __float2bfloat16_rn(o2[j]); }
  }
}

__device__ __forceinline__ float epi_act(float x, int kind) {
  if (kind == 1) return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
  if (kind == 2) { const float k = 0.79788456080286535588f; return 0.5f * x * (1.f + tanhf(k * (x + 0.044715f * x * x * x))); }
  if (kind == 3) return x / (1.f + __expf(-1.702f * x));
  return x;
}

// v[32] = raw accumulators of (row, col0 .. col0+31); caller guarantees row < M and col0 < N
__device__ __forceinline__ void store32(const GemmEpi& epi, int row, int col0, int N, float (&v)[32]) {
  if (epi.mode) { swiglu_store32(epi, row, col0, N, v); return; }
  if (epi.c_f32) {
    float* crow = reinterpret_cast<float*>(epi.C) + (size_t)row * epi.ldc;
    const float* arow = epi.addend ? reinterpret_cast<const float*>(epi.addend) + (size_t)row * epi.ld_add : nullptr;
    const bool vec = ((epi.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(epi.C) & 15) == 0) && (col0 + 32 <= N) &&
                     (!arow || (((epi.ld_add & 3) == 0) && ((reinterpret_cast<uintptr_t>(epi.addend) & 15) == 0)));
    if (vec) {
      // all loads of the chunk are issued before the first store: destination and addend may be the same buffer (gradient
      // accumulation), so the compiler cannot move a load above an earlier store on its own -- eight serialised L2 round trips
      // per chunk would make the epilogue longer than the tile's mainloop
      if (arow) {
        float4 a[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) a[g] = *reinterpret_cast<const float4*>(arow + col0 + 4 * g);
#pragma unroll
        for (int g = 0; g < 8; ++g) { v[4 * g] += a[g].x; v[4 * g + 1] += a[g].y; v[4 * g + 2] += a[g].z; v[4 * g + 3] += a[g].w; }
      }
#pragma unroll
      for (int g = 0; g < 8; ++g)
        *reinterpret_cast<float4*>(crow + col0 + 4 * g) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
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
  const bool vec_ok = ((epi.ldc & 7) == 0) && ((reinterpret_cast<uintptr_t>(epi.C) & 15) == 0) &&
                      (!arow || (((epi.ld_add & 7) == 0) && ((reinterpret_cast<uintptr_t>(epi.addend) & 15) == 0)));
  if (vec_ok && col0 + 32 <= N) {
    if (arow) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        int4 a4 = *reinterpret_cast<const int4*>(arow + col0 + g * 8);
        const bf162* ah = reinterpret_cast<const bf162*>(&a4);
#pragma unroll
        for (int j = 0; j < 4; ++j) { float2 f = __bfloat1622float2(ah[j]); v[g * 8 + 2 * j] += f.x; v[g * 8 + 2 * j + 1] += f.y; }
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      int4 o4; bf162* oh = reinterpret_cast<bf162*>(&o4);
#pragma unroll
      for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
      *reinterpret_cast<int4*>(crow + col0 + g * 8) = o4;
    }
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

}  // namespace gemm_epi

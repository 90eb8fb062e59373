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
};

__device__ __forceinline__ float epi_act(float x, int kind) {
  if (kind == 1) return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
  if (kind == 2) { const float k = 0.79788456080286535588f; return 0.5f * x * (1.f + tanhf(k * (x + 0.044715f * x * x * x))); }
  if (kind == 3) return x / (1.f + __expf(-1.702f * x));
  return x;
}

// v[32] = raw accumulators of (row, col0 .. col0+31); caller guarantees row < M and col0 < N
__device__ __forceinline__ void store32(const GemmEpi& epi, int row, int col0, int N, float (&v)[32]) {
  if (epi.c_f32) {
    float* crow = reinterpret_cast<float*>(epi.C) + (size_t)row * epi.ldc;
    const float* arow = epi.addend ? reinterpret_cast<const float*>(epi.addend) + (size_t)row * epi.ld_add : nullptr;
    const bool vec = ((epi.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(epi.C) & 15) == 0) && (col0 + 32 <= N) &&
                     (!arow || (((epi.ld_add & 3) == 0) && ((reinterpret_cast<uintptr_t>(epi.addend) & 15) == 0)));
    if (vec) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float4 o = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        if (arow) { const float4 a = *reinterpret_cast<const float4*>(arow + col0 + 4 * g); o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
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

// Decode-time (q_len == 1) kernels of generate(): HBM-bound by construction.
//   skinny GEMM  : C[M,N] = X[M,K] W[N,K]^T for M <= 16 rows (one row per sequence of the decode batch).
//                  Every weight byte is read exactly once with 128-bit loads; X is tiny and stays in L1.
//                  (the tcgen05 tile kernel would light up only N/256 CTAs for these shapes)
//   split-KV attention : grid (splits, kv_heads, batch); each CTA streams a slice of the cached keys/values once for all
//                  query heads of its GQA group, then a combine kernel merges the partial (max, sum, out) triples.
//   KV append    : writes the new token's K/V rows into the token-major cache at a per-sequence position.
// These replace DynamicCache's torch.cat growth + SDPA at q_len 1 in the reference stack
// (hf: llama/modeling_llama.py:269-270; mantis/models/mllava/modeling_llava.py:477-519).
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "../../include/mantis_b200.h"
#include <stdlib.h>

namespace {
using mb::Cvt;
using namespace sm100;

// ------------------------------------------------------------------ skinny GEMM
// Up to three weight matrices that share the same input X are served by ONE launch (q/k/v, or gate/up): `seg` picks the
// matrix from the output-row index.  mode 1 (two segments = gate, up): writes silu(gate) * up (SwiGLU fused).
struct SkinnySeg { const bf16* W[3]; bf16* C[3]; int N[3]; long long ldc[3]; int nseg; int mode; };

template <int MT, int RPW, int KU>
__global__ void __launch_bounds__(256)
skinny_gemm_kernel(const bf16* __restrict__ X, SkinnySeg sg, const bf16* __restrict__ bias,
                   const bf16* __restrict__ addend, int M, int K, long long ldx, long long ldw, long long ld_add) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Ntot = (sg.mode == 1) ? sg.N[0] : (sg.N[0] + sg.N[1] + sg.N[2]);
  const int n0 = (blockIdx.x * 8 + warp) * RPW;
  if (n0 >= Ntot) return;
  constexpr int NR = RPW;            // output rows per warp; in SwiGLU mode each output row reads a gate row and an up row
  float acc[NR][MT], acc2[NR][MT];
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int m = 0; m < MT; ++m) { acc[r][m] = 0.f; acc2[r][m] = 0.f; }
  const bf16* wrow[NR]; const bf16* wrow2[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    int n = n0 + r; if (n >= Ntot) n = Ntot - 1;
    if (sg.mode == 1) { wrow[r] = sg.W[0] + (size_t)n * ldw; wrow2[r] = sg.W[1] + (size_t)n * ldw; }
    else {
      int seg = 0, nn = n;
      if (nn >= sg.N[0]) { nn -= sg.N[0]; seg = 1; if (nn >= sg.N[1]) { nn -= sg.N[1]; seg = 2; } }
      wrow[r] = sg.W[seg] + (size_t)nn * ldw; wrow2[r] = wrow[r];
    }
  }
  for (int k = lane * 8; k < K; k += 256 * KU) {
    int4 wv[KU][NR], wv2[KU][NR];
#pragma unroll
    for (int u = 0; u < KU; ++u) {
      const int kk = k + u * 256;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        wv[u][r] = (kk < K) ? mb::ld_stream(reinterpret_cast<const int4*>(wrow[r] + kk)) : make_int4(0, 0, 0, 0);
        if (sg.mode == 1) wv2[u][r] = (kk < K) ? mb::ld_stream(reinterpret_cast<const int4*>(wrow2[r] + kk)) : make_int4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < KU; ++u) {
      const int kk = k + u * 256;
      if (kk >= K) break;
      float xv[MT][8];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        if (m < M) mb::Vec8<bf16>::load(X + (size_t)m * ldx + kk, xv[m]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) xv[m][j] = 0.f;
        }
      }
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const bf162* wh = reinterpret_cast<const bf162*>(&wv[u][r]);
        float wf[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) { float2 t = __bfloat1622float2(wh[j]); wf[2 * j] = t.x; wf[2 * j + 1] = t.y; }
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[r][m] = fmaf(wf[j], xv[m][j], acc[r][m]);
        if (sg.mode == 1) {
          const bf162* wh2 = reinterpret_cast<const bf162*>(&wv2[u][r]);
#pragma unroll
          for (int j = 0; j < 4; ++j) { float2 t = __bfloat1622float2(wh2[j]); wf[2 * j] = t.x; wf[2 * j + 1] = t.y; }
#pragma unroll
          for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc2[r][m] = fmaf(wf[j], xv[m][j], acc2[r][m]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int m = 0; m < MT; ++m) { acc[r][m] = mb::warp_sum(acc[r][m]); if (sg.mode == 1) acc2[r][m] = mb::warp_sum(acc2[r][m]); }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int n = n0 + r;
      if (n >= Ntot) continue;
      int seg = 0, nn = n;
      if (sg.mode != 1 && nn >= sg.N[0]) { nn -= sg.N[0]; seg = 1; if (nn >= sg.N[1]) { nn -= sg.N[1]; seg = 2; } }
      const float b = (bias && seg == 0) ? __bfloat162float(bias[nn]) : 0.f;
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        if (m < M) {
          float v = acc[r][m] + b;
          if (sg.mode == 1) {
            // reference order: act_fn(gate) rounded to bf16, then * up (hf: llama/modeling_llama.py:182-184)
            const float gq = __bfloat162float(__float2bfloat16_rn(v));
            const float uq = __bfloat162float(__float2bfloat16_rn(acc2[r][m]));
            const float sl = __bfloat162float(__float2bfloat16_rn(gq / (1.f + __expf(-gq))));
            v = sl * uq;
          }
          if (addend && seg == 0) v += __bfloat162float(addend[(size_t)m * ld_add + nn]);
          sg.C[seg][(size_t)m * sg.ldc[seg] + nn] = __float2bfloat16_rn(v);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ skinny GEMM, 8 < M <= 16: mma.sync m16n8k16
// At M = 16 the SIMT kernel above is FMA-bound (16 FMAs per weight element); here one warp owns 8 output rows and the
// batch rows form the M = 16 side of an HMMA tile, so the kernel is HBM-bound again.  K is consumed 32 elements per lane
// group at a time with 16-byte loads: the dot product does not care about the order of k, so lane `tid` simply feeds
// memory elements [8 tid, 8 tid + 8) of every 32-chunk to the fragment slots of two consecutive k16 steps -- for A (X)
// and B (W) alike.  (tcgen05 is not used here: a 128-row MMA tile with 16 useful rows buys nothing for an HBM-bound op,
// and mma.sync needs no TMEM / TMA setup per launch.)
__device__ __forceinline__ void mma_16816(float* c, const uint32_t a0, const uint32_t a1, const uint32_t a2, const uint32_t a3,
                                          const uint32_t b0, const uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int KU>
__global__ void __launch_bounds__(128)
skinny_mma_kernel(const bf16* __restrict__ X, SkinnySeg sg, const bf16* __restrict__ bias, const bf16* __restrict__ addend,
                  int M, int K, long long ldx, long long ldw, long long ld_add) {
  __shared__ float red[3][32][8];           // partial accumulators of warps 1..3 (K is split across the 4 warps of a CTA)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gid = lane >> 2, tid = lane & 3;
  const int Ntot = (sg.mode == 1) ? sg.N[0] : (sg.N[0] + sg.N[1] + sg.N[2]);
  const int n0 = blockIdx.x * 8;
  int n = n0 + gid; if (n >= Ntot) n = Ntot - 1;
  const bf16 *w0, *w1;
  if (sg.mode == 1) { w0 = sg.W[0] + (size_t)n * ldw; w1 = sg.W[1] + (size_t)n * ldw; }
  else {
    int seg = 0, nn = n;
    if (nn >= sg.N[0]) { nn -= sg.N[0]; seg = 1; if (nn >= sg.N[1]) { nn -= sg.N[1]; seg = 2; } }
    w0 = sg.W[seg] + (size_t)nn * ldw; w1 = w0;
  }
  const bf16* xlo = X + (size_t)gid * ldx;            // batch row gid
  const bf16* xhi = X + (size_t)(gid + 8) * ldx;      // batch row gid + 8
  const bool lo_ok = gid < M, hi_ok = gid + 8 < M;
  float c[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
  const int4 z = make_int4(0, 0, 0, 0);
  for (int k = (warp * KU) * 32 + tid * 8; k < K; k += 4 * 32 * KU) {
    int4 wv[KU], wv2[KU], xl[KU], xh[KU];
#pragma unroll
    for (int u = 0; u < KU; ++u) {
      const int kk = k + u * 32;
      const bool ok = kk < K;
      wv[u] = ok ? mb::ld_stream(reinterpret_cast<const int4*>(w0 + kk)) : z;
      if (sg.mode == 1) wv2[u] = ok ? mb::ld_stream(reinterpret_cast<const int4*>(w1 + kk)) : z;
      xl[u] = (ok && lo_ok) ? *reinterpret_cast<const int4*>(xlo + kk) : z;
      xh[u] = (ok && hi_ok) ? *reinterpret_cast<const int4*>(xhi + kk) : z;
    }
#pragma unroll
    for (int u = 0; u < KU; ++u) {
      mma_16816(c, xl[u].x, xh[u].x, xl[u].y, xh[u].y, wv[u].x, wv[u].y);
      mma_16816(c, xl[u].z, xh[u].z, xl[u].w, xh[u].w, wv[u].z, wv[u].w);
      if (sg.mode == 1) {
        mma_16816(c2, xl[u].x, xh[u].x, xl[u].y, xh[u].y, wv2[u].x, wv2[u].y);
        mma_16816(c2, xl[u].z, xh[u].z, xl[u].w, xh[u].w, wv2[u].z, wv2[u].w);
      }
    }
  }
  if (warp > 0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { red[warp - 1][lane][e] = c[e]; red[warp - 1][lane][4 + e] = c2[e]; }
  }
  __syncthreads();
  if (warp > 0) return;
#pragma unroll
  for (int w = 0; w < 3; ++w)
#pragma unroll
    for (int e = 0; e < 4; ++e) { c[e] += red[w][lane][e]; c2[e] += red[w][lane][4 + e]; }
  // c[0], c[1]: (row gid, cols n0 + 2 tid, +1) ; c[2], c[3]: (row gid + 8, same cols)
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int m = gid + ((e >> 1) ? 8 : 0);
    const int nn_abs = n0 + tid * 2 + (e & 1);
    if (m >= M || nn_abs >= Ntot) continue;
    int seg = 0, nn = nn_abs;
    if (sg.mode != 1 && nn >= sg.N[0]) { nn -= sg.N[0]; seg = 1; if (nn >= sg.N[1]) { nn -= sg.N[1]; seg = 2; } }
    float v = c[e] + ((bias && seg == 0) ? __bfloat162float(bias[nn]) : 0.f);
    if (sg.mode == 1) {
      const float gq = __bfloat162float(__float2bfloat16_rn(v));
      const float uq = __bfloat162float(__float2bfloat16_rn(c2[e]));
      const float sl = __bfloat162float(__float2bfloat16_rn(gq / (1.f + __expf(-gq))));
      v = sl * uq;
    }
    if (addend && seg == 0) v += __bfloat162float(addend[(size_t)m * ld_add + nn]);
    sg.C[seg][(size_t)m * sg.ldc[seg] + nn] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------ skinny GEMM, TMA-ring variants (the default)
// The register-staged kernels above keep only ~64 KB per SM in flight during their load phases and stall in between
// (ncu: 25 % of DRAM peak, 75 % of issue slots waiting on L1TEX).  Here a producer warp streams weights through a
// shared-memory ring with 1-D bulk copies (cp.async.bulk + mbarrier tx counts, no tensor maps), so ~64 KB per CTA x 3 CTAs
// per SM stay in flight regardless of register pressure, and the stream starts BEFORE griddepcontrol.wait: under
// programmatic dependent launch it overlaps the previous kernel's tail.  CTAs are persistent (a balanced number of 8-row
// groups each), so barriers are set up once and the ring never drains between groups.
//   skinny_rows_kernel (M <= 4): a stage is ONE weight row (up to 4096 elements = one 8 KB bulk copy -- the first ring
//     version moved 1 KB per copy and was bound by the per-SM TMA issue rate); the 8 warps split the row along k, keep
//     8 x M partial sums in registers and reduce across warps once per group.
//   skinny_ring_kernel (M <= 16): a stage is 8 rows x 1024 k (padded rows) feeding mma.sync m16n8k16; the warps split the
//     stage along k and keep their X fragments in registers across 4 consecutive 8-row groups.
constexpr int SK_THREADS = 288;              // 8 consumer warps + 1 producer warp
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ const bf16* skinny_row_ptr(const SkinnySeg& sg, int n, int which, long long ldw) {
  if (sg.mode == 1) return (which ? sg.W[1] : sg.W[0]) + (size_t)n * ldw;
  if (n < sg.N[0]) return sg.W[0] + (size_t)n * ldw;
  n -= sg.N[0];
  if (n < sg.N[1]) return sg.W[1] + (size_t)n * ldw;
  return sg.W[2] + (size_t)(n - sg.N[1]) * ldw;
}
// epilogue for output (m, n): bias, SwiGLU, residual addend, store
__device__ __forceinline__ void skinny_store(const SkinnySeg& sg, int mode, int m, int n, float v, float v2, const bf16* bias,
                                             const bf16* addend, long long ld_add) {
  int seg = 0, nn = n;
  if (mode != 1 && nn >= sg.N[0]) { nn -= sg.N[0]; seg = 1; if (nn >= sg.N[1]) { nn -= sg.N[1]; seg = 2; } }
  bf16* Cs = (seg == 0) ? sg.C[0] : (seg == 1 ? sg.C[1] : sg.C[2]);
  const long long ldc = (seg == 0) ? sg.ldc[0] : (seg == 1 ? sg.ldc[1] : sg.ldc[2]);
  if (bias && seg == 0) v += __bfloat162float(bias[nn]);
  if (mode == 1) {
    // reference order: act_fn(gate) rounded to bf16, then * up (hf: llama/modeling_llama.py:182-184)
    const float gq = __bfloat162float(__float2bfloat16_rn(v));
    const float uq = __bfloat162float(__float2bfloat16_rn(v2));
    const float sl = __bfloat162float(__float2bfloat16_rn(gq / (1.f + __expf(-gq))));
    v = sl * uq;
  }
  if (addend && seg == 0) v += __bfloat162float(addend[(size_t)m * ld_add + nn]);
  Cs[(size_t)m * ldc + nn] = __float2bfloat16_rn(v);
}

constexpr int RS_K = 4096;                   // elements per row stage (8 KB)
constexpr int RS_NS = 8;                     // stages in the ring (64 KB)
constexpr int RS_SMEM = RS_NS * RS_K * 2 + 2 * RS_NS * 8 + 8 * 8 * 4 * 2 * 4;    // ring + barriers + reduction scratch

template <int MT, int MODE>
__global__ void __launch_bounds__(SK_THREADS)
skinny_rows_kernel(const bf16* __restrict__ X, SkinnySeg sg, const bf16* __restrict__ bias, const bf16* __restrict__ addend,
                   int M, int K, long long ldx, long long ldw, long long ld_add, int ngroups,
                   const bf16* __restrict__ gamma, float eps) {
  extern __shared__ __align__(128) unsigned char ring[];
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + RS_NS * RS_K * 2);
  uint64_t* empty = full + RS_NS;
  float* red = reinterpret_cast<float*>(empty + RS_NS);          // [8 warps][8 rows][MT][2]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Ntot = (MODE == 1) ? sg.N[0] : (sg.N[0] + sg.N[1] + sg.N[2]);
  const int nkc = (K + RS_K - 1) / RS_K;
  constexpr int NW = MODE ? 2 : 1;           // matrices per output row (gate, up)
  if (threadIdx.x == 0) {
    for (int i = 0; i < RS_NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
    fence_barrier_init();
  }
  __syncthreads();
  mb::pdl_trigger();
  if (warp == 8) {
    // ---------------- producer (one lane): weights do not depend on the previous kernel -> no pdl_wait
    if (lane != 0) return;
    uint32_t it = 0;
    for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
      for (int kc = 0; kc < nkc; ++kc) {
        const uint32_t bytes = (uint32_t)min(RS_K, K - kc * RS_K) * 2;
        for (int r = 0; r < 8; ++r) {
          int n = g * 8 + r; if (n >= Ntot) n = Ntot - 1;
#pragma unroll
          for (int wh = 0; wh < NW; ++wh, ++it) {
            const uint32_t s = it % RS_NS;
            if (it >= RS_NS) mbar_wait(&empty[s], ((it / RS_NS) - 1) & 1);
            mbar_arrive_expect_tx(&full[s], bytes);
            bulk_g2s(ring + s * (RS_K * 2), skinny_row_ptr(sg, n, wh, ldw) + (size_t)kc * RS_K, bytes, &full[s]);
          }
        }
      }
    }
    return;
  }
  // ---------------- consumers: warp w owns elements [512 w, 512 w + 512) of every row stage
  mb::pdl_wait();                       // X / addend are the previous kernel's outputs
  // Optional fused RMSNorm of X (gamma != null; K = the full hidden size): every CTA recomputes the M row norms from L2
  // (M x K x 2 bytes, while the producer already streams weights) and normalises X on the fly with the standalone
  // kernel's rounding order -- bf16(gamma * bf16(x * rstd)) -- which saves a launch and a round trip of the activations.
  float rstd[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) rstd[m] = 1.f;
  if (gamma) {
    float ss[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) ss[m] = 0.f;
    for (int k = threadIdx.x * 8; k < K; k += 256 * 8) {
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        if (m < M) {
          const int4 xr = __ldg(reinterpret_cast<const int4*>(X + (size_t)m * ldx + k));
          const bf162* xh = reinterpret_cast<const bf162*>(&xr);
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(xh[j]); ss[m] += t.x * t.x + t.y * t.y; }
        }
      }
    }
#pragma unroll
    for (int m = 0; m < MT; ++m) { ss[m] = mb::warp_sum(ss[m]); if (lane == 0) red[warp * 4 + m] = ss[m]; }
    consumer_sync();
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += red[w * 4 + m];
      rstd[m] = rsqrtf(tot / (float)K + eps);
    }
    consumer_sync();
  }
  uint32_t it = 0;
  for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
    float acc[8][MT], acc2[8][MT];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int m = 0; m < MT; ++m) { acc[r][m] = 0.f; acc2[r][m] = 0.f; }
    for (int kc = 0; kc < nkc; ++kc) {
      const int kbase = kc * RS_K, klen = min(RS_K, K - kbase);
      float xf[2][MT][8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int kk = warp * 512 + u * 256 + lane * 8;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const int4 xr = (kk < klen && m < M) ? __ldg(reinterpret_cast<const int4*>(X + (size_t)m * ldx + kbase + kk))
                                               : make_int4(0, 0, 0, 0);
          const bf162* xh = reinterpret_cast<const bf162*>(&xr);
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(xh[j]); xf[u][m][2 * j] = t.x; xf[u][m][2 * j + 1] = t.y; }
        }
        if (gamma && kk < klen) {
          const int4 gr = __ldg(reinterpret_cast<const int4*>(gamma + kbase + kk));
          const bf162* gh = reinterpret_cast<const bf162*>(&gr);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 gg = __bfloat1622float2(gh[j]);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              xf[u][m][2 * j] = mb::rnd<bf16>(gg.x * mb::rnd<bf16>(xf[u][m][2 * j] * rstd[m]));
              xf[u][m][2 * j + 1] = mb::rnd<bf16>(gg.y * mb::rnd<bf16>(xf[u][m][2 * j + 1] * rstd[m]));
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int wh = 0; wh < NW; ++wh, ++it) {
          const uint32_t s = it % RS_NS;
          mbar_wait(&full[s], (it / RS_NS) & 1);
          const unsigned char* st = ring + s * (RS_K * 2) + warp * 1024 + lane * 16;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int kk = warp * 512 + u * 256 + lane * 8;
            if (kk < klen) {
              const int4 wv = *reinterpret_cast<const int4*>(st + u * 512);
              const bf162* wh2 = reinterpret_cast<const bf162*>(&wv);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 t = __bfloat1622float2(wh2[j]);
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                  if (wh == 0) acc[r][m] = fmaf(t.y, xf[u][m][2 * j + 1], fmaf(t.x, xf[u][m][2 * j], acc[r][m]));
                  else acc2[r][m] = fmaf(t.y, xf[u][m][2 * j + 1], fmaf(t.x, xf[u][m][2 * j], acc2[r][m]));
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[s]);
        }
      }
    }
    // cross-lane, then cross-warp reduction; thread t < 8 * MT finishes output (row t / MT, batch t % MT)
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const float a = mb::warp_sum(acc[r][m]);
        const float a2 = MODE ? mb::warp_sum(acc2[r][m]) : 0.f;
        if (lane == 0) { red[((warp * 8 + r) * 4 + m) * 2] = a; red[((warp * 8 + r) * 4 + m) * 2 + 1] = a2; }
      }
    consumer_sync();
    if (threadIdx.x < 8 * MT) {
      const int r = threadIdx.x / MT, m = threadIdx.x % MT;
      float v = 0.f, v2 = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) { v += red[((w * 8 + r) * 4 + m) * 2]; v2 += red[((w * 8 + r) * 4 + m) * 2 + 1]; }
      const int n = g * 8 + r;
      if (n < Ntot && m < M) skinny_store(sg, MODE, m, n, v, v2, bias, addend, ld_add);
    }
    consumer_sync();                    // red is reused by the next group
  }
}

constexpr int SK_KC = 1024;                  // k elements per mma ring stage (2 KB bulk copies)
constexpr int SK_ROWB = SK_KC * 2 + 64;      // padded smem row: 528 words = 16 mod 32 -> conflict-free 16-byte fragment reads
constexpr int SK_NS = 6;
constexpr int SK_RG = 4;                     // 8-row groups per X-fragment load: X is re-read from L2 once per 32 weight rows
constexpr int SK_STAGE_B = 8 * SK_ROWB;      // (with one group per load the X traffic was 2x the weight traffic: L2-bound)
constexpr int SK_SMEM = SK_NS * SK_STAGE_B + 2 * SK_NS * 8;

template <int MODE>
__global__ void __launch_bounds__(SK_THREADS, 2)
skinny_ring_kernel(const bf16* __restrict__ X, SkinnySeg sg, const bf16* __restrict__ bias, const bf16* __restrict__ addend,
                   int M, int K, long long ldx, long long ldw, long long ld_add, int nsuper) {
  extern __shared__ __align__(128) unsigned char ring[];
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + SK_NS * SK_STAGE_B);
  uint64_t* empty = full + SK_NS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Ntot = (MODE == 1) ? sg.N[0] : (sg.N[0] + sg.N[1] + sg.N[2]);
  const int nchunks = (K + SK_KC - 1) / SK_KC;
  constexpr int NW = MODE ? 2 : 1;
  if (threadIdx.x == 0) {
    for (int i = 0; i < SK_NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
    fence_barrier_init();
  }
  __syncthreads();
  mb::pdl_trigger();
  if (warp == 8) {
    // ---------------- producer: lanes 0..7 each own one row of the stage
    uint32_t it = 0;
    for (int su = blockIdx.x; su < nsuper; su += gridDim.x) {
      for (int c = 0; c < nchunks; ++c) {
        const uint32_t bytes = (uint32_t)min(SK_KC, K - c * SK_KC) * 2;
        for (int rg = 0; rg < SK_RG; ++rg) {
          int n = (su * SK_RG + rg) * 8 + (lane & 7); if (n >= Ntot) n = Ntot - 1;
#pragma unroll
          for (int wh = 0; wh < NW; ++wh, ++it) {
            const uint32_t s = it % SK_NS;
            if (it >= SK_NS) mbar_wait(&empty[s], ((it / SK_NS) - 1) & 1);
            if (lane == 0) mbar_arrive_expect_tx(&full[s], bytes * 8);
            __syncwarp();
            if (lane < 8) bulk_g2s(ring + s * SK_STAGE_B + lane * SK_ROWB, skinny_row_ptr(sg, n, wh, ldw) + (size_t)c * SK_KC, bytes, &full[s]);
          }
        }
      }
    }
    return;
  }
  // ---------------- consumers: warp w owns k32-chunks 4w .. 4w+3 of every stage
  mb::pdl_wait();
  const int gid = lane >> 2, tid = lane & 3;
  const bf16* xlo = X + (size_t)gid * ldx;
  const bf16* xhi = X + (size_t)(gid + 8) * ldx;
  const bool lo_ok = gid < M, hi_ok = gid + 8 < M;
  const int4 z = make_int4(0, 0, 0, 0);
  float* scratch = reinterpret_cast<float*>(empty + SK_NS);     // [7 warps][32 lanes][8] partial tiles of one group
  uint32_t it = 0;
  for (int su = blockIdx.x; su < nsuper; su += gridDim.x) {
    float c[SK_RG][4], c2[SK_RG][4];
#pragma unroll
    for (int rg = 0; rg < SK_RG; ++rg)
#pragma unroll
      for (int e = 0; e < 4; ++e) { c[rg][e] = 0.f; c2[rg][e] = 0.f; }
    for (int ch = 0; ch < nchunks; ++ch) {
      const int kbase = ch * SK_KC, kc = min(SK_KC, K - kbase);
      int4 xl[4], xh[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int kk = (warp * 4 + u) * 32 + tid * 8;
        xl[u] = (kk < kc && lo_ok) ? __ldg(reinterpret_cast<const int4*>(xlo + kbase + kk)) : z;
        xh[u] = (kk < kc && hi_ok) ? __ldg(reinterpret_cast<const int4*>(xhi + kbase + kk)) : z;
      }
#pragma unroll
      for (int rg = 0; rg < SK_RG; ++rg) {
#pragma unroll
        for (int wh = 0; wh < NW; ++wh, ++it) {
          const uint32_t s = it % SK_NS;
          mbar_wait(&full[s], (it / SK_NS) & 1);
          const unsigned char* st = ring + s * SK_STAGE_B + gid * SK_ROWB;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int kk = (warp * 4 + u) * 32 + tid * 8;
            if (kk < kc) {
              const int4 wv = *reinterpret_cast<const int4*>(st + kk * 2);
              if (wh == 0) {
                mma_16816(c[rg], xl[u].x, xh[u].x, xl[u].y, xh[u].y, wv.x, wv.y);
                mma_16816(c[rg], xl[u].z, xh[u].z, xl[u].w, xh[u].w, wv.z, wv.w);
              } else {
                mma_16816(c2[rg], xl[u].x, xh[u].x, xl[u].y, xh[u].y, wv.x, wv.y);
                mma_16816(c2[rg], xl[u].z, xh[u].z, xl[u].w, xh[u].w, wv.z, wv.w);
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[s]);
        }
      }
    }
#pragma unroll
    for (int rg = 0; rg < SK_RG; ++rg) {                 // cross-warp reduction, one 8-row group per round (7 KB scratch)
      if (warp > 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          scratch[((warp - 1) * 32 + lane) * 8 + e] = c[rg][e];
          scratch[((warp - 1) * 32 + lane) * 8 + 4 + e] = c2[rg][e];
        }
      }
      consumer_sync();
      if (warp == 0) {
#pragma unroll
        for (int w = 0; w < 7; ++w)
#pragma unroll
          for (int e = 0; e < 4; ++e) { c[rg][e] += scratch[(w * 32 + lane) * 8 + e]; c2[rg][e] += scratch[(w * 32 + lane) * 8 + 4 + e]; }
        // c[0], c[1]: (row gid, cols 2 tid, +1) ; c[2], c[3]: (row gid + 8, same cols)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int m = gid + ((e >> 1) ? 8 : 0);
          const int n = (su * SK_RG + rg) * 8 + tid * 2 + (e & 1);
          if (m < M && n < Ntot) skinny_store(sg, MODE, m, n, c[rg][e], c2[rg][e], bias, addend, ld_add);
        }
      }
      consumer_sync();
    }
  }
}
constexpr int SK_SMEM_TOTAL = SK_SMEM + 7 * 32 * 8 * 4;

// ------------------------------------------------------------------ KV append
// k_new/v_new: [B, Hkv*hd] rows (row stride ld_new) -> cache[b, pos[b], :, :]  (cache: [B, cap, Hkv*hd])
__global__ void __launch_bounds__(256)
kv_append_kernel(const bf16* __restrict__ k_new, const bf16* __restrict__ v_new, bf16* __restrict__ kc,
                 bf16* __restrict__ vc, const int* __restrict__ pos, int pos_const, int B, int row_elems,
                 long long ld_new, long long cap) {
  const int b = blockIdx.y;
  const int p = pos ? pos[b] : pos_const;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= row_elems) return;
  const size_t dst = ((size_t)b * cap + p) * row_elems + i;
  *reinterpret_cast<int4*>(kc + dst) = *reinterpret_cast<const int4*>(k_new + (size_t)b * ld_new + i);
  *reinterpret_cast<int4*>(vc + dst) = *reinterpret_cast<const int4*>(v_new + (size_t)b * ld_new + i);
}

// ------------------------------------------------------------------ paged KV cache (mantis_b200/models/kv_cache.py)
// A page holds KV_PAGE tokens of ALL layers: [L][2 (k,v)][KV_PAGE][Hkv][hd].  The block table is int64 [B, max_blocks] of
// page base ADDRESSES (slabs are separate allocations, so the cache grows without copying or re-laying out old tokens).
constexpr int KV_PAGE_SHIFT = 7;
constexpr int KV_PAGE = 1 << KV_PAGE_SHIFT;

// TO_PAGES: rows [B, S, row] (strides in bytes) -> pages at token offsets start..start+S ; else the reverse (gather)
template <bool TO_PAGES>
__global__ void __launch_bounds__(128)
kv_page_copy_kernel(char* __restrict__ k_lin, char* __restrict__ v_lin, const int64_t* __restrict__ table, int table_stride,
                    long long layer_off_b, long long v_off_b, int start, int row_bytes, long long lin_sb, long long lin_ss) {
  const int s = blockIdx.x, b = blockIdx.y;
  const int t = start + s;
  char* page = reinterpret_cast<char*>(table[(size_t)b * table_stride + (t >> KV_PAGE_SHIFT)]) + layer_off_b +
               (size_t)(t & (KV_PAGE - 1)) * row_bytes;
  char* kl = k_lin + (size_t)b * lin_sb + (size_t)s * lin_ss;
  char* vl = v_lin + (size_t)b * lin_sb + (size_t)s * lin_ss;
  for (int i = threadIdx.x * 16; i < row_bytes; i += blockDim.x * 16) {
    if (TO_PAGES) {
      *reinterpret_cast<int4*>(page + i) = *reinterpret_cast<const int4*>(kl + i);
      *reinterpret_cast<int4*>(page + v_off_b + i) = *reinterpret_cast<const int4*>(vl + i);
    } else {
      *reinterpret_cast<int4*>(kl + i) = *reinterpret_cast<const int4*>(page + i);
      *reinterpret_cast<int4*>(vl + i) = *reinterpret_cast<const int4*>(page + v_off_b + i);
    }
  }
}

// RoPE(q) -> q_out ; RoPE(k) and v -> cache[b, ctx] : one launch per layer for the decode step
__global__ void __launch_bounds__(256)
rope_append_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, bf16* __restrict__ q_out,
                   bf16* __restrict__ kc, bf16* __restrict__ vc, const int64_t* __restrict__ pos, const float* __restrict__ inv_freq,
                   int H, int Hkv, int hd, int ctx, long long cap, float rope_scale, const int64_t* __restrict__ table,
                   int table_stride, long long layer_off, long long v_off) {
  mb::pdl_trigger();
  mb::pdl_wait();       // q/k/v come from the projection kernel just before
  const int b = blockIdx.y;
  const int half = hd >> 1;
  if (table) {          // paged cache: token ctx of sequence b lives in page table[b][ctx / 128], row ctx % 128
    kc = reinterpret_cast<bf16*>(table[(size_t)b * table_stride + (ctx >> KV_PAGE_SHIFT)]) + layer_off +
         (size_t)(ctx & (KV_PAGE - 1)) * Hkv * hd;
    vc = kc + v_off;
  } else {
    kc += ((size_t)b * cap + ctx) * Hkv * hd; vc += ((size_t)b * cap + ctx) * Hkv * hd;
  }
  const int n_q = H * half, n_k = Hkv * half, n_v = Hkv * hd / 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float ps = (float)pos[b];
  if (i < n_q + n_k) {
    const bool isq = i < n_q;
    const int j = isq ? i : i - n_q;
    const int h = j / half, d = j % half;
    float sn, cs; sincosf(ps * inv_freq[d], &sn, &cs);
    cs = __bfloat162float(__float2bfloat16_rn(cs * rope_scale)); sn = __bfloat162float(__float2bfloat16_rn(sn * rope_scale));
    const bf16* src = (isq ? q + (size_t)b * H * hd : k + (size_t)b * Hkv * hd) + (size_t)h * hd;
    bf16* dst = (isq ? q_out + (size_t)b * H * hd : kc) + (size_t)h * hd;
    const float x1 = __bfloat162float(src[d]), x2 = __bfloat162float(src[d + half]);
    const float y1 = __bfloat162float(__float2bfloat16_rn(x1 * cs)) + __bfloat162float(__float2bfloat16_rn(-x2 * sn));
    const float y2 = __bfloat162float(__float2bfloat16_rn(x2 * cs)) + __bfloat162float(__float2bfloat16_rn(x1 * sn));
    dst[d] = __float2bfloat16_rn(y1); dst[d + half] = __float2bfloat16_rn(y2);
  } else if (i < n_q + n_k + n_v) {
    const int j = i - n_q - n_k;
    *reinterpret_cast<int4*>(vc + j * 8) =
        *reinterpret_cast<const int4*>(v + (size_t)b * Hkv * hd + j * 8);
  }
}

// ------------------------------------------------------------------ split-KV decode attention (head_dim 128)
// grid (splits, kv_heads, batch), 4 warps; each warp owns 32-key tiles of the CTA's key range.  A warp puts its whole K and
// V tile in flight at once with 16-byte cp.async (2 x 8 KB per warp, ~200 KB per SM with 3 resident CTAs) and only then
// waits, so the kernel pays one HBM round trip per tile instead of one per dependent load (the first version streamed
// 33 KB per CTA through registers: 0.9 TB/s at bs 1).  Keys of older tokens do not depend on the previous kernel, so under
// programmatic dependent launch their loads are issued before griddepcontrol.wait.
constexpr int DHD = 128;
constexpr int DKT = 32;             // keys per tile (one per lane)
constexpr int DWARPS = 4;
constexpr int DK_ROWB = DHD * 2 + 16;                     // padded K row: lane = key, conflict-free 16-byte reads
constexpr int DV_ROWB = DHD * 2;
constexpr int DWARP_SMEM = DKT * (DK_ROWB + DV_ROWB);     // 16896 B per warp
template <int G> constexpr int dec_smem() { return DWARPS * DWARP_SMEM + G * DHD * 4; }

struct DecP {
  const bf16* q; long long q_sb, q_sh;          // q [B, H, hd]
  const bf16* k; const bf16* v; long long kv_sb, kv_ss, kv_sh;   // cache [B, cap, Hkv, hd]
  const uint32_t* kbits; int kbits_stride;      // key-valid bitmask or null
  float* part;                                  // [B, H, splits, hd + 2]
  int B, H, Hkv, ctx, splits, chunk; float scale;
  const int64_t* table; int table_stride; long long layer_off, v_off;   // paged cache (table != null): k/v unused
  int hk_fast;                                  // grid = (Hkv, splits, B) instead of (splits, Hkv, B)
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int G>
__global__ void __launch_bounds__(DWARPS * 32)
decode_attn_kernel(DecP p) {
  extern __shared__ __align__(16) unsigned char dsm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned char* Kt = dsm + w * DWARP_SMEM;
  unsigned char* Vt = Kt + DKT * DK_ROWB;
  float* Qs = reinterpret_cast<float*>(dsm + DWARPS * DWARP_SMEM);      // [G][DHD], pre-scaled
  const int sp = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
  const int k_begin = sp * p.chunk, k_end = min(p.ctx, k_begin + p.chunk);
  mb::pdl_trigger();
  const bool has_new = (k_end == p.ctx);       // the range holding the token the previous kernel just appended
  if (has_new) mb::pdl_wait();

  // chunk is a multiple of DKT, so a 32-key tile never straddles a 128-token page: one table lookup per tile
  auto issue_tile = [&](int k0) {
    const bf16* kt; const bf16* vt;
    if (p.table) {
      kt = reinterpret_cast<const bf16*>(p.table[(size_t)b * p.table_stride + (k0 >> KV_PAGE_SHIFT)]) + p.layer_off +
           (size_t)(k0 & (KV_PAGE - 1)) * p.kv_ss + (size_t)hk * p.kv_sh;
      vt = kt + p.v_off;
    } else {
      kt = p.k + (size_t)b * p.kv_sb + (size_t)k0 * p.kv_ss + (size_t)hk * p.kv_sh;
      vt = p.v + (size_t)b * p.kv_sb + (size_t)k0 * p.kv_ss + (size_t)hk * p.kv_sh;
    }
    const int nk = min(DKT, k_end - k0);
#pragma unroll 4
    for (int e = lane; e < DKT * 16; e += 32) {          // 16 pieces of 16 B per key; rows past k_end are zero-filled
      const int j = e >> 4, c = e & 15;
      const bool ok = j < nk;
      cp_async16(Kt + j * DK_ROWB + c * 16, ok ? (const void*)(kt + (size_t)j * p.kv_ss + c * 8) : (const void*)kt, ok ? 16 : 0);
    }
    cp_async_commit();
#pragma unroll 4
    for (int e = lane; e < DKT * 16; e += 32) {
      const int j = e >> 4, c = e & 15;
      const bool ok = j < nk;
      cp_async16(Vt + j * DV_ROWB + c * 16, ok ? (const void*)(vt + (size_t)j * p.kv_ss + c * 8) : (const void*)vt, ok ? 16 : 0);
    }
    cp_async_commit();
  };

  int k0 = k_begin + w * DKT;
  if (k0 < k_end) issue_tile(k0);
  if (!has_new) mb::pdl_wait();                // q is the previous kernel's output
  for (int i = threadIdx.x; i < G * DHD; i += blockDim.x) {
    const int g = i / DHD, d = i % DHD;
    Qs[i] = __bfloat162float(p.q[(size_t)b * p.q_sb + (size_t)(hk * G + g) * p.q_sh + d]) * p.scale;
  }
  __syncthreads();

  float m[G], l[G], acc[G][4];
#pragma unroll
  for (int g = 0; g < G; ++g) { m[g] = -INFINITY; l[g] = 0.f; acc[g][0] = acc[g][1] = acc[g][2] = acc[g][3] = 0.f; }
  for (; k0 < k_end; k0 += DWARPS * DKT) {
    cp_async_wait<1>();                        // K tile landed, V may still be in flight
    __syncwarp();
    const int kj = k0 + lane;
    bool vis = kj < k_end;
    if (vis && p.kbits) vis = (p.kbits[(size_t)b * p.kbits_stride + (kj >> 5)] >> (kj & 31)) & 1u;
    float s[G];
#pragma unroll
    for (int g = 0; g < G; ++g) s[g] = 0.f;
    const unsigned char* krow = Kt + lane * DK_ROWB;
#pragma unroll 4
    for (int c = 0; c < 16; ++c) {
      const int4 raw = *reinterpret_cast<const int4*>(krow + c * 16);
      const bf162* kh = reinterpret_cast<const bf162*>(&raw);
      const float2 k01 = __bfloat1622float2(kh[0]), k23 = __bfloat1622float2(kh[1]);
      const float2 k45 = __bfloat1622float2(kh[2]), k67 = __bfloat1622float2(kh[3]);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4 qa = *reinterpret_cast<const float4*>(Qs + g * DHD + c * 8);
        const float4 qb = *reinterpret_cast<const float4*>(Qs + g * DHD + c * 8 + 4);
        float t = s[g];
        t = fmaf(qa.x, k01.x, t); t = fmaf(qa.y, k01.y, t); t = fmaf(qa.z, k23.x, t); t = fmaf(qa.w, k23.y, t);
        t = fmaf(qb.x, k45.x, t); t = fmaf(qb.y, k45.y, t); t = fmaf(qb.z, k67.x, t); t = fmaf(qb.w, k67.y, t);
        s[g] = t;
      }
    }
    float pj[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float sg = vis ? s[g] : -INFINITY;
      const float mt = mb::warp_max(sg);
      const float m_new = fmaxf(m[g], mt);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = (m[g] == -INFINITY) ? 0.f : __expf(m[g] - m_use);
      pj[g] = vis ? __expf(sg - m_use) : 0.f;
      l[g] = l[g] * corr + mb::warp_sum(pj[g]);
      m[g] = m_new;
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[g][t] *= corr;
    }
    cp_async_wait<0>();                        // V tile
    __syncwarp();
    const int nk = min(DKT, k_end - k0);
    for (int j = 0; j < nk; ++j) {             // lanes own 4 consecutive dims
      const uint2 raw = *reinterpret_cast<const uint2*>(Vt + j * DV_ROWB + lane * 8);
      const bf162* vh = reinterpret_cast<const bf162*>(&raw);
      const float2 v01 = __bfloat1622float2(vh[0]), v23 = __bfloat1622float2(vh[1]);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float pb = __shfl_sync(0xffffffffu, pj[g], j);
        acc[g][0] = fmaf(pb, v01.x, acc[g][0]); acc[g][1] = fmaf(pb, v01.y, acc[g][1]);
        acc[g][2] = fmaf(pb, v23.x, acc[g][2]); acc[g][3] = fmaf(pb, v23.y, acc[g][3]);
      }
    }
    __syncwarp();
    if (k0 + DWARPS * DKT < k_end) issue_tile(k0 + DWARPS * DKT);     // only for contexts beyond 64 x 128 keys
  }
  // combine the 4 warps through shared memory (the tiles are dead by now)
  float* red = reinterpret_cast<float*>(dsm);            // [DWARPS][G][DHD + 2] floats (<= 16.6 KB)
  __syncthreads();
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float* r = red + ((size_t)w * G + g) * (DHD + 2);
#pragma unroll
    for (int t = 0; t < 4; ++t) r[lane * 4 + t] = acc[g][t];
    if (lane == 0) { r[DHD] = m[g]; r[DHD + 1] = l[g]; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G * DHD; i += blockDim.x) {
    const int g = i / DHD, d = i % DHD;
    float M_ = -INFINITY;
#pragma unroll
    for (int ww = 0; ww < DWARPS; ++ww) M_ = fmaxf(M_, red[((size_t)ww * G + g) * (DHD + 2) + DHD]);
    float o = 0.f, L = 0.f;
#pragma unroll
    for (int ww = 0; ww < DWARPS; ++ww) {
      const float* r = red + ((size_t)ww * G + g) * (DHD + 2);
      const float sc = (r[DHD] == -INFINITY) ? 0.f : __expf(r[DHD] - M_);
      o += r[d] * sc; L += r[DHD + 1] * sc;
    }
    float* out = p.part + (((size_t)b * p.H + hk * G + g) * p.splits + sp) * (DHD + 2);
    out[d] = o;
    if (d == 0) { out[DHD] = M_; out[DHD + 1] = L; }
  }
}

// ---- tensor-core variant (the default): same tiling and loads, but the per-tile math runs on mma.sync m16n8k16 --
// S = Q K^T with the GQA group's query heads on the M side (rows >= G are zero), P V with the S accumulators re-packed as the
// A operand (their register layouts coincide) and V^T fetched with ldmatrix.trans.  ~220 instructions per 32-key tile instead
// of ~1700 on the FMA pipe, so a warp spends its time waiting for HBM, not issuing (the SIMT kernel reached 3.3 TB/s at
// batch 16).  fp32 softmax, P rounded to bf16 before P V like the reference (hf: llama/modeling_llama.py:216-218).
constexpr int DM_ROWB = DHD * 2 + 16;                       // K and V rows padded to 272 B (conflict-free ldmatrix)
constexpr int DM_WARP_SMEM = DKT * DM_ROWB * 2;             // 17408 B per warp
constexpr int DM_SMEM = DWARPS * DM_WARP_SMEM;
constexpr int DM_RPITCH = DHD + 4;                          // per-warp partial rows (fp32): 16-byte aligned, m and l behind the dims

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(smem_u32(row)));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(smem_u32(row)));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const bf162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void cp_async16_s(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_zs(uint32_t smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}

// The first version of this kernel spent 55 % of the SM's issue slots (ncu, bs 16: 2100 instructions per warp per 32-key
// tile, a quarter of them IMAD address arithmetic of the predicated 16-byte copies) -- an HBM stream cannot hide behind that.
// Now: a lane's 32 copies per tile walk two fixed strides (one 64-bit add each, no predicates on full tiles), Q fragments
// come straight from global memory (no staging tile, no barrier), K fragments through ldmatrix.x4 (16 instead of 64 shared
// loads), full tiles skip the mask arithmetic, and the 4 warps' partials meet in each warp's own (dead) tile with ONE barrier.
__global__ void __launch_bounds__(DWARPS * 32, 3)
decode_attn_mma_kernel(DecP p, int G) {
  extern __shared__ __align__(16) unsigned char dsm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int gid = lane >> 2, tid = lane & 3;
  unsigned char* Kt = dsm + w * DM_WARP_SMEM;
  unsigned char* Vt = Kt + DKT * DM_ROWB;
  // hk_fast: the 8 kv heads of one token range are neighbours in the launch order, so the 256-byte pieces they read out of
  // the same [token][head][dim] rows are in flight together (whole 2 KB rows per DRAM page visit instead of one eighth)
  const int sp = p.hk_fast ? blockIdx.y : blockIdx.x, hk = p.hk_fast ? blockIdx.x : blockIdx.y, b = blockIdx.z;
  const int k_begin = sp * p.chunk, k_end = min(p.ctx, k_begin + p.chunk);
  mb::pdl_trigger();
  const bool has_new = (k_end == p.ctx);
  if (has_new) mb::pdl_wait();

  // lane -> rows (lane >> 4) + 2 i, 16-byte chunk (lane & 15) of every row
  const int r0 = lane >> 4, c16 = lane & 15;
  const uint32_t kdst = smem_u32(Kt) + r0 * DM_ROWB + c16 * 16;
  const uint32_t vdst = kdst + DKT * DM_ROWB;
  const size_t lane_off = (size_t)r0 * p.kv_ss + c16 * 8;
  const size_t step2 = 2 * (size_t)p.kv_ss;

  auto issue_tile = [&](int k0) {
    const bf16* kt; const bf16* vt;
    if (p.table) {
      kt = reinterpret_cast<const bf16*>(p.table[(size_t)b * p.table_stride + (k0 >> KV_PAGE_SHIFT)]) + p.layer_off +
           (size_t)(k0 & (KV_PAGE - 1)) * p.kv_ss + (size_t)hk * p.kv_sh;
      vt = kt + p.v_off;
    } else {
      kt = p.k + (size_t)b * p.kv_sb + (size_t)k0 * p.kv_ss + (size_t)hk * p.kv_sh;
      vt = p.v + (size_t)b * p.kv_sb + (size_t)k0 * p.kv_ss + (size_t)hk * p.kv_sh;
    }
    const int nk = min(DKT, k_end - k0);
    const bf16* ks = kt + lane_off;
    const bf16* vs = vt + lane_off;
    if (nk == DKT) {
#pragma unroll
      for (int i = 0; i < DKT / 2; ++i) cp_async16_s(kdst + i * 2 * DM_ROWB, ks + i * step2);
      cp_async_commit();
#pragma unroll
      for (int i = 0; i < DKT / 2; ++i) cp_async16_s(vdst + i * 2 * DM_ROWB, vs + i * step2);
      cp_async_commit();
    } else {                                   // the last tile of the context: rows >= nk are zero-filled
#pragma unroll 4
      for (int i = 0; i < DKT / 2; ++i) {
        const bool ok = r0 + 2 * i < nk;
        cp_async16_zs(kdst + i * 2 * DM_ROWB, ok ? (const void*)(ks + i * step2) : (const void*)kt, ok ? 16 : 0);
      }
      cp_async_commit();
#pragma unroll 4
      for (int i = 0; i < DKT / 2; ++i) {
        const bool ok = r0 + 2 * i < nk;
        cp_async16_zs(vdst + i * 2 * DM_ROWB, ok ? (const void*)(vs + i * step2) : (const void*)vt, ok ? 16 : 0);
      }
      cp_async_commit();
    }
  };

  int k0 = k_begin + w * DKT;
  if (k0 < k_end) issue_tile(k0);
  if (!has_new) mb::pdl_wait();
  // A fragments of Q (rows = the GQA group's heads, rows >= G zero): (row gid, dims ks*16 + 2 tid .. +1) and (.. + 8)
  uint32_t qa[8][2];
  {
    const bf16* qrow = p.q + (size_t)b * p.q_sb + (size_t)(hk * G + (gid < G ? gid : 0)) * p.q_sh + 2 * tid;
    const bool q32 = !((p.q_sb | p.q_sh) & 1) && !(reinterpret_cast<uintptr_t>(p.q) & 3);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t lo, hi;
      if (q32) {
        lo = *reinterpret_cast<const uint32_t*>(qrow + ks * 16);
        hi = *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 8);
      } else {
        const unsigned short* qs = reinterpret_cast<const unsigned short*>(qrow + ks * 16);
        lo = (uint32_t)qs[0] | ((uint32_t)qs[1] << 16);
        hi = (uint32_t)qs[8] | ((uint32_t)qs[9] << 16);
      }
      qa[ks][0] = gid < G ? lo : 0u;
      qa[ks][1] = gid < G ? hi : 0u;
    }
  }
  float m_run = -INFINITY, l_run = 0.f;          // row gid's running max / sum (replicated over the 4 lanes of a quad)
  float o[16][4];
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }

  for (; k0 < k_end; k0 += DWARPS * DKT) {
    cp_async_wait<1>();
    __syncwarp();
    float sc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
      // ldmatrix.x4: matrices 0..3 = keys nt*8..+7 x dims kp*32 + {0-7, 8-15, 16-23, 24-31}; lanes 8 i .. 8 i + 7 address matrix i
      const unsigned char* kr = Kt + (nt * 8 + (lane & 7)) * DM_ROWB + (lane >> 3) * 16;
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(b0, b1, b2, b3, kr + kp * 64);
        mma_16816(sc[nt], qa[2 * kp][0], 0u, qa[2 * kp][1], 0u, b0, b1);
        mma_16816(sc[nt], qa[2 * kp + 1][0], 0u, qa[2 * kp + 1][1], 0u, b2, b3);
      }
    }
    float mx = -INFINITY;
    if (k0 + DKT <= k_end && !p.kbits) {         // a full tile with no key mask: nothing to test
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) { sc[nt][e] *= p.scale; mx = fmaxf(mx, sc[nt][e]); }
    } else {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int kj = k0 + nt * 8 + 2 * tid + e;
          bool vis = kj < k_end;
          if (vis && p.kbits) vis = (p.kbits[(size_t)b * p.kbits_stride + (kj >> 5)] >> (kj & 31)) & 1u;
          sc[nt][e] = vis ? sc[nt][e] * p.scale : -INFINITY;
          mx = fmaxf(mx, sc[nt][e]);
        }
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float m_new = fmaxf(m_run, mx);
    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
    const float corr = (m_run == -INFINITY) ? 0.f : __expf(m_run - m_use);
    float rs = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float pv = (sc[nt][e] == -INFINITY) ? 0.f : __expf(sc[nt][e] - m_use);
        sc[nt][e] = pv; rs += pv;
      }
    rs += __shfl_xor_sync(0xffffffffu, rs, 1);
    rs += __shfl_xor_sync(0xffffffffu, rs, 2);
    l_run = l_run * corr + rs; m_run = m_new;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) { o[nt][0] *= corr; o[nt][1] *= corr; }
    uint32_t pa[2][2];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      pa[kk][0] = pack_bf16x2(sc[2 * kk][0], sc[2 * kk][1]);
      pa[kk][1] = pack_bf16x2(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
    }
    cp_async_wait<0>();
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      uint32_t v0, v1, v2, v3;          // (keys 0-7, 8-15, 16-23, 24-31) x dims nt*8..+7, transposed
      ldmatrix_x4_trans(v0, v1, v2, v3, Vt + lane * DM_ROWB + nt * 16);
      mma_16816(o[nt], pa[0][0], 0u, pa[0][1], 0u, v0, v1);
      mma_16816(o[nt], pa[1][0], 0u, pa[1][1], 0u, v2, v3);
    }
    __syncwarp();
    if (k0 + DWARPS * DKT < k_end) issue_tile(k0 + DWARPS * DKT);     // only for contexts beyond 64 x 128 keys
  }
  // every warp leaves its (unnormalised) partial rows in its OWN tile -- dead by now, so no barrier before the writes
  float* mine = reinterpret_cast<float*>(Kt);                // [G][DM_RPITCH]: dims, then m, l
  if (gid < G) {
    float* r = mine + gid * DM_RPITCH;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) *reinterpret_cast<float2*>(r + nt * 8 + 2 * tid) = make_float2(o[nt][0], o[nt][1]);
    if (tid == 0) { r[DHD] = m_run; r[DHD + 1] = l_run; }
  }
  __syncthreads();
  for (int g = w; g < G; g += DWARPS) {                      // warp w merges head g: lanes own 4 consecutive dims
    float mw[DWARPS], lw[DWARPS];
    float M_ = -INFINITY;
#pragma unroll
    for (int ww = 0; ww < DWARPS; ++ww) {
      const float* r = reinterpret_cast<const float*>(dsm + ww * DM_WARP_SMEM) + g * DM_RPITCH;
      mw[ww] = r[DHD]; lw[ww] = r[DHD + 1];
      M_ = fmaxf(M_, mw[ww]);
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float L = 0.f;
#pragma unroll
    for (int ww = 0; ww < DWARPS; ++ww) {
      const float* r = reinterpret_cast<const float*>(dsm + ww * DM_WARP_SMEM) + g * DM_RPITCH;
      const float s2 = (mw[ww] == -INFINITY) ? 0.f : __expf(mw[ww] - M_);
      const float4 v = *reinterpret_cast<const float4*>(r + lane * 4);
      acc.x += v.x * s2; acc.y += v.y * s2; acc.z += v.z * s2; acc.w += v.w * s2;
      L += lw[ww] * s2;
    }
    float* out = p.part + (((size_t)b * p.H + hk * G + g) * p.splits + sp) * (DHD + 2);
    *reinterpret_cast<float2*>(out + lane * 4) = make_float2(acc.x, acc.y);
    *reinterpret_cast<float2*>(out + lane * 4 + 2) = make_float2(acc.z, acc.w);
    if (lane == 0) { out[DHD] = M_; out[DHD + 1] = L; }
  }
}

// merges the per-split (max, sum, out) triples: warp 0 turns the <= 64 (m, l) pairs into rescale factors with two parallel
// loads per lane, then every thread sums its output dim over the splits with independent (unrolled) loads
__global__ void __launch_bounds__(DHD)
decode_combine_kernel(const float* __restrict__ part, bf16* __restrict__ o, long long o_sb, long long o_sh, int H,
                      int splits) {
  __shared__ float sc_s[64];
  __shared__ float inv_l;
  mb::pdl_trigger();
  mb::pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y, d = threadIdx.x;
  const float* base = part + ((size_t)b * H + h) * splits * (DHD + 2);
  if (d < 32) {
    const int s0 = d, s1 = d + 32;
    const float m0 = s0 < splits ? base[(size_t)s0 * (DHD + 2) + DHD] : -INFINITY;
    const float m1 = s1 < splits ? base[(size_t)s1 * (DHD + 2) + DHD] : -INFINITY;
    const float l0 = s0 < splits ? base[(size_t)s0 * (DHD + 2) + DHD + 1] : 0.f;
    const float l1 = s1 < splits ? base[(size_t)s1 * (DHD + 2) + DHD + 1] : 0.f;
    const float M_ = mb::warp_max(fmaxf(m0, m1));
    const float c0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - M_);
    const float c1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - M_);
    const float L = mb::warp_sum(l0 * c0 + l1 * c1);
    sc_s[s0] = c0; sc_s[s1] = c1;
    if (d == 0) inv_l = L > 0.f ? 1.f / L : 0.f;
  }
  __syncthreads();
  float acc = 0.f;
#pragma unroll 8
  for (int s = 0; s < splits; ++s) acc = fmaf(base[(size_t)s * (DHD + 2) + d], sc_s[s], acc);
  o[(size_t)b * o_sb + (size_t)h * o_sh + d] = __float2bfloat16_rn(acc * inv_l);
}

template <int MT>
int launch_skinny(const void* X, const SkinnySeg& sg, const void* bias, const void* addend, int M, int K, long long ldx,
                  long long ldw, long long ld_add, cudaStream_t st) {
  constexpr int RPW = (MT <= 2) ? 2 : (MT <= 4 ? 2 : 4);
  constexpr int KU = (MT <= 2) ? 4 : (MT <= 4 ? 2 : 1);
  const int Ntot = (sg.mode == 1) ? sg.N[0] : (sg.N[0] + sg.N[1] + sg.N[2]);
  const int rows_per_block = 8 * RPW;
  const int grid = (Ntot + rows_per_block - 1) / rows_per_block;
  skinny_gemm_kernel<MT, RPW, KU><<<grid, 256, 0, st>>>((const bf16*)X, sg, (const bf16*)bias, (const bf16*)addend, M, K,
                                                       ldx, ldw, ld_add);
  return 0;
}
// persistent grid: every CTA resident at once (occupancy x SMs slots), a balanced number of 8-row groups per CTA
static inline int skinny_grid(int ngroups, int ctas_per_sm) {
  const int slots = mb::num_sms() * (ctas_per_sm > 0 ? ctas_per_sm : 1);
  const int per = (ngroups + slots - 1) / slots;
  return (ngroups + per - 1) / per;
}
template <int MT, int MODE>
int launch_rows(const void* X, const SkinnySeg& sg, const void* bias, const void* addend, int M, int K, long long ldx,
                long long ldw, long long ld_add, cudaStream_t st, const void* gamma = nullptr, float eps = 0.f) {
  static const int occ = [] {             // thread-safe one-time setup
    int o = 0;
    cudaFuncSetAttribute(skinny_rows_kernel<MT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, skinny_rows_kernel<MT, MODE>, SK_THREADS, RS_SMEM);
    return o < 1 ? 1 : o;
  }();
  const int Ntot = (MODE == 1) ? sg.N[0] : (sg.N[0] + sg.N[1] + sg.N[2]);
  const int ngroups = (Ntot + 7) / 8;
  mb::launch_ex(skinny_rows_kernel<MT, MODE>, dim3(skinny_grid(ngroups, occ)), dim3(SK_THREADS), RS_SMEM, st, mb::pdl_mode() != 0,
                (const bf16*)X, sg, (const bf16*)bias, (const bf16*)addend, M, K, ldx, ldw, ld_add, ngroups, (const bf16*)gamma, eps);
  return 0;
}
template <int MODE>
int launch_ring(const void* X, const SkinnySeg& sg, const void* bias, const void* addend, int M, int K, long long ldx,
                long long ldw, long long ld_add, cudaStream_t st) {
  static const int occ = [] {
    int o = 0;
    cudaFuncSetAttribute(skinny_ring_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_TOTAL);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, skinny_ring_kernel<MODE>, SK_THREADS, SK_SMEM_TOTAL);
    return o < 1 ? 1 : o;
  }();
  const int Ntot = (MODE == 1) ? sg.N[0] : (sg.N[0] + sg.N[1] + sg.N[2]);
  const int ngroups = (Ntot + 8 * SK_RG - 1) / (8 * SK_RG);
  mb::launch_ex(skinny_ring_kernel<MODE>, dim3(skinny_grid(ngroups, occ)), dim3(SK_THREADS), SK_SMEM_TOTAL, st, mb::pdl_mode() != 0,
                (const bf16*)X, sg, (const bf16*)bias, (const bf16*)addend, M, K, ldx, ldw, ld_add, ngroups);
  return 0;
}
static int skinny_ring_enabled() {
  static const int v = [] { const char* e = getenv("MB200_SKINNY_RING"); return (e && e[0] == '0') ? 0 : 1; }();
  return v;
}
// can the row-stage kernel (the only one with the fused RMSNorm prologue) take this problem?
static bool skinny_rows_ok(const SkinnySeg& sg, int M, int K) {
  const bool aligned = !((reinterpret_cast<uintptr_t>(sg.W[0]) | reinterpret_cast<uintptr_t>(sg.W[1]) |
                          reinterpret_cast<uintptr_t>(sg.W[2])) & 15);
  return skinny_ring_enabled() && aligned && (M <= 2 || (M <= 4 && (K % 32) != 0));
}
int dispatch_skinny(const void* X, const SkinnySeg& sg, const void* bias, const void* addend, int M, int K, long long ldx,
                    long long ldw, long long ld_add, cudaStream_t st, const void* gamma = nullptr, float eps = 0.f) {
  if (skinny_ring_enabled() && (sg.mode == 1 || sg.mode == 0)) {
    const bool aligned = !((reinterpret_cast<uintptr_t>(sg.W[0]) | reinterpret_cast<uintptr_t>(sg.W[1]) |
                            reinterpret_cast<uintptr_t>(sg.W[2])) & 15);
    if (aligned && (M <= 2 || (M <= 4 && (K % 32) != 0))) {
      if (sg.mode == 1) {
        if (M == 1) return launch_rows<1, 1>(X, sg, bias, addend, M, K, ldx, ldw, ld_add, st, gamma, eps);
        if (M == 2) return launch_rows<2, 1>(X, sg, bias, addend, M, K, ldx, ldw, ld_add, st, gamma, eps);
        return launch_rows<4, 1>(X, sg, bias, addend, M, K, ldx, ldw, ld_add, st, gamma, eps);
      }
      if (M == 1) return launch_rows<1, 0>(X, sg, bias, addend, M, K, ldx, ldw, ld_add, st, gamma, eps);
      if (M == 2) return launch_rows<2, 0>(X, sg, bias, addend, M, K, ldx, ldw, ld_add, st, gamma, eps);
      return launch_rows<4, 0>(X, sg, bias, addend, M, K, ldx, ldw, ld_add, st, gamma, eps);
    }
    if (aligned && (K % 32) == 0) {
      if (sg.mode == 1) return launch_ring<1>(X, sg, bias, addend, M, K, ldx, ldw, ld_add, st);
      return launch_ring<0>(X, sg, bias, addend, M, K, ldx, ldw, ld_add, st);
    }
  }
  if (M > 8 && (K % 32) == 0) {           // tensor-core (mma.sync) variant: HBM-bound instead of FMA-bound
    const int Ntot = (sg.mode == 1) ? sg.N[0] : (sg.N[0] + sg.N[1] + sg.N[2]);
    const int grid = (Ntot + 7) / 8;
    if (sg.mode == 1) skinny_mma_kernel<2><<<grid, 128, 0, st>>>((const bf16*)X, sg, (const bf16*)bias, (const bf16*)addend, M, K, ldx, ldw, ld_add);
    else              skinny_mma_kernel<4><<<grid, 128, 0, st>>>((const bf16*)X, sg, (const bf16*)bias, (const bf16*)addend, M, K, ldx, ldw, ld_add);
    return 0;
  }
  if (M == 1) return launch_skinny<1>(X, sg, bias, addend, M, K, ldx, ldw, ld_add, st);
  if (M == 2) return launch_skinny<2>(X, sg, bias, addend, M, K, ldx, ldw, ld_add, st);
  if (M <= 4) return launch_skinny<4>(X, sg, bias, addend, M, K, ldx, ldw, ld_add, st);
  if (M <= 8) return launch_skinny<8>(X, sg, bias, addend, M, K, ldx, ldw, ld_add, st);
  return launch_skinny<16>(X, sg, bias, addend, M, K, ldx, ldw, ld_add, st);
}
}  // namespace

extern "C" {

// C[M,N] = X[M,K] W[N,K]^T (+bias) (+addend), bf16, M <= 16, K % 8 == 0, 16-byte aligned rows.
int mb200_skinny_gemm_bf16(const void* X, const void* W, void* C, const void* bias, const void* addend, int M, int N,
                           int K, long long ldx, long long ldw, long long ldc, long long ld_add, void* stream) {
  if (M <= 0 || N <= 0) return MB200_OK;
  if (M > 16 || (K & 7) || (ldx & 7) || (ldw & 7)) return -ENOTSUP;
  if ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W)) & 15) return -ENOTSUP;
  SkinnySeg sg; sg.nseg = 1; sg.mode = 0;
  sg.W[0] = (const bf16*)W; sg.C[0] = (bf16*)C; sg.N[0] = N; sg.ldc[0] = ldc;
  sg.W[1] = sg.W[2] = nullptr; sg.C[1] = sg.C[2] = nullptr; sg.N[1] = sg.N[2] = 0; sg.ldc[1] = sg.ldc[2] = 0;
  dispatch_skinny(X, sg, bias, addend, M, K, ldx, ldw, ld_add, (cudaStream_t)stream);
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

// Three projections of the same input in one launch: Ci[M,Ni] = X Wi^T (q/k/v).  W rows share ldw.
int mb200_skinny_gemm3_bf16(const void* X, const void* W0, const void* W1, const void* W2, void* C0, void* C1, void* C2,
                            int M, int N0, int N1, int N2, int K, long long ldx, long long ldw, void* stream) {
  if (M <= 0) return MB200_OK;
  if (M > 16 || (K & 7) || (ldx & 7) || (ldw & 7)) return -ENOTSUP;
  SkinnySeg sg; sg.nseg = 3; sg.mode = 0;
  sg.W[0] = (const bf16*)W0; sg.W[1] = (const bf16*)W1; sg.W[2] = (const bf16*)W2;
  sg.C[0] = (bf16*)C0; sg.C[1] = (bf16*)C1; sg.C[2] = (bf16*)C2;
  sg.N[0] = N0; sg.N[1] = N1; sg.N[2] = N2; sg.ldc[0] = N0; sg.ldc[1] = N1; sg.ldc[2] = N2;
  dispatch_skinny(X, sg, nullptr, nullptr, M, K, ldx, ldw, 0, (cudaStream_t)stream);
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

// C[M,N] = silu(X Wg^T) * (X Wu^T)   (SwiGLU MLP input half, fused)
int mb200_skinny_swiglu_bf16(const void* X, const void* Wg, const void* Wu, void* C, int M, int N, int K, long long ldx,
                             long long ldw, long long ldc, void* stream) {
  if (M <= 0) return MB200_OK;
  if (M > 16 || (K & 7) || (ldx & 7) || (ldw & 7)) return -ENOTSUP;
  SkinnySeg sg; sg.nseg = 2; sg.mode = 1;
  sg.W[0] = (const bf16*)Wg; sg.W[1] = (const bf16*)Wu; sg.W[2] = nullptr;
  sg.C[0] = (bf16*)C; sg.C[1] = sg.C[2] = nullptr; sg.N[0] = N; sg.N[1] = N; sg.N[2] = 0;
  sg.ldc[0] = ldc; sg.ldc[1] = sg.ldc[2] = 0;
  dispatch_skinny(X, sg, nullptr, nullptr, M, K, ldx, ldw, 0, (cudaStream_t)stream);
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

// Fused RMSNorm + projections for the decode step: Ci = rmsnorm(X; gamma, eps) Wi^T.  The norm runs inside the GEMM when
// the row-stage kernel applies (M <= 2); otherwise X is normalised into xn_scratch [M, K] first.
int mb200_skinny_gemm3_norm_bf16(const void* X, const void* gamma, float eps, void* xn_scratch, const void* W0, const void* W1,
                                 const void* W2, void* C0, void* C1, void* C2, int M, int N0, int N1, int N2, int K,
                                 long long ldw, void* stream) {
  if (M <= 0) return MB200_OK;
  if (M > 16 || (K & 7) || (ldw & 7)) return -ENOTSUP;
  SkinnySeg sg; sg.nseg = 3; sg.mode = 0;
  sg.W[0] = (const bf16*)W0; sg.W[1] = (const bf16*)W1; sg.W[2] = (const bf16*)W2;
  sg.C[0] = (bf16*)C0; sg.C[1] = (bf16*)C1; sg.C[2] = (bf16*)C2;
  sg.N[0] = N0; sg.N[1] = N1; sg.N[2] = N2; sg.ldc[0] = N0; sg.ldc[1] = N1; sg.ldc[2] = N2;
  if (skinny_rows_ok(sg, M, K) && !(reinterpret_cast<uintptr_t>(gamma) & 15)) {
    dispatch_skinny(X, sg, nullptr, nullptr, M, K, K, ldw, 0, (cudaStream_t)stream, gamma, eps);
  } else {
    const int rc = mb200_rmsnorm_fwd(X, gamma, xn_scratch, nullptr, M, K, eps, MB200_DTYPE_BF16, stream);
    if (rc) return rc;
    dispatch_skinny(xn_scratch, sg, nullptr, nullptr, M, K, K, ldw, 0, (cudaStream_t)stream);
  }
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

// C[M,N] = silu(xn Wg^T) * (xn Wu^T) with xn = rmsnorm(X; gamma, eps)
int mb200_skinny_swiglu_norm_bf16(const void* X, const void* gamma, float eps, void* xn_scratch, const void* Wg,
                                  const void* Wu, void* C, int M, int N, int K, long long ldw, long long ldc, void* stream) {
  if (M <= 0) return MB200_OK;
  if (M > 16 || (K & 7) || (ldw & 7)) return -ENOTSUP;
  SkinnySeg sg; sg.nseg = 2; sg.mode = 1;
  sg.W[0] = (const bf16*)Wg; sg.W[1] = (const bf16*)Wu; sg.W[2] = nullptr;
  sg.C[0] = (bf16*)C; sg.C[1] = sg.C[2] = nullptr; sg.N[0] = N; sg.N[1] = N; sg.N[2] = 0;
  sg.ldc[0] = ldc; sg.ldc[1] = sg.ldc[2] = 0;
  if (skinny_rows_ok(sg, M, K) && !(reinterpret_cast<uintptr_t>(gamma) & 15)) {
    dispatch_skinny(X, sg, nullptr, nullptr, M, K, K, ldw, 0, (cudaStream_t)stream, gamma, eps);
  } else {
    const int rc = mb200_rmsnorm_fwd(X, gamma, xn_scratch, nullptr, M, K, eps, MB200_DTYPE_BF16, stream);
    if (rc) return rc;
    dispatch_skinny(xn_scratch, sg, nullptr, nullptr, M, K, K, ldw, 0, (cudaStream_t)stream);
  }
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

int mb200_kv_append(const void* k_new, const void* v_new, void* k_cache, void* v_cache, const int* pos_dev, int pos_const,
                    int B, int row_elems, long long ld_new, long long capacity, void* stream) {
  if (B <= 0) return MB200_OK;
  if (row_elems & 7) return -EINVAL;
  dim3 grid((row_elems / 8 + 255) / 256, B);
  kv_append_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)k_new, (const bf16*)v_new, (bf16*)k_cache,
                                                          (bf16*)v_cache, pos_dev, pos_const, B, row_elems, ld_new, capacity);
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

int mb200_rope_append_bf16(const void* q, const void* k, const void* v, void* q_out, void* k_cache, void* v_cache,
                           const int64_t* pos, const float* inv_freq, int B, int H, int Hkv, int hd, int ctx,
                           long long capacity, float rope_scale, void* stream) {
  if (B <= 0) return MB200_OK;
  if ((hd & 7) || ctx >= capacity) return -EINVAL;
  const int total = (H + Hkv) * (hd / 2) + Hkv * hd / 8;
  dim3 grid((total + 255) / 256, B);
  mb::launch_ex(rope_append_kernel, grid, dim3(256), 0, (cudaStream_t)stream, mb::pdl_mode() != 0, (const bf16*)q,
                (const bf16*)k, (const bf16*)v, (bf16*)q_out, (bf16*)k_cache, (bf16*)v_cache, pos, inv_freq, H, Hkv, hd, ctx,
                (long long)capacity, rope_scale, (const int64_t*)nullptr, 0, 0LL, 0LL);
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

// Paged variant: the new token's K/V go to page table[b][ctx / 128] (+ layer_off elements; V at + v_off elements).
int mb200_rope_append_paged_bf16(const void* q, const void* k, const void* v, void* q_out, const int64_t* table,
                                 int table_stride, long long layer_off, long long v_off, const int64_t* pos,
                                 const float* inv_freq, int B, int H, int Hkv, int hd, int ctx, float rope_scale,
                                 void* stream) {
  if (B <= 0) return MB200_OK;
  if ((hd & 7) || !table || (ctx >> KV_PAGE_SHIFT) >= table_stride) return -EINVAL;
  const int total = (H + Hkv) * (hd / 2) + Hkv * hd / 8;
  dim3 grid((total + 255) / 256, B);
  mb::launch_ex(rope_append_kernel, grid, dim3(256), 0, (cudaStream_t)stream, mb::pdl_mode() != 0, (const bf16*)q,
                (const bf16*)k, (const bf16*)v, (bf16*)q_out, (bf16*)nullptr, (bf16*)nullptr, pos, inv_freq, H, Hkv, hd, ctx,
                0LL, rope_scale, table, table_stride, layer_off, v_off);
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

int mb200_kv_page_tokens(void) { return KV_PAGE; }

// Scatter rows k_lin/v_lin [B, S, row_bytes] (byte strides lin_sb / lin_ss) into the pages at token offsets
// start .. start+S (to_pages = 1), or gather them back (to_pages = 0).  Offsets in BYTES; rows 16-byte multiples.
int mb200_kv_page_copy(void* k_lin, void* v_lin, const int64_t* table, int table_stride, long long layer_off_bytes,
                       long long v_off_bytes, int B, int S, int start, int row_bytes, long long lin_sb, long long lin_ss,
                       int to_pages, void* stream) {
  if (B <= 0 || S <= 0) return MB200_OK;
  if ((row_bytes & 15) || (lin_sb & 15) || (lin_ss & 15) || !table) return -EINVAL;
  if (((start + S - 1) >> KV_PAGE_SHIFT) >= table_stride) return -EINVAL;
  if ((reinterpret_cast<uintptr_t>(k_lin) | reinterpret_cast<uintptr_t>(v_lin)) & 15) return -EINVAL;
  dim3 grid(S, B);
  if (to_pages)
    kv_page_copy_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>((char*)k_lin, (char*)v_lin, table, table_stride,
                                                                     layer_off_bytes, v_off_bytes, start, row_bytes, lin_sb, lin_ss);
  else
    kv_page_copy_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>((char*)k_lin, (char*)v_lin, table, table_stride,
                                                                      layer_off_bytes, v_off_bytes, start, row_bytes, lin_sb, lin_ss);
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

int mb200_decode_attn_splits(int ctx) { int s = (ctx + 127) / 128; if (s < 1) s = 1; if (s > 64) s = 64; return s; }

static int decode_attn_launch(DecP& p, void* o, long long o_sb, long long o_sh, void* stream) {
  const int G = p.H / p.Hkv;
  p.splits = mb200_decode_attn_splits(p.ctx);
  p.chunk = ((p.ctx + p.splits - 1) / p.splits + DKT - 1) / DKT * DKT;      // tile-aligned chunks (see the kernel)
  cudaStream_t st = (cudaStream_t)stream;
  static const int use_mma = [] {         // thread-safe one-time setup of every decode-attention variant
    cudaFuncSetAttribute(decode_attn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec_smem<1>());
    cudaFuncSetAttribute(decode_attn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec_smem<2>());
    cudaFuncSetAttribute(decode_attn_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec_smem<4>());
    cudaFuncSetAttribute(decode_attn_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec_smem<8>());
    cudaFuncSetAttribute(decode_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DM_SMEM);
    const char* e = getenv("MB200_DECODE_ATTN_MMA");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  static const int hk_fast = [] { const char* e = getenv("MB200_DECODE_GRID_HK"); return (e && e[0] == '0') ? 0 : 1; }();
  const bool pdl = mb::pdl_mode() != 0;
  const bool mma = use_mma && G <= 8;
  p.hk_fast = (mma && hk_fast && p.splits <= 65535) ? 1 : 0;
  dim3 grid(p.splits, p.Hkv, p.B);
  if (p.hk_fast) grid = dim3(p.Hkv, p.splits, p.B);
  if (mma) mb::launch_ex(decode_attn_mma_kernel, grid, dim3(DWARPS * 32), DM_SMEM, st, pdl, p, G);
  else if (G == 1) mb::launch_ex(decode_attn_kernel<1>, grid, dim3(DWARPS * 32), dec_smem<1>(), st, pdl, p);
  else if (G == 2) mb::launch_ex(decode_attn_kernel<2>, grid, dim3(DWARPS * 32), dec_smem<2>(), st, pdl, p);
  else if (G == 4) mb::launch_ex(decode_attn_kernel<4>, grid, dim3(DWARPS * 32), dec_smem<4>(), st, pdl, p);
  else if (G == 8) mb::launch_ex(decode_attn_kernel<8>, grid, dim3(DWARPS * 32), dec_smem<8>(), st, pdl, p);
  else return -ENOTSUP;
  mb::launch_ex(decode_combine_kernel, dim3(p.H, p.B), dim3(DHD), 0, st, pdl, (const float*)p.part, (bf16*)o, o_sb, o_sh, p.H,
                p.splits);
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

// q [B,H,128] (strides q_sb,q_sh), cache k/v [B,cap,Hkv,128] (strides kv_sb, kv_ss, kv_sh), o [B,H,128].
// part: fp32 scratch of B*H*splits*(130) floats with splits = mb200_decode_attn_splits(ctx).
int mb200_decode_attn_bf16(const void* q, const void* k, const void* v, void* o, float* part, int B, int H, int Hkv,
                           int ctx, int hd, long long q_sb, long long q_sh, long long kv_sb, long long kv_ss,
                           long long kv_sh, long long o_sb, long long o_sh, float scale, const void* kbits,
                           int kbits_stride, void* stream) {
  if (B <= 0 || ctx <= 0) return MB200_OK;
  if (hd != DHD || H % Hkv != 0) return -ENOTSUP;
  DecP p;
  p.q = (const bf16*)q; p.q_sb = q_sb; p.q_sh = q_sh; p.k = (const bf16*)k; p.v = (const bf16*)v;
  p.kv_sb = kv_sb; p.kv_ss = kv_ss; p.kv_sh = kv_sh; p.kbits = (const uint32_t*)kbits; p.kbits_stride = kbits_stride;
  p.part = part; p.B = B; p.H = H; p.Hkv = Hkv; p.ctx = ctx; p.scale = scale;
  p.table = nullptr; p.table_stride = 0; p.layer_off = 0; p.v_off = 0;
  return decode_attn_launch(p, o, o_sb, o_sh, stream);
}

// Same over the paged cache: keys/values of sequence b, token j at table[b][j / 128] + layer_off + (j % 128) * Hkv * 128
// (+ v_off for V), offsets in elements.
int mb200_decode_attn_paged_bf16(const void* q, const int64_t* table, int table_stride, long long layer_off, long long v_off,
                                 void* o, float* part, int B, int H, int Hkv, int ctx, int hd, long long q_sb,
                                 long long q_sh, long long o_sb, long long o_sh, float scale, const void* kbits,
                                 int kbits_stride, void* stream) {
  if (B <= 0 || ctx <= 0) return MB200_OK;
  if (hd != DHD || H % Hkv != 0) return -ENOTSUP;
  if (!table || ((ctx - 1) >> KV_PAGE_SHIFT) >= table_stride) return -EINVAL;
  DecP p;
  p.q = (const bf16*)q; p.q_sb = q_sb; p.q_sh = q_sh; p.k = nullptr; p.v = nullptr;
  p.kv_sb = 0; p.kv_ss = (long long)Hkv * hd; p.kv_sh = hd; p.kbits = (const uint32_t*)kbits; p.kbits_stride = kbits_stride;
  p.part = part; p.B = B; p.H = H; p.Hkv = Hkv; p.ctx = ctx; p.scale = scale;
  p.table = table; p.table_stride = table_stride; p.layer_off = layer_off; p.v_off = v_off;
  return decode_attn_launch(p, o, o_sb, o_sh, stream);
}

}  // extern "C"

// bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
//   C[M,N] (bf16) = act( op(A) * op(B) + bias[N] ) + addend[M,N]      fp32 accumulation in TMEM
//     transA = 0 : A is [M,K] row-major ("K-major")      transA = 1 : A is [K,M] row-major ("MN-major")
//     transB = 1 : B is [N,K] row-major (nn.Linear W)    transB = 0 : B is [K,N] row-major
// which covers the three products of a linear layer without any transposition pass:
//   forward  y  = x  W^T      (A K-major,  B K-major)
//   dgrad    dx = dy W        (A K-major,  B MN-major)
//   wgrad    dW = dy^T x      (A MN-major, B MN-major)  (+ addend = dW for gradient accumulation)
// This is the kernel behind K3/K5/K9/K13/K16 of SURVEY.md section 2.2 (ViT / projector / LLaMA / LM-head
// linears: transformers llama/modeling_llama.py:171-184,238-249 ; siglip/modeling_siglip.py:270-273,320-321 ;
// mantis/models/mllava/modeling_llava.py:110-118).
//
// Structure: persistent CTAs (one per SM), 128 x BN x 64 tiles, warp-specialised:
//   warp 0    : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1    : MMA issuer    (one thread; 4 x tcgen05.mma 128xBNx16 per stage; tcgen05.commit frees the stage)
//   warps 2-5 : epilogue      (tcgen05.ld 32x32b -> bias/act/addend -> bf16 -> 16 B global stores)
// Two TMEM accumulator stages (2 x BN columns) let the epilogue of tile i overlap the MMAs of tile i+1.
// M/N/K tails are handled by TMA zero-fill on loads and predication on stores.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tmap.cuh"
#include "gemm_epi.cuh"

namespace {
using namespace sm100;
using gemm_epi::GemmEpi;

constexpr int BM = 128, BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr int GROUP_M = 16;                  // rasterisation: 16 m-blocks share each sweep over n for L2 reuse


__device__ __forceinline__ void tile_coords(int t, int num_m, int num_n, int& mb_, int& nb_) {
  const int per_group = GROUP_M * num_n;
  const int g = t / per_group, r = t % per_group;
  const int gm0 = g * GROUP_M;
  const int gsz = min(GROUP_M, num_m - gm0);
  mb_ = gm0 + r % gsz;
  nb_ = r / gsz;
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(320, 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const GemmEpi epi, const int M, const int N, const int K) {
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr uint32_t TMEM_COLS = 2 * BN;     // 256 or 512 (power of two)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int nkb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int mb_, nb_; tile_coords(t, num_m, num_n, mb_, nb_);
        const int m0 = mb_ * BM, n0 = nb_ * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], A_STAGE_BYTES + B_STAGE_BYTES);
          uint8_t* a_dst = sA + s * A_STAGE_BYTES;
          uint8_t* b_dst = sB + s * B_STAGE_BYTES;
          if (!A_MN) tma_load_2d(a_dst, &tmA, &full_bar[s], kb * BK, m0);
          else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(a_dst + c * (64 * BK * 2), &tmA, &full_bar[s], m0 + c * 64, kb * BK);
          }
          if (!B_MN) tma_load_2d(b_dst, &tmB, &full_bar[s], kb * BK, n0);
          else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(b_dst + c * (64 * BK * 2), &tmB, &full_bar[s], n0 + c * 64, kb * BK);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      int s = 0; uint32_t ph = 0; int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + s * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(sB + s * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major : 8-row groups 1024 B apart, +32 B per 16-element K step inside the 128 B swizzle atom
            // MN-major: 64-element MN chunks 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO), +2048 B per K step
            const uint64_t ad = A_MN ? make_smem_desc(a_addr + k * 2048, 64 * BK * 2, 1024)
                                     : make_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bd = B_MN ? make_smem_desc(b_addr + k * 2048, 64 * BK * 2, 1024)
                                     : make_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16_ss(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull_bar[as]);
      }
    }
  } else {
    // 8 epilogue warps: two per TMEM lane quarter, each takes one half of the accumulator's BN columns
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      int mb_, nb_; tile_coords(t, num_m, num_n, mb_, nb_);
      const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
      const int row = mb_ * BM + q * 32 + lane;
      const int n0 = nb_ * BN + half * (BN / 2);
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      gemm_epi::epilogue_rows<BN / 64>(epi, tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + half * (BN / 2), row, row < M,
                                       n0, N);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ------------------------------------------------------------------ host side

template <int BN, int STAGES, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmEpi& epi, int M, int N, int K,
                       cudaStream_t st) {
  constexpr int smem = STAGES * (A_STAGE_BYTES + BN * BK * 2) + 1024 + 256;
  auto kern = gemm_sm100_kernel<BN, STAGES, A_MN, B_MN>;
  static const cudaError_t cfg = cudaFuncSetAttribute(gemm_sm100_kernel<BN, STAGES, A_MN, B_MN>,
                                                      cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // once, thread-safe
  if (cfg != cudaSuccess) { mb200_set_last_error("cudaFuncSetAttribute(max dynamic smem) failed"); return -EIO; }
  const int num_tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = num_tiles < mb::num_sms() ? num_tiles : mb::num_sms();
  kern<<<grid, 320, smem, st>>>(tmA, tmB, epi, M, N, K);
  return 0;
}

template <int BN, int STAGES>
static int dispatch_major(int transA, int transB, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmEpi& epi,
                          int M, int N, int K, cudaStream_t st) {
  const bool a_mn = transA != 0, b_mn = transB == 0;
  if (!a_mn && !b_mn) return launch_gemm<BN, STAGES, false, false>(tmA, tmB, epi, M, N, K, st);
  if (!a_mn && b_mn) return launch_gemm<BN, STAGES, false, true>(tmA, tmB, epi, M, N, K, st);
  if (a_mn && b_mn) return launch_gemm<BN, STAGES, true, true>(tmA, tmB, epi, M, N, K, st);
  return launch_gemm<BN, STAGES, true, false>(tmA, tmB, epi, M, N, K, st);
}
}  // namespace

int mb200_gemm_2cta_impl(const void* A, const void* B, void* C, const void* bias, const void* addend, int M, int N, int K,
                         long long lda, long long ldb, long long ldc, long long ld_add, int transA, int transB, int act,
                         int c_f32, void* stream, const gemm_epi::SwigluArgs* swiglu);

static int gemm_1cta_impl(const void* A, const void* B, void* C, const void* bias, const void* addend, int M, int N, int K,
                          long long lda, long long ldb, long long ldc, long long ld_add, int transA, int transB, int act,
                          int c_f32, void* stream, const gemm_epi::SwigluArgs* swiglu = nullptr) {
  if (M <= 0 || N <= 0) return MB200_OK;
  if (K <= 0) return -EINVAL;
  if ((lda & 7) || (ldb & 7) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
    return -ENOTSUP;
  const int BN = (N > 128) ? 256 : 128;
  CUtensorMap tmA, tmB;
  int rc;
  if (!transA) rc = mbtmap::make_2d(&tmA, A, M, K, lda, BK, BM);        // [M,K]: box 64(K) x 128(M)
  else         rc = mbtmap::make_2d(&tmA, A, K, M, lda, 64, BK);        // [K,M]: box 64(M) x 64(K)
  if (rc) return rc;
  if (transB)  rc = mbtmap::make_2d(&tmB, B, N, K, ldb, BK, BN);        // [N,K]: box 64(K) x BN
  else         rc = mbtmap::make_2d(&tmB, B, K, N, ldb, 64, BK);        // [K,N]: box 64(N) x 64(K)
  if (rc) return rc;
  GemmEpi epi;
  epi.pol_a = epi.pol_b = epi.pol_c = 0;              // (2-CTA kernel only)
  epi.C = C; epi.ldc = ldc; epi.bias = (const bf16*)bias; epi.addend = addend;
  epi.ld_add = ld_add; epi.act = act; epi.c_f32 = c_f32; epi.tma_store = 0; epi.mode = 0; epi.aux0 = nullptr; epi.aux1 = nullptr; epi.ld_aux = 0; epi.C2 = nullptr; epi.ldc2 = 0;
  if (swiglu) { epi.mode = swiglu->mode; epi.aux0 = swiglu->aux0; epi.aux1 = swiglu->aux1; epi.ld_aux = swiglu->ld_aux; epi.C2 = swiglu->C2; epi.ldc2 = swiglu->ldc2; }
  cudaStream_t st = (cudaStream_t)stream;
  if (BN == 256) rc = dispatch_major<256, 4>(transA, transB, tmA, tmB, epi, M, N, K, st);
  else           rc = dispatch_major<128, 6>(transA, transB, tmA, tmB, epi, M, N, K, st);
  if (rc) return rc;
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

extern "C" {

// Returns 0 on success, -ENOTSUP when the operands do not meet the TMA alignment rules (caller then uses
// mb200_gemm_generic), other negative errno on failure.
int mb200_gemm_bf16(const void* A, const void* B, void* C, const void* bias, const void* addend, int M, int N, int K,
                    long long lda, long long ldb, long long ldc, long long ld_add, int transA, int transB, int act,
                    void* stream) {
  return gemm_1cta_impl(A, B, C, bias, addend, M, N, K, lda, ldb, ldc, ld_add, transA, transB, act, 0, stream);
}

// C32[M,N] (fp32, leading dimension ldc floats) = (accumulate ? C32 : 0) + op(A) op(B), bf16 operands, fp32 accumulation
// in TMEM and fp32 all the way into memory: the weight-gradient product dW += dy^T x of the fp32 main-gradient buffer.
int mb200_gemm_bf16_acc32(const void* A, const void* B, float* C32, int M, int N, int K, long long lda, long long ldb,
                          long long ldc, int transA, int transB, int accumulate, void* stream) {
  const void* add = accumulate ? (const void*)C32 : nullptr;
  if (M >= 512 && N >= 512)
    return mb200_gemm_2cta_impl(A, B, C32, nullptr, add, M, N, K, lda, ldb, ldc, ldc, transA, transB, 0, 1, stream, nullptr);
  return gemm_1cta_impl(A, B, C32, nullptr, add, M, N, K, lda, ldb, ldc, ldc, transA, transB, 0, 1, stream);
}

// U[M,N] = X[M,K] Wu[N,K]^T and Act[M,N] = silu(G) * U in the same launch (G, U, Act share the leading dimension ld).
int mb200_gemm_bf16_swiglu_fwd(const void* X, const void* Wu, const void* G, void* U, void* Act, int M, int N, int K,
                               long long lda, long long ldb, long long ld, void* stream) {
  gemm_epi::SwigluArgs sw; sw.mode = 1; sw.aux0 = (const bf16*)G; sw.aux1 = nullptr; sw.ld_aux = ld; sw.C2 = (bf16*)Act; sw.ldc2 = ld;
  if (M >= 512 && N >= 512)
    return mb200_gemm_2cta_impl(X, Wu, U, nullptr, nullptr, M, N, K, lda, ldb, ld, 0, 0, 1, 0, 0, stream, &sw);
  return gemm_1cta_impl(X, Wu, U, nullptr, nullptr, M, N, K, lda, ldb, ld, 0, 0, 1, 0, 0, stream, &sw);
}
// d_act[M,N] = dY[M,K] Wd[K,N] (the down-projection's input gradient, never stored) -> dG, dU[M,N] from G, U in the same launch.
int mb200_gemm_bf16_swiglu_bwd(const void* dY, const void* Wd, const void* G, const void* U, void* dG, void* dU, int M, int N,
                               int K, long long lda, long long ldb, long long ld, void* stream) {
  gemm_epi::SwigluArgs sw; sw.mode = 2; sw.aux0 = (const bf16*)G; sw.aux1 = (const bf16*)U; sw.ld_aux = ld; sw.C2 = (bf16*)dU; sw.ldc2 = ld;
  if (M >= 512 && N >= 512)
    return mb200_gemm_2cta_impl(dY, Wd, dG, nullptr, nullptr, M, N, K, lda, ldb, ld, 0, 0, 0, 0, 0, stream, &sw);
  return gemm_1cta_impl(dY, Wd, dG, nullptr, nullptr, M, N, K, lda, ldb, ld, 0, 0, 0, 0, 0, stream, &sw);
}

}  // extern "C"

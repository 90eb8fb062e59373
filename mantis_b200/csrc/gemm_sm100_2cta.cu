// CTA-pair (cta_group::2) variant of the tcgen05 GEMM: two SMs of one TPC cooperate on a 256 x 256 output tile.
// Each CTA stages its own 128 rows of A and its own 128 columns of B (so the L2->SM operand traffic per FLOP drops by
// one third versus the single-CTA 128x256 tile), the even CTA's MMA thread issues tcgen05.mma.cta_group::2 (M = 256,
// N = 256, K = 16) which reads A from each CTA's shared memory and the two halves of B from both, and accumulates rows
// 0-127 in CTA0's TMEM and rows 128-255 in CTA1's.  Same epilogue / majorness options as gemm_sm100.cu.
#include <stdlib.h>
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tmap.cuh"
#include "gemm_epi.cuh"

namespace {
using namespace sm100;
using gemm_epi::GemmEpi;

constexpr int BM = 128, BK = 64, BN_HALF = 128;          // per CTA; the pair covers 256 x 256
constexpr int A_STAGE = BM * BK * 2;                     // 16 KB
constexpr int B_STAGE = BN_HALF * BK * 2;                // 16 KB
constexpr int STAGES = 6;
// Rasterisation: tiles are handed out m-fastest inside groups of `group_m` tile rows, so a wave of 74 clusters works on one
// A panel (group_m x 256 rows) and a few B column tiles.  The A panel is re-used by every wave of the group while B streams
// past once per GROUP: the host sizes the group so that the panel (group_m x 256 x K x 2 bytes) stays L2-resident (<= 48 MB of
// the 126 MB), which takes the DRAM reads of the gate/up forward GEMM from 3.1x the operand bytes (group of 8) towards 1.6x.

__device__ __forceinline__ void tile_coords(int t, int num_m, int num_n, int group_m, int& mb_, int& nb_) {
  const int per_group = group_m * num_n;
  const int g = t / per_group, r = t % per_group;
  const int gm0 = g * group_m;
  const int gsz = min(group_m, num_m - gm0);
  mb_ = gm0 + r % gsz;
  nb_ = r / gsz;
}

template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
gemm_sm100_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmC, const GemmEpi epi, const int M, const int N, const int K,
                       const int group_m) {
  constexpr uint32_t TMEM_COLS = 512;       // 2 accumulator stages x 256 columns
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* sStage = sB + STAGES * B_STAGE + 1024;   // 8 x 4 KB: one 32-row x 64-column staging tile per epilogue warp (TMA store)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_m = (M + 255) / 256, num_n = (N + 255) / 256;
  const int num_tiles = num_m * num_n;
  const int nkb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    if (epi.tma_store) prefetch_tmap(&tmC);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 16); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        int mb_, nb_; tile_coords(t, num_m, num_n, group_m, mb_, nb_);
        const int m0 = mb_ * 256 + (int)rank * BM, n0 = nb_ * 256 + (int)rank * BN_HALF;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * (A_STAGE + B_STAGE));
          uint8_t* a_dst = sA + s * A_STAGE;
          uint8_t* b_dst = sB + s * B_STAGE;
          if (!A_MN) tma_load_2d_2cta_hint(a_dst, &tmA, &full_bar[s], kb * BK, m0, epi.pol_a);
          else {
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_2d_2cta_hint(a_dst + c * 8192, &tmA, &full_bar[s], m0 + c * 64, kb * BK, epi.pol_a);
          }
          if (!B_MN) tma_load_2d_2cta_hint(b_dst, &tmB, &full_bar[s], kb * BK, n0, epi.pol_b);
          else {
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_2d_2cta_hint(b_dst + c * 8192, &tmB, &full_bar[s], n0 + c * 64, kb * BK, epi.pol_b);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, 256, A_MN, B_MN);
      int s = 0; uint32_t ph = 0; int it = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
        const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + s * A_STAGE);
          const uint32_t b_addr = smem_u32(sB + s * B_STAGE);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = A_MN ? make_smem_desc(a_addr + k * 2048, 8192, 1024) : make_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bd = B_MN ? make_smem_desc(b_addr + k * 2048, 8192, 1024) : make_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16_ss_2cta(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2cta(&empty_bar[s], 3);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit_2cta(&tfull_bar[as], 3);
      }
    }
  } else {
    // 8 epilogue warps: two per TMEM lane quarter, each takes one 128-column half of the CTA's 128 x 256 accumulator
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    int it = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
      int mb_, nb_; tile_coords(t, num_m, num_n, group_m, mb_, nb_);
      const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
      const int row = mb_ * 256 + (int)rank * BM + q * 32 + lane;
      const int n0 = nb_ * 256 + half * 128;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256 + half * 128;
      if (epi.tma_store)
        gemm_epi::epilogue_rows_tma<4>(epi, &tmC, sStage + (warp - 2) * 4096, taddr, row, row < M, n0, N, lane);
      else
        gemm_epi::epilogue_rows<4>(epi, taddr, row, row < M, n0, N);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
    }
  }
  if (warp >= 2 && lane == 0 && epi.tma_store) tma_store_wait_all0();     // outstanding bulk stores complete before the CTA exits
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc_2cta(tmem_base, TMEM_COLS); }
}


template <bool A_MN, bool B_MN>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmEpi& epi, int M, int N, int K,
                  cudaStream_t st) {
  constexpr int smem = STAGES * (A_STAGE + B_STAGE) + 1024 /*barriers*/ + 8 * 4096 /*TMA-store staging*/ + 1024 /*alignment*/;
  auto kern = gemm_sm100_2cta_kernel<A_MN, B_MN>;
  // function-local static with a dynamic initialiser: initialised exactly once even when several host threads race here
  // (generate() on a worker thread, mantis/models/mllava/utils.py:100-186)
  static const cudaError_t cfg = cudaFuncSetAttribute(gemm_sm100_2cta_kernel<A_MN, B_MN>,
                                                      cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (cfg != cudaSuccess) { mb200_set_last_error("cudaFuncSetAttribute(max dynamic smem) failed"); return -EIO; }
  const int num_tiles = ((M + 255) / 256) * ((N + 255) / 256);
  int clusters = mb::num_sms() / 2;
  if (num_tiles < clusters) clusters = num_tiles;
  // group height: A panel <= 48 MB, at least 8 tile rows, groups of (almost) equal height
  static const int group_env = [] { const char* e = getenv("MB200_GEMM_GROUP_M"); return e ? atoi(e) : 0; }();
  const int num_m = (M + 255) / 256;
  int gmax = (int)((48ll << 20) / (256ll * 2 * (K > 0 ? K : 1)));
  if (gmax < 8) gmax = 8;
  if (group_env > 0) gmax = group_env;
  const int ngroups = (num_m + gmax - 1) / gmax;
  const int group_m = (num_m + ngroups - 1) / (ngroups > 0 ? ngroups : 1);
  kern<<<clusters * 2, 320, smem, st>>>(tmA, tmB, tmC, epi, M, N, K, group_m > 0 ? group_m : 1);
  return 0;
}
}  // namespace

// shared by mb200_gemm_bf16_2cta (bf16 C) and mb200_gemm_bf16_acc32 (fp32 C); not part of the public header
int mb200_gemm_2cta_impl(const void* A, const void* B, void* C, const void* bias, const void* addend, int M, int N, int K,
                         long long lda, long long ldb, long long ldc, long long ld_add, int transA, int transB, int act,
                         int c_f32, void* stream, const gemm_epi::SwigluArgs* swiglu) {
  if (M <= 0 || N <= 0) return MB200_OK;
  if (K <= 0) return -EINVAL;
  if ((lda & 7) || (ldb & 7) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
    return -ENOTSUP;
  CUtensorMap tmA, tmB;
  int rc;
  if (!transA) rc = mbtmap::make_2d(&tmA, A, M, K, lda, BK, BM); else rc = mbtmap::make_2d(&tmA, A, K, M, lda, 64, BK);
  if (rc) return rc;
  if (transB) rc = mbtmap::make_2d(&tmB, B, N, K, ldb, BK, BN_HALF); else rc = mbtmap::make_2d(&tmB, B, K, N, ldb, 64, BK);
  if (rc) return rc;
  GemmEpi epi;
  epi.C = C; epi.ldc = ldc; epi.bias = (const bf16*)bias; epi.addend = addend; epi.ld_add = ld_add; epi.act = act;
  epi.c_f32 = c_f32; epi.mode = 0; epi.aux0 = nullptr; epi.aux1 = nullptr; epi.ld_aux = 0; epi.C2 = nullptr; epi.ldc2 = 0;
  if (swiglu) { epi.mode = swiglu->mode; epi.aux0 = swiglu->aux0; epi.aux1 = swiglu->aux1; epi.ld_aux = swiglu->ld_aux; epi.C2 = swiglu->C2; epi.ldc2 = swiglu->ldc2; }
  // TMA epilogue (store, or reduce-add for the fp32 gradient accumulation) whenever the alignment rules allow it; the SwiGLU
  // modes (two outputs) and odd alignments keep the per-thread stores
  static const int tma_store_on = [] { const char* e = getenv("MB200_GEMM_TMA_STORE"); return (e && e[0] == '0') ? 0 : 1; }();
  CUtensorMap tmC = tmA;
  epi.tma_store = 0;
  if (tma_store_on && !swiglu && !(reinterpret_cast<uintptr_t>(C) & 15)) {
    if (c_f32) {
      if (!(ldc & 3) && (!addend || addend == C)) {
        if ((rc = mbtmap::make_2d_store(&tmC, C, M, N, ldc, 32, 32, true))) return rc;
        epi.tma_store = 1;
      }
    } else if (!(ldc & 7) && (!addend || (!(ld_add & 7) && !(reinterpret_cast<uintptr_t>(addend) & 15)))) {
      if ((rc = mbtmap::make_2d_store(&tmC, C, M, N, ldc, 64, 32))) return rc;
      epi.tma_store = 1;
    }
  }
  // L2 eviction priority: the A panel of a raster group is what every wave of the group re-reads, so its loads ask to stay
  // (evict_last).  Measured under ncu (profiles/ab_r02.md): DRAM reads 366 -> 361 MB and -1 % time on the K = 14336 shapes;
  // marking B evict_first doubles the reads (its tiles are shared by CTAs a few microseconds apart), C evict_first is neutral.
  static const int l2_hint = [] { const char* e = getenv("MB200_GEMM_L2_HINT"); return e ? atoi(e) : 1; }();
  epi.pol_a = l2_hint >= 1 ? sm100::kL2EvictLast : sm100::kL2EvictNormal;
  epi.pol_b = l2_hint >= 2 ? sm100::kL2EvictFirst : sm100::kL2EvictNormal;
  epi.pol_c = l2_hint >= 3 ? sm100::kL2EvictFirst : sm100::kL2EvictNormal;
  cudaStream_t st = (cudaStream_t)stream;
  const bool a_mn = transA != 0, b_mn = transB == 0;
  if (!a_mn && !b_mn) rc = launch<false, false>(tmA, tmB, tmC, epi, M, N, K, st);
  else if (!a_mn && b_mn) rc = launch<false, true>(tmA, tmB, tmC, epi, M, N, K, st);
  else if (a_mn && b_mn) rc = launch<true, true>(tmA, tmB, tmC, epi, M, N, K, st);
  else rc = launch<true, false>(tmA, tmB, tmC, epi, M, N, K, st);
  if (rc) return rc;
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

extern "C" int mb200_gemm_bf16_2cta(const void* A, const void* B, void* C, const void* bias, const void* addend, int M,
                                    int N, int K, long long lda, long long ldb, long long ldc, long long ld_add,
                                    int transA, int transB, int act, void* stream) {
  return mb200_gemm_2cta_impl(A, B, C, bias, addend, M, N, K, lda, ldb, ldc, ld_add, transA, transB, act, 0, stream, nullptr);
}

// Flash-attention backward on tcgen05 (head_dim 128), as two kernels that share the forward's building blocks
// (TMA -> 128B-swizzled smem, tcgen05.mma with TMEM accumulators, one softmax thread per TMEM lane):
//
//   dKV kernel : CTA owns a 128-key tile of one (batch, kv head); loops over the query heads of the GQA group and
//                over 64-query sub-tiles.  S^T = K Q^T and dP^T = V dO^T (128x64, two TMEM stages), P^T / dS^T are
//                written back into those TMEM columns (bf16) and consumed as TMEM A operands: dV += P^T dO,
//                dK += dS^T Q accumulate in TMEM over the whole loop.
//   dQ kernel  : CTA owns a 128-query tile of one (batch, head); loops over 64-key sub-tiles.
//                S = Q K^T, dP = dO V^T (two TMEM stages), dS -> TMEM (bf16), dQ += dS K accumulates in TMEM.
//   Two softmax warp groups ping-pong over the sub-tiles (one TMEM stage each).
//
// No atomics, deterministic.  S/dP are recomputed in both kernels (7 GEMMs instead of the fused 5) -- the price of
// keeping every accumulator resident in the 512 TMEM columns without a global dQ reduction.
// The same [rows x 64] swizzled tile serves as a K-major operand (rows = M/N) and as an MN-major operand
// (rows = K) -- only the descriptor differs -- so Q, dO, K, V are each loaded once per use.
// Backward of LlamaAttention's SDPA/flash call (transformers llama/modeling_llama.py:199-289).
#include <stdlib.h>
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tmap.cuh"

namespace {
using namespace sm100;

constexpr int HD = 128;
constexpr int FULL_HALF = 128 * 128;     // [128 rows x 64 bf16] swizzled sub-tile (16 KB)
constexpr int FULL_TILE = 2 * FULL_HALF; // [128 rows x 128 hd]
constexpr int SUB_HALF = 64 * 128;       // [64 rows x 64 bf16] (8 KB)
constexpr int SUB_TILE = 2 * SUB_HALF;   // [64 rows x 128 hd] (16 KB)
constexpr float LOG2E = 1.44269504088896340736f;

constexpr int QS = 3;                    // shared-memory stages of the streamed operand tiles (TMEM stages stay 2)
constexpr int SM_WARPS = 8;              // softmax warps (two per TMEM lane quarter, each takes half of the columns)
constexpr int SM_THREADS = SM_WARPS * 32;
constexpr int NTHREADS = 64 + SM_THREADS;

struct BwdParams {
  const float2* ld;      // [B,H,Sq_pad] {lse * log2(e) (+inf for dead / padded rows), delta}
  int Sq_pad;
  bf16* dq; long long dq_sb, dq_ss, dq_sh;
  bf16* dk; long long dk_sb, dk_ss, dk_sh;
  bf16* dv; long long dv_sb, dv_ss, dv_sh;
  const uint32_t* kbits; int kbits_stride;
  int B, H, Hkv, Sq, Sk;
  float scale, scale_log2;
  int causal;
  // single-pass mode: the dK/dV kernel also writes dS^T (bf16) for every (key, query) pair it visits into
  // ds_t[B*H][Sk_pad][Sq_pad]; the dQ kernel then is a plain GEMM over that buffer (S and dP are computed, and read back from
  // TMEM, ONCE instead of once per kernel)
  bf16* ds_t; long long ds_row, ds_head;      // element strides: key row, (batch, head) slab
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  bf162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h);
}
// 1-D bulk copy global -> shared, completion credited to an mbarrier (bytes % 16 == 0, 16-byte aligned)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// store one 128-lane x 128-column fp32 accumulator row as bf16 (256 B) to global
__device__ __forceinline__ void store_acc_row(uint32_t taddr, bf16* dst, bool ok, float mul) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t ov[32];
    tmem_ld_32x32b_x32(taddr + c * 32, ov);
    tmem_ld_wait();
    if (ok) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 o4;
        o4.x = pack_bf16(__uint_as_float(ov[g * 8 + 0]) * mul, __uint_as_float(ov[g * 8 + 1]) * mul);
        o4.y = pack_bf16(__uint_as_float(ov[g * 8 + 2]) * mul, __uint_as_float(ov[g * 8 + 3]) * mul);
        o4.z = pack_bf16(__uint_as_float(ov[g * 8 + 4]) * mul, __uint_as_float(ov[g * 8 + 5]) * mul);
        o4.w = pack_bf16(__uint_as_float(ov[g * 8 + 6]) * mul, __uint_as_float(ov[g * 8 + 7]) * mul);
        *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = o4;
      }
    }
  }
}

// ============================================================================================ dK / dV
// Softmax warps form two groups (warps 2-5 and 6-9) that ping-pong over the 64-query sub-tiles: group g owns TMEM stage g
// (S^T_g, dP^T_g), so while one group waits on barriers / TMEM latency the other computes, and the MMA thread always has
// the other stage's products to issue.  P^T / dS^T (bf16) are written back into the first 32 columns of S^T_g / dP^T_g and
// feed the dV / dK MMAs as TMEM A operands: nothing but Q / dO / K / V tiles ever touches shared memory.
__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_dkv_sm100_kernel(const __grid_constant__ CUtensorMap tmQ64, const __grid_constant__ CUtensorMap tmDO64,
                          const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                          const __grid_constant__ CUtensorMap tmDSst, const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;                           // 32 KB resident
  uint8_t* sV = sK + FULL_TILE;                 // 32 KB resident
  uint8_t* sQ = sV + FULL_TILE;                 // QS stages x 16 KB
  uint8_t* sDO = sQ + QS * SUB_TILE;            // QS stages x 16 KB
  float2* sLD = reinterpret_cast<float2*>(sDO + QS * SUB_TILE);   // [QS][64] {lse2, delta}
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLD + QS * 64);
  // single-pass mode: one 4 KB staging tile per softmax warp ([32 key rows][64 queries] bf16, 128B-swizzled) for the TMA
  // store of dS^T; placed after the barrier block on the next 1 KB boundary
  uint8_t* sDS = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bars) + 256 + 1023) & ~uintptr_t(1023));
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;    // [QS]
  uint64_t* qdo_empty = bars + 4;   // [QS]
  uint64_t* sdp_full = bars + 7;    // [2]
  uint64_t* pds_full = bars + 9;    // [2]
  uint64_t* acc_done = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const mb::LptIdx li = mb::lpt_index();        // early key tiles see the most queries: they go first, across all kv heads
  const int kt = li.rank, hk = li.h, b = li.b;
  const int G = p.H / p.Hkv;
  const int k0 = kt * 128;
  const int off = p.Sk - p.Sq;
  const int n_qs = (p.Sq + 63) / 64;
  int qs_begin = 0;
  if (p.causal) { int qb = k0 - off; if (qb < 0) qb = 0; qs_begin = qb / 64; }
  const int per_head = (n_qs > qs_begin) ? (n_qs - qs_begin) : 0;
  const int n_it = per_head * G;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ64); prefetch_tmap(&tmDO64); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(kv_full, 1);
    for (int s = 0; s < QS; ++s) { mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&sdp_full[s], 1); mbar_init(&pds_full[s], 128); }
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tDK = tmem_base, tDV = tmem_base + 128;
  const uint32_t tST[2] = {tmem_base + 256, tmem_base + 384};
  const uint32_t tDPT[2] = {tmem_base + 320, tmem_base + 448};

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * FULL_TILE);
      tma_load_4d(sK, &tmK, kv_full, 0, hk, k0, b);
      tma_load_4d(sK + FULL_HALF, &tmK, kv_full, 64, hk, k0, b);
      tma_load_4d(sV, &tmV, kv_full, 0, hk, k0, b);
      tma_load_4d(sV + FULL_HALF, &tmV, kv_full, 64, hk, k0, b);
      for (int n = 0; n < n_it; ++n) {
        const int s = n % QS; const uint32_t ph = (n / QS) & 1;
        const int h = hk * G + n / per_head, qs = qs_begin + n % per_head;
        mbar_wait(&qdo_empty[s], ph ^ 1);      // dV/dK of sub-tile n-QS done => Q/dO stage and sLD[s] are free
        mbar_arrive_expect_tx(&qdo_full[s], 2 * SUB_TILE + 64 * 8);
        tma_load_4d(sQ + s * SUB_TILE, &tmQ64, &qdo_full[s], 0, h, qs * 64, b);
        tma_load_4d(sQ + s * SUB_TILE + SUB_HALF, &tmQ64, &qdo_full[s], 64, h, qs * 64, b);
        tma_load_4d(sDO + s * SUB_TILE, &tmDO64, &qdo_full[s], 0, h, qs * 64, b);
        tma_load_4d(sDO + s * SUB_TILE + SUB_HALF, &tmDO64, &qdo_full[s], 64, h, qs * 64, b);
        bulk_load_1d(sLD + s * 64, p.ld + ((size_t)b * p.H + h) * p.Sq_pad + qs * 64, 64 * 8, &qdo_full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_it > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, 128, false, true);
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      auto issue_sdp = [&](int n) {             // TMEM stage n&1 was last read by dV/dK(n-2), issued earlier (in-order tensor pipe)
        const int s = n & 1, ss = n % QS;
        mbar_wait(&qdo_full[ss], (n / QS) & 1);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(sQ + ss * SUB_TILE), do_addr = smem_u32(sDO + ss * SUB_TILE);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t oa = (kk >> 2) * FULL_HALF + (kk & 3) * 32, ob = (kk >> 2) * SUB_HALF + (kk & 3) * 32;
          umma_bf16_ss(tST[s], make_smem_desc(k_addr + oa, 16, 1024), make_smem_desc(q_addr + ob, 16, 1024), idesc_s, kk != 0);
        }
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t oa = (kk >> 2) * FULL_HALF + (kk & 3) * 32, ob = (kk >> 2) * SUB_HALF + (kk & 3) * 32;
          umma_bf16_ss(tDPT[s], make_smem_desc(v_addr + oa, 16, 1024), make_smem_desc(do_addr + ob, 16, 1024), idesc_s, kk != 0);
        }
        umma_commit(&sdp_full[s]);
      };
      mbar_wait(kv_full, 0);
      issue_sdp(0);
      if (n_it > 1) issue_sdp(1);
      for (int n = 0; n < n_it; ++n) {
        const int s = n & 1, ss = n % QS; const uint32_t ph = (n >> 1) & 1;
        mbar_wait(&pds_full[s], ph);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(sQ + ss * SUB_TILE), do_addr = smem_u32(sDO + ss * SUB_TILE);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)     // dV += P^T (TMEM A, K = 64 queries) x dO (MN-major: rows = queries)
          umma_bf16_ts(tDV, tST[s] + kk * 8, make_smem_desc(do_addr + kk * 2048, SUB_HALF, 1024), idesc_acc, (n | kk) != 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)     // dK += dS^T (TMEM A) x Q
          umma_bf16_ts(tDK, tDPT[s] + kk * 8, make_smem_desc(q_addr + kk * 2048, SUB_HALF, 1024), idesc_acc, (n | kk) != 0);
        umma_commit(&qdo_empty[ss]);
        if (n + 2 < n_it) issue_sdp(n + 2);
      }
      umma_commit(acc_done);
    }
  } else {
    const int qd = warp & 3;                     // TMEM lane quarter
    const int grp = (warp - 2) >> 2;             // ping-pong group == TMEM stage it owns
    const int r = qd * 32 + lane;                // key row in tile == TMEM lane
    const int kj = k0 + r;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    bool key_ok = kj < p.Sk;
    if (key_ok && p.kbits) key_ok = (__ldg(p.kbits + (size_t)b * p.kbits_stride + (kj >> 5)) >> (kj & 31)) & 1u;
    const int qlim = kj - off;                   // causal: query qi sees key kj iff qi >= kj - off
    const int s = grp;
    for (int n = grp; n < n_it; n += 2) {
      const uint32_t ph = (n >> 1) & 1;
      const int qs = qs_begin + n % per_head;
      const int ss = n % QS;
      mbar_wait(&qdo_full[ss], (n / QS) & 1);    // lse/delta landed (bulk copy on the same barrier as Q/dO)
      mbar_wait(&sdp_full[s], ph);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {              // two chunks of 32 query columns
        const int q0 = qs * 64 + c * 32;
        float sv[32], dp[32];
        tmem_ld_32x32b_x32(tST[s] + lane_off + c * 32, reinterpret_cast<uint32_t*>(sv));
        tmem_ld_32x32b_x32(tDPT[s] + lane_off + c * 32, reinterpret_cast<uint32_t*>(dp));
        tmem_ld_wait();
        const float2* ldp = sLD + ss * 64 + c * 32;
        uint32_t pk[16], dk_[16];
        const bool full_vis = key_ok && (!p.causal || q0 >= qlim);
        if (full_vis) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 l0 = ldp[2 * i], l1 = ldp[2 * i + 1];
            const float p0 = fast_exp2(fmaf(sv[2 * i], p.scale_log2, -l0.x));
            const float p1 = fast_exp2(fmaf(sv[2 * i + 1], p.scale_log2, -l1.x));
            pk[i] = pack_bf16(p0, p1);
            dk_[i] = pack_bf16(p0 * (dp[2 * i] - l0.y), p1 * (dp[2 * i + 1] - l1.y));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 l0 = ldp[2 * i], l1 = ldp[2 * i + 1];
            const bool v0 = key_ok && (!p.causal || q0 + 2 * i >= qlim);
            const bool v1 = key_ok && (!p.causal || q0 + 2 * i + 1 >= qlim);
            const float p0 = v0 ? fast_exp2(fmaf(sv[2 * i], p.scale_log2, -l0.x)) : 0.f;
            const float p1 = v1 ? fast_exp2(fmaf(sv[2 * i + 1], p.scale_log2, -l1.x)) : 0.f;
            pk[i] = pack_bf16(p0, p1);
            dk_[i] = pack_bf16(p0 * (dp[2 * i] - l0.y), p1 * (dp[2 * i + 1] - l1.y));
          }
        }
        // packed columns [16c, 16c+16) lie inside the fp32 columns this thread has already consumed (lane-private)
        tmem_st_32x32b_x16(tST[s] + lane_off + c * 16, pk);
        tmem_st_32x32b_x16(tDPT[s] + lane_off + c * 16, dk_);
        if (p.ds_t) {
          // this key row's dS for 32 queries = 64 bytes of its row in the warp's staging tile (chunk j of row r at
          // r * 128 + ((j ^ (r & 7)) << 4): conflict-free); rows / columns beyond Sk / Sq hold exact zeros (masked above)
          uint8_t* stg = sDS + (warp - 2) * 4096 + lane * 128;
          if (c == 0) {
            if (lane == 0) tma_store_wait_read0();         // the previous sub-tile's store has finished reading the tile
            __syncwarp();
          }
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(stg + (((c * 4 + g) ^ (lane & 7)) << 4)) =
                make_uint4(dk_[g * 4], dk_[g * 4 + 1], dk_[g * 4 + 2], dk_[g * 4 + 3]);
        }
      }
      if (p.ds_t) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {                                     // 32 key rows x 64 queries of dS^T, written by the TMA unit
          tma_store_3d(&tmDSst, sDS + (warp - 2) * 4096, qs * 64, k0 + qd * 32, b * p.H + hk * G + n / per_head);
          tma_store_commit();
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&pds_full[s]);
    }
    // ---- epilogue: group 0 stores dK (x softmax scale), group 1 stores dV ----
    const bool row_ok = kj < p.Sk;
    bf16* dkp = p.dk + (size_t)b * p.dk_sb + (size_t)(row_ok ? kj : 0) * p.dk_ss + (size_t)hk * p.dk_sh;
    bf16* dvp = p.dv + (size_t)b * p.dv_sb + (size_t)(row_ok ? kj : 0) * p.dv_ss + (size_t)hk * p.dv_sh;
    if (n_it > 0) {
      mbar_wait(acc_done, 0);
      tc_fence_after();
      if (grp == 0) store_acc_row(tDK + lane_off, dkp, row_ok, p.scale);
      else          store_acc_row(tDV + lane_off, dvp, row_ok, 1.f);
    } else if (row_ok) {
      const uint4 z = make_uint4(0, 0, 0, 0);
      bf16* dst = grp == 0 ? dkp : dvp;
#pragma unroll
      for (int c = 0; c < 16; ++c) *reinterpret_cast<uint4*>(dst + c * 8) = z;
    }
  }
  if (p.ds_t && warp >= 2 && lane == 0) tma_store_wait_all0();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ============================================================================================ dQ
// NGQ softmax groups rotate over the key sub-tiles, one TMEM stage (S, dP) each: 128 (dQ) + 3 x 128 = 512 columns.
constexpr int NGQ = 3;
constexpr int QSQ = 4;                   // smem stages of the streamed K/V sub-tiles
constexpr int DQ_THREADS = 64 + NGQ * 128;

__global__ void __launch_bounds__(DQ_THREADS, 1)
attn_bwd_dq_sm100_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                         const __grid_constant__ CUtensorMap tmK64, const __grid_constant__ CUtensorMap tmV64,
                         const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                           // 32 KB resident
  uint8_t* sDO = sQ + FULL_TILE;                // 32 KB resident
  uint8_t* sK = sDO + FULL_TILE;                // QSQ x 16 KB
  uint8_t* sV = sK + QSQ * SUB_TILE;            // QSQ x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + QSQ * SUB_TILE);
  uint64_t* qdo_full = bars + 0;
  uint64_t* kv_full = bars + 1;              // [QSQ]
  uint64_t* kv_empty = kv_full + QSQ;        // [QSQ]
  uint64_t* sdp_full = kv_empty + QSQ;       // [NGQ]
  uint64_t* ds_full = sdp_full + NGQ;        // [NGQ]
  uint64_t* acc_done = ds_full + NGQ;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const mb::LptIdx li = mb::lpt_index();        // heavy (late) query tiles first, across all heads
  const int qt = (int)gridDim.x - 1 - li.rank;
  const int h = li.h, b = li.b;
  const int hk = h / (p.H / p.Hkv);
  const int q0 = qt * 128;
  const int off = p.Sk - p.Sq;
  int kv_end = p.Sk;
  if (p.causal) { kv_end = min(p.Sk, q0 + 128 + off); if (kv_end < 0) kv_end = 0; }
  const int n_it = (kv_end + 63) / 64;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmDO); prefetch_tmap(&tmK64); prefetch_tmap(&tmV64);
    mbar_init(qdo_full, 1);
    for (int s = 0; s < QSQ; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < NGQ; ++s) { mbar_init(&sdp_full[s], 1); mbar_init(&ds_full[s], 128); }
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tDQ = tmem_base;
  const uint32_t tS0 = tmem_base + 128, tDP0 = tmem_base + 192;      // stage g: +128 g

  if (warp == 0) {
    if (lane == 0 && n_it > 0) {
      mbar_arrive_expect_tx(qdo_full, 2 * FULL_TILE);
      tma_load_4d(sQ, &tmQ, qdo_full, 0, h, q0, b);
      tma_load_4d(sQ + FULL_HALF, &tmQ, qdo_full, 64, h, q0, b);
      tma_load_4d(sDO, &tmDO, qdo_full, 0, h, q0, b);
      tma_load_4d(sDO + FULL_HALF, &tmDO, qdo_full, 64, h, q0, b);
      for (int n = 0; n < n_it; ++n) {
        const int s = n % QSQ; const uint32_t ph = (n / QSQ) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * SUB_TILE);
        tma_load_4d(sK + s * SUB_TILE, &tmK64, &kv_full[s], 0, hk, n * 64, b);
        tma_load_4d(sK + s * SUB_TILE + SUB_HALF, &tmK64, &kv_full[s], 64, hk, n * 64, b);
        tma_load_4d(sV + s * SUB_TILE, &tmV64, &kv_full[s], 0, hk, n * 64, b);
        tma_load_4d(sV + s * SUB_TILE + SUB_HALF, &tmV64, &kv_full[s], 64, hk, n * 64, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_it > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, 128, false, true);
      const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sDO);
      auto issue_sdp = [&](int n) {
        const int s = n % NGQ, ss = n % QSQ;
        mbar_wait(&kv_full[ss], (n / QSQ) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + ss * SUB_TILE), v_addr = smem_u32(sV + ss * SUB_TILE);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t oa = (kk >> 2) * FULL_HALF + (kk & 3) * 32, ob = (kk >> 2) * SUB_HALF + (kk & 3) * 32;
          umma_bf16_ss(tS0 + s * 128, make_smem_desc(q_addr + oa, 16, 1024), make_smem_desc(k_addr + ob, 16, 1024), idesc_s, kk != 0);
        }
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t oa = (kk >> 2) * FULL_HALF + (kk & 3) * 32, ob = (kk >> 2) * SUB_HALF + (kk & 3) * 32;
          umma_bf16_ss(tDP0 + s * 128, make_smem_desc(do_addr + oa, 16, 1024), make_smem_desc(v_addr + ob, 16, 1024), idesc_s, kk != 0);
        }
        umma_commit(&sdp_full[s]);
      };
      mbar_wait(qdo_full, 0);
      for (int n = 0; n < NGQ && n < n_it; ++n) issue_sdp(n);
      for (int n = 0; n < n_it; ++n) {
        const int s = n % NGQ, ss = n % QSQ; const uint32_t ph = (n / NGQ) & 1;
        mbar_wait(&ds_full[s], ph);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + ss * SUB_TILE);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)     // dQ += dS (TMEM A, K = 64 keys) x K (MN-major: rows = keys)
          umma_bf16_ts(tDQ, tS0 + s * 128 + kk * 8, make_smem_desc(k_addr + kk * 2048, SUB_HALF, 1024), idesc_acc, (n | kk) != 0);
        umma_commit(&kv_empty[ss]);
        if (n + NGQ < n_it) issue_sdp(n + NGQ);
      }
      umma_commit(acc_done);
    }
  } else {
    const int qd = warp & 3;
    const int grp = (warp - 2) >> 2;             // ping-pong group == stage
    const int r = qd * 32 + lane;
    const int qi = q0 + r;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const bool row_ok = qi < p.Sq;
    const float2 ldv = p.ld[((size_t)b * p.H + h) * p.Sq_pad + qi];        // padded rows hold {+inf, 0}
    const float L = ldv.x, dl = ldv.y;
    const int limit = p.causal ? min(qi + off, p.Sk - 1) : (p.Sk - 1);
    const int s = grp;
    for (int n = grp; n < n_it; n += NGQ) {
      const uint32_t ph = (n / NGQ) & 1;
      mbar_wait(&sdp_full[s], ph);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int k0 = n * 64 + c * 32;
        float sv[32], dp[32];
        tmem_ld_32x32b_x32(tS0 + s * 128 + lane_off + c * 32, reinterpret_cast<uint32_t*>(sv));
        tmem_ld_32x32b_x32(tDP0 + s * 128 + lane_off + c * 32, reinterpret_cast<uint32_t*>(dp));
        tmem_ld_wait();
        uint32_t w = 0xffffffffu;
        if (p.kbits) { const int wi = k0 >> 5; w = (wi < p.kbits_stride) ? __ldg(p.kbits + (size_t)b * p.kbits_stride + wi) : 0u; }
        uint32_t dsk[16];
        if (w == 0xffffffffu && k0 + 31 <= limit) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = fast_exp2(fmaf(sv[2 * i], p.scale_log2, -L));
            const float p1 = fast_exp2(fmaf(sv[2 * i + 1], p.scale_log2, -L));
            dsk[i] = pack_bf16(p0 * (dp[2 * i] - dl), p1 * (dp[2 * i + 1] - dl));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const bool v0 = (k0 + 2 * i <= limit) && ((w >> (2 * i)) & 1u);
            const bool v1 = (k0 + 2 * i + 1 <= limit) && ((w >> (2 * i + 1)) & 1u);
            const float p0 = v0 ? fast_exp2(fmaf(sv[2 * i], p.scale_log2, -L)) : 0.f;
            const float p1 = v1 ? fast_exp2(fmaf(sv[2 * i + 1], p.scale_log2, -L)) : 0.f;
            dsk[i] = pack_bf16(p0 * (dp[2 * i] - dl), p1 * (dp[2 * i + 1] - dl));
          }
        }
        tmem_st_32x32b_x16(tS0 + s * 128 + lane_off + c * 16, dsk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&ds_full[s]);
    }
    // epilogue: groups 0 and 1 each store 64 of the 128 head dims of their dQ row (x softmax scale); group 2 is done
    bf16* dqp = p.dq + (size_t)b * p.dq_sb + (size_t)(row_ok ? qi : 0) * p.dq_ss + (size_t)h * p.dq_sh + grp * 64;
    if (grp >= 2) {
      // nothing to store
    } else if (n_it > 0) {
      mbar_wait(acc_done, 0);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t ov[32];
        tmem_ld_32x32b_x32(tDQ + lane_off + grp * 64 + c * 32, ov);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o4;
            o4.x = pack_bf16(__uint_as_float(ov[g * 8 + 0]) * p.scale, __uint_as_float(ov[g * 8 + 1]) * p.scale);
            o4.y = pack_bf16(__uint_as_float(ov[g * 8 + 2]) * p.scale, __uint_as_float(ov[g * 8 + 3]) * p.scale);
            o4.z = pack_bf16(__uint_as_float(ov[g * 8 + 4]) * p.scale, __uint_as_float(ov[g * 8 + 5]) * p.scale);
            o4.w = pack_bf16(__uint_as_float(ov[g * 8 + 6]) * p.scale, __uint_as_float(ov[g * 8 + 7]) * p.scale);
            *reinterpret_cast<uint4*>(dqp + c * 32 + g * 8) = o4;
          }
        }
      }
    } else if (row_ok) {
      const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(dqp + c * 8) = z;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ============================================================================================ dQ from dS^T (single-pass mode)
// dQ[128 queries x 128] = scale * sum over key sub-tiles of dS[128 q x 64 k] . K[64 k x 128]: a GEMM whose A operand is the
// dS^T buffer the dK/dV kernel wrote ([keys][queries] => MN-major A, fetched by TMA as two 64-query x 64-key boxes) and whose B
// operand is the K sub-tile ([keys][hd] => MN-major B).  No S, no dP, no softmax warps, no TMEM reads besides the epilogue.
constexpr int DQ2_STAGES = 6;
constexpr int DQ2_THREADS = 64 + 128;

__global__ void __launch_bounds__(DQ2_THREADS, 1)
attn_bwd_dq2_sm100_kernel(const __grid_constant__ CUtensorMap tmDS, const __grid_constant__ CUtensorMap tmK64, const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                               // DQ2_STAGES x 16 KB: dS^T  [2 x 64 queries][64 keys][128 B]
  uint8_t* sB = sA + DQ2_STAGES * SUB_TILE;         // DQ2_STAGES x 16 KB: K     [2 x 64 dims   ][64 keys][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + DQ2_STAGES * SUB_TILE);
  uint64_t* full = bars;                            // [DQ2_STAGES]
  uint64_t* empty = full + DQ2_STAGES;              // [DQ2_STAGES]
  uint64_t* acc_done = empty + DQ2_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const mb::LptIdx li = mb::lpt_index();            // heavy (late) query tiles first, across all heads
  const int qt = (int)gridDim.x - 1 - li.rank;
  const int h = li.h, b = li.b;
  const int hk = h / (p.H / p.Hkv);
  const int q0 = qt * 128;
  int kv_end = p.Sk;
  if (p.causal) kv_end = min(p.Sk, q0 + 128);       // single-pass mode is only used for Sq == Sk
  const int n_it = (kv_end + 63) / 64;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmDS); prefetch_tmap(&tmK64);
    for (int s = 0; s < DQ2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tDQ = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int bh = b * p.H + h;
      for (int n = 0; n < n_it; ++n) {
        const int s = n % DQ2_STAGES; const uint32_t ph = (n / DQ2_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], 2 * SUB_TILE);
        tma_load_3d(sA + s * SUB_TILE, &tmDS, &full[s], q0, n * 64, bh);                  // queries q0 .. q0+63
        tma_load_3d(sA + s * SUB_TILE + SUB_HALF, &tmDS, &full[s], q0 + 64, n * 64, bh);  // queries q0+64 .. q0+127
        tma_load_4d(sB + s * SUB_TILE, &tmK64, &full[s], 0, hk, n * 64, b);
        tma_load_4d(sB + s * SUB_TILE + SUB_HALF, &tmK64, &full[s], 64, hk, n * 64, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_it > 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 128, true, true);
      for (int n = 0; n < n_it; ++n) {
        const int s = n % DQ2_STAGES; const uint32_t ph = (n / DQ2_STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + s * SUB_TILE), b_addr = smem_u32(sB + s * SUB_TILE);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)     // 16 keys per step: 16 k-rows x 128 B = 2048 B further into both tiles
          umma_bf16_ss(tDQ, make_smem_desc(a_addr + kk * 2048, SUB_HALF, 1024), make_smem_desc(b_addr + kk * 2048, SUB_HALF, 1024),
                       idesc, (n | kk) != 0);
        umma_commit(&empty[s]);
      }
      umma_commit(acc_done);
    }
  } else {
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const int qi = q0 + r;
    const bool row_ok = qi < p.Sq;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    bf16* dqp = p.dq + (size_t)b * p.dq_sb + (size_t)(row_ok ? qi : 0) * p.dq_ss + (size_t)h * p.dq_sh;
    if (n_it > 0) {
      mbar_wait(acc_done, 0);
      tc_fence_after();
      store_acc_row(tDQ + lane_off, dqp, row_ok, p.scale);
    } else if (row_ok) {
      const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int c = 0; c < 16; ++c) *reinterpret_cast<uint4*>(dqp + c * 8) = z;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tDQ, 128); }
}

// ld[b,h,q] = {lse * log2(e) (+inf if the row is dead or q >= Sq), sum_d dO * O}, rows padded to Sq_pad.
// A warp takes 8 heads of ONE token (adjacent warps the next 8): the [token][head][128] rows of O and dO are read as whole
// contiguous 8 KB pieces with eight 16-byte loads in flight per lane (the one-row-per-warp version, q fastest, touched one
// 256-byte piece per 8 KB row with two 8-byte loads per lane in flight: 2.2 TB/s).  A half-warp owns one head row.
__global__ void __launch_bounds__(256)
attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, const float* __restrict__ lse,
                  float2* __restrict__ ld, int B, int H, int Sq, int Sq_pad, long long o_sb, long long o_ss, long long o_sh,
                  long long g_sb, long long g_ss, long long g_sh) {
  const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
  const int hgroups = (H + 7) / 8;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long tok = wid / hgroups;
  const int hg = (int)(wid % hgroups);
  if (tok >= (long long)B * Sq_pad) return;
  const int b = (int)(tok / Sq_pad), qi = (int)(tok % Sq_pad);
  const bool live = qi < Sq;
  int4 a[4], g[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int h = hg * 8 + it * 2 + half;
    if (live && h < H) {
      a[it] = mb::ld_stream(reinterpret_cast<const int4*>(o + (size_t)b * o_sb + (size_t)qi * o_ss + (size_t)h * o_sh + l16 * 8));
      g[it] = mb::ld_stream(reinterpret_cast<const int4*>(dout + (size_t)b * g_sb + (size_t)qi * g_ss + (size_t)h * g_sh + l16 * 8));
    } else {
      a[it] = make_int4(0, 0, 0, 0); g[it] = make_int4(0, 0, 0, 0);
    }
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int h = hg * 8 + it * 2 + half;
    const bf162* ah = reinterpret_cast<const bf162*>(&a[it]); const bf162* gh = reinterpret_cast<const bf162*>(&g[it]);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 x = __bfloat1622float2(ah[i]), y = __bfloat1622float2(gh[i]); s += x.x * y.x + x.y * y.y; }
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);      // within the 16-lane half
    if (l16 == 0 && h < H) {
      float2 out = make_float2(INFINITY, 0.f);
      if (live) {
        const float L = lse[((size_t)b * H + h) * Sq + qi];
        out = make_float2((L == -INFINITY) ? INFINITY : L * LOG2E, s);
      }
      ld[((size_t)b * H + h) * Sq_pad + qi] = out;
    }
  }
}

}  // namespace

extern "C" {

// strides: 12 entries as in mb200_attn_generic_bwd ({q,k,v,o} x {b,s,h}); dq/dk/dv/dout are contiguous
// [B,Sq,H,hd] / [B,Sk,Hkv,hd] / [B,Sk,Hkv,hd] / [B,Sq,H,hd].
// delta: fp32 scratch of 2 * B * H * mb200_attn_bwd_sq_pad(Sq) floats (written here: {lse*log2e, rowsum(dO*O)} pairs).
long long mb200_attn_bwd_sq_pad(int Sq) { return (long long)((Sq + 127) / 128) * 128; }

// bytes of the dS^T workspace of the single-pass backward (mb200_attn_bwd_bf16_sp): [B*H][Sq_pad][Sq_pad] bf16
long long mb200_attn_bwd_ds_bytes(int B, int H, int Sq) {
  const long long Sp = mb200_attn_bwd_sq_pad(Sq);
  return 2LL * B * H * Sp * Sp;
}

static int attn_bwd_impl(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                        float* delta, void* dq, void* dk, void* dv, int B, int H, int Hkv, int Sq, int Sk, int hd,
                        const long long* strides, float scale, int causal, const int64_t* kmask, long long kmask_sb,
                        const void* kbits, void* ds_ws, void* stream) {
  if (B <= 0 || Sq <= 0) return MB200_OK;
  if (hd != HD || H % Hkv != 0 || Sk <= 0) return -ENOTSUP;
  for (int i = 0; i < 12; ++i) if (strides[i] & 7) return -ENOTSUP;
  if (kmask && !kbits) return -EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const long long dq_ss = (long long)H * hd, dq_sb = (long long)Sq * H * hd;
  const long long dk_ss = (long long)Hkv * hd, dk_sb = (long long)Sk * Hkv * hd;
  const int Sq_pad = (int)mb200_attn_bwd_sq_pad(Sq);
  {
    const long long warps = (long long)B * Sq_pad * ((H + 7) / 8);
    attn_delta_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>((const bf16*)o, (const bf16*)dout, lse, (float2*)delta,
                                                                         B, H, Sq, Sq_pad, strides[9], strides[10], strides[11],
                                                                         dq_sb, dq_ss, hd);
  }
  CUtensorMap tmQ, tmDO, tmK, tmV, tmQ64, tmDO64, tmK64, tmV64;
  int rc;
  if ((rc = mbtmap::make_bshd(&tmQ, q, B, Sq, H, hd, strides[0], strides[1], strides[2], 128))) return rc;
  if ((rc = mbtmap::make_bshd(&tmQ64, q, B, Sq, H, hd, strides[0], strides[1], strides[2], 64))) return rc;
  if ((rc = mbtmap::make_bshd(&tmDO, dout, B, Sq, H, hd, dq_sb, dq_ss, hd, 128))) return rc;
  if ((rc = mbtmap::make_bshd(&tmDO64, dout, B, Sq, H, hd, dq_sb, dq_ss, hd, 64))) return rc;
  if ((rc = mbtmap::make_bshd(&tmK, k, B, Sk, Hkv, hd, strides[3], strides[4], strides[5], 128))) return rc;
  if ((rc = mbtmap::make_bshd(&tmK64, k, B, Sk, Hkv, hd, strides[3], strides[4], strides[5], 64))) return rc;
  if ((rc = mbtmap::make_bshd(&tmV, v, B, Sk, Hkv, hd, strides[6], strides[7], strides[8], 128))) return rc;
  if ((rc = mbtmap::make_bshd(&tmV64, v, B, Sk, Hkv, hd, strides[6], strides[7], strides[8], 64))) return rc;
  BwdParams p;
  p.ld = (const float2*)delta; p.Sq_pad = Sq_pad;
  p.dq = (bf16*)dq; p.dq_sb = dq_sb; p.dq_ss = dq_ss; p.dq_sh = hd;
  p.dk = (bf16*)dk; p.dk_sb = dk_sb; p.dk_ss = dk_ss; p.dk_sh = hd;
  p.dv = (bf16*)dv; p.dv_sb = dk_sb; p.dv_ss = dk_ss; p.dv_sh = hd;
  p.kbits = kmask ? (const uint32_t*)kbits : nullptr; p.kbits_stride = (Sk + 31) / 32;
  p.B = B; p.H = H; p.Hkv = Hkv; p.Sq = Sq; p.Sk = Sk; p.scale = scale; p.scale_log2 = scale * LOG2E; p.causal = causal;
  constexpr int smem_dkv = 2 * FULL_TILE + 2 * QS * SUB_TILE + 1024 + 256 + QS * 512 + 1024 + 8 * 4096;
  constexpr int smem_dq = 2 * FULL_TILE + 2 * QSQ * SUB_TILE + 1024 + 256;
  constexpr int smem_dq2 = 2 * DQ2_STAGES * SUB_TILE + 1024 + 256;
  static const bool cfg_ok =        // thread-safe one-time setup
      cudaFuncSetAttribute(attn_bwd_dkv_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dkv) == cudaSuccess &&
      cudaFuncSetAttribute(attn_bwd_dq_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq) == cudaSuccess &&
      cudaFuncSetAttribute(attn_bwd_dq2_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq2) == cudaSuccess;
  if (!cfg_ok) { mb200_set_last_error("cudaFuncSetAttribute(attn bwd smem) failed"); return -EIO; }
  dim3 gkv((Sk + 127) / 128, Hkv, B), gq((Sq + 127) / 128, H, B);
  p.ds_t = nullptr; p.ds_row = 0; p.ds_head = 0;
  const bool single_pass = ds_ws != nullptr && Sq == Sk;
  CUtensorMap tmDS, tmDSst = tmK;
  if (single_pass) {
    const long long Sp = mb200_attn_bwd_sq_pad(Sq);                  // rows (keys) and columns (queries) padded to 128
    p.ds_t = (bf16*)ds_ws; p.ds_row = Sp; p.ds_head = Sp * Sp;
    if ((rc = mbtmap::make_3d(&tmDS, ds_ws, (long long)B * H, Sp, Sp, 64, 64))) return rc;      // dQ kernel: loads
    if ((rc = mbtmap::make_3d(&tmDSst, ds_ws, (long long)B * H, Sp, Sp, 64, 32))) return rc;    // dK/dV kernel: stores
  }
  attn_bwd_dkv_sm100_kernel<<<gkv, NTHREADS, smem_dkv, st>>>(tmQ64, tmDO64, tmK, tmV, tmDSst, p);
  if (single_pass) attn_bwd_dq2_sm100_kernel<<<gq, DQ2_THREADS, smem_dq2, st>>>(tmDS, tmK64, p);
  else             attn_bwd_dq_sm100_kernel<<<gq, DQ_THREADS, smem_dq, st>>>(tmQ, tmDO, tmK64, tmV64, p);
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

int mb200_attn_bwd_bf16(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                        float* delta, void* dq, void* dk, void* dv, int B, int H, int Hkv, int Sq, int Sk, int hd,
                        const long long* strides, float scale, int causal, const int64_t* kmask, long long kmask_sb,
                        const void* kbits, void* stream) {
  return attn_bwd_impl(q, k, v, o, dout, lse, delta, dq, dk, dv, B, H, Hkv, Sq, Sk, hd, strides, scale, causal, kmask, kmask_sb,
                       kbits, nullptr, stream);
}
// Single-pass variant (self-attention, Sq == Sk): ds_ws = mb200_attn_bwd_ds_bytes(B, H, Sq) bytes of scratch.  The dK/dV kernel
// writes dS^T there once and dQ = scale * dS K is a tensor-core GEMM over it; with ds_ws == NULL or Sq != Sk it is the
// two-kernel recompute form above.
int mb200_attn_bwd_bf16_sp(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                           float* delta, void* dq, void* dk, void* dv, int B, int H, int Hkv, int Sq, int Sk, int hd,
                           const long long* strides, float scale, int causal, const int64_t* kmask, long long kmask_sb,
                           const void* kbits, void* ds_ws, void* stream) {
  return attn_bwd_impl(q, k, v, o, dout, lse, delta, dq, dk, dv, B, H, Hkv, Sq, Sk, hd, strides, scale, causal, kmask, kmask_sb,
                       kbits, ds_ws, stream);
}

}  // extern "C"

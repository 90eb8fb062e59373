// Flash attention on tcgen05 (head_dim 128): S = Q K^T and O += P V run on the 5th-gen tensor cores with the
// S and O accumulators in TMEM; Q/K/V tiles arrive by TMA into 128B-swizzled shared memory; softmax runs on
// 128 threads (one query row per thread == one TMEM lane per thread) with exp2 and fp32 statistics.
//
// Replaces the flash_attn-2 (mma.sync/HMMA) / SDPA call the reference stack makes in
// LlamaAttention.forward (transformers llama/modeling_llama.py:199-289): causal GQA attention with a key padding
// mask, KV-cache offset (query i sees keys j <= i + Sk - Sq) and fp32 softmax.
//
// CTA = one 128-row query tile of one (batch, head).  Warp roles:
//   warp 0     TMA producer: Q once, then K_j / V_j into 2-stage rings
//   warp 1     MMA issuer (one thread): S_{j+1} = Q K_{j+1}^T is issued before P_j V_j so the tensor pipe works
//              while the softmax warps are busy with S_j (S is double buffered in TMEM: 2 x 128 columns)
//   warps 2-5  softmax: tcgen05.ld S row -> mask -> online max/sum -> P (bf16) -> swizzled smem (A operand of
//              the P V MMA) ; rescale O in TMEM only when a row maximum moved ; epilogue O / l -> global
// TMEM columns: [0,128) S0, [128,256) S1, [256,384) O.
#include <stdlib.h>
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tmap.cuh"

namespace {
using namespace sm100;

constexpr int HD = 128;          // head dim
constexpr int BQ = 128, BKV = 128;
constexpr int HALF_BYTES = 128 * 128;   // one [128 rows x 64 bf16] swizzled sub-tile = 16 KB
constexpr int TILE_BYTES = 2 * HALF_BYTES;

struct FwdParams {
  bf16* o; float* lse;
  long long o_sb, o_ss, o_sh;    // element strides of o [B,Sq,H,hd]
  const uint32_t* kbits;         // [B, kbits_stride] key-valid bitmask (bit k%32 of word k/32) or null
  int kbits_stride;
  int B, H, Hkv, Sq, Sk;
  float scale_log2;              // softmax scale * log2(e)
  int causal;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  bf162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int SMW = 8;                    // softmax warps: two per TMEM lane quarter, each owns 64 of the 128 S columns
constexpr int SMT = SMW * 32;
constexpr int FWD_THREADS = 64 + SMT;

// P_TMEM = true : P (bf16) is written back into the TMEM columns of the S tile it came from and the P V MMA reads its A
//                  operand from tensor memory (tcgen05.mma "ts" form): no P store to / P read from shared memory, which
//                  is the scarce resource here (Q, K, P, V operand reads + TMA writes exceed 128 B/clk otherwise).
// P_TMEM = false: P goes through 128B-swizzled shared memory (first version, kept as the cross-check).
template <bool P_TMEM>
__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                          // 32 KB
  uint8_t* sK = sQ + TILE_BYTES;               // 2 x 32 KB
  uint8_t* sV = sK + 2 * TILE_BYTES;           // 2 x 32 KB
  uint8_t* sP = sV + 2 * TILE_BYTES;           // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* s_empty = bars + 11;  // [2]
  uint64_t* p_full = bars + 13;
  uint64_t* pv_done = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* sx = reinterpret_cast<float*>(bars + 18);      // [2][128] row-max / row-sum exchange between the column halves

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const mb::LptIdx li = mb::lpt_index();               // heavy (late) causal tiles first, across ALL heads
  const int qt = (int)gridDim.x - 1 - li.rank;
  const int h = li.h, b = li.b;
  const int hk = h / (p.H / p.Hkv);
  const int q0 = qt * BQ;
  const int off = p.Sk - p.Sq;
  int kv_end = p.Sk;
  if (p.causal) { kv_end = min(p.Sk, q0 + BQ + off); if (kv_end < 0) kv_end = 0; }
  const int n_kv = (kv_end + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], SMT);
    }
    mbar_init(p_full, SMT); mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS[2] = {tmem_base, tmem_base + 128};
  const uint32_t tO = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0 && n_kv > 0) {
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_4d(sQ, &tmQ, q_full, 0, h, q0, b);
      tma_load_4d(sQ + HALF_BYTES, &tmQ, q_full, 64, h, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1; const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
        tma_load_4d(sK + s * TILE_BYTES, &tmK, &k_full[s], 0, hk, j * BKV, b);
        tma_load_4d(sK + s * TILE_BYTES + HALF_BYTES, &tmK, &k_full[s], 64, hk, j * BKV, b);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
        tma_load_4d(sV + s * TILE_BYTES, &tmV, &v_full[s], 0, hk, j * BKV, b);
        tma_load_4d(sV + s * TILE_BYTES + HALF_BYTES, &tmV, &v_full[s], 64, hk, j * BKV, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_kv > 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, false, true);
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
      auto issue_s = [&](int j) {
        const int s = j & 1; const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_full[s], ph);
        if (!P_TMEM) mbar_wait(&s_empty[s], ph ^ 1);     // P_TMEM: buffer s is recycled by P_j V_j, which was issued before
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + s * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t o2 = (kk >> 2) * HALF_BYTES + (kk & 3) * 32;
          umma_bf16_ss(tS[s], make_smem_desc(q_addr + o2, 16, 1024), make_smem_desc(k_addr + o2, 16, 1024),
                       idesc_qk, kk != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[s]);
        umma_commit(&k_empty[s]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_s(j + 1);
        const int s = j & 1; const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&v_full[s], ph);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sV + s * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          const uint64_t bd = make_smem_desc(v_addr + kk * 2048, HALF_BYTES, 1024);
          if (P_TMEM) {
            umma_bf16_ts(tO, tS[s] + kk * 8, bd, idesc_pv, (j | kk) != 0 ? 1u : 0u);   // 16 bf16 of K = 8 TMEM columns
          } else {
            const uint64_t ad = make_smem_desc(p_addr + (kk >> 2) * HALF_BYTES + (kk & 3) * 32, 16, 1024);
            umma_bf16_ss(tO, ad, bd, idesc_pv, (j | kk) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&v_empty[s]);
        umma_commit(pv_done);
      }
    }
  } else {
    const int qd = warp & 3;                        // TMEM lane quarter
    const int half = (warp - 2) >> 2;               // columns [64*half, 64*half + 64) of every S tile / of O
    const int r = qd * 32 + lane;                   // row in tile == TMEM lane
    const int qi = q0 + r;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    float m = -INFINITY, l = 0.f;                   // l: partial row sum over this thread's columns
    const int limit = p.causal ? (qi + off) : (p.Sk - 1);    // last visible key index for this row
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1; const uint32_t ph = (j >> 1) & 1;
      const int k0 = j * BKV + half * 64;
      mbar_wait(&s_full[s], ph);
      tc_fence_after();
      float sv[64];
      tmem_ld_32x32b_x32(tS[s] + lane_off + half * 64, reinterpret_cast<uint32_t*>(sv));
      tmem_ld_32x32b_x32(tS[s] + lane_off + half * 64 + 32, reinterpret_cast<uint32_t*>(sv) + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[s]);
      // ---- masking (only on tiles that need it) ----
      const bool tail = (j * BKV + BKV > p.Sk) || (p.causal && (j * BKV + BKV - 1 > q0 + off));
      uint32_t w[2] = {0xffffffffu, 0xffffffffu};
      bool need = tail;
      if (p.kbits) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int wi = (k0 >> 5) + c;
          w[c] = (wi < p.kbits_stride) ? __ldg(p.kbits + (size_t)b * p.kbits_stride + wi) : 0u;
          need |= (w[c] != 0xffffffffu);
        }
      }
      if (need) {
        const int lim = min(limit, p.Sk - 1);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int kj = k0 + c * 32 + i;
            const bool ok = (kj <= lim) && ((w[c] >> i) & 1u);
            if (!ok) sv[c * 32 + i] = -INFINITY;
          }
        }
      }
      float mx = sv[0];
#pragma unroll
      for (int i = 1; i < 64; ++i) mx = fmaxf(mx, sv[i]);
      // exchange the half-row maxima (the pair of threads that share TMEM lane r)
      sx[half * 128 + r] = mx;
      named_bar_sync(1, SMT);
      mx = fmaxf(mx, sx[(half ^ 1) * 128 + r]);
      // Lazy rescaling: the running maximum only moves when the tile's maximum exceeds it by more than 2^8 (or on a row's first
      // live tile).  Until then P is formed against the stale maximum (values up to 256: exact in fp32 sums, same relative
      // precision in bf16) and O needs no correction -- so the 64 KB read-modify-write of O through the 64 B/clk TMEM port,
      // which with an exact running maximum happens in almost every iteration (some row of the warp always moves), becomes rare.
      // l, O and lse stay consistent because they are all expressed relative to the same m.
      const float m_cand = fmaxf(m, mx * p.scale_log2);
      const float m_new = (m == -INFINITY || m_cand - m > 8.f) ? m_cand : m;
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = (m == -INFINITY) ? 0.f : fast_exp2(m - m_use);
      float rs = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float p0 = fast_exp2(fmaf(sv[2 * i], p.scale_log2, -m_use));
        const float p1 = fast_exp2(fmaf(sv[2 * i + 1], p.scale_log2, -m_use));
        rs += p0 + p1;
        pk[i] = pack_bf16(p0, p1);
      }
      l = l * alpha + rs;
      const bool changed = (m_new != m);
      m = m_new;
      // P buffer and O are free once P_{j-1} V_{j-1} has completed
      if (j > 0) { mbar_wait(pv_done, (j - 1) & 1); tc_fence_after(); }
      if (j > 0 && __any_sync(0xffffffffu, changed)) {
        const float a = changed ? alpha : 1.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t ov[32];
          tmem_ld_32x32b_x32(tO + lane_off + half * 64 + c * 32, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * a);
          tmem_st_32x32b_x16(tO + lane_off + half * 64 + c * 32, ov);
          tmem_st_32x32b_x16(tO + lane_off + half * 64 + c * 32 + 16, ov + 16);
        }
        tmem_st_wait();
      }
      if (P_TMEM) {
        // P (bf16 pairs) overwrites columns [32*half, 32*half + 32) of the S tile: lane = row, 2 elements per column
        tmem_st_32x32b_x16(tS[s] + lane_off + half * 32, pk);
        tmem_st_32x32b_x16(tS[s] + lane_off + half * 32 + 16, pk + 16);
        tmem_st_wait();
      } else {
        // one 128-byte row of half-tile `half` (K-major, 128B swizzle: chunk ^ (r & 7))
        uint8_t* rowp = sP + half * HALF_BYTES + r * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          *reinterpret_cast<uint4*>(rowp + ((ch ^ (r & 7)) << 4)) = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        fence_proxy_async();
      }
      tc_fence_before();
      mbar_arrive(p_full);
      named_bar_sync(2, SMT);      // sx[] is reused by the next tile
    }
    // ---- epilogue ----
    if (n_kv > 0) { mbar_wait(pv_done, (n_kv - 1) & 1); tc_fence_after(); }
    sx[half * 128 + r] = l;
    named_bar_sync(1, SMT);
    l += sx[(half ^ 1) * 128 + r];
    const float inv = (l > 0.f) ? 1.f / l : 0.f;
    const bool row_ok = qi < p.Sq;
    bf16* op = p.o + (size_t)b * p.o_sb + (size_t)(row_ok ? qi : 0) * p.o_ss + (size_t)h * p.o_sh + half * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t ov[32];
      if (n_kv > 0) { tmem_ld_32x32b_x32(tO + lane_off + half * 64 + c * 32, ov); tmem_ld_wait(); }
      else {
#pragma unroll
        for (int i = 0; i < 32; ++i) ov[i] = 0u;
      }
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o4;
          o4.x = pack_bf16(__uint_as_float(ov[g * 8 + 0]) * inv, __uint_as_float(ov[g * 8 + 1]) * inv);
          o4.y = pack_bf16(__uint_as_float(ov[g * 8 + 2]) * inv, __uint_as_float(ov[g * 8 + 3]) * inv);
          o4.z = pack_bf16(__uint_as_float(ov[g * 8 + 4]) * inv, __uint_as_float(ov[g * 8 + 5]) * inv);
          o4.w = pack_bf16(__uint_as_float(ov[g * 8 + 6]) * inv, __uint_as_float(ov[g * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(op + c * 32 + g * 8) = o4;
        }
      }
    }
    if (row_ok && p.lse && half == 0)
      p.lse[((size_t)b * p.H + h) * p.Sq + qi] = (l > 0.f) ? (m * 0.69314718055994530942f + logf(l)) : -INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// key-valid bitmask: bits[b, w] bit i = (kmask[b, 32 w + i] != 0)
__global__ void kmask_bits_kernel(const int64_t* __restrict__ kmask, long long kmask_sb, uint32_t* __restrict__ bits,
                                  int B, int Sk, int words) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int gw = idx >> 5;                    // one warp per word
  if (gw >= B * words) return;
  const int b = gw / words, w = gw % words;
  const int k = w * 32 + lane;
  const bool v = (k < Sk) && (kmask[(size_t)b * kmask_sb + k] != 0);
  const uint32_t word = __ballot_sync(0xffffffffu, v);
  if (lane == 0) bits[(size_t)b * words + w] = word;
}

}  // namespace

extern "C" {

long long mb200_attn_kbits_words(int Sk) { return (Sk + 31) / 32; }

// bits[b, w] bit i = (kmask[b, 32 w + i] != 0): key-valid bitmask consumed by the attention kernels
int mb200_kmask_bits(const int64_t* kmask, long long kmask_sb, void* bits, int B, int Sk, void* stream) {
  if (B <= 0 || Sk <= 0) return MB200_OK;
  const int words = (Sk + 31) / 32;
  const long long threads = (long long)B * words * 32;
  kmask_bits_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kmask, kmask_sb, (uint32_t*)bits, B, Sk, words);
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

// bf16, head_dim 128 only (returns -ENOTSUP otherwise: caller uses mb200_attn_generic_fwd).
// kbits_ws: device scratch of B * mb200_attn_kbits_words(Sk) uint32 (only touched when kmask != null).
int mb200_attn_fwd_bf16(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Hkv,
                        int Sq, int Sk, int hd, const long long* strides, float scale, int causal,
                        const int64_t* kmask, long long kmask_sb, void* kbits_ws, void* stream) {
  if (B <= 0 || Sq <= 0) return MB200_OK;
  if (hd != HD || H % Hkv != 0 || Sk <= 0) return -ENOTSUP;
  for (int i = 0; i < 12; ++i) if (strides[i] & 7) return -ENOTSUP;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
       reinterpret_cast<uintptr_t>(o)) & 15) return -ENOTSUP;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = mbtmap::make_bshd(&tmQ, q, B, Sq, H, hd, strides[0], strides[1], strides[2], BQ))) return rc;
  if ((rc = mbtmap::make_bshd(&tmK, k, B, Sk, Hkv, hd, strides[3], strides[4], strides[5], BKV))) return rc;
  if ((rc = mbtmap::make_bshd(&tmV, v, B, Sk, Hkv, hd, strides[6], strides[7], strides[8], BKV))) return rc;
  FwdParams p;
  p.o = (bf16*)o; p.lse = lse; p.o_sb = strides[9]; p.o_ss = strides[10]; p.o_sh = strides[11];
  p.kbits = nullptr; p.kbits_stride = 0;
  p.B = B; p.H = H; p.Hkv = Hkv; p.Sq = Sq; p.Sk = Sk;
  p.scale_log2 = scale * 1.44269504088896340736f; p.causal = causal;
  if (kmask) {
    if (!kbits_ws) return -EINVAL;
    const int words = (Sk + 31) / 32;
    const long long threads = (long long)B * words * 32;
    kmask_bits_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(kmask, kmask_sb, (uint32_t*)kbits_ws, B, Sk, words);
    p.kbits = (const uint32_t*)kbits_ws; p.kbits_stride = words;
  }
  constexpr int smem = 7 * TILE_BYTES + 1024 + 256 + 1024;
  static const int p_tmem = [] { const char* e = getenv("MB200_ATTN_P_TMEM"); return e ? atoi(e) : 1; }();   // thread-safe once
  static const bool cfg_ok =
      cudaFuncSetAttribute(attn_fwd_sm100_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess &&
      cudaFuncSetAttribute(attn_fwd_sm100_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
  if (!cfg_ok) { mb200_set_last_error("cudaFuncSetAttribute(attn smem) failed"); return -EIO; }
  dim3 grid((Sq + BQ - 1) / BQ, H, B);
  if (p_tmem) attn_fwd_sm100_kernel<true><<<grid, FWD_THREADS, smem, st>>>(tmQ, tmK, tmV, p);
  else        attn_fwd_sm100_kernel<false><<<grid, FWD_THREADS, smem, st>>>(tmQ, tmK, tmV, p);
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

}  // extern "C"

// Flash attention forward, two query tiles per CTA (head_dim 128): the two 128-row tiles A and B of one (batch, head) share
// every K/V tile (one TMA load, half the L2->SM traffic) and ping-pong on the tensor pipe: while the softmax warps of
// tile A work on S_A(j) the MMA thread issues P_B(j-1) V and S_B(j), and vice versa, so the tensor pipe no longer idles
// for the length of a softmax (ncu on the one-tile kernel: tensor 39 %, MUFU 43 %, issue 39 % -- latency-bound).
// TMEM: S_A [0,128) S_B [128,256) O_A [256,384) O_B [384,512); P (bf16) overwrites the first 64 columns of its S tile and
// is consumed as a TMEM A operand.  Same semantics as attn_fwd_sm100_kernel (attn_sm100.cu).
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tmap.cuh"

namespace {
using namespace sm100;

constexpr int HD = 128, BQ = 128, BKV = 128;
constexpr int HALF_BYTES = 128 * 128;
constexpr int TILE_BYTES = 2 * HALF_BYTES;
constexpr int F2_THREADS = 64 + 2 * 128;

struct Fwd2Params {
  bf16* o; float* lse;
  long long o_sb, o_ss, o_sh;
  const uint32_t* kbits; int kbits_stride;
  int B, H, Hkv, Sq, Sk;
  float scale_log2; int causal;
};

__device__ __forceinline__ float fast_exp2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  bf162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(F2_THREADS, 1)
attn_fwd2_sm100_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const Fwd2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                          // 2 tiles x 32 KB
  uint8_t* sK = sQ + 2 * TILE_BYTES;           // 2 stages x 32 KB
  uint8_t* sV = sK + 2 * TILE_BYTES;           // 2 stages x 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2] per query tile
  uint64_t* p_full = bars + 11;   // [2] per query tile
  uint64_t* pv_done = bars + 13;  // [2] per query tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const mb::LptIdx li = mb::lpt_index();                 // heavy (late) causal tiles first, across ALL heads
  const int pair = (int)gridDim.x - 1 - li.rank;
  const int h = li.h, b = li.b;
  const int hk = h / (p.H / p.Hkv);
  const int off = p.Sk - p.Sq;
  int q0t[2], nkv[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    q0t[t] = (pair * 2 + t) * BQ;
    int kv_end = p.Sk;
    if (p.causal) { kv_end = min(p.Sk, q0t[t] + BQ + off); if (kv_end < 0) kv_end = 0; }
    nkv[t] = (q0t[t] < p.Sq) ? (kv_end + BKV - 1) / BKV : 0;
  }
  const int nmax = max(nkv[0], nkv[1]);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 128); mbar_init(&pv_done[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && nmax > 0) {
      mbar_arrive_expect_tx(q_full, 2 * TILE_BYTES);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        tma_load_4d(sQ + t * TILE_BYTES, &tmQ, q_full, 0, h, q0t[t], b);            // rows beyond Sq are zero-filled
        tma_load_4d(sQ + t * TILE_BYTES + HALF_BYTES, &tmQ, q_full, 64, h, q0t[t], b);
      }
      for (int j = 0; j < nmax; ++j) {
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
    if (lane == 0 && nmax > 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, false, true);
      auto issue_s = [&](int t, int j) {
        const uint32_t q_addr = smem_u32(sQ + t * TILE_BYTES), k_addr = smem_u32(sK + (j & 1) * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t o2 = (kk >> 2) * HALF_BYTES + (kk & 3) * 32;
          umma_bf16_ss(tmem_base + t * 128, make_smem_desc(q_addr + o2, 16, 1024), make_smem_desc(k_addr + o2, 16, 1024),
                       idesc_qk, kk != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int j) {
        const uint32_t v_addr = smem_u32(sV + (j & 1) * TILE_BYTES);
        mbar_wait(&p_full[t], j & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk)
          umma_bf16_ts(tmem_base + 256 + t * 128, tmem_base + t * 128 + kk * 8, make_smem_desc(v_addr + kk * 2048, HALF_BYTES, 1024),
                       idesc_pv, (j | kk) != 0 ? 1u : 0u);
        umma_commit(&pv_done[t]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if (nkv[0] > 0) issue_s(0, 0);
      if (nkv[1] > 0) issue_s(1, 0);
      umma_commit(&k_empty[0]);
      for (int j = 0; j < nmax; ++j) {
        const int s = j & 1; const uint32_t ph = (j >> 1) & 1;
        const bool more = j + 1 < nmax;
        mbar_wait(&v_full[s], ph);
        if (more) mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
        tc_fence_after();
        if (j < nkv[0]) issue_pv(0, j);
        if (more && j + 1 < nkv[0]) issue_s(0, j + 1);      // S_A(j+1) overwrites P_A(j): issued after P_A(j) V (in-order pipe)
        if (j < nkv[1]) issue_pv(1, j);
        umma_commit(&v_empty[s]);
        if (more) {
          if (j + 1 < nkv[1]) issue_s(1, j + 1);
          umma_commit(&k_empty[(j + 1) & 1]);
        }
      }
    }
  } else {
    const int t = (warp - 2) >> 2;                  // query tile of this softmax group
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const int q0 = q0t[t];
    const int qi = q0 + r;
    const int n_kv = nkv[t];
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t tS = tmem_base + t * 128, tO = tmem_base + 256 + t * 128;
    float m = -INFINITY, l = 0.f;
    const int limit = p.causal ? (qi + off) : (p.Sk - 1);
    for (int j = 0; j < n_kv; ++j) {
      const int k0 = j * BKV;
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      float sv[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(tS + lane_off + c * 32, reinterpret_cast<uint32_t*>(sv) + c * 32);
      tmem_ld_wait();
      const bool tail = (k0 + BKV > p.Sk) || (p.causal && (k0 + BKV - 1 > q0 + off));
      uint32_t w[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      bool need = tail;
      if (p.kbits) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int wi = (k0 >> 5) + c;
          w[c] = (wi < p.kbits_stride) ? __ldg(p.kbits + (size_t)b * p.kbits_stride + wi) : 0u;
          need |= (w[c] != 0xffffffffu);
        }
      }
      if (need) {
        const int lim = min(limit, p.Sk - 1);
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int kj = k0 + c * 32 + i;
            if (!((kj <= lim) && ((w[c] >> i) & 1u))) sv[c * 32 + i] = -INFINITY;
          }
      }
      float mx = sv[0];
#pragma unroll
      for (int i = 1; i < 128; ++i) mx = fmaxf(mx, sv[i]);
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
      uint32_t pk[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float p0 = fast_exp2(fmaf(sv[2 * i], p.scale_log2, -m_use));
        const float p1 = fast_exp2(fmaf(sv[2 * i + 1], p.scale_log2, -m_use));
        rs += p0 + p1;
        pk[i] = pack_bf16(p0, p1);
      }
      l = l * alpha + rs;
      const bool changed = (m_new != m);
      m = m_new;
      if (j > 0) { mbar_wait(&pv_done[t], (j - 1) & 1); tc_fence_after(); }     // O_t is quiescent
      if (j > 0 && __any_sync(0xffffffffu, changed)) {
        const float a = changed ? alpha : 1.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t ov[32];
          tmem_ld_32x32b_x32(tO + lane_off + c * 32, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * a);
          tmem_st_32x32b_x16(tO + lane_off + c * 32, ov);
          tmem_st_32x32b_x16(tO + lane_off + c * 32 + 16, ov + 16);
        }
      }
      // P (bf16 pairs) -> columns [0, 64) of this tile's S region (lane-private rows: no cross-thread hazard)
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st_32x32b_x16(tS + lane_off + c * 16, pk + c * 16);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
    }
    if (n_kv > 0) { mbar_wait(&pv_done[t], (n_kv - 1) & 1); tc_fence_after(); }
    const float inv = (l > 0.f) ? 1.f / l : 0.f;
    const bool row_ok = qi < p.Sq;
    bf16* op = p.o + (size_t)b * p.o_sb + (size_t)(row_ok ? qi : 0) * p.o_ss + (size_t)h * p.o_sh;
    if (n_kv > 0 || row_ok) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t ov[32];
        if (n_kv > 0) { tmem_ld_32x32b_x32(tO + lane_off + c * 32, ov); tmem_ld_wait(); }
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
    }
    if (row_ok && p.lse)
      p.lse[((size_t)b * p.H + h) * p.Sq + qi] = (l > 0.f) ? (m * 0.69314718055994530942f + logf(l)) : -INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}
}  // namespace

// same contract as mb200_attn_fwd_bf16; kbits must already be filled (mb200_kmask_bits) when kmask semantics are needed
extern "C" int mb200_attn_fwd2_bf16(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Hkv,
                                    int Sq, int Sk, int hd, const long long* strides, float scale, int causal,
                                    const void* kbits, int kbits_stride, void* stream) {
  if (B <= 0 || Sq <= 0) return MB200_OK;
  if (hd != HD || H % Hkv != 0 || Sk <= 0) return -ENOTSUP;
  for (int i = 0; i < 12; ++i) if (strides[i] & 7) return -ENOTSUP;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
       reinterpret_cast<uintptr_t>(o)) & 15) return -ENOTSUP;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = mbtmap::make_bshd(&tmQ, q, B, Sq, H, hd, strides[0], strides[1], strides[2], BQ))) return rc;
  if ((rc = mbtmap::make_bshd(&tmK, k, B, Sk, Hkv, hd, strides[3], strides[4], strides[5], BKV))) return rc;
  if ((rc = mbtmap::make_bshd(&tmV, v, B, Sk, Hkv, hd, strides[6], strides[7], strides[8], BKV))) return rc;
  Fwd2Params p;
  p.o = (bf16*)o; p.lse = lse; p.o_sb = strides[9]; p.o_ss = strides[10]; p.o_sh = strides[11];
  p.kbits = (const uint32_t*)kbits; p.kbits_stride = kbits_stride;
  p.B = B; p.H = H; p.Hkv = Hkv; p.Sq = Sq; p.Sk = Sk;
  p.scale_log2 = scale * 1.44269504088896340736f; p.causal = causal;
  constexpr int smem = 6 * TILE_BYTES + 1024 + 256;
  // thread-safe one-time setup (C++11 static initialisation): generate() may be driven from a worker thread
  static const cudaError_t cfg = cudaFuncSetAttribute(attn_fwd2_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (cfg != cudaSuccess) { mb200_set_last_error("cudaFuncSetAttribute(attn fwd2 smem) failed"); return -EIO; }
  const int n_qt = (Sq + BQ - 1) / BQ;
  dim3 grid((n_qt + 1) / 2, H, B);
  attn_fwd2_sm100_kernel<<<grid, F2_THREADS, smem, (cudaStream_t)stream>>>(tmQ, tmK, tmV, p);
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

// Image-token merge ("scatter") for interleaved multi-image sequences.
//
// Re-implements, bit-exactly for every integer output, what the reference does with ~20 ATen
// launches + 4 host syncs in LlavaForConditionalGeneration._merge_input_ids_with_image_features
// (reference: mantis/models/mllava/modeling_llava.py:293-360), including its quirks:
//   * left/right padding detection from the last column (:296),
//   * `image_to_overwrite = all(final_embedding == 0, -1)` (:344) -- i.e. a *text* row whose
//     embedding is entirely zero (e.g. nn.Embedding padding_idx rows) is treated as an image slot,
//   * the first `nb_image_pad` candidate slots of every row are skipped (:345),
//   * image feature rows are consumed in row-major (batch, position) order (:353),
//   * position_ids = cumsum(mask)-1 with masked positions forced to 1 (:355).
//
// Three kernels:
//   plan  : per-row statistics + zero-text-row flags; the last block to finish folds them into a
//           header {S, left_padding, n_image_slots, ...} and per-row offsets   (1 launch, B CTAs)
//   index : per-row inverse map final position -> source (text token / image row / zero fill),
//           plus final attention_mask, labels and position_ids                  (1 launch, B CTAs)
//   rows  : the HBM-bound part -- coalesced 128-bit row copies driven by the map (persistent grid)
// One host sync (the header read) is needed because the output length S is data dependent.
#include "common.cuh"

namespace {

constexpr int kPlanThreads = 256;
constexpr int kIndexThreads = 1024;

// rowinfo layout (int64 x 4 per row): {n_image_tokens, nb_image_pad, image_rank_offset, n_zero_text_rows}
// header layout  (int64 x 8): {S, left_padding, n_image_slots, max_images, sum_images, sum_last_is_pad, 0, 0}

template <typename T>
__device__ __forceinline__ bool row_is_all_zero(const T* row, int D, int lane) {
  // warp-cooperative; early exit as soon as a non-zero shows up (the common case: first 512 B).
  bool nz = false;
  for (int base = 0; base < D; base += 32 * 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int i = base + j * 32 + lane;
      if (i < D) nz |= !(mb::Cvt<T>::to_f(row[i]) == 0.0f);   // -0.0 == 0, NaN != 0 (as torch)
    }
    if (__any_sync(0xffffffffu, nz)) return false;
  }
  return true;
}

template <typename T>
__global__ void __launch_bounds__(kPlanThreads)
merge_plan_kernel(const int64_t* __restrict__ ids, const T* __restrict__ embeds,
                  int B, int T_len, int D, int P, int64_t image_token, int64_t pad_token,
                  uint8_t* __restrict__ zflag, int64_t* __restrict__ rowinfo,
                  int64_t* __restrict__ header, unsigned int* __restrict__ ticket) {
  __shared__ int s_nimg, s_nzero;
  __shared__ bool s_last;
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = kPlanThreads / 32;
  if (threadIdx.x == 0) { s_nimg = 0; s_nzero = 0; }
  __syncthreads();
  const int64_t* row_ids = ids + (size_t)b * T_len;
  int nimg = 0, nzero = 0;
  for (int t = wid; t < T_len; t += nw) {
    const bool is_img = row_ids[t] == image_token;
    bool z = false;
    if (!is_img) z = row_is_all_zero(embeds + ((size_t)b * T_len + t) * D, D, lane);
    if (lane == 0) { zflag[(size_t)b * T_len + t] = z ? 1 : 0; nimg += is_img; nzero += z; }
  }
  if (lane == 0) { atomicAdd(&s_nimg, nimg); atomicAdd(&s_nzero, nzero); }
  __syncthreads();
  if (threadIdx.x == 0) {
    rowinfo[4 * b + 0] = s_nimg;
    rowinfo[4 * b + 3] = s_nzero;
    rowinfo[4 * b + 1] = (row_ids[T_len - 1] == pad_token) ? 1 : 0;   // temp: last-column-is-pad
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == (unsigned)(B - 1));
  }
  __syncthreads();
  if (!s_last) return;
  // ---- last CTA: fold (B is small; one thread is plenty and keeps the order deterministic) ----
  if (threadIdx.x == 0) {
    __threadfence();
    int64_t max_img = 0, sum_img = 0, sum_last_pad = 0;
    for (int r = 0; r < B; ++r) {
      int64_t n = ((volatile int64_t*)rowinfo)[4 * r + 0];
      max_img = n > max_img ? n : max_img; sum_img += n;
      sum_last_pad += ((volatile int64_t*)rowinfo)[4 * r + 1];
    }
    const int64_t S = max_img * (int64_t)(P - 1) + T_len;                 // :301
    int64_t off = 0;
    for (int r = 0; r < B; ++r) {
      const int64_t n = ((volatile int64_t*)rowinfo)[4 * r + 0];
      const int64_t nz = ((volatile int64_t*)rowinfo)[4 * r + 3];
      const int64_t nb_pad = (max_img - n) * (int64_t)(P - 1);            // :310
      const int64_t cand = (S - (T_len - n)) + nz;                         // rows still all-zero after step 4
      int64_t cnt = cand - nb_pad; if (cnt < 0) cnt = 0;                   // :345
      rowinfo[4 * r + 1] = nb_pad;
      rowinfo[4 * r + 2] = off;
      off += cnt;
    }
    header[0] = S;
    header[1] = (sum_last_pad == 0) ? 1 : 0;                               // left_padding (:296)
    header[2] = off;
    header[3] = max_img;
    header[4] = sum_img;
    header[5] = sum_last_pad;
    header[6] = 0; header[7] = 0;
    *ticket = 0;                                                           // re-arm for the next call
  }
}

// inclusive block scan of an (a, b) pair of int64; returns totals through *ta, *tb
__device__ __forceinline__ void block_scan2(long long& a, long long& b, long long* sm /*>=66*/,
                                            long long& ta, long long& tb) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long ua = __shfl_up_sync(0xffffffffu, a, o), ub = __shfl_up_sync(0xffffffffu, b, o);
    if (lane >= o) { a += ua; b += ub; }
  }
  __syncthreads();
  if (lane == 31) { sm[2 * wid] = a; sm[2 * wid + 1] = b; }
  __syncthreads();
  if (wid == 0) {
    long long wa = lane < nw ? sm[2 * lane] : 0, wb = lane < nw ? sm[2 * lane + 1] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long ua = __shfl_up_sync(0xffffffffu, wa, o), ub = __shfl_up_sync(0xffffffffu, wb, o);
      if (lane >= o) { wa += ua; wb += ub; }
    }
    sm[2 * lane] = wa; sm[2 * lane + 1] = wb;   // inclusive warp totals
  }
  __syncthreads();
  if (wid > 0) { a += sm[2 * (wid - 1)]; b += sm[2 * (wid - 1) + 1]; }
  ta = sm[2 * (nw - 1)]; tb = sm[2 * (nw - 1) + 1];
  __syncthreads();
}

__global__ void __launch_bounds__(kIndexThreads)
merge_index_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ attn,
                   const int64_t* __restrict__ labels /*nullable*/,
                   const uint8_t* __restrict__ zflag, const int64_t* __restrict__ rowinfo,
                   int B, int T_len, int P, int S, int left_padding,
                   int64_t image_token, int64_t ignore_index,
                   int32_t* __restrict__ srcmap, int64_t* __restrict__ out_mask,
                   int64_t* __restrict__ out_labels /*nullable*/, int64_t* __restrict__ out_pos) {
  __shared__ long long sm[66];
  const int b = blockIdx.x;
  const int64_t* row_ids = ids + (size_t)b * T_len;
  int32_t* src = srcmap + (size_t)b * S;
  const long long nb_pad = rowinfo[4 * b + 1];
  const long long img_off = rowinfo[4 * b + 2];
  // 1. nothing written yet
  for (int s = threadIdx.x; s < S; s += blockDim.x) src[s] = -1;
  __syncthreads();
  // 2. new_token_positions = cumsum(is_img*(P-1)+1) - 1 (+ nb_image_pad when left padded) (:309-312)
  long long carry = 0, dummy_carry = 0;
  for (int base = 0; base < T_len; base += blockDim.x) {
    const int t = base + threadIdx.x;
    const bool valid = t < T_len;
    const bool is_img = valid && (row_ids[t] == image_token);
    long long w = valid ? (is_img ? (long long)P : 1ll) : 0ll, z = 0, tw, tz;
    block_scan2(w, z, sm, tw, tz);
    if (valid && !is_img) {
      long long pos = carry + w - 1 + (left_padding ? nb_pad : 0);
      if (pos >= 0 && pos < S) src[pos] = t;                               // text_to_overwrite (:313)
    }
    carry += tw; dummy_carry += tz;
  }
  __syncthreads();
  // 3. image slots, final mask, labels, position ids
  long long c_carry = 0, m_carry = 0;
  const uint8_t* zrow = zflag + (size_t)b * T_len;
  for (int base = 0; base < S; base += blockDim.x) {
    const int s = base + threadIdx.x;
    const bool valid = s < S;
    int t = valid ? src[s] : 0;
    const bool cand = valid && (t < 0 || zrow[t] != 0);                     // all(final_embedding==0) (:344)
    long long c = cand ? 1 : 0, tc, tm_unused;
    long long zero = 0;
    block_scan2(c, zero, sm, tc, tm_unused);
    const long long c_incl = c_carry + c;
    const bool ow = cand && (c_incl - 1 >= nb_pad);                         // (:345)
    long long m = 0;
    if (valid) { m = (t >= 0) ? (long long)attn[(size_t)b * T_len + t] : 0ll; if (ow) m |= 1ll; }   // (:339,:354)
    long long msum = m, z2 = 0, tm, tz2;
    block_scan2(msum, z2, sm, tm, tz2);
    if (valid) {
      const size_t o = (size_t)b * S + s;
      out_mask[o] = m;
      out_pos[o] = (m == 0) ? 1ll : (m_carry + msum - 1);                   // (:355)
      if (out_labels) out_labels[o] = (t >= 0) ? labels[(size_t)b * T_len + t] : ignore_index;   // (:323,:341)
      src[s] = ow ? (int32_t)(-(c_incl - 1 - nb_pad + img_off) - 2) : t;
    }
    c_carry += tc; m_carry += tm;
  }
}

// One warp per destination row; 16-byte vectors, 8 loads in flight per lane.
__global__ void __launch_bounds__(256)
merge_rows_kernel(const int32_t* __restrict__ srcmap, const uint8_t* __restrict__ text,
                  const uint8_t* __restrict__ img, uint8_t* __restrict__ out,
                  long long n_rows, int S, int T_len, long long row_bytes, long long n_img_rows) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nvec = (int)(row_bytes >> 4);
  for (long long r = warp0; r < n_rows; r += nwarps) {
    const int sidx = srcmap[r];
    const long long b = r / S;
    const uint8_t* sp = nullptr;
    if (sidx >= 0) sp = text + ((size_t)b * T_len + sidx) * row_bytes;
    else if (sidx <= -2) { long long k = -(long long)sidx - 2; if (k < n_img_rows) sp = img + (size_t)k * row_bytes; }
    int4* dp = reinterpret_cast<int4*>(out + (size_t)r * row_bytes);
    if (sp) {
      const int4* s4 = reinterpret_cast<const int4*>(sp);
      int i = lane;
      for (; i + 7 * 32 < nvec; i += 8 * 32) {
        int4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = mb::ld_stream(s4 + i + j * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) mb::st_stream(dp + i + j * 32, v[j]);
      }
      for (; i < nvec; i += 32) mb::st_stream(dp + i, mb::ld_stream(s4 + i));
    } else {
      const int4 z = make_int4(0, 0, 0, 0);
      for (int i = lane; i < nvec; i += 32) mb::st_stream(dp + i, z);
    }
  }
}

// Backward of the row copy: route d(final_embedding) rows back to text rows / image feature rows.
// grad_text must be pre-zeroed (image-token rows and overwritten zero-rows receive no gradient).
__global__ void __launch_bounds__(256)
merge_rows_bwd_kernel(const int32_t* __restrict__ srcmap, const uint8_t* __restrict__ gout,
                      uint8_t* __restrict__ gtext, uint8_t* __restrict__ gimg,
                      long long n_rows, int S, int T_len, long long row_bytes, long long n_img_rows) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nvec = (int)(row_bytes >> 4);
  for (long long r = warp0; r < n_rows; r += nwarps) {
    const int sidx = srcmap[r];
    const long long b = r / S;
    uint8_t* dp8 = nullptr;
    if (sidx >= 0) { if (gtext) dp8 = gtext + ((size_t)b * T_len + sidx) * row_bytes; }
    else if (sidx <= -2) { long long k = -(long long)sidx - 2; if (gimg && k < n_img_rows) dp8 = gimg + (size_t)k * row_bytes; }
    if (!dp8) continue;
    const int4* s4 = reinterpret_cast<const int4*>(gout + (size_t)r * row_bytes);
    int4* dp = reinterpret_cast<int4*>(dp8);
    int i = lane;
    for (; i + 7 * 32 < nvec; i += 8 * 32) {
      int4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = mb::ld_stream(s4 + i + j * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) mb::st_stream(dp + i + j * 32, v[j]);
    }
    for (; i < nvec; i += 32) mb::st_stream(dp + i, mb::ld_stream(s4 + i));
  }
}

}  // namespace

extern "C" {

// workspace sizes (bytes) the caller must provide
long long mb200_merge_ws_bytes(int B, int T_len) {
  // ticket (16 B, fixed offset 0 so that it stays armed across calls with different shapes)
  // + rowinfo int64[4B] + zflag [B*T] (rounded to 16)
  long long z = ((long long)B * T_len + 15) / 16 * 16;
  return 16 + (long long)B * 4 * 8 + z;
}

int mb200_merge_plan(const int64_t* ids, const void* embeds, int dtype, int B, int T_len, int D, int P,
                     int64_t image_token, int64_t pad_token, void* ws, int64_t* header_dev,
                     void* stream) {
  if (B <= 0 || T_len <= 0 || D <= 0 || P <= 0) return -EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* ticket = (unsigned int*)ws;
  int64_t* rowinfo = (int64_t*)((uint8_t*)ws + 16);
  uint8_t* zflag = (uint8_t*)ws + 16 + (long long)B * 32;
  // ticket must be zero on first use: caller zero-initialises ws once; the kernel re-arms it.
  if (dtype == MB200_DTYPE_BF16)
    merge_plan_kernel<bf16><<<B, kPlanThreads, 0, st>>>(ids, (const bf16*)embeds, B, T_len, D, P, image_token,
                                                        pad_token, zflag, rowinfo, header_dev, ticket);
  else if (dtype == MB200_DTYPE_F32)
    merge_plan_kernel<float><<<B, kPlanThreads, 0, st>>>(ids, (const float*)embeds, B, T_len, D, P, image_token,
                                                         pad_token, zflag, rowinfo, header_dev, ticket);
  else return -EINVAL;
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

int mb200_merge_index(const int64_t* ids, const int64_t* attn, const int64_t* labels, const void* ws,
                      int B, int T_len, int P, int S, int left_padding, int64_t image_token,
                      int64_t ignore_index, int32_t* srcmap, int64_t* out_mask, int64_t* out_labels,
                      int64_t* out_pos, void* stream) {
  if (B <= 0 || T_len <= 0 || S < T_len) return -EINVAL;
  const int64_t* rowinfo = (const int64_t*)((const uint8_t*)ws + 16);
  const uint8_t* zflag = (const uint8_t*)ws + 16 + (long long)B * 32;
  merge_index_kernel<<<B, kIndexThreads, 0, (cudaStream_t)stream>>>(
      ids, attn, labels, zflag, rowinfo, B, T_len, P, S, left_padding, image_token, ignore_index,
      srcmap, out_mask, labels ? out_labels : nullptr, out_pos);
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

int mb200_merge_rows(const int32_t* srcmap, const void* text, const void* img, void* out,
                     int B, int S, int T_len, long long row_bytes, long long n_img_rows, void* stream) {
  if (row_bytes <= 0 || (row_bytes & 15)) return -EINVAL;   // rows must be 16-byte multiples
  long long n_rows = (long long)B * S;
  int grid = mb::num_sms() * 8;
  long long need = (n_rows + 7) / 8; if (need < grid) grid = (int)(need > 0 ? need : 1);
  merge_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(srcmap, (const uint8_t*)text, (const uint8_t*)img,
                                                           (uint8_t*)out, n_rows, S, T_len, row_bytes, n_img_rows);
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

int mb200_merge_rows_bwd(const int32_t* srcmap, const void* gout, void* gtext, void* gimg,
                         int B, int S, int T_len, long long row_bytes, long long n_img_rows, void* stream) {
  if (row_bytes <= 0 || (row_bytes & 15)) return -EINVAL;
  long long n_rows = (long long)B * S;
  int grid = mb::num_sms() * 8;
  long long need = (n_rows + 7) / 8; if (need < grid) grid = (int)(need > 0 ? need : 1);
  merge_rows_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(srcmap, (const uint8_t*)gout, (uint8_t*)gtext,
                                                               (uint8_t*)gimg, n_rows, S, T_len, row_bytes, n_img_rows);
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

}  // extern "C"

// mantis_b200 -- common device helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <errno.h>

#define MB200_OK 0
#define MB200_DTYPE_F32 0
#define MB200_DTYPE_BF16 1

#define MB200_CHECK_LAUNCH()                                   \
  do {                                                         \
    cudaError_t e__ = cudaGetLastError();                      \
    if (e__ != cudaSuccess) { mb200_set_last_error(cudaGetErrorString(e__)); return -EIO; } \
  } while (0)

extern "C" void mb200_set_last_error(const char* msg);

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

namespace mb {

template <typename T> struct Cvt;
template <> struct Cvt<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct Cvt<bf16> {
  static __device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ bf16 from_f(float v) { return __float2bfloat16_rn(v); }
};

template <typename T> __device__ __forceinline__ float ldf(const T* p) { return Cvt<T>::to_f(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v) { *p = Cvt<T>::from_f(v); }
// round-trip through the storage type (reproduces torch's per-op rounding in bf16 mode)
template <typename T> __device__ __forceinline__ float rnd(float v) { return Cvt<T>::to_f(Cvt<T>::from_f(v)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; `red` must hold >= 33 floats of shared memory. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = (lane < nw) ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = (lane < nw) ? red[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// 128-bit streaming loads / stores (data touched once: keep it out of L1)
__device__ __forceinline__ int4 ld_stream(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(int4* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 8 packed storage elements <-> 8 floats helpers for vectorised elementwise kernels
template <typename T> struct Vec8;
template <> struct Vec8<bf16> {   // 16 bytes
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const bf16* p, float* f) {
    int4 v = *reinterpret_cast<const int4*>(p);
    const bf162* h = reinterpret_cast<const bf162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  static __device__ __forceinline__ void store(bf16* p, const float* f) {
    int4 v; bf162* h = reinterpret_cast<bf162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<int4*>(p) = v;
  }
};
template <> struct Vec8<float> {  // 32 bytes
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const float* p, float* f) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
};

// Longest-processing-time-first order over a (tiles, heads, batch) grid: CTAs are dispatched in linear blockIdx order, so the
// tile rank is taken from the SLOW part of the linear index and (head, batch) from the fast part -- every head's heaviest
// causal tile starts first and the lightest tiles fill the tail (with the natural (x = tile, y = head) mapping the heavy tile
// of the last head starts after 31/32 of the work has been handed out and runs alone at the end).
struct LptIdx { int rank, h, b; };
__device__ __forceinline__ LptIdx lpt_index() {
  const unsigned L = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  const unsigned nhb = gridDim.y * gridDim.z;
  const unsigned hb = L % nhb;
  LptIdx r; r.rank = (int)(L / nhb); r.h = (int)(hb % gridDim.y); r.b = (int)(hb / gridDim.y);
  return r;
}

// ---- programmatic dependent launch (PDL): a kernel launched with the attribute may become resident while its
// predecessor on the stream is still draining; it must call pdl_wait() before touching anything the predecessor writes
// (and before writing anything the predecessor reads).  Weights are constant, so the decode kernels prefetch them first.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// set by mb200_llama_decode_step around its launch sequence (thread-local: the C ABI stays re-entrant)
int& pdl_mode();

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                    Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int num_sms() {
  static const int n = [] {
    int dev = 0, v = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v <= 0 ? 148 : v;
  }();
  return n;
}

}  // namespace mb

// Shape-agnostic SIMT GEMM (fp32 accumulate) -- the "any shape / any dtype" path used for tiny
// configs (hidden 64, vocab 320 ...), fp32 parity runs, and as the on-GPU cross-check of the
// tcgen05 kernels in gemm_sm100.cu.  C[M,N] = alpha * op(A) * op(B) (+ bias[N]) (+ beta * C)
//   transA = 0: A is [M,K] row-major (lda) ; 1: A is [K,M] row-major
//   transB = 0: B is [K,N] row-major (ldb) ; 1: B is [N,K] row-major   (nn.Linear weight layout)
#include "common.cuh"

namespace {
using mb::Cvt;
constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256)
gemm_generic_kernel(const TA* __restrict__ A, const TB* __restrict__ B, TC* __restrict__ C,
                    const TC* __restrict__ bias, int M, int N, int K, long long lda, long long ldb, long long ldc,
                    int transA, int transB, float alpha, float beta,
                    long long strideA, long long strideB, long long strideC) {
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN + 1];
  A += (size_t)blockIdx.z * strideA; B += (size_t)blockIdx.z * strideB; C += (size_t)blockIdx.z * strideC;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // load A tile (BM x BK) and B tile (BK x BN): 1024 elements each, 4 per thread
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const int e = threadIdx.x + l * 256;
      {
        int mm, kk;
        if (transA) { mm = e % BM; kk = e / BM; } else { kk = e % BK; mm = e / BK; }
        const int gm = m0 + mm, gk = k0 + kk;
        float v = 0.f;
        if (gm < M && gk < K) v = Cvt<TA>::to_f(transA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk]);
        As[kk][mm] = v;
      }
      {
        int nn, kk;
        if (transB) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
        const int gn = n0 + nn, gk = k0 + kk;
        float v = 0.f;
        if (gn < N && gk < K) v = Cvt<TB>::to_f(transB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn]);
        Bs[kk][nn] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn >= N) continue;
      float v = alpha * acc[i][j];
      if (bias) v += Cvt<TC>::to_f(bias[gn]);
      TC* cp = C + (size_t)gm * ldc + gn;
      if (beta != 0.f) v += beta * Cvt<TC>::to_f(*cp);
      *cp = Cvt<TC>::from_f(v);
    }
  }
}

template <typename TA, typename TB, typename TC>
int launch(const void* A, const void* B, void* C, const void* bias, int M, int N, int K, long long lda,
           long long ldb, long long ldc, int transA, int transB, float alpha, float beta, int batch,
           long long sA, long long sB, long long sC, cudaStream_t st) {
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batch);
  if (grid.y > 65535 || grid.z > 65535) return -EINVAL;
  gemm_generic_kernel<TA, TB, TC><<<grid, 256, 0, st>>>((const TA*)A, (const TB*)B, (TC*)C, (const TC*)bias, M, N, K,
                                                        lda, ldb, ldc, transA, transB, alpha, beta, sA, sB, sC);
  return 0;
}
}  // namespace

extern "C" int mb200_gemm_generic(const void* A, const void* B, void* C, const void* bias, int M, int N, int K,
                                  long long lda, long long ldb, long long ldc, int transA, int transB,
                                  float alpha, float beta, int batch, long long strideA, long long strideB,
                                  long long strideC, int dtype_ab, int dtype_c, void* stream) {
  if (M <= 0 || N <= 0 || batch <= 0) return MB200_OK;
  if (K < 0) return -EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (dtype_ab == MB200_DTYPE_F32 && dtype_c == MB200_DTYPE_F32)
    rc = launch<float, float, float>(A, B, C, bias, M, N, K, lda, ldb, ldc, transA, transB, alpha, beta, batch, strideA, strideB, strideC, st);
  else if (dtype_ab == MB200_DTYPE_BF16 && dtype_c == MB200_DTYPE_BF16)
    rc = launch<bf16, bf16, bf16>(A, B, C, bias, M, N, K, lda, ldb, ldc, transA, transB, alpha, beta, batch, strideA, strideB, strideC, st);
  else if (dtype_ab == MB200_DTYPE_BF16 && dtype_c == MB200_DTYPE_F32)
    rc = launch<bf16, bf16, float>(A, B, C, bias, M, N, K, lda, ldb, ldc, transA, transB, alpha, beta, batch, strideA, strideB, strideC, st);
  else return -EINVAL;
  if (rc) return rc;
  MB200_CHECK_LAUNCH();
  return MB200_OK;
}

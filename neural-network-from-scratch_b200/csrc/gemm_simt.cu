// fp32 SIMT GEMM (FFMA) -- the fp32-accurate contraction path behind Backend.matmul / dot
// (backends/backend.py:133-140; numpy_backend.py:134-137 = np.matmul / np.dot). It exists to honour the
// north star's 1e-5-relative fp32 contract, which a 10-bit-mantissa tensor-core path cannot; the bf16
// tcgen05 path (gemm_tc.cu) is the throughput path.
//   C[b] = alpha * op(A[b]) * op(B[b]) + beta * C[b] (+ bias[n])      row-major, arbitrary M/N/K/ld
// 128x128x16 block tile, 256 threads, 8x8 register tile per thread, double-buffered smem.
#include "common.cuh"

#define BM 128
#define BN 128
#define BK 16

template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
k_sgemm(int M, int N, int K, float alpha, const float* __restrict__ A, int64_t lda, int64_t sA, const float* __restrict__ B,
        int64_t ldb, int64_t sB, float beta, float* __restrict__ C, int64_t ldc, int64_t sC, const float* __restrict__ bias) {
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  int bz = blockIdx.z;
  A += (int64_t)bz * sA;
  B += (int64_t)bz * sB;
  C += (int64_t)bz * sC;
  int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int tid = threadIdx.x;
  int tx = tid % 16, ty = tid / 16;  // 16 x 16 thread grid, each 8x8 outputs (strided by 16 for conflict-free smem reads)

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // each thread loads 8 A elements and 8 B elements per K-tile
  float ra[8], rb[8];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      int e = tid + l * 256;  // 0 .. 2047
      // A tile element (m, k): choose the mapping whose fastest index follows memory
      int am, ak;
      if (!TA) { ak = e % BK; am = e / BK; } else { am = e % BM; ak = e / BM; }
      int gm = m0 + am, gk = k0 + ak;
      float v = 0.f;
      if (gm < M && gk < K) v = TA ? A[(int64_t)gk * lda + gm] : A[(int64_t)gm * lda + gk];
      ra[l] = v;
      int bn, bk;
      if (!TB) { bn = e % BN; bk = e / BN; } else { bk = e % BK; bn = e / BK; }
      int gn = n0 + bn;
      gk = k0 + bk;
      v = 0.f;
      if (gn < N && gk < K) v = TB ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn];
      rb[l] = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      int e = tid + l * 256;
      int am, ak;
      if (!TA) { ak = e % BK; am = e / BK; } else { am = e % BM; ak = e / BM; }
      As[buf][ak][am] = ra[l];
      int bn, bk;
      if (!TB) { bn = e % BN; bk = e / BN; } else { bk = e % BK; bn = e / BK; }
      Bs[buf][bk][bn] = rb[l];
    }
  };

  int nk = (K + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[buf][kk][ty + i * 16];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Bs[buf][kk][tx + j * 16];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int gm = m0 + ty + i * 16;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int gn = n0 + tx + j * 16;
      if (gn >= N) continue;
      float v = alpha * acc[i][j];
      if (bias) v += bias[gn];
      float* c = C + (int64_t)gm * ldc + gn;
      if (beta != 0.f) v += beta * (*c);
      *c = v;
    }
  }
}

NFB_API int nfb_gemm_f32(int transA, int transB, int M, int N, int K, float alpha, const float* A, int64_t lda, int64_t strideA,
                         const float* B, int64_t ldb, int64_t strideB, float beta, float* C, int64_t ldc, int64_t strideC,
                         int batch, const float* bias) {
  if (M == 0 || N == 0 || batch == 0) return 0;
  if (batch > 65535) return NFB_FAIL("nfb_gemm_f32: batch %d > 65535", batch);
  dim3 grid((unsigned)nfb_cdiv(N, BN), (unsigned)nfb_cdiv(M, BM), (unsigned)batch);
  if (grid.y > 65535) return NFB_FAIL("nfb_gemm_f32: M too large");
  cudaStream_t s = nfb_cur_stream();
#define GO(TA, TB) k_sgemm<TA, TB><<<grid, 256, 0, s>>>(M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC, bias)
  if (!transA && !transB) GO(false, false);
  else if (!transA && transB) GO(false, true);
  else if (transA && !transB) GO(true, false);
  else GO(true, true);
#undef GO
  NFB_LAUNCH_CHECK();
  return 0;
}


// ---- fp64 contraction: Backend.matmul / dot on float64 arrays (the reference's `backend.array(np.random.randn(...))` is
// float64: tests/test_backends.py:131-147).  Not on the training path; a plain 64x64x16 tiled kernel, 4x4 outputs per thread.
#define DBM 64
#define DBN 64
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
k_dgemm(int M, int N, int K, const double* __restrict__ A, int64_t lda, int64_t sA, const double* __restrict__ B, int64_t ldb, int64_t sB,
        double* __restrict__ C, int64_t ldc, int64_t sC) {
  __shared__ double As[BK][DBM + 1];
  __shared__ double Bs[BK][DBN + 1];
  A += (int64_t)blockIdx.z * sA;
  B += (int64_t)blockIdx.z * sB;
  C += (int64_t)blockIdx.z * sC;
  const int m0 = blockIdx.y * DBM, n0 = blockIdx.x * DBN;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const int e = tid + l * 256;   // 0 .. 1023 = 64 x 16
      int am, ak;
      if (!TA) { ak = e % BK; am = e / BK; } else { am = e % DBM; ak = e / DBM; }
      int gm = m0 + am, gk = k0 + ak;
      As[ak][am] = (gm < M && gk < K) ? (TA ? A[(int64_t)gk * lda + gm] : A[(int64_t)gm * lda + gk]) : 0.0;
      int bn, bk;
      if (!TB) { bn = e % DBN; bk = e / DBN; } else { bk = e % BK; bn = e / BK; }
      int gn = n0 + bn;
      gk = k0 + bk;
      Bs[bk][bn] = (gn < N && gk < K) ? (TB ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn]) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty + i * 16];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + j * 16];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty + i * 16;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx + j * 16;
      if (gn < N) C[(int64_t)gm * ldc + gn] = acc[i][j];
    }
  }
}

NFB_API int nfb_gemm_f64(int transA, int transB, int M, int N, int K, const double* A, int64_t lda, int64_t strideA, const double* B,
                         int64_t ldb, int64_t strideB, double* C, int64_t ldc, int64_t strideC, int batch) {
  if (M == 0 || N == 0 || batch == 0) return 0;
  if (batch > 65535) return NFB_FAIL("nfb_gemm_f64: batch %d > 65535", batch);
  dim3 grid((unsigned)nfb_cdiv(N, DBN), (unsigned)nfb_cdiv(M, DBM), (unsigned)batch);
  if (grid.y > 65535) return NFB_FAIL("nfb_gemm_f64: M too large");
  cudaStream_t s = nfb_cur_stream();
#define GO(TA, TB) k_dgemm<TA, TB><<<grid, 256, 0, s>>>(M, N, K, A, lda, strideA, B, ldb, strideB, C, ldc, strideC)
  if (!transA && !transB) GO(false, false);
  else if (!transA && transB) GO(false, true);
  else if (transA && !transB) GO(true, false);
  else GO(true, true);
#undef GO
  NFB_LAUNCH_CHECK();
  return 0;
}


// ---- fp32-accurate contraction ON THE TENSOR CORES: three-way bf16 split -------------------------------------------------
// x = x1 + x2 + x3 with x1 = bf16(x), x2 = bf16(x - x1), x3 = bf16(x - x1 - x2) carries 24 significand bits, and
//   A.B = A1B1 + (A1B2 + A2B1) + (A2B2 + A1B3 + A3B1) + O(2^-24 |A||B|).
// Every bf16 x bf16 product is exact in fp32 and the tcgen05 pipe accumulates in fp32, so the six products are ONE bf16 GEMM
// over a six-fold K: this kernel writes the operand with its K axis concatenated six times in the pattern
//   role 0 (A): 2 1 3 1 2 1        role 1 (B): 2 3 1 2 1 1
// i.e. SMALLEST TERMS FIRST (A2B2, A1B3, A3B1, then A1B2, A2B1, then A1B1).  The tensor pipe aligns the products of an MMA step to
// the accumulator and truncates: second-order products (2^-16) added step after step INTO a large accumulator each lost their low
// bits (measured 3e-6 of |A||B| with the large term first, against 3e-7 for the pipe's own accumulation error); accumulated on
// their own first and merged once, they cost nothing (profiles/r02_fp32_split.txt).
// k_axis 1: K runs along the columns -> out is (R, 6*C) row-major; k_axis 0: K runs along the rows -> out is (6*R, C).
// The host (nfb200/kernels.py gemm_f32) then calls nfb_gemm_bf16 with K' = 6K and an fp32 C.  Replaces the FFMA kernel above
// for large 2-D contractions of the fp32 mode (functional/arithmetic.py:441-575 at the north star's 1e-5 tolerance).
__global__ void __launch_bounds__(256)
k_split_bf16x3(const float* __restrict__ x, int64_t R, int64_t C, int64_t ld, int role, int k_axis, __nv_bfloat16* __restrict__ out) {
  nfb_pdl_enter();
  const int64_t c4n = C >> 2, total = R * c4n;
  const int64_t seg = k_axis ? C : R * C;          // distance between consecutive copies, in elements
  const int64_t pitch = k_axis ? 6 * C : C;        // output row pitch
  const unsigned patA = 0x010201u;                 // copy j (j = 0 .. 5) takes part (pat >> 4j) & 15 (0-based): A = 1 0 2 0 1 0
  const unsigned patB = 0x001021u;                 //                                                             B = 1 2 0 1 0 0
  const unsigned pat = role == 0 ? patA : patB;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c4n, c = (i - r * c4n) << 2;
    const float4 v = *reinterpret_cast<const float4*>(x + r * ld + c);
    const float f[4] = {v.x, v.y, v.z, v.w};
    uint2 part[3];
    __nv_bfloat16 h[3][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat16 a = __float2bfloat16_rn(f[e]);
      const float r1 = f[e] - __bfloat162float(a);
      const __nv_bfloat16 b = __float2bfloat16_rn(r1);
      const float r2 = r1 - __bfloat162float(b);
      h[0][e] = a; h[1][e] = b; h[2][e] = __float2bfloat16_rn(r2);
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) part[q] = *reinterpret_cast<const uint2*>(h[q]);
    __nv_bfloat16* o = out + r * pitch + c;
#pragma unroll
    for (int j = 0; j < 6; ++j) *reinterpret_cast<uint2*>(o + j * seg) = part[(pat >> (4 * j)) & 15u];
  }
}

NFB_API int nfb_split_bf16x3(const float* x, int64_t R, int64_t C, int64_t ld, int role, int k_axis, void* out) {
  if (R == 0 || C == 0) return 0;
  if (C % 4 != 0 || ld % 4 != 0 || (uintptr_t)x % 16 != 0 || (uintptr_t)out % 8 != 0)
    return NFB_FAIL("nfb_split_bf16x3: needs C %% 4 == 0 and 16-byte aligned fp32 rows (C=%lld ld=%lld)", (long long)C, (long long)ld);
  if (role != 0 && role != 1) return NFB_FAIL("nfb_split_bf16x3: role must be 0 (A) or 1 (B)");
  const int64_t total = R * (C >> 2);
  int grid = nfb_grid_for(total, 256);
  nfb_launch_pdl(k_split_bf16x3, dim3(grid), dim3(256), 0, nfb_cur_stream(), x, R, C, ld, role, k_axis ? 1 : 0, (__nv_bfloat16*)out);
  NFB_LAUNCH_CHECK();
  return 0;
}

// Elementwise family: broadcasting binary ops, unary ops (+ their backward forms), casts, where,
// strided gather/scatter copies (transpose / slice / concat), fill / arange.
// Follows the math of functional/arithmetic.py:14-438 and functional/activation.py:14-471 of the
// reference (formulas restated in DESIGN.md / SURVEY App. B). All kernels are HBM-bound:
// 128-bit accesses on the contiguous fast paths, grid = multiple of the SM count.
#include "common.cuh"

struct DimPack {
  int ndim;
  int64_t shape[NFB_MAX_DIMS];
  int64_t sa[NFB_MAX_DIMS];
  int64_t sb[NFB_MAX_DIMS];
  int64_t sc[NFB_MAX_DIMS];
};

enum BinOp { B_ADD = 0, B_SUB, B_MUL, B_DIV, B_POW, B_MAX, B_MIN, B_EQ = 10, B_NE, B_LT, B_LE, B_GT, B_GE };

template <typename T> __device__ __forceinline__ T pow_t(T a, T b);
template <> __device__ __forceinline__ float pow_t<float>(float a, float b) { return powf(a, b); }
template <> __device__ __forceinline__ double pow_t<double>(double a, double b) { return pow(a, b); }
template <> __device__ __forceinline__ int32_t pow_t<int32_t>(int32_t a, int32_t b) {
  int32_t r = 1;
  for (int32_t i = 0; i < b; ++i) r *= a;
  return r;
}
template <> __device__ __forceinline__ int64_t pow_t<int64_t>(int64_t a, int64_t b) {
  int64_t r = 1;
  for (int64_t i = 0; i < b; ++i) r *= a;
  return r;
}
template <> __device__ __forceinline__ uint8_t pow_t<uint8_t>(uint8_t a, uint8_t b) { return b ? a : 1; }

template <typename T> __device__ __forceinline__ T divi_t(T a, T b) { return a / b; }
template <> __device__ __forceinline__ int32_t divi_t<int32_t>(int32_t a, int32_t b) {  // numpy floor division
  if (b == 0) return 0;
  int32_t q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}
template <> __device__ __forceinline__ int64_t divi_t<int64_t>(int64_t a, int64_t b) {
  if (b == 0) return 0;
  int64_t q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}
template <> __device__ __forceinline__ uint8_t divi_t<uint8_t>(uint8_t a, uint8_t b) { return b ? a / b : 0; }

// numpy maximum/minimum propagate NaN
template <typename T> __device__ __forceinline__ T max_t(T a, T b) { return a > b ? a : b; }
template <> __device__ __forceinline__ float max_t<float>(float a, float b) {
  return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
}
template <> __device__ __forceinline__ double max_t<double>(double a, double b) {
  return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
}
template <typename T> __device__ __forceinline__ T min_t(T a, T b) { return a < b ? a : b; }
template <> __device__ __forceinline__ float min_t<float>(float a, float b) {
  return (a != a) ? a : ((b != b) ? b : (a < b ? a : b));
}
template <> __device__ __forceinline__ double min_t<double>(double a, double b) {
  return (a != a) ? a : ((b != b) ? b : (a < b ? a : b));
}

template <int OP, typename T> __device__ __forceinline__ T bin_arith(T a, T b) {
  if (OP == B_ADD) return a + b;
  if (OP == B_SUB) return a - b;
  if (OP == B_MUL) return a * b;
  if (OP == B_DIV) return divi_t<T>(a, b);
  if (OP == B_POW) return pow_t<T>(a, b);
  if (OP == B_MAX) return max_t<T>(a, b);
  return min_t<T>(a, b);
}
template <int OP, typename T> __device__ __forceinline__ uint8_t bin_cmp(T a, T b) {
  if (OP == B_EQ) return a == b;
  if (OP == B_NE) return a != b;
  if (OP == B_LT) return a < b;
  if (OP == B_LE) return a <= b;
  if (OP == B_GT) return a > b;
  return a >= b;
}

__device__ __forceinline__ void unravel2(int64_t idx, const DimPack& d, int64_t& oa, int64_t& ob) {
  oa = 0;
  ob = 0;
#pragma unroll 1
  for (int k = d.ndim - 1; k >= 0; --k) {
    int64_t s = d.shape[k];
    int64_t c = idx % s;
    idx /= s;
    oa += c * d.sa[k];
    ob += c * d.sb[k];
  }
}

// generic broadcasting kernels (any dtype, any strides)
template <int OP, typename T>
__global__ void k_bin_arith_generic(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t n, DimPack d) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t oa, ob;
    unravel2(i, d, oa, ob);
    out[i] = bin_arith<OP, T>(a[oa], b[ob]);
  }
}
template <int OP, typename T>
__global__ void k_bin_cmp_generic(const T* __restrict__ a, const T* __restrict__ b, uint8_t* __restrict__ out, int64_t n, DimPack d) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t oa, ob;
    unravel2(i, d, oa, ob);
    out[i] = bin_cmp<OP, T>(a[oa], b[ob]);
  }
}

// fp32 fast paths. MODE 0: a[i] op b[i]; 1: a[i] op b[i % C] (row-vector broadcast, C % 4 == 0);
// 2: a[i] op b[0]; 3: b-side of 1/2 swapped is handled by the host through SWAP.
template <int OP, int MODE, bool SWAP>
__global__ void k_bin_f32_fast(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t n4, int64_t n, int64_t C) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float bs = (MODE == 2) ? b[0] : 0.f;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n4; v += stride) {
    float4 x = reinterpret_cast<const float4*>(a)[v];
    float4 y;
    if (MODE == 0) y = reinterpret_cast<const float4*>(b)[v];
    else if (MODE == 1) y = *reinterpret_cast<const float4*>(b + ((v * 4) % C));
    else y = make_float4(bs, bs, bs, bs);
    float4 r;
    if (!SWAP) {
      r.x = bin_arith<OP, float>(x.x, y.x); r.y = bin_arith<OP, float>(x.y, y.y);
      r.z = bin_arith<OP, float>(x.z, y.z); r.w = bin_arith<OP, float>(x.w, y.w);
    } else {
      r.x = bin_arith<OP, float>(y.x, x.x); r.y = bin_arith<OP, float>(y.y, x.y);
      r.z = bin_arith<OP, float>(y.z, x.z); r.w = bin_arith<OP, float>(y.w, x.w);
    }
    reinterpret_cast<float4*>(out)[v] = r;
  }
  // scalar tail (only MODE 0 / 2 can have one; MODE 1 requires C % 4 == 0 so n % 4 == 0)
  for (int64_t t = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
  {
    float x = a[t];
    float y = (MODE == 0) ? b[t] : ((MODE == 1) ? b[t % C] : bs);
    out[t] = SWAP ? bin_arith<OP, float>(y, x) : bin_arith<OP, float>(x, y);
  }
}

template <int OP>
static int launch_bin_f32_fast(int mode, bool swap, const float* a, const float* b, float* out, int64_t n, int64_t C) {
  int64_t n4 = n / 4;
  int grid = nfb_grid_for(n4 > 0 ? n4 : 1, 256);
  cudaStream_t s = nfb_cur_stream();
#define L(M, S) k_bin_f32_fast<OP, M, S><<<grid, 256, 0, s>>>(a, b, out, n4, n, C)
  if (mode == 0) { L(0, false); }
  else if (mode == 1) { if (swap) L(1, true); else L(1, false); }
  else { if (swap) L(2, true); else L(2, false); }
#undef L
  NFB_LAUNCH_CHECK();
  return 0;
}

static bool is_contig(const int64_t* shape, const int64_t* st, int ndim) {
  int64_t e = 1;
  for (int k = ndim - 1; k >= 0; --k) {
    if (shape[k] != 1 && st[k] != e) return false;
    e *= shape[k];
  }
  return true;
}
static bool all_zero(const int64_t* shape, const int64_t* st, int ndim) {
  for (int k = 0; k < ndim; ++k) if (shape[k] != 1 && st[k] != 0) return false;
  return true;
}
// b is a contiguous vector over the trailing `t` dims and broadcast (stride 0) over the rest
static bool is_row_vector(const int64_t* shape, const int64_t* st, int ndim, int64_t* C) {
  int64_t e = 1;
  int k = ndim - 1;
  for (; k >= 0; --k) {
    if (shape[k] == 1) continue;
    if (st[k] == e) { e *= shape[k]; continue; }
    break;
  }
  for (; k >= 0; --k) if (shape[k] != 1 && st[k] != 0) return false;
  *C = e;
  return e > 1;
}

template <typename T>
static int bin_dispatch_generic(int op, const void* a, const void* b, void* out, int64_t n, const DimPack& d) {
  int grid = nfb_grid_for(n, 256);
  cudaStream_t s = nfb_cur_stream();
#define A(OPC) case OPC: k_bin_arith_generic<OPC, T><<<grid, 256, 0, s>>>((const T*)a, (const T*)b, (T*)out, n, d); break;
#define C(OPC) case OPC: k_bin_cmp_generic<OPC, T><<<grid, 256, 0, s>>>((const T*)a, (const T*)b, (uint8_t*)out, n, d); break;
  switch (op) {
    A(B_ADD) A(B_SUB) A(B_MUL) A(B_DIV) A(B_POW) A(B_MAX) A(B_MIN)
    C(B_EQ) C(B_NE) C(B_LT) C(B_LE) C(B_GT) C(B_GE)
    default: return NFB_FAIL("nfb_binary: unknown op %d", op);
  }
#undef A
#undef C
  NFB_LAUNCH_CHECK();
  return 0;
}

// out[shape] = a (strides sa) OP b (strides sb); out is contiguous. Comparison ops write bool (u8).
NFB_API int nfb_binary(int op, int dtype, const void* a, const void* b, void* out, int ndim,
                       const int64_t* shape, const int64_t* sa, const int64_t* sb) {
  if (ndim > NFB_MAX_DIMS) return NFB_FAIL("nfb_binary: ndim %d > %d", ndim, NFB_MAX_DIMS);
  int64_t n = 1;
  for (int k = 0; k < ndim; ++k) n *= shape[k];
  if (n == 0) return 0;
  if (dtype == NFB_F32 && op <= B_MIN) {
    bool ca = is_contig(shape, sa, ndim), cb = is_contig(shape, sb, ndim);
    int mode = -1;
    bool swap = false;
    int64_t C = 1;
    const float *pa = (const float*)a, *pb = (const float*)b;
    if (ca && cb) mode = 0;
    else if (ca && all_zero(shape, sb, ndim)) mode = 2;
    else if (cb && all_zero(shape, sa, ndim)) { mode = 2; swap = true; }
    else if (ca && is_row_vector(shape, sb, ndim, &C) && C % 4 == 0 && ((uintptr_t)b % 16 == 0)) mode = 1;
    else if (cb && is_row_vector(shape, sa, ndim, &C) && C % 4 == 0 && ((uintptr_t)a % 16 == 0)) { mode = 1; swap = true; }
    if (mode >= 0 && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) && ((uintptr_t)out % 16 == 0)) {
      if (swap) { const float* t = pa; pa = pb; pb = t; }
#define F(OPC) case OPC: return launch_bin_f32_fast<OPC>(mode, swap, pa, pb, (float*)out, n, C);
      switch (op) { F(B_ADD) F(B_SUB) F(B_MUL) F(B_DIV) F(B_POW) F(B_MAX) F(B_MIN) }
#undef F
    }
  }
  DimPack d;
  d.ndim = ndim;
  for (int k = 0; k < ndim; ++k) { d.shape[k] = shape[k]; d.sa[k] = sa[k]; d.sb[k] = sb[k]; d.sc[k] = 0; }
  switch (dtype) {
    case NFB_F32: return bin_dispatch_generic<float>(op, a, b, out, n, d);
    case NFB_F64: return bin_dispatch_generic<double>(op, a, b, out, n, d);
    case NFB_I32: return bin_dispatch_generic<int32_t>(op, a, b, out, n, d);
    case NFB_I64: return bin_dispatch_generic<int64_t>(op, a, b, out, n, d);
    case NFB_BOOL: return bin_dispatch_generic<uint8_t>(op, a, b, out, n, d);
  }
  return NFB_FAIL("nfb_binary: unsupported dtype %d", dtype);
}

// ------------------------------------------------------------------------------------------------
// unary
// ------------------------------------------------------------------------------------------------
enum UnOp {
  U_NEG = 0, U_EXP, U_LOG, U_SQRT, U_ABS, U_SIGN, U_TANH, U_SIGMOID, U_RELU, U_GELU_ERF, U_GELU_TANH,
  U_SILU, U_SQUARE, U_RECIP, U_CLIP, U_AFFINE, U_POWS, U_RSQRT, U_NAN_TO_NUM, U_SIN, U_COS, U_FLOOR
};

template <typename T> struct Math;
template <> struct Math<float> {
  static __device__ __forceinline__ float exp(float x) { return expf(x); }
  static __device__ __forceinline__ float log(float x) { return logf(x); }
  static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float tanh(float x) { return tanhf(x); }
  static __device__ __forceinline__ float erf(float x) { return erff(x); }
  static __device__ __forceinline__ float pow(float x, float y) { return powf(x, y); }
  static __device__ __forceinline__ float sin(float x) { return sinf(x); }
  static __device__ __forceinline__ float cos(float x) { return cosf(x); }
  static __device__ __forceinline__ float floor(float x) { return floorf(x); }
};
template <> struct Math<double> {
  static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
  static __device__ __forceinline__ double log(double x) { return ::log(x); }
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static __device__ __forceinline__ double tanh(double x) { return ::tanh(x); }
  static __device__ __forceinline__ double erf(double x) { return ::erf(x); }
  static __device__ __forceinline__ double pow(double x, double y) { return ::pow(x, y); }
  static __device__ __forceinline__ double sin(double x) { return ::sin(x); }
  static __device__ __forceinline__ double cos(double x) { return ::cos(x); }
  static __device__ __forceinline__ double floor(double x) { return ::floor(x); }
};

// sigmoid in the reference's split form (functional/activation.py:151-214): never overflows
template <typename T> __device__ __forceinline__ T sigmoid_t(T x) {
  if (x >= T(0)) return T(1) / (T(1) + Math<T>::exp(-x));
  T e = Math<T>::exp(x);
  return e / (T(1) + e);
}

template <int OP, typename T> __device__ __forceinline__ T un_apply(T x, T p0, T p1) {
  typedef Math<T> M;
  switch (OP) {
    case U_NEG: return -x;
    case U_EXP: return M::exp(x);
    case U_LOG: return M::log(x);
    case U_SQRT: return M::sqrt(x);
    case U_ABS: return x < T(0) ? -x : x;
    case U_SIGN: return (x > T(0)) ? T(1) : ((x < T(0)) ? T(-1) : x);  // keeps +-0 and NaN like numpy
    case U_TANH: return M::tanh(x);
    case U_SIGMOID: return sigmoid_t<T>(x);
    case U_RELU: return (x != x) ? x : (x > T(0) ? x : T(0));
    case U_GELU_ERF: return T(0.5) * x * (T(1) + M::erf(x * T(0.70710678118654752440)));
    case U_GELU_TANH: {
      T inner = T(0.79788456080286535588) * (x + T(0.044715) * x * x * x);
      return T(0.5) * x * (T(1) + M::tanh(inner));
    }
    case U_SILU: return x * sigmoid_t<T>(x);
    case U_SQUARE: return x * x;
    case U_RECIP: return T(1) / x;
    case U_CLIP: return (x != x) ? x : (x < p0 ? p0 : (x > p1 ? p1 : x));
    case U_AFFINE: return x * p0 + p1;
    case U_POWS: return M::pow(x, p0);
    case U_RSQRT: return T(1) / M::sqrt(x);
    case U_NAN_TO_NUM: return (x != x) ? T(0) : x;  // +-inf handled by the caller through U_CLIP
    case U_SIN: return M::sin(x);
    case U_COS: return M::cos(x);
    case U_FLOOR: return M::floor(x);
  }
  return x;
}

template <int OP>
__global__ void k_unary_f32(const float* __restrict__ x, float* __restrict__ out, int64_t n4, int64_t n, float p0, float p1) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n4; v += stride) {
    float4 a = reinterpret_cast<const float4*>(x)[v];
    float4 r;
    r.x = un_apply<OP, float>(a.x, p0, p1); r.y = un_apply<OP, float>(a.y, p0, p1);
    r.z = un_apply<OP, float>(a.z, p0, p1); r.w = un_apply<OP, float>(a.w, p0, p1);
    reinterpret_cast<float4*>(out)[v] = r;
  }
  for (int64_t t = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
  out[t] = un_apply<OP, float>(x[t], p0, p1);
}
template <int OP, typename T>
__global__ void k_unary_any(const T* __restrict__ x, T* __restrict__ out, int64_t n, T p0, T p1) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = un_apply<OP, T>(x[i], p0, p1);
}
template <typename T>
__global__ void k_unary_int(int op, const T* __restrict__ x, T* __restrict__ out, int64_t n, T p0, T p1) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T v = x[i];
    T r = v;
    if (op == U_NEG) r = -v;
    else if (op == U_ABS) r = v < 0 ? -v : v;
    else if (op == U_SIGN) r = (v > 0) - (v < 0);
    else if (op == U_SQUARE) r = v * v;
    else if (op == U_CLIP) r = v < p0 ? p0 : (v > p1 ? p1 : v);
    else if (op == U_RELU) r = v > 0 ? v : 0;
    else if (op == U_AFFINE) r = v * p0 + p1;
    out[i] = r;
  }
}

NFB_API int nfb_unary(int op, int dtype, const void* x, void* out, int64_t n, double p0, double p1) {
  if (n == 0) return 0;
  cudaStream_t s = nfb_cur_stream();
  if (dtype == NFB_F32) {
    bool vec = ((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0);
    int64_t n4 = vec ? n / 4 : 0;
    int grid = nfb_grid_for(vec ? (n4 > 0 ? n4 : 1) : n, 256);
    if (!vec) {
#define G(OPC) case OPC: k_unary_any<OPC, float><<<grid, 256, 0, s>>>((const float*)x, (float*)out, n, (float)p0, (float)p1); break;
      switch (op) {
        G(U_NEG) G(U_EXP) G(U_LOG) G(U_SQRT) G(U_ABS) G(U_SIGN) G(U_TANH) G(U_SIGMOID) G(U_RELU) G(U_GELU_ERF)
        G(U_GELU_TANH) G(U_SILU) G(U_SQUARE) G(U_RECIP) G(U_CLIP) G(U_AFFINE) G(U_POWS) G(U_RSQRT) G(U_NAN_TO_NUM)
        G(U_SIN) G(U_COS) G(U_FLOOR)
        default: return NFB_FAIL("nfb_unary: unknown op %d", op);
      }
#undef G
    } else {
#define V(OPC) case OPC: k_unary_f32<OPC><<<grid, 256, 0, s>>>((const float*)x, (float*)out, n4, n, (float)p0, (float)p1); break;
      switch (op) {
        V(U_NEG) V(U_EXP) V(U_LOG) V(U_SQRT) V(U_ABS) V(U_SIGN) V(U_TANH) V(U_SIGMOID) V(U_RELU) V(U_GELU_ERF)
        V(U_GELU_TANH) V(U_SILU) V(U_SQUARE) V(U_RECIP) V(U_CLIP) V(U_AFFINE) V(U_POWS) V(U_RSQRT) V(U_NAN_TO_NUM)
        V(U_SIN) V(U_COS) V(U_FLOOR)
        default: return NFB_FAIL("nfb_unary: unknown op %d", op);
      }
#undef V
    }
    NFB_LAUNCH_CHECK();
    return 0;
  }
  int grid = nfb_grid_for(n, 256);
  if (dtype == NFB_F64) {
#define D(OPC) case OPC: k_unary_any<OPC, double><<<grid, 256, 0, s>>>((const double*)x, (double*)out, n, p0, p1); break;
    switch (op) {
      D(U_NEG) D(U_EXP) D(U_LOG) D(U_SQRT) D(U_ABS) D(U_SIGN) D(U_TANH) D(U_SIGMOID) D(U_RELU) D(U_GELU_ERF)
      D(U_GELU_TANH) D(U_SILU) D(U_SQUARE) D(U_RECIP) D(U_CLIP) D(U_AFFINE) D(U_POWS) D(U_RSQRT) D(U_NAN_TO_NUM)
      D(U_SIN) D(U_COS) D(U_FLOOR)
      default: return NFB_FAIL("nfb_unary: unknown op %d", op);
    }
#undef D
    NFB_LAUNCH_CHECK();
    return 0;
  }
  if (op != U_NEG && op != U_ABS && op != U_SIGN && op != U_SQUARE && op != U_CLIP && op != U_RELU && op != U_AFFINE)
    return NFB_FAIL("nfb_unary: op %d needs a floating dtype (got %d)", op, dtype);
  if (dtype == NFB_I32) k_unary_int<int32_t><<<grid, 256, 0, s>>>(op, (const int32_t*)x, (int32_t*)out, n, (int32_t)p0, (int32_t)p1);
  else if (dtype == NFB_I64) k_unary_int<int64_t><<<grid, 256, 0, s>>>(op, (const int64_t*)x, (int64_t*)out, n, (int64_t)p0, (int64_t)p1);
  else return NFB_FAIL("nfb_unary: unsupported dtype %d", dtype);
  NFB_LAUNCH_CHECK();
  return 0;
}

// ---- backward forms: out = g * f'(.) ; `src` is x for relu/gelu/silu and y for tanh/sigmoid ----
template <int OP> __device__ __forceinline__ float un_grad(float s, float g) {
  switch (OP) {
    case U_RELU: return s > 0.f ? g : 0.f;
    case U_TANH: return g * (1.f - s * s);
    case U_SIGMOID: return g * s * (1.f - s);
    case U_GELU_ERF: {  // functional/activation.py:447-462
      float cdf = 0.5f * (1.f + erff(s * 0.70710678118654752440f));
      float pdf = expf(-0.5f * s * s) * 0.39894228040143267794f;
      return g * (cdf + s * pdf);
    }
    case U_GELU_TANH: {  // functional/activation.py:389-396 ; nn/optimized.py:194-205
      float inner = 0.79788456080286535588f * (s + 0.044715f * s * s * s);
      float t = tanhf(inner);
      float dinner = 0.79788456080286535588f * (1.f + 3.f * 0.044715f * s * s);
      return g * (0.5f * (1.f + t) + 0.5f * s * (1.f - t * t) * dinner);
    }
    case U_SILU: {
      float sg = sigmoid_t<float>(s);
      return g * (sg * (1.f + s * (1.f - sg)));
    }
  }
  return g;
}
template <int OP, typename TS, typename TG, typename TO>
__global__ void k_unary_bwd(const TS* __restrict__ src, const TG* __restrict__ g, TO* __restrict__ out, int64_t n4, int64_t n) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n4; v += stride) {
    f4 a = ld4<TS>(src + v * 4), b = ld4<TG>(g + v * 4);
    f4 r = {un_grad<OP>(a.x, b.x), un_grad<OP>(a.y, b.y), un_grad<OP>(a.z, b.z), un_grad<OP>(a.w, b.w)};
    st4<TO>(out + v * 4, r);
  }
  for (int64_t t = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
  st_f<TO>(out, t, un_grad<OP>(ld_f<TS>(src, t), ld_f<TG>(g, t)));
}

template <typename TS, typename TG, typename TO>
static int unary_bwd_launch(int op, const void* src, const void* g, void* out, int64_t n) {
  bool vec = ((uintptr_t)src % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)out % 16 == 0);
  int64_t n4 = vec ? n / 4 : 0;
  int grid = nfb_grid_for(vec && n4 ? n4 : n, 256);
  cudaStream_t s = nfb_cur_stream();
#define W(OPC) case OPC: k_unary_bwd<OPC, TS, TG, TO><<<grid, 256, 0, s>>>((const TS*)src, (const TG*)g, (TO*)out, n4, n); break;
  switch (op) {
    W(U_RELU) W(U_TANH) W(U_SIGMOID) W(U_GELU_ERF) W(U_GELU_TANH) W(U_SILU)
    default: return NFB_FAIL("nfb_unary_bwd: unsupported op %d", op);
  }
#undef W
  NFB_LAUNCH_CHECK();
  return 0;
}
// dtype applies to src, g and out alike (NFB_F32 or NFB_BF16)
NFB_API int nfb_unary_bwd(int op, int dtype, const void* src, const void* g, void* out, int64_t n) {
  if (n == 0) return 0;
  if (dtype == NFB_F32) return unary_bwd_launch<float, float, float>(op, src, g, out, n);
  if (dtype == NFB_BF16) return unary_bwd_launch<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(op, src, g, out, n);
  return NFB_FAIL("nfb_unary_bwd: unsupported dtype %d", dtype);
}

// ------------------------------------------------------------------------------------------------
// cast
// ------------------------------------------------------------------------------------------------
template <typename S, typename D> __device__ __forceinline__ D cast_one(S v) { return (D)v; }
template <> __device__ __forceinline__ __nv_bfloat16 cast_one<float, __nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ float cast_one<__nv_bfloat16, float>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ uint8_t cast_one<float, uint8_t>(float v) { return v != 0.f; }
template <> __device__ __forceinline__ uint8_t cast_one<double, uint8_t>(double v) { return v != 0.0; }
template <> __device__ __forceinline__ uint8_t cast_one<int32_t, uint8_t>(int32_t v) { return v != 0; }
template <> __device__ __forceinline__ uint8_t cast_one<int64_t, uint8_t>(int64_t v) { return v != 0; }

template <typename S, typename D>
__global__ void k_cast(const S* __restrict__ src, D* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = cast_one<S, D>(src[i]);
}
// vectorised fp32 <-> bf16 (the casts on the training hot path)
__global__ void k_cast_f32_bf16(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n4, int64_t n) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n4; v += stride)
    st4<__nv_bfloat16>(dst + v * 4, ld4<float>(src + v * 4));
  for (int64_t t = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
  dst[t] = __float2bfloat16_rn(src[t]);
}
__global__ void k_cast_bf16_f32(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t n4, int64_t n) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n4; v += stride)
    st4<float>(dst + v * 4, ld4<__nv_bfloat16>(src + v * 4));
  for (int64_t t = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
  dst[t] = __bfloat162float(src[t]);
}

template <typename S>
static int cast_from(int dd, const void* src, void* dst, int64_t n) {
  int grid = nfb_grid_for(n, 256);
  cudaStream_t s = nfb_cur_stream();
  switch (dd) {
    case NFB_F32: k_cast<S, float><<<grid, 256, 0, s>>>((const S*)src, (float*)dst, n); break;
    case NFB_F64: k_cast<S, double><<<grid, 256, 0, s>>>((const S*)src, (double*)dst, n); break;
    case NFB_I32: k_cast<S, int32_t><<<grid, 256, 0, s>>>((const S*)src, (int32_t*)dst, n); break;
    case NFB_I64: k_cast<S, int64_t><<<grid, 256, 0, s>>>((const S*)src, (int64_t*)dst, n); break;
    case NFB_BOOL: k_cast<S, uint8_t><<<grid, 256, 0, s>>>((const S*)src, (uint8_t*)dst, n); break;
    default: return NFB_FAIL("nfb_cast: unsupported destination dtype %d", dd);
  }
  NFB_LAUNCH_CHECK();
  return 0;
}

NFB_API int nfb_cast(int sd, int dd, const void* src, void* dst, int64_t n) {
  if (n == 0) return 0;
  cudaStream_t s = nfb_cur_stream();
  if (sd == NFB_F32 && dd == NFB_BF16) {
    bool vec = ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 8 == 0);
    int64_t n4 = vec ? n / 4 : 0;
    k_cast_f32_bf16<<<nfb_grid_for(vec && n4 ? n4 : n, 256), 256, 0, s>>>((const float*)src, (__nv_bfloat16*)dst, n4, n);
    NFB_LAUNCH_CHECK();
    return 0;
  }
  if (sd == NFB_BF16 && dd == NFB_F32) {
    bool vec = ((uintptr_t)src % 8 == 0) && ((uintptr_t)dst % 16 == 0);
    int64_t n4 = vec ? n / 4 : 0;
    k_cast_bf16_f32<<<nfb_grid_for(vec && n4 ? n4 : n, 256), 256, 0, s>>>((const __nv_bfloat16*)src, (float*)dst, n4, n);
    NFB_LAUNCH_CHECK();
    return 0;
  }
  switch (sd) {
    case NFB_F32: return cast_from<float>(dd, src, dst, n);
    case NFB_F64: return cast_from<double>(dd, src, dst, n);
    case NFB_I32: return cast_from<int32_t>(dd, src, dst, n);
    case NFB_I64: return cast_from<int64_t>(dd, src, dst, n);
    case NFB_BOOL: return cast_from<uint8_t>(dd, src, dst, n);
  }
  return NFB_FAIL("nfb_cast: unsupported source dtype %d", sd);
}

// ------------------------------------------------------------------------------------------------
// where(cond, x, y) with broadcasting strides
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_where(const uint8_t* __restrict__ c, const T* __restrict__ x, const T* __restrict__ y, T* __restrict__ out, int64_t n, DimPack d) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t idx = i, oc = 0, ox = 0, oy = 0;
    for (int k = d.ndim - 1; k >= 0; --k) {
      int64_t s = d.shape[k];
      int64_t q = idx % s;
      idx /= s;
      oc += q * d.sc[k];
      ox += q * d.sa[k];
      oy += q * d.sb[k];
    }
    out[i] = c[oc] ? x[ox] : y[oy];
  }
}
NFB_API int nfb_where(int dtype, const void* cond, const int64_t* scond, const void* x, const int64_t* sx, const void* y,
                      const int64_t* sy, void* out, int ndim, const int64_t* shape) {
  if (ndim > NFB_MAX_DIMS) return NFB_FAIL("nfb_where: ndim %d too large", ndim);
  DimPack d;
  d.ndim = ndim;
  int64_t n = 1;
  for (int k = 0; k < ndim; ++k) { d.shape[k] = shape[k]; d.sa[k] = sx[k]; d.sb[k] = sy[k]; d.sc[k] = scond[k]; n *= shape[k]; }
  if (n == 0) return 0;
  int grid = nfb_grid_for(n, 256);
  cudaStream_t s = nfb_cur_stream();
  switch (dtype) {
    case NFB_F32: case NFB_I32: k_where<uint32_t><<<grid, 256, 0, s>>>((const uint8_t*)cond, (const uint32_t*)x, (const uint32_t*)y, (uint32_t*)out, n, d); break;
    case NFB_F64: case NFB_I64: k_where<uint64_t><<<grid, 256, 0, s>>>((const uint8_t*)cond, (const uint64_t*)x, (const uint64_t*)y, (uint64_t*)out, n, d); break;
    case NFB_BOOL: k_where<uint8_t><<<grid, 256, 0, s>>>((const uint8_t*)cond, (const uint8_t*)x, (const uint8_t*)y, (uint8_t*)out, n, d); break;
    case NFB_BF16: k_where<uint16_t><<<grid, 256, 0, s>>>((const uint8_t*)cond, (const uint16_t*)x, (const uint16_t*)y, (uint16_t*)out, n, d); break;
    default: return NFB_FAIL("nfb_where: unsupported dtype %d", dtype);
  }
  NFB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// strided copies: dst[dst strides] = src[src strides] over `shape` (element size 1/2/4/8 bytes)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_copy_strided(const T* __restrict__ src, T* __restrict__ dst, int64_t n, DimPack d) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t os, od;
    unravel2(i, d, os, od);
    dst[od] = src[os];
  }
}
// batched 2-D transpose through shared memory: src has unit stride along dim `r` (rows of the
// output tile), dst is contiguous with last dim `c`. Used when the source's unit-stride dimension
// is not the output's last dimension (x.T, k.transpose(0,1,3,2), ...).
template <typename T>
__global__ void k_transpose_tiled(const T* __restrict__ src, T* __restrict__ dst, int64_t R, int64_t Ccols,
                                  int64_t src_sr, int64_t src_sc, int64_t dst_sr, int64_t dst_sc, DimPack batch) {
  __shared__ T tile[32][33];
  int64_t bz = blockIdx.z;
  int64_t os, od;
  unravel2(bz, batch, os, od);
  int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  // read: unit stride along r
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t r = r0 + threadIdx.x, c = c0 + j;
    if (r < R && c < Ccols) tile[j][threadIdx.x] = src[os + r * src_sr + c * src_sc];
  }
  __syncthreads();
  // write: unit stride along c
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t r = r0 + j, c = c0 + threadIdx.x;
    if (r < R && c < Ccols) dst[od + r * dst_sr + c * dst_sc] = tile[threadIdx.x][j];
  }
}

template <typename T>
static int copy_strided_t(const void* src, void* dst, int ndim, const int64_t* shape, const int64_t* ss, const int64_t* ds) {
  int64_t n = 1;
  for (int k = 0; k < ndim; ++k) n *= shape[k];
  if (n == 0) return 0;
  cudaStream_t s = nfb_cur_stream();
  // transpose fast path: dst last dim has unit stride, src has unit stride on another dim r with both extents >= 16
  if (ndim >= 2 && ds[ndim - 1] == 1 && ss[ndim - 1] != 1 && shape[ndim - 1] >= 16) {
    int r = -1;
    for (int k = 0; k < ndim - 1; ++k) if (ss[k] == 1 && shape[k] >= 16) { r = k; break; }
    if (r >= 0) {
      DimPack b;
      b.ndim = 0;
      int64_t nb = 1;
      for (int k = 0; k < ndim - 1; ++k) {
        if (k == r) continue;
        b.shape[b.ndim] = shape[k]; b.sa[b.ndim] = ss[k]; b.sb[b.ndim] = ds[k]; b.sc[b.ndim] = 0;
        nb *= shape[k];
        ++b.ndim;
      }
      if (nb <= 65535) {
        dim3 grid((unsigned)nfb_cdiv(shape[ndim - 1], 32), (unsigned)nfb_cdiv(shape[r], 32), (unsigned)nb);
        if (grid.y <= 65535) {
          k_transpose_tiled<T><<<grid, dim3(32, 8), 0, s>>>((const T*)src, (T*)dst, shape[r], shape[ndim - 1], ss[r], ss[ndim - 1], ds[r], ds[ndim - 1], b);
          NFB_LAUNCH_CHECK();
          return 0;
        }
      }
    }
  }
  DimPack d;
  d.ndim = ndim;
  for (int k = 0; k < ndim; ++k) { d.shape[k] = shape[k]; d.sa[k] = ss[k]; d.sb[k] = ds[k]; d.sc[k] = 0; }
  k_copy_strided<T><<<nfb_grid_for(n, 256), 256, 0, s>>>((const T*)src, (T*)dst, n, d);
  NFB_LAUNCH_CHECK();
  return 0;
}

NFB_API int nfb_copy_strided(int elem_size, const void* src, void* dst, int ndim, const int64_t* shape,
                             const int64_t* src_strides, const int64_t* dst_strides) {
  if (ndim > NFB_MAX_DIMS) return NFB_FAIL("nfb_copy_strided: ndim %d too large", ndim);
  switch (elem_size) {
    case 1: return copy_strided_t<uint8_t>(src, dst, ndim, shape, src_strides, dst_strides);
    case 2: return copy_strided_t<uint16_t>(src, dst, ndim, shape, src_strides, dst_strides);
    case 4: return copy_strided_t<uint32_t>(src, dst, ndim, shape, src_strides, dst_strides);
    case 8: return copy_strided_t<uint64_t>(src, dst, ndim, shape, src_strides, dst_strides);
  }
  return NFB_FAIL("nfb_copy_strided: unsupported element size %d", elem_size);
}

// ------------------------------------------------------------------------------------------------
// fill / arange
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_fill(T* __restrict__ p, int64_t n, T v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
template <typename T>
__global__ void k_arange(T* __restrict__ p, int64_t n, double start, double step) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = (T)(start + step * (double)i);
}
NFB_API int nfb_fill(int dtype, void* p, int64_t n, double value) {
  if (n == 0) return 0;
  int grid = nfb_grid_for(n, 256);
  cudaStream_t s = nfb_cur_stream();
  switch (dtype) {
    case NFB_F32: k_fill<float><<<grid, 256, 0, s>>>((float*)p, n, (float)value); break;
    case NFB_F64: k_fill<double><<<grid, 256, 0, s>>>((double*)p, n, value); break;
    case NFB_I32: k_fill<int32_t><<<grid, 256, 0, s>>>((int32_t*)p, n, (int32_t)value); break;
    case NFB_I64: k_fill<int64_t><<<grid, 256, 0, s>>>((int64_t*)p, n, (int64_t)value); break;
    case NFB_BOOL: k_fill<uint8_t><<<grid, 256, 0, s>>>((uint8_t*)p, n, (uint8_t)(value != 0)); break;
    case NFB_BF16: k_fill<__nv_bfloat16><<<grid, 256, 0, s>>>((__nv_bfloat16*)p, n, __float2bfloat16((float)value)); break;
    default: return NFB_FAIL("nfb_fill: unsupported dtype %d", dtype);
  }
  NFB_LAUNCH_CHECK();
  return 0;
}
NFB_API int nfb_arange(int dtype, void* p, int64_t n, double start, double step) {
  if (n == 0) return 0;
  int grid = nfb_grid_for(n, 256);
  cudaStream_t s = nfb_cur_stream();
  switch (dtype) {
    case NFB_F32: k_arange<float><<<grid, 256, 0, s>>>((float*)p, n, start, step); break;
    case NFB_F64: k_arange<double><<<grid, 256, 0, s>>>((double*)p, n, start, step); break;
    case NFB_I32: k_arange<int32_t><<<grid, 256, 0, s>>>((int32_t*)p, n, start, step); break;
    case NFB_I64: k_arange<int64_t><<<grid, 256, 0, s>>>((int64_t*)p, n, start, step); break;
    default: return NFB_FAIL("nfb_arange: unsupported dtype %d", dtype);
  }
  NFB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// fused residual chains and SwiGLU (bandwidth-bound pieces of the Transformer block)
// ------------------------------------------------------------------------------------------------
// out = silu(gate) * up   (models/language/gpt2.py:119-139) ; gate/up are the two halves of one
// (rows, 2*C) buffer when `packed`, else separate (rows, C) buffers.
// sigmoid for the SwiGLU kernels: exact split form for fp32 storage (the 1e-5 contract); for bf16 storage the
// 4-instruction form 1 / (1 + 2^(-x log2 e)) (ex2.approx + rcp.approx, ~2 ulp: invisible after bf16 rounding) -- the exact
// form costs ~30 instructions per element and made these kernels issue-bound instead of HBM-bound
template <typename T> __device__ __forceinline__ float swiglu_sigmoid(float x) { return sigmoid_t<float>(x); }
template <> __device__ __forceinline__ float swiglu_sigmoid<__nv_bfloat16>(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// 16-byte vectors (4 fp32 / 8 bf16 per access), 32-bit index math, one (row, column-vector) per thread and iteration.
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float* v) { float4 a = *reinterpret_cast<const float4*>(p); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
  __device__ static void store(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float* v) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) { float2 f = __bfloat1622float2(h[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
  }
  __device__ static void store(__nv_bfloat16* p, const float* v) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

template <typename T>
__global__ void k_swiglu_fwd(const T* __restrict__ gate, const T* __restrict__ up, T* __restrict__ out, int rows, int CV, int64_t C, int64_t in_ld) {
  nfb_pdl_enter();
  constexpr int N = Vec16<T>::N;
  const unsigned total = (unsigned)rows * (unsigned)CV, stride = gridDim.x * blockDim.x;   // host checks rows * CV < 2^31
  for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < total; v += stride) {
    const int r = (int)(v / (unsigned)CV), c = (int)(v - (unsigned)r * (unsigned)CV) * N;
    float g[N], u[N], o[N];
    Vec16<T>::load(gate + r * in_ld + c, g);
    Vec16<T>::load(up + r * in_ld + c, u);
#pragma unroll
    for (int k = 0; k < N; ++k) o[k] = g[k] * swiglu_sigmoid<T>(g[k]) * u[k];
    Vec16<T>::store(out + (int64_t)r * C + c, o);
  }
}
// dgate = dout * up * silu'(gate) ; dup = dout * silu(gate)
template <typename T>
__global__ void k_swiglu_bwd(const T* __restrict__ gate, const T* __restrict__ up, const T* __restrict__ dout, T* __restrict__ dgate, T* __restrict__ dup,
                             int rows, int CV, int64_t C, int64_t in_ld) {
  nfb_pdl_enter();
  constexpr int N = Vec16<T>::N;
  const unsigned total = (unsigned)rows * (unsigned)CV, stride = gridDim.x * blockDim.x;   // host checks rows * CV < 2^31
  for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < total; v += stride) {
    const int r = (int)(v / (unsigned)CV), c = (int)(v - (unsigned)r * (unsigned)CV) * N;
    float g[N], u[N], d[N], dg[N], du[N];
    Vec16<T>::load(gate + r * in_ld + c, g);
    Vec16<T>::load(up + r * in_ld + c, u);
    Vec16<T>::load(dout + (int64_t)r * C + c, d);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      float sg = swiglu_sigmoid<T>(g[k]);
      du[k] = d[k] * g[k] * sg;
      dg[k] = d[k] * u[k] * (sg * (1.f + g[k] * (1.f - sg)));
    }
    Vec16<T>::store(dgate + r * in_ld + c, dg);
    Vec16<T>::store(dup + r * in_ld + c, du);
  }
}
// scalar fallbacks for shapes / pointers the 16-byte vectors cannot address
template <typename T>
__global__ void k_swiglu_fwd_scalar(const T* __restrict__ gate, const T* __restrict__ up, T* __restrict__ out, int64_t rows, int64_t C, int64_t in_ld) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows * C; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / C, c = i % C;
    float g = ld_f<T>(gate, r * in_ld + c), u = ld_f<T>(up, r * in_ld + c);
    st_f<T>(out, i, g * sigmoid_t<float>(g) * u);
  }
}
template <typename T>
__global__ void k_swiglu_bwd_scalar(const T* __restrict__ gate, const T* __restrict__ up, const T* __restrict__ dout, T* __restrict__ dgate,
                                    T* __restrict__ dup, int64_t rows, int64_t C, int64_t in_ld) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows * C; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / C, c = i % C;
    float g = ld_f<T>(gate, r * in_ld + c), u = ld_f<T>(up, r * in_ld + c), d = ld_f<T>(dout, i);
    float sg = sigmoid_t<float>(g);
    st_f<T>(dup, r * in_ld + c, d * g * sg);
    st_f<T>(dgate, r * in_ld + c, d * u * (sg * (1.f + g * (1.f - sg))));
  }
}
template <typename T>
static bool swiglu_vec_ok(const void* a, const void* b, const void* c, const void* d, const void* e, int64_t rows, int64_t C, int64_t in_ld) {
  constexpr int N = 16 / (int)sizeof(T);
  return C % N == 0 && in_ld % N == 0 && ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d | (uintptr_t)e) & 15) == 0) && rows * (C / N) < (1LL << 31);
}
template <typename T>
static int swiglu_check(const void* a, const void* b, const void* c, const void* d, const void* e, int64_t C, int64_t in_ld, const char* who) {
  constexpr int N = 16 / (int)sizeof(T);
  if (C % N || in_ld % N) return NFB_FAIL("%s: C and in_ld must be multiples of %d elements (16-byte vectors)", who, N);
  if ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d | (uintptr_t)e) & 15) != 0) return NFB_FAIL("%s: buffers must be 16-byte aligned", who);
  return 0;
}
static int swiglu_size_check(int64_t rows, int64_t CV, const char* who) {
  if (rows * CV >= (1LL << 31)) return NFB_FAIL("%s: %lld x %lld vectors exceed the 32-bit index range", who, (long long)rows, (long long)CV);
  return 0;
}
NFB_API int nfb_swiglu_fwd(int dtype, const void* gate, const void* up, void* out, int64_t rows, int64_t C, int64_t in_ld) {
  if (rows == 0 || C == 0) return 0;
  if (dtype == NFB_F32 && !swiglu_vec_ok<float>(gate, up, out, nullptr, nullptr, rows, C, in_ld)) {
    k_swiglu_fwd_scalar<float><<<nfb_grid_for(rows * C, 256), 256, 0, nfb_cur_stream()>>>((const float*)gate, (const float*)up, (float*)out, rows, C, in_ld);
  } else if (dtype == NFB_BF16 && !swiglu_vec_ok<__nv_bfloat16>(gate, up, out, nullptr, nullptr, rows, C, in_ld)) {
    k_swiglu_fwd_scalar<__nv_bfloat16><<<nfb_grid_for(rows * C, 256), 256, 0, nfb_cur_stream()>>>((const __nv_bfloat16*)gate, (const __nv_bfloat16*)up, (__nv_bfloat16*)out, rows, C, in_ld);
  } else if (dtype == NFB_F32) {
    if (swiglu_check<float>(gate, up, out, nullptr, nullptr, C, in_ld, "nfb_swiglu_fwd")) return -1;
    int CV = (int)(C / 4);
    if (swiglu_size_check(rows, CV, "nfb_swiglu_fwd")) return -1;
    nfb_launch_pdl(k_swiglu_fwd<float>, dim3(nfb_grid_for(rows * CV, 256)), dim3(256), 0, nfb_cur_stream(), (const float*)gate, (const float*)up, (float*)out, (int)rows, CV, C, in_ld);
  } else if (dtype == NFB_BF16) {
    if (swiglu_check<__nv_bfloat16>(gate, up, out, nullptr, nullptr, C, in_ld, "nfb_swiglu_fwd")) return -1;
    int CV = (int)(C / 8);
    if (swiglu_size_check(rows, CV, "nfb_swiglu_fwd")) return -1;
    nfb_launch_pdl(k_swiglu_fwd<__nv_bfloat16>, dim3(nfb_grid_for(rows * CV, 256)), dim3(256), 0, nfb_cur_stream(), (const __nv_bfloat16*)gate, (const __nv_bfloat16*)up, (__nv_bfloat16*)out, (int)rows, CV, C, in_ld);
  } else return NFB_FAIL("nfb_swiglu_fwd: unsupported dtype %d", dtype);
  NFB_LAUNCH_CHECK();
  return 0;
}
NFB_API int nfb_swiglu_bwd(int dtype, const void* gate, const void* up, const void* dout, void* dgate, void* dup, int64_t rows, int64_t C, int64_t in_ld) {
  if (rows == 0 || C == 0) return 0;
  if (dtype == NFB_F32 && !swiglu_vec_ok<float>(gate, up, dout, dgate, dup, rows, C, in_ld)) {
    k_swiglu_bwd_scalar<float><<<nfb_grid_for(rows * C, 256), 256, 0, nfb_cur_stream()>>>((const float*)gate, (const float*)up, (const float*)dout, (float*)dgate, (float*)dup, rows, C, in_ld);
  } else if (dtype == NFB_BF16 && !swiglu_vec_ok<__nv_bfloat16>(gate, up, dout, dgate, dup, rows, C, in_ld)) {
    k_swiglu_bwd_scalar<__nv_bfloat16><<<nfb_grid_for(rows * C, 256), 256, 0, nfb_cur_stream()>>>((const __nv_bfloat16*)gate, (const __nv_bfloat16*)up, (const __nv_bfloat16*)dout, (__nv_bfloat16*)dgate, (__nv_bfloat16*)dup, rows, C, in_ld);
  } else if (dtype == NFB_F32) {
    if (swiglu_check<float>(gate, up, dout, dgate, dup, C, in_ld, "nfb_swiglu_bwd")) return -1;
    int CV = (int)(C / 4);
    if (swiglu_size_check(rows, CV, "nfb_swiglu_bwd")) return -1;
    nfb_launch_pdl(k_swiglu_bwd<float>, dim3(nfb_grid_for(rows * CV, 256)), dim3(256), 0, nfb_cur_stream(), (const float*)gate, (const float*)up, (const float*)dout, (float*)dgate, (float*)dup, (int)rows, CV, C, in_ld);
  } else if (dtype == NFB_BF16) {
    if (swiglu_check<__nv_bfloat16>(gate, up, dout, dgate, dup, C, in_ld, "nfb_swiglu_bwd")) return -1;
    int CV = (int)(C / 8);
    if (swiglu_size_check(rows, CV, "nfb_swiglu_bwd")) return -1;
    nfb_launch_pdl(k_swiglu_bwd<__nv_bfloat16>, dim3(nfb_grid_for(rows * CV, 256)), dim3(256), 0, nfb_cur_stream(), (const __nv_bfloat16*)gate, (const __nv_bfloat16*)up, (const __nv_bfloat16*)dout, (__nv_bfloat16*)dgate, (__nv_bfloat16*)dup, (int)rows, CV, C, in_ld);
  } else return NFB_FAIL("nfb_swiglu_bwd: unsupported dtype %d", dtype);
  NFB_LAUNCH_CHECK();
  return 0;
}

// y = a + b (+ c) in one pass; accumulate-in-place when out == a. fp32 residual stream.
__global__ void k_add3(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c, float* __restrict__ out, int64_t n4, int64_t n) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n4; v += stride) {
    float4 x = reinterpret_cast<const float4*>(a)[v], y = reinterpret_cast<const float4*>(b)[v];
    float4 r = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    if (c) {
      float4 z = reinterpret_cast<const float4*>(c)[v];
      r.x += z.x; r.y += z.y; r.z += z.z; r.w += z.w;
    }
    reinterpret_cast<float4*>(out)[v] = r;
  }
  for (int64_t t = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
  out[t] = a[t] + b[t] + (c ? c[t] : 0.f);
}
NFB_API int nfb_add3(const float* a, const float* b, const float* c, float* out, int64_t n) {
  if (n == 0) return 0;
  int64_t n4 = n / 4;
  k_add3<<<nfb_grid_for(n4 ? n4 : 1, 256), 256, 0, nfb_cur_stream()>>>(a, b, c, out, n4, n);
  NFB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// array (op) python-scalar, exact (no reciprocal tricks): out[i] = a[i] op s, or s op a[i] when `reverse`
// ------------------------------------------------------------------------------------------------
template <int OP, typename T, bool REV>
__global__ void k_bin_scalar_arith(const T* __restrict__ a, T s, T* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = REV ? bin_arith<OP, T>(s, a[i]) : bin_arith<OP, T>(a[i], s);
}
template <int OP, typename T, bool REV>
__global__ void k_bin_scalar_cmp(const T* __restrict__ a, T s, uint8_t* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = REV ? bin_cmp<OP, T>(s, a[i]) : bin_cmp<OP, T>(a[i], s);
}
template <int OP, bool REV>
__global__ void k_bin_scalar_f32v(const float* __restrict__ a, float s, float* __restrict__ out, int64_t n4, int64_t n) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n4; v += stride) {
    float4 x = reinterpret_cast<const float4*>(a)[v], r;
    if (!REV) { r.x = bin_arith<OP, float>(x.x, s); r.y = bin_arith<OP, float>(x.y, s); r.z = bin_arith<OP, float>(x.z, s); r.w = bin_arith<OP, float>(x.w, s); }
    else { r.x = bin_arith<OP, float>(s, x.x); r.y = bin_arith<OP, float>(s, x.y); r.z = bin_arith<OP, float>(s, x.z); r.w = bin_arith<OP, float>(s, x.w); }
    reinterpret_cast<float4*>(out)[v] = r;
  }
  for (int64_t t = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += stride)
    out[t] = REV ? bin_arith<OP, float>(s, a[t]) : bin_arith<OP, float>(a[t], s);
}

template <typename T>
static int bin_scalar_t(int op, const void* a, double sd, void* out, int64_t n, int reverse) {
  int grid = nfb_grid_for(n, 256);
  cudaStream_t st = nfb_cur_stream();
  T s = (T)sd;
#define A2(OPC) case OPC: if (reverse) k_bin_scalar_arith<OPC, T, true><<<grid, 256, 0, st>>>((const T*)a, s, (T*)out, n); else k_bin_scalar_arith<OPC, T, false><<<grid, 256, 0, st>>>((const T*)a, s, (T*)out, n); break;
#define C2(OPC) case OPC: if (reverse) k_bin_scalar_cmp<OPC, T, true><<<grid, 256, 0, st>>>((const T*)a, s, (uint8_t*)out, n); else k_bin_scalar_cmp<OPC, T, false><<<grid, 256, 0, st>>>((const T*)a, s, (uint8_t*)out, n); break;
  switch (op) {
    A2(B_ADD) A2(B_SUB) A2(B_MUL) A2(B_DIV) A2(B_POW) A2(B_MAX) A2(B_MIN)
    C2(B_EQ) C2(B_NE) C2(B_LT) C2(B_LE) C2(B_GT) C2(B_GE)
    default: return NFB_FAIL("nfb_binary_scalar: unknown op %d", op);
  }
#undef A2
#undef C2
  NFB_LAUNCH_CHECK();
  return 0;
}

NFB_API int nfb_binary_scalar(int op, int dtype, const void* a, double scalar, void* out, int64_t n, int reverse) {
  if (n == 0) return 0;
  if (dtype == NFB_F32 && op <= B_MIN && ((uintptr_t)a % 16 == 0) && ((uintptr_t)out % 16 == 0)) {
    int64_t n4 = n / 4;
    int grid = nfb_grid_for(n4 ? n4 : 1, 256);
    cudaStream_t st = nfb_cur_stream();
    float s = (float)scalar;
#define V2(OPC) case OPC: if (reverse) k_bin_scalar_f32v<OPC, true><<<grid, 256, 0, st>>>((const float*)a, s, (float*)out, n4, n); else k_bin_scalar_f32v<OPC, false><<<grid, 256, 0, st>>>((const float*)a, s, (float*)out, n4, n); break;
    switch (op) { V2(B_ADD) V2(B_SUB) V2(B_MUL) V2(B_DIV) V2(B_POW) V2(B_MAX) V2(B_MIN) }
#undef V2
    NFB_LAUNCH_CHECK();
    return 0;
  }
  switch (dtype) {
    case NFB_F32: return bin_scalar_t<float>(op, a, scalar, out, n, reverse);
    case NFB_F64: return bin_scalar_t<double>(op, a, scalar, out, n, reverse);
    case NFB_I32: return bin_scalar_t<int32_t>(op, a, scalar, out, n, reverse);
    case NFB_I64: return bin_scalar_t<int64_t>(op, a, scalar, out, n, reverse);
    case NFB_BOOL: return bin_scalar_t<uint8_t>(op, a, scalar, out, n, reverse);
  }
  return NFB_FAIL("nfb_binary_scalar: unsupported dtype %d", dtype);
}

// common.cuh -- shared internals of libhordeb200: the device-resident array
// type (the Concrete-equivalent carrier of CarriersConcrete.hs:150-155,293:
// shape + strides + offset over reference-counted storage), the caching
// allocator, error plumbing and launch helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/hordeb200.h"

#define HB_R HB_MAX_RANK

// ----------------------------------------------------------------- storage
struct hb_storage {
  void *ptr;
  size_t bytes;      // size-class size handed out by the allocator
  int pool;          // allocator pool the block returns to (0 = general; HB_POOL_VIRTUAL / HB_POOL_HOST below)
  std::atomic<int> rc;
  // lazy programs (lazy.cuh): a VIRTUAL block has no memory of its own; `producer` is the recorded node whose
  // out_idx-th output it is, and while the program runs `bound` is the real block it currently stands for
  // (ptr mirrors bound's address).
  void *producer = nullptr;
  int out_idx = 0;
  hb_storage *bound = nullptr;
  // host image of small constant arrays (index tables, the [0.0, 1.0] relu table): what the rewriter reads to
  // recognise affine window tables without a device round trip
  std::shared_ptr<std::vector<uint8_t>> host;
};
#define HB_POOL_VIRTUAL (-1)
#define HB_POOL_HOST (-2)   // trace-only mode (no device): plain host memory

struct hb_array {
  std::atomic<int> rc;
  hb_storage *st;    // may be nullptr for zero-sized arrays
  hb_dtype dt;
  int rank;
  int64_t shape[HB_R];
  int64_t strides[HB_R];  // in elements; 0 = broadcast, <0 = reversed
  int64_t offset;         // in elements from st->ptr
};

// ------------------------------------------------------------ global state
struct hb_global {
  bool inited = false;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::atomic<uint64_t> launches{0};
  std::atomic<uint64_t> live_bytes{0};
  std::atomic<uint64_t> reserved_bytes{0};
  // scratch for two-pass reductions / scatter bookkeeping (grown on demand)
  void *scratch = nullptr;
  size_t scratch_bytes = 0;
  bool capturing = false;
  bool profiling = false;  // hb_prof_begin .. hb_prof_end: one event per launch
  // DMMA convolutions leave out the taps that fall entirely into the zero padding
  // (HB_CONV_FULL_PADDING=1 computes them: only differs for non-finite operands, 0 * Inf)
  bool conv_skip_padding = true;
  // hb_init_trace_only: no device at all; constants live in host memory and every compute entry point must be
  // recorded (hb_lazy_begin) -- lets the contraction pass be exercised on a box without a GPU
  bool trace_only = false;
  cudaEvent_t events[64] = {};
};
extern hb_global g_hb;

// ------------------------------------------------------------------ errors
void hb_set_error(const char *fmt, ...);
#define HB_FAIL(code, ...)        \
  do {                            \
    hb_set_error(__VA_ARGS__);    \
    return (code);                \
  } while (0)
#define HB_REQUIRE(cond, ...)                     \
  do {                                            \
    if (!(cond)) HB_FAIL(HB_ERR_INVALID, __VA_ARGS__); \
  } while (0)
#define HB_CUDA(call)                                                     \
  do {                                                                    \
    cudaError_t e_ = (call);                                              \
    if (e_ != cudaSuccess)                                                \
      HB_FAIL(HB_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                \
              cudaGetErrorString(e_), __FILE__, __LINE__);                \
  } while (0)
#define HB_TRY(call)            \
  do {                          \
    int s_ = (call);            \
    if (s_ != HB_OK) return s_; \
  } while (0)
// Re-entrancy (GHC -threaded, up to -N8 capabilities, finalisers on arbitrary OS threads): every
// compute entry point runs under ONE process-wide recursive lock (entry points nest) and binds the library's device
// to the calling OS thread on first use -- the single stream, the reduction/scatter scratch, the profiler and the
// capture state are therefore only ever touched by one thread at a time, and a foreign thread on rank/device k > 0
// allocates and launches on device k, not on device 0.  hb_release/hb_retain and the metadata/view calls take no
// lock (atomics + the allocator's own mutex).
struct hb_entry_guard {
  hb_entry_guard();
  ~hb_entry_guard();
};
#define HB_CAT2_(a, b) a##b
#define HB_CAT_(a, b) HB_CAT2_(a, b)
#define HB_ENTRY_LOCK() hb_entry_guard HB_CAT_(hb_guard_, __LINE__)
bool hb_lazy_on();  // lazy.cu: a program is being recorded (and recording is not suspended by a replay)
// entry points that can be RECORDED (they branch on hb_lazy_on() themselves) and the constant constructors
#define HB_CHECK_INIT_LAZY()                                   \
  HB_REQUIRE(g_hb.inited, "hb_init has not been called");      \
  HB_ENTRY_LOCK()
// every other entry point: eager only -- refused while a lazy program is being recorded (its result would not be
// part of the program) and in trace-only mode (there is no device to run it on)
#define HB_CHECK_INIT()                                                                                         \
  HB_CHECK_INIT_LAZY();                                                                                         \
  HB_REQUIRE(!hb_lazy_on(), "%s: this entry point cannot be recorded into a lazy program (hb_lazy_begin is active)", \
             __func__);                                                                                         \
  HB_REQUIRE(!g_hb.trace_only, "%s: trace-only mode has no device to run on", __func__)

// per-launch profiler (array.cu): an event after every launch while enabled
void hb_prof_mark(const char *func, int line);
// Names the C-ABI entry point (with operand shapes) the launches belong to;
// the outermost scope wins.  Free when profiling is off.
struct hb_prof_scope {
  bool owner = false;
  hb_prof_scope(const char *name, const struct hb_array *a = nullptr, const struct hb_array *b = nullptr);
  ~hb_prof_scope();
};
// count + check a kernel launch
#define HB_LAUNCHED()                                                  \
  do {                                                                 \
    g_hb.launches.fetch_add(1, std::memory_order_relaxed);             \
    if (g_hb.profiling) hb_prof_mark(__func__, __LINE__);              \
    cudaError_t e_ = cudaPeekAtLastError();                            \
    if (e_ != cudaSuccess) {                                           \
      cudaGetLastError();                                              \
      if (g_hb.trace_only)                                             \
        HB_FAIL(HB_ERR_UNSUPPORTED, "%s: trace-only mode has no device to run on (record the call instead)", __func__); \
      HB_FAIL(HB_ERR_CUDA, "kernel launch failed: %s (%s:%d)",         \
              cudaGetErrorString(e_), __FILE__, __LINE__);             \
    }                                                                  \
  } while (0)

// ---------------------------------------------------------------- helpers
static inline size_t hb_dtype_size(hb_dtype dt) {
  switch (dt) {
    case HB_F64: case HB_I64: return 8;
    case HB_F32: case HB_I32: return 4;
    case HB_I16: return 2;
    default: return 1;
  }
}
// x ^ y for y >= 1 exactly as GHC's (^) computes it (GHC.Real.powImpl / powImplAcc: binary exponentiation, a
// fixed sequence of multiplications) -- updateWithGradientAdam's `betaOne ^ tAdamNew` (OptimizerTools.hs:154-155)
// is therefore reproduced bit for bit on the host, in device code and in the oracle.
#ifdef __CUDACC__
__host__ __device__
#endif
static inline double hb_powi_ghc(double x, long long y) {
  if (y <= 0) return 1.0;
  while ((y & 1) == 0) { x = x * x; y >>= 1; }
  if (y == 1) return x;
  double z = x;               // powImplAcc (x*x) (y `quot` 2) x
  x = x * x; y >>= 1;
  for (;;) {
    if ((y & 1) == 0) { x = x * x; y >>= 1; continue; }
    if (y == 1) return x * z;
    z = x * z; x = x * x; y >>= 1;
  }
}

static inline bool hb_is_float(hb_dtype dt) { return dt == HB_F64 || dt == HB_F32; }
static inline const char *hb_dtype_name(hb_dtype dt) {
  static const char *n[] = {"f64", "f32", "i64", "i32", "i16", "i8", "bool"};
  return (int)dt >= 0 && (int)dt <= 6 ? n[dt] : "?";
}
static inline int64_t hb_numel(const hb_array *a) {
  int64_t n = 1;
  for (int i = 0; i < a->rank; i++) n *= a->shape[i];
  return n;
}
static inline bool hb_contig(const hb_array *a) {
  int64_t s = 1;
  for (int i = a->rank - 1; i >= 0; i--) {
    if (a->shape[i] != 1 && a->strides[i] != s) return false;
    s *= a->shape[i];
  }
  return true;
}
static inline void *hb_data(const hb_array *a) {
  return a->st ? (char *)a->st->ptr + a->offset * (int64_t)hb_dtype_size(a->dt) : nullptr;
}
static inline bool hb_same_shape(const hb_array *a, const hb_array *b) {
  if (a->rank != b->rank) return false;
  for (int i = 0; i < a->rank; i++)
    if (a->shape[i] != b->shape[i]) return false;
  return true;
}
std::string hb_shape_str(const hb_array *a);

// allocator (alloc.cu)
int hb_alloc_storage(size_t bytes, hb_storage **out);
void hb_storage_release(hb_storage *st);
int hb_scratch(size_t bytes, void **ptr);  // persistent scratch, >= bytes

// array construction (array.cu)
int hb_new_array(hb_dtype dt, int rank, const int64_t *shape, hb_array **out);  // fresh contiguous
hb_array *hb_new_view(const hb_array *parent);  // same storage, copies metadata
int hb_zeros(hb_dtype dt, int rank, const int64_t *shape, hb_array **out);
// write one scalar of dtype dt (read from host memory `value`) to device `dst`
int hb_fill_scalar(void *dst, hb_dtype dt, const void *value);

// ------------------------------------------------- strided iteration space
// Collapsed description of up to 3 operands over a common logical shape.
#define HB_MAXOPS 4
struct hb_iter {
  int rank;
  int64_t shape[HB_R];
  int64_t stride[HB_MAXOPS][HB_R];
};
// Build a collapsed iteration space for `nops` arrays of identical logical
// shape: size-1 axes dropped, adjacent axes merged when every operand allows.
void hb_make_iter(int nops, const hb_array *const *ops, hb_iter *it);

// Generic: collapse arbitrary (shape, strides[nops]) lists in place.
void hb_collapse(int *rank, int64_t *shape, int nops, int64_t (*strides)[HB_R]);

static inline int hb_blocks_for(int64_t n, int per_block, int max_blocks = 1 << 30) {
  int64_t b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

// dtype dispatch ---------------------------------------------------------
#define HB_DISPATCH_ALL(dt, T, ...)                                  \
  switch (dt) {                                                      \
    case HB_F64: { typedef double T; __VA_ARGS__; } break;           \
    case HB_F32: { typedef float T; __VA_ARGS__; } break;            \
    case HB_I64: { typedef int64_t T; __VA_ARGS__; } break;          \
    case HB_I32: { typedef int32_t T; __VA_ARGS__; } break;          \
    case HB_I16: { typedef int16_t T; __VA_ARGS__; } break;          \
    case HB_I8: { typedef int8_t T; __VA_ARGS__; } break;            \
    case HB_BOOL: { typedef uint8_t T; __VA_ARGS__; } break;         \
    default: HB_FAIL(HB_ERR_INVALID, "bad dtype %d", (int)(dt));     \
  }
#define HB_DISPATCH_FLOAT(dt, T, ...)                                \
  switch (dt) {                                                      \
    case HB_F64: { typedef double T; __VA_ARGS__; } break;           \
    case HB_F32: { typedef float T; __VA_ARGS__; } break;            \
    default: HB_FAIL(HB_ERR_INVALID, "floating dtype required, got %s", hb_dtype_name(dt)); \
  }
#define HB_DISPATCH_NUM(dt, T, ...)                                  \
  switch (dt) {                                                      \
    case HB_F64: { typedef double T; __VA_ARGS__; } break;           \
    case HB_F32: { typedef float T; __VA_ARGS__; } break;            \
    case HB_I64: { typedef int64_t T; __VA_ARGS__; } break;          \
    case HB_I32: { typedef int32_t T; __VA_ARGS__; } break;          \
    case HB_I16: { typedef int16_t T; __VA_ARGS__; } break;          \
    case HB_I8: { typedef int8_t T; __VA_ARGS__; } break;            \
    default: HB_FAIL(HB_ERR_INVALID, "numeric dtype required, got %s", hb_dtype_name(dt)); \
  }
// byte-size dispatch for pure data movement (bit-exact for every dtype)
#define HB_DISPATCH_BYTES(dt, T, ...)                                \
  switch (hb_dtype_size(dt)) {                                       \
    case 8: { typedef uint64_t T; __VA_ARGS__; } break;              \
    case 4: { typedef uint32_t T; __VA_ARGS__; } break;              \
    case 2: { typedef uint16_t T; __VA_ARGS__; } break;              \
    default: { typedef uint8_t T; __VA_ARGS__; } break;              \
  }

// ---------------------------------------------------- device-side helpers
#ifdef __CUDACC__
struct hb_dims {  // by-value kernel parameter
  int rank;
  int64_t shape[HB_R];
  int64_t stride[HB_R];
};
// linear index (row-major over d.shape) -> element offset with d.stride
__device__ __forceinline__ int64_t hb_offset_of(const hb_dims &d, int64_t lin) {
  int64_t off = 0;
#pragma unroll 1
  for (int i = d.rank - 1; i >= 0; i--) {
    int64_t q = lin / d.shape[i];
    int64_t r = lin - q * d.shape[i];
    off += r * d.stride[i];
    lin = q;
  }
  return off;
}
// xor-shuffle for any trivially copyable type of 1, 2, 4 or 8 bytes
template <typename T>
__device__ __forceinline__ T hb_shfl_xor(T v, int o) {
  if constexpr (sizeof(T) == 8) {
    unsigned long long u;
    memcpy(&u, &v, 8);
    unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)u, o);
    unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)(u >> 32), o);
    u = ((unsigned long long)hi << 32) | lo;
    T r;
    memcpy(&r, &u, 8);
    return r;
  } else if constexpr (sizeof(T) == 4) {
    unsigned u;
    memcpy(&u, &v, 4);
    u = __shfl_xor_sync(0xffffffffu, u, o);
    T r;
    memcpy(&r, &u, 4);
    return r;
  } else {
    int u = (int)v;
    u = __shfl_xor_sync(0xffffffffu, u, o);
    return (T)u;
  }
}
template <typename T>
__device__ __forceinline__ T hb_warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += hb_shfl_xor(v, o);
  return v;
}
#endif

#ifdef __CUDACC__
// tanh for the recurrent layer body (hb_tanh_affine_fwd and the epilogue of the grouped products): branch-free, so
// that the sixteen results of a thread's accumulator tile interleave instead of running libdevice's two-path tanh
// back to back (measured: 23 us per 2 x [512, 1024] epilogue once the pre-activations leave |x| < 0.55).
//   tanh |x| = E / (E + 2),  E = expm1 (2 |x|) = 2^k expm1 (r) + (2^k - 1),  2 |x| = k ln 2 + r, |r| <= ln 2 / 2,
// expm1 (r) by its Taylor polynomial to r^14 (truncation 1.2e-17 relative), the quotient by rcp.approx + two Newton
// steps + one residual correction.  Within 3 ulp of the correctly rounded value over the whole range (tested
// against the oracle); exact for |x| -> 0 (E -> 2 |x|), 1 beyond |x| = 20, NaN for NaN, sign of zero kept.
__device__ __forceinline__ double hb_tanh_f64(double x) {
  double a = fabs(x);
  a = a > 20.0 ? 20.0 : a;                       // NaN stays NaN
  const double y = a + a;
  const double kf = rint(y * 1.4426950408889634);
  double r = fma(-kf, 6.93147180369123816490e-01, y);
  r = fma(-kf, 1.90821492927058770002e-10, r);
  double q = 1.1470745597729725e-11;             // 1/14!
  q = fma(q, r, 1.6059043836821613e-10);         // 1/13!
  q = fma(q, r, 2.08767569878681e-09);           // 1/12!
  q = fma(q, r, 2.505210838544172e-08);          // 1/11!
  q = fma(q, r, 2.755731922398589e-07);          // 1/10!
  q = fma(q, r, 2.7557319223985893e-06);         // 1/9!
  q = fma(q, r, 2.48015873015873e-05);           // 1/8!
  q = fma(q, r, 1.984126984126984e-04);          // 1/7!
  q = fma(q, r, 1.388888888888889e-03);          // 1/6!
  q = fma(q, r, 8.333333333333333e-03);          // 1/5!
  q = fma(q, r, 4.1666666666666664e-02);         // 1/4!
  q = fma(q, r, 1.6666666666666666e-01);         // 1/3!
  q = fma(q, r, 0.5);
  const double em1r = fma(r * r, q, r);          // expm1 (r)
  const double two_k = __longlong_as_double(((long long)kf + 1023ll) << 52);
  const double E = fma(two_k, em1r, two_k - 1.0);
  const double d = E + 2.0;
  double rc;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(d));
  rc = fma(fma(-d, rc, 1.0), rc, rc);
  rc = fma(fma(-d, rc, 1.0), rc, rc);
  double t = E * rc;
  t = fma(fma(-d, t, E), rc, t);
  return copysign(t, x);
}
#endif

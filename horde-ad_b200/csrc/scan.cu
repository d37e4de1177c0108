// scan.cu -- tmapAccumL specialisation for `sfold (*) 1 x` and its reverse
// derivative (LongProdBench: bench/common/BenchProdTools.hs:154-182;
// Ops.hs:188-210; OpsConcrete.hs:990-1019; the transposition rule is
// DeltaMapAccumL, DeltaEval.hs:351-389).
//
// The reference runs n sequential host iterations; here the fold is recognised
// as an associative scan:
//   acc_i = prod_{j<i} x_j        (exclusive prefix product, = the fold states)
//   cot_i = dt * prod_{j>i} x_j   (exclusive suffix product, = the cotangent
//                                  flowing backwards through the steps)
//   grad_i = cot_i * acc_i        (no division: exact also when some x_j = 0)
// Three kernels: tile products -> scan of tile products (one block) -> tile
// rescans producing the gradient.  HBM traffic 24 bytes per element (two
// reads, one write) against the 16-byte minimum.
#include "common.cuh"

#define PT 2048  // elements per tile (256 threads x 8)

template <typename T>
__global__ void __launch_bounds__(256)
    hb_prod_tiles_kernel(const T *__restrict__ x, int64_t stride, int64_t n, T *__restrict__ tileprod) {
  __shared__ T smem[8];
  int64_t base = (int64_t)blockIdx.x * PT + threadIdx.x * 8;
  T p = T(1);
#pragma unroll
  for (int i = 0; i < 8; i++)
    if (base + i < n) p *= x[(base + i) * stride];
  for (int o = 1; o < 32; o <<= 1) {  // ordered warp product (left to right)
    T u = hb_shfl_xor(p, o);
    p = ((threadIdx.x & o) ? u * p : p * u);
  }
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) smem[w] = p;
  __syncthreads();
  if (threadIdx.x == 0) {
    T r = smem[0];
    for (int i = 1; i < 8; i++) r *= smem[i];
    tileprod[blockIdx.x] = r;
  }
}

// exclusive prefix (pre) and suffix (suf) products of the tile products;
// one block, each thread owns a contiguous chunk.
template <typename T>
__global__ void __launch_bounds__(1024)
    hb_prod_tilescan_kernel(const T *__restrict__ tp, int64_t nt, T dt, T *__restrict__ pre,
                            T *__restrict__ suf, T *__restrict__ total) {
  __shared__ T chunk[1024];
  __shared__ T chunk_pre[1024];
  __shared__ T chunk_suf[1024];
  int64_t per = (nt + blockDim.x - 1) / blockDim.x;
  int64_t lo = (int64_t)threadIdx.x * per, hi = lo + per < nt ? lo + per : nt;
  T p = T(1);
  for (int64_t i = lo; i < hi; i++) p *= tp[i];
  chunk[threadIdx.x] = p;
  __syncthreads();
  if (threadIdx.x == 0) {
    T r = T(1);
    for (int i = 0; i < (int)blockDim.x; i++) { chunk_pre[i] = r; r *= chunk[i]; }
    *total = r;
    r = dt;
    for (int i = blockDim.x - 1; i >= 0; i--) { chunk_suf[i] = r; r *= chunk[i]; }
  }
  __syncthreads();
  T r = chunk_pre[threadIdx.x];
  for (int64_t i = lo; i < hi; i++) { pre[i] = r; r *= tp[i]; }
  r = chunk_suf[threadIdx.x];
  for (int64_t i = hi - 1; i >= lo; i--) { suf[i] = r; r *= tp[i]; }
}

// per tile: grad[i] = (suf_tile * prod_{j>i in tile} x_j) * (pre_tile * prod_{j<i in tile} x_j)
template <typename T>
__global__ void __launch_bounds__(256)
    hb_prod_grad_kernel(const T *__restrict__ x, int64_t stride, int64_t n, const T *__restrict__ pre,
                        const T *__restrict__ suf, T *__restrict__ grad) {
  __shared__ T tpre[257], tsuf[257];
  int64_t base = (int64_t)blockIdx.x * PT + threadIdx.x * 8;
  T v[8];
  T p = T(1);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    v[i] = base + i < n ? x[(base + i) * stride] : T(1);
    p *= v[i];
  }
  tpre[threadIdx.x + 1] = p;  // thread products; scanned serially by two threads
  __syncthreads();
  if (threadIdx.x == 0) {
    T r = pre[blockIdx.x];
    T prev = r;
    for (int i = 0; i < 256; i++) { T t = tpre[i + 1]; tpre[i] = prev; prev = prev * t; tsuf[i] = t; }
    // tsuf currently holds thread products; convert to exclusive suffix
    T s = suf[blockIdx.x];
    for (int i = 255; i >= 0; i--) { T t = tsuf[i]; tsuf[i] = s; s = s * t; }
  }
  __syncthreads();
  T a = tpre[threadIdx.x];   // product of everything before this thread's 8
  T c = tsuf[threadIdx.x];   // dt * product of everything after this thread's 8
  T sufl[8];
  T r = c;
#pragma unroll
  for (int i = 7; i >= 0; i--) { sufl[i] = r; r *= v[i]; }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if (base + i < n) grad[base + i] = sufl[i] * a;
    a *= v[i];
  }
}

extern "C" int hb_prod_fold_grad(const hb_array *x, double dt, hb_array **prod_out, hb_array **grad_out) {
  hb_prof_scope prof_("prod_fold_grad", x);
  HB_CHECK_INIT();
  HB_REQUIRE(x && prod_out, "null argument");
  HB_REQUIRE(x->rank == 1 && hb_is_float(x->dt), "prod fold: rank-1 floating vector required");
  int64_t n = x->shape[0];
  int64_t one = 1;
  hb_array *prod, *grad = nullptr;
  HB_TRY(hb_new_array(x->dt, 0, &one, &prod));
  if (grad_out) {
    int st = hb_new_array(x->dt, 1, &n, &grad);
    if (st != HB_OK) { hb_release(prod); return st; }
  }
  if (n == 0) {
    double o64 = 1.0; float o32 = 1.f;
    int st = hb_fill_scalar(hb_data(prod), x->dt, x->dt == HB_F64 ? (void *)&o64 : (void *)&o32);
    if (st != HB_OK) { hb_release(prod); hb_release(grad); return st; }
    *prod_out = prod;
    if (grad_out) *grad_out = grad;
    return HB_OK;
  }
  int64_t nt = (n + PT - 1) / PT;
  size_t es = hb_dtype_size(x->dt);
  void *scr;
  int st = hb_scratch((size_t)nt * 3 * es + 64, &scr);
  if (st != HB_OK) { hb_release(prod); hb_release(grad); return st; }
  if (x->dt == HB_F64) {
    double *tp = (double *)scr, *pre = tp + nt, *suf = pre + nt;
    hb_prod_tiles_kernel<double><<<(unsigned)nt, 256, 0, g_hb.stream>>>((const double *)hb_data(x), x->strides[0], n, tp);
    HB_LAUNCHED();
    hb_prod_tilescan_kernel<double><<<1, 1024, 0, g_hb.stream>>>(tp, nt, dt, pre, suf, (double *)hb_data(prod));
    HB_LAUNCHED();
    if (grad) {
      hb_prod_grad_kernel<double><<<(unsigned)nt, 256, 0, g_hb.stream>>>((const double *)hb_data(x), x->strides[0], n, pre, suf, (double *)hb_data(grad));
      HB_LAUNCHED();
    }
  } else {
    float *tp = (float *)scr, *pre = tp + nt, *suf = pre + nt;
    hb_prod_tiles_kernel<float><<<(unsigned)nt, 256, 0, g_hb.stream>>>((const float *)hb_data(x), x->strides[0], n, tp);
    HB_LAUNCHED();
    hb_prod_tilescan_kernel<float><<<1, 1024, 0, g_hb.stream>>>(tp, nt, (float)dt, pre, suf, (float *)hb_data(prod));
    HB_LAUNCHED();
    if (grad) {
      hb_prod_grad_kernel<float><<<(unsigned)nt, 256, 0, g_hb.stream>>>((const float *)hb_data(x), x->strides[0], n, pre, suf, (float *)hb_data(grad));
      HB_LAUNCHED();
    }
  }
  *prod_out = prod;
  if (grad_out) *grad_out = grad;
  return HB_OK;
}

// reduce.cu -- ssum/ssumN/ssum0, sdot0, generic sdot1In and argmin/argmax.
//
// Reference: OpsConcrete.hs:215-235 (tssum sums the OUTERMOST axis, tssumN
// repeats that for each leading axis, tsdot1In contracts the INNERMOST axis
// of two equal-shape -- possibly stride-0 broadcast -- arrays) and
// :1712-1746 (argmin/argmax over the LAST axis, result Init sh of Int).
//
// One two-operand strided reduction engine serves all of them: an operand
// is described by "kept" axes (the output index space) and "reduced" axes,
// each with its own strides, so transposed/replicated views are consumed in
// place.  Two launch shapes, both deterministic (fixed partition + fixed
// tree, two-pass through a partials buffer when the reduced extent is split):
//   rows : the unit stride lies on a reduced axis (or there is one output):
//          a block per (output, split); threads stride the reduced index;
//          warp-shuffle + shared-memory tree.
//   cols : the unit stride lies on a kept axis: threads own outputs
//          (coalesced across the kept index), blockDim.y interleaves the
//          reduced range, shared-memory tree over y.
// HBM roofline: sizeof(T) bytes per input element (2x for dot), SURVEY 8d.
#include "common.cuh"
#include "lazy.cuh"
#include "iter.cuh"

struct hb_redp {
  int krank, rrank;
  int64_t K, R;
  int64_t kshape[HB_R], ksa[HB_R], ksb[HB_R];
  int64_t rshape[HB_R], rsa[HB_R], rsb[HB_R];
};

template <typename T> struct hb_acc { typedef T type; };
template <> struct hb_acc<float> { typedef double type; };

__device__ __forceinline__ void hb_decomp2(int rank, const int64_t *shape, const int64_t *sa,
                                           const int64_t *sb, int64_t lin, int64_t &oa, int64_t &ob) {
  oa = 0;
  ob = 0;
#pragma unroll 1
  for (int d = rank - 1; d >= 0; d--) {
    int64_t q = lin / shape[d], r = lin - q * shape[d];
    oa += r * sa[d];
    ob += r * sb[d];
    lin = q;
  }
}

template <typename ACC>
__device__ __forceinline__ ACC hb_block_sum(ACC v, ACC *smem) {
  // fixed-shape tree: xor-shuffle within warps, then warp 0 over warp sums
  for (int o = 16; o > 0; o >>= 1) v += hb_shfl_xor(v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  int nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) smem[w] = v;
  __syncthreads();
  if (w == 0) {
    v = l < nw ? smem[l] : ACC(0);
    for (int o = 16; o > 0; o >>= 1) v += hb_shfl_xor(v, o);
  }
  return v;  // valid in thread 0
}


template <typename T, typename ACC, bool DOT>
__global__ void __launch_bounds__(256)
    hb_reduce_rows_kernel(hb_redp p, const T *__restrict__ a, const T *__restrict__ b,
                          ACC *__restrict__ out, int64_t rchunk) {
  __shared__ ACC smem[32];
  int64_t k = blockIdx.x;
  int64_t s = blockIdx.y;
  int64_t ka, kb;
  hb_decomp2(p.krank, p.kshape, p.ksa, p.ksb, k, ka, kb);
  int64_t r0 = s * rchunk;
  int64_t r1 = r0 + rchunk < p.R ? r0 + rchunk : p.R;
  ACC acc = ACC(0);
  if (p.rrank == 1) {
    const int64_t sa = p.rsa[0], sb = p.rsb[0];
    for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
      ACC x = (ACC)a[ka + r * sa];
      if (DOT) x *= (ACC)b[kb + r * sb];
      acc += x;
    }
  } else if (p.rrank == 2) {
    const int64_t n1 = p.rshape[1];
    for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
      int64_t q = r / n1, m = r - q * n1;
      ACC x = (ACC)a[ka + q * p.rsa[0] + m * p.rsa[1]];
      if (DOT) x *= (ACC)b[kb + q * p.rsb[0] + m * p.rsb[1]];
      acc += x;
    }
  } else {
    for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
      int64_t oa, ob;
      hb_decomp2(p.rrank, p.rshape, p.rsa, p.rsb, r, oa, ob);
      ACC x = (ACC)a[ka + oa];
      if (DOT) x *= (ACC)b[kb + ob];
      acc += x;
    }
  }
  acc = hb_block_sum<ACC>(acc, smem);
  if (threadIdx.x == 0) out[s * p.K + k] = acc;
}

// rows kernel for the common layout "the innermost reduced axis is a unit-
// stride run" (ssum0 / sdot0 of contiguous data; bias gradients = ssumN over
// [N][H*W] runs of an NCHW tensor): no per-element index arithmetic, 128-bit
// loads, four independent accumulators per thread (combined in a fixed order).
template <typename T, typename ACC, bool DOT, bool VEC>
__global__ void __launch_bounds__(256)
    hb_reduce_runs_kernel(hb_redp p, const T *__restrict__ a, const T *__restrict__ b, ACC *__restrict__ out,
                          int64_t rchunk) {
  __shared__ ACC smem[32];
  constexpr int V = VEC ? 16 / (int)sizeof(T) : 1;
  typedef hb_vec<T, V> VT;
  const int64_t k = blockIdx.x, s = blockIdx.y;
  int64_t ka, kb;
  hb_decomp2(p.krank, p.kshape, p.ksa, p.ksb, k, ka, kb);
  const int64_t n1 = p.rshape[p.rrank - 1];
  const int64_t s0a = p.rrank == 2 ? p.rsa[0] : 0, s0b = p.rrank == 2 ? p.rsb[0] : 0;
  const int64_t r0 = s * rchunk, r1 = r0 + rchunk < p.R ? r0 + rchunk : p.R;
  ACC acc[4] = {ACC(0), ACC(0), ACC(0), ACC(0)};
  int64_t q = r0 / n1, m0 = r0 - q * n1;
  for (int64_t r = r0; r < r1;) {
    int64_t len = n1 - m0 < r1 - r ? n1 - m0 : r1 - r;
    const VT *pa = reinterpret_cast<const VT *>(a + ka + q * s0a + m0);
    const VT *pb = reinterpret_cast<const VT *>(b + kb + q * s0b + m0);
    const int64_t nv = len / V;
    int64_t i = threadIdx.x;
    for (; i + 3 * 256 < nv; i += 4 * 256) {
      VT x0 = pa[i], x1 = pa[i + 256], x2 = pa[i + 512], x3 = pa[i + 768];
      VT y0, y1, y2, y3;
      if (DOT) { y0 = pb[i]; y1 = pb[i + 256]; y2 = pb[i + 512]; y3 = pb[i + 768]; }
#pragma unroll
      for (int e = 0; e < V; e++) {
        acc[0] += DOT ? (ACC)x0.v[e] * (ACC)y0.v[e] : (ACC)x0.v[e];
        acc[1] += DOT ? (ACC)x1.v[e] * (ACC)y1.v[e] : (ACC)x1.v[e];
        acc[2] += DOT ? (ACC)x2.v[e] * (ACC)y2.v[e] : (ACC)x2.v[e];
        acc[3] += DOT ? (ACC)x3.v[e] * (ACC)y3.v[e] : (ACC)x3.v[e];
      }
    }
    for (; i < nv; i += 256) {
      VT x0 = pa[i], y0;
      if (DOT) y0 = pb[i];
#pragma unroll
      for (int e = 0; e < V; e++) acc[0] += DOT ? (ACC)x0.v[e] * (ACC)y0.v[e] : (ACC)x0.v[e];
    }
    r += len;
    q++;
    m0 = 0;
  }
  ACC t = hb_block_sum<ACC>((acc[0] + acc[1]) + (acc[2] + acc[3]), smem);
  if (threadIdx.x == 0) out[s * p.K + k] = t;
}

// blockDim = (BX, BY); x -> kept index, y interleaves the reduced range
template <typename T, typename ACC, bool DOT>
__global__ void __launch_bounds__(256)
    hb_reduce_cols_kernel(hb_redp p, const T *__restrict__ a, const T *__restrict__ b,
                          ACC *__restrict__ out, int64_t rchunk) {
  __shared__ ACC smem[256];
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t s = blockIdx.y;
  ACC acc = ACC(0);
  if (k < p.K) {
    int64_t ka, kb;
    hb_decomp2(p.krank, p.kshape, p.ksa, p.ksb, k, ka, kb);
    int64_t r0 = s * rchunk;
    int64_t r1 = r0 + rchunk < p.R ? r0 + rchunk : p.R;
    if (p.rrank == 1) {
      const int64_t sa = p.rsa[0], sb = p.rsb[0];
      for (int64_t r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
        ACC x = (ACC)a[ka + r * sa];
        if (DOT) x *= (ACC)b[kb + r * sb];
        acc += x;
      }
    } else {
      for (int64_t r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
        int64_t oa, ob;
        hb_decomp2(p.rrank, p.rshape, p.rsa, p.rsb, r, oa, ob);
        ACC x = (ACC)a[ka + oa];
        if (DOT) x *= (ACC)b[kb + ob];
        acc += x;
      }
    }
  }
  int t = threadIdx.y * blockDim.x + threadIdx.x;
  smem[t] = acc;
  __syncthreads();
  for (int h = blockDim.y >> 1; h > 0; h >>= 1) {  // fixed tree over y
    if (threadIdx.y < h) smem[t] += smem[t + h * blockDim.x];
    __syncthreads();
  }
  if (threadIdx.y == 0 && k < p.K) out[s * p.K + k] = smem[threadIdx.x];
}

template <typename ACC, typename T>
__global__ void hb_finish_partials_kernel(const ACC *__restrict__ part, T *__restrict__ out, int64_t K,
                                          int S) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  ACC acc = ACC(0);
  for (int s = 0; s < S; s++) acc += part[(int64_t)s * K + k];
  out[k] = (T)acc;
}

// Few outputs, many partials (ssum0 / sdot0: K = 1, S ~ 1200): one block per
// output, threads take the partials interleaved, fixed-shape tree at the end
// (the one-thread loop above is a chain of S dependent adds behind L2 loads:
// 40 us of a 130 us ssum0 at n = 2^26).
template <typename ACC, typename T>
__global__ void __launch_bounds__(256)
    hb_finish_partials_block_kernel(const ACC *__restrict__ part, T *__restrict__ out, int64_t K, int S) {
  __shared__ ACC smem[32];
  const int64_t k = blockIdx.x;
  ACC acc = ACC(0);
  for (int s = threadIdx.x; s < S; s += 256) acc += part[(int64_t)s * K + k];
  acc = hb_block_sum<ACC>(acc, smem);
  if (threadIdx.x == 0) out[k] = (T)acc;
}

// kept axes / reduced axes given as index lists into the operands' axes.
template <typename T, bool DOT>
static int reduce_typed(const hb_array *a, const hb_array *b, int nk, const int *kax, int nr,
                        const int *rax, hb_array *out) {
  typedef typename hb_acc<T>::type ACC;
  hb_redp p;
  memset(&p, 0, sizeof p);
  int64_t ks[2][HB_R], rs[2][HB_R];
  p.krank = nk;
  for (int i = 0; i < nk; i++) {
    p.kshape[i] = a->shape[kax[i]];
    ks[0][i] = a->strides[kax[i]];
    ks[1][i] = b ? b->strides[kax[i]] : 0;
  }
  p.rrank = nr;
  for (int i = 0; i < nr; i++) {
    p.rshape[i] = a->shape[rax[i]];
    rs[0][i] = a->strides[rax[i]];
    rs[1][i] = b ? b->strides[rax[i]] : 0;
  }
  hb_collapse(&p.krank, p.kshape, 2, ks);
  hb_collapse(&p.rrank, p.rshape, 2, rs);
  p.K = 1;
  for (int i = 0; i < p.krank; i++) { p.K *= p.kshape[i]; p.ksa[i] = ks[0][i]; p.ksb[i] = ks[1][i]; }
  p.R = 1;
  for (int i = 0; i < p.rrank; i++) { p.R *= p.rshape[i]; p.rsa[i] = rs[0][i]; p.rsb[i] = rs[1][i]; }
  if (p.rrank == 0) { p.rrank = 1; p.rshape[0] = 1; p.rsa[0] = 0; p.rsb[0] = 0; }
  if (p.K == 0) return HB_OK;
  T *po = (T *)hb_data(out);
  if (p.R == 0) {
    HB_CUDA(cudaMemsetAsync(po, 0, (size_t)p.K * sizeof(T), g_hb.stream));
    return HB_OK;
  }
  // which axis group carries the smallest non-zero stride?
  auto minstride = [](int rank, const int64_t *s1, const int64_t *s2) {
    int64_t m = INT64_MAX;
    for (int i = 0; i < rank; i++) {
      int64_t x = s1[i] < 0 ? -s1[i] : s1[i];
      int64_t y = s2[i] < 0 ? -s2[i] : s2[i];
      if (x && x < m) m = x;
      if (y && y < m) m = y;
    }
    return m;
  };
  int64_t mk = p.krank ? minstride(p.krank, p.ksa, p.ksb) : INT64_MAX;
  int64_t mr = minstride(p.rrank, p.rsa, p.rsb);
  bool rows = (p.K == 1) || (mr <= mk && p.R >= 64) || (p.K < 8 && p.R >= 1024);
  const T *pa = (const T *)hb_data(a);
  const T *pb = b ? (const T *)hb_data(b) : pa;
  int target_blocks = g_hb.sm_count * 8;
  int S;
  int64_t rchunk;
  dim3 grid, block;
  if (rows) {
    S = 1;
    if (p.K < target_blocks) {
      int64_t maxs = (p.R + 2047) / 2048;
      S = (int)((target_blocks + p.K - 1) / p.K);
      if (S > maxs) S = (int)maxs;
      if (S > 65535) S = 65535;
      if (S < 1) S = 1;
    }
    rchunk = (p.R + S - 1) / S;
    rchunk = (rchunk + 15) / 16 * 16;  // keeps the 128-bit path of the runs kernel aligned
    S = (int)((p.R + rchunk - 1) / rchunk);
    HB_REQUIRE(p.K < (1ll << 31), "reduce: too many outputs for the rows kernel");
    grid = dim3((unsigned)p.K, S);
    int bx = p.R / S >= 256 ? 256 : (p.R / S >= 128 ? 128 : (p.R / S >= 64 ? 64 : 32));
    block = dim3(bx);
  } else {
    int bx = 1;
    while (bx < 128 && bx < p.K) bx <<= 1;
    int by = 256 / bx;
    while (by > 1 && by > p.R) by >>= 1;
    int64_t kblocks = (p.K + bx - 1) / bx;
    S = 1;
    if (kblocks < target_blocks) {
      int64_t maxs = (p.R + (64 * by) - 1) / (64 * by);
      S = (int)((target_blocks + kblocks - 1) / kblocks);
      if (S > maxs) S = (int)maxs;
      if (S > 65535) S = 65535;
      if (S < 1) S = 1;
    }
    rchunk = (p.R + S - 1) / S;
    S = (int)((p.R + rchunk - 1) / rchunk);
    grid = dim3((unsigned)kblocks, S);
    block = dim3(bx, by);
  }
  ACC *dst;
  bool direct = (S == 1) && std::is_same<ACC, T>::value;
  if (direct) {
    dst = (ACC *)po;
  } else {
    void *scr;
    HB_TRY(hb_scratch((size_t)S * p.K * sizeof(ACC), &scr));
    dst = (ACC *)scr;
  }
  // unit-stride inner run on every operand that moves, 256-thread blocks
  bool runs = rows && block.x == 256 && p.rrank <= 2 && p.rsa[p.rrank - 1] == 1 && (!DOT || p.rsb[p.rrank - 1] == 1);
  if (runs) {
    constexpr int V = 16 / (int)sizeof(T);
    const int64_t n1 = p.rshape[p.rrank - 1];
    auto al = [&](const T *ptr, int64_t s0, const int64_t *ks) {
      bool ok = ((uintptr_t)ptr % 16) == 0 && s0 % V == 0;
      for (int i = 0; i < p.krank; i++) ok = ok && ks[i] % V == 0;
      return ok;
    };
    bool vec = n1 % V == 0 && rchunk % V == 0 && al(pa, p.rrank == 2 ? p.rsa[0] : 0, p.ksa) &&
               (!DOT || al(pb, p.rrank == 2 ? p.rsb[0] : 0, p.ksb));
    if (vec) hb_reduce_runs_kernel<T, ACC, DOT, true><<<grid, block, 0, g_hb.stream>>>(p, pa, pb, dst, rchunk);
    else hb_reduce_runs_kernel<T, ACC, DOT, false><<<grid, block, 0, g_hb.stream>>>(p, pa, pb, dst, rchunk);
  } else if (rows)
    hb_reduce_rows_kernel<T, ACC, DOT><<<grid, block, 0, g_hb.stream>>>(p, pa, pb, dst, rchunk);
  else
    hb_reduce_cols_kernel<T, ACC, DOT><<<grid, block, 0, g_hb.stream>>>(p, pa, pb, dst, rchunk);
  HB_LAUNCHED();
  if (!direct) {
    if (S >= 64 && p.K <= 4096)
      hb_finish_partials_block_kernel<ACC, T><<<(unsigned)p.K, 256, 0, g_hb.stream>>>(dst, po, p.K, S);
    else
      hb_finish_partials_kernel<ACC, T><<<hb_blocks_for(p.K, 256), 256, 0, g_hb.stream>>>(dst, po, p.K, S);
    HB_LAUNCHED();
  }
  return HB_OK;
}

int hb_reduce_generic(const hb_array *a, const hb_array *b, int nk, const int *kax, int nr,
                      const int *rax, hb_array *out) {
  if (b) {
    HB_DISPATCH_NUM(a->dt, T, return (reduce_typed<T, true>(a, b, nk, kax, nr, rax, out)));
  } else {
    HB_DISPATCH_NUM(a->dt, T, return (reduce_typed<T, false>(a, nullptr, nk, kax, nr, rax, out)));
  }
  return HB_OK;
}

extern "C" int hb_sum_leading(const hb_array *a, int n_axes, hb_array **out) {
  hb_prof_scope prof_("sum_leading", a);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(a && out, "null argument");
  HB_REQUIRE(n_axes >= 0 && n_axes <= a->rank, "ssumN: %d axes of rank %d", n_axes, a->rank);
  HB_REQUIRE(a->dt != HB_BOOL, "ssum: numeric dtype required");
  if (hb_lazy_on())
    return hb_lrec(HB_L_SUM_LEADING, "sum_leading").in(a).iarg(0, n_axes).out(out, a->dt, a->rank - n_axes, a->shape + n_axes).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_sum_leading(in[0], (int)n.ia[0], &o[0]); });
  hb_array *o;
  HB_TRY(hb_new_array(a->dt, a->rank - n_axes, a->shape + n_axes, &o));
  int kax[HB_R], rax[HB_R];
  for (int i = 0; i < n_axes; i++) rax[i] = i;
  for (int i = n_axes; i < a->rank; i++) kax[i - n_axes] = i;
  int st = hb_reduce_generic(a, nullptr, a->rank - n_axes, kax, n_axes, rax, o);
  if (st != HB_OK) {
    hb_release(o);
    return st;
  }
  *out = o;
  return HB_OK;
}

extern "C" int hb_sum_all(const hb_array *a, hb_array **out) {
  hb_prof_scope prof_("sum_all", a);
  HB_REQUIRE(a, "null argument");
  return hb_sum_leading(a, a->rank, out);
}

extern "C" int hb_dot_all(const hb_array *a, const hb_array *b, hb_array **out) {
  hb_prof_scope prof_("dot_all", a, b);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(a && b && out, "null argument");
  HB_REQUIRE(a->dt == b->dt && hb_same_shape(a, b), "sdot0: %s %s vs %s %s", hb_dtype_name(a->dt),
             hb_shape_str(a).c_str(), hb_dtype_name(b->dt), hb_shape_str(b).c_str());
  HB_REQUIRE(a->dt != HB_BOOL, "sdot0: numeric dtype required");
  if (hb_lazy_on()) {
    int64_t one_ = 1;
    return hb_lrec(HB_L_GENERIC, "dot_all").in(a).in(b).out(out, a->dt, 0, &one_).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_dot_all(in[0], in[1], &o[0]); });
  }
  hb_array *o;
  int64_t one = 1;
  HB_TRY(hb_new_array(a->dt, 0, &one, &o));
  int rax[HB_R];
  for (int i = 0; i < a->rank; i++) rax[i] = i;
  int st = hb_reduce_generic(a, b, 0, nullptr, a->rank, rax, o);
  if (st != HB_OK) {
    hb_release(o);
    return st;
  }
  *out = o;
  return HB_OK;
}

// generic (non-GEMM) sdot1In; contract.cu routes GEMM-shaped calls elsewhere
int hb_dot_inner_generic(const hb_array *a, const hb_array *b, hb_array *o) {
  int kax[HB_R], rax[1];
  for (int i = 0; i < a->rank - 1; i++) kax[i] = i;
  rax[0] = a->rank - 1;
  return hb_reduce_generic(a, b, a->rank - 1, kax, 1, rax, o);
}

// ---------------------------------------------------------------- argmax
template <typename T, bool MAXI>
__global__ void hb_argext_kernel(hb_dims d /* kept axes */, int64_t n, int64_t stride_last,
                                 const T *__restrict__ a, int64_t *__restrict__ out, int64_t K) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  int64_t off = hb_offset_of(d, k);
  T best = a[off];
  int64_t bi = 0;
  for (int64_t i = 1; i < n; i++) {
    T v = a[off + i * stride_last];
    if (MAXI ? (v > best) : (v < best)) {  // strict: first extremum wins
      best = v;
      bi = i;
    }
  }
  out[k] = bi;
}

template <bool MAXI>
static int argext(const hb_array *a, hb_array **out) {
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(a && out, "null argument");
  HB_REQUIRE(a->rank >= 1, "argMax: rank 0");
  HB_REQUIRE(a->dt != HB_BOOL, "argMax: numeric dtype required");
  if (hb_lazy_on())
    return hb_lrec(HB_L_GENERIC, MAXI ? "argmax_last" : "argmin_last").in(a).out(out, HB_I64, a->rank - 1, a->shape).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) { return argext<MAXI>(in[0], &o[0]); });
  int64_t n = a->shape[a->rank - 1];
  hb_array *o;
  HB_TRY(hb_new_array(HB_I64, a->rank - 1, a->shape, &o));
  int64_t K = hb_numel(o);
  if (K > 0) {
    if (n == 0) {
      HB_CUDA(cudaMemsetAsync(hb_data(o), 0, (size_t)K * 8, g_hb.stream));
    } else {
      hb_dims d;
      d.rank = a->rank - 1;
      for (int i = 0; i < HB_R; i++) {
        d.shape[i] = i < d.rank ? a->shape[i] : 1;
        d.stride[i] = i < d.rank ? a->strides[i] : 0;
      }
      HB_DISPATCH_NUM(a->dt, T,
                      (hb_argext_kernel<T, MAXI><<<hb_blocks_for(K, 128), 128, 0, g_hb.stream>>>(
                          d, n, a->strides[a->rank - 1], (const T *)hb_data(a), (int64_t *)hb_data(o), K)));
      HB_LAUNCHED();
    }
  }
  *out = o;
  return HB_OK;
}

extern "C" int hb_argmax_last(const hb_array *a, hb_array **out) { return argext<true>(a, out); }
extern "C" int hb_argmin_last(const hb_array *a, hb_array **out) { return argext<false>(a, out); }

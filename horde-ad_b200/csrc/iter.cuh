// iter.cuh -- strided N-d iteration kernels shared by the elementwise, cast,
// copy and fused-map families.  An "iteration space" is a collapsed logical
// shape with one stride vector per operand (stride 0 = broadcast view of
// sreplicate, negative = sreverse), so no view is ever materialised before it
// is consumed.
//
// Three launch shapes:
//   vec  : every operand contiguous (or a broadcast scalar): 128-bit loads /
//          stores, 2 independent vectors in flight per thread.
//   rows : innermost collapsed axis long: a block owns (outer index, chunk),
//          the outer index is decomposed once per thread, threads run along
//          the contiguous axis (coalesced for every unit-stride operand).
//   flat : anything else: one element per thread, full index decomposition,
//          32-bit arithmetic when the space is < 2^31 elements.
#pragma once
#include "common.cuh"

template <int NOPS>
struct hb_iterp {
  int rank;
  int64_t shape[HB_R];
  int64_t stride[NOPS][HB_R];
};

template <int NOPS>
static inline void hb_iterp_from(const hb_iter &it, hb_iterp<NOPS> *p) {
  p->rank = it.rank;
  for (int i = 0; i < HB_R; i++) {
    p->shape[i] = i < it.rank ? it.shape[i] : 1;
    for (int o = 0; o < NOPS; o++) p->stride[o][i] = i < it.rank ? it.stride[o][i] : 0;
  }
}

#ifdef __CUDACC__

template <int NOPS, typename I, typename F>
__global__ void __launch_bounds__(256) hb_iter_flat_kernel(hb_iterp<NOPS> it, int64_t n, F f) {
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t lin = gid; lin < n; lin += gstride) {
    int64_t off[NOPS];
#pragma unroll
    for (int o = 0; o < NOPS; o++) off[o] = 0;
    I rem = (I)lin;
#pragma unroll 1
    for (int d = it.rank - 1; d >= 0; d--) {
      I sh = (I)it.shape[d];
      I q = rem / sh;
      I r = rem - q * sh;
#pragma unroll
      for (int o = 0; o < NOPS; o++) off[o] += (int64_t)r * it.stride[o][d];
      rem = q;
    }
    f(lin, off);
  }
}

// blockIdx.x = outer * nchunks + chunk; inner axis = last collapsed axis
template <int NOPS, typename F>
__global__ void __launch_bounds__(256)
    hb_iter_rows_kernel(hb_iterp<NOPS> it, int64_t inner, int nchunks, int chunk, F f) {
  int64_t outer = blockIdx.x / nchunks;
  int c = blockIdx.x - outer * nchunks;
  int64_t base[NOPS];
#pragma unroll
  for (int o = 0; o < NOPS; o++) base[o] = 0;
  int64_t rem = outer;
#pragma unroll 1
  for (int d = it.rank - 2; d >= 0; d--) {
    int64_t q = rem / it.shape[d];
    int64_t r = rem - q * it.shape[d];
#pragma unroll
    for (int o = 0; o < NOPS; o++) base[o] += r * it.stride[o][d];
    rem = q;
  }
  int64_t i0 = (int64_t)c * chunk;
  int64_t i1 = i0 + chunk < inner ? i0 + chunk : inner;
  int last = it.rank - 1;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    int64_t off[NOPS];
#pragma unroll
    for (int o = 0; o < NOPS; o++) off[o] = base[o] + i * it.stride[o][last];
    f(outer * inner + i, off);
  }
}

// blockIdx.x = outer * nchunks + chunk, where "inner" is the product of the last
// `irank` collapsed axes (short innermost axes, e.g. the [14,14,5] tail of a
// broadcast im2col product): the outer index is decomposed once per block and
// each element pays only `irank` small 32-bit divisions.
template <int NOPS, typename F>
__global__ void __launch_bounds__(256)
    hb_iter_tile_kernel(hb_iterp<NOPS> it, int irank, int inner, int nchunks, int chunk, F f) {
  int64_t outer = blockIdx.x / nchunks;
  int c = blockIdx.x - outer * nchunks;
  int64_t base[NOPS];
#pragma unroll
  for (int o = 0; o < NOPS; o++) base[o] = 0;
  int64_t rem = outer;
  const int orank = it.rank - irank;
#pragma unroll 1
  for (int d = orank - 1; d >= 0; d--) {
    int64_t q = rem / it.shape[d];
    int64_t r = rem - q * it.shape[d];
#pragma unroll
    for (int o = 0; o < NOPS; o++) base[o] += r * it.stride[o][d];
    rem = q;
  }
  int i0 = c * chunk;
  int i1 = i0 + chunk < inner ? i0 + chunk : inner;
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    int64_t off[NOPS];
#pragma unroll
    for (int o = 0; o < NOPS; o++) off[o] = base[o];
    unsigned ri = (unsigned)i;
#pragma unroll 1
    for (int d = it.rank - 1; d >= orank; d--) {
      unsigned sh = (unsigned)it.shape[d];
      unsigned q = ri / sh, r = ri - q * sh;
#pragma unroll
      for (int o = 0; o < NOPS; o++) off[o] += (int64_t)r * it.stride[o][d];
      ri = q;
    }
    f(outer * inner + i, off);
  }
}

// Launch f(linear_index, offsets[NOPS]) over the iteration space.
template <int NOPS, typename F>
static int hb_launch_iter(const hb_iter &it, F f) {
  int64_t n = 1;
  for (int i = 0; i < it.rank; i++) n *= it.shape[i];
  if (n == 0) return HB_OK;
  hb_iterp<NOPS> p;
  hb_iterp_from<NOPS>(it, &p);
  int64_t inner = it.rank > 0 ? it.shape[it.rank - 1] : 1;
  if (it.rank >= 2 && inner >= 256) {
    int chunk = 2048;
    int nchunks = (int)((inner + chunk - 1) / chunk);
    int64_t outer = n / inner;
    int64_t blocks = outer * nchunks;
    if (blocks < (1ll << 31)) {
      hb_iter_rows_kernel<NOPS, F><<<(unsigned)blocks, 256, 0, g_hb.stream>>>(p, inner, nchunks, chunk, f);
      HB_LAUNCHED();
      return HB_OK;
    }
  }
  if (it.rank >= 3 && n >= (1 << 16)) {
    // short innermost axes: fold as many trailing axes as needed into a tile of >= 512 elements
    int irank = 0;
    int64_t tile = 1;
    while (irank < it.rank - 1 && tile < 512) { tile *= it.shape[it.rank - 1 - irank]; irank++; }
    int64_t outer = n / tile;
    if (tile >= 64 && tile < (1 << 24)) {
      int chunk = 2048;
      int nchunks = (int)((tile + chunk - 1) / chunk);
      int64_t blocks = outer * nchunks;
      if (blocks < (1ll << 31)) {
        hb_iter_tile_kernel<NOPS, F><<<(unsigned)blocks, 256, 0, g_hb.stream>>>(p, irank, (int)tile, nchunks, chunk, f);
        HB_LAUNCHED();
        return HB_OK;
      }
    }
  }
  int blocks = hb_blocks_for(n, 256, g_hb.sm_count * 32);
  if (n < (1ll << 31))
    hb_iter_flat_kernel<NOPS, uint32_t, F><<<blocks, 256, 0, g_hb.stream>>>(p, n, f);
  else
    hb_iter_flat_kernel<NOPS, int64_t, F><<<blocks, 256, 0, g_hb.stream>>>(p, n, f);
  HB_LAUNCHED();
  return HB_OK;
}

// ------------------------------------------------------------ vector path
template <typename T, int V>
struct alignas(sizeof(T) * V) hb_vec {
  T v[V];
};

// out[i] = f(a[i], b[i]) over contiguous storage; sa/sb: operand is one
// broadcast scalar.  V elements per vector, 2 vectors per thread.
template <typename TO, typename TA, typename TB, int V, typename F>
__global__ void __launch_bounds__(256)
    hb_vec2_kernel(TO *__restrict__ o, const TA *__restrict__ a, const TB *__restrict__ b, bool sa,
                   bool sb, int64_t nvec, int64_t n, F f) {
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  TA as = sa ? a[0] : TA();
  TB bs = sb ? b[0] : TB();
  const hb_vec<TA, V> *av = reinterpret_cast<const hb_vec<TA, V> *>(a);
  const hb_vec<TB, V> *bv = reinterpret_cast<const hb_vec<TB, V> *>(b);
  hb_vec<TO, V> *ov = reinterpret_cast<hb_vec<TO, V> *>(o);
  int64_t i = gid;
  for (; i + gstride < nvec; i += 2 * gstride) {
    hb_vec<TA, V> a0, a1;
    hb_vec<TB, V> b0, b1;
    if (!sa) { a0 = av[i]; a1 = av[i + gstride]; }
    if (!sb) { b0 = bv[i]; b1 = bv[i + gstride]; }
    hb_vec<TO, V> r0, r1;
#pragma unroll
    for (int k = 0; k < V; k++) {
      r0.v[k] = f(sa ? as : a0.v[k], sb ? bs : b0.v[k]);
      r1.v[k] = f(sa ? as : a1.v[k], sb ? bs : b1.v[k]);
    }
    ov[i] = r0;
    ov[i + gstride] = r1;
  }
  for (; i < nvec; i += gstride) {
    hb_vec<TA, V> a0;
    hb_vec<TB, V> b0;
    if (!sa) a0 = av[i];
    if (!sb) b0 = bv[i];
    hb_vec<TO, V> r0;
#pragma unroll
    for (int k = 0; k < V; k++) r0.v[k] = f(sa ? as : a0.v[k], sb ? bs : b0.v[k]);
    ov[i] = r0;
  }
  // scalar tail
  for (int64_t j = nvec * V + gid; j < n; j += gstride) o[j] = f(sa ? as : a[j], sb ? bs : b[j]);
}

template <typename T>
struct hb_vecwidth {
  static constexpr int value = 16 / sizeof(T);
};

static inline bool hb_aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

// Returns true if it handled the operation on the vector path.
// a/b: contiguous or all-stride-0 (scalar) operands of n logical elements.
template <typename TO, typename TA, typename TB, typename F>
static int hb_try_vec2(TO *o, const TA *a, bool sa, const TB *b, bool sb, int64_t n, F f,
                       bool *handled) {
  constexpr size_t big = sizeof(TO) > sizeof(TA) ? (sizeof(TO) > sizeof(TB) ? sizeof(TO) : sizeof(TB))
                                                 : (sizeof(TA) > sizeof(TB) ? sizeof(TA) : sizeof(TB));
  constexpr int V = 16 / big;
  *handled = false;
  if ((((uintptr_t)o) % (sizeof(TO) * V)) || (!sa && ((uintptr_t)a) % (sizeof(TA) * V)) ||
      (!sb && ((uintptr_t)b) % (sizeof(TB) * V)))
    return HB_OK;
  int64_t nvec = n / V;
  int blocks = hb_blocks_for((nvec + 1) / 2 + 1, 256, g_hb.sm_count * 16);
  hb_vec2_kernel<TO, TA, TB, V, F><<<blocks, 256, 0, g_hb.stream>>>(o, a, b, sa, sb, nvec, n, f);
  HB_LAUNCHED();
  *handled = true;
  return HB_OK;
}

// out (contiguous) = f(a, b) where the innermost collapsed axis of every
// operand has stride 1 or 0 and the outer axes are arbitrary (bias add through
// a stride-0 sreplicate view, NCHW tensor x per-channel factor, a row times a
// column): 128-bit accesses on the unit-stride operands, one 32-bit
// decomposition of the outer index per vector, 2 vectors in flight per thread.
// The one-element-per-thread rows kernel moved such a bias add at 57 % of HBM.
template <typename TO, typename TA, typename TB, int V, typename F>
__global__ void __launch_bounds__(256)
    hb_vecrows2_kernel(TO *__restrict__ o, const TA *__restrict__ a, const TB *__restrict__ b,
                       const __grid_constant__ hb_iterp<3> it, unsigned nvrow, int64_t nvec, F f) {
  const int orank = it.rank - 1;
  const bool ua = it.stride[1][orank] != 0, ub = it.stride[2][orank] != 0;
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  hb_vec<TO, V> *ov = reinterpret_cast<hb_vec<TO, V> *>(o);
  auto offsets = [&](int64_t v, int64_t &oa, int64_t &ob) {
    unsigned outer = (unsigned)(v / nvrow);
    unsigned iv = (unsigned)(v - (int64_t)outer * nvrow);
    oa = ua ? (int64_t)iv * V : 0;
    ob = ub ? (int64_t)iv * V : 0;
#pragma unroll 1
    for (int d = orank - 1; d >= 0; d--) {
      unsigned sh = (unsigned)it.shape[d];
      unsigned q = outer / sh, r = outer - q * sh;
      oa += (int64_t)r * it.stride[1][d];
      ob += (int64_t)r * it.stride[2][d];
      outer = q;
    }
  };
  auto load_a = [&](int64_t off, hb_vec<TA, V> &x) {
    if (ua) x = *reinterpret_cast<const hb_vec<TA, V> *>(a + off);
    else {
      TA s = a[off];
#pragma unroll
      for (int k = 0; k < V; k++) x.v[k] = s;
    }
  };
  auto load_b = [&](int64_t off, hb_vec<TB, V> &x) {
    if (ub) x = *reinterpret_cast<const hb_vec<TB, V> *>(b + off);
    else {
      TB s = b[off];
#pragma unroll
      for (int k = 0; k < V; k++) x.v[k] = s;
    }
  };
  int64_t i = gid;
  for (; i + gstride < nvec; i += 2 * gstride) {
    int64_t oa0, ob0, oa1, ob1;
    offsets(i, oa0, ob0);
    offsets(i + gstride, oa1, ob1);
    hb_vec<TA, V> a0, a1;
    hb_vec<TB, V> b0, b1;
    load_a(oa0, a0);
    load_a(oa1, a1);
    load_b(ob0, b0);
    load_b(ob1, b1);
    hb_vec<TO, V> r0, r1;
#pragma unroll
    for (int k = 0; k < V; k++) {
      r0.v[k] = f(a0.v[k], b0.v[k]);
      r1.v[k] = f(a1.v[k], b1.v[k]);
    }
    ov[i] = r0;
    ov[i + gstride] = r1;
  }
  for (; i < nvec; i += gstride) {
    int64_t oa0, ob0;
    offsets(i, oa0, ob0);
    hb_vec<TA, V> a0;
    hb_vec<TB, V> b0;
    load_a(oa0, a0);
    load_b(ob0, b0);
    hb_vec<TO, V> r0;
#pragma unroll
    for (int k = 0; k < V; k++) r0.v[k] = f(a0.v[k], b0.v[k]);
    ov[i] = r0;
  }
}

// Returns handled = true if (out, a, b) fit the vector-rows shape; `it` is the
// collapsed iteration space of (out, a, b) with out contiguous.
template <typename TO, typename TA, typename TB, typename F>
static int hb_try_vecrows2(TO *o, const TA *a, const TB *b, const hb_iter &it, F f, bool *handled) {
  constexpr size_t big = sizeof(TO) > sizeof(TA) ? (sizeof(TO) > sizeof(TB) ? sizeof(TO) : sizeof(TB))
                                                 : (sizeof(TA) > sizeof(TB) ? sizeof(TA) : sizeof(TB));
  constexpr int V = 16 / big;
  *handled = false;
  if (it.rank < 2) return HB_OK;
  const int last = it.rank - 1;
  const int64_t inner = it.shape[last];
  if (inner % V || inner < 4 * V || it.stride[0][last] != 1) return HB_OK;
  int64_t n = 1;
  for (int d = 0; d < it.rank; d++) n *= it.shape[d];
  if (n < (1 << 16) || n / inner >= (1ll << 31) || inner / V >= (1ll << 31)) return HB_OK;
  if (((uintptr_t)o) % (sizeof(TO) * V)) return HB_OK;
  const size_t esz[3] = {sizeof(TO), sizeof(TA), sizeof(TB)};
  const void *ptr[3] = {o, a, b};
  for (int op = 1; op < 3; op++) {
    int64_t sl = it.stride[op][last];
    if (sl != 0 && sl != 1) return HB_OK;
    if (sl == 1) {
      if (((uintptr_t)ptr[op]) % (esz[op] * V)) return HB_OK;
      for (int d = 0; d < last; d++)
        if (it.stride[op][d] % V) return HB_OK;
    }
  }
  hb_iterp<3> p;
  hb_iterp_from<3>(it, &p);
  int64_t nvec = n / V;
  int blocks = hb_blocks_for((nvec + 1) / 2 + 1, 256, g_hb.sm_count * 16);
  hb_vecrows2_kernel<TO, TA, TB, V, F><<<blocks, 256, 0, g_hb.stream>>>(o, a, b, p, (unsigned)(inner / V), nvec, f);
  HB_LAUNCHED();
  *handled = true;
  return HB_OK;
}

#endif  // __CUDACC__

// operand classification for the vector path
static inline bool hb_all_stride0(const hb_array *a) {
  for (int i = 0; i < a->rank; i++)
    if (a->shape[i] != 1 && a->strides[i] != 0) return false;
  return true;
}

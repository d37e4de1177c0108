// elementwise.cu -- map/zipWith families of the device carrier: unary and
// binary arithmetic (CarriersConcrete.hs:333-343,40-92), casts
// (OpsConcrete.hs:455-532), materialising copies (append/stack/contiguous,
// OpsConcrete.hs:125-184,641-643), iota (:536), whole-array comparison
// (CarriersConcrete.hs:310-331) and in-place cotangent accumulation
// (taddTarget in DeltaEval.hs:139-148,317-332).
//
// Roofline: every kernel here is HBM-bound, (k+1)*sizeof(T) algorithmic bytes
// per element for k inputs.
#include "iter.cuh"
#include "lazy.cuh"
#include "ops.cuh"

// ------------------------------------------------------------ fill scalar
__global__ void hb_fill_scalar_kernel(void *dst, uint64_t bits, int size) {
  if (size == 8) *(uint64_t *)dst = bits;
  else if (size == 4) *(uint32_t *)dst = (uint32_t)bits;
  else if (size == 2) *(uint16_t *)dst = (uint16_t)bits;
  else *(uint8_t *)dst = (uint8_t)bits;
}

int hb_fill_scalar(void *dst, hb_dtype dt, const void *value) {
  uint64_t bits = 0;
  memcpy(&bits, value, hb_dtype_size(dt));
  hb_fill_scalar_kernel<<<1, 1, 0, g_hb.stream>>>(dst, bits, (int)hb_dtype_size(dt));
  HB_LAUNCHED();
  return HB_OK;
}

// ------------------------------------------------------------ generic map
// out = f(a, b) with out fresh-contiguous (or a caller-provided view).
template <typename TO, typename TA, typename TB, typename F>
static int map2_into(hb_array *out, const hb_array *a, const hb_array *b, F f) {
  int64_t n = hb_numel(out);
  if (n == 0) return HB_OK;
  TO *po = (TO *)hb_data(out);
  const TA *pa = (const TA *)hb_data(a);
  const TB *pb = (const TB *)hb_data(b);
  bool ca = hb_contig(a), cb = hb_contig(b), sa = hb_all_stride0(a), sb = hb_all_stride0(b);
  if (hb_contig(out) && (ca || sa) && (cb || sb)) {
    bool handled = false;
    HB_TRY((hb_try_vec2<TO, TA, TB, F>(po, pa, sa && !(ca && n == 1), pb, sb && !(cb && n == 1), n, f,
                                       &handled)));
    if (handled) return HB_OK;
  }
  const hb_array *ops[3] = {out, a, b};
  hb_iter it;
  hb_make_iter(3, ops, &it);
  if (hb_contig(out)) {
    bool handled = false;
    HB_TRY((hb_try_vecrows2<TO, TA, TB, F>(po, pa, pb, it, f, &handled)));
    if (handled) return HB_OK;
  }
  auto body = [=] __device__(int64_t, const int64_t *off) {
    po[off[0]] = f(pa[off[1]], pb[off[2]]);
  };
  return hb_launch_iter<3>(it, body);
}

template <typename TO, typename TA, typename F>
static int map1_into(hb_array *out, const hb_array *a, F f) {
  auto f2 = [=] __device__(TA x, TA) { return f(x); };
  return map2_into<TO, TA, TA>(out, a, a, f2);
}

// -------------------------------------------------------------- contiguous
// dst[r, c] (contiguous [R, C]) = src[r + c * ld]: materialising a transposed matrix view (e.g. the
// cotangent of `str` in front of a dense layer) through a 32 x 32 shared tile, so that both the
// reads (along r) and the writes (along c) are coalesced; the generic strided iterator moved the
// 26 MB cotangent of the CNN's dense layer at 1.9 TB/s.
template <typename T>
__global__ void __launch_bounds__(256)
    hb_transpose2d_kernel(const T *__restrict__ src, int64_t ld, T *__restrict__ dst, int64_t R, int64_t C) {
  __shared__ T tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t tiles_r = (R + 31) / 32, tiles_c = (C + 31) / 32;
  for (int64_t t = blockIdx.x; t < tiles_r * tiles_c; t += gridDim.x) {
    const int64_t r0 = (t % tiles_r) * 32, c0 = (t / tiles_r) * 32;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      int64_t c = c0 + ty + 8 * k, r = r0 + tx;
      if (c < C && r < R) tile[ty + 8 * k][tx] = src[r + c * ld];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
      int64_t r = r0 + ty + 8 * k, c = c0 + tx;
      if (r < R && c < C) dst[r * C + c] = tile[tx][ty + 8 * k];
    }
    __syncthreads();
  }
}

static int copy_into(hb_array *dst, const hb_array *src) {
  // [R, c1, .., ck] with unit stride along R and the other axes collapsing to one axis of stride ld
  bool tr2d = src->rank >= 2 && src->rank == dst->rank && hb_contig(dst) && src->strides[0] == 1 && src->shape[0] >= 32;
  int64_t Cc = 1;
  for (int i = src->rank - 1; tr2d && i >= 1; i--) {
    if (i + 1 < src->rank && src->strides[i] != src->strides[i + 1] * src->shape[i + 1]) tr2d = false;
    Cc *= src->shape[i];
  }
  if (tr2d && Cc >= 32 && src->strides[src->rank - 1] >= src->shape[0]) {
    const int64_t R = src->shape[0], C = Cc;
    const int64_t tiles = ((R + 31) / 32) * ((C + 31) / 32);
    const int blocks = (int)(tiles < (int64_t)g_hb.sm_count * 16 ? tiles : (int64_t)g_hb.sm_count * 16);
    HB_DISPATCH_BYTES(src->dt, T,
                      (hb_transpose2d_kernel<T><<<blocks, 256, 0, g_hb.stream>>>((const T *)hb_data(src), src->strides[src->rank - 1],
                                                                                 (T *)hb_data(dst), R, C)));
    HB_LAUNCHED();
    return HB_OK;
  }
  HB_DISPATCH_BYTES(src->dt, T,
                    return (map1_into<T, T>(dst, src, [] __device__(T x) { return x; })));
  return HB_OK;
}

int hb_copy_into(hb_array *dst, const hb_array *src) { return copy_into(dst, src); }

extern "C" int hb_contiguous(const hb_array *a, hb_array **out) {
  hb_prof_scope prof_("contiguous", a);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(a && out, "null argument");
  if (hb_lazy_on())
    return hb_lrec(HB_L_COPY, "contiguous").in(a).out(out, a->dt, a->rank, a->shape).commit(
        [](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_contiguous(in[0], &o[0]); });
  hb_array *o;
  HB_TRY(hb_new_array(a->dt, a->rank, a->shape, &o));
  int st = copy_into(o, a);
  if (st != HB_OK) {
    hb_release(o);
    return st;
  }
  *out = o;
  return HB_OK;
}

extern "C" int hb_append(const hb_array *a, const hb_array *b, hb_array **out) {
  hb_prof_scope prof_("append", a, b);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(a && b && out, "null argument");
  HB_REQUIRE(a->rank >= 1 && a->rank == b->rank && a->dt == b->dt, "append: rank/dtype mismatch");
  for (int i = 1; i < a->rank; i++)
    HB_REQUIRE(a->shape[i] == b->shape[i], "append: %s vs %s", hb_shape_str(a).c_str(),
               hb_shape_str(b).c_str());
  int64_t sh[HB_R];
  memcpy(sh, a->shape, sizeof sh);
  sh[0] = a->shape[0] + b->shape[0];
  if (hb_lazy_on())
    return hb_lrec(HB_L_GENERIC, "append").in(a).in(b).out(out, a->dt, a->rank, sh).commit(
        [](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_append(in[0], in[1], &o[0]); });
  hb_array *o;
  HB_TRY(hb_new_array(a->dt, a->rank, sh, &o));
  hb_array *v1 = nullptr, *v2 = nullptr;
  int st = hb_slice(o, 0, a->shape[0], &v1);
  if (st == HB_OK) st = copy_into(v1, a);
  if (st == HB_OK) st = hb_slice(o, a->shape[0], b->shape[0], &v2);
  if (st == HB_OK) st = copy_into(v2, b);
  hb_release(v1);
  hb_release(v2);
  if (st != HB_OK) {
    hb_release(o);
    return st;
  }
  *out = o;
  return HB_OK;
}

extern "C" int hb_stack(int k, hb_array *const *arrs, hb_array **out) {
  hb_prof_scope prof_("stack");
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(out, "null argument");
  HB_REQUIRE(k > 0 && arrs, "sfromVector: empty vector");  // OpsConcrete.hs:129
  const hb_array *a0 = arrs[0];
  HB_REQUIRE(a0->rank + 1 <= HB_R, "stack: rank too large");
  for (int i = 1; i < k; i++)
    HB_REQUIRE(hb_same_shape(a0, arrs[i]) && arrs[i]->dt == a0->dt, "stack: element %d has shape %s, expected %s",
               i, hb_shape_str(arrs[i]).c_str(), hb_shape_str(a0).c_str());
  int64_t sh[HB_R];
  sh[0] = k;
  for (int i = 0; i < a0->rank; i++) sh[i + 1] = a0->shape[i];
  if (hb_lazy_on()) {
    hb_lrec r(HB_L_GENERIC, "stack");
    for (int i = 0; i < k; i++) r.in(arrs[i]);
    return r.out(out, a0->dt, a0->rank + 1, sh).commit(
        [](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_stack((int)n.in.size(), in, &o[0]); });
  }
  hb_array *o;
  HB_TRY(hb_new_array(a0->dt, a0->rank + 1, sh, &o));
  for (int i = 0; i < k; i++) {
    int64_t ix = i;
    hb_array *v = nullptr;
    int st = hb_index_prefix(o, 1, &ix, &v);
    if (st == HB_OK) st = copy_into(v, arrs[i]);
    hb_release(v);
    if (st != HB_OK) {
      hb_release(o);
      return st;
    }
  }
  *out = o;
  return HB_OK;
}

// Flatten k arrays (row-major) and concatenate them into one vector: the
// contiguous gradient bucket of the data-parallel loop.
// all leaves contiguous (the usual case: gradient leaves): ONE launch moves up
// to 16 of them; blockIdx.y selects the leaf
struct hb_concat_tab {
  const void *src[16];
  int64_t off[16], n[16];
};
template <typename T>
__global__ void __launch_bounds__(256) hb_concat_kernel(const __grid_constant__ hb_concat_tab tab, T *__restrict__ dst) {
  const int k = blockIdx.y;
  const T *s = (const T *)tab.src[k];
  T *d = dst + tab.off[k];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tab.n[k]; i += (int64_t)gridDim.x * blockDim.x)
    d[i] = s[i];
}

extern "C" int hb_flatten_concat(int k, hb_array *const *arrs, hb_array **out) {
  hb_prof_scope prof_("flatten_concat");
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(out && k > 0 && arrs, "flatten_concat: empty list");
  int64_t total = 0, maxn = 0;
  bool all_contig = k <= 16;
  for (int i = 0; i < k; i++) {
    HB_REQUIRE(arrs[i] && arrs[i]->dt == arrs[0]->dt, "flatten_concat: dtype mismatch at %d", i);
    int64_t n = hb_numel(arrs[i]);
    total += n;
    if (n > maxn) maxn = n;
    if (!hb_contig(arrs[i])) all_contig = false;
  }
  if (hb_lazy_on()) {
    hb_lrec r(HB_L_GENERIC, "flatten_concat");
    for (int i = 0; i < k; i++) r.in(arrs[i]);
    return r.out(out, arrs[0]->dt, 1, &total).commit(
        [](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_flatten_concat((int)n.in.size(), in, &o[0]); });
  }
  hb_array *o;
  HB_TRY(hb_new_array(arrs[0]->dt, 1, &total, &o));
  if (all_contig && total > 0) {
    hb_concat_tab tab;
    int64_t p = 0;
    for (int i = 0; i < 16; i++) {
      tab.src[i] = i < k ? hb_data(arrs[i]) : nullptr;
      tab.n[i] = i < k ? hb_numel(arrs[i]) : 0;
      tab.off[i] = p;
      p += tab.n[i];
    }
    dim3 grid((unsigned)hb_blocks_for(maxn, 256 * 4, g_hb.sm_count * 4), (unsigned)k);
    HB_DISPATCH_BYTES(arrs[0]->dt, T, (hb_concat_kernel<T><<<grid, 256, 0, g_hb.stream>>>(tab, (T *)hb_data(o))));
    HB_LAUNCHED();
    *out = o;
    return HB_OK;
  }
  int64_t pos = 0;
  for (int i = 0; i < k; i++) {
    int64_t n = hb_numel(arrs[i]);
    if (n == 0) continue;
    hb_array *v = hb_new_view(o);
    v->offset += pos;
    v->rank = arrs[i]->rank;
    int64_t st = 1;
    for (int d = arrs[i]->rank - 1; d >= 0; d--) {
      v->shape[d] = arrs[i]->shape[d];
      v->strides[d] = st;
      st *= arrs[i]->shape[d];
    }
    int s = copy_into(v, arrs[i]);
    hb_release(v);
    if (s != HB_OK) {
      hb_release(o);
      return s;
    }
    pos += n;
  }
  *out = o;
  return HB_OK;
}

template <typename T>
__global__ void hb_iota_kernel(T *o, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = (T)i;
}

extern "C" int hb_iota(hb_dtype dt, int64_t n, hb_array **out) {
  HB_CHECK_INIT_LAZY();   // a constant: created eagerly even while recording
  HB_REQUIRE(out && n >= 0, "iota: bad argument");
  hb_array *o;
  HB_TRY(hb_new_array(dt, 1, &n, &o));
  if (n > 0 && (n * hb_dtype_size(dt) <= 4096 || g_hb.trace_only) && dt != HB_BOOL) {
    auto img = std::make_shared<std::vector<uint8_t>>((size_t)n * hb_dtype_size(dt));
    for (int64_t i = 0; i < n; i++) {
      HB_DISPATCH_NUM(dt, T, ((T *)img->data())[i] = (T)i);
    }
    o->st->host = img;
    if (g_hb.trace_only) {
      memcpy(hb_data(o), img->data(), img->size());
      *out = o;
      return HB_OK;
    }
  }
  if (n > 0) {
    HB_DISPATCH_ALL(dt, T, hb_iota_kernel<T><<<hb_blocks_for(n, 256), 256, 0, g_hb.stream>>>((T *)hb_data(o), n));
    HB_LAUNCHED();
  }
  *out = o;
  return HB_OK;
}

// ------------------------------------------------------------------- unary
template <typename T, int OP>
static int unary_typed(hb_array *o, const hb_array *a) {
  return map1_into<T, T>(o, a, [] __device__(T x) { return hb_un<OP>(x); });
}

#define UN_CASE(OPC) \
  case OPC: st = unary_typed<T, OPC>(o, a); break;

template <typename T>
static int unary_dispatch(hb_unop op, hb_array *o, const hb_array *a) {
  int st = HB_OK;
  if constexpr (hb_is_fp<T>::value) {
    switch (op) {
      UN_CASE(HB_NEGATE) UN_CASE(HB_ABS) UN_CASE(HB_SIGNUM) UN_CASE(HB_RECIP) UN_CASE(HB_EXP)
      UN_CASE(HB_LOG) UN_CASE(HB_SQRT) UN_CASE(HB_SIN) UN_CASE(HB_COS) UN_CASE(HB_TAN)
      UN_CASE(HB_ASIN) UN_CASE(HB_ACOS) UN_CASE(HB_ATAN) UN_CASE(HB_SINH) UN_CASE(HB_COSH)
      UN_CASE(HB_TANH) UN_CASE(HB_ASINH) UN_CASE(HB_ACOSH) UN_CASE(HB_ATANH)
      default: HB_FAIL(HB_ERR_INVALID, "bad unary op %d", (int)op);
    }
  } else {
    switch (op) {
      UN_CASE(HB_NEGATE) UN_CASE(HB_ABS) UN_CASE(HB_SIGNUM)
      default: HB_FAIL(HB_ERR_INVALID, "unary op %d requires a floating dtype", (int)op);
    }
  }
  return st;
}

extern "C" int hb_unary(hb_unop op, const hb_array *a, hb_array **out) {
  hb_prof_scope prof_("unary", a);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(a && out, "null argument");
  if (hb_lazy_on())
    return hb_lrec(HB_L_UNARY, "unary").in(a).iarg(0, (int)op).out(out, a->dt, a->rank, a->shape).commit(
        [](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_unary((hb_unop)n.ia[0], in[0], &o[0]); });
  hb_array *o;
  HB_TRY(hb_new_array(a->dt, a->rank, a->shape, &o));
  int st = HB_OK;
  switch (a->dt) {
    case HB_F64: st = unary_dispatch<double>(op, o, a); break;
    case HB_F32: st = unary_dispatch<float>(op, o, a); break;
    case HB_I64: st = unary_dispatch<int64_t>(op, o, a); break;
    case HB_I32: st = unary_dispatch<int32_t>(op, o, a); break;
    case HB_I16: st = unary_dispatch<int16_t>(op, o, a); break;
    case HB_I8: st = unary_dispatch<int8_t>(op, o, a); break;
    default: hb_set_error("unary op on dtype %s", hb_dtype_name(a->dt)); st = HB_ERR_INVALID;
  }
  if (st != HB_OK) {
    hb_release(o);
    return st;
  }
  *out = o;
  return HB_OK;
}

// ------------------------------------------------------------------ binary
template <typename T, int OP>
static int binary_typed(hb_array *o, const hb_array *a, const hb_array *b) {
  return map2_into<T, T, T>(o, a, b, [] __device__(T x, T y) { return hb_bin<OP>(x, y); });
}
#define BIN_CASE(OPC) \
  case OPC: st = binary_typed<T, OPC>(o, a, b); break;

template <typename T>
static int binary_dispatch(hb_binop op, hb_array *o, const hb_array *a, const hb_array *b) {
  int st = HB_OK;
  if constexpr (hb_is_fp<T>::value) {
    switch (op) {
      BIN_CASE(HB_ADD) BIN_CASE(HB_SUB) BIN_CASE(HB_MUL) BIN_CASE(HB_DIVIDE) BIN_CASE(HB_POWER)
      BIN_CASE(HB_LOGBASE) BIN_CASE(HB_ATAN2) BIN_CASE(HB_MAX) BIN_CASE(HB_MIN)
      default: HB_FAIL(HB_ERR_INVALID, "binary op %d is not defined on floating arrays", (int)op);
    }
  } else {
    switch (op) {
      BIN_CASE(HB_ADD) BIN_CASE(HB_SUB) BIN_CASE(HB_MUL) BIN_CASE(HB_QUOT) BIN_CASE(HB_REM)
      BIN_CASE(HB_MAX) BIN_CASE(HB_MIN)
      default: HB_FAIL(HB_ERR_INVALID, "binary op %d is not defined on integral arrays", (int)op);
    }
  }
  return st;
}

static int binary_into(hb_binop op, hb_array *o, const hb_array *a, const hb_array *b) {
  switch (a->dt) {
    case HB_F64: return binary_dispatch<double>(op, o, a, b);
    case HB_F32: return binary_dispatch<float>(op, o, a, b);
    case HB_I64: return binary_dispatch<int64_t>(op, o, a, b);
    case HB_I32: return binary_dispatch<int32_t>(op, o, a, b);
    case HB_I16: return binary_dispatch<int16_t>(op, o, a, b);
    case HB_I8: return binary_dispatch<int8_t>(op, o, a, b);
    default: HB_FAIL(HB_ERR_INVALID, "binary op on dtype %s", hb_dtype_name(a->dt));
  }
}

extern "C" int hb_binary(hb_binop op, const hb_array *a, const hb_array *b, hb_array **out) {
  hb_prof_scope prof_("binary", a, b);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(a && b && out, "null argument");
  HB_REQUIRE(a->dt == b->dt, "binary op: dtype %s vs %s", hb_dtype_name(a->dt), hb_dtype_name(b->dt));
  HB_REQUIRE(hb_same_shape(a, b), "binary op: shape %s vs %s", hb_shape_str(a).c_str(),
             hb_shape_str(b).c_str());
  if (hb_lazy_on())
    return hb_lrec(HB_L_BINARY, "binary").in(a).in(b).iarg(0, (int)op).out(out, a->dt, a->rank, a->shape).commit(
        [](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_binary((hb_binop)n.ia[0], in[0], in[1], &o[0]); });
  hb_array *o;
  HB_TRY(hb_new_array(a->dt, a->rank, a->shape, &o));
  int st = binary_into(op, o, a, b);
  if (st != HB_OK) {
    hb_release(o);
    return st;
  }
  *out = o;
  return HB_OK;
}

extern "C" int hb_add_inplace_or_new(hb_array *acc, const hb_array *c, hb_array **out) {
  hb_prof_scope prof_("add_inplace_or_new", acc, c);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(acc && c && out, "null argument");
  HB_REQUIRE(acc->dt == c->dt && hb_same_shape(acc, c), "taddTarget: %s %s vs %s %s",
             hb_dtype_name(acc->dt), hb_shape_str(acc).c_str(), hb_dtype_name(c->dt),
             hb_shape_str(c).c_str());
  if (hb_lazy_on())
    // taddTarget: whether the sum may overwrite `acc` is decided by the program's liveness analysis (acc dies here),
    // not by the reference counts at recording time
    return hb_lrec(HB_L_ADD_ACC, "add_acc").in(acc).in(c).out(out, acc->dt, acc->rank, acc->shape).commit(
        [](hb_lnode &n, hb_array *const *in, hb_array **o) {
          if (n.inplace_ok) {
            HB_TRY(binary_into(HB_ADD, in[0], in[0], in[1]));
            hb_retain(in[0]);
            o[0] = in[0];
            return HB_OK;
          }
          return hb_binary(HB_ADD, in[0], in[1], &o[0]);
        });
  bool unique = acc->rc.load() == 1 && acc->st && acc->st->rc.load() == 1 && hb_contig(acc) &&
                (c->st != acc->st);
  if (unique) {
    HB_TRY(binary_into(HB_ADD, acc, acc, c));
    hb_retain(acc);
    *out = acc;
    return HB_OK;
  }
  return hb_binary(HB_ADD, acc, c, out);
}

// ------------------------------------------------------------------- casts
template <typename TO, typename TA>
static int cast_typed(hb_array *o, const hb_array *a) {
  if constexpr (std::is_same<TO, uint8_t>::value)
    return map1_into<TO, TA>(o, a, [] __device__(TA x) { return (TO)(x != (TA)0); });
  else
    return map1_into<TO, TA>(o, a, [] __device__(TA x) { return (TO)x; });
}

template <typename TA>
static int cast_from(hb_array *o, const hb_array *a) {
  switch (o->dt) {
    case HB_F64: return cast_typed<double, TA>(o, a);
    case HB_F32: return cast_typed<float, TA>(o, a);
    case HB_I64: return cast_typed<int64_t, TA>(o, a);
    case HB_I32: return cast_typed<int32_t, TA>(o, a);
    case HB_I16: return cast_typed<int16_t, TA>(o, a);
    case HB_I8: return cast_typed<int8_t, TA>(o, a);
    case HB_BOOL: return cast_typed<uint8_t, TA>(o, a);
    default: HB_FAIL(HB_ERR_INVALID, "bad dtype");
  }
}

extern "C" int hb_cast(const hb_array *a, hb_dtype dt, hb_array **out) {
  hb_prof_scope prof_("cast", a);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(a && out, "null argument");
  if (hb_lazy_on())
    return hb_lrec(HB_L_GENERIC, "cast").in(a).iarg(0, (int)dt).out(out, dt, a->rank, a->shape).commit(
        [](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_cast(in[0], (hb_dtype)n.ia[0], &o[0]); });
  hb_array *o;
  HB_TRY(hb_new_array(dt, a->rank, a->shape, &o));
  int st;
  switch (a->dt) {
    case HB_F64: st = cast_from<double>(o, a); break;
    case HB_F32: st = cast_from<float>(o, a); break;
    case HB_I64: st = cast_from<int64_t>(o, a); break;
    case HB_I32: st = cast_from<int32_t>(o, a); break;
    case HB_I16: st = cast_from<int16_t>(o, a); break;
    case HB_I8: st = cast_from<int8_t>(o, a); break;
    case HB_BOOL: st = cast_from<uint8_t>(o, a); break;
    default: hb_set_error("bad dtype"); st = HB_ERR_INVALID;
  }
  if (st != HB_OK) {
    hb_release(o);
    return st;
  }
  *out = o;
  return HB_OK;
}

template <typename TO, typename TA>
static int floor_typed(hb_array *o, const hb_array *a) {
  return map1_into<TO, TA>(o, a, [] __device__(TA x) { return (TO)floor(x); });
}

extern "C" int hb_floor(const hb_array *a, hb_dtype dt, hb_array **out) {
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(a && out, "null argument");
  HB_REQUIRE(hb_is_float(a->dt), "floor: source must be floating, got %s", hb_dtype_name(a->dt));
  HB_REQUIRE(!hb_is_float(dt) && dt != HB_BOOL, "floor: target must be integral, got %s", hb_dtype_name(dt));
  if (hb_lazy_on())
    return hb_lrec(HB_L_GENERIC, "floor").in(a).iarg(0, (int)dt).out(out, dt, a->rank, a->shape).commit(
        [](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_floor(in[0], (hb_dtype)n.ia[0], &o[0]); });
  hb_array *o;
  HB_TRY(hb_new_array(dt, a->rank, a->shape, &o));
  int st = HB_OK;
#define FL(TA)                                                         \
  switch (dt) {                                                        \
    case HB_I64: st = floor_typed<int64_t, TA>(o, a); break;           \
    case HB_I32: st = floor_typed<int32_t, TA>(o, a); break;           \
    case HB_I16: st = floor_typed<int16_t, TA>(o, a); break;           \
    default: st = floor_typed<int8_t, TA>(o, a); break;                \
  }
  if (a->dt == HB_F64) { FL(double) } else { FL(float) }
#undef FL
  if (st != HB_OK) {
    hb_release(o);
    return st;
  }
  *out = o;
  return HB_OK;
}

// ------------------------------------------------------ whole-array compare
// first mismatch position (row-major) via a min-reduction: deterministic.
template <typename T>
__global__ void hb_first_diff_kernel(hb_iterp<2> it, const T *a, const T *b, int64_t n,
                                     unsigned long long *first) {
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long best = ~0ull;
  for (int64_t lin = gid; lin < n; lin += gstride) {
    int64_t oa = 0, ob = 0, rem = lin;
    for (int d = it.rank - 1; d >= 0; d--) {
      int64_t q = rem / it.shape[d], r = rem - q * it.shape[d];
      oa += r * it.stride[0][d];
      ob += r * it.stride[1][d];
      rem = q;
    }
    if (!(a[oa] == b[ob])) {
      best = (unsigned long long)lin;
      break;  // later elements of this thread have larger indices
    }
  }
  if (best != ~0ull) atomicMin(first, best);
}

template <typename T>
__global__ void hb_cmp_at_kernel(hb_iterp<2> it, const T *a, const T *b, const unsigned long long *first,
                                 int op, int *result) {
  unsigned long long f = *first;
  if (f == ~0ull) {  // all equal
    *result = (op == 0 || op == 1) ? 1 : 0;
    return;
  }
  int64_t oa = 0, ob = 0, rem = (int64_t)f;
  for (int d = it.rank - 1; d >= 0; d--) {
    int64_t q = rem / it.shape[d], r = rem - q * it.shape[d];
    oa += r * it.stride[0][d];
    ob += r * it.stride[1][d];
    rem = q;
  }
  T x = a[oa], y = b[ob];
  *result = op == 0 ? 0 : (op == 1 ? (x <= y ? 1 : 0) : (x < y ? 1 : 0));
}

extern "C" int hb_compare_all(int op, const hb_array *a, const hb_array *b, int *result) {
  HB_CHECK_INIT();
  HB_REQUIRE(a && b && result, "null argument");
  HB_REQUIRE(op >= 0 && op <= 2, "compare: bad op %d", op);
  HB_REQUIRE(a->dt == b->dt && hb_same_shape(a, b), "compare: operands differ in type or shape");
  int64_t n = hb_numel(a);
  if (n == 0) {
    *result = (op == 0 || op == 1) ? 1 : 0;
    return HB_OK;
  }
  void *scr;
  HB_TRY(hb_scratch(64, &scr));
  unsigned long long *first = (unsigned long long *)scr;
  int *res = (int *)((char *)scr + 16);
  HB_CUDA(cudaMemsetAsync(first, 0xff, 8, g_hb.stream));
  const hb_array *ops[2] = {a, b};
  hb_iter it;
  hb_make_iter(2, ops, &it);
  hb_iterp<2> p;
  hb_iterp_from<2>(it, &p);
  int blocks = hb_blocks_for(n, 256, g_hb.sm_count * 8);
  HB_DISPATCH_ALL(a->dt, T, {
    hb_first_diff_kernel<T><<<blocks, 256, 0, g_hb.stream>>>(p, (const T *)hb_data(a), (const T *)hb_data(b), n, first);
    HB_LAUNCHED();
    hb_cmp_at_kernel<T><<<1, 1, 0, g_hb.stream>>>(p, (const T *)hb_data(a), (const T *)hb_data(b), first, op, res);
    HB_LAUNCHED();
  });
  int h = 0;
  HB_CUDA(cudaMemcpyAsync(&h, res, sizeof(int), cudaMemcpyDeviceToHost, g_hb.stream));
  HB_CUDA(cudaStreamSynchronize(g_hb.stream));
  *result = h;
  return HB_OK;
}

// ------------------------------------------------------- relu / sgd / adam
// reluS (CommonShapedOps.hs:38-44): oneIfGtZero * v with
// oneIfGtZero = ifH (x <=. 0) 0 1.
extern "C" int hb_relu(const hb_array *x, hb_array **out) {
  hb_prof_scope prof_("relu", x);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(x && out, "null argument");
  if (hb_lazy_on())
    return hb_lrec(HB_L_GENERIC, "relu").in(x).out(out, x->dt, x->rank, x->shape).commit(
        [](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_relu(in[0], &o[0]); });
  hb_array *o;
  HB_TRY(hb_new_array(x->dt, x->rank, x->shape, &o));
  int st = HB_OK;
  HB_DISPATCH_FLOAT(x->dt, T, st = (map1_into<T, T>(o, x, [] __device__(T v) {
                      return (v <= (T)0 ? (T)0 : (T)1) * v;
                    })));
  if (st != HB_OK) {
    hb_release(o);
    return st;
  }
  *out = o;
  return HB_OK;
}

template <typename T>
__global__ void hb_sgd_kernel(T *p, const T *g, T gamma, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t s = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += s) p[i] = p[i] - gamma * g[i];  // OptimizerTools.hs:33-36
}

extern "C" int hb_sgd_step(hb_array *p, const hb_array *g, double gamma) {
  hb_prof_scope prof_("sgd_step", p);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(p && g, "null argument");
  HB_REQUIRE(p->dt == g->dt && hb_same_shape(p, g), "sgd: parameter/gradient mismatch");
  HB_REQUIRE(hb_contig(p), "sgd: parameters must be contiguous");
  if (hb_lazy_on())
    return hb_lrec(HB_L_EFFECT, "sgd_step").in(p).in(g).farg(0, gamma).commit(
        [](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_sgd_step(in[0], in[1], n.fa[0]); });
  int64_t n = hb_numel(p);
  if (n == 0) return HB_OK;
  const hb_array *gc = g;
  hb_array *tmp = nullptr;
  if (!hb_contig(g)) {
    HB_TRY(hb_contiguous(g, &tmp));
    gc = tmp;
  }
  int blocks = hb_blocks_for(n, 256, g_hb.sm_count * 16);
  int st = HB_OK;
  switch (p->dt) {
    case HB_F64: hb_sgd_kernel<double><<<blocks, 256, 0, g_hb.stream>>>((double *)hb_data(p), (const double *)hb_data(gc), gamma, n); break;
    case HB_F32: hb_sgd_kernel<float><<<blocks, 256, 0, g_hb.stream>>>((float *)hb_data(p), (const float *)hb_data(gc), (float)gamma, n); break;
    default: hb_set_error("sgd: floating dtype required"); st = HB_ERR_INVALID;
  }
  if (tmp) hb_release(tmp);
  if (st != HB_OK) return st;
  HB_LAUNCHED();
  return HB_OK;
}

// updateWithGradientAdam (OptimizerTools.hs:134-232):
//   alphat = alpha * sqrt(1 - beta2^t) / (1 - beta1^t)
//   m' = beta1*m + (1-beta1)*g ; v' = beta2*v + (1-beta2)*g*g
//   p' = p - alphat * m' / (sqrt v' + epsilon)
template <typename T>
__global__ void hb_adam_kernel(T *p, T *m, T *v, const T *g, T alphat, T b1, T b2, T eps, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t s = (int64_t)gridDim.x * blockDim.x;
  const T one = (T)1;
  for (; i < n; i += s) {
    T gi = g[i];
    T mn = b1 * m[i] + (one - b1) * gi;
    T vn = b2 * v[i] + (one - b2) * (gi * gi);
    m[i] = mn;
    v[i] = vn;
    p[i] = p[i] - alphat * mn / (sqrt(vn) + eps);
  }
}

// 16-byte accesses, two independent vectors per thread (seven streams:
// p, m, v read + written, g read); same per-element arithmetic as above.
template <typename T, int V>
__global__ void __launch_bounds__(256)
    hb_adam_vec_kernel(T *__restrict__ p, T *__restrict__ m, T *__restrict__ v, const T *__restrict__ g, T alphat,
                       T b1, T b2, T eps, int64_t nvec) {
  typedef hb_vec<T, V> VT;
  VT *pv = reinterpret_cast<VT *>(p), *mv = reinterpret_cast<VT *>(m), *vv = reinterpret_cast<VT *>(v);
  const VT *gv = reinterpret_cast<const VT *>(g);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t s = (int64_t)gridDim.x * blockDim.x;
  const T one = (T)1;
  auto upd = [&](VT &pp, VT &mm, VT &vw, const VT &gg) {
#pragma unroll
    for (int k = 0; k < V; k++) {
      T gi = gg.v[k];
      T mn = b1 * mm.v[k] + (one - b1) * gi;
      T vn = b2 * vw.v[k] + (one - b2) * (gi * gi);
      mm.v[k] = mn;
      vw.v[k] = vn;
      pp.v[k] = pp.v[k] - alphat * mn / (sqrt(vn) + eps);
    }
  };
  for (; i + s < nvec; i += 2 * s) {
    VT g0 = gv[i], g1 = gv[i + s], m0 = mv[i], m1 = mv[i + s], v0 = vv[i], v1 = vv[i + s], p0 = pv[i], p1 = pv[i + s];
    upd(p0, m0, v0, g0);
    upd(p1, m1, v1, g1);
    mv[i] = m0; vv[i] = v0; pv[i] = p0;
    mv[i + s] = m1; vv[i + s] = v1; pv[i + s] = p1;
  }
  for (; i < nvec; i += s) {
    VT g0 = gv[i], m0 = mv[i], v0 = vv[i], p0 = pv[i];
    upd(p0, m0, v0, g0);
    mv[i] = m0; vv[i] = v0; pv[i] = p0;
  }
}

template <typename T>
static void adam_launch(T *p, T *m, T *v, const T *g, T alphat, T b1, T b2, T eps, int64_t n) {
  constexpr int V = 16 / (int)sizeof(T);
  int64_t nvec = 0;
  if (hb_aligned16(p) && hb_aligned16(m) && hb_aligned16(v) && hb_aligned16(g) && n >= 4 * V) {
    nvec = n / V;
    int blocks = hb_blocks_for((nvec + 1) / 2, 256, g_hb.sm_count * 16);
    hb_adam_vec_kernel<T, V><<<blocks, 256, 0, g_hb.stream>>>(p, m, v, g, alphat, b1, b2, eps, nvec);
  }
  int64_t done = nvec * V;
  if (done < n) {
    if (done) g_hb.launches.fetch_add(1, std::memory_order_relaxed);  // second launch of this call
    int blocks = hb_blocks_for(n - done, 256, g_hb.sm_count * 16);
    hb_adam_kernel<T><<<blocks, 256, 0, g_hb.stream>>>(p + done, m + done, v + done, g + done, alphat, b1, b2, eps, n - done);
  }
}

extern "C" int hb_adam_step(hb_array *p, hb_array *m, hb_array *v, const hb_array *g, double alpha,
                            double beta1, double beta2, double epsilon, int64_t t) {
  hb_prof_scope prof_("adam_step", p);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(p && m && v && g, "null argument");
  HB_REQUIRE(p->dt == g->dt && p->dt == m->dt && p->dt == v->dt, "adam: dtype mismatch");
  HB_REQUIRE(hb_same_shape(p, g) && hb_same_shape(p, m) && hb_same_shape(p, v), "adam: shape mismatch");
  HB_REQUIRE(hb_contig(p) && hb_contig(m) && hb_contig(v), "adam: state must be contiguous");
  if (hb_lazy_on())
    return hb_lrec(HB_L_EFFECT, "adam_step").in(p).in(m).in(v).in(g).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_adam_step(in[0], in[1], in[2], in[3], alpha, beta1, beta2, epsilon, t); });
  int64_t n = hb_numel(p);
  if (n == 0) return HB_OK;
  const hb_array *gc = g;
  hb_array *tmp = nullptr;
  if (!hb_contig(g)) {
    HB_TRY(hb_contiguous(g, &tmp));
    gc = tmp;
  }
  double alphat = alpha * sqrt(1.0 - hb_powi_ghc(beta2, t)) / (1.0 - hb_powi_ghc(beta1, t));
  int st = HB_OK;
  switch (p->dt) {
    case HB_F64:
      adam_launch<double>((double *)hb_data(p), (double *)hb_data(m), (double *)hb_data(v), (const double *)hb_data(gc),
                          alphat, beta1, beta2, epsilon, n);
      break;
    case HB_F32:
      adam_launch<float>((float *)hb_data(p), (float *)hb_data(m), (float *)hb_data(v), (const float *)hb_data(gc),
                         (float)alphat, (float)beta1, (float)beta2, (float)epsilon, n);
      break;
    default: hb_set_error("adam: floating dtype required"); st = HB_ERR_INVALID;
  }
  if (tmp) hb_release(tmp);
  if (st != HB_OK) return st;
  HB_LAUNCHED();
  return HB_OK;
}

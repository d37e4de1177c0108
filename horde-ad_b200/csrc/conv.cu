// conv.cu -- im2col-free convolution forward and VJPs.
//
// Reference semantics: conv2dPreservingS (CommonShapedOps.hs:136-157, zero
// padding on the HIGH side only via slicezS :225-237) and the handwritten
// adjoints conv2dPreserving_dInp / _dKrn (TestConvQuickCheck.hs:290-349).
//
// No patch tensor is ever built.  The input is copied once into a zero-padded
// buffer; the im2col operand is then a pure strided VIEW of that buffer
// (overlapping windows, for dInp with negative kernel strides) and the
// contraction runs on the strided batched GEMM engine of contract.cu (FP64
// DMMA), which walks views through per-tile offset tables:
//   fwd : M=(n,h,w)  N=o          K=(c,kh,kw)
//   dK  : M=o        N=(c,kh,kw)  K=(n_in,h,w)  batched over image groups,
//         partials reduced in fixed order (deterministic split-K)
//   dA  : M=(n,h,w)  N=c          K=(o,kh,kw)   on the low-side padded dB
// Algorithmic bytes per pass at ConvVjpBench: 205.8 MB,
// 14.80 GFLOP -> FP64 tensor pipe bound.
#include "common.cuh"
#include "lazy.cuh"

int hb_copy_into(hb_array *dst, const hb_array *src);  // elementwise.cu
// conv_dmma.cu: shared-memory-staged DMMA kernels (contiguous f64 operands)
struct hb_pool_fuse {
  const hb_array *bias;
  hb_array *y, *code;
};
int hb_conv_sp_dmma(const hb_array *K, const hb_array *in, hb_array *out, int flip, const hb_pool_fuse *pf,
                    int *done);
int hb_conv_dk_dmma(const hb_array *A, const hb_array *dB, const hb_array *code, hb_array *out, int KH, int KW,
                    int *done);

// conv_strip.cu: warp-autonomous kernels for the one-input-channel layer
int hb_conv_strip_fwd(const hb_array *K, const hb_array *bias, const hb_array *A, hb_array *y, hb_array *code,
                      int *done);
int hb_conv_strip_dk(const hb_array *A, const hb_array *dy, const hb_array *code, int KH, int KW, hb_array *dK,
                     hb_array *dbias, int *done);

// Runs the DMMA fast path into a fresh [n0,n1,n2,n3] array when it applies.
static int try_sp(const hb_array *K, const hb_array *in, const int64_t *osh, int flip, hb_array **out, int *done) {
  *done = 0;
  if (in->dt != HB_F64 || !hb_contig(in)) return HB_OK;
  hb_array *o;
  HB_TRY(hb_new_array(HB_F64, 4, osh, &o));
  int st = hb_conv_sp_dmma(K, in, o, flip, nullptr, done);
  if (st != HB_OK || !*done) {
    hb_release(o);
    return st;
  }
  *out = o;
  return HB_OK;
}
extern "C" int hb_dot_inner_sum_leading(const hb_array *a, const hb_array *b, int n_axes, hb_array **out);

// zero-padded copy of x[N,C,H,W]: lo rows/cols before, hi after
static int pad2d(const hb_array *x, int64_t lo_h, int64_t hi_h, int64_t lo_w, int64_t hi_w, hb_array **out) {
  int64_t sh[4] = {x->shape[0], x->shape[1], x->shape[2] + lo_h + hi_h, x->shape[3] + lo_w + hi_w};
  hb_array *p;
  if (lo_h + hi_h + lo_w + hi_w == 0 && hb_contig(x)) {
    hb_retain((hb_array *)x);
    *out = (hb_array *)x;
    return HB_OK;
  }
  HB_TRY(hb_zeros(x->dt, 4, sh, &p));
  hb_array *v = hb_new_view(p);
  v->shape[2] = x->shape[2];
  v->shape[3] = x->shape[3];
  v->offset += lo_h * p->strides[2] + lo_w * p->strides[3];
  int st = hb_copy_into(v, x);
  hb_release(v);
  if (st != HB_OK) {
    hb_release(p);
    return st;
  }
  *out = p;
  return HB_OK;
}

static hb_array *view6(const hb_array *base, int64_t off, const int64_t *shape, const int64_t *strides, int rank) {
  hb_array *v = hb_new_view(base);
  v->rank = rank;
  v->offset += off;
  for (int i = 0; i < rank; i++) {
    v->shape[i] = shape[i];
    v->strides[i] = strides[i];
  }
  return v;
}

// ---- Float: channels-last staging for the tcgen05 engine (contract_tf32.cu) --------------------------------------
// The strided-view engine fetches 16 bytes per load when a view is contiguous and 16-byte aligned along the reduced
// axis (or along its rows).  In NCHW the im2col window moves by ONE float per tap, which breaks the alignment, and
// the reduced axis (c, kh, kw) has no run of four.  So the Float path copies the activations once into a zero-padded
// CHANNELS-LAST buffer [N, H + pad, W + pad, C] (one pass over the input, ~2 x its bytes) and the weights into
// [Co, KH, KW, Ci]: the reduced axis then ends in the channel run, every window starts on a multiple of C floats,
// and all three contractions run on 128-bit fetches.  Still no patch tensor.  Needs C % 4 == 0 (else: NCHW views).
// x[N,C,H,W] (contiguous) -> the interior of dst[N,Hp,Wp,C]: 32 x 32 tiles through shared memory, reads coalesced
// along the positions, writes coalesced along the channels
template <typename T>
__global__ void __launch_bounds__(256)
    hb_nchw_to_nhwc_kernel(const T *__restrict__ src, T *__restrict__ dst, int C, int H, int W, int Wp, int64_t dimg,
                           int lo_h, int lo_w) {
  __shared__ T tile[32][33];
  const int HW = H * W;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int64_t n = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const T *s = src + n * (int64_t)C * HW;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int c = c0 + ty + 8 * j, p = p0 + tx;
    tile[ty + 8 * j][tx] = (c < C && p < HW) ? s[(int64_t)c * HW + p] : T(0);
  }
  __syncthreads();
  T *d = dst + n * dimg;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int p = p0 + ty + 8 * j, c = c0 + tx;
    if (p < HW && c < C) {
      const int h = p / W, w = p - h * W;
      d[((int64_t)(h + lo_h) * Wp + (w + lo_w)) * C + c] = tile[tx][ty + 8 * j];
    }
  }
}

static int pad2d_nhwc(const hb_array *x, int64_t lo_h, int64_t hi_h, int64_t lo_w, int64_t hi_w, hb_array **out) {
  int64_t N = x->shape[0], C = x->shape[1], H = x->shape[2], W = x->shape[3];
  int64_t sh[4] = {N, H + lo_h + hi_h, W + lo_w + hi_w, C};
  hb_array *p;
  if (lo_h + hi_h + lo_w + hi_w == 0) HB_TRY(hb_new_array(x->dt, 4, sh, &p));
  else HB_TRY(hb_zeros(x->dt, 4, sh, &p));
  if (x->dt == HB_F32 && hb_contig(x) && N > 0 && N < 65536 && C > 0 && H * W > 0 && (C + 31) / 32 < 65536) {
    dim3 grid((unsigned)((H * W + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)N);
    hb_nchw_to_nhwc_kernel<float><<<grid, 256, 0, g_hb.stream>>>((const float *)hb_data(x), (float *)hb_data(p), (int)C,
                                                                (int)H, (int)W, (int)sh[2], p->strides[0], (int)lo_h,
                                                                (int)lo_w);
    HB_LAUNCHED();
    *out = p;
    return HB_OK;
  }
  hb_array *v = hb_new_view(p);   // the interior, indexed as [N, C, H, W]
  v->shape[0] = N; v->strides[0] = p->strides[0];
  v->shape[1] = C; v->strides[1] = 1;
  v->shape[2] = H; v->strides[2] = p->strides[1];
  v->shape[3] = W; v->strides[3] = p->strides[2];
  v->offset += lo_h * p->strides[1] + lo_w * p->strides[2];
  int st = hb_copy_into(v, x);
  hb_release(v);
  if (st != HB_OK) { hb_release(p); return st; }
  *out = p;
  return HB_OK;
}
// K[o,c,kh,kw] as a contiguous [d0, KH, KW, d3] array with (d0, d3) = (o, c) or (c, o)
static int weights_last(const hb_array *K, bool out_first, hb_array **out) {
  hb_array *v = hb_new_view(K);
  int a0 = out_first ? 0 : 1, a3 = out_first ? 1 : 0;
  v->shape[0] = K->shape[a0]; v->strides[0] = K->strides[a0];
  v->shape[1] = K->shape[2];  v->strides[1] = K->strides[2];
  v->shape[2] = K->shape[3];  v->strides[2] = K->strides[3];
  v->shape[3] = K->shape[a3]; v->strides[3] = K->strides[a3];
  int st = hb_contiguous(v, out);
  hb_release(v);
  return st;
}
static bool f32_channels_last(const hb_array *x, int64_t c_reduced) {
  static const bool off = getenv("HB_CONV_F32_NCHW") != nullptr;   // measurement switch: the NCHW views
  return !off && x->dt == HB_F32 && c_reduced >= 4 && c_reduced % 4 == 0;
}

extern "C" int hb_conv2d_preserving(const hb_array *K, const hb_array *A, hb_array **out) {
  hb_prof_scope prof_("conv2d_preserving", K, A);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(K && A && out, "null argument");
  HB_REQUIRE(K->rank == 4 && A->rank == 4 && K->dt == A->dt && hb_is_float(A->dt),
             "conv2dPreserving: rank-4 floating operands required");
  HB_REQUIRE(K->shape[1] == A->shape[1], "conv2dPreserving: kernel %s vs input %s", hb_shape_str(K).c_str(),
             hb_shape_str(A).c_str());
  int64_t N = A->shape[0], Ci = A->shape[1], H = A->shape[2], W = A->shape[3];
  int64_t Co = K->shape[0], KH = K->shape[2], KW = K->shape[3];
  if (hb_lazy_on()) {
    int64_t sh_[4] = {N, Co, H, W};
    return hb_lrec(HB_L_CONV_FWD, "conv2d_preserving").in(K).in(A).out(out, A->dt, 4, sh_).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_conv2d_preserving(in[0], in[1], &o[0]); });
  }
  if (N * Co * H * W == 0 || Ci * KH * KW == 0) {
    int64_t sh[4] = {N, Co, H, W};
    return hb_zeros(A->dt, 4, sh, out);
  }
  {
    int64_t osh[4] = {N, Co, H, W};
    int done = 0;
    HB_TRY(try_sp(K, A, osh, 0, out, &done));
    if (done) return HB_OK;
  }
  if (f32_channels_last(A, Ci)) {
    hb_array *Ap, *Kp;
    HB_TRY(pad2d_nhwc(A, 0, KH - 1, 0, KW - 1, &Ap));
    int st = weights_last(K, true, &Kp);
    if (st != HB_OK) { hb_release(Ap); return st; }
    // leading reduced (kh, kw), kept (n, o, h, w), last reduced c
    int64_t shp[7] = {KH, KW, N, Co, H, W, Ci};
    int64_t sa[7] = {Ap->strides[1], Ap->strides[2], Ap->strides[0], 0, Ap->strides[1], Ap->strides[2], 1};
    int64_t sk[7] = {Kp->strides[1], Kp->strides[2], 0, Kp->strides[0], 0, 0, 1};
    hb_array *va = view6(Ap, 0, shp, sa, 7);
    hb_array *vk = view6(Kp, 0, shp, sk, 7);
    st = hb_dot_inner_sum_leading(va, vk, 2, out);
    hb_release(va); hb_release(vk); hb_release(Ap); hb_release(Kp);
    return st;
  }
  hb_array *Ap;
  HB_TRY(pad2d(A, 0, KH - 1, 0, KW - 1, &Ap));
  // operands over the index space [kept: n, o, h, w | reduced(last... ) ]
  // layout for hb_dot_inner_sum_leading: leading reduced axes (c, kh), kept
  // (n, o, h, w), last reduced axis kw.
  int64_t shp[7] = {Ci, KH, N, Co, H, W, KW};
  int64_t sa[7] = {Ap->strides[1], Ap->strides[2], Ap->strides[0], 0, Ap->strides[2], Ap->strides[3], Ap->strides[3]};
  int64_t sk[7] = {K->strides[1], K->strides[2], 0, K->strides[0], 0, 0, K->strides[3]};
  hb_array *va = view6(Ap, 0, shp, sa, 7);
  hb_array *vk = view6(K, 0, shp, sk, 7);
  int st = hb_dot_inner_sum_leading(va, vk, 2, out);
  hb_release(va);
  hb_release(vk);
  hb_release(Ap);
  return st;
}

extern "C" int hb_conv2d_preserving_dkrn(const hb_array *A, const hb_array *dB, int KH, int KW,
                                         hb_array **out) {
  hb_prof_scope prof_("conv2d_preserving_dkrn", A, dB);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(A && dB && out, "null argument");
  HB_REQUIRE(A->rank == 4 && dB->rank == 4 && A->dt == dB->dt && hb_is_float(A->dt),
             "conv2dPreserving_dKrn: rank-4 floating operands required");
  HB_REQUIRE(A->shape[0] == dB->shape[0] && A->shape[2] == dB->shape[2] && A->shape[3] == dB->shape[3],
             "conv2dPreserving_dKrn: input %s vs cotangent %s", hb_shape_str(A).c_str(), hb_shape_str(dB).c_str());
  HB_REQUIRE(KH >= 0 && KW >= 0, "conv2dPreserving_dKrn: bad kernel size");
  int64_t N = A->shape[0], Ci = A->shape[1], H = A->shape[2], W = A->shape[3], Co = dB->shape[1];
  int64_t ksh[4] = {Co, Ci, KH, KW};
  if (hb_lazy_on())
    return hb_lrec(HB_L_CONV_DK, "conv2d_dkrn").in(A).in(dB).iarg(0, KH).iarg(1, KW).out(out, A->dt, 4, ksh).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_conv2d_preserving_dkrn(in[0], in[1], (int)n.ia[0], (int)n.ia[1], &o[0]); });
  if (N * H * W == 0 || Co * Ci * KH * KW == 0) return hb_zeros(A->dt, 4, ksh, out);
  if (A->dt == HB_F64 && hb_contig(A) && hb_contig(dB)) {
    hb_array *o;
    int done = 0;
    HB_TRY(hb_new_array(HB_F64, 4, ksh, &o));
    int st = hb_conv_dk_dmma(A, dB, nullptr, o, KH, KW, &done);
    if (st == HB_OK && done) {
      *out = o;
      return HB_OK;
    }
    hb_release(o);
    if (st != HB_OK) return st;
  }
  if (f32_channels_last(A, Ci) && Co % 4 == 0) {
    hb_array *Ap, *Bt;
    HB_TRY(pad2d_nhwc(A, 0, KH - 1, 0, KW - 1, &Ap));
    int st = pad2d_nhwc(dB, 0, 0, 0, 0, &Bt);
    if (st != HB_OK) { hb_release(Ap); return st; }
    // leading reduced (n, h), kept (o, kh, kw, c), last reduced w; the engine splits K itself (fixed-order reduce)
    int64_t shp[7] = {N, H, Co, KH, KW, Ci, W};
    int64_t sb[7] = {Bt->strides[0], Bt->strides[1], 1, 0, 0, 0, Bt->strides[2]};
    int64_t sa[7] = {Ap->strides[0], Ap->strides[1], 0, Ap->strides[1], Ap->strides[2], 1, Ap->strides[2]};
    hb_array *vb = view6(Bt, 0, shp, sb, 7);
    hb_array *va = view6(Ap, 0, shp, sa, 7);
    hb_array *part = nullptr;
    st = hb_dot_inner_sum_leading(vb, va, 2, &part);   // [Co, KH, KW, Ci]
    hb_release(vb); hb_release(va); hb_release(Ap); hb_release(Bt);
    if (st != HB_OK) return st;
    hb_array *v = hb_new_view(part);                   // -> [Co, Ci, KH, KW]
    v->shape[1] = Ci; v->strides[1] = part->strides[3];
    v->shape[2] = KH; v->strides[2] = part->strides[1];
    v->shape[3] = KW; v->strides[3] = part->strides[2];
    st = hb_contiguous(v, out);
    hb_release(v);
    hb_release(part);
    return st;
  }
  hb_array *Ap;
  HB_TRY(pad2d(A, 0, KH - 1, 0, KW - 1, &Ap));
  // deterministic split-K over image groups: G partial kernels, summed in order
  int64_t tiles = ((Co + 63) / 64) * ((Ci * KH * KW + 63) / 64);
  int64_t G = 1;
  while (G * 2 <= N && N % (G * 2) == 0 && tiles * G < 2 * (int64_t)g_hb.sm_count) G *= 2;
  int64_t Nin = N / G;
  // index space: leading reduced (n_in, h), kept (g, o, c, kh, kw), last reduced w
  int64_t shp[8] = {Nin, H, G, Co, Ci, KH, KW, W};
  int64_t sa[8] = {Ap->strides[0], Ap->strides[2], Ap->strides[0] * Nin, 0, Ap->strides[1], Ap->strides[2], Ap->strides[3], Ap->strides[3]};
  int64_t sb[8] = {dB->strides[0], dB->strides[2], dB->strides[0] * Nin, dB->strides[1], 0, 0, 0, dB->strides[3]};
  hb_array *va = view6(Ap, 0, shp, sa, 8);
  hb_array *vb = view6(dB, 0, shp, sb, 8);
  hb_array *part = nullptr;
  // a := dB view (M = o), b := A view (N = c,kh,kw)
  int st = hb_dot_inner_sum_leading(vb, va, 2, &part);
  hb_release(va);
  hb_release(vb);
  hb_release(Ap);
  if (st != HB_OK) return st;
  if (G == 1) {
    int64_t zero = 0;
    st = hb_index_prefix(part, 1, &zero, out);
  } else {
    st = hb_sum_leading(part, 1, out);
  }
  hb_release(part);
  return st;
}

extern "C" int hb_conv2d_preserving_dinp(const hb_array *K, const hb_array *dB, hb_array **out) {
  hb_prof_scope prof_("conv2d_preserving_dinp", K, dB);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(K && dB && out, "null argument");
  HB_REQUIRE(K->rank == 4 && dB->rank == 4 && K->dt == dB->dt && hb_is_float(K->dt),
             "conv2dPreserving_dInp: rank-4 floating operands required");
  HB_REQUIRE(K->shape[0] == dB->shape[1], "conv2dPreserving_dInp: kernel %s vs cotangent %s",
             hb_shape_str(K).c_str(), hb_shape_str(dB).c_str());
  int64_t N = dB->shape[0], Co = dB->shape[1], H = dB->shape[2], W = dB->shape[3];
  int64_t Ci = K->shape[1], KH = K->shape[2], KW = K->shape[3];
  int64_t ash[4] = {N, Ci, H, W};
  if (hb_lazy_on())
    return hb_lrec(HB_L_CONV_DI, "conv2d_dinp").in(K).in(dB).out(out, K->dt, 4, ash).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_conv2d_preserving_dinp(in[0], in[1], &o[0]); });
  if (N * Ci * H * W == 0 || Co * KH * KW == 0) return hb_zeros(K->dt, 4, ash, out);
  {
    int done = 0;
    HB_TRY(try_sp(K, dB, ash, 1, out, &done));
    if (done) return HB_OK;
  }
  if (f32_channels_last(dB, Co)) {
    hb_array *Bp, *Kq;
    HB_TRY(pad2d_nhwc(dB, KH - 1, 0, KW - 1, 0, &Bp));
    int st = weights_last(K, false, &Kq);              // [Ci, KH, KW, Co]
    if (st != HB_OK) { hb_release(Bp); return st; }
    // dA[n,c,h,w] = sum_{kh,kw,o} Bp[n, h+(KH-1)-kh, w+(KW-1)-kw, o] Kq[c,kh,kw,o]
    int64_t shp[7] = {KH, KW, N, Ci, H, W, Co};
    int64_t sb[7] = {-Bp->strides[1], -Bp->strides[2], Bp->strides[0], 0, Bp->strides[1], Bp->strides[2], 1};
    int64_t sk[7] = {Kq->strides[1], Kq->strides[2], 0, Kq->strides[0], 0, 0, 1};
    int64_t off = (KH - 1) * Bp->strides[1] + (KW - 1) * Bp->strides[2];
    hb_array *vb = view6(Bp, off, shp, sb, 7);
    hb_array *vk = view6(Kq, 0, shp, sk, 7);
    st = hb_dot_inner_sum_leading(vb, vk, 2, out);
    hb_release(vb); hb_release(vk); hb_release(Bp); hb_release(Kq);
    return st;
  }
  hb_array *Bp;
  HB_TRY(pad2d(dB, KH - 1, 0, KW - 1, 0, &Bp));
  // dA[n,c,h,w] = sum_{o,kh,kw} Bp[n,o,h+(KH-1)-kh, w+(KW-1)-kw] K[o,c,kh,kw]
  // index space: leading reduced (o, kh), kept (n, c, h, w), last reduced kw
  int64_t shp[7] = {Co, KH, N, Ci, H, W, KW};
  int64_t sb[7] = {Bp->strides[1], -Bp->strides[2], Bp->strides[0], 0, Bp->strides[2], Bp->strides[3], -Bp->strides[3]};
  int64_t sk[7] = {K->strides[0], K->strides[2], 0, K->strides[1], 0, 0, K->strides[3]};
  int64_t off = (KH - 1) * Bp->strides[2] + (KW - 1) * Bp->strides[3];
  hb_array *vb = view6(Bp, off, shp, sb, 7);
  hb_array *vk = view6(K, 0, shp, sk, 7);
  int st = hb_dot_inner_sum_leading(vb, vk, 2, out);
  hb_release(vb);
  hb_release(vk);
  hb_release(Bp);
  return st;
}

// ------------------------------------------------ a whole convMnistLayerS
extern "C" int hb_relu_pool_fwd(const hb_array *pre, const hb_array *bias, int ksize, hb_array **out,
                                hb_array **code_out);
extern "C" int hb_relu_pool_bwd(const hb_array *dy, const hb_array *code, int ksize, int64_t H, int64_t W,
                                hb_array **dpre_out, hb_array **dbias_out);

// maxPool k k (relu (conv2dPreserving K A + bias)) (MnistCnnShaped2.hs:42-62).  For
// 2x2/2 windows on contiguous f64 operands the bias add, relu and pooling run in
// the epilogue of the DMMA convolution kernel and the pre-activation never
// reaches HBM; otherwise the two kernels run back to back.
extern "C" int hb_conv_relu_pool_fwd(const hb_array *K, const hb_array *bias, const hb_array *A, int ksize,
                                     hb_array **out, hb_array **code_out) {
  hb_prof_scope prof_("conv_relu_pool_fwd", K, A);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(K && A && out && code_out, "null argument");
  HB_REQUIRE(K->rank == 4 && A->rank == 4 && K->dt == A->dt && hb_is_float(A->dt) && K->shape[1] == A->shape[1],
             "conv_relu_pool: kernel %s vs input %s", hb_shape_str(K).c_str(), hb_shape_str(A).c_str());
  HB_REQUIRE(!bias || (bias->rank == 1 && bias->dt == A->dt && bias->shape[0] == K->shape[0]),
             "conv_relu_pool: bias must be [%lld]", (long long)K->shape[0]);
  HB_REQUIRE(ksize >= 1 && ksize <= 11, "conv_relu_pool: window must be 1..11");
  int64_t N = A->shape[0], Co = K->shape[0], H = A->shape[2], W = A->shape[3];
  if (hb_lazy_on()) {
    int64_t sh_[4] = {N, Co, H / ksize, W / ksize};
    const bool hb_ = bias != nullptr;
    hb_lrec r(HB_L_CRP_FWD, "conv_relu_pool_fwd");
    r.in(K);
    if (hb_) r.in(bias);
    return r.in(A).iarg(0, ksize).iarg(1, hb_).out(out, A->dt, 4, sh_).out(code_out, HB_I8, 4, sh_).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) {
          const bool b_ = n.ia[1] != 0;
          return hb_conv_relu_pool_fwd(in[0], b_ ? in[1] : nullptr, in[b_ ? 2 : 1], (int)n.ia[0], &o[0], &o[1]);
        });
  }
  if (ksize == 2 && A->dt == HB_F64 && N * Co * (H / 2) * (W / 2) > 0 && K->shape[1] * K->shape[2] * K->shape[3] > 0) {
    int64_t osh[4] = {N, Co, H / 2, W / 2};
    hb_pool_fuse pf;
    pf.bias = bias;
    pf.y = pf.code = nullptr;
    int st = hb_new_array(HB_F64, 4, osh, &pf.y);
    if (st == HB_OK) st = hb_new_array(HB_I8, 4, osh, &pf.code);
    int done = 0;
    if (st == HB_OK) st = hb_conv_strip_fwd(K, bias, A, pf.y, pf.code, &done);
    if (st == HB_OK && !done) st = hb_conv_sp_dmma(K, A, nullptr, 0, &pf, &done);
    if (st == HB_OK && done) {
      *out = pf.y;
      *code_out = pf.code;
      return HB_OK;
    }
    hb_release(pf.y);
    hb_release(pf.code);
    if (st != HB_OK) return st;
  }
  hb_array *pre;
  HB_TRY(hb_conv2d_preserving(K, A, &pre));
  int st = hb_relu_pool_fwd(pre, bias, ksize, out, code_out);
  hb_release(pre);
  return st;
}

// Adjoint of hb_conv_relu_pool_fwd: dK, dbias and (optionally) dA from the
// pooled cotangent dy and the code bytes.  When dA is not wanted (first layer:
// the glyphs are not differentiated) the full-size cotangent is never
// materialised: the dKrn kernel expands (dy, code) inside shared memory.
extern "C" int hb_conv_relu_pool_bwd(const hb_array *K, const hb_array *A, const hb_array *dy, const hb_array *code,
                                     int ksize, hb_array **dK_out, hb_array **dbias_out, hb_array **dA_out) {
  hb_prof_scope prof_("conv_relu_pool_bwd", A, dy);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(K && A && dy && code && dK_out, "null argument");
  HB_REQUIRE(K->rank == 4 && A->rank == 4 && dy->rank == 4 && K->dt == A->dt && dy->dt == A->dt && hb_is_float(A->dt),
             "conv_relu_pool_bwd: bad operands");
  int64_t H = A->shape[2], W = A->shape[3];
  int KH = (int)K->shape[2], KW = (int)K->shape[3];
  if (hb_lazy_on()) {
    int64_t co_ = K->shape[0];
    return hb_lrec(HB_L_CRP_BWD, "conv_relu_pool_bwd").in(K).in(A).in(dy).in(code).iarg(0, ksize)
        .out(dK_out, K->dt, 4, K->shape).out(dbias_out, K->dt, 1, &co_).out(dA_out, A->dt, 4, A->shape)
        .commit([=](hb_lnode &n, hb_array *const *in, hb_array **o) {
          return hb_conv_relu_pool_bwd(in[0], in[1], in[2], in[3], (int)n.ia[0], &o[0], n.out[1] ? &o[1] : nullptr,
                                       n.out[2] ? &o[2] : nullptr);
        });
  }
  int st = HB_OK;
  if (!dA_out && ksize == 2 && A->dt == HB_F64 && hb_contig(A) && hb_contig(dy) && hb_numel(dy) > 0 &&
      K->shape[1] * KH * KW > 0) {
    hb_array *dk = nullptr;
    int done = 0;
    int64_t ksh[4] = {K->shape[0], K->shape[1], KH, KW};
    HB_TRY(hb_new_array(HB_F64, 4, ksh, &dk));
    {
      // one input channel: dK and the bias cotangent out of one warp-autonomous kernel
      hb_array *db = nullptr;
      if (dbias_out) st = hb_new_array(HB_F64, 1, &K->shape[0], &db);
      if (st == HB_OK) st = hb_conv_strip_dk(A, dy, code, KH, KW, dk, db, &done);
      if (st == HB_OK && done) {
        *dK_out = dk;
        if (dbias_out) *dbias_out = db;
        return HB_OK;
      }
      hb_release(db);
      if (st != HB_OK) { hb_release(dk); return st; }
    }
    st = hb_conv_dk_dmma(A, dy, code, dk, KH, KW, &done);
    if (st == HB_OK && done) {
      if (dbias_out) st = hb_relu_pool_bwd(dy, code, ksize, H, W, nullptr, dbias_out);
      if (st != HB_OK) { hb_release(dk); return st; }
      *dK_out = dk;
      return HB_OK;
    }
    hb_release(dk);
    if (st != HB_OK) return st;
  }
  hb_array *dpre = nullptr, *dk = nullptr, *da = nullptr, *db = nullptr;
  HB_TRY(hb_relu_pool_bwd(dy, code, ksize, H, W, &dpre, dbias_out ? &db : nullptr));
  st = hb_conv2d_preserving_dkrn(A, dpre, KH, KW, &dk);
  if (st == HB_OK && dA_out) st = hb_conv2d_preserving_dinp(K, dpre, &da);
  hb_release(dpre);
  if (st != HB_OK) { hb_release(dk); hb_release(da); hb_release(db); return st; }
  *dK_out = dk;
  if (dbias_out) *dbias_out = db;
  if (dA_out) *dA_out = da;
  return HB_OK;
}

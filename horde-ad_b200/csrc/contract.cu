// contract.cu -- dense contractions: tsdot1In, tsmatmul2, tsmatvecmul
// (OpsConcrete.hs:234-242) and the fused `ssumN . sdot1In` contraction the
// simplified artifacts contain (TestMnistPP.hs:735: `ssumN @[3,4] (sdot1In ..)`).
//
// The reference expresses matmul -- and, after contraction, convolution -- as
// sdot1In over stride-0 (sreplicate) and permuted (stranspose) VIEWS, which are
// free on the CPU.  Here the views are never materialised either: the axes of
// the two equal-shape operands are classified from their strides into
//   M axes (b has stride 0), N axes (a has stride 0), batch axes (both move),
//   K axes (the reduced ones),
// and one strided batched GEMM engine walks the operands in place through
// per-tile offset tables in shared memory (overlapping im2col views from
// affine gathers included).
//
// Engine: 64x64 output tile, K chunk 16, 256 threads, FP64 DMMA (mma.sync
// m8n8k4 -> SASS DMMA.8x8x4, the only FP64 tensor shape on sm_100a; tcgen05
// has no FP64 kind) for doubles, FFMA register tiles for floats.  Tensor-pipe
// roofline: 2*M*N*K flops against the measured FP64 peak (hb_microbench).
#include "common.cuh"
#include "contract.cuh"
#include "lazy.cuh"

#define TM 64
#define TN 64
#define TK 16

// ---- FP64 tensor-core tile product: warp computes 32(m) x 16(n) of the tile
__device__ __forceinline__ void hb_dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

#define AS_LD (TM + 8)  // k-major tiles; row stride == 8 (mod 16) doubles: conflict-free fragments
#define BS_LD (TN + 8)

template <typename T, bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__(256)
    hb_contract_kernel(const __grid_constant__ hb_contractp P, const T *__restrict__ A,
                       const T *__restrict__ Bm, T *__restrict__ C) {
  __shared__ T As[2][TK][AS_LD];
  __shared__ T Bs[2][TK][BS_LD];
  __shared__ int64_t offAm[TM], offBn[TN], offCm[TM], offCn[TN];
  __shared__ int64_t offAk[2][TK], offBk[2][TK];

  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * TM;
  const int64_t n0 = (int64_t)blockIdx.y * TN;
  int64_t ba, bb, bc;
  hb_cg_offsets(P.B, blockIdx.z, ba, bb, bc);
  if (tid < TM) {
    int64_t m = m0 + tid, oa = 0, ob = 0, oc = 0;
    if (m < P.M.n) hb_cg_offsets(P.M, m, oa, ob, oc);
    offAm[tid] = oa;
    offCm[tid] = oc;
  } else if (tid < TM + TN) {
    int t = tid - TM;
    int64_t n = n0 + t, oa = 0, ob = 0, oc = 0;
    if (n < P.N.n) hb_cg_offsets(P.N, n, oa, ob, oc);
    offBn[t] = ob;
    offCn[t] = oc;
  }
  const int64_t Ktot = P.K.n;
  const int nk = (int)((Ktot + TK - 1) / TK);

  auto load_koffs = [&](int buf, int kc) {
    if (tid < TK) {
      int64_t k = (int64_t)kc * TK + tid, oa = 0, ob = 0, oc = 0;
      if (k < Ktot) hb_cg_offsets(P.K, k, oa, ob, oc);
      offAk[buf][tid] = oa;
      offBk[buf][tid] = ob;
    }
  };
  // each thread stages 4 elements of A and 4 of B per chunk
  T ra[4], rb[4];
  auto fetch = [&](int buf, int kc) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int e = tid + i * 256;
      int am, ak, bn, bk;
      if (A_KFAST) { ak = e % TK; am = e / TK; } else { am = e % TM; ak = e / TM; }
      if (B_KFAST) { bk = e % TK; bn = e / TK; } else { bn = e % TN; bk = e / TN; }
      int64_t kA = (int64_t)kc * TK + ak, kB = (int64_t)kc * TK + bk;
      ra[i] = (m0 + am < P.M.n && kA < Ktot) ? A[ba + offAm[am] + offAk[buf][ak]] : T(0);
      rb[i] = (n0 + bn < P.N.n && kB < Ktot) ? Bm[bb + offBn[bn] + offBk[buf][bk]] : T(0);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int e = tid + i * 256;
      int am, ak, bn, bk;
      if (A_KFAST) { ak = e % TK; am = e / TK; } else { am = e % TM; ak = e / TM; }
      if (B_KFAST) { bk = e % TK; bn = e / TK; } else { bn = e % TN; bk = e / TN; }
      As[buf][ak][am] = ra[i];
      Bs[buf][bk][bn] = rb[i];
    }
  };

  load_koffs(0, 0);
  __syncthreads();
  fetch(0, 0);
  stash(0);
  if (nk > 1) load_koffs(1, 1);
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  if constexpr (std::is_same<T, double>::value) {
    // 8 warps: warp (wm, wn) = (warp % 2, warp / 2) owns rows [32*wm, +32) x cols [16*wn, +16)
    const int wm = (warp & 1) * 32, wn = (warp >> 1) * 16;
    const int fr = lane >> 2, fk = lane & 3;  // fragment row/col (A: row fr, k fk; B: k fk, col fr)
    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kc = 0; kc < nk; kc++) {
      int buf = kc & 1;
      if (kc + 1 < nk) fetch(buf ^ 1, kc + 1);
#pragma unroll
      for (int k4 = 0; k4 < TK; k4 += 4) {
        double af[4], bf[2];
#pragma unroll
        for (int i = 0; i < 4; i++) af[i] = As[buf][k4 + fk][wm + 8 * i + fr];
#pragma unroll
        for (int j = 0; j < 2; j++) bf[j] = Bs[buf][k4 + fk][wn + 8 * j + fr];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 2; j++) hb_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
      if (kc + 1 < nk) {
        stash(buf ^ 1);
        if (kc + 2 < nk) load_koffs(buf, kc + 2);
      }
      __syncthreads();
    }
    // C fragment: row = lane/4, cols = (lane%4)*2 + {0,1}
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 2; j++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          int ml = wm + 8 * i + fr, nl = wn + 8 * j + fk * 2 + e;
          if (m0 + ml < P.M.n && n0 + nl < P.N.n) C[bc + offCm[ml] + offCn[nl]] = acc[i][j][e];
        }
  } else {
    // FP32: 4x4 register tile per thread, thread (tx, ty) = (tid % 16, tid / 16)
    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int kc = 0; kc < nk; kc++) {
      int buf = kc & 1;
      if (kc + 1 < nk) fetch(buf ^ 1, kc + 1);
#pragma unroll
      for (int k = 0; k < TK; k++) {
        float av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) av[i] = As[buf][k][ty + 16 * i];
#pragma unroll
        for (int j = 0; j < 4; j++) bv[j] = Bs[buf][k][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (kc + 1 < nk) {
        stash(buf ^ 1);
        if (kc + 2 < nk) load_koffs(buf, kc + 2);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        int ml = ty + 16 * i, nl = tx + 16 * j;
        if (m0 + ml < P.M.n && n0 + nl < P.N.n) C[bc + offCm[ml] + offCn[nl]] = acc[i][j];
      }
  }
}

static void cg_finish(hb_cgroup *g) {
  // collapse adjacent axes jointly over (sa, sb, sc)
  int64_t s[3][HB_R];
  for (int i = 0; i < g->rank; i++) { s[0][i] = g->sa[i]; s[1][i] = g->sb[i]; s[2][i] = g->sc[i]; }
  hb_collapse(&g->rank, g->shape, 3, s);
  g->n = 1;
  for (int i = 0; i < g->rank; i++) {
    g->sa[i] = s[0][i]; g->sb[i] = s[1][i]; g->sc[i] = s[2][i];
    g->n *= g->shape[i];
  }
}

static int64_t min_abs_stride(const hb_cgroup &g, const int64_t *s) {
  int64_t m = INT64_MAX;
  for (int i = 0; i < g.rank; i++) {
    int64_t x = s[i] < 0 ? -s[i] : s[i];
    if (x && x < m) m = x;
  }
  return m;
}

template <typename T>
static int launch_contract(hb_contractp &P, const hb_array *a, const hb_array *b, hb_array *o) {
  cg_finish(&P.M); cg_finish(&P.N); cg_finish(&P.B); cg_finish(&P.K);
  if (P.M.n == 0 || P.N.n == 0 || P.B.n == 0) return HB_OK;
  if constexpr (std::is_same<T, float>::value) {
    // Float: the tcgen05 engine (3xTF32 on the tensor cores, TMEM accumulators) takes every shape with a tile's
    // worth of work; what it declines stays on the FFMA register-tile kernel below
    int done = 0;
    HB_TRY(hb_contract_tf32(P, (const float *)hb_data(a), (const float *)hb_data(b), (float *)hb_data(o), &done));
    if (done) return HB_OK;
  }
  HB_REQUIRE(P.B.n < 65536 && (P.N.n + TN - 1) / TN < 65536, "contract: grid too large");
  bool a_kfast = min_abs_stride(P.K, P.K.sa) <= min_abs_stride(P.M, P.M.sa);
  bool b_kfast = min_abs_stride(P.K, P.K.sb) <= min_abs_stride(P.N, P.N.sb);
  dim3 grid((unsigned)((P.M.n + TM - 1) / TM), (unsigned)((P.N.n + TN - 1) / TN), (unsigned)P.B.n);
  const T *pa = (const T *)hb_data(a);
  const T *pb = (const T *)hb_data(b);
  T *pc = (T *)hb_data(o);
  if (a_kfast && b_kfast) hb_contract_kernel<T, true, true><<<grid, 256, 0, g_hb.stream>>>(P, pa, pb, pc);
  else if (a_kfast) hb_contract_kernel<T, true, false><<<grid, 256, 0, g_hb.stream>>>(P, pa, pb, pc);
  else if (b_kfast) hb_contract_kernel<T, false, true><<<grid, 256, 0, g_hb.stream>>>(P, pa, pb, pc);
  else hb_contract_kernel<T, false, false><<<grid, 256, 0, g_hb.stream>>>(P, pa, pb, pc);
  HB_LAUNCHED();
  return HB_OK;
}

// gemm_dmma.cu
int hb_gemm_dmma(const double *A, int64_t a_sm, int64_t a_sk, const double *B, int64_t b_sk, int64_t b_sn, double *C,
                 int64_t M, int64_t N, int64_t K, int accum, int *done);
int hb_gemm_dmma_pair(const double *A1, int64_t a1_sm, int64_t a1_sk, const double *B1, int64_t b1_sk, int64_t b1_sn,
                      int64_t K1, const double *A2, int64_t a2_sm, int64_t a2_sk, const double *B2, int64_t b2_sk,
                      int64_t b2_sn, int64_t K2, double *C1, double *C2, int64_t M, int64_t N, int chain, int accum,
                      int *done, const double *epi_bias = nullptr);
int hb_gemm_dmma_list(int n, const double *const *A, int64_t a_sm, int64_t a_sk, const double *const *B, int64_t b_sk,
                      int64_t b_sn, double *C, int64_t M, int64_t N, int64_t K, int accum, int *done);
int hb_reduce_generic(const hb_array *a, const hb_array *b, int nk, const int *kax, int nr,
                      const int *rax, hb_array *out);

// a, b: equal shapes.  Axes [0, n_lead) and the last axis are contracted; the
// axes in between index the result.
// `acc` (hb_matmul2_acc): a uniquely referenced contiguous f64 array of the result's shape that the DMMA GEMM
// may accumulate into (C += A x B); *acc_used reports whether it did (then *out = acc, retained).
static int contract_impl(const hb_array *a, const hb_array *b, int n_lead, hb_array **out, hb_array *acc = nullptr,
                         int *acc_used = nullptr) {
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(a && b && out, "null argument");
  HB_REQUIRE(a->dt == b->dt && hb_same_shape(a, b), "sdot1In: %s %s vs %s %s", hb_dtype_name(a->dt),
             hb_shape_str(a).c_str(), hb_dtype_name(b->dt), hb_shape_str(b).c_str());
  HB_REQUIRE(a->rank >= 1 && n_lead >= 0 && n_lead <= a->rank - 1, "sdot1In: bad rank");
  HB_REQUIRE(a->dt != HB_BOOL, "sdot1In: numeric dtype required");
  if (hb_lazy_on() && !acc)
    return hb_lrec(HB_L_DOT_INNER, n_lead ? "dot_inner_sum_leading" : "dot_inner").in(a).in(b).iarg(0, n_lead)
        .out(out, a->dt, a->rank - 1 - n_lead, a->shape + n_lead).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) { return contract_impl(in[0], in[1], (int)n.ia[0], &o[0]); });
  int r = a->rank;
  hb_array *o;
  HB_TRY(hb_new_array(a->dt, r - 1 - n_lead, a->shape + n_lead, &o));
  int st = HB_OK;
  int64_t K = a->shape[r - 1];
  for (int i = 0; i < n_lead; i++) K *= a->shape[i];
  int64_t outn = hb_numel(o);
  // classify the kept axes
  hb_contractp P;
  memset(&P, 0, sizeof P);
  bool gemm = hb_is_float(a->dt) && outn > 0 && K > 0;
  int64_t nM = 1, nN = 1;
  for (int i = n_lead; i < r - 1 && gemm; i++) {
    if (a->shape[i] == 1) continue;
    int64_t oc = o->strides[i - n_lead];
    hb_cgroup *g;
    if (a->strides[i] != 0 && b->strides[i] == 0) { g = &P.M; nM *= a->shape[i]; }
    else if (a->strides[i] == 0) { g = &P.N; nN *= a->shape[i]; }  // also when b is broadcast too (a constant operand, e.g. srepl 0)
    else g = &P.B;  // both move: batch
    int k = g->rank++;
    g->shape[k] = a->shape[i];
    g->sa[k] = a->strides[i];
    g->sb[k] = b->strides[i];
    g->sc[k] = oc;
  }
  // worth a GEMM only when there is reuse on both sides (the cp.async-staged
  // FP64 DMMA GEMM pads small sides cheaply, so it takes thin matrices too)
  bool thin_ok = a->dt == HB_F64 && nM >= 4 && nN >= 4 && nM * nN >= 512 && K >= 8;
  if (gemm && (nM < 16 || nN < 16 || K < 4) && !thin_ok) gemm = false;
  if (gemm) {
    for (int i = 0; i <= n_lead; i++) {
      int ax = i < n_lead ? i : r - 1;
      if (a->shape[ax] == 1) continue;
      int k = P.K.rank++;
      P.K.shape[k] = a->shape[ax];
      P.K.sa[k] = a->strides[ax];
      P.K.sb[k] = b->strides[ax];
      P.K.sc[k] = 0;
    }
    int done = 0;
    if (a->dt == HB_F64) {
      // plain 2-D (possibly transposed) operands -> the cp.async-staged DMMA GEMM
      cg_finish(&P.M); cg_finish(&P.N); cg_finish(&P.B); cg_finish(&P.K);
      if (P.M.rank == 1 && P.N.rank == 1 && P.K.rank == 1 && P.B.n == 1) {
        const double *pa = (const double *)hb_data(a), *pb = (const double *)hb_data(b);
        double *pc = (double *)hb_data(o);
        // an operand with no unit stride (e.g. a slice of a transposed batch) is
        // packed once: O(rows * K) against the O(M * N * K) product
        hb_array *ta = nullptr, *tb = nullptr;
        auto pack = [&](const hb_array *src, int64_t rows, int64_t s_row, int64_t s_k, hb_array **tmp) -> int {
          hb_array *v = hb_new_view(src);
          v->rank = 2;
          v->shape[0] = rows; v->strides[0] = s_row;
          v->shape[1] = P.K.n; v->strides[1] = s_k;
          int r = hb_contiguous(v, tmp);
          hb_release(v);
          return r;
        };
        if (st == HB_OK && P.M.sa[0] != 1 && P.K.sa[0] != 1) {
          st = pack(a, P.M.n, P.M.sa[0], P.K.sa[0], &ta);
          if (st == HB_OK) { pa = (const double *)hb_data(ta); P.M.sa[0] = P.K.n; P.K.sa[0] = 1; }
        }
        if (st == HB_OK && P.N.sb[0] != 1 && P.K.sb[0] != 1) {
          st = pack(b, P.N.n, P.N.sb[0], P.K.sb[0], &tb);
          if (st == HB_OK) { pb = (const double *)hb_data(tb); P.N.sb[0] = P.K.n; P.K.sb[0] = 1; }
        }
        if (st != HB_OK) { hb_release(ta); hb_release(tb); hb_release(o); return st; }
        const int accum = acc != nullptr;
        if (accum) pc = (double *)hb_data(acc);  // same shape, contiguous: same layout as o
        if (P.N.sc[0] == 1 && P.M.sc[0] == P.N.n)
          st = hb_gemm_dmma(pa, P.M.sa[0], P.K.sa[0], pb, P.K.sb[0], P.N.sb[0], pc, P.M.n, P.N.n, P.K.n, accum, &done);
        else if (P.M.sc[0] == 1 && P.N.sc[0] == P.M.n)  // output is [N, M]: C^T = B^T A^T
          st = hb_gemm_dmma(pb, P.N.sb[0], P.K.sb[0], pa, P.K.sa[0], P.M.sa[0], pc, P.N.n, P.M.n, P.K.n, accum, &done);
        if (accum && done && st == HB_OK) {
          hb_release(o);
          hb_retain(acc);
          o = acc;
          *acc_used = 1;
        }
        hb_release(ta);
        hb_release(tb);
      }
    }
    if (st == HB_OK && !done) {
      if (a->dt == HB_F64) st = launch_contract<double>(P, a, b, o);
      else st = launch_contract<float>(P, a, b, o);
    }
  } else {
    int kax[HB_R], rax[HB_R];
    int nk = 0, nr = 0;
    for (int i = 0; i < n_lead; i++) rax[nr++] = i;
    for (int i = n_lead; i < r - 1; i++) kax[nk++] = i;
    rax[nr++] = r - 1;
    st = hb_reduce_generic(a, b, nk, kax, nr, rax, o);
  }
  if (st != HB_OK) {
    hb_release(o);
    return st;
  }
  *out = o;
  return HB_OK;
}

extern "C" int hb_dot_inner(const hb_array *a, const hb_array *b, hb_array **out) {
  hb_prof_scope prof_("dot_inner", a, b);
  return contract_impl(a, b, 0, out);
}

extern "C" int hb_dot_inner_sum_leading(const hb_array *a, const hb_array *b, int n_axes, hb_array **out) {
  hb_prof_scope prof_("dot_inner_sum_leading", a, b);
  return contract_impl(a, b, n_axes, out);
}

// tsmatmul2 m1 m2 = sdot1In (tr [1,0] (replicate m1)) (tr [0,2,1] (replicate m2))
// (OpsConcrete.hs:237-242): built here directly as the two broadcast views.
extern "C" int hb_matmul2(const hb_array *a, const hb_array *b, hb_array **out) {
  hb_prof_scope prof_("matmul2", a, b);
  HB_REQUIRE(a && b && out, "null argument");
  HB_REQUIRE(a->rank == 2 && b->rank == 2 && a->shape[1] == b->shape[0], "smatmul2: %s x %s",
             hb_shape_str(a).c_str(), hb_shape_str(b).c_str());
  HB_CHECK_INIT_LAZY();
  if (hb_lazy_on()) {
    HB_REQUIRE(a->dt == b->dt, "smatmul2: dtype mismatch");
    int64_t sh_[2] = {a->shape[0], b->shape[1]};
    return hb_lrec(HB_L_MATMUL2, "matmul2").in(a).in(b).out(out, a->dt, 2, sh_).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_matmul2(in[0], in[1], &o[0]); });
  }
  hb_array *av = hb_new_view(a), *bv = hb_new_view(b);
  // av[m, p, n] = a[m, n] ; bv[m, p, n] = b[n, p]
  av->rank = bv->rank = 3;
  av->shape[0] = a->shape[0]; av->strides[0] = a->strides[0];
  av->shape[1] = b->shape[1]; av->strides[1] = 0;
  av->shape[2] = a->shape[1]; av->strides[2] = a->strides[1];
  bv->shape[0] = a->shape[0]; bv->strides[0] = 0;
  bv->shape[1] = b->shape[1]; bv->strides[1] = b->strides[1];
  bv->shape[2] = b->shape[0]; bv->strides[2] = b->strides[0];
  int st = contract_impl(av, bv, 0, out);
  hb_release(av);
  hb_release(bv);
  return st;
}

// acc + a x b.  Cotangent accumulation of a weight (taddTarget of evalRev, DeltaEval.hs:317-332, applied to the
// result of the transposition rule of tsmatmul2): when `acc` is uniquely referenced, contiguous and f64 and the
// DMMA GEMM takes the shape, the product is accumulated straight into acc's storage (the GEMM epilogue reads C) and
// *out = acc; otherwise the product is formed and added as hb_matmul2 + hb_add_inplace_or_new would.
extern "C" int hb_matmul2_acc(hb_array *acc, const hb_array *a, const hb_array *b, hb_array **out) {
  hb_prof_scope prof_("matmul2_acc", a, b);
  HB_REQUIRE(acc && a && b && out, "null argument");
  HB_REQUIRE(a->rank == 2 && b->rank == 2 && a->shape[1] == b->shape[0], "smatmul2: %s x %s",
             hb_shape_str(a).c_str(), hb_shape_str(b).c_str());
  HB_REQUIRE(acc->rank == 2 && acc->shape[0] == a->shape[0] && acc->shape[1] == b->shape[1] && acc->dt == a->dt,
             "matmul2_acc: accumulator %s %s does not match the product", hb_dtype_name(acc->dt),
             hb_shape_str(acc).c_str());
  HB_CHECK_INIT_LAZY();
  if (hb_lazy_on())
    return hb_lrec(HB_L_GENERIC, "matmul2_acc").in(acc).in(a).in(b).out(out, acc->dt, 2, acc->shape).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_matmul2_acc(in[0], in[1], in[2], &o[0]); });
  const bool unique = acc->dt == HB_F64 && acc->rc.load() == 1 && acc->st && acc->st->rc.load() == 1 &&
                      hb_contig(acc) && acc->st != a->st && acc->st != b->st;
  hb_array *av = hb_new_view(a), *bv = hb_new_view(b);
  av->rank = bv->rank = 3;
  av->shape[0] = a->shape[0]; av->strides[0] = a->strides[0];
  av->shape[1] = b->shape[1]; av->strides[1] = 0;
  av->shape[2] = a->shape[1]; av->strides[2] = a->strides[1];
  bv->shape[0] = a->shape[0]; bv->strides[0] = 0;
  bv->shape[1] = b->shape[1]; bv->strides[1] = b->strides[1];
  bv->shape[2] = b->shape[0]; bv->strides[2] = b->strides[0];
  hb_array *prod = nullptr;
  int used = 0;
  int st = contract_impl(av, bv, 0, &prod, unique ? acc : nullptr, &used);
  hb_release(av);
  hb_release(bv);
  if (st != HB_OK) return st;
  if (used) { *out = prod; return HB_OK; }
  st = hb_add_inplace_or_new(acc, prod, out);
  hb_release(prod);
  return st;
}

// ---- pairs of products in one launch (the recurrent layer wX x + wS s of MnistRnnShaped2.hs:55-69 and the
// transposition rules of its two tsmatmul2) ----
namespace {
struct Op2 { const double *p; int64_t s0, s1; };
// a rank-2 f64 array as a GEMM operand with one unit stride; size-1 axes get a benign stride
bool op2(const hb_array *a, Op2 *o) {
  if (!a || a->rank != 2 || a->dt != HB_F64 || a->shape[0] <= 0 || a->shape[1] <= 0) return false;
  o->p = (const double *)hb_data(a);
  o->s0 = a->strides[0];
  o->s1 = a->strides[1];
  if (a->shape[0] == 1) o->s0 = o->s1 == 1 ? a->shape[1] : 1;
  if (a->shape[1] == 1) o->s1 = o->s0 == 1 ? a->shape[0] : 1;
  return o->s0 == 1 || o->s1 == 1;
}
// ... packing an operand with no unit stride first (a row slice of a transposed batch, e.g. the glyph rows of
// MnistRnnShaped2): O(rows * K) against the O(M * N * K) product.  *tmp holds the packed copy (release it).
bool op2_pack(const hb_array *a, Op2 *o, hb_array **tmp) {
  *tmp = nullptr;
  if (op2(a, o)) return true;
  if (!a || a->rank != 2 || a->dt != HB_F64 || a->shape[0] <= 0 || a->shape[1] <= 0) return false;
  if (hb_contiguous(a, tmp) != HB_OK) { *tmp = nullptr; return false; }
  return op2(*tmp, o);
}
bool pair_shapes_ok(const hb_array *a1, const hb_array *b1, const hb_array *a2, const hb_array *b2) {
  return a1->rank == 2 && b1->rank == 2 && a2->rank == 2 && b2->rank == 2 && a1->shape[1] == b1->shape[0] &&
         a2->shape[1] == b2->shape[0] && a1->shape[0] == a2->shape[0] && b1->shape[1] == b2->shape[1];
}
bool acc_unique(const hb_array *acc, const hb_array *a, const hb_array *b) {
  return acc->dt == HB_F64 && acc->rc.load() == 1 && acc->st && acc->st->rc.load() == 1 && hb_contig(acc) &&
         acc->st != a->st && acc->st != b->st;
}
}  // namespace

extern "C" int hb_matmul2_sum(const hb_array *a1, const hb_array *b1, const hb_array *a2, const hb_array *b2,
                              hb_array **out) {
  hb_prof_scope prof_("matmul2_sum", a1, b1);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(a1 && b1 && a2 && b2 && out, "null argument");
  HB_REQUIRE(pair_shapes_ok(a1, b1, a2, b2), "smatmul2 + smatmul2: %s x %s + %s x %s", hb_shape_str(a1).c_str(),
             hb_shape_str(b1).c_str(), hb_shape_str(a2).c_str(), hb_shape_str(b2).c_str());
  if (hb_lazy_on()) {
    int64_t sh_[2] = {a1->shape[0], b1->shape[1]};
    return hb_lrec(HB_L_GENERIC, "matmul2_sum").in(a1).in(b1).in(a2).in(b2).out(out, a1->dt, 2, sh_).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_matmul2_sum(in[0], in[1], in[2], in[3], &o[0]); });
  }
  Op2 A1, B1, A2, B2;
  hb_array *tmp[4] = {nullptr, nullptr, nullptr, nullptr};
  auto drop = [&]() { for (hb_array *t : tmp) hb_release(t); };
  if (op2_pack(a1, &A1, &tmp[0]) && op2_pack(b1, &B1, &tmp[1]) && op2_pack(a2, &A2, &tmp[2]) &&
      op2_pack(b2, &B2, &tmp[3])) {
    int64_t sh[2] = {a1->shape[0], b1->shape[1]};
    hb_array *o;
    int st = hb_new_array(HB_F64, 2, sh, &o);
    if (st != HB_OK) { drop(); return st; }
    int done = 0;
    st = hb_gemm_dmma_pair(A1.p, A1.s0, A1.s1, B1.p, B1.s0, B1.s1, a1->shape[1], A2.p, A2.s0, A2.s1, B2.p, B2.s0,
                           B2.s1, a2->shape[1], (double *)hb_data(o), nullptr, sh[0], sh[1], 1, 0, &done);
    if (st != HB_OK || done) drop();
    if (st != HB_OK) { hb_release(o); return st; }
    if (done) { *out = o; return HB_OK; }
    hb_release(o);
  }
  drop();
  hb_array *t = nullptr;
  HB_TRY(hb_matmul2(a1, b1, &t));
  int st = hb_matmul2_acc(t, a2, b2, out);
  hb_release(t);
  return st;
}

extern "C" int hb_tanh_affine_fwd(const hb_array *u, const hb_array *v, const hb_array *bias, hb_array **y_out);

// tanh (a1 x b1 + a2 x b2 + bias stretched over the columns): the whole rnnMnistLayerS (MnistRnnShaped2.hs:55-69) as
// one launch -- the paired product with the bias + tanh in the epilogue of the last K slice of each tile's chain.
// Bit-identical to hb_matmul2_sum followed by hb_tanh_affine_fwd (same sums, same tanh), which is also the fallback.
extern "C" int hb_matmul2_sum_tanh(const hb_array *a1, const hb_array *b1, const hb_array *a2, const hb_array *b2,
                                   const hb_array *bias, hb_array **out) {
  hb_prof_scope prof_("matmul2_sum_tanh", a1, b1);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(a1 && b1 && a2 && b2 && bias && out, "null argument");
  HB_REQUIRE(pair_shapes_ok(a1, b1, a2, b2), "smatmul2 + smatmul2: %s x %s + %s x %s", hb_shape_str(a1).c_str(),
             hb_shape_str(b1).c_str(), hb_shape_str(a2).c_str(), hb_shape_str(b2).c_str());
  HB_REQUIRE(bias->rank == 1 && bias->shape[0] == a1->shape[0] && bias->dt == a1->dt,
             "matmul2_sum_tanh: bias must be [%lld] of the operands' type", (long long)a1->shape[0]);
  if (hb_lazy_on()) {
    int64_t sh_[2] = {a1->shape[0], b1->shape[1]};
    return hb_lrec(HB_L_GENERIC, "matmul2_sum_tanh").in(a1).in(b1).in(a2).in(b2).in(bias).out(out, a1->dt, 2, sh_).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_matmul2_sum_tanh(in[0], in[1], in[2], in[3], in[4], &o[0]); });
  }
  Op2 A1, B1, A2, B2;
  hb_array *tmp[4] = {nullptr, nullptr, nullptr, nullptr};
  auto drop = [&]() { for (hb_array *t : tmp) hb_release(t); };
  if (a1->dt == HB_F64 && hb_contig(bias) && op2_pack(a1, &A1, &tmp[0]) && op2_pack(b1, &B1, &tmp[1]) &&
      op2_pack(a2, &A2, &tmp[2]) && op2_pack(b2, &B2, &tmp[3])) {
    int64_t sh[2] = {a1->shape[0], b1->shape[1]};
    hb_array *o;
    int st = hb_new_array(HB_F64, 2, sh, &o);
    if (st != HB_OK) { drop(); return st; }
    int done = 0;
    st = hb_gemm_dmma_pair(A1.p, A1.s0, A1.s1, B1.p, B1.s0, B1.s1, a1->shape[1], A2.p, A2.s0, A2.s1, B2.p, B2.s0,
                           B2.s1, a2->shape[1], (double *)hb_data(o), nullptr, sh[0], sh[1], 1, 0, &done,
                           (const double *)hb_data(bias));
    drop();
    if (st != HB_OK) { hb_release(o); return st; }
    if (done) { *out = o; return HB_OK; }
    hb_release(o);
  } else {
    drop();
  }
  hb_array *pre = nullptr;
  HB_TRY(hb_matmul2_sum(a1, b1, a2, b2, &pre));
  int st = hb_tanh_affine_fwd(pre, nullptr, bias, out);
  hb_release(pre);
  return st;
}

extern "C" int hb_matmul2_pair(const hb_array *a1, const hb_array *b1, const hb_array *a2, const hb_array *b2,
                               hb_array **out1, hb_array **out2) {
  hb_prof_scope prof_("matmul2_pair", a1, b1);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(a1 && b1 && a2 && b2 && out1 && out2, "null argument");
  HB_REQUIRE(a1->rank == 2 && b1->rank == 2 && a2->rank == 2 && b2->rank == 2 && a1->shape[1] == b1->shape[0] &&
                 a2->shape[1] == b2->shape[0],
             "smatmul2 pair: %s x %s, %s x %s", hb_shape_str(a1).c_str(), hb_shape_str(b1).c_str(),
             hb_shape_str(a2).c_str(), hb_shape_str(b2).c_str());
  if (hb_lazy_on()) {
    int64_t s1_[2] = {a1->shape[0], b1->shape[1]}, s2_[2] = {a2->shape[0], b2->shape[1]};
    return hb_lrec(HB_L_GENERIC, "matmul2_pair").in(a1).in(b1).in(a2).in(b2).out(out1, a1->dt, 2, s1_).out(out2, a2->dt, 2, s2_)
        .commit([=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_matmul2_pair(in[0], in[1], in[2], in[3], &o[0], &o[1]); });
  }
  Op2 A1, B1, A2, B2;
  if (pair_shapes_ok(a1, b1, a2, b2) && op2(a1, &A1) && op2(b1, &B1) && op2(a2, &A2) && op2(b2, &B2)) {
    int64_t sh[2] = {a1->shape[0], b1->shape[1]};
    hb_array *o1, *o2;
    HB_TRY(hb_new_array(HB_F64, 2, sh, &o1));
    int st = hb_new_array(HB_F64, 2, sh, &o2);
    if (st != HB_OK) { hb_release(o1); return st; }
    int done = 0;
    st = hb_gemm_dmma_pair(A1.p, A1.s0, A1.s1, B1.p, B1.s0, B1.s1, a1->shape[1], A2.p, A2.s0, A2.s1, B2.p, B2.s0,
                           B2.s1, a2->shape[1], (double *)hb_data(o1), (double *)hb_data(o2), sh[0], sh[1], 0, 0, &done);
    if (st == HB_OK && done) { *out1 = o1; *out2 = o2; return HB_OK; }
    hb_release(o1);
    hb_release(o2);
    if (st != HB_OK) return st;
  }
  HB_TRY(hb_matmul2(a1, b1, out1));
  int st = hb_matmul2(a2, b2, out2);
  if (st != HB_OK) { hb_release(*out1); *out1 = nullptr; }
  return st;
}

extern "C" int hb_matmul2_acc_pair(hb_array *acc1, const hb_array *a1, const hb_array *b1, hb_array *acc2,
                                   const hb_array *a2, const hb_array *b2, hb_array **out1, hb_array **out2) {
  hb_prof_scope prof_("matmul2_acc_pair", a1, b1);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(acc1 && acc2 && a1 && b1 && a2 && b2 && out1 && out2, "null argument");
  if (hb_lazy_on())
    return hb_lrec(HB_L_GENERIC, "matmul2_acc_pair").in(acc1).in(a1).in(b1).in(acc2).in(a2).in(b2)
        .out(out1, acc1->dt, acc1->rank, acc1->shape).out(out2, acc2->dt, acc2->rank, acc2->shape)
        .commit([=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_matmul2_acc_pair(in[0], in[1], in[2], in[3], in[4], in[5], &o[0], &o[1]); });
  Op2 A1, B1, A2, B2;
  if (a1->rank == 2 && b1->rank == 2 && a2->rank == 2 && b2->rank == 2 && pair_shapes_ok(a1, b1, a2, b2) &&
      acc1->rank == 2 && acc2->rank == 2 && acc1->shape[0] == a1->shape[0] && acc1->shape[1] == b1->shape[1] &&
      hb_same_shape(acc1, acc2) && acc_unique(acc1, a1, b1) && acc_unique(acc2, a2, b2) && acc1->st != acc2->st &&
      acc1->st != a2->st && acc1->st != b2->st && acc2->st != a1->st && acc2->st != b1->st && op2(a1, &A1) &&
      op2(b1, &B1) && op2(a2, &A2) && op2(b2, &B2)) {
    int done = 0;
    HB_TRY(hb_gemm_dmma_pair(A1.p, A1.s0, A1.s1, B1.p, B1.s0, B1.s1, a1->shape[1], A2.p, A2.s0, A2.s1, B2.p, B2.s0,
                             B2.s1, a2->shape[1], (double *)hb_data(acc1), (double *)hb_data(acc2), acc1->shape[0],
                             acc1->shape[1], 0, 1, &done));
    if (done) {
      hb_retain(acc1);
      hb_retain(acc2);
      *out1 = acc1;
      *out2 = acc2;
      return HB_OK;
    }
  }
  HB_TRY(hb_matmul2_acc(acc1, a1, b1, out1));
  int st = hb_matmul2_acc(acc2, a2, b2, out2);
  if (st != HB_OK) { hb_release(*out1); *out1 = nullptr; }
  return st;
}

// A group of products of one [M, N] with named results, in one launch where the DMMA kernel takes them
// (hb_gemm_dmma_group):  outs[r] = f_r ((acc[r] or 0) + sum over g with res_index[g] == r of a_g x b_g, in order of g),
// f_r = tanh (. + bias[r] stretched over the columns) where bias[r] is given, else the identity.  This is what one
// time step of an unrolled two-layer recurrence asks for, forward (two layers' wX x + wS s + b, tanh of independent
// time steps: the wavefront of MnistRnnShaped2.hs:71-93) and backward (the transposition rules of tsmatmul2 of one
// sweep step, DeltaEval.hs:317-332: three products, two cotangents).  Value semantics: acc[r] is updated in place
// only when the caller holds its single reference.  acc / bias may be NULL (no accumulators / no epilogues).
// Falls back to one product at a time (same sums, possibly another association between K slices).
extern "C" int hb_matmul2_group(int n_res, hb_array *const *acc, const hb_array *const *bias, int n_prod,
                                const hb_array *const *as, const hb_array *const *bs, const int *res_index,
                                hb_array **outs) {
  hb_prof_scope prof_("matmul2_group", n_prod > 0 && as ? as[0] : nullptr, n_prod > 0 && bs ? bs[0] : nullptr);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(n_res >= 1 && n_res <= 16 && n_prod >= 1 && n_prod <= 32 && as && bs && res_index && outs,
             "matmul2_group: %d products into %d results", n_prod, n_res);
  const int64_t M = as[0] ? as[0]->shape[0] : 0, N = bs[0] ? bs[0]->shape[1] : 0;
  int count[16] = {0};
  for (int g = 0; g < n_prod; g++) {
    HB_REQUIRE(as[g] && bs[g] && as[g]->rank == 2 && bs[g]->rank == 2 && as[g]->shape[1] == bs[g]->shape[0] &&
                   as[g]->shape[0] == M && bs[g]->shape[1] == N && as[g]->dt == as[0]->dt && bs[g]->dt == as[0]->dt,
               "matmul2_group: product %d does not match the first one", g);
    HB_REQUIRE(res_index[g] >= 0 && res_index[g] < n_res, "matmul2_group: product %d names result %d of %d", g,
               res_index[g], n_res);
    count[res_index[g]]++;
  }
  for (int r = 0; r < n_res; r++) {
    HB_REQUIRE(count[r] > 0, "matmul2_group: result %d has no product", r);
    HB_REQUIRE(!acc || !acc[r] || (acc[r]->rank == 2 && acc[r]->shape[0] == M && acc[r]->shape[1] == N && acc[r]->dt == as[0]->dt),
               "matmul2_group: accumulator %d does not match the products", r);
    HB_REQUIRE(!bias || !bias[r] || (bias[r]->rank == 1 && bias[r]->shape[0] == M && bias[r]->dt == as[0]->dt),
               "matmul2_group: bias %d must be [%lld] of the operands' type", r, (long long)M);
    HB_REQUIRE(!(acc && acc[r] && bias && bias[r]), "matmul2_group: result %d has both an accumulator and an epilogue", r);
  }
  if (hb_lazy_on()) {
    hb_lrec rec(HB_L_GENERIC, "matmul2_group");
    for (int g = 0; g < n_prod; g++) rec.in(as[g]);
    for (int g = 0; g < n_prod; g++) rec.in(bs[g]);
    std::vector<int> idx(res_index, res_index + n_prod), has_acc(n_res, 0), has_bias(n_res, 0);
    for (int r = 0; r < n_res; r++) if (acc && acc[r]) { rec.in(acc[r]); has_acc[r] = 1; }
    for (int r = 0; r < n_res; r++) if (bias && bias[r]) { rec.in(bias[r]); has_bias[r] = 1; }
    int64_t sh_[2] = {M, N};
    for (int r = 0; r < n_res; r++) rec.out(&outs[r], as[0]->dt, 2, sh_);
    return rec.commit([=](hb_lnode &, hb_array *const *in, hb_array **o) {
      hb_array *ac[16];
      const hb_array *bi[16];
      int k = 2 * n_prod;
      for (int r = 0; r < n_res; r++) ac[r] = has_acc[r] ? in[k++] : nullptr;
      for (int r = 0; r < n_res; r++) bi[r] = has_bias[r] ? in[k++] : nullptr;
      return hb_matmul2_group(n_res, ac, bi, n_prod, in, in + n_prod, idx.data(), o);
    });
  }
  for (int r = 0; r < n_res; r++) outs[r] = nullptr;
  if (as[0]->dt == HB_F64 && M * N > 0 && n_prod <= 4 && n_res <= 4) {
    hb_array *tmp[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    auto drop = [&]() { for (hb_array *t : tmp) hb_release(t); };
    hb_gemm_prod pr[4];
    hb_gemm_res rs[4];
    bool ok = true;
    for (int g = 0; g < n_prod && ok; g++) {
      Op2 A, B;
      const bool bcast = (as[g]->strides[0] == 0 && as[g]->strides[1] == 0) || (bs[g]->strides[0] == 0 && bs[g]->strides[1] == 0);
      ok = !bcast && as[g]->shape[1] > 0 && op2_pack(as[g], &A, &tmp[2 * g]) && op2_pack(bs[g], &B, &tmp[2 * g + 1]);
      if (ok) pr[g] = hb_gemm_prod{A.p, A.s0, A.s1, B.p, B.s0, B.s1, as[g]->shape[1], res_index[g]};
    }
    hb_array *o[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int r = 0; r < n_res && ok; r++) {
      hb_array *a = acc ? acc[r] : nullptr;
      if (a) {
        bool unique = a->dt == HB_F64 && a->rc.load() == 1 && a->st && a->st->rc.load() == 1 && hb_contig(a);
        for (int g = 0; g < n_prod && unique; g++) unique = a->st != as[g]->st && a->st != bs[g]->st;
        for (int q = 0; q < r && unique; q++) unique = acc[q] != a;
        if (!unique) { ok = false; break; }
        hb_retain(a);
        o[r] = a;
      } else {
        int64_t sh[2] = {M, N};
        if (hb_new_array(HB_F64, 2, sh, &o[r]) != HB_OK) { ok = false; break; }
      }
      const hb_array *bi = bias ? bias[r] : nullptr;
      if (bi && !hb_contig(bi)) { ok = false; break; }
      rs[r] = hb_gemm_res{(double *)hb_data(o[r]), a ? 1 : 0, bi ? (const double *)hb_data(bi) : nullptr};
    }
    int done = 0, st = HB_OK;
    if (ok) st = hb_gemm_dmma_group(n_prod, pr, n_res, rs, M, N, &done);
    drop();   // stream-ordered: the packed copies are released behind the launch
    if (st == HB_OK && done) {
      for (int r = 0; r < n_res; r++) outs[r] = o[r];
      return HB_OK;
    }
    for (int r = 0; r < n_res; r++) hb_release(o[r]);
    if (st != HB_OK) return st;
  }
  // one product at a time
  int st = HB_OK;
  for (int r = 0; r < n_res && st == HB_OK; r++) {
    hb_array *cur = acc ? acc[r] : nullptr;
    if (cur) hb_retain(cur);
    for (int g = 0; g < n_prod && st == HB_OK; g++) {
      if (res_index[g] != r) continue;
      hb_array *nx = nullptr;
      st = cur ? hb_matmul2_acc(cur, as[g], bs[g], &nx) : hb_matmul2(as[g], bs[g], &nx);
      hb_release(cur);
      cur = nx;
    }
    if (st == HB_OK && bias && bias[r]) {
      hb_array *y = nullptr;
      st = hb_tanh_affine_fwd(cur, nullptr, bias[r], &y);
      hb_release(cur);
      cur = y;
    }
    outs[r] = st == HB_OK ? cur : nullptr;
  }
  if (st != HB_OK)
    for (int r = 0; r < n_res; r++) { hb_release(outs[r]); outs[r] = nullptr; }
  return st;
}

// acc + sum_t a_t x b_t for a list of products of one shape: the cotangent of a weight that is used at every step
// of an unrolled loop (taddTarget over the transposition rule of tsmatmul2 at every step, DeltaEval.hs:317-332),
// evaluated at the END of the sweep as ONE product over the concatenated reduced axes instead of one small product
// per step.  Deterministic (fixed slice order).  acc may be NULL.  Falls back to a chain of hb_matmul2_acc.
extern "C" int hb_matmul2_list(int n, hb_array *const *as, hb_array *const *bs, hb_array *acc, hb_array **out) {
  hb_prof_scope prof_("matmul2_list", n > 0 ? as[0] : nullptr, n > 0 ? bs[0] : nullptr);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(n >= 1 && as && bs && out, "matmul2_list: empty list");
  for (int t = 0; t < n; t++)
    HB_REQUIRE(as[t] && bs[t] && as[t]->rank == 2 && bs[t]->rank == 2 && as[t]->shape[1] == bs[t]->shape[0] &&
                   as[t]->dt == as[0]->dt && bs[t]->dt == as[0]->dt && hb_same_shape(as[t], as[0]) &&
                   hb_same_shape(bs[t], bs[0]),
               "matmul2_list: entry %d does not match the first product", t);
  const int64_t M = as[0]->shape[0], K = as[0]->shape[1], N = bs[0]->shape[1];
  HB_REQUIRE(!acc || (acc->rank == 2 && acc->shape[0] == M && acc->shape[1] == N && acc->dt == as[0]->dt),
             "matmul2_list: accumulator does not match the products");
  if (hb_lazy_on()) {
    hb_lrec r(HB_L_GENERIC, "matmul2_list");
    for (int t = 0; t < n; t++) r.in(as[t]);
    for (int t = 0; t < n; t++) r.in(bs[t]);
    if (acc) r.in(acc);
    int64_t sh_[2] = {M, N};
    return r.iarg(0, n).iarg(1, acc != nullptr).out(out, as[0]->dt, 2, sh_).commit(
        [=](hb_lnode &nd, hb_array *const *in, hb_array **o) {
          const int n_ = (int)nd.ia[0];
          return hb_matmul2_list(n_, in, in + n_, nd.ia[1] ? in[2 * n_] : nullptr, &o[0]);
        });
  }
  if (as[0]->dt == HB_F64 && M * N > 0 && n <= 32) {
    // the entries that share the stride pattern of the FIRST one (else of the last one) go into the one launch; the
    // odd ones (e.g. the broadcast zero state of time step 0, which arrives last) are accumulated one by one afterwards
    // operands without a unit stride (rows of a transposed batch, e.g. the glyph rows of MnistRnnShaped2) are packed
    // first: O(rows * K) each against the product
    hb_array *tmp[64];
    for (int t = 0; t < 2 * n; t++) tmp[t] = nullptr;
    auto drop_tmp = [&]() { for (int t = 0; t < 2 * n; t++) hb_release(tmp[t]); };
    Op2 OA[32], OB[32];
    bool usable[32];
    for (int t = 0; t < n; t++) {
      const bool bcast_a = as[t]->strides[0] == 0 && as[t]->strides[1] == 0, bcast_b = bs[t]->strides[0] == 0 && bs[t]->strides[1] == 0;
      usable[t] = !bcast_a && !bcast_b && op2_pack(as[t], &OA[t], &tmp[2 * t]) && op2_pack(bs[t], &OB[t], &tmp[2 * t + 1]);
    }
    int ref = -1;
    for (int t = 0; t < n && ref < 0; t++) if (usable[t]) ref = t;
    bool ok = ref >= 0;
    Op2 A0 = OA[ok ? ref : 0], B0 = OB[ok ? ref : 0];
    const double *pa[32], *pb[32];
    int nl = 0;
    bool in_list[32];
    for (int t = 0; t < n && ok; t++) {
      in_list[t] = usable[t] && OA[t].s0 == A0.s0 && OA[t].s1 == A0.s1 && OB[t].s0 == B0.s0 && OB[t].s1 == B0.s1;
      if (in_list[t]) { pa[nl] = OA[t].p; pb[nl] = OB[t].p; nl++; }
    }
    const bool inplace = acc && acc_unique(acc, as[0], bs[0]);
    if (!(ok && nl >= 2)) drop_tmp();
    if (ok && nl >= 2) {
      hb_array *o = nullptr;
      if (inplace) { hb_retain(acc); o = acc; }
      else {
        int64_t sh[2] = {M, N};
        HB_TRY(hb_new_array(HB_F64, 2, sh, &o));
      }
      int done = 0;
      int st = hb_gemm_dmma_list(nl, pa, A0.s0, A0.s1, pb, B0.s0, B0.s1, (double *)hb_data(o), M, N, K, inplace ? 1 : 0, &done);
      drop_tmp();   // stream-ordered: the packed copies are released behind the launch
      if (st != HB_OK) { hb_release(o); return st; }
      if (done) {
        if (acc && !inplace) {
          hb_array *sum = nullptr;
          st = hb_add_inplace_or_new(o, acc, &sum);
          hb_release(o);
          if (st != HB_OK) return st;
          o = sum;
        }
        for (int t = 0; t < n; t++) {
          if (in_list[t]) continue;
          hb_array *nx = nullptr;
          st = hb_matmul2_acc(o, as[t], bs[t], &nx);
          hb_release(o);
          if (st != HB_OK) return st;
          o = nx;
        }
        *out = o;
        return HB_OK;
      }
      hb_release(o);
    }
  }
  // generic: acc (+)= a_t x b_t one by one
  hb_array *cur = acc;
  if (cur) hb_retain(cur);
  for (int t = 0; t < n; t++) {
    hb_array *nx = nullptr;
    int st = cur ? hb_matmul2_acc(cur, as[t], bs[t], &nx) : hb_matmul2(as[t], bs[t], &nx);
    hb_release(cur);
    if (st != HB_OK) return st;
    cur = nx;
  }
  *out = cur;
  return HB_OK;
}

// tsmatvecmul m v = sdot1In m (sreplicate v)  (OpsConcrete.hs:236)
extern "C" int hb_matvecmul(const hb_array *m, const hb_array *v, hb_array **out) {
  hb_prof_scope prof_("matvecmul", m, v);
  HB_REQUIRE(m && v && out, "null argument");
  HB_REQUIRE(m->rank == 2 && v->rank == 1 && m->shape[1] == v->shape[0], "smatvecmul: %s x %s",
             hb_shape_str(m).c_str(), hb_shape_str(v).c_str());
  HB_CHECK_INIT_LAZY();
  if (hb_lazy_on())
    return hb_lrec(HB_L_GENERIC, "matvecmul").in(m).in(v).out(out, m->dt, 1, m->shape).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_matvecmul(in[0], in[1], &o[0]); });
  hb_array *vv = hb_new_view(v);
  vv->rank = 2;
  vv->shape[0] = m->shape[0]; vv->strides[0] = 0;
  vv->shape[1] = v->shape[0]; vv->strides[1] = v->strides[0];
  int st = contract_impl(m, vv, 0, out);
  hb_release(vv);
  return st;
}

// nn.cu -- fused NN blocks emitted by the device contraction pass in place of
// the gather chains of the vectorised artifacts: max-pooling with its adjoint
// (maxPool2dUnpaddedS, CommonShapedOps.hs:241-260; smaximum :33-36; the
// gather/scatter form is at TestMnistPP.hs:735) and the fused
// softmax-cross-entropy loss (lossSoftMaxCrossEntropyS, :89-111).
// All kernels are HBM-bound single passes.
#include "common.cuh"
#include "comm.cuh"
#include "lazy.cuh"
#include "iter.cuh"

// one thread per output window; value at the FIRST maximum of the row-major
// flattened window (kargMax of the flattened slice); positions outside the
// input read as 0 (slicezS, CommonShapedOps.hs:225-237).
template <typename T>
__global__ void __launch_bounds__(256)
    hb_maxpool_kernel(const T *__restrict__ x, hb_dims xd, int ks, int stride, int64_t H, int64_t W,
                      int64_t Ho, int64_t Wo, int64_t total, T *__restrict__ y, int64_t *__restrict__ am) {
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = gid; o < total; o += gs) {
    int64_t j = o % Wo, t = o / Wo;
    int64_t i = t % Ho;
    t /= Ho;
    int64_t c = t % xd.shape[1], n = t / xd.shape[1];
    int64_t base = n * xd.stride[0] + c * xd.stride[1];
    T best = T(0);
    int64_t bi = 0;
    bool first = true;
    for (int a = 0; a < ks; a++)
      for (int b = 0; b < ks; b++) {
        int64_t yy = i * stride + a, xx = j * stride + b;
        T v = (yy < H && xx < W) ? x[base + yy * xd.stride[2] + xx * xd.stride[3]] : T(0);
        if (first || v > best) {
          best = v;
          bi = a * ks + b;
          first = false;
        }
      }
    y[o] = best;
    if (am) am[o] = bi;
  }
}

extern "C" int hb_maxpool2d(const hb_array *x, int ksize, int stride, hb_array **out, hb_array **argmax_out) {
  hb_prof_scope prof_("maxpool2d", x);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(x && out, "null argument");
  HB_REQUIRE(x->rank == 4 && hb_is_float(x->dt), "maxPool2d: rank-4 floating input required");
  HB_REQUIRE(ksize >= 1 && stride >= 1, "maxPool2d: bad window");
  int64_t H = x->shape[2], W = x->shape[3];
  int64_t osh[4] = {x->shape[0], x->shape[1], H / stride, W / stride};
  if (hb_lazy_on())
    return hb_lrec(HB_L_GENERIC, "maxpool2d").in(x).iarg(0, ksize).iarg(1, stride).out(out, x->dt, 4, osh).out(argmax_out, HB_I64, 4, osh)
        .commit([=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_maxpool2d(in[0], (int)n.ia[0], (int)n.ia[1], &o[0], n.out[1] ? &o[1] : nullptr); });
  hb_array *y, *am = nullptr;
  HB_TRY(hb_new_array(x->dt, 4, osh, &y));
  if (argmax_out) {
    int st = hb_new_array(HB_I64, 4, osh, &am);
    if (st != HB_OK) { hb_release(y); return st; }
  }
  int64_t total = hb_numel(y);
  if (total > 0) {
    hb_dims xd;
    xd.rank = 4;
    for (int i = 0; i < HB_R; i++) { xd.shape[i] = i < 4 ? x->shape[i] : 1; xd.stride[i] = i < 4 ? x->strides[i] : 0; }
    int blocks = hb_blocks_for(total, 256, g_hb.sm_count * 32);
    HB_DISPATCH_FLOAT(x->dt, T,
                      (hb_maxpool_kernel<T><<<blocks, 256, 0, g_hb.stream>>>(
                          (const T *)hb_data(x), xd, ksize, stride, H, W, osh[2], osh[3], total, (T *)hb_data(y),
                          am ? (int64_t *)hb_data(am) : nullptr)));
    HB_LAUNCHED();
  }
  *out = y;
  if (argmax_out) *argmax_out = am;
  return HB_OK;
}

// adjoint in gather form: one thread per INPUT position sums, in row-major
// window order, the cotangents of the windows whose argmax selected it.
template <typename T>
__global__ void __launch_bounds__(256)
    hb_maxpool_bwd_kernel(const T *__restrict__ dy, const int64_t *__restrict__ am, int ks, int stride,
                          int64_t C, int64_t H, int64_t W, int64_t Ho, int64_t Wo, int64_t total,
                          T *__restrict__ dx) {
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = gid; p < total; p += gs) {
    int64_t xx = p % W, t = p / W;
    int64_t yy = t % H;
    int64_t nc = t / H;
    T acc = T(0);
    // windows (i, j) with i*stride <= yy < i*stride + ks
    int64_t i_hi = yy / stride, j_hi = xx / stride;
    int64_t i_lo = yy - ks + 1 <= 0 ? 0 : (yy - ks + 1 + stride - 1) / stride;
    int64_t j_lo = xx - ks + 1 <= 0 ? 0 : (xx - ks + 1 + stride - 1) / stride;
    for (int64_t i = i_lo; i <= i_hi && i < Ho; i++)
      for (int64_t j = j_lo; j <= j_hi && j < Wo; j++) {
        int64_t a = yy - i * stride, b = xx - j * stride;
        int64_t o = (nc * Ho + i) * Wo + j;
        if (am[o] == a * ks + b) acc = dy[o] + acc;
      }
    dx[p] = acc;
  }
}

extern "C" int hb_maxpool2d_bwd(const hb_array *dy, const hb_array *argmax, int ksize, int stride, int64_t H,
                                int64_t W, hb_array **out) {
  hb_prof_scope prof_("maxpool2d_bwd", dy);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(dy && argmax && out, "null argument");
  HB_REQUIRE(dy->rank == 4 && hb_is_float(dy->dt) && argmax->dt == HB_I64 && hb_same_shape(dy, argmax),
             "maxPool2d_bwd: bad operands");
  HB_REQUIRE(dy->shape[2] == H / stride && dy->shape[3] == W / stride, "maxPool2d_bwd: output extents do not match");
  if (hb_lazy_on()) {
    int64_t sh_[4] = {dy->shape[0], dy->shape[1], H, W};
    return hb_lrec(HB_L_GENERIC, "maxpool2d_bwd").in(dy).in(argmax).iarg(0, ksize).iarg(1, stride).out(out, dy->dt, 4, sh_).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_maxpool2d_bwd(in[0], in[1], (int)n.ia[0], (int)n.ia[1], n.out[0]->shape[2], n.out[0]->shape[3], &o[0]); });
  }
  const hb_array *dyc = dy, *amc = argmax;
  hb_array *t1 = nullptr, *t2 = nullptr;
  if (!hb_contig(dy)) { HB_TRY(hb_contiguous(dy, &t1)); dyc = t1; }
  if (!hb_contig(argmax)) {
    int st = hb_contiguous(argmax, &t2);
    if (st != HB_OK) { hb_release(t1); return st; }
    amc = t2;
  }
  int64_t xsh[4] = {dy->shape[0], dy->shape[1], H, W};
  hb_array *dx;
  int st = hb_new_array(dy->dt, 4, xsh, &dx);
  if (st == HB_OK) {
    int64_t total = hb_numel(dx);
    if (total > 0) {
      int blocks = hb_blocks_for(total, 256, g_hb.sm_count * 32);
      if (dy->dt == HB_F64)
        hb_maxpool_bwd_kernel<double><<<blocks, 256, 0, g_hb.stream>>>(
            (const double *)hb_data(dyc), (const int64_t *)hb_data(amc), ksize, stride, dy->shape[1], H, W,
            dy->shape[2], dy->shape[3], total, (double *)hb_data(dx));
      else
        hb_maxpool_bwd_kernel<float><<<blocks, 256, 0, g_hb.stream>>>(
            (const float *)hb_data(dyc), (const int64_t *)hb_data(amc), ksize, stride, dy->shape[1], H, W,
            dy->shape[2], dy->shape[3], total, (float *)hb_data(dx));
      g_hb.launches.fetch_add(1);
      cudaError_t e = cudaPeekAtLastError();
      if (e != cudaSuccess) { cudaGetLastError(); hb_set_error("maxpool bwd launch: %s", cudaGetErrorString(e)); st = HB_ERR_CUDA; }
    }
  }
  hb_release(t1);
  hb_release(t2);
  if (st != HB_OK) { if (dx) hb_release(dx); return st; }
  *out = dx;
  return HB_OK;
}

// ------------------------------------------- fused bias + relu + max-pool
// The tail of convMnistLayerS (MnistCnnShaped2.hs:42-62) as ONE pass:
//   yRelu = reluS (yConv + biasStretched);  yPool = maxPool2dUnpaddedS @k @k yRelu
// for non-overlapping windows (ksize == stride, the CNN's 2x2/2).  The full-size
// pre-activation is read once and never written back; what the adjoint needs
// is one byte per pooled element:
//   code = argmax within the window (row-major, first maximum, as kargMax)
//          | 0x80 if the selected pre-activation is > 0 (the relu mask there).
// relu is (x <= 0 ? 0 : 1) * x exactly as CommonShapedOps.hs:38-44 (so
// negative inputs give -0.0 and ties resolve as in the unfused chain).
template <typename T, bool VEC2>
__global__ void __launch_bounds__(256)
    hb_relu_pool_fwd_kernel(const T *__restrict__ pre, const T *__restrict__ bias, int C, int H, int W, int ks,
                            int Ho, int Wo, int64_t total, T *__restrict__ y, uint8_t *__restrict__ code) {
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = gid; o < total; o += gs) {
    int j = (int)(o % Wo);
    int64_t t = o / Wo;
    int i = (int)(t % Ho);
    int64_t nc = t / Ho;
    T b = bias ? bias[nc % C] : T(0);
    const T *p = pre + (nc * H + (int64_t)i * ks) * W + (int64_t)j * ks;
    T best = T(0), bestpre = T(0);
    int bi = 0;
    if (VEC2) {  // ks == 2, W even, 2-element aligned
      typedef hb_vec<T, 2> V;
      V r0 = *reinterpret_cast<const V *>(p), r1 = *reinterpret_cast<const V *>(p + W);
      T v[4] = {r0.v[0] + b, r0.v[1] + b, r1.v[0] + b, r1.v[1] + b};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        T r = (v[k] <= T(0) ? T(0) : T(1)) * v[k];
        if (k == 0 || r > best) { best = r; bestpre = v[k]; bi = k; }
      }
    } else {
      for (int a = 0; a < ks; a++)
        for (int c = 0; c < ks; c++) {
          T v = p[(int64_t)a * W + c] + b;
          T r = (v <= T(0) ? T(0) : T(1)) * v;
          if ((a == 0 && c == 0) || r > best) { best = r; bestpre = v; bi = a * ks + c; }
        }
    }
    y[o] = best;
    code[o] = (uint8_t)(bi | (bestpre > T(0) ? 0x80 : 0));
  }
}

extern "C" int hb_relu_pool_fwd(const hb_array *pre, const hb_array *bias, int ksize, hb_array **out,
                                hb_array **code_out) {
  hb_prof_scope prof_("relu_pool_fwd", pre);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(pre && out && code_out, "null argument");
  HB_REQUIRE(pre->rank == 4 && hb_is_float(pre->dt), "relu_pool: rank-4 floating input required");
  HB_REQUIRE(ksize >= 1 && ksize <= 11, "relu_pool: window must be 1..11");
  HB_REQUIRE(!bias || (bias->rank == 1 && bias->dt == pre->dt && bias->shape[0] == pre->shape[1]),
             "relu_pool: bias must be [%lld]", (long long)pre->shape[1]);
  if (hb_lazy_on()) {
    int64_t sh_[4] = {pre->shape[0], pre->shape[1], pre->shape[2] / ksize, pre->shape[3] / ksize};
    hb_lrec r(HB_L_RELU_POOL_FWD, "relu_pool_fwd");
    r.in(pre);
    if (bias) r.in(bias);
    return r.iarg(0, ksize).out(out, pre->dt, 4, sh_).out(code_out, HB_I8, 4, sh_).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_relu_pool_fwd(in[0], n.in.size() > 1 ? in[1] : nullptr, (int)n.ia[0], &o[0], &o[1]); });
  }
  const hb_array *pc = pre, *bc = bias;
  hb_array *t1 = nullptr, *t2 = nullptr;
  if (!hb_contig(pre)) { HB_TRY(hb_contiguous(pre, &t1)); pc = t1; }
  if (bias && !hb_contig(bias)) {
    int st = hb_contiguous(bias, &t2);
    if (st != HB_OK) { hb_release(t1); return st; }
    bc = t2;
  }
  int64_t H = pre->shape[2], W = pre->shape[3];
  int64_t osh[4] = {pre->shape[0], pre->shape[1], H / ksize, W / ksize};
  hb_array *y = nullptr, *code = nullptr;
  int st = hb_new_array(pre->dt, 4, osh, &y);
  if (st == HB_OK) st = hb_new_array(HB_I8, 4, osh, &code);
  if (st == HB_OK) {
    int64_t total = hb_numel(y);
    if (total > 0) {
      int blocks = hb_blocks_for(total, 256, g_hb.sm_count * 32);
      size_t es = hb_dtype_size(pre->dt);
      bool vec = ksize == 2 && W % 2 == 0 && ((uintptr_t)hb_data(pc) % (2 * es)) == 0;
#define RP_LAUNCH(T, V)                                                                                           \
  hb_relu_pool_fwd_kernel<T, V><<<blocks, 256, 0, g_hb.stream>>>(                                                 \
      (const T *)hb_data(pc), bc ? (const T *)hb_data(bc) : nullptr, (int)pre->shape[1], (int)H, (int)W, ksize,    \
      (int)osh[2], (int)osh[3], total, (T *)hb_data(y), (uint8_t *)hb_data(code))
      if (pre->dt == HB_F64) { if (vec) RP_LAUNCH(double, true); else RP_LAUNCH(double, false); }
      else { if (vec) RP_LAUNCH(float, true); else RP_LAUNCH(float, false); }
#undef RP_LAUNCH
      g_hb.launches.fetch_add(1);
      if (g_hb.profiling) hb_prof_mark(__func__, __LINE__);
      cudaError_t e = cudaPeekAtLastError();
      if (e != cudaSuccess) { cudaGetLastError(); hb_set_error("relu_pool launch: %s", cudaGetErrorString(e)); st = HB_ERR_CUDA; }
    }
  }
  hb_release(t1);
  hb_release(t2);
  if (st != HB_OK) { hb_release(y); hb_release(code); return st; }
  *out = y;
  *code_out = code;
  return HB_OK;
}

// Adjoint: d_pre[window position] = dy if (position == argmax and relu active)
// else 0 (maxpool scatter then mask, DeltaEval.hs:538-562 / :392-396), and the
// bias cotangent  dbias[c] = sum_{n,i,j} of the same values, reduced per
// (channel, image chunk) block in a fixed tree order, then over chunks in
// order by hb_bias_finish_kernel: deterministic.
template <typename T, bool VEC2, bool WRITE>
__global__ void __launch_bounds__(256)
    hb_relu_pool_bwd_kernel(const T *__restrict__ dy, const uint8_t *__restrict__ code, int N, int C, int H, int W,
                            int ks, int Ho, int Wo, int npc, T *__restrict__ dpre, T *__restrict__ part) {
  __shared__ T sm[8];
  const int c = blockIdx.x, chunk = blockIdx.y;
  const int n0 = chunk * npc, n1 = n0 + npc < N ? n0 + npc : N;
  const unsigned HW = (unsigned)(Ho * Wo);
  const unsigned cnt = (unsigned)(n1 - n0) * HW;   // < 2^31 (checked by the host)
  T acc = T(0);
  // one pooled element: 32-bit index arithmetic, the window written with two 16-byte stores
  auto elem = [&](unsigned i, uint8_t cd, T dv) {
    T g = (cd & 0x80) ? dv : T(0);
    acc += g;
    if (!WRITE) return;
    unsigned nn = i / HW, w = i - nn * HW;
    unsigned wi = w / (unsigned)Wo, wj = w - wi * (unsigned)Wo;
    int bi = cd & 0x7f;
    T *p = dpre + (((int64_t)(n0 + (int)nn) * C + c) * H + (int64_t)wi * ks) * W + (int64_t)wj * ks;
    if (VEC2) {
      typedef hb_vec<T, 2> V;
      V r0, r1;
      r0.v[0] = bi == 0 ? g : T(0); r0.v[1] = bi == 1 ? g : T(0);
      r1.v[0] = bi == 2 ? g : T(0); r1.v[1] = bi == 3 ? g : T(0);
      *reinterpret_cast<V *>(p) = r0;
      *reinterpret_cast<V *>(p + W) = r1;
    } else {
      for (int a = 0; a < ks; a++)
        for (int b = 0; b < ks; b++) p[(int64_t)a * W + b] = (a * ks + b == bi) ? g : T(0);
    }
  };
  // element i of image n0 + i / HW sits at ((n0 * C + c) * HW + i) + (i / HW) * (C - 1) * HW
  const T *dyb = dy + ((int64_t)n0 * C + c) * HW;
  const uint8_t *cdb = code + ((int64_t)n0 * C + c) * HW;
  const int64_t skip = (int64_t)(C - 1) * HW;
  unsigned i = threadIdx.x;
  for (; i + 768 < cnt; i += 1024) {           // four independent loads of dy and code in flight
    int64_t o0 = i + (i / HW) * skip, o1 = (i + 256) + ((i + 256) / HW) * skip;
    int64_t o2 = (i + 512) + ((i + 512) / HW) * skip, o3 = (i + 768) + ((i + 768) / HW) * skip;
    uint8_t c0 = cdb[o0], c1 = cdb[o1], c2 = cdb[o2], c3 = cdb[o3];
    T d0 = dyb[o0], d1 = dyb[o1], d2 = dyb[o2], d3 = dyb[o3];
    elem(i, c0, d0);
    elem(i + 256, c1, d1);
    elem(i + 512, c2, d2);
    elem(i + 768, c3, d3);
  }
  for (; i < cnt; i += 256) {
    int64_t o0 = i + (i / HW) * skip;
    elem(i, cdb[o0], dyb[o0]);
  }
  acc = hb_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    T s = sm[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); i++) s += sm[i];
    part[(int64_t)c * gridDim.y + chunk] = s;
  }
}

// one warp per channel: lane l adds chunks l, l + 32, ... in order, the 32 lane sums are combined by a fixed shuffle
// tree (deterministic; a single thread walking all chunks was a chain of dependent loads: 7 us for 74 chunks)
template <typename T>
__global__ void hb_bias_finish_kernel(const T *__restrict__ part, int C, int nchunks, T *__restrict__ out) {
  const int c = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  T s = T(0);
  for (int k = lane; k < nchunks; k += 32) s += part[(int64_t)c * nchunks + k];
  s = hb_warp_sum(s);
  if (lane == 0) out[c] = s;
}

extern "C" int hb_relu_pool_bwd(const hb_array *dy, const hb_array *code, int ksize, int64_t H, int64_t W,
                                hb_array **dpre_out, hb_array **dbias_out) {
  hb_prof_scope prof_("relu_pool_bwd", dy);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(dy && code && (dpre_out || dbias_out), "null argument");
  const bool write = dpre_out != nullptr;  // NULL: only the bias cotangent is wanted
  HB_REQUIRE(dy->rank == 4 && hb_is_float(dy->dt) && code->dt == HB_I8 && hb_same_shape(dy, code),
             "relu_pool_bwd: bad operands");
  HB_REQUIRE(ksize >= 1 && ksize <= 11 && dy->shape[2] == H / ksize && dy->shape[3] == W / ksize,
             "relu_pool_bwd: extents do not match");
  if (hb_lazy_on()) {
    int64_t sh_[4] = {dy->shape[0], dy->shape[1], H, W};
    return hb_lrec(HB_L_RELU_POOL_BWD, "relu_pool_bwd").in(dy).in(code).iarg(0, ksize).iarg(1, H).iarg(2, W)
        .out(dpre_out, dy->dt, 4, sh_).out(dbias_out, dy->dt, 1, &dy->shape[1])
        .commit([=](hb_lnode &n, hb_array *const *in, hb_array **o) {
          return hb_relu_pool_bwd(in[0], in[1], (int)n.ia[0], n.ia[1], n.ia[2], n.out[0] ? &o[0] : nullptr,
                                  n.out[1] ? &o[1] : nullptr);
        });
  }
  const hb_array *dyc = dy, *cc = code;
  hb_array *t1 = nullptr, *t2 = nullptr;
  if (!hb_contig(dy)) { HB_TRY(hb_contiguous(dy, &t1)); dyc = t1; }
  if (!hb_contig(code)) {
    int st = hb_contiguous(code, &t2);
    if (st != HB_OK) { hb_release(t1); return st; }
    cc = t2;
  }
  int64_t N = dy->shape[0], C = dy->shape[1];
  int64_t xsh[4] = {N, C, H, W};
  hb_array *dx = nullptr, *db = nullptr;
  int st = write ? hb_new_array(dy->dt, 4, xsh, &dx) : HB_OK;
  if (st == HB_OK) st = hb_new_array(dy->dt, 1, &C, &db);
  if (st == HB_OK && N * C * H * W > 0) {
    size_t es = hb_dtype_size(dy->dt);
    bool leftover = write && ((H % ksize) || (W % ksize));
    if (leftover) {
      cudaError_t e = cudaMemsetAsync(hb_data(dx), 0, (size_t)hb_numel(dx) * es, g_hb.stream);
      if (e != cudaSuccess) { hb_set_error("memset: %s", cudaGetErrorString(e)); st = HB_ERR_CUDA; }
    }
    int nchunks = (int)((8 * (int64_t)g_hb.sm_count + C - 1) / C);
    if (nchunks > N) nchunks = (int)N;
    if (nchunks < 1) nchunks = 1;
    int npc = (int)((N + nchunks - 1) / nchunks);
    nchunks = (int)((N + npc - 1) / npc);
    if (st == HB_OK && (int64_t)npc * dy->shape[2] * dy->shape[3] >= (1ll << 31)) {
      hb_set_error("relu_pool_bwd: more than 2^31 pooled elements per (channel, image chunk)");
      st = HB_ERR_INVALID;
    }
    void *scr = nullptr;
    if (st == HB_OK) st = hb_scratch((size_t)C * nchunks * es, &scr);
    if (st == HB_OK && hb_numel(dy) > 0) {
      bool vec = write && ksize == 2 && W % 2 == 0 && ((uintptr_t)hb_data(dx) % (2 * es)) == 0;
      dim3 grid((unsigned)C, (unsigned)nchunks);
#define RPB_LAUNCH(T, V, WR)                                                                                \
  hb_relu_pool_bwd_kernel<T, V, WR><<<grid, 256, 0, g_hb.stream>>>(                                         \
      (const T *)hb_data(dyc), (const uint8_t *)hb_data(cc), (int)N, (int)C, (int)H, (int)W, ksize,         \
      (int)dy->shape[2], (int)dy->shape[3], npc, write ? (T *)hb_data(dx) : nullptr, (T *)scr)
      if (dy->dt == HB_F64) { if (!write) RPB_LAUNCH(double, false, false); else if (vec) RPB_LAUNCH(double, true, true); else RPB_LAUNCH(double, false, true); }
      else { if (!write) RPB_LAUNCH(float, false, false); else if (vec) RPB_LAUNCH(float, true, true); else RPB_LAUNCH(float, false, true); }
#undef RPB_LAUNCH
      g_hb.launches.fetch_add(1);
      if (g_hb.profiling) hb_prof_mark(__func__, __LINE__);
      if (dy->dt == HB_F64)
        hb_bias_finish_kernel<double><<<(unsigned)((C + 3) / 4), 128, 0, g_hb.stream>>>((const double *)scr, (int)C, nchunks, (double *)hb_data(db));
      else
        hb_bias_finish_kernel<float><<<(unsigned)((C + 3) / 4), 128, 0, g_hb.stream>>>((const float *)scr, (int)C, nchunks, (float *)hb_data(db));
      g_hb.launches.fetch_add(1);
      if (g_hb.profiling) hb_prof_mark(__func__, __LINE__);
      cudaError_t e = cudaPeekAtLastError();
      if (e != cudaSuccess) { cudaGetLastError(); hb_set_error("relu_pool_bwd launch: %s", cudaGetErrorString(e)); st = HB_ERR_CUDA; }
    } else if (st == HB_OK && C > 0) {
      cudaMemsetAsync(hb_data(db), 0, (size_t)C * es, g_hb.stream);
    }
  }
  hb_release(t1);
  hb_release(t2);
  if (st != HB_OK) { hb_release(dx); hb_release(db); return st; }
  if (dpre_out) *dpre_out = dx;
  if (dbias_out) *dbias_out = db; else hb_release(db);
  return HB_OK;
}

// ---------------------------------------------------------- softmax-xent
// lossSoftMaxCrossEntropyS (CommonShapedOps.hs:89-111), normalising over the
// WHOLE tensor exactly as the reference does:
//   expU = exp (u - minimum u); softMax = recip (sum expU) * expU
//   loss = negate (log softMax `dot` expected); dual = (softMax - expected) * du
// One block (the logits of a minibatch are a few 10^4 elements); three
// block-wide deterministic tree reductions.
template <typename T>
__device__ T hb_block_reduce(T v, T *smem, bool is_min) {
  for (int o = 16; o > 0; o >>= 1) {
    T u = hb_shfl_xor(v, o);
    v = is_min ? (u < v ? u : v) : v + u;
  }
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) smem[w] = v;
  __syncthreads();
  v = smem[0];
  for (int i = 1; i < nw; i++) v = is_min ? (smem[i] < v ? smem[i] : v) : v + smem[i];
  return v;  // every thread
}

// Multi-block version (the [10, batch] logits of a 4096-sample minibatch are
// 40960 elements; one block spent 150 us on them): three grid-wide phases,
// each a kernel whose blocks first combine the previous phase's per-block
// partials IN BLOCK ORDER (so every block derives the identical global value
// and the result is deterministic), then do their share of the next phase.
//   phase 0: partial minima          phase 1: partial sums of exp(u - min)
//   phase 2: dl = scale*(softmax - expected), partial sums of log(softmax)*expected
//   phase 3: loss = scale * -(sum of partials)
template <typename T>
__device__ T hb_combine(const T *part, int n, bool is_min) {
  T v = part[0];
  for (int i = 1; i < n; i++) v = is_min ? (part[i] < v ? part[i] : v) : v + part[i];
  return v;
}

template <typename T, int PHASE>
__global__ void __launch_bounds__(256)
    hb_softmax_xent_phase(const T *__restrict__ u, hb_dims ud, const T *__restrict__ ex, hb_dims ed, int64_t n,
                          T scale, T *__restrict__ pmin, T *__restrict__ psum, T *__restrict__ ploss,
                          T *__restrict__ loss, T *__restrict__ dl) {
  __shared__ T smem[32];
  const int nb = gridDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gs = (int64_t)nb * blockDim.x;
  if (PHASE == 0) {
    T mn = u[hb_offset_of(ud, i0 < n ? i0 : 0)];
    for (int64_t i = i0; i < n; i += gs) {
      T v = u[hb_offset_of(ud, i)];
      mn = v < mn ? v : mn;
    }
    mn = hb_block_reduce<T>(mn, smem, true);
    if (threadIdx.x == 0) pmin[blockIdx.x] = mn;
  } else if (PHASE == 1) {
    T mn = hb_combine<T>(pmin, nb, true);
    T s = T(0);
    for (int64_t i = i0; i < n; i += gs) s += exp(u[hb_offset_of(ud, i)] - mn);
    s = hb_block_reduce<T>(s, smem, false);
    if (threadIdx.x == 0) psum[blockIdx.x] = s;
  } else if (PHASE == 2) {
    T mn = hb_combine<T>(pmin, nb, true);
    T rs = T(1) / hb_combine<T>(psum, nb, false);
    T l = T(0);
    for (int64_t i = i0; i < n; i += gs) {
      T sm = rs * exp(u[hb_offset_of(ud, i)] - mn);
      T e = ex[hb_offset_of(ed, i)];
      l += log(sm) * e;
      if (dl) dl[i] = scale * (sm - e);
    }
    l = hb_block_reduce<T>(l, smem, false);
    if (threadIdx.x == 0) ploss[blockIdx.x] = l;
  } else {
    if (blockIdx.x == 0 && threadIdx.x == 0) *loss = scale * (-hb_combine<T>(ploss, (int)n, false));
  }
}

// One CTA for the whole loss (small minibatches -- a 512-sample shard has 5120 logits -- and every global-objective
// step): min, sum of exponentials, loss and dual in ONE launch instead of four, fixed reduction trees -> deterministic.
// exp (u - min) is computed once (kept in the dual's buffer between the passes when the dual is wanted).
// Global objective (comm.cuh): thread 0 exchanges the shard's minimum and exponential sum with the peer ranks
// through NVLink-mapped slots, so that softMax is normalised over ALL shards as on one device.
__device__ __forceinline__ double hb_sx_exchange(const hb_scalar_xchg &X, int kind, double mine, unsigned long long e,
                                                 bool is_min) {
  const int par = (int)(e & 1);
  for (int p = 0; p < X.world; p++) *(volatile double *)&X.val[p][(par * 2 + kind) * 8 + X.rank] = mine;
  __threadfence_system();
  for (int p = 0; p < X.world; p++)
    if (p != X.rank)
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(X.flag[p] + kind * 8 + X.rank), "l"(e) : "memory");
  for (int q = 0; q < X.world; q++) {
    if (q == X.rank) continue;
    const unsigned long long *f = X.flag[X.rank] + kind * 8 + q;
    unsigned long long v;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v < e) {
        __nanosleep(64);
        if (clock64() - t0 > 20000000000ll) __trap();   // a peer that never arrives must not hang the GPU
      }
    } while (v < e);
  }
  const volatile double *mv = X.val[X.rank] + (par * 2 + kind) * 8;
  double r = mv[0];
  for (int q = 1; q < X.world; q++) r = is_min ? (mv[q] < r ? mv[q] : r) : r + mv[q];   // rank order on every rank
  return r;
}

// linear index -> element offset with 32-bit arithmetic (n < 2^31 elements)
__device__ __forceinline__ int hb_offset_of32(const hb_dims &d, uint32_t lin) {
  int off = 0;
#pragma unroll 1
  for (int i = d.rank - 1; i >= 0; i--) {
    const uint32_t s = (uint32_t)d.shape[i];
    const uint32_t q = lin / s, r = lin - q * s;
    off += (int)r * (int)d.stride[i];
    lin = q;
  }
  return off;
}

// REG: the thread's (at most NE) logits stay in registers through the three passes (n <= NE * 1024): every logit and
// label is read once, every exp evaluated once; otherwise the passes re-read memory (exp kept in the dual's buffer).
template <typename T, bool REG>
__global__ void __launch_bounds__(1024)
    hb_softmax_xent_one_kernel(const T *__restrict__ u, hb_dims ud, const T *__restrict__ ex, hb_dims ed, int64_t n,
                               T scale, T *__restrict__ loss, T *__restrict__ dl, hb_scalar_xchg X) {
  constexpr int NE = 8;
  __shared__ T smem[32];
  __shared__ double bcast;
  const int tid = threadIdx.x, nt = blockDim.x;
  unsigned long long e = 0;
  if (X.world > 1) e = *(volatile unsigned long long *)X.ctl + 1;
  T val[NE];
  T mn;
  if (REG) {
#pragma unroll
    for (int k = 0; k < NE; k++) {
      const int64_t i = tid + (int64_t)k * nt;
      val[k] = u[hb_offset_of32(ud, (uint32_t)(i < n ? i : 0))];   // out-of-range slots repeat element 0 (harmless for min)
    }
    mn = val[0];
#pragma unroll
    for (int k = 1; k < NE; k++) mn = val[k] < mn ? val[k] : mn;
  } else {
    mn = u[hb_offset_of(ud, tid < n ? tid : 0)];
    for (int64_t i = tid; i < n; i += nt) {
      T v = u[hb_offset_of(ud, i)];
      mn = v < mn ? v : mn;
    }
  }
  mn = hb_block_reduce<T>(mn, smem, true);
  if (X.world > 1) {
    if (tid == 0) bcast = hb_sx_exchange(X, 0, (double)mn, e, true);
    __syncthreads();
    mn = (T)bcast;
  }
  T s = T(0);
  if (REG) {
#pragma unroll
    for (int k = 0; k < NE; k++) {
      const int64_t i = tid + (int64_t)k * nt;
      val[k] = i < n ? exp(val[k] - mn) : T(0);
      s += val[k];
    }
  } else {
    for (int64_t i = tid; i < n; i += nt) {
      T ev = exp(u[hb_offset_of(ud, i)] - mn);
      if (dl) dl[i] = ev;
      s += ev;
    }
  }
  s = hb_block_reduce<T>(s, smem, false);
  if (X.world > 1) {
    __syncthreads();
    if (tid == 0) bcast = hb_sx_exchange(X, 1, (double)s, e, false);
    __syncthreads();
    s = (T)bcast;
  }
  const T rs = T(1) / s;
  T l = T(0);
  if (REG) {
#pragma unroll
    for (int k = 0; k < NE; k++) {
      const int64_t i = tid + (int64_t)k * nt;
      if (i < n) {
        T sm = rs * val[k];
        T t = ex[hb_offset_of32(ed, (uint32_t)i)];
        l += log(sm) * t;
        if (dl) dl[i] = scale * (sm - t);
      }
    }
  } else {
    for (int64_t i = tid; i < n; i += nt) {
      T sm = rs * (dl ? dl[i] : exp(u[hb_offset_of(ud, i)] - mn));
      T t = ex[hb_offset_of(ed, i)];
      l += log(sm) * t;
      if (dl) dl[i] = scale * (sm - t);
    }
  }
  l = hb_block_reduce<T>(l, smem, false);
  if (tid == 0) {
    *loss = scale * (-l);
    if (X.world > 1) {
      __threadfence();
      *(volatile unsigned long long *)X.ctl = e;
    }
  }
}

template <typename T>
static int softmax_xent_launch(const hb_array *logits, const hb_array *expected, const hb_dims &ud,
                               const hb_dims &ed, int64_t n, double scale, hb_array *loss, hb_array *dl) {
  const T *u = (const T *)hb_data(logits), *ex = (const T *)hb_data(expected);
  T *lp = (T *)hb_data(loss), *dp = dl ? (T *)hb_data(dl) : nullptr;
  hb_scalar_xchg X;
  memset(&X, 0, sizeof X);
  X.world = 1;
  const bool global = hb_comm_scalar_xchg(&X);
  if (global || n <= 16384) {
    if (n <= 8 * 1024) hb_softmax_xent_one_kernel<T, true><<<1, 1024, 0, g_hb.stream>>>(u, ud, ex, ed, n, (T)scale, lp, dp, X);
    else hb_softmax_xent_one_kernel<T, false><<<1, 1024, 0, g_hb.stream>>>(u, ud, ex, ed, n, (T)scale, lp, dp, X);
    HB_LAUNCHED();
    return HB_OK;
  }
  int nb = (int)((n + 511) / 512);
  if (nb > 128) nb = 128;
  if (nb < 1) nb = 1;
  void *scr;
  HB_TRY(hb_scratch((size_t)3 * 128 * sizeof(T), &scr));
  T *pmin = (T *)scr, *psum = pmin + 128, *ploss = psum + 128;
  hb_softmax_xent_phase<T, 0><<<nb, 256, 0, g_hb.stream>>>(u, ud, ex, ed, n, (T)scale, pmin, psum, ploss, lp, dp);
  HB_LAUNCHED();
  hb_softmax_xent_phase<T, 1><<<nb, 256, 0, g_hb.stream>>>(u, ud, ex, ed, n, (T)scale, pmin, psum, ploss, lp, dp);
  HB_LAUNCHED();
  hb_softmax_xent_phase<T, 2><<<nb, 256, 0, g_hb.stream>>>(u, ud, ex, ed, n, (T)scale, pmin, psum, ploss, lp, dp);
  HB_LAUNCHED();
  hb_softmax_xent_phase<T, 3><<<1, 32, 0, g_hb.stream>>>(u, ud, ex, ed, (int64_t)nb, (T)scale, pmin, psum, ploss, lp, dp);
  HB_LAUNCHED();
  return HB_OK;
}

extern "C" int hb_softmax_xent(const hb_array *logits, const hb_array *expected, double scale,
                               hb_array **loss_out, hb_array **dlogits_out) {
  hb_prof_scope prof_("softmax_xent", logits);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(logits && expected && loss_out, "null argument");
  HB_REQUIRE(hb_is_float(logits->dt) && logits->dt == expected->dt && hb_same_shape(logits, expected),
             "softmax-xent: logits %s vs expected %s", hb_shape_str(logits).c_str(), hb_shape_str(expected).c_str());
  int64_t n = hb_numel(logits);
  HB_REQUIRE(n > 0, "softmax-xent: empty logits");
  if (hb_lazy_on()) {
    int64_t one_ = 1;
    return hb_lrec(HB_L_GENERIC, "softmax_xent").in(logits).in(expected).farg(0, scale).out(loss_out, logits->dt, 0, &one_)
        .out(dlogits_out, logits->dt, logits->rank, logits->shape)
        .commit([=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_softmax_xent(in[0], in[1], n.fa[0], &o[0], n.out[1] ? &o[1] : nullptr); });
  }
  hb_array *loss, *dl = nullptr;
  int64_t one = 1;
  HB_TRY(hb_new_array(logits->dt, 0, &one, &loss));
  if (dlogits_out) {
    int st = hb_new_array(logits->dt, logits->rank, logits->shape, &dl);
    if (st != HB_OK) { hb_release(loss); return st; }
  }
  hb_dims ud, ed;
  ud.rank = ed.rank = logits->rank;
  for (int i = 0; i < HB_R; i++) {
    ud.shape[i] = ed.shape[i] = i < logits->rank ? logits->shape[i] : 1;
    ud.stride[i] = i < logits->rank ? logits->strides[i] : 0;
    ed.stride[i] = i < logits->rank ? expected->strides[i] : 0;
  }
  int st = logits->dt == HB_F64 ? softmax_xent_launch<double>(logits, expected, ud, ed, n, scale, loss, dl)
                                : softmax_xent_launch<float>(logits, expected, ud, ed, n, scale, loss, dl);
  if (st != HB_OK) {
    hb_release(loss);
    hb_release(dl);
    return st;
  }
  *loss_out = loss;
  if (dlogits_out) *dlogits_out = dl;
  return HB_OK;
}

// ------------------------------------------------- MNIST minibatch assembly
// Device twin of the input pipeline of example/MnistData.hs:134-171
// (readMnistData: glyph = fromIntegral pixel / 255, target = one-hot label;
// mkMnistDataBatchS: stack `batch` samples; the epoch's `shuffle` arrives as
// the `order` vector).  The raw IDX bytes stay resident in HBM (47 MB for the
// 60 000 training glyphs), so a training step moves no tensor data over PCIe.
template <typename T, typename I, int VEC>
__global__ void __launch_bounds__(256)
    hb_mnist_batch_kernel(const uint8_t *__restrict__ pix, const uint8_t *__restrict__ lab, const I *__restrict__ order,
                          int64_t n, int64_t start, int64_t batch, int hw, int classes, T *__restrict__ glyphs,
                          T *__restrict__ targets) {
  const int per = hw / VEC;
  const int64_t ng = batch * per, total = ng + batch * classes;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < ng) {
      int64_t b = i / per;
      int k = (int)(i - b * per) * VEC;
      int64_t id = order ? (int64_t)order[start + b] : start + b;
      const bool ok = id >= 0 && id < n;
      T out[VEC];
      if (VEC == 8) {
        uint2 raw = ok ? *reinterpret_cast<const uint2 *>(pix + id * hw + k) : make_uint2(0u, 0u);
#pragma unroll
        for (int e = 0; e < 4; e++) {
          out[e] = (T)((raw.x >> (8 * e)) & 0xffu) / (T)255;
          out[4 + e] = (T)((raw.y >> (8 * e)) & 0xffu) / (T)255;
        }
      } else {
#pragma unroll
        for (int e = 0; e < VEC; e++) out[e] = ok ? (T)pix[id * hw + k + e] / (T)255 : (T)0;
      }
      T *g = glyphs + b * hw + k;
#pragma unroll
      for (int e = 0; e < VEC; e++) g[e] = out[e];
    } else {
      int64_t j = i - ng, b = j / classes;
      int c = (int)(j - b * classes);
      int64_t id = order ? (int64_t)order[start + b] : start + b;
      const bool ok = id >= 0 && id < n;
      targets[j] = (ok && (int)lab[id] == c) ? (T)1 : (T)0;
    }
  }
}

extern "C" int hb_mnist_batch_into(const hb_array *pixels, const hb_array *labels, const hb_array *order,
                                   int64_t start, hb_array *glyphs_dst, hb_array *targets_dst) {
  hb_prof_scope prof_("mnist_batch_into", glyphs_dst);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(pixels && labels && glyphs_dst && targets_dst, "null argument");
  HB_REQUIRE(pixels->dt == HB_I8 && labels->dt == HB_I8 && pixels->rank >= 2 && labels->rank == 1 &&
                 pixels->shape[0] == labels->shape[0] && hb_contig(pixels) && hb_contig(labels),
             "mnist_batch: pixels [n,...] and labels [n] must be contiguous I8 (raw IDX bytes)");
  HB_REQUIRE(hb_is_float(glyphs_dst->dt) && targets_dst->dt == glyphs_dst->dt && hb_contig(glyphs_dst) &&
                 hb_contig(targets_dst) && glyphs_dst->rank == pixels->rank && targets_dst->rank == 2 &&
                 targets_dst->shape[0] == glyphs_dst->shape[0],
             "mnist_batch: glyphs [batch,...] and targets [batch,classes] must be contiguous and of one floating type");
  const int64_t n = pixels->shape[0], batch = glyphs_dst->shape[0];
  int64_t hw = 1;
  for (int i = 1; i < pixels->rank; i++) {
    HB_REQUIRE(pixels->shape[i] == glyphs_dst->shape[i], "mnist_batch: glyph extents %s vs %s",
               hb_shape_str(pixels).c_str(), hb_shape_str(glyphs_dst).c_str());
    hw *= pixels->shape[i];
  }
  const int64_t classes = targets_dst->shape[1];
  HB_REQUIRE(hw < (1 << 30) && classes < (1 << 30), "mnist_batch: extents too large");
  HB_REQUIRE(start >= 0, "mnist_batch: negative start");
  if (order) {
    HB_REQUIRE((order->dt == HB_I32 || order->dt == HB_I64) && order->rank == 1 && hb_contig(order),
               "mnist_batch: order must be a contiguous I32 or I64 vector");
    HB_REQUIRE(start + batch <= order->shape[0], "mnist_batch: batch [%lld, %lld) beyond the order vector (%lld)",
               (long long)start, (long long)(start + batch), (long long)order->shape[0]);
  } else {
    // file order: a batch that runs past the data would silently train on all-zero glyphs (the reference's
    // chunksOf just hands out a short last batch, MnistData.hs:134-171) -- refuse it instead
    HB_REQUIRE(start + batch <= n, "mnist_batch: batch [%lld, %lld) beyond the %lld samples", (long long)start,
               (long long)(start + batch), (long long)n);
  }
  if (hb_lazy_on()) {
    hb_lrec r(HB_L_EFFECT, "mnist_batch_into");
    r.in(pixels).in(labels).in(glyphs_dst).in(targets_dst);
    if (order) r.in(order);
    return r.commit([=](hb_lnode &n, hb_array *const *in, hb_array **o) {
      return hb_mnist_batch_into(in[0], in[1], n.in.size() > 4 ? in[4] : nullptr, start, in[2], in[3]);
    });
  }
  if (batch == 0 || (hw == 0 && classes == 0)) return HB_OK;
  const uint8_t *pp = (const uint8_t *)hb_data(pixels), *lp = (const uint8_t *)hb_data(labels);
  const bool vec = hw % 8 == 0 && ((uintptr_t)pp & 7) == 0 && ((uintptr_t)hb_data(glyphs_dst) & 15) == 0;
  const int64_t total = batch * (vec ? hw / 8 : hw) + batch * classes;
  const int blocks = hb_blocks_for(total, 256, g_hb.sm_count * 16);
#define MB_LAUNCH(T, I, V)                                                                                          \
  hb_mnist_batch_kernel<T, I, V><<<blocks, 256, 0, g_hb.stream>>>(pp, lp, order ? (const I *)hb_data(order) : nullptr, \
                                                                  n, start, batch, (int)hw, (int)classes,           \
                                                                  (T *)hb_data(glyphs_dst), (T *)hb_data(targets_dst))
#define MB_LAUNCH_T(T)                                                        \
  do {                                                                        \
    if (order && order->dt == HB_I64) { if (vec) MB_LAUNCH(T, int64_t, 8); else MB_LAUNCH(T, int64_t, 1); } \
    else { if (vec) MB_LAUNCH(T, int32_t, 8); else MB_LAUNCH(T, int32_t, 1); }                             \
  } while (0)
  if (glyphs_dst->dt == HB_F64) MB_LAUNCH_T(double);
  else MB_LAUNCH_T(float);
#undef MB_LAUNCH_T
#undef MB_LAUNCH
  HB_LAUNCHED();
  return HB_OK;
}

// ------------------------------------------------- recurrent layer body
// rnnMnistLayerS (example/MnistRnnShaped2.hs:55-69) after its two matmuls:
//   y = tanh ((u + v) + stretched bias),   u = wX x, v = wS s  : [width, batch]
// one pass instead of three maps and a broadcast view, and its adjoint
//   dz = (1 - y * y) * c  (the dual of tanh, same association as the generic rule),
//   dbias[i] = sum_j dz[i, j]
// one pass with the row sums reduced in a fixed order (deterministic).
template <typename T>
__global__ void __launch_bounds__(256)
    hb_tanh_affine_fwd_kernel(const T *__restrict__ u, const T *__restrict__ v, const T *__restrict__ bias, int64_t rows,
                              int64_t cols, T *__restrict__ y) {
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T z = v ? u[i] + v[i] : u[i];
    if (bias) z = z + bias[i / cols];
    if constexpr (sizeof(T) == 8) y[i] = hb_tanh_f64(z);   // the same function as the grouped products' epilogue
    else y[i] = tanh(z);
  }
}

// One CTA of 128 threads per row: 16-byte loads when the row length allows, row sum by warp shuffles + one
// shared-memory hop in a fixed order (deterministic).  ACC: dbias[row] += sum (the accumulator of an unrolled
// loop's bias cotangent -- one CTA owns a row, so the update is race-free and ordered by the launches).
template <typename T, bool ACC>
__global__ void __launch_bounds__(128)
    hb_tanh_affine_bwd_kernel(const T *__restrict__ y, const T *__restrict__ c, int64_t cols, T *__restrict__ dz,
                              T *__restrict__ dbias) {
  constexpr int V = 16 / (int)sizeof(T);
  __shared__ T red[4];
  const int64_t row = blockIdx.x;
  const T *yr = y + row * cols, *cr = c + row * cols;
  T *dr = dz + row * cols;
  T s = T(0);
  const bool vec = cols % V == 0 && (((uintptr_t)y | (uintptr_t)c | (uintptr_t)dz) & 15) == 0;
  if (vec) {
    using VT = typename std::conditional<sizeof(T) == 8, double2, float4>::type;
    const int64_t nv = cols / V;
#pragma unroll 4
    for (int64_t j = threadIdx.x; j < nv; j += 128) {
      VT yy = __ldcg(reinterpret_cast<const VT *>(yr) + j), cc = __ldcg(reinterpret_cast<const VT *>(cr) + j), d;
      T *yp = reinterpret_cast<T *>(&yy), *cp = reinterpret_cast<T *>(&cc), *dp = reinterpret_cast<T *>(&d);
#pragma unroll
      for (int k = 0; k < V; k++) {
        dp[k] = (T(1) - yp[k] * yp[k]) * cp[k];
        s += dp[k];
      }
      reinterpret_cast<VT *>(dr)[j] = d;
    }
  } else {
    for (int64_t j = threadIdx.x; j < cols; j += 128) {
      T yy = yr[j];
      T d = (T(1) - yy * yy) * cr[j];
      dr[j] = d;
      s += d;
    }
  }
  if (!dbias) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    T t = (red[0] + red[1]) + (red[2] + red[3]);
    dbias[row] = ACC ? dbias[row] + t : t;
  }
}

extern "C" int hb_tanh_affine_fwd(const hb_array *u, const hb_array *v, const hb_array *bias, hb_array **y_out) {
  hb_prof_scope prof_("tanh_affine_fwd", u);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(u && y_out, "null argument");
  HB_REQUIRE(u->rank == 2 && hb_is_float(u->dt) && (!v || (hb_same_shape(u, v) && u->dt == v->dt)),
             "tanh_affine: floating [width, batch] operands of one shape required");
  HB_REQUIRE(!bias || (bias->rank == 1 && bias->shape[0] == u->shape[0] && bias->dt == u->dt),
             "tanh_affine: bias must be [width] of the operands' type");
  if (hb_lazy_on()) {
    const bool hv_ = v != nullptr, hb_ = bias != nullptr;
    hb_lrec r(HB_L_GENERIC, "tanh_affine_fwd");
    r.in(u);
    if (hv_) r.in(v);
    if (hb_) r.in(bias);
    return r.out(y_out, u->dt, 2, u->shape).commit([=](hb_lnode &n, hb_array *const *in, hb_array **o) {
      return hb_tanh_affine_fwd(in[0], hv_ ? in[1] : nullptr, hb_ ? in[hv_ ? 2 : 1] : nullptr, &o[0]);
    });
  }
  const hb_array *uc = u, *vc = v, *bc = bias;
  hb_array *t1 = nullptr, *t2 = nullptr, *t3 = nullptr, *y = nullptr;
  int st = HB_OK;
  if (!hb_contig(u)) { st = hb_contiguous(u, &t1); uc = t1; }
  if (st == HB_OK && v && !hb_contig(v)) { st = hb_contiguous(v, &t2); vc = t2; }
  if (st == HB_OK && bias && !hb_contig(bias)) { st = hb_contiguous(bias, &t3); bc = t3; }
  if (st == HB_OK) st = hb_new_array(u->dt, 2, u->shape, &y);
  const int64_t rows = u->shape[0], cols = u->shape[1], n = rows * cols;
  if (st == HB_OK && n > 0) {
    const int blocks = hb_blocks_for(n, 256, g_hb.sm_count * 16);
    if (u->dt == HB_F64)
      hb_tanh_affine_fwd_kernel<double><<<blocks, 256, 0, g_hb.stream>>>(
          (const double *)hb_data(uc), vc ? (const double *)hb_data(vc) : nullptr,
          bc ? (const double *)hb_data(bc) : nullptr, rows, cols, (double *)hb_data(y));
    else
      hb_tanh_affine_fwd_kernel<float><<<blocks, 256, 0, g_hb.stream>>>(
          (const float *)hb_data(uc), vc ? (const float *)hb_data(vc) : nullptr,
          bc ? (const float *)hb_data(bc) : nullptr, rows, cols, (float *)hb_data(y));
    HB_LAUNCHED();
  }
  hb_release(t1); hb_release(t2); hb_release(t3);
  if (st != HB_OK) { hb_release(y); return st; }
  *y_out = y;
  return HB_OK;
}

// dz = (1 - y^2) * c; the row sums go to a fresh [rows] array (*dbias_out) or, acc != NULL, are ADDED to acc in
// place (acc is an accumulator the caller owns: hb_tanh_affine_bwd_acc)
static int tanh_affine_bwd_impl(const hb_array *y, const hb_array *c, hb_array *acc, hb_array **dz_out, hb_array **dbias_out) {
  const hb_array *yc = y, *cc = c;
  hb_array *t1 = nullptr, *t2 = nullptr, *dz = nullptr, *db = nullptr;
  int st = HB_OK;
  if (!hb_contig(y)) { st = hb_contiguous(y, &t1); yc = t1; }
  if (st == HB_OK && !hb_contig(c)) { st = hb_contiguous(c, &t2); cc = t2; }
  if (st == HB_OK) st = hb_new_array(y->dt, 2, y->shape, &dz);
  const int64_t rows = y->shape[0], cols = y->shape[1];
  // every row's CTA writes its sum: zero fill only for the degenerate empty row
  if (st == HB_OK && dbias_out) st = cols > 0 ? hb_new_array(y->dt, 1, y->shape, &db) : hb_zeros(y->dt, 1, y->shape, &db);
  if (st == HB_OK && rows > 0 && cols > 0) {
    void *bias = acc ? hb_data(acc) : db ? hb_data(db) : nullptr;
#define HB_TAB_LAUNCH(T, ACC)                                                                                  \
  hb_tanh_affine_bwd_kernel<T, ACC><<<(unsigned)rows, 128, 0, g_hb.stream>>>(                                  \
      (const T *)hb_data(yc), (const T *)hb_data(cc), cols, (T *)hb_data(dz), (T *)bias)
    if (y->dt == HB_F64) {
      if (acc) HB_TAB_LAUNCH(double, true);
      else HB_TAB_LAUNCH(double, false);
    } else {
      if (acc) HB_TAB_LAUNCH(float, true);
      else HB_TAB_LAUNCH(float, false);
    }
#undef HB_TAB_LAUNCH
    HB_LAUNCHED();
  }
  hb_release(t1); hb_release(t2);
  if (st != HB_OK) { hb_release(dz); hb_release(db); return st; }
  *dz_out = dz;
  if (dbias_out) *dbias_out = db;
  return HB_OK;
}

extern "C" int hb_tanh_affine_bwd(const hb_array *y, const hb_array *c, hb_array **dz_out, hb_array **dbias_out) {
  hb_prof_scope prof_("tanh_affine_bwd", y);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(y && c && dz_out, "null argument");
  HB_REQUIRE(y->rank == 2 && hb_same_shape(y, c) && y->dt == c->dt && hb_is_float(y->dt),
             "tanh_affine_bwd: two floating [width, batch] operands of one shape required");
  HB_REQUIRE(y->shape[0] < (1ll << 31), "tanh_affine_bwd: too many rows");
  if (hb_lazy_on())
    return hb_lrec(HB_L_GENERIC, "tanh_affine_bwd").in(y).in(c).out(dz_out, y->dt, 2, y->shape).out(dbias_out, y->dt, 1, y->shape)
        .commit([=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_tanh_affine_bwd(in[0], in[1], &o[0], n.out[1] ? &o[1] : nullptr); });
  return tanh_affine_bwd_impl(y, c, nullptr, dz_out, dbias_out);
}

// dz as above; *acc_out = dbias_acc + row sums.  Value semantics like hb_matmul2_acc: the accumulator's storage is
// updated in place (and handed back retained) only when the caller holds the single reference to it.
extern "C" int hb_tanh_affine_bwd_acc(const hb_array *y, const hb_array *c, hb_array *dbias_acc, hb_array **dz_out,
                                      hb_array **acc_out) {
  hb_prof_scope prof_("tanh_affine_bwd_acc", y);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(y && c && dbias_acc && dz_out && acc_out, "null argument");
  HB_REQUIRE(y->rank == 2 && hb_same_shape(y, c) && y->dt == c->dt && hb_is_float(y->dt),
             "tanh_affine_bwd_acc: two floating [width, batch] operands of one shape required");
  HB_REQUIRE(y->shape[0] < (1ll << 31), "tanh_affine_bwd_acc: too many rows");
  HB_REQUIRE(dbias_acc->rank == 1 && dbias_acc->shape[0] == y->shape[0] && dbias_acc->dt == y->dt,
             "tanh_affine_bwd_acc: the accumulator must be a [width] array of the operands' type");
  if (hb_lazy_on())
    return hb_lrec(HB_L_GENERIC, "tanh_affine_bwd_acc").in(y).in(c).in(dbias_acc).out(dz_out, y->dt, 2, y->shape)
        .out(acc_out, y->dt, 1, y->shape)
        .commit([=](hb_lnode &, hb_array *const *in, hb_array **o) {
          return hb_tanh_affine_bwd_acc(in[0], in[1], in[2], &o[0], &o[1]);
        });
  const bool unique = dbias_acc->rc.load() == 1 && dbias_acc->st && dbias_acc->st->rc.load() == 1 && hb_contig(dbias_acc) &&
                      dbias_acc->st != y->st && dbias_acc->st != c->st;
  if (unique) {
    HB_TRY(tanh_affine_bwd_impl(y, c, dbias_acc, dz_out, nullptr));
    hb_retain(dbias_acc);
    *acc_out = dbias_acc;
    return HB_OK;
  }
  hb_array *db = nullptr;
  HB_TRY(tanh_affine_bwd_impl(y, c, nullptr, dz_out, &db));
  int st = hb_binary(HB_ADD, dbias_acc, db, acc_out);
  hb_release(db);
  if (st != HB_OK) { hb_release(*dz_out); *dz_out = nullptr; }
  return st;
}

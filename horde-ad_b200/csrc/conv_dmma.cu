// conv_dmma.cu -- shared-memory-staged, im2col-free FP64 convolution kernels
// on the DMMA tensor pipe (mma.sync m8n8k4 f64 -> SASS DMMA; tcgen05 has no
// FP64 kind, so this is the FP64 tensor path of sm_100a).
//
// Reference semantics: conv2dPreservingS (CommonShapedOps.hs:136-157) and its
// handwritten adjoints conv2dPreserving_dInp / _dKrn
// (TestConvQuickCheck.hs:290-349).
//
// Why not the generic strided-GEMM engine (contract.cu): ncu showed it LSU/L1
// bound (l1tex 88 %, FP64 pipe 34 %), because every operand element is
// gathered through offset tables.  Here the INPUT TILE ITSELF (with its halo,
// zero-filled at the image border) is staged in shared memory with cp.async
// and the DMMA fragments are read from SHIFTED addresses of that tile -- the
// (kh, kw) taps are address offsets, no patch matrix exists anywhere:
//
//   spatial kernel (forward and dInp):
//     out[n,q,y,x] = sum_{p,i,j} in[n,p,y+i+oy,x+j+ox] * Wt[p,(i,j),q]
//     M = 8 consecutive output positions, N = 8 output channels,
//     K = 4 input channels (or 4 kw taps when there are < 4 input channels)
//   dKrn kernel:
//     dK[o,c,kh,kw] = sum_{n,y,x} dB[n,o,y,x] * A[n,c,y+kh,x+kw]
//     M = 8 output channels, N = 8 input channels (or 8 kw taps),
//     K = 4 positions; split-K over CTAs with a fixed-order second pass
//     (deterministic).
//
// Shared-memory strides are chosen = 4 (mod 8) doubles so that every 64-bit
// fragment load of a half-warp hits 16 distinct bank pairs.  CTAs are
// persistent (grid = #SMs), two cp.async stages, one CTA per SM (up to 227 KB).
#include "common.cuh"
#include <algorithm>
#include <cuda.h>  // CUtensorMap (the encoder is resolved through cudaGetDriverEntryPoint, libcuda is not linked)

namespace {

constexpr int kSmemMax = 227 * 1024;
constexpr int kThreads = 256;

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(void *s, const void *g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp8(void *s, const void *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---- TMA + mbarrier (sm_90+): one thread hands a whole [NI][CC][TR][RS] box of the input tensor to the
// copy engine; rows and columns outside the image, channels beyond P and images beyond N arrive as zeros
// (out-of-bounds fill), so the zero padding of conv2dPreservingS costs no instruction at all
__device__ __forceinline__ void mbar_init(void *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, int parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }\n"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  }
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tm, void *bar, int x, int y, int c, int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(c), "r"(n)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, int bytes, void *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Geometry of one staged input tile: NI images x TR rows x RS columns per
// channel (channel stride CS), holding output rows [y0, y0+R) and their halo.
struct TileGeo {
  int H, W, KH, KW;
  int R, NI;
  int TR, RS, IS, CS;
  int oy, ox;  // tile row t holds input row y0+oy+t; input col x sits at tile col x-ox
  int bands;   // ceil(H / R)
  int nblocks; // ceil(N / NI); item it = band * nblocks + image block (band-major: every CTA sees all bands)
  int vec;     // rows may be moved in 16-byte pieces
  int cpr, cshift;  // pieces per row and log2 of its power-of-two padding
  int dc8, di8;     // 8 / NI and 8 % NI: plane step of a warp (8 warps) in (channel, image)
};

__device__ __forceinline__ void cp16s(unsigned s, const void *g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp8s(unsigned s, const void *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void sts_zero16(unsigned s) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %1};\n" ::"r"(s), "d"(0.0) : "memory");
}
__device__ __forceinline__ void sts_zero8(unsigned s) {
  asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(s), "d"(0.0) : "memory");
}

// channels [c0, c0+CC) of images [n0, n0+NI), input rows of band y0 -> dst.
// A warp owns whole (channel, image) planes.  A lane keeps one piece column
// (pieces per row padded to a power of two) and walks down the rows, so the
// hot loop is two range compares, one cp.async (or zero store) and three adds:
// ncu showed the former per-slot index arithmetic costing as many issue slots
// per item as a third of the DMMA loop.
template <bool VEC>
__device__ __forceinline__ void load_in_tile_t(double *dst, const double *__restrict__ in, const TileGeo &g, int CC,
                                               int c0, int P, int n0, int N, int y0, int64_t sn, int64_t sp,
                                               int tid, int cw, int iw) {
  constexpr int PB = VEC ? 16 : 8, PD = VEC ? 2 : 1;  // bytes / doubles per piece
  const int warp = tid >> 5, lane = tid & 31;
  const int planes = CC * g.NI;
  // rows: lane (tr0, ch) with tstep rows per pass; rows wider than 32 pieces: a lane strides along the row
  const bool wide = g.cshift > 5;
  const int ch0 = wide ? lane : (lane & ((1 << g.cshift) - 1));
  const int tr0 = wide ? 0 : (lane >> g.cshift), tstep = wide ? 1 : (32 >> g.cshift);
  const unsigned sbase = smem_u32(dst);
  const int ybase = y0 + g.oy;
  int c = cw, img = iw;  // = (warp / NI, warp % NI), computed once per kernel
  for (int pl = warp; pl < planes; pl += kThreads / 32) {
    const bool okp = c0 + c < P && n0 + img < N;
    const unsigned d0 = sbase + (unsigned)((c * g.CS + img * g.IS - g.ox + tr0 * g.RS + PD * ch0) * 8);
    const double *s0 = in + (n0 + img) * sn + (c0 + c) * sp + (int64_t)(ybase + tr0) * g.W + PD * ch0;
    for (int cc = ch0; cc < g.cpr; cc += 32) {
      unsigned d = d0 + (unsigned)((cc - ch0) * PB);
      const double *sp_ = s0 + (cc - ch0) * PD;
      for (int t = tr0; t < g.TR; t += tstep) {
        const bool ok = okp && (unsigned)(ybase + t) < (unsigned)g.H;
        if (VEC) {
          if (ok) cp16s(d, sp_);
          else sts_zero16(d);
        } else {
          if (ok) cp8s(d, sp_);
          else sts_zero8(d);
        }
        d += (unsigned)(tstep * g.RS * 8);
        sp_ += tstep * g.W;
      }
    }
    img += g.di8;
    c += g.dc8;
    if (img >= g.NI) { img -= g.NI; c++; }
  }
}
__device__ __forceinline__ void load_in_tile(double *dst, const double *__restrict__ in, const TileGeo &g, int CC,
                                             int c0, int P, int n0, int N, int y0, int64_t sn, int64_t sp,
                                             int tid, int cw, int iw) {
  if (g.vec) load_in_tile_t<true>(dst, in, g, CC, c0, P, n0, N, y0, sn, sp, tid, cw, iw);
  else load_in_tile_t<false>(dst, in, g, CC, c0, P, n0, N, y0, sn, sp, tid, cw, iw);
}

// contiguous n doubles (n even, both 16-byte aligned)
__device__ __forceinline__ void load_linear(double *dst, const double *__restrict__ src, int n, int tid) {
  for (int i = tid * 2; i < n; i += kThreads * 2) cp16(dst + i, src + i);
}

// ===================================================== spatial (fwd / dInp)
struct SpP {
  alignas(64) CUtensorMap tmap;  // input [N][P][H][W], box [NI][CC][TR][RS] (tma != 0)
  int tma, mbar;                 // TMA staging; offset (doubles) of the two stage mbarriers
  int xs;                        // TMA boxes start on a 16-byte boundary: tile column 0 is input column ox - xs
  const double *in, *wp;
  double *out;
  TileGeo g;
  int N, P, Q;
  int CC, NCH;       // input channels per stage, number of stages per item
  int QT, TS, WCS;   // output-channel tile, tap stride, channel stride of the packed weights
  int Ppad;          // NCH * CC
  int items;         // ceil(N / NI) * bands
  int in_doubles, stage_doubles;
  int wfix;          // one channel chunk: the packed weights are staged once per CTA at smem + wfix (doubles), not per stage
  int tail;          // doubles used by the stages (and the fixed weights)
  int64_t sn, sp;
  int KWP;           // KW rounded up to 4 (TAPK)
  // fused tail of convMnistLayerS: out is never written; instead
  //   pool_out = maxPool 2x2/2 (relu (conv + bias)),  code_out = adjoint code
  int pool;
  const double *bias;
  double *pool_out;
  uint8_t *code_out;
  int qg, OPS;       // channels per pooling round, position stride of the output tile
  // Zero-padding taps are not computed: slot s = i * WM + wm of the warp grid holds
  // fragment perm[s]; fragments are dealt so that within a warp the ones whose
  // rows reach furthest into the padding come last, and for tap row kh only the
  // first nk[band & 3][wm][kh] fragments of warp-row wm have any in-image input row.
  unsigned char perm[64];
  unsigned char nk[4][8][16];
};

// Packs K into Wt[qt][p][tap][q] (tap stride TS, channel stride WCS), zero
// padded in p, q and (TAPK) kw.  flip: dInp (Wt[p=o][i',j'][q=c] =
// K[o,c,KH-1-i',KW-1-j']); else forward (Wt[p=c][i,j][q=o] = K[o,c,i,j]).
__global__ void pack_sp_weights(const double *__restrict__ K, double *__restrict__ wp, int KH, int KW, int KWe, int P,
                                int Q, int Ppad, int QT, int TS, int WCS, int nqt, int flip, int64_t k_so,
                                int64_t k_sc, int64_t k_si, int64_t k_sj) {
  int64_t total = (int64_t)nqt * Ppad * WCS;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int w = (int)(idx % WCS);
    int64_t r = idx / WCS;
    int p = (int)(r % Ppad), qt = (int)(r / Ppad);
    int tap = w / TS, ql = w - tap * TS;
    int i = tap / KWe, j = tap - i * KWe;
    int q = qt * QT + ql;
    double v = 0.0;
    if (i < KH && j < KW && ql < QT && p < P && q < Q) {
      if (flip) v = K[p * k_so + q * k_sc + (KH - 1 - i) * k_si + (KW - 1 - j) * k_sj];
      else v = K[q * k_so + p * k_sc + i * k_si + j * k_sj];
    }
    wp[idx] = v;
  }
}

// One stage of the tap-k spatial kernel (k = 4 kw taps of one channel) for a
// warp that owns exactly NF fragments.
template <int NF, int MF, int NQ>
__device__ __forceinline__ void sp_stage_tap(double (&acc)[MF][NQ][2], const int (&aoff)[MF], const char *Sb,
                                             const char *Wb, const SpP &p) {
  const TileGeo &g = p.g;
  const int ng = p.KWP >> 2;
  for (int c = 0; c < p.CC; c++) {
    for (int kh = 0; kh < g.KH; kh++) {
      const char *ia = Sb + (c * g.CS + kh * g.RS) * 8;
      const char *wb = Wb + (c * p.WCS + kh * p.KWP * p.TS) * 8;
      for (int g4 = 0; g4 < ng; g4++) {
        // taps beyond KW exist only as padding of the k dimension: they contribute an
        // exact zero whatever the data holds there (0 * Inf would be NaN)
        const bool tapv = g4 * 4 + (int)(threadIdx.x & 3) < g.KW;
        double b[NQ];
#pragma unroll
        for (int j = 0; j < NQ; j++) b[j] = *reinterpret_cast<const double *>(wb + j * 64);
#pragma unroll
        for (int i = 0; i < NF; i++) {
          double a = *reinterpret_cast<const double *>(ia + aoff[i]);
          a = tapv ? a : 0.0;
#pragma unroll
          for (int j = 0; j < NQ; j++) dmma884(acc[i][j][0], acc[i][j][1], a, b[j]);
        }
        ia += 32;
        wb += 4 * p.TS * 8;
      }
    }
  }
}

// One tap row (kh fixed, all kw) of the channel-k spatial kernel over the first
// NA fragments of the warp.
template <int NA, int MF, int NQ>
__device__ __forceinline__ void sp_row(double (&acc)[MF][NQ][2], const int (&aoff)[MF], const char *ia,
                                       const char *wb, int tsb, int KW) {
  for (int kw = 0; kw < KW; kw++) {
    double b[NQ];
#pragma unroll
    for (int j = 0; j < NQ; j++) b[j] = *reinterpret_cast<const double *>(wb + j * 64);
#pragma unroll
    for (int i = 0; i < NA; i++) {
      double a = *reinterpret_cast<const double *>(ia + aoff[i]);
#pragma unroll
      for (int j = 0; j < NQ; j++) dmma884(acc[i][j][0], acc[i][j][1], a, b[j]);
    }
    ia += 8;
    wb += tsb;
  }
}

// One stage of the channel-k spatial kernel (k = 4 input channels of one tap).
// The number of fragments with an in-image row under tap row kh comes from
// p.nk; predicated-off DMMAs would still occupy the tensor pipe, so the count
// selects an exactly-sized instantiation by a warp-uniform switch.
template <int MF, int NQ>
__device__ __forceinline__ void sp_stage_chan(double (&acc)[MF][NQ][2], const int (&aoff)[MF], const char *Sb,
                                              const char *Wb, const SpP &p, int band, int wm) {
  const TileGeo &g = p.g;
  const int tsb = p.TS * 8;
  for (int c4 = 0; c4 < p.CC; c4 += 4) {
    for (int kh = 0; kh < g.KH; kh++) {
      const char *ia = Sb + (c4 * g.CS + kh * g.RS) * 8;
      const char *wb = Wb + (c4 * p.WCS + kh * g.KW * p.TS) * 8;
      switch (p.nk[band & 3][wm][kh]) {
#define SP_ROW_CASE(k) \
  case k: if (MF >= k) sp_row<(MF >= k ? k : MF), MF, NQ>(acc, aoff, ia, wb, tsb, g.KW); break;
        SP_ROW_CASE(7) SP_ROW_CASE(6) SP_ROW_CASE(5) SP_ROW_CASE(4) SP_ROW_CASE(3) SP_ROW_CASE(2) SP_ROW_CASE(1)
#undef SP_ROW_CASE
        default: break;
      }
    }
  }
}

template <int WM, int WN, int MF, int NQ, bool TAPK, int MINB, bool POOL>
__global__ void __launch_bounds__(kThreads, MINB) conv_sp_kernel(const __grid_constant__ SpP p) {
  extern __shared__ __align__(128) double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Warp w runs on scheduler w % 4, so with WM = 4 the two warps of a scheduler (w, w + 4) would own the same
  // warp row: 25 fragments over 4 rows is 7/6/6/6, i.e. 14 fragment columns on one tensor pipe and 12 on the
  // others.  The second channel column takes its rows rotated by two: 13/12/13/12.
  // (Rotating the rows of the second co-resident CTA of an SM in the one-column layout, where row 0 is the heavy
  // one in every CTA, was measured and changes nothing: 1.416 vs 1.414 ms per CNN step.)
  const int wn = warp / WM, wm = (warp % WM + ((WN == 2 && WM == 4) ? 2 * wn : 0)) % WM;
  const int fr = lane >> 2, fk = lane & 3;
  const TileGeo &g = p.g;
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem + p.mbar);  // [2], TMA staging only

  for (int i = tid; i < 2 * p.stage_doubles; i += kThreads) smem[i] = 0.0;
  if (!TAPK && p.tma && tid == 0) {
    mbar_init(mbar, 1);
    mbar_init(mbar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  const int RW = g.R * g.W, npos = g.NI * RW, HW = g.H * g.W;
  const int F = (npos + 7) >> 3;
  if (POOL) {  // window table: first position, image, pooled row in the band, pooled output offset
    int4 *wl = reinterpret_cast<int4 *>(smem + p.tail + p.qg * p.OPS);
    const int Wo = g.W >> 1, Ho = g.H >> 1, wpi = (g.R >> 1) * Wo, nwin = g.NI * wpi;
    for (int w = tid; w < nwin; w += kThreads) {
      int img = w / wpi, rem = w - img * wpi;
      int i = rem / Wo, j = rem - i * Wo;
      wl[w] = make_int4(img * RW + 2 * i * g.W + 2 * j, img, i, img * p.Q * Ho * Wo + i * Wo + j);
    }
    __syncthreads();
  }
  // slot i * WM + wm holds fragment perm[slot] (dealt over the warps of a column: balanced)
  int nf = (F - wm + WM - 1) / WM;
  nf = nf < 0 ? 0 : (nf > MF ? MF : nf);
  // per fragment: byte offset of this lane's A element inside a stage (k part
  // included), the output offset relative to the item's base, and (img, y)
  int aoff[MF], orel[MF], pimy[MF];
#pragma unroll
  for (int i = 0; i < MF; i++) {
    int pidx = p.perm[(i * WM + wm) & 63] * 8 + fr;
    int off = 0, o = 0, iy = 0x7fff0000;
    if (pidx < npos) {
      int img = pidx / RW, rem = pidx - img * RW;
      int y = rem / g.W, x = rem - y * g.W;
      off = img * g.IS + y * g.RS + x + p.xs;
      o = img * p.Q * HW + y * g.W + x;
      iy = (img << 16) | y;
    }
    aoff[i] = (off + (TAPK ? fk : fk * g.CS)) * 8;
    orel[i] = o;
    pimy[i] = iy;
  }
  double acc[MF][NQ][2];
#pragma unroll
  for (int i = 0; i < MF; i++)
#pragma unroll
    for (int j = 0; j < NQ; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int qt = blockIdx.y;
  const int my_items = (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int T = my_items * p.NCH;
  const double *wbase = p.wp + (int64_t)qt * p.Ppad * p.WCS;
  // this lane's B element inside a weight block (k part included), bytes
  const int wlane = (wn * NQ * 8 + fr + (TAPK ? fk * p.TS : fk * p.WCS)) * 8;

  const int cw = warp / g.NI, iw = warp - cw * g.NI;
  // stage t = (item, channel chunk); both sides of the pipeline step through them without divisions
  int is_ch = 0, is_n0, is_y0;   // next stage to issue
  int cs_ch = 0, cs_n0, cs_y0, cs_band;  // stage being computed
  {
    int band = (int)blockIdx.x / g.nblocks, nblk = (int)blockIdx.x - band * g.nblocks;
    is_n0 = cs_n0 = nblk * g.NI;
    is_y0 = cs_y0 = band * g.R;
    cs_band = band;
  }
  int is_it = blockIdx.x, cs_it = blockIdx.x;
  auto issue = [&](int t) {
    double *st = smem + (t & 1) * p.stage_doubles;
    if (!TAPK && p.tma) {
      if (tid == 0) {  // one thread, a handful of instructions: the copy engine generates every address
        const int wbytes = p.wfix ? 0 : p.CC * p.WCS * 8;
        mbar_expect_tx(mbar + (t & 1), p.in_doubles * 8 + wbytes);
        tma_load_4d(st, &p.tmap, mbar + (t & 1), g.ox - p.xs, is_y0 + g.oy, is_ch * p.CC, is_n0);
        if (wbytes) bulk_load(st + p.in_doubles, wbase + (int64_t)is_ch * p.CC * p.WCS, wbytes, mbar + (t & 1));
      }
    } else {
      load_in_tile(st, p.in, g, p.CC, is_ch * p.CC, p.P, is_n0, p.N, is_y0, p.sn, p.sp, tid, cw, iw);
      if (!p.wfix) load_linear(st + p.in_doubles, wbase + (int64_t)is_ch * p.CC * p.WCS, p.CC * p.WCS, tid);
      cp_commit();
    }
    if (++is_ch == p.NCH) {
      is_ch = 0;
      is_it += gridDim.x;
      int band = is_it / g.nblocks, nblk = is_it - band * g.nblocks;
      is_n0 = nblk * g.NI;
      is_y0 = band * g.R;
    }
  };

  // weights of all channel chunks staged once per CTA (joins the first commit group)
  if (p.wfix && T > 0) load_linear(smem + p.wfix, wbase, p.Ppad * p.WCS, tid);
  if (!TAPK && p.tma) {  // the resident weights are the only cp.async traffic of the TMA path
    cp_commit();
    cp_wait<0>();
    __syncthreads();
  }
  if (T > 0) issue(0);
  for (int t = 0; t < T; t++) {
    if (!TAPK && p.tma) {
      // stage (t + 1) & 1 was last read in iteration t - 1; every warp has passed that iteration's barrier
      if (t + 1 < T) issue(t + 1);
      mbar_wait(mbar + (t & 1), (t >> 1) & 1);
    } else {
      if (t + 1 < T) {
        issue(t + 1);
        cp_wait<1>();
      } else {
        cp_wait<0>();
      }
      __syncthreads();
    }
    const char *Sb = reinterpret_cast<const char *>(smem + (t & 1) * p.stage_doubles);
    const char *Wl = (p.wfix ? reinterpret_cast<const char *>(smem + p.wfix + cs_ch * p.CC * p.WCS) : Sb + p.in_doubles * 8) + wlane;
    // predicated-off DMMAs still occupy the tensor pipe (ncu: dmma cycles =
    // issued slots, not useful ones), so fragment counts are template
    // parameters selected by a warp-uniform switch, never a predicate
    if (TAPK) {
      switch (nf) {
        case 7: if (MF >= 7) sp_stage_tap<(MF >= 7 ? 7 : MF), MF, NQ>(acc, aoff, Sb, Wl, p); break;
        case 6: if (MF >= 6) sp_stage_tap<(MF >= 6 ? 6 : MF), MF, NQ>(acc, aoff, Sb, Wl, p); break;
        case 5: if (MF >= 5) sp_stage_tap<(MF >= 5 ? 5 : MF), MF, NQ>(acc, aoff, Sb, Wl, p); break;
        case 4: if (MF >= 4) sp_stage_tap<(MF >= 4 ? 4 : MF), MF, NQ>(acc, aoff, Sb, Wl, p); break;
        case 3: sp_stage_tap<3, MF, NQ>(acc, aoff, Sb, Wl, p); break;
        case 2: sp_stage_tap<2, MF, NQ>(acc, aoff, Sb, Wl, p); break;
        case 1: sp_stage_tap<1, MF, NQ>(acc, aoff, Sb, Wl, p); break;
        default: break;
      }
    } else {
      sp_stage_chan<MF, NQ>(acc, aoff, Sb, Wl, p, cs_band, wm);
    }
    if (++cs_ch == p.NCH) {
      // epilogue of this item: C fragment row = position, cols = 2 output channels
      const int n0 = cs_n0, y0 = cs_y0;
      {  // the item after this one
        cs_ch = 0;
        cs_it += gridDim.x;
        const int it = cs_it;
        cs_band = it / g.nblocks;
        cs_n0 = (it - cs_band * g.nblocks) * g.NI;
        cs_y0 = cs_band * g.R;
      }
      const int q0 = qt * p.QT + wn * NQ * 8 + fk * 2;
      if (!POOL) {
        double *ob = p.out + ((int64_t)n0 * p.Q + q0) * HW + (int64_t)y0 * g.W;
#pragma unroll
        for (int i = 0; i < MF; i++) {
          if (i < nf) {
            int img = pimy[i] >> 16, y = pimy[i] & 0xffff;
            bool okp = n0 + img < p.N && y0 + y < g.H;
#pragma unroll
            for (int j = 0; j < NQ; j++) {
#pragma unroll
              for (int e = 0; e < 2; e++) {
                if (okp && q0 + j * 8 + e < p.Q) ob[orel[i] + (int64_t)(j * 8 + e) * HW] = acc[i][j][e];
                acc[i][j][e] = 0.0;
              }
            }
          }
        }
      } else {
        // bias + relu + 2x2 max-pool through a shared [channel][position] tile,
        // qg channels per round; the pre-activation never goes to HBM
        double *ot = smem + p.tail;
        const int4 *wl = reinterpret_cast<const int4 *>(ot + p.qg * p.OPS);
        const int Wo = g.W >> 1, Ho = g.H >> 1, nwin = npos >> 2;
        double bv[NQ][2];
#pragma unroll
        for (int j = 0; j < NQ; j++)
#pragma unroll
          for (int e = 0; e < 2; e++) bv[j][e] = (p.bias && q0 + j * 8 + e < p.Q) ? p.bias[q0 + j * 8 + e] : 0.0;
        const int64_t obase = ((int64_t)n0 * p.Q + qt * p.QT) * Ho * Wo + (int64_t)(y0 >> 1) * Wo;
        for (int r0 = 0; r0 < p.QT; r0 += p.qg) {
#pragma unroll
          for (int i = 0; i < MF; i++) {
            if (i < nf) {
              int pidx = p.perm[(i * WM + wm) & 63] * 8 + fr;
#pragma unroll
              for (int j = 0; j < NQ; j++) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                  int ql = wn * NQ * 8 + j * 8 + fk * 2 + e - r0;
                  if (ql >= 0 && ql < p.qg && pidx < npos) ot[ql * p.OPS + pidx] = acc[i][j][e] + bv[j][e];
                }
              }
            }
          }
          __syncthreads();
          // all threads pool: thread (tid % hw) takes windows, tid / hw picks the channel lane
          const int hw = nwin <= 64 ? 64 : (nwin <= 128 ? 128 : 256), lanes = kThreads / hw;
          const int wbeg = tid % hw, qoff = tid / hw;
          for (int ql = qoff; ql < p.qg; ql += lanes) {
            int q = qt * p.QT + r0 + ql;
            if (q >= p.Q) break;
            const double *oq = ot + ql * p.OPS;
            for (int w = wbeg; w < nwin; w += hw) {
              const int4 wv = wl[w];
              const int p00 = wv.x;
              if (n0 + wv.y >= p.N || y0 + 2 * wv.z >= g.H) continue;
              double2 r0v = *reinterpret_cast<const double2 *>(oq + p00);
              double2 r1v = *reinterpret_cast<const double2 *>(oq + p00 + g.W);
              double v[4] = {r0v.x, r0v.y, r1v.x, r1v.y};
              double best = 0.0, bestpre = 0.0;
              int bi = 0;
#pragma unroll
              for (int k = 0; k < 4; k++) {
                double rl = (v[k] <= 0.0 ? 0.0 : 1.0) * v[k];
                if (k == 0 || rl > best) { best = rl; bestpre = v[k]; bi = k; }
              }
              int64_t o = obase + (int64_t)(r0 + ql) * Ho * Wo + wv.w;
              p.pool_out[o] = best;
              p.code_out[o] = (uint8_t)(bi | (bestpre > 0.0 ? 0x80 : 0));
            }
          }
          __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < MF; i++)
#pragma unroll
          for (int j = 0; j < NQ; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
      }
    }
    __syncthreads();
  }
}

// ================================================================== dKrn
struct DkP {
  alignas(64) CUtensorMap tmap;  // A [N][Ci][H][W], box [1][CC][TR][RS] (tma != 0)
  int tma, mbar;                 // TMA staging; offset (doubles) of the two stage mbarriers
  const double *A, *dB;
  double *part;  // [split][Co][Ci][KH][KW]
  TileGeo g;
  int N, Ci, Co;
  int CC;          // input channels per CTA (gridDim.y chunks)
  int NCF;         // CHAN: CC / 8 column fragments; TAPW: kw groups of 8
  int npairs;      // (tap, cf) pairs (CHAN) or (c, kh, kwg) triples (TAPW) per CTA
  int WN, WK, NPW; // warps along pairs / along K; pairs per warp
  int PS, npos4;   // dB tile row stride; positions rounded up to 4
  int items;
  int a_doubles, stage_doubles;
  int64_t a_sn, a_sc, b_sn, b_so;
  int dbvec;
  // pooled cotangent (adjoint of the fused bias + relu + 2x2 max-pool): dB is
  // never materialised; dy [N,Co,H/2,W/2] and its code bytes are staged and
  // expanded into the dB tile in shared memory
  int pooled;
  const double *dy;
  const uint8_t *code;
  int PW, PPS;     // pooled elements per (o, image) plane of a band, and their staging stride
  int dyvec;
  // one image per stage, channel columns: positions whose row y has y + kh >= H only meet
  // zero padding under tap row kh, so the K loop runs in row segments with the pair count
  // shrinking (pairs are dealt round-robin over the warps, ascending in kh within a warp)
  int skip;
};

constexpr int kNP = 9;  // max column pairs per warp

// One stage of the dKrn kernel for a warp that owns exactly NPW column pairs.
template <int NPW, int MO>
__device__ __forceinline__ void dk_stage(double (&acc)[MO][kNP][2], const int (&pair_off)[kNP], const char *As,
                                         const double *Bs, const int *lutk, int kbeg, int kend, int kstep,
                                         const DkP &p) {
  for (int k0 = kbeg; k0 < kend; k0 += kstep) {
    const char *ap = As + lutk[k0];
    double a[MO];
#pragma unroll
    for (int m = 0; m < MO; m++) a[m] = Bs[m * 8 * p.PS + k0];
#pragma unroll
    for (int j = 0; j < NPW; j++) {
      double b = *reinterpret_cast<const double *>(ap + pair_off[j]);
#pragma unroll
      for (int m = 0; m < MO; m++) dmma884(acc[m][j][0], acc[m][j][1], a[m], b);
    }
  }
}

template <int WM, int MO, bool TAPW>
__global__ void __launch_bounds__(kThreads, 2) conv_dk_kernel(const __grid_constant__ DkP p) {
  extern __shared__ __align__(128) double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp % WM, wn = (warp / WM) % p.WN, wk = warp / (WM * p.WN);
  const int fr = lane >> 2, fk = lane & 3;
  const TileGeo &g = p.g;
  // pooled mode: stages hold [A tile | pooled dy staging]; one expanded dB tile follows
  double *dbx = smem + 2 * p.stage_doubles;
  const int otile_rows = WM * MO * 8;
  int *lut = reinterpret_cast<int *>(smem + 2 * p.stage_doubles + (p.pooled ? otile_rows * p.PS : 0));
  int *lutp = lut + p.npos4;  // pooled index -> first position of its window inside one image band

  for (int i = tid; i < 2 * p.stage_doubles + (p.pooled ? otile_rows * p.PS : 0); i += kThreads) smem[i] = 0.0;
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem + p.mbar);  // [2], TMA staging only
  if (!TAPW && p.tma && tid == 0) {
    mbar_init(mbar, 1);
    mbar_init(mbar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  const int RW = g.R * g.W, npos = g.NI * RW;
  if (p.pooled) {
    const int Wo = g.W >> 1;
    for (int i = tid; i < p.PW; i += kThreads) {
      int pi = i / Wo, pj = i - pi * Wo;
      lutp[i] = 2 * pi * g.W + 2 * pj;
    }
  }
  for (int i = tid; i < p.npos4; i += kThreads) {
    int off = 0;
    if (i < npos) {
      int img = i / RW, rem = i - img * RW;
      int y = rem / g.W, x = rem - y * g.W;
      off = img * g.IS + y * g.RS + x;
    }
    lut[i] = off * 8;  // bytes
  }
  __syncthreads();

  // column pairs of this warp (byte offsets into the A tile)
  const int c0 = blockIdx.y * p.CC;
  int pair_off[kNP];
  int npw = (p.npairs - wn + p.WN - 1) / p.WN;  // pair j of this warp is j * WN + wn
  npw = npw < 0 ? 0 : (npw > kNP ? kNP : npw);
#pragma unroll
  for (int j = 0; j < kNP; j++) {
    int pr = j * p.WN + wn, off = 0;
    if (j < npw) {
      if (TAPW) {  // pr = (c * KH + kh) * NCF + kwg ; column = kw
        int kwg = pr % p.NCF, r = pr / p.NCF;
        int kh = r % g.KH, c = r / g.KH;
        off = c * g.CS + kh * g.RS + kwg * 8 + fr;
      } else {     // pr = tap * NCF + cf ; column = channel
        int cf = pr % p.NCF, tap = pr / p.NCF;
        int kh = tap / g.KW, kw = tap - kh * g.KW;
        off = (cf * 8 + fr) * g.CS + kh * g.RS + kw;
      }
    }
    pair_off[j] = off * 8;
  }
  double acc[MO][kNP][2];
#pragma unroll
  for (int m = 0; m < MO; m++)
#pragma unroll
    for (int j = 0; j < kNP; j++) acc[m][j][0] = acc[m][j][1] = 0.0;

  const int o_tile = blockIdx.z * (WM * MO * 8);
  const int T = (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int co = p.Co - o_tile < WM * MO * 8 ? p.Co - o_tile : WM * MO * 8;

  const int cw = warp / g.NI, iw = warp - cw * g.NI;
  auto issue = [&](int t) {
    int it = blockIdx.x + t * gridDim.x;
    int band = it / g.nblocks, nblk = it - band * g.nblocks;
    int n0 = nblk * g.NI, y0 = band * g.R;
    double *st = smem + (t & 1) * p.stage_doubles;
    double *db = st + p.a_doubles;
    if (!TAPW && p.tma) {
      // one thread: the A box (zero fill = the padding) and one contiguous row block of dB per output channel
      if (tid == 0) {
        mbar_expect_tx(mbar + (t & 1), p.CC * g.CS * 8 + co * RW * 8);  // the box [CC][TR][RS] + co row blocks
        tma_load_4d(st, &p.tmap, mbar + (t & 1), 0, y0, c0, n0);
        const double *s0 = p.dB + n0 * p.b_sn + o_tile * p.b_so + (int64_t)y0 * g.W;
        for (int o = 0; o < co; o++) bulk_load(db + o * p.PS, s0 + o * p.b_so, RW * 8, mbar + (t & 1));
      }
      return;
    }
    load_in_tile(st, p.A, g, p.CC, c0, p.Ci, n0, p.N, y0, p.a_sn, p.a_sc, tid, cw, iw);
    int rows_ok = g.H - y0 < g.R ? g.H - y0 : g.R;
    const int planes = co * g.NI;
    if (p.pooled) {
      // pooled cotangent of the band: (rows_ok / 2) * (W / 2) contiguous doubles per (o, image)
      const int Ho = g.H >> 1, Wo = g.W >> 1;
      const int nvalp = (rows_ok >> 1) * Wo;
      for (int pl = warp; pl < planes; pl += kThreads / 32) {
        int o = pl / g.NI, img = pl - o * g.NI;
        double *d0 = db + pl * p.PPS;
        bool okn = n0 + img < p.N;
        const double *s0 = p.dy + (((int64_t)(n0 + img) * p.Co + o_tile + o) * Ho + (y0 >> 1)) * Wo;
        if (p.dyvec) {
          for (int ch = lane * 2; ch < p.PW; ch += 64) {
            if (okn && ch < nvalp) cp16(d0 + ch, s0 + ch);
            else *reinterpret_cast<double2 *>(d0 + ch) = make_double2(0.0, 0.0);
          }
        } else {
          for (int ch = lane; ch < p.PW; ch += 32) {
            if (okn && ch < nvalp) cp8(d0 + ch, s0 + ch);
            else d0[ch] = 0.0;
          }
        }
      }
      cp_commit();
      return;
    }
    // dB tile: [o][img * RW + y * W + x], rows beyond the image zero; a warp
    // owns whole (o, img) planes, lanes run along the contiguous positions
    int nval = rows_ok * g.W;
    for (int pl = warp; pl < planes; pl += kThreads / 32) {
      int o = pl / g.NI, img = pl - o * g.NI;
      double *d0 = db + o * p.PS + img * RW;
      bool okn = n0 + img < p.N;
      const double *s0 = p.dB + (n0 + img) * p.b_sn + (o_tile + o) * p.b_so + (int64_t)y0 * g.W;
      if (p.dbvec) {
        for (int ch = lane * 2; ch < RW; ch += 64) {
          if (okn && ch < nval) cp16(d0 + ch, s0 + ch);
          else *reinterpret_cast<double2 *>(d0 + ch) = make_double2(0.0, 0.0);
        }
      } else {
        for (int ch = lane; ch < RW; ch += 32) {
          if (okn && ch < nval) cp8(d0 + ch, s0 + ch);
          else d0[ch] = 0.0;
        }
      }
    }
    cp_commit();
  };

  if (T > 0) issue(0);
  for (int t = 0; t < T; t++) {
    if (!TAPW && p.tma) {
      if (t + 1 < T) issue(t + 1);
      mbar_wait(mbar + (t & 1), (t >> 1) & 1);
    } else {
      if (t + 1 < T) {
        issue(t + 1);
        cp_wait<1>();
      } else {
        cp_wait<0>();
      }
      __syncthreads();
    }
    if (p.pooled) {
      // expand the staged pooled cotangent of stage t into the dB tile:
      // window position `code & 0x7f` receives dy where the relu was active
      // (code & 0x80), the other three positions of the window zero
      int it = blockIdx.x + t * gridDim.x;
      int band = it / g.nblocks, nblk = it - band * g.nblocks;
      int n0 = nblk * g.NI, y0 = band * g.R;
      int rows_ok = g.H - y0 < g.R ? g.H - y0 : g.R;
      const int Ho = g.H >> 1, Wo = g.W >> 1, nvalp = (rows_ok >> 1) * Wo;
      const double *dyp = smem + (t & 1) * p.stage_doubles + p.a_doubles;
      const int planes = co * g.NI;
      for (int pl = warp; pl < planes; pl += kThreads / 32) {
        int o = pl / g.NI, img = pl - o * g.NI;
        bool okn = n0 + img < p.N;
        const uint8_t *c0p = p.code + (((int64_t)(n0 + img) * p.Co + o_tile + o) * Ho + (y0 >> 1)) * Wo;
        double *x0 = dbx + o * p.PS + img * RW;
        const double *d0 = dyp + pl * p.PPS;
        for (int idx = lane; idx < p.PW; idx += 32) {
          unsigned cd = (okn && idx < nvalp) ? c0p[idx] : 0u;
          double gv = (cd & 0x80u) ? d0[idx] : 0.0;
          int bi = cd & 0x7f;
          double *xp = x0 + lutp[idx];
          *reinterpret_cast<double2 *>(xp) = make_double2(bi == 0 ? gv : 0.0, bi == 1 ? gv : 0.0);
          *reinterpret_cast<double2 *>(xp + g.W) = make_double2(bi == 2 ? gv : 0.0, bi == 3 ? gv : 0.0);
        }
      }
      __syncthreads();
    }
    const char *As = reinterpret_cast<const char *>(smem + (t & 1) * p.stage_doubles);
    const double *Bs = (p.pooled ? dbx : smem + (t & 1) * p.stage_doubles + p.a_doubles) + (wm * MO * 8 + fr) * p.PS + fk;
    const int kbeg = wk * 4, kstep = 4 * p.WK;
    int kb = 0;
    const int ytop = g.H - (int)((blockIdx.x + t * gridDim.x) / g.nblocks) * g.R;  // image rows from the band's first row on
    for (int v = p.skip ? g.KH - 1 : 0; v >= 0; v--) {
      // segment of positions whose row still has an in-image input row under tap rows <= v
      int ke = p.npos4, n = npw;
      if (p.skip) {
        ke = ((ytop - v) * g.W + 3) & ~3;
        ke = ke > p.npos4 ? p.npos4 : (ke < kb ? kb : ke);
        n = ((v + 1) * g.KW * p.NCF - wn + p.WN - 1) / p.WN;  // pairs of this warp with kh <= v
        n = n < 0 ? 0 : (n > npw ? npw : n);
      }
      const int ks = kb + ((kbeg - kb) & (kstep - 1));  // first k0 >= kb of this warp's K slice (kstep is a power of two)
      switch (n) {  // exact pair count: no predicated-off DMMAs (they would still occupy the pipe)
        case 9: dk_stage<9, MO>(acc, pair_off, As, Bs, lut + fk, ks, ke, kstep, p); break;
        case 8: dk_stage<8, MO>(acc, pair_off, As, Bs, lut + fk, ks, ke, kstep, p); break;
        case 7: dk_stage<7, MO>(acc, pair_off, As, Bs, lut + fk, ks, ke, kstep, p); break;
        case 6: dk_stage<6, MO>(acc, pair_off, As, Bs, lut + fk, ks, ke, kstep, p); break;
        case 5: dk_stage<5, MO>(acc, pair_off, As, Bs, lut + fk, ks, ke, kstep, p); break;
        case 4: dk_stage<4, MO>(acc, pair_off, As, Bs, lut + fk, ks, ke, kstep, p); break;
        case 3: dk_stage<3, MO>(acc, pair_off, As, Bs, lut + fk, ks, ke, kstep, p); break;
        case 2: dk_stage<2, MO>(acc, pair_off, As, Bs, lut + fk, ks, ke, kstep, p); break;
        case 1: dk_stage<1, MO>(acc, pair_off, As, Bs, lut + fk, ks, ke, kstep, p); break;
        default: break;
      }
      kb = ke;
    }
    __syncthreads();
  }
  // K-split warps of this CTA are summed in fixed order through shared memory
  if (p.WK > 1) {
    double *red = smem;  // [warp][m][j][lane][2]
#pragma unroll
    for (int m = 0; m < MO; m++)
#pragma unroll
      for (int j = 0; j < kNP; j++)
        if (j < npw) {
          double *r = red + (((warp * MO + m) * kNP + j) * 32 + lane) * 2;
          r[0] = acc[m][j][0];
          r[1] = acc[m][j][1];
        }
    __syncthreads();
    if (wk == 0) {
      const int wstride = WM * p.WN;
#pragma unroll
      for (int m = 0; m < MO; m++)
#pragma unroll
        for (int j = 0; j < kNP; j++)
          if (j < npw) {
            double s0 = acc[m][j][0], s1 = acc[m][j][1];
            for (int k = 1; k < p.WK; k++) {
              const double *r = red + ((((warp + k * wstride) * MO + m) * kNP + j) * 32 + lane) * 2;
              s0 += r[0];
              s1 += r[1];
            }
            acc[m][j][0] = s0;
            acc[m][j][1] = s1;
          }
    }
  }
  if (wk != 0) return;
  // partial result of this CTA: C fragment row = o, cols = 2 columns
  const int KK = g.KH * g.KW;
  double *part = p.part + (int64_t)blockIdx.x * ((int64_t)p.Co * p.Ci * KK);
#pragma unroll
  for (int j = 0; j < kNP; j++) {
    if (j < npw) {
      int pr = j * p.WN + wn;
#pragma unroll
      for (int m = 0; m < MO; m++) {
        int o = o_tile + wm * MO * 8 + m * 8 + fr;
#pragma unroll
        for (int e = 0; e < 2; e++) {
          int col = fk * 2 + e, c, kh, kw;
          if (TAPW) {
            int kwg = pr % p.NCF, r = pr / p.NCF;
            kh = r % g.KH;
            c = c0 + r / g.KH;
            kw = kwg * 8 + col;
          } else {
            int cf = pr % p.NCF, tap = pr / p.NCF;
            kh = tap / g.KW;
            kw = tap - kh * g.KW;
            c = c0 + cf * 8 + col;
          }
          if (o < p.Co && c < p.Ci && kw < g.KW) part[((int64_t)o * p.Ci + c) * KK + kh * g.KW + kw] = acc[m][j][e];
        }
      }
    }
  }
}

// dK[i] = sum_s part[s][i], s ascending within each of 32 interleaved groups,
// groups combined in order (fixed association -> deterministic)
__global__ void __launch_bounds__(1024) dk_reduce_kernel(const double *__restrict__ part, double *__restrict__ out,
                                                         int n, int splits) {
  __shared__ double sm[32][33];
  int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  int i = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (i < n)
    for (int k = grp; k < splits; k += 32) s += part[(int64_t)k * n + i];
  sm[grp][lane] = s;
  __syncthreads();
  if (grp == 0 && i < n) {
    double t = sm[0][lane];
    for (int k = 1; k < 32; k++) t += sm[k][lane];
    out[i] = t;
  }
}

typedef CUresult (*hb_tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled through the runtime (no libcuda link); NULL when unavailable or HB_CONV_TMA=0
static hb_tmap_encode_fn tmap_encoder() {
  static hb_tmap_encode_fn fn = [] {
    const char *e = getenv("HB_CONV_TMA");
    if (e && e[0] == '0') return (hb_tmap_encode_fn) nullptr;
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      f = nullptr;
    }
    return (hb_tmap_encode_fn)f;
  }();
  return fn;
}

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
// smallest s >= n with s = 4 (mod 8)
static inline int stride4mod8(int n) {
  int s = round_up(n, 8) + 4;
  if (s - 8 >= n) s -= 8;
  return s;
}
// position stride of the pooled epilogue's output tile: >= n, = 2 (mod 4), so the
// fragment stores of a half-warp (4 channel pairs x 4 positions) hit distinct banks
static inline int pool_stride(int n) {
  int s = round_up(n, 4) + 2;
  if (s - 4 >= n) s -= 4;
  return s;
}
static inline void set_pieces(TileGeo &g) {
  g.cpr = g.vec ? g.W / 2 : g.W;
  g.cshift = 0;
  while ((1 << g.cshift) < g.cpr) g.cshift++;
  g.dc8 = (kThreads / 32) / g.NI;
  g.di8 = (kThreads / 32) % g.NI;
}

template <typename K>
static int set_smem(K kernel, int bytes) {
  HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return HB_OK;
}

}  // namespace

struct hb_pool_fuse {
  const hb_array *bias;  // [Q] or NULL
  hb_array *y, *code;    // [N,Q,H/2,W/2] f64 / i8
};

// ------------------------------------------------------------------ hosts
// Returns HB_OK and sets *done = 1 when the fast path ran.
// flip = 0: forward (in = A[N,P,H,W], K[Q,P,KH,KW]); flip = 1: dInp
// (in = dB[N,P,H,W], K[P,Q,KH,KW]).  `in` contiguous f64, out contiguous.
// pf != NULL: fused bias + relu + 2x2/2 max-pool epilogue (forward only): `out`
// is ignored, pf->y [N,Q,H/2,W/2] f64 and pf->code (i8) are written instead.
int hb_conv_sp_dmma(const hb_array *K, const hb_array *in, hb_array *out, int flip, const hb_pool_fuse *pf,
                    int *done) {
  *done = 0;
  if (in->dt != HB_F64 || !hb_contig(in)) return HB_OK;
  if (pf ? (flip || !hb_contig(pf->y) || !hb_contig(pf->code) || (pf->bias && !hb_contig(pf->bias)))
         : !hb_contig(out))
    return HB_OK;
  int N = (int)in->shape[0], P = (int)in->shape[1], H = (int)in->shape[2], W = (int)in->shape[3];
  int Q = flip ? (int)K->shape[1] : (int)K->shape[0];
  if (pf && ((H & 1) || (W & 1) || H < 2 || W < 2)) return HB_OK;
  int KH = (int)K->shape[2], KW = (int)K->shape[3];
  if ((int64_t)N * P * H * W >= (1ll << 31) || (int64_t)N * Q * H * W >= (1ll << 31)) return HB_OK;
  if (W > 224 || KH > 16 || KW > 16 || H >= 32768) return HB_OK;
  SpP p;
  memset(&p, 0, sizeof p);
  // k mode
  double eff_chan = (double)P / round_up(P, 4), eff_tap = (double)KW / round_up(KW, 4);
  bool tapk = eff_tap > eff_chan;
  p.KWP = round_up(KW, 4);
  int KWe = tapk ? p.KWP : KW;
  // output-channel tile and warp grid
  int WN, NQ;
  if (Q <= 8) { WN = 1; NQ = 1; }
  else if (Q <= 16) { WN = 1; NQ = 2; }
  else if (Q <= 32) { WN = 2; NQ = 2; }
  else { WN = 2; NQ = 4; }
  if (pf && NQ == 4) return HB_OK;  // 64-channel tile: pooled epilogue not instantiated, unfused path
  const int WM = 8 / WN, MF = 7;
  const bool one_cta = (NQ == 4);  // 56 accumulator doubles per thread: 1 CTA / SM
  p.QT = WN * NQ * 8;
  int nqt = (Q + p.QT - 1) / p.QT;
  p.TS = tapk ? p.QT + 4 : p.QT;
  p.WCS = stride4mod8(KH * KWe * p.TS);
  int slots = WM * MF * 8;
  TileGeo &g = p.g;
  g.H = H; g.W = W; g.KH = KH; g.KW = KW;
  g.oy = flip ? -(KH - 1) : 0;
  g.ox = flip ? -(KW - 1) : 0;
  int halo_w = tapk ? p.KWP - 1 : KW - 1;
  if (flip && tapk) halo_w = (KW - 1) + (p.KWP - 1);  // low-side halo plus the padded taps' overrun
  g.RS = round_up(W + halo_w, 2);
  int Pq = tapk ? P : round_up(P, 4);
  // tile candidates: NI whole images, or bands of R rows of one image; keep the
  // best-balanced one (fragments per warp) that fits two stages in the budget
  int budgets[2] = {one_cta ? kSmemMax : (kSmemMax / 2 - 1024), kSmemMax};
  int bestNI = 0, bestR = 0, bestCC = 0, ctas = 1;
  for (int bi = 0; bi < 2 && !bestCC; bi++) {
    double best_eff = 0;
    for (int cand = 0; cand < 64; cand++) {
      int NI, R;
      if (H * W <= slots) { NI = cand + 1; R = H; if (NI > N || NI * H * W > slots) break; }
      else {
        NI = 1; R = slots / W - cand; if (R < 1) break;
        int nb = (H + R - 1) / R; R = (H + nb - 1) / nb;
        if (pf && (R & 1)) { R++; if (R * W > slots) continue; }  // bands hold whole pooling windows
      }
      int TR = R + KH - 1, IS = TR * g.RS;
      int CS = stride4mod8(NI * IS + (tapk ? 4 : 0));
      int per_c = CS + p.WCS;
      // fused pooling: room for an output tile of at least min(QT, 8) channels + the window table
      int extra = 0;
      if (pf) {
        int np = NI * R * W;
        extra = (p.QT < 8 ? p.QT : 8) * pool_stride(np) * 8 + (np / 4) * 16 + 64;
      }
      if (budgets[bi] - extra < 0) continue;
      // weights per stage next to their channels, or all channels' weights staged once per CTA
      int maxc = ((budgets[bi] - extra) / 8 / 2) / per_c;
      int maxc_fix = ((budgets[bi] - extra) / 8 - round_up(Pq, 4) * p.WCS - 8) / 2 / CS;
      if (maxc_fix > maxc) maxc = maxc_fix;
      if (!tapk) maxc = maxc / 4 * 4;
      if (maxc < (tapk ? 1 : 4)) continue;
      int CC = maxc < Pq ? maxc : Pq;
      int F = (NI * R * W + 7) / 8, per_warp = (F + WM - 1) / WM;
      int nb = (H + R - 1) / R;
      // balance over warps and bands, times the fragment-reuse factor: a warp
      // with m fragments issues m + NQ shared loads per m * NQ DMMAs and pays the
      // per-stage fixed costs over m fragments
      double eff = (double)F / (WM * per_warp) * ((double)H / (nb * R)) * per_warp / (per_warp + 1.0);
      // prefer fewer channel chunks when the balance is (nearly) the same
      int nch = (Pq + CC - 1) / CC;
      eff -= 0.01 * (nch - 1);
      if (eff > best_eff + 1e-9) { best_eff = eff; bestNI = NI; bestR = R; bestCC = CC; }
    }
    if (bestCC) ctas = (bi == 0 && !one_cta) ? 2 : 1;
  }
  if (!bestCC) return HB_OK;
  int CC = bestCC;
  g.R = bestR; g.NI = bestNI; g.TR = g.R + KH - 1; g.IS = g.TR * g.RS;
  g.CS = stride4mod8(g.NI * g.IS + (tapk ? 4 : 0));
  // balance chunks
  p.NCH = (Pq + CC - 1) / CC;
  CC = (Pq + p.NCH - 1) / p.NCH;
  if (!tapk) CC = round_up(CC, 4);
  p.CC = CC;
  p.Ppad = p.NCH * CC;
  g.bands = (H + g.R - 1) / g.R;
  g.nblocks = (N + g.NI - 1) / g.NI;
  p.items = g.nblocks * g.bands;
  p.N = N; p.P = P; p.Q = Q;
  p.sn = in->strides[0]; p.sp = in->strides[1];
  p.in = (const double *)hb_data(in);
  const int budget_used = ctas == 2 ? kSmemMax / 2 - 1024 : kSmemMax;
  int pool_extra = 0;
  if (pf) pool_extra = (p.QT < 8 ? p.QT : 8) * pool_stride(g.NI * g.R * W) * 8 + (g.NI * g.R * W / 4) * 16 + 64;
  // stage layout; all chunks' weights once per CTA when that fits the budget the tile was chosen for
  auto layout = [&](int in_doubles, int align) {
    p.in_doubles = in_doubles;
    const bool wfixed =
        (2 * round_up(in_doubles, align) + round_up(p.Ppad * p.WCS, align)) * 8 + pool_extra + 64 <= budget_used;
    p.stage_doubles = round_up(in_doubles + (wfixed ? 0 : CC * p.WCS), align);
    p.wfix = wfixed ? 2 * p.stage_doubles : 0;
    p.tail = 2 * p.stage_doubles + (wfixed ? round_up(p.Ppad * p.WCS, align) : 0);
    return p.tail * 8 + pool_extra + 64;
  };
  // TMA staging (channel-k mode): the tile is the dense box [NI][CC][TR][RS] the copy engine writes, so
  // the channel stride is TR * RS; the box is widened by up to 6 columns until that is = 4 (mod 8)
  // doubles (conflict-free fragment loads); shapes where no width does it keep the cp.async loader
  p.tma = 0;
  if (!tapk && tmap_encoder() && W % 2 == 0 && ((uintptr_t)p.in & 15) == 0 && p.sn == (int64_t)P * H * W &&
      p.sp == (int64_t)H * W) {
    // an odd first column (dInp with an even kernel width) would start the box off a 16-byte boundary,
    // which the copy engine rejects: start one column further left and shift the fragment reads instead
    const int xs = g.ox & 1, rs_min = round_up(W + halo_w + xs, 2);
    int rs_t = 0;
    for (int add = 0; add <= 6 && !rs_t; add += 2)
      if ((g.TR * (rs_min + add)) % 8 == 4) rs_t = rs_min + add;
    if (rs_t && rs_t <= 256 && g.TR <= 256 && CC <= 256 && g.NI <= 256) {
      const int rs_old = g.RS, cs_old = g.CS, is_old = g.IS;
      g.RS = rs_t; g.CS = g.TR * rs_t; g.IS = CC * g.CS;
      if (layout(g.NI * g.IS, 16) <= budget_used) {
        CUtensorMap tm;
        cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)W * 8, (cuuint64_t)H * W * 8, (cuuint64_t)P * H * W * 8};
        cuuint32_t box[4] = {(cuuint32_t)rs_t, (cuuint32_t)g.TR, (cuuint32_t)CC, (cuuint32_t)g.NI};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = tmap_encoder()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void *)p.in, dims, strides, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS) {
          p.tmap = tm;
          p.tma = 1;
          p.xs = xs;
        }
      }
      if (!p.tma) { g.RS = rs_old; g.CS = cs_old; g.IS = is_old; }
    }
  }
  if (!p.tma) layout(CC * g.CS, 2);
  g.vec = (W % 2 == 0) && (g.ox % 2 == 0) && (((uintptr_t)p.in & 15) == 0) && (p.sn % 2 == 0) && (p.sp % 2 == 0);
  set_pieces(g);
  {
    // fragment order and per-tap-row fragment counts (see SpP::perm).  Under tap row kh
    // a fragment with rows [ymin, ymax] of band b (first row y0) has an in-image input row iff
    // -oy - (y0 + ymax) <= kh <= H - 1 - oy - (y0 + ymin): forward (oy = 0) clips high kh for
    // large ymin, dInp (oy = 1 - KH) clips low kh for small ymax -> ordering the fragments by
    // ymin ascending (forward) or ymax descending (dInp) makes the active ones a prefix, as
    // long as only that one bound is applied
    const int npos = g.NI * g.R * W, F = (npos + 7) / 8, RW = g.R * W;
    int ymin[64], ymax[64], order[64];
    const bool skip = !tapk && g.bands <= 4 && g_hb.conv_skip_padding;
    for (int f = 0; f < F; f++) {
      int p0 = f * 8, p1 = p0 + 7 < npos ? p0 + 7 : npos - 1;
      int i0 = p0 / RW, y0r = (p0 % RW) / W, i1 = p1 / RW, y1r = (p1 % RW) / W;
      ymin[f] = i0 == i1 ? y0r : 0;
      ymax[f] = i0 == i1 ? y1r : g.R - 1;
      order[f] = f;
    }
    if (skip) {
      if (flip) std::stable_sort(order, order + F, [&](int a, int b) { return ymax[a] > ymax[b]; });
      else std::stable_sort(order, order + F, [&](int a, int b) { return ymin[a] < ymin[b]; });
    }
    for (int s = 0; s < 64; s++) p.perm[s] = (unsigned char)(s < F ? order[s] : 0);
    for (int b = 0; b < 4; b++)
      for (int wm = 0; wm < WM; wm++)
        for (int kh = 0; kh < KH; kh++) {
          int n = 0, y0 = b * g.R;
          for (int i = 0; i < MF && i * WM + wm < F; i++) {
            int f = order[i * WM + wm];
            // only the bound the ordering is monotone in (a superset of the needed work: rows of a
            // ragged last band beyond the image may be computed, they are never stored)
            bool on = flip ? -g.oy - (y0 + ymax[f]) <= kh : kh <= H - 1 - g.oy - (y0 + ymin[f]);
            if (on || !skip) n++;
          }
          p.nk[b][wm][kh] = (unsigned char)n;
        }
  }
  int smem_bytes = p.tail * 8;
  if (pf) {
    const int np = g.NI * g.R * W;
    p.pool = 1;
    p.OPS = pool_stride(np);
    p.bias = pf->bias ? (const double *)hb_data(pf->bias) : nullptr;
    p.pool_out = (double *)hb_data(pf->y);
    p.code_out = (uint8_t *)hb_data(pf->code);
    int budget = ctas == 2 ? kSmemMax / 2 - 1024 : kSmemMax;
    int avail = budget - smem_bytes - (np / 4) * 16 - 64;
    p.qg = p.QT;
    while (p.qg > 1 && p.qg * p.OPS * 8 > avail) p.qg >>= 1;
    if (p.qg < 2 || p.qg * p.OPS * 8 > avail) return HB_OK;
    smem_bytes += p.qg * p.OPS * 8 + (np / 4) * 16 + 64;
  } else {
    p.out = (double *)hb_data(out);
  }
  smem_bytes = round_up(smem_bytes, 8);
  p.mbar = smem_bytes / 8;
  smem_bytes += 16;
  if (smem_bytes > kSmemMax) return HB_OK;
  // packed weights in scratch
  int64_t wn = (int64_t)nqt * p.Ppad * p.WCS;
  void *scr;
  HB_TRY(hb_scratch((size_t)wn * 8, &scr));
  p.wp = (const double *)scr;
  {
    int blocks = hb_blocks_for(wn, 256, g_hb.sm_count * 8);
    pack_sp_weights<<<blocks, 256, 0, g_hb.stream>>>((const double *)hb_data(K), (double *)scr, KH, KW, KWe, P, Q,
                                                     p.Ppad, p.QT, p.TS, p.WCS, nqt, flip, K->strides[0],
                                                     K->strides[1], K->strides[2], K->strides[3]);
    HB_LAUNCHED();
  }
  int maxg = g_hb.sm_count * ctas;
  dim3 grid((unsigned)(p.items < maxg ? p.items : maxg), (unsigned)nqt);
#define SP_LAUNCH2(WM_, WN_, NQ_, MINB_, TAPK_, POOL_)                                                            \
  do {                                                                                                           \
    HB_TRY(set_smem(conv_sp_kernel<WM_, WN_, 7, NQ_, TAPK_, MINB_, POOL_>, smem_bytes));                         \
    conv_sp_kernel<WM_, WN_, 7, NQ_, TAPK_, MINB_, POOL_><<<grid, kThreads, smem_bytes, g_hb.stream>>>(p);       \
  } while (0)
#define SP_LAUNCH(WM_, WN_, NQ_, MINB_)                                                \
  do {                                                                                 \
    if (pf) {                                                                          \
      if (tapk) SP_LAUNCH2(WM_, WN_, NQ_, MINB_, true, true);                          \
      else SP_LAUNCH2(WM_, WN_, NQ_, MINB_, false, true);                              \
    } else {                                                                           \
      if (tapk) SP_LAUNCH2(WM_, WN_, NQ_, MINB_, true, false);                         \
      else SP_LAUNCH2(WM_, WN_, NQ_, MINB_, false, false);                             \
    }                                                                                  \
  } while (0)
  if (WN == 1 && NQ == 1) SP_LAUNCH(8, 1, 1, 2);
  else if (WN == 1 && NQ == 2) SP_LAUNCH(8, 1, 2, 2);
  else if (WN == 2 && NQ == 2) SP_LAUNCH(4, 2, 2, 2);
  else {
    if (tapk) SP_LAUNCH2(4, 2, 4, 1, true, false);
    else SP_LAUNCH2(4, 2, 4, 1, false, false);
  }
#undef SP_LAUNCH
#undef SP_LAUNCH2
  HB_LAUNCHED();
  (void)MF;
  *done = 1;
  return HB_OK;
}

// dK[Co,Ci,KH,KW] from A[N,Ci,H,W], dB[N,Co,H,W] (contiguous f64).  With
// `code` != NULL, dB is instead the POOLED cotangent dy [N,Co,H/2,W/2] of the
// fused bias + relu + 2x2 max-pool and `code` its adjoint code bytes.
// measurement switches of the dKrn tile choice (read once): HB_DK_WM=4 -> 64-channel output tiles,
// HB_DK_NCF=n -> n column fragments of 8 input channels per CTA, HB_DK_R=r -> band height r,
// HB_DK_MAXNB=n -> accept two-CTA bands of fewer than 4 rows up to n bands, HB_CONV_DEBUG -> print the choice
namespace {
struct DkKnobs {
  int wm = 0, ncf = 0, r = 0, maxnb = 4, debug = 0;
  DkKnobs() {
    auto num = [](const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; };
    wm = num("HB_DK_WM", 0); ncf = num("HB_DK_NCF", 0); r = num("HB_DK_R", 0); maxnb = num("HB_DK_MAXNB", 4);
    debug = getenv("HB_CONV_DEBUG") != nullptr;
  }
};
const DkKnobs &dk_knobs() { static const DkKnobs k; return k; }
}  // namespace

int hb_conv_dk_dmma(const hb_array *A, const hb_array *dB, const hb_array *code, hb_array *out, int KH, int KW,
                    int *done) {
  *done = 0;
  const DkKnobs &knobs = dk_knobs();
  if (A->dt != HB_F64 || !hb_contig(A) || !hb_contig(dB) || !hb_contig(out)) return HB_OK;
  int N = (int)A->shape[0], Ci = (int)A->shape[1], H = (int)A->shape[2], W = (int)A->shape[3];
  int Co = (int)dB->shape[1];
  const bool pooled = code != nullptr;
  if (pooled && (!hb_contig(code) || (H & 1) || (W & 1) || H < 2 || W < 2)) return HB_OK;
  if ((int64_t)N * Ci * H * W >= (1ll << 31) || (int64_t)N * Co * H * W >= (1ll << 31)) return HB_OK;
  if (W > 256 || KH > 16 || KW > 16) return HB_OK;
  DkP p;
  memset(&p, 0, sizeof p);
  double eff_chan = (double)Ci / round_up(Ci, 8), eff_tap = (double)KW / round_up(KW, 8);
  bool tapw = eff_tap > eff_chan;
  int WM, MO;
  if (Co <= 8) { WM = 1; MO = 1; }
  else if (Co <= 16) { WM = 1; MO = 2; }
  else if (Co <= 32) { WM = 2; MO = 2; }
  else { WM = 2; MO = 2; }  // 32-channel tiles: two CTAs per SM (see the band choice below)
  if (Co > 32 && knobs.wm == 4) WM = 4;
  int otile = WM * MO * 8, not_ = (Co + otile - 1) / otile;
  int wrest = 8 / WM;  // warps left for pairs x K
  int KK = KH * KW;
  int maxpairs = wrest * kNP;
  // channels per CTA
  int CC, NCF, npairs;
  if (tapw) {
    NCF = (KW + 7) / 8;
    int per_c = KH * NCF;
    if (per_c > maxpairs) return HB_OK;
    CC = maxpairs / per_c;
    if (CC > Ci) CC = Ci;
    npairs = CC * per_c;
  } else {
    if (KK > maxpairs) return HB_OK;
    int ncf = maxpairs / KK;  // column fragments of 8 channels
    int need = (Ci + 7) / 8;
    if (ncf > need) ncf = need;
    // 32-channel output tiles take 16 input channels per CTA (K split over the two spare warps) so that two
    // CTAs fit one SM: ConvVjpBench dKrn 0.613 ms (64 x 16 tile, one CTA) -> 0.513 (32 x 32, one CTA, TMA)
    // -> 0.501 ms (32 x 16, two CTAs, TMA)
    if (WM == 2 && ncf > 2) ncf = 2;
    if (knobs.ncf > 0 && knobs.ncf <= maxpairs / KK) ncf = knobs.ncf < need ? knobs.ncf : need;
    // balance the chunks
    int chunks = (need + ncf - 1) / ncf;
    ncf = (need + chunks - 1) / chunks;
    NCF = ncf;
    CC = ncf * 8;
    npairs = KK * ncf;
  }
  int ncc = (Ci + CC - 1) / CC;
  int WN = 1;
  while (WN < wrest && WN * kNP < npairs) WN *= 2;
  if (WN * kNP < npairs) return HB_OK;
  p.WN = WN;
  p.WK = wrest / WN;
  p.NPW = (npairs + WN - 1) / WN;
  p.CC = CC; p.NCF = NCF; p.npairs = npairs;
  TileGeo &g = p.g;
  g.H = H; g.W = W; g.KH = KH; g.KW = KW; g.oy = 0; g.ox = 0;
  int halo_w = tapw ? NCF * 8 - 1 : KW - 1;
  g.RS = round_up(W + halo_w, 2);
  // positions per stage: as many rows as two stages allow, two CTAs per SM
  // when that still leaves at least a quarter of the image per stage
  const int red_bytes = p.WK > 1 ? 8 * MO * kNP * 64 * 8 : 0;  // end-of-kernel K-split reduction area
  int NI = 1, ctas = 1;
  auto smem_need = [&]() {
    int b = 2 * p.stage_doubles * 8 + p.npos4 * 4 + 16;
    if (pooled) b += otile * p.PS * 8 + p.PW * 4 + 16;
    return b;
  };
  auto fits = [&](int R, int budget) {
    if (pooled && (R & 1)) return false;  // bands hold whole pooling windows
    g.R = R; g.NI = NI; g.TR = R + KH - 1; g.IS = g.TR * g.RS;
    g.CS = stride4mod8(NI * g.IS + (tapw ? 8 : 0));
    p.npos4 = round_up(NI * R * W, 4);
    p.PS = stride4mod8(p.npos4);
    p.a_doubles = round_up(CC * g.CS, 2);
    if (pooled) {
      p.PW = (R / 2) * (W / 2);
      p.PPS = round_up(p.PW, 2);
      p.stage_doubles = p.a_doubles + round_up(otile * NI * p.PPS, 2);
    } else {
      p.stage_doubles = p.a_doubles + otile * p.PS;
    }
    return smem_need() <= budget && red_bytes <= budget;
  };
  int R = 0;
  const int maxnb2 = knobs.maxnb, forceR = knobs.r;
  if (forceR > 0 && forceR <= H) {
    if (fits(forceR, kSmemMax / 2 - 1024)) { R = forceR; ctas = 2; }
    else if (fits(forceR, kSmemMax)) { R = forceR; ctas = 1; }
  }
  // Output tiles of at most 16 channels over a whole small image (the second layer of MnistCnnShaped2: 16 -> 16
  // channels, 14 x 14): ONE band per image and one CTA per SM -- no halo rows read twice and half as many partial
  // kernels to write and reduce.  Measured per training step: 0.303 -> 0.298 ms at 512 samples, 0.487 -> 0.484 at
  // 1024, 0.805 -> 0.804 at 2048, equal at 4096; the 64-channel layer (WM = 2) LOSES with it (7.24 -> 7.72 ms).
  if (!R && WM == 1 && ncc == 1 && not_ == 1 && fits(H, kSmemMax)) { R = H; ctas = 1; }
  // Two CTAs per SM first: with one CTA (two warps per scheduler) ncu shows the DMMA pipe idle a third of
  // the time on fixed-latency waits (ConvVjpBench shape: 65 % -> 77 % of the FP64 peak with two).  Fewest
  // bands first, but a band height that divides H (whole bands: TMA staging applies) wins over a ragged one
  // as long as bands keep at least 4 rows.
  int r_any = 0;
  for (int nb = 1; nb <= H && !R; nb++) {
    int r = (H + nb - 1) / nb;
    if (nb > maxnb2 && r < 4) break;
    if (!fits(r, kSmemMax / 2 - 1024)) continue;
    if (H % r == 0 || pooled || tapw) { R = r; ctas = 2; }
    else if (!r_any) r_any = r;
  }
  if (!R && r_any) { R = r_any; ctas = 2; }
  r_any = 0;
  for (int nb = 1; nb <= H && !R; nb++) {  // one CTA per SM: same preference for whole bands
    int r = (H + nb - 1) / nb;
    if (r_any && r < 4) break;
    if (!fits(r, kSmemMax)) continue;
    if (H % r == 0 || pooled || tapw) { R = r; ctas = 1; }
    else if (!r_any) r_any = r;
  }
  if (!R && r_any) { R = r_any; ctas = 1; }
  if (!R) return HB_OK;
  fits(R, kSmemMax);
  g.bands = (H + g.R - 1) / g.R;
  g.nblocks = (N + NI - 1) / NI;
  p.items = g.nblocks * g.bands;
  p.N = N; p.Ci = Ci; p.Co = Co;
  p.A = (const double *)hb_data(A);
  p.dB = (const double *)hb_data(dB);
  p.a_sn = A->strides[0]; p.a_sc = A->strides[1];
  p.b_sn = dB->strides[0]; p.b_so = dB->strides[1];
  g.vec = (W % 2 == 0) && (((uintptr_t)p.A & 15) == 0) && (p.a_sn % 2 == 0) && (p.a_sc % 2 == 0);
  set_pieces(g);
  p.dbvec = ((g.R * W) % 2 == 0) && (W % 2 == 0) && (((uintptr_t)p.dB & 15) == 0) && (p.b_sn % 2 == 0) &&
            (p.b_so % 2 == 0);
  if (pooled) {
    p.pooled = 1;
    p.dy = p.dB;
    p.code = (const uint8_t *)hb_data(code);
    p.dyvec = (((H / 2) * (W / 2)) % 2 == 0) && (p.PW % 2 == 0) && (((uintptr_t)p.dy & 15) == 0);
  }
  p.skip = !tapw && NI == 1 && g_hb.conv_skip_padding;
  // TMA staging (see hb_conv_sp_dmma): the A box by one tensor copy, the dB rows of each output channel by
  // one bulk copy; whole bands only (a ragged last band needs zeroed rows), unpooled cotangent only
  p.tma = 0;
  if (!tapw && !pooled && NI == 1 && tmap_encoder() && H % g.R == 0 && p.dbvec && ((uintptr_t)p.A & 15) == 0 &&
      p.a_sn == (int64_t)Ci * H * W && p.a_sc == (int64_t)H * W) {
    int rs_t = 0;
    for (int add = 0; add <= 6 && !rs_t; add += 2)
      if ((g.TR * (g.RS + add)) % 8 == 4) rs_t = g.RS + add;
    if (rs_t && rs_t <= 256 && g.TR <= 256 && CC <= 256) {
      g.RS = rs_t; g.IS = g.TR * rs_t; g.CS = g.IS;
      p.a_doubles = round_up(CC * g.CS, 16);
      p.stage_doubles = round_up(p.a_doubles + otile * p.PS, 16);
      if (smem_need() + 32 <= (ctas == 2 ? kSmemMax / 2 - 1024 : kSmemMax)) {
        CUtensorMap tm;
        cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Ci, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)W * 8, (cuuint64_t)H * W * 8, (cuuint64_t)Ci * H * W * 8};
        cuuint32_t box[4] = {(cuuint32_t)rs_t, (cuuint32_t)g.TR, (cuuint32_t)CC, 1u};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = tmap_encoder()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void *)p.A, dims, strides, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS) {
          p.tmap = tm;
          p.tma = 1;
        }
      }
      if (!p.tma) {  // back to the cp.async layout
        g.RS = round_up(W + halo_w, 2);
        fits(R, kSmemMax);
        set_pieces(g);
      }
    }
  }
  int smem_bytes = round_up(smem_need(), 8);
  p.mbar = smem_bytes / 8;
  smem_bytes += 16;
  if (smem_bytes < red_bytes) smem_bytes = red_bytes;
  int splits = g_hb.sm_count * ctas / (ncc * not_);
  if (splits < 1) splits = 1;
  if (splits > p.items) splits = p.items;
  int64_t nout = (int64_t)Co * Ci * KK;
  void *scr;
  HB_TRY(hb_scratch((size_t)splits * nout * 8, &scr));
  p.part = (double *)scr;
  dim3 grid((unsigned)splits, (unsigned)ncc, (unsigned)not_);
  if (knobs.debug)
    fprintf(stderr, "dk: WM=%d MO=%d CC=%d npairs=%d WN=%d WK=%d R=%d ctas=%d tma=%d RS=%d smem=%d grid=(%d,%d,%d) items=%d\n", WM, MO,
            CC, npairs, p.WN, p.WK, g.R, ctas, p.tma, g.RS, smem_bytes, splits, ncc, not_, p.items);
#define DK_LAUNCH(WM_, MO_)                                                                  \
  do {                                                                                       \
    if (tapw) {                                                                              \
      HB_TRY(set_smem(conv_dk_kernel<WM_, MO_, true>, smem_bytes));                          \
      conv_dk_kernel<WM_, MO_, true><<<grid, kThreads, smem_bytes, g_hb.stream>>>(p);        \
    } else {                                                                                 \
      HB_TRY(set_smem(conv_dk_kernel<WM_, MO_, false>, smem_bytes));                         \
      conv_dk_kernel<WM_, MO_, false><<<grid, kThreads, smem_bytes, g_hb.stream>>>(p);       \
    }                                                                                        \
  } while (0)
  if (WM == 1 && MO == 1) DK_LAUNCH(1, 1);
  else if (WM == 1) DK_LAUNCH(1, 2);
  else if (WM == 2) DK_LAUNCH(2, 2);
  else DK_LAUNCH(4, 2);
#undef DK_LAUNCH
  HB_LAUNCHED();
  dk_reduce_kernel<<<(unsigned)((nout + 31) / 32), 1024, 0, g_hb.stream>>>(p.part, (double *)hb_data(out), (int)nout,
                                                                          splits);
  HB_LAUNCHED();
  *done = 1;
  return HB_OK;
}

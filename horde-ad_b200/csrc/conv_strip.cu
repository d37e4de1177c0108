// conv_strip.cu -- warp-autonomous DMMA kernels for the ONE-input-channel
// convolution layer (the first layer of MnistCnnShaped2: 1 -> c_out channels,
// (kh+1) x (kw+1) taps on 28 x 28 glyphs, MnistCnnShaped2.hs:42-62), fused with
// its bias + relu + 2x2 max-pool tail and, backwards, with the pooled cotangent.
//
// Up to 64 output channels, processed as groups of 16 over the same staged strip.
//
// Why a separate kernel: with a single input channel a work item is tiny (a
// pair of output rows = one pooled row needs a 6 x 36 input strip), so the
// CTA-wide stage barriers of conv_dmma.cu dominated (ncu: DMMA pipe 39 % / 47 %
// active).  Here nothing is shared between warps except the read-only weights:
// every warp owns a stream of (image, row pair) items, stages its own strips
// with cp.async (double-buffered, __syncwarp only), runs its DMMAs, pools in its
// own shared tile and writes its pooled row.  No __syncthreads in steady state;
// the other 15 warps of the SM hide each warp's latencies.
//
//   forward : M = 8 positions of the row pair, N = 8 output channels,
//             K = 4 taps of the FLATTENED (kh, kw) tap list (25 -> 28, not
//             5 x 8 = 40 as in the tap-k mode of conv_dmma.cu)
//   dKrn    : M = 8 output channels, N = 8 flattened taps, K = 4 positions;
//             the cotangent tile is expanded from (dy pooled, code) in shared
//             memory; tap slot KH*KW carries a column of ONES, so the bias
//             cotangent falls out of the same DMMAs; per-CTA partials are
//             reduced in fixed order (deterministic).
#include "common.cuh"

namespace {

constexpr int WPB = 16;            // warps per block
constexpr int NTHR = WPB * 32;
constexpr int MF = 7;              // position fragments of a row pair (2 * W <= 56)
// fixed shared-memory strides (W <= 28, KW <= 8), so that fragment addresses fold into immediates
constexpr int kRS = 36;            // strip row stride  (>= W + KW - 1, even)
constexpr int kOPS = 58;           // forward pooled-tile position stride (>= 56, = 2 mod 4)
constexpr int kPS = 60;            // dKrn cotangent-tile position stride (>= 56, = 4 mod 8)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(void *s, const void *g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

struct StripP {
  const double *A;       // [N, 1, H, W]
  const double *wp;      // forward: packed weights [NT][TS] (tap-major, zero padded)
  const double *bias;    // forward: [Q] or NULL
  double *y;             // forward: pooled output [N, Q, H/2, W/2]
  uint8_t *code;         // forward: out; dKrn: in
  const double *dy;      // dKrn: pooled cotangent [N, Q, H/2, W/2]
  double *part;          // dKrn: [gridDim.x][Q][KK + 1]
  int N, Q, H, W, KH, KW, KK;
  int QG;                // groups of (up to) 16 output channels processed one after the other per item
  int NT;                // taps rounded up to 4 (forward) / 8 (dKrn)
  int RS, TR, SD;        // strip row stride, rows, doubles
  int TS, OPS, PS;       // weight tap stride; pooled-tile / cotangent-tile strides
  int items;             // N * H / 2
  int w_doubles;         // shared weights block (forward)
  int warp_doubles;      // per-warp shared region
};

// rows [2 rp, 2 rp + TR) of image n -> strip (rows beyond the image zero; the halo columns stay zero)
__device__ __forceinline__ void load_strip(double *dst, const StripP &p, int n, int rp, int lane) {
  const double *src = p.A + (int64_t)n * p.H * p.W + (int64_t)(2 * rp) * p.W;
  const int cpr = p.W >> 1;
  if (lane < cpr) {
    for (int t = 0; t < p.TR; t++) {
      double *d = dst + t * kRS + 2 * lane;
      if (2 * rp + t < p.H) cp16(d, src + t * p.W + 2 * lane);
      else *reinterpret_cast<double2 *>(d) = make_double2(0.0, 0.0);
    }
  }
}

// ------------------------------------------------------------------ forward
// MULTI = false: one group of output channels (weight stride and trip count are compile-time constants)
template <int NQ, bool MULTI>
__global__ void __launch_bounds__(NTHR, 1) strip_fwd_kernel(const __grid_constant__ StripP p) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int fr = lane >> 2, fk = lane & 3;
  double *Wsm = smem;
  int *tapoff = reinterpret_cast<int *>(smem + p.w_doubles);          // [NT] byte offsets into a strip
  double *mine = smem + p.w_doubles + ((p.NT + 1) >> 1) + warp * p.warp_doubles;
  double *ot = mine + 2 * p.SD;                                        // [QT][kOPS]

  for (int i = tid; i < p.w_doubles; i += NTHR) Wsm[i] = p.wp[i];
  for (int i = tid; i < p.NT; i += NTHR) tapoff[i] = i < p.KK ? ((i / p.KW) * kRS + (i % p.KW)) * 8 : 0;
  for (int i = lane; i < p.warp_doubles; i += 32) mine[i] = 0.0;
  __syncthreads();

  const int Wo = p.W >> 1, Ho = p.H >> 1, W2 = 2 * p.W;
  int aoff[MF];
#pragma unroll
  for (int i = 0; i < MF; i++) {
    int pos = i * 8 + fr;
    int pc = pos < W2 ? pos : 0;
    int yy = pc / p.W, x = pc - yy * p.W;
    aoff[i] = (yy * kRS + x) * 8;
  }
  constexpr int QT = NQ * 8;
  const int TS = MULTI ? p.TS : QT + 4, QG = MULTI ? p.QG : 1;
  const int64_t gw = (int64_t)blockIdx.x * WPB + warp, nw = (int64_t)gridDim.x * WPB;
  const int ng = p.NT >> 2;
  const int HoWo = Ho * Wo;
  const int q_lane = lane / Wo, j_lane = lane - q_lane * Wo;   // pooled output `lane` of an item
  const int dq = 32 / Wo, dj = 32 - dq * Wo;                   // ... and the step to output lane + 32
  double *const ot_w = ot + (fk * 2) * kOPS + fr;              // this lane's slot of fragment 0, channel pair 0
  // item -> (image, row pair), advanced incrementally
  int n = (int)(gw / Ho), rp = (int)(gw - (int64_t)n * Ho);
  const int dn = (int)(nw / Ho), drp = (int)(nw - (int64_t)dn * Ho);
  int buf = 0;
  if (gw < p.items) load_strip(mine, p, n, rp, lane);
  cp_commit();
  for (int64_t it = gw; it < p.items; it += nw) {
    int n2 = n + dn, rp2 = rp + drp;
    if (rp2 >= Ho) { rp2 -= Ho; n2++; }
    if (it + nw < p.items) {
      load_strip(mine + (buf ^ 1) * p.SD, p, n2, rp2, lane);
      cp_commit();
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncwarp();
    const char *Sb = reinterpret_cast<const char *>(mine + buf * p.SD);
    // the staged strip serves every group of 16 output channels in turn
    for (int qg = 0; qg < QG; qg++) {
      double acc[MF][NQ][2];
#pragma unroll
      for (int i = 0; i < MF; i++)
#pragma unroll
        for (int j = 0; j < NQ; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
      const double *wl = Wsm + fk * TS + qg * QT + fr;
      for (int g = 0; g < ng - 1; g++) {       // full tap groups: no padding possible
        const char *sa = Sb + tapoff[g * 4 + fk];
        double b[NQ];
#pragma unroll
        for (int j = 0; j < NQ; j++) b[j] = wl[g * 4 * TS + j * 8];
#pragma unroll
        for (int i = 0; i < MF; i++) {
          double a = *reinterpret_cast<const double *>(sa + aoff[i]);
#pragma unroll
          for (int j = 0; j < NQ; j++) dmma884(acc[i][j][0], acc[i][j][1], a, b[j]);
        }
      }
      {                                         // last group: padded taps contribute exact zeros
        const int g = ng - 1, t = g * 4 + fk;
        const char *sa = Sb + tapoff[t];
        const bool tv = t < p.KK;
        double b[NQ];
#pragma unroll
        for (int j = 0; j < NQ; j++) b[j] = wl[g * 4 * TS + j * 8];
#pragma unroll
        for (int i = 0; i < MF; i++) {
          double a = *reinterpret_cast<const double *>(sa + aoff[i]);
          a = tv ? a : 0.0;
#pragma unroll
          for (int j = 0; j < NQ; j++) dmma884(acc[i][j][0], acc[i][j][1], a, b[j]);
        }
      }
      // bias, then the row pair's pre-activations into the warp's [channel][position] tile
      double bv[NQ][2];
#pragma unroll
      for (int j = 0; j < NQ; j++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          int q = qg * QT + j * 8 + fk * 2 + e;
          bv[j][e] = (p.bias && q < p.Q) ? p.bias[q] : 0.0;
        }
#pragma unroll
      for (int i = 0; i < MF; i++) {
        if (i * 8 + fr < W2) {
#pragma unroll
          for (int j = 0; j < NQ; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) ot_w[(j * 8 + e) * kOPS + i * 8] = acc[i][j][e] + bv[j][e];
        }
      }
      __syncwarp();
      {
        const int q0 = qg * QT;
        const int nout = (p.Q - q0 < QT ? p.Q - q0 : QT) * Wo;
        int q = q_lane, jj = j_lane;
        double *yb = p.y + (((int64_t)n * p.Q + q0) * Ho + rp) * Wo;
        uint8_t *cb = p.code + (((int64_t)n * p.Q + q0) * Ho + rp) * Wo;
        for (int o = lane; o < nout; o += 32) {
          const double *oq = ot + q * kOPS + 2 * jj;
          double2 r0 = *reinterpret_cast<const double2 *>(oq), r1 = *reinterpret_cast<const double2 *>(oq + p.W);
          double v[4] = {r0.x, r0.y, r1.x, r1.y};
          double best = 0.0, bestpre = 0.0;
          int bi = 0;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            double rl = (v[k] <= 0.0 ? 0.0 : 1.0) * v[k];
            if (k == 0 || rl > best) { best = rl; bestpre = v[k]; bi = k; }
          }
          int oi = q * HoWo + jj;
          yb[oi] = best;
          cb[oi] = (uint8_t)(bi | (bestpre > 0.0 ? 0x80 : 0));
          q += dq;
          jj += dj;
          if (jj >= Wo) { jj -= Wo; q++; }
        }
      }
      __syncwarp();
    }
    buf ^= 1;
    n = n2;
    rp = rp2;
  }
}

// --------------------------------------------------------------------- dKrn
// G groups of 16 output channels (their accumulators all live in registers, the
// expanded cotangent tile is rebuilt per group), NPN column fragments of 8
// (taps + the ones column): G * NPN <= 8.
template <int G, int NPN>
__global__ void __launch_bounds__(NTHR, 1) strip_dk_kernel(const __grid_constant__ StripP p) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int fr = lane >> 2, fk = lane & 3;
  int *lut = reinterpret_cast<int *>(smem);                            // [64] position -> strip byte offset
  double *mine = smem + 32 + warp * p.warp_doubles;
  double *dbx = mine + 2 * p.SD;                                       // [16][kPS]
  const int Wo = p.W >> 1, Ho = p.H >> 1, W2 = 2 * p.W;

  for (int i = tid; i < 64; i += NTHR) {
    int pc = i < W2 ? i : 0;
    int yy = pc / p.W, x = pc - yy * p.W;
    lut[i] = (yy * kRS + x) * 8;
  }
  for (int i = lane; i < p.warp_doubles; i += 32) mine[i] = 0.0;
  __syncthreads();

  // column t = 8 j + fr of pair j: tap t (< KK), the ones column (t == KK), or padding
  int pair_off[NPN];
  bool ones[NPN];
#pragma unroll
  for (int j = 0; j < NPN; j++) {
    int t = j * 8 + fr;
    pair_off[j] = t < p.KK ? ((t / p.KW) * kRS + (t % p.KW)) * 8 : 0;
    ones[j] = t == p.KK;
  }
  double acc[G][2][NPN][2];
#pragma unroll
  for (int g = 0; g < G; g++)
#pragma unroll
    for (int m = 0; m < 2; m++)
#pragma unroll
      for (int j = 0; j < NPN; j++) acc[g][m][j][0] = acc[g][m][j][1] = 0.0;

  const int64_t gw = (int64_t)blockIdx.x * WPB + warp, nw = (int64_t)gridDim.x * WPB;
  const int HoWo = Ho * Wo;
  // pooled element e = lane + 32 k of a 16-channel group of an item: (channel, window) and its offsets
  int eoff[7], xoff[7], eq[7];
#pragma unroll
  for (int k = 0; k < 7; k++) {
    int e = lane + 32 * k;
    int q = e / Wo, jj = e - q * Wo;
    eq[k] = e < 16 * Wo ? q : 1 << 20;               // channel inside the group (beyond: never valid)
    eoff[k] = q * HoWo + jj;                         // offset inside the group's pooled block
    xoff[k] = q * kPS + 2 * jj;                      // first position of the window in the cotangent tile
  }
  double dyv[7];
  unsigned cdv[7];
  auto fetch = [&](int n, int rp, int g) {  // pooled cotangent + code of (item, group) -> registers
    const int64_t base = (((int64_t)n * p.Q + g * 16) * Ho + rp) * Wo;
    const double *yb = p.dy + base;
    const uint8_t *cb = p.code + base;
    const int qleft = p.Q - g * 16;
#pragma unroll
    for (int k = 0; k < 7; k++) {
      dyv[k] = 0.0;
      cdv[k] = 0;
      if (eq[k] < qleft) {
        dyv[k] = yb[eoff[k]];
        cdv[k] = cb[eoff[k]];
      }
    }
  };
  int n = (int)(gw / Ho), rp = (int)(gw - (int64_t)n * Ho);
  const int dn = (int)(nw / Ho), drp = (int)(nw - (int64_t)dn * Ho);
  int buf = 0;
  if (gw < p.items) {
    load_strip(mine, p, n, rp, lane);
    fetch(n, rp, 0);
  }
  cp_commit();
  for (int64_t it = gw; it < p.items; it += nw) {
    int n2 = n + dn, rp2 = rp + drp;
    if (rp2 >= Ho) { rp2 -= Ho; n2++; }
    const bool more = it + nw < p.items;
    cp_wait<0>();                              // this item's strip (issued one item ago)
    __syncwarp();
    if (more) load_strip(mine + (buf ^ 1) * p.SD, p, n2, rp2, lane);
    cp_commit();
    const char *As = reinterpret_cast<const char *>(mine + buf * p.SD);
    const double *Bs = dbx + fr * kPS + fk;
#pragma unroll
    for (int g = 0; g < G; g++) {
      // expand this (item, group)'s (dy, code) into the cotangent tile dbx[o][position]
#pragma unroll
      for (int k = 0; k < 7; k++) {
        if (eq[k] < 16) {
          double gv = (cdv[k] & 0x80u) ? dyv[k] : 0.0;
          int bi = cdv[k] & 0x7f;
          double *xp = dbx + xoff[k];
          *reinterpret_cast<double2 *>(xp) = make_double2(bi == 0 ? gv : 0.0, bi == 1 ? gv : 0.0);
          *reinterpret_cast<double2 *>(xp + p.W) = make_double2(bi == 2 ? gv : 0.0, bi == 3 ? gv : 0.0);
        }
      }
      // the next (item, group)'s values travel during this group's DMMAs
      if (g + 1 < G) fetch(n, rp, g + 1);
      else if (more) fetch(n2, rp2, 0);
      __syncwarp();
      for (int k0 = 0; k0 < W2; k0 += 4) {
        const char *ap = As + lut[k0 + fk];
        double a0 = Bs[k0], a1 = Bs[8 * kPS + k0];
#pragma unroll
        for (int j = 0; j < NPN; j++) {
          double b = ones[j] ? 1.0 : *reinterpret_cast<const double *>(ap + pair_off[j]);
          dmma884(acc[g][0][j][0], acc[g][0][j][1], a0, b);
          dmma884(acc[g][1][j][0], acc[g][1][j][1], a1, b);
        }
      }
      __syncwarp();
    }
    buf ^= 1;
    n = n2;
    rp = rp2;
  }
  cp_wait<0>();
  // the 16 warps of the CTA are summed in fixed order through shared memory
  __syncthreads();
  double *red = smem + 32;                  // [warp][g][m][j][lane][2]
#pragma unroll
  for (int g = 0; g < G; g++)
#pragma unroll
    for (int m = 0; m < 2; m++)
#pragma unroll
      for (int j = 0; j < NPN; j++)
        *reinterpret_cast<double2 *>(red + ((((warp * G + g) * 2 + m) * NPN + j) * 32 + lane) * 2) =
            make_double2(acc[g][m][j][0], acc[g][m][j][1]);
  __syncthreads();
  const int rowlen = p.KK + 1;
  double *part = p.part + (int64_t)blockIdx.x * p.Q * rowlen;
  // (g, m, j) triples are dealt to the warps; each sums the 16 warps in order
  for (int tr = warp; tr < G * 2 * NPN; tr += WPB) {
    const int g = tr / (2 * NPN), m = (tr / NPN) % 2, j = tr % NPN;
    double s0 = 0.0, s1 = 0.0;
    for (int w = 0; w < WPB; w++) {
      double2 v = *reinterpret_cast<const double2 *>(red + ((((w * G + g) * 2 + m) * NPN + j) * 32 + lane) * 2);
      s0 += v.x;
      s1 += v.y;
    }
    int o = g * 16 + m * 8 + fr, t = j * 8 + fk * 2;   // C fragment: row = o, cols = 2 taps
    if (o < p.Q) {
      if (t < rowlen) part[o * rowlen + t] = s0;
      if (t + 1 < rowlen) part[o * rowlen + t + 1] = s1;
    }
  }
}

// out: dK[o][t] (t < KK) and dbias[o] from the per-CTA partials: one warp per
// output element, lanes take the partials round-robin (each in ascending CTA
// order), then a fixed xor-shuffle tree -> deterministic
__global__ void __launch_bounds__(128) strip_dk_reduce_kernel(const double *__restrict__ part, int nparts, int Q, int KK,
                                                              double *__restrict__ dK, double *__restrict__ dbias) {
  const int rowlen = KK + 1, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (i >= Q * rowlen) return;
  double s = 0.0;
  for (int k = lane; k < nparts; k += 32) s += part[(int64_t)k * Q * rowlen + i];
  s = hb_warp_sum(s);
  if (lane == 0) {
    int o = i / rowlen, t = i - o * rowlen;
    if (t < KK) dK[o * KK + t] = s;
    else if (dbias) dbias[o] = s;
  }
}

// forward weights: Wt[t][q] = K[q, 0, t / KW, t % KW], tap stride TS, zero padded
__global__ void strip_pack_kernel(const double *__restrict__ K, double *__restrict__ wp, int Q, int KK, int NT, int TS,
                                  int64_t k_so, int KW, int64_t k_si, int64_t k_sj) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NT * TS) return;
  int t = i / TS, q = i - t * TS;
  wp[i] = (t < KK && q < Q) ? K[q * k_so + (t / KW) * k_si + (t % KW) * k_sj] : 0.0;
}

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

static bool strip_applicable(int64_t N, int64_t Ci, int64_t Co, int64_t H, int64_t W, int64_t KH, int64_t KW) {
  return Ci == 1 && Co >= 1 && Co <= 64 && (H % 2) == 0 && (W % 2) == 0 && 2 * W > 48 && 2 * W <= 56 && KH >= 1 &&
         KW >= 1 && KH * KW <= 31 && KH <= 8 && KW <= 8 && N * (H / 2) < (1ll << 31) && N * Co * H * W < (1ll << 31);
}

static void strip_geometry(StripP &p, int N, int Q, int H, int W, int KH, int KW) {
  memset(&p, 0, sizeof p);
  p.N = N; p.Q = Q; p.H = H; p.W = W; p.KH = KH; p.KW = KW; p.KK = KH * KW;
  p.RS = kRS;
  p.TR = 2 + KH - 1;
  p.SD = p.TR * kRS;
  p.items = N * (H / 2);
}

}  // namespace

// maxPool 2 2 (relu (conv2dPreserving K A + bias)) for a one-channel A; *done = 1 when it ran.
int hb_conv_strip_fwd(const hb_array *K, const hb_array *bias, const hb_array *A, hb_array *y, hb_array *code,
                      int *done) {
  *done = 0;
  if (A->dt != HB_F64 || !hb_contig(A) || !hb_contig(y) || !hb_contig(code) || (bias && !hb_contig(bias))) return HB_OK;
  if (!strip_applicable(A->shape[0], A->shape[1], K->shape[0], A->shape[2], A->shape[3], K->shape[2], K->shape[3]))
    return HB_OK;
  if (((uintptr_t)hb_data(A) & 15) != 0) return HB_OK;
  StripP p;
  strip_geometry(p, (int)A->shape[0], (int)K->shape[0], (int)A->shape[2], (int)A->shape[3], (int)K->shape[2],
                 (int)K->shape[3]);
  const int NQ = p.Q <= 8 ? 1 : 2, QT = NQ * 8;
  p.QG = (p.Q + QT - 1) / QT;
  p.NT = round_up(p.KK, 4);
  p.TS = p.QG * QT + 4;
  p.w_doubles = round_up(p.NT * p.TS, 2);
  p.OPS = kOPS;
  p.warp_doubles = round_up(2 * p.SD + QT * kOPS, 2);
  p.A = (const double *)hb_data(A);
  p.bias = bias ? (const double *)hb_data(bias) : nullptr;
  p.y = (double *)hb_data(y);
  p.code = (uint8_t *)hb_data(code);
  void *scr;
  HB_TRY(hb_scratch((size_t)p.w_doubles * 8, &scr));
  p.wp = (const double *)scr;
  strip_pack_kernel<<<(p.NT * p.TS + 255) / 256, 256, 0, g_hb.stream>>>((const double *)hb_data(K), (double *)scr, p.Q,
                                                                        p.KK, p.NT, p.TS, K->strides[0], p.KW,
                                                                        K->strides[2], K->strides[3]);
  HB_LAUNCHED();
  int smem = (p.w_doubles + ((p.NT + 1) >> 1) + WPB * p.warp_doubles) * 8;
  if (smem > 227 * 1024) return HB_OK;
  int64_t wneed = ((int64_t)p.items + WPB - 1) / WPB;
  int grid = (int)(wneed < g_hb.sm_count ? wneed : g_hb.sm_count);
#define SFW_LAUNCH(NQ_, M_)                                                                                    \
  do {                                                                                                        \
    HB_CUDA(cudaFuncSetAttribute(strip_fwd_kernel<NQ_, M_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    strip_fwd_kernel<NQ_, M_><<<grid, NTHR, smem, g_hb.stream>>>(p);                                         \
  } while (0)
  if (NQ == 1) SFW_LAUNCH(1, false);
  else if (p.QG == 1) SFW_LAUNCH(2, false);
  else SFW_LAUNCH(2, true);
#undef SFW_LAUNCH
  HB_LAUNCHED();
  *done = 1;
  return HB_OK;
}

// dK [Co,1,KH,KW] and (optionally) dbias [Co] from A [N,1,H,W], the pooled cotangent dy and its code bytes.
int hb_conv_strip_dk(const hb_array *A, const hb_array *dy, const hb_array *code, int KH, int KW, hb_array *dK,
                     hb_array *dbias, int *done) {
  *done = 0;
  if (A->dt != HB_F64 || !hb_contig(A) || !hb_contig(dy) || !hb_contig(code) || !hb_contig(dK) ||
      (dbias && !hb_contig(dbias)))
    return HB_OK;
  if (!strip_applicable(A->shape[0], A->shape[1], dy->shape[1], A->shape[2], A->shape[3], KH, KW)) return HB_OK;
  if (((uintptr_t)hb_data(A) & 15) != 0) return HB_OK;
  StripP p;
  strip_geometry(p, (int)A->shape[0], (int)dy->shape[1], (int)A->shape[2], (int)A->shape[3], KH, KW);
  // (channel groups, column fragments): the accumulators of all groups stay in registers
  const int G = (p.Q + 15) / 16, NPN = (p.KK + 1 + 7) / 8;
  const int Gt = G <= 1 ? 1 : (G <= 2 ? 2 : 4), Nt = NPN <= 2 ? 2 : 4;
  if (Gt * Nt > 8 || NPN > 4) return HB_OK;
  p.QG = Gt;
  p.NT = Nt * 8;
  p.PS = kPS;
  p.warp_doubles = round_up(2 * p.SD + 16 * kPS, 2);
  if (p.warp_doubles < Gt * 2 * Nt * 64) p.warp_doubles = Gt * 2 * Nt * 64;   // end-of-kernel reduction area
  p.A = (const double *)hb_data(A);
  p.dy = (const double *)hb_data(dy);
  p.code = (uint8_t *)hb_data(code);
  int smem = (32 + WPB * p.warp_doubles) * 8;
  if (smem > 227 * 1024) return HB_OK;
  int64_t wneed = ((int64_t)p.items + WPB - 1) / WPB;
  int grid = (int)(wneed < g_hb.sm_count ? wneed : g_hb.sm_count);
  const int rowlen = p.KK + 1;
  void *scr;
  HB_TRY(hb_scratch((size_t)grid * p.Q * rowlen * 8, &scr));
  p.part = (double *)scr;
#define SDK_LAUNCH(G_, N_)                                                                                   \
  do {                                                                                                       \
    HB_CUDA(cudaFuncSetAttribute(strip_dk_kernel<G_, N_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    strip_dk_kernel<G_, N_><<<grid, NTHR, smem, g_hb.stream>>>(p);                                          \
  } while (0)
  if (Gt == 1 && Nt == 4) SDK_LAUNCH(1, 4);
  else if (Gt == 1) SDK_LAUNCH(1, 2);
  else if (Gt == 2 && Nt == 4) SDK_LAUNCH(2, 4);
  else if (Gt == 2) SDK_LAUNCH(2, 2);
  else SDK_LAUNCH(4, 2);
#undef SDK_LAUNCH
  HB_LAUNCHED();
  strip_dk_reduce_kernel<<<(p.Q * rowlen + 3) / 4, 128, 0, g_hb.stream>>>(
      p.part, grid, p.Q, p.KK, (double *)hb_data(dK), dbias ? (double *)hb_data(dbias) : nullptr);
  HB_LAUNCHED();
  *done = 1;
  return HB_OK;
}

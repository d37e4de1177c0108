// gemm_dmma.cu -- dense FP64 GEMM on the DMMA tensor pipe for the plain
// matrix contractions of the hot path: tsmatmul2 / tsmatvecmul-like sdot1In
// over 2-D (possibly transposed) operands (OpsConcrete.hs:234-242), i.e. the
// dense layers of MnistCnnShaped2 / MnistFcnnRanked2 and the recurrence of
// MnistRnnShaped2.
//
//   C[m, n] = sum_k A[m, k] * B[k, n]        (any row/column strides with one
//                                             unit stride per operand)
//
// 64x64 CTA tile, 4 warps (32x32 each = 4x4 DMMA m8n8k4 fragments), K chunks
// of 16 staged by 16-byte cp.async through a 3-stage ring; shared-memory rows
// padded to a stride = 4 (mod 8) doubles so fragment loads are conflict-free
// in either operand orientation.  Small output grids are split along K across
// CTAs (grid.z) with partials reduced in a fixed order (deterministic).
// Anything else (batch axes, non-unit strides) stays on the generic strided
// engine in contract.cu.
#include <algorithm>

#include "common.cuh"
#include "contract.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, STAGES = 3;
constexpr int LDK = BK + 4;   // [row][k] layout (k contiguous in memory)
constexpr int LDM = BM + 4;   // [k][row] layout (row contiguous in memory)
constexpr int TILE = 64 * LDK > BK * LDM ? 64 * LDK : BK * LDM;  // doubles per operand per stage

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
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

constexpr int GMAX = 4;   // products per launch (hb_gemm_dmma_group)

struct GemmP {
  const double *A, *B;
  double *C;            // C or the split-K partial buffer [split][M][N]
  int M, N, K;
  int64_t a_sm, a_sk;   // strides of A[m,k]
  int64_t b_sk, b_sn;   // strides of B[k,n]
  int64_t c_sm;         // row stride of C (columns contiguous)
  int kchunks;          // BK-chunks per split
  int a_vec, b_vec;     // 16-byte cp.async allowed
  int *flags;           // serial split-K: per output tile, the number of K slices already accumulated into C
  int accum;            // C += A x B (taddTarget fused into the product: cotangent accumulation of a weight)
  // A GROUP of up to GMAX products of one [M, N] in one launch (hb_gemm_dmma_group / gemm8_group_kernel; the pair of
  // wX x + wS s, the three input cotangents of one time step of the unrolled recurrence, ...).  Every product is
  // accumulated into the result it names; the K chunks of all products of one result, in product order, are ONE
  // sequence (its chain) that is cut into K slices: a CTA is one (tile, result, slice) and walks from one product
  // into the next inside its slice.  gch[g] = chunks of product g, gnext[g] = the next product of the same result
  // (-1: none), rfirst[r] = the chain's first product, rT[r] its chunks, rkch[r] chunks per slice, rzbeg[r] the
  // first blockIdx.z of result r, rzlen[r] its slices; rC / raccum / rbias: the result, whether the chain starts
  // from its old contents, the bias of the tanh epilogue of the last slice (C = tanh (C + bias[m]): the body of
  // rnnMnistLayerS without a separate elementwise pass).
  int nprod, nres;
  const double *gA[GMAX], *gB[GMAX];
  int64_t ga_sm[GMAX], ga_sk[GMAX], gb_sk[GMAX], gb_sn[GMAX];
  int gK[GMAX], ga_vec[GMAX], gb_vec[GMAX];
  int gch[GMAX], gnext[GMAX];
  int rfirst[GMAX], rT[GMAX], rkch[GMAX], rzbeg[GMAX], rzlen[GMAX], raccum[GMAX];
  double *rC[GMAX];
  const double *rbias[GMAX];
  // A LIST of products of one shape accumulated into one result (hb_gemm_dmma_list: the weight cotangents
  // dW = sum_t d_t x_t^T of an unrolled loop): the K axis is the concatenation of the entries' K axes
  // (K % BK8 == 0, so a chunk never straddles two entries); slices go to partial buffers + the fixed-order reduce.
  int nlist;
  const double *Al[32], *Bl[32];
  // epilogue of an unsplit single product: C = tanh (C + bias[m])
  const double *epi_bias;
};

// Stage one operand tile: `rows` x BK where the operand's (row, k) element is
// at base[row * s_row + k * s_k]; out-of-range elements are zero.  KFAST:
// s_k == 1, shared layout [row][LDK]; else s_row == 1, layout [k][LDM].
template <bool KFAST, int NT>
__device__ __forceinline__ void stage_tile(double *dst, const double *__restrict__ base, int64_t s_row, int64_t s_k,
                                           int row0, int nrows, int k0, int K, int vec, int tid) {
  if (KFAST) {
    // 64 rows x 16 k: 8 x 16-byte pieces per row
    for (int idx = tid; idx < 64 * 8; idx += NT) {
      int r = idx >> 3, kk = (idx & 7) * 2;
      double *d = dst + r * LDK + kk;
      int row = row0 + r, k = k0 + kk;
      const double *s = base + (int64_t)row * s_row + k;
      if (row < nrows && k + 1 < K && vec) cp16(d, s);
      else {
        d[0] = (row < nrows && k < K) ? s[0] : 0.0;
        d[1] = (row < nrows && k + 1 < K) ? s[1] : 0.0;
      }
    }
  } else {
    // 16 k x 64 rows: 32 x 16-byte pieces per k
    for (int idx = tid; idx < 16 * 32; idx += NT) {
      int kk = idx >> 5, r = (idx & 31) * 2;
      double *d = dst + kk * LDM + r;
      int row = row0 + r, k = k0 + kk;
      const double *s = base + (int64_t)k * s_k + row;
      if (k < K && row + 1 < nrows && vec) cp16(d, s);
      else {
        d[0] = (k < K && row < nrows) ? s[0] : 0.0;
        d[1] = (k < K && row + 1 < nrows) ? s[1] : 0.0;
      }
    }
  }
}

// KS = 2: a second group of four warps takes every other k4 step of each
// chunk (intra-CTA split-K, summed through shared memory in fixed order at the
// end): twice the warps per SM when the output grid is too small to fill them.
template <bool A_KFAST, bool B_KFAST, int KS>
__global__ void __launch_bounds__(128 * KS) gemm_dmma_kernel(const __grid_constant__ GemmP p) {
  constexpr int NT = 128 * KS;
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kg = warp >> 2, w4 = warp & 3;
  const int wm = (w4 & 1) * 32, wn = (w4 >> 1) * 32;
  const int fr = lane >> 2, fk = lane & 3;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * p.kchunks * BK;
  int kend = kbeg + p.kchunks * BK;
  if (kend > p.K) kend = p.K;
  const int nk = kend > kbeg ? (kend - kbeg + BK - 1) / BK : 0;

  auto issue = [&](int c) {
    double *as = smem + (c % STAGES) * 2 * TILE, *bs = as + TILE;
    stage_tile<A_KFAST, NT>(as, p.A, p.a_sm, p.a_sk, m0, p.M, kbeg + c * BK, kend, p.a_vec, tid);
    stage_tile<B_KFAST, NT>(bs, p.B, p.b_sn, p.b_sk, n0, p.N, kbeg + c * BK, kend, p.b_vec, tid);
    cp_commit();
  };
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int c = 0; c < STAGES - 1; c++) {
    if (c < nk) issue(c);
    else cp_commit();
  }
  for (int c = 0; c < nk; c++) {
    cp_wait<STAGES - 2>();
    __syncthreads();
    if (c + STAGES - 1 < nk) issue(c + STAGES - 1);
    else cp_commit();
    const double *as = smem + (c % STAGES) * 2 * TILE, *bs = as + TILE;
#pragma unroll
    for (int k4 = kg * 4; k4 < BK; k4 += 4 * KS) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++)
        a[i] = A_KFAST ? as[(wm + i * 8 + fr) * LDK + k4 + fk] : as[(k4 + fk) * LDM + wm + i * 8 + fr];
#pragma unroll
      for (int j = 0; j < 4; j++)
        b[j] = B_KFAST ? bs[(wn + j * 8 + fr) * LDK + k4 + fk] : bs[(k4 + fk) * LDM + wn + j * 8 + fr];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  if (KS == 2) {
    __syncthreads();  // every warp is done with the stage buffers
    double *red = smem + ((w4 * 16) * 32 + lane) * 2;  // [w4][i * 4 + j][lane][2]
    if (kg == 1) {
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
          *reinterpret_cast<double2 *>(red + (i * 4 + j) * 64) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
    __syncthreads();
    if (kg == 1) return;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        double2 t = *reinterpret_cast<const double2 *>(red + (i * 4 + j) * 64);
        acc[i][j][0] += t.x;
        acc[i][j][1] += t.y;
      }
  }
  double *C = p.C + (int64_t)blockIdx.z * p.M * p.c_sm;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int m = m0 + wm + i * 8 + fr, n = n0 + wn + j * 8 + fk * 2;
      if (m < p.M) {
        double *c = C + (int64_t)m * p.c_sm + n;
        const bool addc = p.accum && gridDim.z == 1;  // with K slices the reduce kernel adds the old C
        if (n + 1 < p.N && (p.c_sm & 1) == 0 && ((uintptr_t)C & 15) == 0) {
          double2 o = make_double2(acc[i][j][0], acc[i][j][1]);
          if (addc) {
            double2 t = *reinterpret_cast<const double2 *>(c);
            o.x = t.x + o.x;
            o.y = t.y + o.y;
          }
          *reinterpret_cast<double2 *>(c) = o;
        } else {
          if (n < p.N) c[0] = addc ? c[0] + acc[i][j][0] : acc[i][j][0];
          if (n + 1 < p.N) c[1] = addc ? c[1] + acc[i][j][1] : acc[i][j][1];
        }
      }
    }
}

// ---------------------------------------------------------------------------
// Eight-warp variant: the same 64x64 CTA tile, but every warp owns a 32x16
// piece over the FULL k range of a 32-deep chunk (no intra-CTA split-K, so no
// shared-memory reduction), twice the DMMAs per barrier and per staged byte.
// For output grids of about one tile per SM (the 512-wide recurrence GEMMs of
// MnistRnnShaped2: 128 tiles) it replaces split-K across CTAs + a reduce
// kernel: one launch, partial sums never leave the registers.
constexpr int BK8 = 32, STAGES8 = 3;
constexpr int LDK8 = BK8 + 4, LDM8 = 64 + 4;
constexpr int TILE8 = 64 * LDK8 > BK8 * LDM8 ? 64 * LDK8 : BK8 * LDM8;

template <bool KFAST>
__device__ __forceinline__ void stage_tile8(double *dst, const double *__restrict__ base, int64_t s_row, int64_t s_k,
                                            int row0, int nrows, int k0, int K, int vec, int tid) {
  if (KFAST) {  // 64 rows x 32 k: 16 pieces per row
#pragma unroll
    for (int it = 0; it < 4; it++) {
      int idx = tid + it * 256;
      int r = idx >> 4, kk = (idx & 15) * 2;
      double *d = dst + r * LDK8 + kk;
      int row = row0 + r, k = k0 + kk;
      const double *s = base + (int64_t)row * s_row + k;
      if (row < nrows && k + 1 < K && vec) cp16(d, s);
      else {
        d[0] = (row < nrows && k < K) ? s[0] : 0.0;
        d[1] = (row < nrows && k + 1 < K) ? s[1] : 0.0;
      }
    }
  } else {      // 32 k x 64 rows: 32 pieces per k
#pragma unroll
    for (int it = 0; it < 4; it++) {
      int idx = tid + it * 256;
      int kk = idx >> 5, r = (idx & 31) * 2;
      double *d = dst + kk * LDM8 + r;
      int row = row0 + r, k = k0 + kk;
      const double *s = base + (int64_t)k * s_k + row;
      if (k < K && row + 1 < nrows && vec) cp16(d, s);
      else {
        d[0] = (k < K && row < nrows) ? s[0] : 0.0;
        d[1] = (k < K && row + 1 < nrows) ? s[1] : 0.0;
      }
    }
  }
}

template <bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__(256, 2) gemm8_dmma_kernel(const __grid_constant__ GemmP p) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 16;
  const int fr = lane >> 2, fk = lane & 3;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  // which product this K slice belongs to, and its slice index there
  const double *__restrict__ gA = p.A, *__restrict__ gB = p.B;
  int64_t a_sm = p.a_sm, a_sk = p.a_sk, b_sk = p.b_sk, b_sn = p.b_sn;
  int a_vec = p.a_vec, b_vec = p.b_vec, Kp = p.nlist ? p.nlist * p.K : p.K, kch = p.kchunks;
  const int zl = (int)blockIdx.z;
  const int kbeg = zl * kch * BK8;
  int kend = kbeg + kch * BK8;
  if (kend > Kp) kend = Kp;
  const int nk = kend > kbeg ? (kend - kbeg + BK8 - 1) / BK8 : 0;

  const bool fast = a_vec && b_vec && m0 + BM <= p.M && n0 + BN <= p.N;
  const int64_t toff_a = A_KFAST ? (int64_t)(m0 + (tid >> 4)) * a_sm + (tid & 15) * 2 : (int64_t)(tid >> 5) * a_sk + m0 + (tid & 31) * 2;
  const int64_t toff_b = B_KFAST ? (int64_t)(n0 + (tid >> 4)) * b_sn + (tid & 15) * 2 : (int64_t)(tid >> 5) * b_sk + n0 + (tid & 31) * 2;
  const int64_t step_a = A_KFAST ? 16 * a_sm : 8 * a_sk, step_b = B_KFAST ? 16 * b_sn : 8 * b_sk;
  const int64_t kmul_a = A_KFAST ? 1 : a_sk, kmul_b = B_KFAST ? 1 : b_sk;
  const int da = A_KFAST ? (tid >> 4) * LDK8 + (tid & 15) * 2 : (tid >> 5) * LDM8 + (tid & 31) * 2;
  const int db = B_KFAST ? (tid >> 4) * LDK8 + (tid & 15) * 2 : (tid >> 5) * LDM8 + (tid & 31) * 2;
  constexpr int dstep_a = A_KFAST ? 16 * LDK8 : 8 * LDM8, dstep_b = B_KFAST ? 16 * LDK8 : 8 * LDM8;
  auto issue = [&](int c) {
    double *as = smem + (c % STAGES8) * 2 * TILE8, *bs = as + TILE8;
    const double *__restrict__ cA = gA, *__restrict__ cB = gB;
    int k0 = kbeg + c * BK8, ke = kend;
    if (p.nlist) {   // which entry of the list this chunk lies in
      const int t = k0 / p.K;
      k0 -= t * p.K;
      ke = p.K;
      cA = p.Al[t];
      cB = p.Bl[t];
    }
    if (fast && k0 + BK8 <= ke) {
      // interior chunk of an interior tile with 16-byte aligned rows: this thread's four pieces per operand are
      // base + thread offset + chunk offset + it * step -- eight cp.async and a few adds instead of the
      // bounds-checked generic path
      const double *ga = cA + toff_a + (int64_t)k0 * kmul_a, *gb = cB + toff_b + (int64_t)k0 * kmul_b;
#pragma unroll
      for (int it = 0; it < 4; it++) cp16(as + da + it * dstep_a, ga + it * step_a);
#pragma unroll
      for (int it = 0; it < 4; it++) cp16(bs + db + it * dstep_b, gb + it * step_b);
    } else {
      stage_tile8<A_KFAST>(as, cA, a_sm, a_sk, m0, p.M, k0, ke, a_vec, tid);
      stage_tile8<B_KFAST>(bs, cB, b_sn, b_sk, n0, p.N, k0, ke, b_vec, tid);
    }
    cp_commit();
  };
  double acc[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int c = 0; c < STAGES8 - 1; c++) {
    if (c < nk) issue(c);
    else cp_commit();
  }
  // this lane's fragment element inside a stage (k part included)
  const int aoff = A_KFAST ? (wm + fr) * LDK8 + fk : fk * LDM8 + wm + fr;
  const int boff = B_KFAST ? (wn + fr) * LDK8 + fk : fk * LDM8 + wn + fr;
  constexpr int a_i = A_KFAST ? 8 * LDK8 : 8, a_k = A_KFAST ? 4 : 4 * LDM8;
  constexpr int b_j = B_KFAST ? 8 * LDK8 : 8, b_k = B_KFAST ? 4 : 4 * LDM8;
  for (int c = 0; c < nk; c++) {
    cp_wait<STAGES8 - 2>();
    __syncthreads();
    if (c + STAGES8 - 1 < nk) issue(c + STAGES8 - 1);
    else cp_commit();
    const double *as = smem + (c % STAGES8) * 2 * TILE8 + aoff, *bs = smem + (c % STAGES8) * 2 * TILE8 + TILE8 + boff;
#pragma unroll
    for (int k4 = 0; k4 < BK8 / 4; k4++) {
      double a[4], b[2];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = as[i * a_i + k4 * a_k];
#pragma unroll
      for (int j = 0; j < 2; j++) b[j] = bs[j * b_j + k4 * b_k];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  // K slices of one tile: separate partial buffers (p.flags == NULL, reduced by gemm_reduce_kernel), or
  // accumulated straight into C in slice order ("serial split-K": slice z waits until z slices are in,
  // adds its registers to what they left, publishes z + 1; the last slice resets the flag) -- the order of
  // the additions is fixed, so the result is deterministic, and the partial sums never travel through HBM
  // more than once.  CTAs are dispatched in block-index order, so the slice waited for is always running.
  const bool serial = p.flags != nullptr;
  double *C = serial ? p.C : p.C + (int64_t)blockIdx.z * p.M * p.c_sm;
  int *flag = serial ? p.flags + blockIdx.y * gridDim.x + blockIdx.x : nullptr;
  int z = blockIdx.z, zlen = gridDim.z;   // position in / length of this tile's serial chain
  int accum = p.accum;
  const double *bias = p.epi_bias;
  if (serial && z > 0) {
    if (tid == 0) {
      while (*reinterpret_cast<volatile int *>(flag) != z) __nanosleep(64);
      __threadfence();
    }
    __syncthreads();
  }
  const bool addc = serial ? (z > 0 || accum) : (accum && gridDim.z == 1);
  const bool cvec = (p.c_sm & 1) == 0 && ((uintptr_t)C & 15) == 0;
  const bool epi = bias != nullptr && (serial ? z + 1 == zlen : gridDim.z == 1);
  // three phases so that the eight loads, the sixteen tanh evaluations and the eight stores of a thread each overlap
  // among themselves (a load may not be hoisted over the previous store by the compiler)
  if (addc) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int m = m0 + wm + i * 8 + fr, n = n0 + wn + j * 8 + fk * 2;
        if (m < p.M) {
          const double *c = C + (int64_t)m * p.c_sm + n;
          if (n + 1 < p.N && cvec) {
            const double2 t = __ldcg(reinterpret_cast<const double2 *>(c));
            acc[i][j][0] = t.x + acc[i][j][0];
            acc[i][j][1] = t.y + acc[i][j][1];
          } else {
            if (n < p.N) acc[i][j][0] = __ldcg(c) + acc[i][j][0];
            if (n + 1 < p.N) acc[i][j][1] = __ldcg(c + 1) + acc[i][j][1];
          }
        }
      }
  }
  if (epi) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int m = m0 + wm + i * 8 + fr;
      const double eb = m < p.M ? bias[m] : 0.0;
#pragma unroll
      for (int j = 0; j < 2; j++) {
        acc[i][j][0] = hb_tanh_f64(acc[i][j][0] + eb);
        acc[i][j][1] = hb_tanh_f64(acc[i][j][1] + eb);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int m = m0 + wm + i * 8 + fr, n = n0 + wn + j * 8 + fk * 2;
      if (m < p.M) {
        double *c = C + (int64_t)m * p.c_sm + n;
        if (n + 1 < p.N && cvec) {
          __stcg(reinterpret_cast<double2 *>(c), make_double2(acc[i][j][0], acc[i][j][1]));
        } else {
          if (n < p.N) __stcg(c, acc[i][j][0]);
          if (n + 1 < p.N) __stcg(c + 1, acc[i][j][1]);
        }
      }
    }
  if (serial && zlen > 1) {
    __threadfence();
    __syncthreads();
    if (tid == 0) *reinterpret_cast<volatile int *>(flag) = z + 1 == zlen ? 0 : z + 1;
  }
}

// ---------------------------------------------------------------------------
// The group launch: one CTA per (tile, result, K slice of the result's chain).  Same tile, warp layout, pipeline and
// epilogue as gemm8_dmma_kernel; the difference is that a slice is cut from the CONCATENATED chunks of all products
// of the result, so a thin product (wX x with K = 28 next to wS s with K = 512) costs one more chunk of the CTA
// that also multiplies its neighbour, not a CTA (and a link of the serial chain) of its own.
template <bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__(256, 2) gemm8_group_kernel(const __grid_constant__ GemmP p) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 16;
  const int fr = lane >> 2, fk = lane & 3;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  // result, slice, chunk range inside the chain
  int r = 0;
#pragma unroll
  for (int q = 1; q < GMAX; q++)
    if (q < p.nres && (int)blockIdx.z >= p.rzbeg[q]) r = q;
  int zbeg = p.rzbeg[0], kch = p.rkch[0], T = p.rT[0], zlen = p.rzlen[0], accum = p.raccum[0], g = p.rfirst[0];
  double *C = p.rC[0];
  const double *bias = p.rbias[0];
#pragma unroll
  for (int q = 1; q < GMAX; q++)
    if (q == r) {
      zbeg = p.rzbeg[q]; kch = p.rkch[q]; T = p.rT[q]; zlen = p.rzlen[q]; accum = p.raccum[q]; g = p.rfirst[q];
      C = p.rC[q]; bias = p.rbias[q];
    }
  const int z = (int)blockIdx.z - zbeg;
  const int cbeg = z * kch;
  int cend = cbeg + kch;
  if (cend > T) cend = T;
  const int nk = cend > cbeg ? cend - cbeg : 0;
  // the product the first chunk lies in, and the chunk's index there
  int q0 = cbeg;
#pragma unroll
  for (int it = 0; it < GMAX - 1; it++) {
    int gch = p.gch[0], gnext = p.gnext[0];
#pragma unroll
    for (int h = 1; h < GMAX; h++)
      if (h == g) { gch = p.gch[h]; gnext = p.gnext[h]; }
    if (q0 >= gch && gnext >= 0) { q0 -= gch; g = gnext; }
  }
  // staging state of the product the issue position is in (this thread's four 16-byte pieces per operand and
  // chunk are base + it * step + the chunk's k offset; bounds-checked generic path for edges and unaligned rows)
  int st_g = -1, st_K = 0, st_ch = 0, st_next = -1, a_vec = 0, b_vec = 0;
  bool fast = false;
  const double *gA = nullptr, *gB = nullptr, *pa = nullptr, *pb = nullptr;
  int64_t a_sm = 0, a_sk = 0, b_sk = 0, b_sn = 0, step_a = 0, step_b = 0, kmul_a = 0, kmul_b = 0;
  const int da = A_KFAST ? (tid >> 4) * LDK8 + (tid & 15) * 2 : (tid >> 5) * LDM8 + (tid & 31) * 2;
  const int db = B_KFAST ? (tid >> 4) * LDK8 + (tid & 15) * 2 : (tid >> 5) * LDM8 + (tid & 31) * 2;
  constexpr int dstep_a = A_KFAST ? 16 * LDK8 : 8 * LDM8, dstep_b = B_KFAST ? 16 * LDK8 : 8 * LDM8;
  auto issue = [&](int c) {
    double *as = smem + (c % STAGES8) * 2 * TILE8, *bs = as + TILE8;
    if (g != st_g) {
      st_g = g;
      gA = p.gA[0]; gB = p.gB[0];
      a_sm = p.ga_sm[0]; a_sk = p.ga_sk[0]; b_sk = p.gb_sk[0]; b_sn = p.gb_sn[0];
      a_vec = p.ga_vec[0]; b_vec = p.gb_vec[0]; st_K = p.gK[0]; st_ch = p.gch[0]; st_next = p.gnext[0];
#pragma unroll
      for (int h = 1; h < GMAX; h++)
        if (h == g) {
          gA = p.gA[h]; gB = p.gB[h];
          a_sm = p.ga_sm[h]; a_sk = p.ga_sk[h]; b_sk = p.gb_sk[h]; b_sn = p.gb_sn[h];
          a_vec = p.ga_vec[h]; b_vec = p.gb_vec[h]; st_K = p.gK[h]; st_ch = p.gch[h]; st_next = p.gnext[h];
        }
      fast = a_vec && b_vec && m0 + BM <= p.M && n0 + BN <= p.N;
      if (A_KFAST) { pa = gA + (int64_t)(m0 + (tid >> 4)) * a_sm + (tid & 15) * 2; step_a = 16 * a_sm; kmul_a = 1; }
      else { pa = gA + (int64_t)(tid >> 5) * a_sk + m0 + (tid & 31) * 2; step_a = 8 * a_sk; kmul_a = a_sk; }
      if (B_KFAST) { pb = gB + (int64_t)(n0 + (tid >> 4)) * b_sn + (tid & 15) * 2; step_b = 16 * b_sn; kmul_b = 1; }
      else { pb = gB + (int64_t)(tid >> 5) * b_sk + n0 + (tid & 31) * 2; step_b = 8 * b_sk; kmul_b = b_sk; }
    }
    const int k0 = q0 * BK8;
    if (fast && k0 + BK8 <= st_K) {
      const double *ga = pa + (int64_t)k0 * kmul_a, *gb = pb + (int64_t)k0 * kmul_b;
#pragma unroll
      for (int it = 0; it < 4; it++) cp16(as + da + it * dstep_a, ga + it * step_a);
#pragma unroll
      for (int it = 0; it < 4; it++) cp16(bs + db + it * dstep_b, gb + it * step_b);
    } else {
      stage_tile8<A_KFAST>(as, gA, a_sm, a_sk, m0, p.M, k0, st_K, a_vec, tid);
      stage_tile8<B_KFAST>(bs, gB, b_sn, b_sk, n0, p.N, k0, st_K, b_vec, tid);
    }
    cp_commit();
    if (++q0 == st_ch && st_next >= 0) { q0 = 0; g = st_next; }   // on into the chain's next product
  };
  double acc[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int c = 0; c < STAGES8 - 1; c++) {
    if (c < nk) issue(c);
    else cp_commit();
  }
  const int aoff = A_KFAST ? (wm + fr) * LDK8 + fk : fk * LDM8 + wm + fr;
  const int boff = B_KFAST ? (wn + fr) * LDK8 + fk : fk * LDM8 + wn + fr;
  constexpr int a_i = A_KFAST ? 8 * LDK8 : 8, a_k = A_KFAST ? 4 : 4 * LDM8;
  constexpr int b_j = B_KFAST ? 8 * LDK8 : 8, b_k = B_KFAST ? 4 : 4 * LDM8;
  for (int c = 0; c < nk; c++) {
    cp_wait<STAGES8 - 2>();
    __syncthreads();
    if (c + STAGES8 - 1 < nk) issue(c + STAGES8 - 1);
    else cp_commit();
    const double *as = smem + (c % STAGES8) * 2 * TILE8 + aoff, *bs = smem + (c % STAGES8) * 2 * TILE8 + TILE8 + boff;
#pragma unroll
    for (int k4 = 0; k4 < BK8 / 4; k4++) {
      double a[4], b[2];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = as[i * a_i + k4 * a_k];
#pragma unroll
      for (int j = 0; j < 2; j++) b[j] = bs[j * b_j + k4 * b_k];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  // serial chain over the result's slices (see gemm8_dmma_kernel)
  int *flag = p.flags + r * (int)(gridDim.x * gridDim.y) + blockIdx.y * gridDim.x + blockIdx.x;
  if (z > 0) {
    if (tid == 0) {
      while (*reinterpret_cast<volatile int *>(flag) != z) __nanosleep(64);
      __threadfence();
    }
    __syncthreads();
  }
  const bool addc = z > 0 || accum;
  const bool cvec = (p.c_sm & 1) == 0 && ((uintptr_t)C & 15) == 0;
  const bool epi = bias != nullptr && z + 1 == zlen;
  if (addc) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int m = m0 + wm + i * 8 + fr, n = n0 + wn + j * 8 + fk * 2;
        if (m < p.M) {
          const double *c = C + (int64_t)m * p.c_sm + n;
          if (n + 1 < p.N && cvec) {
            const double2 t = __ldcg(reinterpret_cast<const double2 *>(c));
            acc[i][j][0] = t.x + acc[i][j][0];
            acc[i][j][1] = t.y + acc[i][j][1];
          } else {
            if (n < p.N) acc[i][j][0] = __ldcg(c) + acc[i][j][0];
            if (n + 1 < p.N) acc[i][j][1] = __ldcg(c + 1) + acc[i][j][1];
          }
        }
      }
  }
  if (epi) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int m = m0 + wm + i * 8 + fr;
      const double eb = m < p.M ? bias[m] : 0.0;
#pragma unroll
      for (int j = 0; j < 2; j++) {
        acc[i][j][0] = hb_tanh_f64(acc[i][j][0] + eb);
        acc[i][j][1] = hb_tanh_f64(acc[i][j][1] + eb);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int m = m0 + wm + i * 8 + fr, n = n0 + wn + j * 8 + fk * 2;
      if (m < p.M) {
        double *c = C + (int64_t)m * p.c_sm + n;
        if (n + 1 < p.N && cvec) {
          __stcg(reinterpret_cast<double2 *>(c), make_double2(acc[i][j][0], acc[i][j][1]));
        } else {
          if (n < p.N) __stcg(c, acc[i][j][0]);
          if (n + 1 < p.N) __stcg(c + 1, acc[i][j][1]);
        }
      }
    }
  if (zlen > 1) {
    __threadfence();
    __syncthreads();
    if (tid == 0) *reinterpret_cast<volatile int *>(flag) = z + 1 == zlen ? 0 : z + 1;
  }
}

__global__ void gemm_reduce_kernel(const double *__restrict__ part, double *__restrict__ out, int64_t n, int splits,
                                   int accum) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int k = 0; k < splits; k++) s += part[(int64_t)k * n + i];
  out[i] = accum ? out[i] + s : s;
}

}  // namespace

// per-tile slice counters of the serial split-K epilogue: zero between launches (the last slice
// of a tile resets its counter), allocated once -- outside any graph capture, since every captured
// program is run eagerly first
constexpr int64_t kFlagTiles = 1 << 16;
static int gemm_flags(int **out) {
  static int *flags = nullptr;
  if (!flags) {
    HB_REQUIRE(!g_hb.capturing, "gemm: first use inside a graph capture");
    HB_CUDA(cudaMalloc(&flags, kFlagTiles * sizeof(int)));
    HB_CUDA(cudaMemsetAsync(flags, 0, kFlagTiles * sizeof(int), g_hb.stream));
  }
  *out = flags;
  return HB_OK;
}

// C (contiguous [M,N] f64) = A x B with element strides; *done = 1 if handled.
int hb_gemm_dmma(const double *A, int64_t a_sm, int64_t a_sk, const double *B, int64_t b_sk, int64_t b_sn, double *C,
                 int64_t M, int64_t N, int64_t K, int accum, int *done) {
  *done = 0;
  if (M <= 0 || N <= 0 || K <= 0) return HB_OK;
  if (M >= (1ll << 31) || N >= (1ll << 31) || K >= (1ll << 31)) return HB_OK;
  bool a_kfast = a_sk == 1, b_kfast = b_sk == 1;
  if (!a_kfast && a_sm != 1) return HB_OK;
  if (!b_kfast && b_sn != 1) return HB_OK;
  // a size-1 axis may report any stride: prefer the k-fast reading when both are 1
  GemmP p;
  memset(&p, 0, sizeof p);
  p.A = A; p.B = B;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.a_sm = a_sm; p.a_sk = a_sk; p.b_sk = b_sk; p.b_sn = b_sn;
  p.c_sm = N;
  p.accum = accum;
  p.a_vec = ((uintptr_t)A & 15) == 0 && ((a_kfast ? a_sm : a_sk) % 2 == 0);
  p.b_vec = ((uintptr_t)B & 15) == 0 && ((b_kfast ? b_sn : b_sk) % 2 == 0);
  int gm = (int)((M + BM - 1) / BM), gn = (int)((N + BN - 1) / BN);
  if (gn > 65535) return HB_OK;
  {
    // eight-warp kernel (see gemm8_dmma_kernel); K is sliced across CTAs until two CTAs per SM are busy
    static const int force = getenv("HB_GEMM8") ? atoi(getenv("HB_GEMM8")) : -1;
    const int64_t tiles8 = (int64_t)gm * gn;
    const int chunks8 = (int)((K + BK8 - 1) / BK8);
    int splits = 1;
    if (tiles8 < 2 * (int64_t)g_hb.sm_count) {
      splits = (int)(2 * (int64_t)g_hb.sm_count / tiles8);
      if (splits > chunks8 / 4) splits = chunks8 / 4;  // at least 4 chunks (128 k) per slice
      if (splits < 1) splits = 1;
      if (splits > 64) splits = 64;
    }
    p.kchunks = (chunks8 + splits - 1) / splits;
    splits = (chunks8 + p.kchunks - 1) / p.kchunks;
    // serial split-K up to three slices ([128,4096]x[4096,3136], 98 tiles x 3: 125 vs 168 us with the four-warp
    // kernel + reduce); four slices over 64 tiles lose to it ([512,1024]x[1024,512]: 36 vs 29 us)
    static const int serial_max = getenv("HB_GEMM_SERIAL") ? atoi(getenv("HB_GEMM_SERIAL")) : 3;
    const bool serial = splits >= 2 && splits <= serial_max && tiles8 <= kFlagTiles;
    // measured (tools/gemm_sweep.py): the eight-warp kernel wins on large grids (2048^3: 563 vs 602 us,
    // [512,512]x[512,8192]: 145 vs 159 us), on ~one tile per SM with two K slices ([512,512]x[512,1024]:
    // 29 vs 33 us) and on very deep K over a few tiles ([64,4096]x[4096,784]: 27 vs 32 us); the four-warp
    // kernel with its finer split keeps the 64-tile / K = 1024 shapes (29 vs 36 us)
    const int64_t ctas8 = tiles8 * splits, slots8 = 2 * (int64_t)g_hb.sm_count;
    bool use8 = (splits == 1 && 3 * tiles8 >= 2 * slots8) || (splits == 2 && 5 * ctas8 >= 4 * slots8) ||
                (splits > 4 && chunks8 >= 64);
    if (serial && splits > 2) use8 = true;
    if (force >= 0) use8 = force != 0;
    if (use8) {
      p.C = C;
      p.flags = nullptr;
      if (serial) {
        HB_TRY(gemm_flags(&p.flags));
      } else if (splits > 1) {
        void *scr;
        HB_TRY(hb_scratch((size_t)splits * M * N * 8, &scr));
        p.C = (double *)scr;
      }
      dim3 grid((unsigned)gm, (unsigned)gn, (unsigned)splits);
      size_t smem = (size_t)STAGES8 * 2 * TILE8 * 8;
#define GEMM8_LAUNCH(AK, BKF)                                                                                  \
  do {                                                                                                         \
    HB_CUDA(cudaFuncSetAttribute(gemm8_dmma_kernel<AK, BKF>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                 (int)smem));                                                                  \
    gemm8_dmma_kernel<AK, BKF><<<grid, 256, smem, g_hb.stream>>>(p);                                           \
  } while (0)
      if (a_kfast && b_kfast) GEMM8_LAUNCH(true, true);
      else if (a_kfast) GEMM8_LAUNCH(true, false);
      else if (b_kfast) GEMM8_LAUNCH(false, true);
      else GEMM8_LAUNCH(false, false);
#undef GEMM8_LAUNCH
      HB_LAUNCHED();
      if (splits > 1 && !serial) {
        int64_t n = M * N;
        gemm_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, g_hb.stream>>>(p.C, C, n, splits, accum);
        HB_LAUNCHED();
      }
      *done = 1;
      return HB_OK;
    }
  }
  int chunks = (int)((K + BK - 1) / BK);
  int splits = 1;
  int64_t tiles = (int64_t)gm * gn;
  // fewer than two tiles per SM: split K across CTAs up to ~3 CTAs per SM (measured on
  // [512,512]x[512,1024]: 31 us + 8 us for the fixed-order reduce, against 44 us unsplit)
  if (tiles < 2 * g_hb.sm_count) {
    splits = (int)((3 * (int64_t)g_hb.sm_count + tiles - 1) / tiles);
    if (splits > chunks / 4) splits = chunks / 4;   // at least 4 chunks (64 k) per split
    if (splits < 1) splits = 1;
    if (splits > 64) splits = 64;
  }
  p.kchunks = (chunks + splits - 1) / splits;
  splits = (chunks + p.kchunks - 1) / p.kchunks;
  if (splits > 1) {
    void *scr;
    HB_TRY(hb_scratch((size_t)splits * M * N * 8, &scr));
    p.C = (double *)scr;
  } else {
    p.C = C;
  }
  dim3 grid((unsigned)gm, (unsigned)gn, (unsigned)splits);
  size_t smem = (size_t)STAGES * 2 * TILE * 8;
  // few tiles per SM: eight warps per CTA (intra-CTA split-K) instead of four
  const bool ks2 = tiles * splits < 3 * (int64_t)g_hb.sm_count;
#define GEMM_LAUNCH(AK, BKF)                                                                                  \
  do {                                                                                                        \
    if (ks2) {                                                                                                \
      HB_CUDA(cudaFuncSetAttribute(gemm_dmma_kernel<AK, BKF, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                   (int)smem));                                                               \
      gemm_dmma_kernel<AK, BKF, 2><<<grid, 256, smem, g_hb.stream>>>(p);                                      \
    } else {                                                                                                  \
      HB_CUDA(cudaFuncSetAttribute(gemm_dmma_kernel<AK, BKF, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                   (int)smem));                                                               \
      gemm_dmma_kernel<AK, BKF, 1><<<grid, 128, smem, g_hb.stream>>>(p);                                      \
    }                                                                                                         \
  } while (0)
  if (a_kfast && b_kfast) GEMM_LAUNCH(true, true);
  else if (a_kfast) GEMM_LAUNCH(true, false);
  else if (b_kfast) GEMM_LAUNCH(false, true);
  else GEMM_LAUNCH(false, false);
#undef GEMM_LAUNCH
  HB_LAUNCHED();
  if (splits > 1) {
    int64_t n = M * N;
    gemm_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, g_hb.stream>>>(p.C, C, n, splits, accum);
    HB_LAUNCHED();
  }
  *done = 1;
  return HB_OK;
}

// C (contiguous [M,N] f64) = (accum ? C : 0) + sum_t A_t x B_t for n products of one shape and stride pattern, in ONE
// launch of the eight-warp kernel + the fixed-order reduce: *done = 1 if handled.
int hb_gemm_dmma_list(int n, const double *const *A, int64_t a_sm, int64_t a_sk, const double *const *B, int64_t b_sk,
                      int64_t b_sn, double *C, int64_t M, int64_t N, int64_t K, int accum, int *done) {
  *done = 0;
  if (n < 1 || n > 32 || M <= 0 || N <= 0 || K <= 0 || K % BK8) return HB_OK;
  if (M >= (1ll << 31) || N >= (1ll << 31) || (int64_t)n * K >= (1ll << 31)) return HB_OK;
  const bool a_kfast = a_sk == 1, b_kfast = b_sk == 1;
  if (!a_kfast && a_sm != 1) return HB_OK;
  if (!b_kfast && b_sn != 1) return HB_OK;
  GemmP p;
  memset(&p, 0, sizeof p);
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.a_sm = a_sm; p.a_sk = a_sk; p.b_sk = b_sk; p.b_sn = b_sn;
  p.c_sm = N;
  p.accum = accum;
  p.nlist = n;
  p.a_vec = (a_kfast ? a_sm : a_sk) % 2 == 0;
  p.b_vec = (b_kfast ? b_sn : b_sk) % 2 == 0;
  for (int t = 0; t < n; t++) {
    p.Al[t] = A[t]; p.Bl[t] = B[t];
    if ((uintptr_t)A[t] & 15) p.a_vec = 0;
    if ((uintptr_t)B[t] & 15) p.b_vec = 0;
  }
  p.A = A[0]; p.B = B[0];
  const int gm = (int)((M + BM - 1) / BM), gn = (int)((N + BN - 1) / BN);
  if (gn > 65535) return HB_OK;
  const int64_t tiles = (int64_t)gm * gn, slots = 2 * (int64_t)g_hb.sm_count;
  const int chunks = (int)((int64_t)n * K / BK8);
  // K slices: the count (up to ~4 waves of two CTAs per SM, at least 8 chunks = 256 k per slice) that fills the
  // last wave best
  int splits = 1;
  {
    const int smax = (int)std::min<int64_t>(std::min<int64_t>(512, std::max(1, chunks / 8)), std::max<int64_t>(1, 4 * slots / tiles));
    double best = 0.0;
    for (int sp = 1; sp <= smax; sp++) {
      const int64_t ctas = tiles * sp, waves = (ctas + slots - 1) / slots;
      const double eff = (double)ctas / (double)(waves * slots) - 0.002 * sp;   // mild preference for fewer partials
      if (eff > best) { best = eff; splits = sp; }
    }
  }
  p.kchunks = (chunks + splits - 1) / splits;
  splits = (chunks + p.kchunks - 1) / p.kchunks;
  void *scr;
  HB_TRY(hb_scratch((size_t)splits * M * N * 8, &scr));
  p.C = (double *)scr;
  p.flags = nullptr;
  dim3 grid((unsigned)gm, (unsigned)gn, (unsigned)splits);
  size_t smem = (size_t)STAGES8 * 2 * TILE8 * 8;
#define GEMM8L_LAUNCH(AK, BKF)                                                                                 \
  do {                                                                                                         \
    HB_CUDA(cudaFuncSetAttribute(gemm8_dmma_kernel<AK, BKF>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                 (int)smem));                                                                  \
    gemm8_dmma_kernel<AK, BKF><<<grid, 256, smem, g_hb.stream>>>(p);                                           \
  } while (0)
  if (a_kfast && b_kfast) GEMM8L_LAUNCH(true, true);
  else if (a_kfast) GEMM8L_LAUNCH(true, false);
  else if (b_kfast) GEMM8L_LAUNCH(false, true);
  else GEMM8L_LAUNCH(false, false);
#undef GEMM8L_LAUNCH
  HB_LAUNCHED();
  const int64_t nel = M * N;
  gemm_reduce_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, g_hb.stream>>>(p.C, C, nel, splits, accum);
  HB_LAUNCHED();
  *done = 1;
  return HB_OK;
}

// A GROUP of up to GMAX products of the same [M, N] in ONE launch of the eight-warp kernel, each accumulated into the
// result it names:  C_r = (accum_r ? C_r : 0) + sum over the products with out == r of A_g x B_g, in product order,
// optionally followed by C_r = tanh (C_r + bias_r[m]).  The recurrent layers of MnistRnnShaped2 run ~one 64 x 64 tile
// per SM per product, so one product alone leaves SMs idle and pays a launch + pipeline fill per 0.5 GFLOP:
//   wX x + wS s (+ b, tanh)                      two products, one result                (forward layer)
//   wX^T d, wS^T d                               two products, two results               (input cotangents)
//   dW1 += d x^T, dW2 += d s^T                   two products, two accumulated results   (weight cotangents)
//   wS1^T d1' + wX2^T d2 -> dh1, wS2^T d2 -> dh2 three products, two results             (one backward time step)
// K may differ between the products; all must have the same operand orientations (one template instantiation).
// Slices of one result's tile are combined in chain order through the per-tile flags (deterministic).
// *done = 0: shapes not taken.
int hb_gemm_dmma_group(int nprod, const hb_gemm_prod *pr, int nres, const hb_gemm_res *res, int64_t M, int64_t N,
                       int *done) {
  *done = 0;
  if (nprod < 1 || nprod > GMAX || nres < 1 || nres > GMAX || M <= 0 || N <= 0) return HB_OK;
  if (M >= (1ll << 31) || N >= (1ll << 31)) return HB_OK;
  const bool a_kfast = pr[0].a_sk == 1, b_kfast = pr[0].b_sk == 1;
  for (int g = 0; g < nprod; g++) {
    if (pr[g].K <= 0 || pr[g].K >= (1ll << 31) || pr[g].out < 0 || pr[g].out >= nres) return HB_OK;
    if (a_kfast ? pr[g].a_sk != 1 : pr[g].a_sm != 1) return HB_OK;
    if (b_kfast ? pr[g].b_sk != 1 : pr[g].b_sn != 1) return HB_OK;
  }
  for (int r = 0; r < nres; r++)
    if (res[r].epi_bias && res[r].accum) return HB_OK;   // the epilogue belongs to a chain that starts from zero
  GemmP p;
  memset(&p, 0, sizeof p);
  p.M = (int)M; p.N = (int)N;
  p.c_sm = N;
  const int gm = (int)((M + BM - 1) / BM), gn = (int)((N + BN - 1) / BN);
  if (gn > 65535) return HB_OK;
  const int64_t tiles = (int64_t)gm * gn;
  if (nres * tiles > kFlagTiles) return HB_OK;
  // K slices: about two CTAs per SM over ALL results, every slice about the same depth, cut from each result's
  // concatenated chain (wX x with K = 28 rides in the slice of wS s); at least 4 chunks (128 k) per slice, at most
  // 6 slices per result
  int chunks[GMAX], total = 0;
  for (int g = 0; g < nprod; g++) total += chunks[g] = (int)((pr[g].K + BK8 - 1) / BK8);
  int budget = (int)(2 * (int64_t)g_hb.sm_count / tiles);
  if (budget < nres) budget = nres;
  int per = (total + budget - 1) / budget;
  if (per < 4) per = 4;
  static const int per3 = getenv("HB_GROUP_PER") ? atoi(getenv("HB_GROUP_PER")) : 0;     // measurement switches
  static const int per4 = getenv("HB_GROUP_PER4") ? atoi(getenv("HB_GROUP_PER4")) : 0;
  const int per_forced = nprod == 3 ? per3 : nprod == 4 ? per4 : 0;
  if (per_forced > 0) per = per_forced;
  p.nprod = nprod;
  p.nres = nres;
  int last[GMAX] = {-1, -1, -1, -1};
  for (int r = 0; r < GMAX; r++) { p.rfirst[r] = -1; p.rT[r] = 0; }
  for (int g = 0; g < nprod; g++) {
    p.gA[g] = pr[g].A; p.gB[g] = pr[g].B;
    p.ga_sm[g] = pr[g].a_sm; p.ga_sk[g] = pr[g].a_sk; p.gb_sk[g] = pr[g].b_sk; p.gb_sn[g] = pr[g].b_sn;
    p.gK[g] = (int)pr[g].K;
    p.ga_vec[g] = ((uintptr_t)pr[g].A & 15) == 0 && ((a_kfast ? pr[g].a_sm : pr[g].a_sk) % 2 == 0);
    p.gb_vec[g] = ((uintptr_t)pr[g].B & 15) == 0 && ((b_kfast ? pr[g].b_sn : pr[g].b_sk) % 2 == 0);
    p.gch[g] = chunks[g];
    p.gnext[g] = -1;
    const int r = pr[g].out;
    if (last[r] >= 0) p.gnext[last[r]] = g;
    else p.rfirst[r] = g;
    last[r] = g;
    p.rT[r] += chunks[g];
  }
  int z = 0;
  for (int r = 0; r < nres; r++) {
    if (p.rT[r] == 0) return HB_OK;                    // a result nobody writes
    int sp = (p.rT[r] + per - 1) / per;
    if (sp > 6 && per_forced <= 0) sp = 6;
    if (sp < 1) sp = 1;
    p.rkch[r] = (p.rT[r] + sp - 1) / sp;
    sp = (p.rT[r] + p.rkch[r] - 1) / p.rkch[r];
    p.rzbeg[r] = z;
    p.rzlen[r] = sp;
    z += sp;
    p.rC[r] = res[r].C;
    p.raccum[r] = res[r].accum;
    p.rbias[r] = res[r].epi_bias;
  }
  if (z > 65535) return HB_OK;
  HB_TRY(gemm_flags(&p.flags));
  dim3 grid((unsigned)gm, (unsigned)gn, (unsigned)z);
  size_t smem = (size_t)STAGES8 * 2 * TILE8 * 8;
#define GEMM8G_LAUNCH(AK, BKF)                                                                                 \
  do {                                                                                                         \
    HB_CUDA(cudaFuncSetAttribute(gemm8_group_kernel<AK, BKF>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                 (int)smem));                                                                  \
    gemm8_group_kernel<AK, BKF><<<grid, 256, smem, g_hb.stream>>>(p);                                          \
  } while (0)
  if (a_kfast && b_kfast) GEMM8G_LAUNCH(true, true);
  else if (a_kfast) GEMM8G_LAUNCH(true, false);
  else if (b_kfast) GEMM8G_LAUNCH(false, true);
  else GEMM8G_LAUNCH(false, false);
#undef GEMM8G_LAUNCH
  HB_LAUNCHED();
  *done = 1;
  return HB_OK;
}

// Two products of one [M, N] (see hb_gemm_dmma_group):
//   chain = 1:  C1 = (accum ? C1 : 0) + A1 x B1 + A2 x B2   (C2 ignored), optionally tanh (. + epi_bias[m])
//   chain = 0:  C1 = (accum ? C1 : 0) + A1 x B1,  C2 = (accum ? C2 : 0) + A2 x B2
int hb_gemm_dmma_pair(const double *A1, int64_t a1_sm, int64_t a1_sk, const double *B1, int64_t b1_sk, int64_t b1_sn,
                      int64_t K1, const double *A2, int64_t a2_sm, int64_t a2_sk, const double *B2, int64_t b2_sk,
                      int64_t b2_sn, int64_t K2, double *C1, double *C2, int64_t M, int64_t N, int chain, int accum,
                      int *done, const double *epi_bias) {
  *done = 0;
  if (epi_bias && (!chain || accum)) return HB_OK;   // the epilogue belongs to the one chain of `A1 B1 + A2 B2`
  hb_gemm_prod pr[2] = {{A1, a1_sm, a1_sk, B1, b1_sk, b1_sn, K1, 0}, {A2, a2_sm, a2_sk, B2, b2_sk, b2_sn, K2, chain ? 0 : 1}};
  hb_gemm_res res[2] = {{C1, accum, epi_bias}, {C2, accum, nullptr}};
  return hb_gemm_dmma_group(2, pr, chain ? 1 : 2, res, M, N, done);
}

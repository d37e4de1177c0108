// contract_tf32.cu -- Float (f32) contractions on the 5th-generation tensor cores: tsdot1In / tsmatmul2 /
// `ssumN . sdot1In` (OpsConcrete.hs:234-242) and, through conv.cu's strided im2col views, conv2dPreservingS with its
// two adjoints (CommonShapedOps.hs:136-157, TestConvQuickCheck.hs:290-349) for the reference's Float instantiations
// (tested there at 1e-5: TestConvQuickCheck.hs:430-464).
//
// TF32 alone (10-bit mantissa) cannot meet 1e-5, so every operand element x is split into
//   hi = rn_tf32(x),  lo = rn_tf32(x - hi)            (x - hi is exact in f32)
// and the product is accumulated as  hi_a*hi_b + hi_a*lo_b + lo_a*hi_b  in the f32 TMEM accumulator: the dropped
// term lo_a*lo_b and the rounding of lo are <= 2^-21 |a||b| per product, i.e. f32-level accuracy (the normwise bound
// the tests assert).
//
// Structure of one CTA (one 128 x BN output tile, BN in {32,64,96,128}; 288 threads):
//   warps 0-7  BUILDERS: walk the two operands in place through the same strided-view description as contract.cu
//              (per-row and per-k element offsets; overlapping im2col windows, broadcast and transposed views,
//              negative strides), split hi/lo and write both in the canonical K-major no-swizzle UMMA layout
//              (8 x 16-byte core matrices) into a 3-stage shared-memory ring, fence.proxy.async, arrive on the
//              stage's `ready` mbarrier; after the main loop the same warps are the EPILOGUE: tcgen05.ld of the
//              accumulator (lane = output row) and strided stores.
//   warp 8     allocates TMEM (128 columns), and one elected thread issues tcgen05.mma.cta_group::1.kind::tf32
//              (M = 128, N = BN, K = 8 per instruction; 3 instructions per k-step) from shared-memory descriptors,
//              tcgen05.commit -> the stage's `empty` mbarrier, finally -> `acc_full`.
// Split-K (few output tiles, long K: the dKrn shape) writes per-split partials that a second kernel adds in split
// order -> deterministic.
//
// Roofline: tensor pipe, 2*M*N*K algorithmic flops against 1/3 of the TF32 dense peak (three MMAs per product).
#include "contract.cuh"

namespace {

constexpr int BM = 128;       // UMMA M (TMEM lanes)
constexpr int BK = 32;        // k elements per stage (4 MMA k-steps of 8)
constexpr int STAGES = 3;
constexpr int NBUILD = 256;   // builder threads (8 warps)
constexpr int TILE_BYTES = BM * BK * 4;   // one operand tile (hi or lo), BN <= 128 uses the same footprint
constexpr int STAGE_BYTES = 4 * TILE_BYTES;   // A_hi, A_lo, B_hi, B_lo
constexpr int SBO = 1024;     // bytes between 8-row groups
constexpr int LBO = 128;      // bytes between core matrices adjacent in K
constexpr int TMEM_COLS = 128;

struct tf32_params {
  hb_cgroup M, N, B, K;       // roles as the kernel sees them: "a" = the operand on the TMEM-lane side
  int64_t Mtot, Ntot, Ktot;
  int bn;                     // N tile (32, 64, 96, 128)
  int a_kfast, b_kfast;       // builder thread mapping per operand (coalescing)
  int splits;                 // split-K factor; > 1: compact partial output
  int kb_per_split;           // k blocks (of BK) per split
  float *partial;             // [batch][split][n][m] when splits > 1
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}

// shared-memory matrix descriptor: K-major, no swizzle, version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) |
         (1ull << 46);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// row-major linear index over g.shape -> 32-bit element offsets in a and b
__device__ __forceinline__ void cg_off32(const hb_cgroup &g, uint32_t lin, int &oa, int &ob) {
  oa = ob = 0;
#pragma unroll 1
  for (int d = g.rank - 1; d >= 0; d--) {
    uint32_t s = (uint32_t)g.shape[d];
    uint32_t q = lin / s, r = lin - q * s;
    oa += (int)r * (int)g.sa[d];
    ob += (int)r * (int)g.sb[d];
    lin = q;
  }
}

__global__ void __launch_bounds__(NBUILD + 32, 1)
    hb_contract_tf32_kernel(const __grid_constant__ tf32_params P, const float *__restrict__ A,
                            const float *__restrict__ Bm, float *__restrict__ C) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // [STAGES][A_hi | A_lo | B_hi | B_lo] then tables and barriers
  int *offAm = (int *)(smem + STAGES * STAGE_BYTES);        // [BM]
  int *offBn = offAm + BM;                                  // [128]
  int64_t *offCm = (int64_t *)(offBn + 128);                // [BM]
  int64_t *offCn = offCm + BM;                              // [128]
  uint64_t *bars = (uint64_t *)(offCn + 128);               // ready[STAGES], empty[STAGES], acc_full
  uint32_t *tmem_slot = (uint32_t *)(bars + 2 * STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int64_t n0 = (int64_t)blockIdx.y * P.bn;
  const int split = (int)(blockIdx.z % P.splits);
  const int64_t batch = blockIdx.z / P.splits;
  const int BN = P.bn;

  int64_t ba, bb, bc;
  hb_cg_offsets(P.B, batch, ba, bb, bc);
  A += ba;
  Bm += bb;

  const int nkb_total = (int)((P.Ktot + BK - 1) / BK);
  const int kb0 = split * P.kb_per_split;
  const int kb1 = min(nkb_total, kb0 + P.kb_per_split);
  const int nkb = max(0, kb1 - kb0);

  const uint32_t bar0 = smem_u32(bars);
  auto ready_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  const uint32_t acc_bar = bar0 + 8u * (2 * STAGES);

  // ---- one-time set-up: offset tables, barriers, TMEM
  if (tid < BM) {
    int64_t m = m0 + tid, oa = 0, ob = 0, oc = 0;
    if (m < P.Mtot) hb_cg_offsets(P.M, m, oa, ob, oc);
    offAm[tid] = (int)oa;
    offCm[tid] = oc;
  } else if (tid < BM + 128) {
    int t = tid - BM;
    int64_t n = n0 + t, oa = 0, ob = 0, oc = 0;
    if (t < BN && n < P.Ntot) hb_cg_offsets(P.N, n, oa, ob, oc);
    offBn[t] = (int)ob;
    offCn[t] = oc;
  }
  if (warp == 8) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; s++) {
        mbar_init(ready_bar(s), NBUILD);
        mbar_init(empty_bar(s), 1);
      }
      mbar_init(acc_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *(volatile uint32_t *)tmem_slot;

  if (warp == 8) {
    // ================================================================ MMA issuer
    if (lane == 0 && nkb > 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t sbase = smem_u32(smem);
      for (int i = 0; i < nkb; i++) {
        const int s = i % STAGES;
        mbar_wait(ready_bar(s), (uint32_t)((i / STAGES) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = sbase + (uint32_t)s * STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 8; k++) {
          const uint64_t a_hi = umma_desc(st + k * 2 * LBO);
          const uint64_t a_lo = umma_desc(st + TILE_BYTES + k * 2 * LBO);
          const uint64_t b_hi = umma_desc(st + 2 * TILE_BYTES + k * 2 * LBO);
          const uint64_t b_lo = umma_desc(st + 3 * TILE_BYTES + k * 2 * LBO);
          // small terms first
          umma_tf32(tmem, a_lo, b_hi, idesc, (i | k) != 0);
          umma_tf32(tmem, a_hi, b_lo, idesc, 1u);
          umma_tf32(tmem, a_hi, b_hi, idesc, 1u);
        }
        umma_commit(empty_bar(s));   // implies tcgen05.fence::before_thread_sync
      }
      umma_commit(acc_bar);
    }
  } else {
    // ================================================================ builders
    // thread -> (k chunk of 4, base row in 0..31) per operand; rows r0 + 32 j
    const int a_kc = P.a_kfast ? (lane >> 3) + 4 * (warp & 1) : warp;
    const int a_r0 = P.a_kfast ? (warp >> 1) * 8 + (lane & 7) : lane;
    const int b_kc = P.b_kfast ? (lane >> 3) + 4 * (warp & 1) : warp;
    const int b_r0 = P.b_kfast ? (warp >> 1) * 8 + (lane & 7) : lane;
    // which k of the stage this LANE decomposes for its warp, and where a thread fetches its four from
    const int a_kbase = P.a_kfast ? 16 * (warp & 1) : 4 * warp, a_nk = P.a_kfast ? 16 : 4;
    const int b_kbase = P.b_kfast ? 16 * (warp & 1) : 4 * warp, b_nk = P.b_kfast ? 16 : 4;
    const int a_src = P.a_kfast ? (lane >> 3) * 4 : 0, b_src = P.b_kfast ? (lane >> 3) * 4 : 0;
    const int nrb = BN >> 5;   // row groups of the B tile

    int am_off[4], bn_off[4];
    bool am_ok[4], bn_ok[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      am_off[j] = offAm[a_r0 + 32 * j];
      am_ok[j] = m0 + a_r0 + 32 * j < P.Mtot;
      bn_off[j] = offBn[b_r0 + 32 * j];
      bn_ok[j] = j < nrb && n0 + b_r0 + 32 * j < P.Ntot;
    }
    // byte offset of (row, k chunk) inside a tile
    uint32_t a_st[4], b_st[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int ra = a_r0 + 32 * j, rb = b_r0 + 32 * j;
      a_st[j] = (uint32_t)((ra >> 3) * SBO + a_kc * LBO + (ra & 7) * 16);
      b_st[j] = (uint32_t)((rb >> 3) * SBO + b_kc * LBO + (rb & 7) * 16);
    }

    float ra[4][4], rb[4][4];
    auto load_stage = [&](int kb) {
      // element offsets of this stage's k values: one decomposition per lane, exchanged by shuffles
      int ka_off = 0, kb_off = 0, dummy;
      bool ka_ok = false, kbv_ok = false;
      {
        int64_t k = (int64_t)kb * BK + a_kbase + lane;
        if (lane < a_nk && k < P.Ktot) { cg_off32(P.K, (uint32_t)k, ka_off, dummy); ka_ok = true; }
        k = (int64_t)kb * BK + b_kbase + lane;
        if (lane < b_nk && k < P.Ktot) { cg_off32(P.K, (uint32_t)k, dummy, kb_off); kbv_ok = true; }
      }
      const unsigned ma = __ballot_sync(0xffffffffu, ka_ok), mb = __ballot_sync(0xffffffffu, kbv_ok);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int oa = __shfl_sync(0xffffffffu, ka_off, a_src + i);
        const int ob = __shfl_sync(0xffffffffu, kb_off, b_src + i);
        const bool va = (ma >> (a_src + i)) & 1u, vb = (mb >> (b_src + i)) & 1u;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          ra[j][i] = (va && am_ok[j]) ? __ldg(A + am_off[j] + oa) : 0.f;
          rb[j][i] = (vb && bn_ok[j]) ? __ldg(Bm + bn_off[j] + ob) : 0.f;
        }
      }
    };
    auto split_store = [&](uint8_t *tile_hi, uint32_t off, const float (&v)[4]) {
      float4 hi, lo;
      float h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        h[i] = tf32_rn(v[i]);
        float d = v[i] - h[i];
        l[i] = (fabsf(h[i]) <= 3.0e38f) ? tf32_rn(d) : 0.f;   // Inf/NaN stay in the hi term only
      }
      hi = make_float4(h[0], h[1], h[2], h[3]);
      lo = make_float4(l[0], l[1], l[2], l[3]);
      *(float4 *)(tile_hi + off) = hi;
      *(float4 *)(tile_hi + TILE_BYTES + off) = lo;
    };

    if (nkb > 0) load_stage(kb0);
    for (int i = 0; i < nkb; i++) {
      const int s = i % STAGES;
      mbar_wait(empty_bar(s), (uint32_t)(((i / STAGES) & 1) ^ 1));
      uint8_t *st = smem + (size_t)s * STAGE_BYTES;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        split_store(st, a_st[j], ra[j]);
        if (j < nrb) split_store(st + 2 * TILE_BYTES, b_st[j], rb[j]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(ready_bar(s));
      if (i + 1 < nkb) load_stage(kb0 + i + 1);
    }

    // ================================================================ epilogue
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int row = 32 * q + lane;
    const bool row_ok = m0 + row < P.Mtot;
    const int64_t crow = offCm[row];
    if (nkb > 0) {
      mbar_wait(acc_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    for (int g = warp >> 2; g < (BN >> 5); g += 2) {
      uint32_t v[32];
      if (nkb > 0) {
        const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(g * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int c = 0; c < 32; c++) v[c] = 0u;
      }
      if (P.splits > 1) {
        // compact partial: [batch][split][n][m] (m fastest: coalesced across the warp)
        float *pp = P.partial + ((size_t)blockIdx.z * P.Ntot) * P.Mtot;
#pragma unroll
        for (int c = 0; c < 32; c++) {
          const int64_t n = n0 + g * 32 + c;
          if (row_ok && n < P.Ntot) pp[n * P.Mtot + (m0 + row)] = __uint_as_float(v[c]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; c++) {
          const int64_t n = n0 + g * 32 + c;
          if (row_ok && n < P.Ntot) C[bc + crow + offCn[g * 32 + c]] = __uint_as_float(v[c]);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

// sums the split-K partials in split order (deterministic) into the strided result
__global__ void hb_tf32_splitk_reduce_kernel(const __grid_constant__ tf32_params P, float *__restrict__ C) {
  const int64_t MN = P.Mtot * P.Ntot;
  const int64_t total = MN * P.B.n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t batch = i / MN, r = i - batch * MN;
    const int64_t n = r / P.Mtot, m = r - n * P.Mtot;
    const float *pp = P.partial + (batch * P.splits) * MN + r;
    float acc = pp[0];
    for (int s = 1; s < P.splits; s++) acc = acc + pp[(int64_t)s * MN];
    int64_t oa, ob, ocb, ocm, ocn;
    hb_cg_offsets(P.B, batch, oa, ob, ocb);
    hb_cg_offsets(P.M, m, oa, ob, ocm);
    hb_cg_offsets(P.N, n, oa, ob, ocn);
    C[ocb + ocm + ocn] = acc;
  }
}

int64_t min_abs(const hb_cgroup &g, const int64_t *s) {
  int64_t m = INT64_MAX;
  for (int i = 0; i < g.rank; i++) {
    int64_t x = s[i] < 0 ? -s[i] : s[i];
    if (x && x < m) m = x;
  }
  return m;
}
// largest |element offset| an operand can reach through the four groups
int64_t span_of(const hb_contractp &P, int which) {
  int64_t lo = 0, hi = 0;
  const hb_cgroup *gs[4] = {&P.M, &P.N, &P.B, &P.K};
  for (const hb_cgroup *g : gs)
    for (int i = 0; i < g->rank; i++) {
      int64_t s = which == 0 ? g->sa[i] : g->sb[i];
      int64_t e = s * (g->shape[i] - 1);
      if (e > 0) hi += e; else lo += e;
    }
  return hi > -lo ? hi : -lo;
}
void swap_ab(hb_cgroup &g) {
  for (int i = 0; i < g.rank; i++) std::swap(g.sa[i], g.sb[i]);
}

}  // namespace

static int g_tf32_mode = -1;   // HB_TF32: 0 = off (FFMA engine), 1 = on (default)

int hb_contract_tf32(const hb_contractp &Pin, const float *A, const float *B, float *C, int *done) {
  *done = 0;
  if (g_tf32_mode < 0) {
    const char *e = getenv("HB_TF32");
    g_tf32_mode = (e && e[0] == '0') ? 0 : 1;
  }
  if (!g_tf32_mode || g_hb.trace_only) return HB_OK;
  // groups arrive collapsed (cg_finish): sizes and spans
  const int64_t szM = Pin.M.n, szN = Pin.N.n, K = Pin.K.n, nb = Pin.B.n;
  if (szM < 1 || szN < 1 || K < 8 || nb < 1) return HB_OK;
  // worth a tensor-core tile: enough rows on one side and enough work overall
  if ((szM < 32 && szN < 32) || szM * szN * K < (int64_t)1 << 18) return HB_OK;
  if (span_of(Pin, 0) >= INT32_MAX || span_of(Pin, 1) >= INT32_MAX || K >= INT32_MAX || szM >= INT32_MAX ||
      szN >= INT32_MAX)
    return HB_OK;

  tf32_params P;
  memset(&P, 0, sizeof P);
  P.M = Pin.M; P.N = Pin.N; P.B = Pin.B; P.K = Pin.K;
  // which side rides on the 128 TMEM lanes: the larger one; on a tie-ish (both >= 128) the side the result is
  // contiguous along (coalesced epilogue stores)
  bool swap = false;
  if (szM >= 128 && szN >= 128) swap = min_abs(P.N, P.N.sc) < min_abs(P.M, P.M.sc);
  else swap = szN > szM;
  if (swap) {
    std::swap(P.M, P.N);
    swap_ab(P.M); swap_ab(P.N); swap_ab(P.B); swap_ab(P.K);
    std::swap(A, B);
  }
  P.Mtot = P.M.n; P.Ntot = P.N.n; P.Ktot = K;
  P.bn = (int)std::min<int64_t>(128, ((P.Ntot + 31) / 32) * 32);
  // builder mapping: lanes along k when k is the faster axis of the operand and long enough to fill a 64-byte run
  auto kfast = [&](const int64_t *ks, const hb_cgroup &g, const int64_t *ms) {
    int64_t sk = min_abs(P.K, ks), sm = min_abs(g, ms);
    int64_t inner = P.K.rank ? P.K.shape[P.K.rank - 1] : 1;
    if (sk < sm) return inner >= 8 || sm > 16;
    if (sk == sm) return inner >= 32;
    return false;
  };
  P.a_kfast = kfast(P.K.sa, P.M, P.M.sa);
  P.b_kfast = kfast(P.K.sb, P.N, P.N.sb);

  const int64_t tm = (P.Mtot + BM - 1) / BM, tn = (P.Ntot + P.bn - 1) / P.bn;
  const int nkb = (int)((K + BK - 1) / BK);
  int splits = 1;
  const int64_t ctas = tm * tn * nb;
  if (ctas < g_hb.sm_count && nkb >= 16) {
    splits = (int)std::min<int64_t>((g_hb.sm_count + ctas - 1) / ctas, nkb / 8);
    if (splits < 1) splits = 1;
  }
  P.kb_per_split = (nkb + splits - 1) / splits;
  splits = (nkb + P.kb_per_split - 1) / P.kb_per_split;
  P.splits = splits;
  if (tn > 65535 || nb * splits > 65535) return HB_OK;
  if (splits > 1) {
    void *scratch;
    HB_TRY(hb_scratch((size_t)nb * splits * P.Mtot * P.Ntot * sizeof(float), &scratch));
    P.partial = (float *)scratch;
  }
  const size_t smem_bytes = (size_t)STAGES * STAGE_BYTES + BM * 4 + 128 * 4 + BM * 8 + 128 * 8 + (2 * STAGES + 1) * 8 + 16;
  static bool attr_set = false;
  if (!attr_set) {
    HB_CUDA(cudaFuncSetAttribute(hb_contract_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    attr_set = true;
  }
  dim3 grid((unsigned)tm, (unsigned)tn, (unsigned)(nb * splits));
  hb_contract_tf32_kernel<<<grid, NBUILD + 32, smem_bytes, g_hb.stream>>>(P, A, B, C);
  HB_LAUNCHED();
  if (splits > 1) {
    int64_t total = P.Mtot * P.Ntot * nb;
    hb_tf32_splitk_reduce_kernel<<<hb_blocks_for(total, 256, g_hb.sm_count * 8), 256, 0, g_hb.stream>>>(P, C);
    HB_LAUNCHED();
  }
  *done = 1;
  return HB_OK;
}

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
// Variants measured on 4096^3 (row-major operands / A^T view), TFLOP/s:  8 builder warps with the proxy fence in the
// builders (this version) 119 / 144;  16 builder warps 117 / 122;  16 warps with the fence moved to the MMA thread
// 124 / 127;  8 warps with the fence in the MMA thread 112 / 122;  three k blocks in flight 68 (spills).  More warps or
// deeper prefetch do not help: with everything but the global loads switched off the loop still takes 0.79-0.92 ms of
// the 0.96-1.16 -- every 128 x 128 tile re-reads its operands from L2 (32 KB per MFLOP, 4.3 GB in all), which caps
// this one-CTA-per-tile design near 165 TFLOP/s; the MMAs alone reach 265 TFLOP/s = 0.68 of the TF32 rate (M = 128
// SS-mode operand reads).  Past that: cta_group::2 / cluster multicast (operands shared by a CTA pair), N = 256 tiles.
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
  // per operand (0 = a, 1 = b): shared-memory layout and how a builder thread fetches its 4-element units
  int mn[2];                  // fetch pattern: 0 = k-fast (4 consecutive k of one row), 1 = m-fast (4 rows of one k)
  int contig4[2];             // the 4 elements of a fetch are at consecutive addresses
  int vec[2];                 // ... and 16-byte aligned: one 128-bit load
  int kdelta[4];              // BK written in the mixed radix of the K group (rank <= 4): incremental k decomposition
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

// shared-memory matrix descriptor: K-major, no swizzle, version 1 (sm_100).  Tile = [row/8][k/4][row%8][16 B]:
// core matrices adjacent in K are LBO = 128 B apart, 8-row groups SBO = 1 KB.  (An MN-major no-swizzle tile for the
// m-contiguous operands did not give correct products on the B200 for kind::tf32 with either reading of LBO/SBO --
// measured, round 2 -- so m-contiguous views are transposed in registers into this layout instead.)
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

template <int AMODE, int BMODE>
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
        mbar_init(ready_bar(s), NBUILD / 32);
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
    // Two fetch patterns, one shared-memory layout (K-major core matrices).  Per stage and operand a thread owns
    // 16 elements r[j][i] = (its row j, k = its 4-k chunk + i):
    //   k-fast: lanes = 8 rows x 4 chunks per warp; rows r0 + 32 j (one 128-bit load each when the view is
    //           k-contiguous and aligned);
    //   m-fast: lanes = 32 consecutive 4-row groups, the warp index is the chunk; rows 4 lane .. 4 lane + 3 (one
    //           128-bit load along m per k when m-contiguous and aligned: 512 contiguous bytes per warp), i.e. a
    //           4 x 4 block transposed in registers.  Its four 16-byte stores go to rows 4 lane + ((j + rot) & 3),
    //           rot = (lane >> 1) & 3, so that the eight lanes of a quarter warp hit eight different 16-byte columns
    //           of the 128-byte core matrices (no bank conflicts).
    struct op_thread {
      const float *rp[4];   // address of (row j, k offset 0); m-fast contiguous: only rp[0] (rows are rp[0] + j)
      unsigned rmask;       // bit j: row j exists
      int kq;               // first k (inside the stage) of this thread's 4-k chunk
      uint32_t st[4];       // byte offset inside the tile of the chunk written by store step j
      int srow[4];          // ... and its row (B tile: rows beyond BN are not stored)
      int rot;
    };
    auto make_op = [&](int mode, const float *base, const int *offtab, int rows) {
      op_thread t;
      t.rmask = 0;
      if (!(mode & 1)) {
        const int kc = (lane >> 3) + 4 * (warp & 1), r0 = (warp >> 1) * 8 + (lane & 7);
        t.kq = kc * 4;
        t.rot = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int r = r0 + 32 * j;
          t.srow[j] = r;
          t.st[j] = (uint32_t)((r >> 3) * SBO + kc * LBO + (r & 7) * 16);
          t.rp[j] = base + (r < rows ? offtab[r] : 0);
          if (r < rows) t.rmask |= 1u << j;
        }
      } else {
        t.kq = 4 * warp;
        t.rot = (lane >> 1) & 3;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int r = 4 * lane + ((j + t.rot) & 3);
          t.srow[j] = r;
          t.st[j] = (uint32_t)((r >> 3) * SBO + warp * LBO + (r & 7) * 16);
          const int rr = 4 * lane + j;
          t.rp[j] = base + (rr < rows ? offtab[rr] : 0);
          if (rr < rows) t.rmask |= 1u << j;
        }
      }
      return t;
    };
    const int rows_a = (int)min((int64_t)BM, P.Mtot - m0), rows_b = (int)min((int64_t)BN, P.Ntot - n0);
    const op_thread ta = make_op(AMODE, A, offAm, rows_a), tb = make_op(BMODE, Bm, offBn, rows_b);

    // element offset of reduced index k in operand `which`
    auto koff = [&](int64_t k, int which) {
      int oa, ob;
      if (P.K.rank == 1) return (int)k * (int)(which ? P.K.sb[0] : P.K.sa[0]);
      cg_off32(P.K, (uint32_t)k, oa, ob);
      return which ? ob : oa;
    };
    // Incremental decomposition: a thread's k advances by exactly BK from one call of load_op to the next (per
    // operand), so the multi-index over the K group is carried along with digit-wise adds (no division per stage).
    // Tracked k: the first k of the thread's chunk (k-fast) / the chunk's k of lane & 3 (m-fast, lanes 0..3 are read).
    struct kwalk { int idx[4]; };
    const bool kinc = P.K.rank >= 2 && P.K.rank <= 4;
    auto kwalk_init = [&](int kq, int mn) {
      kwalk w;
      uint32_t lin = (uint32_t)min((int64_t)kb0 * BK + kq + (mn ? (lane & 3) : 0), (int64_t)INT32_MAX);
#pragma unroll
      for (int d = 3; d >= 0; d--) {
        w.idx[d] = 0;
        if (d < P.K.rank) {
          if (d == 0) { w.idx[0] = (int)lin; }
          else { uint32_t s = (uint32_t)P.K.shape[d], q = lin / s; w.idx[d] = (int)(lin - q * s); lin = q; }
        }
      }
      return w;
    };
    auto kwalk_off = [&](const kwalk &w, int which) {
      int off = 0;
#pragma unroll
      for (int d = 0; d < 4; d++)
        if (d < P.K.rank) off += w.idx[d] * (int)(which ? P.K.sb[d] : P.K.sa[d]);
      return off;
    };
    auto kwalk_next = [&](kwalk &w) {
      int carry = 0;
#pragma unroll
      for (int d = 3; d >= 0; d--)
        if (d < P.K.rank) {
          int v = w.idx[d] + P.kdelta[d] + carry;
          carry = 0;
          if (d > 0 && v >= (int)P.K.shape[d]) { v -= (int)P.K.shape[d]; carry = 1; }
          w.idx[d] = v;
        }
    };
    kwalk wa = kwalk_init(ta.kq, AMODE & 1), wb = kwalk_init(tb.kq, BMODE & 1);
    // fetch this thread's 16 elements of one operand for k block kb
    auto load_op = [&](const op_thread &t, auto which_c, int kb, float (&r)[4][4]) {
      constexpr int which = decltype(which_c)::value;
      kwalk &kw = which ? wb : wa;
      constexpr int mode = which ? BMODE : AMODE;
      constexpr bool mn = mode & 1, contig = mode & 2, vec = mode & 4;
      const int64_t k0 = (int64_t)kb * BK + t.kq;
      const bool fullk = (int64_t)(kb + 1) * BK <= P.Ktot;
      int ko[4];
      unsigned kmask = 0;
      if constexpr (!mn) {
        if constexpr (contig) {   // Ktot % 4 == 0 here: the chunk is inside or outside as a whole
          const bool in = fullk || k0 < P.Ktot;
          int o = 0;
          if (kinc) { o = kwalk_off(kw, which); kwalk_next(kw); }
          else if (in) o = koff(k0, which);
#pragma unroll
          for (int i = 0; i < 4; i++) ko[i] = o + i;
          kmask = in ? 15u : 0u;
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const bool in = fullk || k0 + i < P.Ktot;
            ko[i] = in ? koff(k0 + i, which) : 0;
            kmask |= (unsigned)in << i;
          }
        }
        if constexpr (vec) {
#pragma unroll
          for (int j = 0; j < 4; j++) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (((t.rmask >> j) & 1u) && kmask) v = __ldg((const float4 *)(t.rp[j] + ko[0]));
            r[j][0] = v.x; r[j][1] = v.y; r[j][2] = v.z; r[j][3] = v.w;
          }
        } else if (t.rmask == 15u && kmask == 15u) {
#pragma unroll
          for (int j = 0; j < 4; j++) {
            if constexpr (contig) {   // one address per row, the four k at immediate offsets
              const float *q = t.rp[j] + ko[0];
#pragma unroll
              for (int i = 0; i < 4; i++) r[j][i] = __ldg(q + i);
            } else {
#pragma unroll
              for (int i = 0; i < 4; i++) r[j][i] = __ldg(t.rp[j] + ko[i]);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++)
              r[j][i] = (((t.rmask >> j) & 1u) && ((kmask >> i) & 1u)) ? __ldg(t.rp[j] + ko[i]) : 0.f;
        }
      } else {
        // the chunk is warp-uniform: lanes 0..3 decompose one k each
        if (P.K.rank == 1) {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const bool in = fullk || k0 + i < P.Ktot;
            ko[i] = in ? koff(k0 + i, which) : 0;
            kmask |= (unsigned)in << i;
          }
        } else {
          int mine = 0;
          const bool ok = lane < 4 && (fullk || k0 + lane < P.Ktot);
          if (kinc) { mine = kwalk_off(kw, which); kwalk_next(kw); }
          else if (ok) mine = koff(k0 + lane, which);
          kmask = __ballot_sync(0xffffffffu, ok) & 15u;
#pragma unroll
          for (int i = 0; i < 4; i++) ko[i] = __shfl_sync(0xffffffffu, mine, i);
        }
        if constexpr (contig && vec) {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((t.rmask & 1u) && ((kmask >> i) & 1u)) v = __ldg((const float4 *)(t.rp[0] + ko[i]));
            r[0][i] = v.x; r[1][i] = v.y; r[2][i] = v.z; r[3][i] = v.w;
          }
        } else if (t.rmask == 15u && kmask == 15u) {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            if constexpr (contig) {   // one address per k, the four rows at immediate offsets
              const float *q = t.rp[0] + ko[i];
#pragma unroll
              for (int j = 0; j < 4; j++) r[j][i] = __ldg(q + j);
            } else {
#pragma unroll
              for (int j = 0; j < 4; j++) r[j][i] = __ldg(t.rp[j] + ko[i]);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
              r[j][i] = (((t.rmask >> j) & 1u) && ((kmask >> i) & 1u)) ? __ldg(t.rp[j] + ko[i]) : 0.f;
        }
      }
    };
    // x = hi + lo with hi = x rounded to TF32 (bit trick: add half an ulp of TF32, clear the 13 low bits) and
    // lo = x - hi exactly; the tensor core reads the top 19 bits of lo.  Non-finite x: hi = x, lo = NaN.
    auto split_store = [&](uint8_t *tile_hi, uint32_t off, const float (&v)[4]) {
      float h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        h[i] = __uint_as_float((__float_as_uint(v[i]) + 0x1000u) & 0xffffe000u);
        l[i] = v[i] - h[i];
      }
      *(float4 *)(tile_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
      *(float4 *)(tile_hi + TILE_BYTES + off) = make_float4(l[0], l[1], l[2], l[3]);
    };
    auto store_op = [&](uint8_t *tile_hi, const op_thread &t, auto which_c, const float (&r)[4][4], int nrows) {
      constexpr int mode = decltype(which_c)::value ? BMODE : AMODE;
      if constexpr (!(mode & 1)) {
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (t.srow[j] < nrows) split_store(tile_hi, t.st[j], r[j]);
      } else {
      // store step j writes row (j + rot) & 3 of the thread's block: rotate the rows by rot with two select levels
      float s1[4][4], s2[4][4];
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) s1[j][i] = (t.rot & 1) ? r[(j + 1) & 3][i] : r[j][i];
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) s2[j][i] = (t.rot & 2) ? s1[(j + 2) & 3][i] : s1[j][i];
      if (4 * lane < nrows) {
#pragma unroll
        for (int j = 0; j < 4; j++) split_store(tile_hi, t.st[j], s2[j]);
      }
      }
    };
    constexpr std::integral_constant<int, 0> opA{};
    constexpr std::integral_constant<int, 1> opB{};

    auto build = [&](int i, const float (&ra)[4][4], const float (&rb)[4][4]) {
      const int s = i % STAGES;
      mbar_wait(empty_bar(s), (uint32_t)(((i / STAGES) & 1) ^ 1));
      uint8_t *st = smem + (size_t)s * STAGE_BYTES;
      store_op(st, ta, opA, ra, BM);
      store_op(st + 2 * TILE_BYTES, tb, opB, rb, BN);   // rows of the B tile beyond BN are never read by the MMA
      // generic-proxy writes -> visible to the tensor core's async-proxy reads, then one arrival per warp
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(ready_bar(s));
    };

    // two k blocks of loads in flight per thread: registers of block i + 2 are requested before block i + 1 is split
    // (three in flight measured slower: 68 vs 75 TFLOP/s on 4096^3 -- the loop is issue-bound, not latency-bound)
    float ra0[4][4], rb0[4][4], ra1[4][4], rb1[4][4];
    if (nkb > 0) { load_op(ta, opA, kb0, ra0); load_op(tb, opB, kb0, rb0); }
    if (nkb > 1) { load_op(ta, opA, kb0 + 1, ra1); load_op(tb, opB, kb0 + 1, rb1); }
    for (int i = 0; i < nkb; i += 2) {
      build(i, ra0, rb0);
      if (i + 2 < nkb) { load_op(ta, opA, kb0 + i + 2, ra0); load_op(tb, opB, kb0 + i + 2, rb0); }
      if (i + 1 < nkb) {
        build(i + 1, ra1, rb1);
        if (i + 3 < nkb) { load_op(ta, opA, kb0 + i + 3, ra1); load_op(tb, opB, kb0 + i + 3, rb1); }
      }
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
  // per operand: tile layout (K-major / MN-major) and fetch width, from the strides of the view
  auto configure = [&](int which, const hb_cgroup &G, const float *base) {
    auto str = [&](const hb_cgroup &g, int i) { return which ? g.sb[i] : g.sa[i]; };
    auto last_contig = [&](const hb_cgroup &g) {
      return g.rank > 0 && str(g, g.rank - 1) == 1 && g.shape[g.rank - 1] % 4 == 0;
    };
    // every stride other than the unit one a multiple of 4 elements, base 16-byte aligned
    auto aligned = [&](const hb_cgroup &fast) {
      if ((uintptr_t)base % 16) return false;
      const hb_cgroup *gs[3] = {&G, &P.K, &P.B};
      for (const hb_cgroup *g : gs)
        for (int i = 0; i < g->rank; i++) {
          if (g == &fast && i == g->rank - 1) continue;
          if (str(*g, i) % 4) return false;
        }
      return true;
    };
    bool kc = last_contig(P.K), mc = last_contig(G);
    static const char *force = getenv("HB_TF32_LAYOUT");   // measurement / debugging: "k" or "mn" tiles only
    int mn, contig, vec = 0;
    if (force && force[0] == 'k') mc = false;
    if (force && force[0] == 'm') kc = false;
    if (kc && (!mc || aligned(P.K) || !aligned(G))) { mn = 0; contig = 1; vec = aligned(P.K); }
    else if (mc) { mn = 1; contig = 1; vec = aligned(G); }
    else {
      // no 4-element run either way: lanes along the axis with the smaller stride (coalescing), scalar fetches
      auto min_stride = [&](const hb_cgroup &g) {
        int64_t m = INT64_MAX;
        for (int i = 0; i < g.rank; i++) {
          int64_t x = str(g, i) < 0 ? -str(g, i) : str(g, i);
          if (x && x < m) m = x;
        }
        return m;
      };
      mn = min_stride(G) < min_stride(P.K);
      if (force) mn = force[0] == 'm';
      contig = 0;
    }
    P.mn[which] = mn; P.contig4[which] = contig; P.vec[which] = vec;
  };
  configure(0, P.M, A);
  configure(1, P.N, B);
  if (P.K.rank >= 2 && P.K.rank <= 4) {
    int64_t rem = BK;
    for (int d = P.K.rank - 1; d >= 1; d--) { P.kdelta[d] = (int)(rem % P.K.shape[d]); rem /= P.K.shape[d]; }
    P.kdelta[0] = (int)rem;
  }

  const int64_t tm = (P.Mtot + BM - 1) / BM, tn = (P.Ntot + P.bn - 1) / P.bn;
  const int nkb = (int)((K + BK - 1) / BK);
  int splits = 1;
  const int64_t ctas = tm * tn * nb;
  if (ctas < g_hb.sm_count && nkb >= 16) {
    splits = (int)std::min<int64_t>(g_hb.sm_count / ctas, nkb / 8);   // one wave: ctas * splits <= SMs (1 CTA per SM)
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
  // one instantiation per pair of fetch modes (bit 0 = m-fast, bit 1 = 4-element runs, bit 2 = 128-bit loads);
  // vec implies contig: modes 0, 1, 2, 3, 6, 7
  typedef void (*kern_t)(const tf32_params, const float *, const float *, float *);
  static kern_t table[8][8];
  static bool table_set = false;
  if (!table_set) {
#define HB_TF_ROW(a)                                                                                           \
  table[a][0] = hb_contract_tf32_kernel<a, 0>; table[a][1] = hb_contract_tf32_kernel<a, 1>;                    \
  table[a][2] = hb_contract_tf32_kernel<a, 2>; table[a][3] = hb_contract_tf32_kernel<a, 3>;                    \
  table[a][6] = hb_contract_tf32_kernel<a, 6>; table[a][7] = hb_contract_tf32_kernel<a, 7>;
    HB_TF_ROW(0) HB_TF_ROW(1) HB_TF_ROW(2) HB_TF_ROW(3) HB_TF_ROW(6) HB_TF_ROW(7)
#undef HB_TF_ROW
    for (int a = 0; a < 8; a++)
      for (int b = 0; b < 8; b++)
        if (table[a][b])
          HB_CUDA(cudaFuncSetAttribute(table[a][b], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    table_set = true;
  }
  const int am = P.mn[0] | (P.contig4[0] << 1) | (P.vec[0] << 2), bm = P.mn[1] | (P.contig4[1] << 1) | (P.vec[1] << 2);
  kern_t kern = table[am][bm];
  HB_REQUIRE(kern, "tf32 contraction: no kernel for fetch modes %d, %d", am, bm);
  dim3 grid((unsigned)tm, (unsigned)tn, (unsigned)(nb * splits));
  kern<<<grid, NBUILD + 32, smem_bytes, g_hb.stream>>>(P, A, B, C);
  HB_LAUNCHED();
  if (splits > 1) {
    int64_t total = P.Mtot * P.Ntot * nb;
    hb_tf32_splitk_reduce_kernel<<<hb_blocks_for(total, 256, g_hb.sm_count * 8), 256, 0, g_hb.stream>>>(P, C);
    HB_LAUNCHED();
  }
  *done = 1;
  return HB_OK;
}

// ixprog.cu -- index-function programs and the kernels driven by them:
// tsgather (OpsConcrete.hs:321-323,1621-1665), tsscatter (:318-320,
// 1529-1619), tsoneHot (:1516-1525) and fused elementwise maps.
//
// Kernel selection ("kernel generator for index functions"):
//   * affine program, provably in bounds  -> the gather IS a strided view of
//     the source (overlapping windows allowed): zero bytes moved, no launch.
//   * affine program with out-of-bounds    -> masked strided copy, the index
//     arithmetic folded into per-axis strides + per-component bound checks.
//   * general program (table lookups, kargMax, ifH on tensor data ...)
//     -> the VM evaluates the program once per shm position into an offset
//     table, then a pure bandwidth copy moves the shn slices; positions of
//     rank-0 slices are fused into one pass.
// Scatter is deterministic and reproduces the reference's accumulation order
// (row-major over the source positions, OpsConcrete.hs:1594-1601): counting
// pass -> exclusive scan -> fill -> per-destination sort by source position
// -> sequential sum; when no destination is hit twice the sort/sum stages
// exit immediately and a direct store kernel does the work.
// HBM roofline: 2*sizeof(T) bytes per moved element (+ index tables).
#include "iter.cuh"
#include "lazy.cuh"
#include "vm.cuh"
#include "jit.cuh"

#include <mutex>
#include <vector>

// ------------------------------------------------------------------ build
extern "C" int hb_ixprog_build(const hb_insn *code, int n_insns, int n_vars, int n_out,
                               hb_array *const *tensors, int n_tensors, int f32, hb_ixprog **out) {
  HB_REQUIRE(code && out, "null argument");
  HB_REQUIRE(n_insns >= 0 && n_insns <= HB_VM_MAX_INSNS, "index program too long: %d instructions (max %d)",
             n_insns, HB_VM_MAX_INSNS);
  HB_REQUIRE(n_vars >= 0 && n_vars <= HB_R, "index program: %d variables", n_vars);
  HB_REQUIRE(n_out >= 0 && n_out <= HB_VM_MAX_OUT, "index program: %d outputs", n_out);
  HB_REQUIRE(n_tensors >= 0 && n_tensors <= HB_VM_MAX_TENSORS, "index program: %d tensor operands (max %d)",
             n_tensors, HB_VM_MAX_TENSORS);
  hb_ixprog *p = new hb_ixprog;
  memset(&p->vm, 0, sizeof p->vm);
  p->vm.n_insns = n_insns;
  p->vm.n_vars = n_vars;
  p->vm.n_out = n_out;
  p->vm.f32 = f32;
  p->n_tensors = n_tensors;
  bool seen_out[HB_VM_MAX_OUT] = {};
  for (int i = 0; i < n_insns; i++) {
    const hb_insn &I = code[i];
    bool ok = I.op < HB_VM_OP_COUNT && I.dst < HB_VM_MAX_REGS && I.a < HB_VM_MAX_REGS;
    if (I.op == HB_VM_IVAR) ok = ok && I.a < n_vars;
    if (I.op == HB_VM_TBL || I.op == HB_VM_ARGMAX || I.op == HB_VM_ARGMIN) {
      ok = ok && I.a < n_tensors && I.n <= 8;
      if (ok) {
        int need = I.n + (I.op == HB_VM_TBL ? 0 : 1);
        ok = tensors[I.a] && tensors[I.a]->rank == need && need <= HB_VM_TRANK;
      }
    }
    if (I.op == HB_VM_OUT) {
      ok = ok && I.b < n_out;
      if (ok) seen_out[I.b] = true;
    }
    if (!ok) {
      delete p;
      HB_FAIL(HB_ERR_INVALID, "index program: malformed instruction %d (op %d)", i, (int)I.op);
    }
    p->vm.code[i] = I;
  }
  for (int c = 0; c < n_out; c++)
    if (!seen_out[c]) {
      delete p;
      HB_FAIL(HB_ERR_INVALID, "index program: output component %d is never written", c);
    }
  for (int t = 0; t < n_tensors; t++) {
    hb_array *a = tensors[t];
    hb_retain(a);
    p->held[t] = a;
    hb_vm_tensor &vt = p->vm.tensors[t];
    vt.ptr = hb_data(a);
    vt.dtype = a->dt;
    vt.rank = a->rank;
    for (int i = 0; i < a->rank && i < HB_VM_TRANK; i++) {
      vt.shape[i] = a->shape[i];
      vt.stride[i] = a->strides[i];
    }
  }
  // affine classification by abstract interpretation over (c0, coef[])
  struct Aff { bool ok; int64_t c0; int64_t k[HB_R]; };
  Aff regs[HB_VM_MAX_REGS];
  for (auto &r : regs) { r.ok = false; r.c0 = 0; memset(r.k, 0, sizeof r.k); }
  p->affine = true;
  for (int i = 0; i < n_insns; i++) {
    const hb_insn &I = code[i];
    Aff r; r.ok = false; r.c0 = 0; memset(r.k, 0, sizeof r.k);
    switch (I.op) {
      case HB_VM_IVAR: r.ok = true; r.k[I.a] = 1; break;
      case HB_VM_ICONST: r.ok = true; r.c0 = I.imm.i; break;
      case HB_VM_IADD: case HB_VM_ISUB: {
        const Aff &x = regs[I.a], &y = regs[I.b];
        if (x.ok && y.ok) {
          r.ok = true;
          int64_t s = I.op == HB_VM_IADD ? 1 : -1;
          r.c0 = x.c0 + s * y.c0;
          for (int k = 0; k < HB_R; k++) r.k[k] = x.k[k] + s * y.k[k];
        }
      } break;
      case HB_VM_INEG: {
        const Aff &x = regs[I.a];
        if (x.ok) { r.ok = true; r.c0 = -x.c0; for (int k = 0; k < HB_R; k++) r.k[k] = -x.k[k]; }
      } break;
      case HB_VM_IMUL: {
        const Aff &x = regs[I.a], &y = regs[I.b];
        auto isconst = [](const Aff &a) { if (!a.ok) return false; for (int k = 0; k < HB_R; k++) if (a.k[k]) return false; return true; };
        if (x.ok && y.ok && (isconst(x) || isconst(y))) {
          const Aff &c = isconst(x) ? x : y; const Aff &v = isconst(x) ? y : x;
          r.ok = true; r.c0 = c.c0 * v.c0;
          for (int k = 0; k < HB_R; k++) r.k[k] = c.c0 * v.k[k];
        }
      } break;
      case HB_VM_OUT: {
        const Aff &x = regs[I.a];
        if (x.ok) { p->c0[I.b] = x.c0; for (int k = 0; k < HB_R; k++) p->coef[I.b][k] = x.k[k]; }
        else p->affine = false;
      } continue;
      default: break;
    }
    regs[I.dst] = r;
  }
  *out = p;
  return HB_OK;
}

extern "C" int hb_ixprog_free(hb_ixprog *p) {
  if (!p) return HB_OK;
  if (p->rc.fetch_sub(1) != 1) return HB_OK;   // still referenced by a recorded lazy node
  for (int t = 0; t < p->n_tensors; t++) hb_release(p->held[t]);
  delete p;
  return HB_OK;
}

extern "C" int hb_ixprog_is_affine(const hb_ixprog *p, int *flag) {
  HB_REQUIRE(p && flag, "null argument");
  *flag = p->affine ? 1 : 0;
  return HB_OK;
}

// ------------------------------------------------------ generated kernels
// Looks up (or generates and compiles) the kernel of `p` for one set of call
// constants.  nullptr -> the caller runs the VM interpreter kernel instead.
static std::mutex g_gen_mu;
template <typename MakeSource>
static CUfunction hb_generated(const hb_ixprog *p, const std::string &key, MakeSource make) {
  if (!hb_jit_enabled()) return nullptr;
  std::lock_guard<std::mutex> lk(g_gen_mu);
  for (auto &e : p->generated)
    if (e.first == key) return (CUfunction)e.second;
  CUfunction f = hb_jit_kernel(make(), "hbj_k");
  p->generated.emplace_back(key, (void *)f);
  return f;
}
static void key_add(std::string &k, const void *p, size_t n) { k.append((const char *)p, n); }

// index-function kernel (kinds: jit.cuh) for shm -> (pshape, pstride)
static CUfunction hb_generated_index(const hb_ixprog *p, int kind, int shm_rank, const int64_t *shm, int prank,
                                     const int64_t *pshape, const int64_t *pstride, int es) {
  std::string key;
  int hdr[4] = {kind, shm_rank, prank, es};
  key_add(key, hdr, sizeof hdr);
  key_add(key, shm, sizeof(int64_t) * shm_rank);
  key_add(key, pshape, sizeof(int64_t) * prank);
  key_add(key, pstride, sizeof(int64_t) * prank);
  return hb_generated(p, key, [&] {
    return hb_jit_index_source(p->vm, p->n_tensors, kind, shm_rank, shm, prank, pshape, pstride, es);
  });
}
// launches a generated index kernel: `lead` = the kind's own leading arguments, then the tensor operands
static int hb_launch_index(CUfunction f, const hb_ixprog *p, int64_t M, void **lead, int n_lead) {
  void *args[8 + HB_VM_MAX_TENSORS];
  const void *tptr[HB_VM_MAX_TENSORS];
  int n = 0;
  for (int i = 0; i < n_lead; i++) args[n++] = lead[i];
  for (int t = 0; t < p->n_tensors; t++) { tptr[t] = p->vm.tensors[t].ptr; args[n++] = &tptr[t]; }
  return hb_jit_launch(f, (unsigned)hb_blocks_for(M, 256), 256, args);
}

// ------------------------------------------------------------- VM kernels
struct hb_shm {  // the shm index space and the target (shp) bounds/strides
  int mrank, prank;
  int64_t mshape[HB_R];
  int64_t pshape[HB_VM_MAX_OUT];
  int64_t pstride[HB_VM_MAX_OUT];
};

__device__ __forceinline__ int64_t hb_vm_target(const hb_vm_prog &P, const hb_shm &S, int64_t m) {
  int64_t vars[HB_R];
  int64_t rem = m;
#pragma unroll 1
  for (int d = S.mrank - 1; d >= 0; d--) {
    int64_t q = rem / S.mshape[d];
    vars[d] = rem - q * S.mshape[d];
    rem = q;
  }
  uint64_t outs[HB_VM_MAX_OUT];
  hb_vm_run(P, vars, hb_vm_noin(), outs);
  int64_t off = 0;
  bool inb = true;
  for (int c = 0; c < S.prank; c++) {
    int64_t ix = (int64_t)outs[c];
    inb = inb && ix >= 0 && ix < S.pshape[c];
    off += ix * S.pstride[c];
  }
  return inb ? off : INT64_MIN;  // INT64_MIN = out of bounds
}

// offsets[m] = strided offset of prog(m) in the target, or INT64_MIN
__global__ void __launch_bounds__(128)
    hb_vm_offsets_kernel(const __grid_constant__ hb_vm_prog P, const __grid_constant__ hb_shm S,
                         int64_t M, int64_t *__restrict__ offsets) {
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M) offsets[m] = hb_vm_target(P, S, m);
}

// fused gather for rank-0 slices: out[m] = src[prog(m)]
template <typename T>
__global__ void __launch_bounds__(128)
    hb_vm_gather0_kernel(const __grid_constant__ hb_vm_prog P, const __grid_constant__ hb_shm S,
                         int64_t M, const T *__restrict__ src, T *__restrict__ out) {
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  int64_t off = hb_vm_target(P, S, m);
  out[m] = off == INT64_MIN ? T(0) : src[off];
}

// out[m, j] = offsets[m] valid ? src[offsets[m] + noff(j)] : 0
template <typename T>
__global__ void __launch_bounds__(256)
    hb_gather_copy_kernel(const int64_t *__restrict__ offsets, hb_dims nd, int64_t Nn, int64_t total,
                          const T *__restrict__ src, T *__restrict__ out) {
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t lin = gid; lin < total; lin += gs) {
    int64_t m = lin / Nn, j = lin - m * Nn;
    int64_t base = offsets[m];
    T v = T(0);
    if (base != INT64_MIN) v = src[base + hb_offset_of(nd, j)];
    out[lin] = v;
  }
}

// affine gather with bound checks: lin over (shm ++ shn)
struct hb_affp {
  int mrank, prank;
  int64_t c0[HB_VM_MAX_OUT], pshape[HB_VM_MAX_OUT];
  int64_t coef[HB_VM_MAX_OUT][HB_R];
};
template <typename T>
__global__ void __launch_bounds__(256)
    hb_gather_affine_kernel(const __grid_constant__ hb_affp A, hb_dims d /* shm++shn with folded strides */,
                            int64_t base, int64_t total, const T *__restrict__ src, T *__restrict__ out) {
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t lin = gid; lin < total; lin += gs) {
    int64_t rem = lin, off = base;
    int64_t p[HB_VM_MAX_OUT];
#pragma unroll
    for (int c = 0; c < HB_VM_MAX_OUT; c++) p[c] = c < A.prank ? A.c0[c] : 0;
#pragma unroll 1
    for (int k = d.rank - 1; k >= 0; k--) {
      int64_t q = rem / d.shape[k], r = rem - q * d.shape[k];
      off += r * d.stride[k];
      if (k < A.mrank) {
#pragma unroll
        for (int c = 0; c < HB_VM_MAX_OUT; c++)
          if (c < A.prank) p[c] += r * A.coef[c][k];
      }
      rem = q;
    }
    bool inb = true;
#pragma unroll
    for (int c = 0; c < HB_VM_MAX_OUT; c++)
      if (c < A.prank) inb = inb && p[c] >= 0 && p[c] < A.pshape[c];
    out[lin] = inb ? src[off] : T(0);
  }
}

// A run longer than ~2048 elements is cut into equal pieces (a divisor of L) so
// that the rows outnumber the warps; returns the number of pieces (1 = keep).
static int64_t hb_run_pieces(int64_t L) {
  if (L <= 2048) return 1;
  for (int64_t len = 2048; len >= 256; len--)
    if (L % len == 0) return L / len;
  return 1;  // no convenient divisor (e.g. a large prime): one long run per row
}

// grid for the row kernels: 8 warps per block, 32 rows per warp for short runs, one for long ones
static int hb_rows_blocks(int64_t rows, int L) {
  int64_t warps = L <= 32 ? (rows + 31) / 32 : rows;
  return hb_blocks_for(warps, 8, g_hb.sm_count * 16);
}

// ------------------------------------------------------- row-structured copy
// The result of a gather is [shm ++ shn]; its innermost collapsed axis is a
// run of L elements that share ONE index computation.  A warp takes 32 rows
// at a time: lane r decomposes row r (so the integer divisions are paid once
// per row, not once per element), then the warp walks the 32 rows with the
// offsets broadcast by shuffles and moves each run with coalesced accesses
// (several short runs per pass when L < 32).
//   TABLE = false: affine program folded into per-axis strides; the bound
//                  checks of the index components ride along.
//   TABLE = true : axis 0 is the whole shm space, its source offset comes from
//                  the table the VM kernel wrote (INT64_MIN = out of bounds).
struct hb_rowsp {
  int rank;               // outer axes (everything except the inner run), in ENUMERATION order
  int prank;              // affine: number of bound-checked index components
  int idx32;              // rows < 2^31: 32-bit row decomposition
  int64_t rows;
  int shape[HB_R];
  int64_t stride[HB_R];   // source stride of each outer axis
  int64_t ostride[HB_R];  // result stride of each outer axis (the result is contiguous)
  int64_t c0[HB_VM_MAX_OUT], pshape[HB_VM_MAX_OUT];
  int64_t coef[HB_VM_MAX_OUT][HB_R];  // index coefficient of each outer axis (0 for shn axes)
  int64_t base;
  int L, lp, lshift;      // run length, its power-of-two padding (<= 32) and log2
  int tshift;             // table offsets are in units of 2^tshift row elements (vectorised rows)
  int64_t s_in;           // source stride along the run
};

// 16-byte element for the vectorised instantiations (two f64 / four f32 ... per access)
struct __align__(16) hb_b16 {
  uint64_t x, y;
  hb_b16() = default;
  __host__ __device__ explicit hb_b16(int) : x(0), y(0) {}
};

// source / result offsets of one row; source = INT64_MIN when an index component is out of bounds
template <bool TABLE, typename IT>
__device__ __forceinline__ void hb_grow_decode(const hb_rowsp &R, const int64_t *__restrict__ table, int64_t row,
                                               int64_t &off, int64_t &ooff) {
  IT rem = (IT)row;
  int64_t o = R.base, oo = 0;
  int64_t pc[HB_VM_MAX_OUT];
#pragma unroll
  for (int c = 0; c < HB_VM_MAX_OUT; c++) pc[c] = c < R.prank ? R.c0[c] : 0;
#pragma unroll 1
  for (int k = R.rank - 1; k >= (TABLE ? 1 : 0); k--) {
    IT q = rem / (IT)R.shape[k];
    int rr = (int)(rem - q * (IT)R.shape[k]);
    o += rr * R.stride[k];
    oo += rr * R.ostride[k];
    if (!TABLE) {
#pragma unroll
      for (int c = 0; c < HB_VM_MAX_OUT; c++)
        if (c < R.prank) pc[c] += rr * R.coef[c][k];
    }
    rem = q;
  }
  bool inb = true;
  if (TABLE) {
    int64_t t = table[rem];
    inb = t != INT64_MIN;
    o += inb ? (t >> R.tshift) : 0;
    oo += (int64_t)rem * R.ostride[0];
  } else {
#pragma unroll
    for (int c = 0; c < HB_VM_MAX_OUT; c++)
      if (c < R.prank) inb = inb && pc[c] >= 0 && pc[c] < R.pshape[c];
  }
  off = inb ? o : INT64_MIN;
  ooff = oo;
}

template <typename T, bool TABLE>
__global__ void __launch_bounds__(256)
    hb_gather_rows_kernel(const __grid_constant__ hb_rowsp R, const int64_t *__restrict__ table,
                          const T *__restrict__ src, T *__restrict__ out) {
  __shared__ longlong2 s_rows[8][32];  // (source, result) offsets of the warp's 32 rows
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // short runs: 32 rows per warp pass (lane r decodes row r); long runs: one row
  // per warp (every lane decodes it; the cost is amortised over the run)
  const int rpg = R.L <= 32 ? 32 : 1;
  for (int64_t r0 = warp * rpg; r0 < R.rows; r0 += nwarps * rpg) {
    int64_t row = r0 + (rpg == 32 ? lane : 0), off = INT64_MIN, ooff = 0;
    if (row < R.rows) {
      if (R.idx32) hb_grow_decode<TABLE, uint32_t>(R, table, row, off, ooff);
      else hb_grow_decode<TABLE, int64_t>(R, table, row, off, ooff);
    }
    if (R.L <= 32) {
      // walk the 32 rows, 32 / lp of them per pass; four passes are loaded before
      // any is stored, for memory-level parallelism
      const int rpp = 32 >> R.lshift, sub = lane >> R.lshift, j0 = lane & (R.lp - 1);
      s_rows[wib][lane] = make_longlong2(off, ooff);
      __syncwarp();
      const int nrows = (int)min((int64_t)32, R.rows - r0);
      for (int rb = 0; rb < nrows; rb += 4 * rpp) {
        T v[4];
        int64_t dsto[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          int rsel = rb + u * rpp + sub;
          longlong2 oo = s_rows[wib][rsel & 31];
          bool live = rsel < nrows && j0 < R.L;
          dsto[u] = live ? oo.y + j0 : -1;
          v[u] = (live && oo.x != INT64_MIN) ? src[oo.x + (int64_t)j0 * R.s_in] : T(0);
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (dsto[u] >= 0) out[dsto[u]] = v[u];
      }
      __syncwarp();
    } else {
      T *dst = out + ooff;
      const bool inb = off != INT64_MIN;
      const T *sp = src + (inb ? off : 0);
      int j = lane;
      for (; j + 96 < R.L; j += 128) {
        T a = T(0), b = T(0), c = T(0), d = T(0);
        if (inb) {
          a = sp[(int64_t)j * R.s_in];
          b = sp[(int64_t)(j + 32) * R.s_in];
          c = sp[(int64_t)(j + 64) * R.s_in];
          d = sp[(int64_t)(j + 96) * R.s_in];
        }
        dst[j] = a; dst[j + 32] = b; dst[j + 64] = c; dst[j + 96] = d;
      }
      for (; j < R.L; j += 32) dst[j] = inb ? sp[(int64_t)j * R.s_in] : T(0);
    }
  }
}

// Collapses the trailing shn axes (source strides against the contiguous
// result) and fills the row description; the first `nlead` entries of
// shape/stride (the shm axes or the single table axis) are given by the caller.
static void rows_finish(hb_rowsp *R, int nlead, int nrank, const int64_t *nshape, const int64_t *nstride) {
  int64_t sh[HB_R], st[1][HB_R];
  int r = nrank;
  for (int i = 0; i < nrank; i++) { sh[i] = nshape[i]; st[0][i] = nstride[i]; }
  // merge adjacent shn axes that are contiguous in the source (the result is contiguous anyway)
  int w = 0;
  for (int i = 0; i < r; i++) {
    if (sh[i] == 1) continue;
    if (w > 0 && st[0][w - 1] == st[0][i] * sh[i]) { sh[w - 1] *= sh[i]; st[0][w - 1] = st[0][i]; }
    else { sh[w] = sh[i]; st[0][w] = st[0][i]; w++; }
  }
  r = w;
  int64_t L = 1, s_in = 1;
  if (r > 0 && sh[r - 1] <= (1 << 30)) {
    L = sh[r - 1]; s_in = st[0][r - 1]; r--;
    int64_t pieces = hb_run_pieces(L);
    if (pieces > 1 && r < HB_R - nlead - 1) {  // long run: [pieces][L / pieces] keeps every warp busy
      sh[r] = pieces; st[0][r] = (L / pieces) * s_in; r++;
      L /= pieces;
    }
  }
  R->rank = nlead + r;
  for (int i = 0; i < r; i++) { R->shape[nlead + i] = (int)sh[i]; R->stride[nlead + i] = st[0][i]; }
  R->L = (int)L;
  R->s_in = s_in;
  R->lshift = 0;
  while ((1 << R->lshift) < L && R->lshift < 5) R->lshift++;
  R->lp = 1 << R->lshift;
  R->rows = 1;
  for (int i = 0; i < R->rank; i++) R->rows *= R->shape[i];
  R->idx32 = R->rows < (1ll << 31);
  int64_t o = L;  // the result is contiguous over (outer axes ++ run)
  for (int i = R->rank - 1; i >= 0; i--) { R->ostride[i] = o; o *= R->shape[i]; }
}

// Overlapping windows (two shm axes feeding one index component, e.g. the
// i + j of an im2col gather) re-read every source element once per window.
// Walking the result in row-major order puts those re-reads a whole plane of
// the other axes apart -- out of L2 for large tensors.  The rows are therefore
// ENUMERATED with the window axes moved inside the other outer axes, but
// outside a tail of innermost axes that keeps every warp's writes in
// contiguous runs of >= 2 KB: re-reads then hit L1/L2 and the DRAM traffic
// drops to the algorithmic bytes.  The result itself is unchanged (each row
// carries its own result offset).
static void rows_window_order(hb_rowsp *R, int nlead, size_t es) {
  bool win[HB_R] = {};
  bool any = false;
  for (int k = 0; k < nlead; k++)
    for (int k2 = 0; k2 < nlead; k2++)
      if (k != k2 && R->shape[k] > 1 && R->shape[k2] > 1)
        for (int c = 0; c < R->prank; c++)
          if (R->coef[c][k] != 0 && R->coef[c][k2] != 0) { win[k] = true; any = true; }
  if (!any) return;
  int tail = R->rank;  // axes [tail, rank) stay innermost
  int64_t run = (int64_t)R->L * (int64_t)es;
  while (tail > 0 && !win[tail - 1] && run < 2048) {
    const int k = tail - 1;
    const int64_t need = (2048 + run - 1) / run;
    int64_t f = 0;  // smallest divisor of the axis that completes a 2 KB run
    if (R->shape[k] >= 2 * need && R->rank < HB_R)
      for (int64_t d = need; d * 2 <= R->shape[k] && d <= need * 64; d++)
        if (R->shape[k] % d == 0) { f = d; break; }
    if (f) {
      // split axis k into [shape / f][f]: the inner part stays in the tail, the outer part goes outside the windows
      for (int i = R->rank; i > k + 1; i--) {
        R->shape[i] = R->shape[i - 1]; R->stride[i] = R->stride[i - 1]; R->ostride[i] = R->ostride[i - 1];
        win[i] = win[i - 1];
        for (int c = 0; c < HB_VM_MAX_OUT; c++) R->coef[c][i] = R->coef[c][i - 1];
      }
      R->rank++;
      R->shape[k + 1] = (int)f; R->stride[k + 1] = R->stride[k]; R->ostride[k + 1] = R->ostride[k]; win[k + 1] = false;
      R->shape[k] = (int)(R->shape[k] / f); R->stride[k] *= f; R->ostride[k] *= f;
      for (int c = 0; c < HB_VM_MAX_OUT; c++) { R->coef[c][k + 1] = R->coef[c][k]; R->coef[c][k] *= f; }
      tail = k + 1;
      break;
    }
    tail--;
    run *= R->shape[tail];
  }
  int order[HB_R], n = 0;
  for (int k = 0; k < tail; k++) if (!win[k]) order[n++] = k;
  for (int k = 0; k < tail; k++) if (win[k]) order[n++] = k;
  for (int k = tail; k < R->rank; k++) order[n++] = k;
  hb_rowsp Q = *R;
  for (int i = 0; i < R->rank; i++) {
    int k = order[i];
    R->shape[i] = Q.shape[k];
    R->stride[i] = Q.stride[k];
    R->ostride[i] = Q.ostride[k];
    for (int c = 0; c < HB_VM_MAX_OUT; c++) R->coef[c][i] = Q.coef[c][k];
  }
}

// 16-byte accesses: when the run, every source stride and the base are
// multiples of v = 16 / sizeof(T) elements and both pointers are 16-byte
// aligned, the same gather is a gather of 16-byte elements.
static bool rows_vectorise(hb_rowsp *R, size_t es, const void *src, const void *dst, int64_t table_align) {
  if (es >= 16 || R->s_in != 1) return false;
  const int64_t v = 16 / (int64_t)es;
  if (R->L % v || R->base % v || table_align % v) return false;
  if (((uintptr_t)src | (uintptr_t)dst) & 15) return false;
  for (int i = 0; i < R->rank; i++)
    if (R->shape[i] > 1 && (R->stride[i] % v || R->ostride[i] % v)) return false;
  R->L /= (int)v;
  R->base /= v;
  while ((1 << R->tshift) < v) R->tshift++;
  for (int i = 0; i < R->rank; i++) { R->stride[i] /= v; R->ostride[i] /= v; }
  R->lshift = 0;
  while ((1 << R->lshift) < R->L && R->lshift < 5) R->lshift++;
  R->lp = 1 << R->lshift;
  return true;
}

static void fill_shm(hb_shm *S, int shm_rank, const int64_t *shm, int shp_rank, const int64_t *pshape,
                     const int64_t *pstride) {
  memset(S, 0, sizeof *S);
  S->mrank = shm_rank;
  S->prank = shp_rank;
  for (int i = 0; i < shm_rank; i++) S->mshape[i] = shm[i];
  for (int i = 0; i < shp_rank; i++) {
    S->pshape[i] = pshape[i];
    S->pstride[i] = pstride[i];
  }
}

extern "C" int hb_gather(const hb_array *src, int shm_rank, const int64_t *shm, int shp_rank,
                         const hb_ixprog *prog, hb_array **out) {
  hb_prof_scope prof_("gather", src);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(src && prog && out, "null argument");
  HB_REQUIRE(shm_rank >= 0 && shm_rank <= HB_R && (shm_rank == 0 || shm), "gather: bad shm");
  HB_REQUIRE(shp_rank >= 0 && shp_rank <= src->rank, "gather: |shp| = %d exceeds source rank %d", shp_rank,
             src->rank);
  HB_REQUIRE(prog->vm.n_vars == shm_rank && prog->vm.n_out == shp_rank,
             "gather: program maps %d -> %d indices, expected %d -> %d", prog->vm.n_vars, prog->vm.n_out,
             shm_rank, shp_rank);
  int nrank = src->rank - shp_rank;
  HB_REQUIRE(shm_rank + nrank <= HB_R, "gather: result rank too large");
  int64_t osh[HB_R];
  int64_t M = 1, Nn = 1;
  for (int i = 0; i < shm_rank; i++) { osh[i] = shm[i]; M *= shm[i]; }
  for (int i = 0; i < nrank; i++) { osh[shm_rank + i] = src->shape[shp_rank + i]; Nn *= src->shape[shp_rank + i]; }
  int orank = shm_rank + nrank;
  if (hb_lazy_on())
    return hb_lrec(HB_L_GATHER, "gather").in(src).iarg(0, shm_rank).iarg(1, shp_rank).prog(prog).out(out, src->dt, orank, osh)
        .commit([=](hb_lnode &n, hb_array *const *in, hb_array **o) {
          return hb_gather(in[0], (int)n.ia[0], n.out[0]->shape, (int)n.ia[1], n.prog, &o[0]);
        });

  if (prog->affine) {
    // folded strides: stride of shm axis k = sum_c coef[c][k] * src.stride[c]
    int64_t fs[HB_R];
    int64_t base = 0;
    bool inb = true;
    for (int c = 0; c < shp_rank; c++) base += prog->c0[c] * src->strides[c];
    for (int k = 0; k < shm_rank; k++) {
      fs[k] = 0;
      for (int c = 0; c < shp_rank; c++) fs[k] += prog->coef[c][k] * src->strides[c];
    }
    for (int c = 0; c < shp_rank; c++) {  // range of component c over the shm box
      int64_t lo = prog->c0[c], hi = prog->c0[c];
      for (int k = 0; k < shm_rank; k++) {
        int64_t e = prog->coef[c][k] * (shm[k] - 1);
        if (shm[k] == 0) e = 0;
        if (e < 0) lo += e; else hi += e;
      }
      if (lo < 0 || hi >= src->shape[c]) inb = false;
    }
    if (inb || M * Nn == 0) {
      // pure view: zero-copy gather (possibly overlapping windows)
      hb_array *v = hb_new_view(src);
      v->rank = orank;
      v->offset += (M * Nn == 0) ? 0 : base;
      for (int k = 0; k < shm_rank; k++) { v->shape[k] = shm[k]; v->strides[k] = fs[k]; }
      for (int i = 0; i < nrank; i++) {
        v->shape[shm_rank + i] = src->shape[shp_rank + i];
        v->strides[shm_rank + i] = src->strides[shp_rank + i];
      }
      *out = v;
      return HB_OK;
    }
    hb_array *o;
    HB_TRY(hb_new_array(src->dt, orank, osh, &o));
    bool small_axes = true;
    for (int k = 0; k < shm_rank; k++) small_axes = small_axes && shm[k] < (1ll << 31);
    for (int i = 0; i < nrank; i++) small_axes = small_axes && src->shape[shp_rank + i] < (1ll << 31);
    if (small_axes) {
      hb_rowsp R;
      memset(&R, 0, sizeof R);
      R.prank = shp_rank;
      R.base = base;
      for (int c = 0; c < shp_rank; c++) {
        R.c0[c] = prog->c0[c];
        R.pshape[c] = src->shape[c];
        for (int k = 0; k < shm_rank; k++) R.coef[c][k] = prog->coef[c][k];
      }
      for (int k = 0; k < shm_rank; k++) { R.shape[k] = (int)shm[k]; R.stride[k] = fs[k]; }
      rows_finish(&R, shm_rank, nrank, src->shape + shp_rank, src->strides + shp_rank);
      rows_window_order(&R, shm_rank, hb_dtype_size(src->dt));
      int blocks = hb_rows_blocks(R.rows, R.L);
      if (rows_vectorise(&R, hb_dtype_size(src->dt), hb_data(src), hb_data(o), 0)) {
        hb_gather_rows_kernel<hb_b16, false><<<blocks, 256, 0, g_hb.stream>>>(
            R, nullptr, (const hb_b16 *)hb_data(src), (hb_b16 *)hb_data(o));
      } else {
        HB_DISPATCH_BYTES(src->dt, T,
                          (hb_gather_rows_kernel<T, false><<<blocks, 256, 0, g_hb.stream>>>(
                              R, nullptr, (const T *)hb_data(src), (T *)hb_data(o))));
      }
      HB_LAUNCHED();
      *out = o;
      return HB_OK;
    }
    hb_affp A;
    memset(&A, 0, sizeof A);
    A.mrank = shm_rank;
    A.prank = shp_rank;
    for (int c = 0; c < shp_rank; c++) {
      A.c0[c] = prog->c0[c];
      A.pshape[c] = src->shape[c];
      for (int k = 0; k < shm_rank; k++) A.coef[c][k] = prog->coef[c][k];
    }
    hb_dims d;
    d.rank = orank;
    for (int i = 0; i < HB_R; i++) {
      d.shape[i] = i < orank ? osh[i] : 1;
      d.stride[i] = i < shm_rank ? fs[i] : (i < orank ? src->strides[shp_rank + (i - shm_rank)] : 0);
    }
    int64_t total = M * Nn;
    int blocks = hb_blocks_for(total, 256, g_hb.sm_count * 32);
    HB_DISPATCH_BYTES(src->dt, T,
                      (hb_gather_affine_kernel<T><<<blocks, 256, 0, g_hb.stream>>>(
                          A, d, base, total, (const T *)hb_data(src), (T *)hb_data(o))));
    HB_LAUNCHED();
    *out = o;
    return HB_OK;
  }

  hb_array *o;
  HB_TRY(hb_new_array(src->dt, orank, osh, &o));
  int64_t total = M * Nn;
  if (total == 0) { *out = o; return HB_OK; }
  hb_shm S;
  fill_shm(&S, shm_rank, shm, shp_rank, src->shape, src->strides);
  if (Nn == 1) {
    int64_t tail = 0;  // offset of the (all size-1) shn axes is 0
    (void)tail;
    CUfunction gen = hb_generated_index(prog, 1, shm_rank, shm, shp_rank, src->shape, src->strides,
                                        (int)hb_dtype_size(src->dt));
    if (gen) {
      const void *sp = hb_data(src);
      void *op = hb_data(o);
      void *lead[2] = {&sp, &op};
      int st = hb_launch_index(gen, prog, M, lead, 2);
      if (st != HB_OK) { hb_release(o); return st; }
    } else {
      int blocks = hb_blocks_for(M, 128);
      HB_DISPATCH_BYTES(src->dt, T,
                        (hb_vm_gather0_kernel<T><<<blocks, 128, 0, g_hb.stream>>>(
                            prog->vm, S, M, (const T *)hb_data(src), (T *)hb_data(o))));
    }
    HB_LAUNCHED();
  } else {
    void *scr;
    int st = hb_scratch((size_t)M * 8, &scr);
    if (st != HB_OK) { hb_release(o); return st; }
    CUfunction gen = hb_generated_index(prog, 0, shm_rank, shm, shp_rank, src->shape, src->strides, 0);
    if (gen) {
      void *lead[1] = {&scr};
      st = hb_launch_index(gen, prog, M, lead, 1);
      if (st != HB_OK) { hb_release(o); return st; }
    } else {
      hb_vm_offsets_kernel<<<hb_blocks_for(M, 128), 128, 0, g_hb.stream>>>(prog->vm, S, M, (int64_t *)scr);
    }
    HB_LAUNCHED();
    bool small_axes = M < (1ll << 31);
    for (int i = 0; i < nrank; i++) small_axes = small_axes && src->shape[shp_rank + i] < (1ll << 31);
    if (small_axes) {
      hb_rowsp R;
      memset(&R, 0, sizeof R);
      R.shape[0] = (int)M;
      rows_finish(&R, 1, nrank, src->shape + shp_rank, src->strides + shp_rank);
      int blocks = hb_rows_blocks(R.rows, R.L);
      int64_t talign = 0;  // table offsets are combinations of the indexed axes' strides
      for (int c = 0; c < shp_rank; c++)
        if (src->shape[c] > 1) talign |= src->strides[c];
      if (rows_vectorise(&R, hb_dtype_size(src->dt), hb_data(src), hb_data(o), talign & 15)) {
        hb_gather_rows_kernel<hb_b16, true><<<blocks, 256, 0, g_hb.stream>>>(
            R, (const int64_t *)scr, (const hb_b16 *)hb_data(src), (hb_b16 *)hb_data(o));
      } else {
        HB_DISPATCH_BYTES(src->dt, T,
                          (hb_gather_rows_kernel<T, true><<<blocks, 256, 0, g_hb.stream>>>(
                              R, (const int64_t *)scr, (const T *)hb_data(src), (T *)hb_data(o))));
      }
      HB_LAUNCHED();
    } else {
      hb_dims nd;
      nd.rank = nrank;
      for (int i = 0; i < HB_R; i++) {
        nd.shape[i] = i < nrank ? src->shape[shp_rank + i] : 1;
        nd.stride[i] = i < nrank ? src->strides[shp_rank + i] : 0;
      }
      int blocks = hb_blocks_for(total, 256, g_hb.sm_count * 32);
      HB_DISPATCH_BYTES(src->dt, T,
                        (hb_gather_copy_kernel<T><<<blocks, 256, 0, g_hb.stream>>>(
                            (const int64_t *)scr, nd, Nn, total, (const T *)hb_data(src), (T *)hb_data(o))));
      HB_LAUNCHED();
    }
  }
  *out = o;
  return HB_OK;
}

// ------------------------------------------------------------------- scan
// exclusive scan of int32 counts, tiles of 1024 (256 threads x 4)
__global__ void __launch_bounds__(256)
    hb_scan_tiles_kernel(const int *__restrict__ in, int *__restrict__ out, int *__restrict__ tile_sums,
                         int64_t n) {
  __shared__ int warp_sums[8];
  int64_t base = (int64_t)blockIdx.x * 1024 + threadIdx.x * 4;
  int v[4];
  int s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    v[i] = base + i < n ? in[base + i] : 0;
    s += v[i];
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = s;
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[w] = incl;
  __syncthreads();
  if (w == 0) {
    int ws = lane < 8 ? warp_sums[lane] : 0;
    int wi = ws;
    for (int o = 1; o < 8; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < 8) warp_sums[lane] = wi - ws;  // exclusive warp offsets
    if (lane == 7 && tile_sums) tile_sums[blockIdx.x] = wi;
  }
  __syncthreads();
  int excl = incl - s + warp_sums[w];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if (base + i < n) out[base + i] = excl;
    excl += v[i];
  }
}
__global__ void hb_scan_add_kernel(int *__restrict__ out, const int *__restrict__ tile_offs, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  int add = tile_offs[blockIdx.x];
  for (int k = 0; k < 4; k++, i += 256)
    if (i < n) out[i] += add;
}

// out may alias in.  tmp: scratch of >= (n/1024 + 2) * 2 ints
static int exclusive_scan_i32(const int *in, int *out, int64_t n, int *tmp) {
  if (n <= 0) return HB_OK;
  int64_t tiles = (n + 1023) / 1024;
  hb_scan_tiles_kernel<<<(unsigned)tiles, 256, 0, g_hb.stream>>>(in, out, tiles > 1 ? tmp : nullptr, n);
  HB_LAUNCHED();
  if (tiles > 1) {
    HB_TRY(exclusive_scan_i32(tmp, tmp, tiles, tmp + tiles + 1));
    hb_scan_add_kernel<<<(unsigned)tiles, 256, 0, g_hb.stream>>>(out, tmp, n);
    HB_LAUNCHED();
  }
  return HB_OK;
}

// ---------------------------------------------------------------- scatter
// pass 1: dest[m] (linear index in shp, or -1), counts[dest]++, collision flag
__global__ void __launch_bounds__(128)
    hb_scatter_dest_kernel(const __grid_constant__ hb_vm_prog P, const __grid_constant__ hb_shm S,
                           int64_t M, int *__restrict__ dest, int *__restrict__ counts,
                           int *__restrict__ collision) {
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  int64_t off = hb_vm_target(P, S, m);
  int d = off == INT64_MIN ? -1 : (int)off;
  dest[m] = d;
  if (d >= 0) {
    int old = atomicAdd(&counts[d], 1);
    if (old > 0) *collision = 1;
  }
}

// no destination hit twice: out[dest[m], j] = src[m, j]
template <typename T>
__global__ void __launch_bounds__(256)
    hb_scatter_direct_kernel(const int *__restrict__ dest, const int *__restrict__ collision, hb_dims sd,
                             int64_t Nn, int64_t total, const T *__restrict__ src, T *__restrict__ out) {
  if (*collision) return;
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t lin = gid; lin < total; lin += gs) {
    int64_t m = lin / Nn, j = lin - m * Nn;
    int d = dest[m];
    if (d >= 0) out[(int64_t)d * Nn + j] = src[hb_offset_of(sd, lin)];
  }
}

__global__ void __launch_bounds__(128)
    hb_scatter_fill_kernel(const int *__restrict__ dest, const int *__restrict__ collision,
                           const int *__restrict__ starts, int *__restrict__ cursor, int *__restrict__ seg,
                           int64_t M) {
  if (!*collision) return;
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  int d = dest[m];
  if (d >= 0) {
    int slot = atomicAdd(&cursor[d], 1);
    seg[starts[d] + slot] = (int)m;
  }
}

// sort each destination's source list ascending (restores row-major source
// order regardless of the order the atomics resolved in)
__global__ void __launch_bounds__(128)
    hb_scatter_sort_kernel(const int *__restrict__ collision, const int *__restrict__ starts,
                           const int *__restrict__ counts, int *__restrict__ seg, int64_t Pn) {
  if (!*collision) return;
  int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= Pn) return;
  int c = counts[d];
  if (c < 2 || c > 64) return;  // long segments: hb_scatter_sort_long_kernel
  int *s = seg + starts[d];
  for (int i = 1; i < c; i++) {
    int key = s[i], j = i - 1;
    while (j >= 0 && s[j] > key) { s[j + 1] = s[j]; j--; }
    s[j + 1] = key;
  }
}

// one block per destination with a long list: same-direction bitonic network
// with virtual +inf padding
__global__ void __launch_bounds__(256)
    hb_scatter_sort_long_kernel(const int *__restrict__ collision, const int *__restrict__ starts,
                                const int *__restrict__ counts, int *__restrict__ seg, int64_t Pn) {
  if (!*collision) return;
  for (int64_t d = blockIdx.x; d < Pn; d += gridDim.x) {
    int c = counts[d];
    if (c <= 64) continue;
    int *s = seg + starts[d];
    int n = 1;
    while (n < c) n <<= 1;
    for (int k = 2; k <= n; k <<= 1) {
      // first step: mirrored compare
      for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
        int blk = t / (k / 2), o = t % (k / 2);
        int i = blk * k + o, j = blk * k + k - 1 - o;
        if (j < c && s[i] > s[j]) { int x = s[i]; s[i] = s[j]; s[j] = x; }
      }
      __syncthreads();
      for (int h = k / 4; h > 0; h >>= 1) {
        for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
          int blk = t / h, o = t % h;
          int i = blk * 2 * h + o, j = i + h;
          if (j < c && s[i] > s[j]) { int x = s[i]; s[i] = s[j]; s[j] = x; }
        }
        __syncthreads();
      }
    }
  }
}

// out[d, j] = sum over the sorted source list (sequential, reference order)
template <typename T>
__global__ void __launch_bounds__(256)
    hb_scatter_sum_kernel(const int *__restrict__ collision, const int *__restrict__ starts,
                          const int *__restrict__ counts, const int *__restrict__ seg, hb_dims md,
                          hb_dims nd, int64_t Nn, int64_t total, const T *__restrict__ src,
                          T *__restrict__ out) {
  if (!*collision) return;
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t lin = gid; lin < total; lin += gs) {
    int64_t d = lin / Nn, j = lin - d * Nn;
    int c = counts[d];
    if (c == 0) continue;  // stays zero
    const int *s = seg + starts[d];
    int64_t joff = hb_offset_of(nd, j);
    T acc;
    if (Nn == 1) {
      // rank-0 slices: vec starts at 0 and adds u + u2 (OpsConcrete.hs:1594-1601)
      acc = T(0);
      for (int i = 0; i < c; i++) acc = src[hb_offset_of(md, s[i]) + joff] + acc;
    } else {
      // IntMap.insertWith (+): first slice stored as is, later ones new + old
      acc = src[hb_offset_of(md, s[0]) + joff];
      for (int i = 1; i < c; i++) acc = src[hb_offset_of(md, s[i]) + joff] + acc;
    }
    out[lin] = acc;
  }
}

// ------------------------------------------------ row-structured scatter
// Same decomposition as hb_gather_rows_kernel: a row = (leading index, outer
// shn index), its inner run of L elements shares the index work.
struct hb_srowsp {
  int rank;                // axis 0 = the leading index space (M sources or Pn destinations)
  int idx32;               // rows < 2^31: 32-bit row decomposition
  int shape[HB_R];
  int64_t sstride[HB_R];   // source strides of the outer shn axes (entry 0 unused)
  int64_t ostride[HB_R];   // strides of the same axes inside one contiguous result slice
  int64_t rows, Nn;
  int L, lp, lshift;       // run length in accesses of V elements
  int64_t s_in;
};

// moff[m] = source offset of shm position m
__global__ void __launch_bounds__(256)
    hb_scatter_moff_kernel(hb_dims md, int64_t M, int64_t *__restrict__ moff) {
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M) moff[m] = hb_offset_of(md, m);
}

template <typename IT>
__device__ __forceinline__ void hb_srow_decode_t(const hb_srowsp &R, int64_t row, int64_t &lead, int64_t &soff,
                                                 int64_t &ooff) {
  IT rem = (IT)row;
  soff = 0;
  ooff = 0;
#pragma unroll 1
  for (int k = R.rank - 1; k >= 1; k--) {
    IT q = rem / (IT)R.shape[k];
    int rr = (int)(rem - q * (IT)R.shape[k]);
    soff += rr * R.sstride[k];
    ooff += rr * R.ostride[k];
    rem = q;
  }
  lead = (int64_t)rem;
}
__device__ __forceinline__ void hb_srow_decode(const hb_srowsp &R, int64_t row, int64_t &lead, int64_t &soff,
                                               int64_t &ooff) {
  if (R.idx32) hb_srow_decode_t<uint32_t>(R, row, lead, soff, ooff);
  else hb_srow_decode_t<int64_t>(R, row, lead, soff, ooff);
}

// no destination hit twice: out[dest[m], ...] = src[m, ...]
template <typename T>
__global__ void __launch_bounds__(256)
    hb_scatter_direct_rows_kernel(const __grid_constant__ hb_srowsp R, const int *__restrict__ dest,
                                  const int *__restrict__ collision, const int64_t *__restrict__ moff,
                                  const T *__restrict__ src, T *__restrict__ out) {
  if (*collision) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int rpg = R.L <= 32 ? 32 : 1;
  for (int64_t r0 = warp * rpg; r0 < R.rows; r0 += nwarps * rpg) {
    int64_t row = r0 + (rpg == 32 ? lane : 0), so = 0, oo = INT64_MIN;
    if (row < R.rows) {
      int64_t m, soff, ooff;
      hb_srow_decode(R, row, m, soff, ooff);
      int d = dest[m];
      if (d >= 0) { so = moff[m] + soff; oo = (int64_t)d * R.Nn + ooff; }
    }
    const int rpp = 32 >> R.lshift, sub = lane >> R.lshift, j0 = lane & (R.lp - 1);
    for (int rb = 0; rb < rpg; rb += rpp) {
      int rsel = rb + sub;
      int64_t s_ = __shfl_sync(0xffffffffu, so, rsel), o_ = __shfl_sync(0xffffffffu, oo, rsel);
      if (r0 + rsel >= R.rows || o_ == INT64_MIN) continue;
      for (int j = j0; j < R.L; j += R.lp) out[o_ + j] = src[s_ + (int64_t)j * R.s_in];
    }
  }
}

// V elements of T moved by one access
template <typename T, int V>
struct alignas(sizeof(T) * V) hb_vecn {
  T e[V];
};
template <typename T, int V>
__device__ __forceinline__ hb_vecn<T, V> hb_vecn_zero() {
  hb_vecn<T, V> z;
#pragma unroll
  for (int i = 0; i < V; i++) z.e[i] = T(0);
  return z;
}

// out[d, ...] = sum over the sorted source list of d (sequential, reference
// order); destinations nobody writes get their zeros here (the result is NOT
// zero-filled beforehand).  When no destination is hit twice the direct
// kernel has moved the data and only those zeros are left to write.
// V > 1: the run, the slice strides and the source offsets are multiples of
// V elements and 16-byte aligned -> one access moves V elements.
template <typename T, int V>
__global__ void __launch_bounds__(256)
    hb_scatter_sum_rows_kernel(const __grid_constant__ hb_srowsp R, const int *__restrict__ collision,
                               const int *__restrict__ starts, const int *__restrict__ counts,
                               const int *__restrict__ seg, const int64_t *__restrict__ moff,
                               const T *__restrict__ src, T *__restrict__ out) {
  typedef hb_vecn<T, V> VT;
  const bool coll = *collision != 0;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int rpg = R.L <= 32 ? 32 : 1;
  const int64_t Lt = (int64_t)R.L * V;  // run length in elements
  for (int64_t r0 = warp * rpg; r0 < R.rows; r0 += nwarps * rpg) {
    int64_t row = r0 + (rpg == 32 ? lane : 0), so = 0;
    int cnt = 0, st = 0;
    if (row < R.rows) {
      int64_t d, soff, ooff;
      hb_srow_decode(R, row, d, soff, ooff);
      so = soff;
      cnt = counts[d];
      st = starts[d];
    }
    const int rpp = 32 >> R.lshift, sub = lane >> R.lshift, j0 = lane & (R.lp - 1);
    for (int rb = 0; rb < rpg; rb += rpp) {
      int rsel = rb + sub;
      int64_t s_ = __shfl_sync(0xffffffffu, so, rsel);
      int c = __shfl_sync(0xffffffffu, cnt, rsel), st_ = __shfl_sync(0xffffffffu, st, rsel);
      int64_t rr = r0 + rsel;
      if (rr >= R.rows) continue;
      VT *op = (VT *)(out + rr * Lt);
      if (c == 0) {
        const VT z = hb_vecn_zero<T, V>();
        for (int j = j0; j < R.L; j += R.lp) op[j] = z;
        continue;
      }
      if (!coll) continue;  // moved by hb_scatter_direct_rows_kernel
      const int *lst = seg + st_;
      if (c <= 8) {
        // the source offsets of this destination are the same for the whole run:
        // fetch them once, then every access is c independent loads
        int64_t so8[8];
#pragma unroll
        for (int i = 0; i < 8; i++) so8[i] = i < c ? moff[lst[i]] : 0;
        for (int j = j0; j < R.L; j += R.lp) {
          VT v[8];
#pragma unroll
          for (int i = 0; i < 8; i++)
            if (i < c) {
              if (V == 1) v[i].e[0] = src[s_ + so8[i] + (int64_t)j * R.s_in];
              else v[i] = ((const VT *)(src + s_ + so8[i]))[j];
            }
          VT acc;
#pragma unroll
          for (int e = 0; e < V; e++) acc.e[e] = R.Nn == 1 ? v[0].e[e] + T(0) : v[0].e[e];
#pragma unroll
          for (int i = 1; i < 8; i++)
            if (i < c) {
#pragma unroll
              for (int e = 0; e < V; e++) acc.e[e] = v[i].e[e] + acc.e[e];
            }
          op[j] = acc;
        }
        continue;
      }
      for (int j = j0; j < R.L; j += R.lp) {
        VT acc;
        // rank-0 slices: vec starts at 0 and adds u + u2 (OpsConcrete.hs:1594-1601);
        // otherwise IntMap.insertWith (+): first slice stored as is, later ones new + old
        for (int i = 0; i < c; i++) {
          VT v;
          if (V == 1) v.e[0] = src[s_ + moff[lst[i]] + (int64_t)j * R.s_in];
          else v = ((const VT *)(src + s_ + moff[lst[i]]))[j];
#pragma unroll
          for (int e = 0; e < V; e++) {
            if (i == 0) acc.e[e] = R.Nn == 1 ? v.e[e] + T(0) : v.e[e];
            else acc.e[e] = v.e[e] + acc.e[e];
          }
        }
        op[j] = acc;
      }
    }
  }
}

// Rows over [lead] ++ shn with the innermost collapsed shn axis as the run.
static void srows_fill(hb_srowsp *R, int64_t lead, int nrank, const int64_t *nshape, const int64_t *nstride) {
  memset(R, 0, sizeof *R);
  int64_t sh[HB_R], ss[HB_R], os[HB_R];
  int w = 0;
  int64_t Nn = 1;
  for (int i = 0; i < nrank; i++) Nn *= nshape[i];
  R->Nn = Nn;
  for (int i = 0; i < nrank; i++) {
    if (nshape[i] == 1) continue;
    if (w > 0 && ss[w - 1] == nstride[i] * nshape[i]) { sh[w - 1] *= nshape[i]; ss[w - 1] = nstride[i]; }
    else { sh[w] = nshape[i]; ss[w] = nstride[i]; w++; }
  }
  int r = w;
  int64_t L = 1, s_in = 1;
  if (r > 0 && sh[r - 1] <= (1 << 30)) {
    L = sh[r - 1]; s_in = ss[r - 1]; r--;
    int64_t pieces = hb_run_pieces(L);
    if (pieces > 1 && r < HB_R - 2) {
      sh[r] = pieces; ss[r] = (L / pieces) * s_in; r++;
      L /= pieces;
    }
  }
  int64_t o = L;
  for (int i = r - 1; i >= 0; i--) { os[i] = o; o *= sh[i]; }
  R->rank = 1 + r;
  R->shape[0] = (int)lead;
  for (int i = 0; i < r; i++) { R->shape[1 + i] = (int)sh[i]; R->sstride[1 + i] = ss[i]; R->ostride[1 + i] = os[i]; }
  R->L = (int)L;
  R->s_in = s_in;
  R->lshift = 0;
  while ((1 << R->lshift) < L && R->lshift < 5) R->lshift++;
  R->lp = 1 << R->lshift;
  R->rows = lead;
  for (int i = 0; i < r; i++) R->rows *= sh[i];
  R->idx32 = R->rows < (1ll << 31);
}

// can the rows be moved V elements per access?  (run and slice strides multiples of V)
static bool srows_can_vec(const hb_srowsp *R, int V) {
  if (V <= 1 || R->s_in != 1 || R->L % V) return false;
  for (int i = 1; i < R->rank; i++)
    if (R->shape[i] > 1 && R->sstride[i] % V) return false;
  return true;
}
static void srows_set_vec(hb_srowsp *R, int V) {
  R->L /= V;
  R->lshift = 0;
  while ((1 << R->lshift) < R->L && R->lshift < 5) R->lshift++;
  R->lp = 1 << R->lshift;
}

extern "C" int hb_scatter_add(const hb_array *src, int shm_rank, int shp_rank, const int64_t *shp,
                              const hb_ixprog *prog, hb_array **out) {
  hb_prof_scope prof_("scatter_add", src);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(src && prog && out, "null argument");
  HB_REQUIRE(shm_rank >= 0 && shm_rank <= src->rank, "scatter: |shm| = %d exceeds source rank %d", shm_rank,
             src->rank);
  HB_REQUIRE(shp_rank >= 0 && shp_rank <= HB_VM_MAX_OUT && (shp_rank == 0 || shp), "scatter: bad shp");
  HB_REQUIRE(prog->vm.n_vars == shm_rank && prog->vm.n_out == shp_rank,
             "scatter: program maps %d -> %d indices, expected %d -> %d", prog->vm.n_vars, prog->vm.n_out,
             shm_rank, shp_rank);
  HB_REQUIRE(src->dt != HB_BOOL, "scatter: numeric dtype required");
  int nrank = src->rank - shm_rank;
  HB_REQUIRE(shp_rank + nrank <= HB_R, "scatter: result rank too large");
  int64_t osh[HB_R];
  int64_t M = 1, Nn = 1, Pn = 1;
  for (int i = 0; i < shm_rank; i++) M *= src->shape[i];
  for (int i = 0; i < shp_rank; i++) { osh[i] = shp[i]; Pn *= shp[i]; }
  for (int i = 0; i < nrank; i++) { osh[shp_rank + i] = src->shape[shm_rank + i]; Nn *= src->shape[shm_rank + i]; }
  if (hb_lazy_on())
    return hb_lrec(HB_L_SCATTER, "scatter_add").in(src).iarg(0, shm_rank).iarg(1, shp_rank).prog(prog)
        .out(out, src->dt, shp_rank + nrank, osh)
        .commit([=](hb_lnode &n, hb_array *const *in, hb_array **o) {
          return hb_scatter_add(in[0], (int)n.ia[0], (int)n.ia[1], n.out[0]->shape, n.prog, &o[0]);
        });
  hb_array *o;
  bool rows_ok = true;
  for (int i = 0; i < nrank; i++) rows_ok = rows_ok && src->shape[shm_rank + i] < (1ll << 31);
  if (M * Nn == 0 || Pn == 0) {
    HB_TRY(hb_zeros(src->dt, shp_rank + nrank, osh, &o));
    *out = o;
    return HB_OK;
  }
  HB_REQUIRE(M < (1ll << 31) - 1 && Pn < (1ll << 31) - 1,
             "scatter: index spaces of 2^31 positions or more are not supported");
  // the row kernels write every element of the result exactly once (zeros included)
  if (rows_ok) HB_TRY(hb_new_array(src->dt, shp_rank + nrank, osh, &o));
  else HB_TRY(hb_zeros(src->dt, shp_rank + nrank, osh, &o));
  // contiguous strides of the shp space
  int64_t pstr[HB_VM_MAX_OUT];
  { int64_t s = 1; for (int i = shp_rank - 1; i >= 0; i--) { pstr[i] = s; s *= shp[i]; } }
  hb_shm S;
  fill_shm(&S, shm_rank, src->shape, shp_rank, shp, pstr);
  // scratch layout (ints): dest[M] | counts[Pn] | starts[Pn] | cursor[Pn] | seg[M] | flag | scan tmp
  int64_t tiles = (Pn + 1023) / 1024;
  size_t ints = (size_t)M * 2 + (size_t)Pn * 3 + 16 + (size_t)(tiles + 2) * 4;
  ints = (ints + 1) & ~(size_t)1;
  void *scr;
  int st = hb_scratch(ints * 4 + (size_t)M * 8, &scr);  // + moff[M] (int64)
  if (st != HB_OK) { hb_release(o); return st; }
  int64_t *moff = (int64_t *)((int *)scr + ints);
  int *dest = (int *)scr;
  int *counts = dest + M;
  int *starts = counts + Pn;
  int *cursor = starts + Pn;
  int *flag = cursor + Pn;  // 16 ints reserved
  int *seg = flag + 16;
  int *tmp = seg + M;
  // zero counts, starts, cursor, flag in one memset
  cudaError_t e = cudaMemsetAsync(counts, 0, ((size_t)Pn * 3 + 16) * 4, g_hb.stream);
  if (e != cudaSuccess) { hb_release(o); HB_CUDA(e); }
  {
    CUfunction gen = hb_generated_index(prog, 2, shm_rank, src->shape, shp_rank, shp, pstr, 0);
    if (gen) {
      void *lead[3] = {&dest, &counts, &flag};
      st = hb_launch_index(gen, prog, M, lead, 3);
      if (st != HB_OK) { hb_release(o); return st; }
    } else {
      hb_scatter_dest_kernel<<<hb_blocks_for(M, 128), 128, 0, g_hb.stream>>>(prog->vm, S, M, dest, counts, flag);
    }
  }
  HB_LAUNCHED();
  hb_dims sd;  // full source dims (shm ++ shn) for the direct kernel
  sd.rank = src->rank;
  hb_dims md;  // source offset of an shm position
  md.rank = shm_rank;
  hb_dims nd;  // source offset within a slice
  nd.rank = nrank;
  for (int i = 0; i < HB_R; i++) {
    sd.shape[i] = i < src->rank ? src->shape[i] : 1;
    sd.stride[i] = i < src->rank ? src->strides[i] : 0;
    md.shape[i] = i < shm_rank ? src->shape[i] : 1;
    md.stride[i] = i < shm_rank ? src->strides[i] : 0;
    nd.shape[i] = i < nrank ? src->shape[shm_rank + i] : 1;
    nd.stride[i] = i < nrank ? src->strides[shm_rank + i] : 0;
  }
  int64_t total_src = M * Nn, total_dst = Pn * Nn;
  int bsrc = hb_blocks_for(total_src, 256, g_hb.sm_count * 32);
  int bdst = hb_blocks_for(total_dst, 256, g_hb.sm_count * 32);
  st = exclusive_scan_i32(counts, starts, Pn, tmp);
  if (st != HB_OK) { hb_release(o); return st; }
  hb_srowsp Rs, Rd;
  int vec = 1;
  if (rows_ok) {
    srows_fill(&Rs, M, nrank, src->shape + shm_rank, src->strides + shm_rank);
    srows_fill(&Rd, Pn, nrank, src->shape + shm_rank, src->strides + shm_rank);
    hb_scatter_moff_kernel<<<hb_blocks_for(M, 256), 256, 0, g_hb.stream>>>(md, M, moff);
    HB_LAUNCHED();
    // 16-byte accesses in the sum kernel: source offsets of the shm positions, the
    // slice strides and the run are multiples of V elements, pointers 16-byte aligned
    int V = 16 / (int)hb_dtype_size(src->dt);
    bool ok = srows_can_vec(&Rd, V) && ((((uintptr_t)hb_data(src)) | ((uintptr_t)hb_data(o))) & 15) == 0;
    for (int i = 0; i < shm_rank; i++) ok = ok && (src->shape[i] == 1 || src->strides[i] % V == 0);
    if (ok) { vec = V; srows_set_vec(&Rd, V); }
  }
  HB_DISPATCH_NUM(src->dt, T, {
    if (rows_ok)
      hb_scatter_direct_rows_kernel<T><<<hb_rows_blocks(Rs.rows, Rs.L), 256, 0, g_hb.stream>>>(
          Rs, dest, flag, moff, (const T *)hb_data(src), (T *)hb_data(o));
    else
      hb_scatter_direct_kernel<T><<<bsrc, 256, 0, g_hb.stream>>>(dest, flag, sd, Nn, total_src,
                                                                 (const T *)hb_data(src), (T *)hb_data(o));
    HB_LAUNCHED();
    hb_scatter_fill_kernel<<<hb_blocks_for(M, 128), 128, 0, g_hb.stream>>>(dest, flag, starts, cursor, seg, M);
    HB_LAUNCHED();
    hb_scatter_sort_kernel<<<hb_blocks_for(Pn, 128), 128, 0, g_hb.stream>>>(flag, starts, counts, seg, Pn);
    HB_LAUNCHED();
    hb_scatter_sort_long_kernel<<<hb_blocks_for(Pn, 1, g_hb.sm_count * 4), 256, 0, g_hb.stream>>>(flag, starts, counts, seg, Pn);
    HB_LAUNCHED();
    if (rows_ok) {
      constexpr int VV = 16 / (int)sizeof(T);
      if (vec > 1)
        hb_scatter_sum_rows_kernel<T, VV><<<hb_rows_blocks(Rd.rows, Rd.L), 256, 0, g_hb.stream>>>(
            Rd, flag, starts, counts, seg, moff, (const T *)hb_data(src), (T *)hb_data(o));
      else
        hb_scatter_sum_rows_kernel<T, 1><<<hb_rows_blocks(Rd.rows, Rd.L), 256, 0, g_hb.stream>>>(
            Rd, flag, starts, counts, seg, moff, (const T *)hb_data(src), (T *)hb_data(o));
    } else {
      hb_scatter_sum_kernel<T><<<bdst, 256, 0, g_hb.stream>>>(flag, starts, counts, seg, md, nd, Nn, total_dst,
                                                              (const T *)hb_data(src), (T *)hb_data(o));
    }
    HB_LAUNCHED();
  });
  *out = o;
  return HB_OK;
}

extern "C" int hb_onehot(const hb_array *v, int shp_rank, const int64_t *shp, const int64_t *ix,
                         hb_array **out) {
  hb_prof_scope prof_("onehot", v);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(v && out, "null argument");
  HB_REQUIRE(shp_rank >= 0 && shp_rank + v->rank <= HB_R, "oneHot: rank too large");
  int64_t osh[HB_R];
  for (int i = 0; i < shp_rank; i++) osh[i] = shp[i];
  for (int i = 0; i < v->rank; i++) osh[shp_rank + i] = v->shape[i];
  if (hb_lazy_on()) {
    std::vector<int64_t> ix_(ix, ix + shp_rank);
    return hb_lrec(HB_L_GENERIC, "onehot").in(v).iarg(0, shp_rank).out(out, v->dt, shp_rank + v->rank, osh).commit(
        [ix_](hb_lnode &n, hb_array *const *in, hb_array **o) {
          return hb_onehot(in[0], (int)n.ia[0], n.out[0]->shape, ix_.data(), &o[0]);
        });
  }
  hb_array *o;
  HB_TRY(hb_zeros(v->dt, shp_rank + v->rank, osh, &o));
  bool inb = true;
  for (int i = 0; i < shp_rank; i++) inb = inb && ix[i] >= 0 && ix[i] < shp[i];
  if (inb && hb_numel(v) > 0) {
    hb_array *slot = nullptr;
    int st = hb_index_prefix(o, shp_rank, ix, &slot);
    if (st == HB_OK) {
      // copy v into the slot: reuse the generic strided copy through hb_iter
      const hb_array *ops[2] = {slot, v};
      hb_iter it;
      hb_make_iter(2, ops, &it);
      HB_DISPATCH_BYTES(v->dt, T, {
        T *po = (T *)hb_data(slot);
        const T *pi = (const T *)hb_data(v);
        st = hb_launch_iter<2>(it, [=] __device__(int64_t, const int64_t *off) { po[off[0]] = pi[off[1]]; });
      });
    }
    hb_release(slot);
    if (st != HB_OK) { hb_release(o); return st; }
  }
  *out = o;
  return HB_OK;
}

// -------------------------------------------------------------- fused map
struct hb_map_in {
  const hb_mapp *mp;
  const int64_t *off;
  __device__ __forceinline__ uint64_t operator()(int k) const {
    return hb_vm_load(mp->ptr[k], mp->dtype[k], off[k]);
  }
};

template <typename TO>
__global__ void __launch_bounds__(128)
    hb_fused_map_kernel(const __grid_constant__ hb_vm_prog P, const __grid_constant__ hb_mapp mp, int64_t n,
                        TO *__restrict__ out) {
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t lin = gid; lin < n; lin += gs) {
    int64_t off[HB_MAP_MAXIN];
#pragma unroll
    for (int k = 0; k < HB_MAP_MAXIN; k++) off[k] = 0;
    int64_t rem = lin;
#pragma unroll 1
    for (int d = mp.rank - 1; d >= 0; d--) {
      int64_t q = rem / mp.shape[d], r = rem - q * mp.shape[d];
#pragma unroll
      for (int k = 0; k < HB_MAP_MAXIN; k++)
        if (k < mp.n_in) off[k] += r * mp.stride[k][d];
      rem = q;
    }
    uint64_t outs[HB_VM_MAX_OUT];
    hb_map_in in{&mp, off};
    int64_t novars[1] = {lin};
    hb_vm_run(P, novars, in, outs);
    if constexpr (std::is_same<TO, double>::value) out[lin] = hb_bits_f(outs[0]);
    else if constexpr (std::is_same<TO, float>::value) out[lin] = (float)hb_bits_f(outs[0]);
    else out[lin] = (TO)(int64_t)outs[0];
  }
}

extern "C" int hb_fused_map(const hb_ixprog *prog, int n_in, hb_array *const *ins, hb_dtype dt,
                            hb_array **out) {
  hb_prof_scope prof_("fused_map");
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(prog && out && ins, "null argument");
  HB_REQUIRE(n_in >= 1 && n_in <= HB_MAP_MAXIN, "fused map: %d inputs (max %d)", n_in, HB_MAP_MAXIN);
  HB_REQUIRE(prog->vm.n_out == 1, "fused map: program must have exactly one output");
  for (int i = 1; i < n_in; i++)
    HB_REQUIRE(hb_same_shape(ins[0], ins[i]), "fused map: input %d has shape %s, expected %s", i,
               hb_shape_str(ins[i]).c_str(), hb_shape_str(ins[0]).c_str());
  if (hb_lazy_on()) {
    hb_lrec r(HB_L_GENERIC, "fused_map");
    for (int i = 0; i < n_in; i++) r.in(ins[i]);
    return r.prog(prog).out(out, dt, ins[0]->rank, ins[0]->shape).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_fused_map(n.prog, (int)n.in.size(), in, n.out[0]->dt, &o[0]); });
  }
  hb_array *o;
  HB_TRY(hb_new_array(dt, ins[0]->rank, ins[0]->shape, &o));
  int64_t n = hb_numel(o);
  if (n > 0) {
    hb_mapp mp;
    memset(&mp, 0, sizeof mp);
    mp.n_in = n_in;
    mp.rank = ins[0]->rank;
    int64_t str[HB_MAP_MAXIN][HB_R];
    for (int d = 0; d < mp.rank; d++) mp.shape[d] = ins[0]->shape[d];
    for (int k = 0; k < n_in; k++) {
      for (int d = 0; d < mp.rank; d++) str[k][d] = ins[k]->strides[d];
      mp.ptr[k] = hb_data(ins[k]);
      mp.dtype[k] = ins[k]->dt;
    }
    hb_collapse(&mp.rank, mp.shape, n_in, str);
    for (int k = 0; k < n_in; k++)
      for (int d = 0; d < mp.rank; d++) mp.stride[k][d] = str[k][d];
    // generated kernel: shapes / strides / dtypes baked in; 16-byte accesses when every input is
    // contiguous or a broadcast scalar and the pointers are aligned
    int es_max = (int)hb_dtype_size(dt);
    for (int k = 0; k < n_in; k++) es_max = std::max(es_max, (int)hb_dtype_size(ins[k]->dt));
    int vec = 16 / es_max;
    if (mp.rank != 1 || n % vec) vec = 1;
    for (int k = 0; k < n_in && vec > 1; k++) {
      if (mp.stride[k][0] != 0 && mp.stride[k][0] != 1) vec = 1;
      else if (mp.stride[k][0] == 1 && ((uintptr_t)mp.ptr[k] % (hb_dtype_size(ins[k]->dt) * vec))) vec = 1;
    }
    if ((uintptr_t)hb_data(o) % (hb_dtype_size(dt) * (size_t)vec)) vec = 1;
    std::string key;
    int hdr[5] = {3, mp.rank, n_in, (int)dt, vec};
    key_add(key, hdr, sizeof hdr);
    key_add(key, mp.shape, sizeof(int64_t) * mp.rank);
    for (int k = 0; k < n_in; k++) { key_add(key, mp.stride[k], sizeof(int64_t) * mp.rank); key_add(key, &mp.dtype[k], sizeof(int)); }
    CUfunction gen = hb_generated(prog, key, [&] { return hb_jit_map_source(prog->vm, prog->n_tensors, mp, dt, n, vec); });
    if (gen) {
      void *op = hb_data(o);
      void *args[1 + HB_MAP_MAXIN + HB_VM_MAX_TENSORS];
      const void *tptr[HB_VM_MAX_TENSORS];
      int na = 0;
      args[na++] = &op;
      for (int k = 0; k < n_in; k++) args[na++] = &mp.ptr[k];
      for (int t = 0; t < prog->n_tensors; t++) { tptr[t] = prog->vm.tensors[t].ptr; args[na++] = &tptr[t]; }
      int st = hb_jit_launch(gen, (unsigned)hb_blocks_for(n / vec, 256, g_hb.sm_count * 16), 256, args);
      if (st != HB_OK) { hb_release(o); return st; }
    } else {
      int blocks = hb_blocks_for(n, 128, g_hb.sm_count * 32);
      HB_DISPATCH_ALL(dt, T, (hb_fused_map_kernel<T><<<blocks, 128, 0, g_hb.stream>>>(prog->vm, mp, n, (T *)hb_data(o))));
    }
    HB_LAUNCHED();
  }
  *out = o;
  return HB_OK;
}

// scan.cu -- tmapAccumL specialisation for `sfold (*) 1 x` and its reverse
// derivative (LongProdBench: bench/common/BenchProdTools.hs:154-182;
// Ops.hs:188-210; OpsConcrete.hs:990-1019; the transposition rule is
// DeltaMapAccumL, DeltaEval.hs:351-389).
//
// The reference runs n sequential host iterations; here the fold is recognised
// as an associative scan:
//   acc_i = prod_{j<i} x_j        (exclusive prefix product, = the fold states)
//   cot_i = dt * prod_{j>i} x_j   (exclusive suffix product, = the cotangent
//                                  flowing backwards through the steps)
//   grad_i = cot_i * acc_i        (no division: exact also when some x_j = 0)
// Four kernels: tile products -> products of groups of 1024 tiles -> two-level scan of the tile
// products -> tile rescans producing the gradient; in-tile scans are ordered warp-shuffle scans.  HBM traffic 24 bytes per element (two
// reads, one write) against the 16-byte minimum.
#include "common.cuh"
#include "lazy.cuh"
#include "iter.cuh"

#define PT 2048  // elements per tile (256 threads x 8)

// 8 consecutive elements of this thread: two 128-bit loads per 4 doubles when
// the vector is contiguous and aligned (VEC), else strided scalar loads
template <typename T, bool VEC>
__device__ __forceinline__ void hb_load8(const T *__restrict__ x, int64_t stride, int64_t base, int64_t n, T (&v)[8]) {
  if (VEC && base + 8 <= n) {
    typedef hb_vec<T, 16 / sizeof(T)> V;
    constexpr int PER = 16 / sizeof(T), NV = 8 / PER;
    const V *p = reinterpret_cast<const V *>(x + base);
#pragma unroll
    for (int k = 0; k < NV; k++) {
      V t = p[k];
#pragma unroll
      for (int e = 0; e < PER; e++) v[k * PER + e] = t.v[e];
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = base + i < n ? x[(base + i) * stride] : T(1);
  }
}

// ordered (left-to-right) inclusive product scan across the block's 256
// thread values; returns the exclusive prefix of this thread.  `rev` scans
// from the right (exclusive suffix).  Fixed association -> deterministic.
template <typename T>
__device__ __forceinline__ T hb_block_excl_prod(T p, bool rev, T *smem /* 8 */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T incl = p;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T u = rev ? __shfl_down_sync(0xffffffffu, incl, o) : __shfl_up_sync(0xffffffffu, incl, o);
    bool ok = rev ? (lane + o < 32) : (lane >= o);
    if (ok) incl = rev ? incl * u : u * incl;
  }
  T excl = rev ? __shfl_down_sync(0xffffffffu, incl, 1) : __shfl_up_sync(0xffffffffu, incl, 1);
  if (rev ? lane == 31 : lane == 0) excl = T(1);
  __syncthreads();  // smem may still be read by a previous call
  if (rev ? lane == 0 : lane == 31) smem[w] = incl;
  __syncthreads();
  T wp = T(1);  // product of the warps before (after) this one, in order
  if (!rev) { for (int i = 0; i < w; i++) wp = wp * smem[i]; return wp * excl; }
  for (int i = 7; i > w; i--) wp = smem[i] * wp;
  return excl * wp;
}

// both directions at once (exclusive prefix and exclusive suffix of this thread over the block's 256 values) behind
// ONE pair of barriers; same association per direction as hb_block_excl_prod -> bit-identical results
template <typename T>
__device__ __forceinline__ void hb_block_excl_prod2(T p, T *smem /* 16 */, T &pre, T &suf) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T up = p, dn = p;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T u = __shfl_up_sync(0xffffffffu, up, o);
    T d = __shfl_down_sync(0xffffffffu, dn, o);
    if (lane >= o) up = u * up;
    if (lane + o < 32) dn = dn * d;
  }
  T eu = __shfl_up_sync(0xffffffffu, up, 1), ed = __shfl_down_sync(0xffffffffu, dn, 1);
  if (lane == 0) eu = T(1);
  if (lane == 31) ed = T(1);
  __syncthreads();  // smem may still be read by a previous call
  if (lane == 31) smem[w] = up;
  if (lane == 0) smem[8 + w] = dn;
  __syncthreads();
  T wp = T(1);
  for (int i = 0; i < w; i++) wp = wp * smem[i];
  pre = wp * eu;
  T ws = T(1);
  for (int i = 7; i > w; i--) ws = smem[8 + i] * ws;
  suf = ed * ws;
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(256)
    hb_prod_tiles_kernel(const T *__restrict__ x, int64_t stride, int64_t n, T *__restrict__ tileprod) {
  __shared__ T smem[8];
  int64_t base = (int64_t)blockIdx.x * PT + threadIdx.x * 8;
  T v[8];
  hb_load8<T, VEC>(x, stride, base, n, v);
  T p = T(1);
#pragma unroll
  for (int i = 0; i < 8; i++) p *= v[i];
  T before = hb_block_excl_prod<T>(p, false, smem);
  if (threadIdx.x == 255) tileprod[blockIdx.x] = before * p;
}

// ---- scan of the tile products, two levels: groups of 1024 tiles (one block each, coalesced loads)
// and the group totals (<= 1024 of them, rescanned by every block from shared memory).  The former
// single-block version walked 48 strided, dependent global loads per thread at n = 1e8 and cost a
// third of the whole gradient (0.21 of 0.66 ms).
constexpr int GT = 1024;  // tiles per group

// ordered exclusive product scan over the 1024 threads of a block (left to right, or right to left);
// *total (optional) receives the product of all 1024 values.  Fixed association -> deterministic.
template <typename T>
__device__ __forceinline__ T hb_block1024_excl_prod(T v, bool rev, T *smem /* 32 */, T *total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T u = rev ? __shfl_down_sync(0xffffffffu, incl, o) : __shfl_up_sync(0xffffffffu, incl, o);
    bool ok = rev ? (lane + o < 32) : (lane >= o);
    if (ok) incl = rev ? incl * u : u * incl;
  }
  T excl = rev ? __shfl_down_sync(0xffffffffu, incl, 1) : __shfl_up_sync(0xffffffffu, incl, 1);
  if (rev ? lane == 31 : lane == 0) excl = T(1);
  __syncthreads();  // smem may still be read by a previous call
  if (rev ? lane == 0 : lane == 31) smem[w] = incl;
  __syncthreads();
  T wp = T(1);      // product of the warps before (after) this one, in order
  if (!rev) {
    for (int i = 0; i < w; i++) wp = wp * smem[i];
    if (total && threadIdx.x == GT - 1) *total = wp * incl;
    return wp * excl;
  }
  for (int i = 31; i > w; i--) wp = smem[i] * wp;
  if (total && threadIdx.x == 0) *total = incl * wp;
  return excl * wp;
}

template <typename T>
__global__ void __launch_bounds__(GT) hb_prod_groups_kernel(const T *__restrict__ tp, int64_t nt, T *__restrict__ gp) {
  __shared__ T smem[32];
  int64_t i = (int64_t)blockIdx.x * GT + threadIdx.x;
  T v = i < nt ? tp[i] : T(1);
  hb_block1024_excl_prod<T>(v, false, smem, gp + blockIdx.x);
}

// pre[i] = prod_{j<i} tp[j],  suf[i] = dt * prod_{j>i} tp[j],  *total = prod of all (block 0 writes it)
template <typename T>
__global__ void __launch_bounds__(GT)
    hb_prod_tilescan_kernel(const T *__restrict__ tp, int64_t nt, const T *__restrict__ gp, int ng, T dt,
                            T *__restrict__ pre, T *__restrict__ suf, T *__restrict__ total) {
  __shared__ T smem[32];
  __shared__ T gsm[GT];
  __shared__ T edge[2];  // product of the groups before / (dt *) after this one
  const int g = blockIdx.x;
  for (int i = threadIdx.x; i < ng; i += GT) gsm[i] = gp[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    T r = T(1);
    for (int h = 0; h < g; h++) r *= gsm[h];
    edge[0] = r;
    if (g == 0) {
      T all = T(1);
      for (int h = 0; h < ng; h++) all *= gsm[h];
      *total = all;
    }
  } else if (threadIdx.x == 32) {
    T r = dt;
    for (int h = ng - 1; h > g; h--) r *= gsm[h];
    edge[1] = r;
  }
  int64_t i = (int64_t)g * GT + threadIdx.x;
  T v = i < nt ? tp[i] : T(1);
  T ef = hb_block1024_excl_prod<T>(v, false, smem, nullptr);
  T er = hb_block1024_excl_prod<T>(v, true, smem, nullptr);
  // (the scans' barriers have made edge[] visible)
  if (i < nt) {
    pre[i] = edge[0] * ef;
    suf[i] = er * edge[1];
  }
}

// per tile: grad[i] = (suf_tile * prod_{j>i in tile} x_j) * (pre_tile * prod_{j<i in tile} x_j)
template <typename T, bool VEC>
__global__ void __launch_bounds__(256)
    hb_prod_grad_kernel(const T *__restrict__ x, int64_t stride, int64_t n, const T *__restrict__ pre,
                        const T *__restrict__ suf, T *__restrict__ grad) {
  __shared__ T smem[8];
  int64_t base = (int64_t)blockIdx.x * PT + threadIdx.x * 8;
  T v[8];
  hb_load8<T, VEC>(x, stride, base, n, v);
  T p = T(1);
#pragma unroll
  for (int i = 0; i < 8; i++) p *= v[i];
  T a = pre[blockIdx.x] * hb_block_excl_prod<T>(p, false, smem);  // everything before this thread's 8
  T c = hb_block_excl_prod<T>(p, true, smem) * suf[blockIdx.x];   // dt * everything after them
  T out[8];
  T r = c;
#pragma unroll
  for (int i = 7; i >= 0; i--) { out[i] = r; r *= v[i]; }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    out[i] = out[i] * a;
    a *= v[i];
  }
  if (VEC && base + 8 <= n) {
    typedef hb_vec<T, 16 / sizeof(T)> V;
    constexpr int PER = 16 / sizeof(T), NV = 8 / PER;
    V *g = reinterpret_cast<V *>(grad + base);
#pragma unroll
    for (int k = 0; k < NV; k++) {
      V t;
#pragma unroll
      for (int e = 0; e < PER; e++) t.v[e] = out[k * PER + e];
      g[k] = t;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++)
      if (base + i < n) grad[base + i] = out[i];
  }
}

// ------------------------------------------------------------------ one persistent kernel
// Round 2: the four launches above (tile products, group products, scan of the tile products, tile rescans) become
// ONE cooperative persistent kernel.  Every CTA owns a contiguous range of tiles:
//   phase A  products of its tiles (first read of x);
//   grid barrier (cooperative launch: all CTAs are co-resident);
//   phase S  the product of all tiles before / after its range, by an ordered block reduction over the tile products
//            (a few hundred KB out of L2), and the running prefix of its own tiles in shared memory;
//   phase B  its tiles again, LAST TILE FIRST (the end of the range is what phase A read most recently, so for
//            n <= ~1.5e7 the whole second read of x is an L2 hit), exclusive in-tile scans, gradient written once.
// A prefix AND a suffix are needed, so every tile's gradient depends on every other tile's product: x has to be
// read twice unless it fits on chip -- 24 n bytes when n*8 >> L2 (126 MB), 16 n bytes of HBM traffic below that.
// Same ordered in-tile scans as before -> deterministic; no division (exact when some x_j = 0).
// bar[0] counts arrivals, bar[1] departures: the last CTA to LEAVE the spin puts both back to zero, so the two words
// (allocated once, zero between launches) need no memset node in front of every launch
__device__ __forceinline__ void hb_grid_barrier(unsigned *bar, unsigned expected) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (v < expected) __nanosleep(32);
    } while (v < expected);
    if (atomicAdd(bar + 1, 1u) == expected - 1) {   // everybody has seen the full count
      bar[1] = 0;
      __threadfence();
      bar[0] = 0;
    }
  }
  __syncthreads();
}

// ordered product of tp[lo, hi): every thread multiplies a contiguous sub-range left to right, the 256 partial
// products are combined left to right (warp shuffles, then the 8 warp totals); every thread returns the total
template <typename T>
__device__ __forceinline__ T hb_block_range_prod(const T *__restrict__ tp, int64_t lo, int64_t hi, T *smem /* 9 */) {
  const int64_t len = hi - lo;
  const int64_t per = (len + 255) / 256;
  int64_t a = lo + (int64_t)threadIdx.x * per, b = a + per < hi ? a + per : hi;
  T p = T(1);
  for (int64_t i = a; i < b; i++) p = p * __ldcg(tp + i);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T incl = p;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl = u * incl;
  }
  __syncthreads();
  if (lane == 31) smem[w] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    T r = T(1);
    for (int i = 0; i < 8; i++) r = r * smem[i];
    smem[8] = r;
  }
  __syncthreads();
  return smem[8];
}

// Four CTAs per SM (64 registers, no spills; left alone the compiler takes 73 -> three CTAs: 0.428 ms at 1e8 against
// 0.411).  Forcing five or six (48 / 40 registers) spills and loses far more than the occupancy gains (0.54 / 0.67 ms).
// Measured without effect in this configuration: evict-first loads / streaming stores in phase B (0.412-0.417 ms), one
// CTA barrier per tile instead of two (alternating scan scratch: 0.413 ms).
template <typename T, bool VEC>
__global__ void __launch_bounds__(256, 4)
    hb_prod_fused_kernel(const T *__restrict__ x, int64_t stride, int64_t n, int64_t nt, T dt, T *__restrict__ tp,
                         T *__restrict__ cp, unsigned *__restrict__ bar, T *__restrict__ total, T *__restrict__ grad,
                         int own_cap) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  T *own_pre = reinterpret_cast<T *>(dyn_smem);   // [own_cap]: product of everything before own tile k
  T *own_tp = own_pre + own_cap;                  // [own_cap]: products of the own tiles (never re-read from global)
  __shared__ T smem[16];
  __shared__ __align__(16) double stage[8 * 32 * 10];   // per warp: 32 lanes x (8 results + 2 pad)
  const int64_t G = gridDim.x, b = blockIdx.x;
  const int64_t t0 = b * nt / G, t1 = (b + 1) * nt / G;
  // ---- phase A
  T own = T(1);   // thread 255: product of this CTA's tiles, left to right
  // the loads of tile t + 1 are issued before the scan (two CTA barriers) of tile t: twice the bytes in flight per CTA
  T v[8], nx[8];
  if (t0 < t1) hb_load8<T, VEC>(x, stride, t0 * PT + threadIdx.x * 8, n, nx);
  for (int64_t t = t0; t < t1; t++) {
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = nx[i];
    if (t + 1 < t1) hb_load8<T, VEC>(x, stride, (t + 1) * PT + threadIdx.x * 8, n, nx);
    T p = T(1);
#pragma unroll
    for (int i = 0; i < 8; i++) p *= v[i];
    T before = hb_block_excl_prod<T>(p, false, smem);
    if (threadIdx.x == 255) {
      T tile = before * p;
      own_tp[t - t0] = tile;
      own = own * tile;
    }
  }
  if (threadIdx.x == 255) cp[b] = own;
  hb_grid_barrier(bar, (unsigned)G);
  // ---- phase S: G per-CTA products (a few KB), not nt tile products, are what every CTA re-reads
  T e0 = hb_block_range_prod<T>(cp, 0, b, smem);
  T e1 = hb_block_range_prod<T>(cp, b + 1, G, smem);
  if (b == 0 && threadIdx.x == 0) *total = __ldcg(cp) * e1;   // prod x = (own range of CTA 0) * (everything after it)
  if (!grad) return;
  e1 = e1 * dt;
  if (threadIdx.x < 32) {   // exclusive prefix products of the own tiles, by one warp: a run per lane + a warp scan
    const int cnt = (int)(t1 - t0), per = (cnt + 31) / 32, lane = threadIdx.x;
    const int lo = lane * per < cnt ? lane * per : cnt, hi = lo + per < cnt ? lo + per : cnt;
    T run = T(1);
    for (int k = lo; k < hi; k++) run = run * own_tp[k];
    T incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      T u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl = u * incl;
    }
    T r = __shfl_up_sync(0xffffffffu, incl, 1);
    r = lane == 0 ? e0 : e0 * r;
    for (int k = lo; k < hi; k++) {
      own_pre[k] = r;
      r = r * own_tp[k];
    }
  }
  __syncthreads();
  // ---- phase B, last tile first
  T suf = e1;
  if (t0 < t1) hb_load8<T, VEC>(x, stride, (t1 - 1) * PT + threadIdx.x * 8, n, nx);
  for (int64_t t = t1 - 1; t >= t0; t--) {
    int64_t base = t * PT + threadIdx.x * 8;
    const int64_t wbase = t * PT + (threadIdx.x >> 5) * 256;   // this warp's 256 consecutive elements
    double *row = stage + (threadIdx.x >> 5) * (32 * 10);
    const bool coal = VEC && sizeof(T) == 8 && wbase + 256 <= n;
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = nx[i];
    if (t > t0) hb_load8<T, VEC>(x, stride, (t - 1) * PT + threadIdx.x * 8, n, nx);
    T p = T(1);
#pragma unroll
    for (int i = 0; i < 8; i++) p *= v[i];
    T a, c;
    hb_block_excl_prod2<T>(p, smem, a, c);
    a = own_pre[t - t0] * a;
    c = c * suf;
    T out[8];
    T r = c;
#pragma unroll
    for (int i = 7; i >= 0; i--) { out[i] = r; r *= v[i]; }
#pragma unroll
    for (int i = 0; i < 8; i++) {
      out[i] = out[i] * a;
      a *= v[i];
    }
    if (coal) {
      // a thread holds 8 CONSECUTIVE results (64 B): stored directly, one store instruction of the warp touches 32
      // separate half-sectors.  Through the warp's own shared-memory row (padded: a lane's 64 B start 80 B apart, so
      // the four 16-byte stores of a quarter-warp fall into different banks) the warp stores 512 contiguous bytes
      // per instruction instead.
      const int lane = threadIdx.x & 31;
#pragma unroll
      for (int k = 0; k < 4; k++)
        *reinterpret_cast<double2 *>(row + lane * 10 + 2 * k) = make_double2((double)out[2 * k], (double)out[2 * k + 1]);
      __syncwarp();
      double2 *g = reinterpret_cast<double2 *>(grad + wbase);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int e = j * 64 + 2 * lane;            // element of the warp's 256
        g[j * 32 + lane] = *reinterpret_cast<const double2 *>(row + (e >> 3) * 10 + (e & 7));
      }
      __syncwarp();
    } else if (VEC && base + 8 <= n) {
      typedef hb_vec<T, 16 / sizeof(T)> V;
      constexpr int PER = 16 / sizeof(T), NV = 8 / PER;
      V *g = reinterpret_cast<V *>(grad + base);
#pragma unroll
      for (int k = 0; k < NV; k++) {
        V tv;
#pragma unroll
        for (int e = 0; e < PER; e++) tv.v[e] = out[k * PER + e];
        g[k] = tv;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++)
        if (base + i < n) grad[base + i] = out[i];
    }
    suf = own_tp[t - t0] * suf;   // tile t joins the suffix of the tiles before it
  }
}

// 1 when the fused kernel took the call
template <typename T>
static int prod_launch_fused(const hb_array *x, int64_t n, int64_t nt, double dt, void *scr, hb_array *prod,
                             hb_array *grad, int *done) {
  *done = 0;
  if (getenv("HB_PROD_UNFUSED")) return HB_OK;   // measurement switch: the four-launch pipeline of round 1
  const T *xp = (const T *)hb_data(x);
  const bool vec = x->strides[0] == 1 && ((uintptr_t)xp & 15) == 0 && (!grad || ((uintptr_t)hb_data(grad) & 15) == 0);
  void (*kern)(const T *, int64_t, int64_t, int64_t, T, T *, T *, unsigned *, T *, T *, int) =
      vec ? hb_prod_fused_kernel<T, true> : hb_prod_fused_kernel<T, false>;
  int occ = 0;
  const int own_guess = 64;
  HB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 2 * own_guess * sizeof(T)));
  if (occ < 1) return HB_OK;
  if (occ > 8) occ = 8;
  int64_t G = (int64_t)occ * g_hb.sm_count;
  if (G > nt) G = nt;
  int64_t own = (nt + G - 1) / G + 1;
  size_t dyn = 2 * (size_t)own * sizeof(T);   // own_pre + own_tp
  if (dyn > 48 * 1024) return HB_OK;   // absurdly long vectors: the multi-launch pipeline takes them
  HB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, dyn));
  if ((int64_t)occ * g_hb.sm_count < G) {
    G = (int64_t)occ * g_hb.sm_count;
    if (G < 1) return HB_OK;
    own = (nt + G - 1) / G + 1;
    if (2 * (size_t)own * sizeof(T) > dyn) return HB_OK;
  }
  T *tp = (T *)scr, *cp = tp + nt;   // G <= nt per-CTA products in the round-1 pipeline's `pre` region
  static unsigned *bar = nullptr;   // arrival / departure counters of the grid barrier: zero between launches
  if (!bar) {
    HB_REQUIRE(!g_hb.capturing, "prod fold: first use inside a graph capture");
    HB_CUDA(cudaMalloc(&bar, 2 * sizeof(unsigned)));
    HB_CUDA(cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned), g_hb.stream));
  }
  int64_t stride = x->strides[0];
  T dtt = (T)dt;
  T *tot = (T *)hb_data(prod), *gp = grad ? (T *)hb_data(grad) : nullptr;
  int own_cap = (int)own;
  void *args[] = {(void *)&xp, (void *)&stride, (void *)&n, (void *)&nt, (void *)&dtt, (void *)&tp, (void *)&cp, (void *)&bar,
                  (void *)&tot, (void *)&gp, (void *)&own_cap};
  // cooperative launch = the runtime guarantees (or refuses) co-residency of all G CTAs: the barrier cannot hang
  cudaError_t e = cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)G), dim3(256), args, dyn, g_hb.stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return HB_OK;   // not cooperative-capable in this context (e.g. some capture modes): fall back
  }
  HB_LAUNCHED();
  *done = 1;
  return HB_OK;
}

template <typename T>
static int prod_launch(const hb_array *x, int64_t n, int64_t nt, double dt, void *scr, hb_array *prod, hb_array *grad) {
  {
    int done = 0;
    HB_TRY(prod_launch_fused<T>(x, n, nt, dt, scr, prod, grad, &done));
    if (done) return HB_OK;
  }
  T *tp = (T *)scr, *pre = tp + nt, *suf = pre + nt;
  const T *xp = (const T *)hb_data(x);
  bool vec = x->strides[0] == 1 && ((uintptr_t)xp & 15) == 0 && (!grad || ((uintptr_t)hb_data(grad) & 15) == 0);
  if (vec) hb_prod_tiles_kernel<T, true><<<(unsigned)nt, 256, 0, g_hb.stream>>>(xp, 1, n, tp);
  else hb_prod_tiles_kernel<T, false><<<(unsigned)nt, 256, 0, g_hb.stream>>>(xp, x->strides[0], n, tp);
  HB_LAUNCHED();
  const int ng = (int)((nt + GT - 1) / GT);
  T *gp = suf + nt;
  hb_prod_groups_kernel<T><<<ng, GT, 0, g_hb.stream>>>(tp, nt, gp);
  HB_LAUNCHED();
  hb_prod_tilescan_kernel<T><<<ng, GT, 0, g_hb.stream>>>(tp, nt, gp, ng, (T)dt, pre, suf, (T *)hb_data(prod));
  HB_LAUNCHED();
  if (grad) {
    if (vec) hb_prod_grad_kernel<T, true><<<(unsigned)nt, 256, 0, g_hb.stream>>>(xp, 1, n, pre, suf, (T *)hb_data(grad));
    else hb_prod_grad_kernel<T, false><<<(unsigned)nt, 256, 0, g_hb.stream>>>(xp, x->strides[0], n, pre, suf, (T *)hb_data(grad));
    HB_LAUNCHED();
  }
  return HB_OK;
}

extern "C" int hb_prod_fold_grad(const hb_array *x, double dt, hb_array **prod_out, hb_array **grad_out) {
  hb_prof_scope prof_("prod_fold_grad", x);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(x && prod_out, "null argument");
  HB_REQUIRE(x->rank == 1 && hb_is_float(x->dt), "prod fold: rank-1 floating vector required");
  int64_t n = x->shape[0];
  HB_REQUIRE(n <= (int64_t)PT * GT * GT, "prod fold: vector too long (%lld elements)", (long long)n);
  if (hb_lazy_on()) {
    int64_t one_ = 1;
    return hb_lrec(HB_L_GENERIC, "prod_fold_grad").in(x).farg(0, dt).out(prod_out, x->dt, 0, &one_).out(grad_out, x->dt, 1, x->shape)
        .commit([=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_prod_fold_grad(in[0], n.fa[0], &o[0], n.out[1] ? &o[1] : nullptr); });
  }
  int64_t one = 1;
  hb_array *prod, *grad = nullptr;
  HB_TRY(hb_new_array(x->dt, 0, &one, &prod));
  if (grad_out) {
    int st = hb_new_array(x->dt, 1, &n, &grad);
    if (st != HB_OK) { hb_release(prod); return st; }
  }
  if (n == 0) {
    double o64 = 1.0; float o32 = 1.f;
    int st = hb_fill_scalar(hb_data(prod), x->dt, x->dt == HB_F64 ? (void *)&o64 : (void *)&o32);
    if (st != HB_OK) { hb_release(prod); hb_release(grad); return st; }
    *prod_out = prod;
    if (grad_out) *grad_out = grad;
    return HB_OK;
  }
  int64_t nt = (n + PT - 1) / PT;
  size_t es = hb_dtype_size(x->dt);
  void *scr;
  int st = hb_scratch((size_t)(nt * 3 + GT) * es + 128, &scr);
  if (st != HB_OK) { hb_release(prod); hb_release(grad); return st; }
  st = x->dt == HB_F64 ? prod_launch<double>(x, n, nt, dt, scr, prod, grad)
                       : prod_launch<float>(x, n, nt, dt, scr, prod, grad);
  if (st != HB_OK) { hb_release(prod); hb_release(grad); return st; }
  *prod_out = prod;
  if (grad_out) *grad_out = grad;
  return HB_OK;
}

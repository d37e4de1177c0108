// comm.cu -- data-parallel gradient exchange: one process per
// GPU, ONE all-reduce of the flattened gradient bucket per training step over
// NVLink 5 / NVSwitch, followed by the 1/world scaling fused into one pass.
// The reference has no distributed code at all; this is the
// new batch-sharded training loop demanded by BASELINE.json.
//
// NCCL is resolved at run time (dlopen) so that libhordeb200.so carries no
// link-time dependency and shares the libnccl already loaded by the host
// process, if any.
#include <dlfcn.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "comm.cuh"
#include "lazy.cuh"

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { hbNcclFloat32 = 7, hbNcclFloat64 = 8, hbNcclInt64 = 4, hbNcclInt32 = 2, hbNcclSum = 0 };

static struct {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
} g_nccl;

static int load_nccl() {
  if (g_nccl.lib) return HB_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] && !g_nccl.lib; i++) g_nccl.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!g_nccl.lib) HB_FAIL(HB_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                     \
  *(void **)(&g_nccl.field) = dlsym(g_nccl.lib, name);                       \
  if (!g_nccl.field) HB_FAIL(HB_ERR_NCCL, "libnccl: missing symbol %s", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(AllReduce, "ncclAllReduce")
  SYM(AllGather, "ncclAllGather")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return HB_OK;
}

#define HB_NCCL(call)                                                              \
  do {                                                                             \
    ncclResult_t r_ = (call);                                                      \
    if (r_ != 0) HB_FAIL(HB_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); \
  } while (0)

extern "C" int hb_comm_unique_id(void *id_bytes) {
  HB_REQUIRE(id_bytes, "null argument");
  HB_TRY(load_nccl());
  ncclUniqueId id;
  HB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id_bytes, &id, HB_UNIQUE_ID_BYTES);
  return HB_OK;
}

static int fcomm_setup(size_t cap_bytes);

extern "C" int hb_comm_init(int rank, int world, const void *id_bytes) {
  HB_CHECK_INIT();
  HB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "comm: rank %d of %d", rank, world);
  g_nccl.rank = rank;
  g_nccl.world = world;
  if (world == 1) return fcomm_setup(0);
  HB_REQUIRE(id_bytes, "null unique id");
  HB_TRY(load_nccl());
  ncclUniqueId id;
  memcpy(&id, id_bytes, HB_UNIQUE_ID_BYTES);
  HB_NCCL(g_nccl.CommInitRank(&g_nccl.comm, world, id, rank));
  // peer-mapped exchange buffers of the fused all-reduce + optimizer kernel: 2 parities x 8 source slots of 16 MB
  // (every model of the reference's examples fits -- MnistFcnnRanked2 at 1500|500 has 1.93 M parameters = 15.4 MB in
  // Double); 256 MB of the 180 GB
  return fcomm_setup((size_t)16 << 20);
}

template <typename T>
__global__ void hb_scale_kernel(T *p, T s, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t st = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) p[i] *= s;
}

extern "C" int hb_allreduce_sum(hb_array *bucket, double scale) {
  hb_prof_scope prof_("allreduce_sum", bucket);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(bucket, "null argument");
  HB_REQUIRE(hb_contig(bucket), "allreduce: bucket must be contiguous");
  HB_REQUIRE(hb_is_float(bucket->dt), "allreduce: floating bucket required");
  if (hb_lazy_on())
    return hb_lrec(HB_L_EFFECT, "allreduce_sum").in(bucket).farg(0, scale).commit(
        [=](hb_lnode &n, hb_array *const *in, hb_array **o) { return hb_allreduce_sum(in[0], n.fa[0]); });
  int64_t n = hb_numel(bucket);
  if (n == 0) return HB_OK;
  if (g_nccl.world > 1) {
    HB_REQUIRE(g_nccl.comm, "allreduce: hb_comm_init has not been called");
    HB_NCCL(g_nccl.AllReduce(hb_data(bucket), hb_data(bucket), (size_t)n,
                             bucket->dt == HB_F64 ? hbNcclFloat64 : hbNcclFloat32, hbNcclSum, g_nccl.comm,
                             g_hb.stream));
    if (g_hb.profiling) hb_prof_mark("ncclAllReduce", 0);
  }
  if (scale != 1.0) {
    int blocks = hb_blocks_for(n, 256, g_hb.sm_count * 8);
    if (bucket->dt == HB_F64) hb_scale_kernel<double><<<blocks, 256, 0, g_hb.stream>>>((double *)hb_data(bucket), scale, n);
    else hb_scale_kernel<float><<<blocks, 256, 0, g_hb.stream>>>((float *)hb_data(bucket), (float)scale, n);
    HB_LAUNCHED();
  }
  return HB_OK;
}

// ------------------------------------------------- fused exchange + optimizer
// ONE kernel per training step does what ncclAllReduce + hb_scale_kernel + hb_adam_kernel did in three launches
// (65 us + 2 launches for a 0.46 MB bucket at 8 GPUs, the top entry of the strong-scaling step in round 1):
//   PUSH: every rank writes its chunk of the gradient bucket straight into a slot of EVERY peer's exchange buffer
//   (cudaIpc-mapped over NVLink 5 / NVSwitch: posted remote stores, every GPU reaches every peer at full
//   bandwidth), fences, and raises a per-chunk flag in the peer's memory; the receiver waits for its peers' flags
//   and then reads only LOCAL memory: it adds the G contributions in rank order 0..G-1 -- the same order on every
//   rank, so the reduced bucket (and therefore the replicated parameters) are bit-identical across ranks and
//   across runs, independently of any library's algorithm choice -- scales by 1/G (or 1) and applies the Adam update
//   (OptimizerTools.hs:134-232) to its replica in the same pass.  No remote LOAD is on the critical path.
// Chunk c of rank A only ever waits for chunk c of the peers (no grid-wide barrier); slots are double-buffered
// by step parity, so one flag round per step is the only synchronisation: a rank reaches step e+2 (which overwrites
// the parity-e slots it owns in its peers' buffers) only after its step e+1 kernel saw every peer's e+1 flags, i.e.
// after every peer finished its step-e kernel and with it all reads of the parity-e slots.
// The step counter of Adam lives on the device (a CUDA graph replays the launch unchanged).
#define HB_COMM_MAXCHUNKS 256
struct hb_fcomm {
  int rank = 0, world = 1;
  char *buf[8] = {};                  // peer-mapped: [2 parities][8 source ranks][cap bytes]
  unsigned long long *flags[8] = {};  // peer-mapped: [8 source ranks][HB_COMM_MAXCHUNKS] step numbers
  unsigned long long *ctl = nullptr;  // local: [0] = steps completed, [1] = CTAs done in the running step
  size_t cap = 0;
};
// scalar exchange area of every rank's block (global-objective mode, comm.cuh): [HB_SX_BYTES] after the control words
static hb_fcomm g_fc;
static hb_scalar_xchg g_sx;          // peer-mapped scalar slots (world > 1)
static int g_global_objective = 0;
static void *g_fc_block = nullptr;
static void *g_fc_peer[8] = {};
static bool g_fc_ready = false;
static unsigned long long *g_fc_ctl_single = nullptr;   // world == 1: only the control words

__device__ __forceinline__ void hb_st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long hb_ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256)
    hb_allreduce_adam_kernel(hb_fcomm C, T *__restrict__ bucket, int64_t n, int64_t n_params, T *__restrict__ p,
                             T *__restrict__ m, T *__restrict__ v, long long *__restrict__ tcount, double alpha,
                             double beta1, double beta2, double epsilon, T scale) {
  const int c = blockIdx.x, nchunks = gridDim.x;
  const unsigned long long e = *(volatile unsigned long long *)&C.ctl[0] + 1;   // this step's number
  const long long t = *(volatile long long *)tcount + 1;                        // Adam's step count after increment
  const int64_t per = ((n + nchunks - 1) / nchunks + 1) & ~(int64_t)1;
  const int64_t lo = min((int64_t)c * per, n), hi = min(lo + per, n);
  const int par = (int)(e & 1);
  const size_t slot = ((size_t)par * 8 + C.rank) * C.cap;   // where THIS rank's data lives in every peer's buffer
  if (C.world > 1) {
    // push: posted stores into the peers' memory
    for (int q = 1; q < C.world; q++) {
      const int peer = (C.rank + q) % C.world;
      T *dst = (T *)(C.buf[peer] + slot);
      for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = bucket[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) __threadfence_system();
    __syncthreads();
    if (threadIdx.x < C.world && threadIdx.x != C.rank)
      hb_st_release_sys(C.flags[threadIdx.x] + (size_t)C.rank * HB_COMM_MAXCHUNKS + c, e);
    if (threadIdx.x < C.world && threadIdx.x != C.rank) {
      const unsigned long long *f = C.flags[C.rank] + (size_t)threadIdx.x * HB_COMM_MAXCHUNKS + c;
      // a peer that never arrives (crashed rank) must not hang the GPU: give up after ~10 s and fault the launch
      const long long t0 = clock64();
      while (hb_ld_acquire_sys(f) < e) {
        __nanosleep(32);
        if (clock64() - t0 > 20000000000ll) __trap();
      }
    }
    __syncthreads();
  }
  const T one = (T)1;
  const double alphat_d = alpha * sqrt(1.0 - hb_powi_ghc(beta2, t)) / (1.0 - hb_powi_ghc(beta1, t));
  const T alphat = (T)alphat_d, b1 = (T)beta1, b2 = (T)beta2, eps = (T)epsilon;
  const char *mine = C.world > 1 ? C.buf[C.rank] + (size_t)par * 8 * C.cap : nullptr;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    T acc;
    if (C.world > 1) {
      acc = (T)0;
      for (int r = 0; r < C.world; r++)   // fixed rank order: bit-identical on every rank; L2 loads (peers wrote it)
        acc = acc + (r == C.rank ? bucket[i] : __ldcg((const T *)(mine + (size_t)r * C.cap) + i));
    } else {
      acc = bucket[i];
    }
    acc = acc * scale;
    bucket[i] = acc;
    if (i < n_params) {
      T mn = b1 * m[i] + (one - b1) * acc;
      T vn = b2 * v[i] + (one - b2) * (acc * acc);
      m[i] = mn;
      v[i] = vn;
      p[i] = p[i] - alphat * mn / (sqrt(vn) + eps);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&C.ctl[1], 1ull) == (unsigned long long)nchunks - 1) {   // last chunk of this rank: close the step
      C.ctl[1] = 0;
      *tcount = t;
      __threadfence();
      *(volatile unsigned long long *)&C.ctl[0] = e;
    }
  }
}

// Set up the peer-mapped buffers: collective over the communicator (called from hb_comm_init).
static int fcomm_setup(size_t cap_bytes) {
  g_fc = hb_fcomm();
  g_fc.rank = g_nccl.rank;
  g_fc.world = g_nccl.world;
  g_fc.cap = cap_bytes;
  const size_t flag_bytes = (size_t)8 * HB_COMM_MAXCHUNKS * sizeof(unsigned long long);
  const size_t ctl_bytes = 256;
  if (g_nccl.world == 1) {
    if (!g_fc_ctl_single) {
      HB_CUDA(cudaMalloc((void **)&g_fc_ctl_single, ctl_bytes));
      HB_CUDA(cudaMemset(g_fc_ctl_single, 0, ctl_bytes));
    }
    g_fc.ctl = g_fc_ctl_single;
    g_fc_ready = true;
    return HB_OK;
  }
  HB_REQUIRE(g_nccl.world <= 8, "fused exchange: at most 8 ranks of one NVSwitch domain (have %d)", g_nccl.world);
  const size_t total = flag_bytes + ctl_bytes + HB_SX_BYTES + 2 * 8 * cap_bytes;
  HB_CUDA(cudaMalloc(&g_fc_block, total));
  HB_CUDA(cudaMemset(g_fc_block, 0, total));
  // the block is zero before its handle leaves this process: a peer can only raise a flag in it after the exchange
  HB_CUDA(cudaDeviceSynchronize());
  // exchange the IPC handles through the communicator itself
  cudaIpcMemHandle_t mine;
  HB_CUDA(cudaIpcGetMemHandle(&mine, g_fc_block));
  char *dsend = nullptr, *drecv = nullptr;
  HB_CUDA(cudaMalloc((void **)&dsend, sizeof mine));
  HB_CUDA(cudaMalloc((void **)&drecv, sizeof mine * g_nccl.world));
  HB_CUDA(cudaMemcpy(dsend, &mine, sizeof mine, cudaMemcpyHostToDevice));
  HB_NCCL(g_nccl.AllGather(dsend, drecv, sizeof mine, /*ncclInt8*/ 0, g_nccl.comm, g_hb.stream));
  HB_CUDA(cudaStreamSynchronize(g_hb.stream));
  std::vector<cudaIpcMemHandle_t> all(g_nccl.world);
  HB_CUDA(cudaMemcpy(all.data(), drecv, sizeof mine * g_nccl.world, cudaMemcpyDeviceToHost));
  cudaFree(dsend);
  cudaFree(drecv);
  for (int r = 0; r < g_nccl.world; r++) {
    void *base = g_fc_block;
    if (r != g_nccl.rank) {
      HB_CUDA(cudaIpcOpenMemHandle(&base, all[r], cudaIpcMemLazyEnablePeerAccess));
      g_fc_peer[r] = base;
    }
    g_fc.flags[r] = (unsigned long long *)base;
    g_fc.buf[r] = (char *)base + flag_bytes + ctl_bytes + HB_SX_BYTES;
    char *sx = (char *)base + flag_bytes + ctl_bytes;
    g_sx.flag[r] = (unsigned long long *)sx;                       // [2 kinds][8 ranks]
    g_sx.val[r] = (double *)(sx + 2 * 8 * sizeof(unsigned long long));   // [2 parities][2 kinds][8 ranks]
  }
  g_sx.rank = g_nccl.rank;
  g_sx.world = g_nccl.world;
  g_sx.ctl = (unsigned long long *)((char *)g_fc_block + flag_bytes + ctl_bytes + HB_SX_BYTES - 64);
  g_fc.ctl = (unsigned long long *)((char *)g_fc_block + flag_bytes);
  g_fc_ready = true;
  return HB_OK;
}

static void fcomm_teardown() {
  for (int r = 0; r < 8; r++)
    if (g_fc_peer[r]) {
      cudaIpcCloseMemHandle(g_fc_peer[r]);
      g_fc_peer[r] = nullptr;
    }
  if (g_fc_block) cudaFree(g_fc_block);
  g_fc_block = nullptr;
  g_fc_ready = false;
}

extern "C" int hb_allreduce_adam_step(hb_array *bucket, int64_t n_params, hb_array *p, hb_array *m, hb_array *v,
                                      hb_array *t_counter, double alpha, double beta1, double beta2,
                                      double epsilon, double scale) {
  hb_prof_scope prof_("allreduce_adam_step", bucket);
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(bucket && p && m && v && t_counter, "null argument");
  HB_REQUIRE(hb_contig(bucket) && hb_contig(p) && hb_contig(m) && hb_contig(v), "allreduce_adam: contiguous arrays required");
  HB_REQUIRE(hb_is_float(bucket->dt) && p->dt == bucket->dt && m->dt == bucket->dt && v->dt == bucket->dt,
             "allreduce_adam: one floating dtype required");
  HB_REQUIRE(t_counter->dt == HB_I64 && hb_numel(t_counter) == 1, "allreduce_adam: the step counter is a 0-d I64 array");
  int64_t n = hb_numel(bucket);
  HB_REQUIRE(n_params >= 0 && n_params <= n && hb_numel(p) == n_params && hb_numel(m) == n_params && hb_numel(v) == n_params,
             "allreduce_adam: %lld parameters, bucket of %lld, state of %lld", (long long)n_params, (long long)n,
             (long long)hb_numel(p));
  if (hb_lazy_on())
    return hb_lrec(HB_L_EFFECT, "allreduce_adam_step").in(bucket).in(p).in(m).in(v).in(t_counter).commit(
        [=](hb_lnode &, hb_array *const *in, hb_array **) {
          return hb_allreduce_adam_step(in[0], n_params, in[1], in[2], in[3], in[4], alpha, beta1, beta2, epsilon, scale);
        });
  HB_REQUIRE(!g_hb.trace_only, "allreduce_adam: trace-only mode has no device to run on");
  if (n == 0) return HB_OK;
  if (!g_fc_ready) HB_TRY(fcomm_setup(0));   // single process without hb_comm_init
  HB_REQUIRE(g_fc.world == 1 || (size_t)n * hb_dtype_size(bucket->dt) <= g_fc.cap,
             "allreduce_adam: bucket of %lld bytes exceeds the exchange buffer (%zu)", (long long)n * 8, g_fc.cap);
  // chunks: one or two elements per thread for the small buckets of the reference's models (latency is what counts),
  // all CTAs co-resident
  int chunks = (int)std::min<int64_t>(HB_COMM_MAXCHUNKS, std::max<int64_t>(1, (n + 511) / 512));
  if (bucket->dt == HB_F64)
    hb_allreduce_adam_kernel<double><<<chunks, 256, 0, g_hb.stream>>>(
        g_fc, (double *)hb_data(bucket), n, n_params, (double *)hb_data(p), (double *)hb_data(m), (double *)hb_data(v),
        (long long *)hb_data(t_counter), alpha, beta1, beta2, epsilon, scale);
  else
    hb_allreduce_adam_kernel<float><<<chunks, 256, 0, g_hb.stream>>>(
        g_fc, (float *)hb_data(bucket), n, n_params, (float *)hb_data(p), (float *)hb_data(m), (float *)hb_data(v),
        (long long *)hb_data(t_counter), alpha, beta1, beta2, epsilon, (float)scale);
  HB_LAUNCHED();
  return HB_OK;
}

// ---- global-objective mode: lossSoftMaxCrossEntropyS normalises over the WHOLE [10, batch]
// tensor (CommonShapedOps.hs:89-111), so a batch sharded over G ranks equals the reference's one-device step only if
// `minimum u` and `sum (exp (u - min))` are taken over all shards.  When switched on (and world > 1) hb_softmax_xent
// exchanges the two scalars through peer-mapped slots inside its kernel, combined in rank order on every rank.
extern "C" int hb_comm_set_global_objective(int on) {
  HB_CHECK_INIT();
  HB_REQUIRE(!on || g_nccl.world == 1 || g_fc_ready, "global objective: hb_comm_init has not been called");
  g_global_objective = on ? 1 : 0;
  return HB_OK;
}

bool hb_comm_scalar_xchg(hb_scalar_xchg *out) {
  if (!g_global_objective || g_nccl.world <= 1 || !g_fc_ready) return false;
  *out = g_sx;
  return true;
}

extern "C" int hb_comm_destroy(void) {
  g_global_objective = 0;
  fcomm_teardown();
  if (g_nccl.comm) {
    g_nccl.CommDestroy(g_nccl.comm);
    g_nccl.comm = nullptr;
  }
  g_nccl.world = 1;
  g_nccl.rank = 0;
  return HB_OK;
}

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

#include "common.cuh"
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

extern "C" int hb_comm_init(int rank, int world, const void *id_bytes) {
  HB_CHECK_INIT();
  HB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "comm: rank %d of %d", rank, world);
  g_nccl.rank = rank;
  g_nccl.world = world;
  if (world == 1) return HB_OK;
  HB_REQUIRE(id_bytes, "null unique id");
  HB_TRY(load_nccl());
  ncclUniqueId id;
  memcpy(&id, id_bytes, HB_UNIQUE_ID_BYTES);
  HB_NCCL(g_nccl.CommInitRank(&g_nccl.comm, world, id, rank));
  return HB_OK;
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

extern "C" int hb_comm_destroy(void) {
  if (g_nccl.comm) {
    g_nccl.CommDestroy(g_nccl.comm);
    g_nccl.comm = nullptr;
  }
  g_nccl.world = 1;
  g_nccl.rank = 0;
  return HB_OK;
}

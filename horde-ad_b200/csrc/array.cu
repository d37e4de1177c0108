// array.cu -- lifecycle, caching allocator, host<->device transfer, metadata
// and the kernel-free view operations of the device carrier.
//
// Reference counterparts: CarriersConcrete.hs:150-155,293 (carrier),
// OpsConcrete.hs:99-124 (shape queries, tsconcrete), :243-246 (replicate),
// :638-652 (slice/reverse/transpose/reshape), :1451-1490 (tindexZS).
#include <stdarg.h>

#include <algorithm>
#include <map>

#include "common.cuh"

hb_global g_hb;

static std::recursive_mutex g_entry_mu;
static thread_local int tl_bound_device = -1;

hb_entry_guard::hb_entry_guard() {
  g_entry_mu.lock();
  if (tl_bound_device != g_hb.device && g_hb.device >= 0) {
    cudaSetDevice(g_hb.device);  // thread-local in the runtime: cheap, and required for cuLaunchKernel / cudaMalloc
    tl_bound_device = g_hb.device;
  }
}
hb_entry_guard::~hb_entry_guard() { g_entry_mu.unlock(); }

static thread_local char tl_error[512] = "";

void hb_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof tl_error, fmt, ap);
  va_end(ap);
}

std::string hb_shape_str(const hb_array *a) {
  std::string s = "[";
  for (int i = 0; i < a->rank; i++) {
    if (i) s += ",";
    s += std::to_string(a->shape[i]);
  }
  return s + "]";
}

// ------------------------------------------------------- caching allocator
// Single-stream, size-class free lists (as the interpreter allocates one
// result per AST node).  Blocks are never returned to the
// driver before hb_shutdown; frees coming from foreign threads (GHC
// finalisers) are queued and recycled by the next allocation.
static std::mutex g_alloc_mu;
// pool 0 = general; pool k > 0 = private pool of the k-th captured graph:
// blocks allocated while capturing stay with the graph (its kernels have the
// addresses baked in) and are recycled only among that graph's own captures,
// so an eager allocation can never be handed a block a replay will overwrite.
static std::map<int, std::multimap<size_t, void *>> g_pools;
static std::vector<void *> g_all_blocks;
static int g_capture_pool = 0;  // pool new allocations are charged to
static int g_next_pool = 1;

static size_t size_class(size_t bytes) {
  if (bytes < 512) return 512;
  if (bytes < (1u << 20)) {  // next power of two below 1 MiB
    size_t s = 512;
    while (s < bytes) s <<= 1;
    return s;
  }
  const size_t q = 2u << 20;  // 2 MiB granularity above
  return (bytes + q - 1) / q * q;
}

static void *pool_take(int pool, size_t sz) {
  auto &fl = g_pools[pool];
  auto it = fl.find(sz);
  if (it == fl.end()) return nullptr;
  void *p = it->second;
  fl.erase(it);
  return p;
}

int hb_alloc_storage(size_t bytes, hb_storage **out) {
  size_t sz = size_class(bytes);
  void *p = nullptr;
  int pool;
  if (g_hb.trace_only) {  // no device: constants live in host memory (hb_init_trace_only)
    hb_storage *st = new hb_storage;
    st->ptr = calloc(1, sz);
    st->bytes = sz;
    st->pool = HB_POOL_HOST;
    st->rc.store(1);
    *out = st;
    return HB_OK;
  }
  {
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    pool = g_capture_pool;
    p = pool_take(pool, sz);
    // a capture may adopt idle blocks of the general pool (they leave it for good)
    if (!p && pool != 0) p = pool_take(0, sz);
  }
  if (!p) {
    cudaError_t e = cudaMalloc(&p, sz);
    if (e != cudaSuccess) {
      cudaGetLastError();
      // release the general cache and retry once
      if (!g_hb.capturing) {
        std::lock_guard<std::mutex> lk(g_alloc_mu);
        cudaStreamSynchronize(g_hb.stream);
        for (auto &kv : g_pools[0]) {
          cudaFree(kv.second);
          g_hb.reserved_bytes -= kv.first;
          for (auto &b : g_all_blocks)
            if (b == kv.second) b = nullptr;
        }
        g_pools[0].clear();
      }
      e = cudaMalloc(&p, sz);
      if (e != cudaSuccess) {
        cudaGetLastError();
        HB_FAIL(HB_ERR_CUDA, "cudaMalloc(%zu) failed: %s", sz, cudaGetErrorString(e));
      }
    }
    g_hb.reserved_bytes += sz;
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    g_all_blocks.push_back(p);
  }
  hb_storage *st = new hb_storage;
  st->ptr = p;
  st->bytes = sz;
  st->pool = pool;
  st->rc.store(1);
  g_hb.live_bytes += sz;
  *out = st;
  return HB_OK;
}

// pools of graphs that have been freed: their blocks belong to the general pool again
static std::vector<int> g_retired_pools;

void hb_storage_release(hb_storage *st) {
  if (!st) return;
  if (st->rc.fetch_sub(1) == 1) {
    if (st->pool == HB_POOL_VIRTUAL) {  // output block of a lazy node: drop whatever it currently stands for
      if (st->bound) hb_storage_release(st->bound);
      delete st;
      return;
    }
    if (st->pool == HB_POOL_HOST) {
      free(st->ptr);
      delete st;
      return;
    }
    g_hb.live_bytes -= st->bytes;
    {
      // Stream-ordered reuse: every consumer of the block was enqueued on the
      // single library stream before this point, and every later producer is
      // enqueued after it, so the block can be recycled immediately.
      std::lock_guard<std::mutex> lk(g_alloc_mu);
      int pool = st->pool;
      if (pool != 0 && std::find(g_retired_pools.begin(), g_retired_pools.end(), pool) != g_retired_pools.end())
        pool = 0;  // the graph that owned this block is gone: nothing replays into it any more
      g_pools[pool].insert({st->bytes, st->ptr});
    }
    delete st;
  }
}

// Outgrown scratch blocks are retired, not freed: captured graphs have their
// address baked in and may still be replayed (freed at hb_shutdown).
static std::vector<void *> g_retired_scratch;

int hb_scratch(size_t bytes, void **ptr) {
  if (g_hb.scratch_bytes < bytes) {
    HB_REQUIRE(!g_hb.capturing, "scratch growth (%zu bytes) during graph capture: run the step once eagerly first", bytes);
    size_t sz = bytes < (64u << 20) ? (64u << 20) : (bytes + (bytes >> 2));
    void *fresh = nullptr;
    HB_CUDA(cudaMalloc(&fresh, sz));
    if (g_hb.scratch) g_retired_scratch.push_back(g_hb.scratch);
    g_hb.scratch = fresh;
    g_hb.scratch_bytes = sz;
  }
  *ptr = g_hb.scratch;
  return HB_OK;
}

// --------------------------------------------------------------- lifecycle
extern "C" int hb_init(int device) {
  if (g_hb.inited) {
    HB_REQUIRE(!g_hb.trace_only, "hb_init: the library was initialised with hb_init_trace_only (no device)");
    HB_REQUIRE(g_hb.device == device, "hb_init: already initialised on device %d", g_hb.device);
    return HB_OK;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    HB_FAIL(HB_ERR_CUDA,
            "hb_init: no CUDA device available (%s). libhordeb200 has no CPU fallback.",
            e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  HB_REQUIRE(device >= 0 && device < n, "hb_init: device %d out of range (have %d)", device, n);
  HB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  HB_CUDA(cudaGetDeviceProperties(&prop, device));
  g_hb.sm_count = prop.multiProcessorCount;
  {
    const char *e = getenv("HB_CONV_FULL_PADDING");
    g_hb.conv_skip_padding = !(e && e[0] && e[0] != '0');
  }
  HB_CUDA(cudaStreamCreateWithFlags(&g_hb.stream, cudaStreamNonBlocking));
  for (int i = 0; i < 64; i++) HB_CUDA(cudaEventCreate(&g_hb.events[i]));
  g_hb.device = device;
  g_hb.inited = true;
  g_hb.launches = 0;
  return HB_OK;
}

extern "C" int hb_shutdown(void) {
  if (!g_hb.inited) return HB_OK;
  if (g_hb.trace_only) {
    g_hb.inited = false;
    g_hb.trace_only = false;
    g_hb.device = 0;
    return HB_OK;
  }
  cudaStreamSynchronize(g_hb.stream);
  {
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    for (void *p : g_all_blocks)
      if (p) cudaFree(p);
    g_all_blocks.clear();
    g_pools.clear();
  }
  if (g_hb.scratch) cudaFree(g_hb.scratch);
  for (void *p : g_retired_scratch) cudaFree(p);
  g_retired_scratch.clear();
  g_hb.scratch = nullptr;
  g_hb.scratch_bytes = 0;
  for (int i = 0; i < 64; i++) cudaEventDestroy(g_hb.events[i]);
  cudaStreamDestroy(g_hb.stream);
  g_hb.stream = nullptr;
  g_hb.inited = false;
  g_hb.reserved_bytes = 0;
  g_hb.live_bytes = 0;
  return HB_OK;
}

extern "C" int hb_sync(void) {
  HB_CHECK_INIT();
  HB_CUDA(cudaStreamSynchronize(g_hb.stream));
  return HB_OK;
}

extern "C" int hb_last_error(char *buf, size_t len) {
  if (!buf || len == 0) return HB_ERR_INVALID;
  strncpy(buf, tl_error, len - 1);
  buf[len - 1] = 0;
  return HB_OK;
}

extern "C" int hb_device_count(int *count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *count = n;
  return HB_OK;
}

extern "C" int hb_mem_stats(uint64_t *live, uint64_t *reserved) {
  if (live) *live = g_hb.live_bytes.load();
  if (reserved) *reserved = g_hb.reserved_bytes.load();
  return HB_OK;
}

extern "C" int hb_launch_count(uint64_t *c) {
  *c = g_hb.launches.load();
  return HB_OK;
}
extern "C" int hb_reset_launch_count(void) {
  g_hb.launches = 0;
  return HB_OK;
}
extern "C" int hb_stream(void **s) {
  HB_CHECK_INIT();
  *s = (void *)g_hb.stream;
  return HB_OK;
}

extern "C" int hb_event_record(int slot) {
  HB_CHECK_INIT();
  HB_REQUIRE(slot >= 0 && slot < 64, "event slot out of range");
  HB_CUDA(cudaEventRecord(g_hb.events[slot], g_hb.stream));
  return HB_OK;
}
extern "C" int hb_event_elapsed_ms(int a, int b, float *ms) {
  HB_CHECK_INIT();
  HB_REQUIRE(a >= 0 && a < 64 && b >= 0 && b < 64, "event slot out of range");
  HB_CUDA(cudaEventSynchronize(g_hb.events[b]));
  HB_CUDA(cudaEventElapsedTime(ms, g_hb.events[a], g_hb.events[b]));
  return HB_OK;
}

// ----------------------------------------------------- per-launch profiler
// While enabled every HB_LAUNCHED() records an event on the library stream;
// the device time between consecutive events is attributed to the launch site
// (host function:line).  The stream is serial, so as long as the host runs
// ahead of the GPU the difference is the kernel's own duration.  bench.py uses
// it for the live per-kernel roofline; never enabled on the product path.
struct hb_prof_rec {
  cudaEvent_t ev;
  std::string key;
};
static std::string g_prof_tag;

hb_prof_scope::hb_prof_scope(const char *name, const hb_array *a, const hb_array *b) {
  if (!g_hb.profiling || !g_prof_tag.empty()) return;
  owner = true;
  g_prof_tag = name;
  if (a) g_prof_tag += std::string("{") + hb_dtype_name(a->dt) + hb_shape_str(a) + (b ? ";" + hb_shape_str(b) : "") + "}";
}
hb_prof_scope::~hb_prof_scope() {
  if (owner) g_prof_tag.clear();
}
static std::vector<hb_prof_rec> g_prof;
static std::vector<cudaEvent_t> g_prof_free;

static cudaEvent_t prof_event() {
  cudaEvent_t e;
  if (!g_prof_free.empty()) {
    e = g_prof_free.back();
    g_prof_free.pop_back();
  } else {
    cudaEventCreate(&e);
  }
  return e;
}

void hb_prof_mark(const char *func, int line) {
  if (g_hb.capturing) return;
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, g_hb.stream);
  g_prof.push_back({e, (g_prof_tag.empty() ? std::string("-") : g_prof_tag) + " " + func + ":" + std::to_string(line)});
}

extern "C" int hb_prof_begin(void) {
  HB_CHECK_INIT();
  for (auto &r : g_prof) g_prof_free.push_back(r.ev);
  g_prof.clear();
  g_hb.profiling = true;
  hb_prof_mark("<begin>", 0);
  return HB_OK;
}

extern "C" int hb_prof_end(void) {
  HB_CHECK_INIT();
  g_hb.profiling = false;
  HB_CUDA(cudaStreamSynchronize(g_hb.stream));
  return HB_OK;
}

// One line per (entry point{shapes}, launch site), sorted by total time:
// "entry{dtype[shape];[shape]} func:line count total_ms".
extern "C" int hb_prof_report(char *buf, size_t len) {
  HB_CHECK_INIT();
  HB_REQUIRE(buf && len > 0, "null buffer");
  std::map<std::string, std::pair<int64_t, double>> agg;
  for (size_t i = 1; i < g_prof.size(); i++) {
    float ms = 0.f;
    HB_CUDA(cudaEventElapsedTime(&ms, g_prof[i - 1].ev, g_prof[i].ev));
    auto &a = agg[g_prof[i].key];
    a.first++;
    a.second += ms;
  }
  std::vector<std::pair<double, std::string>> rows;
  for (auto &kv : agg) {
    char line[256];
    snprintf(line, sizeof line, "%s %lld %.6f\n", kv.first.c_str(), (long long)kv.second.first, kv.second.second);
    rows.push_back({-kv.second.second, line});
  }
  std::sort(rows.begin(), rows.end());
  std::string out;
  for (auto &r : rows) out += r.second;
  strncpy(buf, out.c_str(), len - 1);
  buf[len - 1] = 0;
  return HB_OK;
}

extern "C" int hb_retain(hb_array *a) {
  HB_REQUIRE(a, "null array");
  a->rc.fetch_add(1);
  return HB_OK;
}

extern "C" int hb_release(hb_array *a) {
  if (!a) return HB_OK;
  if (a->rc.fetch_sub(1) == 1) {
    hb_storage_release(a->st);
    delete a;
  }
  return HB_OK;
}

// ------------------------------------------------------------ construction
int hb_new_array(hb_dtype dt, int rank, const int64_t *shape, hb_array **out) {
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(rank >= 0 && rank <= HB_R, "rank %d out of range (max %d)", rank, HB_R);
  HB_REQUIRE((int)dt >= 0 && (int)dt <= HB_BOOL, "bad dtype %d", (int)dt);
  hb_array *a = new hb_array;
  a->rc.store(1);
  a->dt = dt;
  a->rank = rank;
  a->offset = 0;
  int64_t n = 1;
  for (int i = 0; i < rank; i++) {
    if (shape[i] < 0) {
      delete a;
      HB_FAIL(HB_ERR_INVALID, "negative extent %lld", (long long)shape[i]);
    }
    a->shape[i] = shape[i];
    n *= shape[i];
  }
  int64_t s = 1;
  for (int i = rank - 1; i >= 0; i--) {
    a->strides[i] = s;
    s *= shape[i];
  }
  a->st = nullptr;
  if (n > 0) {
    int st = hb_alloc_storage((size_t)n * hb_dtype_size(dt), &a->st);
    if (st != HB_OK) {
      delete a;
      return st;
    }
  }
  *out = a;
  return HB_OK;
}

hb_array *hb_new_view(const hb_array *p) {
  hb_array *a = new hb_array;
  a->rc.store(1);
  a->st = p->st;
  if (a->st) a->st->rc.fetch_add(1);
  a->dt = p->dt;
  a->rank = p->rank;
  memcpy(a->shape, p->shape, sizeof a->shape);
  memcpy(a->strides, p->strides, sizeof a->strides);
  a->offset = p->offset;
  return a;
}

int hb_zeros(hb_dtype dt, int rank, const int64_t *shape, hb_array **out) {
  HB_TRY(hb_new_array(dt, rank, shape, out));
  int64_t n = hb_numel(*out);
  if (n > 0) HB_CUDA(cudaMemsetAsync(hb_data(*out), 0, (size_t)n * hb_dtype_size(dt), g_hb.stream));
  return HB_OK;
}

// ---------------------------------------------------------------- transfer
extern "C" int hb_from_host(hb_dtype dt, int rank, const int64_t *shape, const void *host,
                            hb_array **out) {
  HB_CHECK_INIT_LAZY();   // constants are created eagerly even while a lazy program is being recorded
  HB_REQUIRE(out, "null out");
  hb_array *a;
  HB_TRY(hb_new_array(dt, rank, shape, &a));
  int64_t n = hb_numel(a);
  if (n > 0 && host && (size_t)n * hb_dtype_size(dt) <= 4096) {
    // small constants (index tables, [0.0,1.0]) keep a host image for the contraction pass
    const uint8_t *hp = (const uint8_t *)host;
    a->st->host = std::make_shared<std::vector<uint8_t>>(hp, hp + (size_t)n * hb_dtype_size(dt));
  }
  if (n > 0 && g_hb.trace_only) {
    HB_REQUIRE(host, "null host pointer");
    memcpy(hb_data(a), host, (size_t)n * hb_dtype_size(dt));
    *out = a;
    return HB_OK;
  }
  if (n > 0) {
    HB_REQUIRE(host, "null host pointer");
    cudaError_t e = cudaMemcpyAsync(hb_data(a), host, (size_t)n * hb_dtype_size(dt),
                                    cudaMemcpyHostToDevice, g_hb.stream);
    if (e != cudaSuccess) {
      hb_release(a);
      HB_CUDA(e);
    }
    // pageable sources are staged synchronously by the runtime; pinned ones
    // complete in stream order, so the caller must keep them alive until the
    // next hb_sync (documented in INTEGRATION.md).
  }
  *out = a;
  return HB_OK;
}

extern "C" int hb_upload_into(hb_array *dst, const void *host) {
  HB_CHECK_INIT();
  HB_REQUIRE(dst && host, "null argument");
  HB_REQUIRE(hb_contig(dst), "hb_upload_into: destination must be contiguous");
  int64_t n = hb_numel(dst);
  if (n > 0)
    // cudaMemcpyDefault: the source may be pinned host memory or a device
    // pointer (a resident batch), resolved through unified addressing
    HB_CUDA(cudaMemcpyAsync(hb_data(dst), host, (size_t)n * hb_dtype_size(dst->dt),
                            cudaMemcpyDefault, g_hb.stream));
  return HB_OK;
}

// ---- prefetch: H2D on a second stream, overlapped with the compute stream
static cudaStream_t g_copy_stream = nullptr;
static cudaEvent_t g_ev_ready = nullptr, g_ev_consumed = nullptr;
static bool g_consumed_recorded = false;

static int prefetch_init() {
  if (g_copy_stream) return HB_OK;
  HB_CUDA(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
  HB_CUDA(cudaEventCreateWithFlags(&g_ev_ready, cudaEventDisableTiming));
  HB_CUDA(cudaEventCreateWithFlags(&g_ev_consumed, cudaEventDisableTiming));
  return HB_OK;
}

extern "C" int hb_prefetch_into(hb_array *dst, const void *host) {
  HB_CHECK_INIT();
  HB_REQUIRE(dst && host, "null argument");
  HB_REQUIRE(hb_contig(dst), "hb_prefetch_into: destination must be contiguous");
  HB_REQUIRE(!g_hb.capturing, "hb_prefetch_into: not inside a graph capture");
  HB_TRY(prefetch_init());
  if (g_consumed_recorded) HB_CUDA(cudaStreamWaitEvent(g_copy_stream, g_ev_consumed, 0));
  int64_t n = hb_numel(dst);
  if (n > 0)
    HB_CUDA(cudaMemcpyAsync(hb_data(dst), host, (size_t)n * hb_dtype_size(dst->dt), cudaMemcpyHostToDevice,
                            g_copy_stream));
  HB_CUDA(cudaEventRecord(g_ev_ready, g_copy_stream));
  return HB_OK;
}

extern "C" int hb_prefetch_ready(void) {
  HB_CHECK_INIT();
  HB_TRY(prefetch_init());
  HB_CUDA(cudaStreamWaitEvent(g_hb.stream, g_ev_ready, 0));
  return HB_OK;
}

extern "C" int hb_prefetch_consumed(void) {
  HB_CHECK_INIT();
  HB_TRY(prefetch_init());
  HB_CUDA(cudaEventRecord(g_ev_consumed, g_hb.stream));
  g_consumed_recorded = true;
  return HB_OK;
}

extern "C" int hb_to_host(const hb_array *a, void *host) {
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(a, "null array");
  int64_t n = hb_numel(a);
  if (n == 0) return HB_OK;
  HB_REQUIRE(host, "null host pointer");
  HB_REQUIRE(hb_data(a), "hb_to_host: the array is a lazy value that has not been computed (only the outputs of a lazy "
                         "program hold values, after hb_lazy_run)");
  if (g_hb.trace_only) {
    HB_REQUIRE(hb_contig(a), "hb_to_host: trace-only mode can only read back contiguous constants");
    memcpy(host, hb_data(a), (size_t)n * hb_dtype_size(a->dt));
    return HB_OK;
  }
  HB_REQUIRE(!hb_lazy_on() || hb_contig(a), "hb_to_host: a strided view cannot be materialised while recording");
  const hb_array *src = a;
  hb_array *tmp = nullptr;
  if (!hb_contig(a)) {
    HB_TRY(hb_contiguous(a, &tmp));
    src = tmp;
  }
  cudaError_t e = cudaMemcpyAsync(host, hb_data(src), (size_t)n * hb_dtype_size(a->dt),
                                  cudaMemcpyDeviceToHost, g_hb.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(g_hb.stream);
  if (tmp) hb_release(tmp);
  HB_CUDA(e);
  return HB_OK;
}

extern "C" int hb_scalar_to_host(const hb_array *a, void *value) {
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(a && value, "null argument");
  HB_REQUIRE(hb_numel(a) == 1, "hb_scalar_to_host: array %s has %lld elements",
             hb_shape_str(a).c_str(), (long long)hb_numel(a));
  HB_REQUIRE(hb_data(a), "hb_scalar_to_host: the array is a lazy value that has not been computed");
  if (g_hb.trace_only) {
    memcpy(value, hb_data(a), hb_dtype_size(a->dt));
    return HB_OK;
  }
  HB_CUDA(cudaMemcpyAsync(value, hb_data(a), hb_dtype_size(a->dt), cudaMemcpyDeviceToHost,
                          g_hb.stream));
  HB_CUDA(cudaStreamSynchronize(g_hb.stream));
  return HB_OK;
}

extern "C" int hb_full(hb_dtype dt, int rank, const int64_t *shape, const void *value,
                       hb_array **out) {
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(value && out, "null argument");
  HB_REQUIRE(rank >= 0 && rank <= HB_R, "rank %d out of range", rank);
  int64_t one = 1;
  hb_array *s;
  HB_TRY(hb_new_array(dt, 0, &one, &s));
  s->st->host = std::make_shared<std::vector<uint8_t>>((const uint8_t *)value, (const uint8_t *)value + hb_dtype_size(dt));
  // the value travels as a kernel parameter (safe under graph capture)
  int st = HB_OK;
  if (g_hb.trace_only) memcpy(hb_data(s), value, hb_dtype_size(dt));
  else st = hb_fill_scalar(hb_data(s), dt, value);
  if (st != HB_OK) {
    hb_release(s);
    return st;
  }
  s->rank = rank;
  for (int i = 0; i < rank; i++) {
    s->shape[i] = shape[i];
    s->strides[i] = 0;
  }
  *out = s;
  return HB_OK;
}

extern "C" int hb_host_alloc(size_t bytes, void **ptr) {
  HB_CHECK_INIT_LAZY();
  if (g_hb.trace_only) {
    *ptr = calloc(1, bytes ? bytes : 1);
    return HB_OK;
  }
  HB_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return HB_OK;
}
extern "C" int hb_host_free(void *ptr) {
  if (ptr && g_hb.trace_only) {
    free(ptr);
    return HB_OK;
  }
  if (ptr) HB_CUDA(cudaFreeHost(ptr));
  return HB_OK;
}

// ---------------------------------------------------------------- metadata
extern "C" int hb_rank(const hb_array *a, int *rank) {
  HB_REQUIRE(a && rank, "null argument");
  *rank = a->rank;
  return HB_OK;
}
extern "C" int hb_shape(const hb_array *a, int64_t *shape) {
  HB_REQUIRE(a && shape, "null argument");
  for (int i = 0; i < a->rank; i++) shape[i] = a->shape[i];
  return HB_OK;
}
extern "C" int hb_strides(const hb_array *a, int64_t *strides) {
  HB_REQUIRE(a && strides, "null argument");
  for (int i = 0; i < a->rank; i++) strides[i] = a->strides[i];
  return HB_OK;
}
extern "C" int hb_dtype_of(const hb_array *a, hb_dtype *dt) {
  HB_REQUIRE(a && dt, "null argument");
  *dt = a->dt;
  return HB_OK;
}
extern "C" int hb_size(const hb_array *a, int64_t *n) {
  HB_REQUIRE(a && n, "null argument");
  *n = hb_numel(a);
  return HB_OK;
}
extern "C" int hb_is_contiguous(const hb_array *a, int *flag) {
  HB_REQUIRE(a && flag, "null argument");
  *flag = hb_contig(a) ? 1 : 0;
  return HB_OK;
}
extern "C" int hb_device_ptr(const hb_array *a, void **ptr) {
  HB_REQUIRE(a && ptr, "null argument");
  *ptr = hb_data(a);
  return HB_OK;
}

// ------------------------------------------------------------------- views
extern "C" int hb_transpose(const hb_array *a, int nperm, const int *perm, hb_array **out) {
  HB_REQUIRE(a && out, "null argument");
  HB_REQUIRE(nperm >= 0 && nperm <= a->rank, "transpose: permutation of length %d on rank %d",
             nperm, a->rank);
  bool seen[HB_R] = {};
  for (int i = 0; i < nperm; i++) {
    HB_REQUIRE(perm[i] >= 0 && perm[i] < nperm && !seen[perm[i]],
               "transpose: not a permutation of 0..%d", nperm - 1);
    seen[perm[i]] = true;
  }
  hb_array *v = hb_new_view(a);
  for (int i = 0; i < nperm; i++) {
    v->shape[i] = a->shape[perm[i]];
    v->strides[i] = a->strides[perm[i]];
  }
  *out = v;
  return HB_OK;
}

extern "C" int hb_replicate(const hb_array *a, int nlead, const int64_t *lead, hb_array **out) {
  HB_REQUIRE(a && out, "null argument");
  HB_REQUIRE(nlead >= 0 && a->rank + nlead <= HB_R, "replicate: rank %d + %d exceeds %d", a->rank,
             nlead, HB_R);
  hb_array *v = hb_new_view(a);
  v->rank = a->rank + nlead;
  for (int i = 0; i < nlead; i++) {
    if (lead[i] < 0) {
      hb_release(v);
      HB_FAIL(HB_ERR_INVALID, "replicate: negative extent");
    }
    v->shape[i] = lead[i];
    v->strides[i] = 0;
  }
  for (int i = 0; i < a->rank; i++) {
    v->shape[nlead + i] = a->shape[i];
    v->strides[nlead + i] = a->strides[i];
  }
  *out = v;
  return HB_OK;
}

extern "C" int hb_slice(const hb_array *a, int64_t i, int64_t n, hb_array **out) {
  HB_REQUIRE(a && out, "null argument");
  HB_REQUIRE(a->rank >= 1, "slice: rank 0");
  HB_REQUIRE(i >= 0 && n >= 0 && i + n <= a->shape[0], "slice [%lld,+%lld) out of extent %lld",
             (long long)i, (long long)n, (long long)a->shape[0]);
  hb_array *v = hb_new_view(a);
  v->shape[0] = n;
  v->offset += i * a->strides[0];
  *out = v;
  return HB_OK;
}

extern "C" int hb_reverse(const hb_array *a, hb_array **out) {
  HB_REQUIRE(a && out, "null argument");
  HB_REQUIRE(a->rank >= 1, "reverse: rank 0");
  hb_array *v = hb_new_view(a);
  if (a->shape[0] > 0) v->offset += (a->shape[0] - 1) * a->strides[0];
  v->strides[0] = -a->strides[0];
  *out = v;
  return HB_OK;
}

// Try to express `shape` as a view of a's storage (numpy's no-copy reshape).
static bool try_view_reshape(const hb_array *a, int rank, const int64_t *shape,
                             int64_t *newstrides) {
  int oldnd = 0;
  int64_t olds[HB_R], oldst[HB_R];
  for (int i = 0; i < a->rank; i++)
    if (a->shape[i] != 1) {
      olds[oldnd] = a->shape[i];
      oldst[oldnd] = a->strides[i];
      oldnd++;
    }
  int oi = 0, oj = 1, ni = 0, nj = 1;
  while (ni < rank && oi < oldnd) {
    int64_t np = shape[ni], op = olds[oi];
    while (np != op) {
      if (np < op)
        np *= shape[nj++];
      else
        op *= olds[oj++];
    }
    for (int ok = oi; ok < oj - 1; ok++)
      if (oldst[ok] != olds[ok + 1] * oldst[ok + 1]) return false;
    newstrides[nj - 1] = oldst[oj - 1];
    for (int nk = nj - 1; nk > ni; nk--) newstrides[nk - 1] = newstrides[nk] * shape[nk];
    ni = nj++;
    oi = oj++;
  }
  int64_t last = ni > 0 ? newstrides[ni - 1] : 1;
  for (int nk = ni; nk < rank; nk++) newstrides[nk] = last;
  return true;
}

extern "C" int hb_reshape(const hb_array *a, int rank, const int64_t *shape, hb_array **out) {
  HB_REQUIRE(a && out, "null argument");
  HB_REQUIRE(rank >= 0 && rank <= HB_R, "reshape: rank %d out of range", rank);
  int64_t n = 1;
  for (int i = 0; i < rank; i++) n *= shape[i];
  HB_REQUIRE(n == hb_numel(a), "reshape: %s (%lld elements) to %lld elements",
             hb_shape_str(a).c_str(), (long long)hb_numel(a), (long long)n);
  int64_t ns[HB_R];
  if (n == 0 || try_view_reshape(a, rank, shape, ns)) {
    hb_array *v = hb_new_view(a);
    v->rank = rank;
    int64_t s = 1;
    for (int i = rank - 1; i >= 0; i--) {
      v->shape[i] = shape[i];
      v->strides[i] = n == 0 ? s : ns[i];
      s *= shape[i];
    }
    *out = v;
    return HB_OK;
  }
  hb_array *c;
  HB_TRY(hb_contiguous(a, &c));
  c->rank = rank;
  int64_t s = 1;
  for (int i = rank - 1; i >= 0; i--) {
    c->shape[i] = shape[i];
    c->strides[i] = s;
    s *= shape[i];
  }
  *out = c;
  return HB_OK;
}

extern "C" int hb_index_prefix(const hb_array *a, int k, const int64_t *ix, hb_array **out) {
  HB_REQUIRE(a && out, "null argument");
  HB_REQUIRE(k >= 0 && k <= a->rank, "index: %d indices into rank %d", k, a->rank);
  bool inb = true;
  int64_t off = 0;
  for (int i = 0; i < k; i++) {
    if (ix[i] < 0 || ix[i] >= a->shape[i]) inb = false;
    off += ix[i] * a->strides[i];
  }
  if (!inb) {
    // OpsConcrete.hs:1481-1490: out of bounds -> def (zeros) of shape shn
    double zero[2] = {0, 0};
    return hb_full(a->dt, a->rank - k, a->shape + k, zero, out);
  }
  hb_array *v = hb_new_view(a);
  v->rank = a->rank - k;
  for (int i = 0; i < v->rank; i++) {
    v->shape[i] = a->shape[k + i];
    v->strides[i] = a->strides[k + i];
  }
  v->offset += off;
  *out = v;
  return HB_OK;
}

// ------------------------------------------------------ iteration collapse
void hb_collapse(int *rank, int64_t *shape, int nops, int64_t (*strides)[HB_R]) {
  // drop size-1 axes
  int r = 0;
  for (int i = 0; i < *rank; i++) {
    if (shape[i] == 1) continue;
    shape[r] = shape[i];
    for (int o = 0; o < nops; o++) strides[o][r] = strides[o][i];
    r++;
  }
  // merge axis i+1 into axis i when stride[i] == stride[i+1] * shape[i+1] for all operands
  int w = 0;
  for (int i = 1; i < r; i++) {
    bool ok = true;
    for (int o = 0; o < nops; o++)
      if (strides[o][w] != strides[o][i] * shape[i]) ok = false;
    if (ok) {
      shape[w] *= shape[i];
      for (int o = 0; o < nops; o++) strides[o][w] = strides[o][i];
    } else {
      w++;
      shape[w] = shape[i];
      for (int o = 0; o < nops; o++) strides[o][w] = strides[o][i];
    }
  }
  *rank = r == 0 ? 0 : w + 1;
}

void hb_make_iter(int nops, const hb_array *const *ops, hb_iter *it) {
  it->rank = ops[0]->rank;
  for (int i = 0; i < it->rank; i++) it->shape[i] = ops[0]->shape[i];
  for (int o = 0; o < nops; o++)
    for (int i = 0; i < it->rank; i++) it->stride[o][i] = ops[o]->strides[i];
  hb_collapse(&it->rank, it->shape, nops, it->stride);
}

// ------------------------------------------------------------- CUDA graphs
struct hb_graph {
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  uint64_t n_kernels;  // kernels recorded during capture (per replay)
  int pool;            // private allocator pool
};
static uint64_t g_capture_start = 0;

extern "C" int hb_graph_begin(void) {
  HB_CHECK_INIT();
  HB_REQUIRE(!g_hb.capturing, "hb_graph_begin: already capturing");
  HB_CUDA(cudaStreamBeginCapture(g_hb.stream, cudaStreamCaptureModeRelaxed));
  g_hb.capturing = true;
  g_capture_start = g_hb.launches.load();
  {
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    g_capture_pool = g_next_pool++;
  }
  return HB_OK;
}
extern "C" int hb_graph_end(hb_graph **out) {
  HB_CHECK_INIT();
  HB_REQUIRE(g_hb.capturing, "hb_graph_end: not capturing");
  g_hb.capturing = false;
  int pool;
  {
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    pool = g_capture_pool;
    g_capture_pool = 0;
  }
  cudaGraph_t gr;
  HB_CUDA(cudaStreamEndCapture(g_hb.stream, &gr));
  cudaGraphExec_t ex;
  cudaError_t e = cudaGraphInstantiate(&ex, gr, 0);
  if (e != cudaSuccess) {
    cudaGraphDestroy(gr);
    HB_CUDA(e);
  }
  hb_graph *g = new hb_graph{gr, ex, g_hb.launches.load() - g_capture_start, pool};
  g_hb.launches = g_capture_start;  // capture enqueued nothing
  *out = g;
  return HB_OK;
}
extern "C" int hb_graph_launch(hb_graph *g) {
  HB_CHECK_INIT();
  HB_REQUIRE(g, "null graph");
  HB_CUDA(cudaGraphLaunch(g->exec, g_hb.stream));
  g_hb.launches.fetch_add(g->n_kernels, std::memory_order_relaxed);
  return HB_OK;
}
extern "C" int hb_graph_free(hb_graph *g) {
  if (!g) return HB_OK;
  cudaGraphExecDestroy(g->exec);
  cudaGraphDestroy(g->graph);
  {
    // idle blocks of the private pool rejoin the general pool; blocks still
    // referenced by live arrays follow when those arrays are released
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    auto it = g_pools.find(g->pool);
    if (it != g_pools.end()) {
      for (auto &kv : it->second) g_pools[0].insert(kv);
      g_pools.erase(it);
    }
    g_retired_pools.push_back(g->pool);
  }
  delete g;
  return HB_OK;
}

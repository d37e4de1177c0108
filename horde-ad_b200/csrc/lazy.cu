// lazy.cu -- recording, liveness and replay of lazy programs (see lazy.cuh); the pattern
// rewriter itself lives in lazy_rewrite.cu.
//
// Reference counterpart: horde-ad interprets a gradient ARTIFACT -- a fixed, shape-specific op
// program (README.md:174-184; AstInterpret.hs:69-366) -- once per minibatch, calling one class
// method per AST node.  Recording those calls once below the C ABI gives the backend the same
// whole-program view `contractAst` has on the Haskell side (AstTraverse.hs:290-699) without any
// change above the ABI.
#include <algorithm>
#include <map>
#include <set>

#include "lazy.cuh"

static hb_lazy *g_rec = nullptr;   // program being recorded (guarded by the entry lock)
static int g_suspend = 0;
static int g_next_node_id = 1;

bool hb_lazy_on() { return g_rec != nullptr && g_suspend == 0; }
hb_lazy_suspend::hb_lazy_suspend() { g_suspend++; }
hb_lazy_suspend::~hb_lazy_suspend() { g_suspend--; }

// ------------------------------------------------------------ virtual blocks
static void unbind(hb_storage *V) {
  if (V && V->bound) {
    hb_storage_release(V->bound);
    V->bound = nullptr;
    V->ptr = nullptr;
  }
}

hb_array *hb_lazy_new_out(hb_lnode *producer, int out_idx, hb_dtype dt, int rank, const int64_t *shape) {
  hb_array *a = new hb_array;
  a->rc.store(1);
  a->dt = dt;
  a->rank = rank;
  a->offset = 0;
  int64_t n = 1;
  for (int i = 0; i < rank; i++) {
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
    hb_storage *V = new hb_storage;
    V->ptr = nullptr;
    V->bytes = (size_t)n * hb_dtype_size(dt);
    V->pool = HB_POOL_VIRTUAL;
    V->rc.store(1);
    V->producer = producer;
    V->out_idx = out_idx;
    a->st = V;
  }
  return a;
}

hb_lnode *hb_lazy_new_node(int op, const char *name) {
  hb_lnode *n = new hb_lnode;
  n->op = op;
  n->name = name;
  n->id = g_next_node_id++;
  return n;
}

void hb_lazy_node_add_in(hb_lnode *n, const hb_array *a) { n->in.push_back(a ? hb_new_view(a) : nullptr); }

void hb_lazy_node_free(hb_lnode *n) {
  for (hb_array *a : n->in) hb_release(a);
  for (hb_array *a : n->out) {
    if (a && a->st && a->st->producer == n) a->st->producer = nullptr;
    hb_release(a);
  }
  if (n->prog) hb_ixprog_free(n->prog);
  delete n;
}

const void *hb_host_copy(const hb_array *a) {
  if (!a || !a->st || !a->st->host) return nullptr;
  return a->st->host->data();
}

// ------------------------------------------------------------------- builder
hb_lrec::hb_lrec(int op, const char *name) { n = hb_lazy_new_node(op, name); }
hb_lrec &hb_lrec::in(const hb_array *a) {
  hb_lazy_node_add_in(n, a);
  return *this;
}
hb_lrec &hb_lrec::prog(const hb_ixprog *p) {
  n->prog = const_cast<hb_ixprog *>(p);
  if (p) n->prog->rc.fetch_add(1);
  return *this;
}
hb_lrec &hb_lrec::out(hb_array **dst, hb_dtype dt, int rank, const int64_t *shape) {
  if (!dst) {
    n->out.push_back(nullptr);
    return *this;
  }
  if (rank < 0 || rank > HB_R) {
    status = HB_ERR_INVALID;
    hb_set_error("%s: result rank %d out of range", n->name, rank);
    n->out.push_back(nullptr);
    return *this;
  }
  hb_array *h = hb_lazy_new_out(n, (int)n->out.size(), dt, rank, shape);
  n->out.push_back(hb_new_view(h));
  *dst = h;
  return *this;
}
int hb_lrec::commit(hb_lrun run) {
  if (status != HB_OK || !g_rec) {
    hb_lazy_node_free(n);
    return status != HB_OK ? status : HB_ERR_INVALID;
  }
  n->run = std::move(run);
  g_rec->nodes.push_back(n);
  return HB_OK;
}

// --------------------------------------------------------- liveness and DCE
static void for_each_operand(const hb_lnode *n, const std::function<void(const hb_array *)> &f) {
  for (const hb_array *a : n->in)
    if (a) f(a);
  if (n->prog)
    for (int t = 0; t < n->prog->n_tensors; t++) f(n->prog->held[t]);
}

// marks dead nodes and computes, per live node index, which virtual blocks see their last use there
static void analyse(hb_lazy *P, std::vector<std::vector<hb_storage *>> *dying) {
  const int N = (int)P->nodes.size();
  std::set<const hb_lnode *> mine(P->nodes.begin(), P->nodes.end());
  std::set<const hb_lnode *> needed;
  std::vector<const hb_lnode *> work;
  auto need = [&](const hb_array *a) {
    hb_lnode *p = hb_producer(a);
    if (p && mine.count(p) && !needed.count(p)) {
      needed.insert(p);
      work.push_back(p);
    }
  };
  for (hb_array *o : P->outputs) need(o);
  for (hb_lnode *n : P->nodes)
    if (n->op == HB_L_EFFECT && !n->dead && !needed.count(n)) {
      needed.insert(n);
      work.push_back(n);
    }
  while (!work.empty()) {
    const hb_lnode *n = work.back();
    work.pop_back();
    for_each_operand(n, need);
  }
  for (hb_lnode *n : P->nodes)
    if (!needed.count(n)) n->dead = true;
  // last use per virtual block
  std::map<hb_storage *, int> last;
  for (int i = 0; i < N; i++) {
    hb_lnode *n = P->nodes[i];
    if (n->dead) continue;
    for (hb_array *o : n->out)
      if (o && o->st) last[o->st] = i;  // produced here: dies here unless somebody reads it later
    for_each_operand(n, [&](const hb_array *a) {
      if (a->st && a->st->pool == HB_POOL_VIRTUAL && hb_producer(a) && mine.count(hb_producer(a))) last[a->st] = i;
    });
  }
  for (hb_array *o : P->outputs)
    if (o && o->st) last.erase(o->st);  // program outputs stay bound
  if (dying) {
    dying->assign(N, {});
    for (auto &kv : last) (*dying)[kv.second].push_back(kv.first);
  }
  // taddTarget in place: operand 0 is a whole virtual block that dies at this node
  for (int i = 0; i < N; i++) {
    hb_lnode *n = P->nodes[i];
    n->inplace_ok = false;
    if (n->dead || n->op != HB_L_ADD_ACC) continue;
    const hb_array *acc = n->in[0], *c = n->in[1];
    if (!acc->st || acc->st->pool != HB_POOL_VIRTUAL || !hb_contig(acc) || acc->offset != 0) continue;
    if ((size_t)hb_numel(acc) * hb_dtype_size(acc->dt) != acc->st->bytes) continue;
    if (c->st == acc->st) continue;
    auto it = last.find(acc->st);
    n->inplace_ok = it != last.end() && it->second == i;
  }
}

// -------------------------------------------------------------------- replay
static int bind_out(hb_array *virt, hb_array *real) {
  if (!virt || !virt->st) {
    hb_release(real);
    return HB_OK;
  }
  if (!real) HB_FAIL(HB_ERR_INVALID, "lazy replay: a recorded output was not produced");
  hb_storage *V = virt->st;
  hb_array *r = real, *tmp = nullptr;
  if (hb_numel(real) != hb_numel(virt) || real->dt != virt->dt) {
    hb_release(real);
    HB_FAIL(HB_ERR_INVALID, "lazy replay: result %s does not match the recorded %s", hb_shape_str(real).c_str(),
            hb_shape_str(virt).c_str());
  }
  if (!hb_contig(real)) {  // e.g. a gather that the eager path answered with a strided view
    int st = hb_contiguous(real, &tmp);
    hb_release(real);
    if (st != HB_OK) return st;
    r = tmp;
  }
  hb_storage *R = r->st;
  void *p = hb_data(r);
  if (R && R->pool == HB_POOL_VIRTUAL) R = R->bound;  // an in-place result over an operand's block
  if (!R) {
    hb_release(r);
    HB_FAIL(HB_ERR_INVALID, "lazy replay: result has no storage");
  }
  R->rc.fetch_add(1);
  unbind(V);
  V->bound = R;
  V->ptr = p;
  hb_release(r);
  return HB_OK;
}

static void rebind_prog(hb_ixprog *p) {
  for (int t = 0; t < p->n_tensors; t++) p->vm.tensors[t].ptr = hb_data(p->held[t]);
}

static int run_program(hb_lazy *P) {
  hb_lazy_suspend guard;
  std::vector<std::vector<hb_storage *>> dying;
  analyse(P, &dying);
  const int N = (int)P->nodes.size();
  int st = HB_OK;
  for (int i = 0; i < N && st == HB_OK; i++) {
    hb_lnode *n = P->nodes[i];
    if (n->dead) continue;
    if (n->prog) rebind_prog(n->prog);
    hb_array *outs[8] = {};
    if (n->out.size() > 8) HB_FAIL(HB_ERR_INVALID, "lazy replay: too many outputs");
    st = n->run(*n, n->in.data(), outs);
    if (st != HB_OK) {
      for (hb_array *o : outs) hb_release(o);
      break;
    }
    for (size_t k = 0; k < n->out.size() && st == HB_OK; k++) st = bind_out(n->out[k], outs[k]);
    for (hb_storage *V : dying[i]) unbind(V);
  }
  if (st != HB_OK)
    for (hb_lnode *n : P->nodes)
      for (hb_array *o : n->out)
        if (o && o->st) unbind(o->st);
  P->runs++;
  return st;
}

// -------------------------------------------------------------------- C ABI
extern "C" int hb_init_trace_only(void) {
  if (g_hb.inited) {
    HB_REQUIRE(g_hb.trace_only, "hb_init_trace_only: the library is already initialised on device %d", g_hb.device);
    return HB_OK;
  }
  g_hb.device = -1;
  g_hb.trace_only = true;
  g_hb.inited = true;
  return HB_OK;
}

extern "C" int hb_lazy_begin(void) {
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(!g_rec, "hb_lazy_begin: a program is already being recorded");
  HB_REQUIRE(!g_hb.capturing, "hb_lazy_begin: not inside hb_graph_begin (record first, then capture hb_lazy_run)");
  g_rec = new hb_lazy;
  return HB_OK;
}

extern "C" int hb_lazy_abort(void) {
  HB_CHECK_INIT_LAZY();
  if (!g_rec) return HB_OK;
  hb_lazy *P = g_rec;
  g_rec = nullptr;
  for (hb_lnode *n : P->nodes) hb_lazy_node_free(n);
  delete P;
  return HB_OK;
}

extern "C" int hb_lazy_end(int n_out, hb_array *const *outs, unsigned flags, hb_lazy **prog) {
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(g_rec, "hb_lazy_end: no program is being recorded");
  HB_REQUIRE(prog && n_out >= 0 && (n_out == 0 || outs), "hb_lazy_end: bad argument");
  hb_lazy *P = g_rec;
  g_rec = nullptr;
  for (int i = 0; i < n_out; i++) P->outputs.push_back(hb_new_view(outs[i]));
  P->n_recorded = (int)P->nodes.size();
  int st = HB_OK;
  if (flags & HB_LAZY_REWRITE) st = hb_lazy_rewrite(P, flags);
  if (st != HB_OK) {
    hb_lazy *tmp = P;
    for (hb_lnode *n : tmp->nodes) hb_lazy_node_free(n);
    for (hb_array *o : tmp->outputs) hb_release(o);
    delete tmp;
    return st;
  }
  analyse(P, nullptr);
  P->n_rewritten = 0;
  for (hb_lnode *n : P->nodes) P->n_rewritten += n->dead ? 0 : 1;
  *prog = P;
  return HB_OK;
}

extern "C" int hb_lazy_run(hb_lazy *P) {
  HB_CHECK_INIT_LAZY();
  HB_REQUIRE(P, "null program");
  HB_REQUIRE(!g_hb.trace_only, "hb_lazy_run: trace-only mode has no device to run on");
  HB_REQUIRE(!hb_lazy_on(), "hb_lazy_run: cannot replay a program while another one is being recorded");
  return run_program(P);
}

extern "C" int hb_lazy_free(hb_lazy *P) {
  if (!P) return HB_OK;
  HB_ENTRY_LOCK();
  for (hb_lnode *n : P->nodes) hb_lazy_node_free(n);
  for (hb_array *o : P->outputs) hb_release(o);
  delete P;
  return HB_OK;
}

extern "C" int hb_lazy_stats(const hb_lazy *P, int *recorded, int *live) {
  HB_REQUIRE(P, "null program");
  if (recorded) *recorded = P->n_recorded;
  if (live) *live = P->n_rewritten;
  return HB_OK;
}

static std::string describe(const hb_array *a) {
  if (!a) return "-";
  std::string s = hb_dtype_name(a->dt) + hb_shape_str(a);
  hb_lnode *p = hb_producer(a);
  if (p) s += "<-%" + std::to_string(p->id) + (a->st->out_idx ? "." + std::to_string(a->st->out_idx) : "");
  else if (a->st && a->st->pool != HB_POOL_VIRTUAL) s += "<-const";
  return s;
}

// One line per LIVE node after the rewriter: "%id name(op args) out-shapes <- operands  # note", then the
// rewriter's log.  For tests and for DESIGN.md's listing of what the pass did to the CNN step.
extern "C" int hb_lazy_dump(const hb_lazy *P, char *buf, size_t len) {
  HB_REQUIRE(P && buf && len > 0, "hb_lazy_dump: bad argument");
  std::string s;
  for (const hb_lnode *n : P->nodes) {
    if (n->dead) continue;
    s += "%" + std::to_string(n->id) + " " + n->name;
    if (n->op == HB_L_BINARY) {
      static const char *bn[] = {"add", "sub", "mul", "divide", "power", "logbase", "atan2", "quot", "rem", "max", "min"};
      s += std::string(".") + (n->ia[0] >= 0 && n->ia[0] < 11 ? bn[n->ia[0]] : "?");
    }
    s += " ";
    for (size_t k = 0; k < n->out.size(); k++) s += (k ? "," : "") + (n->out[k] ? hb_dtype_name(n->out[k]->dt) + hb_shape_str(n->out[k]) : std::string("-"));
    s += " <-";
    for (const hb_array *a : n->in) s += " " + describe(a);
    if (!n->note.empty()) s += "  # " + n->note;
    s += "\n";
  }
  s += "--\n" + P->log;
  strncpy(buf, s.c_str(), len - 1);
  buf[len - 1] = 0;
  return HB_OK;
}

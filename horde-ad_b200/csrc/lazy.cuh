// lazy.cuh -- the recording side of the device contraction pass (the
// reference's sibling is `contractAst`, AstTraverse.hs:290-699, "one such pass per backend").
//
// Between hb_lazy_begin and hb_lazy_end the compute entry points of the C ABI do not
// launch anything: each call appends a node (op, operand views, scalar arguments, index
// program) to the program being recorded and hands back a LAZY array -- an ordinary
// `hb_array` over a *virtual* storage block that knows its producer node.  View calls
// (transpose / replicate / slice / reverse / reshape-as-view / index) keep working on
// lazy arrays unchanged: they are metadata over the virtual block.  hb_lazy_end runs the
// rewriter (lazy_rewrite.cu) over the recorded DAG and returns a program that
// hb_lazy_run replays, eagerly or inside hb_graph_begin / hb_graph_end.
//
// Everything here is host code; it is what lets the UNCHANGED vectorised program that
// horde-ad's interpreter sends (gather chains, sdot1In over broadcast views, kargMax
// gathers, scatters) reach the fused kernels without a `fused=` switch above the ABI.
#pragma once
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "vm.cuh"

enum hb_lop {
  HB_L_GENERIC = 0,  // opaque call: replayed as recorded, never rewritten
  HB_L_BINARY,       // ia[0] = hb_binop
  HB_L_UNARY,        // ia[0] = hb_unop
  HB_L_COPY,         // hb_contiguous (also the copying case of hb_reshape)
  HB_L_SUM_LEADING,  // ia[0] = n_axes
  HB_L_DOT_INNER,    // ia[0] = n_axes summed in front (0 = plain sdot1In)
  HB_L_MATMUL2,
  HB_L_GATHER,       // ia[0] = shm_rank, ia[1] = shp_rank, prog
  HB_L_SCATTER,      // ia[0] = shm_rank, ia[1] = shp_rank, ia[2..] = shp, prog
  HB_L_ADD_ACC,      // taddTarget: out = in0 + in1, in place when in0 dies here
  HB_L_CONV_FWD,     // in: K, A
  HB_L_CONV_DK,      // in: A, dB; ia[0] = KH, ia[1] = KW
  HB_L_CONV_DI,      // in: K, dB
  HB_L_RELU_POOL_FWD,  // in: pre, [bias]; ia[0] = ksize; out: y, code
  HB_L_RELU_POOL_BWD,  // in: dy, code; ia[0] = ksize, ia[1] = H, ia[2] = W; out: dpre, [dbias]
  HB_L_CRP_FWD,      // in: K, [bias], A; ia[0] = ksize, ia[1] = has_bias; out: y, code
  HB_L_CRP_BWD,      // in: K, A, dy, code; ia[0] = ksize; out: dK, [dbias], [dA]
  HB_L_EFFECT,       // in-place on its inputs (all-reduce, optimizer step, upload): a barrier, never dead
  HB_L_OP_COUNT
};

struct hb_lnode;
struct hb_lazy;

// replay thunk: `in` are the node's operand views (their storage is bound by now), `out` receives one fresh
// array per recorded output (nullptr slots for outputs that were not requested)
typedef std::function<int(hb_lnode &, hb_array *const *in, hb_array **out)> hb_lrun;

struct hb_lnode {
  int op = HB_L_GENERIC;
  int id = 0;
  const char *name = "";
  std::vector<hb_array *> in;    // retained views
  std::vector<hb_array *> out;   // retained contiguous views of the virtual output blocks (nullptr = not requested)
  int64_t ia[HB_R + 4] = {};
  double fa[2] = {};
  hb_ixprog *prog = nullptr;     // retained
  hb_lrun run;
  bool dead = false;
  bool inplace_ok = false;       // ADD_ACC: operand 0 dies here and may be overwritten
  std::string note;              // set by the rewriter: what this node replaced
};

// ------------------------------------------------------------------ recording
bool hb_lazy_on();                // a program is being recorded and recording is not suspended
struct hb_lazy_suspend {          // RAII: entry points called while replaying / from inside the library run eagerly
  hb_lazy_suspend();
  ~hb_lazy_suspend();
};

// builder used by the entry points
struct hb_lrec {
  hb_lnode *n;
  int status = HB_OK;
  hb_lrec(int op, const char *name);
  hb_lrec &in(const hb_array *a);            // nullptr operands are recorded as absent (kept as nullptr)
  hb_lrec &iarg(int k, int64_t v) { n->ia[k] = v; return *this; }
  hb_lrec &farg(int k, double v) { n->fa[k] = v; return *this; }
  hb_lrec &prog(const hb_ixprog *p);
  // declares an output; *dst receives the lazy handle (dst == nullptr: output not requested)
  hb_lrec &out(hb_array **dst, hb_dtype dt, int rank, const int64_t *shape);
  int commit(hb_lrun run);
};

// fresh lazy array over a virtual block (numel 0: no storage)
hb_array *hb_lazy_new_out(hb_lnode *producer, int out_idx, hb_dtype dt, int rank, const int64_t *shape);

// ------------------------------------------------------------------ programs
struct hb_lazy {
  std::vector<hb_lnode *> nodes;        // program order
  std::vector<hb_array *> outputs;      // retained
  std::vector<int> last_use;            // per node id -> index of the last node reading any of its outputs
  int n_recorded = 0, n_rewritten = 0;  // nodes before / after the rewriter
  uint64_t runs = 0;
  std::string log;                      // rewriter decisions, one line each
};

// the rewriter (lazy_rewrite.cu): pattern-matches over P->nodes, replaces sub-DAGs by fused nodes, marks dead nodes
int hb_lazy_rewrite(hb_lazy *P, unsigned flags);
// helpers shared by the rewriter
static inline hb_lnode *hb_producer(const hb_array *a) { return a && a->st ? (hb_lnode *)a->st->producer : nullptr; }
hb_lnode *hb_lazy_new_node(int op, const char *name);
void hb_lazy_node_add_in(hb_lnode *n, const hb_array *a);
void hb_lazy_node_free(hb_lnode *n);
const void *hb_host_copy(const hb_array *a);   // host image of a small constant array's storage (or nullptr)

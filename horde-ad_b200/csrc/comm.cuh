// comm.cuh -- what other translation units may use of the peer-memory exchange set up by hb_comm_init (comm.cu).
#pragma once
#include "common.cuh"

// Peer-mapped scalar slots of the global-objective mode: rank r writes its contribution of exchange `kind` (0 = min,
// 1 = sum) of step `e` into val[p][((e & 1) * 2 + kind) * 8 + r] of EVERY rank p, then raises flag[p][kind * 8 + r] = e;
// a rank reads its own slots once all flags show e and combines the values in rank order.  Double-buffered by step
// parity: a rank can only be one exchange ahead of the slowest peer (its next write needs that peer's next flag).
#define HB_SX_BYTES 1024
struct hb_scalar_xchg {
  int rank, world;
  double *val[8];
  unsigned long long *flag[8];
  unsigned long long *ctl;   // local: number of completed softmax-xent steps
};
// false unless hb_comm_set_global_objective(1) was called on a communicator of more than one rank
bool hb_comm_scalar_xchg(hb_scalar_xchg *out);

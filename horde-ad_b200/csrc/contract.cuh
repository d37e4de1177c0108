// contract.cuh -- the axis-group description shared by the contraction engines (contract.cu: FP64 DMMA / FFMA
// over strided views; contract_tf32.cu: the Float engine on tcgen05 tensor cores).
#pragma once
#include "common.cuh"

struct hb_cgroup {
  int rank;
  int64_t n;  // product of extents
  int64_t shape[HB_R];
  int64_t sa[HB_R], sb[HB_R], sc[HB_R];
};

// M axes: only operand a moves; N axes: only operand b moves; B: both move (batch); K: reduced.
struct hb_contractp {
  hb_cgroup M, N, B, K;
};

#ifdef __CUDACC__
__device__ __forceinline__ void hb_cg_offsets(const hb_cgroup &g, int64_t lin, int64_t &oa, int64_t &ob,
                                              int64_t &oc) {
  oa = ob = oc = 0;
#pragma unroll 1
  for (int d = g.rank - 1; d >= 0; d--) {
    int64_t q = lin / g.shape[d], r = lin - q * g.shape[d];
    oa += r * g.sa[d];
    ob += r * g.sb[d];
    oc += r * g.sc[d];
    lin = q;
  }
}
#endif

// contract_tf32.cu: Float contraction over the same strided views on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, three-term split, TMEM accumulators).  *done = 0 when the shape is left to the caller.
int hb_contract_tf32(const hb_contractp &P, const float *A, const float *B, float *C, int *done);

// gemm_dmma.cu: a group of FP64 products of one [M, N] in one launch of the DMMA kernel (hb_gemm_dmma_group).
struct hb_gemm_prod {
  const double *A; int64_t a_sm, a_sk;   // A[m, k] strides
  const double *B; int64_t b_sk, b_sn;   // B[k, n] strides
  int64_t K;
  int out;                               // index of the result the product is accumulated into
};
struct hb_gemm_res {
  double *C;                  // contiguous [M, N]
  int accum;                  // start the chain from the old contents of C
  const double *epi_bias;     // C = tanh (C + bias[m]) after the last product (NULL: none)
};
int hb_gemm_dmma_group(int nprod, const hb_gemm_prod *pr, int nres, const hb_gemm_res *res, int64_t M, int64_t N,
                       int *done);

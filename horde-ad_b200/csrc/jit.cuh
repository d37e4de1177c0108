// jit.cuh -- the kernel generator for index functions and fused maps:
// a reified program (vm.cuh) is translated to straight-line CUDA C with every
// shape, stride, dtype and bound of the call baked in as a constant, compiled
// for sm_100a by NVRTC at first use and cached by source text.  The VM
// interpreter kernels in ixprog.cu are the same semantics executed
// instruction by instruction; they remain the path for HB_JIT=0 and for boxes
// without libnvrtc (both are device code; there is no CPU path).
#pragma once
#include <cuda.h>

#include <string>

#include "vm.cuh"

// Compiles (or finds) the kernel `name` of `source`; nullptr when the
// generator is unavailable or the compilation failed (reason in hb_last_error).
CUfunction hb_jit_kernel(const std::string &source, const char *name);
// Launch on the library stream.
int hb_jit_launch(CUfunction f, unsigned grid, unsigned block, void **args);
bool hb_jit_enabled();

// shared prelude + the statements of the program body.  The body reads the
// variables v0.. (i64), the map inputs in0.. (u64 register bits), the tensor
// operands t0.. (const void*) and leaves the OUT components in o0.. (u64).
std::string hb_jit_prelude(const hb_vm_prog &P);
std::string hb_jit_body(const hb_vm_prog &P);
// `i64 vK = ...` for the row-major decomposition of `m` over `shape`
std::string hb_jit_decode(int rank, const int64_t *shape, const char *m);
// `bool inb; i64 off;` from o0.. against the target bounds / strides
std::string hb_jit_target(int prank, const int64_t *pshape, const int64_t *pstride);
// the parameter list tail `, const void *t0, ...` for the program's tensors
std::string hb_jit_tensor_params(const hb_vm_prog &P, int n_tensors);
const char *hb_jit_ctype(int dtype);

// ---- whole-kernel sources (entry point `hbj_k`) ----------------------------
// kind 0: offsets[m] = strided offset of prog(m) in the target, INT64_MIN when out of bounds
//         (i64 *offsets, tensors...)
// kind 1: rank-0 gather  out[m] = src[prog(m)] or 0   (const ET *src, ET *out, tensors...), es = sizeof(ET)
// kind 2: scatter pass 1 dest[m], counts[dest]++, collision flag   (int *dest, int *counts, int *collision, tensors...)
std::string hb_jit_index_source(const hb_vm_prog &P, int n_tensors, int kind, int shm_rank, const int64_t *shm,
                                int prank, const int64_t *pshape, const int64_t *pstride, int es);

#define HB_MAP_MAXIN 6
struct hb_mapp {
  int rank, n_in;
  int64_t shape[HB_R];
  int64_t stride[HB_MAP_MAXIN][HB_R];
  const void *ptr[HB_MAP_MAXIN];
  int dtype[HB_MAP_MAXIN];
};
// fused elementwise chain over collapsed dims; vec > 1: rank 1, strides in {0, 1}, n % vec == 0,
// pointers aligned to vec elements: one access moves vec elements (16 bytes of the widest operand)
//   (OT *out, const void *in0 .. in{n_in-1}, tensors...)
std::string hb_jit_map_source(const hb_vm_prog &P, int n_tensors, const hb_mapp &mp, int out_dtype, int64_t n, int vec);

/*
 * hordeb200.h -- C ABI of libhordeb200.so, the B200 (sm_100a) tensor backend
 * for horde-ad.
 *
 * This is the drop-in boundary: one exported function
 * per kernel family of horde-ad's tensor class hierarchy.  The Haskell side
 * (`foreign import ccall`, see INTEGRATION.md and haskell/) holds `hb_array*`
 * handles inside ForeignPtrs whose finaliser is `hb_release`.  No torch, no
 * C++ types: plain pointers, sizes and status codes only.
 *
 * Every entry point cites the reference interface it replaces, as
 * `file:line` relative to the horde-ad source tree (v0.4.0.0).
 *
 * Conventions
 *  - Every function returns an `int` status: 0 = HB_OK, non-zero = error;
 *    the message of the last error of the calling thread is available from
 *    hb_last_error().  Programmer errors (shape mismatch, bad dtype) are
 *    errors; OUT-OF-BOUNDS INDICES ARE NOT (OpsConcrete.hs:1448-1450,
 *    1527-1528): index/gather yield zeros, scatter/oneHot drop the update.
 *  - Arrays are immutable values.  Results are returned through `out`
 *    parameters as NEW handles with reference count 1.  Views share storage
 *    with their parent (storage is reference counted separately).
 *  - Shapes and strides are in ELEMENTS, int64, row-major logical order;
 *    strides may be 0 (broadcast, = sreplicate) or negative (= sreverse).
 *  - All work is enqueued on the library's stream for the current device;
 *    calls that return host data (hb_to_host, hb_scalar_to_host,
 *    hb_compare_all) synchronise that stream.
 *  - hb_release may be called from any thread (GHC finalisers).
 */
#ifndef HORDEB200_H
#define HORDEB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_MAX_RANK 12
#define HB_OK 0
#define HB_ERR_INVALID 1   /* bad argument: shape/dtype/rank mismatch   */
#define HB_ERR_CUDA 2      /* a CUDA runtime call failed                  */
#define HB_ERR_UNSUPPORTED 3
#define HB_ERR_NCCL 4

typedef struct hb_array hb_array;   /* opaque, reference counted */
typedef struct hb_ixprog hb_ixprog; /* opaque compiled index / map program */
typedef struct hb_graph hb_graph;   /* opaque captured CUDA graph */

/* Scalar types: the NumScalar whitelist of OpsConcrete.hs:1046-1058
 * (Int = Int64; CInt = Int32; Z1 and () are zero-sized and stay on the host).*/
typedef enum hb_dtype {
  HB_F64 = 0,
  HB_F32 = 1,
  HB_I64 = 2,
  HB_I32 = 3,
  HB_I16 = 4,
  HB_I8 = 5,
  HB_BOOL = 6 /* one byte, 0/1 */
} hb_dtype;

/* ---------------------------------------------------------------- lifecycle */

/* Select the CUDA device for this process and create the stream, the memory
 * pool (stream-ordered caching allocator) and scratch buffers.  Idempotent. */
int hb_init(int device);
int hb_shutdown(void);
/* Wait for all enqueued work. */
int hb_sync(void);
/* Copy the calling thread's last error message into buf (NUL-terminated). */
int hb_last_error(char *buf, size_t len);
int hb_device_count(int *count);
/* Bytes currently held by live arrays / high-water mark of the pool. */
int hb_mem_stats(uint64_t *live_bytes, uint64_t *pool_reserved_bytes);
/* Number of kernels this library has launched since hb_init (or the last
 * hb_reset_launch_count).  bench.py reports it as "gpu_launches". */
int hb_launch_count(uint64_t *count);
int hb_reset_launch_count(void);
/* The cudaStream_t the library enqueues on (for CUDA-event timing). */
int hb_stream(void **cuda_stream);

/* CUDA events on the library stream, 64 slots: record, then elapsed time
 * between two recorded slots (synchronises on the later one). */
int hb_event_record(int slot);
int hb_event_elapsed_ms(int slot_a, int slot_b, float *ms);

/* Per-launch profiler for bench.py's live roofline: between begin and end every
 * kernel launch is bracketed by events on the library stream; the report has
 * one line per launch site, "hostfunction:line count total_ms", sorted by
 * total device time.  Not used on the product path. */
int hb_prof_begin(void);
int hb_prof_end(void);
int hb_prof_report(char *buf, size_t len);

int hb_retain(hb_array *a);
/* ForeignPtr finaliser; stream-ordered free; callable from any thread. */
int hb_release(hb_array *a);

/* ----------------------------------------------------------------- transfer */

/* tsconcrete/trconcrete/txconcrete/tkconcrete (OpsConcrete.hs:120-124): inject
 * a host array (row-major, contiguous, `rank` dims) into the device carrier. */
int hb_from_host(hb_dtype dt, int rank, const int64_t *shape,
                 const void *host, hb_array **out);
/* Overwrite the storage of an existing contiguous array from (pinned) host
 * memory, asynchronously on the library stream: the per-step upload of a
 * minibatch into a resident buffer (no allocation). */
int hb_upload_into(hb_array *dst, const void *host);
/* Prefetch: the same copy on a second (copy-engine) stream so that the upload
 * of the NEXT minibatch overlaps the current step.  hb_prefetch_ready makes the
 * compute stream wait for every prefetch issued so far; hb_prefetch_consumed
 * marks the point of the compute stream after which the prefetch destinations
 * may be overwritten (later prefetches wait for it). */
int hb_prefetch_into(hb_array *dst, const void *host);
int hb_prefetch_ready(void);
int hb_prefetch_consumed(void);
/* Download in logical row-major order (materialises views); synchronises. */
int hb_to_host(const hb_array *a, void *host);
/* kfromS / sunScalar (OpsConcrete.hs:777-812): 0-d (or 1-element) array ->
 * host scalar; `value` points to storage of the array's dtype. Synchronises.*/
int hb_scalar_to_host(const hb_array *a, void *value);
/* srepl / tdefTarget / treplTarget (Ops.hs:114-121, CarriersConcrete.hs:
 * 243-282): a constant array as an all-stride-0 view of one device scalar. */
int hb_full(hb_dtype dt, int rank, const int64_t *shape, const void *value,
            hb_array **out);
/* Pinned host staging buffers for uploads/downloads. */
int hb_host_alloc(size_t bytes, void **ptr);
int hb_host_free(void *ptr);

/* ----------------------------------------------------------------- metadata */

/* sshape/rshape/xshape, tftk (OpsConcrete.hs:99-106). */
int hb_rank(const hb_array *a, int *rank);
int hb_shape(const hb_array *a, int64_t *shape /* HB_MAX_RANK */);
int hb_strides(const hb_array *a, int64_t *strides /* HB_MAX_RANK */);
int hb_dtype_of(const hb_array *a, hb_dtype *dt);
int hb_size(const hb_array *a, int64_t *n_elements);
int hb_is_contiguous(const hb_array *a, int *flag);
/* Raw device pointer of element [0,...,0] (for interop, e.g. NCCL/torch). */
int hb_device_ptr(const hb_array *a, void **ptr);

/* ------------------------------------------------- views (no kernel launch) */

/* tstranspose (OpsConcrete.hs:648-649): out.shape[i] = a.shape[perm[i]] for
 * i < nperm; remaining axes unchanged. */
int hb_transpose(const hb_array *a, int nperm, const int *perm, hb_array **out);
/* tsreplicate / tsreplicateN / tsreplicate0N (OpsConcrete.hs:243-246): new
 * LEADING axes of the given extents with stride 0. */
int hb_replicate(const hb_array *a, int nlead, const int64_t *lead_shape,
                 hb_array **out);
/* tsslice i n (OpsConcrete.hs:638-640): rows [i, i+n) of the first axis. */
int hb_slice(const hb_array *a, int64_t i, int64_t n, hb_array **out);
/* tsreverse (OpsConcrete.hs:644-646): first axis reversed (negative stride).*/
int hb_reverse(const hb_array *a, hb_array **out);
/* tsreshape (OpsConcrete.hs:651-652): view when strides allow it, else a
 * row-major copy. */
int hb_reshape(const hb_array *a, int rank, const int64_t *shape,
               hb_array **out);
/* tsindex with host indices (tindexZS, OpsConcrete.hs:1451-1490): index the k
 * leading axes; any index out of bounds -> zeros of the remaining shape. */
int hb_index_prefix(const hb_array *a, int k, const int64_t *ix,
                    hb_array **out);

/* -------------------------------------------------------------- materialise */

int hb_contiguous(const hb_array *a, hb_array **out);
/* tsappend (OpsConcrete.hs:641-643): concatenate along the first axis. */
int hb_append(const hb_array *a, const hb_array *b, hb_array **out);
/* tsfromVector / tfromList (OpsConcrete.hs:144-149, 960-985): stack k equal-
 * shape arrays along a new leading axis.  k = 0 is an error ("empty vector",
 * OpsConcrete.hs:129). */
int hb_stack(int k, hb_array *const *arrs, hb_array **out);
/* Flatten k arrays (row-major) into one contiguous vector: the gradient
 * bucket of the data-parallel loop (leaves of the TKProduct gradient,
 * MnistCnnShaped2.hs:30-40, in order). */
int hb_flatten_concat(int k, hb_array *const *arrs, hb_array **out);
/* tsiota (OpsConcrete.hs:536): [0..n-1] in dtype dt. */
int hb_iota(hb_dtype dt, int64_t n, hb_array **out);

/* -------------------------------------------------------------- elementwise */

/* Op codes: OpCodeNum1 / OpCode1 / OpCode2 / OpCodeIntegral2 of
 * Ast.hs:707-726 plus + * - max min (AstInterpret.hs:396-435). */
typedef enum hb_unop {
  HB_NEGATE = 0, HB_ABS, HB_SIGNUM,
  HB_RECIP, HB_EXP, HB_LOG, HB_SQRT, HB_SIN, HB_COS, HB_TAN, HB_ASIN, HB_ACOS,
  HB_ATAN, HB_SINH, HB_COSH, HB_TANH, HB_ASINH, HB_ACOSH, HB_ATANH,
  HB_UNOP_COUNT
} hb_unop;

typedef enum hb_binop {
  HB_ADD = 0, HB_SUB, HB_MUL,
  HB_DIVIDE, HB_POWER, HB_LOGBASE, HB_ATAN2,
  HB_QUOT, HB_REM,   /* array semantics: zero divisors replaced by 1
                        (CarriersConcrete.hs:40-56,140-144) */
  HB_MAX, HB_MIN,
  HB_BINOP_COUNT
} hb_binop;

/* Derived Num/Floating instances of Concrete (CarriersConcrete.hs:333-343).
 * Operands may be arbitrary strided views (incl. stride 0); the result is a
 * fresh contiguous array. */
int hb_unary(hb_unop op, const hb_array *a, hb_array **out);
int hb_binary(hb_binop op, const hb_array *a, const hb_array *b,
              hb_array **out);
/* tscast / tsfromIntegral (OpsConcrete.hs:455-532): value-preserving convert
 * to dtype dt (realToFrac / fromIntegral). */
int hb_cast(const hb_array *a, hb_dtype dt, hb_array **out);
/* tsfloor (OpsConcrete.hs:455-470): floor of a floating array into an
 * integral dtype. */
int hb_floor(const hb_array *a, hb_dtype dt, hb_array **out);
/* EqH/OrdH on whole arrays (CarriersConcrete.hs:310-331): ONE host Bool.
 * op: 0 '==', 1 '<=' (lexicographic on the row-major element sequence, the
 * Ord instance of the primitive vector), 2 '<'.  Synchronises. */
int hb_compare_all(int op, const hb_array *a, const hb_array *b, int *result);
/* taddTarget as used by DeltaEval cotangent accumulation (DeltaEval.hs:139-148,
 * 317-332): acc + c.  When `acc` is contiguous, uniquely referenced (handle
 * and storage refcount 1) the sum is written in place and *out = acc (same
 * handle, retained); otherwise a new array is produced. */
int hb_add_inplace_or_new(hb_array *acc, const hb_array *c, hb_array **out);

/* ----------------------------------------- expression programs (VM bytecode)
 *
 * One small register bytecode serves both (i) index functions of gather /
 * scatter (the closures `IxSOf target shm -> IxSOf target shp` of
 * Ops.hs:689-705, reified by applying them to symbolic variables; grammar =
 * SURVEY.md appendix B, Ast.hs:487-528,575-579,654-665) and (ii) fused
 * elementwise map/zipWith chains.  Registers are 64-bit; integer ops read them
 * as int64, real ops as double (rounded to float after every real op when the
 * program is flagged f32).
 */
typedef enum hb_vmop {
  /* integer */
  HB_VM_IVAR = 0, /* dst = loop variable a (k-th component of the shm index) */
  HB_VM_ICONST,   /* dst = imm.i */
  HB_VM_IADD, HB_VM_ISUB, HB_VM_IMUL,
  HB_VM_IQUOT, HB_VM_IREM, /* scalar quotH/remH: divisor 0 -> dividend
                              (Types.hs:394-419) */
  HB_VM_INEG, HB_VM_IABS, HB_VM_ISIGNUM, HB_VM_IMIN, HB_VM_IMAX,
  HB_VM_ILEQ, HB_VM_ILT, HB_VM_IEQ, /* -> 0/1 */
  HB_VM_BAND, HB_VM_BOR, HB_VM_BNOT,
  HB_VM_SEL,      /* dst = a ? b : c   (ifH; raw 64-bit move) */
  /* real */
  HB_VM_FCONST,   /* dst = imm.f */
  HB_VM_FADD, HB_VM_FSUB, HB_VM_FMUL, HB_VM_FDIV, HB_VM_FPOW, HB_VM_FLOGBASE,
  HB_VM_FATAN2, HB_VM_FMAX, HB_VM_FMIN,
  HB_VM_FUN1,     /* dst = unop b applied to a   (b = hb_unop) */
  HB_VM_FLEQ, HB_VM_FLT, HB_VM_FEQ, /* -> 0/1 integer */
  HB_VM_I2F,      /* fromIntegral */
  HB_VM_FLOOR,    /* floor real -> int */
  /* tensor operands */
  HB_VM_TBL,      /* dst = tensors[a][regs[0..n-1]]  (sindex0; OOB -> 0);
                     integer dtypes load as int64, float dtypes as double */
  HB_VM_ARGMAX,   /* dst = argmax over the LAST axis of tensors[a] at prefix
                     regs[0..n-1] (kargMax (t !$ ix)); first maximum wins;
                     OOB prefix -> 0 */
  HB_VM_ARGMIN,
  /* map inputs / outputs */
  HB_VM_IN,       /* dst = a-th map input at the current element */
  HB_VM_OUT,      /* output component b := register a */
  HB_VM_OP_COUNT
} hb_vmop;

typedef struct hb_insn {
  uint8_t op;  /* hb_vmop */
  uint8_t dst;
  uint8_t a, b, c;
  uint8_t n;   /* number of index registers for TBL/ARGMAX/ARGMIN */
  uint8_t pad[2];
  union {
    int64_t i;
    double f;
    uint8_t regs[8];
  } imm;
} hb_insn;

#define HB_VM_MAX_INSNS 96
#define HB_VM_MAX_REGS 32
#define HB_VM_MAX_TENSORS 6
#define HB_VM_MAX_OUT 8

/* Compile a program.  n_vars = number of loop variables (|shm|), n_out =
 * number of OUT components (|shp| for an index function, 1 for a map).
 * `tensors` are the runtime tensor operands of TBL/ARGMAX (retained by the
 * program).  The builder also classifies the program (affine / general). */
int hb_ixprog_build(const hb_insn *code, int n_insns, int n_vars, int n_out,
                    hb_array *const *tensors, int n_tensors, int f32,
                    hb_ixprog **out);
int hb_ixprog_free(hb_ixprog *p);
/* 1 if every output component is an affine function of the variables. */
int hb_ixprog_is_affine(const hb_ixprog *p, int *flag);

/* tsgather (tgatherZS, OpsConcrete.hs:321-323,1621-1665; Ops.hs:701-705):
 *   src : shp ++ shn,  prog : Ix shm -> Ix shp,   out : shm ++ shn
 *   out[i.., j..] = src[prog(i).., j..], zeros where prog(i) is out of bounds.
 * Bit-exact data movement for every dtype. */
int hb_gather(const hb_array *src, int shm_rank, const int64_t *shm,
              int shp_rank, const hb_ixprog *prog, hb_array **out);
/* tsscatter (tscatterZS, OpsConcrete.hs:318-320,1529-1619; Ops.hs:689-693):
 *   src : shm ++ shn,  prog : Ix shm -> Ix shp,   out : shp ++ shn
 *   out[prog(i).., j..] += src[i.., j..]; out-of-bounds destinations dropped.
 * Deterministic: contributions to one destination are added in row-major
 * order of the source positions, exactly the reference's accumulation order
 * (OpsConcrete.hs:1594-1601), so results are bit-identical to the oracle. */
int hb_scatter_add(const hb_array *src, int shm_rank, int shp_rank,
                   const int64_t *shp, const hb_ixprog *prog, hb_array **out);
/* tsoneHot (toneHotS, OpsConcrete.hs:1516-1525): out : shp ++ shn is zero
 * except out[ix.., :] = v; out-of-bounds ix -> all zeros. */
int hb_onehot(const hb_array *v, int shp_rank, const int64_t *shp,
              const int64_t *ix, hb_array **out);
/* Fused elementwise chain: out[e] = program(ins[0][e], ..., ins[n_in-1][e]).
 * All inputs have the same logical shape (any strides); out dtype = dt. */
int hb_fused_map(const hb_ixprog *prog, int n_in, hb_array *const *ins,
                 hb_dtype dt, hb_array **out);

/* The kernel generator behind hb_gather / hb_scatter_add / hb_fused_map:
 * general (non-affine) index programs and map programs are translated to
 * straight-line CUDA C with the call's shapes, strides, dtypes and bounds as
 * constants, compiled for sm_100a by NVRTC (libnvrtc.so.12, resolved with
 * dlopen) on first use and cached; HB_JIT=0 or a box without NVRTC runs the
 * same programs on the bytecode interpreter kernels instead (device code
 * either way).  `available`: 1 when the generator is usable on this device. */
int hb_jit_stats(int *available, uint64_t *compiled, uint64_t *cache_hits,
                 uint64_t *failed);
/* The CUDA C text the generator emits for an index program (entry `hbj_k`).
 * kind 0: offset table of prog over shm into a target of (pshape, pstride);
 * kind 1: rank-0 gather of es-byte elements; kind 2: scatter destinations and
 * collision counts.  Needs no device (used by the CPU-side tests). */
int hb_jit_index_source_text(const hb_ixprog *prog, int kind, int shm_rank,
                             const int64_t *shm, int prank,
                             const int64_t *pshape, const int64_t *pstride,
                             int es, char *buf, size_t len);
/* ... and for a map program over `rank` collapsed dims of `shape`, n_in
 * inputs with element strides strides[k * rank + d] and dtypes dtypes[k];
 * vec > 1 = that many elements per access (rank 1, strides 0/1). */
int hb_jit_map_source_text(const hb_ixprog *prog, int rank,
                           const int64_t *shape, int n_in,
                           const int64_t *strides, const int *dtypes,
                           hb_dtype out_dtype, int vec, char *buf, size_t len);
/* Compile CUDA C text to an sm_100a cubin with NVRTC without loading it
 * (works without a GPU); the compiler log lands in `log`. */
int hb_jit_compile_only(const char *source, size_t *cubin_bytes, char *log,
                        size_t log_len);

/* -------------------------------------------------------- reduce / contract */

/* tssum / tssumN (OpsConcrete.hs:215-228): sum over the n_axes LEADING axes.*/
int hb_sum_leading(const hb_array *a, int n_axes, hb_array **out);
/* tssum0 (OpsConcrete.hs:230): full sum -> 0-d array (scalar stays on the
 * device until hb_scalar_to_host). */
int hb_sum_all(const hb_array *a, hb_array **out);
/* tsdot0 (OpsConcrete.hs:232). */
int hb_dot_all(const hb_array *a, const hb_array *b, hb_array **out);
/* tsdot1In (OpsConcrete.hs:234-235): contract the innermost axis of two
 * equal-shape (possibly stride-0 broadcast) arrays.  Detects (batched) GEMM
 * structure from the strides and routes it to the tensor-core kernel. */
int hb_dot_inner(const hb_array *a, const hb_array *b, hb_array **out);
/* The fused form `ssumN @shm (sdot1In a b)` that the contracted artifacts
 * contain (TestMnistPP.hs:735; AstTraverse.hs:290-699): contract the n_axes
 * LEADING axes together with the innermost one in a single pass. */
int hb_dot_inner_sum_leading(const hb_array *a, const hb_array *b, int n_axes,
                             hb_array **out);
/* tsmatmul2 / tsmatvecmul (OpsConcrete.hs:236-242). */
int hb_matmul2(const hb_array *a, const hb_array *b, hb_array **out);
int hb_matvecmul(const hb_array *m, const hb_array *v, hb_array **out);
/* acc + a x b: the cotangent accumulation `taddTarget acc (tsmatmul2 a b)` of
 * evalRev (DeltaEval.hs:139-148,317-332 applied to the transposition rule of
 * tsmatmul2) as one call.  When `acc` is uniquely referenced, contiguous and
 * f64 and the DMMA GEMM takes the shape, the product is accumulated straight
 * into acc's storage by the GEMM epilogue and *out is acc (retained once
 * more); otherwise it equals hb_matmul2 followed by hb_add_inplace_or_new. */
int hb_matmul2_acc(hb_array *acc, const hb_array *a, const hb_array *b,
                   hb_array **out);
/* Pairs of tsmatmul2 products of one result shape in ONE launch (device
 * contraction-pass nodes for the recurrent layer `wX x + wS s` of
 * rnnMnistLayerS, example/MnistRnnShaped2.hs:55-69, and for the transposition
 * rules of its two products, DeltaEval.hs:274-718): each product alone fills
 * about one 64x64 tile per SM, a pair fills both CTA slots.
 *   hb_matmul2_sum      out  = a1 x b1 + a2 x b2   (K may differ)
 *   hb_matmul2_pair     out1 = a1 x b1, out2 = a2 x b2
 *   hb_matmul2_acc_pair out1 = acc1 + a1 x b1, out2 = acc2 + a2 x b2 (in place
 *                       under the conditions of hb_matmul2_acc)
 * K slices are combined in a fixed order (deterministic).  Operands that the
 * paired kernel does not take (not f64, no unit stride, differing
 * orientations) are computed by the single-product entry points instead. */
int hb_matmul2_sum(const hb_array *a1, const hb_array *b1, const hb_array *a2,
                   const hb_array *b2, hb_array **out);
int hb_matmul2_pair(const hb_array *a1, const hb_array *b1,
                    const hb_array *a2, const hb_array *b2, hb_array **out1,
                    hb_array **out2);
int hb_matmul2_acc_pair(hb_array *acc1, const hb_array *a1,
                        const hb_array *b1, hb_array *acc2,
                        const hb_array *a2, const hb_array *b2,
                        hb_array **out1, hb_array **out2);
/* tanh (a1 x b1 + a2 x b2 + bias[row]): the body of rnnMnistLayerS
 * (MnistRnnShaped2.hs:55-69) in one launch (bias + tanh in the epilogue of the
 * paired product); bit-identical to hb_matmul2_sum + hb_tanh_affine_fwd. */
int hb_matmul2_sum_tanh(const hb_array *a1, const hb_array *b1,
                        const hb_array *a2, const hb_array *b2,
                        const hb_array *bias, hb_array **out);
/* A group of n_prod products of one [M, N] with named results:
 *   outs[r] = f_r ((acc[r] or 0) + sum over g with res_index[g] == r of a[g] x b[g])
 * (products in order of g), f_r = tanh (. + bias[r] stretched over the columns)
 * where bias[r] is given, else the identity; acc and bias may be NULL or hold
 * NULL entries; a result has an accumulator or an epilogue, not both.  One
 * launch of the DMMA kernel when f64, n_prod <= 4, n_res <= 4 and the operands
 * share an orientation: what one step of an unrolled two-layer recurrence
 * needs, forward (the wavefront of rnnMnistTwoS, MnistRnnShaped2.hs:71-93: the
 * two layers' tanh (wX x + wS s + b) of independent time steps) and backward
 * (the transposition rules of tsmatmul2 of one sweep step, DeltaEval.hs:317-332:
 * three products into two cotangents).  acc[r] is updated in place only when
 * the caller holds its single reference.  Otherwise one product at a time.
 * Deterministic (K slices of a result are combined in a fixed order). */
int hb_matmul2_group(int n_res, hb_array *const *acc, const hb_array *const *bias,
                     int n_prod, const hb_array *const *a,
                     const hb_array *const *b, const int *res_index,
                     hb_array **outs);
/* acc + sum_t a[t] x b[t] for n products of one shape (acc may be NULL): the
 * cotangent of a weight used at every step of an unrolled loop (taddTarget
 * over the transposition rule of tsmatmul2, DeltaEval.hs:317-332), evaluated
 * once at the end of the sweep as ONE product over the concatenated reduced
 * axes (f64, n <= 32, K % 32 == 0, one stride pattern; otherwise a chain of
 * hb_matmul2_acc).  In place when `acc` is uniquely referenced.  Fixed
 * reduction order: deterministic. */
int hb_matmul2_list(int n, hb_array *const *a, hb_array *const *b, hb_array *acc,
                    hb_array **out);
/* tsargMax / tsargMin (OpsConcrete.hs:1712-1746): reduce the LAST axis,
 * result shape Init sh, dtype I64; first extremum wins. */
int hb_argmax_last(const hb_array *a, hb_array **out);
int hb_argmin_last(const hb_array *a, hb_array **out);

/* ------------------------------------------------- fused NN blocks (fast path)
 * Emitted by the device-specific contraction pass in place
 * of the gather+dot1In chains the vectoriser produces. */

/* conv2dPreservingS (CommonShapedOps.hs:136-157):
 *   K[Co,Ci,KH,KW], A[N,Ci,H,W] -> B[N,Co,H,W],
 *   B[n,o,h,w] = sum_{c,kh,kw} A[n,c,h+kh,w+kw] K[o,c,kh,kw], A = 0 outside. */
int hb_conv2d_preserving(const hb_array *K, const hb_array *A, hb_array **out);
/* conv2dPreserving_dKrn (TestConvQuickCheck.hs:322-349):
 *   dK[o,c,kh,kw] = sum_{n,h,w} dB[n,o,h,w] A[n,c,h+kh,w+kw]. */
int hb_conv2d_preserving_dkrn(const hb_array *A, const hb_array *dB, int KH,
                              int KW, hb_array **out);
/* conv2dPreserving_dInp (TestConvQuickCheck.hs:290-320):
 *   dA[n,c,h,w] = sum_{o,kh,kw} dB[n,o,h-kh,w-kw] K[o,c,kh,kw], dB = 0 outside.*/
int hb_conv2d_preserving_dinp(const hb_array *K, const hb_array *dB,
                              hb_array **out);
/* reluS (CommonShapedOps.hs:38-44): (x <= 0 ? 0 : 1) * x. */
int hb_relu(const hb_array *x, hb_array **out);
/* maxPool2dUnpaddedS (CommonShapedOps.hs:241-260) on [N,C,H,W]: value at the
 * argmax (first maximum, row-major within the window) of each ksize x ksize
 * window at stride; also returns the window-local argmax as I64 (optional). */
int hb_maxpool2d(const hb_array *x, int ksize, int stride, hb_array **out,
                 hb_array **argmax_out /* may be NULL */);
/* Adjoint of hb_maxpool2d: route dy to the argmax position of each window,
 * accumulating deterministically where windows overlap. */
int hb_maxpool2d_bwd(const hb_array *dy, const hb_array *argmax, int ksize,
                     int stride, int64_t H, int64_t W, hb_array **out);
/* The tail of convMnistLayerS (MnistCnnShaped2.hs:42-62) in one pass, for
 * non-overlapping windows (ksize == stride):
 *   out = maxPool2dUnpaddedS @k @k (reluS (pre + bias stretched over [N,.,H,W]))
 * `code` (I8, shape of out) is what the adjoint needs: argmax within the window
 * (row-major, first maximum) | 0x80 when the selected pre-activation is > 0.
 * bias may be NULL. */
int hb_relu_pool_fwd(const hb_array *pre, const hb_array *bias, int ksize,
                     hb_array **out, hb_array **code_out);
/* Its adjoint: d_pre [N,C,H,W] (dy routed to the argmax where relu is active,
 * zero elsewhere) and, optionally, dbias[c] = sum over n, h, w of d_pre
 * (deterministic two-stage reduction). */
int hb_relu_pool_bwd(const hb_array *dy, const hb_array *code, int ksize,
                     int64_t H, int64_t W, hb_array **dpre_out,
                     hb_array **dbias_out /* may be NULL */);
/* A whole convMnistLayerS (MnistCnnShaped2.hs:42-62) as one node of the device
 * contraction pass:  out = maxPool k k (reluS (conv2dPreservingS K A + bias)).
 * For 2x2/2 windows in Double the bias add, relu and pooling run in the
 * epilogue of the convolution kernel (the pre-activation never reaches HBM);
 * `code` as in hb_relu_pool_fwd. */
int hb_conv_relu_pool_fwd(const hb_array *K, const hb_array *bias,
                          const hb_array *A, int ksize, hb_array **out,
                          hb_array **code_out);
/* Its adjoint: dK, dbias (may be NULL) and dA (NULL when the input is not
 * differentiated: then the full-size cotangent is never materialised, the dKrn
 * kernel expands (dy, code) inside shared memory). */
int hb_conv_relu_pool_bwd(const hb_array *K, const hb_array *A,
                          const hb_array *dy, const hb_array *code, int ksize,
                          hb_array **dK_out, hb_array **dbias_out,
                          hb_array **dA_out);
/* lossSoftMaxCrossEntropyS (CommonShapedOps.hs:89-111) over the WHOLE tensor
 * (the reference normalises over all elements), scaled by `scale`:
 *   loss = scale * -(log softmax . expected);  dlogits = scale*(softmax-expected).
 * loss_out is a 0-d array; dlogits_out may be NULL. */
int hb_softmax_xent(const hb_array *logits, const hb_array *expected,
                    double scale, hb_array **loss_out, hb_array **dlogits_out);

/* The body of rnnMnistLayerS (example/MnistRnnShaped2.hs:55-69) behind its two
 * matmuls, as one node of the device contraction pass:
 *   y = tanh ((u + v) + bias stretched over the batch),  u, v : [width, batch]
 * (bias [width], may be NULL; v may be NULL when u already holds the sum, see
 * hb_matmul2_sum), and its adjoint
 *   dz = (1 - y*y) * c,   dbias[i] = sum_j dz[i,j]   (dbias_out may be NULL)
 * -- dz is the cotangent of both u and v.  Same association as the generic
 * tanh rule; the row sums are reduced in a fixed order. */
int hb_tanh_affine_fwd(const hb_array *u, const hb_array *v,
                       const hb_array *bias, hb_array **y_out);
int hb_tanh_affine_bwd(const hb_array *y, const hb_array *c, hb_array **dz_out,
                       hb_array **dbias_out);
/* The same adjoint with the bias cotangent ACCUMULATED (the taddTarget of an
 * unrolled loop's shared bias, DeltaEval.hs:496-511, fused into the pass that
 * produces the summand): *acc_out = dbias_acc + row sums.  Value semantics like
 * hb_matmul2_acc: dbias_acc's storage is reused only when the caller holds the
 * single reference to it. */
int hb_tanh_affine_bwd_acc(const hb_array *y, const hb_array *c,
                           hb_array *dbias_acc, hb_array **dz_out,
                           hb_array **acc_out);

/* Device twin of the MNIST input pipeline (example/MnistData.hs:134-171,
 * SURVEY.md 8f-3).  `pixels` [n,H,W] and `labels` [n] hold the raw IDX bytes
 * (dtype I8, read as unsigned) resident on the device; the minibatch
 *   glyphs[b,h,w] = fromIntegral pixels[id_b,h,w] / 255      (readMnistData)
 *   targets[b,c]  = if c == labels[id_b] then 1 else 0
 * with id_b = order[start + b] (order: I32 or I64 vector, the epoch's shuffle;
 * NULL = identity) is written INTO glyphs_dst [batch,H,W] / targets_dst
 * [batch,classes] (F64 or F32, contiguous): mkMnistDataBatchS without any
 * host<->device tensor traffic.  Sample ids outside [0,n) give an all-zero
 * glyph and target (the out-of-bounds rule of gather). */
int hb_mnist_batch_into(const hb_array *pixels, const hb_array *labels,
                        const hb_array *order, int64_t start,
                        hb_array *glyphs_dst, hb_array *targets_dst);

/* updateWithGradient (OptimizerTools.hs:20-51): p := p - gamma * g, in place.*/
int hb_sgd_step(hb_array *p, const hb_array *g, double gamma);
/* updateWithGradientAdam (OptimizerTools.hs:134-232, defaults :83-89): one
 * fused in-place pass over p, m, v.  t = step count AFTER increment. */
int hb_adam_step(hb_array *p, hb_array *m, hb_array *v, const hb_array *g,
                 double alpha, double beta1, double beta2, double epsilon,
                 int64_t t);

/* tmapAccumL specialisations.  `sfold (*) 1 x` and its
 * reverse derivative (Ops.hs:188-210, OpsConcrete.hs:990-1019, DeltaEval.hs:
 * 351-389): prod_out (0-d) = prod x_i;  grad_out[i] = dt * prod_{j != i} x_j
 * via exclusive prefix/suffix product scans (no division). */
int hb_prod_fold_grad(const hb_array *x, double dt, hb_array **prod_out,
                      hb_array **grad_out);

/* ------------------------------------------------------------- CUDA graphs
 * Whole-artifact capture: everything enqueued between begin
 * and end is recorded once and replayed with hb_graph_launch. */
int hb_graph_begin(void);
int hb_graph_end(hb_graph **out);
int hb_graph_launch(hb_graph *g);
int hb_graph_free(hb_graph *g);

/* ------------------------------------------- lazy programs + contraction pass
 * The device-specific contraction pass (the reference's sibling
 * is contractAst, AstTraverse.hs:290-699, which anticipates "one such pass per
 * backend").  horde-ad interprets a gradient artifact -- a fixed, shape-specific
 * op program (README.md:174-184, AstInterpret.hs:69-366) -- by calling one class
 * method per AST node.  Between hb_lazy_begin and hb_lazy_end those calls are
 * RECORDED instead of launched: every compute entry point returns a lazy array
 * (shape, dtype and views work as usual; it has no value yet) and appends a node
 * to the program.  hb_lazy_end runs the rewriter over the recorded DAG:
 *   - chains of affine sgather (i + j windows, constant window tables), views,
 *     broadcast `*`, ssum/ssumN, sdot1In and affine sscatter are composed
 *     symbolically into ONE index-affine contraction and matched against
 *     conv2dPreservingS / its dKrn / its dInp (CommonShapedOps.hs:136-157,
 *     TestConvQuickCheck.hs:290-349)             -> hb_conv2d_preserving*;
 *   - the [0.0,1.0]-table gather with `ifH (0 <=. negate x) 0 1` times x  -> relu;
 *     sgather at `kargMax` of the window tensor   -> max-pool; their adjoint
 *     scatters -> hb_relu_pool_bwd; conv + bias + relu + pool -> hb_conv_relu_pool_*;
 *   - dead nodes are dropped, taddTarget runs in place where its operand dies.
 * hb_lazy_run replays the rewritten program (eagerly, or inside
 * hb_graph_begin/end to bake it into a CUDA graph); the handles passed as `outs`
 * hold the results after each run.  Values of lazy arrays other than `outs` are
 * never materialised.  Entry points that need a value on the host
 * (hb_to_host, hb_scalar_to_host, hb_compare_all) are refused while recording. */
typedef struct hb_lazy hb_lazy;
#define HB_LAZY_REWRITE 1u      /* run the contraction pass (0: replay as recorded) */
#define HB_LAZY_NO_FUSE_LAYER 2u /* keep conv / relu+pool as separate nodes */
int hb_lazy_begin(void);
int hb_lazy_end(int n_out, hb_array *const *outs, unsigned flags, hb_lazy **prog);
int hb_lazy_abort(void);
int hb_lazy_run(hb_lazy *prog);
int hb_lazy_free(hb_lazy *prog);
/* nodes recorded / nodes left after the pass */
int hb_lazy_stats(const hb_lazy *prog, int *recorded, int *live);
/* Text listing of the rewritten program and of the pass's decisions. */
int hb_lazy_dump(const hb_lazy *prog, char *buf, size_t len);
/* Initialise WITHOUT a device: constants live in host memory and compute entry
 * points can only be recorded, never run.  For exercising the contraction pass
 * on a box without a GPU (the CPU-side tests); mutually exclusive with hb_init. */
int hb_init_trace_only(void);

/* -------------------------------------------------------------- distributed
 * Data-parallel gradient exchange: NCCL over NVLink/NVSwitch,
 * resolved at run time with dlopen("libnccl.so.2"). */
#define HB_UNIQUE_ID_BYTES 128
int hb_comm_unique_id(void *id_bytes /* HB_UNIQUE_ID_BYTES */);
int hb_comm_init(int rank, int world, const void *id_bytes);
/* In-place sum over all ranks of a contiguous array, then scale by `scale`
 * (1/world for the mean of per-shard mean gradients). */
int hb_allreduce_sum(hb_array *bucket, double scale);
/* The exchange AND the optimizer step in ONE kernel (no NCCL on the data path):
 * every rank publishes `bucket` in a peer-mapped buffer (cudaIpc over NVLink /
 * NVSwitch, set up by hb_comm_init), pulls the peers' buckets and adds them in
 * rank order 0..G-1 -- the same order on every rank, so the result is
 * bit-identical across ranks and runs -- scales by `scale`, writes the reduced
 * bucket back in place, and applies updateWithGradientAdam
 * (OptimizerTools.hs:134-232) to p, m, v with g = bucket[0 : n_params].
 * `t_counter` (0-d I64 device array) is Adam's step count: incremented on the
 * device, so the call can be replayed from a CUDA graph; `betaOne ^ t` is
 * computed as GHC's (^) does.  World size 1 (or no communicator): scale + Adam. */
int hb_allreduce_adam_step(hb_array *bucket, int64_t n_params, hb_array *p,
                           hb_array *m, hb_array *v, hb_array *t_counter,
                           double alpha, double beta1, double beta2,
                           double epsilon, double scale);
/* Global-objective mode.  lossSoftMaxCrossEntropyS (CommonShapedOps.hs:89-111)
 * normalises its softmax over the WHOLE [labels, batch] tensor, so a batch
 * sharded over G ranks reproduces the reference's one-device step only if
 * `minimum u` and `sum (exp (u - minimum u))` range over all shards.  With
 * on != 0 (and world > 1) every hb_softmax_xent call exchanges these two
 * scalars with the peer ranks inside its kernel (peer-mapped slots over NVLink,
 * combined in rank order: identical on every rank); the caller then passes the
 * GLOBAL 1/batch as `scale` and sums (scale 1.0) instead of averaging the
 * gradient buckets.  Every rank must make the same sequence of hb_softmax_xent
 * calls.  Off by default (each rank optimises its own shard's objective). */
int hb_comm_set_global_objective(int on);
int hb_comm_destroy(void);

/* -------------------------------------------------------- micro-benchmarks
 * Measured peaks of the pipes the roofline is quoted against (FP64 FMA, FP64
 * DMMA, FP32 FMA, HBM copy); each returns achieved TFLOP/s or GB/s. */
int hb_microbench(int which, double *result);

#ifdef __cplusplus
}
#endif
#endif /* HORDEB200_H */

// jit.cu -- the kernel generator for index functions (tsgather / tsscatter,
// Ops.hs:689-705; grammar SURVEY.md appendix B) and fused elementwise chains:
// program -> CUDA C (straight-line, all shapes / strides / dtypes constant)
// -> NVRTC (sm_100a) -> module cache.  See jit.cuh.
#include "jit.cuh"

#include <dlfcn.h>
#include <nvrtc.h>
#include <stdarg.h>

#include <mutex>
#include <unordered_map>

namespace {

struct Api {
  bool ok = false;
  std::string why;
  // NVRTC (dlopen: the library is not a link-time dependency)
  nvrtcResult (*create)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *);
  nvrtcResult (*compile)(nvrtcProgram, int, const char *const *);
  nvrtcResult (*cubin_size)(nvrtcProgram, size_t *);
  nvrtcResult (*cubin)(nvrtcProgram, char *);
  nvrtcResult (*log_size)(nvrtcProgram, size_t *);
  nvrtcResult (*log)(nvrtcProgram, char *);
  nvrtcResult (*destroy)(nvrtcProgram *);
  // driver entry points (through the runtime; libcuda is not linked)
  CUresult (*mod_load)(CUmodule *, const void *);
  CUresult (*mod_fn)(CUfunction *, CUmodule, const char *);
  CUresult (*launch)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                     void **, void **);
};

template <typename F>
bool drv(const char *name, F *out) {
  void *f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return false;
  }
  *out = (F)f;
  return true;
}

Api &api() {
  static Api a = [] {
    Api x;
    const char *e = getenv("HB_JIT");
    if (e && e[0] == '0') { x.why = "HB_JIT=0"; return x; }
    void *h = nullptr;
    for (const char *n : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
      h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (h) break;
    }
    if (!h) { x.why = "libnvrtc.so.12 not found"; return x; }
#define SYM(field, name)                                  \
  *(void **)&x.field = dlsym(h, name);                    \
  if (!x.field) { x.why = "libnvrtc: missing " name; return x; }
    SYM(create, "nvrtcCreateProgram");
    SYM(compile, "nvrtcCompileProgram");
    SYM(cubin_size, "nvrtcGetCUBINSize");
    SYM(cubin, "nvrtcGetCUBIN");
    SYM(log_size, "nvrtcGetProgramLogSize");
    SYM(log, "nvrtcGetProgramLog");
    SYM(destroy, "nvrtcDestroyProgram");
#undef SYM
    if (!drv("cuModuleLoadData", &x.mod_load) || !drv("cuModuleGetFunction", &x.mod_fn) ||
        !drv("cuLaunchKernel", &x.launch)) {
      x.why = "driver entry points unavailable";
      return x;
    }
    x.ok = true;
    return x;
  }();
  return a;
}

std::mutex g_mu;
std::unordered_map<std::string, CUfunction> g_cache;  // source text -> kernel (nullptr = failed before)
uint64_t g_compiled = 0, g_hits = 0, g_failed = 0;

void emit(std::string &s, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  s += buf;
}

const char *un_name(int op, bool f32) {
  switch (op) {
    case HB_RECIP: return nullptr;  // handled inline
    case HB_EXP: return f32 ? "expf" : "exp";
    case HB_LOG: return f32 ? "logf" : "log";
    case HB_SQRT: return f32 ? "sqrtf" : "sqrt";
    case HB_SIN: return f32 ? "sinf" : "sin";
    case HB_COS: return f32 ? "cosf" : "cos";
    case HB_TAN: return f32 ? "tanf" : "tan";
    case HB_ASIN: return f32 ? "asinf" : "asin";
    case HB_ACOS: return f32 ? "acosf" : "acos";
    case HB_ATAN: return f32 ? "atanf" : "atan";
    case HB_SINH: return f32 ? "sinhf" : "sinh";
    case HB_COSH: return f32 ? "coshf" : "cosh";
    case HB_TANH: return f32 ? "tanhf" : "tanh";
    case HB_ASINH: return f32 ? "asinhf" : "asinh";
    case HB_ACOSH: return f32 ? "acoshf" : "acosh";
    case HB_ATANH: return f32 ? "atanhf" : "atanh";
    default: return nullptr;
  }
}

const char *ld_name(int dtype) {
  switch (dtype) {
    case HB_F64: return "ld_f64";
    case HB_F32: return "ld_f32";
    case HB_I64: return "ld_i64";
    case HB_I32: return "ld_i32";
    case HB_I16: return "ld_i16";
    case HB_I8: return "ld_i8";
    default: return "ld_u8";
  }
}

// `i64 off; bool inb;` of tensor operand t at index registers regs[0..n)
void emit_tensor_index(std::string &s, const hb_vm_tensor &t, const uint8_t *regs, int n) {
  s += "i64 off = 0; bool inb = true; i64 ix;\n";
  for (int k = 0; k < n; k++)
    emit(s, "    ix = (i64)r%d; inb = inb && ix >= 0 && ix < %lldLL; off += ix * %lldLL;\n", (int)regs[k],
         (long long)t.shape[k], (long long)t.stride[k]);
}

}  // namespace

bool hb_jit_enabled() { return api().ok; }

const char *hb_jit_ctype(int dtype) {
  switch (dtype) {
    case HB_F64: return "double";
    case HB_F32: return "float";
    case HB_I64: return "long long";
    case HB_I32: return "int";
    case HB_I16: return "short";
    case HB_I8: return "signed char";
    default: return "unsigned char";
  }
}

std::string hb_jit_prelude(const hb_vm_prog &P) {
  std::string s;
  s += "typedef long long i64;\ntypedef unsigned long long u64;\n";
  s += "#define I64MIN (-9223372036854775807LL - 1)\n";
  emit(s, "#define F32 %d\n", P.f32 ? 1 : 0);
  s += "__device__ __forceinline__ double bf(u64 b) { return __longlong_as_double((i64)b); }\n"
       "__device__ __forceinline__ u64 fb(double f) { return (u64)__double_as_longlong(f); }\n"
       "__device__ __forceinline__ u64 wf(double v) { if (F32) v = (double)(float)v; return fb(v); }\n"
       "__device__ __forceinline__ u64 ld_f64(const void *p, i64 o) { return fb(((const double *)p)[o]); }\n"
       "__device__ __forceinline__ u64 ld_f32(const void *p, i64 o) { return fb((double)((const float *)p)[o]); }\n"
       "__device__ __forceinline__ u64 ld_i64(const void *p, i64 o) { return (u64)((const i64 *)p)[o]; }\n"
       "__device__ __forceinline__ u64 ld_i32(const void *p, i64 o) { return (u64)(i64)((const int *)p)[o]; }\n"
       "__device__ __forceinline__ u64 ld_i16(const void *p, i64 o) { return (u64)(i64)((const short *)p)[o]; }\n"
       "__device__ __forceinline__ u64 ld_i8(const void *p, i64 o) { return (u64)(i64)((const signed char *)p)[o]; }\n"
       "__device__ __forceinline__ u64 ld_u8(const void *p, i64 o) { return (u64)((const unsigned char *)p)[o]; }\n";
  return s;
}

std::string hb_jit_tensor_params(const hb_vm_prog &P, int n_tensors) {
  (void)P;
  std::string s;
  for (int t = 0; t < n_tensors; t++) emit(s, ", const void *__restrict__ t%d", t);
  return s;
}

std::string hb_jit_decode(int rank, const int64_t *shape, const char *m) {
  std::string s;
  int64_t total = 1;
  for (int d = 0; d < rank; d++) total *= shape[d];
  const char *it = total < (1ll << 31) ? "unsigned" : "i64";
  emit(s, "  %s rem_ = (%s)(%s);\n", it, it, m);
  for (int d = rank - 1; d >= 0; d--) {
    if (d == 0) emit(s, "  const i64 v0 = (i64)rem_;\n");
    else
      emit(s, "  const i64 v%d = (i64)(rem_ %% (%s)%lld); rem_ /= (%s)%lld;\n", d, it, (long long)shape[d], it,
           (long long)shape[d]);
  }
  return s;
}

std::string hb_jit_target(int prank, const int64_t *pshape, const int64_t *pstride) {
  std::string s = "  bool inb = true; i64 off = 0;\n";
  for (int c = 0; c < prank; c++)
    emit(s, "  { i64 ix = (i64)o%d; inb = inb && ix >= 0 && ix < %lldLL; off += ix * %lldLL; }\n", c,
         (long long)pshape[c], (long long)pstride[c]);
  return s;
}

std::string hb_jit_body(const hb_vm_prog &P) {
  std::string s;
  bool used[HB_VM_MAX_REGS] = {};
  for (int pc = 0; pc < P.n_insns; pc++) used[P.code[pc].dst] = true;
  for (int r = 0; r < HB_VM_MAX_REGS; r++)
    if (used[r]) emit(s, "  u64 r%d = 0;\n", r);
  // registers that are read but never written stay 0 like the VM's
  auto R = [&](int r) {
    static thread_local char b[8][16];
    static thread_local int k = 0;
    k = (k + 1) & 7;
    if (used[r]) snprintf(b[k], sizeof b[k], "r%d", r);
    else snprintf(b[k], sizeof b[k], "0ull");
    return b[k];
  };
  for (int c = 0; c < P.n_out; c++) emit(s, "  u64 o%d = 0;\n", c);
  const bool f32 = P.f32 != 0;
  for (int pc = 0; pc < P.n_insns; pc++) {
    const hb_insn &I = P.code[pc];
    const int d = I.dst;
#define RI(x) R(x)
    switch (I.op) {
      case HB_VM_IVAR: emit(s, "  r%d = (u64)v%d;\n", d, (int)I.a); break;
      case HB_VM_ICONST:
        if (I.imm.i == INT64_MIN) emit(s, "  r%d = (u64)I64MIN;\n", d);
        else emit(s, "  r%d = (u64)%lldLL;\n", d, (long long)I.imm.i);
        break;
      case HB_VM_IADD: emit(s, "  r%d = (u64)((i64)%s + (i64)%s);\n", d, RI(I.a), RI(I.b)); break;
      case HB_VM_ISUB: emit(s, "  r%d = (u64)((i64)%s - (i64)%s);\n", d, RI(I.a), RI(I.b)); break;
      case HB_VM_IMUL: emit(s, "  r%d = (u64)((i64)%s * (i64)%s);\n", d, RI(I.a), RI(I.b)); break;
      case HB_VM_IQUOT:
        emit(s, "  { i64 a = (i64)%s, b = (i64)%s; r%d = (u64)(b == 0 ? a : a / b); }\n", RI(I.a), RI(I.b), d);
        break;
      case HB_VM_IREM:
        emit(s, "  { i64 a = (i64)%s, b = (i64)%s; r%d = (u64)(b == 0 ? a : a %% b); }\n", RI(I.a), RI(I.b), d);
        break;
      case HB_VM_INEG: emit(s, "  r%d = (u64)(-(i64)%s);\n", d, RI(I.a)); break;
      case HB_VM_IABS: emit(s, "  { i64 a = (i64)%s; r%d = (u64)(a < 0 ? -a : a); }\n", RI(I.a), d); break;
      case HB_VM_ISIGNUM: emit(s, "  { i64 a = (i64)%s; r%d = (u64)(i64)((a > 0) - (a < 0)); }\n", RI(I.a), d); break;
      case HB_VM_IMIN: emit(s, "  { i64 a = (i64)%s, b = (i64)%s; r%d = (u64)(b < a ? b : a); }\n", RI(I.a), RI(I.b), d); break;
      case HB_VM_IMAX: emit(s, "  { i64 a = (i64)%s, b = (i64)%s; r%d = (u64)(a < b ? b : a); }\n", RI(I.a), RI(I.b), d); break;
      case HB_VM_ILEQ: emit(s, "  r%d = (i64)%s <= (i64)%s;\n", d, RI(I.a), RI(I.b)); break;
      case HB_VM_ILT: emit(s, "  r%d = (i64)%s < (i64)%s;\n", d, RI(I.a), RI(I.b)); break;
      case HB_VM_IEQ: emit(s, "  r%d = (i64)%s == (i64)%s;\n", d, RI(I.a), RI(I.b)); break;
      case HB_VM_BAND: emit(s, "  r%d = (%s != 0) && (%s != 0);\n", d, RI(I.a), RI(I.b)); break;
      case HB_VM_BOR: emit(s, "  r%d = (%s != 0) || (%s != 0);\n", d, RI(I.a), RI(I.b)); break;
      case HB_VM_BNOT: emit(s, "  r%d = %s == 0;\n", d, RI(I.a)); break;
      case HB_VM_SEL: emit(s, "  r%d = %s != 0 ? %s : %s;\n", d, RI(I.a), RI(I.b), RI(I.c)); break;
      case HB_VM_FCONST: {
        uint64_t bits;
        memcpy(&bits, &I.imm.f, 8);
        emit(s, "  r%d = wf(bf(0x%llxull));\n", d, (unsigned long long)bits);
      } break;
      case HB_VM_FADD: emit(s, "  r%d = wf(bf(%s) + bf(%s));\n", d, RI(I.a), RI(I.b)); break;
      case HB_VM_FSUB: emit(s, "  r%d = wf(bf(%s) - bf(%s));\n", d, RI(I.a), RI(I.b)); break;
      case HB_VM_FMUL: emit(s, "  r%d = wf(bf(%s) * bf(%s));\n", d, RI(I.a), RI(I.b)); break;
      case HB_VM_FDIV: emit(s, "  r%d = wf(bf(%s) / bf(%s));\n", d, RI(I.a), RI(I.b)); break;
      case HB_VM_FPOW: emit(s, "  r%d = wf(pow(bf(%s), bf(%s)));\n", d, RI(I.a), RI(I.b)); break;
      case HB_VM_FLOGBASE: emit(s, "  r%d = wf(log(bf(%s)) / log(bf(%s)));\n", d, RI(I.b), RI(I.a)); break;
      case HB_VM_FATAN2: emit(s, "  r%d = wf(atan2(bf(%s), bf(%s)));\n", d, RI(I.a), RI(I.b)); break;
      case HB_VM_FMAX: emit(s, "  { double a = bf(%s), b = bf(%s); r%d = wf(a < b ? b : a); }\n", RI(I.a), RI(I.b), d); break;
      case HB_VM_FMIN: emit(s, "  { double a = bf(%s), b = bf(%s); r%d = wf(b < a ? b : a); }\n", RI(I.a), RI(I.b), d); break;
      case HB_VM_FUN1: {
        const char *ty = f32 ? "float" : "double";
        const char *fn = un_name(I.b, f32);
        emit(s, "  { %s x = (%s)bf(%s); %s y;", ty, ty, RI(I.a), ty);
        if (I.b == HB_NEGATE) s += " y = -x;";
        else if (I.b == HB_ABS) s += f32 ? " y = fabsf(x);" : " y = fabs(x);";
        else if (I.b == HB_SIGNUM) emit(s, " y = x > (%s)0 ? (%s)1 : (x < (%s)0 ? (%s)-1 : x);", ty, ty, ty, ty);
        else if (I.b == HB_RECIP) emit(s, " y = (%s)1 / x;", ty);
        else if (fn) emit(s, " y = %s(x);", fn);
        else s += " y = x;";
        emit(s, " r%d = wf((double)y); }\n", d);
      } break;
      case HB_VM_FLEQ: emit(s, "  r%d = bf(%s) <= bf(%s);\n", d, RI(I.a), RI(I.b)); break;
      case HB_VM_FLT: emit(s, "  r%d = bf(%s) < bf(%s);\n", d, RI(I.a), RI(I.b)); break;
      case HB_VM_FEQ: emit(s, "  r%d = bf(%s) == bf(%s);\n", d, RI(I.a), RI(I.b)); break;
      case HB_VM_I2F: emit(s, "  r%d = wf((double)(i64)%s);\n", d, RI(I.a)); break;
      case HB_VM_FLOOR: emit(s, "  r%d = (u64)(i64)floor(bf(%s));\n", d, RI(I.a)); break;
      case HB_VM_TBL: {
        const hb_vm_tensor &t = P.tensors[I.a];
        s += "  { ";
        emit_tensor_index(s, t, I.imm.regs, I.n);
        emit(s, "    r%d = inb ? %s(t%d, off) : 0ull; }\n", d, ld_name(t.dtype), (int)I.a);
      } break;
      case HB_VM_ARGMAX:
      case HB_VM_ARGMIN: {
        const hb_vm_tensor &t = P.tensors[I.a];
        const bool fl = t.dtype == HB_F64 || t.dtype == HB_F32;
        const char *cmp = I.op == HB_VM_ARGMAX ? ">" : "<";
        const char *cv = fl ? "bf" : "(i64)";
        const char *ct = fl ? "double" : "i64";
        s += "  { ";
        emit_tensor_index(s, t, I.imm.regs, I.n);
        emit(s, "    i64 best = 0;\n    if (inb) {\n      %s bv = %s(%s(t%d, off));\n", ct, cv, ld_name(t.dtype), (int)I.a);
        if (t.shape[I.n] <= 16) s += "#pragma unroll\n";
        emit(s, "      for (i64 i = 1; i < %lldLL; i++) { %s v = %s(%s(t%d, off + i * %lldLL)); if (v %s bv) { bv = v; best = i; } }\n",
             (long long)t.shape[I.n], ct, cv, ld_name(t.dtype), (int)I.a, (long long)t.stride[I.n], cmp);
        emit(s, "    }\n    r%d = (u64)best; }\n", d);
      } break;
      case HB_VM_IN: emit(s, "  r%d = in%d;\n", d, (int)I.a); break;
      case HB_VM_OUT: emit(s, "  o%d = %s;\n", (int)I.b, RI(I.a)); break;
      default: break;
    }
#undef RI
  }
  return s;
}

std::string hb_jit_index_source(const hb_vm_prog &P, int n_tensors, int kind, int shm_rank, const int64_t *shm,
                                int prank, const int64_t *pshape, const int64_t *pstride, int es) {
  std::string s = hb_jit_prelude(P);
  int64_t M = 1;
  for (int d = 0; d < shm_rank; d++) M *= shm[d];
  const char *et = es == 8 ? "u64" : es == 4 ? "unsigned" : es == 2 ? "unsigned short" : "unsigned char";
  s += "extern \"C\" __global__ void __launch_bounds__(256) hbj_k(";
  if (kind == 0) s += "i64 *__restrict__ offsets";
  else if (kind == 1) emit(s, "const %s *__restrict__ src, %s *__restrict__ out", et, et);
  else s += "int *__restrict__ dest, int *__restrict__ counts, int *__restrict__ collision";
  s += hb_jit_tensor_params(P, n_tensors);
  s += ") {\n  const i64 m = (i64)blockIdx.x * 256 + threadIdx.x;\n";
  emit(s, "  if (m >= %lldLL) return;\n", (long long)M);
  s += hb_jit_decode(shm_rank, shm, "m");
  s += hb_jit_body(P);
  s += hb_jit_target(prank, pshape, pstride);
  if (kind == 0) s += "  offsets[m] = inb ? off : I64MIN;\n";
  else if (kind == 1) emit(s, "  out[m] = inb ? src[off] : (%s)0;\n", et);
  else
    s += "  const int d = inb ? (int)off : -1;\n  dest[m] = d;\n"
         "  if (d >= 0) { int old = atomicAdd(&counts[d], 1); if (old > 0) *collision = 1; }\n";
  s += "}\n";
  return s;
}

std::string hb_jit_map_source(const hb_vm_prog &P, int n_tensors, const hb_mapp &mp, int out_dtype, int64_t n, int vec) {
  std::string s = hb_jit_prelude(P);
  s += "__device__ __forceinline__ u64 cv(double x) { return fb(x); }\n"
       "__device__ __forceinline__ u64 cv(float x) { return fb((double)x); }\n"
       "__device__ __forceinline__ u64 cv(i64 x) { return (u64)x; }\n"
       "__device__ __forceinline__ u64 cv(int x) { return (u64)(i64)x; }\n"
       "__device__ __forceinline__ u64 cv(short x) { return (u64)(i64)x; }\n"
       "__device__ __forceinline__ u64 cv(signed char x) { return (u64)(i64)x; }\n"
       "__device__ __forceinline__ u64 cv(unsigned char x) { return (u64)x; }\n";
  const char *ot = hb_jit_ctype(out_dtype);
  auto outcv = [&](const char *o) {
    std::string r;
    if (out_dtype == HB_F64) emit(r, "bf(%s)", o);
    else if (out_dtype == HB_F32) emit(r, "(float)bf(%s)", o);
    else emit(r, "(%s)(i64)%s", ot, o);
    return r;
  };
  for (int k = 0; k < mp.n_in; k++)
    emit(s, "struct alignas(%d) V%d { %s e[%d]; };\n", (int)hb_dtype_size((hb_dtype)mp.dtype[k]) * vec, k,
         hb_jit_ctype(mp.dtype[k]), vec);
  emit(s, "struct alignas(%d) VO { %s e[%d]; };\n", (int)hb_dtype_size((hb_dtype)out_dtype) * vec, ot, vec);
  emit(s, "extern \"C\" __global__ void __launch_bounds__(256) hbj_k(%s *__restrict__ out", ot);
  for (int k = 0; k < mp.n_in; k++) emit(s, ", const void *__restrict__ p%d", k);
  s += hb_jit_tensor_params(P, n_tensors);
  s += ") {\n  const i64 gs = (i64)gridDim.x * 256;\n";
  if (vec > 1) {
    emit(s, "  for (i64 g = (i64)blockIdx.x * 256 + threadIdx.x; g < %lldLL; g += gs) {\n", (long long)(n / vec));
    for (int k = 0; k < mp.n_in; k++) {
      if (mp.stride[k][0] == 0) emit(s, "    const %s s%d = ((const %s *)p%d)[0];\n", hb_jit_ctype(mp.dtype[k]), k, hb_jit_ctype(mp.dtype[k]), k);
      else emit(s, "    const V%d a%d = ((const V%d *)p%d)[g];\n", k, k, k, k);
    }
    s += "    VO res;\n#pragma unroll\n";
    emit(s, "    for (int e = 0; e < %d; e++) {\n      const i64 v0 = g * %d + e;\n", vec, vec);
    for (int k = 0; k < mp.n_in; k++) {
      if (mp.stride[k][0] == 0) emit(s, "      const u64 in%d = cv(s%d);\n", k, k);
      else emit(s, "      const u64 in%d = cv(a%d.e[e]);\n", k, k);
    }
    s += hb_jit_body(P);
    emit(s, "      res.e[e] = %s;\n    }\n    ((VO *)out)[g] = res;\n  }\n}\n", outcv("o0").c_str());
    return s;
  }
  emit(s, "  for (i64 lin = (i64)blockIdx.x * 256 + threadIdx.x; lin < %lldLL; lin += gs) {\n", (long long)n);
  // decompose lin over the collapsed dims (names d0.. to keep v0 = lin for the program)
  {
    const char *it = n < (1ll << 31) ? "unsigned" : "i64";
    emit(s, "    %s rem_ = (%s)lin;\n", it, it);
    for (int d = mp.rank - 1; d >= 0; d--) {
      if (d == 0) emit(s, "    const i64 d0 = (i64)rem_;\n");
      else emit(s, "    const i64 d%d = (i64)(rem_ %% (%s)%lld); rem_ /= (%s)%lld;\n", d, it, (long long)mp.shape[d], it, (long long)mp.shape[d]);
    }
  }
  for (int k = 0; k < mp.n_in; k++) {
    emit(s, "    const u64 in%d = cv(((const %s *)p%d)[0", k, hb_jit_ctype(mp.dtype[k]), k);
    for (int d = 0; d < mp.rank; d++)
      if (mp.stride[k][d] != 0) emit(s, " + d%d * %lldLL", d, (long long)mp.stride[k][d]);
    s += "]);\n";
  }
  s += "    const i64 v0 = lin;\n";
  s += hb_jit_body(P);
  emit(s, "    out[lin] = %s;\n  }\n}\n", outcv("o0").c_str());
  return s;
}

CUfunction hb_jit_kernel(const std::string &source, const char *name) {
  Api &a = api();
  if (!a.ok) return nullptr;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_cache.find(source);
  if (it != g_cache.end()) {
    g_hits++;
    return it->second;
  }
  CUfunction fn = nullptr;
  nvrtcProgram prog = nullptr;
  std::string err;
  do {
    if (a.create(&prog, source.c_str(), "hb_jit.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) { err = "nvrtcCreateProgram failed"; break; }
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-fmad=false", "-default-device", "-lineinfo"};
    nvrtcResult r = a.compile(prog, 5, opts);
    if (r != NVRTC_SUCCESS) {
      size_t n = 0;
      a.log_size(prog, &n);
      std::string log(n, '\0');
      if (n) a.log(prog, &log[0]);
      err = "NVRTC compilation failed: " + log.substr(0, 600);
      break;
    }
    size_t n = 0;
    if (a.cubin_size(prog, &n) != NVRTC_SUCCESS || n == 0) { err = "nvrtcGetCUBINSize failed"; break; }
    std::string bin(n, '\0');
    if (a.cubin(prog, &bin[0]) != NVRTC_SUCCESS) { err = "nvrtcGetCUBIN failed"; break; }
    CUmodule mod;
    CUresult cr = a.mod_load(&mod, bin.data());
    if (cr != CUDA_SUCCESS) { err = "cuModuleLoadData failed (" + std::to_string((int)cr) + ")"; break; }
    cr = a.mod_fn(&fn, mod, name);
    if (cr != CUDA_SUCCESS) { fn = nullptr; err = "cuModuleGetFunction failed"; break; }
  } while (0);
  if (prog) a.destroy(&prog);
  if (!fn) {
    g_failed++;
    hb_set_error("kernel generator: %s", err.c_str());
  } else {
    g_compiled++;
  }
  g_cache.emplace(source, fn);
  return fn;
}

int hb_jit_launch(CUfunction f, unsigned grid, unsigned block, void **args) {
  CUresult r = api().launch(f, grid, 1, 1, block, 1, 1, 0, (CUstream)g_hb.stream, args, nullptr);
  if (r != CUDA_SUCCESS) HB_FAIL(HB_ERR_CUDA, "generated kernel: cuLaunchKernel failed (%d)", (int)r);
  return HB_OK;
}

// Offline entry point of the generator (no device needed): compiles `source`
// to a cubin for sm_100a and reports the size -- the "does it build" check of
// generated kernels on a box without a GPU.
extern "C" int hb_jit_compile_only(const char *source, size_t *cubin_bytes, char *log, size_t log_len) {
  HB_REQUIRE(source, "null argument");
  Api x;
  void *h = dlopen("libnvrtc.so.12", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("/usr/local/cuda/lib64/libnvrtc.so.12", RTLD_NOW | RTLD_LOCAL);
  if (!h) HB_FAIL(HB_ERR_UNSUPPORTED, "libnvrtc.so.12 not found");
  *(void **)&x.create = dlsym(h, "nvrtcCreateProgram");
  *(void **)&x.compile = dlsym(h, "nvrtcCompileProgram");
  *(void **)&x.cubin_size = dlsym(h, "nvrtcGetCUBINSize");
  *(void **)&x.log_size = dlsym(h, "nvrtcGetProgramLogSize");
  *(void **)&x.log = dlsym(h, "nvrtcGetProgramLog");
  *(void **)&x.destroy = dlsym(h, "nvrtcDestroyProgram");
  if (!x.create || !x.compile || !x.cubin_size || !x.log_size || !x.log || !x.destroy)
    HB_FAIL(HB_ERR_UNSUPPORTED, "libnvrtc: missing symbols");
  nvrtcProgram prog = nullptr;
  if (x.create(&prog, source, "hb_jit.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS)
    HB_FAIL(HB_ERR_INVALID, "nvrtcCreateProgram failed");
  const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-fmad=false", "-default-device", "-lineinfo"};
  nvrtcResult r = x.compile(prog, 5, opts);
  size_t n = 0;
  x.log_size(prog, &n);
  if (log && log_len) {
    std::string l(n, '\0');
    if (n) x.log(prog, &l[0]);
    snprintf(log, log_len, "%s", l.c_str());
  }
  size_t bytes = 0;
  if (r == NVRTC_SUCCESS) x.cubin_size(prog, &bytes);
  x.destroy(&prog);
  if (cubin_bytes) *cubin_bytes = bytes;
  if (r != NVRTC_SUCCESS) HB_FAIL(HB_ERR_INVALID, "NVRTC compilation failed");
  return HB_OK;
}

extern "C" int hb_jit_stats(int *available, uint64_t *compiled, uint64_t *cache_hits, uint64_t *failed) {
  if (available) *available = api().ok ? 1 : 0;
  std::lock_guard<std::mutex> lk(g_mu);
  if (compiled) *compiled = g_compiled;
  if (cache_hits) *cache_hits = g_hits;
  if (failed) *failed = g_failed;
  return HB_OK;
}

static int copy_out(const std::string &s, char *buf, size_t len) {
  HB_REQUIRE(buf && len > 0, "null argument");
  HB_REQUIRE(s.size() + 1 <= len, "buffer too small: %zu bytes needed", s.size() + 1);
  memcpy(buf, s.c_str(), s.size() + 1);
  return HB_OK;
}

extern "C" int hb_jit_index_source_text(const hb_ixprog *prog, int kind, int shm_rank, const int64_t *shm, int prank,
                                        const int64_t *pshape, const int64_t *pstride, int es, char *buf,
                                        size_t len) {
  HB_REQUIRE(prog, "null argument");
  HB_REQUIRE(kind >= 0 && kind <= 2 && shm_rank == prog->vm.n_vars && prank == prog->vm.n_out, "bad arguments");
  return copy_out(hb_jit_index_source(prog->vm, prog->n_tensors, kind, shm_rank, shm, prank, pshape, pstride, es), buf,
                  len);
}

extern "C" int hb_jit_map_source_text(const hb_ixprog *prog, int rank, const int64_t *shape, int n_in,
                                      const int64_t *strides, const int *dtypes, hb_dtype out_dtype, int vec,
                                      char *buf, size_t len) {
  HB_REQUIRE(prog && rank >= 0 && rank <= HB_R && n_in >= 1 && n_in <= HB_MAP_MAXIN, "bad arguments");
  hb_mapp mp;
  memset(&mp, 0, sizeof mp);
  mp.rank = rank;
  mp.n_in = n_in;
  int64_t n = 1;
  for (int d = 0; d < rank; d++) { mp.shape[d] = shape[d]; n *= shape[d]; }
  for (int k = 0; k < n_in; k++) {
    mp.dtype[k] = dtypes[k];
    for (int d = 0; d < rank; d++) mp.stride[k][d] = strides[k * rank + d];
  }
  return copy_out(hb_jit_map_source(prog->vm, prog->n_tensors, mp, out_dtype, n, vec), buf, len);
}

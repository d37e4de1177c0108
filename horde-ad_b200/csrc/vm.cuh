// vm.cuh -- the expression VM executed inside gather / scatter / fused-map
// kernels.  Programs are the reified index functions of tsgather/tsscatter
// (Ops.hs:689-705; grammar: SURVEY.md appendix B = the Int/Bool sub-language
// of Ast.hs:487-528,575-579,654-665) and fused elementwise chains.
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"
#include "ops.cuh"

#define HB_VM_TRANK 8

struct hb_vm_tensor {
  const void *ptr;
  int dtype;
  int rank;
  int64_t shape[HB_VM_TRANK];
  int64_t stride[HB_VM_TRANK];
};

struct hb_vm_prog {
  int n_insns, n_vars, n_out, f32;
  hb_insn code[HB_VM_MAX_INSNS];
  hb_vm_tensor tensors[HB_VM_MAX_TENSORS];
};

// host-side compiled program
struct hb_ixprog {
  std::atomic<int> rc{1};               // handle + recorded lazy nodes (lazy.cuh)
  hb_vm_prog vm;
  hb_array *held[HB_VM_MAX_TENSORS];
  int n_tensors;
  bool affine;                          // every OUT is affine in the variables
  int64_t c0[HB_VM_MAX_OUT];            // out_c = c0[c] + sum_k coef[c][k]*var_k
  int64_t coef[HB_VM_MAX_OUT][HB_R];
  // generated kernels of this program, keyed by the call's constants (jit.cuh)
  mutable std::vector<std::pair<std::string, void *>> generated;
};

#ifdef __CUDACC__

__device__ __forceinline__ double hb_bits_f(uint64_t b) { return __longlong_as_double((long long)b); }
__device__ __forceinline__ uint64_t hb_f_bits(double f) { return (uint64_t)__double_as_longlong(f); }

__device__ __forceinline__ uint64_t hb_vm_load(const void *ptr, int dtype, int64_t off) {
  switch (dtype) {
    case HB_F64: return hb_f_bits(((const double *)ptr)[off]);
    case HB_F32: return hb_f_bits((double)((const float *)ptr)[off]);
    case HB_I64: return (uint64_t)((const int64_t *)ptr)[off];
    case HB_I32: return (uint64_t)(int64_t)((const int32_t *)ptr)[off];
    case HB_I16: return (uint64_t)(int64_t)((const int16_t *)ptr)[off];
    case HB_I8: return (uint64_t)(int64_t)((const int8_t *)ptr)[off];
    default: return (uint64_t)((const uint8_t *)ptr)[off];
  }
}

// Run the program.  `vars`: loop variables; `in(k)`: k-th map input as
// register bits; results of OUT land in outs[].
template <typename In>
__device__ __forceinline__ void hb_vm_run(const hb_vm_prog &P, const int64_t *vars, In in,
                                          uint64_t *outs) {
  uint64_t reg[HB_VM_MAX_REGS];
  const bool f32 = P.f32 != 0;
#define RI(x) ((int64_t)reg[(x)])
#define RF(x) hb_bits_f(reg[(x)])
#define WF(v)                                              \
  do {                                                     \
    double v_ = (v);                                       \
    if (f32) v_ = (double)(float)v_;                       \
    reg[I.dst] = hb_f_bits(v_);                            \
  } while (0)
#pragma unroll 1
  for (int pc = 0; pc < P.n_insns; pc++) {
    const hb_insn I = P.code[pc];
    switch (I.op) {
      case HB_VM_IVAR: reg[I.dst] = (uint64_t)vars[I.a]; break;
      case HB_VM_ICONST: reg[I.dst] = (uint64_t)I.imm.i; break;
      case HB_VM_IADD: reg[I.dst] = (uint64_t)(RI(I.a) + RI(I.b)); break;
      case HB_VM_ISUB: reg[I.dst] = (uint64_t)(RI(I.a) - RI(I.b)); break;
      case HB_VM_IMUL: reg[I.dst] = (uint64_t)(RI(I.a) * RI(I.b)); break;
      case HB_VM_IQUOT: { int64_t a = RI(I.a), b = RI(I.b); reg[I.dst] = (uint64_t)(b == 0 ? a : a / b); } break;
      case HB_VM_IREM: { int64_t a = RI(I.a), b = RI(I.b); reg[I.dst] = (uint64_t)(b == 0 ? a : a % b); } break;
      case HB_VM_INEG: reg[I.dst] = (uint64_t)(-RI(I.a)); break;
      case HB_VM_IABS: { int64_t a = RI(I.a); reg[I.dst] = (uint64_t)(a < 0 ? -a : a); } break;
      case HB_VM_ISIGNUM: { int64_t a = RI(I.a); reg[I.dst] = (uint64_t)(int64_t)((a > 0) - (a < 0)); } break;
      case HB_VM_IMIN: { int64_t a = RI(I.a), b = RI(I.b); reg[I.dst] = (uint64_t)(b < a ? b : a); } break;
      case HB_VM_IMAX: { int64_t a = RI(I.a), b = RI(I.b); reg[I.dst] = (uint64_t)(a < b ? b : a); } break;
      case HB_VM_ILEQ: reg[I.dst] = RI(I.a) <= RI(I.b); break;
      case HB_VM_ILT: reg[I.dst] = RI(I.a) < RI(I.b); break;
      case HB_VM_IEQ: reg[I.dst] = RI(I.a) == RI(I.b); break;
      case HB_VM_BAND: reg[I.dst] = (reg[I.a] != 0) && (reg[I.b] != 0); break;
      case HB_VM_BOR: reg[I.dst] = (reg[I.a] != 0) || (reg[I.b] != 0); break;
      case HB_VM_BNOT: reg[I.dst] = reg[I.a] == 0; break;
      case HB_VM_SEL: reg[I.dst] = reg[I.a] != 0 ? reg[I.b] : reg[I.c]; break;
      case HB_VM_FCONST: WF(I.imm.f); break;
      case HB_VM_FADD: WF(RF(I.a) + RF(I.b)); break;
      case HB_VM_FSUB: WF(RF(I.a) - RF(I.b)); break;
      case HB_VM_FMUL: WF(RF(I.a) * RF(I.b)); break;
      case HB_VM_FDIV: WF(RF(I.a) / RF(I.b)); break;
      case HB_VM_FPOW: WF(pow(RF(I.a), RF(I.b))); break;
      case HB_VM_FLOGBASE: WF(log(RF(I.b)) / log(RF(I.a))); break;
      case HB_VM_FATAN2: WF(atan2(RF(I.a), RF(I.b))); break;
      case HB_VM_FMAX: { double a = RF(I.a), b = RF(I.b); WF(a < b ? b : a); } break;
      case HB_VM_FMIN: { double a = RF(I.a), b = RF(I.b); WF(b < a ? b : a); } break;
      case HB_VM_FUN1:
        if (f32) WF((double)hb_un_rt<float>(I.b, (float)RF(I.a)));
        else WF(hb_un_rt<double>(I.b, RF(I.a)));
        break;
      case HB_VM_FLEQ: reg[I.dst] = RF(I.a) <= RF(I.b); break;
      case HB_VM_FLT: reg[I.dst] = RF(I.a) < RF(I.b); break;
      case HB_VM_FEQ: reg[I.dst] = RF(I.a) == RF(I.b); break;
      case HB_VM_I2F: WF((double)RI(I.a)); break;
      case HB_VM_FLOOR: reg[I.dst] = (uint64_t)(int64_t)floor(RF(I.a)); break;
      case HB_VM_TBL: {
        const hb_vm_tensor &t = P.tensors[I.a];
        int64_t off = 0;
        bool inb = true;
        for (int k = 0; k < I.n; k++) {
          int64_t ix = RI(I.imm.regs[k]);
          inb = inb && ix >= 0 && ix < t.shape[k];
          off += ix * t.stride[k];
        }
        reg[I.dst] = inb ? hb_vm_load(t.ptr, t.dtype, off) : 0ull;  // OOB -> def
      } break;
      case HB_VM_ARGMAX:
      case HB_VM_ARGMIN: {
        const hb_vm_tensor &t = P.tensors[I.a];
        int64_t off = 0;
        bool inb = true;
        for (int k = 0; k < I.n; k++) {
          int64_t ix = RI(I.imm.regs[k]);
          inb = inb && ix >= 0 && ix < t.shape[k];
          off += ix * t.stride[k];
        }
        int64_t best = 0;
        if (inb) {
          const int64_t n = t.shape[I.n], s = t.stride[I.n];
          const bool fl = t.dtype == HB_F64 || t.dtype == HB_F32;
          const bool mx = I.op == HB_VM_ARGMAX;
          if (fl) {
            double bv = hb_bits_f(hb_vm_load(t.ptr, t.dtype, off));
            for (int64_t i = 1; i < n; i++) {
              double v = hb_bits_f(hb_vm_load(t.ptr, t.dtype, off + i * s));
              if (mx ? (v > bv) : (v < bv)) { bv = v; best = i; }
            }
          } else {
            int64_t bv = (int64_t)hb_vm_load(t.ptr, t.dtype, off);
            for (int64_t i = 1; i < n; i++) {
              int64_t v = (int64_t)hb_vm_load(t.ptr, t.dtype, off + i * s);
              if (mx ? (v > bv) : (v < bv)) { bv = v; best = i; }
            }
          }
        }
        reg[I.dst] = (uint64_t)best;
      } break;
      case HB_VM_IN: reg[I.dst] = in(I.a); break;
      case HB_VM_OUT: outs[I.b] = reg[I.a]; break;
      default: break;
    }
  }
#undef RI
#undef RF
#undef WF
}

struct hb_vm_noin {
  __device__ __forceinline__ uint64_t operator()(int) const { return 0; }
};

#endif

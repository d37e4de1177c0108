"""ctypes binding of libhordeb200.so -- the same C ABI (include/hordeb200.h) a
Haskell `foreign import ccall` would bind.  No compute lives
here; a missing library or a missing GPU is a hard error (there is no CPU
fallback anywhere in the product path).
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhordeb200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "hordeb200.h")

HB_MAX_RANK = 12
HB_UNIQUE_ID_BYTES = 128

# dtypes -------------------------------------------------------------------
F64, F32, I64, I32, I16, I8, BOOL = range(7)
DTYPE_NAMES = ["f64", "f32", "i64", "i32", "i16", "i8", "bool"]
NP_NAMES = ["float64", "float32", "int64", "int32", "int16", "int8", "bool"]

# op codes (must match the enums of hordeb200.h) ------------------------------
UNOPS = ["negate", "abs", "signum", "recip", "exp", "log", "sqrt", "sin", "cos", "tan", "asin",
         "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh"]
BINOPS = ["add", "sub", "mul", "divide", "power", "logbase", "atan2", "quot", "rem", "max", "min"]
VMOPS = ["IVAR", "ICONST", "IADD", "ISUB", "IMUL", "IQUOT", "IREM", "INEG", "IABS", "ISIGNUM",
         "IMIN", "IMAX", "ILEQ", "ILT", "IEQ", "BAND", "BOR", "BNOT", "SEL", "FCONST", "FADD",
         "FSUB", "FMUL", "FDIV", "FPOW", "FLOGBASE", "FATAN2", "FMAX", "FMIN", "FUN1", "FLEQ",
         "FLT", "FEQ", "I2F", "FLOOR", "TBL", "ARGMAX", "ARGMIN", "IN", "OUT"]
UNOP = {n: i for i, n in enumerate(UNOPS)}
BINOP = {n: i for i, n in enumerate(BINOPS)}
VMOP = {n: i for i, n in enumerate(VMOPS)}


class _Imm(C.Union):
    _fields_ = [("i", C.c_int64), ("f", C.c_double), ("regs", C.c_uint8 * 8)]


class Insn(C.Structure):
    _fields_ = [("op", C.c_uint8), ("dst", C.c_uint8), ("a", C.c_uint8), ("b", C.c_uint8),
                ("c", C.c_uint8), ("n", C.c_uint8), ("pad", C.c_uint8 * 2), ("imm", _Imm)]


assert C.sizeof(Insn) == 16

P = C.c_void_p
PP = C.POINTER(C.c_void_p)
I64P = C.POINTER(C.c_int64)
IP = C.POINTER(C.c_int)

# name -> argtypes (restype is always int).  Checked against the header by
# tests/test_abi.py (every declared symbol must be exported and listed here).
PROTOTYPES = {
    "hb_init": [C.c_int], "hb_shutdown": [], "hb_sync": [],
    "hb_last_error": [C.c_char_p, C.c_size_t], "hb_device_count": [IP],
    "hb_mem_stats": [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)],
    "hb_launch_count": [C.POINTER(C.c_uint64)], "hb_reset_launch_count": [],
    "hb_stream": [PP], "hb_event_record": [C.c_int],
    "hb_event_elapsed_ms": [C.c_int, C.c_int, C.POINTER(C.c_float)],
    "hb_prof_begin": [], "hb_prof_end": [], "hb_prof_report": [C.c_char_p, C.c_size_t],
    "hb_retain": [P], "hb_release": [P],
    "hb_from_host": [C.c_int, C.c_int, I64P, P, PP], "hb_upload_into": [P, P],
    "hb_prefetch_into": [P, P], "hb_prefetch_ready": [], "hb_prefetch_consumed": [],
    "hb_to_host": [P, P], "hb_scalar_to_host": [P, P],
    "hb_full": [C.c_int, C.c_int, I64P, P, PP],
    "hb_host_alloc": [C.c_size_t, PP], "hb_host_free": [P],
    "hb_rank": [P, IP], "hb_shape": [P, I64P], "hb_strides": [P, I64P],
    "hb_dtype_of": [P, IP], "hb_size": [P, I64P], "hb_is_contiguous": [P, IP],
    "hb_device_ptr": [P, PP],
    "hb_transpose": [P, C.c_int, IP, PP], "hb_replicate": [P, C.c_int, I64P, PP],
    "hb_slice": [P, C.c_int64, C.c_int64, PP], "hb_reverse": [P, PP],
    "hb_reshape": [P, C.c_int, I64P, PP], "hb_index_prefix": [P, C.c_int, I64P, PP],
    "hb_contiguous": [P, PP], "hb_append": [P, P, PP], "hb_stack": [C.c_int, PP, PP],
    "hb_iota": [C.c_int, C.c_int64, PP], "hb_flatten_concat": [C.c_int, PP, PP],
    "hb_unary": [C.c_int, P, PP], "hb_binary": [C.c_int, P, P, PP],
    "hb_cast": [P, C.c_int, PP], "hb_floor": [P, C.c_int, PP],
    "hb_compare_all": [C.c_int, P, P, IP], "hb_add_inplace_or_new": [P, P, PP],
    "hb_ixprog_build": [C.POINTER(Insn), C.c_int, C.c_int, C.c_int, PP, C.c_int, C.c_int, PP],
    "hb_ixprog_free": [P], "hb_ixprog_is_affine": [P, IP],
    "hb_gather": [P, C.c_int, I64P, C.c_int, P, PP],
    "hb_scatter_add": [P, C.c_int, C.c_int, I64P, P, PP],
    "hb_onehot": [P, C.c_int, I64P, I64P, PP],
    "hb_fused_map": [P, C.c_int, PP, C.c_int, PP],
    "hb_jit_stats": [IP, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)],
    "hb_jit_index_source_text": [P, C.c_int, C.c_int, I64P, C.c_int, I64P, I64P, C.c_int, C.c_char_p, C.c_size_t],
    "hb_jit_map_source_text": [P, C.c_int, I64P, C.c_int, I64P, IP, C.c_int, C.c_int, C.c_char_p, C.c_size_t],
    "hb_jit_compile_only": [C.c_char_p, C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t],
    "hb_sum_leading": [P, C.c_int, PP], "hb_sum_all": [P, PP], "hb_dot_all": [P, P, PP],
    "hb_dot_inner": [P, P, PP], "hb_dot_inner_sum_leading": [P, P, C.c_int, PP],
    "hb_matmul2": [P, P, PP], "hb_matvecmul": [P, P, PP], "hb_matmul2_acc": [P, P, P, PP],
    "hb_matmul2_sum": [P, P, P, P, PP], "hb_matmul2_pair": [P, P, P, P, PP, PP],
    "hb_matmul2_acc_pair": [P, P, P, P, P, P, PP, PP], "hb_matmul2_list": [C.c_int, C.POINTER(P), C.POINTER(P), P, PP],
    "hb_matmul2_group": [C.c_int, C.POINTER(P), C.POINTER(P), C.c_int, C.POINTER(P), C.POINTER(P), C.POINTER(C.c_int), C.POINTER(P)],
    "hb_matmul2_sum_tanh": [P, P, P, P, P, PP],
    "hb_argmax_last": [P, PP], "hb_argmin_last": [P, PP],
    "hb_conv2d_preserving": [P, P, PP],
    "hb_conv2d_preserving_dkrn": [P, P, C.c_int, C.c_int, PP],
    "hb_conv2d_preserving_dinp": [P, P, PP],
    "hb_relu": [P, PP], "hb_maxpool2d": [P, C.c_int, C.c_int, PP, PP],
    "hb_maxpool2d_bwd": [P, P, C.c_int, C.c_int, C.c_int64, C.c_int64, PP],
    "hb_relu_pool_fwd": [P, P, C.c_int, PP, PP],
    "hb_mnist_batch_into": [P, P, P, C.c_int64, P, P],
    "hb_tanh_affine_fwd": [P, P, P, PP],
    "hb_tanh_affine_bwd": [P, P, PP, PP],
    "hb_tanh_affine_bwd_acc": [P, P, P, PP, PP],
    "hb_relu_pool_bwd": [P, P, C.c_int, C.c_int64, C.c_int64, PP, PP],
    "hb_conv_relu_pool_fwd": [P, P, P, C.c_int, PP, PP],
    "hb_conv_relu_pool_bwd": [P, P, P, P, C.c_int, PP, PP, PP],
    "hb_softmax_xent": [P, P, C.c_double, PP, PP],
    "hb_sgd_step": [P, P, C.c_double],
    "hb_adam_step": [P, P, P, P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64],
    "hb_prod_fold_grad": [P, C.c_double, PP, PP],
    "hb_graph_begin": [], "hb_graph_end": [PP], "hb_graph_launch": [P], "hb_graph_free": [P],
    "hb_comm_unique_id": [P], "hb_comm_init": [C.c_int, C.c_int, P],
    "hb_allreduce_sum": [P, C.c_double], "hb_comm_destroy": [], "hb_comm_set_global_objective": [C.c_int],
    "hb_allreduce_adam_step": [P, C.c_int64, P, P, P, P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double],
    "hb_microbench": [C.c_int, C.POINTER(C.c_double)],
    "hb_lazy_begin": [], "hb_lazy_end": [C.c_int, PP, C.c_uint, PP], "hb_lazy_abort": [],
    "hb_lazy_run": [P], "hb_lazy_free": [P], "hb_lazy_stats": [P, IP, IP],
    "hb_lazy_dump": [P, C.c_char_p, C.c_size_t], "hb_init_trace_only": [],
}

LAZY_REWRITE, LAZY_NO_FUSE_LAYER = 1, 2


def header_symbols(path: str = HEADER_PATH) -> list[str]:
    """Names of every function declared in include/hordeb200.h."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(hb_[a-z0-9_]+)\s*\(", text)))


class HordeError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load the shared library (once).  Fails loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HordeError(
                f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C horde-ad_b200/csrc). There is no CPU fallback.")
        l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, argtypes in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        _lib = l
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib().hb_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(status: int) -> None:
    if status != 0:
        raise HordeError(last_error() or f"libhordeb200 status {status}")


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args))


_inited_device = None


def init(device: int | None = None) -> int:
    """hb_init on LOCAL_RANK (or `device`).  Raises if no GPU is present."""
    global _inited_device
    if _inited_device is not None:
        return _inited_device
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    call("hb_init", device)
    _inited_device = device
    return device


def init_trace_only() -> int:
    """hb_init_trace_only: no device; compute calls can only be recorded (hb_lazy_begin), never run.  For
    exercising the contraction pass on a box without a GPU."""
    global _inited_device
    call("hb_init_trace_only")
    _inited_device = -1
    return -1


def device_count() -> int:
    n = C.c_int(0)
    lib().hb_device_count(C.byref(n))
    return n.value


def shape_arg(shape) -> "C.Array":
    shape = list(shape)
    return (C.c_int64 * max(1, len(shape)))(*shape)

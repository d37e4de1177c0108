"""The device carrier: a host-side mirror of what `instance BaseTensor Device`
would be in Haskell (cf. `instance BaseTensor Concrete`, OpsConcrete.hs:97-875).

Method names follow horde-ad's user-facing shaped API (`rsum = trsum` style
renames of OpsTensor.hs:16-83): ssum, ssumN, sdot1In, sgather, sscatter,
stranspose ...  Every method is ONE call through the C ABI
(include/hordeb200.h), exactly like a `foreign import ccall` wrapper; nothing
is computed on the host.  Index functions are Python closures over the
symbolic index scalar `ix.Ix` and are reified into VM programs
(option A).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Sequence

import numpy as np

from . import ffi, ix as ixm
from .ffi import call

_NP2DT = {np.dtype(n): i for i, n in enumerate(ffi.NP_NAMES)}


def _dt(dtype) -> int:
    if isinstance(dtype, int):
        return dtype
    return _NP2DT[np.dtype(dtype)]


class DeviceArray:
    """A `ForeignPtr hb_array` (finaliser = hb_release)."""

    __slots__ = ("h", "_shape", "_dt", "__weakref__")

    def __init__(self, handle):
        self.h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle
        self._shape = None
        self._dt = None

    def __del__(self):
        h = getattr(self, "h", None)
        if h is not None and h.value and ffi._lib is not None:
            ffi._lib.hb_release(h)
            self.h = None

    @property
    def shape(self) -> tuple:
        if self._shape is None:
            r = C.c_int()
            call("hb_rank", self.h, C.byref(r))
            sh = (C.c_int64 * ffi.HB_MAX_RANK)()
            call("hb_shape", self.h, sh)
            self._shape = tuple(sh[i] for i in range(r.value))
        return self._shape

    @property
    def strides(self) -> tuple:
        st = (C.c_int64 * ffi.HB_MAX_RANK)()
        call("hb_strides", self.h, st)
        return tuple(st[i] for i in range(len(self.shape)))

    @property
    def dt(self) -> int:
        if self._dt is None:
            d = C.c_int()
            call("hb_dtype_of", self.h, C.byref(d))
            self._dt = d.value
        return self._dt

    @property
    def dtype(self) -> np.dtype:
        return np.dtype(ffi.NP_NAMES[self.dt])

    @property
    def size(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n

    def __repr__(self):
        return f"DeviceArray({ffi.DTYPE_NAMES[self.dt]}{list(self.shape)})"


def _out() -> C.c_void_p:
    return C.c_void_p()


def _wrap(p: C.c_void_p) -> DeviceArray:
    return DeviceArray(p.value)


class Program:
    """A built hb_ixprog (retains its tensor operands)."""

    def __init__(self, outs: Sequence[ixm.Ix], n_vars: int, f32: bool = False):
        prog = ixm.compile_program(outs, n_vars)
        self.keep = list(prog.tensors)
        n = len(prog.insns)
        arr = (ffi.Insn * max(1, n))(*prog.insns)
        tens = (C.c_void_p * max(1, len(prog.tensors)))(*[t.h for t in prog.tensors])
        p = _out()
        call("hb_ixprog_build", arr, n, n_vars, prog.n_out, tens, len(prog.tensors), int(f32), C.byref(p))
        self.h = p
        self.n_insns = n

    @property
    def is_affine(self) -> bool:
        f = C.c_int()
        call("hb_ixprog_is_affine", self.h, C.byref(f))
        return bool(f.value)

    def __del__(self):
        h = getattr(self, "h", None)
        if h is not None and h.value and ffi._lib is not None:
            ffi._lib.hb_ixprog_free(h)
            self.h = None


class LazyProgram:
    """A recorded + rewritten op program (hb_lazy_*, include/hordeb200.h): `run()` replays it; the arrays
    given as `outs` hold the results afterwards."""

    def __init__(self, handle, outs):
        self.h = handle
        self.outs = list(outs)

    def run(self):
        call("hb_lazy_run", self.h)

    @property
    def stats(self) -> tuple:
        rec, live = C.c_int(), C.c_int()
        call("hb_lazy_stats", self.h, C.byref(rec), C.byref(live))
        return rec.value, live.value

    def dump(self) -> str:
        buf = C.create_string_buffer(1 << 20)
        call("hb_lazy_dump", self.h, buf, len(buf))
        return buf.value.decode(errors="replace")

    def close(self):
        if self.h is not None and ffi._lib is not None:
            ffi._lib.hb_lazy_free(self.h)
        self.h = None

    def __del__(self):
        self.close()


class DeviceBackend:
    """`target = Device`."""

    name = "device"

    def __init__(self, device: int | None = None, trace_only: bool = False):
        self.trace_only = trace_only
        self.device = ffi.init_trace_only() if trace_only else ffi.init(device)

    # ------------------------------------------------------- lazy programs
    def record(self, build: Callable, rewrite: bool = True, fuse_layers: bool = True) -> LazyProgram:
        """Record the C-ABI calls `build()` makes into a lazy program (nothing is launched), run the device
        contraction pass over it (`rewrite`), and return the program; `build` returns the output array(s)."""
        call("hb_lazy_begin")
        try:
            outs = build()
        except BaseException:
            call("hb_lazy_abort")
            raise
        outs = list(outs) if isinstance(outs, (list, tuple)) else [outs]
        arr = (C.c_void_p * max(1, len(outs)))(*[o.h for o in outs])
        flags = (ffi.LAZY_REWRITE if rewrite else 0) | (0 if fuse_layers else ffi.LAZY_NO_FUSE_LAYER)
        p = _out()
        try:
            call("hb_lazy_end", len(outs), arr, flags, C.byref(p))
        except BaseException:
            call("hb_lazy_abort")
            raise
        return LazyProgram(p, outs)

    # ------------------------------------------------------------ transfer
    def sconcrete(self, a) -> DeviceArray:
        shape = np.shape(a)                 # np.ascontiguousarray turns a 0-d array into shape (1,)
        a = np.ascontiguousarray(a).reshape(shape)
        p = _out()
        call("hb_from_host", _dt(a.dtype), a.ndim, ffi.shape_arg(a.shape),
             a.ctypes.data_as(C.c_void_p), C.byref(p))
        # pageable source: the runtime has staged it before returning
        return _wrap(p)

    def to_numpy(self, t: DeviceArray) -> np.ndarray:
        out = np.empty(t.shape, dtype=t.dtype)
        call("hb_to_host", t.h, out.ctypes.data_as(C.c_void_p))
        return out

    def shape(self, t): return t.shape
    def dtype(self, t): return t.dtype

    def srepl(self, shape, value, dtype=np.float64) -> DeviceArray:
        v = np.array(value, dtype=dtype)
        p = _out()
        call("hb_full", _dt(dtype), len(shape), ffi.shape_arg(shape), v.ctypes.data_as(C.c_void_p), C.byref(p))
        return _wrap(p)

    def sscalar(self, value, dtype=np.float64): return self.srepl((), value, dtype)

    def kfromS(self, t: DeviceArray):
        v = np.empty((), dtype=t.dtype)
        call("hb_scalar_to_host", t.h, v.ctypes.data_as(C.c_void_p))
        return v[()]

    def sync(self): call("hb_sync")

    # ------------------------------------------------------ stack / unstack
    def sfromVector(self, ts: Sequence[DeviceArray]) -> DeviceArray:
        arr = (C.c_void_p * max(1, len(ts)))(*[t.h for t in ts])
        p = _out()
        call("hb_stack", len(ts), arr, C.byref(p))
        return _wrap(p)

    def sunravelToList(self, t): return [self.sindex(t, [i]) for i in range(t.shape[0])]

    def flatten_concat(self, ts: Sequence[DeviceArray]) -> DeviceArray:
        arr = (C.c_void_p * len(ts))(*[t.h for t in ts])
        p = _out()
        call("hb_flatten_concat", len(ts), arr, C.byref(p))
        return _wrap(p)

    # --------------------------------------------------------------- views
    def stranspose(self, perm, t):
        p = _out()
        call("hb_transpose", t.h, len(perm), (C.c_int * max(1, len(perm)))(*perm), C.byref(p))
        return _wrap(p)

    def sreplicateN(self, shm, t):
        p = _out()
        call("hb_replicate", t.h, len(shm), ffi.shape_arg(shm), C.byref(p))
        return _wrap(p)

    def sreplicate(self, k, t): return self.sreplicateN([k], t)
    def sreplicate0N(self, shm, t0): return self.sreplicateN(shm, t0)

    def sslice(self, i, n, t):
        p = _out()
        call("hb_slice", t.h, i, n, C.byref(p))
        return _wrap(p)

    def sreverse(self, t):
        p = _out()
        call("hb_reverse", t.h, C.byref(p))
        return _wrap(p)

    def sreshape(self, sh, t):
        p = _out()
        call("hb_reshape", t.h, len(sh), ffi.shape_arg(sh), C.byref(p))
        return _wrap(p)

    def sindex(self, t, ixs: Sequence[int]):
        p = _out()
        call("hb_index_prefix", t.h, len(ixs), ffi.shape_arg(ixs), C.byref(p))
        return _wrap(p)

    def sappend(self, a, b):
        p = _out()
        call("hb_append", a.h, b.h, C.byref(p))
        return _wrap(p)

    def scontiguous(self, t):
        p = _out()
        call("hb_contiguous", t.h, C.byref(p))
        return _wrap(p)

    def siota(self, n, dtype=np.int64):
        p = _out()
        call("hb_iota", _dt(dtype), n, C.byref(p))
        return _wrap(p)

    # --------------------------------------------------------- elementwise
    def unary(self, op: str, t):
        p = _out()
        call("hb_unary", ffi.UNOP[op], t.h, C.byref(p))
        return _wrap(p)

    def binary(self, op: str, a, b):
        p = _out()
        call("hb_binary", ffi.BINOP[op], a.h, b.h, C.byref(p))
        return _wrap(p)

    def scast(self, t, dtype):
        p = _out()
        call("hb_cast", t.h, _dt(dtype), C.byref(p))
        return _wrap(p)

    sfromIntegral = scast

    def sfloor(self, t, dtype=np.int64):
        p = _out()
        call("hb_floor", t.h, _dt(dtype), C.byref(p))
        return _wrap(p)

    def _cmp(self, op, a, b) -> bool:
        r = C.c_int()
        call("hb_compare_all", op, a.h, b.h, C.byref(r))
        return bool(r.value)

    def all_eq(self, a, b): return self._cmp(0, a, b)
    def all_leq(self, a, b): return self._cmp(1, a, b)
    def all_lt(self, a, b): return self._cmp(2, a, b)

    def add_accumulate(self, acc, c):
        """taddTarget in cotangent accumulation (DeltaEval.hs:139-148): in place
        when `acc` is uniquely referenced, else a fresh array."""
        p = _out()
        call("hb_add_inplace_or_new", acc.h, c.h, C.byref(p))
        if p.value == acc.h.value:  # same handle, retained once more
            ffi._lib.hb_release(p)
            return acc
        return _wrap(p)

    def fused_map(self, outs_expr: ixm.Ix, ins: Sequence[DeviceArray], dtype) -> DeviceArray:
        prog = Program([outs_expr], 0, f32=np.dtype(dtype) == np.float32)
        arr = (C.c_void_p * len(ins))(*[t.h for t in ins])
        p = _out()
        call("hb_fused_map", prog.h, len(ins), arr, _dt(dtype), C.byref(p))
        return _wrap(p)

    # ---------------------------------------------------- reduce / contract
    def ssumN(self, t, n_axes: int):
        p = _out()
        call("hb_sum_leading", t.h, n_axes, C.byref(p))
        return _wrap(p)

    def ssum(self, t): return self.ssumN(t, 1)
    def ssum0(self, t): return self.ssumN(t, len(t.shape))

    def sdot0(self, a, b):
        p = _out()
        call("hb_dot_all", a.h, b.h, C.byref(p))
        return _wrap(p)

    def sdot1In(self, a, b):
        p = _out()
        call("hb_dot_inner", a.h, b.h, C.byref(p))
        return _wrap(p)

    def sdot1InSumN(self, a, b, n_axes: int):
        p = _out()
        call("hb_dot_inner_sum_leading", a.h, b.h, n_axes, C.byref(p))
        return _wrap(p)

    def smatmul2(self, a, b):
        p = _out()
        call("hb_matmul2", a.h, b.h, C.byref(p))
        return _wrap(p)

    def smatmul2_acc(self, acc, a, b):
        """taddTarget acc (smatmul2 a b) in one call: accumulated by the GEMM epilogue straight into `acc`
        when it is uniquely referenced (hb_matmul2_acc)."""
        p = _out()
        call("hb_matmul2_acc", acc.h, a.h, b.h, C.byref(p))
        if p.value == acc.h.value:  # same handle, retained once more
            ffi._lib.hb_release(p)
            return acc
        return _wrap(p)

    def smatmul2_sum(self, a1, b1, a2, b2):
        """smatmul2 a1 b1 + smatmul2 a2 b2 in one launch (hb_matmul2_sum)."""
        p = _out()
        call("hb_matmul2_sum", a1.h, b1.h, a2.h, b2.h, C.byref(p))
        return _wrap(p)

    def smatmul2_pair(self, a1, b1, a2, b2):
        """(smatmul2 a1 b1, smatmul2 a2 b2) in one launch (hb_matmul2_pair)."""
        p1, p2 = _out(), _out()
        call("hb_matmul2_pair", a1.h, b1.h, a2.h, b2.h, C.byref(p1), C.byref(p2))
        return _wrap(p1), _wrap(p2)

    def smatmul2_acc_pair(self, acc1, a1, b1, acc2, a2, b2):
        """(acc1 + a1 b1, acc2 + a2 b2) in one launch, in place when both accumulators are uniquely referenced."""
        p1, p2 = _out(), _out()
        call("hb_matmul2_acc_pair", acc1.h, a1.h, b1.h, acc2.h, a2.h, b2.h, C.byref(p1), C.byref(p2))
        outs = []
        for p, acc in ((p1, acc1), (p2, acc2)):
            if p.value == acc.h.value:
                ffi._lib.hb_release(p)
                outs.append(acc)
            else:
                outs.append(_wrap(p))
        return outs[0], outs[1]

    def smatmul2_sum_tanh(self, a1, b1, a2, b2, bias):
        """tanh (a1 b1 + a2 b2 + bias stretched over the columns): one launch (hb_matmul2_sum_tanh)."""
        p = _out()
        call("hb_matmul2_sum_tanh", a1.h, b1.h, a2.h, b2.h, bias.h, C.byref(p))
        return _wrap(p)

    def smatmul2_list(self, pairs, acc=None):
        """acc + sum_t a_t b_t as ONE product over the concatenated reduced axes (hb_matmul2_list); in place when
        `acc` is uniquely referenced."""
        n = len(pairs)
        arr_a = (C.c_void_p * n)(*[a.h for a, _ in pairs])
        arr_b = (C.c_void_p * n)(*[b.h for _, b in pairs])
        p = _out()
        call("hb_matmul2_list", n, arr_a, arr_b, acc.h if acc is not None else None, C.byref(p))
        if acc is not None and p.value == acc.h.value:
            ffi._lib.hb_release(p)
            return acc
        return _wrap(p)

    def smatmul2_group(self, results, products):
        """results: [(acc or None, bias or None)]; products: [(a, b, result index)] -> [outs]: every result is its
        accumulator plus its products (in order), then tanh (. + bias) if it has one; one launch where the DMMA
        kernel takes the group (hb_matmul2_group).  Accumulators are updated in place when uniquely referenced."""
        nr, npr = len(results), len(products)
        accs = (C.c_void_p * nr)(*[a.h if a is not None else None for a, _ in results])
        bias = (C.c_void_p * nr)(*[b.h if b is not None else None for _, b in results])
        arr_a = (C.c_void_p * npr)(*[a.h for a, _, _ in products])
        arr_b = (C.c_void_p * npr)(*[b.h for _, b, _ in products])
        idx = (C.c_int * npr)(*[int(r) for _, _, r in products])
        outs = (C.c_void_p * nr)()
        call("hb_matmul2_group", nr, accs, bias, npr, arr_a, arr_b, idx, outs)
        res = []
        for (acc, _), h in zip(results, outs):
            if acc is not None and h == acc.h.value:   # same handle, retained once more
                ffi._lib.hb_release(C.c_void_p(h))
                res.append(acc)
            else:
                res.append(_wrap(C.c_void_p(h)))
        return res

    def smatvecmul(self, m, v):
        p = _out()
        call("hb_matvecmul", m.h, v.h, C.byref(p))
        return _wrap(p)

    def sargMax(self, t):
        p = _out()
        call("hb_argmax_last", t.h, C.byref(p))
        return _wrap(p)

    def sargMin(self, t):
        p = _out()
        call("hb_argmin_last", t.h, C.byref(p))
        return _wrap(p)

    # --------------------------------------------------- index programs
    def ix_tbl(self, t: DeviceArray, ixs) -> ixm.Ix:
        return ixm.tbl(t, ixs, t.dt in (ffi.F64, ffi.F32))

    def ix_argmax(self, t: DeviceArray, ixs) -> ixm.Ix:
        return ixm.argmax(t, ixs)

    def ix_argmin(self, t: DeviceArray, ixs) -> ixm.Ix:
        return ixm.argmin(t, ixs)

    def program(self, f: Callable, n_vars: int) -> Program:
        return Program(ixm.reify(f, n_vars), n_vars)

    def sgather(self, shm, t, f, shp_rank: int, prog: Program | None = None):
        """tsgather @shm @shn @shp: t : shp ++ shn, f : Ix shm -> Ix shp."""
        prog = prog or self.program(f, len(shm))
        p = _out()
        call("hb_gather", t.h, len(shm), ffi.shape_arg(shm), shp_rank, prog.h, C.byref(p))
        return _wrap(p)

    def sscatter(self, shp, t, f, shm_rank: int, prog: Program | None = None):
        """tsscatter @shm @shn @shp: t : shm ++ shn, f : Ix shm -> Ix shp."""
        prog = prog or self.program(f, shm_rank)
        p = _out()
        call("hb_scatter_add", t.h, shm_rank, len(shp), ffi.shape_arg(shp), prog.h, C.byref(p))
        return _wrap(p)

    def soneHot(self, shp, v, ixs: Sequence[int]):
        p = _out()
        call("hb_onehot", v.h, len(shp), ffi.shape_arg(shp), ffi.shape_arg(ixs), C.byref(p))
        return _wrap(p)

    # ----------------------------------------------------- fused fast paths
    def conv2dPreserving(self, K, A):
        p = _out()
        call("hb_conv2d_preserving", K.h, A.h, C.byref(p))
        return _wrap(p)

    def conv2d_dkrn(self, A, dB, KH, KW):
        p = _out()
        call("hb_conv2d_preserving_dkrn", A.h, dB.h, KH, KW, C.byref(p))
        return _wrap(p)

    def conv2d_dinp(self, K, dB):
        p = _out()
        call("hb_conv2d_preserving_dinp", K.h, dB.h, C.byref(p))
        return _wrap(p)

    def relu(self, t):
        p = _out()
        call("hb_relu", t.h, C.byref(p))
        return _wrap(p)

    def maxpool2d(self, x, ksize, stride):
        y, am = _out(), _out()
        call("hb_maxpool2d", x.h, ksize, stride, C.byref(y), C.byref(am))
        return _wrap(y), _wrap(am)

    def maxpool2d_bwd(self, dy, am, ksize, stride, H, W):
        p = _out()
        call("hb_maxpool2d_bwd", dy.h, am.h, ksize, stride, H, W, C.byref(p))
        return _wrap(p)

    def relu_pool_fwd(self, pre, bias, ksize):
        y, code = _out(), _out()
        call("hb_relu_pool_fwd", pre.h, bias.h if bias is not None else None, ksize, C.byref(y), C.byref(code))
        return _wrap(y), _wrap(code)

    def relu_pool_bwd(self, dy, code, ksize, H, W, want_bias=True):
        dp, db = _out(), _out()
        call("hb_relu_pool_bwd", dy.h, code.h, ksize, H, W, C.byref(dp), C.byref(db) if want_bias else None)
        return _wrap(dp), (_wrap(db) if want_bias else None)

    def conv_relu_pool_fwd(self, K, bias, A, ksize):
        y, code = _out(), _out()
        call("hb_conv_relu_pool_fwd", K.h, bias.h if bias is not None else None, A.h, ksize, C.byref(y), C.byref(code))
        return _wrap(y), _wrap(code)

    def conv_relu_pool_bwd(self, K, A, dy, code, ksize, want_bias=True, want_dA=True):
        dk, db, da = _out(), _out(), _out()
        call("hb_conv_relu_pool_bwd", K.h, A.h, dy.h, code.h, ksize, C.byref(dk),
             C.byref(db) if want_bias else None, C.byref(da) if want_dA else None)
        return _wrap(dk), (_wrap(db) if want_bias else None), (_wrap(da) if want_dA else None)

    def softmax_xent(self, logits, expected, scale=1.0, want_grad=True):
        l, d = _out(), _out()
        call("hb_softmax_xent", logits.h, expected.h, float(scale), C.byref(l),
             C.byref(d) if want_grad else None)
        return _wrap(l), (_wrap(d) if want_grad else None)

    def prod_fold(self, x):
        """sfold (*) 1 x (forward value; the AD carrier attaches the scan-based reverse rule)."""
        return self.prod_fold_grad(x, 1.0, want_grad=False)[0]

    def prod_fold_grad(self, x, dt=1.0, want_grad=True):
        pr, g = _out(), _out()
        call("hb_prod_fold_grad", x.h, float(dt), C.byref(pr), C.byref(g) if want_grad else None)
        return _wrap(pr), (_wrap(g) if want_grad else None)

    # ------------------------------------------------------------ optimizer
    def sgd_step(self, p, g, gamma):
        call("hb_sgd_step", p.h, g.h, float(gamma))

    def adam_step(self, p, m, v, g, t, alpha=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-7):
        call("hb_adam_step", p.h, m.h, v.h, g.h, alpha, beta1, beta2, epsilon, int(t))

    def upload_into(self, dst: DeviceArray, host_ptr):
        call("hb_upload_into", dst.h, host_ptr)

    def tanh_affine_fwd(self, u: DeviceArray, v: DeviceArray, bias) -> DeviceArray:
        """tanh ((u + v) + bias stretched over the batch): the body of rnnMnistLayerS in one pass."""
        p = _out()
        call("hb_tanh_affine_fwd", u.h, v.h if v is not None else None, bias.h if bias is not None else None, C.byref(p))
        return _wrap(p)

    def tanh_affine_bwd(self, y: DeviceArray, c: DeviceArray, want_bias=True):
        dz, db = _out(), _out()
        call("hb_tanh_affine_bwd", y.h, c.h, C.byref(dz), C.byref(db) if want_bias else None)
        return _wrap(dz), (_wrap(db) if want_bias else None)

    def tanh_affine_bwd_acc(self, y: DeviceArray, c: DeviceArray, acc: DeviceArray):
        """(dz, acc + row sums of dz): the bias cotangent accumulated by the pass that produces the summand, in
        place when `acc` is uniquely referenced (hb_tanh_affine_bwd_acc)."""
        dz, p = _out(), _out()
        call("hb_tanh_affine_bwd_acc", y.h, c.h, acc.h, C.byref(dz), C.byref(p))
        if p.value == acc.h.value:  # same handle, retained once more
            ffi._lib.hb_release(p)
            return _wrap(dz), acc
        return _wrap(dz), _wrap(p)

    def mnist_batch_into(self, pixels: DeviceArray, labels: DeviceArray, order, start: int,
                         glyphs_dst: DeviceArray, targets_dst: DeviceArray):
        """mkMnistDataBatchS on the device (MnistData.hs:134-171): glyph = pixel / 255, target = one-hot
        label, for the samples order[start : start + batch] (order None = file order)."""
        call("hb_mnist_batch_into", pixels.h, labels.h, order.h if order is not None else None, int(start),
             glyphs_dst.h, targets_dst.h)

"""The kernel generator's OUTPUT executed on the CPU: the CUDA C text that csrc/jit.cu emits for an index program
(`hb_jit_index_source_text`, kind 0 = offset table) is compiled with g++ behind a small shim (blockIdx / threadIdx
as globals, the bit-cast intrinsics) and run over every position; the offsets must equal the oracle's independent
evaluation of the same index function (ixsToLinearMaybe, OpsConcrete.hs:944-950: out of bounds -> no offset).
Covers every tensor-free index function of the reference's gather / scatter goldens (tests/programs.py) plus the
grammar corner cases (quotH / remH by zero, ifH, minH / maxH, negative indices).  The same kernels are compared on
the device in tests/test_gpu_parity.py; this is the no-GPU half."""
import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np
import pytest

from horde_ad_b200 import ad, ffi, ix as ixm
from oracle.concrete import ConcreteBackend
from tests import programs as P

SHIM = r"""
#include <cstring>
#include <cmath>
using std::floor; using std::fabs; using std::sqrt; using std::exp; using std::log; using std::tanh; using std::sin; using std::cos;
#define __device__
#define __forceinline__ inline
#define __global__
#define __restrict__
#define __launch_bounds__(x)
struct hb_dim3 { unsigned x, y, z; };
static hb_dim3 blockIdx, threadIdx, gridDim;
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline int atomicAdd(int *p, int v) { int o = *p; *p += v; return o; }
"""
DRIVER = r"""
extern "C" void hbj_run(long long n, long long *offsets) {
  for (long long b = 0; b < (n + 255) / 256; b++)
    for (unsigned t = 0; t < 256; t++) { blockIdx.x = (unsigned)b; threadIdx.x = t; hbj_k(offsets); }
}
"""
I64MIN = -(1 << 63)
_cache = {}


def _host_offsets(text: bytes, n: int) -> np.ndarray:
    key = hashlib.sha1(text).hexdigest()
    if key not in _cache:
        d = tempfile.mkdtemp(prefix="hbj_")
        src, so = os.path.join(d, "k.cpp"), os.path.join(d, "k.so")
        open(src, "w").write(SHIM + text.decode() + DRIVER)
        subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-w", "-o", so, src], check=True)
        _cache[key] = C.CDLL(so)
    out = np.full(max(n, 1), 7, dtype=np.int64)
    _cache[key].hbj_run(C.c_longlong(n), out.ctypes.data_as(C.c_void_p))
    return out[:n]


def _kernel_text(f, shm, shp):
    """CUDA C of the offset-table kernel for index function f : Ix shm -> Ix shp; None if f reads tensors."""
    prog = ixm.compile_program(ixm.reify(f, len(shm)), len(shm))
    if prog.tensors:
        return None
    arr = (ffi.Insn * max(1, len(prog.insns)))(*prog.insns)
    p = C.c_void_p()
    ffi.call("hb_ixprog_build", arr, len(prog.insns), len(shm), prog.n_out, (C.c_void_p * 1)(), 0, 0, C.byref(p))
    try:
        strides = [int(np.prod(shp[k + 1:])) for k in range(len(shp))]
        i64 = lambda xs: (C.c_int64 * max(1, len(xs)))(*xs)
        buf = C.create_string_buffer(1 << 17)
        ffi.call("hb_jit_index_source_text", p, 0, len(shm), i64(shm), len(shp), i64(shp), i64(strides), 0, buf, len(buf))
        return buf.value
    finally:
        ffi.call("hb_ixprog_free", p)


def _check(f, shm, shp, oracle):
    shm, shp = [int(x) for x in shm], [int(x) for x in shp]
    text = _kernel_text(f, shm, shp)
    if text is None:
        return False
    n = int(np.prod(shm)) if shm else 1
    want = ConcreteBackend._targets(oracle, shm, f, shp)          # row-major linear target, -1 = out of bounds
    got = _host_offsets(text, n)
    np.testing.assert_array_equal(got, np.where(want >= 0, want, I64MIN))
    return True


class Recording(ConcreteBackend):
    """Records every (shm, index function, shp) that a program hands to sgather / sscatter."""

    def __init__(self):
        super().__init__()
        self.seen = []

    def _targets(self, shm, f, shp):
        self.seen.append((list(shm), f, list(shp)))
        return super()._targets(shm, f, shp)


def test_generated_offset_kernels_match_the_oracle_on_the_reference_gather_scatter_programs():
    rec = Recording()
    for _name, prog, inp, _expected in P.GATHER_GOLDEN:
        ad.grad(rec, lambda T, t: T.ssum0(prog(T, t)), [rec.sconcrete(inp)])      # forward gathers + adjoint scatters
    for prog, shape, _e in getattr(P, "GATHER_COND_CASES", []):
        prog(rec, rec.sconcrete(np.arange(int(np.prod(shape)), dtype=np.float64).reshape(shape)))
    assert len(rec.seen) >= 20
    plain = ConcreteBackend()
    checked = sum(_check(f, shm, shp, plain) for shm, f, shp in rec.seen)
    assert checked >= 20, checked


@pytest.mark.parametrize("case", ["quot_rem_by_zero_and_negative", "ifh_min_max", "affine_out_of_bounds", "rank0"])
def test_generated_offset_kernels_grammar_corner_cases(case):
    plain = ConcreteBackend()
    if case == "quot_rem_by_zero_and_negative":      # quotH a 0 = a, remH a 0 = a; truncation toward zero (Types.hs:394-419)
        f = lambda i: [(i[0] - 3).quotH(i[1] - 1), (i[0] - 3).remH(i[1] - 1) + 2]
        assert _check(f, [7, 4], [9, 5], plain)
    elif case == "ifh_min_max":
        f = lambda i: [ixm.ifH((i[0] <= 2) & ~(i[1].eq(1)), i[0], 8 - i[0]), ixm.maxH(ixm.minH(i[0] - i[1], 3), 0),
                       abs(i[1] - 2) * (i[0] - 2).signum() + 1]
        assert _check(f, [6, 5], [9, 4, 4], plain)
    elif case == "affine_out_of_bounds":             # windows that run past both ends: those positions have no offset
        f = lambda i: [i[0] + i[1] - 2, 2 * i[2] - 1]
        assert _check(f, [5, 4, 4], [6, 5], plain)
    else:                                            # no index variables: one position
        assert _check(lambda i: [ixm.Ix.lift(2), ixm.Ix.lift(1)], [], [3, 2], plain)
        assert _check(lambda i: [ixm.Ix.lift(3)], [], [3], plain)          # out of bounds


MAP_DRIVER = r"""
extern "C" void hbj_run_map(unsigned nblocks, double *out, const void *p0, const void *p1) {
  gridDim.x = nblocks;
  for (unsigned b = 0; b < nblocks; b++)
    for (unsigned t = 0; t < 256; t++) { blockIdx.x = b; threadIdx.x = t; hbj_k(out, p0, p1); }
}
"""


def _run_map(text: bytes, out, a, b):
    d = tempfile.mkdtemp(prefix="hbj_")
    src, so = os.path.join(d, "k.cpp"), os.path.join(d, "k.so")
    open(src, "w").write(SHIM + text.decode() + MAP_DRIVER)
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-w", "-o", so, src], check=True)
    C.CDLL(so).hbj_run_map(C.c_uint(3), out.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p),
                           b.ctypes.data_as(C.c_void_p))


def test_generated_map_kernels_match_the_oracle_on_the_host():
    """A fused elementwise chain (ifH, maxH, floor, fromIntegral, tanh / exp / sqrt): the strided scalar-access kernel
    over a transposed Double view and a row-major operand, and the 16-byte kernel over a contiguous operand and a
    broadcast scalar, executed on the CPU and compared with the oracle's evaluation of the same expression."""
    oracle = ConcreteBackend()
    x, y = ixm.map_input(0, True), ixm.map_input(1, True)
    e = ixm.fun1("tanh", x * y + 0.5) / (ixm.fun1("exp", -x) + 1.0) - ixm.fun1("sqrt", abs(y))
    e = ixm.ifH(x <= y, e, ixm.maxH(x, y) * ixm.fromIntegral(ixm.floorH(y)))
    prog = ixm.compile_program([e], 0)
    arr = (ffi.Insn * len(prog.insns))(*prog.insns)
    p = C.c_void_p()
    ffi.call("hb_ixprog_build", arr, len(prog.insns), 0, prog.n_out, (C.c_void_p * 1)(), 0, 0, C.byref(p))
    i64 = lambda xs: (C.c_int64 * len(xs))(*xs)
    rng = np.random.default_rng(5)
    try:
        buf = C.create_string_buffer(1 << 17)
        # (a) logical [30, 50]: operand 0 = transposed view of a [50, 30] array, operand 1 = row-major [30, 50]
        a_store = rng.uniform(-2, 2, (50, 30))
        b_store = rng.uniform(-3, 3, (30, 50))
        ffi.call("hb_jit_map_source_text", p, 2, i64([30, 50]), 2, i64([1, 30, 50, 1]), (C.c_int * 2)(ffi.F64, ffi.F64),
                 ffi.F64, 1, buf, len(buf))
        out = np.full((30, 50), np.nan)
        _run_map(buf.value, out, a_store, b_store)
        want = oracle.fused_map(e, [a_store.T, b_store], np.float64)
        np.testing.assert_allclose(out, want, rtol=1e-14, atol=1e-15)
        # (b) contiguous Double [1024] and a broadcast scalar, two elements per access
        a2, s2 = rng.uniform(-2, 2, 1024), np.array([0.75])
        ffi.call("hb_jit_map_source_text", p, 1, i64([1024]), 2, i64([1, 0]), (C.c_int * 2)(ffi.F64, ffi.F64), ffi.F64, 2,
                 buf, len(buf))
        out2 = np.full(1024, np.nan)
        _run_map(buf.value, out2, a2, s2)
        want2 = oracle.fused_map(e, [a2, np.full(1024, 0.75)], np.float64)
        np.testing.assert_allclose(out2, want2, rtol=1e-14, atol=1e-15)
    finally:
        ffi.call("hb_ixprog_free", p)


GATHER_DRIVER = r"""
extern "C" void hbj_run_gather(long long n, const void *src, void *out) {
  for (long long b = 0; b < (n + 255) / 256; b++)
    for (unsigned t = 0; t < 256; t++) { blockIdx.x = (unsigned)b; threadIdx.x = t; hbj_k((const ELEM *)src, (ELEM *)out); }
}
"""
SCATTER_DRIVER = r"""
extern "C" void hbj_run_scatter(long long n, int *dest, int *counts, int *collision) {
  for (long long b = 0; b < (n + 255) / 256; b++)
    for (unsigned t = 0; t < 256; t++) { blockIdx.x = (unsigned)b; threadIdx.x = t; hbj_k(dest, counts, collision); }
}
"""


def _so(text: str, driver: str):
    d = tempfile.mkdtemp(prefix="hbj_")
    src, so = os.path.join(d, "k.cpp"), os.path.join(d, "k.so")
    open(src, "w").write(SHIM + text + driver)
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-w", "-o", so, src], check=True)
    return C.CDLL(so)


def _index_kernel(f, shm, shp, kind, es):
    prog = ixm.compile_program(ixm.reify(f, len(shm)), len(shm))
    arr = (ffi.Insn * max(1, len(prog.insns)))(*prog.insns)
    p = C.c_void_p()
    ffi.call("hb_ixprog_build", arr, len(prog.insns), len(shm), prog.n_out, (C.c_void_p * 1)(), 0, 0, C.byref(p))
    try:
        strides = [int(np.prod(shp[k + 1:])) for k in range(len(shp))]
        i64 = lambda xs: (C.c_int64 * max(1, len(xs)))(*xs)
        buf = C.create_string_buffer(1 << 17)
        ffi.call("hb_jit_index_source_text", p, kind, len(shm), i64(shm), len(shp), i64(shp), i64(strides), es, buf, len(buf))
        return buf.value.decode()
    finally:
        ffi.call("hb_ixprog_free", p)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int16, np.int8])
def test_generated_element_gather_kernels_on_the_host(dtype):
    """kind 1: the rank-0 gather (one element of 8 / 4 / 2 / 1 bytes per position, zero when out of bounds,
    OpsConcrete.hs:1488-1490) executed on the CPU = the oracle's sgather, bit for bit."""
    import re
    oracle = ConcreteBackend()
    f = lambda i: [i[0] + i[1] - 1, ixm.ifH(i[0].eq(3), -4, (2 * i[1]).remH(5))]      # row 3: out of bounds
    shm, shp = [6, 4], [7, 5]
    es = np.dtype(dtype).itemsize
    text = _index_kernel(f, shm, shp, 1, es)
    elem = re.search(r"hbj_k\(const (.+?) \*__restrict__ src", text).group(1)
    lib = _so(text, GATHER_DRIVER.replace("ELEM", elem))
    rng = np.random.default_rng(6)
    src = (rng.uniform(-100, 100, shp)).astype(dtype)
    out = np.full(shm, 1, dtype=dtype)
    lib.hbj_run_gather(C.c_longlong(24), src.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    want = oracle.sgather(shm, src, f, 2)
    assert out.tobytes() == np.ascontiguousarray(want).tobytes()


def test_generated_scatter_destination_kernels_on_the_host():
    """kind 2: destination of every source position (-1 = dropped, OpsConcrete.hs:1527-1528), per-destination
    counts and the collision flag that selects the deterministic ordered path."""
    oracle = ConcreteBackend()
    for f, shm, shp, collide in ((lambda i: [i[0] + i[1] - 1], [5, 3], [6], True),
                                 (lambda i: [i[0], i[1]], [3, 4], [3, 5], False)):
        text = _index_kernel(f, shm, shp, 2, 0)
        lib = _so(text, SCATTER_DRIVER)
        n, P_ = int(np.prod(shm)), int(np.prod(shp))
        dest, counts, coll = np.full(n, 99, np.int32), np.zeros(P_, np.int32), np.zeros(1, np.int32)
        lib.hbj_run_scatter(C.c_longlong(n), dest.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p),
                            coll.ctypes.data_as(C.c_void_p))
        want = ConcreteBackend._targets(oracle, shm, f, shp)
        np.testing.assert_array_equal(dest, want)
        np.testing.assert_array_equal(counts, np.bincount(want[want >= 0], minlength=P_))
        assert bool(coll[0]) == collide


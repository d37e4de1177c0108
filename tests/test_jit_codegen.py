"""The kernel generator (csrc/jit.cu) without a GPU: programs built through the C ABI are translated to CUDA C and
compiled for sm_100a by NVRTC -- the "does it build" check of generated kernels.  Their results are compared with
the oracle (and with the bytecode interpreter kernels) in tests/test_gpu_parity.py on the device."""
import ctypes as C

import numpy as np
import pytest

from horde_ad_b200 import ffi, ix as ixm


def _build(outs, n_vars, f32=False):
    prog = ixm.compile_program(outs, n_vars)
    assert not prog.tensors
    arr = (ffi.Insn * max(1, len(prog.insns)))(*prog.insns)
    p = C.c_void_p()
    ffi.call("hb_ixprog_build", arr, len(prog.insns), n_vars, prog.n_out, (C.c_void_p * 1)(), 0, int(f32), C.byref(p))
    return p


def _compile(text):
    n = C.c_size_t()
    log = C.create_string_buffer(1 << 14)
    try:
        ffi.call("hb_jit_compile_only", text, C.byref(n), log, len(log))
    except ffi.HordeError as e:  # pragma: no cover
        if "not found" in str(e):
            pytest.skip("libnvrtc.so.12 is not on this machine")
        raise AssertionError(log.value.decode() + "\n" + text.decode())
    assert n.value > 0
    return n.value


def _i64(xs):
    return (C.c_int64 * max(1, len(xs)))(*xs)


def _index_text(p, kind, shm, pshape, pstride, es=8):
    buf = C.create_string_buffer(1 << 16)
    ffi.call("hb_jit_index_source_text", p, kind, len(shm), _i64(shm), len(pshape), _i64(pshape), _i64(pstride), es, buf, len(buf))
    return buf.value


def test_index_program_kernels_compile_for_sm_100a():
    i = [ixm.Ix.var(k) for k in range(3)]
    outs = [ixm.ifH((i[0] <= 2) & ~(i[1].eq(1)), i[0], 8 - i[0]), abs(i[1] - 2).quotH(2) + i[2].remH(3),
            ixm.maxH(ixm.minH(i[0] - i[1], 5), 0) * (-i[2]).signum()]
    p = _build(outs, 3)
    try:
        for kind, es in ((0, 0), (1, 8), (1, 4), (1, 2), (1, 1), (2, 0)):
            text = _index_text(p, kind, [6, 4, 5], [9, 7, 5], [35, 5, 1], es)
            assert b"hbj_k" in text and b"% (unsigned)5" in text      # shapes are constants: no runtime division
            _compile(text)
        # index spaces of 2^31 positions or more decompose in 64 bits
        text = _index_text(p, 0, [1 << 20, 1 << 10, 4], [9, 7, 5], [35, 5, 1])
        assert b"i64 rem_" in text
        _compile(text)
    finally:
        ffi.call("hb_ixprog_free", p)


@pytest.mark.parametrize("f32", [False, True])
def test_map_program_kernels_compile_for_sm_100a(f32):
    x, y = ixm.map_input(0, True), ixm.map_input(1, True)
    e = ixm.fun1("tanh", x * y + 0.5) / (ixm.fun1("exp", -x) + 1.0) - ixm.fun1("sqrt", abs(y))
    e = ixm.ifH(x <= y, e, ixm.maxH(x, y) * ixm.fromIntegral(ixm.floorH(y)))
    p = _build([e], 0, f32)
    dt = ffi.F32 if f32 else ffi.F64
    try:
        buf = C.create_string_buffer(1 << 16)
        # contiguous operands + a broadcast scalar: 16-byte accesses
        vec = 4 if f32 else 2
        ffi.call("hb_jit_map_source_text", p, 1, _i64([1 << 20]), 2, _i64([1, 0]), (C.c_int * 2)(dt, dt), dt, vec, buf, len(buf))
        assert b"alignas(16) VO" in buf.value
        _compile(buf.value)
        # strided operands (a transposed view): scalar accesses, constant strides
        ffi.call("hb_jit_map_source_text", p, 2, _i64([30, 50]), 2, _i64([1, 30, 50, 1]), (C.c_int * 2)(dt, ffi.I32), dt, 1, buf, len(buf))
        assert b"d1 * 30LL" in buf.value
        _compile(buf.value)
    finally:
        ffi.call("hb_ixprog_free", p)


def test_generator_reports_unavailable_without_a_device():
    if ffi.device_count() > 0:
        pytest.skip("a GPU is present")
    avail = C.c_int(-1)
    z = [C.c_uint64() for _ in range(3)]
    ffi.call("hb_jit_stats", C.byref(avail), C.byref(z[0]), C.byref(z[1]), C.byref(z[2]))
    assert avail.value == 0

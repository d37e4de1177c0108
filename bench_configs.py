#!/usr/bin/env python
"""bench_configs.py -- the OTHER BASELINE.json configs on the device backend
(the headline config lives in bench.py): one JSON line per config, with the
achieved fraction of the roofline that bounds its dominant kernel.

  LongProdBench   gradient of the product of an n-vector (n = 1e6 .. 1e8),
                  sfold (*) 1 + its reverse rule = prefix/suffix product scans
                  (bench/LongProdBench.hs): HBM bound, 16 n bytes
  ShortMnistForCI MnistFcnnRanked2 (Double, Float hidden layer), batch 1, SGD
                  gamma = 0.02, 40 samples (bench/ShortMnistForCI.hs:11-17):
                  latency bound; eager C-ABI calls vs one CUDA graph per step
  MnistRnnShaped2 width 512, sequence 28, per-GPU batch 1024 (8192 / 8 GPUs):
                  FP64 DMMA GEMMs (tensor pipe)
  GEMM            smatmul2 [512,512] x [512,8192] alone (the RNN's recurrence)

Usage: python bench_configs.py [--quick]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from horde_ad_b200 import ad, ffi, mnist_fcnn, mnist_rnn  # noqa: E402
from horde_ad_b200.device import DeviceBackend  # noqa: E402


# fusion level of the RNN layer body: "wave" (pair nodes in wavefront order: two layer bodies per launch), "pair" (smatmul2_sum + tanh_affine nodes), "1" (tanh_affine only), "0" (literal)
RNN_FUSION = {"wave": "wave", "pair": "pair", "1": True, "0": False}[os.environ.get("HB_RNN_FUSION", "wave")]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return (float(json.load(open(p))["hbm_gbs"]), "measured") if os.path.exists(p) else (6650.0, "fallback")


def timed(dev, f, reps, warm=3):
    for _ in range(warm):
        f()
    dev.sync()
    ms = C.c_float()
    ffi.call("hb_event_record", 4)
    for _ in range(reps):
        f()
    ffi.call("hb_event_record", 5)
    ffi.call("hb_event_elapsed_ms", 4, 5, C.byref(ms))
    return ms.value / reps


def families(dev, rng, hbm, src, quick, emit=None):
    emit = emit or (lambda d: print(json.dumps(d), flush=True))
    """Achieved HBM GB/s of the data-parallel kernel families against their
    algorithmic bytes (SURVEY.md 8d table)."""
    from horde_ad_b200 import nn
    n = 1 << (24 if quick else 26)            # 512 MiB of doubles per operand
    a = dev.sconcrete(rng.uniform(0.5, 1.5, n))
    b = dev.sconcrete(rng.uniform(0.5, 1.5, n))

    def report(name, ms, by, extra=None):
        gbs = by / (ms * 1e-3) / 1e9
        d = {"config": "kernel family", "kernel": name, "ms": ms, "algorithmic_bytes": by,
             "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "peak_source": src}}
        if extra:
            d.update(extra)
        emit(d)

    report("binary add, contiguous f64 [2^26]", timed(dev, lambda: dev.binary("add", a, b), 10), 3 * 8 * n)
    report("unary exp, contiguous f64", timed(dev, lambda: dev.unary("exp", a), 10), 2 * 8 * n)
    report("taddTarget in place (cotangent accumulation)", timed(dev, lambda: dev.add_accumulate(dev.binary("add", a, b), b), 10) -
           timed(dev, lambda: dev.binary("add", a, b), 10), 3 * 8 * n)
    from horde_ad_b200 import ix as ixm
    from horde_ad_b200.device import Program
    ea, eb = ixm.map_input(0, True), ixm.map_input(1, True)
    chain = (ea * eb + ea) * ixm.maxH(ea, eb)
    cprog = Program([chain], 0)
    arr2 = (C.c_void_p * 2)(a.h, b.h)

    def fused():
        p = C.c_void_p()
        ffi.call("hb_fused_map", cprog.h, 2, arr2, ffi.F64, C.byref(p))
        ffi._lib.hb_release(p)
    report("fused map chain (a*b + a) * max(a, b): one generated kernel, 16-byte accesses", timed(dev, fused, 10), 3 * 8 * n,
           {"unfused_passes": 4})
    report("ssum0 (full reduction)", timed(dev, lambda: dev.ssum0(a), 10), 8 * n)
    report("sdot0", timed(dev, lambda: dev.sdot0(a, b), 10), 2 * 8 * n)
    m = dev.sconcrete(np.zeros(n)), dev.sconcrete(np.zeros(n))
    report("Adam step (p, m, v in place, g read)", timed(dev, lambda: dev.adam_step(a, m[0], m[1], b, 3), 10), 7 * 8 * n)
    del m
    # the CNN's bias-gradient reduction: ssumN @[N,H,W] of [N,H,W,C] (a transposed view of [N,C,H,W])
    N, Cc, H, W = (1024 if quick else 4096), 16, 28, 28
    x = dev.sconcrete(rng.uniform(-.5, .5, (N, Cc, H, W)))
    xt = dev.stranspose([0, 2, 3, 1], x)
    report("ssumN @[N,H,W] over a transposed view [N,H,W,C] (bias gradient)", timed(dev, lambda: dev.ssumN(xt, 3), 10), 8 * N * Cc * H * W)
    bias = dev.sconcrete(rng.uniform(-.5, .5, (Cc,)))
    bst = dev.stranspose([0, 3, 1, 2], dev.sreplicateN([N, H, W], bias))
    report("bias add through a stride-0 broadcast view", timed(dev, lambda: dev.binary("add", x, bst), 10), 2 * 8 * N * Cc * H * W)
    report("relu_pool_fwd (bias + relu + 2x2 max-pool)", timed(dev, lambda: dev.relu_pool_fwd(x, bias, 2), 10),
           8 * N * Cc * H * W + 9 * N * Cc * H * W // 4)
    y, code = dev.relu_pool_fwd(x, bias, 2)
    report("relu_pool_bwd (cotangent routing + bias cotangent)", timed(dev, lambda: dev.relu_pool_bwd(y, code, 2, H, W), 10),
           8 * N * Cc * H * W + 9 * N * Cc * H * W // 4)
    # gathers / scatters compiled from index functions
    # (a) the vectorised convolution's first im2col gather: rows [H, KH] of A as [H, N, C, W]
    xh = dev.stranspose([2, 0, 1, 3], x)
    f = lambda ix: [ix[0] + ix[1]]
    prog = dev.program(f, 2)
    KH = 5
    g1 = dev.sgather([H, KH], xh, f, 1, prog)
    report("sgather, affine index i+j (im2col rows, KH = 5)", timed(dev, lambda: dev.sgather([H, KH], xh, f, 1, prog), 10),
           8 * N * Cc * H * W * (KH + 1), {"shape": [H, KH, N, Cc, W]})
    report("sscatter, affine index i+j (its adjoint, deterministic order)",
           timed(dev, lambda: dev.sscatter([H], g1, f, 2, prog), 5), 8 * N * Cc * H * W * (KH + 1))
    del g1
    # (b) data-dependent gather: max-pool as a gather at kargMax of the window
    Ho, Wo = H // 2, W // 2
    win = dev.sconcrete(rng.uniform(-.5, .5, (N, Cc, Ho, Wo, 4)))
    fa = lambda ix: [ix[0], ix[1], ix[2], ix[3], dev.ix_argmax(win, ix)]
    proga = dev.program(fa, 4)
    report("sgather with kargMax in the index (max-pool of the vectorised program)",
           timed(dev, lambda: dev.sgather([N, Cc, Ho, Wo], win, fa, 5, proga), 10), 8 * N * Cc * Ho * Wo * 5)
    # (c) the reference's gather48 chain (bench/ConvVjpBench.hs:428-439): [50,3,3,50] -> [48,3,3,48,3,3]
    t = dev.sconcrete(rng.uniform(-.5, .5, (50, 3, 3, 50)))

    def gather48():
        q = dev.sgather([48, 3], t, f, 1)                      # [48,3,3,3,50]
        q = dev.stranspose([4, 0, 1, 2, 3], q)                 # [50,48,3,3,3]
        return dev.sgather([48, 3], q, f, 1)                   # [48,3,48,3,3,3]
    eager_ms = timed(dev, gather48, 20)
    # the same chain recorded once through the C ABI (hb_lazy_*) and replayed inside a CUDA graph, as the trainer
    # replays its step: what is left is the launch floor of the chain's kernels (1.5 MB, L2-resident)
    lp = dev.record(gather48)
    n0, n1 = C.c_uint64(), C.c_uint64()
    ffi.call("hb_launch_count", C.byref(n0))
    lp.run()
    ffi.call("hb_launch_count", C.byref(n1))
    dev.sync()
    ffi.call("hb_graph_begin")
    try:
        lp.run()
    finally:
        gph = C.c_void_p()
        ffi.call("hb_graph_end", C.byref(gph))
    graph_ms = timed(dev, lambda: ffi.call("hb_graph_launch", gph), 50)
    ffi.call("hb_graph_free", gph)
    rec, live = lp.stats
    lp.close()
    report("gather48 chain of the reference (two affine gathers, 1.5 MB)", graph_ms,
           2 * 8 * (48 * 3 * 3 * 3 * 50 + 48 * 3 * 48 * 3 * 3 * 3),
           {"note": "both gathers are provably in-bounds affine programs: called eagerly they ARE strided views of the "
                    "source (ixprog.cu: no launch, ms_eager_calls is host time); the recorded replay materialises every "
                    "node, so the timed CUDA graph holds the two gather kernels (launch floor, L2-resident data)",
            "ms_eager_calls": eager_ms, "nodes_recorded": rec, "nodes_replayed": live,
            "kernel_launches_per_replay": int(n1.value - n0.value)})


import contextlib


@contextlib.contextmanager
def _no_section(_name):
    yield


class _Sections:
    """begin(name) closes the running section and opens the next one (the blocks below stay flat)."""

    def __init__(self, factory):
        self.factory, self.cur = factory, None

    def begin(self, name):
        self.end()
        self.cur = self.factory(name)
        self.cur.__enter__()

    def end(self):
        if self.cur is not None:
            self.cur.__exit__(None, None, None)
            self.cur = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--skip-families", action="store_true")
    ap.add_argument("--only-families", action="store_true")
    args = ap.parse_args()
    dev = DeviceBackend(int(os.environ.get("LOCAL_RANK", "0")))
    run_configs(dev, args.quick, args.skip_families, args.only_families)


def run_configs(dev, quick=False, skip_families=False, only_families=False, emit=None, section=None):
    """Runs every config; `emit(record)` receives one dict per measurement (default: print a JSON line), `section(name)`
    is a context manager entered around each group (bench.py samples the GPU clocks inside it)."""
    import types
    args = types.SimpleNamespace(quick=quick, skip_families=skip_families, only_families=only_families)
    emit = emit or (lambda d: print(json.dumps(d), flush=True))
    section = section or _no_section
    hbm, src = peaks()
    fp64 = C.c_double()
    ffi.call("hb_microbench", 1, C.byref(fp64))
    rng = np.random.Generator(np.random.Philox(42))
    if args.only_families:
        with section("kernel families"):
            families(dev, rng, hbm, src, args.quick, emit)
        return

    # ------------------------------------------ HBM-bound kernel families (8a rows)
    if not args.skip_families:
        with section("kernel families"):
            families(dev, rng, hbm, src, args.quick, emit)

    sec = _Sections(section)
    # ---------------------------------------------------------------- LongProd
    sec.begin("LongProdBench")
    for n in ([10**6, 10**7] if args.quick else [10**6, 10**7, 10**8]):
        x = dev.sconcrete(np.exp(rng.uniform(-1e-3, 1e-3, n)))      # mean-zero log: the product stays finite
        ms = timed(dev, lambda: dev.prod_fold_grad(x, 1.0), 10 if n < 10**8 else 5)
        pr, g = dev.prod_fold_grad(x, 1.0)
        # size-independent property: g_i * x_i = prod for every i
        chk = dev.binary("mul", g, x)
        p = float(dev.kfromS(pr))
        err = float(np.max(np.abs(dev.to_numpy(chk) / p - 1.0)))
        gbs = 16.0 * n / (ms * 1e-3) / 1e9
        # the same through the PROGRAM the reference writes -- cgrad of `sfold (*) 1 x` -- whose step function the
        # carrier recognises (nn.sfold); includes the probe recording and the host read of the 0-d cotangent
        from horde_ad_b200 import nn

        def via_program():
            return ad.grad(dev, lambda T, t: nn.sfold(T, lambda a, e: T.binary("mul", a, e), T.sscalar(1.0), t), [x])
        ms_prog = timed(dev, via_program, 5, warm=1)
        emit(({"config": "LongProdBench", "n": n, "ms": ms, "ms_via_sfold_program": ms_prog, "algorithmic_bytes": 16 * n,
               "hbm_traffic_note": "a prefix AND a suffix product are needed: x is read twice unless it fits in L2 "
                                   "(16 n bytes of HBM traffic for n <= ~1.5e7, 24 n above)",
                          "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                                       "peak_source": src},
                          "max_rel_err_of_g_i_times_x_i_vs_prod": err}))
        del x, pr, g, chk

    # ----------------------------------------------------------- ShortMnistForCI
    sec.begin("ShortMnistForCI")
    for (h1, h2) in ([(300, 100)] if args.quick else [(30, 10), (300, 100), (500, 150), (1500, 500)]):
        params = [dev.sconcrete(p) for p in mnist_fcnn.init_params(h1, h2)]
        datum = dev.sconcrete(rng.uniform(0, 1, 784))
        target = dev.sconcrete(np.eye(10)[3])

        def step():
            _, grads = ad.grad(dev, lambda T, *ps: mnist_fcnn.afcnnMnistLoss2(T, datum, target, list(ps)), params)
            for p, g in zip(params, grads):
                dev.sgd_step(p, g, 0.02)
        step()
        dev.sync()
        t0 = time.perf_counter()
        for _ in range(40):
            step()
        dev.sync()
        eager_ms = (time.perf_counter() - t0) / 40 * 1e3
        # the same step captured once
        ffi.call("hb_graph_begin")
        try:
            step()
        finally:
            gph = C.c_void_p()
            ffi.call("hb_graph_end", C.byref(gph))
        graph_ms = timed(dev, lambda: ffi.call("hb_graph_launch", gph), 40)
        ffi.call("hb_graph_free", gph)
        nparam = sum(int(np.prod(p.shape)) for p in params)
        by = 3 * sum(int(np.prod(p.shape)) * p.dtype.itemsize for p in params)   # W read fwd + bwd, dW write
        emit(({"config": "ShortMnistForCI / MnistFcnnRanked2", "widths": [h1, h2], "batch": 1,
                          "parameters": nparam, "ms_per_sample_eager": eager_ms, "ms_per_sample_graph": graph_ms,
                          "samples_per_s_graph": 1e3 / graph_ms,
                          "roofline": {"bound": "hbm", "achieved": by / (graph_ms * 1e-3) / 1e9, "peak": hbm,
                                       "unit": "GB/s", "frac": by / (graph_ms * 1e-3) / 1e9 / hbm, "peak_source": src,
                                       "note": "batch 1: launch-latency bound by construction"}}))

    # ---------------------------------- secondary CNN config
    sec.begin("MnistCnnShaped2 secondary")
    # (kh,kw,c_out,n_hidden) = (2,2,64,128): 3x3 kernels, 64 channels (aligned with ConvVjpBench)
    from horde_ad_b200.train import CnnTrainer
    B2 = 512 if args.quick else 4096
    tr = CnnTrainer(dev, B2, kh=2, kw=2, c_out=64, n_hidden=128)
    gl = np.ascontiguousarray(rng.uniform(0, 1, (B2, 28, 28)))
    lb = np.ascontiguousarray(np.eye(10)[rng.integers(0, 10, B2)])
    tr.upload(gl.ctypes.data_as(C.c_void_p), lb.ctypes.data_as(C.c_void_p))
    for _ in range(3):
        tr.step_enqueue()
    tr.reset()
    cnt = [0]

    def cnn2_step():
        cnt[0] += 1
        if cnt[0] % 8 == 0:
            tr.reset()               # the reference objective diverges on random labels
        tr.step_enqueue()
    ms = timed(dev, cnn2_step, 10)
    # conv flops (in-bounds taps ~ nominal): layer 1 1->64 at 28x28, layer 2 64->64 at 14x14, 3x3; fwd + dK (+ dInp for layer 2)
    fl = 2.0 * B2 * 9 * (2 * 64 * 784 + 3 * 64 * 64 * 196) + 3 * 2.0 * B2 * (128 * 64 * 49 + 10 * 128)
    tf = fl / (ms * 1e-3) / 1e12
    emit(({"config": "MnistCnnShaped2 secondary", "hyper": {"kh": 2, "kw": 2, "c_out": 64, "n_hidden": 128},
                      "per_gpu_batch": B2, "parameters": tr.n_params, "ms_per_step": ms, "samples_per_s": B2 / (ms * 1e-3),
                      "roofline": {"bound": "tensor", "achieved": tf, "peak": fp64.value, "unit": "TFLOP/s",
                                   "frac": tf / fp64.value, "peak_source": "hb_microbench FP64 DMMA, this run",
                                   "algorithmic_flops": fl}}))
    tr.close()
    del tr

    # -------------------------------------------------------------------- GEMM
    sec.begin("smatmul2")
    m, k, n = 512, 512, (2048 if args.quick else 8192)
    a, b = dev.sconcrete(rng.uniform(-.5, .5, (m, k))), dev.sconcrete(rng.uniform(-.5, .5, (k, n)))
    ms = timed(dev, lambda: dev.smatmul2(a, b), 20)
    tf = 2.0 * m * k * n / (ms * 1e-3) / 1e12
    emit(({"config": "smatmul2", "shape": [m, k, n], "ms": ms,
                      "roofline": {"bound": "tensor", "achieved": tf, "peak": fp64.value, "unit": "TFLOP/s",
                                   "frac": tf / fp64.value, "peak_source": "hb_microbench FP64 DMMA, this run"}}))

    # --------------------------------------------------------------------- RNN
    sec.begin("MnistRnnShaped2")
    width, batch = 512, (256 if args.quick else 1024)
    # parameters, moments and the gradient bucket as flat arrays (the trainer's layout, train.py): Adam with its
    # device-side step counter then lives INSIDE the step's CUDA graph, as one launch
    from horde_ad_b200.train import BucketLayout, exchange_scale
    host = mnist_rnn.init_params(width, rng_range=0.05)
    lay = BucketLayout([p.shape for p in host])
    flat = lay.flatten(host, 0.0)[:lay.n_params]
    fp, fm, fv = dev.sconcrete(flat), dev.sconcrete(np.zeros_like(flat)), dev.sconcrete(np.zeros_like(flat))
    t_dev = dev.sconcrete(np.zeros((), np.int64))
    params = [dev.sreshape(list(sh), dev.sslice(off, n, fp)) for sh, n, off in zip(lay.shapes, lay.sizes, lay.offsets)]
    glyphs = dev.sconcrete(rng.uniform(0, 1, (batch, 28, 28)))
    labels = dev.sconcrete(np.eye(10)[rng.integers(0, 10, batch)])

    def rnn_grad():
        return ad.grad(dev, lambda T, *ps: mnist_rnn.rnnMnistLossFusedS(T, glyphs, labels, list(ps), fused=RNN_FUSION), params)

    def rnn_apply(val, grads):
        bucket = dev.flatten_concat(list(grads) + [val])
        ffi.call("hb_allreduce_adam_step", bucket.h, lay.n_params, fp.h, fm.h, fv.h, t_dev.h, 1e-3, 0.9, 0.999, 1e-7,
                 exchange_scale(1))
        return bucket
    # eager: one C-ABI call per op, as the interpreter would issue them
    val, grads = rnn_grad()
    rnn_apply(val, grads)
    dev.sync()
    t0 = time.perf_counter()
    for _ in range(3):
        val, grads = rnn_grad()
        rnn_apply(val, grads)
    dev.sync()
    eager_ms = (time.perf_counter() - t0) / 3 * 1e3
    # the whole step -- gradient program, bucket, Adam -- captured once into a CUDA graph
    del val, grads
    ffi.call("hb_graph_begin")
    try:
        val, grads = rnn_grad()
        bucket = rnn_apply(val, grads)
    finally:
        gph = C.c_void_p()
        ffi.call("hb_graph_end", C.byref(gph))

    def rnn_step():
        ffi.call("hb_graph_launch", gph)
    step_ms = timed(dev, rnn_step, 10)
    loss = float(dev.kfromS(val))
    # flops: per time step 2 layers x (W_x x + W_s s), fwd + 2x bwd
    fl = 28 * 3 * 2.0 * batch * (28 * width + 3 * width * width)
    tf = fl / (step_ms * 1e-3) / 1e12
    emit(({"config": "MnistRnnShaped2", "width": width, "seq": 28, "per_gpu_batch": batch, "ms_per_step": step_ms,
                      "ms_per_step_eager": eager_ms,
                      "program": "rnnMnistLossFusedS with the paired-product / tanh_affine nodes selected ABOVE the C ABI "
                                 "(fused='pair'; weight cotangents as hb_matmul2_list, bias cotangents accumulated by the tanh adjoint; not the "
                                 "contraction pass); flat parameters + Adam (device step counter) inside the same CUDA graph",
                      "samples_per_s": batch / (step_ms * 1e-3), "loss": loss if np.isfinite(loss) else None,
                      "roofline": {"bound": "tensor", "achieved": tf, "peak": fp64.value, "unit": "TFLOP/s",
                                   "frac": tf / fp64.value, "peak_source": "hb_microbench FP64 DMMA, this run",
                                   "algorithmic_flops": fl}}))
    sec.end()



def rnn_dp(dev, world, rank, barrier, max_over_ranks, width=512, batch=1024, steps=8):
    """BASELINE config 5 data-parallel: MnistRnnShaped2 (width 512, 28 steps), `batch` samples PER GPU, the gradient
    program + flat bucket + the fused peer-memory all-reduce + Adam kernel captured into one CUDA graph (the same
    exchange as the CNN trainer).  Called by every rank; returns the record (max over ranks of the device time)."""
    from horde_ad_b200.train import BucketLayout, exchange_scale
    rng = np.random.Generator(np.random.Philox(4200 + rank))
    host = mnist_rnn.init_params(width, rng_range=0.05)
    lay = BucketLayout([p.shape for p in host])
    flat = lay.flatten(host, 0.0)[:lay.n_params]
    fp, fm, fv = dev.sconcrete(flat), dev.sconcrete(np.zeros_like(flat)), dev.sconcrete(np.zeros_like(flat))
    t_dev = dev.sconcrete(np.zeros((), np.int64))
    params = [dev.sreshape(list(sh), dev.sslice(off, n, fp)) for sh, n, off in zip(lay.shapes, lay.sizes, lay.offsets)]
    glyphs = dev.sconcrete(rng.uniform(0, 1, (batch, 28, 28)))
    labels = dev.sconcrete(np.eye(10)[rng.integers(0, 10, batch)])

    def step():
        val, grads = ad.grad(dev, lambda T, *ps: mnist_rnn.rnnMnistLossFusedS(T, glyphs, labels, list(ps), fused=RNN_FUSION), params)
        bucket = dev.flatten_concat(list(grads) + [val])
        ffi.call("hb_allreduce_adam_step", bucket.h, lay.n_params, fp.h, fm.h, fv.h, t_dev.h, 1e-3, 0.9, 0.999, 1e-7,
                 exchange_scale(world))
        return bucket
    b = step()                      # allocations, scratch
    dev.sync()
    del b
    barrier()
    ffi.call("hb_graph_begin")
    try:
        bucket = step()
    finally:
        gph = C.c_void_p()
        ffi.call("hb_graph_end", C.byref(gph))
    for _ in range(2):
        ffi.call("hb_graph_launch", gph)
    barrier()
    ms = C.c_float()
    ffi.call("hb_event_record", 6)
    for _ in range(steps):
        ffi.call("hb_graph_launch", gph)
    ffi.call("hb_event_record", 7)
    ffi.call("hb_event_elapsed_ms", 6, 7, C.byref(ms))
    barrier()
    step_ms = max_over_ranks(float(ms.value)) / steps
    digest = float(np.sum(dev.to_numpy(fp)))
    fl = 28 * 3 * 2.0 * batch * (28 * width + 3 * width * width)
    ffi.call("hb_graph_free", gph)
    return {"config": "MnistRnnShaped2 data-parallel", "width": width, "seq": 28, "per_gpu_batch": batch,
            "n_gpus": world, "scaling": "weak", "ms_per_step": step_ms, "samples_per_s": batch * world / (step_ms * 1e-3),
            "tflops_per_gpu": fl / (step_ms * 1e-3) / 1e12, "param_sum": digest,
            "program": "grouped products in wavefront order / accumulating tanh adjoint, selected above the C ABI (fused='wave')",
            "exchange": "hb_allreduce_adam_step (fused peer-memory all-reduce + Adam) inside the step's CUDA graph"}


if __name__ == "__main__":
    main()

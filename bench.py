#!/usr/bin/env python
"""bench.py -- the headline measurement (BASELINE.json: "MNIST CNN train
samples/s at 1/2/4/8 B200; conv VJP achieved HBM GB/s").

  python bench.py [--gpus N --steps K --warmup W] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic MNIST-shaped
input: the MnistCnnShaped2 gradient program (example/MnistCnnShaped2.hs:111-136)
interpreted over the device carrier + one all-reduce of the gradient bucket +
the Adam step (OptimizerTools.hs:134-232).  The program handed to the C ABI is
the UNCHANGED vectorised one (gather chains, sdot1In over broadcast views,
kargMax gathers, scatters: `CnnTrainer(program="vectorised")`); the library
records it once, its contraction pass rewrites it into fused nodes, and the
step replays the result inside a CUDA graph.  Batch 4096 (BASELINE config 4) is
SHARDED over the N GPUs (`--scaling strong`, the default: 4096/N samples per
rank, as SURVEY.md 8d-4 prescribes); `--scaling weak` gives every rank 4096
samples, and at N > 1 the default run reports that line too under
`weak_scaling`.  One JSON line on stdout (rank 0).

`value`  : samples/s with the minibatch already resident in HBM.
`e2e`    : the same through the public call `CnnTrainer.train_step(host ptrs)`:
           H2D of glyphs+labels from pinned host memory and the loss D2H are
           inside the timed region, every step.
`roofline`: the dominant kernel of the step, timed live with CUDA events on the
           library stream (hb_prof_*), against MEASURED_PEAKS.json.
`conv_vjp`: ConvVjpBench at BASELINE config 3 (batch 256, 64 ch, 3x3, 28x28,
           Double): fwd / dKrn / dInp achieved HBM GB/s and FP64 TFLOP/s.
`cpu_baseline`: the oracle (CPU restatement of the reference's Concrete
           backend) on a bounded sample of the same workload, on this box's
           host cores.

--impl reference: only the CPU oracle leg (the reference itself is Haskell and
cannot be built here: no GHC), same metric/config.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L2_BYTES = 126e6
PER_GPU_BATCH = 4096
HYPER = dict(kh=4, kw=4, c_out=16, n_hidden=64)     # 5x5 kernels: TestMnistCNNS.hs:422-424
METRIC = "mnist_cnn_train_samples_per_s"
# The reference's objective (whole-tensor softmax with the custom dual
# `softMax - expected`, CommonShapedOps.hs:89-111) pushes all labelled logits up
# without bound on a batch of random labels and overflows exp after ~50 Adam
# steps -- identically on the CPU oracle.  The bench
# therefore restores the initial parameters every RESET_EVERY steps (three
# 0.46 MB device-to-device copies inside the timed region) so reported losses
# stay finite; the arithmetic per step is unchanged.
RESET_EVERY = 16


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.rows = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark(self):
        """Samples taken before this point (warm-up) are dropped."""
        self.rows = self.rows[-1:]

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------ algorithmic work
def _shapes(tag: str):
    m = re.match(r"(\w+)\{(\w+)\[([\d,]*)\](?:;\[([\d,]*)\])?\}", tag)
    if not m:
        return tag, None, None, None
    f = lambda s: [int(x) for x in s.split(",") if x] if s is not None else None
    return m.group(1), m.group(2), f(m.group(3)), f(m.group(4))


def alg_work(tag: str, kh1: int, kw1: int):
    """(algorithmic bytes, flops) of one call of the C-ABI entry `tag`
    (per-unit figures x units per launch)."""
    name, dt, a, b = _shapes(tag)
    if a is None:
        return None
    s = {"f64": 8, "f32": 4, "i64": 8, "i32": 4, "i16": 2, "i8": 1, "bool": 1}[dt]
    n = lambda sh: int(np.prod(sh)) if sh is not None else 0
    if name == "conv2d_preserving":          # K ; A  -> B
        Co, Ci, KH, KW = a; N, _, H, W = b
        return (n(a) + n(b) + N * Co * H * W) * s, 2 * N * Co * H * W * Ci * KH * KW
    if name == "conv2d_preserving_dkrn":     # A ; dB -> dK
        N, Ci, H, W = a; Co = b[1]
        return (n(a) + n(b) + Co * Ci * kh1 * kw1) * s, 2 * N * Co * H * W * Ci * kh1 * kw1
    if name == "conv2d_preserving_dinp":     # K ; dB -> dA
        Co, Ci, KH, KW = a; N, _, H, W = b
        return (n(a) + n(b) + N * Ci * H * W) * s, 2 * N * Co * H * W * Ci * KH * KW
    if name == "conv_relu_pool_bwd":         # A ; dy (pooled) -> dK (+ dA): tagged launches are the dKrn / dInp kernels
        N, Ci, H, W = a; Co = b[1]
        return (n(a) + N * Co * H * W + Co * Ci * kh1 * kw1) * s, 2 * N * Co * H * W * Ci * kh1 * kw1
    if name == "conv_relu_pool_fwd":         # K ; A -> pooled out
        Co, Ci, KH, KW = a; N, _, H, W = b
        return (n(a) + n(b)) * s + N * Co * (H // 2) * (W // 2) * (s + 1), 2 * N * Co * H * W * Ci * KH * KW
    if name == "matmul2":
        return (n(a) + n(b) + a[0] * b[1]) * s, 2 * a[0] * a[1] * b[1]
    if name in ("binary", "add_inplace_or_new"):
        return 3 * n(a) * s, n(a)
    if name in ("unary", "relu", "cast", "contiguous"):
        return 2 * n(a) * s, n(a)
    if name == "maxpool2d":                  # x -> y (1/4) + argmax i64 (1/4)
        return n(a) * s + n(a) // 4 * (s + 8), n(a)
    if name == "maxpool2d_bwd":              # dy + argmax -> dx (4x)
        return n(a) * (s + 8) + 4 * n(a) * s, n(a)
    if name in ("sum_leading", "sum_all"):
        return n(a) * s, n(a)
    if name == "adam_step":                  # p, m, v read+write, g read
        return 7 * n(a) * s, 12 * n(a)
    return None


# ------------------------------------------------------- CPU oracle workers
def _cpu_worker(args):
    """One host core: `steps` oracle training steps (dual-number gradient of
    convMnistLossFusedS + Adam) on its own shard of `batch` samples."""
    seed, batch, steps = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from horde_ad_b200 import ad, mnist_cnn
    from oracle.concrete import ConcreteBackend
    orc = ConcreteBackend()
    rng = np.random.Generator(np.random.Philox(seed))
    gl = rng.uniform(0, 1, (batch, 28, 28))
    lb = np.eye(10)[rng.integers(0, 10, batch)]
    ps = mnist_cnn.init_params(HYPER["kh"], HYPER["kw"], HYPER["c_out"], HYPER["n_hidden"])
    ms = [np.zeros_like(p) for p in ps]
    vs = [np.zeros_like(p) for p in ps]
    val = 0.0
    for t in range(steps):
        val, grads = ad.grad(orc, lambda T, *p: mnist_cnn.convMnistLossFusedS(T, gl, lb, list(p), fused=True), ps)
        for p, m, v, g in zip(ps, ms, vs, grads):
            orc.adam_step(p, m, v, np.asarray(g), t + 1)
    return float(val)


def cpu_oracle_rate(per_worker_batch: int, steps: int, workers: int, warm: int = 0):
    """samples/s of the oracle with `workers` processes, each owning a shard
    (the data-parallel decomposition of the GPU arm, on host cores); the
    reference's own interpreter is single-threaded, this is every core it
    could be given."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        pool.map(_cpu_worker, [(100 + i, 8, 1) for i in range(workers)])          # start-up (imports, first touch)
        if warm:
            pool.map(_cpu_worker, [(42 + i, per_worker_batch, warm) for i in range(workers)])   # untimed warm-up steps
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(42 + i, per_worker_batch, steps) for i in range(workers)])
        dt = time.perf_counter() - t0
    return workers * per_worker_batch * steps / dt, dt


# ---------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's own CPU path: the oracle restatement of the Concrete
    backend (numpy; the Haskell reference cannot be built here), same program
    (dual-number gradient of convMnistLossFusedS + Adam), bounded sample."""
    if rank != 0:
        return
    workers = args.cpu_workers or (os.cpu_count() or 1)
    # EXACTLY --steps timed and --warmup untimed steps; the per-step sample is bounded so that the whole run stays
    # within a couple of minutes (~150 samples/s per core for this model on the boxes seen so far)
    steps, warm = max(1, args.steps), max(0, args.warmup)
    per = max(8, min(args.cpu_sample // workers, int(120 * 150 / (steps + warm))))
    val, dt_total = cpu_oracle_rate(per, steps, workers, warm)
    dt = dt_total / steps
    sample = per * workers
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, world),
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": workers, "kind": "port",
                         "sample": f"{steps} steps of {workers} shards x {per} samples (of {PER_GPU_BATCH}); numpy oracle, "
                                   f"one process per host core"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "host": {"nproc": os.cpu_count()},
    }
    print(json.dumps(out), flush=True)


def config_dict(args, world):
    local = PER_GPU_BATCH if args.scaling == "weak" else PER_GPU_BATCH // world
    return {"workload": "MnistCnnShaped2 training step (convMnistLossFusedS gradient + Adam), synthetic MNIST 28x28; "
                        "batch 4096 sharded over the GPUs" + ("" if args.scaling == "strong" else " -- WEAK: 4096 per GPU"),
            "per_gpu_batch": local, "global_batch": local * world,
            "hyper": "kh=kw=4 (5x5 kernels), c_out=16, n_hidden=64",
            "parallelism": f"dp{world}", "optimizer": "adam",
            "param_reset_every_steps": RESET_EVERY,
            "l2": "per-step working set (activations ~1.6 GB at batch 4096) >> 126 MB L2; no explicit flush"}


# -------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--program", default="vectorised", choices=["vectorised", "vectorised-eager", "fused"],
                    help="what the step hands to the C ABI (see CnnTrainer)")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="gradient exchange: the one-kernel peer-memory all-reduce + Adam, or ncclAllReduce + scale + Adam")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the other BASELINE configs (LongProd, FCNN, RNN, secondary CNN, kernel families)")
    ap.add_argument("--no-extras", action="store_true", help="skip the drop-in / hand-fused / weak-scaling side measurements")
    ap.add_argument("--cpu-sample", type=int, default=8192, help="samples per step of the bounded CPU-baseline sample")
    ap.add_argument("--cpu-workers", type=int, default=0, help="host processes of the CPU baseline (0 = all cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-conv-vjp", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--breakdown-file", default=None, help="write the full per-launch-site table of one step here")
    ap.add_argument("--shard-batch", type=int, default=0,
                    help="diagnostic: run THIS many samples per GPU (e.g. 512 on one GPU = the N=8 strong-scaling shard)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # convenience: relaunch under torchrun
            port = os.environ.get("MASTER_PORT", "29517")
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", port, __file__] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)   # control plane only

    from horde_ad_b200 import ffi
    from horde_ad_b200.device import DeviceBackend
    from horde_ad_b200.train import CnnTrainer, PinnedBuffer, comm_init

    def dbg(msg):
        if os.environ.get("HB_BENCH_DEBUG"):
            print(f"[rank {rank}] {time.strftime('%H:%M:%S')} {msg}", file=sys.stderr, flush=True)

    dev = DeviceBackend(local_rank)      # raises without a GPU / without the built library
    dbg("device up")

    def bcast(b):
        objs = [b]
        dist.broadcast_object_list(objs, src=0)
        return objs[0]
    # NCCL writes its version banner to stdout at communicator creation; stdout
    # carries exactly one JSON line, so fd 1 points at stderr while NCCL initialises
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        comm_init(rank, world, bcast if world > 1 else None)
        if world > 1:       # first collective (lazy channel set-up may print too)
            warm = dev.sconcrete(np.zeros(8))
            ffi.call("hb_allreduce_sum", warm.h, 1.0)
            dev.sync()
    finally:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(saved_fd)
    dbg("comm up")

    def barrier():
        dev.sync()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    local_batch = PER_GPU_BATCH if args.scaling == "weak" else PER_GPU_BATCH // world
    if args.shard_batch:
        local_batch = args.shard_batch
    global_batch = local_batch * world
    tr = CnnTrainer(dev, local_batch, **HYPER, world=world, use_graph=not args.no_graph, program=args.program,
                    exchange=args.exchange)
    rng = np.random.Generator(np.random.Philox(42 + 1000 * rank))
    gl = PinnedBuffer((local_batch, 28, 28), np.float64)
    lb = PinnedBuffer((local_batch, 10), np.float64)
    gl.array[...] = rng.uniform(0, 1, gl.array.shape)
    lb.array[...] = np.eye(10)[rng.integers(0, 10, local_batch)]
    h2d = gl.nbytes + lb.nbytes

    K, W = args.steps, max(3, args.warmup)
    n0 = C.c_uint64()

    # ---- device-resident: inputs uploaded once, K steps timed with CUDA events
    clocks = ClockSampler(local_rank)
    clocks.start()
    tr.upload(gl.ptr, lb.ptr)
    for i in range(W):
        tr.step_enqueue()
        if i == 0:
            dev.sync()
            dbg("first step done")
    barrier()
    dbg("warm-up done")
    clocks.mark()
    ffi.call("hb_launch_count", C.byref(n0))
    ffi.call("hb_event_record", 0)
    for i in range(K):
        if i % RESET_EVERY == 0:
            tr.reset()
        bucket = tr.step_enqueue()
    ffi.call("hb_event_record", 1)
    ms = C.c_float()
    ffi.call("hb_event_elapsed_ms", 0, 1, C.byref(ms))
    n1 = C.c_uint64()
    ffi.call("hb_launch_count", C.byref(n1))
    barrier()
    dbg("resident timing done")
    loss_resident = tr.loss_of(bucket)
    dev_ms = max_over_ranks(float(ms.value))
    value = global_batch * K / (dev_ms * 1e-3)

    # ---- end to end through the public call: H2D + step + loss D2H per step
    # two pinned host batches alternate, as a data loader would hand them out;
    # step k uploads batch k+1 on the copy stream while it computes (every
    # step still performs one full H2D of a minibatch and one loss D2H)
    gl2 = PinnedBuffer((local_batch, 28, 28), np.float64)
    lb2 = PinnedBuffer((local_batch, 10), np.float64)
    gl2.array[...] = gl.array[::-1]
    lb2.array[...] = lb.array[::-1]
    hostb = [(gl.ptr, lb.ptr), (gl2.ptr, lb2.ptr)]
    for i in range(2):
        tr.train_step(*hostb[i % 2], *hostb[(i + 1) % 2])
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        if i % RESET_EVERY == 0:
            tr.reset()
        loss = tr.train_step(*hostb[i % 2], *hostb[(i + 1) % 2])
    dev.sync()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    dbg("e2e timing done")
    clk = clocks.stop()
    e2e_value = global_batch * K / e2e_s

    # ---- the same step fed by the device-resident input pipeline: 60 000 synthetic
    # glyphs as raw bytes in HBM, every minibatch assembled by hb_mnist_batch_into from the epoch's
    # shuffle; only the 8-byte loss crosses PCIe (reported next to e2e, not instead of it)
    from horde_ad_b200 import mnist_data
    drng = np.random.Generator(np.random.Philox(4242 + rank))
    dmn = mnist_data.DeviceMnist(dev, drng.integers(0, 256, (60000, 28, 28), dtype=np.uint8),
                                 drng.integers(0, 10, 60000, dtype=np.uint8))
    dmn.shuffle(42 + rank)
    nb = dmn.n_batches(local_batch)
    for i in range(2):
        tr.train_step_device(dmn, i % nb)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        if i % RESET_EVERY == 0:
            tr.reset()
        tr.train_step_device(dmn, i % nb)
    dev.sync()
    resident_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    del dmn

    # ---- cross-rank consistency: after the timed loops every rank must hold bit-identical parameters
    # (replicated weights + identical all-reduced bucket + identical Adam step)
    import hashlib
    digest = hashlib.sha256(np.concatenate([p.reshape(-1) for p in tr.get_params()]).tobytes()).hexdigest()
    params_identical = True
    if dist is not None:
        objs = [None] * world
        dist.all_gather_object(objs, digest)
        params_identical = len(set(objs)) == 1
        if not params_identical:
            raise SystemExit(f"[rank {rank}] parameters differ across ranks after the timed loop: {objs}")
    pass_stats = tr.prog.stats if tr.prog is not None else None

    # ---- side measurements (every rank runs them: the step contains the all-reduce)
    def time_program(program, steps, warm, batch=None, graph=True, exchange=None, objective="shard"):
        t2 = CnnTrainer(dev, batch or local_batch, **HYPER, world=world, use_graph=graph, program=program,
                        exchange=exchange or args.exchange, objective=objective)
        g2, l2 = gl, lb
        if batch and batch != local_batch:
            g2 = PinnedBuffer((batch, 28, 28), np.float64)
            l2 = PinnedBuffer((batch, 10), np.float64)
            g2.array[...] = rng.uniform(0, 1, g2.array.shape)
            l2.array[...] = np.eye(10)[rng.integers(0, 10, batch)]
        t2.upload(g2.ptr, l2.ptr)
        for _ in range(warm):
            t2.step_enqueue()
        barrier()
        ffi.call("hb_event_record", 4)
        for i in range(steps):
            if i % RESET_EVERY == 0:
                t2.reset()
            b2 = t2.step_enqueue()
        ffi.call("hb_event_record", 5)
        m2 = C.c_float()
        ffi.call("hb_event_elapsed_ms", 4, 5, C.byref(m2))
        t2.reset()                      # loss of the FIRST step from the initial parameters: comparable across programs
        l = t2.loss_of(t2.step_enqueue())
        t2.close()
        del t2
        return max_over_ranks(float(m2.value)) / steps, l
    extras = {}
    if not args.no_extras and args.program == "vectorised":
        # the same recorded calls launched op by op as they arrive, WITHOUT the contraction pass: the raw drop-in step
        extras["dropin_unfused_ms"], l_raw = time_program("vectorised-eager", 2, 1, graph=False)
        # round 1's hand-fused model (fused=True above the ABI): what the pass has to reach
        extras["hand_fused_ms"], l_fused = time_program("fused", min(K, 20), 3)
        tr.upload(gl.ptr, lb.ptr)       # the e2e / device-pipeline loops left other batches in the input buffers
        tr.reset()
        l_rew = tr.loss_of(tr.step_enqueue())
        extras["first_step_loss"] = {"rewritten": l_rew, "hand_fused": l_fused, "unfused_eager": l_raw,
                                     "what": "same inputs, initial parameters: the three programs compute one value"}
        if world > 1:
            # the same step with ncclAllReduce + scale kernel + Adam kernel instead of the fused exchange kernel
            other = "nccl" if args.exchange == "fused" else "fused"
            extras["exchange_comparison"] = {"this_run": args.exchange, "other": other,
                                             "other_ms_per_step": time_program("vectorised", min(K, 20), 3, exchange=other)[0]}
        if world > 1:
            # the step whose result equals the reference's one-device step on the whole batch: min u / sum expU of
            # the loss exchanged across the ranks inside the loss kernel (train.py objective="global")
            gms, gl_ = time_program("vectorised", min(K, 20), 3, objective="global")
            extras["global_objective"] = {"ms_per_step": gms, "value": global_batch / (gms * 1e-3), "unit": "samples/s",
                                          "first_step_loss": gl_,
                                          "what": "N-rank step == reference step on the whole batch (2-GPU test: 1e-12 vs oracle)"}
        if world > 1:
            import bench_configs
            try:
                extras["rnn_dp"] = bench_configs.rnn_dp(dev, world, rank, barrier, max_over_ranks)
                sums = [None] * world
                dist.all_gather_object(sums, extras["rnn_dp"]["param_sum"])
                extras["rnn_dp"]["params_identical_across_ranks"] = len(set(sums)) == 1
            except Exception as e:      # a side measurement must not take the headline line down with it
                extras["rnn_dp"] = {"error": repr(e)[:300]}
        if world > 1 and args.scaling == "strong":
            wms, _ = time_program("vectorised", min(K, 20), 3, batch=PER_GPU_BATCH)
            extras["weak_scaling"] = {"per_gpu_batch": PER_GPU_BATCH, "ms_per_step": wms,
                                      "value": PER_GPU_BATCH * world / (wms * 1e-3), "unit": "samples/s"}

    # ---- live per-kernel timing of the same step (eager, events per launch);
    # every rank runs it (the step contains the all-reduce), rank 0 reports
    prof = profile_step(dev, ffi, CnnTrainer, local_batch, gl, lb, world, program=args.program, exchange=args.exchange)
    barrier()
    dbg("profile done")
    if rank != 0:
        tr.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    hbm_peak, peak_src, _ = peaks()
    fp64_dmma, fp64_fma = C.c_double(), C.c_double()
    ffi.call("hb_microbench", 1, C.byref(fp64_dmma))
    ffi.call("hb_microbench", 0, C.byref(fp64_fma))
    fp64_peak = max(fp64_dmma.value, fp64_fma.value)
    roof, breakdown = roofline_of(prof, hbm_peak, peak_src, fp64_peak)
    if args.breakdown_file:
        with open(args.breakdown_file, "w") as f:
            for r in prof:
                f.write(f'{r["ms_per_step"]:9.4f} ms {100 * r["share"]:5.1f}% n={r["launches_per_step"]:4.1f} {r["entry"]} @ {r["site"]}\n')

    out = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_dict(args, world),
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": 8 * world, "ms_per_step": e2e_s / K * 1e3},
        "device_input_pipeline": {"value": global_batch * K / resident_s, "unit": "samples/s",
                                  "ms_per_step": resident_s / K * 1e3, "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": 8 * world,
                                  "what": "minibatches assembled on the device from resident IDX bytes (hb_mnist_batch_into)"},
        "gpu_launches": int(n1.value - n0.value),
        "clocks": clk, "loss": loss if np.isfinite(loss) else None,
        "loss_resident": loss_resident if np.isfinite(loss_resident) else None,
        "roofline": roof, "breakdown": breakdown,
        "fp64_peak_tflops": {"dmma": fp64_dmma.value, "dfma": fp64_fma.value, "how": "hb_microbench, this run"},
        "program": args.program, "exchange": args.exchange,
        "contraction_pass": ({"recorded_nodes": pass_stats[0], "nodes_after_pass": pass_stats[1]} if pass_stats else None),
        "params_identical_across_ranks": params_identical,
    }
    out.update(extras)
    if not args.no_conv_vjp:
        out["conv_vjp"] = conv_vjp_bench(dev, ffi, hbm_peak, fp64_peak)
        f32_peak, f32_src = f32_tensor_peak()
        out["conv_vjp_f32"] = conv_vjp_bench(dev, ffi, hbm_peak, f32_peak, dtype=np.float32)
        out["conv_vjp_f32"]["engine"] = "tcgen05.mma kind::tf32, three-term split (csrc/contract_tf32.cu)"
        out["conv_vjp_f32"]["tensor_peak"] = {"tflops": f32_peak, "source": f32_src}
        out["smatmul2_f32"] = f32_matmul_bench(dev, ffi)
    if not args.no_other_configs and world == 1:
        # the other BASELINE configs (LongProdBench, ShortMnistForCI, RNN, secondary CNN, kernel families), each group
        # with its own nvidia-smi clock samples: the numbers bench_configs.py prints, on the driver's record
        import contextlib
        import bench_configs
        recs = []
        cur = {"clk": None, "name": None, "first": 0}

        @contextlib.contextmanager
        def section(name):
            smp = ClockSampler(local_rank)
            smp.start()
            first = len(recs)
            try:
                yield
            finally:
                dev.sync()
                c = smp.stop()
                for r in recs[first:]:
                    r["section"] = name
                    r["clocks"] = c
        t0 = time.perf_counter()
        try:
            bench_configs.run_configs(dev, quick=False, emit=recs.append, section=section)
        except Exception as e:        # a side measurement must not take the headline line down with it
            recs.append({"config": "error", "error": repr(e)[:300]})
        out["other_configs"] = recs
        out["other_configs_wall_s"] = time.perf_counter() - t0
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(args.cpu_sample, args.cpu_workers)
    print(json.dumps(out), flush=True)
    tr.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def profile_step(dev, ffi, CnnTrainer, local_batch, gl, lb, world, steps=3, program="vectorised", exchange="fused"):
    """Per-launch CUDA-event timing of the step replayed node by node (no CUDA graph)."""
    tr = CnnTrainer(dev, local_batch, **HYPER, world=world, use_graph=False, program=program, exchange=exchange)
    tr.upload(gl.ptr, lb.ptr)
    for _ in range(2):
        tr.step_enqueue()
    dev.sync()
    ffi.call("hb_prof_begin")
    for _ in range(steps):
        tr.step_enqueue()
    ffi.call("hb_prof_end")
    buf = C.create_string_buffer(1 << 16)
    ffi.call("hb_prof_report", buf, len(buf))
    rows = []
    for line in buf.value.decode().splitlines():
        tag, site, cnt, tot = line.rsplit(" ", 3)
        rows.append({"entry": tag, "site": site, "launches_per_step": int(cnt) / steps, "ms_per_step": float(tot) / steps})
    return rows


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed
# `ncu --set full` capture (profiles/r01_conv_kernels_v5_full_raw.csv), keyed
# by (launch-site function, operand shapes of the entry)
NCU_TRAFFIC = {   # (launch site, C-ABI entry) -> dram__bytes_read.sum + dram__bytes_write.sum per launch
    ("hb_conv_dk_dmma", "conv_relu_pool_bwd"): 242468096 + 5267200,    # conv2 dKrn (two row bands per image: halo rows read twice)
    ("hb_conv_sp_dmma", "conv_relu_pool_fwd"): 102900992 + 14155520,   # conv2 fwd + bias/relu/pool
    ("hb_conv_sp_dmma", "conv_relu_pool_bwd"): 102893824 + 58643200,   # conv2 dInp (rest of the write still in L2)
    ("hb_conv_strip_fwd", "conv_relu_pool_fwd"): 25750528 + 58959616,  # conv1 fwd + bias/relu/pool
    ("hb_conv_strip_dk", "conv_relu_pool_bwd"): 141343232 + 3762944,   # conv1 dKrn + bias cotangent
}
NCU_TRAFFIC_SOURCE = "profiles/r01_conv_kernels_v5_full_raw.csv (ncu --set full, per launch)"


def ncu_traffic(entry: str, site: str):
    for (fn, name), by in NCU_TRAFFIC.items():
        if site.startswith(fn) and entry.startswith(name + "{") and "[4096," in entry:   # captured at per-GPU batch 4096
            return by
    return None


def issued_ratio(tag: str, kh1: int) -> float:
    """Fraction of the nominal conv flops whose tap row meets the image at all (the DMMA kernels skip the others)."""
    name, dt, a, b = _shapes(tag)
    if a is None or not name.startswith("conv"):
        return 1.0
    H = a[2] if name in ("conv_relu_pool_bwd", "conv2d_preserving_dkrn") else (b[2] if b else a[2])
    KH = a[2] if name in ("conv2d_preserving", "conv2d_preserving_dinp", "conv_relu_pool_fwd") else kh1
    if not H or not KH:
        return 1.0
    return sum(max(0, H - k) for k in range(KH)) / (H * KH)


def roofline_of(rows, hbm_peak, peak_src, fp64_peak):
    kh1 = kw1 = HYPER["kh"] + 1
    total = sum(r["ms_per_step"] for r in rows)
    for r in rows:
        r["share"] = r["ms_per_step"] / total if total else 0.0
    # group by entry (an entry may launch a pad/copy kernel besides its main kernel)
    top = rows[0]
    w = alg_work(top["entry"], kh1, kw1)
    per_launch_ms = top["ms_per_step"] / max(1.0, top["launches_per_step"])
    roof = {"kernel": f'{top["entry"]} @ {top["site"]}', "share_of_step": top["share"],
            "launch_ms": per_launch_ms}
    if w is not None:
        by, fl = w
        gbs = by / (per_launch_ms * 1e-3) / 1e9
        tfs = fl / (per_launch_ms * 1e-3) / 1e12
        ai = fl / by
        ridge = fp64_peak * 1e12 / (hbm_peak * 1e9)
        if ai > ridge:
            roof.update({"bound": "tensor", "achieved": tfs, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": tfs / fp64_peak, "frac_issued": tfs * issued_ratio(top["entry"], kh1) / fp64_peak,
                         "frac_issued_note": "flops of the DMMAs actually issued (tap rows that only meet the zero "
                                             "padding are skipped: sum_kh (H - kh) / (H KH) of the nominal ones) / peak", "peak_source": "hb_microbench FP64 (DMMA/DFMA) this run; "
                         "MEASURED_PEAKS.json has no FP64 figure", "algorithmic_flops": fl,
                         "hbm_view": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                      "algorithmic_bytes": by, "peak_source": peak_src}})
        else:
            roof.update({"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                         "peak_source": peak_src, "algorithmic_bytes": by})
        roof["traffic"] = ncu_traffic(top["entry"], top["site"])
        roof["traffic_source"] = NCU_TRAFFIC_SOURCE
    breakdown = [{"entry": r["entry"], "site": r["site"], "ms": round(r["ms_per_step"], 4),
                  "share": round(r["share"], 4), "n": r["launches_per_step"]} for r in rows[:12]]
    return roof, breakdown


def conv_vjp_bench(dev, ffi, hbm_peak, fp64_peak, reps=10, dtype=np.float64):
    """ConvVjpBench at BASELINE config 3: K[64,64,3,3], A,dB[256,64,28,28], Double (the reference's benchmark type)
    or Float (the tcgen05 3xTF32 engine; `fp64_peak` is then the Float tensor peak to compare with)."""
    rng = np.random.Generator(np.random.Philox(42))
    K = dev.sconcrete(rng.uniform(-.5, .5, (64, 64, 3, 3)).astype(dtype))
    A = dev.sconcrete(rng.uniform(-.5, .5, (256, 64, 28, 28)).astype(dtype))
    dB = dev.sconcrete(rng.uniform(-.5, .5, (256, 64, 28, 28)).astype(dtype))
    by = (2 * 256 * 64 * 28 * 28 + 64 * 64 * 9) * np.dtype(dtype).itemsize
    fl = 2 * 256 * 64 * 28 * 28 * 64 * 9
    out = {"config": f"K[64,64,3,3], A,dB[256,64,28,28] {'f64' if dtype == np.float64 else 'f32'}",
           "algorithmic_bytes_per_pass": by, "algorithmic_flops_per_pass": fl}
    ms = C.c_float()
    for name, f in (("fwd", lambda: dev.conv2dPreserving(K, A)), ("dkrn", lambda: dev.conv2d_dkrn(A, dB, 3, 3)),
                    ("dinp", lambda: dev.conv2d_dinp(K, dB))):
        for _ in range(3):
            f()
        dev.sync()
        ffi.call("hb_event_record", 2)
        for _ in range(reps):
            f()
        ffi.call("hb_event_record", 3)
        ffi.call("hb_event_elapsed_ms", 2, 3, C.byref(ms))
        t = ms.value / reps * 1e-3
        out[name] = {"ms": t * 1e3, "hbm_gbs": by / t / 1e9, "hbm_frac": by / t / 1e9 / hbm_peak,
                     "tflops": fl / t / 1e12,
                     ("fp64_frac" if dtype == np.float64 else "tensor_frac"): fl / t / 1e12 / fp64_peak}
    return out


def f32_tensor_peak():
    """Float contractions run as three kind::tf32 MMAs per product: peak = (TF32 dense = 1/2 of the measured bf16
    figure of MEASURED_PEAKS.json) / 3."""
    _, _, d = peaks()
    bf16 = float(d.get("bf16_tflops", 1590.0))
    return bf16 / 2 / 3, ("measured bf16 burst (MEASURED_PEAKS.json) / 2 (TF32) / 3 (three-term split)"
                          if d else "fallback bf16 1590 / 2 / 3")


def f32_matmul_bench(dev, ffi, reps=10):
    """smatmul2 in Float on the tcgen05 engine: [4096,4096] x [4096,4096]."""
    rng = np.random.Generator(np.random.Philox(43))
    n = 4096
    a = dev.sconcrete(rng.uniform(-.5, .5, (n, n)).astype(np.float32))
    b = dev.sconcrete(rng.uniform(-.5, .5, (n, n)).astype(np.float32))
    ms = C.c_float()
    for _ in range(3):
        dev.smatmul2(a, b)
    dev.sync()
    ffi.call("hb_event_record", 2)
    for _ in range(reps):
        dev.smatmul2(a, b)
    ffi.call("hb_event_record", 3)
    ffi.call("hb_event_elapsed_ms", 2, 3, C.byref(ms))
    t = ms.value / reps * 1e-3
    peak, src = f32_tensor_peak()
    tf = 2.0 * n ** 3 / t / 1e12
    return {"shape": [n, n, n], "ms": t * 1e3,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                         "peak_source": src}}


def cpu_baseline(sample: int, workers: int = 0):
    """The oracle on the box's host cores, bounded sample (same program)."""
    workers = workers or (os.cpu_count() or 1)
    per = max(8, sample // workers)
    rate, dt = cpu_oracle_rate(per, 2, workers)
    return {"value": rate, "unit": "samples/s", "cores": workers, "kind": "port",
            "sample": f"2 steps of {workers} shards x {per} samples (of {PER_GPU_BATCH}); numpy oracle, one process per "
                      f"host core ({dt:.1f} s)", "host_nproc": os.cpu_count()}


if __name__ == "__main__":
    main()

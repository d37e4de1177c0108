"""Batch-sharded MNIST CNN training loop over the device carrier (SURVEY.md 8e;
BASELINE.json config 4): every rank owns batch/world samples, interprets the
same gradient program (the unchanged vectorised program, rewritten into fused nodes by the library's
contraction pass) over its shard, and ONE all-reduce of the flattened gradient bucket (plus the loss
in its last slot) over NVLink precedes the identical Adam step on every rank.

The whole step (gradient program + bucket + all-reduce + optimizer) is
recorded once into a CUDA graph; a training step is then
  hb_upload_into (glyphs, labels from pinned host memory)  ->  hb_graph_launch
  ->  one scalar read-back.
Parameters, Adam moments and the bucket are leaves/views of flat device
buffers so the optimizer is a single fused pass (OptimizerTools.hs:134-232).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import ad, ffi, mnist_cnn
from .device import DeviceArray, DeviceBackend
from .ffi import call


def shard_range(rank: int, world: int, batch: int) -> tuple[int, int]:
    """Rank r owns samples [r*B/G, (r+1)*B/G); the batch must split evenly
    (the loss is a mean over the shard, so unequal shards would bias it)."""
    if batch % world != 0:
        raise ValueError(f"batch {batch} does not split evenly over {world} ranks")
    per = batch // world
    return rank * per, (rank + 1) * per


def exchange_scale(world: int, objective: str = "shard") -> float:
    """Factor applied to the all-reduce SUM.  objective="shard": the mean of per-shard mean losses / gradients
    (MnistCnnShaped2.hs:136 applied to every shard).  objective="global": every shard already carries the global
    `recip batch` factor and the globally normalised softmax, so the buckets are summed."""
    return 1.0 if objective == "global" else 1.0 / world


class BucketLayout:
    """Offsets of the gradient leaves of the TKProduct (MnistCnnShaped2.hs:30-40, in parameter order) inside
    the one flat bucket that is exchanged; one extra trailing slot carries the loss.  The same object lays out
    the flat parameter / Adam-moment buffers of `CnnTrainer` (its leaves are views at these offsets), and the
    world_size-2 gloo test drives it on the CPU."""

    def __init__(self, shapes):
        self.shapes = [tuple(s) for s in shapes]
        self.sizes = [int(np.prod(s)) if len(s) else 1 for s in self.shapes]
        self.offsets = [0]
        for n in self.sizes:
            self.offsets.append(self.offsets[-1] + n)
        self.n_params = self.offsets[-1]
        self.n_total = self.n_params + 1

    def flatten(self, leaves, loss: float) -> np.ndarray:
        assert len(leaves) == len(self.shapes)
        out = np.empty(self.n_total, dtype=np.asarray(leaves[0]).dtype)
        for leaf, off, n, sh in zip(leaves, self.offsets, self.sizes, self.shapes):
            leaf = np.asarray(leaf)
            assert tuple(leaf.shape) == sh, (leaf.shape, sh)
            out[off:off + n] = leaf.reshape(-1)
        out[self.n_params] = loss
        return out

    def unflatten(self, bucket: np.ndarray):
        leaves = [bucket[o:o + n].reshape(sh) for o, n, sh in zip(self.offsets, self.sizes, self.shapes)]
        return leaves, float(bucket[self.n_params])


class PinnedBuffer:
    """Pinned host staging memory exposed as a numpy array."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        call("hb_host_alloc", self.nbytes, C.byref(p))
        self.ptr = p
        buf = (C.c_char * self.nbytes).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr is not None:
            self.array = None
            ffi.lib().hb_host_free(self.ptr)
            self.ptr = None


def _addr(p) -> int:
    return p.value if isinstance(p, C.c_void_p) else int(p)


def _devptr(t: DeviceArray) -> C.c_void_p:
    p = C.c_void_p()
    call("hb_device_ptr", t.h, C.byref(p))
    return p


class CnnTrainer:
    """`program` selects what is handed to the C ABI each step:
      "vectorised" (default) -- the UNCHANGED program horde-ad sends: chained i+j gathers, sdot1In over broadcast
          views, [0.0,1.0]-table relu masks, kargMax gathers and their adjoint scatters (`fused=False` model +
          the dual-number sweep of ad.py).  It is RECORDED once through the C ABI (hb_lazy_begin/end), rewritten by
          the library's contraction pass into whole-layer nodes, and replayed (inside a CUDA graph);
      "vectorised-eager"     -- the same calls launched op by op as they arrive (no pass): the raw drop-in cost;
      "fused"                -- round 1's hand-fused model (`fused=True`): the pass's target, kept as a cross-check."""

    def __init__(self, dev: DeviceBackend, local_batch: int, kh=4, kw=4, c_out=16, n_hidden=64,
                 dtype=np.float64, world=1, seed=44, use_graph=True, optimizer="adam", gamma=0.02,
                 program="vectorised", exchange="fused", objective="shard"):
        assert program in ("vectorised", "vectorised-eager", "fused")
        assert exchange in ("fused", "nccl")
        assert objective in ("shard", "global")
        self.program = program
        # "shard": every rank optimises the reference objective of ITS shard (softmax normalised over the shard) and
        #          the gradients are averaged -- the decomposition SURVEY.md 8e prescribes;
        # "global": the N-rank step IS the reference's step on the whole batch: `minimum u` and `sum expU` of
        #          lossSoftMaxCrossEntropyS (CommonShapedOps.hs:89-111) are exchanged across the ranks inside the loss
        #          kernel (hb_comm_set_global_objective), every shard uses 1 / global batch, the buckets are summed
        self.objective = objective if world > 1 else "shard"
        if self.objective == "global":
            call("hb_comm_set_global_objective", 1)
        # "fused": gradient exchange + 1/world + Adam in ONE kernel over peer-mapped buffers (hb_allreduce_adam_step,
        # deterministic rank-order sum, step counter on the device, so the whole step is one graph launch);
        # "nccl": ncclAllReduce, then the scale kernel, then the Adam kernel (round 1's path, the comparison arm)
        self.exchange = exchange if optimizer == "adam" else "nccl"
        self.prog = None
        self.dev, self.B, self.world, self.dtype = dev, local_batch, world, np.dtype(dtype)
        self.optimizer, self.gamma = optimizer, gamma
        self.hyper = (kh, kw, c_out, n_hidden)
        host_params = mnist_cnn.init_params(kh, kw, c_out, n_hidden, seed=seed, dtype=dtype)
        self.layout = BucketLayout([p.shape for p in host_params])
        self.shapes, self.sizes, self.n_params = self.layout.shapes, self.layout.sizes, self.layout.n_params
        flat = self.layout.flatten(host_params, 0.0)[:self.n_params]
        self.flat_p = dev.sconcrete(flat)
        self.flat_m = dev.sconcrete(np.zeros_like(flat))
        self.flat_v = dev.sconcrete(np.zeros_like(flat))
        self._p0 = dev.sconcrete(flat)                    # initial parameters, for reset()
        self._z0 = dev.sconcrete(np.zeros_like(flat))
        self.params = self._leaves(self.flat_p)
        self.glyphs = dev.sconcrete(np.zeros((local_batch, 28, 28), dtype))
        self.labels = dev.sconcrete(np.zeros((local_batch, 10), dtype))
        self.t = 0
        self.t_dev = dev.sconcrete(np.zeros((), np.int64))     # Adam's step count on the device (fused exchange)
        self._t0 = dev.sconcrete(np.zeros((), np.int64))
        self.graph = None
        # the graph holds the gradient program + bucket + all-reduce; Adam's
        # bias-corrected step size depends on t, so it is one eager launch after it
        self.use_graph = use_graph
        self.loss_arr = None
        self.bucket = None
        self._prefetched = None
        self._stage_g = self._stage_l = None

    def _leaves(self, flat: DeviceArray):
        return [self.dev.sreshape(list(sh), self.dev.sslice(off, n, flat))
                for sh, n, off in zip(self.layout.shapes, self.layout.sizes, self.layout.offsets)]

    # ---- one step, enqueued op by op -------------------------------------
    def _gradient_program(self):
        dev = self.dev
        fused = self.program == "fused"
        loss, grads = ad.grad(
            dev, lambda T, *ps: mnist_cnn.convMnistLossFusedS(
                T, self.glyphs, self.labels, list(ps), fused=fused,
                loss_scale=(1.0 / (self.B * self.world) if self.objective == "global" else None)),
            self.params)
        bucket = dev.flatten_concat(list(grads) + [loss])
        if self.exchange == "fused":
            call("hb_allreduce_adam_step", bucket.h, self.n_params, self.flat_p.h, self.flat_m.h, self.flat_v.h,
                 self.t_dev.h, 1e-3, 0.9, 0.999, 1e-7, exchange_scale(self.world, self.objective))
        else:
            call("hb_allreduce_sum", bucket.h, exchange_scale(self.world, self.objective))
        return bucket

    def _apply(self, bucket):
        dev = self.dev
        if self.exchange == "fused":
            self.t += 1                  # the update itself happened inside the exchange kernel
            return
        g = dev.sslice(0, self.n_params, bucket)
        if self.optimizer == "adam":
            self.t += 1
            dev.adam_step(self.flat_p, self.flat_m, self.flat_v, g, self.t)
        else:
            dev.sgd_step(self.flat_p, g, self.gamma)

    def step_enqueue(self):
        """Enqueue one training step on the inputs currently in self.glyphs /
        self.labels; returns the bucket whose last element is the mean loss."""
        if self.use_graph and self.program != "vectorised-eager":
            if self.graph is None:
                self._capture()
            call("hb_graph_launch", self.graph)
            if self.optimizer == "adam":
                self._apply(self.bucket)
            return self.bucket
        if self.program == "vectorised":        # the rewritten program, replayed node by node (no CUDA graph)
            if self.prog is None:
                self.prog = self.dev.record(self._gradient_program)
                self.bucket = self.prog.outs[0]
            self.prog.run()
            self._apply(self.bucket)
            return self.bucket
        bucket = self._gradient_program()
        self._apply(bucket)
        self.bucket = bucket
        return bucket

    def _state_snapshot(self):
        """The warm-up run before a capture executes the whole step; with the fused exchange that includes the Adam
        update, so the optimizer state is put back afterwards (four small device-to-device copies, once)."""
        if self.exchange != "fused":
            return None
        return [self.dev.scontiguous(x) for x in (self.flat_p, self.flat_m, self.flat_v, self.t_dev)]   # fresh copies

    def _state_restore(self, snap):
        if snap is None:
            return
        for dst, src in zip((self.flat_p, self.flat_m, self.flat_v, self.t_dev), snap):
            self.dev.upload_into(dst, _devptr(src))

    def _capture(self):
        snap = self._state_snapshot()
        try:
            self._capture_inner()
        finally:
            self._state_restore(snap)

    def _capture_inner(self):
        if self.program == "vectorised":
            # record the unchanged program once (nothing is launched), let the library's contraction pass rewrite
            # it, run the result once eagerly (allocations, scratch, NCCL channels), then bake it into a graph
            self.prog = self.dev.record(self._gradient_program)
            self.bucket = self.prog.outs[0]
            self.prog.run()
            self.dev.sync()
            call("hb_graph_begin")
            try:
                self.prog.run()
                if self.optimizer != "adam":
                    self._apply(self.bucket)
            finally:
                g = C.c_void_p()
                call("hb_graph_end", C.byref(g))
            self.graph = g
            return
        # run once eagerly so every size class and the scratch buffer exist
        # (and NCCL has set up its channels before the all-reduce is captured)
        b = self._gradient_program()
        self.dev.sync()
        del b
        call("hb_graph_begin")
        try:
            self.bucket = self._gradient_program()
            if self.optimizer != "adam":
                self._apply(self.bucket)
        finally:
            g = C.c_void_p()
            call("hb_graph_end", C.byref(g))
        self.graph = g

    # ---- public API --------------------------------------------------------
    def upload(self, glyphs_ptr, labels_ptr):
        self.dev.upload_into(self.glyphs, glyphs_ptr)
        self.dev.upload_into(self.labels, labels_ptr)

    def train_step(self, glyphs_ptr, labels_ptr, next_glyphs_ptr=None, next_labels_ptr=None) -> float:
        """The end-to-end user call: H2D of the minibatch, step, loss D2H.

        With `next_*` (the host buffers of the FOLLOWING minibatch, as a data
        loader hands them out) that batch's H2D copy is started on the copy
        stream into staging buffers and overlaps this step; the next call then
        finds its inputs already on the device and pays one D2D copy."""
        key = (_addr(glyphs_ptr), _addr(labels_ptr))
        if self._prefetched == key:
            call("hb_prefetch_ready")
            self.dev.upload_into(self.glyphs, _devptr(self._stage_g))
            self.dev.upload_into(self.labels, _devptr(self._stage_l))
            call("hb_prefetch_consumed")
        else:
            self.upload(glyphs_ptr, labels_ptr)
        self._prefetched = None
        bucket = self.step_enqueue()
        if next_glyphs_ptr is not None:
            if self._stage_g is None:
                self._stage_g = self.dev.sconcrete(np.zeros((self.B, 28, 28), self.dtype))
                self._stage_l = self.dev.sconcrete(np.zeros((self.B, 10), self.dtype))
            call("hb_prefetch_into", self._stage_g.h, next_glyphs_ptr)
            call("hb_prefetch_into", self._stage_l.h, next_labels_ptr)
            self._prefetched = (_addr(next_glyphs_ptr), _addr(next_labels_ptr))
        return float(self.dev.kfromS(self.dev.sslice(self.n_params, 1, bucket)))

    def train_step_device(self, data, k: int) -> float:
        """One step on minibatch k of a `mnist_data.DeviceMnist`: the batch is
        assembled on the device from the resident IDX bytes; only the loss crosses PCIe."""
        data.batch_into(k, self.glyphs, self.labels)
        bucket = self.step_enqueue()
        return float(self.dev.kfromS(self.dev.sslice(self.n_params, 1, bucket)))

    def loss_of(self, bucket) -> float:
        return float(self.dev.kfromS(self.dev.sslice(self.n_params, 1, bucket)))

    def reset(self):
        """Back to the initial parameters and zero Adam state (three small
        device-to-device copies on the library stream)."""
        self.dev.upload_into(self.flat_p, _devptr(self._p0))
        self.dev.upload_into(self.flat_m, _devptr(self._z0))
        self.dev.upload_into(self.flat_v, _devptr(self._z0))
        self.dev.upload_into(self.t_dev, _devptr(self._t0))
        self.t = 0

    def get_params(self):
        return [self.dev.to_numpy(p) for p in self.params]

    def close(self):
        if self.graph is not None:
            call("hb_graph_free", self.graph)
            self.graph = None
        if self.prog is not None:
            self.prog.close()
            self.prog = None
        if self.objective == "global":
            call("hb_comm_set_global_objective", 0)
            self.objective = "shard"


def comm_init(rank: int, world: int, broadcast_bytes) -> None:
    """hb_comm_init with the NCCL unique id created on rank 0 and shipped by
    `broadcast_bytes(bytes_or_None) -> bytes` (the host process' own control
    plane, e.g. torch.distributed over gloo)."""
    if world == 1:
        call("hb_comm_init", 0, 1, None)
        return
    buf = C.create_string_buffer(ffi.HB_UNIQUE_ID_BYTES)
    if rank == 0:
        call("hb_comm_unique_id", buf)
    raw = broadcast_bytes(bytes(buf.raw) if rank == 0 else None)
    idb = C.create_string_buffer(raw, ffi.HB_UNIQUE_ID_BYTES)
    call("hb_comm_init", rank, world, idb)

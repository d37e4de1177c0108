"""The multi-GPU data path ON HARDWARE (needs >= 2 GPUs; skipped otherwise): two ranks, one process per GPU, drive
`train.py`'s real path -- the recorded + rewritten gradient program, the flat bucket, and the fused peer-memory
all-reduce + Adam kernel (hb_allreduce_adam_step) -- and check
  * the exchanged bucket == mean of the two shard buckets (rtol 1e-14) and == the NCCL + scale path bit for bit
    (two summands commute),
  * parameters bit-identical across ranks after 5 steps, and equal to the NCCL-path trainer's,
  * Adam's device-side step counter advanced identically,
  * objective="global" (min u and sum expU of the loss exchanged across the ranks inside the loss kernel): the summed
    bucket of the two shards == the ORACLE's gradient and loss of the unfused reference program on the WHOLE batch
    (1e-12 normwise) -- the N-GPU step is the reference's one-device step on the same inputs."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import ctypes as C
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["LOCAL_RANK"] = str(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from horde_ad_b200 import ffi
    from horde_ad_b200.device import DeviceBackend
    from horde_ad_b200.train import CnnTrainer, PinnedBuffer, comm_init, exchange_scale
    dev = DeviceBackend(rank)

    def bcast(b):
        objs = [b]
        dist.broadcast_object_list(objs, src=0)
        return objs[0]
    comm_init(rank, world, bcast)
    rng = np.random.Generator(np.random.Philox(100 + rank))
    B, hyper = 12, dict(kh=2, kw=2, c_out=4, n_hidden=8)
    gl = PinnedBuffer((B, 28, 28), np.float64)
    lb = PinnedBuffer((B, 10), np.float64)
    gl.array[...] = rng.uniform(0, 1, gl.array.shape)
    lb.array[...] = np.eye(10)[rng.integers(0, 10, B)]
    out = {}
    # (1) one exchange: fused kernel vs NCCL + scale on the same shard buckets
    x = rng.uniform(-1, 1, 5000)
    a, b = dev.sconcrete(x), dev.sconcrete(x)
    z = lambda n, dt=np.float64: dev.sconcrete(np.zeros(n, dt))
    p1, m1, v1, p2, m2, v2 = (z(4000) for _ in range(6))
    t1 = dev.sconcrete(np.zeros((), np.int64))
    for step in range(1, 4):
        ffi.call("hb_allreduce_adam_step", a.h, 4000, p1.h, m1.h, v1.h, t1.h, 1e-3, 0.9, 0.999, 1e-7, exchange_scale(world))
        ffi.call("hb_allreduce_sum", b.h, exchange_scale(world))
        dev.adam_step(p2, m2, v2, dev.sslice(0, 4000, b), step)
    out["exchange_equal"] = bool(np.array_equal(dev.to_numpy(a), dev.to_numpy(b)))
    out["adam_equal"] = bool(np.array_equal(dev.to_numpy(p1), dev.to_numpy(p2)) and np.array_equal(dev.to_numpy(v1), dev.to_numpy(v2)))
    out["t"] = int(dev.to_numpy(t1))
    gathered = [None] * world
    dist.all_gather_object(gathered, x)
    mean = sum(gathered) * exchange_scale(world)
    first = x.copy()
    # after three exchanges of the already-averaged bucket the value is the mean (every rank holds it from step 1 on)
    out["mean_ok"] = bool(np.allclose(dev.to_numpy(a), mean, rtol=1e-14, atol=0))
    # (2) five training steps through the real trainer, both exchanges
    res = {}
    for exch in ("fused", "nccl"):
        tr = CnnTrainer(dev, B, **hyper, world=world, exchange=exch)
        losses = [tr.train_step(gl.ptr, lb.ptr) for _ in range(5)]
        res[exch] = (np.concatenate([p.reshape(-1) for p in tr.get_params()]), losses, dev.to_numpy(tr.bucket))
        tr.close()
    out["fused_vs_nccl_params_equal"] = bool(np.array_equal(res["fused"][0], res["nccl"][0]))
    out["losses"] = res["fused"][1]
    pall = [None] * world
    dist.all_gather_object(pall, res["fused"][0])
    out["params_identical_across_ranks"] = bool(all(np.array_equal(pall[0], p) for p in pall))
    # the exchanged bucket of the LAST step == mean of what each rank computes alone on its shard with the same weights
    from horde_ad_b200 import ad, mnist_cnn
    tr0 = CnnTrainer(dev, B, **hyper, world=world, exchange="fused")
    for _ in range(4):
        tr0.train_step(gl.ptr, lb.ptr)
    tr0.upload(gl.ptr, lb.ptr)
    loss, grads = ad.grad(dev, lambda T, *ps: mnist_cnn.convMnistLossFusedS(T, tr0.glyphs, tr0.labels, list(ps), fused=True),
                          tr0.params)
    local = dev.to_numpy(dev.flatten_concat(list(grads) + [loss]))       # no exchange: this rank's shard bucket
    tr0.train_step(gl.ptr, lb.ptr)
    exchanged = dev.to_numpy(tr0.bucket)
    locs = [None] * world
    dist.all_gather_object(locs, local)
    out["bucket_is_mean_of_shards"] = bool(np.allclose(exchanged, sum(locs) * exchange_scale(world), rtol=1e-13, atol=1e-300))
    tr0.close()
    # (3) global objective: sharded step == the oracle's step on the concatenated batch
    trg = CnnTrainer(dev, B, **hyper, world=world, exchange="nccl", objective="global", use_graph=False)
    trg.upload(gl.ptr, lb.ptr)
    p_init = [p.copy() for p in trg.get_params()]
    gbucket = dev.to_numpy(trg._gradient_program())          # recorded program is not needed here: one eager step
    trg.close()
    glall, lball = [None] * world, [None] * world
    dist.all_gather_object(glall, np.array(gl.array))
    dist.all_gather_object(lball, np.array(lb.array))
    if rank == 0:
        from oracle.concrete import ConcreteBackend
        orc = ConcreteBackend()
        G, L = orc.sconcrete(np.concatenate(glall)), orc.sconcrete(np.concatenate(lball))
        oval, ograds = ad.grad(orc, lambda T, *ps: mnist_cnn.convMnistLossFusedS(T, G, L, list(ps), fused=False),
                               [orc.sconcrete(p) for p in p_init])
        ref = np.concatenate([orc.to_numpy(g).reshape(-1) for g in ograds] + [np.array([float(orc.kfromS(oval))])])
        out["global_objective_rel_err"] = float(np.max(np.abs(gbucket - ref)) / np.max(np.abs(ref)))
        out["global_loss"] = (float(gbucket[-1]), float(ref[-1]))
    gb = [None] * world
    dist.all_gather_object(gb, gbucket)
    out["global_bucket_identical_across_ranks"] = bool(all(np.array_equal(gb[0], b) for b in gb))
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_and_training_on_hardware():
    from horde_ad_b200 import ffi
    if ffi.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, world = _free_port(), 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out["exchange_equal"] and out["adam_equal"] and out["t"] == 3 and out["mean_ok"], out
    assert out["fused_vs_nccl_params_equal"] and out["params_identical_across_ranks"], out
    assert out["bucket_is_mean_of_shards"], out
    assert all(np.isfinite(out["losses"])), out
    assert out["global_objective_rel_err"] <= 1e-12, out
    assert abs(out["global_loss"][0] - out["global_loss"][1]) <= 1e-12 * abs(out["global_loss"][1]), out
    assert out["global_bucket_identical_across_ranks"], out

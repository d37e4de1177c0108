"""MNIST input pipeline: IDX decoding and the device batch
assembly against the oracle's restatement of example/MnistData.hs:127-178 and
against facts computed from the reference's own sample data
(tests/golden/make_mnist_fixture.py)."""
import gzip
import json
import os
import struct

import numpy as np
import pytest

from horde_ad_b200 import mnist_data as MD
from oracle import concrete as OC

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIX = json.load(open(os.path.join(GOLDEN, "mnist_fixture.json")))


def fixture_pair():
    return MD.load_idx_pair(os.path.join(GOLDEN, "t10k-first64-images-idx3-ubyte.gz"),
                            os.path.join(GOLDEN, "t10k-first64-labels-idx1-ubyte.gz"))


# ------------------------------------------------------------------ CPU side
def test_idx_decoder_on_the_reference_sample_data():
    glyphs, labels = fixture_pair()
    s = FIX["slice"]
    assert glyphs.shape == (s["n"], FIX["full"]["height"], FIX["full"]["width"]) and glyphs.dtype == np.uint8
    assert labels.tolist() == s["labels"]
    assert glyphs.reshape(s["n"], -1).sum(axis=1, dtype=np.int64).tolist() == s["pixel_sums"]
    assert glyphs[0, 14].tolist() == s["glyph0_row14"]


def test_idx_decoder_rejects_malformed_files():
    ok = struct.pack(">HBB", 0, 8, 1) + struct.pack(">I", 3) + bytes([1, 2, 3])
    assert MD.decode_idx(ok).tolist() == [1, 2, 3]
    assert MD.decode_idx(gzip.compress(ok)).tolist() == [1, 2, 3]
    for bad in (b"", b"\x00\x01\x08\x01" + ok[4:],           # bad magic
                struct.pack(">HBB", 0, 0x0D, 1) + ok[4:],    # float payload: not an MNIST file
                ok[:-1], ok + b"\x00",                        # payload size vs extents
                ok[:6]):                                      # truncated header
        with pytest.raises(ValueError, match="wrong MNIST file"):
            MD.decode_idx(bad)
    empty = struct.pack(">HBB", 0, 8, 3) + struct.pack(">III", 0, 28, 28)
    assert MD.decode_idx(empty).shape == (0, 28, 28)


def test_oracle_batches_follow_readMnistData():
    glyphs, labels = fixture_pair()
    g, t = OC.mnist_batch(glyphs, labels, None, 8, 4)
    assert g.shape == (4, 28, 28) and t.shape == (4, 10)
    # glyph = fromIntegral pixel / 255: exactly the IEEE quotient, not pixel * (1 / 255)
    assert g[0, 14, 14] == glyphs[8, 14, 14] / 255.0 and g.max() <= 1.0 and g.min() == 0.0
    assert np.array_equal(np.argmax(t, axis=1), labels[8:12]) and np.all(t.sum(axis=1) == 1)
    g32, _ = OC.mnist_batch(glyphs, labels, None, 8, 4, dtype=np.float32)
    assert g32.dtype == np.float32 and np.array_equal(g32, glyphs[8:12].astype(np.float32) / np.float32(255))
    # an order vector selects the samples; ids outside [0, n) give zeros
    order = np.array([5, 63, -1, 64, 0], np.int64)
    g, t = OC.mnist_batch(glyphs, labels, order, 1, 4)
    assert np.array_equal(g[0], glyphs[63] / 255.0) and not g[1].any() and not g[2].any() and not t[1:3].any()
    assert np.array_equal(g[3], glyphs[0] / 255.0)


def test_shuffle_is_a_stable_sort_by_random_keys():
    o = MD.shuffle_order(1000, 7)
    assert o.dtype == np.int32 and sorted(o.tolist()) == list(range(1000))
    assert np.array_equal(o, MD.shuffle_order(1000, 7)) and not np.array_equal(o, MD.shuffle_order(1000, 8))
    assert not np.array_equal(o, np.arange(1000))


# ------------------------------------------------------------------ GPU side
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("order_dtype", [None, np.int32, np.int64])
def test_device_batches_equal_the_oracle_bit_for_bit(dev, dtype, order_dtype):
    glyphs, labels = fixture_pair()
    data = MD.DeviceMnist(dev, glyphs, labels)
    order = None
    if order_dtype is not None:
        order = MD.shuffle_order(64, 3).astype(order_dtype)
        order[5], order[17] = -3, 64                      # out-of-range ids -> zeros (the kernel's own rule)
        with pytest.raises(ValueError, match="sample ids"):
            data.set_order(order)                         # the host-side check refuses them ...
        data.set_order(order, validate=False)             # ... unless it is switched off
    for batch in (1, 7, 16):
        for k in range(data.n_batches(batch)):
            g, t = data.batch(k, batch, dtype)
            go, to = OC.mnist_batch(glyphs, labels, order, k * batch, batch, dtype=dtype)
            assert np.array_equal(dev.to_numpy(g), go) and np.array_equal(dev.to_numpy(t), to)


@pytest.mark.gpu
def test_device_batches_odd_glyph_sizes_and_errors(dev):
    rng = np.random.default_rng(5)
    pix = rng.integers(0, 256, (9, 5, 3), dtype=np.uint8)       # 15 pixels per glyph: scalar path
    lab = rng.integers(0, 4, 9, dtype=np.uint8)
    data = MD.DeviceMnist(dev, pix, lab, classes=4)
    g, t = data.batch(1, 4)
    go, to = OC.mnist_batch(pix, lab, None, 4, 4, classes=4)
    assert np.array_equal(dev.to_numpy(g), go) and np.array_equal(dev.to_numpy(t), to)
    data.set_order(np.arange(9, dtype=np.int32))
    with pytest.raises(RuntimeError, match="beyond the order vector"):
        data.batch(2, 4)                                          # samples 8..11 of 9
    with pytest.raises(RuntimeError, match="glyph extents"):
        dev.mnist_batch_into(data.pixels, data.labels, None, 0, dev.sconcrete(np.zeros((2, 5, 4))),
                             dev.sconcrete(np.zeros((2, 4))))


@pytest.mark.gpu
def test_training_from_resident_data_equals_training_from_host_batches(dev):
    """CnnTrainer.train_step_device (batch assembled on the device) follows the same
    trajectory as train_step on the host-made batches of the same samples."""
    from horde_ad_b200.train import CnnTrainer
    import ctypes as C
    glyphs, labels = fixture_pair()
    order = MD.shuffle_order(64, 11)
    data = MD.DeviceMnist(dev, glyphs, labels)
    data.set_order(order)
    B = 16
    a = CnnTrainer(dev, B, kh=2, kw=2, c_out=4, n_hidden=8, use_graph=True)
    b = CnnTrainer(dev, B, kh=2, kw=2, c_out=4, n_hidden=8, use_graph=False)
    for k in range(data.n_batches(B)):
        la = a.train_step_device(data, k)
        g, t = OC.mnist_batch(glyphs, labels, order, k * B, B)
        g, t = np.ascontiguousarray(g), np.ascontiguousarray(t)
        lb = b.train_step(g.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p))
        assert la == lb
    for pa, pb in zip(a.get_params(), b.get_params()):
        assert np.array_equal(pa, pb)
    a.close()
    b.close()


@pytest.mark.gpu
def test_pipelined_train_step_equals_plain_uploads(dev):
    """CnnTrainer.train_step(cur, next): the next minibatch is copied to the device during the step and moved into the
    step's input arrays right behind it -- the same losses and parameters as uploading every batch in its own call;
    an upload or another batch in between invalidates what was staged."""
    from horde_ad_b200.train import CnnTrainer, PinnedBuffer
    rng = np.random.default_rng(5)
    B, nb = 16, 5
    bufs = []
    for _ in range(nb):
        g, t = PinnedBuffer((B, 28, 28), np.float64), PinnedBuffer((B, 10), np.float64)
        g.array[...] = rng.uniform(0, 1, (B, 28, 28))
        t.array[...] = np.eye(10)[rng.integers(0, 10, B)]
        bufs.append((g, t))
    a = CnnTrainer(dev, B, kh=2, kw=2, c_out=4, n_hidden=8)
    b = CnnTrainer(dev, B, kh=2, kw=2, c_out=4, n_hidden=8)
    seq = [0, 1, 2, 3, 4, 0, 2]
    la = [a.train_step(bufs[i][0].ptr, bufs[i][1].ptr) for i in seq]
    lb = []
    for k, i in enumerate(seq):
        nxt = bufs[seq[k + 1]] if k + 1 < len(seq) else (None, None)
        if k == 4:       # the caller changes its mind: batch 0 was announced, an explicit upload of batch 3 comes first
            b.upload(bufs[3][0].ptr, bufs[3][1].ptr)
        lb.append(b.train_step(bufs[i][0].ptr, bufs[i][1].ptr, *(x.ptr if x is not None else None for x in nxt)))
    assert la == lb
    for pa, pb in zip(a.get_params(), b.get_params()):
        assert np.array_equal(pa, pb)
    a.close()
    b.close()
    for g, t in bufs:
        g.free()
        t.free()

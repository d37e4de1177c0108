"""MNIST input pipeline with the data resident on the device.

Host-side mirror of example/MnistData.hs: `decodeIDX`/`decodeIDXLabels` +
`readMnistData` (:134-152: glyph = fromIntegral pixel / 255, one-hot target),
`shuffle` (:174-178: sort the samples by a stream of random keys), `chunksOf` +
`mkMnistDataBatchS` (:127-133).  The IDX bytes are uploaded ONCE; every
minibatch is then assembled by `hb_mnist_batch_into` from the resident bytes and
the epoch's order vector, so a training step moves no tensor over PCIe (the
8-byte loss comes back).

The reference's `shuffle` draws its keys from `System.Random`'s StdGen, whose
stream cannot be reproduced without that package (unpinned);
here the keys come from numpy's Philox generator and the sort is stable, as
`sortBy (comparing snd)` is.
"""
from __future__ import annotations

import gzip
import struct

import numpy as np

from .device import DeviceArray, DeviceBackend

IDX_UBYTE = 0x08


def decode_idx(raw: bytes) -> np.ndarray:
    """IDX container (big-endian): 0x0000, type byte, rank byte, rank x uint32
    extents, then the payload.  Only the unsigned-byte type MNIST uses."""
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    if len(raw) < 4 or raw[0] != 0 or raw[1] != 0:
        raise ValueError("wrong MNIST file: bad IDX magic")
    if raw[2] != IDX_UBYTE:
        raise ValueError(f"wrong MNIST file: IDX element type 0x{raw[2]:02x}, expected unsigned byte")
    rank = raw[3]
    if len(raw) < 4 + 4 * rank:
        raise ValueError("wrong MNIST file: truncated IDX header")
    dims = struct.unpack(">" + "I" * rank, raw[4:4 + 4 * rank])
    n = int(np.prod(dims)) if rank else 1
    body = raw[4 + 4 * rank:]
    if len(body) != n:
        raise ValueError(f"wrong MNIST file: {len(body)} payload bytes for extents {dims}")
    return np.frombuffer(body, dtype=np.uint8).reshape(dims)


def load_idx_pair(glyphs_path: str, labels_path: str):
    """loadMnistData (MnistData.hs:160-171): (pixels uint8 [n,H,W], labels uint8 [n])."""
    with open(glyphs_path, "rb") as f:
        glyphs = decode_idx(f.read())
    with open(labels_path, "rb") as f:
        labels = decode_idx(f.read())
    if glyphs.ndim != 3 or labels.ndim != 1:
        raise ValueError("wrong MNIST file: expected a rank-3 glyph file and a rank-1 label file")
    if glyphs.shape[0] != labels.shape[0]:
        raise ValueError("can't decode MNIST file into integers: glyph and label counts differ")
    return glyphs, labels


def shuffle_order(n: int, seed: int) -> np.ndarray:
    """`shuffle g l`: positions sorted (stably) by one random key each."""
    keys = np.random.Generator(np.random.Philox(seed)).integers(np.iinfo(np.int64).min, np.iinfo(np.int64).max,
                                                                 size=n, dtype=np.int64, endpoint=True)
    return np.argsort(keys, kind="stable").astype(np.int32)


class DeviceMnist:
    """The decoded data set resident in HBM + the current epoch's order."""

    def __init__(self, dev: DeviceBackend, pixels: np.ndarray, labels: np.ndarray, classes: int = 10):
        if pixels.dtype != np.uint8 or labels.dtype != np.uint8:
            raise ValueError("DeviceMnist: raw IDX bytes (uint8) expected")
        self.dev, self.n, self.classes = dev, int(pixels.shape[0]), classes
        self.sample_shape = tuple(pixels.shape[1:])
        self.pixels = dev.sconcrete(np.ascontiguousarray(pixels).view(np.int8))
        self.labels = dev.sconcrete(np.ascontiguousarray(labels).view(np.int8))
        self.order = None            # file order until shuffled
        self.order_len = self.n

    def shuffle(self, seed: int):
        self.set_order(shuffle_order(self.n, seed))

    def set_order(self, order: np.ndarray, validate: bool = True):
        """Any sample-id vector: an epoch's shuffle, or this rank's shard of it."""
        order = np.ascontiguousarray(order)
        if validate and order.size and (int(order.min()) < 0 or int(order.max()) >= self.n):
            # validated once on the host: a bad id would otherwise yield an all-zero glyph and target on the device
            raise ValueError(f"set_order: sample ids must lie in [0, {self.n})")
        if order.dtype not in (np.int32, np.int64):
            order = order.astype(np.int64)
        self.order = self.dev.sconcrete(order)
        self.order_len = int(order.shape[0])

    def n_batches(self, batch: int) -> int:
        return self.order_len // batch

    def batch_into(self, k: int, glyphs: DeviceArray, targets: DeviceArray):
        """Minibatch k of `chunksOf batch` into resident [batch,H,W] / [batch,classes] arrays."""
        batch = glyphs.shape[0]
        self.dev.mnist_batch_into(self.pixels, self.labels, self.order, k * batch, glyphs, targets)

    def batch(self, k: int, batch: int, dtype=np.float64):
        g = self.dev.sconcrete(np.zeros((batch,) + self.sample_shape, dtype))
        t = self.dev.sconcrete(np.zeros((batch, self.classes), dtype))
        self.batch_into(k, g, t)
        return g, t

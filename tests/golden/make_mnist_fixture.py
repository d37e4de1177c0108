"""Writes the MNIST fixtures of tests/golden/ from the reference's own sample
data (samplesData/t10k-*-ubyte.gz, the files example/MnistData.hs:154-158
names).  Run in the build container from the repository root:

    python tests/golden/make_mnist_fixture.py

* t10k-first64-images-idx3-ubyte.gz / t10k-first64-labels-idx1-ubyte.gz: the
  first 64 samples re-wrapped as IDX files (what the decoder and the device
  batch assembly are tested on; /root/reference does not exist on the GPU box);
* mnist_fixture.json: facts about the FULL files (counts, label histogram,
  pixel sum) and about the slice (labels, per-glyph pixel sums), computed with
  plain struct/zlib parsing that shares no code with horde-ad_b200/mnist_data.py.
"""
import gzip
import json
import os
import struct
import sys

import numpy as np

REF = "/root/reference/samplesData"
OUT = os.path.dirname(os.path.abspath(__file__))
N = 64


def read_idx(path):
    raw = gzip.open(path, "rb").read()
    zero, ty, rank = struct.unpack(">HBB", raw[:4])
    assert zero == 0 and ty == 8
    dims = struct.unpack(">" + "I" * rank, raw[4:4 + 4 * rank])
    return np.frombuffer(raw[4 + 4 * rank:], np.uint8).reshape(dims)


def write_idx(path, a):
    hdr = struct.pack(">HBB", 0, 8, a.ndim) + struct.pack(">" + "I" * a.ndim, *a.shape)
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(hdr + a.tobytes())


if not os.path.isdir(REF):
    sys.exit("the reference's samplesData is not available here")
glyphs = read_idx(f"{REF}/t10k-images-idx3-ubyte.gz")
labels = read_idx(f"{REF}/t10k-labels-idx1-ubyte.gz")
write_idx(f"{OUT}/t10k-first{N}-images-idx3-ubyte.gz", glyphs[:N])
write_idx(f"{OUT}/t10k-first{N}-labels-idx1-ubyte.gz", labels[:N])
json.dump({
    "source": "samplesData/t10k-images-idx3-ubyte.gz, t10k-labels-idx1-ubyte.gz (example/MnistData.hs:154-158)",
    "full": {"n": int(glyphs.shape[0]), "height": int(glyphs.shape[1]), "width": int(glyphs.shape[2]),
             "label_histogram": np.bincount(labels, minlength=10).tolist(), "pixel_sum": int(glyphs.sum(dtype=np.int64))},
    "slice": {"n": N, "labels": labels[:N].tolist(),
              "pixel_sums": glyphs[:N].reshape(N, -1).sum(axis=1, dtype=np.int64).tolist(),
              "glyph0_row14": glyphs[0, 14].tolist()},
}, open(f"{OUT}/mnist_fixture.json", "w"), indent=1)
print("wrote fixtures for", N, "samples of", glyphs.shape[0])

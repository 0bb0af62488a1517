"""MNIST IDX loader for BASELINE.json configs[0] (the reference's `mnist` source, bench/time/Main.hs:113-125).

The reference streams `assets/mnist/train-images-idx3-ubyte` with mnist-idx-conduit's `sourceIdxSparse fp (Just n)`: every
image becomes an SVector of dimension rows*cols (784) holding only its NONZERO pixels, each mapped through
`toUnitRange w8 = fromIntegral w8 / 255` (bench/time/Main.hs:124-125).  The data file itself is absent from the reference
checkout (.MISSING_LARGE_BLOBS); this loader reads any IDX3 unsigned-byte file in that format.
"""
import struct

import numpy as np

from .api import SparseRows


def read_idx_ubyte(path, n=None):
    """IDX file of unsigned bytes -> uint8 array of shape dims (first n items).  Big-endian header:
    0x00 0x00 0x08 ndim, then ndim uint32 sizes."""
    with open(path, "rb") as fh:
        head = fh.read(4)
        if len(head) != 4 or head[0] != 0 or head[1] != 0:
            raise ValueError("%s: not an IDX file" % path)
        if head[2] != 0x08:
            raise ValueError("%s: IDX element type 0x%02x is not unsigned byte" % (path, head[2]))
        ndim = head[3]
        dims = struct.unpack(">%dI" % ndim, fh.read(4 * ndim))
        if ndim < 1:
            raise ValueError("%s: IDX file without dimensions" % path)
        items = dims[0] if n is None else min(int(n), dims[0])
        per = int(np.prod(dims[1:], dtype=np.int64)) if ndim > 1 else 1
        buf = fh.read(items * per)
        if len(buf) != items * per:
            raise ValueError("%s: truncated IDX file" % path)
    return np.frombuffer(buf, np.uint8).reshape((items,) + tuple(dims[1:]))


def mnistSparse(path, n=None):
    """`mnist fp n` (bench/time/Main.hs:113-122): SparseRows of dimension rows*cols with the nonzero pixels / 255."""
    img = read_idx_ubyte(path, n)
    flat = img.reshape(img.shape[0], -1)
    r, c = np.nonzero(flat)
    off = np.zeros(flat.shape[0] + 1, np.int64)
    np.add.at(off, r + 1, 1)
    return SparseRows(np.cumsum(off), c.astype(np.int32), flat[r, c].astype(np.float64) / 255.0, flat.shape[1])

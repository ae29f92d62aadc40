"""TEST INFRASTRUCTURE: reader for the record stream written by oracle/ref_dump.cpp."""
import struct

import numpy as np

_DT = {0: "<f8", 1: "<i8", 2: "<i4", 3: "i1", 4: "<u8"}


def load(path):
    b = open(path, "rb").read()
    o, d = 0, {}
    while o < len(b):
        (nl,) = struct.unpack_from("<I", b, o)
        o += 4
        name = b[o:o + nl].decode()
        o += nl
        ty, nd = struct.unpack_from("<II", b, o)
        o += 8
        shape = struct.unpack_from("<%dQ" % nd, b, o)
        o += 8 * nd
        n = int(np.prod(shape)) if nd else 1
        a = np.frombuffer(b, dtype=_DT[ty], count=n, offset=o).reshape(shape)
        o += a.nbytes
        d[name] = a.copy()
    return d

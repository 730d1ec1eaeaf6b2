"""WFC1 record container (reader/writer) -- Python twin of oracle/wfc_io.h.

file   := b"WFC1" , record*
record := name[32] (NUL padded) , int32 dtype (0 = float64, 1 = int32, 2 = int64) , int64 count , data
"""
import struct
import numpy as np

_DT = {0: np.float64, 1: np.int32, 2: np.int64}
_CODE = {np.dtype(np.float64): 0, np.dtype(np.int32): 1, np.dtype(np.int64): 2}


def write(path, records):
    with open(path, "wb") as fp:
        fp.write(b"WFC1")
        for name, val in records.items():
            a = np.asarray(val)
            if a.dtype not in _CODE:
                a = a.astype(np.float64) if a.dtype.kind == "f" else a.astype(np.int32)
            a = np.ascontiguousarray(a).reshape(-1)
            fp.write(name.encode().ljust(32, b"\0")[:32])
            fp.write(struct.pack("<iq", _CODE[a.dtype], a.size))
            fp.write(a.tobytes())


def read(path):
    out = {}
    with open(path, "rb") as fp:
        if fp.read(4) != b"WFC1":
            raise ValueError("%s: not a WFC1 file" % path)
        while True:
            name = fp.read(32)
            if len(name) < 32:
                break
            dt, cnt = struct.unpack("<iq", fp.read(12))
            dtype = np.dtype(_DT[dt])
            out[name.rstrip(b"\0").decode()] = np.frombuffer(fp.read(cnt * dtype.itemsize), dtype=dtype).copy()
    return out

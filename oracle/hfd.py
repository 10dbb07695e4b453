"""TEST INFRASTRUCTURE. Reader for the .hfd array container written by oracle/ref_dump.cpp and by the host
library's own dump call: records of [i32 name_len][name][i32 dtype 0=f64,1=i32][i32 ndim][i64 dims][data],
data column-major as hf_array stores it (reference include/hf_array.h:303-325)."""
import struct
import numpy as np


def read_hfd(path):
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    off = 0
    n = len(buf)
    while off < n:
        (nl,) = struct.unpack_from("<i", buf, off); off += 4
        name = buf[off:off + nl].decode(); off += nl
        dt, nd = struct.unpack_from("<ii", buf, off); off += 8
        dims = struct.unpack_from("<%dq" % nd, buf, off); off += 8 * nd
        cnt = int(np.prod(dims)) if nd else 1
        dtype = np.float64 if dt == 0 else np.int32
        arr = np.frombuffer(buf, dtype=dtype, count=cnt, offset=off).reshape(dims, order="F")
        off += cnt * (8 if dt == 0 else 4)
        out[name] = arr
    return out


def write_hfd(path, arrays):
    with open(path, "wb") as f:
        for name, a in arrays.items():
            a = np.asarray(a)
            if a.dtype.kind == "f":
                a = a.astype(np.float64); dt = 0
            else:
                a = a.astype(np.int32); dt = 1
            nb = name.encode()
            f.write(struct.pack("<i", len(nb))); f.write(nb)
            f.write(struct.pack("<ii", dt, a.ndim))
            f.write(struct.pack("<%dq" % a.ndim, *a.shape))
            f.write(np.asfortranarray(a).tobytes(order="F"))

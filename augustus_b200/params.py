"""Reader/writer for the AUGB2PAR parameter blob (include/augb200_params.h)."""
import lzma
import struct
import numpy as np

MAGIC = b"AUGB2PAR"
_HDR = struct.Struct("<8sII")
_ENT = struct.Struct("<40sII4QQQ")


def load_bytes(path: str) -> bytes:
    """Read a blob file; ``.xz`` files are decompressed transparently."""
    if path.endswith(".xz"):
        with lzma.open(path, "rb") as f:
            return f.read()
    with open(path, "rb") as f:
        return f.read()


def parse(blob: bytes) -> dict:
    magic, version, n = _HDR.unpack_from(blob, 0)
    if magic != MAGIC:
        raise ValueError("not an AUGB2PAR blob")
    if version != 1:
        raise ValueError("unsupported blob version %d" % version)
    out = {}
    off = _HDR.size
    for _ in range(n):
        name, dtype, ndim, d0, d1, d2, d3, offset, nbytes = _ENT.unpack_from(blob, off)
        off += _ENT.size
        name = name.split(b"\0", 1)[0].decode()
        dims = (d0, d1, d2, d3)[:ndim]
        dt = np.float64 if dtype == 0 else np.int32
        out[name] = np.frombuffer(blob, dtype=dt, count=int(np.prod(dims)), offset=offset).reshape(dims)
    return out

"""CPU tests of the C-ABI library: it loads, exports every symbol include/augb200.h declares, and fails
loudly (no CPU fallback) when no GPU is present."""
import ctypes
import os
import re

import pytest

from tests import util

LIB = os.path.join(util.ROOT, "augustus_b200", "libaugb200.so")


def _declared_symbols():
    hdr = open(os.path.join(util.ROOT, "include", "augb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(augb200_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build the CUDA extension first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(LIB)
    syms = _declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), s


def test_model_create_validates_blob_and_needs_a_device():
    import torch
    from augustus_b200 import AugB200Error, Decoder
    with pytest.raises(AugB200Error) as ei:
        Decoder(b"not a blob at all" * 10)
    assert ei.value.code == 1
    if not torch.cuda.is_available():
        with pytest.raises(AugB200Error) as ei:
            Decoder(util.blob_bytes())
        assert ei.value.code in (3, 4)          # AUGB200_ERR_NO_DEVICE / _CUDA: no silent CPU path


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(util.ROOT, "augustus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cc", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "ghmm_oracle" not in txt and "hostemu" not in txt and "oracle/" not in txt, f


def test_host_side_binding_is_not_under_oracle():
    """the drop-in shim and the blob exporter are product integration code (host/), the oracle directory holds the checker only"""
    assert os.path.exists(os.path.join(util.ROOT, "host", "augshim.cc")) and os.path.exists(os.path.join(util.ROOT, "host", "aug_export.h"))
    assert not os.path.exists(os.path.join(util.ROOT, "oracle", "augshim.cc"))
    for f in ("augshim.cc", "aug_export.h"):
        txt = open(os.path.join(util.ROOT, "host", f)).read()
        assert "ghmm_oracle" not in txt and "hostemu" not in txt


def test_library_path_cannot_be_redirected(monkeypatch):
    """the Python binding loads the in-tree library and nothing else (no environment override)"""
    from augustus_b200 import decoder
    monkeypatch.setenv("AUGB200_LIB", "/tmp/other.so")
    assert decoder.library_path() == os.path.join(util.ROOT, "augustus_b200", "libaugb200.so")


def test_blob_parser_rejects_corrupt_blobs_without_crashing():
    """offsets that wrap, short scalars, wrong dtypes and wild geometry integers must give AUGB200_ERR_BAD_BLOB / _UNSUPPORTED"""
    import struct
    he = util.HostEmu(util.blob_bytes())
    blob = bytearray(util.blob_bytes())
    n_entries = struct.unpack_from("<I", blob, 12)[0]
    ent = 16
    esz = 40 + 4 + 4 + 32 + 8 + 8

    def entry(name):
        for i in range(n_entries):
            o = ent + i * esz
            if bytes(blob[o:o + 40]).split(b"\0")[0] == name:
                return o
        raise KeyError(name)

    def rejected(b):
        err = ctypes.create_string_buffer(512)
        h = he.lib.hostemu_model_create(bytes(b), len(b), err, 512)
        if h:
            he.lib.hostemu_model_destroy(ctypes.c_void_p(h))
        return not h

    he.lib.hostemu_model_destroy.argtypes = [ctypes.c_void_p]
    b = bytearray(blob); o = entry(b"exon_emi"); struct.pack_into("<Q", b, o + 80, 2 ** 64 - 64)      # offset wraps
    assert rejected(b)
    b = bytearray(blob); o = entry(b"statecount"); struct.pack_into("<Q", b, o + 88, 2)                # 2-byte int32
    assert rejected(b)
    b = bytearray(blob); o = entry(b"intron_d"); struct.pack_into("<I", b, o + 40, 0)                  # int declared as f64
    assert rejected(b)
    for name, bad in ((b"ass_start", 2000), (b"dss_end", -3), (b"tis_motif_k", 17), (b"trans_init_window", 1 << 30)):
        b = bytearray(blob); o = entry(name); off = struct.unpack_from("<Q", b, o + 80)[0]; struct.pack_into("<i", b, off, bad)
        assert rejected(b), name
    assert not rejected(blob)

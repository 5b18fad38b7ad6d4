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

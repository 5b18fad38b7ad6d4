"""BASELINE.json configs[2] at full size (pytest -m gpu): every 200 kb window of examples/chr2L under --species=fly defaults
(71 states, softmasking, sample=100) decoded on the GPU in one batch; the Viterbi path and all 99 sampled paths of each window are
compared with digests of the reference's own result (tests/golden/ref_chr2L_digests.json, written by make_golden_chr2L.py from
oracle/_ref/augdump = the unmodified reference objects, one process per window).

The 31 Mbp of input are reference data, copied by oracle/Makefile to oracle/_ref/data/ (travels with the built checker, not
committed); without it the test skips."""
import hashlib
import json
import os

import numpy as np
import pytest

import bench
from augustus_b200 import Decoder
from tests import util

pytestmark = pytest.mark.gpu

DIGESTS = os.path.join(util.GOLDEN, "ref_chr2L_digests.json")


def _rows(raw, i):
    n, status, logp, offset, b, e, t, tr = raw
    o, k = int(offset[i]), int(n[i])
    return np.stack([t[o:o + k].astype("<i4"), b[o:o + k].astype("<i4"), e[o:o + k].astype("<i4"), tr[o:o + k].astype("<i4")], axis=1)


def test_all_chr2L_windows_match_reference_digests():
    wins = bench.chr2l_windows()
    if wins is None or not os.path.exists(DIGESTS):
        pytest.skip("oracle/_ref/data/chr2L.sm.fa.gz or the digests are not present")
    gold = json.load(open(DIGESTS))["windows"]
    idx = sorted(int(k) for k in gold)
    assert len(idx) >= 8
    dec = Decoder(util.blob_bytes("fly_softmask_utr"), 0)
    vit, samp = dec.decode_batch_sampling_raw([wins[i].encode() for i in idx], 100)
    assert not vit[1].any() and not samp[1].any()
    bad = []
    for k, i in enumerate(idx):
        g = gold[str(i)]
        assert g["length"] == len(wins[i])
        rows = _rows(vit, k)
        ok = hashlib.sha1(np.ascontiguousarray(rows).tobytes()).hexdigest() == g["viterbi_sha1"]
        ok &= abs(float(vit[2][k]) - g["log_prob"]) <= 1e-6 * abs(g["log_prob"])          # LLDouble posterior scores within 1e-6 relative
        h = hashlib.sha1()
        for s in range(99):
            r = _rows(samp, k * 99 + s)
            h.update(np.asarray([len(r)], dtype="<i4").tobytes())
            h.update(np.ascontiguousarray(r).tobytes())
        ok &= h.hexdigest() == g["samples_sha1"]
        if not ok:
            bad.append(i)
    assert not bad, "windows that differ from the reference: %s" % bad

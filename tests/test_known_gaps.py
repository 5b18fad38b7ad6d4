"""The documented differences from the reference, PINNED: each test states exactly how the kernel source differs
today, so that a fix and a regression are both noticed.  The oracle (which restates the reference's call-order dependent memos
sequentially) equals the reference in every one of these cases; the CUDA path runs the kernel source checked here.

CPU tests: the host build of the kernel source (tests/hostemu) against the oracle and the reference's own results."""
import gzip
import json

import numpy as np

from tests import util


def test_aedes_three_gc_classes_sampling_steps_at_class_boundaries():
    """--species=aedes (three GC classes) on the soft-masked chr2L window, 99 sampled paths.  Round 1 drew 96 of the reference's 99: a
    sampling step of lessD inside the columns around a GC-class boundary took the plain prefix difference, the reference the value its
    SnippetProbs memo holds since the forward pass (statemodel.cc:312-342).  The forward fill now keeps every memo value that is not the
    plain difference (SnipX) and the sampling steps look it up: 99 of 99, like the oracle."""
    blob = util.blob_bytes("aedes")
    dna = util.read_fasta(util.GOLDEN + "/fly_softmask_window.fa")[0][1]
    ref = [[tuple(x) for x in s["states"]] for s in json.load(gzip.open(util.GOLDEN + "/ref_samples_aedes.json.gz", "rt"))["samples"]]
    assert len(ref) == 99
    o = util.Oracle(blob).sample(dna, 100)
    assert [s["states"] for s in o["samples"]] == ref                                   # the oracle = the reference
    for simt32 in (False, True):
        e = util.HostEmu(blob, simt32=simt32).sample(dna, 99)
        assert e["status"] == 0
        assert [k for k, (a, b) in enumerate(zip(e["samples"], ref)) if a["states"] != b] == []
    v = util.HostEmu(blob).decode(dna, want_cells=True)
    r = util.Oracle(blob).viterbi(dna, want_matrix=True)
    assert v["states"] == r["condensed"] and (np.where(r["V"] <= util.NEGT, -(1 << 61), r["V"]) == v["cells"]).all()     # every Viterbi cell agrees


def test_human_utr_two_gc_classes_cells_behind_a_class_boundary():
    """human --UTR=on on real DNA with two GC classes: IntronModel::aSSProb memoises per position and forgets after 1000 positions
    (intronmodel.cc:1119-1136); with UTR states asking for acceptor sites far behind the current column a memoised value can carry the
    motif of the other class.  The kernels score a site with the class of its own column: a handful of longass / exon cells right
    behind a boundary differ from the oracle (= reference), the paths and scores do not."""
    blob = util.blob_bytes("human_utr")
    emu, orc = util.HostEmu(blob), util.Oracle(blob)
    want = {"chr2L_5005000_30000": [(25109, 13), (25109, 18), (25109, 23), (25188, 8)],
            "chr2L_12010000_30000": [(2928, 13), (2928, 18), (2928, 23), (2931, 8), (2954, 5), (2963, 10), (3503, 11)],
            "chr2L_21470000_20000": []}
    for name, dna in util.read_fasta(util.GOLDEN + "/real_windows.fa"):
        r, e = orc.viterbi(dna, want_matrix=True), emu.decode(dna, want_cells=True)
        assert e["status"] == 0 and e["states"] == r["condensed"] and e["log_prob"] == r["log_prob"]
        V, E = r["V"], e["cells"]
        zv, ze = V <= util.NEGT, E <= util.NEGT
        diff = (zv != ze) | (~zv & ~ze & (V != E))
        assert [tuple(int(x) for x in ij) for ij in np.argwhere(diff)] == want[name], name


def test_human_parameters_on_multi_class_real_dna_sampled_paths():
    """human parameters on 100 kb of chr2L (13 / 38 GC-class boundaries), 99 sampled paths: kernel source == oracle (round 1: one path of
    the first window differed, which moved two posterior probabilities of the GFF by 0.01).  Needs the chr2L data of the built checker."""
    import os
    import pytest
    path = os.path.join(util.ROOT, "oracle", "_ref", "data", "chr2L.sm.fa.gz")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/data/chr2L.sm.fa.gz not present")
    seq = "".join(l.strip() for l in gzip.open(path, "rt") if not l.startswith(">"))
    blob = util.blob_bytes("human")
    emu, orc = util.HostEmu(blob), util.Oracle(blob)
    for off in (5000000, 12000000):
        dna = seq[off:off + 100000].upper()
        o, e = orc.sample(dna, 100), emu.sample(dna, 99)
        assert e["status"] == 0 and [a["states"] for a in e["samples"]] == [b["states"] for b in o["samples"]]

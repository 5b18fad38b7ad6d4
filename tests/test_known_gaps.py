"""The documented differences from the reference, PINNED: each test states exactly how the kernel source differs
today, so that a fix and a regression are both noticed.  The oracle (which restates the reference's call-order dependent memos
sequentially) equals the reference in every one of these cases; the CUDA path runs the kernel source checked here.

CPU tests: the host build of the kernel source (tests/hostemu) against the oracle and the reference's own results."""
import gzip
import json

import numpy as np

from tests import util


def test_aedes_three_gc_classes_sampling_steps_at_class_boundaries():
    """--species=aedes (three GC classes) on the soft-masked chr2L window, 99 sampled paths: the oracle draws the reference's 99, the
    kernel source 96 — a sampling step of lessD inside the columns around a GC-class boundary takes the plain prefix difference, the
    reference the SnippetProbs memo as the forward pass, the backtracking and the earlier walks left it (statemodel.cc:312-342)."""
    blob = util.blob_bytes("aedes")
    dna = util.read_fasta(util.GOLDEN + "/fly_softmask_window.fa")[0][1]
    ref = [[tuple(x) for x in s["states"]] for s in json.load(gzip.open(util.GOLDEN + "/ref_samples_aedes.json.gz", "rt"))["samples"]]
    assert len(ref) == 99
    o = util.Oracle(blob).sample(dna, 100)
    assert [s["states"] for s in o["samples"]] == ref                                   # the oracle = the reference
    e = util.HostEmu(blob).sample(dna, 99)
    assert e["status"] == 0
    bad = [k for k, (a, b) in enumerate(zip(e["samples"], ref)) if a["states"] != b]
    assert bad == [7, 35, 41], "the pinned gap changed (a fix shortens this list; anything else is a regression): %s" % bad
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

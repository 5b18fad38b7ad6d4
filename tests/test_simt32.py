"""The DEVICE flavour of the kernel source on the CPU: tests/hostemu/simt32.h runs ghmm_sweep.h / ghmm_sample.h with 32 lanes
(32 fibers per warp, every ballot / shuffle / warp barrier a rendezvous), so the lane-group paths the one-lane host build never
takes — three reading frames of an exon kind or of lessD in one pass with 8 lanes each, fixed-length states with
lane = (frame, ancestor), per-group arg-best — are checked cell for cell against the oracle without a GPU.  The executor also aborts
when lanes meet in different collectives or read a flag another lane is about to clear (it found one such race in apply_pending)."""
import numpy as np
import pytest

from augustus_b200 import synth
from tests import util


def _same_cells(orc, emu, dna):
    r, e = orc.viterbi(dna, want_matrix=True), emu.decode(dna, want_cells=True)
    assert e["status"] == 0 and e["states"] == r["condensed"] and e["log_prob"] == r["log_prob"]
    V, E = r["V"], e["cells"]
    assert ((V <= util.NEGT) == (E <= util.NEGT)).all() and (V[V > util.NEGT] == E[V > util.NEGT]).all()


def test_47_state_human_two_gc_classes_and_synthetic():
    blob = util.blob_bytes()
    orc, emu = util.Oracle(blob), util.HostEmu(blob, simt32=True)
    for _, dna in util.read_fasta(util.GOLDEN + "/example.fa"):           # 9453 + 2344 bases, both GC classes: memo-emulated columns too
        _same_cells(orc, emu, dna)
    for i, n in enumerate([12000, 3001, 257, 40, 3]):
        _same_cells(orc, emu, synth.window(900 + i, n))
    real = util.read_fasta(util.GOLDEN + "/real_windows.fa")[0][1]
    _same_cells(orc, emu, real[:15000])


def test_fly_47_states_real_dna():
    blob = util.blob_bytes("fly_noutr")
    orc, emu = util.Oracle(blob), util.HostEmu(blob, simt32=True)
    dna = util.read_fasta(util.GOLDEN + "/fly_window.fa")[0][1]
    _same_cells(orc, emu, dna[:20000])


def test_71_state_utr_models():
    for name, fa, n in (("human_utr", "example.fa", 9453), ("fly_softmask_utr", "fly_softmask_window.fa", 14000)):
        blob = util.blob_bytes(name)
        orc, emu = util.Oracle(blob), util.HostEmu(blob, simt32=True)
        dna = util.read_fasta(util.GOLDEN + "/" + fa)[0][1][:n]
        _same_cells(orc, emu, dna)


def test_forward_and_sampling_kernels_32_lanes():
    blob = util.blob_bytes()
    orc, emu = util.Oracle(blob), util.HostEmu(blob, simt32=True)
    dna = util.read_fasta(util.GOLDEN + "/example.fa")[1][1]
    o, e = orc.sample(dna, 30)["samples"], emu.sample(dna, 29)
    assert e["status"] == 0 and all(a["states"] == b["states"] for a, b in zip(e["samples"], o))

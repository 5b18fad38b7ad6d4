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
    """forward fill with the frames of a kind in one pass (per-group log-sum-exp) + the sampling walks: sampled paths == oracle"""
    for name, fa, sl, ns in (("human", "example.fa", slice(0, 2344), 30), ("human", "example.fa", slice(0, 6000), 12),
                             ("fly_softmask_utr", "fly_softmask_window.fa", slice(2000, 11000), 20)):
        blob = util.blob_bytes(name)
        orc, emu = util.Oracle(blob), util.HostEmu(blob, simt32=True)
        seqs = util.read_fasta(util.GOLDEN + "/" + fa)
        dna = (seqs[1][1] if sl.stop == 2344 else seqs[0][1])[sl]
        o, e = orc.sample(dna, ns)["samples"], emu.sample(dna, ns - 1)
        assert e["status"] == 0 and len(e["samples"]) == ns - 1 and all(a["states"] == b["states"] for a, b in zip(e["samples"], o))


def test_random_windows_with_class_shifts_n_runs_and_masks():
    """Small fixed-seed slice of tools/fuzz_simt32.py (composition shifts = GC-class boundaries, N runs, soft-masked runs, tiny
    windows): 32-lane device flavour == one-lane build on every cell, == oracle on every cell for the models without class-history
    memos (47 states; fly has one GC class)."""
    import random
    for name, cases in (("human", 10), ("fly_softmask_utr", 6)):
        blob = util.blob_bytes(name)
        orc, e1, e32 = util.Oracle(blob), util.HostEmu(blob), util.HostEmu(blob, simt32=True)
        rng = random.Random(31)
        for _ in range(cases):
            L = rng.choice([2, 7, 150, 1200, 5000, 8000])
            parts = []
            while sum(map(len, parts)) < L:
                gc = rng.choice([0.3, 0.41, 0.62, 0.7]); at = (1 - gc) / 2
                parts.append("".join(rng.choices("ACGT", weights=(at, gc / 2, gc / 2, at), k=rng.randint(1, max(1, L // 2)))))
            s = list("".join(parts)[:L])
            for _ in range(rng.randint(0, 2)):
                a = rng.randrange(L)
                for i in range(a, min(L, a + rng.randint(1, 30))):
                    s[i] = "N"
            if "softmask" in name:
                for _ in range(rng.randint(0, 4)):
                    a = rng.randrange(L)
                    for i in range(a, min(L, a + rng.randint(1, 400))):
                        s[i] = s[i].lower()
            dna = "".join(s)
            a, b = e1.decode(dna, want_cells=True), e32.decode(dna, want_cells=True)
            assert a["status"] == b["status"] == 0 and a["states"] == b["states"] and a["log_prob"] == b["log_prob"] and (a["cells"] == b["cells"]).all()
            r = orc.viterbi(dna, want_matrix=True)
            assert b["states"] == r["condensed"] and b["log_prob"] == r["log_prob"]
            if set(dna.upper()) - {"N"}:
                V, E = r["V"], b["cells"]
                assert ((V <= util.NEGT) == (E <= util.NEGT)).all() and (V[V > util.NEGT] == E[V > util.NEGT]).all()


def test_forward_matrix_32_lanes_matches_oracle():
    """ln forward of every cell from the grouped forward fill against the oracle's forward matrix (the reference's LLDouble values are
    pinned to the oracle's in test_oracle): same zero pattern, values within 1e-9 (they differ by summation order only)."""
    for name, fa, n in (("human", "example.fa", 5000), ("human_utr", "example.fa", 4000)):
        blob = util.blob_bytes(name)
        orc, emu = util.Oracle(blob), util.HostEmu(blob, simt32=True)
        dna = util.read_fasta(util.GOLDEN + "/" + fa)[0][1][:n]
        Fo = orc.sample(dna, 2, want_forward=True)["F"]
        Fo = np.where(Fo < -1e300, -np.inf, Fo)
        r = emu.forward(dna)
        assert r["status"] == 0
        F = r["F"]; fin = np.isfinite(F)
        assert (np.isfinite(Fo) == fin).all()
        assert np.abs(F[fin] - Fo[fin]).max() <= 1e-9


def test_reverse_lane_order(monkeypatch):
    """the executor runs the lanes of a phase in ascending order by default (lane 0 first); with SIMT32_REVERSE lane 0 runs last, which
    exposes a read that needs a barrier after lane 0's write: same cells, same sampled paths"""
    monkeypatch.setenv("SIMT32_REVERSE", "1")
    for name in ("human", "human_nc"):
        blob = util.blob_bytes(name)
        orc, emu = util.Oracle(blob), util.HostEmu(blob, simt32=True)
        dna = util.read_fasta(util.GOLDEN + "/example.fa")[1][1]
        _same_cells(orc, emu, dna)
        o, e = orc.sample(dna, 16)["samples"], emu.sample(dna, 15)
        assert e["status"] == 0 and all(a["states"] == b["states"] for a, b in zip(e["samples"], o))

"""NcModel states (--nc=on: 83 states).

Without hints the six nc states that begin or end with a transcript boundary are dead (precomputeTxEndProbs leaves their tss / tts
probabilities at zero, ncmodel.cc:744-826), but ncintron / rncintron (alive in every column) and ncinternal / rncinternal live on the
initial probabilities of column 0, share the aSSProb memo with the intron model and compete for the path at the last column.  The C
restatement (oracle/ghmm_oracle.c: nc_eval) is pinned here against the reference's own paths, scores, per-state cell counts, last
matrix column and sampled paths (tests/golden/make_golden_nc.py); the kernel source (two more lazily evaluated chains, two table-driven
exon states, the dead states as column-0 cells) is compared with the oracle cell for cell on the CPU, the GPU test runs the same
windows through the C ABI."""
import gzip
import json
import os

import numpy as np
import pytest

from augustus_b200 import synth
from tests import util


@pytest.fixture(scope="module")
def blob():
    return util.blob_bytes("human_nc")


@pytest.fixture(scope="module")
def oracle(blob):
    return util.Oracle(blob)


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(util.GOLDEN, "ref_paths_nc.json")))


def _check(oracle, dna, ref):
    r = oracle.viterbi(dna, want_matrix=True)
    assert r["condensed"] == [tuple(s) for s in ref["states"]]
    assert abs(r["log_prob"] - ref["log_prob"]) <= 1e-9 * abs(ref["log_prob"])
    V = r["V"]
    assert [int((V[:, s] > util.NEGT).sum()) for s in range(71, 83)] == ref["nc_cells"]          # same zero pattern per nc state
    last = V[-1].astype(np.float64) / 2.0 ** 40
    for s, want in enumerate(ref["last_column"]):
        if want is None:
            assert V[-1, s] <= util.NEGT
        else:
            assert abs(last[s] - want) <= 1e-7


def test_blob_has_83_states(blob, oracle):
    from augustus_b200 import params
    p = params.parse(blob)
    assert int(p["nc_option_on"][0]) == 1 and int(p["statecount"][0]) == 83 and oracle.S == 83


def test_oracle_matches_reference_on_example_fa(oracle, golden):
    for (name, dna), ref in zip(util.read_fasta(util.GOLDEN + "/example.fa"), golden["example"]):
        _check(oracle, dna, ref)
        assert ref["nc_cells"][2] == len(dna) and ref["nc_cells"][4] > 0          # ncintron everywhere, ncinternal somewhere


def test_oracle_matches_reference_on_real_and_synthetic_windows(oracle, golden):
    for (name, dna), ref in zip(util.read_fasta(util.GOLDEN + "/real_windows.fa"), golden["real"]):
        _check(oracle, dna, ref)
    for i, ref in enumerate(golden["synthetic50k"][:1]):
        _check(oracle, synth.window(i, 50000), ref)
    for spec, ref in zip(golden["synthetic_short_spec"], golden["synthetic_short"]):
        _check(oracle, synth.window(*spec), ref)


def test_oracle_sampling_matches_reference(oracle):
    (name, rec), = json.load(gzip.open(os.path.join(util.GOLDEN, "ref_samples_nc.json.gz"), "rt")).items()
    dna = dict(util.read_fasta(util.GOLDEN + "/example.fa"))[name]
    r = oracle.sample(dna, 100)
    assert len(rec["samples"]) == 99
    for mine, theirs in zip(r["samples"], rec["samples"]):
        assert mine["states"] == [tuple(s) for s in theirs["states"]]


def _same_cells(oracle, emu, dna):
    r, e = oracle.viterbi(dna, want_matrix=True), emu.decode(dna, want_cells=True)
    assert e["status"] == 0 and e["states"] == r["condensed"] and e["log_prob"] == r["log_prob"]
    V, E = r["V"], e["cells"]
    assert ((V <= util.NEGT) == (E <= util.NEGT)).all() and (V[V > util.NEGT] == E[V > util.NEGT]).all()


def test_kernel_source_matches_oracle_cells_one_lane_and_32_lanes(blob, oracle):
    """the sweep with the two nc chains and the two table-driven nc exon states (ghmm_sweep.h, ghmm_model.cc), host build with one lane
    and device flavour on the 32-lane executor: every cell of the 83-state matrix, paths, scores"""
    seqs = util.read_fasta(util.GOLDEN + "/example.fa")
    real = util.read_fasta(util.GOLDEN + "/real_windows.fa")[2][1]
    for simt in (False, True):
        emu = util.HostEmu(blob, simt32=simt)
        _same_cells(oracle, emu, seqs[1][1])
        _same_cells(oracle, emu, seqs[0][1][:6000] if simt else seqs[0][1])
        _same_cells(oracle, emu, synth.window(7, 4000))
        if not simt:
            _same_cells(oracle, emu, real)
            _same_cells(oracle, emu, synth.window(0, 50000))


def test_kernel_source_sampling_matches_oracle(blob, oracle):
    dna = util.read_fasta(util.GOLDEN + "/example.fa")[1][1]
    o = oracle.sample(dna, 40)["samples"]
    for simt in (False, True):
        e = util.HostEmu(blob, simt32=simt).sample(dna, 39)
        assert e["status"] == 0 and all(a["states"] == b["states"] for a, b in zip(e["samples"], o))


@pytest.mark.gpu
def test_gpu_nc_model_matches_reference_and_oracle(blob, oracle, golden):
    from augustus_b200 import Decoder
    dec = Decoder(blob, 0)
    seqs = util.read_fasta(util.GOLDEN + "/example.fa")
    wins = [s for _, s in seqs] + [synth.window(0, 50000)]
    refs = golden["example"] + golden["synthetic50k"][:1]
    for dna, p, ref in zip(wins, dec.decode_batch(wins), refs):
        assert p.status == 0 and p.as_tuples() == [tuple(s) for s in ref["states"]]
        assert abs(p.log_prob - ref["log_prob"]) <= 1e-6 * abs(ref["log_prob"])
        o = oracle.viterbi(dna)
        assert p.as_tuples() == o["condensed"] and p.log_prob == o["log_prob"]
    (name, rec), = json.load(gzip.open(os.path.join(util.GOLDEN, "ref_samples_nc.json.gz"), "rt")).items()
    dna = dict(seqs)[name]
    vit, samples = dec.decode_batch_sampling([dna], 100)
    assert [s.as_tuples() for s in samples[0]] == [[tuple(t) for t in sm["states"]] for sm in rec["samples"]]

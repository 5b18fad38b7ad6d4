"""The 71-state model (--UTR=on; BASELINE.json configs[3]): UtrModel states next to the coding ones.

CPU part: the oracle against the reference's OWN golden vector for this configuration (aug_utr_on.gff of its test suite),
against state paths / scores / sampled paths dumped from the unmodified reference, and the host build of the kernel source
against the oracle cell for cell.  GPU part (pytest -m gpu): the CUDA path through the C ABI against all of these."""
import numpy as np
import pytest

from augustus_b200 import params, synth
from tests import util


@pytest.fixture(scope="module")
def blob():
    return util.blob_bytes("human_utr")


@pytest.fixture(scope="module")
def oracle(blob):
    return util.Oracle(blob)


@pytest.fixture(scope="module")
def emu(blob):
    return util.HostEmu(blob)


@pytest.fixture(scope="module")
def golden():
    return util.golden_paths(utr=True)


def _inputs(golden):
    """(reference record, dna) of every fixture window"""
    out = []
    for key, fa in (("example", "example.fa"), ("real", "real_windows.fa")):
        for (name, dna), ref in zip(util.read_fasta(util.GOLDEN + "/" + fa), golden[key]):
            assert name == ref["name"]
            out.append((ref, dna))
    out += [(ref, synth.window(i, 50000)) for i, ref in enumerate(golden["synthetic50k"])]
    out += [(ref, synth.window(*spec)) for spec, ref in zip(golden["synthetic_short_spec"], golden["synthetic_short"])]
    return out


def _check(r, ref):
    assert r["condensed"] == [tuple(s) for s in ref["states"]]
    assert abs(r["log_prob"] - ref["log_prob"]) <= 1e-9 * abs(ref["log_prob"])


def test_blob(blob):
    p = params.parse(blob)
    assert int(p["statecount"][0]) == 71 and int(p["utr_option_on"][0]) == 1
    assert p["utr5_emi"].shape == (2, 1024) and p["lendist_utr3single"].shape[0] == int(p["utr_max3singlelength"][0]) + 1


def test_oracle_reproduces_the_reference_golden_gff(oracle, blob):
    """tests/short/examples/expected_results/test_utr_on/aug_utr_on.gff: CDS parts, transcription start and end sites."""
    par = params.parse(blob)
    for name, dna in util.read_fasta(util.GOLDEN + "/example.fa"):
        want = util.gff_features(util.GOLDEN + "/aug_utr_on.gff", name)
        assert len(want["CDS"]) >= 5
        assert util.path_features(oracle.viterbi(dna)["condensed"], par) == want


def test_oracle_matches_reference_paths(oracle, golden):
    for ref, dna in _inputs(golden):
        r = oracle.viterbi(dna)
        assert (r["gc"] == util.gc_from_runs(ref["gc"], len(dna))).all()
        _check(r, ref)


def test_oracle_matches_reference_on_a_200k_window(oracle, golden):
    _check(oracle.viterbi(synth.window(0, 200000)), golden["synthetic200k"][0])


def test_oracle_sampling_matches_reference(oracle):
    (name, rec), = util.golden_samples_utr().items()
    dna = dict(util.read_fasta(util.GOLDEN + "/example.fa"))[name]
    r = oracle.sample(dna, 100)
    assert r["viterbi"] == [tuple(s) for s in rec["viterbi"]["states"]]
    assert len(r["samples"]) == len(rec["samples"]) == 99
    for mine, theirs in zip(r["samples"], rec["samples"]):
        assert mine["states"] == [tuple(s) for s in theirs["states"]]
        assert abs(mine["log_prob"] - theirs["log_prob"]) <= 1e-6 * max(1.0, abs(theirs["log_prob"]))


def _cells_equal(oracle, emu, dna):
    r, e = oracle.viterbi(dna, want_matrix=True), emu.decode(dna, want_cells=True)
    assert e["status"] == 0 and e["states"] == r["condensed"] and e["log_prob"] == r["log_prob"]
    V, E = r["V"], e["cells"]
    if not set(dna.upper()) & set("ACGT"):
        V, E = V[1:], E[1:]             # a window without nucleotides keeps only the intergenic state (namgene.cc:205-226); column 0 is never read
    assert ((V <= util.NEGT) == (E <= util.NEGT)).all()
    assert (V[V > util.NEGT] == E[V > util.NEGT]).all()


def test_kernel_source_on_host_matches_oracle_cells(oracle, emu, golden):
    for name, dna in util.read_fasta(util.GOLDEN + "/example.fa"):
        _cells_equal(oracle, emu, dna)
    _cells_equal(oracle, emu, synth.window(1, 50000))
    _cells_equal(oracle, emu, util.read_fasta(util.GOLDEN + "/real_windows.fa")[2][1])         # one GC class, a run of N
    for spec in golden["synthetic_short_spec"]:
        _cells_equal(oracle, emu, synth.window(*spec))
    for dna in ("N" * 500, "ACGT" * 300, synth.window(7, 2000)[:61], "A" * 3000 + synth.window(8, 3000)):
        _cells_equal(oracle, emu, dna)


def test_kernel_source_on_host_across_gc_class_boundaries(oracle, emu):
    """Two GC classes in one window.  IntronModel::aSSProb memoises per position (intronmodel.cc:1119-1136); with UTR states
    asking for splice sites far behind the current column a kept value can carry the class of an earlier column.  The oracle
    restates the memo, the kernels use the class of the column: paths agree, a handful of cells right behind a boundary differ."""
    for name, dna in util.read_fasta(util.GOLDEN + "/real_windows.fa")[:2]:
        r, e = oracle.viterbi(dna, want_matrix=True), emu.decode(dna, want_cells=True)
        assert e["states"] == r["condensed"]
        V, E = r["V"], e["cells"]
        assert ((V <= util.NEGT) == (E <= util.NEGT)).all()
        assert int((V != E).sum()) <= 64


def test_sampler_source_on_host_matches_oracle(oracle, emu):
    dna = util.read_fasta(util.GOLDEN + "/example.fa")[0][1]
    o, e = oracle.sample(dna, 100)["samples"], emu.sample(dna, 99)
    assert e["status"] == 0
    for a, b in zip(e["samples"], o):
        assert a["states"] == b["states"] and abs(a["log_prob"] - b["log_prob"]) <= 1e-9 * abs(b["log_prob"])


# ----------------------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def dec(blob):
    from augustus_b200 import Decoder
    d = Decoder(blob, 0)
    assert d.statecount == 71
    return d


@pytest.mark.gpu
def test_gpu_matches_reference_paths_and_golden_gff(dec, oracle, golden, blob):
    inputs = _inputs(golden)
    paths = dec.decode_batch([dna for _, dna in inputs])
    for (ref, dna), p in zip(inputs, paths):
        assert p.status == 0
        assert p.as_tuples() == [tuple(s) for s in ref["states"]]
        assert abs(p.log_prob - ref["log_prob"]) <= 1e-6 * abs(ref["log_prob"])
    par = params.parse(blob)
    for (name, dna), p in zip(util.read_fasta(util.GOLDEN + "/example.fa"), paths):
        assert util.path_features(p.as_tuples(), par) == util.gff_features(util.GOLDEN + "/aug_utr_on.gff", name)


@pytest.mark.gpu
def test_gpu_matches_oracle_bit_for_bit(dec, oracle):
    wins = [synth.window(300 + i, n) for i, n in enumerate((20000, 9000, 33333, 1200, 64, 300))] + ["N" * 500, "ACGT" * 300]
    for p, dna in zip(dec.decode_batch(wins), wins):
        o = oracle.viterbi(dna)
        assert p.status == 0 and p.as_tuples() == o["condensed"] and p.log_prob == o["log_prob"]


@pytest.mark.gpu
def test_gpu_200k_window_matches_reference(dec, golden):
    p, = dec.decode_batch([synth.window(0, 200000)])
    ref = golden["synthetic200k"][0]
    assert p.status == 0 and p.as_tuples() == [tuple(s) for s in ref["states"]]
    assert abs(p.log_prob - ref["log_prob"]) <= 1e-6 * abs(ref["log_prob"])


@pytest.mark.gpu
def test_gpu_sampling_matches_reference(dec):
    (name, rec), = util.golden_samples_utr().items()
    dna = dict(util.read_fasta(util.GOLDEN + "/example.fa"))[name]
    vit, samples = dec.decode_batch_sampling([dna], 100)
    assert vit[0].as_tuples() == [tuple(s) for s in rec["viterbi"]["states"]]
    assert len(samples[0]) == 99
    for mine, theirs in zip(samples[0], rec["samples"]):
        assert mine.as_tuples() == [tuple(s) for s in theirs["states"]]
        assert abs(mine.log_prob - theirs["log_prob"]) <= 1e-6 * max(1.0, abs(theirs["log_prob"]))


def test_acceptor_pattern_running_past_the_window_end():
    """a UTR exon may begin with an acceptor site whose downstream pattern ends one base past the window (the exon is the last base;
    IntronModel::aSSProb then uses the invalid-pattern probability): found by tools/fuzz_simt32.py, 24 cells of utr5term were missing"""
    dna = "TTCGGCTTCCGCCCACTAATAAATAANNNNNNNNNNNNNNNNNNNNNNGNNNNGGCTAGC"
    blob = util.blob_bytes("human_utr")
    orc = util.Oracle(blob)
    r = orc.viterbi(dna, want_matrix=True)
    assert (r["V"][36:, 29] > util.NEGT).all()                     # the reference holds utr5term cells in the last 24 columns (augdump)
    for simt in (False, True):
        e = util.HostEmu(blob, simt32=simt).decode(dna, want_cells=True)
        V, E = r["V"], e["cells"]
        assert e["status"] == 0 and e["states"] == r["condensed"] and e["log_prob"] == r["log_prob"]
        assert ((V <= util.NEGT) == (E <= util.NEGT)).all() and (V[V > util.NEGT] == E[V > util.NEGT]).all()


def test_acceptor_pattern_past_the_window_end_with_softmasking_fly():
    """the same edge with ass_end = 4 (pattern up to two bases past the end) and the softmasking bonus of the site's intron bases"""
    dna = "TTTCTACacggatcnnnnnnnnnnnnnnnnnnnnnnnnnnnnnnnnnnnnnntaaaaaagca"
    blob = util.blob_bytes("fly_softmask_utr")
    orc = util.Oracle(blob)
    for seq in (dna, dna + "t", dna.upper()):
        r = orc.viterbi(seq, want_matrix=True)
        for simt in (False, True):
            e = util.HostEmu(blob, simt32=simt).decode(seq, want_cells=True)
            V, E = r["V"], e["cells"]
            assert e["status"] == 0 and e["states"] == r["condensed"] and e["log_prob"] == r["log_prob"]
            assert ((V <= util.NEGT) == (E <= util.NEGT)).all() and (V[V > util.NEGT] == E[V > util.NEGT]).all()

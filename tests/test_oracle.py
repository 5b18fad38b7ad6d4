"""CPU tests: the oracle (oracle/ghmm_oracle.c) against the reference's own outputs (tests/golden/),
and the host build of the kernel source against the oracle.  No GPU needed."""
import numpy as np
import pytest

from augustus_b200 import params, synth
from tests import util


@pytest.fixture(scope="module")
def blob():
    return util.blob_bytes()


@pytest.fixture(scope="module")
def oracle(blob):
    return util.Oracle(blob)


@pytest.fixture(scope="module")
def emu(blob):
    return util.HostEmu(blob)


@pytest.fixture(scope="module")
def golden():
    return util.golden_paths()


def test_blob_parses(blob):
    p = params.parse(blob)
    assert int(p["statecount"][0]) == 47 and int(p["num_gc_classes"][0]) == 2
    assert p["trans"].shape == (2, 47, 47) and p["exon_emi"].shape == (2, 3, 1024)
    assert np.isneginf(p["lendist_intron"][0])


def test_synth_is_reproducible():
    a, b = synth.window(3, 1000), synth.window(3, 1000)
    assert a == b and set(a) <= set("ACGT") and synth.window(4, 1000) != a
    gc = (a.count("G") + a.count("C")) / 1000
    assert 0.33 < gc < 0.49


def _check_against_ref(oracle, dna, ref):
    r = oracle.viterbi(dna)
    assert r["n"] > 0
    # GC-class stairs as ContentStairs::computeStairs (reference motif.cc:543-614)
    assert (r["gc"] == util.gc_from_runs(ref["gc"], len(dna))).all()
    # state path: bit-exact
    assert r["condensed"] == [tuple(s) for s in ref["states"]]
    # LLDouble score vs Q40 log score: 1e-6 relative is the contract, we are ~1e-13
    assert abs(r["log_prob"] - ref["log_prob"]) <= 1e-9 * abs(ref["log_prob"])


def test_oracle_matches_reference_on_example_fa(oracle, golden):
    seqs = util.read_fasta(util.GOLDEN + "/example.fa")
    for (name, dna), ref in zip(seqs, golden["example"]):
        assert name == ref["name"]
        _check_against_ref(oracle, dna, ref)


def test_oracle_matches_reference_on_multiclass_real_dna(oracle, golden):
    """Windows with several GC classes and an N run: exercises the restated SnippetProbs memo."""
    seqs = util.read_fasta(util.GOLDEN + "/real_windows.fa")
    assert any(len(r["gc"]) > 1 for r in golden["real"])
    for (name, dna), ref in zip(seqs, golden["real"]):
        _check_against_ref(oracle, dna, ref)


@pytest.mark.parametrize("i", [0, 1, 2, 3])
def test_oracle_matches_reference_on_synthetic_50k(oracle, golden, i):
    _check_against_ref(oracle, synth.window(i, 50000), golden["synthetic50k"][i])


def test_oracle_matches_reference_on_short_windows(oracle, golden):
    for (idx, n), ref in zip(golden["synthetic_short_spec"], golden["synthetic_short"]):
        _check_against_ref(oracle, synth.window(idx, n), ref)


def test_oracle_forward_and_sampling_match_reference(oracle):
    """Config 5 path (--sample=100 --alternatives-from-sampling=true): the 99 sampled state paths per window are identical
    to the reference's (same glibc rand() stream, one process per window) and forward cells agree to 1e-9."""
    gold = util.golden_samples()
    wins = {"example_HS08198": util.read_fasta(util.GOLDEN + "/example.fa")[1][1],
            "synthetic_301_20000": synth.window(301, 20000),
            "real_chr2L_5005000": util.read_fasta(util.GOLDEN + "/real_windows.fa")[0][1]}
    for name, dna in wins.items():
        ref = gold[name]
        r = oracle.sample(dna, 100, want_forward=True)
        assert len(r["samples"]) == len(ref["samples"]) == 99
        for mine, theirs in zip(r["samples"], ref["samples"]):
            assert mine["states"] == [tuple(x) for x in theirs["states"]]
            assert abs(mine["log_prob"] - theirs["log_prob"]) <= 1e-6 * max(1.0, abs(theirs["log_prob"]))
        for j, s, v in ref["forward_cells"]:
            assert abs(r["F"][j, s] - v) <= 1e-9 * abs(v) + 1e-8
        assert int(np.isfinite(r["F"][[c[0] for c in ref["forward_cells"]]]).sum()) >= len(ref["forward_cells"])


def test_second_parameter_set_fly(golden):
    """--species=fly --UTR=off: other signal geometry (ass_start 1, ass_end 4, trans_init_window 25, no exon-terminal content,
    d = 929).  Oracle vs the reference's path and samples; host build of the kernel source vs the oracle, cell for cell."""
    blob = util.blob_bytes("fly_noutr")
    orc, emu = util.Oracle(blob), util.HostEmu(blob)
    dna = util.read_fasta(util.GOLDEN + "/fly_window.fa")[0][1]
    ref = golden["fly"][0]
    r = orc.viterbi(dna, want_matrix=True)
    assert r["condensed"] == [tuple(s) for s in ref["states"]]
    assert abs(r["log_prob"] - ref["log_prob"]) <= 1e-9 * abs(ref["log_prob"])
    e = emu.decode(dna, want_cells=True)
    assert e["states"] == r["condensed"] and e["log_prob"] == r["log_prob"]
    V, E = r["V"], e["cells"]
    assert ((V <= util.NEGT) == (E <= util.NEGT)).all() and (V[V > util.NEGT] == E[V > util.NEGT]).all()
    gs = util.golden_samples()["fly_chr2L_5000000"]["samples"]
    os_ = orc.sample(dna, 100)["samples"]
    es = emu.sample(dna, 99)["samples"]
    for a, b, c in zip(os_, gs, es):
        assert a["states"] == [tuple(x) for x in b["states"]] == c["states"]
        assert abs(a["log_prob"] - b["log_prob"]) <= 1e-6 * max(1.0, abs(b["log_prob"]))


# ---------------------------------------------------------------- host build of the kernel source
def _cells_equal(oracle, emu, dna):
    o = oracle.viterbi(dna, want_matrix=True)
    e = emu.decode(dna, want_cells=True)
    assert e["status"] == 0
    assert (o["gc"] == e["gc"]).all()
    assert e["states"] == o["condensed"]
    assert e["log_prob"] == o["log_prob"]
    V, E = o["V"], e["cells"]
    assert ((V <= util.NEGT) == (E <= util.NEGT)).all()
    ok = V > util.NEGT
    assert (V[ok] == E[ok]).all()            # bit-exact fixed-point scores in every DP cell


def test_kernel_source_on_host_matches_oracle_cells(oracle, emu):
    for name, dna in util.read_fasta(util.GOLDEN + "/example.fa"):
        _cells_equal(oracle, emu, dna)
    _cells_equal(oracle, emu, synth.window(5, 20000))
    _cells_equal(oracle, emu, synth.window(103, 120))


def test_kernel_source_on_host_matches_oracle_cells_across_gc_class_boundaries(oracle, emu):
    """Real DNA with several GC classes: the windowed restatement of the SnippetProbs memo in the sweep must give
    the same cells as the oracle's whole-window restatement (which is pinned to the reference)."""
    for name, dna in util.read_fasta(util.GOLDEN + "/real_windows.fa"):
        _cells_equal(oracle, emu, dna)


def test_sampler_source_on_host_matches_oracle(oracle, emu):
    """ghmm_sample.h + the forward mode of ghmm_sweep.h compiled for the host: every one of the 99 sampled paths (and
    its ln pathemiProb) equals the oracle's, which is pinned to the reference's; also checks the glibc rand() restatement."""
    import ctypes
    libc = ctypes.CDLL("libc.so.6"); libc.srand(1)
    ref = np.array([libc.rand() for _ in range(20000)], dtype=np.uint32)
    mine = np.zeros(20000, dtype=np.uint32)
    emu.lib.hostemu_rand_stream.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int]
    emu.lib.hostemu_rand_stream(1, mine.ctypes.data, 20000)
    assert (ref == mine).all()
    wins = [util.read_fasta(util.GOLDEN + "/example.fa")[1][1], synth.window(300, 12000), util.read_fasta(util.GOLDEN + "/real_windows.fa")[0][1],
            "N" * 300, synth.window(9, 400)]
    for dna in wins:
        o = oracle.sample(dna, 100)
        e = emu.sample(dna, 99)
        assert e["status"] == 0 and len(e["samples"]) == 99
        for a, b in zip(e["samples"], o["samples"]):
            assert a["states"] == b["states"]
            assert abs(a["log_prob"] - b["log_prob"]) <= 1e-9 * max(1.0, abs(b["log_prob"]))


def test_kernel_source_on_host_edge_cases(oracle, emu):
    base = synth.window(7, 4000)
    cases = [base[:2], base[:3], base[:9], base[:41], base[:600], "N" * 500, base[:1000] + "N" * 300 + base[1000:2000],
             base.lower(), "ACGT" * 300, "GT" * 700, "AG" * 700, "ATG" * 400 + "TAA" * 3]
    for dna in cases:
        o = oracle.viterbi(dna)
        e = emu.decode(dna)
        assert e["status"] == 0 and o["n"] >= 0, len(dna)
        assert e["states"] == o["condensed"], len(dna)
        assert e["log_prob"] == o["log_prob"]


def test_heated_sampling_matches_the_reference():
    """--temperature=3: the forward summands are transEmiProb^(5/8) (LLDouble::heated, lldouble.cc:209-264; exonmodel.cc:1095 and the other
    state models).  The kernel source raises them in the log domain; all 99 sampled paths of both sequences of example.fa equal the
    reference's (one process: the second sequence continues the rand() stream), one lane and 32 lanes."""
    import gzip
    import json
    ref = json.load(gzip.open(util.GOLDEN + "/ref_samples_heated.json.gz", "rt"))["sequences"]
    blob = util.blob_bytes("human_t3")                  # the human parameters exported with --temperature=3
    dnas = [s for _, s in util.read_fasta(util.GOLDEN + "/example.fa")]
    for simt32 in (False, True):
        emu = util.HostEmu(blob, simt32=simt32)
        pos = 0
        for dna, r in zip(dnas, ref):
            got, used = util.hostemu_samples_at(emu, dna, 99, pos)
            assert got == [[tuple(x) for x in s] for s in r["samples"]]
            pos += used
    cold = util.HostEmu(util.blob_bytes()).sample(dnas[1], 99)["samples"]
    assert [s["states"] for s in cold] != [[tuple(x) for x in s] for s in ref[1]["samples"]]          # the temperature does change the paths


def test_carried_rand_generator_matches_libc_at_any_position():
    """the library generates the rand() stream window by window from a carried generator state (ring + 64-bit position) instead
    of from position 0.  Whatever the history of seeks (forwards, backwards = restart), the window at a position equals what an unseeded
    libc process draws there."""
    import ctypes
    import ctypes.util
    emu = util.HostEmu(util.blob_bytes())
    libc = ctypes.CDLL(ctypes.util.find_library("c"))
    libc.srand(1)
    N = 300000
    ref = np.array([libc.rand() for _ in range(N)], dtype=np.uint32)
    emu.lib.hostemu_rand_window.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int]
    for seeks, pos in (([], 0), ([5, 100000, 7], 123456), ([299000], 10), ([1, 2, 3, 250000, 4], 299000)):
        sk = np.array(seeks, dtype=np.uint64)
        out = np.zeros(1000, dtype=np.uint32)
        emu.lib.hostemu_rand_window(sk.ctypes.data if len(seeks) else None, len(seeks), pos, out.ctypes.data, 1000)
        assert (out == ref[pos:pos + 1000]).all(), (seeks, pos)

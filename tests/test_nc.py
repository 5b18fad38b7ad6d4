"""NcModel states (--nc=on: 83 states) — ORACLE ONLY this round.

Without hints the six nc states that begin or end with a transcript boundary are dead (precomputeTxEndProbs leaves their tss / tts
probabilities at zero, ncmodel.cc:744-826), but ncintron / rncintron (alive in every column) and ncinternal / rncinternal live on the
initial probabilities of column 0, share the aSSProb memo with the intron model and compete for the path at the last column.  The C
restatement (oracle/ghmm_oracle.c: nc_eval) is pinned here against the reference's own paths, scores, per-state cell counts, last
matrix column and sampled paths (tests/golden/make_golden_nc.py); the CUDA path rejects such models (AUGB200_ERR_UNSUPPORTED) until
the kernels learn the four live states."""
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


def test_product_rejects_nc_models_loudly(blob):
    emu_err = None
    try:
        util.HostEmu(blob)
    except Exception as ex:          # the model builder of the product (ghmm_model.cc) refuses: no approximation, no fallback
        emu_err = str(ex)
    assert emu_err is not None

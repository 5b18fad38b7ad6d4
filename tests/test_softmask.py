"""Softmasking (the reference's default mode): lower-case runs of the input are nonexonpart hints whose bonus multiplies every
intergenic / intronic position they cover (extrinsicinfo.cc:1696-1724, igenicmodel.cc:306-316, intronmodel.cc:1011-1036,
utrmodel.cc:1143-1158,1521-1545).  Fixture: --species=fly --UTR=on --softmasking=1 (the defaults BASELINE.json configs[2] names) on a
chr2L window that is one third lower case."""
import json
import gzip
import os

import pytest

from augustus_b200 import params
from tests import util


@pytest.fixture(scope="module")
def blob():
    return util.blob_bytes("fly_softmask_utr")


@pytest.fixture(scope="module")
def oracle(blob):
    return util.Oracle(blob)


@pytest.fixture(scope="module")
def emu(blob):
    return util.HostEmu(blob)


@pytest.fixture(scope="module")
def window():
    (name, dna), = util.read_fasta(util.GOLDEN + "/fly_softmask_window.fa")
    assert 0.25 < sum(c.islower() for c in dna) / len(dna) < 0.4
    return name, dna


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(util.GOLDEN, "ref_paths_softmask.json")))


def _samples():
    (name, rec), = json.load(gzip.open(os.path.join(util.GOLDEN, "ref_samples_softmask.json.gz"), "rt")).items()
    return rec


def test_blob(blob):
    p = params.parse(blob)
    assert int(p["softmasking"][0]) == 1 and int(p["extrinsic_malus_all_one"][0]) == 1 and int(p["statecount"][0]) == 71
    assert abs(float(p["softmask_bonus"][0]) - 0.13976194) < 1e-6          # ln 1.15 (config/extrinsic/extrinsic.cfg, source RM)


def test_oracle_matches_reference(oracle, window, golden):
    name, dna = window
    for key, seq in (("masked", dna), ("unmasked", dna.upper())):
        r, ref = oracle.viterbi(seq), golden[key]
        assert r["condensed"] == [tuple(s) for s in ref["states"]]
        assert abs(r["log_prob"] - ref["log_prob"]) <= 1e-9 * abs(ref["log_prob"])
    assert golden["masked"]["states"] != golden["unmasked"]["states"] or golden["masked"]["log_prob"] != golden["unmasked"]["log_prob"]


def test_oracle_sampling_matches_reference(oracle, window):
    rec = _samples()
    r = oracle.sample(window[1], 100)
    for mine, theirs in zip(r["samples"], rec["samples"]):
        assert mine["states"] == [tuple(s) for s in theirs["states"]]


def test_kernel_source_on_host_matches_oracle_cells(oracle, emu, window):
    for dna in (window[1], window[1][5000:17000], window[1].swapcase()[:9000]):
        r, e = oracle.viterbi(dna, want_matrix=True), emu.decode(dna, want_cells=True)
        assert e["status"] == 0 and e["states"] == r["condensed"] and e["log_prob"] == r["log_prob"]
        V, E = r["V"], e["cells"]
        assert ((V <= util.NEGT) == (E <= util.NEGT)).all() and (V[V > util.NEGT] == E[V > util.NEGT]).all()
    o, e = oracle.sample(window[1][:15000], 30)["samples"], emu.sample(window[1][:15000], 29)
    assert e["status"] == 0 and all(a["states"] == b["states"] for a, b in zip(e["samples"], o))


@pytest.mark.gpu
def test_gpu_softmasking_matches_reference_and_oracle(blob, oracle, window, golden):
    from augustus_b200 import Decoder
    dec = Decoder(blob, 0)
    name, dna = window
    seqs = [dna, dna.upper(), dna[5000:17000], dna.swapcase()[:9000]]
    vit, samples = dec.decode_batch_sampling(seqs, 100)
    for p, ref in zip(vit[:2], (golden["masked"], golden["unmasked"])):
        assert p.status == 0 and p.as_tuples() == [tuple(s) for s in ref["states"]]
        assert abs(p.log_prob - ref["log_prob"]) <= 1e-6 * abs(ref["log_prob"])
    for p, seq in zip(vit, seqs):
        o = oracle.viterbi(seq)
        assert p.as_tuples() == o["condensed"] and p.log_prob == o["log_prob"]
    for mine, theirs in zip(samples[0], _samples()["samples"]):
        assert mine.as_tuples() == [tuple(s) for s in theirs["states"]]
    dec.close()

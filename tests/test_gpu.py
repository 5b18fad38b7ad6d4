"""GPU parity tests (pytest -m gpu): the CUDA path through the C ABI against the oracle on the same
inputs, against the committed reference paths, and through size-independent properties at full size."""
import ctypes

import numpy as np
import pytest

from augustus_b200 import AugB200Error, Decoder, synth
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dec():
    return Decoder(util.blob_bytes(), 0)


@pytest.fixture(scope="module")
def oracle():
    return util.Oracle(util.blob_bytes())


@pytest.fixture(scope="module")
def golden():
    return util.golden_paths()


def _same(path, o):
    assert path.status == 0
    assert path.as_tuples() == o["condensed"]
    assert path.log_prob == o["log_prob"]          # identical Q40 integers on both sides


def test_example_fa_matches_reference_and_oracle(dec, oracle, golden):
    seqs = util.read_fasta(util.GOLDEN + "/example.fa")
    paths = dec.decode_batch([s for _, s in seqs])
    for (name, dna), p, ref in zip(seqs, paths, golden["example"]):
        assert p.as_tuples() == [tuple(s) for s in ref["states"]]          # the reference's own path
        assert abs(p.log_prob - ref["log_prob"]) <= 1e-6 * abs(ref["log_prob"])
        _same(p, oracle.viterbi(dna))


def test_synthetic_50k_windows_match_reference(dec, golden):
    wins = synth.windows(16, 50000)
    paths = dec.decode_batch(wins)
    for p, ref in zip(paths, golden["synthetic50k"]):
        assert p.as_tuples() == [tuple(s) for s in ref["states"]]
        assert abs(p.log_prob - ref["log_prob"]) <= 1e-6 * abs(ref["log_prob"])


def test_config2_64_windows_match_reference_digests(dec):
    """the first 64 windows of BASELINE.json configs[1] against the reference (tests/golden/ref_config2_digests.json,
    written by make_golden_config2.py from the unmodified reference; bench.py checks its timed outputs against the same file)."""
    import json
    import bench
    gold = json.load(open(util.GOLDEN + "/ref_config2_digests.json"))["windows"]
    paths = dec.decode_batch(synth.windows(64, 50000))
    for i, p in enumerate(paths):
        g = gold[str(i)]
        assert bench.path_digest(p.as_tuples()) == g["sha1"] and len(p.states) == g["n"]
        assert abs(p.log_prob - g["log_prob"]) <= 1e-6 * abs(g["log_prob"])


def test_real_dna_with_gc_class_boundaries_matches_reference(dec, oracle, golden):
    seqs = util.read_fasta(util.GOLDEN + "/real_windows.fa")
    paths = dec.decode_batch([s for _, s in seqs])
    for (name, dna), p, ref in zip(seqs, paths, golden["real"]):
        assert p.as_tuples() == [tuple(s) for s in ref["states"]]
        assert abs(p.log_prob - ref["log_prob"]) <= 1e-6 * abs(ref["log_prob"])
        _same(p, oracle.viterbi(dna))


def test_seeded_windows_match_oracle(dec, oracle):
    wins = [synth.window(200 + i, n) for i, n in enumerate([50000, 31000, 12345, 7000, 2048, 999, 120, 41, 9, 3, 2])]
    paths = dec.decode_batch(wins)
    for dna, p in zip(wins, paths):
        _same(p, oracle.viterbi(dna))


def test_edge_cases_match_oracle(dec, oracle):
    base = synth.window(7, 4000)
    cases = ["N" * 500, base[:1000] + "N" * 300 + base[1000:2000], base.lower(), "ACGT" * 300, "GT" * 700, "AG" * 700,
             "ATG" * 400 + "TAA" * 3, base[:600]]
    paths = dec.decode_batch(cases)
    for dna, p in zip(cases, paths):
        _same(p, oracle.viterbi(dna))


def test_host_supplied_gc_classes_equal_device_computed(dec, oracle):
    dna = synth.window(11, 20000)
    o = oracle.viterbi(dna)
    a = dec.decode_batch([dna])[0]
    b = dec.decode_batch([dna], gc=[o["gc"]])[0]
    assert a.as_tuples() == b.as_tuples() == o["condensed"]


def test_single_window_seam(dec, oracle):
    dna = util.read_fasta(util.GOLDEN + "/example.fa")[1][1]
    dec.viterbiAndForward(dna)
    _same(dec.getViterbiPath(), oracle.viterbi(dna))


def test_posterior_sampling_matches_reference_and_oracle(dec, oracle):
    """Config 5 path: forward fill + 99 sampled paths per window through the C ABI."""
    gold = util.golden_samples()
    wins = {"example_HS08198": util.read_fasta(util.GOLDEN + "/example.fa")[1][1],
            "synthetic_301_20000": synth.window(301, 20000),
            "real_chr2L_5005000": util.read_fasta(util.GOLDEN + "/real_windows.fa")[0][1]}
    names = list(wins)
    vit, samples = dec.decode_batch_sampling([wins[k] for k in names], 100)
    for k, v, ss in zip(names, vit, samples):
        o = oracle.viterbi(wins[k])
        assert v.as_tuples() == o["condensed"]
        ref = gold[k]["samples"]
        assert len(ss) == len(ref) == 99
        for mine, theirs in zip(ss, ref):
            assert mine.status == 0
            assert mine.as_tuples() == [tuple(x) for x in theirs["states"]]                 # the reference's own sampled path
            assert abs(mine.log_prob - theirs["log_prob"]) <= 1e-6 * max(1.0, abs(theirs["log_prob"]))
    # more windows against the oracle (ragged lengths)
    more = [synth.window(700 + i, n) for i, n in enumerate([30000, 7777, 2000, 300])] + ["N" * 200]
    vit, samples = dec.decode_batch_sampling(more, 100)
    for dna, ss in zip(more, samples):
        o = oracle.sample(dna, 100)
        for mine, theirs in zip(ss, o["samples"]):
            assert mine.as_tuples() == theirs["states"]
            assert abs(mine.log_prob - theirs["log_prob"]) <= 1e-6 * max(1.0, abs(theirs["log_prob"]))


def test_second_parameter_set_fly(golden):
    """--species=fly --UTR=off parameters (different splice-site geometry) through the C ABI: Viterbi path and the 99 samples
    equal the reference's."""
    dec2 = Decoder(util.blob_bytes("fly_noutr"), 0)
    dna = util.read_fasta(util.GOLDEN + "/fly_window.fa")[0][1]
    vit, samples = dec2.decode_batch_sampling([dna, dna[:25000]], 100)
    ref = golden["fly"][0]
    assert vit[0].as_tuples() == [tuple(s) for s in ref["states"]]
    assert abs(vit[0].log_prob - ref["log_prob"]) <= 1e-6 * abs(ref["log_prob"])
    for mine, theirs in zip(samples[0], util.golden_samples()["fly_chr2L_5000000"]["samples"]):
        assert mine.as_tuples() == [tuple(x) for x in theirs["states"]]
    orc = util.Oracle(util.blob_bytes("fly_noutr"))
    assert vit[1].as_tuples() == orc.viterbi(dna[:25000])["condensed"]
    dec2.close()


def test_full_size_properties(dec):
    """Size-independent checks on a larger batch: batch order independence, idempotence, path structure."""
    wins = synth.windows(96, 50000, start=1000)
    p1 = dec.decode_batch(wins)
    p2 = dec.decode_batch(wins[::-1])[::-1]
    for a, b, w in zip(p1, p2, wins):
        assert a.as_tuples() == b.as_tuples() and a.log_prob == b.log_prob
        st = a.states
        assert st[0].begin <= 1 and st[-1].end == len(w) - 1     # a left-truncated first exon may start before the window
        for x, y in zip(st, st[1:]):
            assert y.begin == x.end + 1                 # the path tiles columns 1..L-1
        assert a.log_prob < 0


def test_staged_batch_larger_than_the_arena_runs_in_waves(dec, monkeypatch):
    """augb200_stage_batch keeps DNA + descriptors of all windows on the device and hands the workspace arena from wave to wave
    when the batch does not fit (bench.py stages 10 000 windows): same paths as the plain call, run twice to check re-use."""
    wins = [synth.window(700 + i, n) for i, n in enumerate([20000, 18000, 9000, 21000, 15000, 8000, 22000, 5000, 17000, 19000, 4000, 20000])]
    want = dec.decode_batch(wins)
    monkeypatch.setenv("AUGB200_ARENA_MB", "24")          # a 20 kb window needs ~7 MB: 2-3 windows per wave
    small = Decoder(util.blob_bytes(), 0)
    try:
        small.stage(wins)
        for _ in range(2):
            small.run_staged()
        got = small.fetch_staged()
        assert small.last_launch_count >= 4 * 4             # at least four waves of (prep, sweep, backtrace, pack)
        for a, b in zip(got, want):
            assert a.status == 0 and a.as_tuples() == b.as_tuples() and a.log_prob == b.log_prob
        assert small.last_sweep_ms > 0
        # the plain call on the same small arena goes through sub-batches
        for a, b in zip(small.decode_batch(wins), want):
            assert a.as_tuples() == b.as_tuples()
    finally:
        small.close()


def test_multi_device_call_returns_input_order(dec):
    """augb200_decode_batch_multi: windows dealt block-cyclically to several models (here two models on device 0, which take turns;
    on a node one per GPU), results in input order — Viterbi paths and sampled paths identical to the single-model calls."""
    wins = [synth.window(1300 + i, n) for i, n in enumerate([9000, 4000, 12000, 3000, 7000, 50, 8000])]
    want = dec.decode_batch(wins)
    wv, ws = dec.decode_batch_sampling(wins, 12)
    other = Decoder(util.blob_bytes(), 0)
    try:
        got = Decoder.decode_batch_multi([dec, other], wins)
        for a, b in zip(got, want):
            assert a.status == 0 and a.as_tuples() == b.as_tuples() and a.log_prob == b.log_prob
        gv, gs = Decoder.decode_batch_multi([dec, other], wins, nsample=12)
        for a, b in zip(gv, wv):
            assert a.as_tuples() == b.as_tuples()
        for sa, sb in zip(gs, ws):
            assert [x.as_tuples() for x in sa] == [x.as_tuples() for x in sb]
        with pytest.raises(AugB200Error):
            Decoder.decode_batch_multi([dec, dec], wins)           # the same model twice
    finally:
        other.close()


def test_plain_call_invalidates_a_staged_batch(dec):
    """stage(100) -> decode_batch(10) -> run_staged used to launch over stale descriptors; the staged batch is gone now"""
    d2 = Decoder(util.blob_bytes(), 0)
    try:
        d2.stage([synth.window(40 + i, 3000) for i in range(12)])
        d2.decode_batch([synth.window(7, 2000)])
        with pytest.raises(AugB200Error):
            d2.run_staged()
    finally:
        d2.close()


def test_rand_position_is_a_64_bit_offset_and_may_move_backwards(dec):
    """the rand() stream is generated from a carried generator state, window by window; positions beyond 2^31 are legal,
    a smaller position restarts the generator.  Sampled paths at a position == the host build of the kernel source at that position."""
    dna = util.read_fasta(util.GOLDEN + "/example.fa")[1][1]
    emu = util.HostEmu(util.blob_bytes())
    emu.lib.hostemu_sample_at.restype = ctypes.c_int
    emu.lib.hostemu_sample_at.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 7 + [ctypes.c_uint64, ctypes.c_void_p]

    def emu_at(pos, ns=20):
        L = len(dna); cap = ns * (L // 8 + 64)
        sb, se = np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int32)
        st, tr = np.zeros(cap, dtype=np.uint8), np.zeros(cap, dtype=np.uint8)
        cnt, lp, status, used = np.zeros(ns, dtype=np.int32), np.zeros(ns), ctypes.c_int32(), ctypes.c_int32()
        emu.lib.hostemu_sample_at(emu.m, dna.encode(), L, None, ns, cap, sb.ctypes.data, se.ctypes.data, st.ctypes.data, tr.ctypes.data,
                                  cnt.ctypes.data, lp.ctypes.data, ctypes.byref(status), pos, ctypes.byref(used))
        out, p = [], 0
        for k in range(ns):
            c = int(cnt[k]); out.append([(int(st[p + q]), int(sb[p + q]), int(se[p + q]), int(tr[p + q])) for q in range(c)]); p += c
        return out, used.value
    try:
        for pos in (5_000_000, 1234, 5_000_000 + 77):          # forward, backwards (restart), forward inside the buffered window
            dec.set_rand_position(pos)
            _, samples = dec.decode_batch_sampling([dna], 21)
            want, used = emu_at(pos)
            assert [s.as_tuples() for s in samples[0]] == want
            assert dec.last_rand_consumed == used
        dec.set_rand_position(2 ** 33 + 5)                        # accepted (nothing is generated until a decode asks for it)
    finally:
        dec.set_rand_position(0)


def test_sampled_paths_are_deduplicated_on_the_device(dec):
    """SURVEY.md §8f next-1, first step: a sampled path that repeats an earlier sample of its window shares that sample's arrays and is
    not copied to the host again; first[i, k] names the first occurrence.  The paths themselves are what they were (the reference
    comparisons above run through the same call)."""
    wins = [synth.window(820 + i, n) for i, n in enumerate([30000, 9000, 2500, 400])]
    ns = 100
    vit, samples = dec.decode_batch_sampling(wins, ns)
    first, total = dec.sample_first_occurrence(len(wins), ns)
    held = 0
    for i, ss in enumerate(samples):
        tup = [s.as_tuples() for s in ss]
        seen = {}
        for k, t in enumerate(tup):
            key = tuple(t)
            assert first[i, k] == seen.setdefault(key, k)            # the earliest sample with the same states
        held += sum(len(t) for k, t in enumerate(tup) if first[i, k] == k)
    assert total == sum(len(s.states) for ss in samples for s in ss)
    raw_v, raw_s = dec.decode_batch_sampling_raw(wins, ns)
    assert len(raw_s[4]) == held and held < total                   # the store (and the device->host copy) holds the unique paths only


def test_heated_sampling_matches_the_reference_on_the_gpu():
    """--temperature=3 through the C ABI: the 99 sampled paths of both sequences of example.fa == the reference's (one process: the second
    sequence starts where the first left the rand() stream), fixture tests/golden/ref_samples_heated.json.gz."""
    import gzip
    import json
    ref = json.load(gzip.open(util.GOLDEN + "/ref_samples_heated.json.gz", "rt"))["sequences"]
    d3 = Decoder(util.blob_bytes("human_t3"), 0)
    try:
        pos = 0
        for (_, dna), r in zip(util.read_fasta(util.GOLDEN + "/example.fa"), ref):
            d3.set_rand_position(pos)
            _, samples = d3.decode_batch_sampling([dna], 100)
            assert [s.as_tuples() for s in samples[0]] == [[tuple(x) for x in s] for s in r["samples"]]
            pos += d3.last_rand_consumed
    finally:
        d3.close()

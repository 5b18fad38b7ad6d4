"""Chunk planning and joining of chunk predictions (augustus_b200/chromosome.py) against outputs of the reference's own scripts
createAugustusJoblist.pl and join_aug_pred.pl (tests/golden/make_golden_join.py): synthetic runs covering the overlap rules, and the
first 650 kb of chr2L predicted by the unmodified reference in four overlapping chunks."""
import gzip
import io
import json
import os

import pytest

from augustus_b200 import chromosome as ch
from tests import util


def _load(name):
    with gzip.open(os.path.join(util.GOLDEN, name), "rt") as f:
        return json.load(f)


def test_chunk_borders_and_file_names_equal_the_joblist_script():
    plans = _load("join_cases.json.gz")["plans"]
    assert len(plans) >= 10
    for p in plans:
        got = ch.plan_chunks(p["start"], p["end"], p["chunksize"], p["overlap"] or None, p["padding"])
        names = ch.chunk_names(1, "/data/x/chrT.fa", got, "out")
        assert [[a, b, n] for (a, b), n in zip(got, names)] == p["rows"], p
    # BASELINE.json configs[2]: chr2L in 200 kb windows stepping 150 kb = the joblist for chunksize 200000, overlap 50000
    chunks = ch.plan_chunks(1, 23513712, 200000, 50000)
    assert len(chunks) == 157 and chunks[1] == (150001, 350000) and chunks[-1] == (23400001, 23513712)


def test_chunk_planner_rejects_what_the_script_cannot_run():
    with pytest.raises(ValueError):
        ch.plan_chunks(1, 1000, 300)                 # no overlap below 3 Mbp chunks: the script dies
    with pytest.raises(ValueError):
        ch.plan_chunks(1, 1000, 300, 300)            # overlap >= chunksize: the script loops forever
    with pytest.raises(ValueError):
        ch.plan_chunks(1, 1000, 0, 10)


def test_join_equals_the_perl_script_on_synthetic_runs():
    cases = _load("join_cases.json.gz")["cases"]
    removed = dropped = multi = 0
    for c in cases:
        err = io.StringIO()
        got = ch.join_predictions(c["input"], c["droplist"], err)
        assert got == c["joined"]
        n_in, n_out = c["input"].count("# start gene"), got.count("# start gene")
        removed += n_out < n_in and not c["droplist"]
        dropped += "dropping" in err.getvalue()
        multi += c["input"].count(ch.RUN_SEPARATOR) > 2
    assert removed >= 30 and dropped >= 5 and multi >= 40          # the cases do exercise overlaps, the drop list and chains of runs


def test_join_renumbers_genes_and_transcripts_in_output_order():
    cases = _load("join_cases.json.gz")["cases"]
    for c in cases[:40]:
        ids = [l.split()[3] for l in c["joined"].splitlines() if l.startswith("# start gene ")]
        assert ids == ["g%d" % (i + 1) for i in range(len(ids))]


def test_join_of_reference_chunk_outputs_equals_the_reference_pipeline():
    d = _load("chr2L_chunks.json.gz")
    assert [tuple(c) for c in d["chunks"]] == ch.plan_chunks(1, d["region"], d["chunksize"], d["overlap"])
    got = ch.join_predictions(d["concat"])
    assert got == d["joined"]
    n_in, n_out = d["concat"].count("# start gene"), got.count("# start gene")
    assert n_in > n_out > 50                                        # genes predicted twice in the overlaps were removed
    # streaming input (an iterable of lines) gives the same text
    assert ch.join_predictions(iter(d["concat"].splitlines(keepends=True))) == got


def test_chunk_runs_through_the_shim_equal_the_reference_pipeline(tmp_path):
    """run_chunks + join on the first 650 kb of chr2L with the CPU twin of the drop-in front end (oracle/_ref/augustus_emu: reference
    front end + host/augshim.cc + host build of the kernel source, tests/test_dropin_emu.py) — the per-chunk outputs and the joined
    text equal what the unmodified reference and its Perl scripts produced.  The GPU run of the same pipeline: tests/test_whole_chromosome.py."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("twc", os.path.join(util.ROOT, "tests", "test_whole_chromosome.py"))
    twc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(twc)
    emu = os.path.join(twc.REFDIR, "augustus_emu")
    if not (os.path.exists(emu) and os.path.exists(twc.CHR2L) and os.path.isdir(twc.CFG)):
        pytest.skip("oracle/_ref/augustus_emu or chr2L not present (make -C oracle ref dropin_emu)")
    d = _load("chr2L_chunks.json.gz")
    fa = str(tmp_path / "chr2L_650k.fa")
    twc._write_region(fa, d["region"])
    chunks = ch.plan_chunks(1, d["region"], d["chunksize"], d["overlap"])
    concat = ch.run_chunks(emu, fa, chunks, ["--species=fly"], env=dict(os.environ, AUGUSTUS_CONFIG_PATH=twc.CFG), jobs=4)
    assert twc._body(concat) == twc._body(d["concat"])
    assert twc._body(ch.join_predictions(concat)) == twc._body(d["joined"])

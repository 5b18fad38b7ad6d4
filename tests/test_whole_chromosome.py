"""Chunked prediction end to end on the GPU (pytest -m gpu): the first 650 kb of chr2L, `--species=fly` defaults, four overlapping
200 kb chunks decoded by the drop-in front end (oracle/_ref/augustus_b200 -> libaugb200.so), concatenated and joined by
augustus_b200/chromosome.py.  Expected: the text the unmodified reference + its Perl scripts produced for the same region
(tests/golden/chr2L_chunks.json.gz, made by tests/golden/make_golden_join.py) — every gene line, every posterior probability."""
import gzip
import json
import os

import pytest

from augustus_b200 import chromosome as ch
from tests import util

pytestmark = pytest.mark.gpu

REFDIR = os.path.join(util.ROOT, "oracle", "_ref")
DROPIN = os.path.join(REFDIR, "augustus_b200")
REF = os.path.join(REFDIR, "augustus")
CFG = os.path.join(REFDIR, "config")
CHR2L = os.path.join(REFDIR, "data", "chr2L.sm.fa.gz")


def _body(text):
    """Drop what depends on where the binaries and files live: header lines naming paths, and the echoed command lines."""
    keep, skip = [], False
    for l in text.splitlines():
        if l.startswith("# command line:"):
            skip = True
            continue
        if skip:
            skip = False
            continue
        if l.startswith("# Initializing the parameters using config directory") or l.startswith("# Looks like "):
            continue
        keep.append(l)
    return keep


needs_files = pytest.mark.skipif(not (os.path.exists(DROPIN) and os.path.isdir(CFG) and os.path.exists(CHR2L)),
                                 reason="oracle/_ref/augustus_b200 or oracle/_ref/data/chr2L.sm.fa.gz not present (make -C oracle ref dropin)")


def _write_region(path, length):
    seq, n = [], 0
    with gzip.open(CHR2L, "rt") as f:
        for line in f:
            if line.startswith(">"):
                continue
            seq.append(line.strip())
            n += len(seq[-1])
            if n >= length:
                break
    dna = "".join(seq)[:length]
    with open(path, "w") as f:
        f.write(">chr2L\n")
        for i in range(0, len(dna), 60):
            f.write(dna[i:i + 60] + "\n")


@needs_files
def test_chr2L_650kb_in_chunks_equals_reference_pipeline(tmp_path):
    with gzip.open(os.path.join(util.GOLDEN, "chr2L_chunks.json.gz"), "rt") as f:
        d = json.load(f)
    fa = str(tmp_path / "chr2L_650k.fa")
    _write_region(fa, d["region"])
    chunks = ch.plan_chunks(1, d["region"], d["chunksize"], d["overlap"])
    concat = ch.run_chunks(DROPIN, fa, chunks, ["--species=fly"], env=dict(os.environ, AUGUSTUS_CONFIG_PATH=CFG))
    assert _body(concat) == _body(d["concat"])
    joined = ch.join_predictions(concat)
    assert _body(joined) == _body(d["joined"])
    assert joined.count("# start gene") == d["joined"].count("# start gene") > 50


@needs_files
@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/augustus not built")
def test_chr2L_650kb_in_one_process_is_decoded_in_pieces_like_the_reference(tmp_path):
    """The same region in ONE front-end process: longer than fly's maxDNAPieceSize, so doViterbiPiecewise cuts it into four pieces
    at points its cut search finds (namgene.cc:973-1133); initial / terminal vectors swap at the cuts, sampling runs on every piece.
    Drop-in (GPU) and unmodified reference (CPU, the checker) must print the same GFF and cut at the same points."""
    import concurrent.futures as cf
    import subprocess
    fa = str(tmp_path / "chr2L_650k.fa")
    _write_region(fa, 650000)
    env = dict(os.environ, AUGUSTUS_CONFIG_PATH=CFG)

    def run(exe):
        r = subprocess.run([exe, "--species=fly", "--progress=true", fa], env=env, capture_output=True, text=True, timeout=3000)
        assert r.returncode == 0, r.stderr[-2000:]
        return _body(r.stdout), [l for l in r.stderr.splitlines() if l.startswith("examining piece")]
    with cf.ThreadPoolExecutor(2) as ex:
        fw, fg = ex.submit(run, REF), ex.submit(run, DROPIN)
        (want, wp), (got, gp) = fw.result(), fg.result()
    assert len(wp) == 4 and gp == wp
    assert sum(1 for l in want if "\tCDS\t" in l) > 200
    assert got == want


@needs_files
@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/augustus not built")
def test_example_fa_in_pieces_with_utr_and_sampling():
    """Small pieces on example.fa (--maxDNAPieceSize below the sequence length): four model variants (initial / terminal vectors at
    the cuts), UTR states, and --sample=100 whose rand() stream runs on across pieces while the cut search must not touch it.
    The same command lines pass against the CPU twin in tests/test_dropin_emu.py."""
    import subprocess
    fa = os.path.join(util.GOLDEN, "example.fa")
    env = dict(os.environ, AUGUSTUS_CONFIG_PATH=CFG)
    for args in (["--species=human", "--softmasking=0", "--maxDNAPieceSize=3000"],
                 ["--species=human", "--softmasking=0", "--UTR=on", "--maxDNAPieceSize=2500"],
                 ["--species=human", "--softmasking=0", "--maxDNAPieceSize=4000", "--sample=100", "--alternatives-from-sampling=true"]):
        outs = []
        for exe in (REF, DROPIN):
            r = subprocess.run([exe] + args + [fa], env=env, capture_output=True, text=True, timeout=1200)
            assert r.returncode == 0, r.stderr[-2000:]
            outs.append(_body(r.stdout))
        assert any("\tCDS\t" in l for l in outs[0])
        assert outs[0] == outs[1], args

"""Drop-in check (pytest -m gpu): the reference's own front end with its three DP entry points bound to
libaugb200.so (host/augshim.cc -> oracle/_ref/augustus_b200) must print the same GFF as the unmodified
reference binary (oracle/_ref/augustus) for the same command line.

Both binaries are built by oracle/Makefile in the build container (they need /root/reference) and travel to
the GPU box inside oracle/_ref/; without them the tests skip.  The reference binary is the checker here,
the GPU library is what runs the DP in augustus_b200."""
import os
import subprocess

import pytest

from augustus_b200 import synth
from tests import util

pytestmark = pytest.mark.gpu

REFDIR = os.path.join(util.ROOT, "oracle", "_ref")
REF = os.path.join(REFDIR, "augustus")
DROPIN = os.path.join(REFDIR, "augustus_b200")
CFG = os.path.join(REFDIR, "config")

needs_binaries = pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(DROPIN) and os.path.isdir(CFG)),
                                    reason="oracle/_ref/augustus{,_b200} not built (make -C oracle ref dropin)")


def _run(exe, args, fasta):
    env = dict(os.environ, AUGUSTUS_CONFIG_PATH=CFG)
    r = subprocess.run([exe] + list(args) + [fasta], env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stderr[-2000:] + r.stdout[-2000:]
    lines = r.stdout.splitlines()
    # the trailer echoes argv[0]; everything else must be identical
    if "# command line:" in lines:
        lines = lines[: lines.index("# command line:")]
    return lines


def _same_gff(args, fasta):
    want = _run(REF, args, fasta)
    got = _run(DROPIN, args, fasta)
    assert any(l.split("\t")[2:3] == ["CDS"] for l in want if not l.startswith("#")), "reference predicted nothing: weak test"
    assert got == want
    return want


@needs_binaries
def test_example_fa_human_ab_initio_gff_identical():
    """BASELINE.json configs[0]: examples/example.fa --species=human (two sequences, two GC classes)."""
    _same_gff(["--species=human", "--softmasking=0"], os.path.join(util.GOLDEN, "example.fa"))


@needs_binaries
def test_example_fa_utr_on_gff_identical_and_equals_reference_golden():
    """--UTR=on (71 states); the CDS / UTR features also equal the reference's own expected file
    tests/short/examples/expected_results/test_utr_on/aug_utr_on.gff (committed as tests/golden/aug_utr_on.gff)."""
    got = _same_gff(["--species=human", "--UTR=on", "--softmasking=0"], os.path.join(util.GOLDEN, "example.fa"))
    feats = [tuple(l.split("\t")[:8]) for l in got if not l.startswith("#")]
    gold = [tuple(l.rstrip("\n").split("\t")[:8]) for l in open(os.path.join(util.GOLDEN, "aug_utr_on.gff")) if not l.startswith("#") and "\t" in l]
    assert feats == gold


@needs_binaries
def test_sampling_posteriors_gff_identical():
    """--sample=100 --alternatives-from-sampling=true: the posterior probabilities and alternative transcripts in the GFF
    come from the Viterbi path + 99 sampled paths per sequence.  Two sequences in one process share one rand() stream
    (vitmatrix.cc:300); the shim carries the stream position across calls (augb200_set_rand_position)."""
    _same_gff(["--species=human", "--softmasking=0", "--sample=100", "--alternatives-from-sampling=true"],
              os.path.join(util.GOLDEN, "example.fa"))


@needs_binaries
def test_fly_defaults_softmasked_window_gff_identical(tmp_path):
    """--species=fly with its defaults (UTR on, softmasking on, sample=100) on a soft-masked chr2L window (config 3 shape)."""
    name, dna = util.read_fasta(os.path.join(util.GOLDEN, "fly_softmask_window.fa"))[0]
    fa = str(tmp_path / "fly.fa")
    synth.write_fasta(fa, [dna], [name])
    _same_gff(["--species=fly"], fa)


@needs_binaries
def test_synthetic_50k_window_gff_identical(tmp_path):
    fa = str(tmp_path / "syn.fa")
    synth.write_fasta(fa, [synth.window(3, 50000)], ["w3"])
    _same_gff(["--species=human", "--softmasking=0"], fa)

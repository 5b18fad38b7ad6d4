"""The drop-in shim's own logic on the CPU: oracle/_ref/augustus_emu is the reference front end + host/augshim.cc (exactly as in
the drop-in oracle/_ref/augustus_b200) linked against tests/hostemu/augb200_emu.cc — the C ABI served by the host build of the
kernel source — instead of libaugb200.so.  Its GFF must equal the unmodified reference binary's.  Covers what the shim adds on
top of the library: one model per (initProbs, termProbs) pair for the pieces of sequences longer than maxDNAPieceSize
(doViterbiPiecewise, namgene.cc:516-676, cut search :973-1133), sampled paths drawn lazily at the first getSampledPath call, the
rand() stream position carried across pieces and sequences.  Build container only (needs oracle/_ref/, made by oracle/Makefile);
the same command lines run against the GPU library in tests/test_dropin.py."""
import os
import subprocess

import pytest

from tests import util

REFDIR = os.path.join(util.ROOT, "oracle", "_ref")
REF = os.path.join(REFDIR, "augustus")
EMU = os.path.join(REFDIR, "augustus_emu")
CFG = os.path.join(REFDIR, "config")
EXAMPLE = os.path.join(util.GOLDEN, "example.fa")

pytestmark = pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(EMU) and os.path.isdir(CFG)),
                                reason="oracle/_ref/augustus{,_emu} not built (make -C oracle ref dropin_emu)")


def _run(exe, args, extra_env=None, fasta=EXAMPLE):
    env = dict(os.environ, AUGUSTUS_CONFIG_PATH=CFG, **(extra_env or {}))
    r = subprocess.run([exe] + list(args) + [fasta], env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stderr[-2000:] + r.stdout[-2000:]
    lines = r.stdout.splitlines()
    if "# command line:" in lines:                     # the trailer echoes argv[0]
        lines = lines[: lines.index("# command line:")]
    return lines, r.stderr


def _same(args):
    want, _ = _run(REF, args)
    got, err = _run(EMU, args, {"AUGSHIM_VERBOSE": "1"})
    assert sum(1 for l in want if "\tCDS\t" in l) >= 5, "reference predicted nothing: weak test"
    assert got == want
    return err


def test_single_piece_sequences():
    err = _same(["--species=human", "--softmasking=0"])
    assert err.count("augshim: model") == 1


def test_pieces_use_one_model_per_initial_terminal_vector_pair():
    """maxDNAPieceSize below the sequence length: first / inner / last pieces swap initProbs and termProbs for the synch-state
    vectors (namgene.cc:594-603) -> four model variants; the cut search decodes exam chunks with the vectors of the previous piece."""
    err = _same(["--species=human", "--softmasking=0", "--maxDNAPieceSize=3000"])
    assert err.count("augshim: model") == 4
    _same(["--species=human", "--softmasking=0", "--UTR=on", "--maxDNAPieceSize=2500"])


def test_pieces_with_sampling_keep_the_rand_stream():
    """--sample=100 on pieces: the cut search runs the DP without sampling, so the sampled paths (and the posterior probabilities in
    the GFF) only stay identical if the shim draws them lazily and carries the stream position across pieces and sequences."""
    _same(["--species=human", "--softmasking=0", "--maxDNAPieceSize=4000", "--sample=100", "--alternatives-from-sampling=true"])


def test_fly_defaults_260kb_two_pieces_cut_by_the_cut_search(tmp_path):
    """--species=fly defaults (UTR states, softmasking, sample=100) on the first 260 kb of chr2L: longer than fly's maxDNAPieceSize
    (200 000), so the cut search decodes an exam chunk around the end of the range and cuts inside a predicted intergenic region."""
    import concurrent.futures as cf
    import gzip
    src = os.path.join(REFDIR, "data", "chr2L.sm.fa.gz")
    if not os.path.exists(src):
        pytest.skip("oracle/_ref/data/chr2L.sm.fa.gz not present")
    seq, n = [], 0
    with gzip.open(src, "rt") as f:
        for line in f:
            if not line.startswith(">"):
                seq.append(line.strip())
                n += len(seq[-1])
                if n >= 260000:
                    break
    fa = str(tmp_path / "chr2L_260k.fa")
    with open(fa, "w") as f:
        f.write(">chr2L\n" + "".join(seq)[:260000] + "\n")
    with cf.ThreadPoolExecutor(2) as ex:
        fw = ex.submit(_run, REF, ["--species=fly", "--progress=true"], None, fa)
        fg = ex.submit(_run, EMU, ["--species=fly", "--progress=true"], {"AUGSHIM_VERBOSE": "1"}, fa)
        (want, werr), (got, gerr) = fw.result(), fg.result()
    pieces = [l for l in werr.splitlines() if l.startswith("examining piece")]
    assert len(pieces) == 2 and not pieces[0].startswith("examining piece 1..200000 "), pieces      # a cut found by the search, not the fallback
    assert [l for l in gerr.splitlines() if l.startswith("examining piece")] == pieces
    assert sum(1 for l in want if "\tCDS\t" in l) >= 50
    assert got == want


def test_ragged_input_softmasked_runs_unknown_bases_tiny_sequences(tmp_path):
    """One multi-FASTA with what real input holds: lower-case (soft-masked) runs with softmasking at its default (on), runs of N, IUPAC
    codes, a 57-base and a 6-base sequence, an all-N sequence — through the front end, the shim and the kernel source, compared with the
    unmodified reference (Viterbi, UTR states, fly defaults with sampling, and pieces with sampling)."""
    import random
    seqs = util.read_fasta(EXAMPLE)
    rng = random.Random(5)
    d = list(seqs[0][1])
    for _ in range(25):
        a = rng.randrange(len(d))
        for i in range(a, min(len(d), a + rng.randrange(10, 400))):
            d[i] = d[i].lower()
    for _ in range(6):
        a = rng.randrange(len(d))
        for i in range(a, min(len(d), a + rng.randrange(1, 60))):
            d[i] = "N"
    fa = str(tmp_path / "ragged.fa")
    with open(fa, "w") as f:
        f.write(">masked_with_N\n" + "".join(d) + "\n>short\n" + seqs[1][1][:57] + "\n>tiny\nACGTAC\n>allN\n" + "N" * 300 + "\n>iupac\n"
                + seqs[1][1][:1200].replace("A", "R", 3).replace("C", "y", 4) + "\n")
    for args in (["--species=human"], ["--species=human", "--UTR=on"], ["--species=fly"],
                 ["--species=human", "--softmasking=0", "--sample=50", "--alternatives-from-sampling=true", "--maxDNAPieceSize=2000"]):
        want, _ = _run(REF, args, fasta=fa)
        got, _ = _run(EMU, args, fasta=fa)
        assert any("\tCDS\t" in l for l in want)
        assert got == want, args


@pytest.mark.parametrize("args,needle", [
    (["--genemodel=intronless"], "not exportable"),                                          # no intron tables in that model
    (["--genemodel=exactlyone"], "duplicate state role"),                                    # two intergenic states
    (["--genemodel=bacterium"], "overlap mode"),
    (["--emiprobs=on"], "--emiprobs needs the DP matrices"),                                 # re-scores paths against the CPU matrices
])
def test_configurations_outside_the_scope_end_with_an_error_not_with_other_output(args, needle):
    """No silent approximation: what the library does not decode stops the front end with a message (the reference's ProjectError path)."""
    env = dict(os.environ, AUGUSTUS_CONFIG_PATH=CFG)
    r = subprocess.run([EMU, "--species=human", "--softmasking=0"] + args + [EXAMPLE], env=env, capture_output=True, text=True, timeout=600)
    assert needle in r.stderr + r.stdout
    assert "\tCDS\t" not in r.stdout


def test_sampling_temperature_gff_identical():
    """--temperature=3 (heated forward sums, LLDouble::heated): posterior probabilities and alternative transcripts of the GFF through the
    twin == the reference (round 1 refused the option)."""
    args = ["--species=human", "--softmasking=0", "--sample=100", "--alternatives-from-sampling=true", "--temperature=3"]
    want, _ = _run(REF, args)
    got, _ = _run(EMU, args)
    assert any("\tCDS\t" in l for l in want) and got == want

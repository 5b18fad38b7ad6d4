#!/usr/bin/env python
"""Golden vectors for augustus_b200/chromosome.py (build container only: needs /root/reference and oracle/_ref/augustus).

1. chr2L_chunks.json.gz — the first 650 kb of examples/chr2L run the way the reference runs genomes: chunk borders from the joblist
   of scripts/createAugustusJoblist.pl (chunksize 200000, overlap 50000), one UNMODIFIED `augustus --species=fly` process per chunk
   (defaults: UTR on, softmasking on, sample=100; --predictionStart/--predictionEnd as in the joblist), outputs concatenated in order,
   joined with scripts/join_aug_pred.pl.  Stored: the chunk borders, the concatenation and the joined text.
2. join_cases.json.gz — synthetic runs (gene blocks with random coordinates: disjoint, nested, chained overlaps, several sequences,
   empty runs, a drop list, gff3) pushed through the same Perl script, plus joblists of createAugustusJoblist.pl for a list of
   (start, end, chunksize, overlap, padding).
"""
import concurrent.futures as cf
import gzip
import json
import os
import random
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
AUG = os.path.join(ROOT, "oracle", "_ref", "augustus")
CFG = os.path.join(ROOT, "oracle", "_ref", "config")
CHR2L = os.path.join(ROOT, "oracle", "_ref", "data", "chr2L.sm.fa.gz")
JOIN = os.path.join(REF, "scripts", "join_aug_pred.pl")
JOBLIST = os.path.join(REF, "scripts", "createAugustusJoblist.pl")
REGION = 650000


def perl_join(text, droplist=None):
    args = ["perl", JOIN]
    with tempfile.TemporaryDirectory() as td:
        if droplist is not None:
            dl = os.path.join(td, "drop")
            open(dl, "w").write("".join(g + "\n" for g in droplist))
            args.append("--droplist=" + dl)
        r = subprocess.run(args, input=text, capture_output=True, text=True, check=True)
    return r.stdout


def perl_chunks(start, end, chunksize, overlap, padding):
    with tempfile.TemporaryDirectory() as td:
        lst, jobs = os.path.join(td, "seqs.lst"), os.path.join(td, "jobs.lst")
        open(lst, "w").write("/data/x/chrT.fa\t%d\t%d\n" % (start, end))
        args = ["perl", JOBLIST, "--sequences=" + lst, "--wrap=", "--chunksize=%d" % chunksize, "--outputdir=out", "--joblist=" + jobs, "--command=augustus"]
        if overlap:
            args.append("--overlap=%d" % overlap)
        if padding:
            args.append("--padding=%d" % padding)
        subprocess.run(args, check=True, capture_output=True)
        rows = []
        for line in open(jobs):
            m = re.search(r"--predictionStart=(-?\d+) --predictionEnd=(-?\d+) .* --outfile=(\S+)", line)
            rows.append([int(m.group(1)), int(m.group(2)), m.group(3)])
    return rows


def make_chunks():
    head, seq = None, []
    n = 0
    with gzip.open(CHR2L, "rt") as f:
        for line in f:
            if line.startswith(">"):
                head = line.strip()
                continue
            seq.append(line.strip())
            n += len(seq[-1])
            if n >= REGION:
                break
    dna = "".join(seq)[:REGION]
    chunks = perl_chunks(1, REGION, 200000, 50000, 0)
    td = tempfile.mkdtemp()
    fa = os.path.join(td, "chr2L_650k.fa")
    with open(fa, "w") as f:
        f.write(">chr2L\n")
        for i in range(0, len(dna), 60):
            f.write(dna[i:i + 60] + "\n")

    def run(c):
        r = subprocess.run([AUG, "--species=fly", "--predictionStart=%d" % c[0], "--predictionEnd=%d" % c[1], fa],
                           env=dict(os.environ, AUGUSTUS_CONFIG_PATH=CFG), capture_output=True, text=True, check=True)
        return r.stdout
    with cf.ThreadPoolExecutor(8) as ex:
        outs = list(ex.map(run, chunks))
    concat = "".join(outs).replace(fa, "chr2L_650k.fa")
    return {"region": REGION, "chunksize": 200000, "overlap": 50000, "chunks": [c[:2] for c in chunks], "concat": concat, "joined": perl_join(concat)}


def synth_run(rng, seqname, genes, gff3=False, with_header=True, first_id=1):
    """Text of one run in the layout of the reference's output: separator + header, per gene a `# start gene` block."""
    t = []
    if with_header:
        t.append("# This output was generated with AUGUSTUS (version 3.5.0).\n# some header line %d\n" % rng.randrange(1000))
        if gff3:
            t.append("##gff-version 3\n")
        t.append("# ----- prediction on sequence number 1 (length = 1000, name = %s) -----\n#\n# Predicted genes for sequence number 1 on both strands\n" % seqname)
    for k, (b, e) in enumerate(genes):
        g = "g%d" % (first_id + k)
        t.append("# start gene %s\n" % g)
        t.append("%s\tAUGUSTUS\tgene\t%d\t%d\t0.5\t+\t.\t%s\n" % (seqname, b, e, g))
        ntx = rng.choice([1, 1, 2])
        for x in range(1, ntx + 1):
            t.append('%s\tAUGUSTUS\ttranscript\t%d\t%d\t0.4\t+\t.\t%s.t%d\n' % (seqname, b, e, g, x))
            t.append('%s\tAUGUSTUS\tCDS\t%d\t%d\t0.9\t+\t0\ttranscript_id "%s.t%d"; gene_id "%s";\n' % (seqname, b, e, g, x, g))
            t.append("# protein sequence = [MAg%dX]\n" % rng.randrange(50))
        t.append("# end gene %s\n###\n" % g)
    t.append("# command line:\n# augustus --species=x file\n")
    return "".join(t)


def make_cases():
    rng = random.Random(20260923)
    cases = []

    def genes(lo, hi, n):
        out = []
        for _ in range(n):
            b = rng.randrange(lo, hi)
            out.append((b, b + rng.randrange(1, max(2, (hi - lo) // 3))))
        return sorted(out) if rng.random() < 0.8 else out
    for ci in range(150):
        nruns = rng.choice([1, 2, 2, 3, 4, 6])
        gff3 = rng.random() < 0.15
        text, pos, name = [], 0, "chrA"
        if rng.random() < 0.2:
            text.append("stray line before the first run\n")
        for r in range(nruns):
            if rng.random() < 0.15:
                name = "chr" + rng.choice("ABC")
            n = rng.choice([0, 1, 2, 3, 5]) if rng.random() < 0.9 else 0
            text.append(synth_run(rng, name, genes(pos, pos + 1000, n), gff3=gff3))
            pos += rng.choice([0, 300, 600, 900, 1200])
        drop = None
        if rng.random() < 0.2:
            drop = ["g%d%s" % (rng.randrange(1, 4), rng.choice(["", ".t1"])) for _ in range(rng.randrange(1, 3))]
        t = "".join(text)
        if rng.random() < 0.1:
            t = t.rstrip("\n")
        cases.append({"input": t, "droplist": drop, "joined": perl_join(t, drop)})
    cases.append({"input": "", "droplist": None, "joined": perl_join("")})
    plans = []
    for (s, e, cs, ov, pad) in [(1, 650000, 200000, 50000, 0), (1, 23513712, 200000, 50000, 0), (1, 1000, 300, 100, 0), (1, 1000, 1000, 10, 0),
                                (1, 1000, 999, 1, 0), (1, 1001, 500, 250, 0), (5001, 9000, 1000, 999, 0), (1, 20000000, 4000000, 0, 0),
                                (1, 1000, 300, 100, 50), (1, 10, 300, 100, 0), (1, 1200, 400, 200, 7)]:
        plans.append({"start": s, "end": e, "chunksize": cs, "overlap": ov, "padding": pad, "rows": perl_chunks(s, e, cs, ov, pad)})
    return {"cases": cases, "plans": plans}


if __name__ == "__main__":
    what = sys.argv[1:] or ["cases", "chunks"]
    if "cases" in what:
        with gzip.open(os.path.join(HERE, "join_cases.json.gz"), "wt", compresslevel=9) as f:
            json.dump(make_cases(), f)
    if "chunks" in what:
        with gzip.open(os.path.join(HERE, "chr2L_chunks.json.gz"), "wt", compresslevel=9) as f:
            json.dump(make_chunks(), f)

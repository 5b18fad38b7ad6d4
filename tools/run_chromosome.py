#!/usr/bin/env python
"""Predict a long sequence in overlapping chunks and join the chunk predictions (augustus_b200/chromosome.py): what the reference does
with createAugustusJoblist.pl + a cluster + join_aug_pred.pl, in one command.  Needs a GPU for the default front end.
usage: run_chromosome.py [--exe=oracle/_ref/augustus_b200] [--chunksize=200000] [--overlap=50000] [--jobs=4] seq.fa [augustus options ...] > joined.gff"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from augustus_b200 import chromosome as ch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
opt = {"exe": os.path.join(ROOT, "oracle", "_ref", "augustus_b200"), "chunksize": "200000", "overlap": "50000", "jobs": "4"}
rest = []
for a in sys.argv[1:]:
    k = a[2:].split("=", 1)[0] if a.startswith("--") else None
    if k in opt:
        opt[k] = a.split("=", 1)[1]
    else:
        rest.append(a)
fasta = [a for a in rest if not a.startswith("--")]
if len(fasta) != 1:
    sys.exit(__doc__)
length = sum(len(l.strip()) for l in open(fasta[0]) if not l.startswith(">"))
env = dict(os.environ)
env.setdefault("AUGUSTUS_CONFIG_PATH", os.path.join(ROOT, "oracle", "_ref", "config"))
sys.stdout.write(ch.predict_chromosome(opt["exe"], fasta[0], length, int(opt["chunksize"]), int(opt["overlap"]),
                                       [a for a in rest if a.startswith("--")], env, jobs=int(opt["jobs"])))

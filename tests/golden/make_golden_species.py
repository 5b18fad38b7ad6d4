#!/usr/bin/env python
"""Golden fixtures for parameter sets whose splice-signal geometry differs from human / fly (build container only).

tools/species_sweep.py runs all 167 species directories of the reference through the oracle and the kernel source; two of the sets that
exposed a defect (an exon following a splice-site state of column 0, possible when the exon part of the signal is shorter than two
bases) are kept as fixtures:  nasonia (ass_end = 0, 5 GC classes)  and  Monosiga_brevicollis (dss_start = 1);  zebrafish stands for the
TRANSINITBIN sets (start codon x TIS motif probability mapped to bins, exonmodel.cc:1321-1326).
Writes <species>.params.xz and ref_paths_species.json (the reference's paths + scores on example.fa, --UTR=off --softmasking=0)."""
import json
import lzma
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden import AUGDUMP, ENV, REF, condense  # noqa: E402

SPECIES = ["nasonia", "Monosiga_brevicollis", "zebrafish"]


def main():
    out = {}
    for sp in SPECIES:
        with tempfile.TemporaryDirectory() as td:
            blob, pf = os.path.join(td, "b"), os.path.join(td, "p")
            subprocess.run([AUGDUMP, "--species=" + sp, "--UTR=off", "--softmasking=0", REF + "/examples/example.fa"],
                           env=dict(ENV, AUGDUMP_PARAMS=blob, AUGDUMP_PATH=pf), check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            with lzma.open(os.path.join(HERE, sp + ".params.xz"), "wb", preset=9) as f:
                f.write(open(blob, "rb").read())
            seqs, cur = [], None
            for line in open(pf):
                t = line.split()
                if t[0] == "seq":
                    cur = {"name": t[1], "length": int(t[2]), "log_prob": float(t[4]), "states": []}; seqs.append(cur)
                elif t[0] == "state":
                    cur["states"].append([int(v) for v in t[1:5]])
            for s in seqs:
                s["states"] = condense(s["states"])
            out[sp] = seqs
    json.dump(out, open(os.path.join(HERE, "ref_paths_species.json"), "w"), separators=(",", ":"))
    print("species fixtures written to", HERE)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Fixtures that PIN the known, documented differences from the reference (build container only).

  aedes.params.xz + ref_samples_aedes.json.gz
      --species=aedes --UTR=off (three GC classes, softmasking and sample=100 by the species' defaults) on the soft-masked 40 kb
      chr2L window of fly_softmask_window.fa: the reference's 99 sampled paths.  The oracle draws all 99; the kernel source draws 96
      (sampling steps of lessD inside the columns around a GC-class boundary do not see the SnippetProbs memo of the forward pass).
"""
import gzip
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
from make_golden import AUGDUMP, ENV, condense  # noqa: E402


def main():
    fa = os.path.join(HERE, "fly_softmask_window.fa")
    with tempfile.TemporaryDirectory() as td:
        blob, pf = os.path.join(td, "b"), os.path.join(td, "p")
        subprocess.run([AUGDUMP, "--species=aedes", "--UTR=off", fa], env=dict(ENV, AUGDUMP_PARAMS=blob, AUGDUMP_PATH=pf), check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        with lzma.open(os.path.join(HERE, "aedes.params.xz"), "wb", preset=9) as f:
            f.write(open(blob, "rb").read())
        samples = []
        for line in open(pf):
            if not line.startswith("s"):
                continue
            t = line.split()
            if t[0] == "sample":
                samples.append({"log_prob": float(t[3]), "states": []})
            elif t[0] == "sstate":
                samples[-1]["states"].append([int(v) for v in t[1:5]])
        for s in samples:
            s["states"] = condense(s["states"])
    with gzip.open(os.path.join(HERE, "ref_samples_aedes.json.gz"), "wt") as f:
        json.dump({"command": "augustus --species=aedes --UTR=off fly_softmask_window.fa", "samples": samples}, f, separators=(",", ":"))
    print("wrote aedes fixtures:", len(samples), "sampled paths")


if __name__ == "__main__":
    main()

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


def heated():
    """ref_samples_heated.json.gz: the reference's 99 sampled paths of both sequences of example.fa with --temperature=3 (LLDouble::heated in
    the forward sums; one process, the second sequence continues the rand() stream of the first)."""
    fa = os.path.join(HERE, "example.fa")
    with tempfile.TemporaryDirectory() as td:
        pf, blob = os.path.join(td, "p"), os.path.join(td, "b")
        subprocess.run([AUGDUMP, "--species=human", "--softmasking=0", "--sample=100", "--temperature=3", fa], env=dict(ENV, AUGDUMP_PATH=pf, AUGDUMP_PARAMS=blob),
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        with lzma.open(os.path.join(HERE, "human_t3.params.xz"), "wb", preset=9) as f:       # the human blob with temperature = 3
            f.write(open(blob, "rb").read())
        seqs = []
        for line in open(pf):
            t = line.split()
            if t[0] == "seq":
                seqs.append({"name": t[1], "samples": []})
            elif t[0] == "sample":
                seqs[-1]["samples"].append([])
            elif t[0] == "sstate":
                seqs[-1]["samples"][-1].append([int(v) for v in t[1:5]])
    for s in seqs:
        s["samples"] = [condense(x) for x in s["samples"]]
    with gzip.open(os.path.join(HERE, "ref_samples_heated.json.gz"), "wt") as f:
        json.dump({"command": "augustus --species=human --softmasking=0 --sample=100 --temperature=3 example.fa (one process, both sequences)",
                   "sequences": seqs}, f, separators=(",", ":"))


if __name__ == "__main__" and "--heated" in sys.argv:
    heated()

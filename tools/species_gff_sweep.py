"""GFF-level parity sweep over the species parameter sets of the reference (build container only): for every config/species/<name>
run the unmodified `augustus` and the CPU twin of the drop-in (oracle/_ref/augustus_emu: reference front end + host/augshim.cc + host
build of the kernel source) on tests/golden/example.fa with the species' own defaults (most sample 100 paths) and compare the output
byte for byte.  Complements species_sweep.py (Viterbi cells) with the sampled paths / posterior probabilities per species.
usage: species_gff_sweep.py [--utr] [--fasta=file.fa --softmask] [species ...]   (--softmask: softmasking at its default, on)"""
import concurrent.futures as cf
import os
import subprocess
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import util

REF = "/root/reference"
BIN = os.path.join(util.ROOT, "oracle", "_ref")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
extra = ["--UTR=on"] if "--utr" in sys.argv else ["--UTR=off"]
extra += [] if "--softmask" in sys.argv else ["--softmasking=0"]
FASTA = ([a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--fasta=")] or [util.GOLDEN + "/example.fa"])[0]
species = args or sorted(os.listdir(REF + "/config/species"))
env = dict(os.environ, AUGUSTUS_CONFIG_PATH=REF + "/config")


def body(text):
    lines = text.splitlines()
    return lines[: lines.index("# command line:")] if "# command line:" in lines else lines


def one(sp):
    cmd = ["--species=" + sp] + extra + [FASTA]
    try:
        r = subprocess.run([BIN + "/augustus"] + cmd, env=env, capture_output=True, text=True, timeout=600)
        e = subprocess.run([BIN + "/augustus_emu"] + cmd, env=env, capture_output=True, text=True, timeout=600)
    except subprocess.TimeoutExpired:
        return sp, "timeout", ""
    if r.returncode != 0:
        return sp, "reference failed", (r.stderr.strip().splitlines() or ["?"])[-1][:100]
    if e.returncode != 0:
        return sp, "rejected", (e.stderr.strip().splitlines() or ["?"])[-1][:100]
    ncds = sum(1 for l in body(r.stdout) if "\tCDS\t" in l)
    return (sp, "ok", "%d CDS lines" % ncds) if body(r.stdout) == body(e.stdout) else (sp, "MISMATCH", "")


tot = {}
with cf.ThreadPoolExecutor(8) as ex:
    for sp, st, msg in ex.map(one, species):
        print(sp, st, msg, flush=True)
        tot[st] = tot.get(st, 0) + 1
print(tot)

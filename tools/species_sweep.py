"""Parity sweep over the species parameter sets of the reference (build container only: needs /root/reference and oracle/_ref/augdump).
For every config/species/<name>: export the blob, dump the reference's Viterbi matrix on a short sequence, load the blob into the oracle
and into the kernel source (host build), compare every cell.  Prints one line per species: ok / rejected (why) / MISMATCH.
usage: species_sweep.py [--utr | --nc] [species ...]"""
import os, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import util

REF = "/root/reference"
AUGDUMP = os.path.join(util.ROOT, "oracle", "_ref", "augdump")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
utr = "--utr" in sys.argv or "--nc" in sys.argv
nc = "--nc" in sys.argv
species = args or sorted(os.listdir(REF + "/config/species"))
name, dna = util.read_fasta(util.GOLDEN + "/example.fa")[1]
tot = {"ok": 0, "rejected": 0, "MISMATCH": 0, "reference failed": 0}
for sp in species:
    with tempfile.TemporaryDirectory() as td:
        blobf, mat, fa = os.path.join(td, "b"), os.path.join(td, "m"), os.path.join(td, "w.fa")
        open(fa, "w").write(">%s\n%s\n" % (name, dna))        # ONE sequence per process: the reference keeps per-class memos across sequences
        cmd = [AUGDUMP, "--species=" + sp, "--softmasking=0"] + (["--UTR=on"] if utr else ["--UTR=off"]) + (["--nc=on"] if nc else []) + [fa]
        env = dict(os.environ, AUGUSTUS_CONFIG_PATH=REF + "/config", AUGDUMP_PARAMS=blobf, AUGDUMP_MATRIX=mat, AUGDUMP_PATH=os.path.join(td, "p"))
        try:
            r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=120)
        except subprocess.TimeoutExpired:
            print(sp, "reference failed: timeout"); tot["reference failed"] += 1; continue
        if r.returncode != 0 or not os.path.exists(blobf) or not os.path.exists(mat + ".1.vit"):
            print(sp, "reference failed:", (r.stderr.strip().splitlines() or ["?"])[-1][:100]); tot["reference failed"] += 1; continue
        blob = open(blobf, "rb").read()
        from augustus_b200 import params as _p
        pp = _p.parse(blob)
        key = " ".join("%s=%d" % (k2, int(pp[k][0])) for k2, k in (("C", "num_gc_classes"), ("k", "exon_k"), ("ds", "dss_start"), ("de", "dss_end"), ("as", "ass_start"), ("ae", "ass_end"), ("up", "ass_upwindow_size"), ("tiw", "trans_init_window"), ("il", "init_coding_len"), ("et", "et_coding_len"), ("mel", "min_exon_length"), ("d", "intron_d"), ("tm", "tis_motif_memory")))
        try:
            emu = util.HostEmu(blob)
        except Exception as ex:
            print(sp, "rejected by the product's model builder:", str(ex)[:90]); tot["rejected"] += 1; continue
        S = emu.lib.hostemu_statecount(__import__("ctypes").c_void_p(emu.m))
        V = np.fromfile(mat + ".1.vit").reshape(-1, S)
        try:
            orc = util.Oracle(blob)
        except Exception as ex:
            print(sp, "rejected by the oracle:", str(ex)[:90]); tot["rejected"] += 1; continue
        o = orc.viterbi(dna, want_matrix=True)
        e = emu.decode(dna, want_cells=True)
        fo, fv = o["V"] > util.NEGT, np.isfinite(V)
        ok_ref = (fo == fv).all() and np.abs((o["V"][fo & fv] / 2.0 ** 40) - V[fo & fv]).max() <= 1e-7
        ok_emu = e["status"] == 0 and e["states"] == o["condensed"] and e["log_prob"] == o["log_prob"] and ((e["cells"] > util.NEGT) == fo).all() and (e["cells"][fo] == o["V"][fo]).all()
        if ok_ref and ok_emu:
            print(sp, "ok  (S=%d, path states %d)" % (S, len(o["condensed"])), key); tot["ok"] += 1
        else:
            print(sp, "MISMATCH oracle-vs-reference ok=%s kernel-vs-oracle ok=%s" % (ok_ref, ok_emu), key); tot["MISMATCH"] += 1
print(tot)

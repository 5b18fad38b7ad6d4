"""Test helpers: golden fixtures, the oracle (CPU restatement) and the host emulator of the kernel source.

The oracle and the emulator are TEST INFRASTRUCTURE: nothing in augustus_b200/ imports this module.
"""
import ctypes
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_SO = os.path.join(ROOT, "oracle", "libghmm_oracle.so")
HOSTEMU_SO = os.path.join(ROOT, "tests", "hostemu", "libaugb200_hostemu.so")
NEGT = -(1 << 60)
FRAC = 40


def blob_bytes(name="human"):
    """Parameter blobs exported from the reference: 'human' (--species=human) or 'fly_noutr' (--species=fly --UTR=off)."""
    from augustus_b200 import params
    return params.load_bytes(os.path.join(GOLDEN, name + ".params.xz"))


def golden_samples():
    import gzip
    return json.load(gzip.open(os.path.join(GOLDEN, "ref_samples.json.gz"), "rt"))


def golden_paths(utr=False):
    return json.load(open(os.path.join(GOLDEN, "ref_paths_utr.json" if utr else "ref_paths.json")))


def golden_samples_utr():
    import gzip
    return json.load(gzip.open(os.path.join(GOLDEN, "ref_samples_utr.json.gz"), "rt"))


def path_features(states, par):
    """Biological coordinates (1-based, as in GFF) of what a condensed state path predicts: CDS parts, transcription start and
    end sites.  The state geometry is the reference's (ExonModel::ExonModel exonmodel.cc:231-279: a coding-exon state begins
    beginPartLen - innerPartOffset bases before / after the exon and ends baseOffset bases before it; UtrModel::getEndPositions
    utrmodel.cc:1572-1643).  `par` = parsed parameter blob."""
    tiw, dss_start, ass_end, tuw = (int(par[k][0]) for k in ("trans_init_window", "dss_start", "ass_end", "tss_upwindow_size"))
    cds, tss, tts = [], [], []
    for t, b, e, tr in states:
        if t in (1, 2, 3, 4):            # single, initial: the state starts trans_init_window bases before the start codon
            cds.append((b + tiw + 1, e + (0 if t == 1 else dss_start) + 1, "+"))
        elif 5 <= t <= 8:                # internal, terminal
            cds.append((b - ass_end + 1, e + (0 if t == 8 else dss_start) + 1, "+"))
        elif t in (36, 37):              # rsingle, rinitial end trans_init_window bases behind the (reverse) start codon
            cds.append((b - (0 if t == 36 else dss_start) + 1, e - tiw + 1, "-"))
        elif 38 <= t <= 43:              # rinternal, rterminal
            cds.append((b - (dss_start if t <= 40 else 0) + 1, e + ass_end + 1, "-"))
        elif t in (24, 25) and not (tr & 1) and b >= -tuw + 1:      # utr5single, utr5init begin with the TSS window
            tss.append((b + tuw + 1, "+"))
        elif t in (30, 35) and not (tr & 2):
            tts.append((e + 1, "+"))
        elif t in (59, 60):
            tss.append((e - tuw + 1, "-"))
        elif t in (65, 70) and not (tr & 1) and b > 0:
            tts.append((b + 1, "-"))
    return {"CDS": cds, "tss": tss, "tts": tts}


def gff_features(path, seqname):
    out = {"CDS": [], "tss": [], "tts": []}
    for line in open(path):
        f = line.rstrip("\n").split("\t")
        if len(f) < 8 or f[0] != seqname or f[2] not in out:
            continue
        out[f[2]].append((int(f[3]), int(f[4]), f[6]) if f[2] == "CDS" else (int(f[3]), f[6]))
    return out


def read_fasta(path):
    seqs, name, buf = [], None, []
    for line in open(path):
        if line.startswith(">"):
            if name is not None:
                seqs.append((name, "".join(buf)))
            name, buf = line[1:].split()[0], []
        else:
            buf.append(line.strip())
    seqs.append((name, "".join(buf)))
    return seqs


def condense(states):
    """StatePath::condenseStatePath (reference gene.cc:977-1000) on (type, begin, end, trunc) tuples."""
    out = []
    for t, b, e, tr in states:
        coding = (1 <= t <= 8) or (36 <= t <= 43)
        if out and out[-1][0] == t and not coding:
            out[-1][2] = e
            out[-1][3] |= tr
        else:
            out.append([t, b, e, tr])
    return [tuple(x) for x in out]


def gc_from_runs(runs, length):
    gc = np.zeros(length, dtype=np.int32)
    for pos, idx in runs:
        gc[pos:] = idx
    return gc


def _ensure(path, cmd, cwd):
    if not os.path.exists(path):
        subprocess.run(cmd, cwd=cwd, check=True)


class Oracle:
    def __init__(self, blob: bytes):
        _ensure(ORACLE_SO, ["make", "liboracle"], os.path.join(ROOT, "oracle"))
        self.lib = ctypes.CDLL(ORACLE_SO)
        self.lib.orc_model_load.restype = ctypes.c_void_p
        self.lib.orc_model_load.argtypes = [ctypes.c_char_p]
        self.lib.orc_model_statecount.argtypes = [ctypes.c_void_p]
        self.lib.orc_decode.restype = ctypes.c_int
        self.lib.orc_decode.argtypes = ([ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
                                        + [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 6)
        self.lib.orc_viterbi.restype = ctypes.c_int
        self.lib.orc_viterbi.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 5
        import tempfile
        with tempfile.NamedTemporaryFile(suffix=".blob", delete=False) as f:
            f.write(blob)
            self._tmp = f.name
        self.m = self.lib.orc_model_load(self._tmp.encode())
        os.unlink(self._tmp)
        assert self.m
        self.S = self.lib.orc_model_statecount(self.m)

    def viterbi(self, dna: str, gc=None, want_matrix=False):
        L = len(dna)
        V = np.zeros((L, self.S), dtype=np.int64) if want_matrix else None
        gco = np.zeros(L, dtype=np.int32)
        cap = L + 16
        pt, pb, pe, ptr = (np.zeros(cap, dtype=np.int32) for _ in range(4))
        lp = ctypes.c_double()
        gci = None if gc is None else np.ascontiguousarray(gc, dtype=np.int32)
        n = self.lib.orc_viterbi(self.m, dna.encode(), L, None if gci is None else gci.ctypes.data,
                                 None if V is None else V.ctypes.data, gco.ctypes.data, cap,
                                 pt.ctypes.data, pb.ctypes.data, pe.ctypes.data, ptr.ctypes.data, ctypes.byref(lp))
        states = [(int(pt[k]), int(pb[k]), int(pe[k]), int(ptr[k])) for k in range(max(n, 0))]
        return {"n": n, "states": states, "condensed": condense(states), "log_prob": lp.value, "gc": gco, "V": V}


    def sample(self, dna: str, nsample=100, gc=None, want_forward=False):
        """Viterbi + forward fill + (nsample-1) sampled paths, as NAMGene::findGenes with sample=nsample."""
        L = len(dna)
        cap = L + 16
        pa = [np.zeros(cap, dtype=np.int32) for _ in range(4)]
        lp = ctypes.c_double()
        F = np.zeros((L, self.S)) if want_forward else None
        scap = nsample * (L // 8 + 64)
        sa = [np.zeros(scap, dtype=np.int32) for _ in range(4)]
        sc = np.zeros(nsample, dtype=np.int32)
        slp = np.zeros(nsample)
        gci = None if gc is None else np.ascontiguousarray(gc, dtype=np.int32)
        n = self.lib.orc_decode(self.m, dna.encode(), L, None if gci is None else gci.ctypes.data, None, None, cap,
                                *[a.ctypes.data for a in pa], ctypes.byref(lp), nsample, None if F is None else F.ctypes.data,
                                scap, *[a.ctypes.data for a in sa], sc.ctypes.data, slp.ctypes.data)
        samples, pos = [], 0
        for k in range(nsample - 1):
            c = int(sc[k])
            samples.append({"log_prob": float(slp[k]),
                            "states": [(int(sa[0][pos + q]), int(sa[1][pos + q]), int(sa[2][pos + q]), int(sa[3][pos + q])) for q in range(max(c, 0))]})
            pos += max(c, 0)
        viterbi = condense([(int(pa[0][k]), int(pa[1][k]), int(pa[2][k]), int(pa[3][k])) for k in range(max(n, 0))])
        return {"n": n, "viterbi": viterbi, "log_prob": lp.value, "samples": samples, "F": F}


class HostEmu:
    NCHAIN = 13

    def __init__(self, blob: bytes, simt32: bool = False):
        """simt32=True: the 32-lane flavour of the kernel source (the lane-group paths of the device) on 32 fibers per warp,
        tests/hostemu/simt32.h; default: one lane."""
        srcs = ["tests/hostemu/hostemu.cc", "augustus_b200/csrc/ghmm_model.cc"]
        hdrs = [os.path.join(ROOT, "augustus_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "augustus_b200", "csrc"))]
        hdrs.append(os.path.join(ROOT, "tests", "hostemu", "simt32.h"))
        newest = max(os.path.getmtime(p) for p in hdrs + [os.path.join(ROOT, s) for s in srcs])
        so = HOSTEMU_SO.replace("hostemu.so", "hostemu32.so") if simt32 else HOSTEMU_SO
        if not os.path.exists(so) or os.path.getmtime(so) < newest:
            subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC"] + (["-DAUGB_SIMT32"] if simt32 else []) + ["-o", so] + srcs, cwd=ROOT, check=True)
        self.lib = ctypes.CDLL(so)
        self.lib.hostemu_model_create.restype = ctypes.c_void_p
        self.lib.hostemu_model_create.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_int]
        self.lib.hostemu_decode.restype = ctypes.c_int
        self.lib.hostemu_decode.argtypes = ([ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
                                            + [ctypes.c_void_p] * 6 + [ctypes.c_int] + [ctypes.c_void_p] * 6)
        err = ctypes.create_string_buffer(512)
        self.m = self.lib.hostemu_model_create(blob, len(blob), err, 512)
        if not self.m:
            raise RuntimeError(err.value.decode())

    def sample(self, dna: str, nsamples=99, gc=None):
        """Forward fill + sampled paths with the host build of ghmm_sweep.h / ghmm_sample.h."""
        self.lib.hostemu_sample.restype = ctypes.c_int
        self.lib.hostemu_sample.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 7
        L = len(dna)
        cap = nsamples * (L // 8 + 64)
        sb, se = np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int32)
        st, tr = np.zeros(cap, dtype=np.uint8), np.zeros(cap, dtype=np.uint8)
        cnt, lp, status = np.zeros(nsamples, dtype=np.int32), np.zeros(nsamples), ctypes.c_int32()
        gci = None if gc is None else np.ascontiguousarray(gc, dtype=np.int32)
        rc = self.lib.hostemu_sample(self.m, dna.encode(), L, None if gci is None else gci.ctypes.data, nsamples, cap, sb.ctypes.data,
                                     se.ctypes.data, st.ctypes.data, tr.ctypes.data, cnt.ctypes.data, lp.ctypes.data, ctypes.byref(status))
        out, pos = [], 0
        for k in range(nsamples if rc >= 0 else 0):
            c = int(cnt[k])
            out.append({"log_prob": float(lp[k]), "states": [(int(st[pos + q]), int(sb[pos + q]), int(se[pos + q]), int(tr[pos + q])) for q in range(c)]})
            pos += c
        return {"status": status.value, "samples": out}

    def forward(self, dna: str, gc=None):
        """ln forward of every cell as a dense (L, S) matrix (-inf = zero) from the forward fill of the kernel source."""
        self.lib.hostemu_forward.restype = ctypes.c_int
        self.lib.hostemu_forward.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 6
        self.lib.hostemu_chain_state.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L = len(dna); S = self.lib.hostemu_statecount(ctypes.c_void_p(self.m))
        cap = 48 * L + 256
        col, st, F = np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int32), np.zeros(cap)
        nev, status = ctypes.c_int32(), ctypes.c_int32()
        chainF = np.zeros((L, self.NCHAIN))
        gci = None if gc is None else np.ascontiguousarray(gc, dtype=np.int32)
        self.lib.hostemu_forward(self.m, dna.encode(), L, None if gci is None else gci.ctypes.data, cap, col.ctypes.data, st.ctypes.data, F.ctypes.data,
                                 ctypes.byref(nev), chainF.ctypes.data, ctypes.byref(status))
        out = np.full((L, S), -np.inf)
        n = nev.value
        out[col[:n], st[:n]] = F[:n]
        for ch in range(self.NCHAIN):
            cs = self.lib.hostemu_chain_state(ctypes.c_void_p(self.m), ch)
            if cs >= 0:
                v = chainF[:, ch]; out[:, cs] = np.where(v < -1e300, -np.inf, v)
        return {"status": status.value, "F": out}

    def decode(self, dna: str, gc=None, want_cells=False, S=None):
        L = len(dna)
        S = S or self.lib.hostemu_statecount(ctypes.c_void_p(self.m))
        cap = L // 2 + 64
        eb, ee = np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int32)
        et, etr = np.zeros(cap, dtype=np.uint8), np.zeros(cap, dtype=np.uint8)
        lp, st, nev = ctypes.c_double(), ctypes.c_int32(), ctypes.c_int32()
        evcap = 8 * L + 256
        evc, evs = np.zeros(evcap, dtype=np.int32), np.zeros(evcap, dtype=np.int32)
        evV = np.zeros(evcap, dtype=np.int64)
        chV = np.zeros((L, self.NCHAIN), dtype=np.int64) if want_cells else None
        gco = np.zeros(L, dtype=np.uint8)
        gci = None if gc is None else np.ascontiguousarray(gc, dtype=np.int32)
        n = self.lib.hostemu_decode(self.m, dna.encode(), L, None if gci is None else gci.ctypes.data, cap,
                                    eb.ctypes.data, ee.ctypes.data, et.ctypes.data, etr.ctypes.data,
                                    ctypes.byref(lp), ctypes.byref(st), evcap, evc.ctypes.data, evs.ctypes.data,
                                    evV.ctypes.data, ctypes.byref(nev), None if chV is None else chV.ctypes.data, gco.ctypes.data)
        res = {"n": n, "status": st.value, "log_prob": lp.value, "gc": gco,
               "states": [(int(et[k]), int(eb[k]), int(ee[k]), int(etr[k])) for k in range(n)]}
        if want_cells:
            E = np.full((L, S), -(1 << 61), dtype=np.int64)
            ne = nev.value
            E[evc[:ne], evs[:ne]] = evV[:ne]
            for c in range(self.NCHAIN):
                s = self.lib.hostemu_chain_state(ctypes.c_void_p(self.m), c)
                if s >= 0:
                    E[:, s] = chV[:, c]
            res["cells"] = E
        return res


def blob_with_int(blob: bytes, name: str, value: int) -> bytes:
    """Copy of an AUGB2PAR blob with the int32 scalar `name` set to `value` (e.g. the sampling temperature)."""
    import struct
    b = bytearray(blob)
    n_entries = struct.unpack_from("<I", b, 12)[0]
    for i in range(n_entries):
        o = 16 + i * 96
        if bytes(b[o:o + 40]).split(b"\0")[0] == name.encode():
            off = struct.unpack_from("<Q", b, o + 80)[0]
            struct.pack_into("<i", b, off, value)
            return bytes(b)
    raise KeyError(name)


def hostemu_samples_at(emu, dna, nsamples, rand_pos):
    """Sampled paths of the host build with the rand() stream started `rand_pos` draws in; returns (paths, draws consumed)."""
    emu.lib.hostemu_sample_at.restype = ctypes.c_int
    emu.lib.hostemu_sample_at.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 7 + [ctypes.c_uint64, ctypes.c_void_p]
    L = len(dna); cap = nsamples * (L // 8 + 64)
    sb, se = np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int32)
    st, tr = np.zeros(cap, dtype=np.uint8), np.zeros(cap, dtype=np.uint8)
    cnt, lp, status, used = np.zeros(nsamples, dtype=np.int32), np.zeros(nsamples), ctypes.c_int32(), ctypes.c_int32()
    emu.lib.hostemu_sample_at(emu.m, dna.encode(), L, None, nsamples, cap, sb.ctypes.data, se.ctypes.data, st.ctypes.data, tr.ctypes.data,
                              cnt.ctypes.data, lp.ctypes.data, ctypes.byref(status), rand_pos, ctypes.byref(used))
    assert status.value == 0
    out, p = [], 0
    for k in range(nsamples):
        c = int(cnt[k]); out.append([(int(st[p + q]), int(sb[p + q]), int(se[p + q]), int(tr[p + q])) for q in range(c)]); p += c
    return out, used.value

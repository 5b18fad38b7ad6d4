"""ctypes binding of the C ABI declared in include/augb200.h.

Naming follows the reference seam: ``Decoder.viterbiAndForward`` + ``getViterbiPath``
stand for NAMGene::viterbiAndForward / getViterbiPath (src/namgene.cc:168, :432) on one window,
``decode_batch`` is the batched form the throughput configurations use.
"""
import ctypes
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import params as _params

_HERE = os.path.dirname(os.path.abspath(__file__))

# stateTypeIdentifiers, reference src/types.cc:150-171 (index = StateType value)
STATE_TYPE_NAMES = (
    "igenic single initial0 initial1 initial2 internal0 internal1 internal2 terminal "
    "lessD0 longdss0 equalD0 geometric0 longass0 lessD1 longdss1 equalD1 geometric1 longass1 "
    "lessD2 longdss2 equalD2 geometric2 longass2 "
    "utr5single utr5init utr5intron utr5intronvar utr5internal utr5term "
    "utr3single utr3init utr3intron utr3intronvar utr3internal utr3term "
    "rsingle rinitial rinternal0 rinternal1 rinternal2 rterminal0 rterminal1 rterminal2 "
    "rlessD0 rlongdss0 requalD0 rgeometric0 rlongass0 rlessD1 rlongdss1 requalD1 rgeometric1 rlongass1 "
    "rlessD2 rlongdss2 requalD2 rgeometric2 rlongass2 "
    "rutr5single rutr5init rutr5intron rutr5intronvar rutr5internal rutr5term "
    "rutr3single rutr3init rutr3intron rutr3intronvar rutr3internal rutr3term "
    "intron rintron exon ncsingle ncinit ncintron ncintronvar ncinternal ncterm "
    "rncsingle rncinit rncintron rncintronvar rncinternal rncterm").split()

ERR_NO_PATH, ERR_STUCK = 6, 7


class AugB200Error(RuntimeError):
    """Mirrors the reference's ProjectError for this path (include/types.hh:449-483)."""

    def __init__(self, code: int, msg: str):
        super().__init__("augb200 error %d: %s" % (code, msg))
        self.code = code


def library_path() -> str:
    return os.path.join(_HERE, "libaugb200.so")


class _Window(ctypes.Structure):
    _fields_ = [("dna", ctypes.c_char_p), ("length", ctypes.c_int32), ("gc_class", ctypes.POINTER(ctypes.c_int32))]


class _Path(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("status", ctypes.c_int32),
                ("begin", ctypes.POINTER(ctypes.c_int32)), ("end", ctypes.POINTER(ctypes.c_int32)),
                ("type", ctypes.POINTER(ctypes.c_uint8)), ("truncated", ctypes.POINTER(ctypes.c_uint8)),
                ("log_prob", ctypes.c_double)]


@dataclass
class State:            # include/gene.hh:101-245 (State)
    type: int
    begin: int
    end: int
    truncated: int

    @property
    def name(self) -> str:
        return STATE_TYPE_NAMES[self.type]


@dataclass
class StatePath:        # include/gene.hh (StatePath), condensed as gene.cc:977-1000
    states: List[State]
    log_prob: float
    status: int = 0

    def as_tuples(self):
        return [(s.type, s.begin, s.end, s.truncated) for s in self.states]


_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise AugB200Error(4, "CUDA extension %s is missing — build it with __graft_entry__.build(); "
                              "there is no CPU fallback" % path)
    lib = ctypes.CDLL(path)
    lib.augb200_model_create.restype = ctypes.c_int
    lib.augb200_model_create.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    lib.augb200_model_destroy.argtypes = [ctypes.c_void_p]
    lib.augb200_model_statecount.argtypes = [ctypes.c_void_p]
    lib.augb200_model_num_gc_classes.argtypes = [ctypes.c_void_p]
    for f in ("augb200_decode_batch",):
        getattr(lib, f).restype = ctypes.c_int
        getattr(lib, f).argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(_Window), ctypes.POINTER(_Path)]
    lib.augb200_decode_batch_sampling.restype = ctypes.c_int
    lib.augb200_decode_batch_sampling.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(_Window), ctypes.c_int32, ctypes.POINTER(_Path), ctypes.POINTER(_Path)]
    lib.augb200_decode_batch_multi.restype = ctypes.c_int
    lib.augb200_decode_batch_multi.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(_Window), ctypes.c_int32,
                                               ctypes.POINTER(_Path), ctypes.POINTER(_Path)]
    lib.augb200_decode.restype = ctypes.c_int
    lib.augb200_decode.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Window), ctypes.POINTER(_Path)]
    lib.augb200_stage_batch.restype = ctypes.c_int
    lib.augb200_stage_batch.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(_Window)]
    lib.augb200_run_staged.restype = ctypes.c_int
    lib.augb200_run_staged.argtypes = [ctypes.c_void_p]
    lib.augb200_fetch_staged.restype = ctypes.c_int
    lib.augb200_fetch_staged.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Path)]
    lib.augb200_model_stream.restype = ctypes.c_void_p
    lib.augb200_model_stream.argtypes = [ctypes.c_void_p]
    lib.augb200_last_launch_count.restype = ctypes.c_int64
    lib.augb200_last_launch_count.argtypes = [ctypes.c_void_p]
    lib.augb200_set_rand_position.restype = ctypes.c_int
    lib.augb200_set_rand_position.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
    lib.augb200_last_rand_consumed.restype = ctypes.c_int64
    lib.augb200_last_rand_consumed.argtypes = [ctypes.c_void_p]
    lib.augb200_last_sweep_ms.restype = ctypes.c_double
    lib.augb200_last_sweep_ms.argtypes = [ctypes.c_void_p]
    lib.augb200_result_store.restype = ctypes.c_int64
    lib.augb200_result_store.argtypes = [ctypes.c_void_p] + [ctypes.POINTER(ctypes.c_void_p)] * 4
    lib.augb200_sample_store.restype = ctypes.c_int64
    lib.augb200_sample_store.argtypes = [ctypes.c_void_p] + [ctypes.POINTER(ctypes.c_void_p)] * 4
    lib.augb200_sample_first_occurrence.restype = ctypes.c_int64
    lib.augb200_sample_first_occurrence.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
    lib.augb200_strerror.restype = ctypes.c_char_p
    lib.augb200_strerror.argtypes = [ctypes.c_int]
    lib.augb200_last_cuda_error.restype = ctypes.c_char_p
    _lib = lib
    return lib


class Decoder:
    """One model (species parameter set) on one GPU."""

    def __init__(self, blob: bytes, device: int = 0):
        self._lib = _load()
        self._h = ctypes.c_void_p()
        self._blob = blob
        rc = self._lib.augb200_model_create(blob, len(blob), device, ctypes.byref(self._h))
        if rc:
            raise AugB200Error(rc, "%s (%s)" % (self._lib.augb200_strerror(rc).decode(),
                                                self._lib.augb200_last_cuda_error().decode()))
        self._keep = None
        self._n_staged = 0
        self._last_path = None

    @classmethod
    def from_file(cls, path: str, device: int = 0) -> "Decoder":
        return cls(_params.load_bytes(path), device)

    def close(self):
        if self._h:
            self._lib.augb200_model_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def statecount(self) -> int:
        return self._lib.augb200_model_statecount(self._h)

    # ------------------------------------------------------------------ helpers
    def _windows(self, seqs: Sequence, gc: Optional[Sequence] = None):
        n = len(seqs)
        arr = (_Window * n)()
        keep = []
        for i, s in enumerate(seqs):
            b = s if isinstance(s, (bytes, bytearray)) else s.encode()
            keep.append(b)
            arr[i].dna = b
            arr[i].length = len(b)
            if gc is not None and gc[i] is not None:
                g = np.ascontiguousarray(gc[i], dtype=np.int32)
                keep.append(g)
                arr[i].gc_class = g.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        return arr, keep

    def _check(self, rc):
        if rc:
            raise AugB200Error(rc, "%s (%s)" % (self._lib.augb200_strerror(rc).decode(),
                                                self._lib.augb200_last_cuda_error().decode()))

    @staticmethod
    def _paths(out, n) -> List[StatePath]:
        res = []
        for i in range(n):
            p = out[i]
            k = p.n
            if k > 0:
                b = np.ctypeslib.as_array(p.begin, (k,)); e = np.ctypeslib.as_array(p.end, (k,))
                t = np.ctypeslib.as_array(p.type, (k,)); tr = np.ctypeslib.as_array(p.truncated, (k,))
                states = [State(int(t[j]), int(b[j]), int(e[j]), int(tr[j])) for j in range(k)]
            else:
                states = []
            res.append(StatePath(states, p.log_prob, p.status))
        return res

    # ------------------------------------------------------------------ the reference seam
    def decode_batch(self, seqs: Sequence, gc: Optional[Sequence] = None, raise_on_dp_error: bool = True) -> List[StatePath]:
        """viterbiAndForward + getViterbiPath for every window (host buffers in, host paths out)."""
        arr, keep = self._windows(seqs, gc)
        out = (_Path * len(seqs))()
        self._check(self._lib.augb200_decode_batch(self._h, len(seqs), arr, out))
        paths = self._paths(out, len(seqs))
        if raise_on_dp_error:
            for p in paths:
                if p.status:
                    raise AugB200Error(p.status, self._lib.augb200_strerror(p.status).decode())
        return paths

    def decode_batch_raw(self, seqs: Sequence, gc: Optional[Sequence] = None):
        """Same call as :meth:`decode_batch`, results as flat numpy arrays (no per-state Python objects):
        returns (n[i], status[i], log_prob[i], offset[i], begin, end, type, truncated)."""
        arr, keep = self._windows(seqs, gc)
        nw = len(seqs)
        out = (_Path * nw)()
        self._check(self._lib.augb200_decode_batch(self._h, nw, arr, out))
        return self._raw(out, nw)

    def _raw(self, out, nw, store=None):
        ptrs = [ctypes.c_void_p() for _ in range(4)]
        total = (store or self._lib.augb200_result_store)(self._h, *[ctypes.byref(p) for p in ptrs])
        rec = np.frombuffer(out, dtype=np.dtype([("n", "<i4"), ("status", "<i4"), ("begin", "<u8"), ("end", "<u8"),
                                                 ("type", "<u8"), ("trunc", "<u8"), ("log_prob", "<f8")]), count=nw)
        base = ptrs[0].value or 0
        offset = ((rec["begin"].astype(np.int64) - base) // 4) if total else np.zeros(nw, dtype=np.int64)

        def arr_of(p, ct, dt):
            if not total:
                return np.zeros(0, dtype=dt)
            return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ct)), (total,)).copy()
        return (rec["n"].copy(), rec["status"].copy(), rec["log_prob"].copy(), offset,
                arr_of(ptrs[0], ctypes.c_int32, np.int32), arr_of(ptrs[1], ctypes.c_int32, np.int32),
                arr_of(ptrs[2], ctypes.c_uint8, np.uint8), arr_of(ptrs[3], ctypes.c_uint8, np.uint8))

    def decode_batch_sampling(self, seqs: Sequence, nsample: int = 100, gc: Optional[Sequence] = None):
        """findGenes with --sample=nsample: returns (viterbi_paths, samples) where samples[i] is the list of the nsample-1
        posterior state paths of window i (NAMGene::getSampledPath, namgene.cc:367)."""
        arr, keep = self._windows(seqs, gc)
        nw, ns = len(seqs), nsample - 1
        out = (_Path * nw)()
        samp = (_Path * (nw * ns))()
        self._check(self._lib.augb200_decode_batch_sampling(self._h, nw, arr, nsample, out, samp))
        vit = self._paths(out, nw)
        allp = self._paths(samp, nw * ns)
        return vit, [allp[i * ns:(i + 1) * ns] for i in range(nw)]

    def decode_batch_sampling_raw(self, seqs: Sequence, nsample: int = 100, gc: Optional[Sequence] = None):
        """As decode_batch_sampling, results as flat numpy arrays: (viterbi_raw, samples_raw), each the tuple of _raw();
        sample k of window i is row i*(nsample-1)+k."""
        arr, keep = self._windows(seqs, gc)
        nw, ns = len(seqs), nsample - 1
        out = (_Path * nw)()
        samp = (_Path * (nw * ns))()
        self._check(self._lib.augb200_decode_batch_sampling(self._h, nw, arr, nsample, out, samp))
        return self._raw(out, nw), self._raw(samp, nw * ns, self._lib.augb200_sample_store)

    @staticmethod
    def decode_batch_multi(decoders: Sequence["Decoder"], seqs: Sequence, nsample: int = 0):
        """augb200_decode_batch_multi: window i is decoded by decoders[i mod len(decoders)] (one model per device, same blob), results in
        input order.  Returns the Viterbi paths, or (paths, samples) when nsample >= 2."""
        lib = decoders[0]._lib
        arr, keep = decoders[0]._windows(seqs)
        nw, ns = len(seqs), (nsample - 1 if nsample else 0)
        out = (_Path * nw)()
        samp = (_Path * max(1, nw * ns))()
        hs = (ctypes.c_void_p * len(decoders))(*[d._h for d in decoders])
        decoders[0]._check(lib.augb200_decode_batch_multi(hs, len(decoders), nw, arr, nsample, out, samp if ns else None))
        vit = Decoder._paths(out, nw)
        if not ns:
            return vit
        allp = Decoder._paths(samp, nw * ns)
        return vit, [allp[i * ns:(i + 1) * ns] for i in range(nw)]

    def sample_first_occurrence(self, n_windows: int, nsample: int):
        """(first, total_states): first[i, k] = first sample of window i with the same states as sample k (k itself for a new path) for the
        last sampling call; total_states counts duplicates, the sample store holds the unique paths only."""
        p = ctypes.c_void_p()
        total = self._lib.augb200_sample_first_occurrence(self._h, ctypes.byref(p))
        ns = nsample - 1
        arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_int32)), (n_windows * ns,)).copy() if n_windows * ns else np.zeros(0, np.int32)
        return arr.reshape(n_windows, ns), int(total)

    def set_rand_position(self, draws_consumed: int):
        """Where in the process-wide rand() stream (vitmatrix.cc:300) the next sampling call starts; 0 = a fresh process."""
        self._check(self._lib.augb200_set_rand_position(self._h, draws_consumed))

    @property
    def last_rand_consumed(self) -> int:
        """rand() draws the sampled paths of window 0 of the last sampling call took."""
        return self._lib.augb200_last_rand_consumed(self._h)

    def viterbiAndForward(self, dna, gc=None):
        """NAMGene::viterbiAndForward (namgene.cc:168) for one window; the path is kept for getViterbiPath."""
        self._last_path = self.decode_batch([dna], None if gc is None else [gc])[0]

    def getViterbiPath(self) -> StatePath:
        """NAMGene::getViterbiPath (namgene.cc:432)."""
        if self._last_path is None:
            raise AugB200Error(5, "viterbiAndForward has not been called")
        return self._last_path

    # ------------------------------------------------------------------ device-resident timing path
    def stage(self, seqs: Sequence, gc: Optional[Sequence] = None):
        arr, keep = self._windows(seqs, gc)
        self._check(self._lib.augb200_stage_batch(self._h, len(seqs), arr))
        self._n_staged = len(seqs)

    def run_staged(self):
        self._check(self._lib.augb200_run_staged(self._h))

    def fetch_staged(self) -> List[StatePath]:
        out = (_Path * self._n_staged)()
        self._check(self._lib.augb200_fetch_staged(self._h, out))
        return self._paths(out, self._n_staged)

    @property
    def stream(self) -> int:
        return self._lib.augb200_model_stream(self._h) or 0

    @property
    def last_launch_count(self) -> int:
        return self._lib.augb200_last_launch_count(self._h)

    @property
    def last_sweep_ms(self) -> float:
        return self._lib.augb200_last_sweep_ms(self._h)

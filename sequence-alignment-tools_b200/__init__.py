"""sequence-alignment-tools_b200 -- B200-native PrimerMatch database scan.

The product is the CUDA library `lib/libpm_b200.so` (C ABI: include/pm_b200.h)
and the C++ host code over it.  This module is the thin ctypes binding used by
the tests and bench.py; `primer_match()` mirrors the reference's main loop
(primer_match.cc:943-1268) on top of the C ABI so that parity tests read like
runs of the reference tool.

There is no CPU fallback: importing works anywhere, but every call that scans
needs the library and a CUDA device and raises otherwise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpm_b200.so")

PM_OK = 0
PM_E_ARG, PM_E_CUDA, PM_E_UNSUPPORTED, PM_E_NOMEM, PM_E_BOGUS = -1, -2, -3, -4, -5
PM_ALIGN_PRIMER_MATCH, PM_ALIGN_PCR_MATCH = 0, 1
PM_INDEX_AUTO, PM_INDEX_KEYWORD_TREE, PM_INDEX_SHIFT_AND, PM_INDEX_FILTER_BITVEC = 0, 3, 4, 5
PM_INDEX_EXACT_BASES, PM_INDEX_EXACT_HALVES = 9, 13
PM_SCAN_AUTO, PM_SCAN_PLANES, PM_SCAN_INDEX = 0, 1, 2
PM_MAX_CODES = 72
INT64_MAX = (1 << 63) - 1
EOS = 10

# alignment_code.h:5-36 -> pattern_alignment::alignment_string (pattern_alignment.h:124-166)
CODE_CHARS = {1: "|", 2: "+", 3: "*", 4: "^", 5: "v", 6: "!"}

# every symbol include/pm_b200.h declares
EXPORTS = ["pm_last_error", "pm_version", "pm_create", "pm_destroy", "pm_set_stream", "pm_add_pattern",
           "pm_init", "pm_set_text_host", "pm_set_text_device", "pm_set_text_sqz_host", "pm_set_text_sqn_host", "pm_sqz_bits", "pm_set_text_fasta_host", "pm_get_sequences", "pm_get_text", "pm_host_alloc", "pm_host_free", "pm_shard_overlap", "pm_find_patterns",
           "pm_find_patterns_device", "pm_scan_candidates", "pm_reset", "pm_get_stats",
           "pm_normalize_fasta", "pm_reverse_comp", "pm_bench_int_peak", "pm_set_scan_path", "pm_pcr_pairs", "pm_free", "pm_set_alignment_rules", "pm_set_alphabet_size", "pm_set_text_database", "pm_get_header", "pm_gather_text", "pm_last_candidates", "pm_hit_order"]


class PmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pm_b200 error %d: %s" % (code, msg))
        self.code = code


class Options(C.Structure):
    _fields_ = [("device", C.c_int), ("index", C.c_int), ("k", C.c_int), ("indels", C.c_int),
                ("wildcards", C.c_int), ("text_n", C.c_int), ("eos", C.c_ubyte)]


class Hit(C.Structure):
    _fields_ = [("start", C.c_int64), ("end", C.c_int64), ("first", C.c_int64), ("pattern_id", C.c_uint32),
                ("editdist", C.c_int32), ("value", C.c_int32), ("ncodes", C.c_uint32),
                ("codes", C.c_uint8 * PM_MAX_CODES)]


class Candidate(C.Structure):
    _fields_ = [("pos", C.c_int64), ("pattern_id", C.c_uint32), ("reserved", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("text_len", C.c_uint64), ("n_patterns", C.c_uint64), ("n_seeds", C.c_uint64),
                ("n_candidates", C.c_uint64), ("n_chains", C.c_uint64), ("n_hits", C.c_uint64),
                ("n_bogus", C.c_uint64), ("kernel_launches", C.c_uint64), ("scan_items", C.c_uint64),
                ("scan_warp_steps", C.c_uint64), ("scan_path", C.c_uint32), ("index_q", C.c_uint32),
                ("index_nd", C.c_uint32), ("index_pieces", C.c_uint32), ("index_lookups", C.c_uint64),
                ("index_entries", C.c_uint64), ("n_ambiguous", C.c_uint64), ("host_syncs", C.c_uint64), ("pack_ms", C.c_float), ("scan_ms", C.c_float), ("total_ms", C.c_float)]


class PcrOptions(C.Structure):
    _fields_ = [("k", C.c_int), ("indels", C.c_int), ("allorient", C.c_int), ("betweenlen", C.c_int),
                ("deviation", C.c_int), ("reserved", C.c_int), ("mindist", C.c_int64), ("maxdist", C.c_int64),
                ("sts_sizelb", C.c_void_p), ("sts_sizeub", C.c_void_p), ("cands", C.c_void_p), ("ncands", C.c_uint64)]


class PcrPair(C.Structure):
    _fields_ = [("first", C.c_uint64), ("second", C.c_uint64), ("amplicon_len", C.c_int64), ("pair", C.c_uint32),
                ("reserved", C.c_uint32)]


HIT_DTYPE = np.dtype([("start", "<i8"), ("end", "<i8"), ("first", "<i8"), ("pattern_id", "<u4"), ("editdist", "<i4"),
                      ("value", "<i4"), ("ncodes", "<u4"), ("codes", "u1", (PM_MAX_CODES,))])
CAND_DTYPE = np.dtype([("pos", "<i8"), ("pattern_id", "<u4"), ("reserved", "<u4")])
assert HIT_DTYPE.itemsize == C.sizeof(Hit)

_lib = None


def lib():
    """The C-ABI library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PmError(PM_E_CUDA, "%s is missing: run `python -m build` equivalent "
                          "(__graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        u8p = C.POINTER(C.c_uint8)
        L.pm_last_error.restype = C.c_char_p
        L.pm_version.restype = C.c_char_p
        L.pm_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Options)]
        L.pm_destroy.argtypes = [C.c_void_p]
        L.pm_destroy.restype = None
        L.pm_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.pm_set_scan_path.argtypes = [C.c_void_p, C.c_int]
        L.pm_set_alignment_rules.argtypes = [C.c_void_p, C.c_int]
        L.pm_set_alphabet_size.argtypes = [C.c_void_p, C.c_int]
        L.pm_add_pattern.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_int, C.c_int]
        L.pm_init.argtypes = [C.c_void_p]
        L.pm_set_text_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64]
        L.pm_set_text_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64]
        L.pm_set_text_sqz_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_char_p, C.c_uint32,
                                           C.c_uint64, C.c_int64]
        L.pm_sqz_bits.argtypes = [C.c_uint32]
        L.pm_set_text_sqn_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_char_p, C.c_uint32, C.c_int64]
        L.pm_set_text_fasta_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
        i64pp = C.POINTER(C.POINTER(C.c_int64))
        L.pm_get_sequences.argtypes = [C.c_void_p, i64pp, i64pp, i64pp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.pm_get_text.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        L.pm_gather_text.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.pm_set_text_database.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        L.pm_get_header.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.pm_host_alloc.argtypes = [C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]
        L.pm_host_free.argtypes = [C.c_void_p]
        L.pm_shard_overlap.argtypes = [C.c_void_p]
        L.pm_find_patterns.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.POINTER(Hit)),
                                       C.POINTER(C.c_uint64)]
        L.pm_find_patterns_device.argtypes = [C.c_void_p, C.c_int64, C.c_int64,
                                              C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.pm_last_candidates.argtypes = [C.c_void_p, C.POINTER(C.POINTER(Candidate)), C.POINTER(C.c_uint64)]
        L.pm_hit_order.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        L.pm_scan_candidates.argtypes = [C.c_void_p, C.c_int64, C.c_int64,
                                         C.POINTER(C.POINTER(Candidate)), C.POINTER(C.c_uint64)]
        L.pm_reset.argtypes = [C.c_void_p]
        L.pm_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.pm_normalize_fasta.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_ubyte, C.c_void_p,
                                         C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_uint64, C.POINTER(C.c_uint64)]
        L.pm_reverse_comp.argtypes = [C.c_char_p, C.c_char_p]
        L.pm_pcr_pairs.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64,
                                   C.POINTER(PcrOptions), C.POINTER(C.POINTER(PcrPair)), C.POINTER(C.c_uint64)]
        L.pm_free.argtypes = [C.c_void_p]
        L.pm_free.restype = None
        L.pm_bench_int_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _check(rc):
    if rc != PM_OK:
        raise PmError(rc, lib().pm_last_error().decode("latin-1"))


def reverse_comp(s: str) -> str:
    buf = C.create_string_buffer(len(s) + 1)
    _check(lib().pm_reverse_comp(s.encode("latin-1"), buf))
    return buf.value.decode("latin-1")


@dataclass
class Stream:
    """Normalised character stream of a FASTA file (fasta_io.t:478-629)."""
    chars: np.ndarray
    seq_key: np.ndarray
    headers: list

    def locate(self, p: int):
        """get_header_data / get_seq_pos (fasta_io.t:163-202)."""
        i = int(np.searchsorted(self.seq_key, p - 1, side="right")) - 1
        if i < 0:
            return 0, 0
        return i + 1, int(self.seq_key[i])


def normalize_fasta(raw: bytes, uc: bool = False, eos: int = EOS) -> Stream:
    n = len(raw)
    rawa = np.frombuffer(raw, dtype=np.uint8) if n else np.zeros(1, np.uint8)
    out = np.zeros(n + 4, np.uint8)
    maxseq = raw.count(b">") + 2
    key = np.zeros(maxseq, np.int64)
    off = np.zeros(maxseq, np.int64)
    ln = np.zeros(maxseq, np.int64)
    olen, nseq = C.c_uint64(0), C.c_uint64(0)
    _check(lib().pm_normalize_fasta(rawa.ctypes.data, n, int(uc), eos, out.ctypes.data, C.byref(olen),
                                    key.ctypes.data, off.ctypes.data, ln.ctypes.data, maxseq,
                                    C.byref(nseq)))
    ns = nseq.value
    headers = [raw[off[i]:off[i] + max(ln[i], 0)].decode("latin-1") for i in range(ns)]
    return Stream(out[:olen.value].copy(), key[:ns].copy(), headers)


class Scanner:
    """One PatternMatch index on one GPU (pattern_match.h:83-157)."""

    def __init__(self, k=0, indels=True, wildcards=False, text_n=False, index=PM_INDEX_AUTO,
                 device=0, eos=EOS, scan_path=PM_SCAN_AUTO):
        self._h = C.c_void_p()
        opt = Options(device, index, k, int(indels), int(wildcards), int(text_n), eos)
        _check(lib().pm_create(C.byref(self._h), C.byref(opt)))
        if scan_path != PM_SCAN_AUTO:
            _check(lib().pm_set_scan_path(self._h, scan_path))
        self.k = k
        self.indels = bool(indels)
        self._keep = None

    def close(self):
        if self._h:
            lib().pm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_pattern(self, pat: str, pid: int, exact_start_bases=0, exact_end_bases=0):
        _check(lib().pm_add_pattern(self._h, pat.encode("latin-1"), pid, exact_start_bases,
                                    exact_end_bases))

    def init(self):
        _check(lib().pm_init(self._h))

    def set_alignment_rules(self, rules: int):
        """PM_ALIGN_PCR_MATCH: the alignment pcr_match makes of every hit (differs from
        primer_match's for k = 0 with wildcards)."""
        _check(lib().pm_set_alignment_rules(self._h, rules))

    def set_alphabet_size(self, size: int):
        """CharacterProducer::size() of the database (select.cc:120-123); before init()."""
        _check(lib().pm_set_alphabet_size(self._h, size))

    def set_stream(self, cuda_stream: int):
        _check(lib().pm_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_text(self, chars, offset: int = 0):
        """chars: numpy uint8 array (host) or an object with .data_ptr()/.is_cuda (torch)."""
        if hasattr(chars, "data_ptr"):
            n = chars.numel()
            if chars.is_cuda:
                self._keep = chars
                _check(lib().pm_set_text_device(self._h, C.c_void_p(chars.data_ptr()), n, offset))
            else:
                self._keep = chars     # page-locked memory is read by DMA after the call returns
                _check(lib().pm_set_text_host(self._h, C.c_void_p(chars.data_ptr()), n, offset))
            return
        a = np.ascontiguousarray(chars, np.uint8)
        self._keep = a
        _check(lib().pm_set_text_host(self._h, C.c_void_p(a.ctypes.data), a.size, offset))

    def set_text_sqz(self, sqz, table: bytes, n_chars: int = 0, offset: int = 0):
        """The text as the .sqz bit stream of compress_seq -z (host memory: numpy uint8
        array or a CPU torch tensor, pinned or not) with its .tbz character table."""
        if hasattr(sqz, "data_ptr"):
            assert not sqz.is_cuda
            self._keep = sqz
            ptr, nb = sqz.data_ptr(), sqz.numel()
        else:
            a = np.ascontiguousarray(sqz, np.uint8)
            self._keep = a
            ptr, nb = a.ctypes.data, a.size
        _check(lib().pm_set_text_sqz_host(self._h, C.c_void_p(ptr), nb, bytes(table), len(table), n_chars, offset))

    def set_text_sqn(self, sqn, table: bytes, offset: int = 0):
        """The text as the .sqn file of compress_seq -n (one .tbl index per character)."""
        a = np.ascontiguousarray(sqn, np.uint8)
        self._keep = a
        _check(lib().pm_set_text_sqn_host(self._h, C.c_void_p(a.ctypes.data), a.size, bytes(table), len(table), offset))

    def set_text_fasta(self, raw, uc: bool = False, fetch_chars: bool = False) -> "Stream":
        """The text as the bytes of a FASTA file (bytes, numpy uint8 array or CPU torch tensor,
        pinned or not): normalised on the device (pm_fasta.cuh).  Returns the Stream with the
        sequence table; its .chars are fetched from the device only when asked for."""
        if hasattr(raw, "data_ptr"):
            assert not raw.is_cuda
            self._keep = raw
            ptr, nb = raw.data_ptr(), raw.numel()
            view = memoryview(raw.numpy())
        else:
            a = np.frombuffer(raw, np.uint8) if isinstance(raw, (bytes, bytearray, memoryview)) else np.ascontiguousarray(raw, np.uint8)
            self._keep = a
            ptr, nb = (a.ctypes.data if a.size else 0), a.size
            view = memoryview(a)
        _check(lib().pm_set_text_fasta_host(self._h, C.c_void_p(ptr), nb, int(uc)))
        kp, op, lp = C.POINTER(C.c_int64)(), C.POINTER(C.c_int64)(), C.POINTER(C.c_int64)()
        ns, sl = C.c_uint64(0), C.c_uint64(0)
        _check(lib().pm_get_sequences(self._h, C.byref(kp), C.byref(op), C.byref(lp), C.byref(ns), C.byref(sl)))
        n = ns.value
        key = np.array([kp[i] for i in range(n)], np.int64)
        headers = [bytes(view[op[i]:op[i] + max(lp[i], 0)]).decode("latin-1") for i in range(n)]
        chars = self.get_text(0, sl.value) if fetch_chars else None
        return Stream(chars, key, headers)

    def set_text_database(self, database: str, dbind: int = 0, uc: bool = False) -> "Stream":
        """The text read from a database on disk the way primer_match opens it for -D `dbind`
        (pick_fasta_file): the FASTA file, or compress_seq's .seq / .sqn / .sqz, through the library's
        own reader (ring of page-locked chunk buffers).  Returns the Stream with the sequence table."""
        _check(lib().pm_set_text_database(self._h, database.encode(), dbind, int(uc)))
        kp, op, lp = C.POINTER(C.c_int64)(), C.POINTER(C.c_int64)(), C.POINTER(C.c_int64)()
        ns, sl = C.c_uint64(0), C.c_uint64(0)
        _check(lib().pm_get_sequences(self._h, C.byref(kp), C.byref(op), C.byref(lp), C.byref(ns), C.byref(sl)))
        n = ns.value
        key = np.array([kp[i] for i in range(n)], np.int64)
        headers = []
        for i in range(n):
            tp, tl = C.c_void_p(), C.c_uint64(0)
            _check(lib().pm_get_header(self._h, i, C.byref(tp), C.byref(tl)))
            headers.append(C.string_at(tp, tl.value).decode("latin-1") if tl.value else "")
        return Stream(None, key, headers)

    def get_text(self, begin: int, length: int) -> np.ndarray:
        out = np.zeros(length, np.uint8)
        if length:
            _check(lib().pm_get_text(self._h, begin, length, C.c_void_p(out.ctypes.data)))
        return out

    def shard_overlap(self) -> int:
        return lib().pm_shard_overlap(self._h)

    def find_patterns(self, begin=0, end=INT64_MAX, bogus_ok=False) -> np.ndarray:
        """bogus_ok: PM_E_BOGUS (a hit whose own alignment has more than k errors, on which
        primer_match stops) still delivers the hits; pcr_match goes on with them."""
        hp = C.POINTER(Hit)()
        n = C.c_uint64(0)
        rc = lib().pm_find_patterns(self._h, begin, end, C.byref(hp), C.byref(n))
        if not (bogus_ok and rc == PM_E_BOGUS):
            _check(rc)
        if n.value == 0:
            return np.zeros(0, HIT_DTYPE)
        buf = C.string_at(hp, n.value * C.sizeof(Hit))
        return np.frombuffer(buf, dtype=HIT_DTYPE).copy()

    def find_patterns_device(self, begin=0, end=INT64_MAX):
        dp = C.c_void_p()
        n = C.c_uint64(0)
        _check(lib().pm_find_patterns_device(self._h, begin, end, C.byref(dp), C.byref(n)))
        return dp.value, n.value

    def scan_candidates(self, begin=0, end=INT64_MAX) -> np.ndarray:
        cp = C.POINTER(Candidate)()
        n = C.c_uint64(0)
        _check(lib().pm_scan_candidates(self._h, begin, end, C.byref(cp), C.byref(n)))
        out = np.zeros((n.value, 2), np.int64)
        for i in range(n.value):
            out[i, 0] = cp[i].pos
            out[i, 1] = cp[i].pattern_id
        return out

    def last_candidates(self):
        """pm_last_candidates: the raw candidates behind the hits of the last find_patterns as a
        structured array (pos, pattern_id, reserved), or None for exact_halves / exact_bases."""
        cp = C.POINTER(Candidate)()
        n = C.c_uint64(0)
        rc = lib().pm_last_candidates(self._h, C.byref(cp), C.byref(n))
        if rc == PM_E_UNSUPPORTED:
            return None
        _check(rc)
        if not n.value:
            return np.zeros(0, CAND_DTYPE)
        return np.ctypeslib.as_array(C.cast(cp, C.POINTER(C.c_uint8)), (n.value * C.sizeof(Candidate),)).view(CAND_DTYPE).copy()

    def stats(self) -> dict:
        s = Stats()
        _check(lib().pm_get_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in Stats._fields_}


def int_peak(device=0) -> float:
    """Measured LOP3 throughput of the device, 1e12 thread-level ops/s."""
    v = C.c_double(0)
    _check(lib().pm_bench_int_peak(device, C.byref(v)))
    return v.value


# ALU-pipe instructions per level-1 lookup of pair_scan_kernel (SASS of the inner
# loop: SHF window, LOP3 key, SHF word index, LEA address, SHF.L.W rotate, SHF.L.W
# collect); the hash multiply is an IMAD on the FMA pipe and the LDS is on the LSU.
PAIR_ALU_OPS_PER_LOOKUP = 6


def int_pipe_report(sc: "Scanner", scan_ms: float, device=0) -> dict:
    """Integer-pipe roofline of the last scan of `sc`: algorithmic logic
    operations per second against the measured LOP3 peak.
      PLANES  counter / AND-chain updates only (one LOP3 per counter bit per word)
      INDEX   level-1 lookups x PAIR_ALU_OPS_PER_LOOKUP"""
    st = sc.stats()
    peak = int_peak(device)
    if st["scan_path"] == PM_SCAN_INDEX:
        useful = st["index_lookups"] * PAIR_ALU_OPS_PER_LOOKUP
        ach = useful / (scan_ms * 1e-3) / 1e12
        return {"unit": "Tlop/s", "achieved": ach, "peak": peak, "frac": ach / peak if peak else None,
                "useful_lops_per_launch": useful, "lookups": st["index_lookups"],
                "alu_ops_per_lookup": PAIR_ALU_OPS_PER_LOOKUP, "level2_records": st["scan_warp_steps"],
                "peak_source": "pm_bench_int_peak (LOP3 microbenchmark, same run)"}
    # LOP3 per word per character: one per counter bit (pm_scan.cuh Counter<K>), one AND for seeds
    levels = {0: 1, 1: 2, 2: 2, 3: 3}[sc.k] if (not sc.indels or sc.k == 0) else 1
    useful = st["scan_warp_steps"] * 8 * 32 * levels
    ach = useful / (scan_ms * 1e-3) / 1e12
    return {"unit": "Tlop/s", "achieved": ach, "peak": peak, "frac": ach / peak if peak else None,
            "useful_lops_per_launch": useful, "warp_steps": st["scan_warp_steps"],
            "peak_source": "pm_bench_int_peak (LOP3 microbenchmark, same run)"}


class HostBuffer:
    """Page-locked host memory from pm_host_alloc, as a numpy uint8 view (.array)."""

    def __init__(self, nbytes: int, write_combined: bool = False):
        p = C.c_void_p()
        _check(lib().pm_host_alloc(nbytes, int(write_combined), C.byref(p)))
        self.ptr, self.nbytes = p.value, nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(nbytes, 1)).from_address(self.ptr))[:nbytes]

    def data_ptr(self):
        return self.ptr

    def numel(self):
        return self.nbytes

    is_cuda = False

    def numpy(self):
        return self.array

    def free(self):
        if self.ptr:
            lib().pm_host_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def sqz_table(chars: np.ndarray, eos: int = 10, dna_opt: bool = True) -> bytes:
    """The .tbz character table compress_seq writes for a stream (compress_seq.cc:683-722):
    the observed characters (the eos character counts as observed) in byte order, with
    A, C, G, T swapped to the front when -D (the default) is on."""
    obs = np.zeros(256, bool)
    obs[np.unique(np.asarray(chars, np.uint8))] = True
    obs[eos] = True
    order = list(range(256))
    if dna_opt:
        for i, ch in enumerate(b"ACGT"):
            order[i], order[ch] = ch, i
    return bytes(o for o in order if obs[o])


def sqz_encode(chars: np.ndarray, table: bytes, eos: int = 10) -> np.ndarray:
    """compress_seq -z (compress_seq.cc:745-870): the stream as `bits` bits per character,
    most significant bit first, padded with eos characters to a whole buffer of
    lcm(bits, 8) bytes.  Vectorised numpy; the oracle holds the independent restatement."""
    chars = np.asarray(chars, np.uint8)
    inv = np.full(256, -1, np.int16)
    inv[np.frombuffer(table, np.uint8)] = np.arange(len(table), dtype=np.int16)
    bits = 1
    while (1 << bits) < len(table):
        bits += 1
    bufsize = (bits * 8 // int(np.gcd(bits, 8))) // 8 * 8      # bytes per writer buffer
    per_buf = bufsize * 8 // bits
    n = len(chars)
    nbuf = max(1, -(-n // per_buf))
    if inv[eos] < 0 or (n and int(inv[chars].min()) < 0):
        raise ValueError("character outside the table")
    codes = np.full(nbuf * per_buf, inv[eos], np.uint8)
    codes[:n] = inv[chars]
    b = ((codes[:, None] >> np.arange(bits - 1, -1, -1, dtype=np.uint8)) & 1).astype(np.uint8)
    return np.packbits(b.reshape(-1))


def sqz_encode_torch(chars, table: bytes, eos: int = 10):
    """sqz_encode for a (possibly device-resident) torch uint8 tensor; returns a uint8
    tensor on the same device.  Used to prepare large synthetic databases."""
    import torch
    dev = chars.device
    bits = 1
    while (1 << bits) < len(table):
        bits += 1
    inv = torch.full((256,), 255, dtype=torch.uint8, device=dev)
    inv[torch.tensor(list(table), dtype=torch.long, device=dev)] = torch.arange(len(table), dtype=torch.uint8, device=dev)
    bufsize = (bits * 8 // int(np.gcd(bits, 8))) // 8 * 8
    per_buf = bufsize * 8 // bits
    n = chars.numel()
    nbuf = max(1, -(-n // per_buf))
    out = torch.empty(nbuf * bufsize, dtype=torch.uint8, device=dev)
    shifts = torch.arange(bits - 1, -1, -1, dtype=torch.uint8, device=dev)
    weights = torch.tensor([128, 64, 32, 16, 8, 4, 2, 1], dtype=torch.uint8, device=dev)
    step = per_buf * (1 << 21)
    for a in range(0, nbuf * per_buf, step):
        b = min(nbuf * per_buf, a + step)
        codes = torch.full((b - a,), int(inv[eos]), dtype=torch.uint8, device=dev)
        if a < n:
            codes[:min(b, n) - a] = inv[chars[a:min(b, n)].long()]
        bm = (codes[:, None] >> shifts) & 1
        out[a * bits // 8:b * bits // 8] = (bm.reshape(-1, 8) * weights).sum(dim=1, dtype=torch.uint8)
    return out


@dataclass(frozen=True, order=True)
class HitRecord:
    """One output record of primer_match, the fields alignformat() receives
    (primer_match.cc:1174-1200)."""
    seq_index: int
    seq_start: int
    seq_end: int
    strand: str
    pattern_index: int
    editdist: int
    stream_start: int
    stream_end: int
    alignment: str


def pattern_array(primers, rev_comp: bool):
    """primer_match.cc:943-1080: ids 1..n forward, n+1..2n reverse complements."""
    pats = list(primers)
    if rev_comp:
        pats += [reverse_comp(p) for p in primers]
    return pats


def pattern_constraints(primers, rev_comp: bool, stlen=0, edlen=0, fplen=0, tplen=0):
    """primer_match.cc:990-1060: exact_start_bases / exact_end_bases of every pattern of
    pattern_array() from the -s / -e / -5 / -3 options (negative = the `~n` form: all but n
    characters).  For a reverse complement the 5' and 3' ends swap."""
    def one(ln, five, three):
        f = max(stlen if stlen > 0 else 0, five if five > 0 else 0)
        if edlen < 0:
            f = max(f, ln + edlen)
        if three < 0:
            f = max(f, ln + three)
        s_ = max(edlen if edlen > 0 else 0, three if three > 0 else 0)
        if stlen < 0:
            s_ = max(s_, ln + stlen)
        if five < 0:
            s_ = max(s_, ln + five)
        return f, s_
    cons = [one(len(p), fplen, tplen) for p in primers]
    if rev_comp:
        cons += [one(len(p), tplen, fplen) for p in primers]
    return [c[0] for c in cons], [c[1] for c in cons]


def hits_to_records(stream: Stream, hits: np.ndarray, n_primers: int):
    """primer_match.cc:1156-1200: coordinates handed to alignformat()."""
    recs = []
    for h in hits:
        start, end, pid = int(h["start"]), int(h["end"]), int(h["pattern_id"])
        idx, key = stream.locate(end)
        length = end - start + 1          # pattern_alignment::length()
        spe = end - key
        sps = spe - length + 1
        rc = pid > n_primers
        astr = "".join(CODE_CHARS.get(int(c), " ") for c in h["codes"][:int(h["ncodes"])])
        recs.append(HitRecord(idx, sps, spe, "R" if rc else "F", pid - (n_primers if rc else 0),
                              int(h["editdist"]), end - length + 1, end, astr))
    return sorted(recs)


def primer_match(stream: Stream, primers, k=0, indels=True, rev_comp=False, wc=False, tn=False,
                 eos=EOS, device=0, index=PM_INDEX_AUTO, scan_path=PM_SCAN_AUTO, sqz=None,
                 stlen=0, edlen=0, fplen=0, tplen=0):
    """The reference's main loop on the GPU path: returns sorted HitRecords.
    index: a -N selection; PM_INDEX_AUTO follows pick_pattern_index (exact_halves for k == 1).
    sqz = (sqz bytes, tbz table, n_chars): load the text from compress_seq's compressed
    form instead of stream.chars (stream still supplies the sequence index)."""
    pats = pattern_array(primers, rev_comp)
    sc = Scanner(k=k, indels=indels, wildcards=wc, text_n=tn, index=index, device=device, eos=eos,
                 scan_path=scan_path)
    try:
        esb, eeb = pattern_constraints(primers, rev_comp, stlen, edlen, fplen, tplen)   # -s / -e / -5 / -3
        for i, p in enumerate(pats):
            sc.add_pattern(p, i + 1, esb[i], eeb[i])
        sc.init()
        if sqz is not None and len(sqz) == 2:
            sc.set_text_sqn(*sqz)        # (sqn bytes, tbl table)
        elif sqz is not None:
            sc.set_text_sqz(*sqz)
        else:
            sc.set_text(stream.chars)
        hits = sc.find_patterns()
        return hits_to_records(stream, hits, len(primers))
    finally:
        sc.close()


# ---------------------------------------------------------------------------
# pcr_match: primer pairs, hits paired into amplicons (pcr_match.cc)
@dataclass(frozen=True, order=True)
class PcrRecord:
    """One output record of pcr_match: the positional arguments of its alignformat()
    (pcr_match.cc:1179-1196)."""
    seq_index: int        # %f
    start: int            # %>s
    start1: int           # %<s
    end: int              # %>e
    end1: int             # %<e
    stream_start: int     # %>S
    stream_start1: int    # %<S
    stream_end: int       # %>E
    stream_end1: int      # %<E
    pair: int             # %i
    editdist: int         # %>d
    editdist1: int        # %<d
    strand: str           # %>r
    strand1: str          # %<r
    orient: str           # %r
    alignment: str        # %>A
    alignment1: str       # %<A
    ncount: int           # %N
    amplicon: str         # %@


def pcr_pattern_array(primers, rev_comp: bool):
    """pcr_match.cc:772-784, 827, 868: pairs (1,2), (3,4) ...; -r reverse complements the second
    primer of every pair first; ids 1..n as given, n+1..2n their reverse complements."""
    pr = [reverse_comp(p) if (rev_comp and i % 2 == 1) else p for i, p in enumerate(primers)]
    return pr + [reverse_comp(p) for p in pr]


def pcr_pairs(hits: np.ndarray, n_primers: int, patlen, seq_key, k=0, indels=True, allorient=False,
              betweenlen=False, mindist=0, maxdist=2000, sts_sizes=None, deviation=-1, cands=None):
    """pm_pcr_pairs: [(index of the first hit, index of its partner, amplicon length, pair)] in
    the order pcr_match prints them."""
    hits = np.ascontiguousarray(hits, HIT_DTYPE)
    pl = np.ascontiguousarray(patlen, np.uint32)
    keys = np.ascontiguousarray(seq_key, np.int64)
    o = PcrOptions(k=k, indels=int(indels), allorient=int(allorient), betweenlen=int(betweenlen),
                   deviation=deviation if sts_sizes is not None else -1, mindist=mindist, maxdist=maxdist)
    if sts_sizes is not None:
        lb = np.array([s[0] for s in sts_sizes], np.uint64)
        ub = np.array([s[1] for s in sts_sizes], np.uint64)
        o.sts_sizelb, o.sts_sizeub = lb.ctypes.data, ub.ctypes.data
    if cands is not None and len(cands):
        cands = np.ascontiguousarray(cands, CAND_DTYPE)
        o.cands, o.ncands = cands.ctypes.data, len(cands)
    pp = C.POINTER(PcrPair)()
    n = C.c_uint64(0)
    _check(lib().pm_pcr_pairs(C.c_void_p(hits.ctypes.data), len(hits), n_primers, C.c_void_p(pl.ctypes.data),
                              C.c_void_p(keys.ctypes.data), len(keys), C.byref(o), C.byref(pp), C.byref(n)))
    out = [(pp[i].first, pp[i].second, pp[i].amplicon_len, pp[i].pair) for i in range(n.value)]
    lib().pm_free(pp)
    return out


def pcr_match(stream: Stream, primers, k=0, indels=True, rev_comp=False, wc=False, tn=False, eos=EOS, device=0,
              index=PM_INDEX_AUTO, stlen=0, edlen=0, fplen=0, tplen=0, mindist=0, maxdist=2000,
              betweenlen=False, allorient=False, sts_sizes=None, deviation=-1):
    """pcr_match's main loop on the GPU path: the scan and the alignments of every hit
    (pm_find_patterns), the pairing (pm_pcr_pairs), and the record each pair prints
    (pcr_match.cc:1116-1196).  sts_sizes: [(sizelb, sizeub)] per pair for primers from a UniSTS
    file (-S, which implies -r); deviation = -d.  Returns PcrRecords in output order."""
    if sts_sizes is not None:
        rev_comp = True
    n = len(primers)
    pats = pcr_pattern_array(primers, rev_comp)
    sc = Scanner(k=k, indels=indels, wildcards=wc, text_n=tn, index=index, device=device, eos=eos)
    try:
        for i in range(n):
            # -5 and -3 swap for the second primer of a pair (pcr_match.cc:803-808)
            fp, tp = (fplen, tplen) if i % 2 == 0 else (tplen, fplen)
            esb, eeb = pattern_constraints([pats[i]], True, stlen, edlen, fp, tp)
            sc.add_pattern(pats[i], i + 1, esb[0], eeb[0])
            sc.add_pattern(pats[i + n], i + n + 1, esb[1], eeb[1])
        sc.init()
        sc.set_alignment_rules(PM_ALIGN_PCR_MATCH)
        sc.set_text(stream.chars)
        hits = sc.find_patterns(bogus_ok=True)
        cands = sc.last_candidates()
        patlen = [0] + [len(p) for p in pats]
        T = stream.chars
        recs = []
        for a, b, amplicon_len, pind in pcr_pairs(hits, n, patlen, stream.seq_key, k, indels, allorient, betweenlen,
                                                  mindist, maxdist, sts_sizes, deviation, cands):
            h, h1 = hits[a], hits[b]
            pid, pid1 = int(h["pattern_id"]), int(h1["pattern_id"])
            pe, pe1 = int(h["end"]), int(h1["end"])
            ps, ps1 = int(h["start"]), int(h1["start"])
            idx, key = stream.locate(pe)
            _, key1 = stream.locate(pe1)
            spe, spe1 = pe - key, pe1 - key1
            sps, sps1 = spe - (pe - ps), spe1 - (pe1 - ps1)
            rc, rc1 = pid > n, pid1 > n
            ind, ind1 = pid - (n if rc else 0), pid1 - (n if rc1 else 0)
            if rev_comp:                                     # pcr_match.cc:1143-1149
                if ind % 2 == 0:
                    rc = not rc
                elif ind1 % 2 == 0:
                    rc1 = not rc1
            amp = bytes(T[ps:ps + amplicon_len]).decode("latin-1") if amplicon_len > 0 else ""
            astr = "".join(CODE_CHARS.get(int(c), " ") for c in h["codes"][:int(h["ncodes"])])
            astr1 = "".join(CODE_CHARS.get(int(c), " ") for c in h1["codes"][:int(h1["ncodes"])])
            recs.append(PcrRecord(idx, sps, sps1, spe, spe1, ps, ps1, pe, pe1, pind, int(h["editdist"]),
                                  int(h1["editdist"]), "R" if rc else "F", "R" if rc1 else "F",
                                  "F" if ind < ind1 else "R", astr, astr1, amp.count("N") + amp.count("n"), amp))
        return recs
    finally:
        sc.close()

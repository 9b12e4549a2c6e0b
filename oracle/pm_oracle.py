"""ctypes front-end of oracle/libpm_oracle.so -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never from the product package.

`primer_match()` mirrors the reference's main loop (primer_match.cc:1100-1268):
pattern array with reverse complements appended (lines 943-1080), the index
chosen by pick_pattern_index (select.cc:39-298; the oracle covers selections
2/3 keyword_tree, 4 shift_and and 5 filter_bitvec), one final alignment per
returned hit (lines 1132-1153) and the coordinates handed to alignformat
(lines 1156-1200).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

EOS = 10  # '\n', primer_match.cc:99

# alignment_code.h:5-36 -> the characters of pattern_alignment::alignment_string
# (pattern_alignment.h:124-166)
CODE_CHARS = {1: "|", 2: "+", 3: "*", 4: "^", 5: "v", 6: "!"}


class _Hit(C.Structure):
    _fields_ = [("pos", C.c_int64), ("pid", C.c_int32), ("val", C.c_int32)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libpm_oracle.so")
    src = os.path.join(_HERE, "pm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libpm_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p = C.POINTER(C.c_uint8)
        L.pmo_normalize_fasta.restype = C.c_long
        L.pmo_normalize_fasta.argtypes = [u8p, C.c_long, C.c_int, C.c_uint8, u8p,
                                          C.POINTER(C.c_long), C.POINTER(C.c_long),
                                          C.POINTER(C.c_long), C.c_long, C.POINTER(C.c_long)]
        L.pmo_wm_scan.restype = C.c_long
        L.pmo_wm_scan.argtypes = [u8p, C.c_long, C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_uint8, C.POINTER(_Hit), C.c_long]
        L.pmo_exact_scan.restype = C.c_long
        L.pmo_exact_scan.argtypes = [u8p, C.c_long, C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int,
                                     C.POINTER(_Hit), C.c_long]
        L.pmo_editdist_align.restype = C.c_int
        L.pmo_editdist_align.argtypes = [u8p, C.c_long, C.c_char_p, C.c_int64, C.c_int64, C.c_int,
                                         C.c_uint8, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                         C.POINTER(C.c_int), C.POINTER(C.c_int), u8p,
                                         C.POINTER(C.c_int)]
        L.pmo_exact_halves.restype = C.c_long
        L.pmo_exact_halves.argtypes = [u8p, C.c_long, C.c_int, C.POINTER(C.c_char_p), C.POINTER(_Hit), C.c_long,
                                       C.c_int, C.c_uint8, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                       C.POINTER(C.c_int), C.POINTER(_Hit), C.c_long]
        L.pmo_exact_bases.restype = C.c_long
        L.pmo_exact_bases.argtypes = [u8p, C.c_long, C.c_int, C.POINTER(C.c_char_p), C.POINTER(_Hit), C.c_long,
                                      C.c_int, C.c_uint8, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                      C.POINTER(C.c_int), C.POINTER(_Hit), C.c_long]
        L.pmo_filter_chains.restype = C.c_long
        L.pmo_filter_chains.argtypes = [u8p, C.c_long, C.c_int, C.POINTER(C.c_char_p),
                                        C.POINTER(_Hit), C.c_long, C.c_int, C.c_uint8, C.c_int,
                                        C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(_Hit), C.c_long]
        L.pmo_exact_wc_align.restype = C.c_int
        L.pmo_exact_wc_align.argtypes = [u8p, C.c_long, C.c_char_p, C.c_int64, C.c_int, u8p]
        L.pmo_scan_match.restype = C.c_int
        L.pmo_scan_match.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        _LIB = L
    return _LIB


def _u8(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _pats(pats):
    arr = (C.c_char_p * len(pats))(*[p.encode("latin-1") for p in pats])
    return arr


# --- util.cc:319-348 iupac_revcomp, util.cc:374-381 reverse_comp -------------
_RC = {i: i for i in range(256)}
for a, b in ["AT", "CG", "GC", "TA", "UA", "MK", "RY", "WW", "SS", "YR", "KM", "VB", "HD", "DH", "BV"]:
    _RC[ord(a)] = ord(b)
    _RC[ord(a.lower())] = ord(b.lower())


def reverse_comp(s: str) -> str:
    return "".join(chr(_RC[ord(c)]) for c in reversed(s))


@dataclass
class Stream:
    """Normalised character stream of a FASTA file (fasta_io.t:478-629)."""
    chars: np.ndarray      # uint8, length L
    seq_key: np.ndarray    # int64: stream index of the first base of each sequence
    headers: list          # header strings

    def locate(self, p: int):
        """get_header_data / get_seq_pos (fasta_io.t:163-202): the sequence whose
        key is the last one <= p-1.  Returns (ordinal (1-based), key)."""
        i = int(np.searchsorted(self.seq_key, p - 1, side="right")) - 1
        if i < 0:
            return 0, 0
        return i + 1, int(self.seq_key[i])


def normalize_fasta(raw: bytes, uc: bool = False, eos: int = EOS) -> Stream:
    L = lib()
    n = len(raw)
    rawa = np.frombuffer(raw, dtype=np.uint8) if n else np.zeros(0, np.uint8)
    rawa = np.ascontiguousarray(rawa)
    out = np.zeros(n + 4, np.uint8)
    maxseq = raw.count(b">") + 2
    key = (C.c_long * maxseq)()
    off = (C.c_long * maxseq)()
    ln = (C.c_long * maxseq)()
    nseq = C.c_long(0)
    slen = L.pmo_normalize_fasta(_u8(rawa) if n else None, n, int(uc), eos, _u8(out), key, off, ln,
                                 maxseq, C.byref(nseq))
    ns = nseq.value
    headers = [raw[off[i]:off[i] + max(ln[i], 0)].decode("latin-1") for i in range(ns)]
    return Stream(out[:slen].copy(), np.array([key[i] for i in range(ns)], np.int64), headers)


def wm_scan(chars: np.ndarray, pats, k: int, indels: bool, wc=False, tn=False, eos=EOS):
    """shift_and_inexact candidates as an (n,3) int64 array [pos, pid, level]."""
    L = lib()
    chars = np.ascontiguousarray(chars, np.uint8)
    cap = 1 << 16
    while True:
        buf = (_Hit * cap)()
        n = L.pmo_wm_scan(_u8(chars), len(chars), len(pats), _pats(pats), k, int(indels), int(wc),
                          int(tn), eos, buf, cap)
        if n < 0:
            raise ValueError("pmo_wm_scan: unsupported pattern length or k")
        if n <= cap:
            break
        cap = n
    return np.array([(buf[i].pos, buf[i].pid, buf[i].val) for i in range(n)], np.int64).reshape(-1, 3)


def exact_scan(chars: np.ndarray, pats, wc=False, tn=False):
    L = lib()
    chars = np.ascontiguousarray(chars, np.uint8)
    cap = 1 << 16
    while True:
        buf = (_Hit * cap)()
        n = L.pmo_exact_scan(_u8(chars), len(chars), len(pats), _pats(pats), int(wc), int(tn), buf, cap)
        if n <= cap:
            break
        cap = n
    return np.array([(buf[i].pos, buf[i].pid, 0) for i in range(n)], np.int64).reshape(-1, 3)


@dataclass
class Alignment:
    ok: bool
    start: int
    end: int
    val: int
    editdist: int   # -1 == MAXINT (constraint violation)
    codes: bytes

    @property
    def string(self) -> str:
        return "".join(CODE_CHARS.get(c, " ") for c in self.codes)


def editdist_align(chars, pat: str, e1: int, e2: int, k: int, indels: bool, wc=False, tn=False,
                   lconst=0, rconst=0, eos=EOS) -> Alignment:
    L = lib()
    chars = np.ascontiguousarray(chars, np.uint8)
    s, e = C.c_int64(0), C.c_int64(0)
    val, ed, nc = C.c_int(0), C.c_int(0), C.c_int(0)
    codes = np.zeros(len(pat) * 2 + 4 * k + 64 + int(e2 - e1), np.uint8)
    rc = L.pmo_editdist_align(_u8(chars), len(chars), pat.encode("latin-1"), e1, e2, k, eos, int(wc),
                              int(tn), int(indels), lconst, rconst, C.byref(s), C.byref(e),
                              C.byref(val), C.byref(ed), _u8(codes), C.byref(nc))
    if rc < 0:
        raise AssertionError("editdist_alignment traceback left the band (reference assert(0))")
    return Alignment(bool(rc), s.value, e.value, val.value, ed.value, bytes(codes[:nc.value]))


def filter_chains(chars, pats, cands: np.ndarray, k: int, indels: bool, wc=False, tn=False,
                  lconst=None, rconst=None, eos=EOS):
    """filter_bitvec: (n,3) candidates (sorted by position) -> (h,3) [end,pid,val]."""
    L = lib()
    chars = np.ascontiguousarray(chars, np.uint8)
    n = len(cands)
    cb = (_Hit * max(n, 1))()
    for i in range(n):
        cb[i].pos, cb[i].pid, cb[i].val = int(cands[i, 0]), int(cands[i, 1]), int(cands[i, 2])
    ob = (_Hit * max(n, 1))()
    lc = (C.c_int * len(pats))(*lconst) if lconst is not None else None
    rc = (C.c_int * len(pats))(*rconst) if rconst is not None else None
    h = L.pmo_filter_chains(_u8(chars), len(chars), len(pats), _pats(pats), cb, n, k, eos, int(wc),
                            int(tn), int(indels), lc, rc, ob, max(n, 1))
    return np.array([(ob[i].pos, ob[i].pid, ob[i].val) for i in range(h)], np.int64).reshape(-1, 3)


def exact_wc_align(chars, pat: str, end: int, tn=False):
    L = lib()
    chars = np.ascontiguousarray(chars, np.uint8)
    codes = np.zeros(len(pat), np.uint8)
    sub = L.pmo_exact_wc_align(_u8(chars), len(chars), pat.encode("latin-1"), end, int(tn), _u8(codes))
    return sub, bytes(codes)


@dataclass(frozen=True, order=True)
class HitRecord:
    """One output record of primer_match, the fields alignformat() receives
    (primer_match.cc:1174-1200)."""
    seq_index: int      # %f  ordinal of the FASTA entry (1-based)
    seq_start: int      # %s
    seq_end: int        # %e
    strand: str         # %r  "F" / "R"
    pattern_index: int  # %i  index of the primer (1-based, before rev-comp)
    editdist: int       # %d
    stream_start: int   # %S
    stream_end: int     # %E
    alignment: str      # %A


def pattern_array(primers, rev_comp: bool):
    """primer_match.cc:943-1080: ids 1..n forward, n+1..2n reverse complements."""
    pats = list(primers)
    if rev_comp:
        pats += [reverse_comp(p) for p in primers]
    return pats


def exact_halves(chars, pats, k: int, indels: bool, wc=False, tn=False, eos=EOS, lconst=None, rconst=None):
    """exact_halves::find_patterns: (end, pattern id, value) of every reported hit."""
    L = lib()
    chars = np.ascontiguousarray(chars, np.uint8)
    halves = []
    for p in pats:
        halves += [p[:len(p) // 2], p[len(p) // 2:]]      # inner ids 2i-1, 2i (exact_halves.cc:201-224)
    hh = exact_scan(chars, halves, wc, tn)
    n = len(hh)
    hb = (_Hit * max(n, 1))()
    for i, (pos, pid, _) in enumerate(hh):
        hb[i].pos, hb[i].pid, hb[i].val = int(pos), int(pid), 0
    out = (_Hit * max(n, 1))()
    lc = (C.c_int * len(pats))(*lconst) if lconst is not None else None
    rc = (C.c_int * len(pats))(*rconst) if rconst is not None else None
    m = L.pmo_exact_halves(_u8(chars), len(chars), len(pats), _pats(pats), hb, n, k, eos, int(wc), int(tn),
                           int(indels), lc, rc, out, max(n, 1))
    return np.array([[out[i].pos, out[i].pid, out[i].val] for i in range(m)], np.int64).reshape(-1, 3)


def pattern_constraints(primers, rev_comp: bool, stlen=0, edlen=0, fplen=0, tplen=0):
    """exact_start_bases / exact_end_bases of every pattern of pattern_array(): the -s / -e / -5 / -3
    options of primer_match (a negative value is the `~n` form: all but n characters),
    primer_match.cc:990-1060.  For a reverse complement 5' and 3' swap ends."""
    first, second = [], []
    for p in primers:
        ln = len(p)
        f = stlen if stlen > 0 else 0
        f = max(f, fplen)
        if edlen < 0:
            f = max(f, ln + edlen)
        if tplen < 0:
            f = max(f, ln + tplen)
        s_ = edlen if edlen > 0 else 0
        s_ = max(s_, tplen)
        if stlen < 0:
            s_ = max(s_, ln + stlen)
        if fplen < 0:
            s_ = max(s_, ln + fplen)
        first.append(f)
        second.append(s_)
    if rev_comp:
        for p in primers:
            ln = len(p)
            f = stlen if stlen > 0 else 0
            f = max(f, tplen)
            if edlen < 0:
                f = max(f, ln + edlen)
            if fplen < 0:
                f = max(f, ln + fplen)
            s_ = edlen if edlen > 0 else 0
            s_ = max(s_, fplen)
            if stlen < 0:
                s_ = max(s_, ln + stlen)
            if tplen < 0:
                s_ = max(s_, ln + tplen)
            first.append(f)
            second.append(s_)
    return first, second


def exact_bases(chars, pats, k: int, indels: bool, lconst, rconst, wc=False, tn=False, eos=EOS):
    """exact_bases::find_patterns: (end, pattern id, value) of every hit."""
    L = lib()
    chars = np.ascontiguousarray(chars, np.uint8)
    parts = [p[:a] if a >= b else p[len(p) - b:] for p, a, b in zip(pats, lconst, rconst)]
    hh = exact_scan(chars, parts, wc, tn)
    n = len(hh)
    hb = (_Hit * max(n, 1))()
    for i, (pos, pid, _) in enumerate(hh):
        hb[i].pos, hb[i].pid, hb[i].val = int(pos), int(pid), 0
    out = (_Hit * max(n, 1))()
    lc = (C.c_int * len(pats))(*lconst)
    rc = (C.c_int * len(pats))(*rconst)
    m = L.pmo_exact_bases(_u8(chars), len(chars), len(pats), _pats(pats), hb, n, k, eos, int(wc), int(tn),
                          int(indels), lc, rc, out, max(n, 1))
    return np.array([[out[i].pos, out[i].pid, out[i].val] for i in range(m)], np.int64).reshape(-1, 3)


def auto_index(primers, k: int, lconst=None, rconst=None) -> str:
    """pick_pattern_index (select.cc:32-141) for a streamed FASTA database (CharacterProducer::size()
    == 256), no -x seed length and a build with -DNOPRIMEGEN.  primers = the pattern array
    (reverse complements included when -r)."""
    if k == 0:
        return "exact"
    n = len(primers)
    lconst = lconst or [0] * n
    rconst = rconst or [0] * n
    ex = [max(a, b) for a, b in zip(lconst, rconst)]
    min_length = min(len(p) for p in primers)
    cumdiff = sum(e - len(p) // 2 for e, p in zip(ex, primers))
    cumbooldiff = sum(1 for e, p in zip(ex, primers) if e - len(p) // 2 >= 0)
    if k == 1 and min_length >= 8 and (cumbooldiff <= 0 or cumdiff <= 0):
        return "exact_halves"
    if min(ex) >= 6:
        return "exact_bases"
    return "filter_bitvec"


def aligned_hits(stream: Stream, pats, lconst, rconst, k=0, indels=True, wc=False, tn=False, eos=EOS,
                 index="filter_bitvec", bogus_ok=False, exact_align=True, keep_bogus=False):
    """PatternMatch::find_patterns of the selected index followed, per hit, by the alignment
    both command line tools make at the hit's position (primer_match.cc:1132-1153,
    pcr_match.cc:1096-1113): [(pattern id, start, end, editdist, alignment string)] in the
    order the index reports the hits.  end == the hit's key.  A hit whose alignment at its own
    position fails ("bogus") makes primer_match abort; pcr_match just never pairs it
    (pcr_match.cc:1114-1115), which bogus_ok=True stands for: such hits are left out.
    exact_align: primer_match's exact_alignment / exact_wc_alignment for k == 0; pcr_match
    makes an editdist_alignment there too (pcr_match.cc:1096-1100)."""
    T = stream.chars
    out = []
    if index == "auto":
        index = auto_index(pats, k, lconst, rconst)
    if k == 0:
        hits = exact_scan(T, pats, wc, tn)
    elif index == "exact_halves":
        hits = exact_halves(T, pats, k, indels, wc, tn, eos, lconst, rconst)
    elif index == "exact_bases":
        hits = exact_bases(T, pats, k, indels, lconst, rconst, wc, tn, eos)
    else:
        cands = wm_scan(T, pats, k, indels, wc, tn, eos)
        if len(cands) > 16:
            # filter_bitvec.cc:91: l.normalize() -- std::sort by position on the candidates as
            # shift_and_inexact reports them (position, then pattern): equal positions come out
            # in the order introsort leaves them, and chains are taken in that order
            cands = np.array(std_sort_by_key([tuple(r) for r in cands.tolist()], lambda r: r[0]), np.int64).reshape(-1, 3)
        hits = filter_chains(T, pats, cands, k, indels, wc, tn, lconst, rconst, eos)
    for pos, pid, _ in hits:
        pos, pid = int(pos), int(pid)
        pat = pats[pid - 1]
        if k == 0 and exact_align:
            start, end, ed = pos - len(pat), pos, 0
            if wc:
                sub, codes = exact_wc_align(T, pat, pos, tn)
                ed = sub
            else:
                codes = bytes([1] * len(pat))
            if ed > 0:
                raise AssertionError("Bogus hit (primer_match.cc:1249)")
            astr = "".join(CODE_CHARS[c] for c in codes)
        else:
            al = editdist_align(T, pat, pos, pos, k, indels, wc, tn, lconst[pid - 1], rconst[pid - 1], eos)
            if al.editdist < 0 or al.editdist > k:
                if bogus_ok:
                    if keep_bogus:      # pcr_match: stays in the list (and in its sort), never pairs
                        out.append((pid, pos, pos, -1, ""))
                    continue
                raise AssertionError("Bogus hit (primer_match.cc:1249)")
            start, end, ed, astr = al.start, al.end, al.editdist, al.string
        out.append((pid, start, end, ed, astr))
    return out


def primer_match(stream: Stream, primers, k=0, indels=True, rev_comp=False, wc=False, tn=False,
                 eos=EOS, index="filter_bitvec", stlen=0, edlen=0, fplen=0, tplen=0):
    """Whole path, as the reference runs it with `-N 5` for k>0 (index="filter_bitvec"), with its
    exact_halves / exact_bases indexes (index="exact_halves" / "exact_bases"; index="auto" follows
    pick_pattern_index), and its automatic choice for k==0 (keyword_tree, or shift_and under
    -w/-W).  stlen / edlen / fplen / tplen = the -s / -e / -5 / -3 options.
    Returns a sorted list of HitRecord."""
    n = len(primers)
    pats = pattern_array(primers, rev_comp)
    lconst, rconst = pattern_constraints(primers, rev_comp, stlen, edlen, fplen, tplen)
    recs = []
    for pid, start, end, ed, astr in aligned_hits(stream, pats, lconst, rconst, k, indels, wc, tn, eos, index):
        # primer_match.cc:1158-1179; pattern_alignment::length() is end-start+1
        idx, key = stream.locate(end)
        spe = end - key
        sps = spe - (end - start + 1) + 1
        rc = pid > n
        recs.append(HitRecord(idx, sps, spe, "R" if rc else "F", pid - (n if rc else 0), ed,
                              end - (end - start + 1) + 1, end, astr))
    return sorted(recs)


# ---------------------------------------------------------------------------
def std_sort_by_key(items, key):
    """std::sort of libstdc++ (bits/stl_algo.h: introsort with median-of-three pivots down to
    ranges of 16, then one insertion sort) on a comparison that looks at the key alone
    (sortedvector.t:38-44, :502-504).  The reference leaves hits with equal positions in whatever
    order this leaves them, and pcr_match's pairing depends on that order (a hit only pairs with
    partners that come later), so the order is part of the reference's answer."""
    v = list(items)
    n = len(v)
    if n < 2:
        return v
    k = [key(x) for x in v]

    def swap(i, j):
        v[i], v[j] = v[j], v[i]
        k[i], k[j] = k[j], k[i]

    def linear_insert(last):           # __unguarded_linear_insert
        val, kv = v[last], k[last]
        nxt = last - 1
        while kv < k[nxt]:
            v[last], k[last] = v[nxt], k[nxt]
            last = nxt
            nxt -= 1
        v[last], k[last] = val, kv

    def insertion_sort(first, last):   # __insertion_sort
        for i in range(first + 1, last):
            if k[i] < k[first]:
                val, kv = v[i], k[i]
                v[first + 1:i + 1] = v[first:i]
                k[first + 1:i + 1] = k[first:i]
                v[first], k[first] = val, kv
            else:
                linear_insert(i)

    def heap_sort(first, last):        # __partial_sort(first, last, last): make_heap + sort_heap
        def adjust(hole, length, val, kv):
            top = hole
            child = hole
            while child < (length - 1) // 2:
                child = 2 * (child + 1)
                if k[first + child] < k[first + child - 1]:
                    child -= 1
                v[first + hole], k[first + hole] = v[first + child], k[first + child]
                hole = child
            if (length & 1) == 0 and child == (length - 2) // 2:
                child = 2 * (child + 1)
                v[first + hole], k[first + hole] = v[first + child - 1], k[first + child - 1]
                hole = child - 1
            parent = (hole - 1) // 2
            while hole > top and k[first + parent] < kv:
                v[first + hole], k[first + hole] = v[first + parent], k[first + parent]
                hole = parent
                parent = (hole - 1) // 2
            v[first + hole], k[first + hole] = val, kv
        length = last - first
        if length >= 2:
            parent = (length - 2) // 2
            while True:
                adjust(parent, length, v[first + parent], k[first + parent])
                if parent == 0:
                    break
                parent -= 1
        while last - first > 1:
            last -= 1
            val, kv = v[last], k[last]
            v[last], k[last] = v[first], k[first]
            adjust(0, last - first, val, kv)

    def introsort(first, last, depth):
        while last - first > 16:
            if depth == 0:
                heap_sort(first, last)
                return
            depth -= 1
            mid = first + (last - first) // 2
            a, b, c = first + 1, mid, last - 1       # __move_median_to_first(first, a, b, c)
            if k[a] < k[b]:
                if k[b] < k[c]:
                    swap(first, b)
                elif k[a] < k[c]:
                    swap(first, c)
                else:
                    swap(first, a)
            elif k[a] < k[c]:
                swap(first, a)
            elif k[b] < k[c]:
                swap(first, c)
            else:
                swap(first, b)
            lo, hi = first + 1, last                 # __unguarded_partition(first + 1, last, first)
            while True:
                while k[lo] < k[first]:
                    lo += 1
                hi -= 1
                while k[first] < k[hi]:
                    hi -= 1
                if not lo < hi:
                    break
                swap(lo, hi)
                lo += 1
            introsort(lo, last, depth)
            last = lo

    introsort(0, n, 2 * (n.bit_length() - 1))
    if n > 16:                                       # __final_insertion_sort
        insertion_sort(0, 16)
        for i in range(16, n):
            linear_insert(i)
    else:
        insertion_sort(0, n)
    return v


# pcr_match: the second caller of the scan -- primer PAIRS, hits paired into amplicons
def pcr_pattern_array(primers, rev_comp: bool):
    """pcr_match.cc:772-784, 827, 868: pairs are (1,2), (3,4) ...; with -r the second primer of
    every pair is reverse complemented first; ids 1..n as given, n+1..2n their reverse
    complements (always both strands)."""
    pr = list(primers)
    if rev_comp:
        pr = [reverse_comp(p) if i % 2 == 1 else p for i, p in enumerate(pr)]
    return pr + [reverse_comp(p) for p in pr]


def pcr_pattern_constraints(pats_fwd, stlen=0, edlen=0, fplen=0, tplen=0):
    """pcr_match.cc:803-909: primer_match's rules, -5 and -3 swapped for the second primer of a pair."""
    n = len(pats_fwd)
    first, second = [0] * (2 * n), [0] * (2 * n)
    for i, p in enumerate(pats_fwd):
        fp, tp = (fplen, tplen) if i % 2 == 0 else (tplen, fplen)
        f, s_ = pattern_constraints([p], True, stlen, edlen, fp, tp)
        first[i], first[i + n], second[i], second[i + n] = f[0], f[1], s_[0], s_[1]
    return first, second


U64 = 1 << 64


@dataclass(frozen=True, order=True)
class PcrRecord:
    """One output record of pcr_match: the arguments its alignformat() receives
    (pcr_match.cc:1179-1196) that do not follow from the others."""
    seq_index: int        # %f
    start: int            # %>s  sps
    start1: int           # %<s  sps1
    end: int              # %>e  spe
    end1: int             # %<e  spe1
    stream_start: int     # %>S  ps
    stream_start1: int    # %<S  ps1
    stream_end: int       # %>E  pe
    stream_end1: int      # %<E  pe1
    pair: int             # %i   pind
    editdist: int         # %>d
    editdist1: int        # %<d
    strand: str           # %>r
    strand1: str          # %<r
    orient: str           # %r   "F" when ppo (ind < ind1)
    alignment: str        # %>A
    alignment1: str       # %<A
    ncount: int           # %N
    amplicon: str         # %@


def pcr_match(stream: Stream, primers, k=0, indels=True, rev_comp=False, wc=False, tn=False, eos=EOS,
              index="auto", stlen=0, edlen=0, fplen=0, tplen=0, mindist=0, maxdist=2000,
              betweenlen=False, allorient=False, sts_sizes=None, deviation=-1):
    """pcr_match.cc:921-1240 with every hit in one batch (what the reference does on any
    database whose hits it does not have to hand over in several find_patterns batches):
    hits in position order; a hit looks for hits of its partner pattern(s) that are still
    unprocessed -- later in the order -- and end inside its stretch window; both alignments
    and the amplicon length decide.  sts_sizes: [(sizelb, sizeub)] per pair when the primers
    came from a UniSTS file (-S, which implies -r); deviation = -d.  Returns the records in
    output order."""
    if sts_sizes is not None:
        rev_comp = True
    n = len(primers)
    pats = pcr_pattern_array(primers, rev_comp)
    lconst, rconst = pcr_pattern_constraints(pats[:n], stlen, edlen, fplen, tplen)
    hits = aligned_hits(stream, pats, lconst, rconst, k, indels, wc, tn, eos, index, bogus_ok=True,
                        exact_align=False, keep_bogus=True)
    hits = std_sort_by_key(hits, lambda h: h[2])    # l.normalize() (pcr_match.cc:958): std::sort on the key alone
    bypid = {}
    for i, h in enumerate(hits):
        bypid.setdefault(h[0], []).append(i)
    done = [False] * len(hits)
    T = stream.chars
    err = k if indels else 1
    out = []
    for i, (pid, start, pos, ed, astr) in enumerate(hits):
        if ed < 0:                                      # its own alignment fails for every partner (pcr_match.cc:1114)
            done[i] = True
            continue
        pid1 = pid2 = 0                                 # pcr_match.cc:968-997
        if pid <= n and pid % 2 == 1:
            pid1 = pid + 1
        elif pid > n and (pid - n) % 2 == 0:
            pid1 = pid - 1
        if allorient:
            if pid <= n:
                if pid % 2 == 1:
                    pid2 = pid + n + 1
                else:
                    pid1, pid2 = pid - 1, pid + n - 1
            else:
                if pid % 2 == 0:
                    pid2 = pid - n - 1
                else:
                    pid1, pid2 = pid + 1, pid - n + 1
        smax, smin = maxdist, mindist
        if betweenlen:
            plen = len(pats[pid1 - 1]) if pid1 else 0
            if pid2 and len(pats[pid2 - 1]) > plen:
                plen = len(pats[pid2 - 1])
            smax += plen + len(pats[pid - 1])
        if sts_sizes is not None and deviation >= 0:            # pcr_match.cc:1032-1039, unsigned compares
            lb, ub = sts_sizes[(pid - (n if pid > n else 0) + 1) // 2 - 1]
            if (smax % U64) > (ub + deviation) % U64:
                smax = ub + deviation
            if (smin % U64) < (lb - deviation) % U64:
                smin = lb - deviation
        smax += pos - len(pats[pid - 1]) + err
        smin += pos - len(pats[pid - 1]) - err
        partners = []
        for q in (pid1, pid2):
            if not q:
                continue
            for j in bypid.get(q, ()):
                if hits[j][2] < smin:
                    continue
                if not done[j] and hits[j][2] > smax:
                    break
                if not done[j] and hits[j][3] >= 0:
                    partners.append(j)
        for j in partners:
            pidb, startb, posb, edb, astrb = hits[j]
            idx, key = stream.locate(pos)
            idxb, keyb = stream.locate(posb)
            spe, spe1 = pos - key, posb - keyb
            length, length1 = pos - start + 1, posb - startb + 1      # pattern_alignment::length()
            sps, sps1 = spe - length + 1, spe1 - length1 + 1
            ps, ps1 = pos - length + 1, posb - length1 + 1
            rc, rc1 = pid > n, pidb > n
            ind, ind1 = pid - (n if rc else 0), pidb - (n if rc1 else 0)
            pind = (ind if ind < ind1 else ind1) // 2 + 1
            if rev_comp:
                if ind % 2 == 0:
                    rc = not rc
                elif ind1 % 2 == 0:
                    rc1 = not rc1
            amplicon_len = (ps1 - pos) if betweenlen else (posb - ps)
            # FastaFile::is_subseq(ps, pe1), fasta_io.t:203-213
            same = stream.locate(ps + 1)[0] == stream.locate(posb)[0] and stream.locate(ps + 1)[0] > 0
            if not (same and mindist <= amplicon_len <= maxdist):
                continue
            if sts_sizes is not None and deviation >= 0:        # pcr_match.cc:1163-1165
                lb, ub = sts_sizes[pind - 1]
                if not ((amplicon_len + deviation) % U64 >= lb and amplicon_len <= ub + deviation):
                    continue
            amp = bytes(T[ps:ps + amplicon_len]).decode("latin-1")
            out.append(PcrRecord(idx, sps, sps1, spe, spe1, ps, ps1, pos, posb, pind, ed, edb,
                                 "R" if rc else "F", "R" if rc1 else "F", "F" if ind < ind1 else "R",
                                 astr, astrb, amp.count("N") + amp.count("n"), amp))
        done[i] = True
    return out


PCR_REF_BIN = os.path.join(_HERE, "_ref", "pcr_match")
PCR_REF_FORMAT = (r"%f\t%>s\t%<s\t%>e\t%<e\t%>S\t%<S\t%>E\t%<E\t%i\t%>d\t%<d\t%>r\t%<r\t%r\t%>A\t%<A"
                  r"\t%N\t%@\n")


def pcr_cli_args(fasta_path, primers, k=0, indels=True, rev_comp=False, wc=False, tn=False, select=0,
                 stlen=0, edlen=0, fplen=0, tplen=0, mindist=0, maxdist=2000, betweenlen=False,
                 allorient=False, sts_path=None, deviation=-1):
    """Command line of pcr_match (the reference's and the drop-in's) for these parameters;
    sts_path: the primers are read from that UniSTS file (-S) instead of -p."""
    args = ["-i", fasta_path] + (["-S", sts_path] if sts_path else ["-p", " ".join(primers)])
    if deviation >= 0:
        args += ["-d", str(deviation)]
    if k > 0:
        args += ["-k" if indels else "-K", str(k)]
    if rev_comp:
        args.append("-r")
    if wc:
        args.append("-W" if tn else "-w")
    if select:
        args += ["-N", str(select)]
    for opt, v in (("-s", stlen), ("-e", edlen), ("-5", fplen), ("-3", tplen)):
        if v:
            args += [opt, ("~%d" % -v) if v < 0 else str(v)]
    if mindist:
        args += ["-m", str(mindist)]
    if maxdist != 2000:
        args += ["-M", str(maxdist)]
    if betweenlen:
        args.append("-b")
    if allorient:
        args.append("-a")
    return args


def parse_pcr_records(text: str):
    recs = []
    for line in text.splitlines():
        f = line.split("\t")
        recs.append(PcrRecord(*[int(x) for x in f[:12]], f[12], f[13], f[14], f[15], f[16], int(f[17]), f[18]))
    return recs


def run_reference_pcr(fasta_path: str, primers, fmt=PCR_REF_FORMAT, timeout=600, **params):
    """Run the unmodified reference pcr_match; returns its stdout (fmt=None: its default format)."""
    args = [PCR_REF_BIN] + pcr_cli_args(fasta_path, primers, **params) + (["-A", fmt] if fmt else [])
    p = subprocess.run(args, capture_output=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError("reference failed rc=%d: %s" % (p.returncode, p.stderr.decode()[-500:]))
    return p.stdout.decode("latin-1")


# ---------------------------------------------------------------------------
# compress_seq's compressed database (.tbz table + .sqz bit stream) and the
# Compressed<> reader.  Plain loops, small inputs only.
def sqz_table(chars, eos: int = EOS, dna_opt: bool = True) -> bytes:
    """compress_seq.cc:683-722: observed characters (eos always) in byte order,
    A C G T swapped to the front with -D true."""
    obs = [False] * 256
    obs[eos] = True
    for ch in bytes(bytearray(chars)):
        obs[ch] = True
    order = list(range(256))
    if dna_opt:
        for i, ch in enumerate(b"ACGT"):
            order[i] = ch
            order[ch] = i
    return bytes(order[i] for i in range(256) if obs[order[i]])


def sqz_bits(table_size: int) -> int:
    """char_io.t:58-61 / compress_seq.cc:766-769."""
    bits = 1
    while (1 << bits) < table_size:
        bits += 1
    return bits


def sqz_encode(chars, table: bytes, eos: int = EOS) -> bytes:
    """compress_seq.cc:745-870: codes of `bits` bits, most significant bit first, buffers of
    lcm(bits, 8) bytes; the last buffer is filled up with eos characters."""
    inv = {ch: i for i, ch in enumerate(table)}
    bits = sqz_bits(len(table))
    lcm = bits * 8 // math.gcd(bits, 8)
    bufsize = lcm // 8 * 8
    out = bytearray()
    buf = bytearray(bufsize)
    bytepos = bitpos = 0
    data = list(bytes(bytearray(chars)))
    i = 0
    partial = False
    while i < len(data) or partial:
        partial = True
        ch = data[i] if i < len(data) else eos
        i += 1
        code = inv[ch]
        for b in range(bits - 1, -1, -1):
            if (code >> b) & 1:
                buf[bytepos] |= 0x80 >> bitpos
            bitpos += 1
            if bitpos == 8:
                bitpos = 0
                bytepos += 1
        if bytepos == bufsize:
            out += buf
            buf = bytearray(bufsize)
            bytepos = 0
            partial = False
    return bytes(out)


def sqz_decode(sqz: bytes, table: bytes, n_chars: int = 0) -> np.ndarray:
    """Compressed<>::getnch / getch, char_io.t:105-151: character i = table[bits
    i*bits .. (i+1)*bits of the byte stream, most significant first]."""
    bits = sqz_bits(len(table))
    total = len(sqz) * 8 // bits
    if n_chars == 0:
        n_chars = total
    out = np.zeros(n_chars, np.uint8)
    for i in range(n_chars):
        code = 0
        for b in range(bits):
            pos = i * bits + b
            code = (code << 1) | ((sqz[pos >> 3] >> (7 - (pos & 7))) & 1)
        out[i] = table[code]
    return out


# ---------------------------------------------------------------------------
# the real reference, when oracle/_ref/primer_match exists
REF_BIN = os.path.join(_HERE, "_ref", "primer_match")
REF_FORMAT = r"%f\t%s\t%e\t%r\t%i\t%d\t%S\t%E\t%A\n"


def ref_available() -> bool:
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run_reference(fasta_path: str, primers, k=0, indels=True, rev_comp=False, wc=False, tn=False,
                  select=None, extra=(), timeout=600):
    """Run the unmodified reference binary and parse its records."""
    args = [REF_BIN, "-i", fasta_path, "-p", " ".join(primers), "-A", REF_FORMAT]
    if k > 0:
        args += ["-k" if indels else "-K", str(k)]
    if rev_comp:
        args.append("-r")
    if wc:
        args.append("-W" if tn else "-w")
    if select is None and k > 0:
        select = 5   # force filter_bitvec, the path under test (select.cc:234-237)
    if select:
        args += ["-N", str(select)]
    args += list(extra)
    p = subprocess.run(args, capture_output=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError("reference failed rc=%d: %s" % (p.returncode, p.stderr.decode()[-500:]))
    recs = []
    for line in p.stdout.decode("latin-1").splitlines():
        f = line.split("\t")
        recs.append(HitRecord(int(f[0]), int(f[1]), int(f[2]), f[3], int(f[4]), int(f[5]),
                              int(f[6]), int(f[7]), f[8]))
    return sorted(recs)

"""Host-side pieces of the product that need no GPU: the C-ABI library loads and
exports every symbol include/pm_b200.h declares, the FASTA normaliser and
reverse complement agree with the oracle, and argument errors are reported the
way the header promises.  No scan is run here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol(pkg, built_lib):
    hdr = open(os.path.join(ROOT, "include", "pm_b200.h")).read()
    declared = set(re.findall(r"\b(pm_[a-z_]+)\s*\(", hdr))
    declared -= {"pm_ctx"}
    assert declared == set(pkg.EXPORTS)
    for name in declared:
        assert hasattr(built_lib, name), name
    assert b"sm_100a" in built_lib.pm_version()


def test_struct_layouts_match_header(pkg):
    assert C.sizeof(pkg.Hit) == 8 + 8 + 8 + 4 + 4 + 4 + 4 + pkg.PM_MAX_CODES
    assert C.sizeof(pkg.Candidate) == 16
    assert C.sizeof(pkg.PcrOptions) == 6 * 4 + 2 * 8 + 2 * 8
    assert C.sizeof(pkg.PcrPair) == 8 + 8 + 8 + 4 + 4


def test_reverse_comp_matches_oracle(pkg, built_lib, oracle):
    for s in ["ACGT", "AAGGGYAATAA", "WKGKWRGGRWW", "acgtnRYKMSWBDHVNX", "U", "TTKGRRGGTAC"]:
        assert pkg.reverse_comp(s) == oracle.reverse_comp(s)


@pytest.mark.parametrize("seed", range(12))
def test_normalize_fasta_matches_oracle(pkg, built_lib, oracle, seed):
    import casegen
    fasta, _ = casegen.make_case(seed, nseq=1 + seed % 5, seqlen=50 + 37 * seed, line=7 + seed,
                                 crlf=bool(seed % 3 == 0), lower=0.05, text_iupac=0.02)
    raw = fasta.encode()
    for uc in (False, True):
        a = pkg.normalize_fasta(raw, uc=uc)
        b = oracle.normalize_fasta(raw, uc=uc)
        assert np.array_equal(a.chars, b.chars)
        assert np.array_equal(a.seq_key, b.seq_key)
        assert a.headers == b.headers


@pytest.mark.parametrize("raw", [b"", b">h\n", b">h\nACGT", b">h\nAC GT\n\n>g\n\nTT\n", b"ACGT\n>x\nGG\n",
                                 b">a\r\nAC\r\nGT\r\n>b\r\nTT\r\n", b">only header"])
def test_normalize_fasta_edge_cases(pkg, built_lib, oracle, raw):
    a = pkg.normalize_fasta(raw)
    b = oracle.normalize_fasta(raw)
    assert np.array_equal(a.chars, b.chars)
    assert np.array_equal(a.seq_key, b.seq_key)


def test_create_rejects_bad_options_or_missing_gpu(pkg, built_lib):
    import torch
    h = C.c_void_p()
    bad = pkg.Options(0, 0, 7, 1, 0, 0, 10)
    assert built_lib.pm_create(C.byref(h), C.byref(bad)) == pkg.PM_E_UNSUPPORTED
    bad = pkg.Options(0, pkg.PM_INDEX_KEYWORD_TREE, 2, 1, 0, 0, 10)
    assert built_lib.pm_create(C.byref(h), C.byref(bad)) == pkg.PM_E_ARG
    if not torch.cuda.is_available():
        ok = pkg.Options(0, 0, 1, 1, 0, 0, 10)
        # no device: the product fails loudly, it does not fall back to a CPU scan
        assert built_lib.pm_create(C.byref(h), C.byref(ok)) == pkg.PM_E_CUDA
        assert b"CUDA" in built_lib.pm_last_error()
        with pytest.raises(pkg.PmError):
            pkg.Scanner(k=1)

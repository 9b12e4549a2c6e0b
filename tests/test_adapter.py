"""The drop-in proof: the reference's OWN primer_match.cc / pcr_match.cc, compiled unchanged
from /root/reference and linked with tests/adapter/select_b200.cc (a pick_pattern_index that
returns tests/adapter/b200_match.h, the PatternMatch adapter of INTEGRATION.md) instead of its
select.cc.  The text its main() prints through the adapter must be the text the unmodified
reference printed (the goldens).

  not gpu   the adapter binary linked with the CPU stand-in of the C ABI (tests/mock), serving
            the oracle's hits: the adapter's own logic (database hand-over from the file the
            CharacterProducer names, prefix check, batches of report_interval hits) inside the
            reference's main loop
  gpu       the same binary linked with libpm_b200.so: scan on the device, everything else
            (FASTA reader, final alignments, alignformat) the reference's

The binaries need the reference's headers: they are built where /root/reference is mounted
(tests/adapter/Makefile, called by __graft_entry__.build()) and travel to the GPU box."""
import os
import subprocess
import sys

import pytest

from conftest import PKG_NAME, ROOT
from test_cli_host_logic import write_mock_inputs, pcr_case_inputs

ADAPTER = os.path.join(ROOT, "tests", "adapter")
BUILD = os.path.join(ADAPTER, "_build")


def _blocks(text, n):
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    assert len(lines) % n == 0
    return sorted(tuple(lines[i:i + n]) for i in range(0, len(lines), n))


@pytest.fixture(scope="module")
def adapter(built_lib):
    subprocess.check_call(["make", "-C", ADAPTER, "all"], stdout=subprocess.DEVNULL)
    exe = os.path.join(BUILD, "primer_match_b200")
    if not os.path.exists(exe):
        pytest.skip("adapter binaries unavailable: no /root/reference here and no prebuilt tests/adapter/_build")
    return {n: os.path.join(BUILD, n) for n in ("primer_match_b200", "pcr_match_b200", "primer_match_b200_mock",
                                                "pcr_match_b200_mock")}


def _run(exe, args, env=None):
    p = subprocess.run([exe] + args, capture_output=True, timeout=300, env=env)
    assert p.returncode == 0, p.stderr.decode("latin-1")[-800:]
    return p.stdout.decode("latin-1")


def _golden_args(fa, prm, par, extra=()):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    return make_golden.cli_args(fa, prm, par, extra)


def test_reference_main_through_the_adapter_on_the_mock(adapter, oracle, golden, tmp_path):
    real = os.path.join(ROOT, PKG_NAME, "lib", "libpm_b200.so")
    checked = 0
    for case in golden["cases"]:
        if "stdout_default" not in case:
            continue
        d = str(tmp_path / ("a%s" % str(case["seed"]).replace(":", "_")))
        os.makedirs(d)
        par, prm = case["params"], case["primers"]
        fasta = case["fasta"].encode("latin-1")
        st = oracle.normalize_fasta(fasta)
        pats = oracle.pattern_array(prm, par["rev_comp"])
        lc, rc = oracle.pattern_constraints(prm, par["rev_comp"])
        hits = oracle.aligned_hits(st, pats, lc, rc, par["k"], par["indels"], par["wc"], par["tn"])
        write_mock_inputs(d, fasta, st, hits)
        fa = os.path.join(d, "db.fasta")
        with open(fa, "wb") as f:
            f.write(fasta)
        env = dict(os.environ, PM_MOCK_DIR=d, PM_REAL_LIB=real)
        got = _run(adapter["primer_match_b200_mock"], _golden_args(fa, prm, par), env)
        assert _blocks(got, 4) == _blocks(case["stdout_default"], 4), "default format, case %s" % case["seed"]
        # -R 2: the hits arrive in batches of two find_patterns calls each
        got = _run(adapter["primer_match_b200_mock"], _golden_args(fa, prm, par, ["-R", "2", "-c"]), env)
        assert got == case["stdout_counts"], "counts, case %s" % case["seed"]
        checked += 1
    assert checked >= 15


def test_adapter_moves_the_producer_to_the_scan_frontier(adapter, oracle, tmp_path):
    """A database longer than the adapter's prefix probe, several sequences, exact search: the
    reference's main() does not read the text for an exact_alignment, so the sequence index and
    header of a hit in a later sequence are only right when find_patterns has moved the producer
    past it (what the reference's own indexes do by scanning)."""
    import casegen
    real = os.path.join(ROOT, PKG_NAME, "lib", "libpm_b200.so")
    for seed, k in ((71, 0), (72, 2)):
        fasta_text, primers = casegen.make_case(seed, nseq=5, seqlen=12000, nprimers=8, plen=(18, 22), k=k, indels=True)
        par = dict(k=k, indels=True, wc=False, tn=False, rev_comp=True)
        fasta = fasta_text.encode()
        d = str(tmp_path / ("f%d" % seed))
        os.makedirs(d)
        st = oracle.normalize_fasta(fasta)
        pats = oracle.pattern_array(primers, True)
        lc, rc = oracle.pattern_constraints(primers, True)
        hits = oracle.aligned_hits(st, pats, lc, rc, k, True, False, False)
        assert len(st.chars) > 20000 and any(h[2] > 18000 for h in hits)
        from test_cli_host_logic import write_mock_inputs as wmi
        wmi(d, fasta, st, hits)
        fa = os.path.join(d, "db.fasta")
        with open(fa, "wb") as f:
            f.write(fasta)
        ref = os.path.join(ROOT, "oracle", "_ref", "primer_match")
        if not os.path.exists(ref):
            pytest.skip("needs oracle/_ref/primer_match")
        args = _golden_args(fa, primers, par, ["-A", oracle.REF_FORMAT])
        want = subprocess.run([ref] + args, capture_output=True, timeout=300).stdout.decode("latin-1")
        env = dict(os.environ, PM_MOCK_DIR=d, PM_REAL_LIB=real)
        got = _run(adapter["primer_match_b200_mock"], args, env)
        assert sorted(got.splitlines()) == sorted(want.splitlines()) and len(want.splitlines()) >= 5


def test_reference_pcr_match_through_the_adapter_on_the_mock(adapter, oracle, pcr_golden, tmp_path):
    real = os.path.join(ROOT, PKG_NAME, "lib", "libpm_b200.so")
    n = 0
    for i, case in enumerate(pcr_golden["cases"]):
        d = str(tmp_path / ("q%d" % i))
        os.makedirs(d)
        fa, run = pcr_case_inputs(oracle, case, d)
        env = dict(os.environ, PM_MOCK_DIR=d, PM_REAL_LIB=real)
        args = oracle.pcr_cli_args(fa, case["primers"], **run)
        got = _run(adapter["pcr_match_b200_mock"], args + ["-A", oracle.PCR_REF_FORMAT], env)
        assert got == case["stdout_records"], "records, case %s" % case["seed"]
        got = _run(adapter["pcr_match_b200_mock"], args, env)
        assert got == case["stdout_default"], "default format, case %s" % case["seed"]
        n += len(case["stdout_records"].splitlines())
    assert n > 300


@pytest.mark.gpu
def test_reference_main_through_the_adapter_on_the_gpu(adapter, golden, tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device")
    checked = 0
    for case in golden["cases"]:
        if "stdout_default" not in case:
            continue
        checked += 1
        if checked % 2 == 0:          # a process start costs a CUDA context: every other case
            continue
        fa = tmp_path / ("g%s.fasta" % str(case["seed"]).replace(":", "_"))
        fa.write_bytes(case["fasta"].encode("latin-1"))
        par, prm = case["params"], case["primers"]
        got = _run(adapter["primer_match_b200"], _golden_args(str(fa), prm, par))
        assert _blocks(got, 4) == _blocks(case["stdout_default"], 4), "default format, case %s" % case["seed"]
        if checked % 6 == 1:
            got = _run(adapter["primer_match_b200"], _golden_args(str(fa), prm, par, ["-c"]))
            assert got == case["stdout_counts"], "counts, case %s" % case["seed"]
    assert checked >= 15


@pytest.mark.gpu
def test_reference_pcr_match_through_the_adapter_on_the_gpu(adapter, oracle, pcr_golden, tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device")
    n = 0
    for i, case in enumerate(pcr_golden["cases"]):
        if i % 5 != 0:                # a process start costs a CUDA context
            continue
        d = str(tmp_path / ("h%d" % i))
        os.makedirs(d)
        fa = os.path.join(d, "db.fasta")
        with open(fa, "wb") as f:
            f.write(case["fasta"].encode("latin-1"))
        run = dict(case["params"])
        if "sts" in case:
            run["sts_path"] = os.path.join(d, "primers.sts")
            with open(run["sts_path"], "w") as f:
                f.write(case["sts"])
        args = oracle.pcr_cli_args(fa, case["primers"], **run)
        got = _run(adapter["pcr_match_b200"], args + ["-A", oracle.PCR_REF_FORMAT])
        assert got == case["stdout_records"], "records, case %s" % case["seed"]
        n += 1
    assert n >= 10


@pytest.mark.gpu
def test_reference_main_through_the_adapter_on_large_databases(adapter, oracle, tmp_path):
    """Databases of 1-2.5 Mbp (several sequences, N runs, lower case, IUPAC text) with the automatic
    index choice: the records the reference's main() prints through the adapter against the ones the
    unmodified reference printed (tests/golden/large_golden.json).  The database is larger than the
    adapter's prefix probe, so the producer is left mid-stream, as a real run leaves it."""
    import json
    import make_large_golden as M
    with open(os.path.join(ROOT, "tests", "golden", "large_golden.json")) as f:
        large = json.load(f)
    n = 0
    for entry in large["cases"]:
        c = entry["case"]
        if c["seed"] not in (1, 2, 4, 5, 6):
            continue
        fasta, primers = M.large_case(c)
        fa = tmp_path / ("large%d.fasta" % c["seed"])
        fa.write_bytes(fasta)
        pf = tmp_path / ("primers%d.txt" % c["seed"])
        pf.write_text("\n".join(primers) + "\n")
        args = ["-i", str(fa), "-P", str(pf), "-A", oracle.REF_FORMAT, "-r"]
        if c["k"] > 0:
            args += ["-k" if c["indels"] else "-K", str(c["k"])]
        if c["wc"]:
            args.append("-W" if c["tn"] else "-w")
        if c["select"]:
            args += ["-N", str(c["select"])]
        got = _run(adapter["primer_match_b200"], args)
        recs = []
        for line in got.splitlines():
            f = line.split("\t")
            recs.append((int(f[0]), int(f[1]), int(f[2]), f[3], int(f[4]), int(f[5]), int(f[6]), int(f[7]), f[8]))
        want = sorted(tuple(r) for r in entry["records"])
        assert sorted(recs) == want, "case %d" % c["seed"]
        n += len(want)
    assert n > 1000

"""The drop-in command line tool (bin/primer_match, C++ host code over the C
ABI) against the text the UNMODIFIED reference printed for the same command
lines (tests/golden/make_golden.py: default format, a custom -A format, -c and
-C/-a counts).  The text is compared record by record IN ORDER: primer_match prints
its hits as the index reports them, and the tool hands them over in that order
(pm_hit_order: filter_bitvec's std::sort of the candidates, chains in that order)."""
import importlib
import os
import subprocess
import sys

import pytest

from conftest import PKG_NAME, ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cli(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device; the tool has no CPU fallback")
    build = importlib.import_module(PKG_NAME + ".build")
    path = build.build_cli()
    assert path and os.path.exists(path)
    return path


@pytest.fixture(scope="module")
def pcr_cli(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device; the tool has no CPU fallback")
    build = importlib.import_module(PKG_NAME + ".build")
    path = build.build_cli(tool="pcr_match")
    assert path and os.path.exists(path)
    return path


def _args(fasta_path, primers, params, extra=()):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    return make_golden.cli_args(fasta_path, primers, params, extra)


def _run(cli, args):
    p = subprocess.run([cli] + args, capture_output=True, timeout=300)
    assert p.returncode == 0, p.stderr.decode()[-500:]
    return p.stdout.decode("latin-1")


def _blocks(text, n):
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    assert len(lines) % n == 0
    return [tuple(lines[i:i + n]) for i in range(0, len(lines), n)]


def test_cli_output_matches_reference(cli, golden, tmp_path):
    # every process start pays for a CUDA context (~3 s): the default format for every case, the
    # other formats for every third; tests/test_cli_host_logic.py runs all of them on the CPU
    checked = 0
    for case in golden["cases"]:
        if "stdout_default" not in case:
            continue
        fa = tmp_path / ("c%s.fasta" % str(case["seed"]).replace(":", "_"))
        fa.write_bytes(case["fasta"].encode("latin-1"))
        prm, par = case["primers"], case["params"]
        got = _run(cli, _args(str(fa), prm, par))
        assert _blocks(got, 4) == _blocks(case["stdout_default"], 4), "default format, case %s" % case["seed"]
        checked += 1
        if checked % 3 != 1:
            continue
        got = _run(cli, _args(str(fa), prm, par, ["-A", "%H %f %5 %3 %l %q %p %t|%T|%Q|%A %| %^ %v %* %+ %%\\n"]))
        assert _blocks(got, 1) == _blocks(case["stdout_custom"], 1), "custom format, case %s" % case["seed"]
        got = _run(cli, _args(str(fa), prm, par, ["-c"]))
        assert got == case["stdout_counts"], "counts, case %s" % case["seed"]
        got = _run(cli, _args(str(fa), prm, par, ["-C", "%i\\t%p\\t%c\\t%C\\n", "-a"]))
        assert sorted(got.split("\n")) == sorted(case["stdout_counts_agg"].split("\n")), "aggregate counts, case %s" % case["seed"]
    assert checked >= 15


def test_cli_pattern_files_and_output_file(cli, golden, tmp_path):
    case = [c for c in golden["cases"] if "stdout_default" in c and len(c["records"]) > 3][0]
    fa = tmp_path / "db.fasta"
    fa.write_bytes(case["fasta"].encode("latin-1"))
    pf = tmp_path / "primers.txt"
    pf.write_text("\n".join(case["primers"]) + "\n")
    pfa = tmp_path / "primers.fasta"
    pfa.write_text("".join(">p%d some primer\n%s\n" % (i, p) for i, p in enumerate(case["primers"])))
    base = _args(str(fa), case["primers"], case["params"])
    i = base.index("-p")
    want = _blocks(case["stdout_default"], 4)
    outf = tmp_path / "out.txt"
    _run(cli, base[:i] + ["-P", str(pf)] + base[i + 2:] + ["-o", str(outf)])
    assert _blocks(outf.read_text(), 4) == want
    got = _run(cli, base[:i] + ["-F", str(pfa)] + base[i + 2:])
    assert _blocks(got, 4) == want


REF_FORMAT = "%f\\t%s\\t%e\\t%r\\t%i\\t%d\\t%S\\t%E\\t%A\\n"


def _records(text):
    out = []
    for line in text.splitlines():
        f = line.split("\t")
        out.append((int(f[0]), int(f[1]), int(f[2]), f[3], int(f[4]), int(f[5]), int(f[6]), int(f[7]), f[8]))
    return sorted(out)


def test_cli_automatic_index_and_constraints(cli, halves_golden, constraints_golden, tmp_path):
    """The same command lines the reference got in make_halves_golden.py / make_constraints_golden.py:
    no -N (pick_pattern_index chooses), -N 13/14 forced, and the -s / -e / -5 / -3 options.
    Every third case (a process start costs ~3 s of CUDA context); test_gpu_parity.py checks all
    of them through the C ABI in one process."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_constraints_golden
    n = 0
    for gi, g in enumerate((halves_golden, constraints_golden)):
        for case in g["cases"][gi::3]:
            fa = tmp_path / ("g%d_%s.fasta" % (gi, case["seed"]))
            fa.write_bytes(case["fasta"].encode("latin-1"))
            par = case["params"]
            args = ["-i", str(fa), "-p", " ".join(case["primers"]), "-A", REF_FORMAT]
            if par["k"] > 0:
                args += ["-k" if par["indels"] else "-K", str(par["k"])]
            if par["rev_comp"]:
                args.append("-r")
            if par["wc"]:
                args.append("-W" if par["tn"] else "-w")
            if case["select"]:
                args += ["-N", str(case["select"])]
            if "constraints" in case:
                args += make_constraints_golden.cli_constraints(case["constraints"])
            got = _records(_run(cli, args))
            assert got == sorted(tuple(r) for r in case["records"]), "golden %d case %s" % (gi, case["seed"])
            n += 1
    assert n >= 30


def test_pcr_match_output_matches_reference(pcr_cli, oracle, pcr_golden, tmp_path):
    """bin/pcr_match on the GPU library, command lines the reference got in make_pcr_golden.py:
    records (every alignformat argument) in the reference's order, and its default output
    format.  Every process start pays for a CUDA context, so this takes a spread of the cases
    with records; tests/test_cli_host_logic.py runs all of them (three formats each) on the
    oracle's hits."""
    with_records = [c for c in pcr_golden["cases"] if c["stdout_records"]]
    picked = with_records[::5]
    assert len(picked) >= 10 and any("sts" in c for c in picked)
    n = 0
    for ci, case in enumerate(picked):
        fa = tmp_path / ("pcr%d.fasta" % case["seed"])
        fa.write_bytes(case["fasta"].encode("latin-1"))
        run = dict(case["params"])
        if "sts" in case:
            sts = tmp_path / ("pcr%d.sts" % case["seed"])
            sts.write_text(case["sts"])
            run["sts_path"] = str(sts)
        args = oracle.pcr_cli_args(str(fa), case["primers"], **run)
        got = _run(pcr_cli, args + ["-A", oracle.PCR_REF_FORMAT])
        assert got == case["stdout_records"], "records, case %d" % case["seed"]
        if ci % 4 == 0:
            got = _run(pcr_cli, args)
            assert got == case["stdout_default"], "default format, case %d" % case["seed"]
        n += len(case["stdout_records"].splitlines())
    assert n > 40


def test_cli_opens_the_database_like_pick_fasta_file(cli, golden, tmp_path):
    """-D: the FASTA file itself (1), or the .sqn (3) / .sqz (4) database the reference's compress_seq
    wrote next to it, or whichever exists (0) -- read by the library's own reader.  The expectation
    is what the UNMODIFIED reference prints for the same -D on the same files (run here: a
    compressed database is not the same text as its FASTA file -- compress_seq upper-cases it -- and
    its small alphabet changes the automatic index choice, select.cc:120-123)."""
    csq = os.path.join(ROOT, "oracle", "_ref", "compress_seq")
    ref = os.path.join(ROOT, "oracle", "_ref", "primer_match")
    if not (os.path.exists(csq) and os.path.exists(ref)):
        pytest.skip("needs oracle/_ref/compress_seq and oracle/_ref/primer_match")
    checked = differ = 0
    for case in golden["cases"]:
        if "stdout_default" not in case or case["params"]["wc"]:
            continue
        checked += 1
        if checked % 4 != 1:
            continue
        d = tmp_path / ("db%d" % checked)
        d.mkdir()
        fa = d / "db.fasta"
        fa.write_bytes(case["fasta"].encode("latin-1"))
        prm, par = case["primers"], case["params"]
        want = _blocks(case["stdout_default"], 4)
        assert _blocks(_run(cli, _args(str(fa), prm, par, ["-D", "1"])), 4) == want, "FASTA, case %s" % case["seed"]
        p = subprocess.run([csq, "-i", str(fa), "-z", "true", "-n", "true"], capture_output=True, timeout=120)
        assert p.returncode == 0 and (d / "db.fasta.sqz").exists() and (d / "db.fasta.idb").exists()
        for dbind in ("4", "3", "0"):
            args = _args(str(fa), prm, par, ["-D", dbind])
            r = subprocess.run([ref] + args, capture_output=True, timeout=300)
            assert r.returncode == 0
            want_d = _blocks(r.stdout.decode("latin-1"), 4)
            differ += want_d != want
            got = _run(cli, args)
            assert _blocks(got, 4) == want_d, "-D %s, case %s" % (dbind, case["seed"])
    assert checked >= 8 and differ >= 1


def test_cli_uppercase_maxcount_and_eos_options(cli, golden, tmp_path):
    """-u (primers and database upper-cased), -M (a primer stops being reported once it has that
    many hits; the counts carry a '+'), -E (another end-of-sequence character): the same command
    line on the UNMODIFIED reference (oracle/_ref/primer_match, run here) and on the drop-in tool."""
    ref = os.path.join(ROOT, "oracle", "_ref", "primer_match")
    if not os.path.exists(ref):
        pytest.skip("needs oracle/_ref/primer_match")
    variants = (["-u"], ["-c", "-M", "2"], ["-C", "%i %r %p %+ %c", "-a", "-M", "1"],
                ["-c", "-M", "1", "-A", "%h %s %e %i\n"], ["-E", "36"])
    checked = with_hits = 0
    for case in golden["cases"]:
        if "stdout_default" not in case or case["params"]["wc"]:
            continue
        checked += 1
        if checked % 3 != 1:
            continue
        d = tmp_path / ("c%d" % checked)
        d.mkdir()
        fa = d / "db.fasta"
        txt = case["fasta"]
        mixed = "\n".join(l if l.startswith(">") else "".join(c.lower() if i % 3 == 0 else c for i, c in enumerate(l))
                          for l in txt.split("\n"))
        for extra in variants:
            lower = extra == ["-u"]
            fa.write_bytes((mixed if lower else txt).encode("latin-1"))
            prm = [p.lower() for p in case["primers"]] if lower else case["primers"]
            args = _args(str(fa), prm, case["params"], extra)
            r = subprocess.run([ref] + args, capture_output=True, timeout=300)
            assert r.returncode == 0
            want = sorted(r.stdout.decode("latin-1").split("\n"))
            got = sorted(_run(cli, args).split("\n"))
            with_hits += len(want) > 3
            assert got == want, "%s, case %s" % (extra, case["seed"])
    assert checked >= 8 and with_hits >= 5

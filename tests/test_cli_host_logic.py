"""Host logic of the command line tools without a GPU: bin/pcr_match's option handling,
primer input (-p / -P / -F / -S), pairing of hits into amplicons and output formats,
against the text the UNMODIFIED reference pcr_match printed (tests/golden/make_pcr_golden.py).

The tools are compiled against tests/mock/pm_mock.cc, a stand-in for the C ABI that serves
the hits the oracle computes for the case (the scan itself is what the `-m gpu` tests
cover: tests/test_gpu_cli.py runs the same command lines on the real library)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import PKG_NAME, ROOT

MOCK = os.path.join(ROOT, "tests", "mock")
CSRC = os.path.join(ROOT, PKG_NAME, "csrc")
CODE_OF = {"|": 1, "+": 2, "*": 3, "^": 4, "v": 5, "!": 6}


@pytest.fixture(scope="module")
def mock_tools(built_lib):
    out = os.path.join(MOCK, "_build")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libpm_b200.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                           "-o", lib, os.path.join(MOCK, "pm_mock.cc"), "-ldl"])
    tools = {}
    for tool in ("pcr_match", "primer_match"):
        exe = os.path.join(out, tool)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", exe,
                               os.path.join(CSRC, tool + "_main.cc"), "-L", out, "-lpm_b200", "-Wl,-rpath," + out])
        tools[tool] = exe
    real = os.path.join(ROOT, PKG_NAME, "lib", "libpm_b200.so")
    assert os.path.exists(real)
    return tools, real


def write_mock_inputs(d, fasta: bytes, stream, hits):
    """stream.bin / seqs.bin / hits.bin for tests/mock/pm_mock.cc; hits = [(pid, start, end, ed, astr)]."""
    stream.chars.astype(np.uint8).tofile(os.path.join(d, "stream.bin"))
    hoff, hlen, at = [], [], 0
    for h in stream.headers:
        at = fasta.index(b">" + h.encode("latin-1"), at) + 1
        hoff.append(at)
        hlen.append(len(h))
    n = len(stream.seq_key)
    with open(os.path.join(d, "seqs.bin"), "wb") as f:
        f.write(struct.pack("<q", n))
        f.write(np.asarray(stream.seq_key, np.int64).tobytes())
        f.write(np.asarray(hoff, np.int64).tobytes())
        f.write(np.asarray(hlen, np.int64).tobytes())
    with open(os.path.join(d, "hits.bin"), "wb") as f:
        for pid, start, end, ed, astr in hits:
            codes = bytes(CODE_OF[c] for c in astr)
            f.write(struct.pack("<qqqIiiI72s", start, end, end, pid, ed, 0, len(codes), codes))


def run_tool(exe, args, mock_dir, real_lib, stdin=None):
    env = dict(os.environ, PM_MOCK_DIR=mock_dir, PM_REAL_LIB=real_lib)
    p = subprocess.run([exe] + args, capture_output=True, timeout=120, env=env, input=stdin)
    assert p.returncode == 0, p.stderr.decode()[-500:]
    return p.stdout.decode("latin-1")


def pcr_case_inputs(oracle, case, d):
    par = dict(case["params"])
    select = par.pop("select")
    par.pop("deviation", None)
    fasta = case["fasta"].encode("latin-1")
    st = oracle.normalize_fasta(fasta)
    rev_comp = par["rev_comp"] or "sts" in case
    pats = oracle.pcr_pattern_array(case["primers"], rev_comp)
    lc, rc = oracle.pcr_pattern_constraints(pats[:len(case["primers"])], par["stlen"], par["edlen"], par["fplen"],
                                            par["tplen"])
    hits = oracle.aligned_hits(st, pats, lc, rc, par["k"], par["indels"], par["wc"], par["tn"],
                               index="auto" if select == 0 else "filter_bitvec", bogus_ok=True, exact_align=False,
                               keep_bogus=True)
    write_mock_inputs(d, fasta, st, hits)
    fa = os.path.join(d, "db.fasta")
    with open(fa, "wb") as f:
        f.write(fasta)
    run = dict(case["params"])
    if "sts" in case:
        run["sts_path"] = os.path.join(d, "primers.sts")
        with open(run["sts_path"], "w") as f:
            f.write(case["sts"])
    return fa, run


def test_pcr_match_host_logic_matches_reference_output(oracle, pcr_golden, mock_tools, tmp_path):
    import make_pcr_golden
    tools, real = mock_tools
    n = 0
    for case in pcr_golden["cases"]:
        d = str(tmp_path / ("c%d" % case["seed"]))
        os.makedirs(d)
        fa, run = pcr_case_inputs(oracle, case, d)
        args = oracle.pcr_cli_args(fa, case["primers"], **run)
        got = run_tool(tools["pcr_match"], args + ["-A", oracle.PCR_REF_FORMAT], d, real)
        assert got == case["stdout_records"], "records, case %d" % case["seed"]
        got = run_tool(tools["pcr_match"], args, d, real)
        assert got == case["stdout_default"], "default format, case %d" % case["seed"]
        fmt = make_pcr_golden.CUSTOM_FORMAT + ("%*\\n" if "sts" in case else "")
        got = run_tool(tools["pcr_match"], args + ["-A", fmt], d, real)
        assert got == case["stdout_custom"], "custom format, case %d" % case["seed"]
        n += len(case["stdout_records"].splitlines())
    assert n > 300


def test_pcr_match_host_logic_orders_simultaneous_hits_like_the_reference(oracle, pcr_ties_golden, mock_tools, tmp_path):
    """pm_pcr_pairs runs the std::sort the reference runs (csrc/pm_pcr.hpp): the 15 cases in which
    hits of different primers end at one position and the order decides the pairs."""
    tools, real = mock_tools
    for case in pcr_ties_golden["cases"]:
        d = str(tmp_path / ("t%d" % case["seed"]))
        os.makedirs(d)
        fa, run = pcr_case_inputs(oracle, case, d)
        args = oracle.pcr_cli_args(fa, case["primers"], **run)
        got = run_tool(tools["pcr_match"], args + ["-A", oracle.PCR_REF_FORMAT], d, real)
        assert got == case["stdout_records"], "records, case %d" % case["seed"]


def test_pcr_match_primer_files(oracle, pcr_golden, mock_tools, tmp_path):
    """-P (words), -F (FASTA, %P = the deflines), "-" = standard input, -o appends."""
    tools, real = mock_tools
    case = [c for c in pcr_golden["cases"] if "sts" not in c and len(c["stdout_records"].splitlines()) > 3][0]
    d = str(tmp_path)
    fa, run = pcr_case_inputs(oracle, case, d)
    base = oracle.pcr_cli_args(fa, case["primers"], **run)
    i = base.index("-p")
    words = "\n".join(case["primers"]) + "\n"
    with open(os.path.join(d, "p.txt"), "w") as f:
        f.write(words)
    fmt = ["-A", oracle.PCR_REF_FORMAT]
    assert run_tool(tools["pcr_match"], base[:i] + ["-P", os.path.join(d, "p.txt")] + base[i + 2:] + fmt, d,
                    real) == case["stdout_records"]
    assert run_tool(tools["pcr_match"], base[:i] + ["-P", "-"] + base[i + 2:] + fmt, d, real,
                    stdin=words.encode()) == case["stdout_records"]
    with open(os.path.join(d, "p.fa"), "w") as f:
        f.write("".join(">primer%d of pair %d\n%s\n" % (j, j // 2, p) for j, p in enumerate(case["primers"])))
    got = run_tool(tools["pcr_match"], base[:i] + ["-F", os.path.join(d, "p.fa")] + base[i + 2:] +
                   ["-A", "%i %>P|%<P\\n"], d, real)
    want = oracle.parse_pcr_records(case["stdout_records"])
    assert len(got.splitlines()) == len(want)
    for line, r in zip(got.splitlines(), want):
        pair = int(line.split(" ")[0])
        assert pair == r.pair and "of pair %d|" % (pair - 1) in line and line.endswith("of pair %d" % (pair - 1))
    outf = os.path.join(d, "out.txt")
    with open(outf, "w") as f:
        f.write("kept\n")
    run_tool(tools["pcr_match"], base + fmt + ["-o", outf], d, real)
    assert open(outf).read() == "kept\n" + case["stdout_records"]


def test_pcr_match_odd_number_of_primers_is_refused(mock_tools, tmp_path):
    tools, real = mock_tools
    env = dict(os.environ, PM_MOCK_DIR=str(tmp_path), PM_REAL_LIB=real)
    p = subprocess.run([tools["pcr_match"], "-i", "x.fasta", "-p", "ACGTACGTACGT ACGTTTGACCA GGGTTTAAACCC"],
                       capture_output=True, env=env)
    assert p.returncode == 1 and b"Odd number of primers!" in p.stderr   # pcr_match.cc:768


def _blocks(text, n):
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    assert len(lines) % n == 0
    return [tuple(lines[i:i + n]) for i in range(0, len(lines), n)]


def test_primer_match_host_logic_matches_reference_output(oracle, golden, mock_tools, tmp_path):
    """bin/primer_match's formats and counts on the oracle's hits, against the reference's text
    (records as multisets: the reference orders simultaneous hits by an unstable sort)."""
    import make_golden
    tools, real = mock_tools
    checked = 0
    for case in golden["cases"]:
        if "stdout_default" not in case:
            continue
        d = str(tmp_path / ("p%s" % str(case["seed"]).replace(":", "_")))
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
        exe = tools["primer_match"]
        got = run_tool(exe, make_golden.cli_args(fa, prm, par), d, real)
        assert _blocks(got, 4) == _blocks(case["stdout_default"], 4), "default format, case %s" % case["seed"]
        got = run_tool(exe, make_golden.cli_args(fa, prm, par, ["-A", "%H %f %5 %3 %l %q %p %t|%T|%Q|%A %| %^ %v %* %+ %%\\n"]),
                       d, real)
        assert _blocks(got, 1) == _blocks(case["stdout_custom"], 1), "custom format, case %s" % case["seed"]
        got = run_tool(exe, make_golden.cli_args(fa, prm, par, ["-c"]), d, real)
        assert got == case["stdout_counts"], "counts, case %s" % case["seed"]
        got = run_tool(exe, make_golden.cli_args(fa, prm, par, ["-C", "%i\\t%p\\t%c\\t%C\\n", "-a"]), d, real)
        assert sorted(got.split("\n")) == sorted(case["stdout_counts_agg"].split("\n")), "aggregate counts, case %s" % case["seed"]
        checked += 1
    assert checked >= 15


def test_pm_pcr_pairs_on_oracle_hits(pkg, built_lib, oracle, pcr_golden):
    """pm_pcr_pairs (host code of the library, no device) through the Python binding, on the
    oracle's hits: the pairs behind every record the reference printed, in its order."""
    n = 0
    for case in pcr_golden["cases"]:
        par = dict(case["params"])
        select = par.pop("select")
        st = oracle.normalize_fasta(case["fasta"].encode("latin-1"))
        rev_comp = par["rev_comp"] or "sts" in case
        pats = oracle.pcr_pattern_array(case["primers"], rev_comp)
        assert pats == pkg.pcr_pattern_array(case["primers"], rev_comp)
        lc, rc = oracle.pcr_pattern_constraints(pats[:len(case["primers"])], par["stlen"], par["edlen"], par["fplen"],
                                                par["tplen"])
        hits = oracle.aligned_hits(st, pats, lc, rc, par["k"], par["indels"], par["wc"], par["tn"],
                                   index="auto" if select == 0 else "filter_bitvec", bogus_ok=True, exact_align=False)
        arr = np.zeros(len(hits), pkg.HIT_DTYPE)
        for i, (pid, start, end, ed, astr) in enumerate(hits):
            arr[i]["start"], arr[i]["end"], arr[i]["first"] = start, end, end
            arr[i]["pattern_id"], arr[i]["editdist"], arr[i]["ncodes"] = pid, ed, len(astr)
        got = pkg.pcr_pairs(arr, len(case["primers"]), [0] + [len(p) for p in pats], st.seq_key, par["k"],
                            par["indels"], par["allorient"], par["betweenlen"], par["mindist"], par["maxdist"],
                            case.get("sts_sizes"), par.get("deviation", -1))
        want = oracle.parse_pcr_records(case["stdout_records"])
        assert [(int(arr[a]["end"]), int(arr[b]["end"]), pind, alen) for a, b, alen, pind in got] == \
               [(r.stream_end, r.stream_end1, r.pair, len(r.amplicon)) for r in want], "case %s" % case["seed"]
        n += len(got)
    assert n > 300


def test_primer_match_constraint_options_and_index_selections(oracle, halves_golden, constraints_golden, mock_tools,
                                                              tmp_path):
    """The command lines of make_halves_golden.py / make_constraints_golden.py (-N 0 / 5 / 13,
    -s / -e / -5 / -3 with the `~n` form) through bin/primer_match on the oracle's hits: every
    record field the reference printed."""
    import make_constraints_golden
    tools, real = mock_tools
    fmt = "%f\\t%s\\t%e\\t%r\\t%i\\t%d\\t%S\\t%E\\t%A\\n"
    n = 0
    for gi, g in enumerate((halves_golden, constraints_golden)):
        for case in g["cases"]:
            d = str(tmp_path / ("g%d_%s" % (gi, case["seed"])))
            os.makedirs(d)
            par, prm = case["params"], case["primers"]
            cons = case.get("constraints", dict(stlen=0, edlen=0, fplen=0, tplen=0))
            fasta = case["fasta"].encode("latin-1")
            st = oracle.normalize_fasta(fasta)
            pats = oracle.pattern_array(prm, par["rev_comp"])
            lc, rc = oracle.pattern_constraints(prm, par["rev_comp"], **cons)
            index = case.get("index", "exact_halves") if gi else "exact_halves"
            hits = oracle.aligned_hits(st, pats, lc, rc, par["k"], par["indels"], par["wc"], par["tn"], index=index)
            write_mock_inputs(d, fasta, st, hits)
            fa = os.path.join(d, "db.fasta")
            with open(fa, "wb") as f:
                f.write(fasta)
            args = ["-i", fa, "-p", " ".join(prm), "-A", fmt]
            if par["k"] > 0:
                args += ["-k" if par["indels"] else "-K", str(par["k"])]
            if par["rev_comp"]:
                args.append("-r")
            if par["wc"]:
                args.append("-W" if par["tn"] else "-w")
            if case["select"]:
                args += ["-N", str(case["select"])]
            args += make_constraints_golden.cli_constraints(cons)
            got = run_tool(tools["primer_match"], args, d, real)
            recs = sorted((int(f[0]), int(f[1]), int(f[2]), f[3], int(f[4]), int(f[5]), int(f[6]), int(f[7]), f[8])
                          for f in (line.split("\t") for line in got.splitlines()))
            assert recs == sorted(tuple(r) for r in case["records"]), "golden %d case %s" % (gi, case["seed"])
            n += 1
    assert n >= 90


def test_primer_match_unists_input(oracle, mock_tools, tmp_path):
    """primer_match -S: primers and the %I %L %a %O %& %X codes from a UniSTS file, in the
    alignment and the count formats (tests/golden/make_sts_golden.py)."""
    import json
    import make_sts_golden
    tools, real = mock_tools
    with open(os.path.join(ROOT, "tests", "golden", "primer_match_sts_golden.json")) as f:
        g = json.load(f)
    assert len(g["cases"]) >= 8
    for case in g["cases"]:
        d = str(tmp_path / ("s%d" % case["seed"]))
        os.makedirs(d)
        par, prm = case["params"], case["primers"]
        fasta = case["fasta"].encode("latin-1")
        st = oracle.normalize_fasta(fasta)
        pats = oracle.pattern_array(prm, True)              # -S implies -r (primer_match.cc:137-141)
        lc, rc = oracle.pattern_constraints(prm, True)
        hits = oracle.aligned_hits(st, pats, lc, rc, par["k"], par["indels"])
        write_mock_inputs(d, fasta, st, hits)
        fa, sp = os.path.join(d, "db.fasta"), os.path.join(d, "primers.sts")
        with open(fa, "wb") as f:
            f.write(fasta)
        with open(sp, "w") as f:
            f.write(case["sts"])
        args = make_sts_golden.sts_cli_args(fa, sp, par)
        got = run_tool(tools["primer_match"], args + ["-A", make_sts_golden.ALIGN_FORMAT], d, real)
        assert sorted(got.splitlines()) == sorted(case["stdout_align"].splitlines()), "case %d" % case["seed"]
        got = run_tool(tools["primer_match"], args + ["-C", make_sts_golden.COUNT_FORMAT], d, real)
        assert got == case["stdout_counts"], "counts, case %d" % case["seed"]
        got = run_tool(tools["primer_match"], args + ["-C", make_sts_golden.COUNT_FORMAT, "-a"], d, real)
        assert got == case["stdout_counts_agg"], "aggregate counts, case %d" % case["seed"]
        with open(sp) as f:
            got = run_tool(tools["primer_match"], ["-S", "-"] + args[args.index("-S") + 2:] + ["-i", fa, "-A",
                           make_sts_golden.ALIGN_FORMAT], d, real, stdin=f.read().encode())
        assert sorted(got.splitlines()) == sorted(case["stdout_align"].splitlines())

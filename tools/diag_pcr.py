"""Diagnostic: every pcr golden case through the Python binding; prints the cases that differ."""
import dataclasses, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pm_oracle as O
pkg = importlib.import_module("sequence-alignment-tools_b200")
g = json.load(open(os.path.join(ROOT, "tests", "golden", "pcr_match_golden.json")))
for case in g["cases"]:
    st = pkg.normalize_fasta(case["fasta"].encode("latin-1"))
    par = dict(case["params"]); select = par.pop("select")
    want = [dataclasses.astuple(r) for r in O.parse_pcr_records(case["stdout_records"])]
    try:
        got = [dataclasses.astuple(r) for r in pkg.pcr_match(st, case["primers"], index=pkg.PM_INDEX_AUTO if select == 0 else pkg.PM_INDEX_FILTER_BITVEC, sts_sizes=case.get("sts_sizes"), **par)]
    except Exception as e:
        print("case", case["seed"], "EXC", repr(e)[:300], par, select); continue
    if got != want:
        print("case", case["seed"], par, "select", select, "sts" in case, "got", len(got), "want", len(want))
        sg, sw = set(got), set(want)
        for r in sorted(sg - sw)[:3]: print("   extra  ", r[:18])
        for r in sorted(sw - sg)[:3]: print("   missing", r[:18])
        if sg == sw: print("   same set, different order/multiplicity")
print("done")

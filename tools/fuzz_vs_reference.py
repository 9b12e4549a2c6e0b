#!/usr/bin/env python
"""Differential campaign: seeded random databases and primer sets (tests/golden/casegen.py, the
generator of the goldens) through the UNMODIFIED reference binary (oracle/_ref/primer_match) and
through the GPU library on both scan paths; every record compared.  Not part of the test suite: a
way to spend spare GPU minutes looking for parity bugs beyond the committed goldens.

    python tools/fuzz_vs_reference.py --first 10000 --count 2000 > gpurun_out/fuzz.json
"""
import argparse
import importlib
import json
import os
import random
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--first", type=int, default=10000)
    ap.add_argument("--count", type=int, default=500)
    ap.add_argument("--seconds", type=float, default=600.0)
    ap.add_argument("--oracle", action="store_true",
                    help="compare the CPU oracle (oracle/pm_oracle.py) with the reference instead of the GPU library: "
                         "no GPU needed")
    ap.add_argument("--long", action="store_true", help="primers of 33-64 characters (primer mode)")
    ap.add_argument("--mode", default="primer", choices=["primer", "constraints", "pcr"],
                    help="primer_match; primer_match with -s/-e/-5/-3 (exact_bases); pcr_match")
    a = ap.parse_args()
    if a.mode != "primer":
        return other_modes(a)
    import casegen
    import make_golden
    import pm_oracle as O
    from conftest import as_tuples
    pkg = importlib.import_module("sequence-alignment-tools_b200")
    tmp = tempfile.mkdtemp()
    t0 = time.time()
    done = skipped = records = 0
    failures = []
    by_kind = {}
    O_MAX_K = 3   # the oracle's recurrence is written for k <= 3 in one machine word
    for seed in range(a.first, a.first + a.count):
        if time.time() - t0 > a.seconds:
            break
        params, kw = make_golden.case_params(seed)
        rng = random.Random(seed * 7 + 1)
        select = rng.choice([None, None, 0])       # -N 5, or the reference's automatic choice
        if rng.random() < 0.1:
            params["k"] = kw["k"] = rng.choice([4, 5])
        if a.long:
            kw["plen"] = rng.choice([(33, 40), (33, 64), (48, 64), (64, 64)])
        if rng.random() < 0.3:                       # longer databases: several tiles' worth of rows
            kw["seqlen"] = rng.choice([5000, 20000])
        fasta, primers = casegen.make_case(seed, **kw)
        path = os.path.join(tmp, "f.fasta")
        with open(path, "w", newline="") as f:
            f.write(fasta)
        try:
            want = O.run_reference(path, primers, select=select, **params)
        except RuntimeError:
            skipped += 1                             # the reference aborts on this input
            continue
        kind = "k%d%s%s%s" % (params["k"], "e" if params["indels"] else "s", "w" if params["wc"] else "",
                              "-auto" if select == 0 else "")
        index = pkg.PM_INDEX_AUTO if (select == 0 or params["k"] == 0) else pkg.PM_INDEX_FILTER_BITVEC
        st = pkg.normalize_fasta(fasta.encode("latin-1"))
        if a.oracle:
            try:
                got = as_tuples(O.primer_match(O.normalize_fasta(fasta.encode("latin-1")), primers,
                                               index="auto" if select == 0 else "filter_bitvec", **params))
                ok = got == as_tuples(want)
                why = None if ok else "records differ: %d against %d" % (len(got), len(want))
            except Exception as e:   # noqa: BLE001
                ok, why = False, "%s: %s" % (type(e).__name__, e)
            if not ok:
                failures.append(dict(seed=seed, path="oracle", kind=kind, select=select, why=why))
            done += 1
            records += len(want)
            by_kind[kind] = by_kind.get(kind, 0) + 1
            continue
        for sp_name, sp in (("auto", pkg.PM_SCAN_AUTO), ("planes", pkg.PM_SCAN_PLANES)):
            try:
                got = as_tuples(pkg.primer_match(st, primers, scan_path=sp, index=index, **params))
                ok = got == as_tuples(want)
                why = None if ok else "records differ: %d against %d" % (len(got), len(want))
            except Exception as e:   # noqa: BLE001
                ok, why = False, "%s: %s" % (type(e).__name__, e)
            if not ok:
                failures.append(dict(seed=seed, path=sp_name, kind=kind, select=select, why=why))
        done += 1
        records += len(want)
        by_kind[kind] = by_kind.get(kind, 0) + 1
    print(json.dumps(dict(first=a.first, cases=done, reference_aborts=skipped, records=records,
                          seconds=round(time.time() - t0, 1), failures=failures, by_kind=by_kind)))


def other_modes(a):
    import dataclasses
    import casegen
    import make_constraints_golden
    import make_pcr_golden
    import pm_oracle as O
    from conftest import as_tuples
    pkg = importlib.import_module("sequence-alignment-tools_b200")
    tmp = tempfile.mkdtemp()
    t0 = time.time()
    done = skipped = records = 0
    failures = []
    for seed in range(a.first, a.first + a.count):
        if time.time() - t0 > a.seconds:
            break
        rng = random.Random(seed * 13 + 5)
        path = os.path.join(tmp, "f.fasta")
        if a.mode == "constraints":
            k = rng.choice([1, 1, 2, 3])
            indels = rng.random() < 0.5
            wc = rng.random() < 0.25
            params = dict(k=k, indels=indels, wc=wc, tn=wc and rng.random() < 0.5, rev_comp=rng.random() < 0.7)
            kw = dict(nseq=rng.choice([1, 3]), seqlen=rng.choice([400, 1500, 6000]), nprimers=rng.choice([2, 6, 12]),
                      plen=rng.choice([(14, 22), (20, 20), (25, 32)]), k=k, indels=indels, wc=wc,
                      text_iupac=rng.choice([0, 0, 0.02]), lower=0, repeats=rng.random() < 0.4, line=60, crlf=False)
            fasta, primers = casegen.make_case(seed, **kw)
            cons = dict(stlen=0, edlen=0, fplen=0, tplen=0)
            which = rng.choice(["fplen", "tplen", "stlen", "edlen", "both"])
            if which == "both":
                cons["fplen"], cons["tplen"] = rng.choice([3, 6, 8]), rng.choice([3, 6, 9])
            else:
                cons[which] = rng.choice([3, 6, 8, 10, -4, -8])
            select = rng.choice([0, 5])
            with open(path, "w", newline="") as f:
                f.write(fasta)
            try:
                want = as_tuples(O.run_reference(path, primers, select=select,
                                                 extra=make_constraints_golden.cli_constraints(cons), **params))
            except RuntimeError:
                skipped += 1
                continue
            st = pkg.normalize_fasta(fasta.encode("latin-1"))
            index = pkg.PM_INDEX_AUTO if select == 0 else pkg.PM_INDEX_FILTER_BITVEC

            def run(sp):
                if a.oracle:
                    return as_tuples(O.primer_match(O.normalize_fasta(fasta.encode("latin-1")), primers,
                                                    index="auto" if select == 0 else "filter_bitvec", **params, **cons))
                return as_tuples(pkg.primer_match(st, primers, scan_path=sp, index=index, **params, **cons))
            label = dict(params=params, cons=cons, select=select)
        else:
            k = rng.choice([0, 1, 1, 2, 2, 3])
            indels = rng.random() < 0.5
            wc = rng.random() < 0.2
            rev_comp = rng.random() < 0.7
            params = dict(k=k, indels=indels, wc=wc, tn=wc and rng.random() < 0.5, rev_comp=rev_comp,
                          select=rng.choice([0, 0, 5]) if k else 0,
                          mindist=rng.choice([0, 0, 0, 60, 200]), maxdist=rng.choice([2000, 2000, 700, 250]),
                          betweenlen=rng.random() < 0.25, allorient=rng.random() < 0.35,
                          stlen=0, edlen=0, fplen=0, tplen=0)
            if k and rng.random() < 0.3:
                params[rng.choice(["stlen", "edlen", "fplen", "tplen"])] = rng.choice([3, 6, 8, -5])
            fasta, primers = make_pcr_golden.make_pcr_case(seed, k, indels, wc, rev_comp)
            with open(path, "w", newline="") as f:
                f.write(fasta)
            try:
                want = [dataclasses.astuple(r) for r in O.parse_pcr_records(O.run_reference_pcr(path, primers, **params))]
            except RuntimeError:
                skipped += 1
                continue
            st = pkg.normalize_fasta(fasta.encode("latin-1"))
            par = dict(params)
            select = par.pop("select")
            index = pkg.PM_INDEX_AUTO if select == 0 else pkg.PM_INDEX_FILTER_BITVEC

            def run(sp):
                if a.oracle:
                    return [dataclasses.astuple(r) for r in O.pcr_match(O.normalize_fasta(fasta.encode("latin-1")), primers,
                                                                        index="auto" if select == 0 else "filter_bitvec", **par)]
                return [dataclasses.astuple(r) for r in pkg.pcr_match(st, primers, index=index, sts_sizes=None, **par)]
            label = dict(params=params)
        for sp_name, sp in (("auto", pkg.PM_SCAN_AUTO), ("planes", pkg.PM_SCAN_PLANES)):
            if (a.mode == "pcr" or a.oracle) and sp_name == "planes":
                continue                       # pcr_match() has no scan-path argument; the oracle has one path
            try:
                got = run(sp)
                ok = got == want
                why = None if ok else "records differ: %d against %d" % (len(got), len(want))
            except Exception as e:   # noqa: BLE001
                ok, why = False, "%s: %s" % (type(e).__name__, e)
            if not ok:
                failures.append(dict(seed=seed, path=sp_name, why=why, **label))
        done += 1
        records += len(want)
    print(json.dumps(dict(mode=a.mode, first=a.first, cases=done, reference_aborts=skipped, records=records,
                          seconds=round(time.time() - t0, 1), failures=failures)))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""File -> stdout through the drop-in command line tool: a synthetic FASTA database (and the
.sqz database compress_seq would write for it) in /dev/shm, bin/primer_match on it with the
headline workload, stage times from PM_TIMING.  One JSON line.
  python tools/cli_timing.py --bases 3000000000"""
import argparse
import json
import os
import struct
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bases", type=int, default=3_000_000_000)
    ap.add_argument("--dir", default="/dev/shm")
    a = ap.parse_args()
    args = argparse.Namespace(host_mem="pinned", bases=0)
    B = bench.Bench(args)
    torch = B.torch
    cfg = bench.CONFIGS["headline"]
    primers, pats = B.patterns("headline")
    d = B.make_database(a.bases, 1234)
    B.plant(d, pats, cfg["k"], 99)
    base = os.path.join(a.dir, "pm_cli_db.fasta")
    # the FASTA file
    d_fa = B.fasta_of(d)
    fa_bytes = int(d_fa.numel())
    h = torch.empty(fa_bytes, dtype=torch.uint8, pin_memory=True)
    h.copy_(d_fa)
    torch.cuda.synchronize()
    h.numpy().tofile(base)
    del d_fa, h
    # the compressed database of compress_seq -z: .sqz + .tbz, sequence index .idb, headers .hdr
    d_sqz = B.pkg.sqz_encode_torch(d, B.sqz_table)
    d_sqz.cpu().numpy().tofile(base + ".sqz")
    with open(base + ".tbz", "wb") as f:
        f.write(B.sqz_table)
    n = int(d.numel())
    keys = list(range(1, n - 1, bench.SEQ_LEN))
    hdr = b"".join(b"synthetic%d random DNA\n" % i for i in range(len(keys)))
    offs, at = [], 0
    for i in range(len(keys)):
        offs.append(at)
        at += len(b"synthetic%d random DNA\n" % i)
    with open(base + ".hdr", "wb") as f:
        f.write(hdr)
    with open(base + ".idb", "wb") as f:
        f.write(struct.pack("<Q", len(keys) + 1))
        for k, o in zip(keys, offs):
            f.write(struct.pack("<qq", k, o))
        f.write(struct.pack("<qq", n + 1, len(hdr)))
    del d, d_sqz
    torch.cuda.empty_cache()
    pf = os.path.join(a.dir, "pm_cli_primers.txt")
    with open(pf, "w") as f:
        f.write("\n".join(primers) + "\n")
    exe = os.path.join(ROOT, bench.PKG, "bin", "primer_match")
    out = {"bases": a.bases, "fasta_bytes": fa_bytes, "runs": []}
    for dbind, label in ((1, "FASTA file"), (4, ".sqz database")):
        for rep in range(2):
            t0 = time.perf_counter()
            p = subprocess.run([exe, "-i", base, "-P", pf, "-K", "2", "-r", "-D", str(dbind), "-A", bench.REF_FORMAT],
                               capture_output=True, env=dict(os.environ, PM_TIMING="1"), timeout=600)
            dt = time.perf_counter() - t0
            tim = None
            for ln in p.stderr.decode().splitlines():
                if ln.startswith("PM_TIMING "):
                    tim = json.loads(ln[len("PM_TIMING "):])
            out["runs"].append({"input": label, "rc": p.returncode, "wall_s": dt, "stdout_lines": p.stdout.count(b"\n"),
                                "stages": tim, "stderr_tail": p.stderr.decode()[-200:] if p.returncode else ""})
    for f in (base, base + ".sqz", base + ".tbz", base + ".hdr", base + ".idb", pf):
        try:
            os.unlink(f)
        except OSError:
            pass
    print(json.dumps(out))


if __name__ == "__main__":
    main()

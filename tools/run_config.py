#!/usr/bin/env python
"""One BASELINE configuration, resident leg only: the command ncu wraps.
  python tools/run_config.py --config edit25_k3 --bases 1000000000 --steps 2 --warmup 1"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="headline", choices=list(bench.CONFIGS))
    ap.add_argument("--bases", type=int, default=1_000_000_000)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--plant", type=int, default=64, help="planted primer copies (hit-dense runs: 1000000)")
    a = ap.parse_args()
    args = argparse.Namespace(host_mem="wc", bases=0)
    B = bench.Bench(args)
    cfg = bench.CONFIGS[a.config]
    primers, pats = B.patterns(a.config)
    d = B.make_database(a.bases, 1234)
    if a.plant > 10000:
        B.plant_dense(d, pats, cfg["k"], 99, a.plant)
    else:
        B.plant(d, pats, cfg["k"], 99, count=a.plant)
    sc = B.scanner(cfg, pats)
    out = []

    def step():
        sc.set_text(d)
        ptr, nh = sc.find_patterns_device()
        out.append((nh, sc.stats()))

    ms = B.timed(step, a.steps, a.warmup)
    nh, st = out[-1]
    print(json.dumps({"config": a.config, "bases": a.bases, "ms_per_step": ms / a.steps,
                      "gbp_s": (a.bases + 1) * a.steps / (ms * 1e-3) / 1e9, "hits": nh,
                      "stats": {k: st[k] for k in ("scan_ms", "pack_ms", "total_ms", "n_seeds", "n_candidates", "n_chains",
                                                    "kernel_launches", "host_syncs", "index_q", "index_nd", "index_entries",
                                                    "scan_warp_steps", "index_lookups")}}))
    sc.close()


if __name__ == "__main__":
    main()

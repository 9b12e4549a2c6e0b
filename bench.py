#!/usr/bin/env python
"""bench.py -- PrimerMatch database scan throughput on B200.

Metric (BASELINE.json): Gbp/s scanned, 1k 20-mer primers, k=2, both strands
(2000 patterns), bit-exact hit lists.  One "step" = one pass of the hot path
over one synthetic random-DNA database shard per GPU:

  value  stream bytes already resident in HBM -> pack kernel -> scan ->
         chains / alignment -> sorted hit list (left on the device)
  e2e    the same through the C ABI with HOST buffers: page-locked host copy of
         the database (.sqz form) -> H2D -> unpack -> pack -> scan -> ... ->
         hit records written back to host memory

One command covers every BASELINE.json configuration (`--config`, default
"all": the headline in full, the others as entries of `configs`) and both
scaling modes:

  weak    every rank scans its own shard (a multi-genome database partitioned
          at sequence boundaries); no step waits for another rank, NCCL sums
          the hit counts once per run
  strong  ONE database (3 Gbp for the headline) is cut with shard_plan() and the
          documented overlap, every rank scans its piece, the hit records are
          gathered over NCCL (all_gather of counts, then of the records) and
          rank 0 checks the gathered list byte for byte against its own scan of
          the whole database (`strong` object; `--scaling strong` makes it the
          headline value)

At N = 1 the CPU baseline runs the UNMODIFIED reference (oracle/_ref/primer_match)
on windows cut out of the benchmark shard itself and the GPU hit records inside
those windows must be identical (`parity_checked_bp`).

`--impl reference` times the reference on the box's host cores, one process per
core, each on its own slice of a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "sequence-alignment-tools_b200"

SEQ_LEN = 50_000_000          # the synthetic database: sequences of 50 Mbp
REF_FORMAT = "%f\\t%s\\t%e\\t%r\\t%i\\t%d\\n"

# BASELINE.json configs 1-4 (config 0 is the reference's own CPU case).  ref_bp: bases per
# host core of the bounded CPU sample (sized for ~5-10 s of the reference per core).
CONFIGS = {
    "headline": dict(nprimers=1000, plen=20, k=2, indels=0, iupac=0, select=0, bases=3_000_000_000, ref_bp=3_000_000,
                     metric="Gbp/s scanned, 1k 20-mer primers k=2 both strands",
                     baseline="BASELINE.json metric; configs[2] without wildcards"),
    "iupac_k2": dict(nprimers=1000, plen=20, k=2, indels=0, iupac=1, select=0, bases=3_000_000_000, ref_bp=1_500_000,
                     metric="Gbp/s scanned, 1k IUPAC 20-mer primers (-w) k=2 both strands",
                     baseline="BASELINE.json configs[2]"),
    "edit25_k3": dict(nprimers=1000, plen=25, k=3, indels=1, iupac=0, select=0, bases=3_000_000_000, ref_bp=600_000,
                      metric="Gbp/s scanned, 1k 25-mer primers k=3 edits both strands",
                      baseline="BASELINE.json configs[3]"),
    "p100_k1": dict(nprimers=100, plen=20, k=1, indels=0, iupac=0, select=0, bases=100_000_000, ref_bp=6_000_000,
                    metric="Gbp/s scanned, 100 20-mer primers k=1 both strands, 100 Mbp",
                    baseline="BASELINE.json configs[1]"),
    "oligo100k_exact": dict(nprimers=100_000, plen=20, k=0, indels=0, iupac=0, select=0, bases=3_750_000_000,
                            ref_bp=4_000_000,
                            ref_skip="the reference's keyword_tree build grows ~n^2.5 with the pattern count (0.5 s for 2 000 "
                                     "patterns, 2.6 s for 4 000, 15 s for 8 000, measured with oracle/_ref/primer_match): "
                                     "200 000 patterns do not finish in bounded time",
                            metric="Gbp/s scanned, 100k 20-mer oligos exact both strands",
                            baseline="BASELINE.json configs[4] (keyword_tree leg), 30 Gbp / 8 GPUs per shard"),
    "oligo100k_k1": dict(nprimers=100_000, plen=20, k=1, indels=0, iupac=0, select=5, bases=3_750_000_000,
                         ref_bp=20_000, metric="Gbp/s scanned, 100k 20-mer oligos k=1 (-N 5) both strands",
                         baseline="BASELINE.json configs[4] (shift_and_inexact leg), 30 Gbp / 8 GPUs per shard"),
}
# not a BASELINE configuration: the 20-mer edit scan (A/B runs of tools/run_config.py)
CONFIGS["edit20_k2"] = dict(nprimers=1000, plen=20, k=2, indels=1, iupac=0, select=0, bases=3_000_000_000, ref_bp=600_000,
                            metric="Gbp/s scanned, 1k 20-mer primers k=2 edits both strands", baseline="(none)")
# the headline workload on a database whose repeats contain primer sites: 10^6 planted copies per Gbp
CONFIGS["hit_dense"] = dict(nprimers=1000, plen=20, k=2, indels=0, iupac=0, select=0, bases=1_000_000_000, ref_bp=1_000_000,
                            metric="Gbp/s scanned, 1k 20-mer primers k=2 both strands, 10^6 hits per Gbp",
                            baseline="(none: the headline workload, hit-dense database)", dense=1_000_000)
ALL_ORDER = ["headline", "iupac_k2", "edit25_k3", "p100_k1", "oligo100k_exact", "oligo100k_k1"]


def make_primers(cfg, seed=12345):
    import numpy as np
    n, m = cfg["nprimers"], cfg["plen"]
    g = np.random.default_rng(seed)
    a = np.frombuffer(b"ACGT", np.uint8)[g.integers(0, 4, (n, m))]
    if cfg["iupac"]:  # one to three ambiguity codes per primer, like the reference's db/patwc.txt
        codes = np.frombuffer(b"RYKMSWBDHVN", np.uint8)
        for r in a:
            for _ in range(int(g.integers(1, 4))):
                r[int(g.integers(0, m))] = codes[int(g.integers(0, len(codes)))]
    return [bytes(r).decode() for r in a]


def ref_cmd(cfg, fasta, pfile, fmt=REF_FORMAT):
    ref = os.path.join(ROOT, "oracle", "_ref", "primer_match")
    return [ref, "-i", fasta, "-P", pfile, ("-k" if cfg["indels"] else "-K"), str(cfg["k"]), "-r",
            *(["-w"] if cfg["iupac"] else []), *(["-N", str(cfg["select"])] if cfg["select"] else []), "-A", fmt]


def workload_text(cfg):
    return "%d random %d-mer primers%s, k=%d %s (%s %d), both strands (-r), synthetic random-DNA database" % (
        cfg["nprimers"], cfg["plen"], " with 1-3 IUPAC codes each (-w)" if cfg["iupac"] else "", cfg["k"],
        "edits" if cfg["indels"] else "mismatches", "-k" if cfg["indels"] else "-K", cfg["k"])


def index_text(cfg):
    if cfg["select"]:
        return "-N %d" % cfg["select"]
    return "automatic (pick_pattern_index): " + ("exact index" if cfg["k"] == 0 else
                                                  "exact_halves" if (cfg["k"] == 1 and cfg["plen"] >= 8) else "filter_bitvec")


# ---------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index: int, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_ev = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._stop_ev.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def visible_gpu_index(local_rank: int) -> int:
    cvd = os.environ.get("CUDA_VISIBLE_DEVICES")
    if cvd:
        try:
            return int(cvd.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_available():
    p = os.path.join(ROOT, "oracle", "_ref", "primer_match")
    return os.path.exists(p) and os.access(p, os.X_OK)


# ---------------------------------------------------------------------------
# the reference on the host cores
def write_fasta_records(path, records):
    """records: [(name, uint8 array of bases)]; 80 bases per line."""
    with open(path, "wb") as f:
        for name, seq in records:
            f.write(b">" + name + b"\n")
            full = len(seq) // 80 * 80
            if full:
                f.write(b"\n".join(bytes(r) for r in seq[:full].reshape(-1, 80)))
                f.write(b"\n")
            if len(seq) > full:
                f.write(bytes(seq[full:]) + b"\n")


REF_DEADLINE_S = 240.0


def run_reference_files(cfg, primers, files, workdir, capture=False):
    """One timed pass of the unmodified reference, one process per file.  Returns
    (wall seconds over all processes, [stdout bytes per file] when capture).  The processes
    are killed (by PID) when they exceed REF_DEADLINE_S."""
    pfile = os.path.join(workdir, "primers_%d_%d.txt" % (cfg["nprimers"], cfg["plen"]))
    with open(pfile, "w") as f:
        f.write("\n".join(primers) + "\n")
    t0 = time.perf_counter()
    ps = [subprocess.Popen(ref_cmd(cfg, fp, pfile), stdout=subprocess.PIPE if capture else subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL) for fp in files]
    outs = []
    if capture:
        # drain the pipes concurrently (the outputs are small: a few hits per Mbp)
        bufs = [None] * len(ps)

        def drain(i, p):
            bufs[i] = p.stdout.read()
        ths = [threading.Thread(target=drain, args=(i, p)) for i, p in enumerate(ps)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        outs = bufs
    for p in ps:
        try:
            p.wait(timeout=max(1.0, REF_DEADLINE_S - (time.perf_counter() - t0)))
        except subprocess.TimeoutExpired:
            for q in ps:
                if q.poll() is None:
                    q.kill()
            raise RuntimeError("reference did not finish within %.0f s" % REF_DEADLINE_S)
    dt = time.perf_counter() - t0
    bad = [p.returncode for p in ps if p.returncode != 0]
    if bad:
        raise RuntimeError("reference exited with %s" % bad)
    return dt, outs


def run_reference_sample(cfg, primers, per_proc: int, procs: int, seed: int, workdir: str):
    """Random DNA of the workload's shape, `per_proc` bases per process.  Returns seconds."""
    import numpy as np
    g = np.random.default_rng(seed)
    files = []
    for i in range(procs):
        seq = np.frombuffer(b"ACGT", np.uint8)[g.integers(0, 4, per_proc)]
        path = os.path.join(workdir, "s%d_%d.fasta" % (seed, i))
        write_fasta_records(path, [(b"sample%d" % i, seq)])
        files.append(path)
    dt, _ = run_reference_files(cfg, primers, files, workdir)
    for fp in files:
        os.unlink(fp)
    return dt


def run_oracle_sample(cfg, primers, total_bases: int, seed: int):
    """Fallback CPU baseline: the oracle port (single thread)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pm_oracle as O
    g = np.random.default_rng(seed)
    seq = np.frombuffer(b"ACGT", np.uint8)[g.integers(0, 4, total_bases)]
    T = np.concatenate([[10], seq, [10]]).astype(np.uint8)
    st = O.Stream(T, np.array([1], np.int64), ["s"])
    t0 = time.perf_counter()
    O.primer_match(st, primers, k=cfg["k"], indels=bool(cfg["indels"]), rev_comp=True, wc=bool(cfg["iupac"]))
    return time.perf_counter() - t0


def parse_ref_records(out: bytes):
    """(%f, %s, %e, %r, %i, %d) records of REF_FORMAT."""
    recs = []
    for ln in out.decode("latin-1").splitlines():
        p = ln.split("\t")
        if len(p) != 6:
            continue
        recs.append((int(p[0]), int(p[1]), int(p[2]), p[3].strip(), int(p[4]), int(p[5])))
    return recs


def parity_windows(cfg, primers, d_stream, n_chars, hits, procs, per_proc, workdir, planted):
    """The reference on `procs` windows of `per_proc` bases cut out of the benchmark shard itself
    (some around planted primer copies, one across a sequence separator) against the GPU hit
    records that lie inside those windows.  Returns (seconds of the reference, bp, compared, ok, detail)."""
    import numpy as np
    margin = 200
    w = per_proc
    starts = []
    sep = SEQ_LEN if n_chars > SEQ_LEN + w else None
    if sep is not None:
        starts.append(max(1, sep - w // 2))                       # across a sequence separator
    starts.append(1)                                              # the start of the database
    for at in planted:                                            # around planted copies
        a = max(1, min(n_chars - 1 - w, at - w // 3))
        if all(abs(a - b) >= w for b in starts):
            starts.append(a)
        if len(starts) >= procs:
            break
    j = 0
    while len(starts) < procs:                                    # the rest spread over the shard
        a = 1 + (j * 7919 * 104729) % max(1, n_chars - 1 - w)
        j += 1
        if j > 20000 or all(abs(a - b) >= w for b in starts):
            starts.append(a)
    files, metas = [], []
    for i, a in enumerate(starts):
        b = min(n_chars - 1, a + w)
        win = d_stream[a:b].cpu().numpy()
        eos = np.flatnonzero(win == 10)
        cuts = [-1] + [int(x) for x in eos] + [len(win)]
        recs, g = [], []
        for c0, c1 in zip(cuts, cuts[1:]):
            if c1 - c0 - 1 > 0:
                recs.append((b"w%d_%d" % (i, len(recs)), win[c0 + 1:c1]))
                g.append(a + c0 + 1)                              # stream index of the record's first base
        path = os.path.join(workdir, "win%d.fasta" % i)
        write_fasta_records(path, recs)
        files.append(path)
        metas.append((a, b, g))
    dt, outs = run_reference_files(cfg, primers, files, workdir, capture=True)
    for fp in files:
        os.unlink(fp)
    npr = cfg["nprimers"]
    want, got = set(), set()
    for (a, b, g), out in zip(metas, outs):
        for (f, s, e, r, i, d) in parse_ref_records(out):
            gs, ge = s + g[f - 1], e + g[f - 1]
            if gs >= a + margin and ge <= b - margin:
                want.add((gs, ge, r, i, d))
        sel = hits[(hits["start"] >= a + margin) & (hits["end"] <= b - margin)]
        for h in sel:
            pid = int(h["pattern_id"])
            rc = pid > npr
            got.add((int(h["start"]), int(h["end"]), "R" if rc else "F", pid - (npr if rc else 0), int(h["editdist"])))
    bp = sum(b - a for a, b, _ in metas)
    ok = want == got
    detail = None
    if not ok:
        detail = {"only_reference": sorted(want - got)[:5], "only_gpu": sorted(got - want)[:5]}
    return dt, bp, len(want), ok, detail


# ---------------------------------------------------------------------------
def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = "headline" if args.config == "all" else args.config
    cfg = dict(CONFIGS[name])
    if cfg.get("ref_skip"):
        print(json.dumps({"impl": "reference", "unavailable": "config %s: %s" % (name, cfg["ref_skip"])}))
        return
    primers = make_primers(cfg)
    cores = cpu_cores()
    if reference_available():
        kind = "reference"
        procs = cores
        per_proc = args.ref_bases_per_core or cfg["ref_bp"]
        work = tempfile.mkdtemp(prefix="pmref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        for w in range(args.warmup):
            run_reference_sample(cfg, primers, max(2000, per_proc // 8), procs, 100 + w, work)
        t = 0.0
        for s in range(args.steps):
            t += run_reference_sample(cfg, primers, per_proc, procs, 200 + s, work)
        total = per_proc * procs * args.steps
        sample = "%d processes x %d bp per step, %d steps (throughput of the streaming scan, not the %d bp shard)" % (
            procs, per_proc, args.steps, cfg["bases"])
    else:
        kind = "port"
        procs = 1
        per = 200_000 if cfg["nprimers"] <= 1000 else 2000
        for w in range(min(args.warmup, 1)):
            run_oracle_sample(cfg, primers, per // 10, 100 + w)
        t = 0.0
        for s in range(args.steps):
            t += run_oracle_sample(cfg, primers, per, 200 + s)
        total = per * args.steps
        sample = "oracle port, 1 thread x %d bp per step, %d steps" % (per, args.steps)
    value = total / t / 1e9
    conf = config_dict(cfg, cfg["bases"], args)
    conf["reference_sample"] = sample
    line = {"impl": "reference", "metric": cfg["metric"], "value": value, "unit": "Gbp/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": conf,
            "cpu_baseline": {"value": value, "unit": "Gbp/s", "cores": procs, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def config_dict(cfg, shard_bases, args):
    return {"workload": workload_text(cfg), "shard_bases_per_gpu": shard_bases, "patterns": 2 * cfg["nprimers"],
            "index": index_text(cfg), "baseline_config": cfg["baseline"],
            "l2": "inputs larger than L2 (%.2f GB stream + %.2f GB packed text per pass)" % (shard_bases / 1e9, shard_bases / 2e9),
            "partition": "one shard per GPU, cut at sequence boundaries (weak); one database cut by shard_plan with overlap (strong)",
            "host_mem": args.host_mem}


# ---------------------------------------------------------------------------
class Bench:
    """Everything one rank needs: the synthetic database on the device, its host copies,
    the timing helpers."""

    def __init__(self, args):
        import numpy as np
        import torch
        import torch.distributed as dist
        self.np, self.torch, self.dist, self.args = np, torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        # this rank's threads (and with them the pages of its page-locked buffers, first touch) on
        # the CPUs next to its GPU: NVML knows them
        self.affinity = None
        self.all_cpus = os.sched_getaffinity(0)
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(visible_gpu_index(self.local_rank))
            pynvml.nvmlDeviceSetCpuAffinity(h)
            self.affinity = {"cpus": len(os.sched_getaffinity(0)), "set_by": "nvmlDeviceSetCpuAffinity"}
        except Exception as ex:
            self.affinity = {"error": str(ex)[:100]}
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.pkg = importlib.import_module(PKG)
        self.pkg.lib()
        self.sharding = importlib.import_module(PKG + ".sharding")
        self.stream = torch.cuda.current_stream()
        self.launches = 0
        self.sqz_table = b"ACGT\n"
        self._pats = {}

    def patterns(self, name):
        """(primers, pattern array with the reverse complements) of a configuration, cached."""
        if name not in self._pats:
            primers = make_primers(CONFIGS[name])
            self._pats[name] = (primers, self.pkg.pattern_array(primers, True))
        return self._pats[name]

    # -- data ---------------------------------------------------------------
    def make_database(self, bases, seed):
        """eos, random ACGT in sequences of SEQ_LEN, eos: bases + 1 characters with an
        end-of-sequence character at every multiple of SEQ_LEN (generated on the device)."""
        torch = self.torch
        n = bases + 1
        gen = torch.Generator(device=self.dev)
        gen.manual_seed(seed)
        lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=self.dev)
        d = torch.empty(n, dtype=torch.uint8, device=self.dev)
        chunk = 1 << 28
        for a in range(0, n, chunk):
            b = min(n, a + chunk)
            d[a:b] = lut[torch.randint(0, 4, (b - a,), device=self.dev, generator=gen, dtype=torch.uint8).long()]
        d[0::SEQ_LEN] = 10
        d[n - 1] = 10
        return d

    def plant(self, d, pats, k, seed, count=64, limit=None):
        """A few exact / mutated primer copies so that the hit path is exercised."""
        np, torch = self.np, self.torch
        rs = np.random.default_rng(seed)
        n = int(limit or d.numel())
        at_list = []
        for _ in range(count):
            p = bytearray(pats[int(rs.integers(0, len(pats)))].encode())
            p = bytearray(c if c in b"ACGT" else b"ACGT"[int(rs.integers(0, 4))] for c in p)
            for _ in range(int(rs.integers(0, k + 1))):
                p[int(rs.integers(0, len(p)))] = b"ACGT"[int(rs.integers(0, 4))]
            at = int(rs.integers(1000, n - 1000))
            if (at % SEQ_LEN) > SEQ_LEN - 100 or (at % SEQ_LEN) < 100:
                continue
            d[at:at + len(p)] = torch.tensor(list(p), dtype=torch.uint8, device=self.dev)
            at_list.append(at)
        return at_list

    def plant_dense(self, d, pats, k, seed, count, limit=None):
        """`count` primer copies (up to k substitutions each) at evenly spread positions, written with
        one scatter: the hit-dense shape of a genome whose repeats contain primer sites."""
        torch = self.torch
        n = int(limit or d.numel())
        m = len(pats[0])
        g = torch.Generator(device=self.dev)
        g.manual_seed(seed)
        pm_ = torch.tensor([list(p.encode()) for p in pats if len(p) == m], dtype=torch.uint8, device=self.dev)
        lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=self.dev)
        pm_ = torch.where(((pm_ == 65) | (pm_ == 67) | (pm_ == 71) | (pm_ == 84)), pm_, torch.full_like(pm_, 65))
        step = max(m + 8, (n - 4000) // count)
        pos = 2000 + torch.arange(count, device=self.dev, dtype=torch.int64) * step
        pos = pos[(pos % SEQ_LEN > 200) & (pos % SEQ_LEN < SEQ_LEN - 200) & (pos < n - 200)]
        which = torch.randint(0, pm_.shape[0], (pos.numel(),), device=self.dev, generator=g)
        copies = pm_[which].clone()
        for _ in range(k):      # up to k substitutions (a drawn base may equal the original)
            col = torch.randint(0, m, (pos.numel(),), device=self.dev, generator=g)
            base = lut[torch.randint(0, 4, (pos.numel(),), device=self.dev, generator=g)]
            hit = torch.rand(pos.numel(), device=self.dev, generator=g) < 0.5
            rows = torch.arange(pos.numel(), device=self.dev)[hit]
            copies[rows, col[hit]] = base[hit]
        d[(pos[:, None] + torch.arange(m, device=self.dev)[None, :]).reshape(-1)] = copies.reshape(-1)
        return int(pos.numel())

    def host_copy(self, d_t):
        """The tensor's bytes in page-locked host memory (what a pinned file reader would hold)."""
        torch = self.torch
        if self.args.host_mem == "wc":
            hb = self.pkg.HostBuffer(d_t.numel(), write_combined=True)
            torch.from_numpy(hb.array).copy_(d_t)
            return hb
        h = torch.empty(d_t.numel(), dtype=torch.uint8, pin_memory=True)
        h.copy_(d_t)
        return h

    def fasta_of(self, d_stream):
        """The FASTA file the stream came from: one record per sequence, 80 bases per line."""
        torch = self.torch
        n = d_stream.numel()
        parts = []
        bounds = list(range(0, n - 1, SEQ_LEN)) + [n - 1]
        nl = torch.tensor([10], dtype=torch.uint8, device=self.dev)
        for si in range(len(bounds) - 1):
            a, b = bounds[si] + 1, bounds[si + 1]
            if b <= a:
                continue
            body = d_stream[a:b]
            full = (b - a) // 80 * 80
            lines = torch.cat([body[:full].view(-1, 80), nl.expand(full // 80, 1)], dim=1).reshape(-1)
            parts += [torch.tensor(list(b">synthetic%d random DNA\n" % si), dtype=torch.uint8, device=self.dev), lines]
            if b - a > full:
                parts += [body[full:], nl]
        return torch.cat(parts)

    # -- timing -------------------------------------------------------------
    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def sum_over_ranks(self, v):
        if self.world == 1:
            return int(v)
        t = self.torch.tensor([int(v)], dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(t)                  # NCCL: the hit counts of all shards, summed
        return int(t.item())

    def timed(self, fn, steps, warmup):
        """W warm-up calls, then `steps` calls between two CUDA events on the stream the
        library launches on, bracketed by barrier + synchronize.  Returns ms (this rank)."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            fn()
        e1.record(self.stream)
        self.barrier()
        return e0.elapsed_time(e1)

    def scanner(self, cfg, pats):
        sc = self.pkg.Scanner(k=cfg["k"], indels=bool(cfg["indels"]), wildcards=bool(cfg["iupac"]),
                              device=self.local_rank, index=cfg["select"])
        for i, p in enumerate(pats):
            sc.add_pattern(p, i + 1)
        sc.init()
        sc.set_stream(self.stream.cuda_stream)
        return sc


def kernel_name(cfg, st):
    if st["scan_path"] != 2:
        return ("hamming_scan_kernel<%d>" % cfg["k"]) if not cfg["indels"] else (
            "edit_seed_kernel<%d> + edit_verify_kernel<%d>" % (cfg["k"], cfg["k"]))
    if cfg["indels"] and cfg["k"]:
        return "pair_scan_edit_kernel<%d, %d> + index_verify_kernel<%d>" % (st["index_q"], cfg["k"], cfg["k"])
    ent = st["index_entries"]
    # third template argument: the level-1b variant (24 000 < entries <= 100 000, csrc/pm_b200.cu)
    return "pair_scan_sub_kernel<%d, %d, %s>" % (st["index_q"], st["index_nd"], "true" if 24000 < ent <= 100000 else "false")


def roofline_of(B, cfg, sc, n_chars, n_pats, scan_ms, clocks, peaks):
    """The dominant kernel (the scan) against both roofs; the binding one (the slower of the
    integer pipe and HBM, per the north star) is the top-level `bound`."""
    pkg = B.pkg
    st = sc.stats()
    on_index = st["scan_path"] == pkg.PM_SCAN_INDEX
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    cells = float(n_chars) * n_pats * (cfg["k"] + 1)
    if on_index:
        alg_bytes = n_chars * 0.5      # the scan reads the packed text once: one uint2 per 16 characters
        # dram__bytes_read+write of one launch over 3.0e9 bases: 1504.0 + 6.3 MB (ncu --set full,
        # profiles/README.md, prof_pairsub_r1): a constant taken from that capture, scaled by the
        # bases of this launch -- not measured in this run
        traffic = 0.5034 * n_chars if (not cfg["indels"] and cfg["nprimers"] <= 1000) else None
        tsrc = "constant from the ncu --set full capture prof_pairsub_r1 (profiles/README.md): bytes per base x bases; not measured in this run"
    else:
        n_planes = 5 if not cfg["iupac"] else 16
        alg_bytes = n_chars / 8.0 * n_planes * (1088.0 / 1024.0)
        traffic, tsrc = None, None
    hbm_ach = alg_bytes / (scan_ms * 1e-3) / 1e9
    hbm = {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
           "algorithmic_bytes_per_base": alg_bytes / n_chars,
           "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s"}
    try:
        ip = pkg.int_pipe_report(sc, scan_ms, B.local_rank)
    except Exception as ex:  # diagnostic only
        ip = {"error": str(ex)[:200]}
    roof = {"bound": "int_pipe", "achieved": ip.get("achieved"), "peak": ip.get("peak"), "unit": "Tlop/s",
            "frac": ip.get("frac"), "traffic": traffic, "traffic_source": tsrc,
            "kernel": kernel_name(cfg, st), "kernel_ms": scan_ms,
            "note": "bit-parallel integer work: the integer pipe binds (north star: the slower of the two roofs); "
                    "the HBM view of the same kernel is under `hbm`",
            "int_pipe": ip, "hbm": hbm, "cell_updates_per_s": cells / (scan_ms * 1e-3),
            "scan_path": "index" if on_index else "planes"}
    if on_index:
        roof["index"] = {"piece_len": st["index_q"], "pieces": st["index_pieces"],
                         "distances": st["index_nd"], "entries": st["index_entries"]}
    return roof


def measure_config(B, name, d_db, h_sqz, steps, warmup, full, clocks_out=None):
    """One configuration on this rank's shard: resident leg, end-to-end leg (.sqz host buffer),
    and for `full` the stream / FASTA host legs.  Returns the (rank-0) result dict."""
    np, torch, pkg, args = B.np, B.torch, B.pkg, B.args
    cfg = dict(CONFIGS[name])
    if args.bases and name == "headline":
        cfg["bases"] = args.bases
    bases = min(cfg["bases"], d_db.numel() - 1)
    n_chars = bases + 1
    primers, pats = B.patterns(name)
    d_stream = d_db[:n_chars]
    sc = B.scanner(cfg, pats)
    scan_ms, res = [], {}

    def step_resident():
        sc.set_text(d_stream)                      # pack kernel (1 launch)
        ptr, nh = sc.find_patterns_device()
        st = sc.stats()
        B.launches += 1 + st["kernel_launches"]
        scan_ms.append(st["scan_ms"])
        res["nh"], res["st"] = nh, st

    sqz_bytes = (3 * n_chars + 7) // 8
    h_sqz_view = None
    if h_sqz is not None:
        h_sqz_view = h_sqz.array[:sqz_bytes] if hasattr(h_sqz, "array") else h_sqz[:sqz_bytes]

    def step_e2e_sqz():
        sc.set_text_sqz(h_sqz_view, B.sqz_table, n_chars)   # H2D of the .sqz bytes + unpack + pack
        res["hits"] = sc.find_patterns()                    # ... + the hit records written to host memory
        B.launches += 1 + sc.stats()["kernel_launches"]

    sampler = ClockSampler(visible_gpu_index(B.local_rank)) if clocks_out is not None else None
    if sampler:
        sampler.start()
    B.launches = 0
    for _ in range(warmup):
        step_resident()
    if clocks_out is not None:
        # the headline is measured first: keep warming up until the GPUs of ALL ranks have run for
        # ~0.7 s (clocks and power state settled; with 8 GPUs starting from idle at once, three 6 ms
        # steps are not enough and the ramp would sit inside the timed region)
        t_w = time.perf_counter()
        n_w = 0
        while time.perf_counter() - t_w < 0.7 and n_w < 200:
            step_resident()
            n_w += 1
        res["extra_warmup_steps"] = n_w
    B.launches = 0
    scan_ms.clear()
    ms = B.timed(step_resident, steps, 0)
    n_launch = B.launches
    res["total"] = B.sum_over_ranks(res["nh"])     # NCCL: the hit counts of all shards, summed (once per run:
    #                                                the shards are independent, no step waits for another rank)
    if sampler:
        clocks_out.update(sampler.stop())
    mean_scan = sum(scan_ms) / max(1, len(scan_ms))
    if h_sqz_view is not None:
        ms_e2e = B.timed(step_e2e_sqz, steps, min(warmup, 2))
    else:   # no host leg: the hit records of the resident text, for the parity check
        sc.set_text(d_stream)
        res["hits"] = sc.find_patterns()
        ms_e2e = 0.0
    hits = res["hits"]
    legs = {}
    if full:
        h_stream = B.host_copy(d_stream)
        d_fasta = B.fasta_of(d_stream)
        h_fasta = B.host_copy(d_fasta)
        fasta_bytes = int(d_fasta.numel())
        del d_fasta

        def step_stream():
            sc.set_text(h_stream)
            res["hits_stream"] = sc.find_patterns()

        def step_fasta():
            sc.set_text_fasta(h_fasta)
            res["hits_fasta"] = sc.find_patterns()

        legs["stream"] = (B.timed(step_stream, steps, min(warmup, 2)), n_chars)
        legs["fasta"] = (B.timed(step_fasta, steps, min(warmup, 2)), fasta_bytes)
        assert res["hits_stream"].tobytes() == hits.tobytes() == res["hits_fasta"].tobytes(), "the three input forms disagree"
        del h_stream, h_fasta
    copy_gbs = None
    if full and h_sqz_view is not None:
        # the host-to-device copy alone, all ranks at once: what the uplinks give this rank while its
        # neighbours copy too (the end-to-end legs are bound by it; the ratio says how close they come)
        h_t = torch.from_numpy(h_sqz_view) if not hasattr(h_sqz_view, "is_pinned") else h_sqz_view
        d_tmp = torch.empty(int(sqz_bytes), dtype=torch.uint8, device=B.dev)

        def step_copy():
            d_tmp.copy_(h_t, non_blocking=True)
        try:
            ms_copy = B.timed(step_copy, steps, 1)
            copy_gbs = sqz_bytes * steps / (ms_copy * 1e-3) / 1e9
        except Exception:
            copy_gbs = None
        del d_tmp
    ms, ms_e2e, mean_scan = B.max_over_ranks([ms, ms_e2e, mean_scan])
    leg_ms = {k: B.max_over_ranks([v[0]])[0] for k, v in legs.items()}
    out = None
    if B.rank == 0:
        world = B.world
        bases_total = float(n_chars) * world * steps
        st = res["st"]
        out = {"metric": cfg["metric"], "value": bases_total / (ms * 1e-3) / 1e9, "unit": "Gbp/s",
               "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
               "config": config_dict(cfg, bases, args), "hits_per_step": res["total"],
               "host_syncs_per_call": st["host_syncs"], "extra_warmup_steps": res.get("extra_warmup_steps", 0),
               "cell_updates_per_s": float(n_chars) * world * len(pats) * (cfg["k"] + 1) / (ms / steps * 1e-3),
               "tail_ms_per_step": st["total_ms"] - st["scan_ms"],
               "e2e": None if h_sqz_view is None else {"value": bases_total / (ms_e2e * 1e-3) / 1e9, "unit": "Gbp/s",
                       "h2d_bytes_per_step": int(sqz_bytes) * world,
                       "d2h_bytes_per_step": (hits.nbytes + 256) * world, "ms_per_step": ms_e2e / steps,
                       "h2d_gbs_per_rank": sqz_bytes / (ms_e2e / steps * 1e-3) / 1e9,
                       "h2d_copy_only_gbs_this_rank": copy_gbs,
                       "input": "pinned host buffer holding the database as compress_seq -z stores it "
                                "(.sqz, 3 bits per character; pm_set_text_sqz_host)"},
               "gpu_launches": n_launch}
        if full:
            out["e2e_stream"] = {"value": bases_total / (leg_ms["stream"] * 1e-3) / 1e9, "unit": "Gbp/s",
                                 "h2d_bytes_per_step": n_chars * world, "ms_per_step": leg_ms["stream"] / steps,
                                 "h2d_gbs_per_rank": n_chars / (leg_ms["stream"] / steps * 1e-3) / 1e9,
                                 "input": "pinned host buffer holding the normalised character stream, 1 byte per character (pm_set_text_host)"}
            out["e2e_fasta"] = {"value": bases_total / (leg_ms["fasta"] * 1e-3) / 1e9, "unit": "Gbp/s",
                                "h2d_bytes_per_step": legs["fasta"][1] * world, "ms_per_step": leg_ms["fasta"] / steps,
                                "h2d_gbs_per_rank": legs["fasta"][1] / (leg_ms["fasta"] / steps * 1e-3) / 1e9,
                                "input": "pinned host buffer holding the FASTA file (80 bases per line), normalised on the "
                                         "device chunk by chunk while the copy runs (pm_set_text_fasta_host); pack and scan follow"}
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        out["roofline"] = roofline_of(B, cfg, sc, n_chars, len(pats), mean_scan, clocks_out or {}, peaks)
    res["cfg"], res["primers"], res["n_chars"] = cfg, primers, n_chars
    sc.close()
    return out, res


def cpu_baseline_leg(B, res, d_db, planted, line, per_proc=None):
    """N = 1, rank 0: the reference on windows of the shard, GPU records compared inside them."""
    cfg, primers, n_chars = res["cfg"], res["primers"], res["n_chars"]
    try:
        os.sched_setaffinity(0, B.all_cpus)       # the reference gets every host core, not only the GPU's neighbours
    except Exception:
        pass
    cores = cpu_cores()
    if cfg.get("ref_skip"):
        line["cpu_baseline"] = {"value": None, "unit": "Gbp/s", "cores": cores, "kind": "reference", "sample": "not run: " + cfg["ref_skip"]}
        return True
    try:
        if reference_available():
            work = tempfile.mkdtemp(prefix="pmref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            per = per_proc or cfg["ref_bp"]
            per = min(per, max(1000, (n_chars - 2) // (2 * cores)))
            dt, bp, compared, ok, detail = parity_windows(cfg, primers, d_db, n_chars, res["hits"], cores, per, work, planted)
            line["cpu_baseline"] = {"value": bp / dt / 1e9, "unit": "Gbp/s", "cores": cores, "kind": "reference",
                                    "sample": "%d processes x %d bp windows of the benchmark shard itself, %.1f s" % (cores, per, dt)}
            line["parity_checked_bp"] = bp
            line["parity_records_compared"] = compared
            line["parity_ok"] = ok
            if not ok:
                line["parity_detail"] = detail
            return ok
        dt = run_oracle_sample(cfg, primers, 200_000 if cfg["nprimers"] <= 1000 else 2000, 300)
        line["cpu_baseline"] = {"value": (200_000 if cfg["nprimers"] <= 1000 else 2000) / dt / 1e9, "unit": "Gbp/s",
                                "cores": 1, "kind": "port", "sample": "oracle port, %.1f s" % dt}
    except Exception as ex:
        line["cpu_baseline"] = {"error": str(ex)[:300]}
    return True


def strong_scaling(B, cfg_name, steps, warmup):
    """ONE database of cfg.bases, cut with shard_plan + overlap; every rank scans its piece
    (resident and from its .sqz host shard) and the hit records are gathered over NCCL.  Rank 0
    then scans the whole database alone and compares byte for byte."""
    np, torch, pkg, args, dist = B.np, B.torch, B.pkg, B.args, B.dist
    cfg = dict(CONFIGS[cfg_name])
    if args.bases and cfg_name == "headline":
        cfg["bases"] = args.bases
    bases = cfg["bases"]
    n_chars = bases + 1
    primers, pats = B.patterns(cfg_name)
    d_db = B.make_database(bases, 777)           # the same database on every rank
    B.plant(d_db, pats, cfg["k"], 778)
    sc = B.scanner(cfg, pats)
    plan = B.sharding.shard_plan(n_chars, B.world, sc.shard_overlap())
    sh = plan[B.rank]
    lo = sh.load_lo & ~15                        # a .sqz shard starts on a byte boundary of the bit stream
    d_shard = d_db[lo:sh.load_hi]
    h_sqz = B.host_copy(pkg.sqz_encode_torch(d_shard, B.sqz_table))
    nsh = int(d_shard.numel())
    sqz_bytes = (3 * nsh + 7) // 8
    h_view = h_sqz.array[:sqz_bytes] if hasattr(h_sqz, "array") else h_sqz[:sqz_bytes]
    item = pkg.HIT_DTYPE.itemsize
    res = {}

    def step_resident():
        sc.set_text(d_shard, offset=lo)
        ptr, nh = sc.find_patterns_device(sh.begin, sh.end)
        counts, recs = B.sharding.gather_hits_device(ptr, nh, item, dist if B.world > 1 else None, B.dev)
        res["counts"], res["recs"] = counts, recs

    def step_e2e():
        sc.set_text_sqz(h_view, B.sqz_table, nsh, offset=lo)
        ptr, nh = sc.find_patterns_device(sh.begin, sh.end)
        counts, recs = B.sharding.gather_hits_device(ptr, nh, item, dist if B.world > 1 else None, B.dev)
        res["host"] = recs.cpu() if B.rank == 0 else None      # the gathered records on rank 0's host

    ms = B.timed(step_resident, steps, warmup)
    ms_e2e = B.timed(step_e2e, steps, min(warmup, 2))
    ms, ms_e2e = B.max_over_ranks([ms, ms_e2e])
    gathered = res["recs"].cpu().numpy().tobytes()
    parity = None
    if B.rank == 0:
        sc.set_text(d_db)
        whole = sc.find_patterns()
        parity = (whole.tobytes() == gathered) and (res["host"].numpy().tobytes() == gathered)
    sc.close()
    del d_db, h_sqz
    torch.cuda.empty_cache()
    if B.rank != 0:
        return None
    return {"database_bases": bases, "shards": B.world, "overlap": plan[0].load_hi - plan[0].end if B.world > 1 else 0,
            "left_overlap_chars": (plan[-1].begin - plan[-1].load_lo) if B.world > 1 else 0,
            "value": n_chars * steps / (ms * 1e-3) / 1e9, "unit": "Gbp/s", "ms_per_step": ms / steps,
            "e2e": {"value": n_chars * steps / (ms_e2e * 1e-3) / 1e9, "unit": "Gbp/s", "ms_per_step": ms_e2e / steps,
                    "input": "every rank's shard as a .sqz host buffer; gathered records copied to rank 0's host"},
            "hits": sum(res["counts"]), "hits_per_rank": res["counts"],
            "gather": "NCCL all_gather_into_tensor of the counts, then of the hit records (padded to the largest count)",
            "parity": "gathered hit list byte-equal to rank 0's scan of the whole database" if parity else "MISMATCH against the whole-database scan",
            "parity_ok": bool(parity)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="all", choices=["all"] + ALL_ORDER,
                    help="BASELINE.json configuration; all = the headline in full + the others under `configs`")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: ONE database partitioned over the ranks with overlap, hit records gathered over NCCL")
    ap.add_argument("--bases", type=int, default=0, help="override the headline's database shard per GPU")
    ap.add_argument("--host-mem", default="wc", choices=["pinned", "wc"],
                    help="host buffers of the end-to-end legs: torch pinned memory, or write-combined pages (pm_host_alloc)")
    ap.add_argument("--ref-bases-per-core", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling block of a multi-GPU run")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)

    B = Bench(args)
    torch, pkg = B.torch, B.pkg
    main_name = "headline" if args.config == "all" else args.config
    names = ALL_ORDER if args.config == "all" else [args.config]
    bases_max = max((args.bases if (args.bases and n == "headline") else CONFIGS[n]["bases"]) for n in names)
    d_db = B.make_database(bases_max, 1234 + B.rank)
    main_cfg = CONFIGS[main_name]
    main_bases = args.bases if (args.bases and main_name == "headline") else main_cfg["bases"]
    # planted primer copies of every configuration (before the host copies are made)
    planted = {}
    for nm in names:
        nb = args.bases if (args.bases and nm == "headline") else CONFIGS[nm]["bases"]
        planted[nm] = B.plant(d_db, B.patterns(nm)[1], CONFIGS[nm]["k"], 99 + B.rank + 17 * len(planted), limit=min(nb, bases_max))
    h_sqz = B.host_copy(pkg.sqz_encode_torch(d_db, B.sqz_table))
    torch.cuda.synchronize()

    clocks = {}
    line, res = measure_config(B, main_name, d_db, h_sqz, args.steps, args.warmup, full=True, clocks_out=clocks)
    ok = True
    if B.rank == 0:
        line = {**{"metric": line.pop("metric"), "value": line.pop("value"), "unit": line.pop("unit"),
                   "n_gpus": B.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": line.pop("ms_per_step"),
                   "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
                   "data": "synthetic"}, **line, "clocks": clocks, "cpu_affinity": B.affinity}
        if not args.no_cpu_baseline and B.world == 1:
            ok = cpu_baseline_leg(B, res, d_db, planted[main_name], line)
    # the other BASELINE configurations: short runs on the same shard
    if len(names) > 1:
        others = {}
        for nm in names[1:]:
            steps = max(1, min(args.steps, 2))
            o, r = measure_config(B, nm, d_db, h_sqz, steps, min(args.warmup, 2), full=False)
            if B.rank == 0:
                if not args.no_cpu_baseline and B.world == 1:
                    ok = cpu_baseline_leg(B, r, d_db, planted[nm], o) and ok
                others[nm] = o
        if B.rank == 0:
            line["configs"] = others
    del d_db, h_sqz
    torch.cuda.empty_cache()
    if args.config == "all":
        # the hit-dense database: 10^6 primer sites per Gbp (a genome whose repeats contain primer sites)
        cfg = CONFIGS["hit_dense"]
        d_dense = B.make_database(cfg["bases"], 4321 + B.rank)
        n_planted = B.plant_dense(d_dense, B.patterns("hit_dense")[1], cfg["k"], 77 + B.rank, cfg["dense"])
        o, r = measure_config(B, "hit_dense", d_dense, None, max(1, min(args.steps, 2)), min(args.warmup, 2), full=False)
        if B.rank == 0:
            o["planted_copies"] = n_planted
            if not args.no_cpu_baseline and B.world == 1:
                ok = cpu_baseline_leg(B, r, d_dense, [], o) and ok
            line["configs"]["hit_dense"] = o
        del d_dense
        torch.cuda.empty_cache()
    # strong scaling: one database, partitioned with overlap, hit records gathered
    if B.world > 1 and not args.no_strong:
        s = strong_scaling(B, main_name, args.steps, args.warmup)
        if B.rank == 0:
            line["strong"] = s
            ok = ok and s["parity_ok"]
            if args.scaling == "strong":
                line["weak_value"] = line["value"]
                line["value"], line["ms_per_step"], line["scaling"] = s["value"], s["ms_per_step"], "strong"
                line["e2e_weak"] = line["e2e"]
                line["e2e"] = {**s["e2e"], "h2d_bytes_per_step": int(3 * (main_bases + 1) / 8),
                               "d2h_bytes_per_step": s["hits"] * pkg.HIT_DTYPE.itemsize}
    if B.rank == 0:
        print(json.dumps(line))
        if not ok:
            print("bench.py: PARITY FAILED (see parity_ok / strong.parity in the line)", file=sys.stderr)
    if B.world > 1:
        B.dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- PrimerMatch database scan throughput on B200.

Metric (BASELINE.json): Gbp/s scanned, 1k 20-mer primers, k=2, both strands
(2000 patterns), bit-exact hit lists.  One "step" = one pass of the hot path
over one synthetic random-DNA database shard per GPU:

  value  stream bytes already resident in HBM -> pack kernel -> dense scan ->
         chains / alignment -> sorted hit list (left on the device)
  e2e    the same through the C ABI with HOST buffers: pinned host stream ->
         H2D -> pack -> scan -> ... -> hit list copied back to the host

Weak scaling: every rank scans its own shard (a multi-genome database
partitioned at sequence boundaries); the only communication is the NCCL
gather of hit counts and hit records at the end of each step.

`--impl reference` times the UNMODIFIED reference (oracle/_ref/primer_match,
compiled from the reference's own sources by oracle/Makefile) on the box's
host cores, one process per core, each on its own slice of a bounded sample of
the same workload.
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

N_PRIMERS = 1000
PRIMER_LEN = 20
K = 2
INDELS = 0
IUPAC = 0
SELECT = 0


def make_primers(seed=12345, n=None, m=None):
    import numpy as np
    n = N_PRIMERS if n is None else n
    m = PRIMER_LEN if m is None else m
    g = np.random.default_rng(seed)
    a = np.frombuffer(b"ACGT", np.uint8)[g.integers(0, 4, (n, m))]
    if IUPAC:  # one to three ambiguity codes per primer, like the reference's db/patwc.txt
        codes = np.frombuffer(b"RYKMSWBDHVN", np.uint8)
        for r in a:
            for _ in range(int(g.integers(1, 4))):
                r[int(g.integers(0, m))] = codes[int(g.integers(0, len(codes)))]
    return [bytes(r).decode() for r in a]


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


# ---------------------------------------------------------------------------
def visible_gpu_index(local_rank: int) -> int:
    cvd = os.environ.get("CUDA_VISIBLE_DEVICES")
    if cvd:
        try:
            return int(cvd.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def run_reference_sample(primers, total_bases: int, procs: int, seed: int, workdir: str):
    """One timed pass of the unmodified reference over `total_bases` of random DNA,
    split over `procs` processes.  Returns seconds (wall, all processes)."""
    import numpy as np
    ref = os.path.join(ROOT, "oracle", "_ref", "primer_match")
    per = total_bases // procs
    g = np.random.default_rng(seed)
    files = []
    pfile = os.path.join(workdir, "primers.txt")
    with open(pfile, "w") as f:
        f.write("\n".join(primers) + "\n")
    for i in range(procs):
        seq = np.frombuffer(b"ACGT", np.uint8)[g.integers(0, 4, per)]
        path = os.path.join(workdir, "s%d_%d.fasta" % (seed, i))
        with open(path, "wb") as f:
            f.write(b">sample%d\n" % i)
            body = seq.reshape(-1, 80) if per % 80 == 0 else None
            if body is not None:
                f.write(b"\n".join(bytes(r) for r in body))
            else:
                f.write(bytes(seq))
            f.write(b"\n")
        files.append(path)
    t0 = time.perf_counter()
    ps = [subprocess.Popen([ref, "-i", fp, "-P", pfile, ("-k" if INDELS else "-K"), str(K), "-r", *(["-w"] if IUPAC else []), *(["-N", str(SELECT)] if SELECT else []), "-A", "%f\\t%s\\t%e\\t%r\\t%i\\t%d\\n"],
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for fp in files]
    for p in ps:
        p.wait()
    dt = time.perf_counter() - t0
    for fp in files:
        os.unlink(fp)
    bad = [p.returncode for p in ps if p.returncode != 0]
    if bad:
        raise RuntimeError("reference exited with %s" % bad)
    return dt


def reference_available():
    p = os.path.join(ROOT, "oracle", "_ref", "primer_match")
    return os.path.exists(p) and os.access(p, os.X_OK)


def run_oracle_sample(primers, total_bases: int, seed: int):
    """Fallback CPU baseline: the oracle port (single thread)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pm_oracle as O
    g = np.random.default_rng(seed)
    seq = np.frombuffer(b"ACGT", np.uint8)[g.integers(0, 4, total_bases)]
    T = np.concatenate([[10], seq, [10]]).astype(np.uint8)
    st = O.Stream(T, np.array([1], np.int64), ["s"])
    t0 = time.perf_counter()
    O.primer_match(st, primers, k=K, indels=bool(INDELS), rev_comp=True)
    return time.perf_counter() - t0


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def config_dict(args, shard_bases):
    return {"workload": "%d random %d-mer primers%s, k=%d %s (%s %d), both strands (-r), "
                        "synthetic random-DNA database" % (N_PRIMERS, PRIMER_LEN,
                                                           " with 1-3 IUPAC codes each (-w)" if IUPAC else "", K,
                                                           "edits" if INDELS else "mismatches",
                                                           "-k" if INDELS else "-K", K),
            "shard_bases_per_gpu": shard_bases, "patterns": 2 * N_PRIMERS,
            "index": ("-N %d" % SELECT) if SELECT else ("automatic (pick_pattern_index): " + ("exact index" if K == 0 else "exact_halves" if (K == 1 and PRIMER_LEN >= 8) else "filter_bitvec")),
            "l2": "inputs larger than L2 (%.1f GB stream + %.2f GB packed text per pass)" % (shard_bases / 1e9, shard_bases / 2e9),
            "partition": "one shard per GPU, cut at sequence boundaries",
            "host_mem": args.host_mem}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    primers = make_primers()
    cores = cpu_cores()
    metric = "Gbp/s scanned, 1k 20-mer primers k=2 both strands"
    if reference_available():
        kind = "reference"
        procs = cores
        per_proc = args.ref_bases_per_core
        work = tempfile.mkdtemp(prefix="pmref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        for w in range(args.warmup):
            run_reference_sample(primers, max(200_000, per_proc // 8) * procs, procs, 100 + w, work)
        t = 0.0
        for s in range(args.steps):
            t += run_reference_sample(primers, per_proc * procs, procs, 200 + s, work)
        total = per_proc * procs * args.steps
        sample = "%d processes x %d bp per step, %d steps" % (procs, per_proc, args.steps)
    else:
        kind = "port"
        procs = 1
        per = 200_000
        for w in range(min(args.warmup, 1)):
            run_oracle_sample(primers, 20_000, 100 + w)
        t = 0.0
        for s in range(args.steps):
            t += run_oracle_sample(primers, per, 200 + s)
        total = per * args.steps
        sample = "oracle port, 1 thread x %d bp per step, %d steps" % (per, args.steps)
    value = total / t / 1e9
    line = {"impl": "reference", "metric": metric, "value": value, "unit": "Gbp/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": config_dict(args, args.bases),
            "cpu_baseline": {"value": value, "unit": "Gbp/s", "cores": procs, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bases", type=int, default=3_000_000_000, help="database shard per GPU")
    ap.add_argument("--indels", type=int, default=0, help="1: -k (edits) instead of -K")
    ap.add_argument("--iupac", type=int, default=0, help="1: IUPAC wildcard primers (db/patwc.txt style), scanned with -w")
    ap.add_argument("--k", type=int, default=K, help="errors allowed (headline: 2)")
    ap.add_argument("--plen", type=int, default=PRIMER_LEN, help="primer length (headline: 20)")
    ap.add_argument("--nprimers", type=int, default=N_PRIMERS, help="primers before reverse complements (headline: 1000)")
    ap.add_argument("--N", type=int, default=0, dest="select",
                    help="primer_match -N: pattern index (0 automatic, 5 filter_bitvec, 13 exact_halves)")
    ap.add_argument("--host-mem", default="wc", choices=["pinned", "wc"],
                    help="host buffers of the end-to-end legs: torch pinned memory, or write-combined pages (pm_host_alloc)")
    ap.add_argument("--cpu-bases-per-core", type=int, default=3_000_000)
    ap.add_argument("--ref-bases-per-core", type=int, default=4_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    globals().update(K=args.k, PRIMER_LEN=args.plen, N_PRIMERS=args.nprimers, INDELS=args.indels, IUPAC=args.iupac,
                     SELECT=args.select)
    if args.impl == "reference":
        return main_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module(PKG)
    pkg.lib()

    primers = make_primers()
    pats = pkg.pattern_array(primers, True)
    n = args.bases
    # synthetic shard: eos, random ACGT in sequences of 50 Mbp, eos  (generated on the device,
    # then copied to pinned host memory for the end-to-end leg)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    d_stream = torch.empty(n + 2, dtype=torch.uint8, device=dev)
    chunk = 1 << 28
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        d_stream[1 + a:1 + b] = lut[torch.randint(0, 4, (b - a,), device=dev, generator=gen, dtype=torch.uint8).long()]
    d_stream[0] = 10
    d_stream[n + 1] = 10
    seq_len = 50_000_000
    d_stream[seq_len::seq_len] = 10     # sequence separators
    # plant a few exact / mutated primer copies so that the hit path is exercised
    rs = np.random.default_rng(99 + rank)
    for j in range(64):
        p = bytearray(pats[int(rs.integers(0, len(pats)))].encode())
        p = bytearray(c if c in b"ACGT" else b"ACGT"[int(rs.integers(0, 4))] for c in p)
        for _ in range(int(rs.integers(0, K + 1))):
            p[int(rs.integers(0, len(p)))] = b"ACGT"[int(rs.integers(0, 4))]
        at = int(rs.integers(1000, n - 1000))
        d_stream[at:at + len(p)] = torch.tensor(list(p), dtype=torch.uint8, device=dev)
    def host_copy(d_t):
        """The tensor's bytes in page-locked host memory (what a pinned file reader would hold)."""
        if args.host_mem == "wc":
            hb = pkg.HostBuffer(d_t.numel(), write_combined=True)
            torch.from_numpy(hb.array).copy_(d_t)
            return hb
        h = torch.empty(d_t.numel(), dtype=torch.uint8, pin_memory=True)
        h.copy_(d_t)
        return h

    h_stream = host_copy(d_stream)
    # the same database as compress_seq -z would store it (.tbz table + .sqz bit stream,
    # 3 bits per character for A C G T eos): what primer_match opens when the files exist
    sqz_table = b"ACGT\n"
    d_sqz = pkg.sqz_encode_torch(d_stream, sqz_table)
    h_sqz = host_copy(d_sqz)
    del d_sqz
    # ... and as the FASTA file it came from: one record per sequence, 80 bases per line
    parts = []
    bounds = [0] + list(range(seq_len, n + 1, seq_len)) + [n + 1]
    nl = torch.tensor([10], dtype=torch.uint8, device=dev)
    for si in range(len(bounds) - 1):
        a, b = bounds[si] + 1, bounds[si + 1]          # bases of this sequence: d_stream[a:b]
        if b <= a:
            continue
        body = d_stream[a:b]
        full = (b - a) // 80 * 80
        lines = torch.cat([body[:full].view(-1, 80), nl.expand(full // 80, 1)], dim=1).reshape(-1)
        parts += [torch.tensor(list(b">synthetic%d random DNA\n" % si), dtype=torch.uint8, device=dev), lines]
        if b - a > full:
            parts += [body[full:], nl]
    d_fasta = torch.cat(parts)
    h_fasta = host_copy(d_fasta)
    del d_fasta, parts
    torch.cuda.synchronize()

    sc = pkg.Scanner(k=K, indels=bool(args.indels), wildcards=bool(args.iupac), device=local_rank, index=args.select)
    for i, p in enumerate(pats):
        sc.add_pattern(p, i + 1)
    sc.init()
    stream = torch.cuda.current_stream()
    sc.set_stream(stream.cuda_stream)

    launches = {"n": 0}
    scan_ms = []

    def step_resident():
        sc.set_text(d_stream)                      # pack kernel (1 launch)
        ptr, nh = sc.find_patterns_device()
        st = sc.stats()
        launches["n"] += 1 + st["kernel_launches"]
        scan_ms.append(st["scan_ms"])
        return nh, st

    def step_e2e(source):
        if source == "sqz":
            sc.set_text_sqz(h_sqz, sqz_table, n + 2)   # H2D of the .sqz bytes + unpack + pack
        elif source == "fasta":
            sc.set_text_fasta(h_fasta)             # H2D of the file + normalisation on the device + pack
        else:
            sc.set_text(h_stream)                  # H2D of the stream from pinned memory + pack
        hits = sc.find_patterns()                  # ... + D2H of the hit list
        st = sc.stats()
        launches["n"] += 1 + st["kernel_launches"]
        return hits, st

    def gather_counts(nh):
        if world == 1:
            return nh
        t = torch.tensor([nh], dtype=torch.int64, device=dev)
        dist.all_reduce(t)                      # NCCL: the hit counts of all shards, summed
        return int(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        nh, st = step_resident()
        gather_counts(nh)
    launches["n"] = 0
    scan_ms.clear()
    sampler = ClockSampler(visible_gpu_index(local_rank))
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    total_hits = 0
    for _ in range(args.steps):
        nh, st = step_resident()
        total_hits = gather_counts(nh)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    n_launch = launches["n"]
    mean_scan_ms = sum(scan_ms) / max(1, len(scan_ms))

    # end-to-end legs: host buffers in, hit list out.  "e2e" reads the database in the
    # reference's compressed form, "e2e_stream" as the plain character stream.
    e2e_ms, d2h, hit_sets = {}, 0, {}
    for source in ("sqz", "stream", "fasta"):
        for _ in range(min(args.warmup, 2)):
            step_e2e(source)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for _ in range(args.steps):
            hits, st = step_e2e(source)
            d2h = hits.nbytes + 64
            gather_counts(len(hits))
        e3.record(stream)
        barrier()
        e2e_ms[source] = e2.elapsed_time(e3)
        hit_sets[source] = hits.tobytes()
    assert hit_sets["sqz"] == hit_sets["stream"] == hit_sets["fasta"], "the three input forms disagree"
    ms_e2e, ms_e2e_stream, ms_e2e_fasta = e2e_ms["sqz"], e2e_ms["stream"], e2e_ms["fasta"]

    if world > 1:
        t = torch.tensor([ms, ms_e2e, mean_scan_ms, ms_e2e_stream, ms_e2e_fasta], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, mean_scan_ms, ms_e2e_stream, ms_e2e_fasta = (float(x) for x in t.tolist())

    if rank == 0:
        bases_total = float(n + 2) * world * args.steps
        value = bases_total / (ms * 1e-3) / 1e9
        e2e = bases_total / (ms_e2e * 1e-3) / 1e9
        e2e_stream = bases_total / (ms_e2e_stream * 1e-3) / 1e9
        e2e_fasta = bases_total / (ms_e2e_fasta * 1e-3) / 1e9
        # roofline of the dominant kernel (the dense scan), per launch
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        st_last = sc.stats()
        on_index = st_last["scan_path"] == pkg.PM_SCAN_INDEX
        cells = float(n + 2) * len(pats) * (K + 1)
        if on_index:
            # pair_scan_kernel reads the packed text once: one uint2 (codes, flags) per 16 characters
            alg_bytes = (n + 2) * 0.5
            # dram__bytes_read+write of one launch over 3.0e9 bases: 1504.0 + 6.3 MB
            # (ncu --set full, profiles/README.md, prof_pairsub_r1), scaled to this launch's bases
            traffic = None if args.indels else 0.5034 * (n + 2)
            if args.indels and K:
                kname = "pair_scan_kernel<true> + index_verify_kernel<%d>" % K
            else:
                # third template argument: the level-1b variant (24 000 < entries <= 100 000, csrc/pm_b200.cu)
                kname = "pair_scan_sub_kernel<%d, %d, %s>" % (
                    st_last["index_q"], st_last["index_nd"],
                    "true" if 24000 < st_last["index_entries"] <= 100000 else "false")
            tsrc = "ncu --set full capture prof_pairsub_r1 (profiles/README.md), bytes per base x bases"
        else:
            n_planes = 5 if not args.iupac else 16
            alg_bytes = (n + 2) / 8.0 * n_planes * (1088.0 / 1024.0)      # planes read once per launch
            # (268 Mbp: 567.6 MB with 4 pattern chunks), scaled to this launch's bases
            traffic = None if args.indels else 2.115 * (n + 2)
            kname = ("hamming_scan_kernel<%d>" % K) if not args.indels else ("edit_seed_kernel<%d> + edit_verify_kernel<%d>" % (K, K))
            tsrc = "ncu --set full capture prof_hamming_r1b (profiles/README.md), bytes per base x bases"
        hbm_achieved = alg_bytes / (mean_scan_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": hbm_achieved / hbm_peak, "traffic": traffic,
                "traffic_source": tsrc,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s",
                "kernel": kname,
                "kernel_ms": mean_scan_ms,
                "algorithmic_bytes_per_base": alg_bytes / (n + 2),
                "note": "the scan is integer-pipe bound, not HBM bound (north star: the slower of the two): see int_pipe",
                "cell_updates_per_s": cells / (mean_scan_ms * 1e-3),
                "scan_path": "index" if on_index else "planes"}
        if on_index:
            roof["index"] = {"piece_len": st_last["index_q"], "pieces": st_last["index_pieces"],
                             "distances": st_last["index_nd"], "entries": st_last["index_entries"]}
        if on_index and not args.indels and K > 0 and st_last["index_pieces"] == K + 2:
            # the pipe that actually binds pair_scan_sub_kernel: shared-memory wavefronts.  One level-1
            # lookup is one lane of an LDS; 32 lanes hitting 32 random banks take 3.76 wavefronts
            # (l1tex__data_pipe_lsu_wavefronts_mem_shared / smsp__sass_inst_executed_op_shared_ld of the
            # ncu capture prof_pairsub_r1); an SM retires one wavefront per clock.
            wf = st_last["index_lookups"] * 3.76 / 32.0
            sm_clock = (clocks.get("sm_mhz") or 1965) * 1e6
            lsu_peak = 148 * sm_clock
            roof["smem_wavefronts"] = {"unit": "Gwavefront/s", "achieved": wf / (mean_scan_ms * 1e-3) / 1e9,
                                       "peak": lsu_peak / 1e9, "frac": wf / (mean_scan_ms * 1e-3) / lsu_peak,
                                       "wavefronts_per_warp_lds": 3.76,
                                       "note": "ncu (profiles/README.md): l1tex data pipe 93.9 % busy incl. queue traffic"}
        try:
            ip = pkg.int_pipe_report(sc, mean_scan_ms, local_rank)
            roof["int_pipe"] = ip
        except Exception as ex:  # diagnostic only
            roof["int_pipe"] = {"error": str(ex)[:200]}
        line = {"metric": "Gbp/s scanned, 1k 20-mer primers k=2 both strands", "value": value, "unit": "Gbp/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
                "data": "synthetic", "config": config_dict(args, n),
                "hits_per_step": total_hits, "clocks": clocks,
                "e2e": {"value": e2e, "unit": "Gbp/s", "h2d_bytes_per_step": int(h_sqz.numel()) * world,
                        "d2h_bytes_per_step": d2h * world, "ms_per_step": ms_e2e / args.steps,
                        "input": "pinned host buffer holding the database as compress_seq -z stores it "
                                 "(.sqz, 3 bits per character; pm_set_text_sqz_host)"},
                "e2e_stream": {"value": e2e_stream, "unit": "Gbp/s", "h2d_bytes_per_step": (n + 2) * world,
                               "d2h_bytes_per_step": d2h * world, "ms_per_step": ms_e2e_stream / args.steps,
                               "input": "pinned host buffer holding the normalised character stream, "
                                        "1 byte per character (pm_set_text_host)"},
                "e2e_fasta": {"value": e2e_fasta, "unit": "Gbp/s", "h2d_bytes_per_step": int(h_fasta.numel()) * world,
                              "d2h_bytes_per_step": d2h * world, "ms_per_step": ms_e2e_fasta / args.steps,
                              "input": "pinned host buffer holding the FASTA file (80 bases per line), normalised "
                                       "on the device chunk by chunk while the copy runs (pm_set_text_fasta_host); pack and scan follow"},
                "gpu_launches": n_launch, "roofline": roof}
        if not args.no_cpu_baseline and world == 1:
            cores = cpu_cores()
            try:
                if reference_available():
                    work = tempfile.mkdtemp(prefix="pmref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
                    per = args.cpu_bases_per_core
                    dt = run_reference_sample(primers, per * cores, cores, 300, work)
                    line["cpu_baseline"] = {"value": per * cores / dt / 1e9, "unit": "Gbp/s", "cores": cores,
                                            "kind": "reference",
                                            "sample": "%d processes x %d bp of the same workload, %.1f s" % (cores, per, dt)}
                else:
                    dt = run_oracle_sample(primers, 200_000, 300)
                    line["cpu_baseline"] = {"value": 200_000 / dt / 1e9, "unit": "Gbp/s", "cores": 1, "kind": "port",
                                            "sample": "oracle port, 200000 bp, %.1f s" % dt}
            except Exception as ex:
                line["cpu_baseline"] = {"error": str(ex)[:200]}
        print(json.dumps(line))
    sc.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

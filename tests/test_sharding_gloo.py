"""The N>1 host logic on CPU: two gloo ranks partition one stream with
shard_plan(), each produces the hits it owns (the oracle stands in for the GPU
scan here -- this test checks partition, ownership and gather, not the
kernels), and the gathered list must equal the whole-stream answer."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import PKG_NAME, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _owned_hits(O, pkg, T, lo, pats, k, indels, begin, end):
    """Hits of the shard T (global offset lo) owned by [begin, end): chains whose
    first candidate position lies in the range (filter_bitvec.cc:157-168)."""
    cands = O.wm_scan(T, pats, k, indels)
    if len(cands) == 0:
        return np.zeros(0, pkg.HIT_DTYPE)
    order = np.lexsort((cands[:, 0], cands[:, 1]))
    cands = cands[order]
    out = []
    i = 0
    n = len(cands)
    while i < n:
        j = i
        while j + 1 < n and cands[j + 1, 1] == cands[i, 1] and cands[j + 1, 0] - cands[j, 0] <= 2 * k + 1:
            j += 1
        first = int(cands[i, 0]) + lo
        if begin <= first < end:
            chain = cands[i:j + 1]
            for e, pid, _ in O.filter_chains(T, pats, chain, k, indels):
                al = O.editdist_align(T, pats[int(pid) - 1], int(e), int(e), k, indels)
                rec = np.zeros(1, pkg.HIT_DTYPE)
                rec["start"], rec["end"], rec["pattern_id"] = al.start + lo, al.end + lo, int(pid)
                rec["first"] = first
                rec["editdist"], rec["value"], rec["ncodes"] = al.editdist, al.val, len(al.codes)
                rec["codes"][0, :len(al.codes)] = list(al.codes)
                out.append(rec)
        i = j + 1
    return np.concatenate(out) if out else np.zeros(0, pkg.HIT_DTYPE)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import torch.distributed as dist
    import casegen
    import pm_oracle as O
    pkg = importlib.import_module(PKG_NAME)
    sharding = importlib.import_module(PKG_NAME + ".sharding")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    k, indels = 2, True
    fasta, primers = casegen.make_case(321, nseq=4, seqlen=1500, nprimers=8, plen=(16, 22), k=k, indels=indels)
    st = O.normalize_fasta(fasta.encode())
    pats = O.pattern_array(primers, True)
    T = st.chars
    overlap = max(len(p) for p in pats) + 3 * k + 2
    plan = sharding.shard_plan(len(T), world, overlap, right_guard=200)
    sh = plan[rank]
    local = _owned_hits(O, pkg, T[sh.load_lo:sh.load_hi], sh.load_lo, pats, k, indels, sh.begin, sh.end)
    merged = sharding.gather_hits(local, dist)
    if rank == 0:
        whole = _owned_hits(O, pkg, T, 0, pats, k, indels, 0, len(T) + 1)
        whole = whole[np.lexsort((whole["pattern_id"], whole["first"]))]
        q.put((len(whole), merged.tobytes() == whole.tobytes(), [len(local)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_and_gather_gloo(oracle):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    n, same, _ = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert n > 5
    assert same


def test_shard_plan_covers_everything_once(pkg):
    sharding = importlib.import_module(PKG_NAME + ".sharding")
    for n, world in ((1000, 1), (1000, 2), (12345, 8), (7, 4)):
        plan = sharding.shard_plan(n, world, overlap=30, right_guard=50)
        assert plan[0].begin == 0 and plan[-1].end == n + 1
        for a, b in zip(plan, plan[1:]):
            assert a.end == b.begin
        for s in plan:
            assert s.load_lo <= max(0, s.begin - 30) or s.rank == 0
            assert s.load_hi >= min(n, s.end)
    cuts = np.array([100, 480, 900])
    plan = sharding.shard_plan(1000, 2, 30, 50, cut_points=cuts)
    assert plan[0].end == 900 or plan[0].end == 480 or plan[0].end in cuts

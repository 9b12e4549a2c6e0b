"""One database cut with shard_plan() over several GPUs, one context per GPU (one process per
GPU), hit records gathered over NCCL: the gathered list must equal the single-GPU scan of the
whole database byte for byte.  Needs >= 2 GPUs (gpurun --gpus 2); the same check runs inside
bench.py's strong-scaling block at every N > 1 of the driver's scaling run."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

from conftest import PKG_NAME, ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        _work(rank, world, port, q)
    except Exception as ex:     # the parent must not wait for a worker that died
        import traceback
        q.put(("error", rank, traceback.format_exc()[-1500:]))


def _work(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import random
    import torch
    import torch.distributed as dist
    import casegen
    pkg = importlib.import_module(PKG_NAME)
    sharding = importlib.import_module(PKG_NAME + ".sharding")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=dev)
    results = []
    for k, indels in ((2, False), (2, True)):
        rng = random.Random(900 + k + indels)
        primers = [casegen.rand_dna(rng, 20) for _ in range(200)] + ["A" * 18]
        pats = pkg.pattern_array(primers, True)
        n = 50_000_000
        g = torch.Generator(device=dev)
        g.manual_seed(4321)                      # the same database on every rank
        lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
        d = lut[torch.randint(0, 4, (n + 1,), device=dev, generator=g, dtype=torch.uint8).long()]
        d[0::10_000_000] = 10
        d[n] = 10
        plan0 = sharding.shard_plan(n + 1, world, 40)
        cuts = [s.end for s in plan0[:-1]]
        for j in range(400):                     # planted copies, many of them straddling the cuts
            q_ = casegen.mutate(rng, rng.choice(primers[:200]), rng.randint(0, k), indels)
            at = (cuts[j % len(cuts)] - 30 + (j // len(cuts)) % 60) if j < 120 else rng.randrange(1000, n - 1000)
            d[at:at + len(q_)] = torch.tensor(list(q_.encode()), dtype=torch.uint8, device=dev)
        d[cuts[0] - 40:cuts[0] + 40] = ord("A")  # a poly-A run across the first cut: one chain, one owner
        sc = pkg.Scanner(k=k, indels=indels, index=pkg.PM_INDEX_FILTER_BITVEC, device=rank)
        for i, p in enumerate(pats):
            sc.add_pattern(p, i + 1)
        sc.init()
        plan = sharding.shard_plan(n + 1, world, sc.shard_overlap())
        sh = plan[rank]
        lo = sh.load_lo & ~3                     # pm_set_text_device wants a 4-byte aligned pointer
        sc.set_text(d[lo:sh.load_hi], offset=lo)
        ptr, nh = sc.find_patterns_device(sh.begin, sh.end)
        counts, recs = sharding.gather_hits_device(ptr, nh, pkg.HIT_DTYPE.itemsize, dist, dev)
        gathered = recs.cpu().numpy().tobytes()
        # the host route (numpy records through gather_hits) must agree
        local = sc.find_patterns(sh.begin, sh.end)
        merged = sharding.gather_hits(local, dist)
        if rank == 0:
            sc.set_text(d)
            whole = sc.find_patterns()
            results.append((k, indels, len(whole), counts, whole.tobytes() == gathered, whole.tobytes() == merged.tobytes()))
        sc.close()
        dist.barrier()
    if rank == 0:
        q.put(results)
    dist.destroy_process_group()


def test_shards_over_gpus_gathered_over_nccl_equal_the_whole_scan(built_lib):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs (run under gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = q.get(timeout=400)
    if results and results[0] == "error":
        for p in procs:
            p.kill()
        pytest.fail("rank %d: %s" % (results[1], results[2]))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert len(results) == 2
    for k, indels, n, counts, same_dev, same_host in results:
        assert n >= 300 and all(c > 0 for c in counts), (k, indels, n, counts)
        assert same_dev and same_host, (k, indels)

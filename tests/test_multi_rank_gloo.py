"""N>1 host path on CPU: world_size-2 gloo run of the key broadcast + batch sharding used by bench.py / a multi-GPU
deployment.  Each rank evaluates its shard (with the oracle standing in for the GPU, since this box has none) and the
gathered result must equal the single-process result: the path has no data-path collective."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    from tfhe_gpu_b200.dist import broadcast_keys, shard_range

    pd = bk = ksk = None
    p = po.Port.params_named(po.TOY, po.GINX)
    port_ = po.Port(p)
    if rank == 0:
        sk, bk, ksk = port_.keygen(3)
        pd = p.as_dict()
    pd, bk_t, ksk_t = broadcast_keys(pd, bk, ksk, torch.device("cpu"), src=0)
    assert pd == p.as_dict()
    bk_r = bk_t.numpy().view(np.uint64)
    ksk_r = ksk_t.numpy().view(np.uint64)
    batch = 11                                      # ragged on purpose
    rng = np.random.default_rng(42)                 # same inputs on every rank
    c1 = rng.integers(0, p.q, (batch, p.n + 1), dtype=np.uint64)
    c2 = rng.integers(0, p.q, (batch, p.n + 1), dtype=np.uint64)
    start, count = shard_range(batch, world, rank)
    mine = port_.eval_bin_gate(bk_r, ksk_r, po.GATES["NAND"], c1[start:start + count], c2[start:start + count], p.q)
    gathered = [None] * world
    dist.all_gather_object(gathered, (start, mine))
    if rank == 0:
        full = np.concatenate([g[1] for g in sorted(gathered, key=lambda t: t[0])])
        want = port_.eval_bin_gate(bk, ksk, po.GATES["NAND"], c1, c2, p.q)
        q.put(bool(np.array_equal(full, want)))
    dist.destroy_process_group()


def test_two_rank_key_broadcast_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert ok and all(p.exitcode == 0 for p in procs)

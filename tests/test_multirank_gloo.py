"""The N>1 path on CPU: world_size-2 gloo processes shard a batch, receive the key by broadcast, encrypt
their shard (kernels run in the TEST-ONLY host simulation), and all-gather the result limbs; the gathered
ciphertexts must equal the reference-generated golden ciphertexts."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, queue):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import importlib
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    from oracle.golden import H, load_golden
    pb = importlib.import_module("python-paillier_b200")
    par = importlib.import_module("python-paillier_b200.parallel")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = pb.Engine(ge.build_hostsim())
        fx = load_golden("vectors_256.json")
        key = [H(fx["n"]), H(fx["p"]), H(fx["q"])] if rank == 0 else [0, 0, 0]
        n, p, q = par.broadcast_ints(key, 8)
        enc = [e for e in fx["encrypt"] if 0 <= H(e["m"]) < n][:21]
        batch = len(enc)
        lo, hi = par.shard_range(batch, rank, world)
        pub = pb.PublicContext(n, engine=eng)
        m = pb.ints_to_limbs([H(e["m"]) for e in enc[lo:hi]], pub.n_limbs)
        r = pb.ints_to_limbs([H(e["r"]) for e in enc[lo:hi]], pub.n_limbs)
        c_local = torch.from_numpy(pub.encrypt_host(m, r).view(np.int32).copy())
        full = par.all_gather_rows(c_local, batch)
        got = pb.limbs_to_ints(full.numpy().view(np.uint32))
        priv = pb.PrivateContext(p, q, engine=eng)
        dec = priv.raw_decrypt(got[lo:hi])
        ok = got == [H(e["c"]) for e in enc] and dec == [H(e["d"]) for e in enc[lo:hi]]
        queue.put((rank, bool(ok), hi - lo))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_broadcast_allgather():
    import torch.multiprocessing as mp
    import __graft_entry__ as ge
    ge.build_hostsim()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [True, True]
    assert sum(r[2] for r in res) == 21

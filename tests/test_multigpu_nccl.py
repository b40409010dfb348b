"""N>1 on real GPUs: two ranks over NCCL shard a batch, receive the key by broadcast, encrypt their shard with
the CUDA engine and all-gather the ciphertext limbs.  Needs >= 2 visible GPUs (skipped otherwise; the same logic
runs on gloo + the simulation engine in test_multirank_gloo.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, queue):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), NCCL_DEBUG="WARN")
    import importlib
    import torch
    import torch.distributed as dist
    from oracle.golden import H, load_golden
    pb = importlib.import_module("python-paillier_b200")
    par = importlib.import_module("python-paillier_b200.parallel")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        fx = load_golden("vectors_1024.json")
        key = [H(fx["n"]), H(fx["p"]), H(fx["q"])] if rank == 0 else [0, 0, 0]
        n, p, q = par.broadcast_ints(key, 32, device="cuda:%d" % rank)
        enc = [e for e in fx["encrypt"] if 0 <= H(e["m"]) < n]
        batch = len(enc)
        lo, hi = par.shard_range(batch, rank, world)
        pub = pb.PublicContext(n, device=rank)
        m = torch.from_numpy(pb.ints_to_limbs([H(e["m"]) for e in enc[lo:hi]], pub.n_limbs).view(np.int32).copy()).cuda(rank)
        r = torch.from_numpy(pb.ints_to_limbs([H(e["r"]) for e in enc[lo:hi]], pub.n_limbs).view(np.int32).copy()).cuda(rank)
        c = torch.empty((hi - lo, pub.c_limbs), dtype=torch.int32, device="cuda:%d" % rank)
        pub.encrypt_dev(m, r, c, hi - lo)
        full = par.all_gather_rows(c, batch)
        got = pb.limbs_to_ints(full.cpu().numpy().view(np.uint32))
        priv = pb.PrivateContext(p, q, device=rank)
        dec = priv.raw_decrypt(got[lo:hi])
        queue.put((rank, got == [H(e["c"]) for e in enc] and dec == [H(e["d"]) for e in enc[lo:hi]], hi - lo))
    finally:
        dist.destroy_process_group()


def test_two_gpu_shard_broadcast_allgather():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [True, True]

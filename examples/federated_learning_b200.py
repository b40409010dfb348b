#!/usr/bin/env python3
"""BASELINE.json configs[4]: the protocol of the reference's examples/federated_learning_with_encryption.py
(:195-231) at 2048 bits with a D-element gradient, on the batched engine.

Per round: every client encrypts its gradient (encrypt_vector, :122-123), the encrypted gradients are summed
client by client (sum_encrypted_vectors, :130-133 -- with the reference's exponent alignment), the server decrypts
the aggregate and divides by the number of clients (decrypt_aggregate, :143-144).  Gradients are drawn like the
reference's synthetic noise (seed 43, :76-77).  Prints one JSON line with the wall-clock of each phase and checks
the decrypted aggregate against the plaintext mean.

    python examples/federated_learning_b200.py [--dim 100000] [--clients 5] [--key-bits 2048] [--cpu-sample 200]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=100000)
    ap.add_argument("--clients", type=int, default=5)
    ap.add_argument("--key-bits", type=int, default=2048)
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="measurement only: also time this many elements of the same round on one CPU core with the oracle port "
                         "of the reference (oracle/, test infrastructure) and scale; 0 = GPU round only")
    args = ap.parse_args()

    import torch
    import paillier_b200 as phe
    import importlib
    fixtures = importlib.import_module("python-paillier_b200.fixtures")
    H, load_golden = fixtures.H, fixtures.load_golden

    fx = load_golden("vectors_%d.json" % args.key_bits)           # a fixed key keeps runs comparable
    pk = phe.PaillierPublicKey(H(fx["n"]))
    sk = phe.PaillierPrivateKey(pk, H(fx["p"]), H(fx["q"]))
    grads = [np.random.RandomState(43 + i).randn(args.dim) * 0.1 for i in range(args.clients)]
    pk.encrypt_batch(grads[0][:256])                               # warm-up: contexts, workspaces
    torch.cuda.synchronize()

    t = {}
    t0 = time.perf_counter()
    enc = [pk.encrypt_batch(g) for g in grads]                     # clients: encode + encrypt (fresh random r each)
    torch.cuda.synchronize()
    t["encrypt_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    acc = enc[0]
    for e in enc[1:]:
        acc = acc + e                                              # ring sum with exponent alignment
    torch.cuda.synchronize()
    t["sum_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    agg = np.array(sk.decrypt_batch(acc)) / args.clients           # server: decrypt + decode + average
    t["decrypt_s"] = time.perf_counter() - t0
    t["round_s"] = sum(t.values())
    ok = bool(np.allclose(agg, np.mean(grads, axis=0), rtol=0, atol=1e-12))

    # the same round on ONE CPU core with the oracle port (libgmp = what gmpy2 wraps), small sample, scaled
    cpu = None
    if args.cpu_sample:
        from oracle import paillier_oracle as orc
        import random
        orc.BACKEND = "gmp" if orc.have_gmp() else "python"
        opub = orc.PublicConsts(pk.n)
        opriv = orc.PrivateConsts(opub, sk.p, sk.q)
        S = args.cpu_sample
        rng = random.Random(1)
        t0 = time.perf_counter()
        encs = [[phe.EncodedNumber.encode(pk, float(x)) for x in g[:S]] for g in grads]
        cts = [[orc.raw_encrypt(opub, e.encoding, rng.randrange(1, pk.n)) for e in row] for row in encs]
        te = time.perf_counter() - t0
        t0 = time.perf_counter()
        accc, acce = cts[0], [e.exponent for e in encs[0]]
        for row, erow in zip(cts[1:], encs[1:]):
            nxt, nxe = [], []
            for a, ea, b, eb in zip(accc, acce, row, erow):
                ex = min(ea, eb.exponent)
                if ea > ex:
                    a = orc.raw_mul(opub, a, 16 ** (ea - ex))
                if eb.exponent > ex:
                    b = orc.raw_mul(opub, b, 16 ** (eb.exponent - ex))
                nxt.append(orc.raw_add(opub, a, b)); nxe.append(ex)
            accc, acce = nxt, nxe
        ts = time.perf_counter() - t0
        t0 = time.perf_counter()
        dec = [orc.raw_decrypt(opriv, c) for c in accc]
        td = time.perf_counter() - t0
        scale = args.dim / S
        cpu = {"cores": 1, "engine": "oracle port of phe on " + orc.BACKEND, "sample_elements": S,
               "round_s_scaled_to_dim": (te + ts + td) * scale,
               "encrypt_s": te * scale, "sum_s": ts * scale, "decrypt_s": td * scale}
        assert [phe.EncodedNumber(pk, d, e).decode() for d, e in zip(dec, acce)] == sk.decrypt_batch(
            phe.EncryptedVector(pk, acc.limbs[:S].contiguous(), acc.exponents[:S]))

    print(json.dumps({"workload": "configs[4] federated round", "key_bits": args.key_bits, "dim": args.dim, "clients": args.clients,
                      "gpu": t, "aggregate_matches_plaintext_mean": ok, "cpu_reference_1core": cpu,
                      "ops": {"encrypts": args.clients * args.dim, "adds": (args.clients - 1) * args.dim, "decrypts": args.dim}}))


if __name__ == "__main__":
    main()

"""Tensor-core encrypt path (pai_tc.cuh) against the integer-pipe digit path and the oracle, with timings.
   python bench_micro/tc_check.py [key_bits ...]      (env PAI_TC_STAGGER=<cycles> is read at context creation)"""
import json, os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paillier_b200 as pb
import importlib
_fx = importlib.import_module("python-paillier_b200.fixtures")


def ctx(n, tc, stagger=None, pq=None):
    os.environ["PAI_TC"] = "2" if tc else "0"
    if stagger is not None:
        os.environ["PAI_TC_STAGGER"] = str(stagger)
    c = pb.PublicContext(n) if pq is None else pb.PrivateContext(*pq)
    os.environ.pop("PAI_TC", None)
    return c


def run(kb, waves, staggers):
    n, p, q = _fx.fixed_key(kb)
    ref = ctx(n, False)
    res = {"key_bits": kb}
    first = True
    for stg in staggers:
        pub = ctx(n, True, stg)
        wave = pub.wave()
        dwave = ctx(n, True, stg, (p, q)).wave()
        B = max(int(waves * wave), dwave) + 77                 # encrypt: `waves` full waves + a 77-row tail
        Bd = (B - 77) // dwave * dwave + 77                     # decrypt: whole waves of ITS kernel + the same tail
        if first:
            d_m = torch.empty((B, pub.n_limbs), dtype=torch.int32, device="cuda")
            d_r = torch.empty_like(d_m)
            pub.random_lt_n_dev(d_m, B, seed=b"\x01" * 32, nonce=0)
            pub.random_lt_n_dev(d_r, B, seed=b"\x01" * 32, nonce=1)
            d_c = torch.empty((B, pub.c_limbs), dtype=torch.int32, device="cuda")
            d_ref = torch.empty_like(d_c)
            ref.encrypt_dev(d_m, d_r, d_ref, B)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ref.encrypt_dev(d_m, d_r, d_ref, B); e1.record(); torch.cuda.synchronize()
            res["imad_path"] = {"rows": B, "ms": e0.elapsed_time(e1), "per_s": B / e0.elapsed_time(e1) * 1e3, "wave": ref.wave()}
        d_c.zero_()
        pub.encrypt_dev(d_m, d_r, d_c, B)
        torch.cuda.synchronize()
        same = bool((d_c == d_ref).all().item())
        bad = int((d_c != d_ref).any(dim=1).sum().item())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pub.encrypt_dev(d_m, d_r, d_c, B); e1.record(); torch.cuda.synchronize()
        res["tc_stagger_%d" % stg] = {"rows": B, "wave": wave, "ms": e0.elapsed_time(e1), "per_s": B / e0.elapsed_time(e1) * 1e3,
                                       "equal_to_imad_path": same, "rows_differing": bad}
        if first:
            from oracle import paillier_oracle as orc
            orc.BACKEND = "gmp" if orc.have_gmp() else "python"
            opub = orc.PublicConsts(n)
            idx = [0, 1, 127, 128, 255, 256, wave - 1, wave, B - 1] + [random.Random(1).randrange(B) for _ in range(7)]
            ti = torch.tensor(idx, device="cuda")
            mi, ri, ci = (pb.limbs_to_ints(t[ti].cpu().numpy().view(np.uint32)) for t in (d_m, d_r, d_c))
            res["oracle_rows_equal"] = ci == [orc.raw_encrypt(opub, a, b) for a, b in zip(mi, ri)]
        # decrypt: tensor-core path vs integer-pipe path on the ciphertexts just produced
        priv = ctx(n, True, stg, (p, q))
        if first:
            pref = ctx(n, False, None, (p, q))
            d_dref = torch.empty((B, pub.n_limbs), dtype=torch.int32, device="cuda")
            pref.decrypt_dev(d_ref, d_dref, B); torch.cuda.synchronize()
            e0.record(); pref.decrypt_dev(d_ref, d_dref, B); e1.record(); torch.cuda.synchronize()
            res["imad_decrypt"] = {"ms": e0.elapsed_time(e1), "per_s": B / e0.elapsed_time(e1) * 1e3, "wave": pref.wave(),
                                   "roundtrip": bool((d_dref == d_m).all().item())}
        d_d = torch.zeros((B, pub.n_limbs), dtype=torch.int32, device="cuda")
        priv.decrypt_dev(d_ref, d_d, B); torch.cuda.synchronize()
        e0.record(); priv.decrypt_dev(d_ref, d_d, Bd); e1.record(); torch.cuda.synchronize()
        res["tc_decrypt_stagger_%d" % stg] = {"rows": Bd, "ms": e0.elapsed_time(e1), "per_s": Bd / e0.elapsed_time(e1) * 1e3, "wave": priv.wave(),
                                               "roundtrip": bool((d_d == d_m).all().item()), "rows_differing": int((d_d != d_m).any(dim=1).sum().item())}
        first = False
        print(json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    kbs = [int(x) for x in sys.argv[1:]] or [1024, 2048]
    stg = [int(x) for x in os.environ.get("TC_STAGGERS", "0,40000").split(",")]
    for kb in kbs:
        run(kb, 2.0, stg)

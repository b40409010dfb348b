#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 Paillier engine (BASELINE.json configs[1]).

Workload (one "step"): raw_encrypt of a batch of 2048-bit-key plaintexts followed by raw_decrypt of
the resulting ciphertexts (BASELINE.json configs[1]: 2048-bit key, batch 1M, bit-exact round trip).
`value` is encrypts/s of the whole job with inputs resident in HBM; the decrypt leg of the same
steps is reported under "decrypt".  `e2e` runs the same step through the host-pointer C ABI with
pinned host buffers (H2D + kernels + D2H inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

Multi-GPU (torchrun, one rank per GPU): the batch shards across ranks (weak scaling: `--batch` is the
per-GPU batch), no data-path collective; the key limbs are broadcast from rank 0 over NCCL.
`--impl reference` times the reference's CPU path (oracle port of phe bound to libgmp -- the routine
gmpy2.powmod wraps -- fanned over all host cores) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KEY_BITS = 2048
DEFAULT_BATCH = 1 << 20


def executed_macs_2048():
    """MACs the kernels actually execute per op at 2048-bit keys on the base-n digit path (pai_digit.cuh): 64 MACs
    per tile product, 36 per truncated quotient product; sliding-window encrypt (w = 6), fixed-window CRT decrypt."""
    def dmul(nth):
        return 64 * (5 * nth * nth + 2 * nth) + 36 * 2 * nth
    def dsqr(nth):
        return 64 * (nth * (nth - 1) // 2 + nth + nth * nth + nth + 2 * nth * nth + nth) + 36 * 2 * nth
    enc = 2048 * dsqr(8) + (293 + 32 + 2) * dmul(8) + 64 * 64
    dec = 2 * (1024 * dsqr(4) + (205 + 30 + 6) * dmul(4))
    return enc, dec


def canonical_macs(kb):
    """canonical 32x32->64 MAC counts (schoolbook CIOS, window 5, no squaring credit)."""
    def modmul(L):
        return 2 * L * L + L

    def modexp(e, L):
        return (e + -(-e // 5) + 30 + 2) * modmul(L)
    enc = modexp(kb, kb // 16) + 2 * modmul(kb // 16) + (kb // 32) ** 2
    dec = 2 * modexp(kb // 2, kb // 32)
    return enc, dec


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------- CPU side
def _cpu_worker(args):
    """Encrypt+decrypt `count` elements with the oracle port bound to libgmp.  Returns (enc_s, dec_s)."""
    n, p, q, seed, count = args
    import random
    from oracle import paillier_oracle as orc
    orc.BACKEND = "gmp" if orc.have_gmp() else "python"
    pub = orc.PublicConsts(n)
    priv = orc.PrivateConsts(pub, p, q)
    rng = random.Random(seed)
    ms = [rng.randrange(n) for _ in range(count)]
    rs = [rng.randrange(1, n) for _ in range(count)]
    t0 = time.perf_counter()
    cs = [orc.raw_encrypt(pub, m, r) for m, r in zip(ms, rs)]
    t1 = time.perf_counter()
    ds = [orc.raw_decrypt(priv, c) for c in cs]
    t2 = time.perf_counter()
    assert ds == ms
    return t1 - t0, t2 - t1, orc.BACKEND


def cpu_sample(n, p, q, per_core, cores):
    """All host cores, `per_core` elements each.  Returns dict(enc_per_s, dec_per_s, cores, backend, sample)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(n, p, q, 1000 + i, per_core) for i in range(cores)])
    wall = time.perf_counter() - t0
    total = per_core * cores
    # whole-host throughput = sum of the per-process rates (the processes overlap; summing rates is generous to
    # the CPU side, it ignores process start-up skew)
    enc_rate = sum(per_core / r[0] for r in res)
    dec_rate = sum(per_core / r[1] for r in res)
    return {"enc_per_s": enc_rate, "dec_per_s": dec_rate, "cores": cores, "backend": res[0][2],
            "sample": "%d encrypt + %d decrypt (2048-bit) over %d processes" % (total, total, cores), "wall_s": wall}


def run_reference(args, key):
    """--impl reference: the reference's CPU path on this box's host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, p, q = key
    cores = os.cpu_count() or 1
    per_core = max(4, args.ref_per_core)
    for _ in range(args.warmup):
        cpu_sample(n, p, q, 2, cores)
    t = []
    last = None
    for _ in range(args.steps):
        last = cpu_sample(n, p, q, per_core, cores)
        t.append(last)
    enc = sum(x["enc_per_s"] for x in t) / len(t)
    dec = sum(x["dec_per_s"] for x in t) / len(t)
    total = per_core * cores
    line = {
        "impl": "reference", "metric": "paillier_raw_encrypts_per_sec_2048", "value": enc, "unit": "encrypts/s",
        "decrypt": {"value": dec, "unit": "decrypts/s"},
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / enc, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (exact integer)", "data": "synthetic",
        "config": {"workload": "configs[1]: 2048-bit key raw_encrypt + raw_decrypt, bounded sample of %d elements per step" % total,
                   "key_bits": KEY_BITS},
        "cpu_baseline": {"value": enc, "unit": "encrypts/s", "decrypts_per_s": dec, "cores": cores, "kind": "port",
                         "sample": last["sample"], "engine": "oracle port of phe bound to libgmp mpz_powm (what gmpy2.powmod wraps)"
                         if last["backend"] == "gmp" else "oracle port of phe, Python pow"},
        "e2e": {"value": enc, "unit": "encrypts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# --------------------------------------------------------------------------------------------- GPU side
def measured_int_peak():
    """Peak 32x32->64 MAC rate of the integer pipe, measured by bench_micro/imad_peak (IMAD.WIDE.U32.X chains)."""
    exe = os.path.join(ROOT, "bench_micro", "imad_peak")
    fallback = {"mac_per_clk_sm": 25.1, "source": "profiles/r01_imad_peak.json (earlier measurement on this pool)"}
    if not os.path.exists(exe):
        return fallback
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
        js = json.loads(out)
        best = max((r for r in js["results"] if "wide" in r["op"]), key=lambda r: r["thread_ops_per_clk_per_sm"])
        return {"mac_per_clk_sm": best["thread_ops_per_clk_per_sm"], "op": best["op"], "mhz": best["eff_mhz"], "sms": js["sms"],
                "source": "bench_micro/imad_peak run inside this bench"}
    except Exception as e:     # noqa: BLE001
        fallback["error"] = str(e)[:100]
        return fallback


def _ncu_traffic(batch):
    """DRAM bytes (read + write) of one encrypt launch of `batch` rows, from the committed `ncu --set full` capture
    (profiles/r01_ncu_traffic.json: measured on one wave of 33 152 rows, scaled per row).  Algorithmic bytes are 1 KB
    per encrypt; the rest is the per-thread window table spilling from L2."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_ncu_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        return {"bytes_per_launch": t["bytes_per_ciphertext"] * batch, "bytes_per_ciphertext": t["bytes_per_ciphertext"],
                "algorithmic_bytes_per_ciphertext": 1024, "source": t["source"]}
    except (OSError, KeyError, ValueError):
        return None


_OUT = None


def _emit(obj):
    out = _OUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=DEFAULT_BATCH, help="elements per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-per-core", type=int, default=48)
    ap.add_argument("--cpu-per-core", type=int, default=48)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    # the contract is ONE JSON line on stdout: native libraries (NCCL's version banner, ...) write to fd 1 as well, so
    # everything but the final line is sent to stderr
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import importlib
    H, load_golden = (lambda m: (m.H, m.load_golden))(importlib.import_module("python-paillier_b200.fixtures"))
    fx = load_golden("vectors_%d.json" % KEY_BITS)
    key = (H(fx["n"]), H(fx["p"]), H(fx["q"]))
    if args.impl == "reference":
        run_reference(args, key)
        return

    import numpy as np
    import torch
    import paillier_b200 as pb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")          # keep stdout to the one JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, p, q = key
    # key limbs travel from rank 0 to every rank over NCCL (a few hundred bytes); batches never move
    if world > 1:
        kl = torch.from_numpy(pb.ints_to_limbs([n, p, q], KEY_BITS // 32).view(np.int32).copy()).cuda()
        dist.broadcast(kl, 0)
        n, p, q = pb.limbs_to_ints(kl.cpu().numpy().view(np.uint32))
    pub = pb.PublicContext(n, device=local)
    priv = pb.PrivateContext(p, q, device=local)
    ln, lc = pub.n_limbs, pub.c_limbs
    B = args.batch
    g = torch.Generator(device="cuda")
    g.manual_seed(1234 + rank)
    top = KEY_BITS // 32

    def rand_lt_n(rows):
        t = torch.randint(-2 ** 31, 2 ** 31 - 1, (rows, ln), dtype=torch.int32, device="cuda", generator=g)
        t[:, top - 1:] = 0          # < 2^(kb-32) < n: uniform enough for throughput, always a valid plaintext / r
        return t
    d_m, d_r = rand_lt_n(B), rand_lt_n(B)
    d_r[:, 0] |= 1
    d_c = torch.empty((B, lc), dtype=torch.int32, device="cuda")
    d_d = torch.empty((B, ln), dtype=torch.int32, device="cuda")
    l2_flush = torch.empty(256 << 20, dtype=torch.int8, device="cuda")
    eng = pb.get_engine()

    def step():
        pub.encrypt_dev(d_m, d_r, d_c, B)
        priv.decrypt_dev(d_c, d_d, B)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    assert bool((d_d == d_m).all().item()), "decrypt(encrypt(m)) != m on device"
    # parity spot check against the oracle (sampled indices)
    if rank == 0:
        from oracle import paillier_oracle as orc
        orc.BACKEND = "gmp" if orc.have_gmp() else "python"
        idx = [0, 1, B // 2, B - 1]
        mi = pb.limbs_to_ints(d_m[idx].cpu().numpy().view(np.uint32))
        ri = pb.limbs_to_ints(d_r[idx].cpu().numpy().view(np.uint32))
        ci = pb.limbs_to_ints(d_c[idx].cpu().numpy().view(np.uint32))
        opub = orc.PublicConsts(n)
        assert ci == [orc.raw_encrypt(opub, a, b) for a, b in zip(mi, ri)], "device ciphertexts differ from the oracle"

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        l2_flush.zero_()                       # flush L2 between timed iterations (256 MiB > 126 MB L2)
        ev[i][0].record()
        pub.encrypt_dev(d_m, d_r, d_c, B)
        ev[i][1].record()
        priv.decrypt_dev(d_c, d_d, B)
        ev[i][2].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = eng.launch_count() - launches0
    enc_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    dec_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    t = torch.tensor([enc_ms, dec_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    enc_ms, dec_ms = t.tolist()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the host-pointer C ABI, pinned host buffers
    e2e = None
    if not args.no_e2e:
        Be = min(B, 1 << 18)
        h_m = torch.empty((Be, ln), dtype=torch.int32).pin_memory(); h_m.copy_(d_m[:Be])
        h_r = torch.empty((Be, ln), dtype=torch.int32).pin_memory(); h_r.copy_(d_r[:Be])
        h_c = torch.empty((Be, lc), dtype=torch.int32).pin_memory()
        h_d = torch.empty((Be, ln), dtype=torch.int32).pin_memory()

        def e2e_step():
            eng.check(eng.lib.pai_encrypt_host(pub.h, h_m.data_ptr(), h_r.data_ptr(), h_c.data_ptr(), Be))
            t1 = time.perf_counter()
            eng.check(eng.lib.pai_decrypt_host(priv.h, h_c.data_ptr(), h_d.data_ptr(), Be))
            return t1
        e2e_step()
        barrier()
        te, td = 0.0, 0.0
        for _ in range(max(1, args.steps)):
            t0 = time.perf_counter()
            t1 = e2e_step()
            t2 = time.perf_counter()
            te += t1 - t0; td += t2 - t1
        assert bool((h_d == h_m).all().item())
        te /= max(1, args.steps); td /= max(1, args.steps)
        tt = torch.tensor([te, td], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te, td = tt.tolist()
        e2e = {"value": world * Be / te, "unit": "encrypts/s", "decrypts_per_s": world * Be / td,
               "batch_per_gpu": Be, "h2d_bytes_per_step": Be * (2 * ln + lc) * 4, "d2h_bytes_per_step": Be * (lc + ln) * 4,
               "api": "pai_encrypt_host / pai_decrypt_host (C ABI, pinned host buffers)"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    enc_per_s = world * B / (enc_ms * 1e-3)
    dec_per_s = world * B / (dec_ms * 1e-3)
    mac_enc, mac_dec = canonical_macs(KEY_BITS)
    peak = measured_int_peak()
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback"
    if os.path.exists(peaks_file):
        hbm_peak, hbm_src = json.load(open(peaks_file))["hbm_gbs"], "measured"
    sm_mhz = (clocks or {}).get("sm_mhz") or peak.get("mhz") or 1965.0
    peak_mac_s = peak["mac_per_clk_sm"] * 148 * sm_mhz * 1e6
    ach_mac_s = enc_per_s / world * mac_enc
    roofline = {
        "bound": "int_pipe", "kernel": "k_body<EncDigitBody<8>> (raw_encrypt, r^n mod n^2 on base-n digits)",
        "achieved": ach_mac_s / 1e12, "peak": peak_mac_s / 1e12, "unit": "TMAC/s (32x32->64, canonical count, per GPU)",
        "frac": ach_mac_s / peak_mac_s,
        "note": "canonical MAC counts (SURVEY 8d) give no credit for squaring / sliding windows, so frac may exceed 1; "
                "executed_* counts the MACs the kernels really issue",
        "executed_achieved": enc_per_s / world * executed_macs_2048()[0] / 1e12,
        "executed_frac": enc_per_s / world * executed_macs_2048()[0] / peak_mac_s,
        "peak_source": "measured IMAD.WIDE.U32.X rate %.1f MAC/clk/SM (%s) x 148 SMs x %.0f MHz (SM clock sampled under load)"
                       % (peak["mac_per_clk_sm"], peak.get("source"), sm_mhz),
        "decrypt": {"achieved": dec_per_s / world * mac_dec / 1e12, "frac": dec_per_s / world * mac_dec / peak_mac_s,
                    "executed_frac": dec_per_s / world * executed_macs_2048()[1] / peak_mac_s},
        "hbm": {"achieved_gbs": enc_per_s / world * (ln * 2 + lc) * 4 / 1e9, "peak_gbs": hbm_peak, "peak_source": hbm_src,
                "frac": enc_per_s / world * (ln * 2 + lc) * 4 / 1e9 / hbm_peak},
        "traffic": _ncu_traffic(args.batch),
    }
    cpu = None
    if not args.no_cpu:
        cores = os.cpu_count() or 1
        c = cpu_sample(key[0], key[1], key[2], args.cpu_per_core, cores)
        cpu = {"value": c["enc_per_s"], "unit": "encrypts/s", "decrypts_per_s": c["dec_per_s"], "cores": cores, "kind": "port",
               "sample": c["sample"], "engine": "oracle port of phe bound to libgmp mpz_powm (what gmpy2.powmod wraps)"
               if c["backend"] == "gmp" else "oracle port of phe, Python pow"}
    line = {
        "metric": "paillier_raw_encrypts_per_sec_2048", "value": enc_per_s, "unit": "encrypts/s",
        "decrypt": {"value": dec_per_s, "unit": "decrypts/s", "ms_per_step": dec_ms},
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": enc_ms,
        "step_ms_enc_plus_dec": enc_ms + dec_ms, "wall_s_timed_region": t_wall,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (exact integer)", "data": "synthetic",
        "config": {"workload": "configs[1]: 2048-bit key, raw_encrypt + raw_decrypt, batch %d per GPU" % B, "key_bits": KEY_BITS,
                   "batch_per_gpu": B, "parallelism": "batch sharded over %d GPU(s), no data-path collective" % world,
                   "l2": "256 MiB flush between timed iterations; inputs (%.0f MB) exceed L2" % (B * (2 * ln + lc) * 4 / 1e6)},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "targets": {"encrypts_per_s_1gpu": 1e5, "decrypts_per_s_1gpu": 2e5},
    }
    _emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

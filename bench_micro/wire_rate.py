"""Decimal wire format: GPU radix kernels vs CPython str()/int() on 4096-bit ciphertext rows (prints one JSON line)."""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("python-paillier_b200")
eng = importlib.import_module("python-paillier_b200.engine")

batch, lc = int(os.environ.get("WIRE_BATCH", 200000)), 128
rng = np.random.default_rng(0)
c = rng.integers(0, 2 ** 32, size=(batch, lc), dtype=np.uint32)
d_c = torch.from_numpy(c.view(np.int32)).cuda()
width = eng.decimal_width(lc)
d_text = torch.empty((batch, width), dtype=torch.uint8, device="cuda")
d_back = torch.empty_like(d_c)
d_status = torch.zeros((batch,), dtype=torch.int32, device="cuda")
for _ in range(2):
    eng.limbs_to_decimal_dev(d_c, lc, d_text, batch)
    eng.decimal_to_limbs_dev(d_text, width, d_back, lc, d_status, batch)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); eng.limbs_to_decimal_dev(d_c, lc, d_text, batch)
e[1].record(); eng.decimal_to_limbs_dev(d_text, width, d_back, lc, d_status, batch)
e[2].record(); torch.cuda.synchronize()
ok = bool((d_back == d_c).all().item())
sample = pkg.limbs_to_ints(c[:3000])
t = time.perf_counter(); ss = [str(v) for v in sample]; t_str = (time.perf_counter() - t) / len(sample)
t = time.perf_counter(); vv = [int(s) for s in ss]; t_int = (time.perf_counter() - t) / len(sample)
# whole path a caller sees: device rows -> list of Python strings
t = time.perf_counter()
raw = d_text.cpu().numpy().tobytes()
strs = [(raw[i * width:(i + 1) * width].lstrip(b"0") or b"0").decode("ascii") for i in range(batch)]
t_host = (time.perf_counter() - t) / batch
print(json.dumps({"workload": "4096-bit rows <-> decimal text", "batch": batch, "roundtrip_ok": ok and strs[:3000] == ss,
                  "gpu_to_decimal_us_per_row": e[0].elapsed_time(e[1]) * 1e3 / batch,
                  "gpu_from_decimal_us_per_row": e[1].elapsed_time(e[2]) * 1e3 / batch,
                  "d2h_plus_python_slicing_us_per_row": t_host * 1e6,
                  "cpython_str_us_per_row": t_str * 1e6, "cpython_int_us_per_row": t_int * 1e6}))

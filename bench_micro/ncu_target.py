"""Target for ncu captures: one wave of raw_encrypt and raw_decrypt at 2048 bit (device-resident inputs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paillier_b200 as pb
import importlib
_fx = importlib.import_module("python-paillier_b200.fixtures")
H, load_golden = _fx.H, _fx.load_golden

kb = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
waves = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
fx = load_golden("vectors_%d.json" % kb)
pub = pb.PublicContext(H(fx["n"])); priv = pb.PrivateContext(H(fx["p"]), H(fx["q"]))
batch = int(pub.wave() * waves)
ln, lc = pub.n_limbs, pub.c_limbs
rng = np.random.default_rng(1)
m = rng.integers(0, 2**32, size=(batch, ln), dtype=np.uint32); m[:, kb // 32 - 1:] = 0
r = rng.integers(0, 2**32, size=(batch, ln), dtype=np.uint32); r[:, kb // 32 - 1:] = 0
dm, dr = torch.from_numpy(m.view(np.int32)).cuda(), torch.from_numpy(r.view(np.int32)).cuda()
dc = torch.empty((batch, lc), dtype=torch.int32, device="cuda"); dd = torch.empty((batch, ln), dtype=torch.int32, device="cuda")
for _ in range(2):
    pub.encrypt_dev(dm, dr, dc, batch)
    priv.decrypt_dev(dc, dd, batch)
torch.cuda.synchronize()
assert bool((dd == dm).all().item())
print("ok", batch)

"""Where does the time of the federated ring-sum go?  (development probe)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paillier_b200 as phe, importlib
fx = importlib.import_module("python-paillier_b200.fixtures")
n, p, q = fx.fixed_key(2048)
pk = phe.PaillierPublicKey(n); sk = phe.PaillierPrivateKey(pk, p, q)
D = 100000
grads = [np.random.RandomState(43 + i).randn(D) * 0.1 for i in range(3)]
enc = [pk.encrypt_batch(g) for g in grads]
torch.cuda.synchronize()
vec = importlib.import_module("python-paillier_b200.vector")
orig = vec.EncryptedVector.decrease_exponent_to
tt = {"dec": 0.0, "n": 0}
def timed(self, e):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = orig(self, e)
    torch.cuda.synchronize(); tt["dec"] += time.perf_counter() - t0; tt["n"] += 1
    return r
vec.EncryptedVector.decrease_exponent_to = timed
for rep in range(3):
    tt["dec"] = 0.0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    acc = enc[0] + enc[1]
    torch.cuda.synchronize(); t1 = time.perf_counter()
    acc = acc + enc[2]
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("add1 %.3f add2 %.3f of which decrease_exponent_to %.3f" % (t1 - t0, t2 - t1, tt["dec"]))
ctx = pk.engine_context()
idx = np.nonzero(enc[0].exponents > enc[0].exponents.min())[0]
sub = enc[0].limbs[torch.from_numpy(idx).cuda()].contiguous()
scal = torch.zeros((len(idx), ctx.n_limbs), dtype=torch.int32, device="cuda"); scal[:, 0] = 16
out = torch.empty_like(sub); st = torch.zeros(len(idx), dtype=torch.int32, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.raw_mul_dev(sub, scal, out, st, len(idx))
    torch.cuda.synchronize(); print("raw_mul of %d rows by 16: %.4f s" % (len(idx), time.perf_counter() - t0))

"""Phase profile of the tensor-core encrypt kernel: PAI_TC_PROF dump summarised (cycles per op per warp)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, json
import paillier_b200 as pb, importlib
_fx = importlib.import_module("python-paillier_b200.fixtures")
kb = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n, p, q = _fx.fixed_key(kb)
pub = pb.PublicContext(n)
B = pub.wave()
m = torch.empty((B, pub.n_limbs), dtype=torch.int32, device="cuda"); r = torch.empty_like(m)
pub.random_lt_n_dev(m, B, seed=b"\x02" * 32, nonce=0); pub.random_lt_n_dev(r, B, seed=b"\x02" * 32, nonce=1)
c = torch.empty((B, pub.c_limbs), dtype=torch.int32, device="cuda")
pub.encrypt_dev(m, r, c, B); torch.cuda.synchronize()
os.environ["PAI_TC_PROF"] = "/tmp/tc_prof.txt"
pub.encrypt_dev(m, r, c, B); torch.cuda.synchronize()
del os.environ["PAI_TC_PROF"]
a = np.loadtxt("/tmp/tc_prof.txt")
names = ["P1sqr", "gemm1_wait", "epi_m", "gemm2_wait", "epi_t", "P2sqr", "epi_z", "z0copy", "P1mul", "P2mul", "n_sqr", "n_mul"]
tot = a[:, :10].sum()
nsq, nmul = a[:, 10].mean(), a[:, 11].mean()
out = {"warps": int(a.shape[0]), "ops_sqr": nsq, "ops_mul": nmul, "cycles_per_warp_total": a[:, :10].sum(axis=1).mean()}
for i, nm in enumerate(names[:10]):
    out[nm] = {"share": a[:, i].sum() / tot, "cycles_per_op": a[:, i].mean() / (nsq + nmul if i in (1, 2, 3, 4, 6, 7) else (nsq if i in (0, 5) else nmul))}
print(json.dumps(out, indent=1))

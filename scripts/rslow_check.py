"""Is there a bias in the FP32 production chain of the hidden-rate sampler?  Means of the slow statistics with
batch-means standard errors: oracle (R order), GPU f64, GPU f32, long chains."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import cases
import phylomap_b200 as pb
from oracle import bridge
Q, pid = cases.q4(), np.full(4, 0.25)
z = cases.tree_hidden(Q, T=24, S=2, seed=4, mean_branch=0.5)
N = int(os.environ.get("N", 200000))
def summ(a):
    a = a[2000:]
    out = {}
    for name, x in [("R_slow", a[:, 0:2].sum(1)), ("R_even", a[:, 0:4:2].sum(1)), ("l01", a[:, 20]), ("gamma", a[:, 24]), ("kap>", a[:, 22]),
                    ("changes", a[:, [5, 6, 7, 8, 10, 11, 12, 13, 15, 16, 17, 18]].sum(1)), ("virtual", a[:, [4, 9, 14, 19]].sum(1))]:
        nb = 50
        b = x[: len(x) // nb * nb].reshape(nb, -1).mean(1)
        out[name] = (round(float(x.mean()), 5), round(float(b.std(ddof=1) / np.sqrt(nb)), 5))
    return out
res = {}
res["oracle_a"] = summ(bridge.OracleRun(bridge.KS, [z.oracle_dict()], Q.copy(), pid, 4.0, N, prior=cases.PRIOR_KS, rng_mode=bridge.SEQUENTIAL, seed=5).run())
res["oracle_b"] = summ(bridge.OracleRun(bridge.KS, [z.oracle_dict()], Q.copy(), pid, 4.0, N, prior=cases.PRIOR_KS, rng_mode=bridge.SEQUENTIAL, seed=6).run())
for prec in ("f64", "f32"):
    for seed in (31, 32):
        res["gpu_%s_%d" % (prec, seed)] = summ(pb.sumstatMCMCks(z, np.asfortranarray(Q.copy()), pid, 4.0, N, cases.PRIOR_KS, seed=seed, precision=prec))
for k, v in res.items():
    print(k, v)

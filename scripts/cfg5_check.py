"""BASELINE configs[4] shapes on one GPU: the rate-updating samplers (ks, ksmt, mt) on the 10 000-tip tree with a
GPU's share of the sites.  Checks the invariants of every row and reports the time per sweep."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import cases
import phylomap_b200 as pb
from phylomap_b200 import synth
S = int(os.environ.get("PM_BENCH_SITES", 125000))
Q4, pid4 = cases.q4(), np.full(4, 0.25)
tree = synth.yule_tree(10000, seed=4, mean_branch=0.1 / 1.2)
t0 = time.time()
zk = synth.simulate_4_state_tree(7, tree, Q4, pid4, n_sites=S, device="cuda", segments=2)
print("data", round(time.time() - t0, 1), "s", flush=True)
out = {}
def timed(name, fn, N):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    out[name] = {"s_per_sweep": dt / N, "histories_per_s": tree.E * S * N / dt}
    return r
N = 6
ks = timed("ks", lambda: pb.sumstatMCMCks(zk, np.asfortranarray(Q4.copy()), pid4, 4.0, N, cases.PRIOR_KS, precision="f32", seed=3), N)
assert np.allclose(ks[:, :4].sum(1), S * tree.edge_length.sum(), rtol=2e-4)
assert np.all(ks[:, 4:20] >= 0) and np.all(ks[:, 20:25] > 0)
out["ks"]["rates_last"] = ks[-1, 20:25].tolist()
dic = timed("ksDICt", lambda: pb.sumstatMCMCksDICt(zk, np.asfortranarray(Q4.copy()), pid4, 4.0, N, cases.PRIOR_KS, precision="f32", seed=3), N)
assert np.all(np.isfinite(dic[:, -1])) and np.all(dic[:, -1] < 0)
out["ksDICt"]["loglik_last"] = float(dic[-1, -1])
trees = [zk, pb.PhyloTree(zk.edge, zk.edge_length * 1.1).with_states(zk.states, segments=2)]
half = [pb.PhyloTree(t.edge, t.edge_length, t.states[: S // 2], t.maps, t.mapnames) for t in trees]
ksmt = timed("ksmt_2trees_half_sites", lambda: pb.sumstatMCMCksmt(half, np.asfortranarray(Q4.copy()), pid4, 4.0, N, cases.PRIOR_KSMT, precision="f32", seed=3), N)
assert set(np.unique(ksmt[:, -1])) <= {0.0, 1.0}
Q2 = np.array([[-0.1, 0.1], [0.1, -0.1]])
z2 = synth.simulate_2_state_tree(9, tree, Q2, cases.PID2, n_sites=S, device="cuda", segments=2)
bf = timed("bf", lambda: pb.sumstatMCMCbf(z2, np.asfortranarray(Q2.copy()), cases.PID2, 0.5, N, cases.PRIOR_BF, precision="f32", seed=3), N)
assert np.allclose(bf[:, :2].sum(1), S * tree.edge_length.sum(), rtol=2e-4)
out["bf"]["rates_last"] = bf[-1, 6:8].tolist()
print(json.dumps(out))

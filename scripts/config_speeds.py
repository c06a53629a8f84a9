"""Sweep times at the other BASELINE configurations (bench.py times configs[3] at one GPU's share): configs[1] = 1 000 tips x
10 000 sites, 4 states, sumstatMCMC; configs[2] = the Squamate tree (3 951 tips) x 100 000 sites, 2 states,
SPARSEsumstatMCMC.  Resident chain, FP32 and FP64 production arithmetic."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import cases
import phylomap_b200 as pb
from phylomap_b200 import capi, synth

def timed(variant, z, Q, pid, Om, prec, n=10):
    ch = pb.Chain(variant, z, Q, pid, Om, 3 * n + 20, precision=prec, seed=3)
    ch.run(20)
    best = 1e30
    for _ in range(2):
        torch.cuda.synchronize(); t = time.perf_counter(); ch.run(n); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    ch.enable_timing(True); ch.run(n); kt, _ = ch.kernel_times()
    E, S = z.E, z.n_sites()
    r = {"ms_per_sweep": 1e3 * best / n, "histories_per_s": E * S * n / best, "in_sweep_ms": {k: round(v / n, 3) for k, v in kt.items()},
         "device_bytes_per_branch_site": ch.device_bytes() / (E * S)}
    ch.close()
    capi.lib().pm_release_cached_memory()
    return r

out = {}
Q, pid = cases.jc(4, 0.1), np.full(4, 0.25)
tree = synth.yule_tree(1000, seed=2, mean_branch=1.0)
st = synth.simulate_tip_states(tree, Q, pid, 10000, seed=202, device="cuda").cpu().numpy()
z = tree.with_states(st, segments=2)
for prec in ("f32", "f64"):
    out["configs1_1000tips_10000sites_%s" % prec] = timed(capi.PM_V_PLAIN, z, Q, pid, 0.6, prec)
Q2 = np.array([[-0.001, 0.001], [0.006, -0.006]])
zs = synth.simulate_2_state_tree(5, cases.squamate_tree(), Q2, cases.PID2, n_sites=100000, device="cuda", segments=8)
for prec in ("f32", "f64"):
    out["configs2_squamate_100000sites_%s" % prec] = timed(capi.PM_V_SPARSE, zs, Q2, cases.PID2, 0.06, prec)
print(json.dumps(out))

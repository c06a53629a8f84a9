"""SURVEY H6 on one B200: (i) 300 000 sites of the 10 000-tip configuration resident at once; (ii) configs[3]'s full
1 000 000 sites through ONE call of the drop-in entry: the state does not fit (214 GB), the call cuts itself into site
tiles (pm_host.cu, one_call) instead of failing in cudaMalloc."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import cases
import phylomap_b200 as pb
from phylomap_b200 import capi, synth
Q, pid = cases.q4(), np.full(4, 0.25)
tree = synth.yule_tree(10000, seed=4, mean_branch=0.1 / 1.2)
out = {}
def data(S, seed):
    host = torch.empty((S, tree.T), dtype=torch.uint8)
    for s0 in range(0, S, 100000):
        n = min(100000, S - s0)
        host[s0:s0 + n].copy_(synth.simulate_tip_states(tree, Q, pid, n, seed=seed + s0, device="cuda", batch_sites=32768))
    torch.cuda.empty_cache()
    return tree.with_states(host.numpy(), segments=2)
S = int(os.environ.get("BIG_RESIDENT", 300000))
z = data(S, 11)
ch = pb.Chain(capi.PM_V_BIGTREE, z, Q.copy(), pid, 2.4, 6, precision="f32", seed=3)
ch.run(3)
torch.cuda.synchronize(); t = time.perf_counter(); rows = ch.run(3); torch.cuda.synchronize(); dt = time.perf_counter() - t
assert np.allclose(rows[:, :4].sum(1), S * tree.edge_length.sum(), rtol=2e-4)
out["resident_sites"] = S
out["resident_device_GB"] = ch.device_bytes() / 1e9
out["resident_ms_per_sweep"] = 1e3 * dt / 3
out["resident_histories_per_s"] = tree.E * S * 3 / dt
ch.close(); del ch, z
capi.lib().pm_release_cached_memory()
S = int(os.environ.get("BIG_TILED", 1000000))
z = data(S, 12)
N = 3
torch.cuda.synchronize(); t = time.perf_counter()
rows = pb.sumstatMCMC_bigtree(z, Q.copy(), pid, 2.4, N, precision="f32", seed=5)
torch.cuda.synchronize(); dt = time.perf_counter() - t
assert np.allclose(rows[:, :4].sum(1), S * tree.edge_length.sum(), rtol=2e-4)
out["tiled_sites"] = S
out["tiled_call_seconds"] = dt
out["tiled_histories_per_s_e2e"] = tree.E * S * N / dt
print(json.dumps(out))

"""Times the production pruning kernel alone (pm_chain_time_prune) for its variants (PHYLOMAP_B200_K1_UNROLL:
42 = clade order [default], 21 = level order, one node per round, 20 = two; see pm_launch_impl.cuh) at the benchmark size, and checks that they give identical rows."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import phylomap_b200 as pb
from phylomap_b200 import capi, synth
tree, Q, pid = bench.workload_tree()
S = int(os.environ.get("PM_BENCH_SITES", 125000))
st = synth.simulate_tip_states(tree, Q, pid, S, seed=101, device="cuda", batch_sites=32768).cpu().numpy()
z = tree.with_states(st, segments=2)
T, E = tree.T, tree.E
bytes_site = (T - 1) * 16 + (T - 2) * 16 + T + 4 * E
for v in sys.argv[1:] or ["21", "20"]:
    os.environ["PHYLOMAP_B200_K1_UNROLL"] = v
    ch = pb.Chain(capi.PM_V_BIGTREE, z, Q.copy(), pid, 2.4, 20, precision="f32", seed=1)
    rows = ch.run(6)
    if "ref_rows" not in globals():
        ref_rows = rows
    same = bool(np.array_equal(rows, ref_rows))
    ms = ch.time_prune(reps=10)
    print(json.dumps({"same_rows_as_first_variant": same, "unroll": v, "k1_ms": ms, "GBps": bytes_site * S / ms / 1e6, "frac": bytes_site * S / ms / 1e6 / 6530.3}))
    ch.close(); del ch
    torch.cuda.empty_cache()

"""Where the end-to-end call spends its time: chain build (allocation, upload, init), sweeps, tear-down."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import phylomap_b200 as pb
from phylomap_b200 import capi, synth
tree, Q, pid = bench.workload_tree()
S = int(os.environ.get("PM_BENCH_SITES", 125000))
st = synth.simulate_tip_states(tree, Q, pid, S, seed=101, device="cuda", batch_sites=32768)
host = torch.empty((S, tree.T), dtype=torch.uint8, pin_memory=True); host.copy_(st); del st; torch.cuda.empty_cache()
z = tree.with_states(host.numpy(), segments=2)
order = [z.order()]
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ch = pb.Chain(capi.PM_V_BIGTREE, z, Q.copy(), pid, 2.4, 10, order=order, precision="f32", seed=1)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ch.run(10)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    ch.close()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(json.dumps({"create_s": t1 - t0, "run10_s": t2 - t1, "destroy_s": t3 - t2}))

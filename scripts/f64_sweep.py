"""The benchmark's sweep in FP64 production arithmetic (the library's DEFAULT precision; bench.py times FP32): ms per sweep and
the split over the kernels, same tree / model / sites."""
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
out = {}
for prec in ("f64", "f32"):
    ch = pb.Chain(capi.PM_V_BIGTREE, z, Q.copy(), pid, 2.4, 16, order=[z.order()], precision=prec, seed=1)
    ch.run(6)
    ch.enable_timing(True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); ch.run(8); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    kt, _ = ch.kernel_times()
    out[prec] = {"ms_per_sweep": 1e3 * dt / 8, "histories_per_s": tree.E * S * 8 / dt, "in_sweep_ms": {k: v / 8 for k, v in kt.items()},
                 "device_bytes_per_branch_site": ch.device_bytes() / (tree.E * S)}
    ch.close(); del ch
    capi.lib().pm_release_cached_memory()
print(json.dumps(out))

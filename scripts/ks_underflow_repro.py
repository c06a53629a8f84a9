"""Reproduces the FP32 zero-vector failure seen on rank 6 of the 8-GPU bench (rate_sampler block): the ks chain on that
rank's 125 000 sites (seed 101 + 6)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import phylomap_b200 as pb
from phylomap_b200 import capi, synth
import bench
rank = int(os.environ.get("REPRO_RANK", 6))
tree, Q, pid = bench.workload_tree()
S = 125000
st = synth.simulate_tip_states(tree, Q, pid, S, seed=101 + rank, device="cuda", batch_sites=32768).cpu().numpy()
par = ((st.astype(np.int64) - 1) % 2 + 1).astype(np.uint8)
zk = tree.with_states(par, segments=2)
prior = np.array([1.0, 10.0, 2.0, 10.0, 20.0, 2.0])
ch = pb.Chain(capi.PM_V_KS, zk, np.asfortranarray(Q.copy()), pid, 4.0, 12, prior=prior, precision="f32", seed=2026, site_offset=rank * S)
try:
    out = ch.run(12)
    print("ok", out[-1, 20:25])
except capi.PhylomapError as e:
    print("FAILED", e)

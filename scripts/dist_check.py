"""Multi-GPU check (run under torchrun, one rank per GPU): site-sharded runs must reproduce the unsharded single-GPU run --
integer counts exactly, dwell times / rates to rounding -- with the library's own NCCL all-reduce (pm_options.nccl_*) and
with the torch.distributed callback; a device error on ONE rank must stop ALL ranks in the same sweep (no hang).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/dist_check.py
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, torch.distributed as dist
import cases
import phylomap_b200 as pb
from phylomap_b200 import capi, synth, dist as pdist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
stream = torch.cuda.Stream()
report = {}

def run(variant, z, Q, pid, Om, N, prior=None, native=True, **kw):
    local, start = pdist.shard_tree(z, rank, world)
    if native:   # the library joins its own NCCL clique and calls ncclAllReduce on its private stream
        opts = dict(device=rank, site_offset=start, nccl=pdist.nccl_clique(rank, world), **kw)
    else:
        opts = dict(device=rank, site_offset=start, stream=stream.cuda_stream, allreduce=pdist.allreduce_callback(stream.cuda_stream), **kw)
    sharded = pb.Chain(variant, local, np.asfortranarray(Q.copy()), pid, Om, N, prior=prior, **opts).run()
    torch.cuda.synchronize()
    full = None
    if rank == 0:
        full = pb.Chain(variant, z, np.asfortranarray(Q.copy()), pid, Om, N, prior=prior, device=0, **kw).run()
    return sharded, full

Q4, pid4 = cases.q4(), np.full(4, 0.25)
det = dict(mode="deterministic", precision="f64", seed=5)
z = cases.tree_n(Q4, T=60, S=64 * world + 3, seed=3, mean_branch=0.4, segments=3)
sh, full = run(capi.PM_V_BIGTREE, z, Q4, pid4, 2.4, 6, **det)
if rank == 0:
    report["bigtree_counts_equal"] = bool(np.array_equal(sh[:, 4:], full[:, 4:]))
    report["bigtree_dwell_maxrel"] = float(np.max(np.abs(sh[:, :4] / full[:, :4] - 1)))
zk = cases.tree_hidden(Q4, T=40, S=32 * world + 1, seed=4, mean_branch=0.5)
sh, full = run(capi.PM_V_KS, zk, Q4, pid4, 4.0, 8, prior=cases.PRIOR_KS, **det)
gathered = [None] * world
dist.all_gather_object(gathered, sh.tolist())
if rank == 0:
    report["ks_rows_identical_on_all_ranks"] = all(np.array_equal(np.array(g), sh) for g in gathered)
    report["ks_counts_equal"] = bool(np.array_equal(sh[:, 4:20], full[:, 4:20]))
    report["ks_rates_maxrel"] = float(np.max(np.abs(sh[:, 20:25] / full[:, 20:25] - 1)))
sh2, _ = run(capi.PM_V_KS, zk, Q4, pid4, 4.0, 8, prior=cases.PRIOR_KS, native=False, **det)
if rank == 0:
    report["ks_callback_equals_native"] = bool(np.array_equal(sh2, sh))
# a device error on rank 1 only (its record slices are too small): every rank must raise in the same sweep
local, start = pdist.shard_tree(zk, rank, world)
try:
    pb.Chain(capi.PM_V_KS, local, np.asfortranarray(Q4.copy()), pid4, 4.0, 8, prior=cases.PRIOR_KS, device=rank, site_offset=start,
             nccl=pdist.nccl_clique(rank, world), path_capacity=(1 if rank == 1 else 0), **det).run()
    code = 0
except capi.PhylomapError as e:
    code = e.code
codes = [None] * world
dist.all_gather_object(codes, code)
if rank == 0:
    report["error_codes_per_rank"] = codes
    report["error_stops_all_ranks"] = bool(all(c != 0 for c in codes) and codes[1] == capi.PM_ERR_CAPACITY)
prod = dict(precision="f32", seed=9)
zp = cases.tree_n(Q4, T=300, S=512 * world, seed=6, mean_branch=0.3, segments=2)
sh, full = run(capi.PM_V_BIGTREE, zp, Q4, pid4, 2.4, 6, **prod)
if rank == 0:
    report["production_counts_equal"] = bool(np.array_equal(sh[:, 4:], full[:, 4:]))
    report["production_dwell_maxrel"] = float(np.max(np.abs(sh[:, :4] / full[:, :4] - 1)))
    ok = (report["ks_callback_equals_native"] and report["error_stops_all_ranks"] and report["bigtree_counts_equal"] and report["ks_rows_identical_on_all_ranks"] and report["ks_counts_equal"] and
          report["production_counts_equal"] and report["bigtree_dwell_maxrel"] < 1e-9 and report["ks_rates_maxrel"] < 1e-6 and
          report["production_dwell_maxrel"] < 1e-4)
    report["world"] = world
    report["ok"] = bool(ok)
    print(json.dumps(report))
dist.barrier()
dist.destroy_process_group()

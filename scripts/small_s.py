"""The reference's literal usage, one character per call: seconds per sweep of the GPU library
against the CPU port on ONE core, for BASELINE configs[0] (100 tips, 1 000 sweeps) and the Squamate vignette tree."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import cases
import phylomap_b200 as pb
from phylomap_b200 import synth
from oracle import bridge
out = {}
def gpu(fn, *a, **k):
    fn(*a, **k)  # warm
    best = 1e30
    for _ in range(3):   # best of three calls (a call of 1 000 sweeps of a 100-tip tree lasts 0.1 s: host jitter shows)
        torch.cuda.synchronize(); t = time.perf_counter(); r = fn(*a, **k); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best, r
def cpu(variant, z, Q, pid, Om, N, fast):
    o = bridge.OracleRun(variant, [z.oracle_dict()], Q, pid, Om, N, rng_mode=bridge.SEQUENTIAL, seed=3)
    o.set_fast_lookup(fast)
    t = time.perf_counter(); o.run(); return time.perf_counter() - t
# configs[0]
z = synth.simulate_2_state_tree(101, synth.yule_tree(100, seed=1, mean_branch=5.0), cases.Q2, cases.PID2)
N = 1000
def resident(zz, Q, pid, Om, n_sweeps, prec, variant=None):
    """us per sweep of a chain that already exists (creation, upload and the first call excluded): best of three calls"""
    from phylomap_b200 import capi
    ch = pb.Chain(capi.PM_V_PLAIN if variant is None else variant, zz, Q, pid, Om, 5 * n_sweeps, precision=prec, seed=5)
    ch.run(n_sweeps)
    best = 1e30
    for _ in range(3):
        torch.cuda.synchronize(); t = time.perf_counter(); ch.run(n_sweeps); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    ch.close()
    return 1e6 * best / n_sweeps
# the one-block-per-site chain kernel (pm_small.cuh, the default at this size) ...
for prec in ("f32", "f64"):
    dt, _ = gpu(pb.sumstatMCMC, z, cases.Q2, cases.PID2, 0.2, N, seed=5, precision=prec)
    out["cfg0_gpu_%s_one_block_per_site_us_per_sweep_whole_call" % prec] = 1e6 * dt / N
    out["cfg0_gpu_%s_one_block_per_site_us_per_sweep_resident" % prec] = resident(z, cases.Q2, cases.PID2, 0.2, N, prec)
# ... against the 32-sites-per-warp kernels
os.environ["PHYLOMAP_B200_SMALL"] = "0"
for prec in ("f32", "f64"):
    for graph in ("1", "0"):   # replay of a captured sweep (their default at this size) against launch-by-launch
        os.environ["PHYLOMAP_B200_GRAPH"] = graph
        dt, _ = gpu(pb.sumstatMCMC, z, cases.Q2, cases.PID2, 0.2, N, seed=5, precision=prec)
        out["cfg0_gpu_%s_wide_%s_us_per_sweep_whole_call" % (prec, "graph" if graph == "1" else "launches")] = 1e6 * dt / N
del os.environ["PHYLOMAP_B200_GRAPH"]
out["cfg0_gpu_f32_wide_us_per_sweep_resident"] = resident(z, cases.Q2, cases.PID2, 0.2, N, "f32")
for S in (8, 32, 148, 592, 2048):   # a few characters at once on the same tree, either set of kernels
    zS = synth.simulate_2_state_tree(101, synth.yule_tree(100, seed=1, mean_branch=5.0), cases.Q2, cases.PID2, n_sites=S)
    for small in ("1", "0"):
        os.environ["PHYLOMAP_B200_SMALL"] = small
        os.environ["PHYLOMAP_B200_SMALL_SITES"] = "1000000"
        out["cfg0_gpu_f32_%d_sites_%s_us_per_sweep_resident" % (S, "one_block_per_site" if small == "1" else "wide")] = \
            resident(zS, cases.Q2, cases.PID2, 0.2, 200, "f32")
del os.environ["PHYLOMAP_B200_SMALL"], os.environ["PHYLOMAP_B200_SMALL_SITES"]
out["cfg0_cpu_port_faithful_us_per_sweep"] = 1e6 * cpu(bridge.PLAIN, z, cases.Q2, cases.PID2, 0.2, N, False) / N
out["cfg0_cpu_port_optimised_us_per_sweep"] = 1e6 * cpu(bridge.PLAIN, z, cases.Q2, cases.PID2, 0.2, N, True) / N
t = time.perf_counter(); bridge.ref_run(bridge.PLAIN, [z.oracle_dict()], cases.Q2, cases.PID2, 0.2, N, seed=3)
out["cfg0_cpu_reference_us_per_sweep"] = 1e6 * (time.perf_counter() - t) / N
# the rate-updating samplers with one character (the tutorial's usage): one launch per sweep, the host update in between
from phylomap_b200 import capi
zb = synth.simulate_2_state_tree(101, synth.yule_tree(100, seed=1, mean_branch=5.0), cases.Q2, cases.PID2)
zk = cases.tree_hidden(cases.q4(), T=100, S=1, seed=4, mean_branch=1.0)
for name, var, ovar, zz, Qm, pidm, Om, prior in (("bf", capi.PM_V_BF, bridge.BF, zb, cases.Q2, cases.PID2, 0.4, cases.PRIOR_BF),
                                                  ("ks", capi.PM_V_KS, bridge.KS, zk, cases.q4(), np.full(4, 0.25), 4.0, cases.PRIOR_KS)):
    for small in ("1", "0"):
        os.environ["PHYLOMAP_B200_SMALL"] = small
        ch = pb.Chain(var, zz, np.asfortranarray(Qm.copy()), pidm, Om, 4 * N, prior=prior, precision="f32", seed=5)
        ch.run(N)
        b = 1e30
        for _ in range(3):
            torch.cuda.synchronize(); t = time.perf_counter(); ch.run(N); torch.cuda.synchronize()
            b = min(b, time.perf_counter() - t)
        ch.close()
        out["%s_gpu_f32_%s_us_per_sweep_resident" % (name, "one_block_per_site" if small == "1" else "wide")] = 1e6 * b / N
    del os.environ["PHYLOMAP_B200_SMALL"]
    o = bridge.OracleRun(ovar, [zz.oracle_dict()], np.asfortranarray(Qm.copy()), pidm, Om, N, prior=prior, rng_mode=bridge.SEQUENTIAL, seed=3)
    o.set_fast_lookup(True)
    t = time.perf_counter(); o.run(); out["%s_cpu_port_optimised_us_per_sweep" % name] = 1e6 * (time.perf_counter() - t) / N
    t = time.perf_counter(); bridge.ref_run(ovar, [zz.oracle_dict()], np.asfortranarray(Qm.copy()), pidm, Om, N, prior=prior, seed=3)
    out["%s_cpu_reference_us_per_sweep" % name] = 1e6 * (time.perf_counter() - t) / N
# the Squamate vignette: 3 951 tips x 100 segments, Omega = 10, one character
Q = np.array([[-0.001, 0.001], [0.006, -0.006]])
zs = synth.simulate_2_state_tree(101, cases.squamate_tree(), Q, cases.PID2)
N2 = 40
dt, _ = gpu(pb.sumstatMCMC_bigtree, zs, Q, cases.PID2, 10.0, N2, seed=5, precision="f64")
out["squamate_gpu_f64_one_block_per_site_ms_per_sweep"] = 1e3 * dt / N2
os.environ["PHYLOMAP_B200_SMALL"] = "0"
dt, _ = gpu(pb.sumstatMCMC_bigtree, zs, Q, cases.PID2, 10.0, N2, seed=5, precision="f64")
out["squamate_gpu_f64_wide_ms_per_sweep"] = 1e3 * dt / N2
del os.environ["PHYLOMAP_B200_SMALL"]
out["squamate_cpu_port_faithful_ms_per_sweep"] = 1e3 * cpu(bridge.BIGTREE, zs, Q, cases.PID2, 10.0, 4, False) / 4
out["squamate_cpu_port_optimised_ms_per_sweep"] = 1e3 * cpu(bridge.BIGTREE, zs, Q, cases.PID2, 10.0, 4, True) / 4
print(json.dumps(out))

"""k_paths_easy2 (two-run branches in a clean-up pass) against k_paths_easy (in line, PHYLOMAP_B200_TUNE=4): same seed,
same rows -- integer counts identical, dwell times to FP32 summation order."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
if len(sys.argv) > 1:
    import cases, phylomap_b200 as pb
    from phylomap_b200 import synth
    Q, pid = cases.q4(), np.full(4, 0.25)
    tree = synth.yule_tree(3000, seed=4, mean_branch=0.1 / 1.2)
    st = synth.simulate_tip_states(tree, Q, pid, 1000, seed=7, device="cuda").cpu().numpy()
    z = tree.with_states(st, segments=2)
    a = pb.sumstatMCMC_bigtree(z, Q, pid, 2.4, 8, seed=21, precision="f32")
    zk = synth.simulate_4_state_tree(9, tree, Q, pid, n_sites=777, device="cuda", segments=2)
    b = pb.sumstatMCMCks(zk, np.asfortranarray(Q.copy()), pid, 4.0, 6, cases.PRIOR_KS, seed=5, precision="f64")
    np.savez(sys.argv[1], a=a, b=b)
else:
    for tune, f in (("0", "/tmp/ab_new.npz"), ("4", "/tmp/ab_old.npz")):
        subprocess.check_call([sys.executable, __file__, f], env=dict(os.environ, PHYLOMAP_B200_TUNE=tune))
    n, o = np.load("/tmp/ab_new.npz"), np.load("/tmp/ab_old.npz")
    print(json.dumps({"bigtree_counts_equal": bool(np.array_equal(n["a"][:, 4:], o["a"][:, 4:])),
                      "bigtree_dwell_maxrel": float(np.max(np.abs(n["a"][:, :4] / o["a"][:, :4] - 1))),
                      "ks_counts_equal": bool(np.array_equal(n["b"][:, 4:20], o["b"][:, 4:20])),
                      "ks_rates_maxrel": float(np.max(np.abs(n["b"][:, 20:25] / o["b"][:, 20:25] - 1)))}))

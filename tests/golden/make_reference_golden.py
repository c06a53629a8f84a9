"""Writes tests/golden/reference/*.json: outputs of THE REFERENCE ITSELF, run in this container.

    python tests/golden/make_reference_golden.py

oracle/_ref/libphylomap_ref.so is the unmodified /root/reference/src/phylomap.cpp + RcppExports.cpp compiled against
the stand-in Rcpp / RcppArmadillo headers of oracle/standin/ (oracle/Makefile, target `ref`).  Each case calls one of the
ten `.Call` entry points (src/RcppExports.cpp:11-237) with R's generator seeded like set.seed(seed) and stores the
inputs, the returned matrix and Q / B as the call left them.  One character per call: the reference has no site axis.
The fixtures pin the oracle restatement (tests/test_reference_pin.py, CPU) and the CUDA library's deterministic mode
replaying the same uniform stream (tests/test_gpu_parity.py::test_gpu_replays_the_reference).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)


def eig_of(Q):
    """lefts, rights, d of R/sumstatEXP.R:21-33 (Q = left diag(d) right)."""
    w, V = np.linalg.eig(Q)
    return V.real, np.linalg.inv(V).real, np.diag(w.real)


def case_list():
    import cases
    from make_golden import _tree_json
    rng = np.random.default_rng(9)
    Q4 = cases.q4()
    Q6 = cases.q6()
    z = cases.tree2(T=12, S=1, seed=31)
    z4 = cases.tree_n(Q4, T=10, S=1, seed=32, mean_branch=0.6, segments=3)
    zk = cases.tree_hidden(Q4, T=10, S=1, seed=33, mean_branch=0.5)
    z6 = cases.tree_hidden(Q6, T=9, S=1, seed=35, mean_branch=0.5)
    Q20 = rng.uniform(0.01, 0.05, size=(20, 20))  # asymmetric: no exact ties (tie order for n > 16)
    np.fill_diagonal(Q20, 0)
    np.fill_diagonal(Q20, -Q20.sum(1))
    z20 = cases.tree_n(Q20, T=8, S=1, seed=36, mean_branch=2.0, segments=4)
    pid20 = rng.dirichlet(np.ones(20))
    sc2 = [rng.uniform(0.7, 1.3, size=z.E).tolist()]
    sck = [rng.uniform(0.7, 1.3, size=zk.E).tolist()]
    p4 = [0.25] * 4

    def c(variant, Q, pid, Om, N, seed, z, prior=None, scales=None, eig=False):
        d = {"variant": variant, "Q": np.asarray(Q).tolist(), "pid": list(map(float, pid)), "Omega": Om, "N": N, "seed": seed,
             "tree": _tree_json(z)}
        if prior is not None:
            d["prior"] = list(map(float, prior))
        if scales:
            d["extra_tree_scales"] = scales
        if eig:
            d["eig"] = [m.tolist() for m in eig_of(np.asarray(Q))]
        return d

    return [
        ("plain_2state", c("PLAIN", cases.Q2, cases.PID2, 0.2, 12, 101, z)),
        ("sparse_2state", c("SPARSE", cases.Q2, cases.PID2, 0.2, 12, 102, z)),
        ("bigtree_2state", c("BIGTREE", cases.Q2, cases.PID2, 0.2, 12, 103, z)),
        ("plain_4state", c("PLAIN", Q4, p4, 2.4, 8, 104, z4)),
        ("sparse_4state", c("SPARSE", Q4, p4, 2.4, 8, 105, z4)),
        ("bigtree_4state", c("BIGTREE", Q4, p4, 2.4, 8, 106, z4)),
        ("plain_20state", c("PLAIN", Q20, pid20, 1.5, 5, 107, z20)),
        ("bf_2state", c("BF", cases.Q2, cases.PID2, 0.5, 12, 108, z, prior=cases.PRIOR_BF)),
        ("ks_4state", c("KS", Q4, p4, 4.0, 10, 109, zk, prior=cases.PRIOR_KS)),
        ("ks_6state", c("KS", Q6, [1 / 6] * 6, 8.0, 6, 110, z6, prior=cases.PRIOR_KS)),
        ("mt_2state", c("MT", cases.Q2, cases.PID2, 0.5, 10, 111, z, prior=cases.PRIOR_BF, scales=sc2)),
        ("ksmt_4state", c("KSMT", Q4, p4, 4.0, 8, 112, zk, prior=cases.PRIOR_KSMT, scales=sck)),
        ("dic2s_2state", c("DIC2S", cases.Q2, cases.PID2, 0.5, 8, 113, z, prior=cases.PRIOR_BF)),
        ("dicks_4state", c("DICKS", Q4, p4, 4.0, 6, 114, zk, prior=cases.PRIOR_KS)),
        ("exp_2state", c("EXP", cases.Q2, cases.PID2, 0.2, 10, 115, z, eig=True)),
        ("exp_4state", c("EXP", Q4, p4, 2.4, 8, 116, z4, eig=True)),
    ]


def trees_of(case):
    from make_golden import tree_from_case
    from phylomap_b200 import PhyloTree
    z = tree_from_case(case)
    trees = [z]
    for sc in case.get("extra_tree_scales", []):
        sc = np.array(sc)
        trees.append(PhyloTree(z.edge, z.edge_length * sc, z.states, [m * s for m, s in zip(z.maps, sc)], z.mapnames))
    return trees


def run_reference(bridge, case):
    eig = [np.array(m) for m in case["eig"]] if "eig" in case else None
    return bridge.ref_run(getattr(bridge, case["variant"]), [t.oracle_dict() for t in trees_of(case)], np.array(case["Q"]),
                          np.array(case["pid"]), case["Omega"], case["N"], prior=case.get("prior"), seed=case["seed"], eig=eig)


def main():
    from oracle import bridge
    if bridge.ref_lib() is None:
        raise SystemExit("oracle/_ref is not built: needs /root/reference")
    os.makedirs(os.path.join(HERE, "reference"), exist_ok=True)
    for name, case in case_list():
        rows, Q, B = run_reference(bridge, case)
        with open(os.path.join(HERE, "reference", name + ".json"), "w") as f:
            json.dump({"case": case, "rows": rows.tolist(), "Q_after": Q.tolist(), "B_after": B.tolist()}, f)
        print(name, rows.shape)


if __name__ == "__main__":
    main()

"""Writes tests/golden/*.json: small self-contained cases (inputs + the oracle's output rows).

    python tests/golden/make_golden.py

The reference has no golden vectors for this path and cannot be run here (R/Rcpp absent), so these rows come from
the in-repo oracle; they freeze its behaviour (regression guard) and give the GPU tests fixed vectors to hit.
The GPU parity tests replay the same cases through the CUDA library in deterministic mode.
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


def tree_from_case(case):
    from phylomap_b200 import PhyloTree
    t = case["tree"]
    return PhyloTree(np.array(t["edge"]), np.array(t["edge_length"]), np.array(t["states"], dtype=np.int32),
                     [np.array(m) for m in t["maps"]], [np.array(m, dtype=np.int32) for m in t["mapnames"]])


def run_case(oracle, case):
    z = tree_from_case(case)
    trees = [z]
    if case.get("extra_tree_scales"):
        from phylomap_b200 import PhyloTree
        for sc in case["extra_tree_scales"]:
            el = z.edge_length * np.array(sc)
            maps = [m * s for m, s in zip(z.maps, sc)]
            trees.append(PhyloTree(z.edge, el, z.states, maps, z.mapnames))
    o = oracle.OracleRun(getattr(oracle, case["variant"]), [t.oracle_dict() for t in trees], np.array(case["Q"]),
                         np.array(case["pid"]), case["Omega"], case["N"], prior=case.get("prior"),
                         rng_mode=oracle.KEYED, seed=case["seed"])
    return o.run()


def _tree_json(z):
    st = z.states if z.states.ndim == 2 else z.states[None, :]
    return {"edge": z.edge.tolist(), "edge_length": z.edge_length.tolist(), "states": st.astype(int).tolist(),
            "maps": [m.tolist() for m in z.maps], "mapnames": [m.astype(int).tolist() for m in z.mapnames]}


def main():
    import cases
    from oracle import bridge
    bridge.build()
    out = []
    z = cases.tree2(T=12, S=3, seed=31)
    out.append(("plain_2state", {"variant": "PLAIN", "Q": cases.Q2.tolist(), "pid": [0.5, 0.5], "Omega": 0.2, "N": 8,
                                 "seed": 11, "tree": _tree_json(z)}))
    out.append(("sparse_2state", {"variant": "SPARSE", "Q": cases.Q2.tolist(), "pid": [0.5, 0.5], "Omega": 0.2, "N": 8,
                                  "seed": 12, "tree": _tree_json(z)}))
    Q4 = cases.q4()
    z4 = cases.tree_n(Q4, T=10, S=2, seed=32, mean_branch=0.6, segments=3)
    out.append(("bigtree_4state", {"variant": "BIGTREE", "Q": Q4.tolist(), "pid": [0.25] * 4, "Omega": 2.4, "N": 6,
                                   "seed": 13, "tree": _tree_json(z4)}))
    out.append(("bf_2state", {"variant": "BF", "Q": cases.Q2.tolist(), "pid": [0.5, 0.5], "Omega": 0.5, "N": 10,
                              "seed": 14, "prior": cases.PRIOR_BF.tolist(), "tree": _tree_json(z)}))
    zk = cases.tree_hidden(Q4, T=10, S=2, seed=33, mean_branch=0.5)
    out.append(("ks_4state", {"variant": "KS", "Q": Q4.tolist(), "pid": [0.25] * 4, "Omega": 4.0, "N": 8, "seed": 15,
                              "prior": cases.PRIOR_KS.tolist(), "tree": _tree_json(zk)}))
    rng = np.random.default_rng(9)
    scales = [rng.uniform(0.7, 1.3, size=z.E).tolist()]
    out.append(("mt_2state", {"variant": "MT", "Q": cases.Q2.tolist(), "pid": [0.5, 0.5], "Omega": 0.5, "N": 8, "seed": 16,
                              "prior": cases.PRIOR_BF.tolist(), "tree": _tree_json(z), "extra_tree_scales": scales}))
    scales = [rng.uniform(0.7, 1.3, size=zk.E).tolist()]
    out.append(("ksmt_4state", {"variant": "KSMT", "Q": Q4.tolist(), "pid": [0.25] * 4, "Omega": 4.0, "N": 6, "seed": 17,
                                "prior": cases.PRIOR_KSMT.tolist(), "tree": _tree_json(zk), "extra_tree_scales": scales}))
    out.append(("dic2s_2state", {"variant": "DIC2S", "Q": cases.Q2.tolist(), "pid": [0.5, 0.5], "Omega": 0.5, "N": 8, "seed": 18,
                                 "prior": cases.PRIOR_BF.tolist(), "tree": _tree_json(z)}))
    out.append(("dicks_4state", {"variant": "DICKS", "Q": Q4.tolist(), "pid": [0.25] * 4, "Omega": 4.0, "N": 6, "seed": 19,
                                 "prior": cases.PRIOR_KS.tolist(), "tree": _tree_json(zk)}))
    for name, case in out:
        rows = run_case(bridge, case)
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump({"case": case, "rows": rows.tolist()}, f)
        print(name, rows.shape)


if __name__ == "__main__":
    main()

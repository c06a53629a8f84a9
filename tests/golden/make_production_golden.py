"""Freezes rows of the PRODUCTION arithmetic (Philox keying, fast kernels) so that a refactor that is meant to keep the
draws cannot change them unnoticed.  These are not parity vectors (the production mode is checked against the oracle
statistically): regenerate on a B200 after a deliberate change of the random-number mapping.

    python tests/golden/make_production_golden.py        # writes tests/golden/production/rows.json (needs the GPU)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402
import phylomap_b200 as pb  # noqa: E402


def rows():
    out = {}
    Q4, pid4 = cases.q4(), np.full(4, 0.25)
    z = cases.tree_n(Q4, T=50, S=33, seed=3, mean_branch=0.8, segments=3)
    out["bigtree_f64"] = pb.sumstatMCMC_bigtree(z, Q4, pid4, 2.4, 6, precision="f64", seed=21)
    out["plain_f32"] = pb.sumstatMCMC(z, Q4, pid4, 2.4, 6, precision="f32", seed=22)
    zk = cases.tree_hidden(Q4, T=30, S=5, seed=4, mean_branch=0.5)
    out["ks_f64"] = pb.sumstatMCMCks(zk, np.asfortranarray(Q4.copy()), pid4, 4.0, 6, cases.PRIOR_KS, precision="f64", seed=23)
    z2 = cases.tree2(T=40, S=17, seed=5, mean_branch=4.0)
    out["bf_f64"] = pb.sumstatMCMCbf(z2, np.asfortranarray(cases.Q2.copy()), cases.PID2, 0.5, 6, cases.PRIOR_BF, precision="f64", seed=24)
    out["exp_f64"] = pb.sumstatEXP(z2, cases.Q2, cases.PID2, 4, precision="f64", seed=25)
    return out


if __name__ == "__main__":
    dst = os.path.join(os.environ.get("PM_GOLDEN_OUT", os.path.join(HERE, "production")), "rows.json")
    json.dump({k: v.tolist() for k, v in rows().items()}, open(dst, "w"))
    print("wrote", dst)

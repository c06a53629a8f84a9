"""Seeded inputs shared by the oracle tests (CPU) and the parity tests (GPU)."""
import numpy as np

import phylomap_b200 as pb
from phylomap_b200 import synth

Q2 = np.array([[-0.1, 0.1], [0.1, -0.1]])
PID2 = np.array([0.5, 0.5])
PRIOR_BF = np.array([0.55, 1.0, 0.56, 1.01])          # phylomap_tutorial.Rnw:218
PRIOR_KS = np.array([1.0, 10.0, 2.0, 10.0, 20.0, 2.0])  # phylomap_tutorial.Rnw:261
PRIOR_KSMT = np.array([1.0, 10.0, 1.1, 11.0, 2.0, 10.0, 20.0, 2.0])  # phylomap_tutorial.Rnw:303


def q4():
    return synth.make2sQ(0.1, 0.1, 0.2, 0.2, 10.0)  # phylomap_tutorial.Rnw:248


def q6():
    return synth.make2sQ(0.1, 0.15, [0.2, 0.1], [0.2, 0.3], [5.0, 0.5])


def jc(n, r=0.1):
    Q = np.full((n, n), r)
    np.fill_diagonal(Q, 0)
    np.fill_diagonal(Q, -Q.sum(1))
    return Q


def tree2(T=20, S=1, seed=1, mean_branch=5.0):
    t = synth.yule_tree(T, seed, mean_branch=mean_branch)
    return synth.simulate_2_state_tree(101 + seed, t, Q2, PID2, n_sites=S)


def tree_n(Q, T=20, S=1, seed=1, mean_branch=1.0, segments=None):
    t = synth.yule_tree(T, seed, mean_branch=mean_branch)
    n = Q.shape[0]
    st = synth.simulate_tip_states(t, Q, np.full(n, 1.0 / n), S, 300 + seed).numpy()
    return t.with_states(st[0].astype(np.int32) if S == 1 else st, segments=segments)


def tree_hidden(Q, T=20, S=1, seed=1, mean_branch=1.0, segments=3):
    t = synth.yule_tree(T, seed, mean_branch=mean_branch)
    n = Q.shape[0]
    return synth.simulate_4_state_tree(500 + seed, t, Q, np.full(n, 1.0 / n), n_sites=S, segments=segments)


def squamate_tree():
    """The reference's Squamate tree (3 951 tips, 100 segments per branch) from the derived fixture
    tests/golden/squamate_tree.npz (tests/golden/make_squamate_fixture.py).  Like the file it comes from it is a template:
    its `states` hold the placeholder -10 for most tips (R/Squamate_tree_setup.R:60) and the vignette simulates tip data
    onto it (`simulate_2_state_tree(seed = 101, atree, Q2, pid2)`, Squamate_DIC_model_selection.Rnw:93)."""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "squamate_tree.npz"))
    off, ml, ms = d["maps_off"], d["maps_len"], d["maps_state"].astype(np.int32)   # (each access decompresses: once)
    maps = [ml[off[e]:off[e + 1]] for e in range(len(off) - 1)]
    names = [ms[off[e]:off[e + 1]] for e in range(len(off) - 1)]
    return pb.PhyloTree(d["edge"], d["edge_length"], d["states"].astype(np.int32), maps, names)

"""CPU tests of the oracle (the restatement of src/phylomap.cpp): the reference ships no tests, golden vectors or
known-answer fixtures for this path ("parity unpinned"), so the oracle is pinned by
(i) invariants of the algorithm, (ii) closed-form 2-state expectations, (iii) agreement between its two
independent samplers (the uniformization MCMC and the matrix-exponential direct sampler, maketreelistEXP),
and (iv) the committed golden rows in tests/golden/ which freeze its output against regressions.
"""
import os

import numpy as np
import pytest

import cases
from phylomap_b200 import PhyloTree, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(oracle, variant, z, Q, pid, Om, N, **kw):
    trees = z if isinstance(z, list) else [z]
    o = oracle.OracleRun(variant, [t.oracle_dict() for t in trees], Q, pid, Om, N, **kw)
    return o, o.run()


def test_total_dwell_time_is_tree_length(oracle):
    for S in (1, 6):
        z = cases.tree2(T=30, S=S, seed=4)
        _, out = _run(oracle, oracle.PLAIN, z, cases.Q2, cases.PID2, 0.2, 30, seed=3)
        np.testing.assert_allclose(out[:, :2].sum(1), S * z.edge_length.sum(), rtol=1e-12)
        assert np.all(out[:, 2:] >= 0) and np.array_equal(out[:, 2:], np.round(out[:, 2:]))


def test_sparse_equals_dense_when_nothing_is_dropped(oracle):
    z = cases.tree2(T=25, S=3, seed=5)
    _, a = _run(oracle, oracle.PLAIN, z, cases.Q2, cases.PID2, 0.2, 20, seed=11)
    _, b = _run(oracle, oracle.SPARSE, z, cases.Q2, cases.PID2, 0.2, 20, seed=11)
    assert np.array_equal(a, b)


def test_bigtree_is_plain_with_rescaled_partials(oracle):
    """Per-node renormalisation (src/phylomap.cpp:525) only rescales the weights of each categorical draw."""
    z = cases.tree2(T=40, S=2, seed=6)
    oa, a = _run(oracle, oracle.PLAIN, z, cases.Q2, cases.PID2, 0.2, 15, seed=5)
    ob, b = _run(oracle, oracle.BIGTREE, z, cases.Q2, cases.PID2, 0.2, 15, seed=5)
    assert np.array_equal(a[:, 2:], b[:, 2:])
    np.testing.assert_allclose(a[:, :2], b[:, :2], rtol=1e-12)
    pa, pb_ = oa.partials(0), ob.partials(0)
    T = z.T
    scale = pa[T:].sum(1, keepdims=True)   # plain partials are the normalised ones times a per-node factor
    rescaled = pa[T:] / np.where(scale > 0, scale, 1.0)
    rescaled /= rescaled.sum(1, keepdims=True)
    np.testing.assert_allclose(rescaled, pb_[T:], rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(pb_[T:].sum(1), 1.0, rtol=1e-12)


def test_replay_table_reproduces_the_sequential_run(oracle):
    z = cases.tree2(T=14, S=2, seed=8)
    o, a = _run(oracle, oracle.BF, z, cases.Q2.copy(), cases.PID2, 0.5, 12, prior=cases.PRIOR_BF,
                rng_mode=oracle.SEQUENTIAL, seed=42, want_log=True)
    table, host = o.export_log()
    o2, b = _run(oracle, oracle.BF, z, cases.Q2.copy(), cases.PID2, 0.5, 12, prior=cases.PRIOR_BF,
                 rng_mode=oracle.TABLE, seed=42, table=table, host_table=host)
    assert np.array_equal(a, b)
    assert np.array_equal(o.node_states(), o2.node_states())


def test_single_branch_pair_closed_form(oracle):
    """Cherry (2 tips), symmetric 2-state chain with rate a: E[N] and E[R] follow from the endpoint-conditioned
    closed forms.  For tips (1, 2) and a root drawn from (.5, .5) the two tip branches carry an odd number of jumps
    in total; with a*t small the expected number of jumps is close to 1 and both states share the time."""
    a, t = 0.1, 1.0
    Q = np.array([[-a, a], [a, -a]])
    tree = PhyloTree(np.array([[3, 1], [3, 2]]), np.array([t, t])).with_states(np.array([1, 2]))
    N = 20000
    _, out = _run(oracle, oracle.PLAIN, tree, Q, cases.PID2, 2 * a, N, seed=9)
    out = out[200:]
    # path of length 2t from tip 1 (state 0) to tip 2 (state 1) through the root: a 2-state chain conditioned on
    # its end points; root prior is uniform = stationary, so the pair is a stationary bridge of length 2t.
    L = 2 * t
    p01 = 0.5 * (1 - np.exp(-2 * a * L))
    # E[number of jumps | X0=0, XL=1]: sum over odd k of k * Pois(k; aL) / p01
    lam = a * L
    ks = np.arange(1, 60, 2)
    from scipy.stats import poisson
    en = (ks * poisson.pmf(ks, lam)).sum() / p01
    njumps = out[:, 2] + out[:, 3]
    assert abs(njumps.mean() - en) < 4 * njumps.std() / np.sqrt(len(out) / 5)
    np.testing.assert_allclose(out[:, 0].mean(), L / 2, rtol=0.03)   # symmetry: equal expected time in both states
    # jumps along the path from tip 1 to tip 2 start in state 0 and end in state 1; reading the tree from the root,
    # 0->1 and 1->0 counts differ by at most one
    assert np.all(np.abs(out[:, 2] - out[:, 3]) <= 1)


@pytest.mark.parametrize("n", [2, 4])
def test_mcmc_matches_direct_sampler(oracle, n):
    """sumstatMCMC vs sumstatEXP (the reference's own cross-check, phylomap_tutorial.Rnw:119-135): expected
    transition counts and dwell times of the two independent samplers agree within 1 %+MC error."""
    if n == 2:
        Q, pid = cases.Q2, cases.PID2
        z = cases.tree2(T=30, S=8, seed=3, mean_branch=4.0)
    else:
        Q, pid = cases.jc(4, 0.1), np.full(4, 0.25)
        z = cases.tree_n(Q, T=30, S=8, seed=3, mean_branch=2.0)
    Om = 2 * np.max(-np.diag(Q))
    N = 3000
    _, mc = _run(oracle, oracle.PLAIN, z, Q, pid, Om, N, seed=21)
    w, V = np.linalg.eig(Q)
    eig = (V.real, np.linalg.inv(V).real, np.diag(w.real))
    _, ex = _run(oracle, oracle.EXP, z, Q, pid, Om, N, rng_mode=oracle.SEQUENTIAL, seed=22, eig=eig)
    mc = mc[300:]
    tot_mc, tot_ex = mc[:, n:].sum(1), ex[:, n:].sum(1)
    se = np.sqrt(tot_mc.var() * 8 / len(mc) + tot_ex.var() / len(ex))
    assert abs(tot_mc.mean() - tot_ex.mean()) < 4 * se + 0.01 * tot_ex.mean()
    np.testing.assert_allclose(mc[:, :n].mean(0), ex[:, :n].mean(0), rtol=0.03)


def test_rate_chains_run_and_record_layout(oracle):
    z = cases.tree2(T=20, S=2, seed=7)
    Q = cases.Q2.copy()
    o, bf = _run(oracle, oracle.BF, z, Q, cases.PID2, 0.5, 50, prior=cases.PRIOR_BF, seed=5)
    assert bf.shape == (50, 9)
    assert bf[0, 6] == 0.1 and bf[0, 7] == 0.1              # recordQ runs before the sweep (:1296)
    assert np.all(bf[:, 6:8] <= 0.5) and set(np.unique(bf[:, 8])) <= {0.0, 1.0}
    assert not np.allclose(o.Q, cases.Q2)                   # Q updated in place
    Q4 = cases.q4()
    zk = cases.tree_hidden(Q4, T=20, S=2, seed=3, mean_branch=0.5)
    o, ks = _run(oracle, oracle.KS, zk, Q4.copy(), np.full(4, .25), 4.0, 30, prior=cases.PRIOR_KS, seed=5)
    assert ks.shape == (30, 4 + 16 + 2 + 3 + 1)
    np.testing.assert_allclose(ks[0, 20:25], [0.1, 0.1, 0.2, 0.2, 10.0])
    np.testing.assert_allclose(o.Q.sum(1), 0, atol=1e-12)   # rows of Q keep summing to zero


def test_golden_rows(oracle):
    """Frozen oracle output (tests/golden/make_golden.py wrote these with this same oracle; they guard against
    accidental changes of the restated algorithm, they do not pin it to a real R run)."""
    import json
    import sys
    sys.path.insert(0, GOLD)
    import make_golden
    names = [f for f in sorted(os.listdir(GOLD)) if f.endswith(".json")]
    assert len(names) >= 9
    for name in names:
        g = json.load(open(os.path.join(GOLD, name)))
        out = make_golden.run_case(oracle, g["case"])
        np.testing.assert_allclose(out, np.array(g["rows"]), rtol=1e-13, atol=0, err_msg=name)


def test_dic_loglik_against_scipy(oracle):
    """The DIC column (log p(y|Q) by matrix exponentiation, src/phylomap.cpp:3242-3250) against an independent
    evaluation with scipy.linalg.expm; and it must not disturb the bf chain it rides on."""
    from scipy.linalg import expm
    z = cases.tree2(T=14, S=2, seed=5)
    o, dic = _run(oracle, oracle.DIC2S, z, cases.Q2.copy(), cases.PID2, 0.5, 6, prior=cases.PRIOR_BF, seed=4)
    _, bf = _run(oracle, oracle.BF, z, cases.Q2.copy(), cases.PID2, 0.5, 6, prior=cases.PRIOR_BF, seed=4)
    assert np.array_equal(dic[:, :9], bf)
    nen, _, root = z.order()
    for i in (0, 5):
        Q = np.array([[-dic[i, 6], dic[i, 6]], [dic[i, 7], -dic[i, 7]]])
        tot = 0.0
        for s in range(2):
            PL = np.zeros((2 * z.T - 1, 2))
            PL[np.arange(z.T), z.states[s] - 1] = 1
            S = 0.0
            for k in range(z.T - 1):
                ea, eb = nen[2 * k] - 1, nen[2 * k + 1] - 1
                v = (expm(Q * z.edge_length[ea]) @ PL[z.edge[ea, 1] - 1]) * (expm(Q * z.edge_length[eb]) @ PL[z.edge[eb, 1] - 1])
                S += np.log(v.sum())
                PL[z.edge[ea, 0] - 1] = v / v.sum()
            tot += np.log(PL[root - 1] @ cases.PID2) + S
        np.testing.assert_allclose(dic[i, 9], tot, rtol=1e-12)


def test_squamate_tree_vignette_configuration(oracle):
    """The DIC vignette's 2-state setup on the reference's own tree (Squamate_DIC_model_selection.Rnw:78-104: Q2, pid2,
    prior2r, Omega = 10): a sweep adds ~Omega x tree length = 877 000 jump points; every row keeps the invariants."""
    Q2 = np.array([[-0.001, 0.001], [0.006, -0.006]])
    z = synth.simulate_2_state_tree(101, cases.squamate_tree(), Q2, cases.PID2, segments=100)
    N = 3
    run = oracle.OracleRun(oracle.DIC2S, [z.oracle_dict()], Q2, cases.PID2, 10.0, N, prior=np.array([0.55, 1, 0.55, 1.0]), seed=3)
    rows = run.run()
    assert rows.shape == (N, 10)
    np.testing.assert_allclose(rows[:, :2].sum(1), z.edge_length.sum(), rtol=1e-9)
    jumps = rows[:, 2:6].sum(1)                       # all jump points, virtual ones included
    assert np.all(np.abs(jumps[1:] / (10.0 * z.edge_length.sum()) - 1) < 0.01)
    assert np.all(rows[:, 6:8] > 0) and np.all(rows[:, 9] < 0) and np.all(np.isfinite(rows[:, 9]))

"""BASELINE.json's configurations at their STATED sizes, as driver-run GPU tests.

configs[1]  simulate_4_state_tree 1k-tip tree, 10k-site nucleotide alignment, sumstatMCMC vs sumstatEXP expected counts
configs[2]  SPARSEsumstatMCMC on the Squamate tree with 100k synthetic sites
configs[4]  sumstatMCMCks / ksmt / mt at one GPU's share of the 10k-tip, 1M-site run (125 000 sites)
(configs[0] is tests/test_gpu_parity.py::test_config0_hundred_tips_thousand_sweeps, configs[3] is bench.py.)

Bar (north_star): expected counts and dwell times of the MCMC within 1 % of the matrix-exponentiation sampler; the
direct sampler itself is checked against the oracle's restatement of maketreelistEXP on a 64-site subsample.
"""
import numpy as np
import pytest

import cases
import phylomap_b200 as pb
from phylomap_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _within_one_percent(mc, ex, n, S, what):
    """mc, ex: per-sweep rows (sums over S sites).  Means of every dwell time, of the total number of changes and of the
    individual N_ij that carry at least 2 % of the changes: within 1 % (+ 4 standard errors of the difference)."""
    def chk(a, b, name):
        se = np.sqrt(a.var() / max(len(a) / 4, 1) + b.var() / len(b))   # MCMC rows are autocorrelated: effective n / 4
        assert abs(a.mean() - b.mean()) <= 0.01 * abs(b.mean()) + 4 * se, "%s %s: mcmc %.6g vs exp %.6g (se %.3g)" % (what, name, a.mean(), b.mean(), se)
    for s in range(n):
        chk(mc[:, s], ex[:, s], "R_%d" % s)
    tm, te = mc[:, n:].sum(1), ex[:, n:].sum(1)
    chk(tm, te, "changes")
    for c in range(n, mc.shape[1]):
        if ex[:, c].mean() > 0.02 * te.mean():
            chk(mc[:, c], ex[:, c], "N col %d" % c)


def test_config1_thousand_tips_ten_thousand_sites(oracle):
    Q, pid = cases.jc(4, 0.1), np.full(4, 0.25)
    T, S = 1000, 10000
    tree = synth.yule_tree(T, seed=2, mean_branch=1.0)
    st = synth.simulate_tip_states(tree, Q, pid, S, seed=202, device="cuda").cpu().numpy()
    z = tree.with_states(st, segments=2)
    mc = pb.sumstatMCMC(z, Q, pid, 0.6, 70, seed=4, precision="f32")[30:]
    ex = pb.sumstatEXP(z, Q, pid, 10, seed=3, precision="f32")
    np.testing.assert_allclose(mc[:, :4].sum(1), S * tree.edge_length.sum(), rtol=2e-4)
    _within_one_percent(mc, ex, 4, S, "cfg1 mcmc/exp")
    # The oracle on a 16-site subsample.  The reference's own direct sampler cannot run at this size: makePLold / makePLexp
    # (src/phylomap.cpp:2877-2907) never rescale, so with 1 000 tips every partial underflows FP64 and
    # RcppArmadillo::sample throws -- the oracle restatement reproduces that, hence the comparison is with the oracle's
    # maketreelistMCMC_bigtree chain (:942, rescaled partials) instead.
    sub = tree.with_states(st[:16], segments=2)
    w, V = np.linalg.eig(Q)
    eig = (V.real, np.linalg.inv(V).real, np.diag(w.real))
    with pytest.raises(oracle.OracleError, match="Not enough positive probabilities"):
        oracle.OracleRun(oracle.EXP, [sub.oracle_dict()], Q, pid, 0.6, 1, rng_mode=oracle.SEQUENTIAL, seed=22, eig=eig).run()
    ref = oracle.OracleRun(oracle.BIGTREE, [sub.oracle_dict()], Q, pid, 0.6, 120, rng_mode=oracle.KEYED, seed=22).run()[30:]
    got = pb.sumstatMCMC_bigtree(sub, Q, pid, 0.6, 240, seed=5, precision="f32")[30:]
    ex16 = pb.sumstatEXP(sub, Q, pid, 60, seed=6, precision="f32")
    for name, f in [("changes", lambda a: a[:, 4:].sum(1)), ("R_0", lambda a: a[:, 0]), ("R_3", lambda a: a[:, 3])]:
        b = f(ref)
        for what, a in (("gpu mcmc", f(got)), ("gpu exp", f(ex16))):
            se = np.sqrt(a.var() / (len(a) / 4) + b.var() / (len(b) / 4))
            assert abs(a.mean() - b.mean()) <= 0.01 * abs(b.mean()) + 4 * se, (what, name, a.mean(), b.mean(), se)


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_config2_squamate_hundred_thousand_sites(precision):
    """The Squamate tree itself (3 951 tips, tree length 87 740; fixture derived from the package's .RData), the
    vignette's Q (Squamate_DIC_model_selection.Rnw:83), SPARSE sampler, 100 000 synthetic sites.
    FP32 and FP64 production arithmetic.  With rates this slow on a tree this large a few sites in 1e5 have sister clades
    that each all but settle a DIFFERENT state (1e-20 x 1e-20): the FP32 product underflows although both factors are
    representable; the pruning kernel redoes such a node in double (round 1 reported PM_ERR_SAMPLE for them)."""
    Q = np.array([[-0.001, 0.001], [0.006, -0.006]])
    tree = cases.squamate_tree()
    S = 100000
    z = synth.simulate_2_state_tree(5, tree, Q, cases.PID2, n_sites=S, device="cuda", segments=8)
    # The initial maps of simulate_2_state_tree (every internal branch in state 1) are far from equilibrium, and a branch
    # without a jump point pins its two end states together, so the chain forgets them at a speed set by Omega x t: with
    # Omega = 0.012 (0.13 jump points per branch) the gap to the direct sampler still shrinks by a third per 120 sweeps
    # after 400 sweeps.  Omega = 0.06 (0.66 per branch, ~5 300 jump points per site and sweep) mixes five times faster.
    Om = 0.06
    mc = pb.SPARSEsumstatMCMC(z, Q, cases.PID2, Om, 300, precision=precision, seed=3)[230:]
    np.testing.assert_allclose(mc[:, :2].sum(1), S * tree.edge_length.sum(), rtol=2e-4)
    assert np.all(mc[:, 2:] >= 0) and np.array_equal(mc[:, 2:], np.round(mc[:, 2:]))
    ex = pb.sumstatEXP(z, Q, cases.PID2, 8, seed=9, precision="f64")
    _within_one_percent(mc, ex, 2, S, "cfg2 sparse/exp")


def test_config4_rate_samplers_full_size():
    """ks, ksDICt, ksmt (2 trees x half the sites) and bf on the 10 000-tip tree with 125 000 sites, FP32: every row's
    invariants, rates stay finite and positive, the model indicator is a tree index, acceptance counters are kept."""
    S = 125000
    Q4, pid4 = cases.q4(), np.full(4, 0.25)
    tree = synth.yule_tree(10000, seed=4, mean_branch=0.1 / 1.2)
    zk = synth.simulate_4_state_tree(7, tree, Q4, pid4, n_sites=S, device="cuda", segments=2)
    total = S * tree.edge_length.sum()
    N = 5
    ch = pb.Chain(capi.PM_V_KS, zk, np.asfortranarray(Q4.copy()), pid4, 4.0, N, prior=cases.PRIOR_KS, seed=3, precision="f32")
    ks = ch.run()
    np.testing.assert_allclose(ks[:, :4].sum(1), total, rtol=2e-4)
    assert np.all(ks[:, 4:20] >= 0) and np.array_equal(ks[:, 4:20], np.round(ks[:, 4:20]))
    assert np.all(ks[:, 20:25] > 0) and np.all(np.isfinite(ks))
    prop, acc = ch.acceptance()
    assert len(prop) == 5 and np.all(prop == N)
    del ch
    dic = pb.sumstatMCMCksDICt(zk, np.asfortranarray(Q4.copy()), pid4, 4.0, 3, cases.PRIOR_KS, precision="f32", seed=3)
    assert np.all(np.isfinite(dic[:, -1])) and np.all(dic[:, -1] < 0)
    # the log-likelihood of 125 000 sites under the simulating model: about -S x (a few nats per site), same scale every sweep
    assert np.ptp(dic[:, -1]) < 0.02 * abs(dic[:, -1].mean())
    half = S // 2
    trees = [pb.PhyloTree(zk.edge, zk.edge_length * f).with_states(zk.states[:half], segments=2) for f in (1.0, 1.1)]
    ksmt = pb.sumstatMCMCksmt(trees, np.asfortranarray(Q4.copy()), pid4, 4.0, N, cases.PRIOR_KSMT, precision="f32", seed=3)
    assert set(np.unique(ksmt[:, -1])) <= {0.0, 1.0}
    for i in range(N):
        f = 1.0 if ksmt[i, -1] == 0 else 1.1
        np.testing.assert_allclose(ksmt[i, :4].sum(), half * tree.edge_length.sum() * f, rtol=2e-4)
    del trees, zk
    Q2 = np.array([[-0.1, 0.1], [0.1, -0.1]])
    z2 = synth.simulate_2_state_tree(9, tree, Q2, cases.PID2, n_sites=S, device="cuda", segments=2)
    mt_trees = [pb.PhyloTree(z2.edge, z2.edge_length * f).with_states(z2.states[:half], segments=2) for f in (1.0, 0.9)]
    mt = pb.sumstatMCMCmt(mt_trees, np.asfortranarray(Q2.copy()), cases.PID2, 0.5, N, cases.PRIOR_BF, precision="f32", seed=3)
    assert set(np.unique(mt[:, -1])) <= {0.0, 1.0} and np.all(mt[:, 6:8] > 0)
    capi.lib().pm_release_cached_memory()

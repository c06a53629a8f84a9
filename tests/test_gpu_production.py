"""GPU tests of the production arithmetic (FP32 / FP64 fast path, Philox streams): tier 2 of BASELINE.json.

Posterior distributions of N_ij and R_i must agree with the oracle's chains (two-sample KS, p > 0.01), expected
counts with the matrix-exponential sampler sumstatEXP within 1 %, and — at the full tree size of the benchmark —
the size-independent invariants of a sweep must hold.
"""
import numpy as np
import pytest
from scipy import stats

import cases
import phylomap_b200 as pb
from phylomap_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_ks_against_oracle_chain(oracle, precision):
    """One site, long chains, thinned: the GPU chain (production mode) and the oracle chain (R-order Mersenne-Twister
    stream) sample the same posterior of (N01, N10, R0)."""
    z = cases.tree2(T=20, S=1, seed=3, mean_branch=6.0)
    N, thin, burn = 12000, 12, 600
    ref = oracle.OracleRun(oracle.PLAIN, [z.oracle_dict()], cases.Q2, cases.PID2, 0.2, N, rng_mode=oracle.SEQUENTIAL,
                           seed=123).run()[burn::thin]
    got = pb.sumstatMCMC(z, cases.Q2, cases.PID2, 0.2, N, seed=99, precision=precision)[burn::thin]
    np.testing.assert_allclose(got[:, :2].sum(1), z.edge_length.sum(), rtol=1e-5)
    for col, name in [(2, "N01"), (3, "N10"), (0, "R0")]:
        p = stats.ks_2samp(got[:, col], ref[:, col]).pvalue
        assert p > 0.01, "%s: KS p = %.4f" % (name, p)


def test_ks_rate_traces_bf(oracle):
    """Rate-updating sampler: posterior of (lambda01, lambda10) from the GPU chain vs the oracle chain."""
    z = cases.tree2(T=40, S=1, seed=8, mean_branch=4.0)
    N, thin, burn = 24000, 40, 1000
    ref = oracle.OracleRun(oracle.BF, [z.oracle_dict()], cases.Q2.copy(), cases.PID2, 1.0, N, prior=cases.PRIOR_BF,
                           rng_mode=oracle.SEQUENTIAL, seed=5).run()[burn::thin]
    got = pb.sumstatMCMCbf(z, cases.Q2.copy(), cases.PID2, 1.0, N, cases.PRIOR_BF, seed=77, precision="f32")[burn::thin]
    for col, name in [(6, "l01"), (7, "l10"), (3, "n01")]:
        p = stats.ks_2samp(got[:, col], ref[:, col]).pvalue
        assert p > 0.01, "%s: KS p = %.4f" % (name, p)


def test_expected_counts_match_sumstatEXP(oracle):
    """configs[1] of BASELINE.json in miniature: 4-state nucleotide-like model, many sites; mean transition counts
    and dwell times per site from the GPU MCMC within 1 % of the direct sampler (plus its Monte-Carlo error)."""
    Q, pid = cases.jc(4, 0.1), np.full(4, 0.25)
    S = 1500
    z = cases.tree_n(Q, T=100, S=S, seed=2, mean_branch=1.5)
    Om = 2 * 0.3
    N = 60
    got = pb.sumstatMCMC(z, Q, pid, Om, N, seed=4, precision="f32")[20:] / S
    w, V = np.linalg.eig(Q)
    eig = (V.real, np.linalg.inv(V).real, np.diag(w.real))
    ex = oracle.OracleRun(oracle.EXP, [z.oracle_dict()], Q, pid, Om, 6, rng_mode=oracle.SEQUENTIAL, seed=22, eig=eig).run() / S
    tot_g, tot_e = got[:, 4:].sum(1), ex[:, 4:].sum(1)
    se = np.sqrt(tot_g.var() / 4 + tot_e.var() / len(ex))
    assert abs(tot_g.mean() - tot_e.mean()) < 0.01 * tot_e.mean() + 4 * se
    np.testing.assert_allclose(got[:, :4].mean(0), ex[:, :4].mean(0), rtol=0.02)
    np.testing.assert_allclose(got[:, 4:].mean(0), ex[:, 4:].mean(0), rtol=0.06)   # 12 individual N_ij, noisier


def test_f32_and_f64_production_agree_statistically():
    Q, pid = cases.q4(), np.full(4, 0.25)
    z = cases.tree_n(Q, T=200, S=512, seed=9, mean_branch=0.3, segments=2)
    a = pb.sumstatMCMC_bigtree(z, Q, pid, 2.4, 40, seed=1, precision="f32")[15:]
    b = pb.sumstatMCMC_bigtree(z, Q, pid, 2.4, 40, seed=2, precision="f64")[15:]
    np.testing.assert_allclose(a[:, :4].mean(0), b[:, :4].mean(0), rtol=0.02)
    np.testing.assert_allclose(a[:, 4:].sum(1).mean(), b[:, 4:].sum(1).mean(), rtol=0.02)


def test_full_tree_size_invariants():
    """The benchmark's tree (10 000 tips, 4 states, sumstatMCMC_bigtree) at a reduced site count: properties that do
    not need the oracle.  (i) total dwell time = sites x tree length; (ii) counts are non-negative integers;
    (iii) same seed -> identical rows; (iv) two site shards keyed by global site index sum to the unsharded run."""
    Q, pid = cases.q4(), np.full(4, 0.25)
    tree = synth.yule_tree(10000, seed=4, mean_branch=0.1 / 1.2)
    S = 4096
    st = synth.simulate_tip_states(tree, Q, pid, S, seed=7, device="cuda").cpu().numpy()
    z = tree.with_states(st, segments=2)
    kw = dict(precision="f32", seed=11)
    N = 4
    a = pb.sumstatMCMC_bigtree(z, Q, pid, 2.4, N, **kw)
    np.testing.assert_allclose(a[:, :4].sum(1), S * tree.edge_length.sum(), rtol=2e-4)
    assert np.all(a[:, 4:] >= 0) and np.array_equal(a[:, 4:], np.round(a[:, 4:]))
    assert a[:, 4:].sum() > 0
    b = pb.sumstatMCMC_bigtree(z, Q, pid, 2.4, N, **kw)
    assert np.array_equal(a, b)
    h = S // 2
    lo = pb.sumstatMCMC_bigtree(tree.with_states(st[:h], segments=2), Q, pid, 2.4, N, site_offset=0, **kw)
    hi = pb.sumstatMCMC_bigtree(tree.with_states(st[h:], segments=2), Q, pid, 2.4, N, site_offset=h, **kw)
    assert np.array_equal(lo[:, 4:] + hi[:, 4:], a[:, 4:])
    np.testing.assert_allclose(lo[:, :4] + hi[:, :4], a[:, :4], rtol=1e-5)


def test_squamate_tree_sparse_run():
    """configs[2]: the Squamate tree itself (3 951 tips, tree length 87 740; fixture derived from the package's .RData),
    2-state SPARSE sampler with the vignette's Q (Squamate_DIC_model_selection.Rnw:83) and Omega = 0.012 (so that
    Omega x tree length ~ 1 000 jump points per site), many synthetic sites."""
    Q = np.array([[-0.001, 0.001], [0.006, -0.006]])
    tree = cases.squamate_tree()
    S = 2048
    # initial maps: every branch in 8 equal pieces (the Squamate set-up script uses 100, R/Squamate_tree_setup.R:54-82);
    # with the one-piece internal branches of simulate_2_state_tree a 3 951-tip tree underflows even FP64 partials
    z = synth.simulate_2_state_tree(5, tree, Q, cases.PID2, n_sites=S, device="cuda", segments=8)
    out = pb.SPARSEsumstatMCMC(z, Q, cases.PID2, 0.012, 8, precision="f32", seed=3)
    np.testing.assert_allclose(out[:, :2].sum(1), S * tree.edge_length.sum(), rtol=2e-4)
    assert np.all(out[:, 2:] >= 0)
    assert out[4:, 2:].sum() > 0


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_long_branches_gap_mode(oracle, precision):
    """Runs with (Omega + Q_ss) L > 16 draw their virtual jumps from exponential gaps (pm_device.cuh); with
    Omega * t ~ 30 most branches do, every branch goes through the general path kernel, and the power table
    is longer than its shared-memory part.  Posterior of the jump counts against the oracle chain."""
    z = cases.tree2(T=12, S=1, seed=6, mean_branch=75.0)   # ~8 real changes per branch: well below the 63-change limit
    N, thin, burn = 6000, 10, 300
    ref = oracle.OracleRun(oracle.PLAIN, [z.oracle_dict()], cases.Q2, cases.PID2, 0.4, N, rng_mode=oracle.SEQUENTIAL,
                           seed=11).run()[burn::thin]
    got = pb.sumstatMCMC(z, cases.Q2, cases.PID2, 0.4, N, seed=5, precision=precision)[burn::thin]
    np.testing.assert_allclose(got[:, :2].sum(1), z.edge_length.sum(), rtol=1e-5)
    for col, name in [(2, "N01"), (3, "N10"), (0, "R0")]:
        p = stats.ks_2samp(got[:, col], ref[:, col]).pvalue
        assert p > 0.01, "%s: KS p = %.4f" % (name, p)


def test_ragged_site_counts_and_chunk_edges():
    """Site counts that are not multiples of the 32-site tiles / 128-site blocks, on a tree whose branch count is not a
    multiple of the chunk size: the invariants must hold for every S."""
    Q, pid = cases.q4(), np.full(4, 0.25)
    tree = synth.yule_tree(333, seed=8, mean_branch=0.2)
    for S in (1, 31, 33, 127, 129, 1000):
        st = synth.simulate_tip_states(tree, Q, pid, S, seed=S).numpy()
        z = tree.with_states(st if S > 1 else st[0].astype(np.int32), segments=2)
        out = pb.sumstatMCMC_bigtree(z, Q, pid, 2.4, 5, seed=3, precision="f32")
        np.testing.assert_allclose(out[:, :4].sum(1), S * tree.edge_length.sum(), rtol=2e-4)
        assert np.array_equal(out[:, 4:], np.round(out[:, 4:])) and np.all(out[:, 4:] >= 0)


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_direct_sampler_matches_oracle_exp(oracle, precision):
    """maketreelistEXP on the GPU (independent samples) against the oracle's restatement of it: same distribution of
    the per-iteration statistics (two-sample KS) and the same means."""
    Q, pid = cases.jc(4, 0.1), np.full(4, 0.25)
    z = cases.tree_n(Q, T=30, S=1, seed=3, mean_branch=2.0)
    N = 4000
    w, V = np.linalg.eig(Q)
    eig = (V.real, np.linalg.inv(V).real, np.diag(w.real))
    ref = oracle.OracleRun(oracle.EXP, [z.oracle_dict()], Q, pid, 0.3, N, rng_mode=oracle.SEQUENTIAL, seed=22, eig=eig).run()
    got = pb.sumstatEXP(z, Q, pid, N, seed=5, precision=precision)
    assert got.shape == ref.shape == (N, 16)
    np.testing.assert_allclose(got[:, :4].sum(1), z.edge_length.sum(), rtol=1e-5)
    assert np.array_equal(got[:, 4:], np.round(got[:, 4:]))
    tg, tr = got[:, 4:].sum(1), ref[:, 4:].sum(1)
    assert stats.ks_2samp(tg, tr).pvalue > 0.01
    assert stats.ks_2samp(got[:, 0], ref[:, 0]).pvalue > 0.01
    np.testing.assert_allclose(got[:, 4:].mean(0), ref[:, 4:].mean(0), rtol=0.12, atol=0.02)
    np.testing.assert_allclose(got[:, :4].mean(0), ref[:, :4].mean(0), rtol=0.02)


def test_direct_sampler_and_mcmc_agree_on_the_gpu():
    """The reference's own cross-check (phylomap_tutorial.Rnw:119-135) at scale, entirely on the GPU: sumstatEXP vs
    sumstatMCMC, many sites, expected counts within 1 %."""
    Q, pid = cases.jc(4, 0.1), np.full(4, 0.25)
    S = 2000
    z = cases.tree_n(Q, T=100, S=S, seed=2, mean_branch=1.5)
    ex = pb.sumstatEXP(z, Q, pid, 8, seed=3, precision="f32") / S
    mc = pb.sumstatMCMC(z, Q, pid, 0.6, 60, seed=4, precision="f32")[20:] / S
    te, tm = ex[:, 4:].sum(1), mc[:, 4:].sum(1)
    se = np.sqrt(te.var() / len(te) + tm.var() / 4)
    assert abs(te.mean() - tm.mean()) < 0.01 * te.mean() + 4 * se
    np.testing.assert_allclose(ex[:, :4].mean(0), mc[:, :4].mean(0), rtol=0.02)


def test_tutorial_twenty_state_model_three_samplers(oracle):
    """phylomap_tutorial.Rnw:67-135: 20-state tridiagonal Q, 50 tips, Omega = 0.2 -- sumstatEXP, sumstatMCMC and
    SPARSEsumstatMCMC must give the same distribution of the number of jumps (the tutorial overlays their histograms).
    Regression: the direct sampler's pruning kernel kept its partials in arrays of 8 states."""
    Q = np.zeros((20, 20))
    for j in range(19):
        Q[j, j + 1] = Q[j + 1, j] = 0.003
    np.fill_diagonal(Q, -Q.sum(1))
    pid = np.full(20, 0.05)
    phy = synth.yule_tree(50, seed=3, mean_branch=12.0)
    tips = synth.simulate_tip_states(phy, Q, np.eye(20)[0], 1, seed=11).numpy()[0].astype(np.int32)
    z = phy.with_states(tips)
    N = 3000
    ex = pb.sumstatEXP(z, Q, pid, N, seed=1)
    mc = pb.sumstatMCMC(z, Q, pid, 0.2, N, seed=2)[200::4]
    sp = pb.SPARSEsumstatMCMC(z, Q, pid, 0.2, N, seed=3)[200::4]
    w, V = np.linalg.eig(Q)
    ref = oracle.OracleRun(oracle.EXP, [z.oracle_dict()], Q, pid, 0.2, 1500, rng_mode=oracle.SEQUENTIAL, seed=22,
                           eig=(V.real, np.linalg.inv(V).real, np.diag(w.real))).run()
    je, jm, js, jr = (a[:, 20:].sum(1) for a in (ex, mc, sp, ref))
    np.testing.assert_allclose(ex[:, :20].sum(1), z.edge_length.sum(), rtol=1e-5)
    assert abs(je.mean() - jr.mean()) < 0.15 and abs(jm.mean() - jr.mean()) < 0.2 and abs(js.mean() - jr.mean()) < 0.2
    assert stats.ks_2samp(je, jr).pvalue > 0.005 and stats.ks_2samp(je, jm).pvalue > 0.002 and stats.ks_2samp(jm, js).pvalue > 0.002


def test_record_capacity_overflow_is_reported():
    """A path with two or more real jumps keeps its runs as records in a per-(site, chunk) slice; when the caller
    forces the slices too small the sweep must stop with PM_ERR_CAPACITY, not corrupt memory."""
    Q = np.array([[-1.0, 1.0], [1.0, -1.0]])
    z = cases.tree2(T=64, S=40, seed=3, mean_branch=6.0)     # ~6 real jumps per branch
    with pytest.raises(capi.PhylomapError) as e:
        pb.sumstatMCMC(z, Q, cases.PID2, 2.0, 10, seed=1, precision="f32", path_capacity=1)
    assert e.value.code == capi.PM_ERR_CAPACITY
    ok = pb.sumstatMCMC(z, Q, cases.PID2, 2.0, 10, seed=1, precision="f32")  # default sizing copes
    np.testing.assert_allclose(ok[:, :2].sum(1), 40 * z.edge_length.sum(), rtol=2e-4)
    assert ok[3:, 2:].mean() > 0.5 * 40 * z.edge_length.sum() * 0.9 / 2
    # the deterministic mode holds at most 64 runs per path and says so on a saturated branch
    zs = cases.tree2(T=6, S=8, seed=3, mean_branch=150.0)
    with pytest.raises(capi.PhylomapError) as e:
        pb.sumstatMCMC(zs, Q, cases.PID2, 2.0, 10, seed=1, mode="deterministic")
    assert e.value.code == capi.PM_ERR_CAPACITY and "63" in e.value.msg


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_saturated_branches_long_paths(oracle, precision):
    """Branches with hundreds of state changes (rate 1 on lengths ~150): paths of 64 runs and more keep their run count in
    a header record and are written in a second, replayed pass.  Posterior of the jump counts against the oracle chain."""
    Q = np.array([[-1.0, 1.0], [1.0, -1.0]])
    z = cases.tree2(T=6, S=1, seed=3, mean_branch=150.0)
    N, thin, burn = 3000, 6, 200
    ref = oracle.OracleRun(oracle.PLAIN, [z.oracle_dict()], Q, cases.PID2, 2.0, N, rng_mode=oracle.SEQUENTIAL, seed=11).run()[burn::thin]
    ch = pb.Chain(capi.PM_V_PLAIN, z, Q, cases.PID2, 2.0, N, seed=5, precision=precision)
    got = ch.run(N)[burn::thin]
    np.testing.assert_allclose(got[:, :2].sum(1), z.edge_length.sum(), rtol=1e-5)
    assert got[:, 2:].sum(1).mean() > 0.8 * z.edge_length.sum()       # ~ one change per unit length
    for col, name in [(2, "N01"), (3, "N10"), (0, "R0")]:
        p = stats.ks_2samp(got[:, col], ref[:, col]).pvalue
        assert p > 0.005, "%s: KS p = %.4f" % (name, p)
    e_long = int(np.argmax(z.edge_length))                            # the stored path of the longest branch adds up
    ln, st = ch.path(0, e_long, cap=4096)
    assert len(ln) > 64 and np.all(st[1:] != st[:-1])
    np.testing.assert_allclose(ln.sum(), z.edge_length[e_long], rtol=1e-4)
    big = pb.sumstatMCMC(cases.tree2(T=6, S=300, seed=3, mean_branch=150.0), Q, cases.PID2, 2.0, 12, seed=2, precision=precision)
    np.testing.assert_allclose(big[:, :2].sum(1), 300 * z.edge_length.sum(), rtol=1e-4)


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_partials_match_exact_pruning_on_peaked_data(precision):
    """Hidden-rate model with parity tip partials on a large tree: sub-clades are certain of their state, sibling
    partials multiply numbers like 1e-15 with structural zeros.  The stored partials must equal an exact (FP64, host)
    pruning for the same jump counts — in particular impossible states must stay at exactly zero.  (Regression: a
    floor on the stored partials once inflated them to 1e-18, i.e. to 1e-4 of a legitimate 1e-14 competitor.)"""
    Q, pid = cases.q4(), np.full(4, 0.25)
    T, S = 3000, 6
    tree = synth.yule_tree(T, seed=4, mean_branch=0.1 / 1.2)
    z = synth.simulate_4_state_tree(7, tree, Q, pid, n_sites=S, device="cuda", segments=2)
    ch = pb.Chain(capi.PM_V_KS, z, np.asfortranarray(Q.copy()), pid, 4.0, 3, prior=cases.PRIOR_KS, precision=precision, seed=3)
    ch.run(1)
    m = ch.piece_counts()          # jump counts the next pruning pass will use
    ch.time_prune(reps=1)          # K1 alone on the current state
    nen, _, root = z.order()
    par, chi = z.edge[:, 0] - 1, z.edge[:, 1] - 1
    B = np.array(ch.B)             # rewritten in place by the rate update that ended the sweep
    pows = [np.linalg.matrix_power(B, k) for k in range(int(m.max()) + 1)]
    for s in range(S):
        got = ch.partials(s)
        PL = np.zeros((2 * T - 1, 4))
        for i in range(T):
            PL[i, (0 if z.states[s, i] == 1 else 1)::2] = 1
        for i in range(T - 1):
            ea, eb = nen[2 * i] - 1, nen[2 * i + 1] - 1
            v = (pows[m[s, ea] - 1] @ PL[chi[ea]]) * (pows[m[s, eb] - 1] @ PL[chi[eb]])
            PL[par[ea]] = v / v.sum()
        exact = PL[T:]
        g = got[T:]
        assert np.all(g[exact == 0] == 0), "structural zeros must stay exact zeros"
        big = exact > 1e-6
        np.testing.assert_allclose(g[big], exact[big], rtol=2e-3 if precision == "f32" else 1e-9)
        assert np.all(g[exact < 1e-30] < 1e-20)


@pytest.mark.parametrize("what", ["bigtree_f32", "ks_f64", "ksmt_deterministic"])
def test_checkpoint_resume_continues_bit_for_bit(what):
    """pm_chain_export_state / import_state: a run cut in two (state exported, chain destroyed, a new chain created from
    the original inputs and resumed) returns the rows of the uninterrupted run, rate traces included."""
    N, cut = 14, 6
    if what == "bigtree_f32":
        Q = cases.q4()
        z = cases.tree_n(Q, T=300, S=200, seed=5, mean_branch=1.5, segments=3)   # long branches: records in use
        mk = lambda Qa: pb.Chain(capi.PM_V_BIGTREE, z, Qa, np.full(4, 0.25), 2.4, N, precision="f32", seed=11)
    elif what == "ks_f64":
        Q = cases.q4()
        z = cases.tree_hidden(Q, T=60, S=40, seed=4, mean_branch=0.5)
        mk = lambda Qa: pb.Chain(capi.PM_V_KS, z, Qa, np.full(4, 0.25), 4.0, N, prior=cases.PRIOR_KS, precision="f64", seed=11)
    else:
        Q = cases.q4()
        base = cases.tree_hidden(Q, T=20, S=3, seed=4, mean_branch=0.5)
        trees = [base, pb.PhyloTree(base.edge, base.edge_length * 1.2).with_states(base.states, segments=3)]
        mk = lambda Qa: pb.Chain(capi.PM_V_KSMT, trees, Qa, np.full(4, 0.25), 4.0, N, prior=cases.PRIOR_KSMT, seed=11,
                                 mode="deterministic")
    full = mk(np.asfortranarray(Q.copy())).run(N)
    Qa = np.asfortranarray(Q.copy())
    a = mk(Qa)
    head = a.run(cut)
    blob = a.export_state()
    Q_cut, B_cut = np.array(a.Q), np.array(a.B)
    a.close()
    Qb = np.asfortranarray(Q.copy())
    b = mk(Qb)
    b.import_state(blob)
    assert b.done == cut
    np.testing.assert_array_equal(np.array(b.Q), Q_cut)
    np.testing.assert_array_equal(np.array(b.B), B_cut)
    tail = b.run(N - cut)
    np.testing.assert_array_equal(np.vstack([head, tail]), full)
    # a blob written under another state format / buffer layout is refused, not reinterpreted (header word 10)
    other = blob.copy()
    other[8 + 4 * 9: 8 + 4 * 10] = np.frombuffer(np.int32(0x01abcdef).tobytes(), dtype=np.uint8)
    c = mk(np.asfortranarray(Q.copy()))
    with pytest.raises(capi.PhylomapError) as ei:
        c.import_state(other)
    assert ei.value.code == capi.PM_ERR_ARG and "format" in str(ei.value)
    # a state of another shape is refused
    other = pb.Chain(capi.PM_V_BIGTREE, cases.tree_n(cases.q4(), T=10, S=2, seed=1), np.asfortranarray(cases.q4()), np.full(4, 0.25), 2.4, 3)
    with pytest.raises(capi.PhylomapError) as ei:
        other.import_state(blob)
    assert ei.value.code == capi.PM_ERR_ARG
    with pytest.raises(capi.PhylomapError):
        other.import_state(np.zeros(16, dtype=np.uint8))


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_many_real_jumps_mixed_gap_and_count_runs(oracle, precision):
    """Fast-regime 4-state model on long branches: paths with three and more real jumps whose runs mix both ways of
    drawing virtual jumps (exponential gaps above lambda = 16, Poisson count below).  Regression: a gap-mode run used to
    skip its count word while the next sweep's regeneration consumed one per run (PM_DE_INCONSISTENT at sweep 2).
    Posterior of the sufficient statistics against the oracle chain, and a large-S run for the invariants."""
    Q, pid = cases.q4(), np.full(4, 0.25)
    z = cases.tree_n(Q, T=10, S=1, seed=5, mean_branch=5.0, segments=3)
    N, thin, burn = 8000, 10, 500
    ref = oracle.OracleRun(oracle.PLAIN, [z.oracle_dict()], Q, pid, 2.4, N, rng_mode=oracle.SEQUENTIAL, seed=11).run()[burn::thin]
    got = pb.sumstatMCMC(z, Q, pid, 2.4, N, seed=5, precision=precision)[burn::thin]
    np.testing.assert_allclose(got[:, :4].sum(1), z.edge_length.sum(), rtol=1e-5)
    cols = {"R0": got[:, 0], "R2": got[:, 2], "Ntot": got[:, 4:].sum(1)}
    refc = {"R0": ref[:, 0], "R2": ref[:, 2], "Ntot": ref[:, 4:].sum(1)}
    for name in cols:
        p = stats.ks_2samp(cols[name], refc[name]).pvalue
        assert p > 0.005, "%s: KS p = %.4f" % (name, p)
    zz = cases.tree_n(Q, T=300, S=200, seed=5, mean_branch=4.0, segments=2)
    big = pb.sumstatMCMC_bigtree(zz, Q, pid, 2.4, 12, seed=3, precision=precision)
    np.testing.assert_allclose(big[:, :4].sum(1), 200 * zz.edge_length.sum(), rtol=1e-4)


def _ladder_tree(T, mean_branch, rng):
    """Caterpillar: the deepest schedule a tree of T tips can have."""
    edge, node = [], T + 1
    for t in range(T, 2, -1):
        edge.append((node, t)); edge.append((node, node + 1)); node += 1
    edge.append((node, 1)); edge.append((node, 2))
    return pb.PhyloTree(np.array(edge, dtype=np.int32), rng.exponential(mean_branch, size=len(edge)) + 1e-3)


@pytest.mark.parametrize("case", range(12))
def test_random_models_against_oracle_chain(oracle, case):
    """Randomly drawn models (state count, dense generator, Omega, tree size, branch-length scale, sampler, precision):
    the production chain and the oracle chain must sample the same posterior of dwell times and jump counts.  Fixed
    seeds, so the outcome is deterministic."""
    rng = np.random.default_rng(1000 + case)
    n = int(rng.choice([2, 3, 5]))
    Q = rng.uniform(0.05, 0.5, size=(n, n))
    np.fill_diagonal(Q, 0)
    np.fill_diagonal(Q, -Q.sum(1))
    pid = np.full(n, 1.0 / n)
    Om = np.abs(np.diag(Q)).max() * float(rng.choice([1.0, 1.7, 4.0]))
    T = int(rng.choice([4, 9]))
    mb = float(rng.choice([0.3, 2.0, 10.0])) / np.abs(np.diag(Q)).max()   # expected changes per branch: 0.3 .. 10
    variant, fn = [(oracle.PLAIN, pb.sumstatMCMC), (oracle.SPARSE, pb.SPARSEsumstatMCMC), (oracle.BIGTREE, pb.sumstatMCMC_bigtree)][case % 3]
    precision = ["f32", "f64"][case % 2]
    seg = [None, 3][int(rng.integers(2))]  # None: the reference's initial maps
    if case < 8:
        z = cases.tree_n(Q, T=T, S=1, seed=40 + case, mean_branch=mb, segments=seg)
    else:  # other shapes: ladder (cases 8, 9: 12 and 25 tips) and balanced (10, 11: 8 and 32 tips)
        tree = _ladder_tree(12 if case == 8 else 25, mb, rng) if case < 10 else synth.balanced_tree(3 if case == 10 else 5, branch=mb)
        st = synth.simulate_tip_states(tree, Q, pid, 1, 300 + case).numpy()
        z = tree.with_states(st[0].astype(np.int32), segments=seg)
        T = z.T
    N, thin, burn = 12000, 12, 600
    if case >= 8:   # the ladder mixes slowly (lag-12 autocorrelation of R0 ~ 0.6 in case 8): thin much harder
        N, thin, burn = 40000, 100, 1000
    ref = oracle.OracleRun(variant, [z.oracle_dict()], Q, pid, Om, N, rng_mode=oracle.SEQUENTIAL, seed=7 + case).run()[burn::thin]
    got = fn(z, Q, pid, Om, N, seed=70 + case, precision=precision)[burn::thin]
    np.testing.assert_allclose(got[:, :n].sum(1), z.edge_length.sum(), rtol=1e-5)
    checks = {"R0": (got[:, 0], ref[:, 0]), "R_last": (got[:, n - 1], ref[:, n - 1]),
              "N_first": (got[:, n], ref[:, n]), "N_total": (got[:, n:].sum(1), ref[:, n:].sum(1))}
    for name, (a, b) in checks.items():
        # a dwell time has an atom at "no change anywhere": equal up to summation order, so compare on a 1e-7 grid
        p = stats.ks_2samp(np.round(a, 7), np.round(b, 7)).pvalue
        assert p > 0.001, "case %d (n=%d T=%d Omega=%.2f mean_branch=%.2f %s): %s KS p = %.5f" % (case, n, T, Om, mb, precision, name, p)


def test_hundred_thousand_tips():
    """A tree ten times the benchmark's (100 000 tips, 199 998 branches): schedules, 64-bit offsets and capacities hold;
    fixed-Q FP32 and the hidden-rate sampler in FP64."""
    Q, pid = cases.q4(), np.full(4, 0.25)
    T, S = 100000, 256
    tree = synth.yule_tree(T, seed=4, mean_branch=0.1 / 1.2)
    st = synth.simulate_tip_states(tree, Q, pid, S, seed=7, device="cuda").cpu().numpy()
    a = pb.sumstatMCMC_bigtree(tree.with_states(st, segments=2), Q, pid, 2.4, 4, precision="f32", seed=11)
    np.testing.assert_allclose(a[:, :4].sum(1), S * tree.edge_length.sum(), rtol=3e-4)
    assert np.all(a[:, 4:] >= 0) and np.array_equal(a[:, 4:], np.round(a[:, 4:])) and a[:, 4:].sum() > 0
    zk = synth.simulate_4_state_tree(7, tree, Q, pid, n_sites=S, device="cuda", segments=8)
    ks = pb.sumstatMCMCks(zk, np.asfortranarray(Q.copy()), pid, 4.0, 3, cases.PRIOR_KS, precision="f64", seed=3)
    np.testing.assert_allclose(ks[:, :4].sum(1), S * tree.edge_length.sum(), rtol=1e-9)


@pytest.mark.parametrize("small", [0, 1])
def test_production_rows_are_frozen(small, monkeypatch):
    """Regression fixture of the production arithmetic (tests/golden/make_production_golden.py): the same seeds give the
    same rows as when the fixture was written -- integer columns exactly, sums to rounding.  Regenerate it only after a
    deliberate change of the random-number mapping."""
    import json
    import os
    sys_path_golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_production_golden", os.path.join(sys_path_golden, "make_production_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = json.load(open(os.path.join(sys_path_golden, "production", "rows.json")))
    # the fixture was written by the 32-sites-per-warp kernels; the one-block-per-site kernel (the default at these sizes,
    # pm_small.cuh) draws the same histories and adds the dwell times up in another order (FP32 partial sums: 2e-9)
    monkeypatch.setenv("PHYLOMAP_B200_SMALL", str(small))
    got = mod.rows()
    assert sorted(got) == sorted(want)
    for k, ref in want.items():
        ref = np.asarray(ref)
        np.testing.assert_allclose(got[k], ref, rtol=1e-8 if small else 1e-12, atol=0, err_msg=k)
        whole = ref == np.round(ref)
        assert np.array_equal(got[k][whole], ref[whole]), k


def test_site_tiles_equal_one_call(monkeypatch):
    """A fixed-Q call that does not fit in device memory is cut into site tiles run one after the other (pm_host.cu,
    one_call).  Forced here with PHYLOMAP_B200_SITE_TILE: the tiles are keyed by global site index, so the summed
    rows equal the untiled call -- integer counts exactly, dwell times to FP32 summation order."""
    Q, pid = cases.q4(), np.full(4, 0.25)
    z = cases.tree_n(Q, T=120, S=333, seed=12, mean_branch=0.4, segments=2)
    whole = pb.sumstatMCMC_bigtree(z, Q, pid, 2.4, 6, seed=21, precision="f32")
    ex_whole = pb.sumstatEXP(z, Q, pid, 4, seed=22, precision="f32")
    monkeypatch.setenv("PHYLOMAP_B200_SITE_TILE", "100")
    tiled = pb.sumstatMCMC_bigtree(z, Q, pid, 2.4, 6, seed=21, precision="f32")
    assert np.array_equal(tiled[:, 4:], whole[:, 4:])
    np.testing.assert_allclose(tiled[:, :4], whole[:, :4], rtol=1e-5)
    ex_tiled = pb.sumstatEXP(z, Q, pid, 4, seed=22, precision="f32")
    assert np.array_equal(ex_tiled[:, 4:], ex_whole[:, 4:])


@pytest.mark.parametrize("what", ["bigtree_f32", "bigtree_f32_many_sites", "ks_f32", "ksmt_f64"])
def test_rows_do_not_depend_on_the_memory_layout(what, monkeypatch):
    """The fused prune + node-draw kernel (partials in per-block slots claimed per SM, PHYLOMAP_B200_FUSED) and record
    slices shared by groups of 32 sites (PHYLOMAP_B200_REC_POOL) only change where scratch and records live: every site
    draws from its own Philox keys, so the rows -- counts, dwell times, rate traces -- must be identical bit for bit to the
    two separate kernels with the full partials array / per-site slices, also with a ragged last block and with more site
    blocks than slots (slots are claimed and handed on), and a stored path must read back the same from either layout."""
    N = 10
    if what.startswith("bigtree_f32"):
        Q = cases.q4()
        many = what.endswith("many_sites")     # 470 site blocks > 444 slots on a B200
        z = cases.tree_n(Q, T=40 if many else 300, S=15013 if many else 203, seed=5, mean_branch=1.5, segments=3)   # long branches: records in use
        mk = lambda: pb.Chain(capi.PM_V_BIGTREE, z, Q.copy(), np.full(4, 0.25), 2.4, N, precision="f32", seed=11)
    elif what == "ks_f32":
        Q = cases.q4()
        z = cases.tree_hidden(Q, T=90, S=150, seed=4, mean_branch=0.8)
        mk = lambda: pb.Chain(capi.PM_V_KS, z, np.asfortranarray(Q.copy()), np.full(4, 0.25), 4.0, N, prior=cases.PRIOR_KS,
                              precision="f32", seed=11)
    else:
        Q = cases.q4()
        base = cases.tree_hidden(Q, T=40, S=70, seed=4, mean_branch=0.5)
        trees = [base, pb.PhyloTree(base.edge, base.edge_length * 1.2).with_states(base.states, segments=3)]
        mk = lambda: pb.Chain(capi.PM_V_KSMT, trees, np.asfortranarray(Q.copy()), np.full(4, 0.25), 4.0, N,
                              prior=cases.PRIOR_KSMT, precision="f64", seed=11)

    monkeypatch.setenv("PHYLOMAP_B200_SMALL", "0")   # the layouts of the 32-sites-per-warp kernels (test_gpu_small.py covers the other)

    def run(fused, pool):
        monkeypatch.setenv("PHYLOMAP_B200_FUSED", str(fused))
        monkeypatch.setenv("PHYLOMAP_B200_REC_POOL", str(pool))
        ch = mk()
        rows = ch.run(N)
        t0 = z if what != "ksmt_f64" else trees[0]
        sites = (0, 33, 69) if t0.n_sites() < 1000 else (0, 7000, 15012)
        paths = [ch.path(s, e, cap=256) for s in sites for e in range(0, t0.E, max(1, t0.E // 12))]
        pc = ch.piece_counts()
        ch.time_prune(reps=1)                     # K1 alone on the final state: what partials() reports in either layout
        pls = [ch.partials(s) for s in sites]
        ch.close()
        return rows, paths, pc, pls

    ref_rows, ref_paths, ref_pc, ref_pls = run(0, 0)   # separate kernels + full partials array, one record slice per (site, chunk)
    for fused, pool in [(1, 0), (0, 1), (1, 1)]:
        rows, paths, pc, pls = run(fused, pool)
        for a, b in zip(ref_pls, pls):
            assert np.array_equal(a, b)
        assert np.array_equal(rows, ref_rows), "rows differ with fused %d, pooled records %d" % (fused, pool)
        assert np.array_equal(pc, ref_pc)
        for (l0, s0), (l1, s1) in zip(ref_paths, paths):
            assert np.array_equal(l0, l1) and np.array_equal(s0, s1)


@pytest.mark.parametrize("what", ["plain_f32_one_site", "bigtree_f64", "plain_deterministic", "sparse_f32_generic_n"])
def test_graph_replay_gives_the_rows_of_plain_launches(what, monkeypatch):
    """Small problems replay one captured sweep (CUDA graph, sweep index and output row in device memory) instead of
    launching every kernel from the host: the rows must be those of the launch-by-launch run, bit for bit, across several
    pm_chain_run calls of different lengths (the statistics buffer moves: the graph is captured again)."""
    if what == "plain_f32_one_site":
        z = cases.tree2(T=100, S=1, seed=1, mean_branch=5.0)
        mk = lambda: pb.Chain(capi.PM_V_PLAIN, z, cases.Q2, cases.PID2, 0.2, 64, precision="f32", seed=5)
    elif what == "bigtree_f64":
        Q = cases.q4()
        z = cases.tree_n(Q, T=60, S=45, seed=5, mean_branch=1.5, segments=3)
        mk = lambda: pb.Chain(capi.PM_V_BIGTREE, z, Q.copy(), np.full(4, 0.25), 2.4, 64, precision="f64", seed=11)
    elif what == "plain_deterministic":
        z = cases.tree2(T=30, S=3, seed=2, mean_branch=3.0)
        mk = lambda: pb.Chain(capi.PM_V_PLAIN, z, cases.Q2, cases.PID2, 0.2, 64, precision="f64", mode="deterministic", seed=5)
    else:
        Q = cases.jc(5, 0.1)
        z = cases.tree_n(Q, T=40, S=9, seed=3, mean_branch=1.0, segments=2)
        mk = lambda: pb.Chain(capi.PM_V_SPARSE, z, Q.copy(), np.full(5, 0.2), 1.0, 64, precision="f32", seed=7)

    monkeypatch.setenv("PHYLOMAP_B200_SMALL", "0")   # (the one-block-per-site kernel has no per-sweep launches to replay)

    def run(graph):
        monkeypatch.setenv("PHYLOMAP_B200_GRAPH", str(graph))
        ch = mk()
        rows = np.vstack([ch.run(5), ch.run(1), ch.run(30), ch.run(28)])
        ns = ch.node_states()
        ch.close()
        return rows, ns

    ref, ns0 = run(0)
    got, ns1 = run(1)
    assert np.array_equal(got, ref)
    assert np.array_equal(ns0, ns1)

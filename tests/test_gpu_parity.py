"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.

Deterministic mode (FP64, reference order of operations, keyed or replayed uniforms): node states, piece counts,
merged paths and integer transition counts must be IDENTICAL; dwell-time sums within 1e-9 relative (the sums
are accumulated in a different order; the bar in BASELINE.json is 1e-6).
"""
import numpy as np
import pytest

import cases
import phylomap_b200 as pb
from phylomap_b200 import capi, synth

pytestmark = pytest.mark.gpu

DET = dict(mode="deterministic", precision="f64")


def _oracle(oracle, variant, trees, Q, pid, Omega, N, prior=None, seed=7, **kw):
    o = oracle.OracleRun(variant, [t.oracle_dict() for t in trees], Q, pid, Omega, N, prior=prior,
                         rng_mode=kw.pop("rng_mode", oracle.KEYED), seed=seed, **kw)
    return o, o.run()


def _compare_rows(got, ref, n, int_cols, tol=1e-9):
    assert got.shape == ref.shape
    for c in range(ref.shape[1]):
        if c in int_cols:
            assert np.array_equal(got[:, c], ref[:, c]), "integer column %d differs" % c
        else:
            np.testing.assert_allclose(got[:, c], ref[:, c], rtol=tol, atol=1e-12, err_msg="column %d" % c)


def _compare_state(chain, orc, tree, S, E, n_paths=40, tree_idx=0, exact_lengths=True):
    assert np.array_equal(chain.node_states(tree_idx), orc.node_states(tree_idx))
    assert np.array_equal(chain.piece_counts(tree_idx), orc.piece_counts(tree_idx))
    rng = np.random.default_rng(0)
    for _ in range(n_paths):
        s, e = int(rng.integers(S)), int(rng.integers(E))
        gl, gs = chain.path(s, e, tree_idx)
        ol, os_ = orc.path(s, e, tree_idx)
        assert np.array_equal(gs, os_)
        if exact_lengths:
            np.testing.assert_array_equal(gl, ol)  # same additions in the same order: bit-exact
        else:
            # rate-updating samplers: the proposals see dwell-time SUMS, which the GPU accumulates in another order
            # (1e-13 relative), so rates and hence piece lengths agree to rounding, not bit for bit
            np.testing.assert_allclose(gl, ol, rtol=1e-9)


@pytest.mark.parametrize("variant,name", [(capi.PM_V_PLAIN, "PLAIN"), (capi.PM_V_SPARSE, "SPARSE"),
                                          (capi.PM_V_BIGTREE, "BIGTREE")])
@pytest.mark.parametrize("S", [1, 37])
def test_fixed_q_two_state(oracle, variant, name, S):
    z = cases.tree2(T=24, S=S, seed=3)
    N, Om = 25, 0.2
    orc, ref = _oracle(oracle, getattr(oracle, name), [z], cases.Q2, cases.PID2, Om, N)
    ch = pb.Chain(variant, z, cases.Q2.copy(), cases.PID2, Om, N, seed=7, **DET)
    got = ch.run()
    _compare_rows(got, ref, 2, int_cols={2, 3})
    _compare_state(ch, orc, z, S, z.E)


@pytest.mark.parametrize("variant,name", [(capi.PM_V_PLAIN, "PLAIN"), (capi.PM_V_BIGTREE, "BIGTREE")])
def test_fixed_q_four_state(oracle, variant, name):
    Q = cases.q4()
    z = cases.tree_n(Q, T=40, S=19, seed=5, mean_branch=0.8, segments=4)
    N, Om = 20, 2.4
    pid = np.full(4, 0.25)
    orc, ref = _oracle(oracle, getattr(oracle, name), [z], Q, pid, Om, N)
    ch = pb.Chain(variant, z, Q.copy(), pid, Om, N, seed=7, **DET)
    got = ch.run()
    _compare_rows(got, ref, 4, int_cols=set(range(4, 16)))
    _compare_state(ch, orc, z, 19, z.E)


def test_sparse_threshold_active(oracle):
    """SPARSE with B entries <= 1e-7 dropped (matTospmat, src/phylomap.cpp:811)."""
    Q = np.array([[-0.1, 0.1, 0.0], [1e-9, -0.1 - 1e-9, 0.1], [0.05, 0.05, -0.1]])
    z = cases.tree_n(cases.jc(3), T=16, S=5, seed=2, mean_branch=3.0, segments=4)
    N, Om = 15, 0.4
    pid = np.full(3, 1 / 3)
    orc, ref = _oracle(oracle, oracle.SPARSE, [z], Q, pid, Om, N)
    ch = pb.Chain(capi.PM_V_SPARSE, z, Q.copy(), pid, Om, N, seed=7, **DET)
    _compare_rows(ch.run(), ref, 3, int_cols=set(range(3, 9)))
    _compare_state(ch, orc, z, 5, z.E)


@pytest.mark.parametrize("n", [3, 6, 20])
def test_generic_state_count(oracle, n):
    """Run-time state count (the tutorial's 20-state tridiagonal-like chain, phylomap_tutorial.Rnw:119-135)."""
    Q = cases.jc(n, 0.02)
    z = cases.tree_n(Q, T=12, S=3, seed=n, mean_branch=4.0)
    N, Om = 6, 2 * 0.02 * (n - 1)
    pid = np.full(n, 1.0 / n)
    orc, ref = _oracle(oracle, oracle.PLAIN, [z], Q, pid, Om, N)
    ch = pb.Chain(capi.PM_V_PLAIN, z, Q.copy(), pid, Om, N, seed=7, **DET)
    _compare_rows(ch.run(), ref, n, int_cols=set(range(n, n * n)))
    _compare_state(ch, orc, z, 3, z.E, n_paths=20)


def test_bf(oracle):
    z = cases.tree2(T=30, S=4, seed=9)
    N, Om = 40, 0.5
    Qo, Qg = cases.Q2.copy(), np.asfortranarray(cases.Q2.copy())
    orc, ref = _oracle(oracle, oracle.BF, [z], Qo, cases.PID2, Om, N, prior=cases.PRIOR_BF)
    ch = pb.Chain(capi.PM_V_BF, z, Qg, cases.PID2, Om, N, prior=cases.PRIOR_BF, seed=7, **DET)
    got = ch.run()
    _compare_rows(got, ref, 2, int_cols={2, 3, 4, 5, 8}, tol=1e-8)
    np.testing.assert_allclose(Qg, orc.Q, rtol=1e-8)      # Q updated in place like the reference
    np.testing.assert_allclose(ch.B, orc.B, rtol=1e-8)
    _compare_state(ch, orc, z, 4, z.E, exact_lengths=False)


def test_ks(oracle):
    Q = cases.q4()
    z = cases.tree_hidden(Q, T=30, S=3, seed=4, mean_branch=0.5)
    N, Om = 30, 4.0
    pid = np.full(4, 0.25)
    orc, ref = _oracle(oracle, oracle.KS, [z], Q.copy(), pid, Om, N, prior=cases.PRIOR_KS)
    Qg = np.asfortranarray(Q.copy())
    ch = pb.Chain(capi.PM_V_KS, z, Qg, pid, Om, N, prior=cases.PRIOR_KS, seed=7, **DET)
    got = ch.run()
    n = 4
    _compare_rows(got, ref, n, int_cols=set(range(n, n + n * n)) | {ref.shape[1] - 1}, tol=1e-7)
    np.testing.assert_allclose(Qg, orc.Q, rtol=1e-7, atol=1e-12)
    _compare_state(ch, orc, z, 3, z.E, exact_lengths=False)


def test_ks_six_state(oracle):
    Q = cases.q6()
    z = cases.tree_hidden(Q, T=16, S=2, seed=6, mean_branch=0.5)
    N, Om = 12, 8.0
    pid = np.full(6, 1 / 6)
    orc, ref = _oracle(oracle, oracle.KS, [z], Q.copy(), pid, Om, N, prior=cases.PRIOR_KS)
    ch = pb.Chain(capi.PM_V_KS, z, np.asfortranarray(Q.copy()), pid, Om, N, prior=cases.PRIOR_KS, seed=7, **DET)
    n = 6
    _compare_rows(ch.run(), ref, n, int_cols=set(range(n, n + n * n)) | {ref.shape[1] - 1}, tol=1e-7)


def _loglik_scipy(z, Q, pid, parity=False):
    """log p(y | Q) summed over sites: pruning with scipy's expm, independent of both the oracle and the library."""
    from scipy.linalg import expm
    T, n = z.T, Q.shape[0]
    nen, _, root = z.order()
    st = z.states if z.states.ndim == 2 else z.states[None, :]
    P = [expm(Q * t) for t in z.edge_length]
    tot = 0.0
    for s in range(st.shape[0]):
        PL = np.zeros((2 * T - 1, n))
        for i in range(T):
            if parity:
                PL[i, (0 if st[s, i] % 2 == 1 else 1)::2] = 1
            else:
                PL[i, st[s, i] - 1] = 1
        S = 0.0
        for i in range(T - 1):
            ea, eb = nen[2 * i] - 1, nen[2 * i + 1] - 1
            v = (P[ea] @ PL[z.edge[ea, 1] - 1]) * (P[eb] @ PL[z.edge[eb, 1] - 1])
            S += np.log(v.sum())
            PL[z.edge[ea, 0] - 1] = v / v.sum()
        tot += np.log(PL[root - 1] @ pid) + S
    return tot


def test_dic_two_state(oracle):
    """maketreelistMCMC2sDICt: the bf chain + log p(y|Q) by matrix exponentiation (src/phylomap.cpp:3183-3264)."""
    z = cases.tree2(T=30, S=4, seed=9)
    N, Om = 25, 0.5
    orc, ref = _oracle(oracle, oracle.DIC2S, [z], cases.Q2.copy(), cases.PID2, Om, N, prior=cases.PRIOR_BF)
    got = pb.sumstatMCMC2sDICt(z, np.asfortranarray(cases.Q2.copy()), cases.PID2, Om, N, cases.PRIOR_BF, seed=7, **DET)
    assert got.shape == (N, 10)
    _compare_rows(got, ref, 2, int_cols={2, 3, 4, 5, 8}, tol=1e-8)
    np.testing.assert_allclose(got[:, 9], ref[:, 9], rtol=1e-9)       # bar: 1e-6 relative in FP64
    for i in (0, 7, N - 1):                                            # and against an independent scipy evaluation
        Q = np.array([[-got[i, 6], got[i, 6]], [got[i, 7], -got[i, 7]]])
        np.testing.assert_allclose(got[i, 9], _loglik_scipy(z, Q, cases.PID2), rtol=1e-9)
    f32 = pb.sumstatMCMC2sDICt(z, np.asfortranarray(cases.Q2.copy()), cases.PID2, Om, 6, cases.PRIOR_BF, seed=7, precision="f32")
    for i in range(6):                                                 # FP32 production mode: 1e-4 relative
        Q = np.array([[-f32[i, 6], f32[i, 6]], [f32[i, 7], -f32[i, 7]]])
        np.testing.assert_allclose(f32[i, 9], _loglik_scipy(z, Q, cases.PID2), rtol=1e-4)


def test_dic_hidden_rates(oracle):
    """maketreelistMCMCksDICt (src/phylomap.cpp:3300-3403): ks chain + log-likelihood with parity tip partials."""
    Q = cases.q4()
    z = cases.tree_hidden(Q, T=24, S=3, seed=4, mean_branch=0.5)
    N, Om = 20, 4.0
    pid = np.full(4, 0.25)
    orc, ref = _oracle(oracle, oracle.DICKS, [z], Q.copy(), pid, Om, N, prior=cases.PRIOR_KS)
    got = pb.sumstatMCMCksDICt(z, np.asfortranarray(Q.copy()), pid, Om, N, cases.PRIOR_KS, seed=7, **DET)
    n = 4
    assert got.shape == (N, n + n * n + 2 + 3 + 2)
    _compare_rows(got[:, :-1], ref[:, :-1], n, int_cols=set(range(n, n + n * n)) | {ref.shape[1] - 2}, tol=1e-7)
    np.testing.assert_allclose(got[:, -1], ref[:, -1], rtol=1e-8)
    np.testing.assert_allclose(got[0, -1], _loglik_scipy(z, Q, pid, parity=True), rtol=1e-9)


def test_loglik_and_dic_helpers():
    """pm_loglik and make{2,4}stateDIC(big) (R/sourceme.R:141-177, 248-284, 445-516): D(Q-hat) at the posterior-mean rates
    against scipy, DIC = D + 2 pD from a DIC trace."""
    z = cases.tree2(T=40, S=5, seed=3)
    for prec, tol in (("f64", 1e-10), ("f32", 1e-4)):
        np.testing.assert_allclose(pb.loglik(z, cases.Q2, cases.PID2, precision=prec), _loglik_scipy(z, cases.Q2, cases.PID2), rtol=tol)
    mat = pb.sumstatMCMC2sDICt(z, np.asfortranarray(cases.Q2.copy()), cases.PID2, 0.5, 40, cases.PRIOR_BF, seed=7)
    Qh = np.array([[-mat[:, 6].mean(), mat[:, 6].mean()], [mat[:, 7].mean(), -mat[:, 7].mean()]])
    D = -2 * _loglik_scipy(z, Qh, cases.PID2)
    want = D + 2 * (np.mean(-2 * mat[:, 9]) - D)
    np.testing.assert_allclose(pb.make2stateDIC(mat, z, cases.PID2), want, rtol=1e-10)
    np.testing.assert_allclose(pb.make2stateDICbig(mat, z, cases.PID2, z.order()[0]), want, rtol=1e-10)
    import pandas as pd                      # the reference addresses the columns by name
    frame = pd.DataFrame(mat, columns=["t0", "t1", "n00", "n01", "n10", "n11", "l01", "l10", "root_state", "log(p(y|Q))"])
    np.testing.assert_allclose(pb.make2stateDIC(frame, z, cases.PID2), want, rtol=1e-10)

    Q = cases.q4()
    pid = np.full(4, 0.25)
    zk = cases.tree_hidden(Q, T=24, S=3, seed=4, mean_branch=0.5)
    np.testing.assert_allclose(pb.loglik(zk, Q, pid, parity_tips=True), _loglik_scipy(zk, Q, pid, parity=True), rtol=1e-10)
    mk = pb.sumstatMCMCksDICt(zk, np.asfortranarray(Q.copy()), pid, 4.0, 30, cases.PRIOR_KS, seed=7)
    Qk = synth.make2sQ(*[mk[:, 20 + i].mean() for i in range(5)])
    Dk = -2 * _loglik_scipy(zk, Qk, pid, parity=True)
    np.testing.assert_allclose(pb.make4stateDICbig(mk, zk, pid), Dk + 2 * (np.mean(-2 * mk[:, -1]) - Dk), rtol=1e-10)
    with pytest.raises(capi.PhylomapError) as ei:
        pb.loglik(zk, np.zeros((3, 3)), np.full(3, 1 / 3), parity_tips=True)
    assert ei.value.code == capi.PM_ERR_ARG


def _tree_set(maker, k, **kw):
    base = maker(seed=11, **kw)
    out = [base]
    rng = np.random.default_rng(5)
    for _ in range(k - 1):
        el = base.edge_length * rng.uniform(0.7, 1.3, size=base.E)
        t = pb.PhyloTree(base.edge, el).with_states(base.states)
        out.append(t)
    return out


def test_mt(oracle):
    trees = _tree_set(lambda seed, **kw: cases.tree2(T=18, S=3, seed=seed), 3)
    N, Om = 25, 0.5
    orc, ref = _oracle(oracle, oracle.MT, trees, cases.Q2.copy(), cases.PID2, Om, N, prior=cases.PRIOR_BF)
    ch = pb.Chain(capi.PM_V_MT, trees, np.asfortranarray(cases.Q2.copy()), cases.PID2, Om, N, prior=cases.PRIOR_BF,
                  seed=7, **DET)
    _compare_rows(ch.run(), ref, 2, int_cols={2, 3, 4, 5, 8}, tol=1e-8)
    for ti in range(3):
        _compare_state(ch, orc, trees[ti], 3, trees[ti].E, n_paths=10, tree_idx=ti, exact_lengths=False)


def test_ksmt(oracle):
    Q = cases.q4()
    trees = _tree_set(lambda seed, **kw: cases.tree_hidden(Q, T=14, S=2, seed=seed, mean_branch=0.5), 2)
    N, Om = 15, 4.0
    pid = np.full(4, 0.25)
    orc, ref = _oracle(oracle, oracle.KSMT, trees, Q.copy(), pid, Om, N, prior=cases.PRIOR_KSMT)
    ch = pb.Chain(capi.PM_V_KSMT, trees, np.asfortranarray(Q.copy()), pid, Om, N, prior=cases.PRIOR_KSMT, seed=7, **DET)
    n = 4
    _compare_rows(ch.run(), ref, n, int_cols=set(range(n, n + n * n)) | {ref.shape[1] - 1}, tol=1e-7)


@pytest.mark.parametrize("variant,name,prior", [(capi.PM_V_PLAIN, "PLAIN", None), (capi.PM_V_BF, "BF", cases.PRIOR_BF)])
def test_replay_of_sequential_r_stream(oracle, variant, name, prior):
    """Tier 1 of BASELINE.json: the kernels consume a uniform stream exported from a sequential (R-order,
    Mersenne-Twister) run of the reference algorithm and reproduce its histories and counts."""
    z = cases.tree2(T=16, S=2, seed=13)
    N, Om = 12, 0.4
    orc = oracle.OracleRun(getattr(oracle, name), [z.oracle_dict()], cases.Q2.copy(), cases.PID2, Om, N, prior=prior,
                           rng_mode=oracle.SEQUENTIAL, seed=101, want_log=True)
    ref = orc.run()
    table, host = orc.export_log()
    ch = pb.Chain(variant, z, np.asfortranarray(cases.Q2.copy()), cases.PID2, Om, N, prior=prior, seed=101,
                  table=table, host_table=host, **DET)
    got = ch.run()
    _compare_rows(got, ref, 2, int_cols={2, 3} if prior is None else {2, 3, 4, 5, 8}, tol=1e-8)
    _compare_state(ch, orc, z, 2, z.E, n_paths=20, exact_lengths=prior is None)


def _reference_fixtures():
    import glob
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    return sorted(p for p in glob.glob(os.path.join(here, "golden", "reference", "*.json")) if "/exp_" not in p)


@pytest.mark.parametrize("path", _reference_fixtures(), ids=lambda p: p.split("/")[-1][:-5])
def test_gpu_replays_the_reference(oracle, path):
    """Tier 1 of BASELINE.json against THE REFERENCE ITSELF: tests/golden/reference/*.json are outputs of the unmodified
    src/phylomap.cpp (oracle/_ref, made by tests/golden/make_reference_golden.py).  The kernels consume the uniform
    stream of that R-order run (recorded by the oracle restatement, which reproduces the reference bit for bit:
    tests/test_reference_pin.py) and must return the reference's rows: integer counts, root / tree columns identical,
    dwell times 1e-9 (bar 1e-6), rates 1e-7, log-likelihood 1e-9; Q and B end where the reference left them."""
    import json
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_reference_golden as mrg
    with open(path) as f:
        d = json.load(f)
    case, want = d["case"], np.array(d["rows"])
    trees = mrg.trees_of(case)
    Q0, pid, n = np.array(case["Q"]), np.array(case["pid"]), len(case["pid"])
    orc = oracle.OracleRun(getattr(oracle, case["variant"]), [t.oracle_dict() for t in trees], Q0.copy(), pid, case["Omega"],
                           case["N"], prior=case.get("prior"), rng_mode=oracle.SEQUENTIAL, seed=case["seed"], want_log=True)
    orc.run()
    table, host = orc.export_log()
    variant = {"PLAIN": capi.PM_V_PLAIN, "SPARSE": capi.PM_V_SPARSE, "BIGTREE": capi.PM_V_BIGTREE, "BF": capi.PM_V_BF,
               "KS": capi.PM_V_KS, "MT": capi.PM_V_MT, "KSMT": capi.PM_V_KSMT, "DIC2S": capi.PM_V_DIC2S,
               "DICKS": capi.PM_V_DICKS}[case["variant"]]
    Qg = np.asfortranarray(Q0.copy())
    ch = pb.Chain(variant, trees if len(trees) > 1 else trees[0], Qg, pid, case["Omega"], case["N"], prior=case.get("prior"),
                  seed=case["seed"], table=table, host_table=host, **DET)
    got = ch.run()
    v = case["variant"]
    if v in ("PLAIN", "SPARSE", "BIGTREE"):
        ints, ll = set(range(n, n * n)), None
    elif v in ("BF", "MT"):
        ints, ll = {2, 3, 4, 5, 8}, None
    elif v == "DIC2S":
        ints, ll = {2, 3, 4, 5, 8}, 9
    elif v in ("KS", "KSMT"):
        ints, ll = set(range(n, n + n * n)) | {want.shape[1] - 1}, None
    else:
        ints, ll = set(range(n, n + n * n)) | {want.shape[1] - 2}, want.shape[1] - 1
    assert got.shape == want.shape
    for c in range(want.shape[1]):
        if c in ints:
            assert np.array_equal(got[:, c], want[:, c]), "integer column %d differs from the reference" % c
        elif c == ll:
            np.testing.assert_allclose(got[:, c], want[:, c], rtol=1e-9)
        else:
            np.testing.assert_allclose(got[:, c], want[:, c], rtol=1e-9 if c < n else 1e-7, atol=1e-12, err_msg="column %d" % c)
    np.testing.assert_allclose(Qg, np.array(d["Q_after"]), rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(ch.B, np.array(d["B_after"]), rtol=1e-7, atol=1e-12)


def test_one_call_entries_match_chain(oracle):
    """The seven drop-in entries (pm_maketreelist*) are the chain interface run in one go."""
    z = cases.tree2(T=20, S=3, seed=21)
    N, Om = 10, 0.2
    _, ref = _oracle(oracle, oracle.PLAIN, [z], cases.Q2, cases.PID2, Om, N)
    got = pb.sumstatMCMC(z, cases.Q2, cases.PID2, Om, N, seed=7, **DET)
    _compare_rows(got, ref, 2, int_cols={2, 3})
    _, ref = _oracle(oracle, oracle.BF, [z], cases.Q2.copy(), cases.PID2, 0.5, N, prior=cases.PRIOR_BF)
    got = pb.sumstatMCMCbf(z, cases.Q2.copy(), cases.PID2, 0.5, N, cases.PRIOR_BF, seed=7, **DET)
    _compare_rows(got, ref, 2, int_cols={2, 3, 4, 5, 8}, tol=1e-8)


def test_chunked_run_equals_single_run():
    z = cases.tree2(T=20, S=5, seed=2)
    a = pb.Chain(capi.PM_V_PLAIN, z, cases.Q2.copy(), cases.PID2, 0.2, 12, seed=3, **DET).run()
    c = pb.Chain(capi.PM_V_PLAIN, z, cases.Q2.copy(), cases.PID2, 0.2, 12, seed=3, **DET)
    b = np.vstack([c.run(5), c.run(4), c.run(3)])
    assert np.array_equal(a, b)


def test_errors_like_the_reference():
    z = cases.tree2(T=10, S=1, seed=2)
    with pytest.raises(capi.PhylomapError) as e:  # zero root prior -> RcppArmadillo::sample throws
        pb.sumstatMCMC(z, cases.Q2, np.array([0.0, 0.0]), 0.2, 2, **DET)
    assert e.value.code == capi.PM_ERR_SAMPLE
    bad = pb.PhyloTree(z.edge, z.edge_length, np.array([1, 2, 3] + [1] * 7), z.maps, z.mapnames)
    with pytest.raises(capi.PhylomapError) as e:
        pb.sumstatMCMC(bad, cases.Q2, cases.PID2, 0.2, 2, **DET)
    assert e.value.code == capi.PM_ERR_ARG


def test_golden_vectors():
    """The committed fixtures (tests/golden/*.json: inputs + expected rows) through the CUDA library."""
    import json
    import os
    import sys
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gold)
    import make_golden
    var = {"PLAIN": capi.PM_V_PLAIN, "SPARSE": capi.PM_V_SPARSE, "BIGTREE": capi.PM_V_BIGTREE, "BF": capi.PM_V_BF,
           "KS": capi.PM_V_KS, "MT": capi.PM_V_MT, "KSMT": capi.PM_V_KSMT, "DIC2S": capi.PM_V_DIC2S,
           "DICKS": capi.PM_V_DICKS}
    names = [f for f in sorted(os.listdir(gold)) if f.endswith(".json")]
    assert len(names) >= 9
    for name in names:
        g = json.load(open(os.path.join(gold, name)))
        case = g["case"]
        z = make_golden.tree_from_case(case)
        trees = [z]
        for sc in case.get("extra_tree_scales", []):
            trees.append(pb.PhyloTree(z.edge, z.edge_length * np.array(sc), z.states,
                                      [m * s for m, s in zip(z.maps, sc)], z.mapnames))
        Q = np.asfortranarray(np.array(case["Q"]))
        n = Q.shape[0]
        ch = pb.Chain(var[case["variant"]], trees if len(trees) > 1 else z, Q, np.array(case["pid"]), case["Omega"],
                      case["N"], prior=case.get("prior"), seed=case["seed"], **DET)
        got, ref = ch.run(), np.array(g["rows"])
        fixed = case["variant"] in ("PLAIN", "SPARSE", "BIGTREE")
        dic = case["variant"].startswith("DIC")
        ints = set(range(n, ref.shape[1])) if fixed else set(range(n, n + n * n)) | {ref.shape[1] - (2 if dic else 1)}
        _compare_rows(got, ref, n, ints, tol=1e-7)


def test_smallest_tree_and_zero_iterations(oracle):
    """A cherry (2 tips, empty nodelist) and N = 0."""
    tree = pb.PhyloTree(np.array([[3, 1], [3, 2]]), np.array([1.0, 2.0])).with_states(np.array([1, 2]))
    _, ref = _oracle(oracle, oracle.PLAIN, [tree], cases.Q2, cases.PID2, 0.2, 30)
    got = pb.sumstatMCMC(tree, cases.Q2, cases.PID2, 0.2, 30, seed=7, **DET)
    _compare_rows(got, ref, 2, int_cols={2, 3})
    assert pb.sumstatMCMC(tree, cases.Q2, cases.PID2, 0.2, 0, seed=7, **DET).shape == (0, 4)


def test_long_branches_many_pieces(oracle):
    """Omega * t ~ 30 per branch: hundreds of pieces per branch, powers of B beyond the shared-memory table."""
    z = cases.tree2(T=10, S=3, seed=4, mean_branch=150.0)
    N, Om = 6, 0.2
    orc, ref = _oracle(oracle, oracle.PLAIN, [z], cases.Q2, cases.PID2, Om, N)
    ch = pb.Chain(capi.PM_V_PLAIN, z, cases.Q2.copy(), cases.PID2, Om, N, seed=7, **DET)
    _compare_rows(ch.run(), ref, 2, int_cols={2, 3})
    _compare_state(ch, orc, z, 3, z.E, n_paths=15)
    assert orc.piece_counts().max() > 40


def test_medium_tree_many_chunks(oracle):
    """300 tips x 70 sites, 4 states: several branch chunks per site tile, a partial last site block, record slices
    shared by many branches — still identical to the oracle in deterministic mode."""
    Q = cases.q4()
    z = cases.tree_n(Q, T=300, S=70, seed=12, mean_branch=0.5, segments=3)
    N, Om = 5, 2.4
    pid = np.full(4, 0.25)
    orc, ref = _oracle(oracle, oracle.BIGTREE, [z], Q, pid, Om, N)
    ch = pb.Chain(capi.PM_V_BIGTREE, z, Q.copy(), pid, Om, N, seed=7, **DET)
    _compare_rows(ch.run(), ref, 4, int_cols=set(range(4, 16)))
    _compare_state(ch, orc, z, 70, z.E, n_paths=60)


def test_squamate_vignette_configuration(oracle):
    """The reference's own tree in the DIC vignette's setup (Squamate_DIC_model_selection.Rnw:78-104): 3 951 tips, every
    branch cut into 100 segments, Q2, prior2r, Omega = 10 -- about 877 000 jump points per site and sweep, jump counts in
    the thousands on the long branches.  Deterministic mode against the oracle, two sites."""
    Q2 = np.array([[-0.001, 0.001], [0.006, -0.006]])
    prior = np.array([0.55, 1.0, 0.55, 1.0])
    z = synth.simulate_2_state_tree(101, cases.squamate_tree(), Q2, cases.PID2, n_sites=2, segments=100)
    N, Om = 3, 10.0
    orc, ref = _oracle(oracle, oracle.DIC2S, [z], Q2.copy(), cases.PID2, Om, N, prior=prior)
    ch = pb.Chain(capi.PM_V_DIC2S, z, np.asfortranarray(Q2.copy()), cases.PID2, Om, N, prior=prior, seed=7, **DET)
    got = ch.run(N)
    _compare_rows(got, ref, 2, int_cols={2, 3, 4, 5, 8}, tol=1e-8)
    np.testing.assert_allclose(got[:, 9], ref[:, 9], rtol=1e-9)
    assert np.array_equal(ch.node_states(), orc.node_states(0))
    assert np.array_equal(ch.piece_counts(), orc.piece_counts(0))
    assert ch.piece_counts().max() > 1000
    # production arithmetic on the same input: the invariants of a row
    fast = pb.sumstatMCMC2sDICt(z, np.asfortranarray(Q2.copy()), cases.PID2, Om, 4, prior, seed=7, precision="f64")
    np.testing.assert_allclose(fast[:, :2].sum(1), 2 * z.edge_length.sum(), rtol=1e-9)
    assert np.all(np.abs(fast[1:, 2:6].sum(1) / (2 * Om * z.edge_length.sum()) - 1) < 0.01)


def test_config0_hundred_tips_thousand_sweeps(oracle):
    """configs[0] of BASELINE.json, the reference's own CPU-runnable case: simulate_2_state_tree on a 100-tip tree,
    sumstatMCMC for 1000 iterations, one site.  Deterministic mode against the oracle over the whole run."""
    z = cases.tree2(T=100, S=1, seed=1, mean_branch=10.0)
    N, Om = 1000, 0.2
    orc, ref = _oracle(oracle, oracle.PLAIN, [z], cases.Q2, cases.PID2, Om, N)
    got = pb.sumstatMCMC(z, cases.Q2, cases.PID2, Om, N, seed=7, **DET)
    assert got.shape == (N, 4)
    _compare_rows(got, ref, 2, int_cols={2, 3}, tol=1e-9)
    fast = pb.sumstatMCMC(z, cases.Q2, cases.PID2, Om, N, seed=7)          # production arithmetic, same posterior
    from scipy import stats
    for col in (2, 3, 0):
        assert stats.ks_2samp(fast[100::10, col], ref[100::10, col]).pvalue > 0.005

"""The one-block-per-site chain kernel of small problems (pm_small.cuh: a block owns ONE site, its state lives in shared
memory, the sweeps of a fixed-Q call run inside one launch) against the 32-sites-per-warp kernels (PHYLOMAP_B200_SMALL=0).
Every draw takes the same Philox key and every item runs the same arithmetic, so node states, piece counts, stored paths
and transition counts must agree bit for bit; dwell-time sums are accumulated in another order (1e-6 in FP32)."""
import numpy as np
import pytest

import cases
import phylomap_b200 as pb
from phylomap_b200 import capi

pytestmark = pytest.mark.gpu


def _case(what):
    """-> (make_chain, first tree, N, runs, number of dwell columns)"""
    if what == "plain_f32_one_character":          # BASELINE configs[0]: the reference's literal call
        z = cases.tree2(T=100, S=1, seed=1, mean_branch=5.0)
        return (lambda: pb.Chain(capi.PM_V_PLAIN, z, cases.Q2, cases.PID2, 0.2, 300, precision="f32", seed=5)), z, (120, 1, 179), 2
    if what == "bigtree_f32_records":               # long branches: paths with records, three- and four-piece shapes
        Q = cases.q4()
        z = cases.tree_n(Q, T=300, S=37, seed=5, mean_branch=1.5, segments=3)
        return (lambda: pb.Chain(capi.PM_V_BIGTREE, z, Q.copy(), np.full(4, 0.25), 2.4, 24, precision="f32", seed=11)), z, (5, 19), 4
    if what == "sparse_f64_gap_mode":               # lambda > 16 on most branches: exponential gaps, long paths, many segments
        Q = cases.q4()
        z = cases.tree_n(Q, T=40, S=3, seed=7, mean_branch=4.0, segments=40)
        return (lambda: pb.Chain(capi.PM_V_SPARSE, z, Q.copy(), np.full(4, 0.25), 12.0, 12, precision="f64", seed=3)), z, (12,), 4
    if what == "bf_f64":
        z = cases.tree2(T=60, S=2, seed=3, mean_branch=3.0)
        return (lambda: pb.Chain(capi.PM_V_BF, z, np.asfortranarray(cases.Q2.copy()), cases.PID2, 0.4, 30, prior=cases.PRIOR_BF,
                                 precision="f64", seed=9)), z, (30,), 2
    if what == "ks_f64":                            # hidden-rate model: parity tips in the pruning, tips redrawn
        Q = cases.q4()
        z = cases.tree_hidden(Q, T=90, S=20, seed=4, mean_branch=0.8)
        return (lambda: pb.Chain(capi.PM_V_KS, z, np.asfortranarray(Q.copy()), np.full(4, 0.25), 4.0, 16, prior=cases.PRIOR_KS,
                                 precision="f64", seed=11)), z, (16,), 4
    if what == "ksmt_f64":
        Q = cases.q4()
        base = cases.tree_hidden(Q, T=40, S=7, seed=4, mean_branch=0.5)
        trees = [base, pb.PhyloTree(base.edge, base.edge_length * 1.2).with_states(base.states, segments=3)]
        return (lambda: pb.Chain(capi.PM_V_KSMT, trees, np.asfortranarray(Q.copy()), np.full(4, 0.25), 4.0, 16,
                                 prior=cases.PRIOR_KSMT, precision="f64", seed=11)), base, (16,), 4
    if what == "bf_f64_one_character":              # one site + rate updates: the block writes its row straight into mapped host memory
        z = cases.tree2(T=60, S=1, seed=3, mean_branch=3.0)
        return (lambda: pb.Chain(capi.PM_V_BF, z, np.asfortranarray(cases.Q2.copy()), cases.PID2, 0.4, 40, prior=cases.PRIOR_BF,
                                 precision="f64", seed=9)), z, (25, 15), 2
    if what == "ksmt_f64_one_character":
        Q = cases.q4()
        base = cases.tree_hidden(Q, T=40, S=1, seed=4, mean_branch=0.5)
        trees = [base, pb.PhyloTree(base.edge, base.edge_length * 1.2).with_states(base.states, segments=3)]
        return (lambda: pb.Chain(capi.PM_V_KSMT, trees, np.asfortranarray(Q.copy()), np.full(4, 0.25), 4.0, 30,
                                 prior=cases.PRIOR_KSMT, precision="f64", seed=11)), base, (30,), 4
    if what == "dic_ks_f64":
        Q = cases.q4()
        z = cases.tree_hidden(Q, T=50, S=5, seed=6, mean_branch=0.6)
        return (lambda: pb.Chain(capi.PM_V_DICKS, z, np.asfortranarray(Q.copy()), np.full(4, 0.25), 4.0, 10, prior=cases.PRIOR_KS,
                                 precision="f64", seed=2)), z, (10,), 4
    raise ValueError(what)


@pytest.mark.parametrize("what", ["plain_f32_one_character", "bigtree_f32_records", "sparse_f64_gap_mode", "bf_f64", "ks_f64",
                                  "ksmt_f64", "dic_ks_f64", "bf_f64_one_character", "ksmt_f64_one_character"])
def test_one_block_per_site_gives_the_rows_of_the_wide_kernels(what, monkeypatch):
    mk, z, runs, nd = _case(what)

    def run(small):
        monkeypatch.setenv("PHYLOMAP_B200_SMALL", str(small))
        monkeypatch.setenv("PHYLOMAP_B200_SMALL_WORK", "1e15")   # (some cases carry longer paths than the kernel is chosen for by default)
        ch = mk()
        rows = np.vstack([ch.run(c) for c in runs])
        ns, pc = ch.node_states(), ch.piece_counts()
        S = z.n_sites()
        paths = [ch.path(s, e, cap=4096) for s in sorted({0, S // 2, S - 1}) for e in range(0, z.E, max(1, z.E // 16))]
        ch.close()
        return rows, ns, pc, paths

    wide = run(0)
    small = run(1)
    assert np.array_equal(small[1], wide[1]), "node states differ"
    assert np.array_equal(small[2], wide[2]), "piece counts differ"
    f32 = "f32" in what
    for (l0, s0), (l1, s1) in zip(wide[3], small[3]):
        assert np.array_equal(s0, s1)
        np.testing.assert_allclose(l1, l0, rtol=0 if f32 or what.startswith(("plain", "bigtree", "sparse")) else 1e-9, atol=0)
    a, b = small[0], wide[0]
    assert a.shape == b.shape
    if what in ("plain_f32_one_character", "bigtree_f32_records", "sparse_f64_gap_mode"):   # fixed Q: [dwell | counts]
        assert np.array_equal(a[:, nd:], b[:, nd:]), "transition counts differ"
        np.testing.assert_allclose(a[:, :nd], b[:, :nd], rtol=2e-6 if f32 else 1e-12)
    else:   # rate traces consume the dwell-time sums: every column to the rounding of those sums
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)


def test_small_limit_and_switch(monkeypatch):
    """More sites than PHYLOMAP_B200_SMALL_SITES: the wide kernels run (same rows either way, so only the launch count tells)."""
    Q = cases.q4()
    z = cases.tree_n(Q, T=50, S=9, seed=2, mean_branch=1.0, segments=2)

    def launches(env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ch = pb.Chain(capi.PM_V_BIGTREE, z, Q.copy(), np.full(4, 0.25), 2.4, 40, precision="f32", seed=1)
        rows = ch.run(40)
        n = ch.kernel_times()[1]
        ch.close()
        return rows, n

    r_small, n_small = launches({"PHYLOMAP_B200_SMALL": "1", "PHYLOMAP_B200_SMALL_SITES": "592"})
    r_wide, n_wide = launches({"PHYLOMAP_B200_SMALL": "1", "PHYLOMAP_B200_SMALL_SITES": "8"})
    assert n_small == 2            # the chain kernel + the row reduction, for all 40 sweeps
    assert n_wide >= 40 * 4
    assert np.array_equal(r_small[:, 4:], r_wide[:, 4:])
